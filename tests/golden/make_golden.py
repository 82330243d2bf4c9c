"""Generates the committed golden fixtures from the COMPILED, UNMODIFIED reference
(oracle/_ref/refdump, built by oracle/Makefile from /root/reference).  Run here (the container
that has the reference); the fixtures travel with the repo.

    python tests/golden/make_golden.py

Each fixture holds, for a seeded synthetic scenario (re-creatable anywhere with
cityflow_b200.scenario), a digest per checkpoint of the full dynamic state after that step:
per-lane vehicle counts, per-lane waiting counts and, for every running vehicle sorted by id,
(flow, index, drivable, leader, blocker, raw IEEE-754 bits of distance / speed / gap).
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

SCENARIOS = {
    "g1_default": dict(rows=1, cols=1, dense=None, steps=400, every=20),
    "g2_dense": dict(rows=2, cols=2, dense=dict(frac=1.0, interval=3.0, seed=5), steps=400, every=20),
    "g3_dense": dict(rows=3, cols=3, dense=dict(frac=1.0, interval=3.0, seed=3), steps=600, every=25),
    "g6_dense": dict(rows=6, cols=6, dense=dict(frac=1.0, interval=4.0, seed=1), steps=1000, every=50),
}


# laneChange=true fixtures: generated from oracle/_ref/refdump_lcorder -- the reference with the
# worker vehicle set ordered by priority and two uninitialised members zeroed (oracle/lc_order_patch.sh
# says why the unmodified build cannot be a pin) -- single worker thread.
LC_SCENARIOS = {
    "lc3_default": dict(rows=3, cols=3, dense=None, steps=600, every=25),
    "lc4_dense": dict(rows=4, cols=4, dense=dict(frac=1.0, interval=3.0, seed=2), steps=500, every=25),
}


def lc_state_digest(st) -> str:
    """Digest of a lane-change StepState (harness.parse_runlc / PortOracle.lc_snapshot): every
    running vehicle including shadows in vehiclePool order, plus the list order of every drivable."""
    v = st.vehicles
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(st.lane_count, dtype="<i4").tobytes())
    for f in ("flow", "cnt", "priority", "partner_type", "partner", "drivable", "leader", "blocker", "flags", "last_dir"):
        h.update(np.ascontiguousarray(v[f], dtype="<i4").tobytes())
    for f in ("dis", "speed", "gap", "offset", "waiting_time", "last_change_time"):
        a = np.ascontiguousarray(v[f], dtype="<f8").copy()
        a[a == 0] = 0.0
        h.update(a.tobytes())
    for o in st.order:
        h.update(np.ascontiguousarray(o, dtype="<i4").tobytes())
    return h.hexdigest()


def make_lc_config(name, directory):
    from cityflow_b200 import scenario
    s = LC_SCENARIOS[name]
    return scenario.make_grid_scenario(directory, s["rows"], s["cols"], dense=s["dense"], name=name, lane_change=True)


def state_digest(st) -> str:
    """Digest of a harness.StepState (only fields every engine can produce)."""
    v = st.key_sorted()
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(st.lane_count, dtype="<i4").tobytes())
    h.update(np.ascontiguousarray(st.lane_waiting, dtype="<i4").tobytes())
    for f in ("flow", "cnt", "drivable", "leader_flow", "leader_cnt", "blocker_flow", "blocker_cnt"):
        h.update(np.ascontiguousarray(v[f], dtype="<i4").tobytes())
    for f in ("dis", "speed", "gap"):
        a = np.ascontiguousarray(v[f], dtype="<f8").copy()
        a[a == 0] = 0.0  # -0.0 == +0.0
        h.update(a.tobytes())
    return h.hexdigest()


def make_config(name, directory):
    from cityflow_b200 import scenario
    s = SCENARIOS[name]
    return scenario.make_grid_scenario(directory, s["rows"], s["cols"], dense=s["dense"], name=name)


def main():
    from oracle import harness as H
    assert H.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    for name, s in SCENARIOS.items():
        with tempfile.TemporaryDirectory() as d:
            cfg = make_config(name, d)
            o = H.PortOracle(cfg)  # only for the table sizes
            ref = H.RefDump.run(cfg, s["steps"], 1, s["every"], n_inter=o.n_inter, n_drivables=o.n_drivables)
        out = {"scenario": {k: s[k] for k in ("rows", "cols", "dense", "steps", "every")}, "generator": "oracle/_ref/refdump (unmodified reference)",
               "checkpoints": [{"step": st.step, "vehicles": st.vehicle_count, "finished": st.finished,
                                "lane_count_sum": int(st.lane_count.sum()), "digest": state_digest(st)} for st in ref]}
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(out, f, indent=1)
        print(name, len(ref), "checkpoints, final vehicles", ref[-1].vehicle_count, "finished", ref[-1].finished)
    assert H.have_lc_ref(), "build oracle/_ref/refdump_lcorder first (make -C oracle ref)"
    for name, s in LC_SCENARIOS.items():
        with tempfile.TemporaryDirectory() as d:
            cfg = make_lc_config(name, d)
            o = H.PortOracle(cfg)
            ref = H.RefDump.runlc(cfg, s["steps"], s["every"], n_inter=o.n_inter, n_drivables=o.n_drivables)
        out = {"scenario": {k: s[k] for k in ("rows", "cols", "dense", "steps", "every")},
               "generator": "oracle/_ref/refdump_lcorder (reference + oracle/lc_order_patch.sh), thread_num=1, laneChange=true",
               "checkpoints": [{"step": st.step, "vehicles": st.vehicle_count, "finished": st.finished,
                                "shadows": int((st.vehicles["partner_type"] == 2).sum()),
                                "lane_count_sum": int(st.lane_count.sum()), "digest": lc_state_digest(st)} for st in ref]}
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(out, f, indent=1)
        print(name, len(ref), "checkpoints, final vehicles", ref[-1].vehicle_count, "finished", ref[-1].finished,
              "shadows seen", sum(c["shadows"] for c in out["checkpoints"]))


if __name__ == "__main__":
    main()
