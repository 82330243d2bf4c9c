"""Lane change on the GPU against the restatement, every field of every vehicle (shadows included), every step:

    python tools/lc_gpu_check.py [rows cols steps]

Compares the engine's running vehicles INCLUDING shadows (cfb_debug_lc_vehicles: partner, offset, waiting time, leader,
blocker, ...) and the per-lane counts with the restatement's (PortOracle.lc_snapshot), which is pinned against
oracle/_ref/refdump_lcorder.  Prints the first difference per field and stops.  CITYFLOW_B200_LC_SERIAL=1 selects the
one-thread forms of the scheduling / control-tail kernels."""
import ctypes
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cityflow_b200 import scenario  # noqa: E402
from cityflow_b200.capi import CEngine  # noqa: E402
from oracle import harness as H  # noqa: E402


def main():
    rows, cols, steps = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (3, 3, 600)))
    d = tempfile.mkdtemp()
    cfg = scenario.make_grid_scenario(d, rows, cols, dense=dict(frac=1.0, interval=3.0, seed=2), name="lc", lane_change=True)
    eng = CEngine(cfg)
    lib = eng.lib
    lib.cfb_debug_lc_vehicles.restype = ctypes.c_int64
    lib.cfb_debug_lc_vehicles.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
    ora = H.PortOracle(cfg)
    shadows = 0
    for s in range(1, steps + 1):
        eng.next_step()
        ora.next_step()
        want = ora.lc_snapshot()
        n = int(lib.cfb_debug_lc_vehicles(eng.h, None, 0))
        got = np.zeros(n, H.LC_DTYPE)
        if n:
            lib.cfb_debug_lc_vehicles(eng.h, got.ctypes.data, n)
        bad = []
        if eng.vehicle_count() != want.vehicle_count:
            bad.append("vehicle count %d vs %d" % (eng.vehicle_count(), want.vehicle_count))
        if not np.array_equal(eng.lane_vehicle_count(), want.lane_count):
            bad.append("lane counts differ on %d lanes" % int((eng.lane_vehicle_count() != want.lane_count).sum()))
        if len(got) != len(want.vehicles):
            bad.append("running vehicles %d vs %d" % (len(got), len(want.vehicles)))
        else:
            for f in H.LC_DTYPE.names:
                ne = np.nonzero(got[f] != want.vehicles[f])[0]
                if len(ne):
                    k = ne[0]
                    bad.append("%s differs for %d vehicles, first flow_%d_%d prio %d: engine %r restatement %r" % (
                        f, len(ne), want.vehicles["flow"][k], want.vehicles["cnt"][k], want.vehicles["priority"][k],
                        got[f][k], want.vehicles[f][k]))
        shadows += int((want.vehicles["partner_type"] == 2).sum())
        if bad:
            print("step %d:\n  %s" % (s, "\n  ".join(bad[:12])))
            return 1
    print("%d steps equal, %d vehicles at the end, %d shadow-steps" % (steps, len(got), shadows))
    return 0


if __name__ == "__main__":
    sys.exit(main())
