"""One GPU: the sharded engine in loop-back mode (W ranks, device copies instead of NCCL) AND the unsharded
engine against the compiled reference's observables (`refdump counts`) on the bench scenario of N GPUs
(30 x 30*W grid), through the bench's timed window.  Tells apart "the seam protocol diverges" from "the
engine diverges from the reference at this size".

    python tools/shard_loopback_ref_check.py [rows cols world steps every]"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cityflow_b200 import scenario  # noqa: E402
from cityflow_b200.capi import CEngine, CShardGroup  # noqa: E402
from oracle import harness as H  # noqa: E402


def main():
    a = [int(x) for x in sys.argv[1:]]
    rows, cols, world, steps, every = (a + [30, 60, 2, 1225, 25][len(a):])[:5]
    d = tempfile.mkdtemp()
    cfg = scenario.make_grid_scenario(d, rows, cols, name="bench", dense=dict(frac=0.5, interval=10.0, seed=1))
    t0 = time.time()
    ref = H.RefDump.counts(cfg, steps, os.cpu_count() or 8, every)
    print("reference: %d steps in %.1f s, final count %d" % (steps, time.time() - t0, ref["vehicle_count"][-1]), flush=True)
    one = CEngine(cfg)
    grp = CShardGroup(cfg, world)
    ok = True
    for s in range(1, steps + 1):
        one.next_step()
        grp.next_step()
        if s % every and s != steps:
            continue
        rc, rw, _ = ref["dumps"][s]
        n1, ng, nr = one.vehicle_count(), grp.vehicle_count(), int(ref["vehicle_count"][s - 1])
        l1, lg = one.lane_vehicle_count(), grp.lane_counts(one.n_lanes)
        w1, wg = one.lane_waiting_count(), grp.lane_counts(one.n_lanes, True)
        e1 = n1 == nr and np.array_equal(l1, rc) and np.array_equal(w1, rw)
        eg = ng == nr and np.array_equal(lg, rc) and np.array_equal(wg, rw)
        if not (e1 and eg) or s % (every * 8) == 0 or s == steps:
            print("step %d: ref %d | unsharded %d (%s, %d lanes differ) | loop-back x%d %d (%s, %d lanes differ)" % (
                s, nr, n1, "equal" if e1 else "DIFF", int((l1 != rc).sum()), world, ng, "equal" if eg else "DIFF",
                int((lg != rc).sum())), flush=True)
        ok = ok and e1 and eg
        if not (e1 and eg):
            break
    print("SHARD_LOOPBACK_REF", "OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
