"""2..N-rank check of the sharded engine over NCCL (run under torch.distributed.run):
every rank steps its strip of ONE simulation; the global lane counts must equal those of an
unsharded engine that rank 0 runs next to it."""
import os
import sys
import tempfile
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cityflow_b200  # noqa: E402
from cityflow_b200 import scenario  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    rows, cols, steps = (int(x) for x in (sys.argv[1:4] + ["6", "6", "400"][len(sys.argv) - 1:]))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d = tempfile.mkdtemp()
    cfg = scenario.make_grid_scenario(d, rows, cols, dense=dict(frac=1.0 if rows <= 10 else 0.5, interval=4.0 if rows <= 10 else 10.0, seed=1), name="sh")
    ids = [cityflow_b200.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    eng = cityflow_b200.Engine(cfg, thread_num=1, device=local, shard_rank=rank, shard_world=world, nccl_id=ids[0])
    ref = cityflow_b200.Engine(cfg, thread_num=1, device=local) if rank == 0 else None
    ok = True
    t_sh = 0.0
    for s in range(1, steps + 1):
        t0 = time.perf_counter()
        eng.next_step()
        if s % 50 == 0 or s == steps:
            n = eng.get_vehicle_count()          # collective
            lanes = eng.get_lane_vehicle_count()  # collective
        torch.cuda.synchronize()
        t_sh += time.perf_counter() - t0
        if ref is not None:
            ref.next_step()
            if s % 50 == 0 or s == steps:
                rn, rl = ref.get_vehicle_count(), ref.get_lane_vehicle_count()
                same = (n == rn) and (lanes == rl)
                ok = ok and same
                print("step %d: vehicles sharded %d / unsharded %d, lane counts %s" % (s, n, rn, "equal" if same else "DIFFER"), flush=True)
    dist.barrier()
    if rank == 0:
        print("SHARD_CHECK", "OK" if ok else "FAILED", "world", world, "ms/step(sharded, incl. checks)", 1e3 * t_sh / steps)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
