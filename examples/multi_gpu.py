"""One simulation cut into column strips across the GPUs of a node:

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 examples/multi_gpu.py
"""
import os
import sys
import tempfile

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cityflow_b200  # noqa: E402
from cityflow_b200 import scenario  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl")
box = [None]
if rank == 0:
    d = tempfile.mkdtemp()
    box[0] = (scenario.make_grid_scenario(d, 12, 12 * world, dense=dict(frac=0.5, interval=10.0, seed=1)), cityflow_b200.nccl_unique_id())
dist.broadcast_object_list(box, src=0)          # every rank loads the same files and joins the same NCCL group
cfg, nccl_id = box[0]
eng = cityflow_b200.Engine(cfg, device=local, shard_rank=rank, shard_world=world, nccl_id=nccl_id)
for _ in range(600):
    eng.next_step()                              # every rank, in lock step
n = eng.get_vehicle_count()                      # collective: network-wide on every rank
lanes = eng.get_lane_vehicle_count()
if rank == 0:
    print("%d GPUs, %d vehicles in the whole network, %d on lanes" % (world, n, sum(lanes.values())))
dist.destroy_process_group()
