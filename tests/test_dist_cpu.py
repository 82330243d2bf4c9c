"""world_size-2 gloo test (CPU) of the N>1 path of bench.py: barrier + max-over-ranks timing + summed
units, with each rank stepping its own replica (here the CPU restatement stands in for the engine)."""
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, cfg, out):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from cityflow_b200.distutil import Ranks, aggregate_throughput
    from oracle import harness as H
    r = Ranks("gloo")
    o = H.PortOracle(cfg)
    o.next_step(50 + 10 * rank)          # replicas may differ in progress: weak scaling, no exchange
    units = 0
    for _ in range(20):
        o.next_step()
        units += o.vehicle_count()
    r.barrier()
    secs = 1.0 + rank                   # deterministic "timings": the slowest rank (2.0 s) must win
    agg = aggregate_throughput(r, units, secs)
    tot = r.sum(units)
    out.put((rank, units, tot, agg))
    r.close()


def test_two_rank_replica_aggregation(cfg_3x3_dense):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, cfg_3x3_dense, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, u0, t0, a0), (r1, u1, t1, a1) = res
    assert t0 == t1 == u0 + u1
    assert a0 == a1 == pytest.approx((u0 + u1) / 2.0)
    assert u0 > 0 and u1 > 0


# ------------------------------------------------------------------------------------------
# The seam index tables of the peer-memory transport (partition.h seamTables, uploaded by DeviceSim::shardConnect): every
# rank derives its own from the road network, independently, in its own process; a message only arrives where the receiver
# expects it if the tables of the two ends meet.  Each rank runs the product's table code (tests/partition_probe.cpp) for
# itself, the tables are exchanged over gloo, and every rank checks every seam lane of every pair it takes part in.
def _seam_worker(rank, world, port, cfg, probe, out):
    import json
    import subprocess
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    dist.init_process_group("gloo")
    mine = json.loads(subprocess.check_output([probe, cfg, str(world), str(rank)]))
    tables = [None] * world
    dist.all_gather_object(tables, {k: mine[k] for k in ("nbr", "out_peer", "out_dst", "in_peer", "in_dst", "out_lanes", "in_lanes")})
    bad = []
    me = tables[rank]
    for j, lane in enumerate(me["out_lanes"]):            # a mover message: my out entry j -> the owner's in list
        q, d = me["out_peer"][j], me["out_dst"][j]
        if not (0 <= d < len(tables[q]["in_lanes"])) or tables[q]["in_lanes"][d] != lane or tables[q]["in_peer"][d] != rank:
            bad.append(("mover", rank, q, lane))
    for j, lane in enumerate(me["in_lanes"]):             # a tail message: my in entry j -> the feeder's out list
        q, d = me["in_peer"][j], me["in_dst"][j]
        if not (0 <= d < len(tables[q]["out_lanes"])) or tables[q]["out_lanes"][d] != lane or tables[q]["out_peer"][d] != rank:
            bad.append(("tail", rank, q, lane))
    for q in me["nbr"]:                                     # neighbourhood is symmetric (both ends wait for each other's flags)
        if rank not in tables[q]["nbr"]:
            bad.append(("nbr", rank, q, -1))
    n_msgs = len(me["out_lanes"]) + len(me["in_lanes"])
    dist.barrier()
    out.put((rank, bad, n_msgs, me["nbr"]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_seam_tables_of_independent_ranks_meet(cfg_6x6, world):
    import subprocess
    probe = os.path.join(ROOT, "oracle", "_build", "partition_probe")
    csrc = os.path.join(ROOT, "cityflow_b200", "csrc")
    os.makedirs(os.path.dirname(probe), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "partition_probe.cpp"),
                           os.path.join(csrc, "partition.cpp"), os.path.join(csrc, "roadnet.cpp"), "-o", probe])
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_seam_worker, args=(r, world, port, cfg_6x6, probe, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=180) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, bad, n_msgs, nbr in res:
        assert not bad, bad[:5]
        assert n_msgs > 0
        assert nbr == [r for r in (rank - 1, rank + 1) if 0 <= r < world]     # column strips: a rank only touches its neighbours
