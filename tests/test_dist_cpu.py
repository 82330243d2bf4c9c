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
