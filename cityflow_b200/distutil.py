"""Multi-process plumbing for bench.py (one process per GPU, launched by torch.distributed.run): the barrier around the
timed regions and the max / sum reductions of the per-rank timings and vehicle-step counts.  The engine's own cross-rank
traffic (seam records of a sharded run) does not go through here.  The helpers work on any backend
(nccl on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import os
from typing import Optional


class Ranks:
    def __init__(self, backend: Optional[str] = None, device=None):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        self.device = device
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            if not dist.is_initialized():
                kw = {}
                if backend == "nccl" and device is not None:
                    kw["device_id"] = device
                dist.init_process_group(backend or "gloo", **kw)
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def _reduce(self, x: float, op) -> float:
        if self.dist is None:
            return float(x)
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64, device=self.device if self.device is not None else "cpu")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x: float) -> float:
        return self._reduce(x, self.dist.ReduceOp.MAX) if self.dist is not None else float(x)

    def sum(self, x: float) -> float:
        return self._reduce(x, self.dist.ReduceOp.SUM) if self.dist is not None else float(x)

    def close(self):
        if self.dist is not None and self.dist.is_initialized():
            self.dist.destroy_process_group()


def aggregate_throughput(ranks: Ranks, local_units: float, local_seconds: float) -> float:
    """Whole-job throughput: units of all ranks / slowest rank's time (task contract: max over ranks)."""
    total = ranks.sum(local_units)
    slowest = ranks.max(local_seconds)
    return total / slowest if slowest > 0 else 0.0
