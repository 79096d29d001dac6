"""Index-range sharding of a batch of independent pairs.

No collective is on the data path: rank/GPU k scores pairs [k*n/G, (k+1)*n/G) and the
scores are gathered by writing disjoint slices of one host array.  The same partition is
used inside the C ABI (one host thread per GPU, csrc/swb200_api.cu score_host) and by
bench.py under torchrun (one process per GPU)."""
from __future__ import annotations

from typing import Tuple


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range of shard `rank` of `world` over n pairs; the ranges tile [0, n)."""
    if world < 1 or not (0 <= rank < world) or n < 0:
        raise ValueError("bad shard arguments")
    return n * rank // world, n * (rank + 1) // world


def max_over_ranks(value: float, dist=None) -> float:
    """Multi-GPU timings are the MAX over ranks (a job is as slow as its slowest shard)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, dist=None) -> float:
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
