"""CPU tests of the multi-GPU host logic: index-range partition and the
max-over-ranks reduction, with a world_size-2 gloo run standing in for two GPU ranks.
The scorer in the workers is the ORACLE (this is a test of the plumbing, not of the kernel)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_tile_the_batch():
    from sharding import shard_range
    for n in (0, 1, 2, 7, 1_000_000, 100_000_001):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for (lo0, hi0), (lo1, hi1) in zip(edges, edges[1:]):
                assert hi0 == lo1 and lo0 <= hi0
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "smith-waterman-simd_b200"))
    from oracle import oracle as O
    from sharding import max_over_ranks, shard_range, sum_over_ranks
    import swb200
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = shard_range(n, rank, world)
    a, b = swb200.counter_pairs(lo, hi - lo)          # any rank can produce any index range
    scores = np.memmap(out_path, dtype=np.int32, mode="r+", shape=(n,))
    scores[lo:hi] = O.score_batch(a, b, O.MATRIX_SPEEDTEST, 15)   # host gather = disjoint slices of one array
    scores.flush()
    t = max_over_ranks(1.0 + rank, dist)
    total = sum_over_ranks(hi - lo, dist)
    assert t == float(world) and total == float(n)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shards_equal_single_rank(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "smith-waterman-simd_b200"))
    from oracle import oracle as O
    import swb200
    O.build()
    n, world = 1001, 2
    out_path = str(tmp_path / "scores.i32")
    np.memmap(out_path, dtype=np.int32, mode="w+", shape=(n,)).flush()
    mp.spawn(_worker, args=(world, _free_port(), n, out_path), nprocs=world, join=True)
    got = np.array(np.memmap(out_path, dtype=np.int32, mode="r", shape=(n,)))
    a, b = swb200.counter_pairs(0, n)
    assert np.array_equal(got, O.score_batch(a, b, O.MATRIX_SPEEDTEST, 15))


def test_bench_has_no_collective_inside_rank0_only_code():
    """A collective that only rank 0 executes hangs every multi-GPU run.  Static check of bench.py."""
    import ast
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    collectives = {"max_over_ranks", "sum_over_ranks", "barrier", "all_reduce", "all_gather", "broadcast"}
    bad = []
    for node in ast.walk(tree):
        if isinstance(node, ast.If) and "rank == 0" in ast.unparse(node.test):
            for sub in ast.walk(ast.Module(body=node.body, type_ignores=[])):
                if isinstance(sub, ast.Call):
                    name = sub.func.attr if isinstance(sub.func, ast.Attribute) else getattr(sub.func, "id", "")
                    if name in collectives:
                        bad.append((node.lineno, name))
    assert not bad, bad


def test_cpulist_parser_and_numa_binding_is_best_effort(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "smith-waterman-simd_b200"))
    import swb200
    assert swb200._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert swb200._parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    r = swb200.bind_to_gpu_numa_node(0, sysfs=str(tmp_path))    # no GPU / no sysfs entry: reports, never raises
    assert r is not None and r["bound"] is False
    assert os.sched_getaffinity(0) == before
