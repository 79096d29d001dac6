"""Streaming mode (BASELINE.json configs[2] and configs[4]).

Host threads generate pairs of the counter-based stream straight into a ring of PINNED
buffers; each full buffer is handed to `swb200_submit` (C ABI), which moves it through the
GPU in chunks (H2D, kernel, D2H overlapped) while the producers fill the next buffers.
With several GPUs the pair-index space is split in contiguous ranges: in one process the
context shards every batch over its GPUs; under torchrun every rank streams its own range
(`shard_range`).  No collective is on the data path.

The report names the bottleneck: `produce_s` (host generation), `wait_s` (time the driver
thread spent blocked on the GPU pipeline) and the wall time.
"""
from __future__ import annotations

import time
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field
from typing import Callable, List, Optional

import numpy as np

import swb200


@dataclass
class StreamReport:
    pairs: int = 0
    wall_s: float = 0.0
    produce_s: float = 0.0       # summed over batches: time spent generating (all producer threads, wall per batch)
    wait_s: float = 0.0          # driver thread blocked on the GPU pipeline (a slot it needs is still in flight)
    batches: int = 0
    bytes_h2d: int = 0
    bytes_d2h: int = 0
    score_sum: int = 0
    fnv_of_batch_fnvs: int = 1469598103934665603
    sampled: List = field(default_factory=list)

    @property
    def alignments_per_s(self) -> float:
        return self.pairs / self.wall_s if self.wall_s else 0.0

    @property
    def gcups(self) -> float:
        return self.alignments_per_s * 16384 / 1e9

    def bottleneck(self) -> str:
        busy_gen = self.produce_s / self.wall_s if self.wall_s else 0.0
        busy_gpu = self.wait_s / self.wall_s if self.wall_s else 0.0
        if busy_gen > 0.85 and busy_gpu < 0.3:
            return "host generation"
        return "PCIe/kernel pipeline" if busy_gpu >= busy_gen else "host generation"


class StreamRunner:
    """Ring of `n_buffers` pinned (seq1, seq2, scores) triples of `batch_pairs` pairs."""

    def __init__(self, ctx: "swb200.Context", batch_pairs: int = 1 << 21, n_buffers: int = 3,
                 packed: bool = False, gen_threads: int = 8):
        self.ctx = ctx
        self.batch = int(batch_pairs)
        self.packed = packed
        self.width = 32 if packed else 128
        self.gen_threads = max(1, gen_threads)
        self.bufs = []
        for _ in range(n_buffers):
            a = swb200.PinnedArray((self.batch, self.width), np.uint8)
            b = swb200.PinnedArray((self.batch, self.width), np.uint8)
            s = swb200.PinnedArray((self.batch,), np.int32)
            self.bufs.append((a, b, s))
        self.pool = ThreadPoolExecutor(max_workers=1)       # generates one batch ahead of the GPU
        self.finisher = ThreadPoolExecutor(max_workers=1)   # waits for tickets and checksums finished batches, in order

    def close(self):
        self.pool.shutdown(wait=True)
        self.finisher.shutdown(wait=True)
        for a, b, s in self.bufs:
            a.free(); b.free(); s.free()
        self.bufs = []

    def _produce(self, slot: int, first: int, m: int, seed: int) -> float:
        a, b, _ = self.bufs[slot]
        t = time.perf_counter()
        swb200.counter_pairs(first, m, seed=seed, packed=self.packed, out=(a.array[:m], b.array[:m]), threads=self.gen_threads)
        return time.perf_counter() - t

    def run(self, first: int, total: int, score_matrix, gap_penalty, seed: int = 10000,
            on_batch: Optional[Callable[[int, np.ndarray], None]] = None) -> StreamReport:
        """Scores pairs [first, first+total).  `on_batch(batch_first_index, scores_view)` is called
        (on the finisher thread, in batch order) with each finished batch; the view is only valid
        during the call.  Three stages overlap: generation of batch i+1, GPU pipeline of batch i,
        checksum/callback of batch i-1."""
        rep = StreamReport()
        nb = len(self.bufs)
        starts = list(range(first, first + total, self.batch))
        sizes = [min(self.batch, first + total - s0) for s0 in starts]
        done = {}    # slot -> future of the finisher for the batch that last used it
        t0 = time.perf_counter()
        fut = self.pool.submit(self._produce, 0, starts[0], sizes[0], seed) if starts else None
        for i, (s0, m) in enumerate(zip(starts, sizes)):
            slot = i % nb
            rep.produce_s += fut.result()
            if i + 1 < len(starts):     # the next batch's slot must have been fully consumed before it is overwritten
                nslot = (i + 1) % nb
                if nslot in done:
                    t = time.perf_counter()
                    done.pop(nslot).result()
                    rep.wait_s += time.perf_counter() - t
                fut = self.pool.submit(self._produce, nslot, starts[i + 1], sizes[i + 1], seed)
            a, b, s = self.bufs[slot]
            ticket = self.ctx.submit(a.array[:m], b.array[:m], score_matrix, gap_penalty, s.array[:m], packed=self.packed)
            done[slot] = self.finisher.submit(self._finish, (ticket, s0, m), slot, rep, on_batch)
        t = time.perf_counter()
        for f in done.values():
            f.result()
        rep.wait_s += time.perf_counter() - t
        rep.wall_s = time.perf_counter() - t0
        return rep

    def _finish(self, job, slot, rep: StreamReport, on_batch):
        ticket, s0, m = job
        self.ctx.wait(ticket)
        scores = self.bufs[slot][2].array[:m]
        rep.pairs += m
        rep.batches += 1
        rep.bytes_h2d += 2 * m * self.width
        rep.bytes_d2h += 4 * m
        rep.score_sum += int(scores.sum(dtype=np.int64))
        h = swb200.fnv1a64(scores)
        rep.fnv_of_batch_fnvs = ((rep.fnv_of_batch_fnvs ^ h) * 1099511628211) & 0xFFFFFFFFFFFFFFFF   # checksum of checksums
        if on_batch is not None:
            on_batch(s0, scores)
