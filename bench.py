#!/usr/bin/env python
"""bench.py -- the hot path's benchmark (BASELINE.json: GCUPS and alignments/s on the
1 M-pair batch).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one pass of the kernel over one batch of 1 000 000 pairs per GPU (weak scaling:
rank r scores its own 1 M pairs; no collective is on the data path).  Rank 0's batch is the
reference's own seeded stream (source.cpp:2944-2953), so every run re-checks the reference's
checksum (FNV-1a-64 ae56a1e6a1d57492) on the scores it just timed.

Prints ONE JSON line (rank 0):
  value      whole-job GCUPS with inputs resident in HBM (cells = pairs * 128 * 128)
  e2e        the same metric through the C-ABI host call swb200_score_batch with pinned HOST
             buffers: H2D of both sequence arrays and D2H of the scores inside the timed region
  roofline   integer-ALU issue roofline of the dominant kernel, measured live
             with CUDA events on the launching stream; `hbm` sub-object = the sequence stream
  cpu_baseline  the reference's simd4 (oracle/_ref) or the oracle port, timed on the host cores
`--impl reference` times the reference's own CPU path (all host threads) on the same config.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # before torch creates the CUDA context: see smith-waterman-simd_b200/swb200.py

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "smith-waterman-simd_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

CELLS_PER_PAIR = 128 * 128
PAIRS_PER_GPU = 1_000_000
ALGO_INSTR_PER_CELL = 2.0        # 4 packed int16x2 instructions per 2 cells
ALGO_BYTES_PER_PAIR = 128 + 128 + 4


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self, t0: float, t1: float) -> dict:
        sm, mx, pw, reasons = [], [], [], set()
        rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-3:]
        for _, line in rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- peaks
def load_peaks() -> dict:
    peaks = {"hbm_gbs": 6650.0, "hbm_src": "fallback (B200_PROFILING.md)"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        peaks["hbm_gbs"] = float(m["hbm_gbs"]); peaks["hbm_src"] = "measured (MEASURED_PEAKS.json)"
        peaks["sm_max_mhz"] = float(m.get("sm_max_mhz", 1965.0))
    except Exception:
        pass
    # integer-ALU issue peak: measured by tools/pipebench.cu on this pool (profiles/INT_PEAK.json)
    try:
        with open(os.path.join(ROOT, "profiles", "INT_PEAK.json")) as f:
            ip = json.load(f)
        peaks["alu_lanes_per_clk_per_sm"] = float(ip["alu_lanes_per_clk_per_sm"])
        peaks["alu_src"] = "measured (profiles/INT_PEAK.json, tools/pipebench.cu)"
    except Exception:
        peaks["alu_lanes_per_clk_per_sm"] = 64.0
        peaks["alu_src"] = "nominal 16 lanes/clk/SMSP (no profiles/INT_PEAK.json)"
    return peaks


# --------------------------------------------------------------------------- reference arm / CPU baseline
WORKLOAD = ("configs[1]: 1M seeded random 128-mer pairs per GPU (rank 0 = reference stream source.cpp:2944-2953; rank r = counter stream "
            "pairs [r*1M,(r+1)*1M)), matrix +10/-30, gap 15")
MATRIX_SPEEDTEST = (10, -30, -30, -30, -30, 10, -30, -30, -30, -30, 10, -30, -30, -30, -30, 10)   # source.cpp:3041-3045
GAP_SPEEDTEST = 15                                                                            # source.cpp:3046


def cpu_reference_run(a, b, matrix, gap, steps, warmup, budget_s=150.0):
    """Times the reference's own AVX2 path (oracle/_ref, the unmodified source) -- or the oracle port when that build is
    absent -- with all host threads.  simd4 is the README's reference point, simd9 its fastest variant on most CPUs
    (README.md:12): BOTH are timed on 200 000 pairs and the faster one runs the measurement, so the baseline is the
    reference at its best on THIS box.  Returns (kind, variant, cores, ms_per_step, sample_pairs, probe)."""
    from oracle import oracle as O   # allowed here: cpu_baseline / --impl reference legs only
    cores = os.cpu_count() or 1
    n = a.shape[0]
    probe_n = min(n, 200_000)
    probe = {}
    if O.have_ref():
        best = None
        for v, nm in ((4, "SmithWaterman_simd4 (source.cpp:462-571)"), (9, "SmithWaterman_simd9 (source.cpp:953-1071)")):
            O.ref_score_batch(v, a[:20_000], b[:20_000], matrix, gap, threads=cores)
            dtv = min(_timed(lambda: O.ref_score_batch(v, a[:probe_n], b[:probe_n], matrix, gap, threads=cores)) for _ in range(2))
            probe[f"simd{v}_gcups"] = probe_n * CELLS_PER_PAIR / dtv / 1e9
            if best is None or dtv < best[0]:
                best = (dtv, v, nm)
        dt, vbest, nm = best
        kind, variant = "reference", nm + f", unmodified, g++ -O3 -mavx2 (faster of simd4/simd9 on {probe_n} pairs on this host)"
        run = lambda m: O.ref_score_batch(vbest, a[:m], b[:m], matrix, gap, threads=cores)
    else:
        O.build()
        kind, variant = "port", "oracle/sw_oracle.c scalar restatement of source.cpp:35-60, gcc -O2"
        run = lambda m: O.score_batch(a[:m], b[:m], matrix, gap, threads=cores)
        probe_n = min(n, 20_000)
        dt = _timed(lambda: run(probe_n))
    per_pair = dt / probe_n
    # a step is the whole batch unless the whole run would blow the budget
    sample = n
    total = (steps + warmup) * n * per_pair
    if total > budget_s:
        sample = max(1000, int(n * budget_s / total))
    for _ in range(warmup):
        run(sample)
    times = [_timed(lambda: run(sample)) for _ in range(steps)]
    return kind, variant, cores, 1e3 * sum(times) / len(times), sample, probe


def _timed(fn) -> float:
    t = time.perf_counter()
    fn()
    return time.perf_counter() - t


def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_cpu_table(args):
    """the reference's scalar, simd4, simd7 and simd9 on the 1M-pair batch
    (scalar on a 50 000-pair sample), one thread and all host cores, both harness matrices.
    `python bench.py --impl reference --cpu-table` -> one JSON line per row."""
    from oracle import oracle as O
    if not O.have_ref():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libswref.so not built (no /root/reference here and no prebuilt file)"}))
        return
    cores = os.cpu_count() or 1
    a, b = O.reference_stream(PAIRS_PER_GPU)
    names = {0: "scalar SmithWaterman (source.cpp:35-60)", 4: "SmithWaterman_simd4 (462-571)", 7: "SmithWaterman_simd7 (758-850)", 9: "SmithWaterman_simd9 (953-1071)"}
    m111 = (1, -1, -1, -1, -1, 1, -1, -1, -1, -1, 1, -1, -1, -1, -1, 1)
    for label, matrix, gap in (("+10/-30 gap 15 (SpeedTest, source.cpp:3041-3046)", MATRIX_SPEEDTEST, 15),
                               ("+1/-1 gap 1 (speedtest111x32, source.cpp:3202-3207)", m111, 1)):
        for variant in (0, 4, 7, 9):
            for threads in (1, cores):
                m = (50_000 if variant == 0 else 250_000) * (1 if threads == 1 else min(4, cores))
                m = min(m, PAIRS_PER_GPU)
                O.ref_score_batch(variant, a[:2000], b[:2000], matrix, gap, threads=threads)
                dt = _timed(lambda: O.ref_score_batch(variant, a[:m], b[:m], matrix, gap, threads=threads))
                print(json.dumps({"impl": "reference", "kernel": names[variant], "scoring": label, "threads": threads, "host_cores": cores,
                                  "cpu_model": cpu_model(), "sample_pairs": m, "ms_per_1M_pairs": dt / m * 1e9, "gcups": m * CELLS_PER_PAIR / dt / 1e9,
                                  "alignments_per_s": m / dt, "build": "g++ -std=c++17 -O3 -mavx2, unmodified source"}), flush=True)


def run_reference_arm(args):
    """The reference's own CPU implementation on the box's host cores.  Nothing of the product is loaded in this process:
    the inputs come from the oracle's restatement of the reference's generator (source.cpp:2944-2953), the scores from
    oracle/_ref (the unmodified source.cpp)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.cpu_table:
        return run_cpu_table(args)
    from oracle import oracle as O
    O.build()
    a, b = O.reference_stream(PAIRS_PER_GPU)
    kind, variant, cores, ms, sample, probe = cpu_reference_run(a, b, MATRIX_SPEEDTEST, GAP_SPEEDTEST, args.steps, args.warmup)
    gcups = sample * CELLS_PER_PAIR / (ms * 1e-3) / 1e9
    line = {
        "impl": "reference", "metric": "GCUPS", "value": gcups, "unit": "GCUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int16x2", "data": "synthetic",
        "alignments_per_s": sample / (ms * 1e-3),
        "config": {"workload": WORKLOAD, "pairs_per_gpu": PAIRS_PER_GPU, "pairs_per_step": sample, "cells_per_pair": CELLS_PER_PAIR, "cpu_model": cpu_model(),
                   "note": "the CPU arm scores rank 0's batch (the reference stream) whatever --gpus says; it has no GPUs to scale over"},
        "cpu_baseline": {"value": gcups, "unit": "GCUPS", "cores": cores, "kind": kind, "variant": variant, "probe_200k": probe,
                         "sample": f"{sample} of 1000000 pairs per step, {args.steps} steps, {cores} threads over contiguous index ranges"},
        "e2e": {"value": gcups, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- B200 arm: the legs
class Ranks:
    """The ranks of this run: NCCL for the device-side barrier and the reductions of timing scalars (no data-path
    collective), and a gloo group for the legs in which ranks must wait WITHOUT touching their GPU."""

    def __init__(self):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(self.world)))
        self.dist = None
        self.cpu_group = None
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
            try:
                self.cpu_group = dist.new_group(backend="gloo")
            except Exception as ex:   # the NCCL group still gives a (GPU-polling) barrier
                log(f"bench.py: no gloo group ({ex}); CPU waits fall back to the NCCL barrier")

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def cpu_barrier(self):
        """Ranks meet without any GPU work (a blocked socket read), so a rank that is still measuring is not disturbed."""
        if self.dist is None:
            return
        if self.cpu_group is not None:
            self.dist.barrier(group=self.cpu_group)
        else:
            self.dist.barrier()

    def max(self, v):
        from sharding import max_over_ranks
        return max_over_ranks(v, self.dist)

    def sum(self, v):
        from sharding import sum_over_ranks
        return sum_over_ranks(v, self.dist)

    def close(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()


def timed_host_calls(R: "Ranks", fn, steps: int, warmup: int = 3):
    """ms per call of a blocking host call, max over ranks, every rank between the same two barriers."""
    for _ in range(warmup):
        fn()
    R.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    R.barrier()
    return R.max(1e3 * (time.perf_counter() - t0) / steps)


HOST_PROBE_MB = 1024


def leg_host_ceiling(R: "Ranks", swb200, cpus: int) -> dict:
    """What the HOST can deliver, all ranks at once: (a) pinned H2D copies alone, (b) the cores' streaming reads alone,
    (c) both together.  A byte-coded batch has to leave host DRAM once -- read by a core that packs it or by the DMA
    engine -- so (c) / 256 B is the roofline of the end-to-end number for byte-coded input and (a) / 64 B that of the
    2-bit wire format.  The probe buffer (1 GiB of pinned memory per rank) is larger than any last-level cache, and every
    phase runs for a few hundred milliseconds.  Aggregates are sums over ranks of bytes / the slowest rank's time."""
    torch = R.torch
    probe_mb = HOST_PROBE_MB
    buf = swb200.PinnedArray((probe_mb << 20,), np.uint8)
    buf.array[:] = 1                           # first touch on this rank's cores
    src = torch.from_numpy(buf.array)          # pinned (cudaHostAlloc): the copies below are plain DMA
    piece = min(256, probe_mb // 2) << 20
    dst = torch.empty(piece, dtype=torch.uint8, device="cuda")
    n_piece = src.numel() // piece
    stream = torch.cuda.current_stream()

    def dma(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for k in range(reps):
            dst.copy_(src[(k % n_piece) * piece:(k % n_piece + 1) * piece], non_blocking=True)
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3
    dma(2)
    R.barrier()
    reps = 3 * n_piece                                         # 3 GiB: ~60 ms per rank alone, longer when the ranks share the host
    t = R.max(dma(reps))
    h2d = R.sum(piece * reps) / t / 1e9
    rate = swb200.host_read_bandwidth(buf.array, cpus, 1)       # one pass: warms nothing (1 GiB), gives the pass count for ~0.3 s
    passes = max(2, int(R.max(0.3 * rate / buf.nbytes)) + 1)    # the same count on every rank
    R.barrier()
    t0 = time.perf_counter()
    swb200.host_read_bandwidth(buf.array, cpus, passes)
    t = R.max(time.perf_counter() - t0)
    cpu_read = R.sum(buf.nbytes * passes) / t / 1e9
    # both at once: the DMA loop runs on a second thread until the cores have finished their passes
    stop = threading.Event()
    ready = threading.Event()
    copied = [0]

    def dma_loop():
        torch.cuda.set_device(R.local_rank)
        s2 = torch.cuda.Stream()
        with torch.cuda.stream(s2):
            k = 0
            ready.set()
            while not stop.is_set():
                dst.copy_(src[(k % n_piece) * piece:(k % n_piece + 1) * piece], non_blocking=True)
                s2.synchronize()
                copied[0] += piece
                k += 1
    th = threading.Thread(target=dma_loop)
    th.start()
    ready.wait()
    R.barrier()
    c0 = copied[0]
    t0 = time.perf_counter()
    swb200.host_read_bandwidth(buf.array, max(1, cpus - 1), passes)
    dt = time.perf_counter() - t0
    c1 = copied[0]
    stop.set()
    th.join()
    tmax = R.max(dt)
    both = R.sum(buf.nbytes * passes + (c1 - c0)) / tmax / 1e9
    both_dma = R.sum(c1 - c0) / tmax / 1e9
    del dst, src
    buf.free()
    return {"h2d_pinned_gbs": h2d, "host_read_gbs": cpu_read, "h2d_and_read_together_gbs": both, "h2d_share_of_together_gbs": both_dma,
            "threads_per_rank": cpus, "buffer_mb_per_rank": probe_mb, "read_passes": passes,
            "how": "all ranks at once between barriers; sum of bytes over ranks / slowest rank's time: torch pinned->device copies of 256 MiB slices "
                   "(CUDA events), swb200_host_read_bandwidth (AVX2 streaming reads on all of the rank's cores), then both concurrently"}


def ceiling_gcups(bytes_per_s_g: float, bytes_per_pair: float) -> float:
    return bytes_per_s_g * 1e9 / bytes_per_pair * CELLS_PER_PAIR / 1e9


def leg_stream(R: "Ranks", swb200, ctx, pairs: int, packed: bool, gen_threads: int, matrix, gap) -> dict:
    """BASELINE.json configs[2] + configs[4]: the index space [0, pairs) of the counter stream split in contiguous ranges
    over the ranks (shard_range); every rank generates its pairs on host threads into pinned ring buffers and streams
    them through swb200_submit[_packed].  End-to-end alignments/s = pairs / the slowest rank's wall time."""
    from sharding import shard_range
    from streaming import StreamRunner
    lo, hi = shard_range(pairs, R.rank, R.world)
    batch = 1 << 21
    runner = StreamRunner(ctx, batch_pairs=batch, n_buffers=3, packed=packed, gen_threads=gen_threads)
    runner.run(lo, min(hi - lo, 2 * batch), matrix, gap)          # warm-up: first touch of the pinned ring, staging allocation
    R.barrier()
    launches0 = ctx.launch_count
    rep = runner.run(lo, hi - lo, matrix, gap)
    R.torch.cuda.synchronize()
    wall = R.max(rep.wall_s)
    total = R.sum(rep.pairs)
    score_sum = int(R.sum(rep.score_sum))
    gen_s = R.max(rep.produce_s)
    wait_s = R.max(rep.wait_s)
    launches = R.sum(ctx.launch_count - launches0)
    h2d, d2h = R.sum(rep.bytes_h2d), R.sum(rep.bytes_d2h)
    runner.close()
    if gen_s > 0.8 * wall:
        bottleneck = "host generation (the producer threads were busy for > 80 % of the wall time)"
    elif wait_s > 0.5 * wall:
        bottleneck = "PCIe / kernel pipeline (the driver thread was blocked on in-flight batches for > 50 % of the wall time)"
    else:
        bottleneck = "mixed: host generation and the GPU pipeline alternate"
    return {"alignments_per_s": total / wall, "gcups": total * CELLS_PER_PAIR / wall / 1e9, "wall_s": wall, "pairs": int(total),
            "wire_format": "2-bit packed (source.cpp:1580-1583), 64 B/pair" if packed else "byte codes, 256 B/pair",
            "batch_pairs": batch, "gen_threads_per_rank": gen_threads, "h2d_bytes": int(h2d), "d2h_bytes": int(d2h), "gpu_launches": int(launches),
            "breakdown": {"host_generation_s_max_rank": gen_s, "blocked_on_gpu_pipeline_s_max_rank": wait_s, "bottleneck": bottleneck},
            "score_sum": score_sum, "score_sum_equals_reference": stream_sum_check(pairs, score_sum)}


def leg_inproc(R: "Ranks", swb200, n_per_gpu: int, matrix, gap, steps: int, restore_affinity) -> dict:
    """The library's OWN sharding layer (north_star (c)): rank 0 alone opens one context on all N GPUs and pushes N x 1M
    pairs through ONE swb200_score_batch call (contiguous index ranges, one host thread per GPU, host gather by direct
    stores), byte-coded and 2-bit packed.  The other ranks wait on a CPU barrier; their GPUs are idle."""
    out = None
    if R.rank == 0:
        try:
            before = sorted(os.sched_getaffinity(0))
            if restore_affinity:
                os.sched_setaffinity(0, restore_affinity)     # this leg owns the whole box
            G = R.world
            n = G * n_per_gpu
            ctxN = swb200.Context(devices=list(range(G)))
            pa, pb = swb200.PinnedArray((n, 128), np.uint8), swb200.PinnedArray((n, 128), np.uint8)
            ka, kb = swb200.PinnedArray((n, 32), np.uint8), swb200.PinnedArray((n, 32), np.uint8)
            ps = swb200.PinnedArray((n,), np.int32)
            swb200.counter_pairs(0, n, out=(pa.array, pb.array))
            swb200.counter_pairs(0, n, packed=True, out=(ka.array, kb.array))
            want = counter_prefix_sum(n)
            res = {}
            for name, (x, y, packed) in (("bytes", (pa, pb, False)), ("packed", (ka, kb, True))):
                for _ in range(3 if packed else 10):     # byte-coded: the lane tuner's exploration calls come first
                    ctxN.score_batch(x.array, y.array, matrix, gap, out=ps.array, packed=packed)
                l0 = ctxN.launch_count
                t0 = time.perf_counter()
                for _ in range(steps):
                    ctxN.score_batch(x.array, y.array, matrix, gap, out=ps.array, packed=packed)
                ms = 1e3 * (time.perf_counter() - t0) / steps
                ssum = int(ps.array.sum(dtype=np.int64))
                res[name] = {"value": n * CELLS_PER_PAIR / (ms * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": ms, "alignments_per_s": n / (ms * 1e-3),
                             "h2d_bytes_per_step": 2 * n * (32 if packed else 128), "d2h_bytes_per_step": 4 * n,
                             "gpu_launches_per_step": (ctxN.launch_count - l0) / steps,
                             "score_sum_equals_reference": (None if want is None else bool(ssum == want))}
            res["bytes"]["host_pack"] = dict(ctxN.host_pack_stats(), auto_tuner=[ctxN.host_pack_tuning(k) for k in range(G)])
            out = {"api": f"ONE swb200_score_batch[_packed] call per step on a context of {G} GPUs, {n} pairs in pinned host arrays, scores gathered into one pinned host array",
                   "n_gpus": G, "pairs_per_step": n, "steps": steps, "process": "rank 0 alone; the other ranks wait on a gloo barrier with idle GPUs", **res}
            ctxN.close()
            for p in (pa, pb, ka, kb, ps):
                p.free()
            os.sched_setaffinity(0, before)
        except Exception as ex:   # an extra leg: report and carry on, and never leave the other ranks at the barrier
            out = {"error": f"{type(ex).__name__}: {ex}"}
    R.cpu_barrier()
    return out


def counter_prefix_sum(n: int):
    """Sum of the reference's scores over counter-stream pairs [0, n) at +10/-30/15, from the committed golden file."""
    try:
        with open(os.path.join(ROOT, "tests", "golden", "counter_stream_sums.json")) as f:
            g = json.load(f)
        v = g["sum_of_scores_over_prefix"]["speedtest_10_-30_15"].get(str(n))
        if v is None and n % 1_000_000 == 0:
            blocks = g["sum_of_scores_block_1M"]["speedtest_10_-30_15"]
            if all(str(k) in blocks for k in range(n // 1_000_000)):
                v = sum(int(blocks[str(k)]) for k in range(n // 1_000_000))
        return None if v is None else int(v)
    except (OSError, KeyError, ValueError):
        return None


def leg_per_pair(swb200, ctx, a, b, matrix, gap, calls: int = 10_000) -> dict:
    """The reference's literal metric shape (SpeedTest, source.cpp:3036-3054): ONE fixed pair scored again and again,
    "ms / 1M".  Here: microseconds per swb200_score_pair call on the first pair of the reference stream, timed around a
    loop of direct C-ABI calls (ctypes overhead included)."""
    import ctypes as C
    lib, h = ctx._lib, ctx._h
    m = np.asarray(matrix, dtype=np.int8)
    s1, s2 = np.ascontiguousarray(a[0]), np.ascontiguousarray(b[0])
    out = np.zeros(1, np.int32)
    fn = lib.swb200_score_pair
    args = (h, s1.ctypes.data, s2.ctypes.data, m.ctypes.data, C.c_int8(int(gap)), out.ctypes.data)
    for _ in range(200):
        fn(*args)
    l0 = ctx.launch_count
    t0 = time.perf_counter()
    for _ in range(calls):
        fn(*args)
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    for _ in range(calls):
        pass
    loop_overhead = time.perf_counter() - t1
    try:
        pstats = ctx.pair_path_stats()     # the resident server kernel: launched once, rung `calls` times
    except Exception:
        pstats = None
    res = {"us_per_call": 1e6 * dt / calls, "ms_per_1M_calls": 1e3 * dt / calls * 1e6, "calls": calls, "score": int(out[0]), "score_expected": 80,
           "gpu_launches_per_call": (ctx.launch_count - l0) / calls, "resident_server": pstats, "python_loop_overhead_us": 1e6 * loop_overhead / calls,
           "timed_from": "a Python loop of direct ctypes calls of swb200_score_pair",
           "api": "swb200_score_pair: the call writes the pair into a 320-byte doorbell in mapped pinned memory; a resident one-warp server kernel polls it across PCIe, scores the pair and stores a tagged word into mapped pinned memory the call spins on (no launch per call; the server leaves by itself after 200 us without a call)",
           "shape": "SpeedTest (source.cpp:3036-3054): one fixed pair, repeated calls"}
    # the same loop in C++ (tools/speedtest_b200.cu, built by __graft_entry__.build()): no interpreter in the timed region,
    # and beside it the floor of any per-call GPU path on this box (empty kernel + tagged mapped word + spin)
    exe = os.path.join(ROOT, "tools", "speedtest_b200")
    if os.path.exists(exe):
        try:
            p = subprocess.run([exe, "20000"], capture_output=True, text=True, timeout=120)
            res["cpp_loop"] = json.loads(p.stdout.strip().splitlines()[-1])
        except Exception as ex:
            res["cpp_loop"] = {"error": f"{type(ex).__name__}: {ex}"}
    return res


SG_LEG_PAIRS, SG_LEG_LEN = 148 * 256, 16384      # the semi-global leg of the default line (the dry-run test shrinks it)


def leg_semiglobal(R: "Ranks", swb200, ctx, steps: int = 5) -> dict:
    """SURVEY.md 8(f4) in the default line: the adaptive-banded X-drop semi-global aligner (score + traceback) on 37888
    pairs of 16384-mers with TestSemiGlobal's 10/10/10 % edits (source.cpp:2750-2771) -- device-resident, and end to end
    through swb200_semiglobal_xdrop_batch with host arrays (both directions cross the link four to a byte, csrc/sg_pipe.inc);
    the whole batch is compared with the oracle's committed sums (tests/golden/semiglobal_batch_sums.json).
    `bench.py --workload semiglobal` is the full arm (roofline of the forward kernel, CPU reference beside it)."""
    torch = R.torch
    n, L = SG_LEG_PAIRS, SG_LEG_LEN
    pa, pb = swb200.PinnedArray((n, L), np.uint8), swb200.PinnedArray((n, L), np.uint8)
    swb200.related_pairs(0, n, L, out=(pa.array, pb.array))
    h_meta = [np.zeros(n, np.int32) for _ in range(4)]
    h_ops = swb200.PinnedArray((n, 2 * L), np.uint8)
    d_a, d_b = torch.from_numpy(pa.array).cuda(), torch.from_numpy(pb.array).cuda()
    d_meta = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(4)]
    d_ops = torch.empty((n, 2 * L), dtype=torch.uint8, device="cuda")
    l0 = ctx.launch_count
    for _ in range(2):
        ctx.semiglobal_xdrop_device(d_a, d_b, d_meta[0], d_meta[1], d_meta[2], d_meta[3], d_ops)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ctx.semiglobal_xdrop_device(d_a, d_b, d_meta[0], d_meta[1], d_meta[2], d_meta[3], d_ops)
    e1.record()
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1) / steps
    dev_scores, dev_nops = d_meta[0].cpu().numpy(), d_meta[3].cpu().numpy()
    del d_ops, d_a, d_b
    torch.cuda.empty_cache()

    def e2e():
        ctx._check(ctx._lib.swb200_semiglobal_xdrop_batch(ctx._h, pa.array.ctypes.data, pb.array.ctypes.data, L, n, h_meta[0].ctypes.data,
                                                          h_meta[1].ctypes.data, h_meta[2].ctypes.data, h_meta[3].ctypes.data, h_ops.array.ctypes.data))
    for _ in range(2):
        e2e()
    l1 = ctx.launch_count
    per_call = []
    for _ in range(steps):
        t0 = time.perf_counter()
        e2e()
        per_call.append(1e3 * (time.perf_counter() - t0))
    e2e_ms = sum(per_call) / steps
    launches_e2e = (ctx.launch_count - l1) / steps
    sums_ok = None
    try:
        with open(os.path.join(ROOT, "tests", "golden", "semiglobal_batch_sums.json")) as f:
            want = json.load(f)["prefix"].get(str(n))
        if want is not None:
            got = {k: int(h_meta[j].sum(dtype=np.int64)) for j, k in enumerate(("score", "end_y", "end_x", "n_ops"))}
            got["ops_weighted"] = sg_weighted_ops_sum(h_ops.array, h_meta[3])
            sums_ok = bool(got == {k: int(v) for k, v in want.items()})
    except Exception as ex:
        sums_ok = f"{type(ex).__name__}: {ex}"
    return {"workload": f"SURVEY.md 8(f4): adaptive-banded X-drop semi-global aligner (band 32, X 70, 1/1/1), score + traceback, {n} pairs of {L}-mers, "
                        "10/10/10 % mismatch/insert/delete (source.cpp:2750-2771)",
            "pairs": n, "seq_len": L,
            "device_resident": {"alignments_per_s": n / (dev_ms * 1e-3), "ms_per_step": dev_ms, "gpu_launches_per_step": 3,
                                "api": "swb200_semiglobal_xdrop_batch_device (forward, traceback, left-align kernels)"},
            "e2e": {"alignments_per_s": n / (e2e_ms * 1e-3), "ms_per_step": e2e_ms, "ms_per_call": [round(t, 2) for t in per_call], "gpu_launches_per_step": launches_e2e,
                    "host_bytes_in_per_step": 2 * n * L, "host_bytes_out_per_step": n * (16 + 2 * L), "h2d_bytes_per_step": n * L // 2, "d2h_bytes_per_step": n * (16 + L // 2),
                    "api": "swb200_semiglobal_xdrop_batch (host byte arrays in; scores, end cells and move strings out; host lanes pack the sequences to 2 bits "
                           "and expand the 2-bit move strings, chunks of one forward warp per SM round 16 slots)",
                    "bound": "latency: a forward warp is a serial chain of 32768 rounds (15-18 ms) that cannot start before its chunk has been packed and copied"},
            "verified": {"e2e_scores_and_lengths_equal_device": bool(np.array_equal(h_meta[0], dev_scores) and np.array_equal(h_meta[3], dev_nops)),
                         "whole_batch_sums_equal_oracle": sums_ok}}


def leg_sweep(R: "Ranks", swb200, ctx, matrix, gap, steps: int, peak_tinstr: float) -> list:
    """BASELINE.json configs[3], in the default line: L = 128 / 256 / 512 on whole-wave batches of the counter stream
    re-cut to length L, device-resident, each with the oracle's committed score sum."""
    torch = R.torch
    rows = []
    for L in swb200.SWEEP_LENGTHS:
        info = ctx.kernel_info(matrix, gap, seq_len=L)
        n = sweep_pairs(L, info)
        h_a, h_b = swb200.counter_pairs(0, n * (L // 128))
        d_a = torch.from_numpy(h_a.reshape(n, L)).cuda()
        d_b = torch.from_numpy(h_b.reshape(n, L)).cuda()
        del h_a, h_b
        d_s = torch.empty(n, dtype=torch.int32, device="cuda")
        for _ in range(3):
            ctx.score_batch_device(d_a, d_b, matrix, gap, d_s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            ctx.score_batch_device(d_a, d_b, matrix, gap, d_s)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        gcups = n * L * L / (ms * 1e-3) / 1e9
        score_sum = int(d_s.sum(dtype=torch.int64).item())
        sum_ok = None
        try:
            with open(os.path.join(ROOT, "tests", "golden", "sweep_sums.json")) as f:
                want = json.load(f)["by_length"][str(L)]["sum_of_scores_by_pairs"].get(str(n))
            if want is not None:
                sum_ok = bool(int(want) == score_sum)
        except (OSError, KeyError, ValueError):
            pass
        rows.append({"seq_len": L, "pairs": n, "waves": n / max(1, info["sm_count"] * info["blocks_per_sm"] * info["threads_per_block"] * 2),
                     "ms_per_launch": ms, "gcups": gcups, "alignments_per_s": n / (ms * 1e-3),
                     "roofline_frac": gcups * 1e9 * ALGO_INSTR_PER_CELL / 1e12 / peak_tinstr, "vs_L128": None,
                     "score_sum": score_sum, "score_sum_equals_oracle": sum_ok,
                     "kernel": {k: info[k] for k in ("regs_per_thread", "threads_per_block", "blocks_per_sm", "smem_bytes_per_block")}})
        del d_a, d_b, d_s
    for r in rows:
        r["vs_L128"] = r["gcups"] / rows[0]["gcups"]
    return rows


SASS_ALU_PER_STEP = 57.3          # ALU-pipe instructions per anti-diagonal step (16 words = 32 cells of two pairs) of the shipped L = 128 kernel,
SASS_CELLS_RATIO = 16640 / 16384  # counted from the built library by tools/sass_hist.py (profiles/r02/sass_hot_loops.json); steps x 16 / real cells


# --------------------------------------------------------------------------- B200 arm
def run_b200_arm(args):
    import swb200
    R = Ranks()
    torch = R.torch
    world, rank, local_rank = R.world, R.rank, R.local_rank
    matrix, gap = swb200.MATRIX_SPEEDTEST, swb200.GAP_SPEEDTEST
    # several ranks on one box: each gets its own share of the cores (the library sizes its PACK-lane pool from the CPUs the
    # process may run on), on the NUMA node its GPU hangs off; pinned buffers are first-touched after this
    cpu_share = swb200.bind_rank_cpus(local_rank, R.local_world) if world > 1 else {"before": sorted(os.sched_getaffinity(0)), "numa": None,
                                                                                     "cpus": len(os.sched_getaffinity(0)), "share": None}
    my_cpus = cpu_share["cpus"]
    ctx = swb200.Context(devices=[local_rank])
    if args.pack_threads is not None:
        ctx.set_host_pack_threads(args.pack_threads)
    info = ctx.kernel_info(matrix, gap)
    n = PAIRS_PER_GPU

    # ---- this rank's batch, in PINNED host memory (the e2e legs copy from here every step)
    pa, pb = swb200.PinnedArray((n, 128), np.uint8), swb200.PinnedArray((n, 128), np.uint8)
    ps = swb200.PinnedArray((n,), np.int32)
    if rank == 0:
        swb200.reference_stream(n, out=(pa.array, pb.array))        # the reference's own stream
    else:
        swb200.counter_pairs(rank * n, n, out=(pa.array, pb.array))  # index range [rank*n, (rank+1)*n)

    d_a = torch.from_numpy(pa.array).cuda(non_blocking=False)
    d_b = torch.from_numpy(pb.array).cuda(non_blocking=False)
    d_s = torch.empty(n, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream()

    # ================= leg 1: device-resident (value, roofline)
    for _ in range(args.warmup):
        ctx.score_batch_device(d_a, d_b, matrix, gap, d_s)
    sampler = ClockSampler(local_rank)
    sampler.start()
    R.barrier()
    launches0 = ctx.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    t_wall0 = time.perf_counter()
    ev[0].record(stream)
    for i in range(args.steps):
        ctx.score_batch_device(d_a, d_b, matrix, gap, d_s)   # 256 MB of sequence data per launch: larger than the 126 MB L2
        ev[i + 1].record(stream)
    R.barrier()
    t_wall1 = time.perf_counter()
    launches_dev = ctx.launch_count - launches0
    ms_total = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    ms_step = R.max(ms_total / args.steps)
    total_pairs = R.sum(n)
    gcups = total_pairs * CELLS_PER_PAIR / (ms_step * 1e-3) / 1e9
    # the roofline's denominator, measured on this GPU right after the timed launches (clocks are still up)
    try:
        peak_live = ctx.measure_alu_peak(target_ms=50.0)
    except Exception as ex:
        peak_live = {"error": f"{type(ex).__name__}: {ex}"}

    scores = d_s.cpu().numpy()
    verified = None
    if rank == 0:
        verified = (f"{swb200.fnv1a64(scores):016x}" == "ae56a1e6a1d57492") and int(scores.sum()) == 75_478_815
    # the same batch resident in HBM in the 2-bit layout: the persistent consumer kernel with every tile already there
    # (expansion fused into the scoring launch) -- what the end-to-end packed leg could reach if PCIe cost nothing
    packed_resident = None
    try:
        d_ka = torch.from_numpy(swb200.pack2bit(pa.array)).cuda()
        d_kb = torch.from_numpy(swb200.pack2bit(pb.array)).cuda()
        d_s2 = torch.empty(n, dtype=torch.int32, device="cuda")
        for _ in range(3):
            ctx.score_batch_device(d_ka, d_kb, matrix, gap, d_s2, n=n, packed=True)
        l0 = ctx.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            ctx.score_batch_device(d_ka, d_kb, matrix, gap, d_s2, n=n, packed=True)
        e1.record(stream)
        torch.cuda.synchronize()
        pr_ms = e0.elapsed_time(e1) / args.steps
        packed_resident = {"ms_per_launch": pr_ms, "gcups": n * CELLS_PER_PAIR / (pr_ms * 1e-3) / 1e9, "gpu_launches_per_step": (ctx.launch_count - l0) / args.steps,
                           "scores_equal": bool(np.array_equal(d_s2.cpu().numpy(), scores)),
                           "api": "swb200_score_batch_packed_device: sw_feed_kernel on 2-bit arrays resident in HBM (this rank)"}
        del d_ka, d_kb, d_s2
    except Exception as ex:
        packed_resident = {"error": f"{type(ex).__name__}: {ex}"}
    # ranks >= 1 score counter-stream pairs [rank*1M, (rank+1)*1M): their score sum against the value computed with the
    # unmodified reference (tests/golden/counter_stream_sums.json).  The two reductions run on EVERY rank.
    block_state = 0.0        # 1.0 = checked and equal, -1.0 = checked and different, 0.0 = no golden entry / rank 0
    if rank > 0:
        try:
            with open(os.path.join(ROOT, "tests", "golden", "counter_stream_sums.json")) as f:
                want = json.load(f)["sum_of_scores_block_1M"]["speedtest_10_-30_15"].get(str(rank))
            if want is not None and n == 1_000_000:
                block_state = 1.0 if int(scores.sum(dtype=np.int64)) == int(want) else -1.0
        except (OSError, KeyError, ValueError):
            pass
    blocks_equal = R.sum(1.0 if block_state > 0 else 0.0)
    blocks_differ = R.sum(1.0 if block_state < 0 else 0.0)

    # ================= leg 2: end to end through the C-ABI host call, byte-coded input (the reference's own layout)
    def e2e_bytes():
        ctx.score_batch(pa.array, pb.array, matrix, gap, out=ps.array)   # H2D + kernel + scores back; returns when they are on the host
    plain_ms = None
    if world == 1 and not args.no_plain_e2e:
        # the same call with the packing lanes off (every byte crosses PCIe), for the record
        ctx.set_host_pack_threads(0)
        plain_ms = timed_host_calls(R, e2e_bytes, 5, warmup=2)
        ctx.set_host_pack_threads(args.pack_threads if args.pack_threads is not None else -1)
    for _ in range(10):          # the library tries its four lane configurations (twice each) on the first calls and then keeps the fastest
        e2e_bytes()
    pack0 = ctx.host_pack_stats()
    launches1 = ctx.launch_count
    e2e_ms = timed_host_calls(R, e2e_bytes, args.steps, warmup=1)
    launches_e2e = (ctx.launch_count - launches1) * args.steps / float(args.steps + 1)
    pack1 = ctx.host_pack_stats()
    packed_frac = (pack1["packed_pairs"] - pack0["packed_pairs"]) / float(n * (args.steps + 1))
    try:
        tuning = ctx.host_pack_tuning()
    except Exception as ex:
        tuning = {"error": f"{type(ex).__name__}: {ex}"}
    e2e_gcups = total_pairs * CELLS_PER_PAIR / (e2e_ms * 1e-3) / 1e9
    e2e_ok = bool(np.array_equal(ps.array, scores))
    all_ok = R.sum(1.0 if e2e_ok else 0.0) == world

    # ================= leg 3: the same call fed with the reference's 2-bit wire format (source.cpp:1580-1583), at EVERY N
    pka, pkb = swb200.PinnedArray((n, 32), np.uint8), swb200.PinnedArray((n, 32), np.uint8)
    psp = swb200.PinnedArray((n,), np.int32)
    pka.array[...] = swb200.pack2bit(pa.array)
    pkb.array[...] = swb200.pack2bit(pb.array)

    def e2e_pk():
        ctx.score_batch(pka.array, pkb.array, matrix, gap, out=psp.array, packed=True)
    launches2 = ctx.launch_count
    pk_ms = timed_host_calls(R, e2e_pk, args.steps, warmup=3)
    launches_pk = (ctx.launch_count - launches2) / float(args.steps + 3)
    pk_ok = R.sum(1.0 if np.array_equal(psp.array, scores) else 0.0) == world
    clocks = sampler.summary(t_wall0, t_wall1)
    sampler.stop()
    e2e_packed = {"value": total_pairs * CELLS_PER_PAIR / (pk_ms * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": pk_ms,
                  "alignments_per_s": total_pairs / (pk_ms * 1e-3), "fraction_of_device_resident_rate": ms_step / pk_ms,
                  "h2d_bytes_per_step": 2 * n * 32, "d2h_bytes_per_step": 4 * n, "gpu_launches_per_step": launches_pk,
                  "api": "swb200_score_batch_packed (pinned host arrays [n][32], 2 bits per base; one persistent kernel consumes the tiles as the copy engine lands them)",
                  "scores_equal_device_leg": bool(pk_ok)}

    # ================= leg 4: what the host can deliver (the roofline of the two numbers above)
    host_ceiling = None
    if not args.quick:
        try:
            host_ceiling = leg_host_ceiling(R, swb200, my_cpus)
            host_ceiling["byte_input_ceiling_gcups"] = ceiling_gcups(host_ceiling["h2d_and_read_together_gbs"], 256.0)
            host_ceiling["byte_input_e2e_fraction_of_ceiling"] = e2e_gcups / host_ceiling["byte_input_ceiling_gcups"]
            host_ceiling["packed_input_pcie_ceiling_gcups"] = ceiling_gcups(host_ceiling["h2d_pinned_gbs"], 64.0)
            host_ceiling["packed_input_bound"] = ("kernel" if host_ceiling["packed_input_pcie_ceiling_gcups"] > gcups else "host -> device link (PCIe / host DRAM)")
        except Exception as ex:
            host_ceiling = {"error": f"{type(ex).__name__}: {ex}"}
            R.barrier()

    # ================= leg 5: configs[2] / configs[4] -- 100 M pairs sharded by index range, streamed from host generators
    stream_legs = None
    if not args.quick:
        stream_legs = {}
        gen_threads = max(1, my_cpus - 1)
        for name, packed in (("packed", True), ("bytes", False)):
            try:
                stream_legs[name] = leg_stream(R, swb200, ctx, args.stream_pairs, packed, gen_threads, matrix, gap)
            except Exception as ex:
                stream_legs[name] = {"error": f"{type(ex).__name__}: {ex}"}
                R.barrier()

    # ================= leg 6: the library's own multi-GPU sharding layer, in ONE process (N > 1)
    inproc = None
    if world > 1 and not args.quick:
        inproc = leg_inproc(R, swb200, n, matrix, gap, steps=min(args.steps, 10), restore_affinity=cpu_share["before"])

    # ================= N = 1 only: length sweep, per-pair call, CPU baseline
    sweep = per_pair = semiglobal = None
    peak_t = peak_live.get("tinstr_per_s") if isinstance(peak_live, dict) else None
    if world == 1 and not args.quick:
        try:
            sweep = leg_sweep(R, swb200, ctx, matrix, gap, steps=max(5, args.steps // 2), peak_tinstr=peak_t or 18.46)
        except Exception as ex:
            sweep = {"error": f"{type(ex).__name__}: {ex}"}
        try:
            per_pair = leg_per_pair(swb200, ctx, pa.array, pb.array, matrix, gap)
        except Exception as ex:
            per_pair = {"error": f"{type(ex).__name__}: {ex}"}
        if not getattr(args, "no_semiglobal", False):
            try:
                semiglobal = leg_semiglobal(R, swb200, ctx)
            except Exception as ex:
                semiglobal = {"error": f"{type(ex).__name__}: {ex}"}

    # ---- roofline of the dominant kernel, from this rank's live CUDA-event launch times
    peaks = load_peaks()
    avg_launch_ms = sum(per_launch_ms) / len(per_launch_ms)
    sm_mhz = clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
    alu_peak_file = peaks["alu_lanes_per_clk_per_sm"] * info["sm_count"] * sm_mhz * 1e6 / 1e12
    alu_peak_tinstr = peak_t if peak_t else alu_peak_file
    achieved_tinstr = n * CELLS_PER_PAIR * ALGO_INSTR_PER_CELL / (avg_launch_ms * 1e-3) / 1e12
    executed_alu_tinstr = n * CELLS_PER_PAIR * SASS_CELLS_RATIO * (SASS_ALU_PER_STEP / 32.0) / (avg_launch_ms * 1e-3) / 1e12
    hbm_achieved = n * ALGO_BYTES_PER_PAIR / (avg_launch_ms * 1e-3) / 1e9
    ncu = {}
    try:
        with open(os.path.join(ROOT, "profiles", "NCU_SUMMARY.json")) as f:
            ncu = json.load(f)
    except Exception:
        pass
    roofline = {
        "bound": "int_alu", "kernel": "swb::sw_kernel<FAST=%d, L=128, NT=%d, MINB=%d>" % (info["fast_path"], info["threads_per_block"], info["blocks_per_sm"]),
        "achieved": achieved_tinstr, "peak": alu_peak_tinstr, "unit": "Tinstr/s (thread-level packed int16x2 ALU instructions)",
        "frac": achieved_tinstr / alu_peak_tinstr,
        "algorithmic_instr_per_cell": ALGO_INSTR_PER_CELL, "cells_per_launch": n * CELLS_PER_PAIR, "avg_launch_ms": avg_launch_ms,
        "peak_live": peak_live,
        "peak_src": ("measured live in this run by swb200_measure_alu_peak (independent VIADDMNMX.S16x2 chains on every SM, CUDA events)" if peak_t else
                     f"{peaks['alu_src']}: {peaks['alu_lanes_per_clk_per_sm']} lanes/clk/SM x {info['sm_count']} SMs x {sm_mhz:.0f} MHz"),
        "peak_from_file": {"tinstr_per_s": alu_peak_file, "src": f"{peaks['alu_src']} x {info['sm_count']} SMs x {sm_mhz:.0f} MHz (median SM clock sampled during the run)"},
        "alu_pipe_busy": {"model": executed_alu_tinstr / alu_peak_tinstr,
                          "how": f"ALU-pipe instructions the kernel EXECUTES ({SASS_ALU_PER_STEP} per step of 16 words = 32 cells by SASS count = {SASS_ALU_PER_STEP / 32:.2f} per computed cell, "
                                 f"{SASS_CELLS_RATIO:.4f} computed cells per real cell) / launch time / live peak",
                          "ncu_pct": ncu.get("alu_pipe_pct"), "ncu_src": ncu.get("source")},
        "traffic": ncu.get("dram_bytes_per_launch"),
        "hbm": {"bound": "hbm", "achieved": hbm_achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm_achieved / peaks["hbm_gbs"],
                "algorithmic_bytes_per_pair": ALGO_BYTES_PER_PAIR, "peak_src": peaks["hbm_src"],
                "note": "evidence that the sequence stream is not limiting (SURVEY.md 8d)"},
    }

    # ---- CPU baseline beside it (rank 0, N=1 only)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        kind, variant, cores, ms, sample, probe = cpu_reference_run(pa.array, pb.array, matrix, gap, steps=3, warmup=1, budget_s=40.0)
        cpu_baseline = {"value": sample * CELLS_PER_PAIR / (ms * 1e-3) / 1e9, "unit": "GCUPS", "cores": cores, "kind": kind,
                        "variant": variant, "cpu_model": cpu_model(), "probe_200k": probe,
                        "sample": f"{sample} of 1000000 pairs per pass, 3 timed passes, {cores} threads over contiguous index ranges",
                        "ms_per_1M_pairs": ms * 1e6 / sample}
        if per_pair and "us_per_call" in per_pair:
            try:
                from oracle import oracle as O
                per_pair["reference_simd4_us_per_call"] = O.ref_score_repeat(4, pa.array[0], pb.array[0], matrix, gap, 200_000) * 1e6
            except Exception as ex:
                per_pair["reference_simd4_us_per_call"] = f"{type(ex).__name__}: {ex}"

    if rank == 0:
        line = {
            "metric": "GCUPS", "value": gcups, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int16x2", "data": "synthetic",
            "alignments_per_s": total_pairs / (ms_step * 1e-3),
            "config": {"workload": WORKLOAD,
                       "pairs_per_gpu": n, "cells_per_pair": CELLS_PER_PAIR, "sharding": "contiguous index ranges, no collective",
                       "l2": "inputs 256 MB per launch > 126 MB L2, no flush needed", "host_cpus": os.cpu_count(),
                       "cpus_of_this_rank": {k: cpu_share[k] for k in ("cpus", "share", "numa")}, "kernel": info},
            "clocks": clocks,
            "e2e": {"value": e2e_gcups, "unit": "GCUPS", "ms_per_step": e2e_ms, "alignments_per_s": total_pairs / (e2e_ms * 1e-3),
                    "h2d_bytes_per_step": int(round(2 * n * (128 * (1.0 - packed_frac) + 32 * packed_frac))), "d2h_bytes_per_step": 4 * n,
                    "host_input_bytes_per_step": 2 * n * 128, "gpu_launches_per_step": launches_e2e / float(args.steps),
                    "api": "swb200_score_batch (C ABI, pinned host byte arrays [n][128] in, int32 scores out): one persistent kernel per call; the calling thread DMA-copies "
                           "raw pieces while PACK lanes compress others to 2 bits on host cores; scores are stored straight into the caller's pinned array",
                    "host_pack": {"threads_per_gpu": pack1["pack_threads_per_gpu"], "auto_tuner": tuning, "fraction_of_pairs_sent_packed": packed_frac,
                                  "plain_pipeline_ms_per_step": plain_ms,
                                  "note": "rank 0's split; wire compression only, no scoring on the host"},
                    "bound": "host memory: every byte of the 256 B/pair input is read from host DRAM once (by a packing core or the DMA engine); see host_ceiling",
                    "scores_equal_device_leg": all_ok, "packed_input": e2e_packed},
            "packed_resident": packed_resident,
            "host_ceiling": host_ceiling,
            "stream": (None if stream_legs is None else dict(stream_legs, pairs=args.stream_pairs,
                       workload=f"configs[2]/[4]: {args.stream_pairs} counter-stream pairs sharded by contiguous index range over {world} rank(s); host threads -> pinned ring -> swb200_submit[_packed]")),
            "e2e_inproc": inproc,
            "sweep": sweep, "per_pair": per_pair, "semiglobal": semiglobal,
            "gpu_launches": int(launches_dev), "gpu_launches_e2e": int(launches_e2e),
            "roofline": roofline,
            "verified": {"fnv1a64_ae56a1e6a1d57492_and_sum_75478815": verified, "e2e_equals_device": all_ok, "e2e_packed_equals_device": bool(pk_ok),
                         "other_ranks_score_sums_equal_reference": (None if blocks_equal + blocks_differ == 0 else bool(blocks_differ == 0)),
                         "other_ranks_checked": int(blocks_equal + blocks_differ)},
        }
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line), flush=True)

    ctx.close()
    R.close()


# --------------------------------------------------------------------------- streaming / 100M-pair mode
def stream_sum_check(pairs: int, score_sum: int):
    """The sum of all scores of counter-stream pairs [0, pairs) against the committed value computed with the
    reference (tests/golden/make_counter_sums.py); independent of batch size, wire format and sharding."""
    try:
        with open(os.path.join(ROOT, "tests", "golden", "counter_stream_sums.json")) as f:
            want = json.load(f)["sum_of_scores_over_prefix"]["speedtest_10_-30_15"].get(str(pairs))
    except (OSError, KeyError, ValueError):
        return None
    return None if want is None else bool(int(want) == int(score_sum))


def run_stream_arm(args):
    """BASELINE.json configs[2]/[4]: `--workload stream --pairs 100000000` -- the pair-index space
    [0, pairs) is split in contiguous ranges over the ranks (STRONG scaling: total work fixed);
    every rank generates its pairs on host threads into pinned ring buffers and streams them
    through swb200_submit/_wait.  End-to-end alignments/s, max over ranks."""
    import torch
    import torch.distributed as dist
    import swb200
    from sharding import max_over_ranks, shard_range, sum_over_ranks
    from streaming import StreamRunner

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ddist = dist if world > 1 else None
    matrix, gap = swb200.MATRIX_SPEEDTEST, swb200.GAP_SPEEDTEST
    numa = swb200.bind_to_gpu_numa_node(local_rank) if world > 1 else None   # before any pinned allocation
    ctx = swb200.Context(devices=[local_rank])
    lo, hi = shard_range(args.pairs, rank, world)
    threads = max(1, (os.cpu_count() or 8) // world - 1)
    runner = StreamRunner(ctx, batch_pairs=args.batch_pairs, n_buffers=3, packed=args.packed, gen_threads=threads)
    runner.run(lo, min(hi - lo, 2 * args.batch_pairs), matrix, gap)      # warm-up: first-touch of pinned buffers, staging allocation
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = ctx.launch_count
    rep = runner.run(lo, hi - lo, matrix, gap)
    torch.cuda.synchronize()
    # every collective is executed by EVERY rank, before any rank-0-only code
    wall = max_over_ranks(rep.wall_s, ddist)
    total = sum_over_ranks(rep.pairs, ddist)
    score_sum = sum_over_ranks(rep.score_sum, ddist)
    gen_s = max_over_ranks(rep.produce_s, ddist)
    wait_s = max_over_ranks(rep.wait_s, ddist)
    launches = sum_over_ranks(ctx.launch_count - launches0, ddist)
    h2d = sum_over_ranks(rep.bytes_h2d, ddist)
    d2h = sum_over_ranks(rep.bytes_d2h, ddist)
    if rank == 0:
        line = {
            "metric": "alignments_per_s_streaming_e2e", "value": total / wall, "unit": "alignments/s", "gcups": total * CELLS_PER_PAIR / wall / 1e9,
            "n_gpus": world, "steps": 1, "warmup": 1, "ms_per_step": wall * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int16x2", "data": "synthetic",
            "config": {"workload": f"streaming: {args.pairs} counter-stream pairs sharded by contiguous index range over {world} rank(s); host threads -> pinned ring buffers -> swb200_submit{'_packed' if args.packed else ''}",
                       "pairs": args.pairs, "batch_pairs": args.batch_pairs, "wire_format": "2-bit packed (source.cpp:1580-1583), 64 B/pair" if args.packed else "byte codes, 256 B/pair",
                       "gen_threads_per_rank": threads, "host_cores": os.cpu_count(), "numa": numa},
            "e2e": {"value": total / wall, "unit": "alignments/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "breakdown": {"wall_s": wall, "host_generation_s_max_rank": gen_s, "blocked_on_gpu_pipeline_s_max_rank": wait_s,
                          "bottleneck": "host generation" if gen_s > 0.8 * wall else "PCIe/kernel pipeline"},
            "gpu_launches": int(launches), "score_sum": int(score_sum), "mean_score": score_sum / total,
            "verified": {"score_sum_equals_reference": stream_sum_check(args.pairs, int(score_sum)),
                         "golden": "tests/golden/counter_stream_sums.json (sums of the unmodified reference's simd9 scores over prefixes of the counter stream; null = no entry for this --pairs)"},
        }
        print(json.dumps(line), flush=True)
    runner.close()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# --------------------------------------------------------------------------- semi-global X-drop aligner
SG_LEN = 16384
SG_ROUNDS_NOMINAL = 2 * SG_LEN       # a pair aligned end to end runs one round per anti-diagonal
SG_TRACE_BYTES_PER_ROUND = 8.125     # 64 direction bits + 1 move bit, written once and read once by the traceback
SG_RECORD_BYTES_PER_ROUND = 16       # what the kernels move: one 16-byte record per round, written by the forward kernel, read by the traceback


SG_KERNEL_SHAPES = {16: {"pairs_per_warp": 32, "instr": 360.0, "alu": 222.0}, 8: {"pairs_per_warp": 16, "instr": 224.0, "alu": 127.0}}


def sg_kernel_shape(n, sm_count):
    """The forward kernel the library launches for a batch of n pairs (csrc/sg_kernel.cuh, sg_words_for): words per lane,
    pairs per warp, and warp instructions per warp-round (all pairs of the warp advance one round) in total and on the ALU
    pipe -- from the committed ncu capture when it is of this width, else the SASS count of the loop body."""
    words = 16 if n >= sm_count * 128 else 8
    shape = dict(SG_KERNEL_SHAPES[words], words_per_lane=words, src="SASS count of the loop body (cuobjdump)")
    path = os.path.join(ROOT, "profiles", "r01", "ncu_full_semiglobal_v11_summary.json")
    try:
        with open(path) as f:
            d = json.load(f)
        if int(d.get("words_per_lane", 0)) == words:
            shape.update(instr=float(d["warp_instr_per_warp_round"]), alu=float(d["alu_pipe_instr_per_warp_round"]),
                         src="profiles/r01/ncu_full_semiglobal_v11_summary.json")
    except (OSError, KeyError, ValueError):
        pass
    return shape


def sg_weighted_ops_sum(ops: np.ndarray, n_ops: np.ndarray, rows_per_block: int = 512) -> int:
    """sum over pairs p and k < n_ops[p] of (k+1) * ops[p][k] -- moves when any op of any alignment moves."""
    total = 0
    w = np.arange(1, ops.shape[1] + 1, dtype=np.uint64)
    for r0 in range(0, ops.shape[0], rows_per_block):
        blk = ops[r0:r0 + rows_per_block].astype(np.uint64)
        blk *= (np.arange(ops.shape[1])[None, :] < n_ops[r0:r0 + rows_per_block, None])
        total += int((blk * w[None, :]).sum(dtype=np.uint64))
    return total


def sg_cpu_reference(a, b, budget_s=20.0):
    """The reference's aligner on all host threads (the fastest of its four AVX2 forms on this box)."""
    from oracle import oracle as O   # allowed: cpu_baseline / --impl reference legs only
    cores = os.cpu_count() or 1
    if not O.have_ref():
        return None
    names = {1: "_simd (source.cpp:1978-2165)", 2: "_simd_mark2 (2167-2353)", 3: "_simd_mark3 (2355-2541)", 4: "_simd_mark4 (2543-2725)"}
    probe = min(a.shape[0], 32 * cores)
    best = None
    for v in (1, 2, 3, 4):
        O.ref_semiglobal_batch(v, a[:cores], b[:cores], threads=cores)
        t = time.perf_counter(); O.ref_semiglobal_batch(v, a[:probe], b[:probe], threads=cores); dt = time.perf_counter() - t
        if best is None or dt < best[0]:
            best = (dt, v)
    dt, v = best
    m = int(min(a.shape[0], max(probe, probe * budget_s / 3 / dt)))
    times = []
    for _ in range(3):
        t = time.perf_counter(); sc = O.ref_semiglobal_batch(v, a[:m], b[:m], threads=cores); times.append(time.perf_counter() - t)
    return {"variant": "SemiGlobal_AdaptiveBanded_XDrop_111_32_70" + names[v] + ", unmodified, g++ -O3 -mavx2 (fastest of the four AVX2 forms here)",
            "cores": cores, "sample_pairs": m, "s_per_pass": min(times), "alignments_per_s": m / min(times), "scores": sc}


def run_semiglobal_arm(args):
    """`--workload semiglobal`: n pairs of 16384-mers related as in the reference's TestSemiGlobal (10 % mismatch /
    insert / delete, source.cpp:2750-2771) through the adaptive-banded X-drop aligner, score + traceback.
    value = alignments/s device-resident; e2e = swb200_semiglobal_xdrop_batch with pinned host arrays."""
    import swb200
    # default batch: 32 pairs x 2 warps per scheduler x 4 schedulers x 148 SMs = 37888 pairs -- the forward kernel runs one warp per
    # 32 pairs, and a batch that is not a multiple of 148 x 128 pairs leaves some schedulers a warp short
    n = args.pairs if args.pairs != 100_000_000 else 148 * 256
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        a, b = swb200.related_pairs(0, min(n, 4096), SG_LEN)
        r = sg_cpu_reference(a, b, budget_s=30.0)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libswref.so not built"}))
            return
        v = r["alignments_per_s"]
        print(json.dumps({"impl": "reference", "metric": "alignments_per_s", "value": v, "unit": "alignments/s", "n_gpus": args.gpus, "steps": 3, "warmup": 1,
                          "ms_per_step": r["s_per_pass"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                          "config": {"workload": "semi-global X-drop aligner, 16384-mer pairs, TestSemiGlobal-style 10/10/10 % edits", "pairs_per_step": r["sample_pairs"], "cpu_model": cpu_model()},
                          "cpu_baseline": {"value": v, "unit": "alignments/s", "cores": r["cores"], "kind": "reference", "variant": r["variant"], "sample": f"{r['sample_pairs']} pairs per pass, best of 3"},
                          "e2e": {"value": v, "unit": "alignments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
        return
    import torch
    import torch.distributed as dist
    from sharding import max_over_ranks, sum_over_ranks
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ddist = dist if world > 1 else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    ctx = swb200.Context(devices=[local_rank])
    info = ctx.semiglobal_kernel_info()
    pa, pb = swb200.PinnedArray((n, SG_LEN), np.uint8), swb200.PinnedArray((n, SG_LEN), np.uint8)
    swb200.related_pairs(rank * n, n, SG_LEN, out=(pa.array, pb.array))       # rank r aligns pairs [r n, (r+1) n): weak scaling, no collective
    h_meta = [swb200.PinnedArray((n,), np.int32) for _ in range(4)]
    h_ops = swb200.PinnedArray((n, 2 * SG_LEN), np.uint8)
    d_a, d_b = torch.from_numpy(pa.array).cuda(), torch.from_numpy(pb.array).cuda()
    d_meta = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(4)]
    d_ops = torch.empty((n, 2 * SG_LEN), dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()

    def launch():
        ctx.semiglobal_xdrop_device(d_a, d_b, d_meta[0], d_meta[1], d_meta[2], d_meta[3], d_ops)

    def launch_forward_only():              # score and end cell only: the forward kernel alone
        ctx.semiglobal_xdrop_device(d_a, d_b, d_meta[0], d_meta[1], d_meta[2])
    for _ in range(max(3, args.warmup)):
        launch()
    torch.cuda.synchronize()
    fwd_ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    launch_forward_only()
    fwd_ev[0].record(stream)
    for i in range(args.steps):
        launch_forward_only()
        fwd_ev[i + 1].record(stream)
    torch.cuda.synchronize()
    fwd_ms = fwd_ev[0].elapsed_time(fwd_ev[-1]) / args.steps
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    launches0 = ctx.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    t_wall0 = time.perf_counter()
    ev[0].record(stream)
    for i in range(args.steps):
        launch()                 # 268 MB of sequence + ~2.1 GB of trace written and read per launch: far beyond L2
        ev[i + 1].record(stream)
    barrier()
    t_wall1 = time.perf_counter()
    launches_dev = ctx.launch_count - launches0
    ms = max_over_ranks(ev[0].elapsed_time(ev[-1]) / args.steps, ddist)
    n_total = sum_over_ranks(n, ddist)
    dev_scores = d_meta[0].cpu().numpy()
    dev_nops = d_meta[3].cpu().numpy()
    rounds = (d_meta[1].cpu().numpy().astype(np.int64) + d_meta[2].cpu().numpy()).sum()   # >= end_y + end_x rounds per pair (a lower bound: the band runs on to the edge)

    # e2e: host arrays in, scores + tracebacks out
    def e2e():
        ctx._check(ctx._lib.swb200_semiglobal_xdrop_batch(ctx._h, pa.array.ctypes.data, pb.array.ctypes.data, SG_LEN, n,
                                                          h_meta[0].array.ctypes.data, h_meta[1].array.ctypes.data, h_meta[2].array.ctypes.data,
                                                          h_meta[3].array.ctypes.data, h_ops.array.ctypes.data))
    # The clock sampler (nvidia-smi -lms 100) covers the device-resident timed region above and stops here: every NVML query
    # holds up the CUDA calls of the moment for 10-40 ms, and this call's calling thread enqueues chunks and polls events all
    # the way through -- with the sampler running, calls took 41 or 113 ms at random (the default line's semi-global leg,
    # which runs after its sampler has stopped: 39-48 ms).
    sampler.stop()
    for _ in range(2):
        e2e()
    barrier()
    launches1 = ctx.launch_count
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e()
    e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0) / args.steps, ddist)
    launches_e2e = ctx.launch_count - launches1
    if rank != 0:
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return
    clocks = sampler.summary(t_wall0, t_wall1)
    e2e_ok = bool(np.array_equal(h_meta[0].array, dev_scores) and np.array_equal(h_meta[3].array, dev_nops))

    # The cpu_baseline leg (the only place this file runs anything under oracle/): the reference's aligner timed on the
    # host cores, its scores compared with the GPU's, and -- the oracle as checker -- score and traceback of a sample of
    # pairs.  With --no-cpu-baseline nothing under oracle/ is loaded and the committed whole-batch sums below stand alone.
    ok = None
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import oracle as O
        O.build()
        ok = True
        for i in range(0, n, max(1, n // 24)):
            s, ey, ex, ops = O.semiglobal_xdrop(pa.array[i], pb.array[i])
            ok = ok and s == h_meta[0].array[i] and ops.size == h_meta[3].array[i] and np.array_equal(h_ops.array[i, :ops.size], ops)
        cpu = sg_cpu_reference(pa.array, pb.array, budget_s=20.0)
        if cpu is not None:
            ok = ok and np.array_equal(cpu["scores"], h_meta[0].array[:cpu["sample_pairs"]])
    # the WHOLE batch against sums computed with the oracle restatement (tests/golden/make_semiglobal_batch_sums.py):
    # scores, end cells, op counts and a position-weighted sum of every op string.  None = no entry for this batch size.
    full_ok = None
    try:
        with open(os.path.join(ROOT, "tests", "golden", "semiglobal_batch_sums.json")) as f:
            want = json.load(f)["prefix"].get(str(n))
        if want is not None:
            got = {k: int(h_meta[j].array.sum(dtype=np.int64)) for j, k in enumerate(("score", "end_y", "end_x", "n_ops"))}
            got["ops_weighted"] = sg_weighted_ops_sum(h_ops.array, h_meta[3].array)
            full_ok = bool(got == {k: int(v) for k, v in want.items()})
    except Exception as ex:       # a check, not the measurement: report and carry on
        full_ok = f"{type(ex).__name__}: {ex}"
    peaks = load_peaks()
    sm_mhz = clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)
    # roofline of the forward kernel (the dominant one): its ALU pipe.  A warp-round advances sixteen pairs by one round;
    # it needs `alu_wr` ALU-pipe warp instructions (ncu), and an SM retires 2 of those per clock (63.5 lanes, INT_PEAK.json).
    shape = sg_kernel_shape(n, info["sm_count"])
    instr_wr, alu_wr, instr_src, ppw = shape["instr"], shape["alu"], shape["src"], float(shape["pairs_per_warp"])
    int_peak = json.load(open(os.path.join(ROOT, "profiles", "INT_PEAK.json")))
    lanes = float(int_peak.get("alu_lanes_per_clk_per_sm", 63.5))
    alu_peak = lanes * info["sm_count"] * sm_mhz * 1e6            # thread-level ALU-pipe instructions per second
    warp_rounds_per_s = (n / ppw) * SG_ROUNDS_NOMINAL / (fwd_ms * 1e-3)
    alu_achieved = warp_rounds_per_s * alu_wr * 32.0
    rounds_per_s = n * SG_ROUNDS_NOMINAL / (ms * 1e-3)
    trace_gbs = n * SG_ROUNDS_NOMINAL * SG_RECORD_BYTES_PER_ROUND * 2 / (ms * 1e-3) / 1e9
    line = {
        "metric": "alignments_per_s", "value": n_total / (ms * 1e-3), "unit": "alignments/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "band_gcups": rounds_per_s * 32 / 1e9,
        "config": {"workload": "SURVEY.md 8(f4): adaptive-banded X-drop semi-global aligner (band 32, X 70, 1/1/1), score + traceback, "
                               f"{n} pairs of 16384-mers with 10/10/10 % mismatch/insert/delete (TestSemiGlobal's construction, source.cpp:2750-2771)",
                   "pairs": n, "pairs_per_gpu": n, "sharding": "contiguous index ranges, no collective", "seq_len": SG_LEN, "l2": f"inputs {2 * n * SG_LEN / 1e6:.0f} MB + {n * SG_ROUNDS_NOMINAL * 16 / 1e9:.1f} GB of round records per launch > 126 MB L2, no flush needed",
                   "kernel": info},
        "clocks": clocks,
        "e2e": {"value": n_total / (e2e_ms * 1e-3), "unit": "alignments/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": 2 * n * SG_LEN,
                "d2h_bytes_per_step": n * (16 + 2 * SG_LEN), "api": "swb200_semiglobal_xdrop_batch (C ABI, pinned host arrays; scores, end cells and move strings back)",
                "equals_device_leg": e2e_ok},
        "gpu_launches": int(launches_dev), "gpu_launches_e2e": int(launches_e2e),
        "roofline": {"bound": "int_alu", "kernel": "swb::sg2_xdrop_kernel (forward pass; timed alone as the score-only call)",
                     "achieved": alu_achieved / 1e12, "peak": alu_peak / 1e12, "unit": "Tinstr/s (thread-level ALU-pipe instructions)",
                     "frac": alu_achieved / alu_peak, "avg_launch_ms": fwd_ms, "share_of_step": fwd_ms / ms,
                     "alu_instr_per_warp_round": alu_wr, "instr_per_warp_round": instr_wr, "instr_src": instr_src,
                     "warp_rounds_per_launch": (n / ppw) * SG_ROUNDS_NOMINAL, "words_per_lane": shape["words_per_lane"], "pairs_per_warp": ppw,
                     "note": "one lane per pair (two below 18944 pairs): one warp instruction serves 32 (16) pairs; a round is one serial chain per pair, so the kernel "
                             "is bound by how many ALU-pipe instructions a round needs and by how many warps there are to overlap the chains, not by HBM",
                     "traffic": None,
                     "hbm": {"achieved": trace_gbs + n * 2 * SG_LEN / (ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                             "algorithmic_bytes_per_round": SG_TRACE_BYTES_PER_ROUND * 2, "record_bytes_per_round": SG_RECORD_BYTES_PER_ROUND * 2}},
        "verified": {"sample_equals_oracle_score_and_traceback": (None if ok is None else bool(ok)), "e2e_equals_device": e2e_ok, "min_end_rounds": int(rounds // n),
                     "whole_batch_sums_equal_oracle": full_ok},
    }
    if cpu is not None:
        line["cpu_baseline"] = {"value": cpu["alignments_per_s"], "unit": "alignments/s", "cores": cpu["cores"], "kind": "reference", "variant": cpu["variant"],
                                "cpu_model": cpu_model(), "sample": f"{cpu['sample_pairs']} of {n} pairs per pass, best of 3 passes"}
    print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------- length sweep
def sweep_pairs(L: int, info: dict) -> int:
    """Pairs per launch of the length sweep: 2^34 cells, never fewer than 262144 pairs, rounded UP to whole waves of the
    kernel on this GPU (resident pairs = SMs x blocks per SM x threads x 2).  A pair is one indivisible work item of
    L^2 cells, so a launch that ends in a partly filled wave times the tail, not the kernel: at L = 512 the plain
    262144 pairs are 2.31 waves of 113664 and cannot read above 0.89 of the kernel's steady rate."""
    base = max((1 << 34) // (L * L), 262144)
    wave = info["sm_count"] * info["blocks_per_sm"] * info["threads_per_block"] * 2
    return -(-base // wave) * wave if wave > 0 else base


def run_sweep_arm(args):
    """BASELINE.json configs[3]: `--workload sweep` -- square pairs of 128, 256 and 512 bases on one
    GPU, device-resident, 2^34 cells per launch and never fewer than 262144 pairs, rounded up to whole waves of the
    kernel (sweep_pairs)."""
    import torch
    import swb200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- no CPU fallback")
    torch.cuda.set_device(0)
    matrix, gap = swb200.MATRIX_SPEEDTEST, swb200.GAP_SPEEDTEST
    ctx = swb200.Context(devices=[0])
    peaks = load_peaks()
    rows = []
    for L in swb200.SWEEP_LENGTHS:
        info = ctx.kernel_info(matrix, gap, seq_len=L)
        n = sweep_pairs(L, info)
        # the counter stream re-cut to length L (sequence i = rows i*L/128 .. of counter_pairs(0, n*L/128)): reproducible on
        # the host, so the whole batch has a committed score sum from the oracle (tests/golden/make_sweep_sums.py)
        h_a, h_b = swb200.counter_pairs(0, n * (L // 128))
        d_a = torch.from_numpy(h_a.reshape(n, L)).cuda()
        d_b = torch.from_numpy(h_b.reshape(n, L)).cuda()
        del h_a, h_b
        d_s = torch.empty(n, dtype=torch.int32, device="cuda")
        for _ in range(max(3, args.warmup)):
            ctx.score_batch_device(d_a, d_b, matrix, gap, d_s)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            ctx.score_batch_device(d_a, d_b, matrix, gap, d_s)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        gcups = n * L * L / (ms * 1e-3) / 1e9
        peak_t = peaks["alu_lanes_per_clk_per_sm"] * info["sm_count"] * peaks.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
        score_sum = int(d_s.sum(dtype=torch.int64).item())
        sum_ok = None
        try:
            with open(os.path.join(ROOT, "tests", "golden", "sweep_sums.json")) as f:
                want = json.load(f)["by_length"][str(L)]["sum_of_scores_by_pairs"].get(str(n))
            if want is not None:
                sum_ok = bool(int(want) == score_sum)
        except (OSError, KeyError, ValueError):
            pass
        rows.append({"seq_len": L, "pairs": n, "ms_per_launch": ms, "gcups": gcups, "alignments_per_s": n / (ms * 1e-3),
                     "roofline_frac": gcups * 1e9 * ALGO_INSTR_PER_CELL / 1e12 / peak_t, "mean_score": score_sum / n,
                     "score_sum": score_sum, "score_sum_equals_oracle": sum_ok, "kernel": info})
        del d_a, d_b, d_s
    line = {"metric": "GCUPS", "unit": "GCUPS", "value": rows[0]["gcups"], "n_gpus": 1, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": rows[0]["ms_per_launch"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16x2", "data": "synthetic",
            "config": {"workload": "configs[3]: sequence-length sweep 128/256/512 (templated kernels), max(2^34 cells, 262144 pairs) per launch rounded up to whole waves of resident pairs, iid pairs (counter stream re-cut to length L), matrix +10/-30, gap 15",
                       "roofline_peak": f"{peaks['alu_src']}, at sm_max clock"},
            "sweep": rows}
    print(json.dumps(line), flush=True)
    ctx.close()


# --------------------------------------------------------------------------- 8(f2), 8(f3): one-vs-many and fixed 1/1/1
def run_variants_arm(args):
    """`--workload variants`: the two call shapes of SURVEY.md 8(f2)/(f3) -- 1 M queries against ONE target
    (SmithWaterman_8b111x32mark*, source.cpp:1227-1234) and 1 M independent pairs at the fixed 1/-1/1 scoring
    (SmithWaterman_111 / _8bit111simd, source.cpp:1073-1225) -- through the C ABI with pinned host arrays, plus the
    1/-1/1 kernel device-resident.  CPU beside them: the reference's x32mark3 (one thread: it has no batch driver) and its
    general simd9 on all cores at the same scoring."""
    import torch
    import swb200
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- no CPU fallback")
    torch.cuda.set_device(0)
    ctx = swb200.Context(devices=[0])
    n = PAIRS_PER_GPU
    pa, pb = swb200.PinnedArray((n, 128), np.uint8), swb200.PinnedArray((n, 128), np.uint8)
    ps = swb200.PinnedArray((n,), np.int32)
    swb200.reference_stream(n, out=(pa.array, pb.array))
    target = pb.array[0].copy()

    def timed(fn, reps):
        for _ in range(3):
            fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return 1e3 * (time.perf_counter() - t0) / reps
    rows = []
    ms = timed(lambda: ctx.score_batch_111(pa.array, pb.array, out=ps.array), args.steps)
    s111 = ps.array.copy()
    rows.append({"call": "swb200_score_batch_111 (host arrays, 1 M pairs)", "ms_per_step": ms, "gcups": n * CELLS_PER_PAIR / ms / 1e6,
                 "alignments_per_s": n / (ms * 1e-3), "h2d_bytes_per_step": 2 * n * 128, "d2h_bytes_per_step": 4 * n})
    ms = timed(lambda: ctx.score_one_vs_many(pa.array, target, out=ps.array), args.steps)
    sx = ps.array.copy()
    rows.append({"call": "swb200_score_one_vs_many (host arrays, 1 M queries x 1 target, 1/-1/1)", "ms_per_step": ms, "gcups": n * CELLS_PER_PAIR / ms / 1e6,
                 "alignments_per_s": n / (ms * 1e-3), "h2d_bytes_per_step": n * 128 + 128, "d2h_bytes_per_step": 4 * n})
    d_a, d_b = torch.from_numpy(pa.array).cuda(), torch.from_numpy(pb.array).cuda()
    d_s = torch.empty(n, dtype=torch.int32, device="cuda")
    for _ in range(3):
        ctx.score_batch_device(d_a, d_b, swb200.MATRIX_111, swb200.GAP_111, d_s)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        ctx.score_batch_device(d_a, d_b, swb200.MATRIX_111, swb200.GAP_111, d_s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    rows.append({"call": "swb200_score_batch_device at 1/-1/1 (device-resident, 1 M pairs)", "ms_per_step": ms, "gcups": n * CELLS_PER_PAIR / ms / 1e6,
                 "alignments_per_s": n / (ms * 1e-3), "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    ok = bool(np.array_equal(d_s.cpu().numpy(), s111))
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import oracle as O   # allowed: cpu_baseline leg only
        if O.have_ref():
            cores = os.cpu_count() or 1
            m = 4096                      # x32mark3: 128 calls of 32 queries against the same target, one thread
            t0 = time.perf_counter()
            got = np.concatenate([O.ref_x32(3, pa.array[k:k + 32], target) for k in range(0, m, 32)])
            dt = time.perf_counter() - t0
            ok = ok and bool(np.array_equal(got, sx[:m]))
            m9 = 200_000
            O.ref_score_batch(9, pa.array[:m9], pb.array[:m9], swb200.MATRIX_111, 1, threads=cores)
            t0 = time.perf_counter()
            g9 = O.ref_score_batch(9, pa.array[:m9], pb.array[:m9], swb200.MATRIX_111, 1, threads=cores)
            dt9 = time.perf_counter() - t0
            ok = ok and bool(np.array_equal(g9, s111[:m9]))
            cpu = {"x32mark3_1_thread": {"gcups": m * CELLS_PER_PAIR / dt / 1e9, "sample": f"{m} queries, one thread, ctypes call per 32"},
                   "simd9_all_cores": {"gcups": m9 * CELLS_PER_PAIR / dt9 / 1e9, "cores": cores, "sample": f"{m9} pairs"}, "cpu_model": cpu_model()}
    line = {"metric": "GCUPS", "unit": "GCUPS", "value": rows[2]["gcups"], "n_gpus": 1, "steps": args.steps, "warmup": 3, "ms_per_step": rows[2]["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16x2", "data": "synthetic",
            "config": {"workload": "SURVEY.md 8(f2)/(f3): one-vs-many and fixed 1/-1/1 scoring, 1 M 128-mers of the reference stream"},
            "e2e": {"value": rows[0]["gcups"], "unit": "GCUPS", "h2d_bytes_per_step": rows[0]["h2d_bytes_per_step"], "d2h_bytes_per_step": rows[0]["d2h_bytes_per_step"]},
            "variants": rows, "verified": {"equal_to_device_and_reference": ok}, "gpu_launches": int(ctx.launch_count)}
    if cpu is not None:
        line["cpu_baseline"] = dict(cpu, kind="reference", unit="GCUPS", value=cpu["simd9_all_cores"]["gcups"], cores=cpu["simd9_all_cores"]["cores"],
                                    sample=cpu["simd9_all_cores"]["sample"])
    print(json.dumps(line), flush=True)
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pack-threads", type=int, default=None, help="host 2-bit packing lanes per GPU in the e2e leg (default: library auto; 0 = off)")
    ap.add_argument("--no-plain-e2e", action="store_true", help="skip the lanes-off comparison run of the e2e leg")
    ap.add_argument("--cpu-table", action="store_true", help="with --impl reference: scalar/simd4/simd7/simd9, 1 thread and all cores")
    ap.add_argument("--workload", choices=["batch1m", "stream", "sweep", "semiglobal", "variants"], default="batch1m",
                    help="batch1m = the headline 1M-pair batch (default); stream = configs[2]/[4] streaming of --pairs pairs")
    ap.add_argument("--pairs", type=int, default=100_000_000)
    ap.add_argument("--batch-pairs", type=int, default=1 << 21)
    ap.add_argument("--packed", action="store_true", help="stream the 2-bit packed wire format (64 B/pair)")
    ap.add_argument("--no-semiglobal", action="store_true", help="batch1m at N = 1: skip the semi-global aligner leg (SURVEY.md 8(f4))")
    ap.add_argument("--quick", action="store_true", help="batch1m: only the device-resident and the two end-to-end legs (no host ceiling, stream, in-process, sweep, per-pair)")
    ap.add_argument("--stream-pairs", type=int, default=100_000_000, help="batch1m: pairs of the sharded streaming leg (configs[2]/[4])")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        log("bench.py: warmup raised to 3 (timing rules)")
        args.warmup = 3
    if args.workload == "semiglobal":
        run_semiglobal_arm(args)
    elif args.impl == "reference":
        run_reference_arm(args)
    elif args.workload == "stream":
        run_stream_arm(args)
    elif args.workload == "sweep":
        run_sweep_arm(args)
    elif args.workload == "variants":
        run_variants_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
