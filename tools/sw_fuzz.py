"""Differential fuzzing of the Smith-Waterman kernel's per-thread code (host emulator, tests/emu/emu_main.cpp) against
the oracle: random score matrices and gaps over the whole reference domain, low-complexity and repetitive sequences,
at 128 / 256 / 512 bases, all four tuning variants (bit 0 best on the FMA pipe, bit 1 FIFO pre-offset) and every FIFO read-ahead distance.
Development tool:  python tools/sw_fuzz.py [seconds] [processes]"""
import ctypes as C
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_batch(rng, n=64, L=128):
    alpha = int(rng.choice([1, 2, 3, 4, 4]))
    a = rng.integers(0, alpha, (n, L)).astype(np.uint8)
    b = a.copy()
    mode = int(rng.integers(0, 4))
    if mode == 0:
        b = rng.integers(0, alpha, (n, L)).astype(np.uint8)
    elif mode == 1:
        hit = rng.random((n, L)) < rng.random()
        b[hit] = rng.integers(0, 4, int(hit.sum()))
    elif mode == 2:
        k = int(rng.integers(1, 64))
        b = np.roll(a, k, axis=1)
    else:
        period = int(rng.integers(1, 12))
        a = np.tile(rng.integers(0, 4, (n, period)), (1, L // period + 1))[:, :L].astype(np.uint8)
        b = np.roll(a, int(rng.integers(0, period + 3)), axis=1)
    kind = int(rng.integers(0, 4))
    if kind == 0:
        m = rng.integers(-127, 128, 16)
    elif kind == 1:
        m = np.where(np.eye(4, dtype=bool), rng.integers(1, 128), -rng.integers(0, 128)).reshape(16)
    elif kind == 2:
        m = rng.integers(-5, 6, 16)
    else:
        m = rng.choice([-127, -1, 0, 1, 127], 16)
    gap = int(rng.choice([0, 1, 2, 15, 63, 64, 126, 127, int(rng.integers(0, 128))]))
    return np.ascontiguousarray(a), np.ascontiguousarray(b), m.astype(np.int8), gap


def worker(args):
    seed, seconds = args
    from oracle import oracle as O
    O.build()
    lib = C.CDLL(os.path.join(ROOT, "tests", "emu", "libswemu.so"))
    lib.swemu_score_batch_len.restype = C.c_int
    lib.swemu_score_batch_len.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_int]
    rng = np.random.default_rng(seed)
    t0 = time.time()
    n = 0
    while time.time() - t0 < seconds:
        # every fourth batch at 256 or 512 bases (the length sweep), with the tuning variant and the FIFO read-ahead
        # distance of the global-memory FIFO kernel drawn at random; the rest at the reference's 128
        L = int(rng.choice([256, 512])) if n % 4 == 3 else 128
        a, b, m, gap = make_batch(rng, n=64 if L == 128 else 12, L=L)
        m = np.clip(m, -127, 32767 // L).astype(np.int8)            # int16 stays exact while L * max(S) <= 32767
        exp = O.score_batch(a, b, m, gap)
        variant, ahead = (0, 0) if L == 128 and n % 2 == 0 else (int(rng.integers(0, 4)), int(rng.integers(0, 3)))
        if os.environ.get("SWFUZZ_VARIANTS"):       # e.g. SWFUZZ_VARIANTS=2,3: a campaign on chosen variant bits only
            variant = int(rng.choice([int(x) for x in os.environ["SWFUZZ_VARIANTS"].split(",")]))
        lib.swemu_set_variant(variant)
        lib.swemu_set_prefetch(ahead)
        for fg in (0, 1):
            out = np.empty(a.shape[0], np.int32)
            rc = lib.swemu_score_batch_len(L, a.ctypes.data, b.ctypes.data, m.ctypes.data, gap, out.ctypes.data, a.shape[0], fg)
            if rc < 0 or not np.array_equal(out, exp):
                np.savez(f"/tmp/swfuzz_{seed}_{n}.npz", a=a, b=b, m=m, gap=gap)
                return ("MISMATCH", seed, n, rc, fg, L, variant, ahead, m.tolist(), gap)
        n += 1
    return ("ok", seed, n)


if __name__ == "__main__":
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60
    procs = int(sys.argv[2]) if len(sys.argv) > 2 else max(1, (os.cpu_count() or 2) - 1)
    with mp.Pool(procs) as pool:
        for r in pool.imap_unordered(worker, [(2000 + i, seconds) for i in range(procs)]):
            print(r, flush=True)
