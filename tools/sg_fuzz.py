"""Differential fuzzing of the semi-global kernel's per-lane code (host emulator, tests/emu/sg2_emu.cpp) against the
oracle on adversarial inputs: low-complexity alphabets, periodic repeats, runs of mismatches and gaps of every length
around the band width (32) and the X-drop (70).  Development tool:  python tools/sg_fuzz.py [seconds] [processes]"""
import ctypes as C
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_case(rng):
    length = int(rng.choice([rng.integers(1, 40), rng.integers(40, 200), rng.integers(200, 700)]))
    alpha = int(rng.choice([1, 2, 2, 3, 4, 4]))
    kind = int(rng.integers(0, 5))
    if kind == 0:                                   # periodic
        period = int(rng.integers(1, 9))
        a = np.resize(rng.integers(0, alpha, period), length).astype(np.uint8)
    else:
        a = rng.integers(0, alpha, length).astype(np.uint8)
    b = a.copy()
    for _ in range(int(rng.integers(0, 5))):        # edits: mismatch runs, deletions, insertions
        if length < 4:
            break
        pos = int(rng.integers(0, length))
        span = int(rng.choice([1, 2, 15, 16, 17, 31, 32, 33, 34, 35, 36, 68, 69, 70, 71, 72, int(rng.integers(1, 120))]))
        span = min(span, length - pos)
        what = int(rng.integers(0, 4))
        if what == 0:
            b[pos:pos + span] = (b[pos:pos + span] + 1 + rng.integers(0, 3, span)) % 4
        elif what == 1:
            a[pos:pos + span] = 0; b[pos:pos + span] = 1
        elif what == 2:
            b = np.concatenate([b[:pos], b[pos + span:], rng.integers(0, alpha, span).astype(np.uint8)])
        else:
            b = np.concatenate([b[:pos], rng.integers(0, alpha, span).astype(np.uint8), b[pos:]])[:length]
    return np.ascontiguousarray(a), np.ascontiguousarray(b.astype(np.uint8))


def worker(args):
    seed, seconds = args
    from oracle import oracle as O
    O.build()
    lib = C.CDLL(os.path.join(ROOT, "tests", "emu", "libsg2emu.so"))
    lib.swemu_sg2.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_int]
    rng = np.random.default_rng(seed)
    t0 = time.time()
    n = 0
    while time.time() - t0 < seconds:
        a, b = make_case(rng)
        L = a.size
        meta = np.zeros(4, np.int32)
        ops = np.zeros(2 * L, np.uint8)
        rc = lib.swemu_sg2(a.ctypes.data, b.ctypes.data, L, meta[0:].ctypes.data, meta[1:].ctypes.data, meta[2:].ctypes.data,
                           ops.ctypes.data, meta[3:].ctypes.data, (16, 8, 4)[n % 3])
        exp = O.semiglobal_xdrop(a, b)
        if rc != 0 or (int(meta[0]), int(meta[1]), int(meta[2])) != exp[:3] or not np.array_equal(ops[:meta[3]], exp[3]):
            np.save(f"/tmp/sgfuzz_{seed}_{n}_a.npy", a); np.save(f"/tmp/sgfuzz_{seed}_{n}_b.npy", b)
            return ("MISMATCH", seed, n, rc, meta.tolist(), exp[:3])
        n += 1
    return ("ok", seed, n)


if __name__ == "__main__":
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60
    procs = int(sys.argv[2]) if len(sys.argv) > 2 else max(1, (os.cpu_count() or 2) - 1)
    with mp.Pool(procs) as pool:
        for r in pool.imap_unordered(worker, [(1000 + i, seconds) for i in range(procs)]):
            print(r, flush=True)
