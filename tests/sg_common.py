"""Shared helpers of the semi-global aligner tests (fixture loading, hashing)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def fnv1a64_bytes(buf: bytes) -> int:
    """FNV-1a-64 over a byte string, vectorised by 4096-byte blocks would change the value, so: plain loop in numpy ints."""
    h = 1469598103934665603
    for c in np.frombuffer(buf, np.uint8).tolist():
        h = ((h ^ c) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def unpack_codes(packed: np.ndarray) -> np.ndarray:
    """2-bit packed -> codes, the layout of source.cpp:1580-1583."""
    p = np.asarray(packed, np.uint8)
    return ((p[:, None] >> (2 * np.arange(4, dtype=np.uint8))[None, :]) & 3).reshape(-1).astype(np.uint8)


def load_cases():
    with open(os.path.join(GOLDEN, "semiglobal.json")) as f:
        meta = json.load(f)
    z = np.load(os.path.join(GOLDEN, "semiglobal.npz"))
    cases = []
    for c in meta["cases"]:
        name = c["name"]
        ops = unpack_codes(z[name + "_ops"])[:c["n_ops"]]
        cases.append(dict(c, seq1=unpack_codes(z[name + "_seq1"]), seq2=unpack_codes(z[name + "_seq2"]), ops=ops))
    return cases


def ops_to_traceback(ops: np.ndarray) -> np.ndarray:
    ops = np.asarray(ops)
    tb = np.zeros((ops.size + 1, 2), np.int32)
    tb[1:, 0] = np.cumsum(ops != 2)
    tb[1:, 1] = np.cumsum(ops != 1)
    return tb


def xdrop_edge_cases(rng):
    """Pairs built to sit ON the aligner's thresholds: a run of mismatches (or a gap) of exactly the length at which the
    X-drop of 70 does or does not end the alignment, after a prefix that has or has not lifted the score above 70 (the
    threshold T = max(best - 70, 1) is pinned at 1 until then), and again near the end of the sequences."""
    cases = []
    for length in (160, 400, 1025):
        for prefix in (0, 1, 2, 35, 69, 70, 71, 72, 73, 120):
            for run in (33, 34, 35, 36, 37, 68, 69, 70, 71, 72):
                if prefix + run + 10 > length:
                    continue
                a = rng.integers(0, 4, length, dtype=np.uint8)
                b = a.copy()
                a[prefix:prefix + run] = 0          # a stretch that matches on NO diagonal: the alignment can only pay its way through
                b[prefix:prefix + run] = 1
                cases.append((a, b))
        for prefix in (5, 80, 300):
            for gap in (1, 15, 16, 17, 31, 32, 33, 40):
                if prefix + gap + 40 > length:
                    continue
                a = rng.integers(0, 4, length, dtype=np.uint8)
                b = np.concatenate([a[:prefix], a[prefix + gap:], rng.integers(0, 4, gap, dtype=np.uint8)])      # a deletion: the band must drift
                cases.append((a, b))
                cases.append((b, a))                                                                       # ... and an insertion
    return cases
