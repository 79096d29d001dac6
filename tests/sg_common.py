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
