"""Generates tests/golden/semiglobal_batch_sums.json: size-independent checks of the semi-global X-drop aligner on the
bench batch -- `swb200.related_pairs(0, N, 16384)` (TestSemiGlobal-style 10/10/10 % edits, source.cpp:2750-2771) for
N = 2048 and N = 37888 (bench.py --workload semiglobal's default: 148 SMs x 256 pairs).

    python tests/golden/make_semiglobal_batch_sums.py          (about a minute on 8 cores)

Every alignment comes from oracle/sg_oracle.c, the plain-C restatement of the reference's scalar aligner
(source.cpp:1836-1976) that tests/test_semiglobal_oracle.py pins to the reference build; the first 256 scores are
also compared here with the unmodified reference's scalar and _simd_mark4 where oracle/_ref exists.
Per prefix: the sums of score, end_y, end_x and n_ops, and the position-weighted sum of every op string
(sum over pairs and k < n_ops of (k+1) * op[k]), which moves when any op moves."""
import json
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "smith-waterman-simd_b200"))
from oracle import oracle as O  # noqa: E402
import swb200  # noqa: E402  (the generator only)

PREFIXES = [2048, 37888]


def weighted_ops_sum(ops: np.ndarray) -> int:
    return int((ops.astype(np.uint64) * np.arange(1, ops.size + 1, dtype=np.uint64)).sum())


def main():
    O.build()
    n = PREFIXES[-1]
    out = {}
    acc = dict(score=0, end_y=0, end_x=0, n_ops=0, ops_weighted=0)
    chunk = 1024
    with ThreadPoolExecutor(os.cpu_count() or 1) as pool:
        for c0 in range(0, n, chunk):
            m = min(chunk, n - c0)
            a, b = swb200.related_pairs(c0, m, 16384)
            res = list(pool.map(lambda i: O.semiglobal_xdrop(a[i], b[i]), range(m)))     # ctypes releases the GIL
            if c0 == 0 and O.have_ref():
                sc = np.array([r[0] for r in res[:256]], np.int32)
                assert np.array_equal(sc, O.ref_semiglobal_batch(0, a[:256], b[:256], threads=os.cpu_count() or 1))
                assert np.array_equal(sc, O.ref_semiglobal_batch(4, a[:256], b[:256], threads=os.cpu_count() or 1))
            for s, ey, ex, ops in res:
                acc["score"] += s; acc["end_y"] += ey; acc["end_x"] += ex; acc["n_ops"] += int(ops.size)
                acc["ops_weighted"] += weighted_ops_sum(ops)
            if c0 + m in PREFIXES:
                out[str(c0 + m)] = dict(acc)
                print(c0 + m, acc, flush=True)
    doc = {"input": "swb200.related_pairs(0, N, 16384), seed 10000, 10/10/10 % substitutions / insertions / deletions",
           "aligned_by": "oracle/sg_oracle.c (restatement of source.cpp:1836-1976; pinned by tests/test_semiglobal_oracle.py)",
           "prefix": out}
    with open(os.path.join(HERE, "semiglobal_batch_sums.json"), "w") as f:
        json.dump(doc, f, indent=1)


if __name__ == "__main__":
    main()
