"""Generates tests/golden/sweep_sums.json: the sum of the oracle's scores over the batches of
`bench.py --workload sweep` (BASELINE.json configs[3]): L = 128 / 256 / 512, max(2^34 / L^2, 262144) pairs each -- as such and rounded up to whole waves of a B200 --, matrix
+10/-30, gap 15.  Inputs: the counter stream re-cut to length L -- sequence i of a batch is rows i*L/128 .. of
swb200.counter_pairs(0, n*L/128) laid end to end (`sweep_inputs` below; bench.py builds them the same way).

    python tests/golden/make_sweep_sums.py        (about a minute on 8 cores)

Scores come from oracle/sw_oracle.c, the plain-C restatement of source.cpp:35-60 for any length, pinned by
tests/test_oracle.py; at L = 128 the unmodified reference's simd9 must give the same sum where oracle/_ref exists."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "smith-waterman-simd_b200"))
from oracle import oracle as O  # noqa: E402
import swb200  # noqa: E402  (the generator only)


# Batch sizes: the plain rule max(2^34 / L^2, 262144), and the same rounded up to whole waves of a 148-SM B200 (what
# bench.py --workload sweep launches there: resident pairs = SMs x blocks per SM x threads x 2 = 113664 / 56832 / 113664
# at L = 128 / 256 / 512, csrc/swb200_api.cu LenCfg).
def sweep_pairs(L: int):
    base = max((1 << 34) // (L * L), 262144)
    wave = {128: 148 * 6 * 64 * 2, 256: 148 * 3 * 64 * 2, 512: 148 * 6 * 64 * 2}[L]
    return [base, -(-base // wave) * wave]


def sweep_inputs(L: int, first: int, n: int):
    """Sequences [first, first+n) of the length-L batch: uint8 [n][L] x 2."""
    k = L // 128
    a, b = swb200.counter_pairs(first * k, n * k)
    return a.reshape(n, L), b.reshape(n, L)


def main():
    O.build()
    threads = os.cpu_count() or 1
    out = {}
    for L in (128, 256, 512):
        sizes = sweep_pairs(L)
        sums, total, head = {}, 0, None
        for c0 in range(0, max(sizes), 16384):
            m = min(16384, max(sizes) - c0)
            a, b = sweep_inputs(L, c0, m)
            s = O.score_batch(a, b, O.MATRIX_SPEEDTEST, 15, threads=threads)
            if L == 128 and c0 == 0 and O.have_ref():
                assert np.array_equal(s, O.ref_score_batch(9, a, b, O.MATRIX_SPEEDTEST, 15, threads=threads))
            if c0 == 0:
                head = [int(x) for x in s[:8]]
            for n in sizes:                       # a batch size may end inside this chunk
                if c0 < n <= c0 + m:
                    sums[str(n)] = total + int(s[:n - c0].sum(dtype=np.int64))
            total += int(s.sum(dtype=np.int64))
        out[str(L)] = {"sum_of_scores_by_pairs": sums, "first_8_scores": head}
        print(L, out[str(L)], flush=True)
    with open(os.path.join(HERE, "sweep_sums.json"), "w") as f:
        json.dump({"input": "counter stream (seed 10000) re-cut to length L: sequence i = rows i*L/128 .. (i+1)*L/128 of counter_pairs(0, n*L/128)",
                   "scored_by": "oracle/sw_oracle.c (source.cpp:35-60 restated for any length), matrix +10/-30, gap 15", "by_length": out}, f, indent=1)


if __name__ == "__main__":
    main()
