"""Generates tests/golden/reference_stream_10m.json: the reference's own differential test at its own size --
TestSimdSmithWaterman runs 10 000 000 iterations of the mt19937_64(10000) stream (source.cpp:2944-2953) -- as
known answers: sum, min, max, first arg-max and FNV-1a-64 of the 10 M scores, for the harness matrix +10/-30 gap 15
(source.cpp:2954-2959) and for 1/-1/1.

    python tests/golden/make_reference_stream_10m.py      (authoring container: needs oracle/_ref; about a minute)

Scores come from the unmodified reference: simd9 and simd4 over all 10 M pairs (asserted equal), the scalar
SmithWaterman over the first 2 M.  The first million reproduces SURVEY.md 8(c) (75 478 815, ae56a1e6a1d57492)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "smith-waterman-simd_b200"))
from oracle import oracle as O  # noqa: E402
import swb200  # noqa: E402  (generator and checksum only)

N = 10_000_000


def main():
    assert O.have_ref()
    th = os.cpu_count() or 1
    a, b = swb200.reference_stream(N)
    ar, br = O.reference_stream(4096, use_ref=True)          # the reference's own std::uniform_int_distribution draws
    assert np.array_equal(a[:4096], ar) and np.array_equal(b[:4096], br)
    out = {}
    for key, (m, g) in {"speedtest_10_-30_15": (O.MATRIX_SPEEDTEST, 15), "x32_1_-1_1": (O.MATRIX_111, 1)}.items():
        s = O.ref_score_batch(9, a, b, m, g, threads=th)
        assert np.array_equal(s, O.ref_score_batch(4, a, b, m, g, threads=th))
        assert np.array_equal(s[:2_000_000], O.ref_score_batch(0, a[:2_000_000], b[:2_000_000], m, g, threads=th))
        out[key] = {"pairs": N, "sum": int(s.sum(dtype=np.int64)), "min": int(s.min()), "max": int(s.max()),
                    "first_argmax": int(s.argmax()), "fnv1a64": f"{swb200.fnv1a64(s):016x}",
                    "first_million": {"sum": int(s[:1_000_000].sum()), "fnv1a64": f"{swb200.fnv1a64(s[:1_000_000]):016x}"}}
        print(key, out[key], flush=True)
    assert out["speedtest_10_-30_15"]["first_million"] == {"sum": 75478815, "fnv1a64": "ae56a1e6a1d57492"}
    with open(os.path.join(HERE, "reference_stream_10m.json"), "w") as f:
        json.dump({"stream": "std::mt19937_64(10000), base = draw >> 62, a[i] and b[i] drawn alternately (source.cpp:2944-2953)",
                   "scored_by": "unmodified reference: simd9 == simd4 on all pairs, == scalar on the first 2 000 000", "by_scoring": out}, f, indent=1)


if __name__ == "__main__":
    main()
