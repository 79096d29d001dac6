"""Generates tests/golden/*.  Run in the authoring container, where /root/reference exists:

    python tests/golden/make_golden.py

Every expected value comes from the UNMODIFIED reference compiled by oracle/Makefile
(oracle/_ref/libswref.so): variant 0 = scalar `SmithWaterman` (source.cpp:35-60), and the
AVX2 variants 4/7/9 are asserted equal to it where their domain allows.
The fixtures travel to the GPU box; /root/reference does not.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402


def mm(match, mismatch):
    return [match if i == j else mismatch for i in range(4) for j in range(4)]


def structured_pairs(seed=20261018, per_class=96):
    """identical / 5% substitutions / indel-mutated (source.cpp:2753-2771 style) / 90% homopolymer / iid."""
    rng = np.random.default_rng(seed)
    n = per_class * 5
    a = rng.integers(0, 4, (n, 128), dtype=np.uint8)
    b = a.copy()
    c1 = slice(per_class, 2 * per_class)
    mut = rng.random((per_class, 128)) < 0.05
    blk = b[c1]
    blk[mut] = (blk[mut] + rng.integers(1, 4, int(mut.sum()))) % 4
    b[c1] = blk
    for i in range(2 * per_class, 3 * per_class):
        out = []
        for x in a[i]:
            r = rng.random()
            if r < 0.1:
                out.append(int(rng.integers(0, 4)))          # mismatch
            elif r < 0.2:
                out.extend([int(x), int(rng.integers(0, 4))])  # insertion
            elif r < 0.3:
                continue                                     # deletion
            else:
                out.append(int(x))
        out = (out + list(rng.integers(0, 4, 128)))[:128]
        b[i] = np.array(out, dtype=np.uint8)
    c3 = slice(3 * per_class, 4 * per_class)
    a[c3] = np.where(rng.random((per_class, 128)) < 0.9, 2, a[c3])
    b[c3] = np.where(rng.random((per_class, 128)) < 0.9, 2, rng.integers(0, 4, (per_class, 128)))
    b[4 * per_class:] = rng.integers(0, 4, (per_class, 128), dtype=np.uint8)
    return a, b


PARAM_SETS = [
    ("speedtest_10_-30_15", mm(10, -30), 15),          # source.cpp:3041-3046
    ("x32_1_-1_1", mm(1, -1), 1),                      # source.cpp:3202-3207
    ("corner_127_-127_127", mm(127, -127), 127),
    ("corner_127_-1_1", mm(127, -1), 1),
    ("corner_gap0", mm(5, -4), 0),
    ("corner_zero_diag", mm(0, -3), 2),
    ("corner_positive_offdiag", mm(2, 1), 3),
    ("corner_all_negative", mm(-5, -7), 3),
    ("fast_edge_63_-127_32", mm(63, -127), 32),
    ("fast_edge_25_-60_31", mm(25, -60), 31),
    ("fast_edge_1_-127_63", mm(1, -127), 63),
    ("asymmetric_a", [int(x) for x in np.random.default_rng(7).integers(-127, 128, 16)], 7),
    ("asymmetric_b", [int(x) for x in np.random.default_rng(8).integers(-20, 21, 16)], 5),
    ("asymmetric_c", [int(x) for x in np.random.default_rng(9).integers(-127, 128, 16)], 100),
]


def simd9_domain(sm, g):
    return all(0 <= e + g + 100 <= 255 for e in sm)     # source.cpp:982


def main():
    assert O.have_ref(), "oracle/_ref/libswref.so missing: run `make -C oracle` where /root/reference exists"
    meta = {"generated_by": "tests/golden/make_golden.py", "source": "oracle/_ref/libswref.so = unmodified /root/reference/source.cpp"}

    # 1. the reference's own test stream (source.cpp:2944-2953)
    n_stream = 1_000_000
    a, b = O.reference_stream(n_stream, use_ref=True)
    stream = {}
    for name, sm, g in PARAM_SETS[:2]:
        s0 = O.ref_score_batch(0, a, b, sm, g, threads=8)
        for v in (4, 7, 9):
            assert np.array_equal(O.ref_score_batch(v, a, b, sm, g, threads=8), s0), (name, v)
        entry = {"matrix": sm, "gap": g, "first16": s0[:16].tolist()}
        for m in (100_000, 1_000_000):
            part = s0[:m]
            entry[str(m)] = {"sum": int(part.sum()), "min": int(part.min()), "max": int(part.max()),
                             "argmax": int(np.argmax(part)), "fnv1a64": f"{O.fnv1a64(part):016x}"}
        stream[name] = entry
        np.save(os.path.join(HERE, f"stream_first4096_{name}.npy"), s0[:4096].astype(np.int16))
    meta["reference_stream"] = {"seed": 10000, "n": n_stream, "sets": stream,
                                "pair0_seq1_head": a[0, :8].tolist(), "pair0_seq2_head": b[0, :8].tolist()}

    # 2. structured inputs x parameter corners
    sa, sb = structured_pairs()
    exp = {}
    for name, sm, g in PARAM_SETS:
        s0 = O.ref_score_batch(0, sa, sb, sm, g)
        for v in (4, 7):
            assert np.array_equal(O.ref_score_batch(v, sa, sb, sm, g), s0), (name, v)
        if simd9_domain(sm, g):
            assert np.array_equal(O.ref_score_batch(9, sa, sb, sm, g), s0), (name, 9)
        exp[name] = s0.astype(np.int16)
    np.savez_compressed(os.path.join(HERE, "structured.npz"), seq1=sa, seq2=sb, **exp)
    meta["structured"] = {"n": int(sa.shape[0]), "param_sets": [{"name": n_, "matrix": m_, "gap": g_} for n_, m_, g_ in PARAM_SETS]}

    # 3. one-vs-many: the reference's batched functions on their own test inputs
    #    (TestSimdSmithWaterman111x32, source.cpp:3003-3030; fixed +1/-1/1 scoring)
    qs, ts = O.x32_stream(40)
    x32 = np.stack([O.ref_x32(1, qs[i], ts[i]) for i in range(40)])
    for i in range(40):
        for mark in (2, 3):
            assert np.array_equal(O.ref_x32(mark, qs[i], ts[i]), x32[i])
        assert O.ref_111(qs[i][0], ts[i]) == x32[i][0]
    np.savez_compressed(os.path.join(HERE, "x32_first40.npz"), queries=qs, targets=ts, scores=x32.astype(np.int16))
    meta["x32"] = {"iterations": 40, "source": "SmithWaterman_8b111x32mark1 == mark2 == mark3 (source.cpp:1227,1299,1383)", "sum": int(x32.sum())}

    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(json.dumps(meta["reference_stream"]["sets"], indent=1))


if __name__ == "__main__":
    main()
