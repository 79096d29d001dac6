"""Generates tests/golden/semiglobal.npz + semiglobal.json.  Run in the authoring container, where
/root/reference exists:

    python tests/golden/make_golden_semiglobal.py

Every expected value comes from the UNMODIFIED reference compiled by oracle/Makefile
(oracle/_ref/libswref.so): SemiGlobal_AdaptiveBanded_XDrop_111_32_70 (source.cpp:1836-1976), and its four
AVX2 forms (source.cpp:1978-2725) are asserted to return the same score and traceback here, as
TestSemiGlobal does (source.cpp:2774-2784).  Inputs are stored 2-bit packed (source.cpp:1580-1583 layout),
expected results as score / end cell / traceback length / FNV-1a-64 of the traceback's (y,x) int32 stream,
plus the complete move string of every case (2 bits per move).  The fixtures travel to the GPU box.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402

L = O.SG_LEN


def fnv1a64_bytes(buf: bytes) -> int:
    h = 1469598103934665603
    for chunk in np.frombuffer(buf, np.uint8):
        h = ((h ^ int(chunk)) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def tb_to_ops(tb: np.ndarray) -> np.ndarray:
    dy = np.diff(tb[:, 0])
    dx = np.diff(tb[:, 1])
    assert np.all((dy | dx) == 1) and np.all((dy >= 0) & (dx >= 0))
    return np.where((dy == 1) & (dx == 1), 0, np.where(dy == 1, 1, 2)).astype(np.uint8)


def cases():
    rng = np.random.default_rng(20261018)
    out = []
    a, b = O.ref_semiglobal_test_inputs(6)                      # TestSemiGlobal, source.cpp:2734-2771
    for i in range(6):
        out.append((f"testsemiglobal_it{i}", a[i], b[i]))
    sa, sb = O.ref_semiglobal_speedtest_input()                 # SpeedtestSemiGlobal, source.cpp:2804-2813
    out.append(("speedtest_pair", sa, sb))
    x = rng.integers(0, 4, L, dtype=np.uint8)
    out.append(("identical", x, x.copy()))
    out.append(("unrelated_iid", x, rng.integers(0, 4, L, dtype=np.uint8)))          # dies on the X-drop after a few dozen rounds
    out.append(("homopolymer", np.zeros(L, np.uint8), np.zeros(L, np.uint8)))        # every cell ties: exercises the tie-breaks
    y = np.concatenate([x[40:], rng.integers(0, 4, 40, dtype=np.uint8)])            # needs 40 gaps up front: beyond the band's reach? (X-drop 70 allows it)
    out.append(("shifted_by_40", x, y))
    y = np.concatenate([rng.integers(0, 4, 25, dtype=np.uint8), x[:-25]])
    out.append(("shifted_other_way_25", x, y))
    y = x.copy()
    y[8000:8100] = rng.integers(0, 4, 100, dtype=np.uint8)                          # a 100-base unrelated island: X-drop stops there
    out.append(("island_of_noise", x, y))
    z = x.copy()
    mut = rng.random(L) < 0.35
    z[mut] = rng.integers(0, 4, int(mut.sum()), dtype=np.uint8)
    out.append(("65pct_identity", x, z))
    out.append(("no_match_at_all", np.zeros(L, np.uint8), np.ones(L, np.uint8)))       # score 0 at (0,0); every cell drops after 70 rounds
    a2, b2 = x.copy(), x.copy()
    a2[6000:] = 1
    b2[6000:] = 0                                                                   # identical up to base 6000, then no match is possible: X-drop ends it
    out.append(("match_then_nothing", a2, b2))
    rep = np.tile(np.array([0, 1], np.uint8), L // 2)                               # dinucleotide repeat vs itself shifted by one
    out.append(("repeat_shift1", rep, np.roll(rep, 1)))
    return out


def main():
    meta = {"generated_by": "tests/golden/make_golden_semiglobal.py",
            "source": "oracle/_ref/libswref.so = /root/reference/source.cpp unmodified, g++ -O3 -mavx2, asserts live",
            "function": "SemiGlobal_AdaptiveBanded_XDrop_111_32_70 (source.cpp:1836-1976); _simd, _simd_mark2/3/4 asserted equal",
            "ops": "0 = diagonal, 1 = down (y+1), 2 = right (x+1); forward from (0,0); stored 4 per byte, move k at bits 2(k%4)",
            "cases": []}
    arrays = {}
    for name, a, b in cases():
        score, tb = O.ref_semiglobal(0, a, b)
        for v in (1, 2, 3, 4):
            sv, tbv = O.ref_semiglobal(v, a, b)
            assert sv == score and np.array_equal(tbv, tb), (name, v)
        assert tb[0, 0] == 0 and tb[0, 1] == 0
        ops = tb_to_ops(tb)
        ps, pey, pex, pops = O.semiglobal_xdrop(a, b)           # the restatement, checked here too
        assert ps == score and np.array_equal(pops, ops) and (pey, pex) == (int(tb[-1, 0]), int(tb[-1, 1])), name
        pad = (-ops.size) % 4
        arrays[name + "_seq1"] = O.pack2bit(a.reshape(-1, 128)).reshape(-1)
        arrays[name + "_seq2"] = O.pack2bit(b.reshape(-1, 128)).reshape(-1)
        q = np.concatenate([ops, np.zeros(pad, np.uint8)]).reshape(-1, 4)
        arrays[name + "_ops"] = (q[:, 0] | (q[:, 1] << 2) | (q[:, 2] << 4) | (q[:, 3] << 6)).astype(np.uint8)
        meta["cases"].append({"name": name, "score": score, "end_y": int(tb[-1, 0]), "end_x": int(tb[-1, 1]),
                              "traceback_len": int(tb.shape[0]), "n_ops": int(ops.size),
                              "traceback_fnv1a64": f"{fnv1a64_bytes(np.ascontiguousarray(tb, np.int32).tobytes()):016x}"})
        print(meta["cases"][-1])
    np.savez_compressed(os.path.join(HERE, "semiglobal.npz"), **arrays)
    with open(os.path.join(HERE, "semiglobal.json"), "w") as f:
        json.dump(meta, f, indent=1)


if __name__ == "__main__":
    main()
