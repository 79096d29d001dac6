"""CPU tests of the KERNEL'S ALGORITHM: csrc/sw_core.cuh compiled for the host with emulated
packed instructions (tests/emu) must equal the oracle.  This validates the offset frames,
the strip wrap, the FIFO re-basing and the PRMT selectors without a GPU; the GPU parity
tests (test_parity_gpu.py) then validate the real instructions."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_SRC = os.path.join(ROOT, "tests", "emu", "emu_main.cpp")
EMU_LIB = os.path.join(ROOT, "tests", "emu", "libswemu.so")
CSRC = os.path.join(ROOT, "smith-waterman-simd_b200", "csrc")


@pytest.fixture(scope="module")
def emu():
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", f"-I{CSRC}", "-o", EMU_LIB, EMU_SRC], check=True)
    lib = C.CDLL(EMU_LIB)
    lib.swemu_score_batch_len.restype = C.c_int
    lib.swemu_score_batch_len.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_int]

    def run(a, b, sm, gap, force_general=0, allow_refusal=False):
        a = np.ascontiguousarray(a, dtype=np.uint8)
        b = np.ascontiguousarray(b, dtype=np.uint8)
        m = np.asarray(sm, dtype=np.int8)
        out = np.empty(a.shape[0], dtype=np.int32)
        rc = lib.swemu_score_batch_len(a.shape[1], a.ctypes.data, b.ctypes.data, m.ctypes.data, gap, out.ctypes.data, a.shape[0], force_general)
        assert rc >= 0 or allow_refusal
        return rc, out
    return run


@pytest.mark.parametrize("variant", [0, 1, 2, 3])   # bit 0: best on the FMA pipe; bit 1: FIFO pre-offset (experimental)
def test_emu_structured_all_param_sets(emu, golden, variant):
    z = golden["structured_npz"]
    fast_seen = general_seen = 0
    C.CDLL(EMU_LIB).swemu_set_variant(variant)     # tuning variants of sw_core.cuh must all be bit-exact
    for ps in golden["structured"]["param_sets"]:
        exp = z[ps["name"]].astype(np.int32)
        for force_general in (0, 1):
            path, got = emu(z["seq1"], z["seq2"], ps["matrix"], ps["gap"], force_general)
            assert np.array_equal(got, exp), (ps["name"], "fast" if path else "general")
            fast_seen += path
            general_seen += 1 - path
    C.CDLL(EMU_LIB).swemu_set_variant(0)
    assert fast_seen >= 5 and general_seen >= 5   # both kernels' algorithms were exercised


def test_emu_reference_stream(emu, oracle):
    a, b = oracle.reference_stream(4097)   # odd count: the last thread scores one pair
    for sm, g in ((oracle.MATRIX_SPEEDTEST, 15), (oracle.MATRIX_111, 1)):
        exp = oracle.score_batch(a, b, sm, g, threads=os.cpu_count())
        _, got = emu(a, b, sm, g)
        assert np.array_equal(got, exp)


def test_emu_fast_path_domain_boundary(emu, oracle):
    # parameter sets on both sides of the fast/general switch in sw_params.h
    rng = np.random.default_rng(11)
    a = rng.integers(0, 4, (64, 128), dtype=np.uint8)
    b = a.copy()
    b[:, 40:] = rng.integers(0, 4, (64, 88), dtype=np.uint8)
    b[:8] = a[:8]

    def mm(m, x):
        return [m if i == j else x for i in range(4) for j in range(4)]
    for sm, g, want_fast in ((mm(67, -90), 30, 1), (mm(68, -90), 30, 0), (mm(97, -100), 15, 1), (mm(98, -100), 15, 0),
                             (mm(1, -127), 63, 1), (mm(0, -127), 64, 0), (mm(127, -127), 0, 1)):
        path, got = emu(a, b, sm, g)
        assert path == want_fast, (sm[0], g)
        assert np.array_equal(got, oracle.score_batch(a, b, sm, g)), (sm[0], g)


def _related_pairs(rng, n, L):
    """half iid pairs, half pairs sharing long gapped matches (so scores grow with L)"""
    a = rng.integers(0, 4, (n, L), dtype=np.uint8)
    b = rng.integers(0, 4, (n, L), dtype=np.uint8)
    for i in range(n // 2):
        keep = rng.random(L) > 0.08
        sub = a[i][keep]
        ins = rng.integers(0, 4, L, dtype=np.uint8)
        b[i] = np.concatenate([sub, ins])[:L]
        mut = rng.random(L) < 0.05
        b[i][mut] = (b[i][mut] + 1) % 4
    a[0] = b[0]   # identical pair: score = L * match
    return a, b


@pytest.mark.parametrize("variant", [1, 3])
@pytest.mark.parametrize("prefetch", [0, 1, 2])
@pytest.mark.parametrize("L", [256, 512])
def test_emu_length_sweep(emu, oracle, L, prefetch, variant):
    # BASELINE.json configs[3]: 2x and 4x the built-in shape; oracle = source.cpp:35-60 restated for any length
    rng = np.random.default_rng(L)
    a, b = _related_pairs(rng, 41, L)
    C.CDLL(EMU_LIB).swemu_set_prefetch(prefetch)   # 1 = the read-ahead FIFO path of the global-memory FIFO kernels
    C.CDLL(EMU_LIB).swemu_set_variant(variant)     # 1 = shipped; 3 = with the FIFO pre-offset (experimental)

    def mm(m, x):
        return [m if i == j else x for i in range(4) for j in range(4)]
    for sm, g in ((mm(10, -30), 15), (mm(1, -1), 1), (mm(5, -4), 0), (mm(40, -50), 30), (mm(60, -127), 3)):
        exp = oracle.score_batch(a, b, sm, g, threads=os.cpu_count())
        for force_general in (0, 1):
            path, got = emu(a, b, sm, g, force_general)
            assert np.array_equal(got, exp), (L, sm[0], g, path)
    assert exp.max() >= L * 60 * 0.5          # long alignments really were exercised
    # 512 * 127 does not fit int16: refused, not wrong
    rc, _ = emu(a, b, mm(127, -127), 127, allow_refusal=True)
    assert rc == (-2 if L == 512 else 0)
    C.CDLL(EMU_LIB).swemu_set_prefetch(0)
    C.CDLL(EMU_LIB).swemu_set_variant(0)


def test_emu_prefetching_fifo_at_128(emu, golden):
    z = golden["structured_npz"]
    lib = C.CDLL(EMU_LIB)
    try:
        for ahead in (1, 2):
            lib.swemu_set_prefetch(ahead)
            for ps in golden["structured"]["param_sets"][:6]:
                for force_general in (0, 1):
                    _, got = emu(z["seq1"], z["seq2"], ps["matrix"], ps["gap"], force_general)
                    assert np.array_equal(got, z[ps["name"]].astype(np.int32)), (ahead, ps["name"])
    finally:
        lib.swemu_set_prefetch(0)


def test_emu_one_vs_many_matches_x32_golden():
    # the kernel's algorithm with a shared target (stride 0) vs the fixture made by the reference's
    # SmithWaterman_8b111x32mark1 (tests/golden/x32_first40.npz, written by tests/test_oracle.py)
    z = np.load(os.path.join(ROOT, "tests", "golden", "x32_first40.npz"))
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", f"-I{CSRC}", "-o", EMU_LIB, EMU_SRC], check=True)
    lib = C.CDLL(EMU_LIB)
    lib.swemu_one_vs_many.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_int]
    m = np.asarray([1 if i == j else -1 for i in range(4) for j in range(4)], dtype=np.int8)
    for it in range(40):
        q = np.ascontiguousarray(z["queries"][it]); t = np.ascontiguousarray(z["targets"][it])
        for n in (32, 31):
            out = np.empty(n, np.int32)
            for fg in (0, 1):
                assert lib.swemu_one_vs_many(q.ctypes.data, t.ctypes.data, m.ctypes.data, 1, out.ctypes.data, n, fg) >= 0
                assert np.array_equal(out, z["scores"][it][:n].astype(np.int32)), (it, n, fg)


# --------------------------------------------------------------------------- semi-global aligner (csrc/sg2_core.cuh)
SG2_SRC = os.path.join(ROOT, "tests", "emu", "sg2_emu.cpp")
SG2_LIB = os.path.join(ROOT, "tests", "emu", "libsg2emu.so")


@pytest.fixture(scope="module")
def sg2():
    """The per-lane round of the semi-global kernel, run on the host: the lanes of a pair are coroutines in lock step."""
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", f"-I{CSRC}", "-o", SG2_LIB, SG2_SRC], check=True)
    lib = C.CDLL(SG2_LIB)
    lib.swemu_sg2.restype = C.c_int
    lib.swemu_sg2.argtypes = [C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 5 + [C.c_int]

    def run(a, b):
        a = np.ascontiguousarray(a, dtype=np.uint8)
        b = np.ascontiguousarray(b, dtype=np.uint8)
        n = a.size
        out = []
        for words in (16, 8, 4):        # one, two and four lanes per pair: the same templated code (the library ships 16 and 8)
            meta = np.zeros(4, np.int32)
            ops = np.zeros(2 * n, np.uint8)
            rc = lib.swemu_sg2(a.ctypes.data, b.ctypes.data, n, meta[0:].ctypes.data, meta[1:].ctypes.data, meta[2:].ctypes.data,
                               ops.ctypes.data, meta[3:].ctypes.data, words)
            assert rc == 0, (rc, words)
            out.append((int(meta[0]), int(meta[1]), int(meta[2]), ops[:meta[3]].copy()))
        for o in out[1:]:
            assert out[0][:3] == o[:3] and np.array_equal(out[0][3], o[3])
        return out[0]
    return run


def test_sg2_emu_equals_the_reference_fixtures(sg2):
    from sg_common import load_cases
    for c in load_cases():
        score, ey, ex, ops = sg2(c["seq1"], c["seq2"])
        assert (score, ey, ex, ops.size) == (c["score"], c["end_y"], c["end_x"], c["n_ops"]), c["name"]
        assert np.array_equal(ops, c["ops"]), c["name"]


def test_sg2_emu_equals_the_oracle_at_many_lengths(sg2, oracle):
    rng = np.random.default_rng(20261018)
    for length in (1, 2, 3, 7, 8, 9, 30, 31, 32, 33, 63, 64, 65, 100, 255, 256, 700, 1500):
        for kind in range(4):
            a = rng.integers(0, 4, length, dtype=np.uint8)
            if kind == 0:
                b = rng.integers(0, 4, length, dtype=np.uint8)                 # unrelated: the X-drop ends the pair early
            elif kind == 1:
                b = a.copy()                                                    # identical: the band runs down the diagonal
            else:
                rate = 0.03 if kind == 2 else 0.25
                b = a.copy()
                hit = rng.random(length) < rate
                b[hit] = rng.integers(0, 4, int(hit.sum()), dtype=np.uint8)
                if length > 8:
                    cut = int(rng.integers(1, length // 2))
                    b = np.concatenate([b[cut:], rng.integers(0, 4, cut, dtype=np.uint8)])   # a shift: gaps at the start
            exp = oracle.semiglobal_xdrop(a, b)
            got = sg2(a, b)
            assert got[:3] == exp[:3], (length, kind)
            assert np.array_equal(got[3], exp[3]), (length, kind)


def test_sg2_emu_on_the_xdrop_thresholds(sg2, oracle):
    rng = np.random.default_rng(7070)
    from sg_common import xdrop_edge_cases
    cases = xdrop_edge_cases(rng)
    assert len(cases) > 300
    ended_early = 0
    for k, (a, b) in enumerate(cases):
        exp = oracle.semiglobal_xdrop(a, b)
        got = sg2(a, b)
        assert got[:3] == exp[:3], k
        assert np.array_equal(got[3], exp[3]), k
        ended_early += exp[1] < a.size // 2
    assert 20 < ended_early < len(cases) - 20      # both outcomes occur: the run ended the alignment / the alignment got past it
