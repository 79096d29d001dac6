"""CPU tests: the oracle is pinned against the reference's known answers,
against the committed golden fixtures, and -- where oracle/_ref was built -- against the
unmodified reference itself."""
import os

import numpy as np
import pytest

N_STREAM = 100_000


@pytest.fixture(scope="module")
def stream(oracle):
    return oracle.reference_stream(N_STREAM)


def test_stream_matches_survey_head(oracle, stream, golden):
    a, b = stream
    rs = golden["reference_stream"]
    assert a[0, :8].tolist() == rs["pair0_seq1_head"] == [2, 2, 2, 0, 1, 2, 2, 3]
    assert b[0, :8].tolist() == rs["pair0_seq2_head"] == [0, 3, 2, 0, 3, 0, 0, 3]
    assert a.max() <= 3 and b.max() <= 3


def test_known_answers_speedtest_matrix(oracle, stream, golden):
    # first 16 scores, and sum/min/max/argmax/FNV of the first 100 000
    a, b = stream
    s = oracle.score_batch(a, b, oracle.MATRIX_SPEEDTEST, oracle.GAP_SPEEDTEST, threads=os.cpu_count())
    assert s[:16].tolist() == [80, 80, 70, 95, 70, 80, 80, 75, 80, 70, 75, 65, 65, 80, 100, 70]
    assert int(s.sum()) == 7_550_735 and int(s.min()) == 50 and int(s.max()) == 195
    assert int(np.argmax(s)) == 28971
    assert f"{oracle.fnv1a64(s):016x}" == "ca3723235bbaa0fc"
    g = golden["reference_stream"]["sets"]["speedtest_10_-30_15"]["100000"]
    assert g == {"sum": 7550735, "min": 50, "max": 195, "argmax": 28971, "fnv1a64": "ca3723235bbaa0fc"}


def test_known_answers_111_matrix(oracle, stream, golden):
    a, b = stream
    s = oracle.score_batch(a, b, oracle.MATRIX_111, oracle.GAP_111, threads=os.cpu_count())
    g = golden["reference_stream"]["sets"]["x32_1_-1_1"]
    assert s[:16].tolist() == g["first16"]
    assert int(s.sum()) == g["100000"]["sum"]
    assert f"{oracle.fnv1a64(s):016x}" == g["100000"]["fnv1a64"]


def test_first4096_fixture(oracle, stream):
    a, b = stream
    for name, sm, gp in (("speedtest_10_-30_15", oracle.MATRIX_SPEEDTEST, 15), ("x32_1_-1_1", oracle.MATRIX_111, 1)):
        exp = np.load(os.path.join(os.path.dirname(__file__), "golden", f"stream_first4096_{name}.npy")).astype(np.int32)
        got = oracle.score_batch(a[:4096], b[:4096], sm, gp)
        assert np.array_equal(got, exp)


def test_structured_fixture_all_param_sets(oracle, golden):
    z = golden["structured_npz"]
    a, b = z["seq1"], z["seq2"]
    for ps in golden["structured"]["param_sets"]:
        got = oracle.score_batch(a, b, ps["matrix"], ps["gap"], threads=os.cpu_count())
        assert np.array_equal(got, z[ps["name"]].astype(np.int32)), ps["name"]


def test_port_equals_reference_build(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built here (no /root/reference and no prebuilt file)")
    a, b = oracle.reference_stream(3000)
    a2, b2 = oracle.reference_stream(3000, use_ref=True)   # std::uniform_int_distribution itself
    assert np.array_equal(a, a2) and np.array_equal(b, b2)
    exp = oracle.ref_score_batch(0, a, b, oracle.MATRIX_SPEEDTEST, 15)
    assert np.array_equal(oracle.score_batch(a, b, oracle.MATRIX_SPEEDTEST, 15), exp)
    for v in range(1, 10):   # the reference's own differential test, source.cpp:2961-2979
        assert np.array_equal(oracle.ref_score_batch(v, a, b, oracle.MATRIX_SPEEDTEST, 15), exp), v


def test_pack_unpack_roundtrip(oracle):
    rng = np.random.default_rng(3)
    codes = rng.integers(0, 4, (257, 128), dtype=np.uint8)
    packed = oracle.pack2bit(codes)
    assert packed.shape == (257, 32)
    assert np.array_equal(oracle.unpack2bit(packed), codes)
    # layout of source.cpp:1580-1583: dest[i*4+j] = (src[i] >> 2j) & 3
    i, j = 5, 2
    assert codes[0, i * 4 + j] == (packed[0, i] >> (2 * j)) & 3
    if oracle.have_ref():
        import ctypes as C
        out = np.empty(128, dtype=np.uint8)
        oracle.ref().swref_unpack(packed[7].ctypes.data_as(C.POINTER(C.c_uint8)), out.ctypes.data_as(C.POINTER(C.c_uint8)))
        assert np.array_equal(out, codes[7])


def test_rectangular_and_long(oracle):
    # the templated-length restatement used by the length sweep (config 4): L=256 identical pair
    a = np.tile(np.arange(4, dtype=np.uint8), 64)[None, :]
    assert oracle.score_batch(a, a, oracle.MATRIX_SPEEDTEST, 15)[0] == 2560


def test_one_vs_many_row_is_pinned_by_the_reference_x32_functions(oracle):
    """the reference's batched functions SmithWaterman_8b111x32mark1/2/3
    (source.cpp:1227, 1299, 1383) and SmithWaterman_111 (source.cpp:1073) equal the oracle with a
    +1/-1 matrix and gap 1, on the inputs of TestSimdSmithWaterman111x32 (source.cpp:3004-3013)."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built here")
    qs, ts = oracle.x32_stream(40)
    for it in range(40):
        exp = oracle.score_batch(qs[it], np.repeat(ts[it][None, :], 32, axis=0), oracle.MATRIX_111, oracle.GAP_111)
        for mark in (1, 2, 3):
            assert np.array_equal(oracle.ref_x32(mark, qs[it], ts[it]), exp), (it, mark)
        assert oracle.ref_111(qs[it][5], ts[it]) == exp[5]
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "x32_first40.npz"))   # committed fixture = the same run
    assert np.array_equal(z["queries"], qs) and np.array_equal(z["targets"], ts)
    assert np.array_equal(z["scores"][7].astype(np.int32), oracle.ref_x32(3, qs[7], ts[7]))


def test_fixed_111_functions_of_the_reference_equal_the_general_scalar(oracle, stream):
    # SmithWaterman_111 (source.cpp:1073-1103) and SmithWaterman_8bit111simd (1105-1225)
    # are the general recurrence with +1/-1, gap 1 -- that equality is what lets one kernel serve both.
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    a, b = stream
    exp = oracle.score_batch(a[:300], b[:300], oracle.MATRIX_111, 1)
    for i in range(300):
        assert oracle.ref_111(a[i], b[i]) == exp[i]
        assert oracle.ref_8bit111(a[i], b[i]) == exp[i]


def test_counter_stream_golden_sums_against_the_port(oracle):
    # tests/golden/counter_stream_sums.json was written from the unmodified reference's simd9 (make_counter_sums.py);
    # the plain-C restatement must reproduce its 1 M-pair block sums (what ranks >= 1 of a multi-GPU bench score).
    import json
    import swb200
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "counter_stream_sums.json")) as f:
        g = json.load(f)
    want = g["sum_of_scores_block_1M"]
    a, b = swb200.counter_pairs(3_000_000, 1_000_000)
    s = oracle.score_batch(a, b, oracle.MATRIX_SPEEDTEST, 15, threads=os.cpu_count() or 1)
    assert int(s.sum(dtype=np.int64)) == want["speedtest_10_-30_15"]["3"]
    s = oracle.score_batch(a, b, oracle.MATRIX_111, 1, threads=os.cpu_count() or 1)
    assert int(s.sum(dtype=np.int64)) == want["x32_1_-1_1"]["3"]
    pre = g["sum_of_scores_over_prefix"]["speedtest_10_-30_15"]
    assert pre["1000000"] == want["speedtest_10_-30_15"]["0"]
    assert pre["2000000"] == want["speedtest_10_-30_15"]["0"] + want["speedtest_10_-30_15"]["1"]
    assert pre["8000000"] == sum(want["speedtest_10_-30_15"][str(r)] for r in range(8))


def test_reference_stream_10m_golden_is_consistent_with_the_survey_pins(oracle, stream):
    # the 10 M-pair known answers (the reference test's own size) contain the 1 M-pair pins of SURVEY.md 8(c)
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_stream_10m.json")) as f:
        g = json.load(f)["by_scoring"]
    assert g["speedtest_10_-30_15"]["first_million"] == {"sum": 75478815, "fnv1a64": "ae56a1e6a1d57492"}
    assert g["x32_1_-1_1"]["first_million"]["sum"] == 18154767
    assert g["speedtest_10_-30_15"]["pairs"] == g["x32_1_-1_1"]["pairs"] == 10_000_000
    # and the port agrees with the file on a window in the middle of the stream's first 100 000 pairs
    a, b = stream
    s = oracle.score_batch(a, b, oracle.MATRIX_SPEEDTEST, 15, threads=os.cpu_count() or 1)
    assert int(s.min()) >= g["speedtest_10_-30_15"]["min"] and int(s.max()) <= g["speedtest_10_-30_15"]["max"]
