"""The semi-global X-drop aligner's oracle (oracle/sg_oracle.c, a restatement of source.cpp:1836-1976)
against the committed fixtures and, where oracle/_ref travelled, against the reference itself."""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from sg_common import fnv1a64_bytes, load_cases, ops_to_traceback


@pytest.fixture(scope="module")
def cases():
    return load_cases()


def test_fixture_is_self_consistent(cases):
    assert len(cases) >= 16
    for c in cases[:4] + cases[-3:]:
        assert c["seq1"].size == 16384 and c["seq2"].size == 16384
        tb = ops_to_traceback(c["ops"])
        assert tb.shape[0] == c["traceback_len"] and tuple(tb[-1]) == (c["end_y"], c["end_x"])
        assert f"{fnv1a64_bytes(tb.tobytes()):016x}" == c["traceback_fnv1a64"]


def test_restatement_equals_golden(oracle, cases):
    for c in cases:
        score, ey, ex, ops = oracle.semiglobal_xdrop(c["seq1"], c["seq2"])
        assert (score, ey, ex) == (c["score"], c["end_y"], c["end_x"]), c["name"]
        assert np.array_equal(ops, c["ops"]), c["name"]


def test_restatement_equals_reference_build(oracle):
    # TestSemiGlobal's own differential test (source.cpp:2774-2784), with the restatement beside it
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    a, b = oracle.ref_semiglobal_test_inputs(24, seed=777)

    def one(i):
        s0, tb0 = oracle.ref_semiglobal(0, a[i], b[i])
        s, ey, ex, ops = oracle.semiglobal_xdrop(a[i], b[i])
        ok = s == s0 and np.array_equal(oracle.ops_to_traceback(ops), tb0)
        for v in (1, 4):
            sv, tbv = oracle.ref_semiglobal(v, a[i], b[i])
            ok = ok and sv == s0 and np.array_equal(tbv, tb0)
        return ok
    with ThreadPoolExecutor(8) as ex:
        assert all(ex.map(one, range(24)))
    assert np.array_equal(oracle.ref_semiglobal_batch(4, a[:6], b[:6], threads=3), [oracle.ref_semiglobal(0, a[i], b[i])[0] for i in range(6)])


def test_short_lengths_have_the_obvious_answers(oracle):
    # the restatement takes the length as a parameter (the reference fixes 16384): sanity at the small end
    rng = np.random.default_rng(5)
    for n in (1, 2, 3, 31, 32, 33, 100, 257):
        x = rng.integers(0, 4, n, dtype=np.uint8)
        score, ey, ex, ops = oracle.semiglobal_xdrop(x, x)
        assert (score, ey, ex) == (n, n, n) and np.all(ops == 0) and ops.size == n
        score, ey, ex, ops = oracle.semiglobal_xdrop(np.zeros(n, np.uint8), np.ones(n, np.uint8))
        assert (score, ey, ex, ops.size) == (0, 0, 0, 0)
    # one deletion in the middle: 99 matches, one gap
    x = rng.integers(0, 4, 100, dtype=np.uint8)
    y = np.concatenate([np.delete(x, 50), [x[49] ^ 1]])
    score, ey, ex, ops = oracle.semiglobal_xdrop(x, y)
    assert score == 98 and (ops != 0).sum() == 1 and ey - ex == 1


def test_on_the_xdrop_threshold_the_scalar_reference_is_the_specification(oracle):
    """A stretch that matches on no diagonal, of exactly the length at which the X-drop decides: the restatement must
    equal the reference's SCALAR aligner (source.cpp:1836-1976).  The reference's AVX2 forms are asserted equal to it
    only on TestSemiGlobal's inputs (source.cpp:2774-2784); on these inputs some of them give a different answer
    (e.g. _simd_mark4 ends at (1,1) where the scalar aligns to the end), so they cannot all be matched -- the B200
    kernels follow the scalar, as this oracle does."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    rng = np.random.default_rng(5)
    cases = []
    for prefix in (0, 1, 2, 3):
        for run in (68, 69, 70, 71, 72):
            a = rng.integers(0, 4, 16384, dtype=np.uint8)
            b = a.copy()
            a[prefix:prefix + run] = 0
            b[prefix:prefix + run] = 1
            cases.append((a, b))

    def one(c):
        a, b = c
        s0, tb0 = oracle.ref_semiglobal(0, a, b)
        s, ey, ex, ops = oracle.semiglobal_xdrop(a, b)
        same = s == s0 and np.array_equal(oracle.ops_to_traceback(ops), tb0)
        s4, _ = oracle.ref_semiglobal(4, a, b)
        return same, s4 == s0, s0
    with ThreadPoolExecutor(8) as ex:
        out = list(ex.map(one, cases))
    assert all(o[0] for o in out)                      # restatement == scalar reference, score and traceback
    assert any(not o[1] for o in out)                  # ... while _simd_mark4 is NOT equal to the scalar on some of them
    assert any(o[2] > 16000 for o in out) and any(o[2] < 10 for o in out)      # both sides of the threshold occur
