"""GPU tests against the full-size pins under tests/golden/*_sums.json and reference_stream_10m.json (each written by
the script beside it from the unmodified reference or the oracle): the length-sweep batches, the reference test's own
10 M pairs, the counter-stream prefixes of the streaming mode and the semi-global bench batch.  The file sorts last on
purpose: these are the newest tests, and `pytest -x` should reach every older one first."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("L", [128, 256, 512])
def test_sweep_batch_head_equals_the_golden(ctx, swb, L):
    # tests/golden/sweep_sums.json (make_sweep_sums.py): the batches of `bench.py --workload sweep` -- the counter stream
    # re-cut to length L -- scored by the oracle; the bench compares the whole batch's score sum, this test its head.
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sweep_sums.json")) as f:
        want = json.load(f)["by_length"][str(L)]
    n = 4096
    a, b = swb.counter_pairs(0, n * (L // 128))
    got = ctx.score_batch(a.reshape(n, L), b.reshape(n, L), swb.MATRIX_SPEEDTEST, 15)
    assert [int(x) for x in got[:8]] == want["first_8_scores"]


def test_the_reference_tests_own_size_10m_pairs(ctx, swb):
    # TestSimdSmithWaterman runs 10 000 000 iterations of its stream (source.cpp:2947-2948).  Known answers for exactly
    # that: tests/golden/reference_stream_10m.json, written from the unmodified reference (make_reference_stream_10m.py).
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_stream_10m.json")) as f:
        want = json.load(f)["by_scoring"]
    n = 10_000_000
    a, b = swb.reference_stream(n)
    for key, (m, g) in (("speedtest_10_-30_15", (swb.MATRIX_SPEEDTEST, 15)), ("x32_1_-1_1", (swb.MATRIX_111, 1))):
        s = ctx.score_batch(a, b, m, g)
        w = want[key]
        assert w["pairs"] == n
        assert (int(s.sum(dtype=np.int64)), int(s.min()), int(s.max()), int(s.argmax())) == (w["sum"], w["min"], w["max"], w["first_argmax"])
        assert f"{swb.fnv1a64(s):016x}" == w["fnv1a64"]


@pytest.mark.parametrize("packed", [False, True])
def test_stream_prefix_sum_equals_the_reference_golden(ctx, swb, packed):
    # tests/golden/counter_stream_sums.json: sums of the UNMODIFIED reference's scores over prefixes of the counter
    # stream (make_counter_sums.py).  A sum does not depend on batch size, wire format or sharding, so the same file
    # pins the full 100 M-pair configuration (bench.py --workload stream; profiles/r01/stream_100m_*_1gpu.json).
    import json
    from streaming import StreamRunner
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "counter_stream_sums.json")) as f:
        want = json.load(f)["sum_of_scores_over_prefix"]
    r = StreamRunner(ctx, batch_pairs=700_000, n_buffers=3, packed=packed, gen_threads=min(16, os.cpu_count() or 1))
    try:
        a = r.run(0, 2_000_000, swb.MATRIX_SPEEDTEST, 15)
        b = r.run(0, 2_000_000, swb.MATRIX_111, 1)
    finally:
        r.close()
    assert a.score_sum == want["speedtest_10_-30_15"]["2000000"]
    assert b.score_sum == want["x32_1_-1_1"]["2000000"]


def test_prefix_2048_of_the_bench_batch_equals_the_golden_sums(ctx, swb):
    # tests/golden/semiglobal_batch_sums.json (make_semiglobal_batch_sums.py): sums over the first 2048 pairs of the bench
    # batch computed with the oracle restatement -- scores, end cells, op counts and the position-weighted op sum.
    # bench.py --workload semiglobal makes the same comparison on its whole 37888-pair batch.
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "semiglobal_batch_sums.json")) as f:
        want = json.load(f)["prefix"]["2048"]
    a, b = swb.related_pairs(0, 2048, 16384)
    r = ctx.semiglobal_xdrop(a, b)
    got = {k: int(np.asarray(r[k]).sum(dtype=np.int64)) for k in ("score", "end_y", "end_x", "n_ops")}
    w = np.arange(1, r["ops"].shape[1] + 1, dtype=np.uint64)
    total = 0
    for r0 in range(0, 2048, 256):          # in blocks: the whole op array as uint64 would be half a gigabyte
        mask = np.arange(r["ops"].shape[1])[None, :] < np.asarray(r["n_ops"])[r0:r0 + 256, None]
        total += int((r["ops"][r0:r0 + 256].astype(np.uint64) * mask * w[None, :]).sum(dtype=np.uint64))
    got["ops_weighted"] = total
    assert got == want
