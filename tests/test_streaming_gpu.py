"""GPU tests of the streaming mode: pairs generated on host
threads into pinned ring buffers, streamed through swb200_submit/_wait, verified on a strided
sample against the reference's simd4 (oracle/_ref) or the oracle port."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("packed", [False, True])
def test_stream_8m_pairs_sampled_against_oracle(ctx, swb, oracle, packed):
    from streaming import StreamRunner
    total, first, batch = 8_000_000 + 12_345, 3_000_000_000, 1 << 20   # indices beyond 2^31: 64-bit pair counters
    stride = 4001
    picked = {}

    def on_batch(s0, scores):
        idx = np.arange((-s0) % stride, scores.size, stride)
        picked[s0] = (idx + s0, scores[idx].copy())

    runner = StreamRunner(ctx, batch_pairs=batch, n_buffers=3, packed=packed, gen_threads=min(16, os.cpu_count() or 1))
    try:
        rep = runner.run(first, total, swb.MATRIX_SPEEDTEST, 15, on_batch=on_batch)
    finally:
        runner.close()
    assert rep.pairs == total and rep.batches == -(-total // batch)
    gidx = np.concatenate([v[0] for v in picked.values()])
    got = np.concatenate([v[1] for v in picked.values()])
    assert gidx.size >= total // stride
    a = np.empty((gidx.size, 128), np.uint8)
    b = np.empty((gidx.size, 128), np.uint8)
    for j, k in enumerate(gidx):      # any index is addressable on its own: regenerate just the sampled pairs
        aj, bj = swb.counter_pairs(int(k), 1, threads=1)
        a[j], b[j] = aj[0], bj[0]
    if oracle.have_ref():
        exp = oracle.ref_score_batch(4, a, b, swb.MATRIX_SPEEDTEST, 15, threads=os.cpu_count() or 1)   # reference simd4
    else:
        exp = oracle.score_batch(a, b, swb.MATRIX_SPEEDTEST, 15, threads=os.cpu_count() or 1)
    assert np.array_equal(got, exp)
    # iid 128-mers under +10/-30/15 average 75.5
    assert 75.0 < rep.score_sum / rep.pairs < 76.0


def test_stream_is_deterministic_and_order_independent(ctx, swb):
    from streaming import StreamRunner
    r1 = StreamRunner(ctx, batch_pairs=1 << 19, n_buffers=2, gen_threads=4)
    r2 = StreamRunner(ctx, batch_pairs=300_000, n_buffers=3, gen_threads=7, packed=True)
    try:
        a = r1.run(10, 2_000_000, swb.MATRIX_111, 1)
        b = r2.run(10, 2_000_000, swb.MATRIX_111, 1)
    finally:
        r1.close(); r2.close()
    assert a.score_sum == b.score_sum and a.pairs == b.pairs == 2_000_000   # batch size, wire format, thread count: no effect
