"""The wavefront schedule of the one-warp-per-pair latency kernel (csrc/sw_pair_kernel.cuh), restated in numpy and
checked against the oracle in the GPU-less container: lane l owns rows 4l..4l+3, step t computes cells (4l+k, t-4l-k),
the boundary row travels to the next lane one step later, and columns outside the matrix carry a substitution score of
-128 instead of a predicate.  This pins the ALGORITHM (in particular: that the -128 padding can never change a score,
at gap 0 and at the domain's corners); the CUDA text itself is checked on the GPU (tests/test_parity_gpu.py).
A test tool, not a fallback: nothing in the product imports this."""
import numpy as np
import pytest

PAD = 124


def wavefront_score(a, b, sm, g):
    S = np.full((4, 5), -128, dtype=np.int64)
    S[:, :4] = np.asarray(sm, dtype=np.int64).reshape(4, 4)
    sel = np.full(384, 4, dtype=np.int64)            # code 4 = "outside the matrix"
    sel[PAD:PAD + 128] = b & 3
    lanes = np.arange(32)
    rows = (a & 3).reshape(32, 4)
    h1 = np.zeros((32, 4), np.int64)
    h2 = np.zeros((32, 4), np.int64)
    up0 = np.zeros(32, np.int64)
    best = np.zeros(32, np.int64)
    w = np.full((32, 4), 4, np.int64)
    nxt = sel[PAD - 4 * lanes]
    for t in range(256):
        w[:, 1:] = w[:, :-1].copy()
        w[:, 0] = nxt
        nxt = sel[PAD - 4 * lanes + t + 1]
        dg0 = up0
        up0 = np.concatenate([[0], h1[:-1, 3]])      # __shfl_up of the previous step's bottom row; lane 0 sees row -1 = 0
        hn = np.empty((32, 4), np.int64)
        for k in range(4):
            s = S[rows[:, k], w[:, k]]
            up = h1[:, k - 1] if k else up0
            dg = h2[:, k - 1] if k else dg0
            hn[:, k] = np.maximum(np.maximum(dg + s, np.maximum(up, h1[:, k]) - g), 0)
        best = np.maximum(best, hn.max(axis=1))
        h2, h1 = h1, hn
    return int(best.max())


def mm(match, mismatch):
    return [match if i == j else mismatch for i in range(4) for j in range(4)]


@pytest.mark.parametrize("sm,g", [(mm(10, -30), 15), (mm(1, -1), 1), (mm(127, -127), 127), (mm(127, -1), 0), (mm(5, 3), 0),
                                  (mm(0, -5), 3), (list(range(-8, 8)), 2)])
def test_wavefront_schedule_equals_the_oracle(oracle, sm, g):
    rng = np.random.default_rng(20261018)
    for it in range(12):
        a = rng.integers(0, 4, 128).astype(np.uint8)
        if it % 4 == 0:
            b = rng.integers(0, 4, 128).astype(np.uint8)
        elif it % 4 == 1:
            b = a.copy()
        elif it % 4 == 2:
            b = a.copy()
            b[rng.integers(0, 128, 6)] = rng.integers(0, 4, 6)
        else:
            b = np.roll(a, 5)
        assert wavefront_score(a, b, sm, g) == int(oracle.score_batch(a[None], b[None], sm, g)[0])
