"""The wavefront schedule of the one-warp-per-pair latency kernels (csrc/sw_pair_kernel.cuh), restated in numpy and
checked against the oracle in the GPU-less container: lane l owns rows 4l..4l+3, step t computes cells (4l+k, t-4l-k),
the boundary row travels to the next lane one step later, columns outside the matrix carry a substitution score of
-128 instead of a predicate, and all values live in the anti-diagonal offset frame.  This pins the ALGORITHM (in
particular: that the -128 padding and the frame can never change a score, at gap 0 and at the domain's corners); the CUDA
text itself is checked on the GPU (tests/test_parity_gpu.py).
A test tool, not a fallback: nothing in the product imports this."""
import numpy as np
import pytest

PAD = 124
COLS = 384
OUTSIDE = -128


def wavefront_score(a, b, sm, g):
    """The kernel's sweep, statement by statement: int32, the anti-diagonal OFFSET frame H^ = H + g*t (t = row + column =
    the step), so that a gap step costs nothing and a cell is max3(max(diag^ + s'', up^), left^, Z) with s'' = s + 2g
    and Z = g*t the value of a true zero; the running best and Z move with the frame (+g per step)."""
    S = np.asarray(sm, dtype=np.int64).reshape(4, 4)
    tab = np.full((4, COLS), OUTSIDE + 2 * g, dtype=np.int64)          # score table: row = query base, column = target position
    tab[:, PAD:PAD + 128] = S[:, b & 3] + 2 * g
    lanes = np.arange(32)
    rows = (a & 3).reshape(32, 4)
    top = (lanes == 0).astype(np.int64)
    h1 = np.full((32, 4), -g, np.int64)               # step -1: true zeros in the frame of step -1 ...
    h2 = np.full((32, 4), -2 * g, np.int64)           # ... and of step -2
    up0 = -g * (1 - top)
    dg0 = np.full(32, -2 * g, np.int64)
    zm = -2 * g * top                                 # lane 0 only: Z(t-1), what row -1 holds as next step's diagonal neighbour
    best = np.full(32, -g, np.int64)
    Z = -g
    for t in range(256):
        Z += g
        best = best + g
        zm = zm + g * top
        hn = np.empty((32, 4), np.int64)
        for k in (3, 2, 1, 0):
            s = tab[rows[:, k], PAD - 4 * lanes + t - k]
            up = h1[:, k - 1] if k else up0
            dg = h2[:, k - 1] if k else dg0
            hn[:, k] = np.maximum(np.maximum(np.maximum(dg + s, up), h1[:, k]), Z)
        best = np.maximum(best, hn.max(axis=1))
        sh = np.concatenate([hn[:1, 3], hn[:-1, 3]])  # __shfl_up of this step's bottom row (lane 0 gets its own value back)
        dg0 = up0 + zm                                # next step's diagonal neighbour = this step's upper one; lane 0: exactly Z(t-1)
        up0 = sh * (1 - top)                          # lane 0: row -1; as an UPPER neighbour any value <= Z will do: 0
        h2, h1 = h1, hn
    return int((best - Z).max())


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
