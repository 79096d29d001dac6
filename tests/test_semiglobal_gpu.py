"""GPU parity of the adaptive-banded X-drop semi-global aligner (csrc/sg_kernel.cuh) through the C ABI:
against the fixtures generated from the reference (scalar + its four AVX2 forms), against the oracle on
seeded inputs at many lengths, and -- at sizes the oracle does not cover in seconds -- through properties."""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from sg_common import fnv1a64_bytes, load_cases, ops_to_traceback, xdrop_edge_cases

pytestmark = pytest.mark.gpu
NCPU = min(os.cpu_count() or 1, 16)


def related_pairs(rng, n, length, sub=0.1, ins=0.1, dele=0.1):
    """TestSemiGlobal-style inputs (source.cpp:2750-2771) at any length, from numpy's generator."""
    a = rng.integers(0, 4, (n, length), dtype=np.uint8)
    b = np.empty_like(a)
    for i in range(n):
        p = rng.random(3 * length)
        out, j, k = [], 0, 0
        while len(out) < length:
            if j >= length:
                out.append(rng.integers(0, 4))
                continue
            q = p[k]; k += 1
            if q < sub: out.append(rng.integers(0, 4)); j += 1
            elif q < sub + ins: out.append(rng.integers(0, 4))
            elif q < sub + ins + dele: j += 1
            else: out.append(a[i, j]); j += 1
        b[i] = out
    return a, b


def oracle_batch(oracle, a, b):
    with ThreadPoolExecutor(NCPU) as ex:
        return list(ex.map(lambda i: oracle.semiglobal_xdrop(a[i], b[i]), range(a.shape[0])))


def check_against_oracle(oracle, got, a, b, idx=None):
    idx = range(a.shape[0]) if idx is None else idx
    exp = oracle_batch(oracle, a[list(idx)], b[list(idx)])
    for (score, ey, ex, ops), i in zip(exp, idx):
        assert (got["score"][i], got["end_y"][i], got["end_x"][i]) == (score, ey, ex), i
        if "ops" in got:
            assert got["n_ops"][i] == ops.size, i
            assert np.array_equal(got["ops"][i, :ops.size], ops), i


def test_golden_cases_from_the_reference(ctx):
    cases = load_cases()
    a = np.stack([c["seq1"] for c in cases])
    b = np.stack([c["seq2"] for c in cases])
    launches0 = ctx.launch_count
    r = ctx.semiglobal_xdrop(a, b)
    assert ctx.launch_count == launches0 + 3                    # the sm_100a kernels ran: forward, traceback, left-align, once for the whole batch
    for i, c in enumerate(cases):
        assert (r["score"][i], r["end_y"][i], r["end_x"][i], r["n_ops"][i]) == (c["score"], c["end_y"], c["end_x"], c["n_ops"]), c["name"]
        ops = r["ops"][i, :r["n_ops"][i]]
        assert np.array_equal(ops, c["ops"]), c["name"]
        assert f"{fnv1a64_bytes(ops_to_traceback(ops).tobytes()):016x}" == c["traceback_fnv1a64"], c["name"]
    # score-only mode skips the traceback and agrees
    launches0 = ctx.launch_count
    s = ctx.semiglobal_xdrop(a, b, traceback=False)
    assert ctx.launch_count == launches0 + 1                    # forward kernel only
    assert "ops" not in s and all(np.array_equal(s[k], r[k]) for k in ("score", "end_y", "end_x"))


def test_reference_build_beside_it(ctx, oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref did not travel")
    a, b = oracle.ref_semiglobal_test_inputs(40, seed=4242)
    r = ctx.semiglobal_xdrop(a, b)
    for i in range(40):
        score, tb = oracle.ref_semiglobal(4 if i % 2 else 0, a[i], b[i])     # scalar and _simd_mark4 alternately
        assert r["score"][i] == score
        assert np.array_equal(ops_to_traceback(r["ops"][i, :r["n_ops"][i]]), tb), i


@pytest.mark.parametrize("length", [1, 2, 3, 31, 32, 33, 64, 100, 1000, 4097])
def test_lengths_against_oracle(ctx, oracle, length):
    rng = np.random.default_rng(100 + length)
    n = 64
    a, b = related_pairs(rng, n, length)
    a[0] = b[0]                                   # identical
    a[1] = 0; b[1] = 1                            # nothing matches
    a[2] = 2; b[2] = 2                            # homopolymer: all ties
    b[3] = rng.integers(0, 4, length)             # unrelated
    r = ctx.semiglobal_xdrop(a, b)
    check_against_oracle(oracle, r, a, b)
    assert r["score"][0] == length and r["score"][1] == 0 and r["n_ops"][1] == 0


def test_more_pairs_than_resident_warps(ctx, oracle):
    # the forward grid is persistent (SMs x resident blocks of one warp; a batch this large runs 32 pairs per warp): with
    # more pairs than that every warp takes a second batch; 777 is not a multiple of 32, so the last warp has lanes beyond
    # the batch (they shadow the last pair into the spare group) and the last traceback warp is partial
    info = ctx.semiglobal_kernel_info()
    resident = 32 * info["sm_count"] * info["blocks_per_sm"] * info["threads_per_block"] // 32
    n = resident + 777
    rng = np.random.default_rng(9)
    length = 96
    a = rng.integers(0, 4, (n, length), dtype=np.uint8)
    b = np.where(rng.random((n, length)) < 0.85, a, rng.integers(0, 4, (n, length), dtype=np.uint8)).astype(np.uint8)
    b[::7] = np.roll(a[::7], 5, axis=1)           # shifted copies: the band has to wander
    r = ctx.semiglobal_xdrop(a, b)
    idx = np.r_[0:300, resident - 150:resident + 150, n - 300:n]
    check_against_oracle(oracle, r, a, b, idx.tolist())
    # properties on all pairs: a path from (0,0) to the end cell whose moves re-score to the reported score
    ops, n_ops = r["ops"], r["n_ops"]
    valid = np.arange(ops.shape[1])[None, :] < n_ops[:, None]
    assert np.array_equal(((ops != 2) & valid).sum(1), r["end_y"]) and np.array_equal(((ops != 1) & valid).sum(1), r["end_x"])
    for i in range(0, n, 97):
        o = ops[i, :n_ops[i]]
        y = np.cumsum(o != 2); x = np.cumsum(o != 1)
        d = o == 0
        sc = np.where(a[i, y[d] - 1] == b[i, x[d] - 1], 1, -1).sum() - int((~d).sum())
        assert sc == r["score"][i], i


def test_staging_slots_are_reused(ctx, oracle):
    # len = 2048: the host path cuts a batch of more than four times 128 pairs per SM into five chunks over four
    # independent slots; the fifth chunk reuses slot 0 after its first chunk is back on the host
    n = 4 * 128 * ctx.semiglobal_kernel_info()["sm_count"] + 5003
    length = 2048
    rng = np.random.default_rng(78)
    a = rng.integers(0, 4, (n, length), dtype=np.uint8)
    b = np.where(rng.random((n, length)) < 0.88, a, rng.integers(0, 4, (n, length), dtype=np.uint8)).astype(np.uint8)
    b[::9] = np.roll(a[::9], -7, axis=1)
    r = ctx.semiglobal_xdrop(a, b)
    c = -(-n // 5)                                   # the chunk size the library arrives at: five equal parts
    idx = np.r_[0:40, c - 20:c + 20, 3 * c - 20:3 * c + 20, 4 * c - 20:4 * c + 60, n - 40:n].tolist()
    check_against_oracle(oracle, r, a, b, idx)
    ops, n_ops = r["ops"], r["n_ops"]
    valid = np.arange(ops.shape[1])[None, :] < n_ops[:, None]
    assert np.array_equal(((ops != 2) & valid).sum(1), r["end_y"]) and np.array_equal(((ops != 1) & valid).sum(1), r["end_x"])


def test_reference_shape_batch_spans_chunks(ctx, oracle):
    # len = 16384: two staging chunks (half a wave of the forward kernel each), the second one partial
    info = ctx.semiglobal_kernel_info()
    wave = info["sm_count"] * 32
    n = 2 * wave + 211
    rng = np.random.default_rng(77)
    a = rng.integers(0, 4, (n, 16384), dtype=np.uint8)
    b = np.where(rng.random((n, 16384)) < 0.9, a, rng.integers(0, 4, (n, 16384), dtype=np.uint8)).astype(np.uint8)
    for i in range(0, n, 5):
        k = int(rng.integers(1, 30))
        b[i] = np.concatenate([a[i, k:], rng.integers(0, 4, k, dtype=np.uint8)])      # a deletion of k bases up front
    r = ctx.semiglobal_xdrop(a, b)
    idx = np.r_[0:30, wave // 2 - 20:wave // 2 + 20, wave - 20:wave + 20, 2 * wave - 50:2 * wave - 10, n - 30:n].tolist()
    check_against_oracle(oracle, r, a, b, idx)
    s = ctx.semiglobal_xdrop(a, b, traceback=False)
    assert np.array_equal(s["score"], r["score"]) and np.array_equal(s["end_y"], r["end_y"])
    assert r["score"].min() > 8000                 # 90 % identity over 16384 bases


def test_device_resident_entry_and_errors(ctx, swb, oracle):
    import torch
    cases = load_cases()[:5]
    a = torch.from_numpy(np.stack([c["seq1"] for c in cases])).cuda()
    b = torch.from_numpy(np.stack([c["seq2"] for c in cases])).cuda()
    n = len(cases)
    sc, ey, ex, no = (torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(4))
    ops = torch.empty((n, 2 * 16384), dtype=torch.uint8, device="cuda")
    ctx.semiglobal_xdrop_device(a, b, sc, ey, ex, no, ops)
    torch.cuda.synchronize()
    for i, c in enumerate(cases):
        assert (int(sc[i]), int(ey[i]), int(ex[i]), int(no[i])) == (c["score"], c["end_y"], c["end_x"], c["n_ops"])
        assert np.array_equal(ops[i, :c["n_ops"]].cpu().numpy(), c["ops"])
    with pytest.raises(swb.SwbError) as e:
        ctx.semiglobal_xdrop(np.zeros((2, 40000), np.uint8), np.zeros((2, 40000), np.uint8))
    assert e.value.code == swb.ERR_ARG
    score, tb = swb.SemiGlobal_AdaptiveBanded_XDrop_111_32_70_b200(cases[0]["seq1"], cases[0]["seq2"])
    assert score == cases[0]["score"] and np.array_equal(np.array(tb, np.int32), ops_to_traceback(cases[0]["ops"]))


def test_pairs_on_the_xdrop_thresholds_beside_live_pairs(ctx, oracle):
    # Pairs built to sit on the aligner's thresholds (sg_common.xdrop_edge_cases): many of them end early, in the same
    # warp as pairs that run on.  A finished pair must stay finished while its warp keeps running rounds -- the round
    # after its last one could revive a cell through the diagonal, which the reference never computes (source.cpp:1938).
    rng = np.random.default_rng(7070)
    by_len = {}
    for a, b in xdrop_edge_cases(rng):
        by_len.setdefault(a.size, []).append((a, b))
    assert len(by_len) >= 3
    for length, cs in by_len.items():
        a = np.stack([c[0] for c in cs]); b = np.stack([c[1] for c in cs])
        perm = rng.permutation(len(cs))                      # early enders and survivors side by side
        a, b = a[perm], b[perm]
        r = ctx.semiglobal_xdrop(a, b)
        check_against_oracle(oracle, r, a, b)
        assert (r["end_y"] < length // 2).sum() > 5 and (r["end_y"] > length // 2).sum() > 5


def test_adversarial_inputs_against_oracle(ctx, oracle):
    # low-complexity alphabets, periodic repeats, runs of mismatches and indels of every length around the band width
    # (32) and the X-drop (70) -- the generator of tools/sg_fuzz.py, cut to fixed lengths; every pair checked
    import importlib.util
    spec = importlib.util.spec_from_file_location("sg_fuzz", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "sg_fuzz.py"))
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    rng = np.random.default_rng(424242)
    for length in (37, 150, 420):
        aa, bb = [], []
        while len(aa) < 1500:
            a, b = fz.make_case(rng)
            if a.size >= length:
                aa.append(a[:length]); bb.append(b[:length])
            elif a.size >= 8:
                reps = -(-length // a.size)
                aa.append(np.tile(a, reps)[:length]); bb.append(np.tile(b, reps)[:length])
        a = np.stack(aa); b = np.stack(bb)
        r = ctx.semiglobal_xdrop(a, b)
        check_against_oracle(oracle, r, a, b)


def test_all_visible_gpus_share_a_batch(swb, oracle):
    # pairs are independent -- contiguous index ranges per GPU, every GPU writes its own slice of the
    # caller's arrays, no collective.  (With one visible GPU this is the single-GPU path.)
    import torch
    g = torch.cuda.device_count()
    rng = np.random.default_rng(31337)
    n, length = 2501, 640
    a = rng.integers(0, 4, (n, length), dtype=np.uint8)
    b = np.where(rng.random((n, length)) < 0.9, a, rng.integers(0, 4, (n, length), dtype=np.uint8)).astype(np.uint8)
    b[::11] = np.roll(a[::11], 9, axis=1)
    with swb.Context(n_devices=g) as c:
        assert c.n_devices == g
        r = c.semiglobal_xdrop(a, b)
    check_against_oracle(oracle, r, a, b, list(range(0, n, 7)) + list(range(n - 40, n)))


def test_score_only_calls_allocate_no_round_record_scratch(swb):
    # A score-only call (ops == NULL) runs the RECORD = false forward kernel, which never touches the round records:
    # on a FRESH context neither the host entry nor the device entry may allocate that scratch (512 KiB per pair at
    # 16384 bases -- 2 GiB for this batch, per slot).  Device memory in use may grow by the input staging only (~0.7 GiB).
    import torch
    n, length = 4096, 16384
    a, b = swb.related_pairs(0, n, length)
    with swb.Context(n_devices=1) as c:
        torch.cuda.synchronize()
        free0, _ = torch.cuda.mem_get_info()
        s = c.semiglobal_xdrop(a, b, traceback=False)
        da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        out = [torch.empty(n, dtype=torch.int32, device="cuda") for _ in range(3)]
        c.semiglobal_xdrop_device(da, db, out[0], out[1], out[2])
        torch.cuda.synchronize()
        free1, _ = torch.cuda.mem_get_info()
        assert np.array_equal(out[0].cpu().numpy(), s["score"])
        used = free0 - free1
        assert used < (3 << 29), f"{used / 2**20:.0f} MiB taken by a score-only batch: the round-record scratch was allocated"


def test_host_pipeline_on_the_compressed_wire(ctx, swb, oracle, monkeypatch):
    # Large host batches cross the link four to a byte in both directions (csrc/sg_pipe.inc): lanes pack the sequences,
    # the device expands them, aligns, packs the move strings, lanes expand those into the caller's rows.  Test knobs make
    # a 700-pair batch of 1024-mers run as eleven chunks round a ring of three slots (every staging buffer is reused),
    # from PAGEABLE arrays; results equal the oracle, the plain chunk pipeline, and the score-only form.
    a, b = swb.related_pairs(4242, 700, 1024)
    monkeypatch.setenv("SWB200_SG_CHUNK_PAIRS", "64")
    monkeypatch.setenv("SWB200_SG_SLOTS", "3")
    l0 = ctx.launch_count
    r = ctx.semiglobal_xdrop(a, b)
    assert ctx.launch_count - l0 == 11 * 6          # per chunk: two expansions, forward, traceback, left-align, move packing
    check_against_oracle(oracle, r, a, b)
    l1 = ctx.launch_count
    s = ctx.semiglobal_xdrop(a, b, traceback=False)
    assert ctx.launch_count - l1 == 11 * 3          # two expansions and the forward kernel
    for k in ("score", "end_y", "end_x"):
        assert np.array_equal(s[k], r[k]), k
    monkeypatch.setenv("SWB200_SG_PIPE", "0")
    l2 = ctx.launch_count
    q = ctx.semiglobal_xdrop(a, b)
    assert ctx.launch_count - l2 < 11 * 6           # the plain chunk pipeline: a few large chunks
    for k in ("score", "end_y", "end_x", "n_ops"):
        assert np.array_equal(q[k], r[k]), k
    for i in range(700):
        assert np.array_equal(q["ops"][i, :q["n_ops"][i]], r["ops"][i, :r["n_ops"][i]]), i
    # a ragged last chunk and a last piece shorter than 32 pairs; two lanes only
    monkeypatch.delenv("SWB200_SG_PIPE")
    monkeypatch.setenv("SWB200_SG_THREADS", "2")
    a2, b2 = a[:64 * 2 + 13], b[:64 * 2 + 13]
    check_against_oracle(oracle, ctx.semiglobal_xdrop(a2, b2), a2, b2)


def test_all_visible_gpus_run_the_compressed_wire_pipeline(swb, oracle, monkeypatch):
    # the host pipeline (csrc/sg_pipe.inc) under the library's own sharding layer: every visible GPU takes a contiguous index
    # range and runs its own ring of slots with its own lanes, all writing slices of the caller's (pageable) arrays
    import torch
    g = torch.cuda.device_count()
    monkeypatch.setenv("SWB200_SG_CHUNK_PAIRS", "96")
    monkeypatch.setenv("SWB200_SG_SLOTS", "2")
    monkeypatch.setenv("SWB200_SG_THREADS", "3")
    n, length = 96 * 5 * g + 37, 512
    a, b = swb.related_pairs(99, n, length)
    with swb.Context(n_devices=g) as c:
        l0 = c.launch_count
        r = c.semiglobal_xdrop(a, b)
        assert c.launch_count - l0 >= 6 * 5 * g              # the pipeline ran: six launches per chunk, at least five chunks per GPU
    check_against_oracle(oracle, r, a, b, list(range(0, n, 5)) + list(range(n - 40, n)))


def test_pipeline_ring_reuse_at_its_default_geometry(ctx, swb, oracle):
    # the default chunk size (one forward warp per SM) and slot count: twenty chunks round sixteen slots, so the ring is reused
    # with real events in flight; short sequences keep it quick.  Pinned and pageable callers' arrays give the same results.
    info = ctx.semiglobal_kernel_info()
    chunk = info["sm_count"] * 32
    n, length = chunk * 20 + 333, 64
    a, b = swb.related_pairs(777, n, length)
    l0 = ctx.launch_count
    r = ctx.semiglobal_xdrop(a, b)
    assert ctx.launch_count - l0 == 21 * 6
    idx = list(range(0, n, 997)) + list(range(chunk * 16 - 20, chunk * 16 + 20)) + list(range(n - 50, n))
    check_against_oracle(oracle, r, a, b, idx)
    pa, pb = swb.PinnedArray((n, length), np.uint8), swb.PinnedArray((n, length), np.uint8)
    pa.array[:] = a
    pb.array[:] = b
    q = ctx.semiglobal_xdrop(pa.array, pb.array)
    for k in ("score", "end_y", "end_x", "n_ops"):
        assert np.array_equal(q[k], r[k]), k
    rows = np.arange(n)[:, None]
    mask = np.arange(2 * length)[None, :] < r["n_ops"][:, None]
    assert np.array_equal(np.where(mask, q["ops"], 0), np.where(mask, r["ops"], 0))
    pa.free(); pb.free()
