"""GPU parity tests (run on the B200 box with -m gpu).  Every score comes out of the CUDA
kernel THROUGH THE C ABI (libswb200.so via ctypes) and is compared bit-exactly with
  * the committed golden fixtures (generated from the unmodified reference),
  * the oracle (oracle/sw_oracle.c) on the same seeded inputs,
  * the reference build itself (oracle/_ref) where it travelled with the snapshot,
and, at the full 1 M-pair size of BASELINE.json's configs[1], with the checksum the
reference's scalar kernel produces (FNV-1a-64 ae56a1e6a1d57492)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NCPU = os.cpu_count() or 1


def mm(match, mismatch):
    return [match if i == j else mismatch for i in range(4) for j in range(4)]


def test_extension_is_loaded_and_targets_b200(ctx, swb):
    import torch
    assert torch.cuda.is_available()
    assert torch.cuda.get_device_capability(0)[0] == 10
    info = ctx.kernel_info(swb.MATRIX_SPEEDTEST, swb.GAP_SPEEDTEST)
    assert info["fast_path"] == 1 and info["sm_count"] >= 100 and info["blocks_per_sm"] >= 1
    assert ctx.kernel_info(mm(127, -127), 127)["fast_path"] == 0


def test_per_pair_call_matches_known_answers(ctx, swb, oracle):
    # the reference's per-pair signature (source.cpp:462-466) on the first pairs of its stream
    a, b = oracle.reference_stream(16)
    got = [ctx.smith_waterman(a[i], b[i], swb.MATRIX_SPEEDTEST, swb.GAP_SPEEDTEST) for i in range(16)]
    assert got == [80, 80, 70, 95, 70, 80, 80, 75, 80, 70, 75, 65, 65, 80, 100, 70]
    assert swb.SmithWaterman_b200(a[3], b[3], swb.MATRIX_SPEEDTEST, 15) == 95


def test_first4096_golden_both_matrices(ctx, swb, oracle):
    a, b = oracle.reference_stream(4096)
    for name, sm, g in (("speedtest_10_-30_15", swb.MATRIX_SPEEDTEST, 15), ("x32_1_-1_1", swb.MATRIX_111, 1)):
        exp = np.load(os.path.join(os.path.dirname(__file__), "golden", f"stream_first4096_{name}.npy")).astype(np.int32)
        assert np.array_equal(ctx.score_batch(a, b, sm, g), exp), name


def test_structured_golden_all_param_sets_all_three_kernels(ctx, golden):
    # 480 structured pairs x 14 parameter sets (domain corners included) through every kernel that can score them:
    # the one-warp-per-pair latency kernel (what a host batch this small takes by default), and the throughput
    # kernel in its fast (offset) and general form.
    z = golden["structured_npz"]
    launches0 = ctx.launch_count
    try:
        for latency, force_general in ((True, False), (False, False), (False, True)):
            ctx.set_latency_path(latency)
            ctx.set_force_general(force_general)
            for ps in golden["structured"]["param_sets"]:
                got = ctx.score_batch(z["seq1"], z["seq2"], ps["matrix"], ps["gap"])
                assert np.array_equal(got, z[ps["name"]].astype(np.int32)), (ps["name"], latency, force_general)
    finally:
        ctx.set_force_general(False)
        ctx.set_latency_path(True)
    assert ctx.launch_count - launches0 == 3 * len(golden["structured"]["param_sets"])


def test_per_pair_call_over_the_whole_domain(ctx, swb, oracle, golden):
    # swb200_score_pair (the pair rides in the launch parameters) and a 2-pair batch (mapped slot) on structured pairs,
    # every parameter set of the golden file: equal to the fixture
    z = golden["structured_npz"]
    idx = list(range(0, z["seq1"].shape[0], 37))
    for ps in golden["structured"]["param_sets"]:
        want = z[ps["name"]].astype(np.int32)
        got = [ctx.smith_waterman(z["seq1"][i], z["seq2"][i], ps["matrix"], ps["gap"]) for i in idx]
        assert got == [int(want[i]) for i in idx], ps["name"]
        two = ctx.score_batch(z["seq1"][5:7], z["seq2"][5:7], ps["matrix"], ps["gap"])
        assert np.array_equal(two, want[5:7]), ps["name"]
    # codes above 3 are the caller's error; the kernels mask them and never read out of bounds
    bad = np.full(128, 7, np.uint8)
    assert ctx.smith_waterman(bad, bad, swb.MATRIX_SPEEDTEST, 15) == 1280


def test_per_pair_doorbell_resident_server(ctx, swb, oracle):
    # swb200_score_pair rings the doorbell of a resident one-warp server kernel (csrc/sw_pair_kernel.cuh, pairpath.inc):
    # a run of calls launches it once; after a pause longer than its linger time it has left and the next call launches
    # a new one; matrices and gaps change from call to call; several threads may call at once; and every score equals
    # the launch-per-call form of the same kernel (mode 2) and the oracle.
    import threading, time
    a, b = swb.counter_pairs(777, 600)
    sets = [(swb.MATRIX_SPEEDTEST, 15), (swb.MATRIX_111, 1), (mm(127, -127), 127), (mm(5, 3), 0), (list(range(-8, 8)), 2)]
    want = {k: oracle.score_batch(a, b, sm, g, threads=NCPU) for k, (sm, g) in enumerate(sets)}
    ctx.smith_waterman(a[0], b[0], *sets[0])                  # a server is resident from here on
    l0 = ctx.launch_count
    got = [ctx.smith_waterman(a[i], b[i], *sets[0]) for i in range(600)]
    assert got == [int(x) for x in want[0]]
    assert ctx.launch_count - l0 <= 20, "a run of per-pair calls must not launch per call"
    for i in range(120):                                      # parameters change with every call
        k = i % len(sets)
        assert ctx.smith_waterman(a[i], b[i], *sets[k]) == int(want[k][i]), (i, k)
    l1 = ctx.launch_count
    for i in range(12):                                       # the server leaves during the pause; the call launches the next one
        time.sleep(0.004)
        assert ctx.smith_waterman(a[i], b[i], *sets[1]) == int(want[1][i])
    assert ctx.launch_count - l1 >= 6
    errors = []
    def worker(k):
        try:
            for i in range(k, 400, 4):
                if ctx.smith_waterman(a[i], b[i], *sets[0]) != int(want[0][i]):
                    errors.append(i)
        except Exception as e:                                # pragma: no cover
            errors.append(repr(e))
    ths = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert not errors
    try:
        ctx.set_latency_path(2)                               # one launch per call, the pair in the launch parameters
        l2 = ctx.launch_count
        assert [ctx.smith_waterman(a[i], b[i], *sets[2]) for i in range(50)] == [int(x) for x in want[2][:50]]
        assert ctx.launch_count - l2 == 50
    finally:
        ctx.set_latency_path(True)
    # a batch call right after per-pair calls (the resident server is still there) is not held up by it for long
    t0 = time.perf_counter()
    assert np.array_equal(ctx.score_batch(a, b, *sets[0]), want[0])
    assert time.perf_counter() - t0 < 0.5


def test_latency_kernels_on_random_parameters_of_the_whole_domain(ctx, swb, oracle):
    # the one-warp sweep (offset frame in int32, score table in shared memory) over the reference's whole parameter domain:
    # random 4 x 4 matrices with entries in [-127, 127] (asymmetric, all-positive, all-negative included), gaps in [0, 127],
    # on random, identical, shifted and low-complexity pairs -- as a batch (sw_pair_kernel) and pair by pair through the
    # doorbell (sw_pair_server), whose table of out-of-matrix scores depends on the gap and is rebuilt when it changes
    rng = np.random.default_rng(20261018)
    n = 384
    a = rng.integers(0, 4, (n, 128), dtype=np.uint8)
    b = rng.integers(0, 4, (n, 128), dtype=np.uint8)
    b[:64] = a[:64]
    b[64:128] = np.roll(a[64:128], 7, axis=1)
    a[128:160] = rng.integers(0, 2, (32, 128), dtype=np.uint8)
    b[128:160] = rng.integers(0, 2, (32, 128), dtype=np.uint8)
    for k in range(160, 224):
        keep = rng.random(128) > 0.1
        b[k] = np.concatenate([a[k][keep], rng.integers(0, 4, 128, dtype=np.uint8)])[:128]
    for trial in range(60):
        kind = trial % 4
        if kind == 0:
            sm = rng.integers(-127, 128, 16)
        elif kind == 1:
            sm = rng.integers(0, 128, 16)
        elif kind == 2:
            sm = rng.integers(-127, 1, 16)
        else:
            m, x = int(rng.integers(1, 128)), int(rng.integers(-127, 1))
            sm = np.array(mm(m, x))
        g = int(rng.integers(0, 128)) if trial % 5 else (0, 127)[trial % 2]
        sm = [int(v) for v in sm]
        want = oracle.score_batch(a, b, sm, g, threads=NCPU)
        got = ctx.score_batch(a, b, sm, g)
        assert np.array_equal(got, want), (trial, sm, g, int(np.flatnonzero(got != want)[0]))
        for i in range(trial % 7, n, 29):
            assert ctx.smith_waterman(a[i], b[i], sm, g) == int(want[i]), (trial, i, sm, g)


def test_latency_kernel_at_its_largest_batch(ctx, swb, oracle):
    # 2048 pairs = the most the one-warp-per-pair kernel takes; 2049 is the first batch of the throughput kernel
    a, b = swb.counter_pairs(424242, 2049)
    want = oracle.score_batch(a, b, swb.MATRIX_SPEEDTEST, 15, threads=NCPU)
    l0 = ctx.launch_count
    assert np.array_equal(ctx.score_batch(a[:2048], b[:2048], swb.MATRIX_SPEEDTEST, 15), want[:2048])
    assert np.array_equal(ctx.score_batch(a, b, swb.MATRIX_SPEEDTEST, 15), want)
    assert ctx.launch_count - l0 == 2
    sm = mm(127, -127)
    assert np.array_equal(ctx.score_batch(a[:2048], b[:2048], sm, 127), oracle.score_batch(a[:2048], b[:2048], sm, 127, threads=NCPU))


def test_full_1m_batch_checksum(ctx, swb, oracle, golden):
    # BASELINE.json configs[1]: the 1 M-pair batch, bit-exact vs reference scalar/simd4/simd9
    n = 1_000_000
    a, b = oracle.reference_stream(n)
    for name, sm, g in (("speedtest_10_-30_15", swb.MATRIX_SPEEDTEST, 15), ("x32_1_-1_1", swb.MATRIX_111, 1)):
        s = ctx.score_batch(a, b, sm, g)
        gold = golden["reference_stream"]["sets"][name]["1000000"]
        assert int(s.sum()) == gold["sum"] and int(s.min()) == gold["min"] and int(s.max()) == gold["max"]
        assert int(np.argmax(s)) == gold["argmax"]
        assert f"{oracle.fnv1a64(s):016x}" == gold["fnv1a64"]
    assert golden["reference_stream"]["sets"]["speedtest_10_-30_15"]["1000000"]["fnv1a64"] == "ae56a1e6a1d57492"
    # element-wise against the reference's own AVX2 kernels where the build travelled
    if oracle.have_ref():
        s = ctx.score_batch(a, b, swb.MATRIX_SPEEDTEST, 15)
        for variant in (4, 9):
            assert np.array_equal(s, oracle.ref_score_batch(variant, a, b, swb.MATRIX_SPEEDTEST, 15, threads=NCPU)), variant


def test_general_kernel_on_stream(ctx, swb, oracle):
    a, b = oracle.reference_stream(20_000)
    exp = oracle.score_batch(a, b, swb.MATRIX_SPEEDTEST, 15, threads=NCPU)
    ctx.set_force_general(True)
    try:
        assert np.array_equal(ctx.score_batch(a, b, swb.MATRIX_SPEEDTEST, 15), exp)
    finally:
        ctx.set_force_general(False)
    # out-of-fast-domain parameters take the general kernel on their own
    sm = mm(127, -127)
    assert np.array_equal(ctx.score_batch(a[:5000], b[:5000], sm, 127), oracle.score_batch(a[:5000], b[:5000], sm, 127, threads=NCPU))


@pytest.mark.parametrize("latency", [True, False])
@pytest.mark.parametrize("n", [0, 1, 2, 3, 127, 128, 255, 256, 257, 1023])
def test_ragged_batch_sizes(ctx, swb, oracle, n, latency):
    a, b = swb.counter_pairs(77, n)
    ctx.set_latency_path(latency)           # the latency kernel (default for a batch this small) and the throughput kernel
    try:
        got = ctx.score_batch(a, b, swb.MATRIX_SPEEDTEST, 15)
    finally:
        ctx.set_latency_path(True)
    assert got.shape == (n,)
    if n:
        assert np.array_equal(got, oracle.score_batch(a, b, swb.MATRIX_SPEEDTEST, 15))


@pytest.mark.parametrize("n", [8192, 8193, 12287, 12289, 4096 * 5 + 129, 65536 + 1])
@pytest.mark.parametrize("packed", [False, True])
def test_ragged_batches_through_the_persistent_kernel(ctx, swb, oracle, n, packed):
    # sizes around the tile (4096 pairs), work-item (128 pairs) and odd-pair boundaries of the consumer kernel, byte-coded and
    # 2-bit input, pageable host arrays (device score staging + one D2H) -- equal to the oracle on every pair
    a, b = swb.counter_pairs(1_234_567, n)
    want = oracle.score_batch(a, b, swb.MATRIX_SPEEDTEST, 15, threads=NCPU)
    l0 = ctx.launch_count
    if packed:
        got = ctx.score_batch(oracle.pack2bit(a), oracle.pack2bit(b), swb.MATRIX_SPEEDTEST, 15, packed=True)
    else:
        got = ctx.score_batch(a, b, swb.MATRIX_SPEEDTEST, 15)
    assert 1 <= ctx.launch_count - l0 <= 3 + n // 65536     # a few windows of the persistent kernel, however many copies fed them
    assert np.array_equal(got, want)


@pytest.mark.parametrize("n", [8192, 40_001, 300_007])
def test_pinned_2bit_input_is_pulled_by_the_kernel(ctx, swb, oracle, n):
    # 2-bit input in PINNED arrays: no copy at all, the blocks load their pairs across PCIe themselves (one launch);
    # also from a row offset into the pinned arrays (the pointers are then interior to the allocation)
    a, b = swb.counter_pairs(31_000_000, n + 5)
    ka, kb = swb.PinnedArray((n + 5, 32), np.uint8), swb.PinnedArray((n + 5, 32), np.uint8)
    ka.array[...] = oracle.pack2bit(a)
    kb.array[...] = oracle.pack2bit(b)
    out = swb.PinnedArray((n,), np.int32)
    want = oracle.score_batch(a, b, swb.MATRIX_SPEEDTEST, 15, threads=NCPU)
    for off in (0, 5):
        out.array[:] = -1
        l0 = ctx.launch_count
        ctx.score_batch(ka.array[off:off + n], kb.array[off:off + n], swb.MATRIX_SPEEDTEST, 15, out=out.array, packed=True)
        assert ctx.launch_count - l0 == 1
        assert np.array_equal(out.array, want[off:off + n]), off
    for p in (ka, kb, out):
        p.free()


def test_persistent_kernel_scores_into_an_unaligned_pinned_slice(ctx, swb, oracle):
    # pinned arrays: the kernel stores scores straight into host memory; an odd offset takes the 4-byte store form
    n = 20_000
    a, b = swb.counter_pairs(55, n)
    pa, pb, ps = swb.PinnedArray((n, 128), np.uint8), swb.PinnedArray((n, 128), np.uint8), swb.PinnedArray((n + 3,), np.int32)
    pa.array[:] = a
    pb.array[:] = b
    ps.array[:] = -1
    want = oracle.score_batch(a, b, swb.MATRIX_SPEEDTEST, 15, threads=NCPU)
    for off in (0, 1):
        ps.array[:] = -1
        ctx.score_batch(pa.array, pb.array, swb.MATRIX_SPEEDTEST, 15, out=ps.array[off:off + n])
        assert np.array_equal(ps.array[off:off + n], want) and (ps.array[off + n:] == -1).all() and (ps.array[:off] == -1).all()
    for p in (pa, pb, ps):
        p.free()


def test_multi_chunk_batch_with_pinned_buffers(ctx, swb, oracle):
    # 400 000 pairs, pinned in, pinned out: the persistent kernel stores straight into the pinned score array
    n = 3 * 131072 + 12345
    a, b = swb.counter_pairs(5_000_000, n)
    pa, pb, ps = swb.PinnedArray((n, 128), np.uint8), swb.PinnedArray((n, 128), np.uint8), swb.PinnedArray((n,), np.int32)
    pa.array[:] = a
    pb.array[:] = b
    got = ctx.score_batch(pa.array, pb.array, swb.MATRIX_SPEEDTEST, 15, out=ps.array)
    sample = np.r_[0:4096, n - 4096:n, np.arange(0, n, 997)]
    assert np.array_equal(got[sample], oracle.score_batch(a[sample], b[sample], swb.MATRIX_SPEEDTEST, 15, threads=NCPU))
    # property at full size: the batch result is a pure function of each pair (order-independent)
    perm = np.random.default_rng(5).permutation(n)[:200_000]
    again = ctx.score_batch(a[perm], b[perm], swb.MATRIX_SPEEDTEST, 15)
    assert np.array_equal(again, got[perm])
    for p in (pa, pb, ps):
        p.free()


def test_host_pack_lanes_equal_plain_pipeline(ctx, swb, oracle):
    # swb200_score_batch on a large byte-coded batch: RAW + PACK lanes (2-bit wire compression on host
    # cores) must give the very scores of the plain chunk pipeline, with any lane count.
    n = 20 * 16384 + 4321
    a, b = swb.counter_pairs(9_000_000, n)
    try:
        ctx.set_host_pack_threads(0)
        s0 = ctx.host_pack_stats()
        plain = ctx.score_batch(a, b, swb.MATRIX_SPEEDTEST, 15)
        s1 = ctx.host_pack_stats()
        assert s1["packed_pairs"] == s0["packed_pairs"] and s1["raw_pairs"] == s0["raw_pairs"] + n   # every pair travelled as bytes
        sample = np.r_[0:2048, n - 2048:n, np.arange(0, n, 1009)]
        assert np.array_equal(plain[sample], oracle.score_batch(a[sample], b[sample], swb.MATRIX_SPEEDTEST, 15, threads=NCPU))
        for t in (1, 5):
            ctx.set_host_pack_threads(t)
            s1 = ctx.host_pack_stats()
            assert np.array_equal(ctx.score_batch(a, b, swb.MATRIX_SPEEDTEST, 15), plain), t
            assert np.array_equal(ctx.score_batch(a, b, mm(127, -127), 127), ctx.score_batch(a[::-1].copy(), b[::-1].copy(), mm(127, -127), 127)[::-1])
            s2 = ctx.host_pack_stats()
            assert s2["packed_pairs"] + s2["raw_pairs"] - s1["packed_pairs"] - s1["raw_pairs"] == 3 * n
            assert s2["packed_pairs"] > s1["packed_pairs"]          # the pack lanes did take work
            assert s2["pack_threads_per_gpu"] == t
    finally:
        ctx.set_host_pack_threads(-1)


def test_lanes_epoch_makes_progress_without_the_relay_warp(ctx, swb, oracle, monkeypatch):
    # The relay warp mirrors the PACK lanes' host flags into device memory; should it find no room on an SM, every waiting
    # block looks at its tile's host flag itself now and then.  With the relay switched off that path carries the whole batch.
    n = 400_000
    a, b = swb.counter_pairs(77_000_000, n)
    want = ctx.score_batch(a, b, swb.MATRIX_SPEEDTEST, 15)
    monkeypatch.setenv("SWB200_FEED_NO_RELAY", "1")
    try:
        ctx.set_host_pack_threads(6)
        s0 = ctx.host_pack_stats()
        got = ctx.score_batch(a, b, swb.MATRIX_SPEEDTEST, 15)
        s1 = ctx.host_pack_stats()
    finally:
        ctx.set_host_pack_threads(-1)
    assert s1["packed_pairs"] > s0["packed_pairs"]
    assert np.array_equal(got, want)
    sample = np.r_[0:1024, n - 1024:n]
    assert np.array_equal(got[sample], oracle.score_batch(a[sample], b[sample], swb.MATRIX_SPEEDTEST, 15, threads=NCPU))


def test_domain_properties_at_scale(ctx, swb):
    # size-independent properties on 300 000 pairs (no oracle needed)
    n = 300_000
    a, b = swb.counter_pairs(123, n)
    s = ctx.score_batch(a, b, swb.MATRIX_SPEEDTEST, 15)
    # identical sequences: 128 matches
    assert np.all(ctx.score_batch(a, a, swb.MATRIX_SPEEDTEST, 15) == 1280)
    # symmetric matrix: swapping the roles of seq1 and seq2 transposes the table
    assert np.array_equal(ctx.score_batch(b, a, swb.MATRIX_SPEEDTEST, 15), s)
    # reversing both sequences reverses every alignment
    assert np.array_equal(ctx.score_batch(a[:, ::-1].copy(), b[:, ::-1].copy(), swb.MATRIX_SPEEDTEST, 15), s)
    # scaling matrix and gap by k scales the score by k (also crosses fast -> general kernel)
    assert np.array_equal(ctx.score_batch(a, b, mm(2, -6), 3) * 5, s)
    assert np.array_equal(ctx.score_batch(a, b, mm(40, -120), 60), s * 4)
    # relabelling the alphabet with a permutation leaves a match/mismatch score unchanged
    perm = np.array([2, 0, 3, 1], dtype=np.uint8)
    assert np.array_equal(ctx.score_batch(perm[a], perm[b], swb.MATRIX_SPEEDTEST, 15), s)
    # bounds
    assert s.min() >= 0 and s.max() <= 1280


def test_packed_2bit_input(ctx, swb, oracle):
    n = 50_001
    a, b = swb.counter_pairs(9, n)
    pa, pb = oracle.pack2bit(a), oracle.pack2bit(b)
    got = ctx.score_batch(pa, pb, swb.MATRIX_SPEEDTEST, 15, packed=True)
    assert np.array_equal(got, ctx.score_batch(a, b, swb.MATRIX_SPEEDTEST, 15))
    assert np.array_equal(got[:3000], oracle.score_batch(a[:3000], b[:3000], swb.MATRIX_SPEEDTEST, 15, threads=NCPU))


def test_device_resident_entry(ctx, swb, oracle):
    import torch
    n = 40_000
    a, b = swb.counter_pairs(31337, n)
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    ds = torch.empty(n, dtype=torch.int32, device="cuda")
    ctx.score_batch_device(da, db, swb.MATRIX_111, 1, ds)
    torch.cuda.synchronize()
    assert np.array_equal(ds.cpu().numpy(), oracle.score_batch(a, b, swb.MATRIX_111, 1, threads=NCPU))
    # packed device entry
    dpa, dpb = torch.from_numpy(oracle.pack2bit(a)).cuda(), torch.from_numpy(oracle.pack2bit(b)).cuda()
    ds2 = torch.empty(n, dtype=torch.int32, device="cuda")
    ctx.score_batch_device(dpa, dpb, swb.MATRIX_111, 1, ds2, n=n, packed=True)
    torch.cuda.synchronize()
    assert torch.equal(ds, ds2)
    assert ctx.count_bad_codes_device(da) == 0
    da[17, 5] = 9
    assert ctx.count_bad_codes_device(da) == 1


def test_submit_wait_streaming(ctx, swb, oracle):
    n = 30_000
    batches = [swb.counter_pairs(k * n, n) for k in range(3)]
    outs = [np.empty(n, np.int32) for _ in batches]
    tickets = [ctx.submit(a, b, swb.MATRIX_SPEEDTEST, 15, o) for (a, b), o in zip(batches, outs)]
    for t in tickets:
        ctx.wait(t)
    for (a, b), o in zip(batches, outs):
        assert np.array_equal(o[:2000], oracle.score_batch(a[:2000], b[:2000], swb.MATRIX_SPEEDTEST, 15))
    with pytest.raises(swb.SwbError) as e:
        ctx.wait(tickets[0])
    assert e.value.code == swb.ERR_TICKET


def test_errors_are_codes_not_aborts(ctx, swb):
    a, b = swb.counter_pairs(0, 4)
    with pytest.raises(swb.SwbError) as e:
        ctx.score_batch(a, b, mm(10, -128), 15)
    assert e.value.code == swb.ERR_DOMAIN
    with pytest.raises(swb.SwbError) as e:
        ctx.score_batch(a, b, swb.MATRIX_SPEEDTEST, -1)
    assert e.value.code == swb.ERR_DOMAIN
    with pytest.raises(ValueError):
        ctx.smith_waterman(a[0][:100], b[0], swb.MATRIX_SPEEDTEST, 15)
    # the context is still usable afterwards
    assert ctx.score_batch(a, a, swb.MATRIX_SPEEDTEST, 15).tolist() == [1280] * 4


def test_all_visible_gpus_shard_by_index_range(swb, oracle):
    import torch
    g = torch.cuda.device_count()
    n = 300_001
    a, b = swb.counter_pairs(42, n)
    with swb.Context(n_devices=g) as c:
        assert c.n_devices == g
        got = c.score_batch(a, b, swb.MATRIX_SPEEDTEST, 15)
    sample = np.r_[0:2000, n - 2000:n, np.arange(0, n, 499)]
    assert np.array_equal(got[sample], oracle.score_batch(a[sample], b[sample], swb.MATRIX_SPEEDTEST, 15, threads=NCPU))


@pytest.mark.parametrize("L", [256, 512])
def test_length_sweep_parity(ctx, swb, oracle, L):
    # BASELINE.json configs[3]: 2x / 4x the built-in shape, templated kernels, same C ABI family
    rng = np.random.default_rng(1000 + L)
    n = 3001
    a = rng.integers(0, 4, (n, L), dtype=np.uint8)
    b = rng.integers(0, 4, (n, L), dtype=np.uint8)
    for i in range(n // 2):                       # related pairs: long gapped alignments
        keep = rng.random(L) > 0.08
        b[i] = np.concatenate([a[i][keep], rng.integers(0, 4, L, dtype=np.uint8)])[:L]
    a[0] = b[0]
    for sm, g in ((swb.MATRIX_SPEEDTEST, 15), (swb.MATRIX_111, 1), (mm(40, -50), 30), (mm(60, -127), 3)):
        exp = oracle.score_batch(a, b, sm, g, threads=NCPU)
        for force_general in (False, True):
            ctx.set_force_general(force_general)
            try:
                got = ctx.score_batch(a, b, sm, g)
            finally:
                ctx.set_force_general(False)
            assert np.array_equal(got, exp), (L, sm[0], g, force_general)
    assert ctx.score_batch(a[:1], b[:1], swb.MATRIX_SPEEDTEST, 15)[0] == 10 * L
    info = ctx.kernel_info(swb.MATRIX_SPEEDTEST, 15, seq_len=L)
    assert info["fast_path"] == 1
    if info["smem_bytes_per_block"] > 4096:       # FIFO in shared memory: L words per thread must fit the SM
        assert info["threads_per_block"] * info["blocks_per_sm"] * L * 4 <= 227 * 1024
    else:
        # FIFO in global memory / L2 (L = 512): a persistent grid.  Enough pairs that every resident thread
        # walks more than one work item, checked against the oracle on both sides of the first pass.
        resident_pairs = 2 * info["threads_per_block"] * info["blocks_per_sm"] * info["sm_count"]
        n2 = resident_pairs + 9001
        a2 = rng.integers(0, 4, (n2, L), dtype=np.uint8)
        b2 = np.where(rng.random((n2, L)) < 0.9, a2, rng.integers(0, 4, (n2, L), dtype=np.uint8)).astype(np.uint8)
        got2 = ctx.score_batch(a2, b2, swb.MATRIX_SPEEDTEST, 15)
        sample = np.r_[0:1500, resident_pairs - 1500:resident_pairs + 1500, n2 - 1500:n2]
        assert np.array_equal(got2[sample], oracle.score_batch(a2[sample], b2[sample], swb.MATRIX_SPEEDTEST, 15, threads=NCPU))
        assert np.array_equal(ctx.score_batch(b2, a2, swb.MATRIX_SPEEDTEST, 15), got2)      # symmetric matrix: transpose invariance, all pairs
    if L == 512:
        with pytest.raises(swb.SwbError) as e:    # 512 * 127 overflows packed int16: refused, not wrong
            ctx.score_batch(a[:4], b[:4], mm(127, -127), 127)
        assert e.value.code == swb.ERR_DOMAIN
    # device-resident entry at this length
    import torch
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    ds = torch.empty(n, dtype=torch.int32, device="cuda")
    ctx.score_batch_device(da, db, swb.MATRIX_SPEEDTEST, 15, ds)
    torch.cuda.synchronize()
    assert np.array_equal(ds.cpu().numpy(), oracle.score_batch(a, b, swb.MATRIX_SPEEDTEST, 15, threads=NCPU))


def test_fixed_111_entry(ctx, swb, oracle, golden):
    # swb200_score_batch_111 stands in for SmithWaterman_111 / SmithWaterman_8bit111simd
    n = 100_000
    a, b = oracle.reference_stream(n)
    s = ctx.score_batch_111(a, b)
    gold = golden["reference_stream"]["sets"]["x32_1_-1_1"]
    assert list(s[:16]) == gold["first16"]
    assert int(s.sum()) == gold["100000"]["sum"] and f"{oracle.fnv1a64(s):016x}" == gold["100000"]["fnv1a64"]
    assert np.array_equal(s, ctx.score_batch(a, b, swb.MATRIX_111, 1))
    z = golden["structured_npz"]                       # identical / mutated / indel / homopolymer pairs
    got = ctx.score_batch_111(z["seq1"], z["seq2"])
    assert np.array_equal(got, oracle.score_batch(z["seq1"], z["seq2"], oracle.MATRIX_111, 1))
    if oracle.have_ref():
        for i in range(0, z["seq1"].shape[0], 7):
            assert got[i] == oracle.ref_111(z["seq1"][i], z["seq2"][i]) == oracle.ref_8bit111(z["seq1"][i], z["seq2"][i])
    assert swb.SmithWaterman_111_b200(a[0], b[0]) == gold["first16"][0]


def test_one_vs_many_equals_reference_x32(ctx, swb, oracle):
    # many queries vs one target, the shape of SmithWaterman_8b111x32mark1
    # (source.cpp:1227-1234), on the inputs of TestSimdSmithWaterman111x32 (source.cpp:3004-3013)
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "x32_first40.npz"))
    for it in range(40):
        got = ctx.score_one_vs_many(z["queries"][it], z["targets"][it])          # defaults: +1/-1, gap 1
        assert np.array_equal(got, z["scores"][it].astype(np.int32)), it
        if oracle.have_ref():
            for mark in (1, 2, 3):
                assert np.array_equal(got, oracle.ref_x32(mark, z["queries"][it], z["targets"][it]))
    # any n, any matrix of the domain; equals the pairwise entry with the target repeated
    n = 200_001
    q, _ = swb.counter_pairs(7, n)
    t = z["targets"][3]
    got = ctx.score_one_vs_many(q, t, swb.MATRIX_SPEEDTEST, 15)
    assert np.array_equal(got, ctx.score_batch(q, np.repeat(t[None, :], n, axis=0), swb.MATRIX_SPEEDTEST, 15))
    assert np.array_equal(got[:3000], oracle.score_batch(q[:3000], np.repeat(t[None, :], 3000, axis=0), swb.MATRIX_SPEEDTEST, 15, threads=NCPU))


def test_random_matrices_and_adversarial_sequences(ctx, oracle):
    # the generator of tools/sw_fuzz.py: score matrices and gaps drawn over the whole reference domain (both kernels get
    # picked), low-complexity / periodic / shifted sequences; every score compared with the oracle
    import importlib.util
    spec = importlib.util.spec_from_file_location("sw_fuzz", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "sw_fuzz.py"))
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    rng = np.random.default_rng(20261018)
    fast = general = 0
    for _ in range(250):
        a, b, m, gap = fz.make_batch(rng, n=66)
        got = ctx.score_batch(a, b, m, gap)
        assert np.array_equal(got, oracle.score_batch(a, b, m, gap)), (m.tolist(), gap)
        if ctx.kernel_info(m, gap)["fast_path"]:
            fast += 1
        else:
            general += 1
    assert fast > 20 and general > 20
