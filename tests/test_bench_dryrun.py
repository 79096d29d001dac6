"""Dry run of bench.py's B200 arm on the CPU: the GPU, the library context and pinned memory are replaced by
stand-ins (scores come from the oracle -- this is a test, the oracle is the checker), so every line of the
measurement / JSON-assembly code runs in the GPU-less container and a Python error cannot hide until the real run.
Nothing here measures anything."""
import contextlib
import ctypes as C
import importlib.util
import io
import json
import os
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Event:
    def __init__(self, enable_timing=False):
        pass

    def record(self, stream=None):
        pass

    def elapsed_time(self, other):
        return 2.0


class _Stream:
    def synchronize(self):
        pass


class _FakePinned:
    def __init__(self, shape, dtype):
        self.array = np.zeros(shape, dtype)
        self.nbytes = self.array.nbytes

    def free(self):
        pass


def _fake_context_class(oracle, swb200):
    class FakeContext:
        def __init__(self, devices=None, n_devices=1):
            self.launch_count = 0
            self._pack = 5
            self._packed_pairs = 0

        def kernel_info(self, m, g, device_index=0, seq_len=128):
            return {"fast_path": 1, "regs_per_thread": 168, "threads_per_block": 64, "blocks_per_sm": 6,
                    "smem_bytes_per_block": 32784, "sm_count": 148, "sm_clock_khz": 1965000}

        def score_batch_device(self, d_a, d_b, m, g, d_s, **kw):
            import torch
            d_s.copy_(torch.from_numpy(oracle.score_batch(d_a.numpy(), d_b.numpy(), m, g, threads=os.cpu_count() or 1)))
            self.launch_count += 1
            return d_s

        def score_batch(self, a, b, m, g, out=None, packed=False):
            if packed:
                a, b = oracle.unpack2bit(a), oracle.unpack2bit(b)
            else:
                self._packed_pairs += a.shape[0] // 2 if self._pack else 0
            out[:a.shape[0]] = oracle.score_batch(a, b, m, g, threads=os.cpu_count() or 1)
            self.launch_count += 3
            return out[:a.shape[0]]

        def submit(self, a, b, m, g, out, packed=False):
            self.score_batch(a, b, m, g, out=out, packed=packed)
            return 1

        def wait(self, ticket):
            pass

        # ---- semi-global aligner stand-ins (oracle per pair)
        def semiglobal_kernel_info(self, device_index=0):
            return {"fast_path": 0, "regs_per_thread": 142, "threads_per_block": 32, "blocks_per_sm": 8,
                    "smem_bytes_per_block": 0, "sm_count": 148, "sm_clock_khz": 1965000}

        def _sg(self, a, b):
            res = [oracle.semiglobal_xdrop(a[i], b[i]) for i in range(a.shape[0])]
            return res

        def semiglobal_xdrop_device(self, d_a, d_b, d_score, d_ey, d_ex, d_nops=None, d_ops=None, **kw):
            for i, (sc, ey, ex, ops) in enumerate(self._sg(d_a.numpy(), d_b.numpy())):
                d_score[i], d_ey[i], d_ex[i] = sc, ey, ex
                if d_ops is not None:
                    d_nops[i] = ops.size
                    d_ops[i, :ops.size] = __import__("torch").from_numpy(ops)
            self.launch_count += 1 if d_ops is None else 3

        def _check(self, rc):
            assert rc == 0

        @property
        def _h(self):
            return None

        @property
        def _lib(self):
            ctx = self

            class Lib:
                @staticmethod
                def swb200_score_pair(h, p1, p2, pm, gap, pout):
                    a = np.frombuffer((C.c_uint8 * 128).from_address(p1), dtype=np.uint8)[None, :]
                    b = np.frombuffer((C.c_uint8 * 128).from_address(p2), dtype=np.uint8)[None, :]
                    m = np.frombuffer((C.c_int8 * 16).from_address(pm), dtype=np.int8)
                    np.frombuffer((C.c_int32 * 1).from_address(pout), dtype=np.int32)[0] = oracle.score_batch(a, b, m, gap.value)[0]
                    ctx.launch_count += 1
                    return 0

                @staticmethod
                def swb200_semiglobal_xdrop_batch(h, pa, pb, length, n, p_score, p_ey, p_ex, p_nops, p_ops):
                    def view(ptr, count, ctype, dtype):
                        return np.frombuffer((ctype * count).from_address(ptr), dtype=dtype)
                    a = view(pa, n * length, C.c_uint8, np.uint8).reshape(n, length)
                    b = view(pb, n * length, C.c_uint8, np.uint8).reshape(n, length)
                    score, ey, ex, nops = (view(p, n, C.c_int32, np.int32) for p in (p_score, p_ey, p_ex, p_nops))
                    ops = view(p_ops, n * 2 * length, C.c_uint8, np.uint8).reshape(n, 2 * length)
                    for i, (sc, y, x, o) in enumerate(ctx._sg(a, b)):
                        score[i], ey[i], ex[i], nops[i] = sc, y, x, o.size
                        ops[i, :o.size] = o
                    ctx.launch_count += 3
                    return 0
            return Lib

        def set_host_pack_threads(self, t):
            self._pack = 5 if t < 0 else t

        def host_pack_tuning(self, device_index=0):
            return {"lanes_in_use": self._pack, "pairs_per_s": {"all_lanes": 1.0, "half": 0.5, "none": 0.25}}

        def measure_alu_peak(self, device_index=0, target_ms=50.0):
            return {"tinstr_per_s": 18.4, "elapsed_ms": target_ms}

        def host_pack_stats(self):
            return {"packed_pairs": self._packed_pairs, "raw_pairs": 0, "pack_threads_per_gpu": self._pack}

        def close(self):
            pass
    return FakeContext


@pytest.fixture
def bench(monkeypatch, oracle):
    import torch
    sys.path.insert(0, os.path.join(ROOT, "smith-waterman-simd_b200"))
    import swb200
    spec = importlib.util.spec_from_file_location("bench_dry", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: types.SimpleNamespace(cuda_stream=0))
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "Stream", _Stream)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    real_empty = torch.empty
    monkeypatch.setattr(torch, "empty", lambda *a, **k: real_empty(*a, **{kk: v for kk, v in k.items() if kk != "device"}))
    monkeypatch.setattr(swb200, "Context", _fake_context_class(oracle, swb200))
    monkeypatch.setattr(swb200, "PinnedArray", _FakePinned)
    monkeypatch.setattr(mod, "PAIRS_PER_GPU", 3000)
    monkeypatch.setattr(mod, "HOST_PROBE_MB", 8)
    monkeypatch.setattr(mod, "SG_LEG_PAIRS", 6)
    monkeypatch.setattr(mod, "SG_LEG_LEN", 256)
    mod.real_sweep_pairs = mod.sweep_pairs
    monkeypatch.setattr(mod, "sweep_pairs", lambda L, info: 64)       # the default line carries the sweep: keep the oracle's share small
    for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    return mod


def test_b200_arm_assembles_its_line(bench):
    args = types.SimpleNamespace(gpus=1, steps=2, warmup=3, impl="b200", no_cpu_baseline=False, pack_threads=None,
                                 no_plain_e2e=False, cpu_table=False, workload="batch1m", pairs=0, batch_pairs=0, packed=False,
                                 quick=False, stream_pairs=5000)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        bench.run_b200_arm(args)
    line = json.loads(buf.getvalue().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "verified"):
        assert k in line, k
    assert line["gpu_launches"] == 2 and line["e2e"]["scores_equal_device_leg"] is True
    assert line["e2e"]["packed_input"]["scores_equal_device_leg"] is True and line["e2e"]["packed_input"]["h2d_bytes_per_step"] == 2 * 3000 * 32
    assert line["roofline"]["bound"] == "int_alu" and line["roofline"]["frac"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["verified"]["fnv1a64_ae56a1e6a1d57492_and_sum_75478815"] is False      # 3000 pairs, not the 1 M batch
    # the legs the default line carries since round 2
    assert line["roofline"]["peak_live"]["tinstr_per_s"] == 18.4 and line["roofline"]["peak"] == 18.4 and 0 < line["roofline"]["alu_pipe_busy"]["model"]
    assert [r["seq_len"] for r in line["sweep"]] == [128, 256, 512] and all(r["pairs"] == 64 for r in line["sweep"])
    assert line["per_pair"]["score"] == 80 and line["per_pair"]["us_per_call"] > 0 and line["per_pair"]["gpu_launches_per_call"] == 1
    sg = line["semiglobal"]                                          # SURVEY.md 8(f4) in the default line (shrunk to 6 pairs of 256 here)
    assert "error" not in sg, sg
    assert sg["pairs"] == 6 and sg["device_resident"]["alignments_per_s"] > 0 and sg["e2e"]["alignments_per_s"] > 0
    assert sg["verified"]["e2e_scores_and_lengths_equal_device"] is True and sg["verified"]["whole_batch_sums_equal_oracle"] is None
    assert line["stream"]["packed"]["pairs"] == 5000 and line["stream"]["bytes"]["score_sum"] == line["stream"]["packed"]["score_sum"]
    hc = line["host_ceiling"]
    assert "error" not in hc, hc
    assert hc["h2d_pinned_gbs"] > 0 and hc["host_read_gbs"] > 0 and hc["byte_input_ceiling_gcups"] > 0
    assert line["e2e_inproc"] is None and line["e2e"]["packed_input"]["gpu_launches_per_step"] > 0
    assert line["verified"]["other_ranks_score_sums_equal_reference"] is None           # one rank


def _args(**kw):
    base = dict(gpus=1, steps=2, warmup=3, impl="b200", no_cpu_baseline=True, pack_threads=None, no_plain_e2e=False,
                cpu_table=False, workload="batch1m", pairs=0, batch_pairs=0, packed=False, quick=False, stream_pairs=4000)
    base.update(kw)
    return types.SimpleNamespace(**base)


def _run(fn, args):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        fn(args)
    return json.loads(buf.getvalue().strip().splitlines()[-1])


def test_sweep_arm_assembles_its_line(bench, monkeypatch):
    info = {"sm_count": 148, "blocks_per_sm": 6, "threads_per_block": 64}
    assert [bench.real_sweep_pairs(L, dict(info, blocks_per_sm=b)) for L, b in ((128, 6), (256, 3), (512, 6))] == [1136640, 284160, 340992]
    with open(os.path.join(ROOT, "tests", "golden", "sweep_sums.json")) as f:
        g = json.load(f)["by_length"]
    assert all(str(n) in g[L]["sum_of_scores_by_pairs"] for L, n in (("128", 1136640), ("256", 284160), ("512", 340992)))   # the B200's batches are pinned
    monkeypatch.setattr(bench, "sweep_pairs", lambda L, info: 96)
    line = _run(bench.run_sweep_arm, _args(workload="sweep"))
    assert [r["seq_len"] for r in line["sweep"]] == [128, 256, 512]
    for r in line["sweep"]:
        assert r["pairs"] == 96 and r["score_sum"] > 0 and r["score_sum_equals_oracle"] is None   # not the golden's batch size


@pytest.mark.parametrize("packed", [False, True])
def test_stream_arm_assembles_its_line(bench, packed):
    line = _run(bench.run_stream_arm, _args(workload="stream", pairs=5000, batch_pairs=2000, packed=packed))
    assert line["unit"] == "alignments/s" and line["scaling"] == "strong" and line["config"]["pairs"] == 5000
    assert line["verified"]["score_sum_equals_reference"] is None            # 5000 pairs: no golden entry
    assert 70.0 < line["mean_score"] < 81.0 and line["e2e"]["h2d_bytes_per_step"] == 2 * 5000 * (32 if packed else 128)


def test_semiglobal_arm_assembles_its_line(bench, monkeypatch):
    monkeypatch.setattr(bench, "SG_LEN", 512)              # short sequences: the oracle aligns them in milliseconds
    monkeypatch.setattr(bench, "SG_ROUNDS_NOMINAL", 1024)
    line = _run(bench.run_semiglobal_arm, _args(workload="semiglobal", pairs=12, no_cpu_baseline=True))
    assert line["verified"]["sample_equals_oracle_score_and_traceback"] is None and "cpu_baseline" not in line   # nothing under oracle/ ran
    monkeypatch.setattr(bench, "sg_cpu_reference", lambda a, b, budget_s=20.0: None)      # the reference build aligns 16384-mers only
    line = _run(bench.run_semiglobal_arm, _args(workload="semiglobal", pairs=12, no_cpu_baseline=False))
    assert line["unit"] == "alignments/s" and line["config"]["pairs"] == 12
    v = line["verified"]
    assert v["sample_equals_oracle_score_and_traceback"] is True and v["e2e_equals_device"] is True
    assert v["whole_batch_sums_equal_oracle"] is None       # 12 pairs: no golden entry


def _two_rank_worker(rank, world, port, out_path):
    """One rank of a 2-process dry run: the same stand-ins as the `bench` fixture, applied by hand (no pytest here),
    and the process group on gloo instead of nccl."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "smith-waterman-simd_b200"))
    import swb200
    from oracle import oracle as O
    os.environ.update(WORLD_SIZE=str(world), RANK=str(rank), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    spec = importlib.util.spec_from_file_location("bench_dry2", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda d: None
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.current_stream = lambda *a, **k: types.SimpleNamespace(cuda_stream=0)
    torch.cuda.Event = _Event
    torch.cuda.Stream = _Stream
    torch.cuda.stream = lambda s: contextlib.nullcontext()
    torch.Tensor.cuda = lambda self, *a, **k: self
    real_empty = torch.empty
    torch.empty = lambda *a, **k: real_empty(*a, **{kk: v for kk, v in k.items() if kk != "device"})
    real_init = dist.init_process_group
    dist.init_process_group = lambda backend, **k: real_init("gloo", rank=rank, world_size=world)
    swb200.Context = _fake_context_class(O, swb200)
    swb200.PinnedArray = _FakePinned
    swb200.bind_to_gpu_numa_node = lambda *a, **k: None
    swb200.bind_rank_cpus = lambda lr, lw, device_index=None: {"before": sorted(os.sched_getaffinity(0)), "numa": None, "cpus": 2, "share": [lr, lr]}
    mod.PAIRS_PER_GPU = 1500
    mod.HOST_PROBE_MB = 8
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        mod.run_b200_arm(_args())
    if rank == 0:
        with open(out_path, "w") as f:
            f.write(buf.getvalue())
    else:
        assert buf.getvalue().strip() == ""          # only rank 0 prints


def test_b200_arm_two_ranks_on_gloo(tmp_path, oracle):
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out_path = str(tmp_path / "line.json")
    mp.spawn(_two_rank_worker, args=(2, port, out_path), nprocs=2, join=True)
    line = json.loads(open(out_path).read().strip().splitlines()[-1])
    assert line["n_gpus"] == 2 and line["config"]["pairs_per_gpu"] == 1500
    assert line["value"] == pytest.approx(2 * 1500 * 16384 / (line["ms_per_step"] * 1e-3) / 1e9)
    assert line["verified"]["e2e_equals_device"] is True
    assert line["verified"]["other_ranks_checked"] == 0 and line["verified"]["other_ranks_score_sums_equal_reference"] is None   # 1500-pair blocks have no golden
    assert "cpu_baseline" not in line and line["sweep"] is None and line["per_pair"] is None and line["semiglobal"] is None     # N = 1 only
    assert line["e2e"]["packed_input"]["scores_equal_device_leg"] is True                           # at every N
    inproc = line["e2e_inproc"]                                                                      # rank 0 alone drives both "GPUs"
    assert inproc["n_gpus"] == 2 and inproc["pairs_per_step"] == 3000 and inproc["bytes"]["value"] > 0 and inproc["packed"]["value"] > 0
    assert line["stream"]["packed"]["pairs"] == 4000 and line["stream"]["bytes"]["pairs"] == 4000   # the index space is sharded, not replicated
    assert line["host_ceiling"]["h2d_pinned_gbs"] > 0
