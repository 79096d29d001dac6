"""Python host side of swb200: a ctypes binding of include/swb200.h.

It mirrors the reference's interface for the hot path -- the per-pair call
`SmithWaterman_simdN(seq1, seq2, score_matrix, gap_penalty) -> int`
(/root/reference/source.cpp:462-466) and the batch loop its harnesses run
(source.cpp:2947-2970) -- on top of the C ABI.  torch is used for device memory and
streams only.  Nothing here computes a score on the CPU: if libswb200.so or a B200 is
missing, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

# Streams are multiplexed onto this many in-order hardware queues (default 8).  The batch packer uses a stream per PACK
# lane plus kernel / copy streams per staging region; with more queues fewer of them share one, so copies and the
# consumer kernel overlap as intended.  (Correctness never depends on it -- see csrc/sw_feed_kernel.cuh -- and the
# variable only takes effect if it is set before the process creates its CUDA context.)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libswb200.so")
SEQ_LEN = 128
SWEEP_LENGTHS = (128, 256, 512)   # BASELINE.json configs[3]: 1x, 2x, 4x the built-in shape

# reference harness constants
MATRIX_SPEEDTEST = (10, -30, -30, -30, -30, 10, -30, -30, -30, -30, 10, -30, -30, -30, -30, 10)  # source.cpp:3041-3045
GAP_SPEEDTEST = 15                                                                           # source.cpp:3046
MATRIX_111 = (1, -1, -1, -1, -1, 1, -1, -1, -1, -1, 1, -1, -1, -1, -1, 1)                    # source.cpp:3202-3206
GAP_111 = 1                                                                                  # source.cpp:3207

# every symbol include/swb200.h declares (tests check the .so exports each one)
ABI_SYMBOLS = (
    "swb200_device_count", "swb200_init", "swb200_shutdown", "swb200_n_devices", "swb200_last_error",
    "swb200_strerror", "swb200_alloc_pinned", "swb200_free_pinned", "swb200_score_pair",
    "swb200_score_batch", "swb200_score_batch_packed", "swb200_submit", "swb200_submit_packed", "swb200_wait",
    "swb200_score_batch_device", "swb200_score_batch_packed_device", "swb200_validate_codes_device",
    "swb200_kernel_info_for", "swb200_launch_count", "swb200_set_force_general",
    "swb200_gen_reference_stream", "swb200_gen_counter_pairs", "swb200_gen_counter_pairs_packed",
    "swb200_fnv1a64_i32", "swb200_score_batch_len", "swb200_score_batch_len_device", "swb200_kernel_info_len", "swb200_score_one_vs_many",
    "swb200_score_batch_111", "swb200_semiglobal_xdrop_batch", "swb200_semiglobal_xdrop_batch_device",
    "swb200_semiglobal_kernel_info", "swb200_gen_related_pairs", "swb200_set_host_pack_threads", "swb200_host_pack_stats", "swb200_pack2bit_host", "swb200_set_latency_path", "swb200_host_read_bandwidth", "swb200_measure_alu_peak", "swb200_host_pack_tuning", "swb200_pair_path_stats", "swb200_unpack2bit_host",
)

ERR_ARG, ERR_DOMAIN, ERR_NO_DEVICE, ERR_CUDA, ERR_NOMEM, ERR_TICKET = -1, -2, -3, -4, -5, -6


class SwbError(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__(f"swb200 error {code}: {text}")
        self.code = code


class KernelInfo(C.Structure):
    _fields_ = [("fast_path", C.c_int), ("regs_per_thread", C.c_int), ("threads_per_block", C.c_int),
                ("blocks_per_sm", C.c_int), ("smem_bytes_per_block", C.c_int), ("sm_count", C.c_int),
                ("sm_clock_khz", C.c_int)]


_lib = None


def load_library():
    """dlopen libswb200.so (no CUDA call is made by loading)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} is missing: build it with `python smith-waterman-simd_b200/build.py` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, u64, i32 = C.c_void_p, C.c_uint64, C.c_int
    lib.swb200_device_count.restype = i32
    lib.swb200_init.restype = i32
    lib.swb200_init.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), i32]
    lib.swb200_shutdown.restype = None
    lib.swb200_shutdown.argtypes = [vp]
    lib.swb200_n_devices.restype = i32
    lib.swb200_n_devices.argtypes = [vp]
    lib.swb200_last_error.restype = C.c_char_p
    lib.swb200_last_error.argtypes = [vp]
    lib.swb200_strerror.restype = C.c_char_p
    lib.swb200_strerror.argtypes = [i32]
    lib.swb200_alloc_pinned.restype = i32
    lib.swb200_alloc_pinned.argtypes = [C.POINTER(vp), C.c_size_t]
    lib.swb200_free_pinned.restype = i32
    lib.swb200_free_pinned.argtypes = [vp]
    lib.swb200_score_pair.restype = i32
    lib.swb200_score_pair.argtypes = [vp, vp, vp, vp, C.c_int8, vp]
    for name in ("swb200_score_batch", "swb200_score_batch_packed"):
        f = getattr(lib, name)
        f.restype = i32
        f.argtypes = [vp, vp, vp, vp, C.c_int8, vp, u64]
    for name in ("swb200_submit", "swb200_submit_packed"):
        f = getattr(lib, name)
        f.restype = i32
        f.argtypes = [vp, vp, vp, vp, C.c_int8, vp, u64, C.POINTER(u64)]
    lib.swb200_wait.restype = i32
    lib.swb200_wait.argtypes = [vp, u64]
    for name in ("swb200_score_batch_device", "swb200_score_batch_packed_device"):
        f = getattr(lib, name)
        f.restype = i32
        f.argtypes = [vp, i32, vp, vp, vp, C.c_int8, vp, u64, vp]
    lib.swb200_score_batch_len.restype = i32
    lib.swb200_score_batch_len.argtypes = [vp, i32, vp, vp, vp, C.c_int8, vp, u64]
    lib.swb200_score_batch_len_device.restype = i32
    lib.swb200_score_batch_len_device.argtypes = [vp, i32, i32, vp, vp, vp, C.c_int8, vp, u64, vp]
    lib.swb200_kernel_info_len.restype = i32
    lib.swb200_kernel_info_len.argtypes = [vp, i32, i32, vp, C.c_int8, C.POINTER(KernelInfo)]
    lib.swb200_score_one_vs_many.restype = i32
    lib.swb200_score_one_vs_many.argtypes = [vp, vp, vp, vp, C.c_int8, vp, u64]
    lib.swb200_validate_codes_device.restype = i32
    lib.swb200_validate_codes_device.argtypes = [vp, i32, vp, u64, C.POINTER(u64), vp]
    lib.swb200_kernel_info_for.restype = i32
    lib.swb200_kernel_info_for.argtypes = [vp, i32, vp, C.c_int8, C.POINTER(KernelInfo)]
    lib.swb200_launch_count.restype = u64
    lib.swb200_launch_count.argtypes = [vp]
    lib.swb200_set_force_general.restype = i32
    lib.swb200_set_force_general.argtypes = [vp, i32]
    lib.swb200_measure_alu_peak.restype = i32
    lib.swb200_measure_alu_peak.argtypes = [vp, i32, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.swb200_host_read_bandwidth.restype = i32
    lib.swb200_host_read_bandwidth.argtypes = [vp, u64, i32, i32, C.POINTER(C.c_double)]
    lib.swb200_set_latency_path.restype = i32
    lib.swb200_set_latency_path.argtypes = [vp, i32]
    lib.swb200_gen_reference_stream.restype = i32
    lib.swb200_gen_reference_stream.argtypes = [u64, u64, vp, vp]
    for name in ("swb200_gen_counter_pairs", "swb200_gen_counter_pairs_packed"):
        f = getattr(lib, name)
        f.restype = i32
        f.argtypes = [u64, u64, u64, vp, vp, i32]
    lib.swb200_gen_related_pairs.restype = i32
    lib.swb200_gen_related_pairs.argtypes = [u64, u64, u64, i32, i32, i32, i32, vp, vp, i32]
    lib.swb200_fnv1a64_i32.restype = u64
    lib.swb200_fnv1a64_i32.argtypes = [vp, u64]
    lib.swb200_score_batch_111.restype = i32
    lib.swb200_score_batch_111.argtypes = [vp, vp, vp, vp, u64]
    lib.swb200_semiglobal_xdrop_batch.restype = i32
    lib.swb200_semiglobal_xdrop_batch.argtypes = [vp, vp, vp, i32, u64, vp, vp, vp, vp, vp]
    lib.swb200_semiglobal_xdrop_batch_device.restype = i32
    lib.swb200_semiglobal_xdrop_batch_device.argtypes = [vp, i32, vp, vp, i32, u64, vp, vp, vp, vp, vp, vp]
    lib.swb200_semiglobal_kernel_info.restype = i32
    lib.swb200_semiglobal_kernel_info.argtypes = [vp, i32, C.POINTER(KernelInfo)]
    lib.swb200_set_host_pack_threads.restype = i32
    lib.swb200_set_host_pack_threads.argtypes = [vp, i32]
    lib.swb200_host_pack_stats.restype = i32
    lib.swb200_host_pack_stats.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(i32)]
    lib.swb200_pair_path_stats.restype = i32
    lib.swb200_pair_path_stats.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
    lib.swb200_host_pack_tuning.restype = i32
    lib.swb200_host_pack_tuning.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(C.c_double)]
    lib.swb200_unpack2bit_host.restype = i32
    lib.swb200_unpack2bit_host.argtypes = [vp, vp, u64]
    lib.swb200_pack2bit_host.restype = i32
    lib.swb200_pack2bit_host.argtypes = [vp, vp, u64]
    _lib = lib
    return lib


def _matrix(score_matrix) -> np.ndarray:
    m = np.ascontiguousarray(np.asarray(score_matrix).reshape(-1))
    if m.size != 16:
        raise ValueError("score_matrix must have 16 entries (4x4, index seq1*4+seq2)")
    if m.min() < -128 or m.max() > 127:
        raise ValueError("score_matrix entries must fit int8")
    return m.astype(np.int8)


def _gap(gap_penalty) -> int:
    g = int(gap_penalty)
    if g < -128 or g > 127:
        raise SwbError(ERR_DOMAIN, "gap_penalty does not fit the reference's int8 argument")
    return g


class PinnedArray:
    """A numpy view of page-locked host memory from swb200_alloc_pinned."""

    def __init__(self, shape, dtype):
        lib = load_library()
        self._lib = lib
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        rc = lib.swb200_alloc_pinned(C.byref(p), self.nbytes)
        if rc != 0:
            raise SwbError(rc, lib.swb200_last_error(None).decode())
        self._ptr = p
        buf = (C.c_uint8 * max(self.nbytes, 1)).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self._ptr is not None:
            self.array = None
            self._lib.swb200_free_pinned(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """Owns the library handle: streams, staging buffers and the per-GPU workers."""

    def __init__(self, devices: Optional[Sequence[int]] = None, n_devices: int = 1):
        self._lib = load_library()
        self._h = C.c_void_p()
        if devices is not None:
            arr = (C.c_int * len(devices))(*devices)
            rc = self._lib.swb200_init(C.byref(self._h), arr, len(devices))
        else:
            rc = self._lib.swb200_init(C.byref(self._h), None, n_devices)
        if rc != 0:
            raise SwbError(rc, self._lib.swb200_last_error(None).decode() or self._lib.swb200_strerror(rc).decode())

    # -- plumbing
    def close(self):
        if self._h:
            self._lib.swb200_shutdown(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise SwbError(rc, self._lib.swb200_last_error(self._h).decode() or self._lib.swb200_strerror(rc).decode())

    @property
    def n_devices(self) -> int:
        return self._lib.swb200_n_devices(self._h)

    @property
    def launch_count(self) -> int:
        return int(self._lib.swb200_launch_count(self._h))

    def set_host_pack_threads(self, threads_per_gpu: int):
        """Host cores per GPU that compress byte-coded batches to 2 bits before PCIe (-1 auto, 0 off)."""
        self._check(self._lib.swb200_set_host_pack_threads(self._h, int(threads_per_gpu)))

    def host_pack_stats(self) -> dict:
        p, r, t = C.c_uint64(), C.c_uint64(), C.c_int()
        self._check(self._lib.swb200_host_pack_stats(self._h, C.byref(p), C.byref(r), C.byref(t)))
        return {"packed_pairs": int(p.value), "raw_pairs": int(r.value), "pack_threads_per_gpu": int(t.value)}

    def host_pack_tuning(self, device_index: int = 0) -> dict:
        """The auto-tuner of the PACK-lane count: lanes it currently prefers, and the pairs/s it saw with all / half / none."""
        lanes, raw = C.c_int(), C.c_int()
        rates = (C.c_double * 4)()
        self._check(self._lib.swb200_host_pack_tuning(self._h, device_index, C.byref(lanes), C.byref(raw), rates))
        return {"lanes_in_use": int(lanes.value), "raw_lane_in_use": bool(raw.value),
                "pairs_per_s": {"all_lanes_and_raw_lane": rates[0], "all_lanes_no_dma": rates[1], "half_the_lanes_and_raw_lane": rates[2], "raw_lane_only": rates[3]}}

    def set_force_general(self, on: bool):
        self._check(self._lib.swb200_set_force_general(self._h, int(on)))

    def measure_alu_peak(self, device_index: int = 0, target_ms: float = 50.0) -> dict:
        """Integer-ALU issue peak of the GPU, measured now (VIADDMNMX.S16x2 chains on every SM): Tinstr/s."""
        t, ms = C.c_double(), C.c_double()
        self._check(self._lib.swb200_measure_alu_peak(self._h, device_index, float(target_ms), C.byref(t), C.byref(ms)))
        return {"tinstr_per_s": float(t.value), "elapsed_ms": float(ms.value)}

    def pair_path_stats(self) -> dict:
        """The per-pair call's resident server: kernels launched, calls through its doorbell, the last call's sweep time."""
        n, c, ns = C.c_uint64(), C.c_uint64(), C.c_uint32()
        self._check(self._lib.swb200_pair_path_stats(self._h, C.byref(n), C.byref(c), C.byref(ns)))
        return {"server_launches": int(n.value), "doorbell_calls": int(c.value), "last_sweep_us": ns.value * 1e-3}

    def set_latency_path(self, on):
        """False: small host batches go through the throughput kernel instead of the one-warp-per-pair kernels; 2: those
        kernels with one launch per call even for a single pair (no resident server); True: the default (test hook)."""
        self._check(self._lib.swb200_set_latency_path(self._h, int(on)))

    def kernel_info(self, score_matrix, gap_penalty, device_index: int = 0, seq_len: int = SEQ_LEN) -> dict:
        m = _matrix(score_matrix)
        info = KernelInfo()
        self._check(self._lib.swb200_kernel_info_len(self._h, device_index, seq_len, m.ctypes.data, _gap(gap_penalty), C.byref(info)))
        return {k: getattr(info, k) for k, _ in KernelInfo._fields_}

    # -- the reference's per-pair call (source.cpp:462-466)
    def smith_waterman(self, seq1, seq2, score_matrix, gap_penalty) -> int:
        a = np.ascontiguousarray(seq1, dtype=np.uint8).reshape(-1)
        b = np.ascontiguousarray(seq2, dtype=np.uint8).reshape(-1)
        if a.size != SEQ_LEN or b.size != SEQ_LEN:
            raise ValueError("sequences must hold exactly 128 codes (std::array<uint8_t,128>)")
        m = _matrix(score_matrix)
        out = np.zeros(1, dtype=np.int32)
        self._check(self._lib.swb200_score_pair(self._h, a.ctypes.data, b.ctypes.data, m.ctypes.data, _gap(gap_penalty), out.ctypes.data))
        return int(out[0])

    # -- the batch loop (source.cpp:2947-2970), host arrays
    def score_batch(self, seq1: np.ndarray, seq2: np.ndarray, score_matrix, gap_penalty,
                    out: Optional[np.ndarray] = None, packed: bool = False) -> np.ndarray:
        a = np.ascontiguousarray(seq1, dtype=np.uint8)
        b = np.ascontiguousarray(seq2, dtype=np.uint8)
        if a.ndim != 2 or a.shape != b.shape or (packed and a.shape[1] != 32) or (not packed and a.shape[1] not in SWEEP_LENGTHS):
            raise ValueError("seq1 and seq2 must both be uint8 [n][128] ([n][256] / [n][512] for the length sweep; [n][32] packed)")
        n, width = a.shape
        m = _matrix(score_matrix)
        if out is None:
            out = np.empty(n, dtype=np.int32)
        assert out.dtype == np.int32 and out.size >= n and out.flags.c_contiguous
        if packed:
            rc = self._lib.swb200_score_batch_packed(self._h, a.ctypes.data, b.ctypes.data, m.ctypes.data, _gap(gap_penalty), out.ctypes.data, n)
        elif width == SEQ_LEN:
            rc = self._lib.swb200_score_batch(self._h, a.ctypes.data, b.ctypes.data, m.ctypes.data, _gap(gap_penalty), out.ctypes.data, n)
        else:
            rc = self._lib.swb200_score_batch_len(self._h, width, a.ctypes.data, b.ctypes.data, m.ctypes.data, _gap(gap_penalty), out.ctypes.data, n)
        self._check(rc)
        return out[:n]

    # -- fixed 1/1/1 scoring (SmithWaterman_111 source.cpp:1073-1103, SmithWaterman_8bit111simd 1105-1225)
    def score_batch_111(self, seq1: np.ndarray, seq2: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        a = np.ascontiguousarray(seq1, dtype=np.uint8).reshape(-1, SEQ_LEN)
        b = np.ascontiguousarray(seq2, dtype=np.uint8).reshape(-1, SEQ_LEN)
        if a.shape != b.shape:
            raise ValueError("seq1 and seq2 must both be uint8 [n][128]")
        n = a.shape[0]
        if out is None:
            out = np.empty(n, dtype=np.int32)
        self._check(self._lib.swb200_score_batch_111(self._h, a.ctypes.data, b.ctypes.data, out.ctypes.data, n))
        return out[:n]

    # -- many queries vs one target (SmithWaterman_8b111x32mark1, source.cpp:1227-1234)
    def score_one_vs_many(self, seq1s: np.ndarray, seq2: np.ndarray, score_matrix=MATRIX_111, gap_penalty=GAP_111,
                          out: Optional[np.ndarray] = None) -> np.ndarray:
        a = np.ascontiguousarray(seq1s, dtype=np.uint8).reshape(-1, SEQ_LEN)
        b = np.ascontiguousarray(seq2, dtype=np.uint8).reshape(-1)
        if b.size != SEQ_LEN:
            raise ValueError("seq2 must hold exactly 128 codes")
        n = a.shape[0]
        m = _matrix(score_matrix)
        if out is None:
            out = np.empty(n, dtype=np.int32)
        self._check(self._lib.swb200_score_one_vs_many(self._h, a.ctypes.data, b.ctypes.data, m.ctypes.data, _gap(gap_penalty), out.ctypes.data, n))
        return out[:n]

    # -- adaptive-banded X-drop semi-global aligner (SemiGlobal_AdaptiveBanded_XDrop_111_32_70, source.cpp:1836-1976)
    def semiglobal_xdrop(self, seq1: np.ndarray, seq2: np.ndarray, traceback: bool = True) -> dict:
        """seq1, seq2: uint8 [n][len].  Returns score/end_y/end_x int32 [n] and, with traceback, n_ops [n] and
        ops uint8 [n][2*len] (0 = diagonal, 1 = down, 2 = right, forward order from (0,0))."""
        a = np.ascontiguousarray(seq1, dtype=np.uint8)
        b = np.ascontiguousarray(seq2, dtype=np.uint8)
        if a.ndim != 2 or a.shape != b.shape:
            raise ValueError("seq1 and seq2 must both be uint8 [n][len]")
        n, length = a.shape
        out = {k: np.empty(n, np.int32) for k in ("score", "end_y", "end_x")}
        n_ops = np.empty(n, np.int32) if traceback else None
        ops = np.empty((n, 2 * length), np.uint8) if traceback else None
        self._check(self._lib.swb200_semiglobal_xdrop_batch(
            self._h, a.ctypes.data, b.ctypes.data, length, n, out["score"].ctypes.data, out["end_y"].ctypes.data, out["end_x"].ctypes.data,
            n_ops.ctypes.data if traceback else None, ops.ctypes.data if traceback else None))
        if traceback:
            out["n_ops"], out["ops"] = n_ops, ops
        return out

    def semiglobal_xdrop_device(self, d_seq1, d_seq2, d_score, d_end_y, d_end_x, d_n_ops=None, d_ops=None,
                                device_index: int = 0, stream: Optional[int] = None):
        import torch
        n, length = d_seq1.shape
        if stream is None:
            stream = torch.cuda.current_stream(d_seq1.device).cuda_stream
        self._check(self._lib.swb200_semiglobal_xdrop_batch_device(
            self._h, device_index, d_seq1.data_ptr(), d_seq2.data_ptr(), length, n, d_score.data_ptr(), d_end_y.data_ptr(), d_end_x.data_ptr(),
            d_n_ops.data_ptr() if d_n_ops is not None else None, d_ops.data_ptr() if d_ops is not None else None, stream))

    def semiglobal_kernel_info(self, device_index: int = 0) -> dict:
        info = KernelInfo()
        self._check(self._lib.swb200_semiglobal_kernel_info(self._h, device_index, C.byref(info)))
        return {k: getattr(info, k) for k, _ in KernelInfo._fields_}

    def submit(self, seq1: np.ndarray, seq2: np.ndarray, score_matrix, gap_penalty, out: np.ndarray, packed: bool = False) -> int:
        """Asynchronous batch: returns a ticket; the arrays must stay alive and untouched until wait(ticket)."""
        assert seq1.flags.c_contiguous and seq2.flags.c_contiguous and out.flags.c_contiguous
        assert seq1.dtype == np.uint8 and seq2.dtype == np.uint8 and out.dtype == np.int32
        assert seq1.shape == seq2.shape and seq1.shape[1] == (32 if packed else SEQ_LEN) and out.size >= seq1.shape[0]
        m = _matrix(score_matrix)
        t = C.c_uint64()
        fn = self._lib.swb200_submit_packed if packed else self._lib.swb200_submit
        self._check(fn(self._h, seq1.ctypes.data, seq2.ctypes.data, m.ctypes.data, _gap(gap_penalty),
                       out.ctypes.data, seq1.shape[0], C.byref(t)))
        return int(t.value)

    def wait(self, ticket: int):
        self._check(self._lib.swb200_wait(self._h, ticket))

    # -- device-resident arrays (torch tensors on the context's GPU `device_index`)
    def score_batch_device(self, d_seq1, d_seq2, score_matrix, gap_penalty, d_scores, n: Optional[int] = None,
                           device_index: int = 0, stream: Optional[int] = None, packed: bool = False):
        import torch
        if n is None:
            n = d_seq1.shape[0]
        m = _matrix(score_matrix)
        if stream is None:
            stream = torch.cuda.current_stream(d_seq1.device).cuda_stream
        if packed:
            rc = self._lib.swb200_score_batch_packed_device(self._h, device_index, d_seq1.data_ptr(), d_seq2.data_ptr(), m.ctypes.data,
                                                            _gap(gap_penalty), d_scores.data_ptr(), n, stream)
        else:
            rc = self._lib.swb200_score_batch_len_device(self._h, device_index, int(d_seq1.shape[1]), d_seq1.data_ptr(), d_seq2.data_ptr(),
                                                         m.ctypes.data, _gap(gap_penalty), d_scores.data_ptr(), n, stream)
        self._check(rc)
        return d_scores

    def count_bad_codes_device(self, d_codes, device_index: int = 0) -> int:
        import torch
        bad = C.c_uint64()
        stream = torch.cuda.current_stream(d_codes.device).cuda_stream
        self._check(self._lib.swb200_validate_codes_device(self._h, device_index, d_codes.data_ptr(), d_codes.numel(), C.byref(bad), stream))
        return int(bad.value)


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            lo, hi = part.split("-")
            cpus.update(range(int(lo), int(hi) + 1))
        else:
            cpus.add(int(part))
    return cpus


def bind_to_gpu_numa_node(device_index: int, sysfs: str = "/sys") -> Optional[dict]:
    """Pins the calling process to the CPUs of the NUMA node the GPU is attached to, so that the
    pinned host buffers allocated afterwards (first touch) and the copy-issuing threads are local
    to that GPU's PCIe root.  One process per GPU on a multi-socket box otherwise sends half of
    its H2D traffic across the socket interconnect.  Returns what was done, or None."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        addr = f"{dom:04x}:{bus:02x}:{dev:02x}.0"
        with open(f"{sysfs}/bus/pci/devices/{addr}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return {"pci": addr, "numa_node": node, "bound": False}
        with open(f"{sysfs}/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        target = (cpus & allowed) or allowed
        os.sched_setaffinity(0, target)
        return {"pci": addr, "numa_node": node, "cpus": len(target), "bound": True}
    except Exception as e:   # best effort: never fail a run because sysfs looks different
        return {"bound": False, "error": str(e)[:80]}


def bind_rank_cpus(local_rank: int, local_world: int, device_index: Optional[int] = None) -> dict:
    """One process per GPU on a shared box: gives rank `local_rank` of `local_world` its own contiguous share of the CPUs
    this process may run on (of the GPU's NUMA node when sysfs names one), so that the library's auto-sized lane pool
    -- (CPUs available - GPUs) / GPUs PACK lanes plus the calling thread -- adds up to the box instead of every rank
    claiming all of it.  Returns what was done; undo with os.sched_setaffinity(0, result["before"])."""
    before = sorted(os.sched_getaffinity(0))
    numa = bind_to_gpu_numa_node(device_index if device_index is not None else local_rank) if local_world > 1 else None
    allowed = sorted(os.sched_getaffinity(0))
    out = {"before": before, "numa": numa, "cpus": len(allowed), "share": None}
    if local_world > 1 and len(allowed) >= local_world:
        # ranks whose GPUs sit on the same node split that node's CPUs among themselves; with one node that is everybody
        peers = local_world if (numa is None or not numa.get("bound")) else max(1, local_world * len(allowed) // max(len(before), 1))
        k = local_rank % peers
        lo, hi = len(allowed) * k // peers, len(allowed) * (k + 1) // peers
        mine = allowed[lo:hi]
        if mine:
            os.sched_setaffinity(0, mine)
            out["share"] = [mine[0], mine[-1]]
            out["cpus"] = len(mine)
    return out


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(n_devices=1)
    return _default_ctx


def SmithWaterman_b200(seq1, seq2, score_matrix, gap_penalty) -> int:
    """Drop-in for the reference's per-pair call (same argument order and meaning)."""
    return default_context().smith_waterman(seq1, seq2, score_matrix, gap_penalty)


def SemiGlobal_AdaptiveBanded_XDrop_111_32_70_b200(seq1, seq2):
    """Drop-in for the reference's aligner (source.cpp:1836-1838): returns (score, [(y, x), ...]) with the
    traceback from (0,0) to the best cell, exactly the reference's pair<int, vector<pair<int,int>>>."""
    r = default_context().semiglobal_xdrop(np.asarray(seq1, np.uint8)[None, :], np.asarray(seq2, np.uint8)[None, :])
    ops = r["ops"][0, :r["n_ops"][0]]
    ys = np.concatenate([[0], np.cumsum(ops != 2)])
    xs = np.concatenate([[0], np.cumsum(ops != 1)])
    return int(r["score"][0]), list(zip(ys.tolist(), xs.tolist()))


def SmithWaterman_111_b200(seq1, seq2) -> int:
    """Drop-in for SmithWaterman_111 / SmithWaterman_8bit111simd (source.cpp:1073-1076, 1105-1107)."""
    return int(default_context().score_batch_111(np.asarray(seq1, np.uint8)[None, :], np.asarray(seq2, np.uint8)[None, :])[0])


# ------------------------------------------------------------------ synthetic pairs
def reference_stream(n: int, seed: int = 10000, out=None):
    """Pairs [0,n) of the reference's own test stream (source.cpp:2944-2953)."""
    lib = load_library()
    a, b = out if out is not None else (np.empty((n, SEQ_LEN), np.uint8), np.empty((n, SEQ_LEN), np.uint8))
    rc = lib.swb200_gen_reference_stream(seed, n, a.ctypes.data, b.ctypes.data)
    if rc != 0:
        raise SwbError(rc, "swb200_gen_reference_stream")
    return a, b


def counter_pairs(first: int, n: int, seed: int = 10000, packed: bool = False, out=None, threads: int = 0):
    """Pairs [first, first+n) of the counter-based stream: pair k depends only on (seed, k),
    so any index range can be produced by any worker or rank."""
    lib = load_library()
    width = 32 if packed else SEQ_LEN
    a, b = out if out is not None else (np.empty((n, width), np.uint8), np.empty((n, width), np.uint8))
    assert a.shape == (n, width) and b.shape == (n, width) and a.flags.c_contiguous and b.flags.c_contiguous
    fn = lib.swb200_gen_counter_pairs_packed if packed else lib.swb200_gen_counter_pairs
    rc = fn(seed, first, n, a.ctypes.data, b.ctypes.data, threads or min(os.cpu_count() or 1, 16))
    if rc != 0:
        raise SwbError(rc, "swb200_gen_counter_pairs")
    return a, b


def related_pairs(first: int, n: int, seq_len: int = 16384, sub_pct: int = 10, ins_pct: int = 10, del_pct: int = 10,
                  seed: int = 10000, out=None, threads: int = 0):
    """Pairs [first, first+n) of TestSemiGlobal-style related sequences (source.cpp:2750-2771), counter-based."""
    lib = load_library()
    a, b = out if out is not None else (np.empty((n, seq_len), np.uint8), np.empty((n, seq_len), np.uint8))
    assert a.shape == (n, seq_len) and b.shape == (n, seq_len) and a.flags.c_contiguous and b.flags.c_contiguous
    rc = lib.swb200_gen_related_pairs(seed, first, n, seq_len, sub_pct, ins_pct, del_pct, a.ctypes.data, b.ctypes.data,
                                      threads or min(os.cpu_count() or 1, 16))
    if rc != 0:
        raise SwbError(rc, "swb200_gen_related_pairs")
    return a, b


def pack2bit(codes: np.ndarray) -> np.ndarray:
    """Byte codes [..., 4k] -> the reference's 2-bit packing [..., k] (inverse of `unpack`, source.cpp:1580-1583)."""
    c = np.ascontiguousarray(codes, dtype=np.uint8)
    if c.shape[-1] % 8:
        raise ValueError("the last dimension must be a multiple of 8 codes")
    out = np.empty(c.shape[:-1] + (c.shape[-1] // 4,), np.uint8)
    rc = load_library().swb200_pack2bit_host(c.ctypes.data, out.ctypes.data, c.size)
    if rc != 0:
        raise SwbError(rc, "swb200_pack2bit_host")
    return out


def unpack2bit(packed: np.ndarray, n_codes: int | None = None) -> np.ndarray:
    """The reference's `unpack` (source.cpp:1580-1583): [..., k] packed bytes -> [..., 4k] byte codes (1-D: the first n_codes)."""
    p = np.ascontiguousarray(packed, dtype=np.uint8)
    if n_codes is not None:
        if p.ndim != 1 or n_codes > 4 * p.size:
            raise ValueError("n_codes needs a 1-D input holding at least n_codes/4 bytes")
        out = np.empty(n_codes, np.uint8)
    else:
        out = np.empty(p.shape[:-1] + (p.shape[-1] * 4,), np.uint8)
    rc = load_library().swb200_unpack2bit_host(p.ctypes.data, out.ctypes.data, out.size)
    if rc != 0:
        raise SwbError(rc, "swb200_unpack2bit_host")
    return out


def host_read_bandwidth(buf: np.ndarray, threads: int, passes: int = 1) -> float:
    """Bytes per second at which `threads` host threads stream-read `buf` (benchmark helper, bench.py host_ceiling)."""
    assert buf.flags.c_contiguous
    out = C.c_double()
    rc = load_library().swb200_host_read_bandwidth(buf.ctypes.data, buf.nbytes, int(threads), int(passes), C.byref(out))
    if rc != 0:
        raise SwbError(rc, "swb200_host_read_bandwidth")
    return float(out.value)


def fnv1a64(scores: np.ndarray) -> int:
    s = np.ascontiguousarray(scores, dtype=np.int32)
    return int(load_library().swb200_fnv1a64_i32(s.ctypes.data, s.size))


def counter_pairs_numpy(first: int, n: int, seed: int = 10000):
    """The same stream restated in numpy (used by tests to pin the C++ generator)."""
    k = (np.arange(first, first + n, dtype=np.uint64)[:, None] * np.uint64(8) + np.arange(8, dtype=np.uint64)[None, :])
    with np.errstate(over="ignore"):
        x = k + np.uint64((int(seed) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)
        x = (x + np.uint64(0x9E3779B97F4A7C15))
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    shifts = (np.arange(32, dtype=np.uint64) * np.uint64(2))[None, None, :]
    codes = ((x[:, :, None] >> shifts) & np.uint64(3)).astype(np.uint8).reshape(n, 256)
    return np.ascontiguousarray(codes[:, :128]), np.ascontiguousarray(codes[:, 128:])
