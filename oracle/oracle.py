"""ctypes loader for the test oracle.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product (smith-waterman-simd_b200/) never does.

Two libraries:
  * libsworacle.so  -- oracle/sw_oracle.c, the plain-C restatement of
    /root/reference/source.cpp:35-60 (+ generator 2944-2953, unpack 1580-1583).
  * _ref/libswref.so -- oracle/ref_wrap.cpp, the unmodified reference compiled where it
    lies; present when it was built in the authoring container (it travels to the GPU box
    as a built file).  `have_ref()` says whether it is loadable.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_PORT = os.path.join(HERE, "libsworacle.so")
_REF = os.path.join(HERE, "_ref", "libswref.so")

_u8p = C.POINTER(C.c_uint8)
_i8p = C.POINTER(C.c_int8)
_i32p = C.POINTER(C.c_int32)


def build(quiet: bool = True) -> None:
    """Compile the oracle (and, when /root/reference exists, oracle/_ref)."""
    subprocess.run(["make", "-C", HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


_port = None
_ref = None


def port():
    global _port
    if _port is None:
        if not os.path.exists(_PORT):
            build()
        lib = C.CDLL(_PORT)
        lib.swo_score.restype = C.c_int32
        lib.swo_score.argtypes = [_u8p, C.c_int, _u8p, C.c_int, _i8p, C.c_int]
        lib.swo_score_batch.restype = None
        lib.swo_score_batch.argtypes = [_u8p, _u8p, C.c_int, _i8p, C.c_int, _i32p, C.c_uint64]
        lib.swo_score_batch_mt.restype = None
        lib.swo_score_batch_mt.argtypes = [_u8p, _u8p, C.c_int, _i8p, C.c_int, _i32p, C.c_uint64, C.c_int]
        lib.swo_reference_stream.restype = None
        lib.swo_reference_stream.argtypes = [C.c_uint64, C.c_uint64, _u8p, _u8p]
        lib.swo_fnv1a64_scores.restype = C.c_uint64
        lib.swo_fnv1a64_scores.argtypes = [_i32p, C.c_uint64]
        lib.swo_unpack2bit.restype = None
        lib.swo_unpack2bit.argtypes = [_u8p, _u8p, C.c_uint64]
        lib.swo_pack2bit.restype = None
        lib.swo_pack2bit.argtypes = [_u8p, _u8p, C.c_uint64]
        lib.swo_semiglobal_xdrop.restype = C.c_int
        lib.swo_semiglobal_xdrop.argtypes = [_u8p, _u8p, C.c_int, _i32p, _i32p, _i32p, _u8p, _i32p]
        _port = lib
    return _port


def have_ref() -> bool:
    try:
        ref()
        return True
    except OSError:
        return False


def ref():
    global _ref
    if _ref is None:
        if not os.path.exists(_REF) and os.path.exists("/root/reference/source.cpp"):
            build()
        lib = C.CDLL(_REF)
        lib.swref_score_batch.restype = C.c_int
        lib.swref_score_batch.argtypes = [C.c_int, _u8p, _u8p, _i8p, C.c_int, _i32p, C.c_uint64, C.c_int]
        lib.swref_score_repeat.restype = C.c_int
        lib.swref_score_repeat.argtypes = [C.c_int, _u8p, _u8p, _i8p, C.c_int, C.c_uint64]
        lib.swref_reference_stream.restype = None
        lib.swref_reference_stream.argtypes = [C.c_uint64, C.c_uint64, _u8p, _u8p]
        lib.swref_unpack.restype = None
        lib.swref_unpack.argtypes = [_u8p, _u8p]
        lib.swref_111.restype = C.c_int
        lib.swref_111.argtypes = [_u8p, _u8p]
        lib.swref_8bit111.restype = C.c_int
        lib.swref_8bit111.argtypes = [_u8p, _u8p]
        lib.swref_semiglobal.restype = C.c_int
        lib.swref_semiglobal.argtypes = [C.c_int, _u8p, _u8p, _i32p, _i32p, C.c_int64, _i32p]
        lib.swref_semiglobal_batch.restype = C.c_int
        lib.swref_semiglobal_batch.argtypes = [C.c_int, _u8p, _u8p, C.c_uint64, _i32p, C.c_int]
        lib.swref_semiglobal_test_inputs.restype = None
        lib.swref_semiglobal_test_inputs.argtypes = [C.c_uint64, C.c_uint64, _u8p, _u8p]
        lib.swref_semiglobal_speedtest_input.restype = None
        lib.swref_semiglobal_speedtest_input.argtypes = [C.c_uint64, _u8p, _u8p]
        lib.swref_x32.restype = C.c_int
        lib.swref_x32.argtypes = [C.c_int, _u8p, _u8p, _i32p]
        lib.swref_hardware_threads.restype = C.c_int
        _ref = lib
    return _ref


def _p(a, t):
    return a.ctypes.data_as(t)


def _mat(score_matrix) -> np.ndarray:
    m = np.ascontiguousarray(np.asarray(score_matrix, dtype=np.int8).reshape(16))
    return m


def score_batch(seq1: np.ndarray, seq2: np.ndarray, score_matrix, gap: int, threads: int = 1) -> np.ndarray:
    """Restatement (sw_oracle.c). seq1/seq2: uint8 [n][L] codes 0..3."""
    seq1 = np.ascontiguousarray(seq1, dtype=np.uint8)
    seq2 = np.ascontiguousarray(seq2, dtype=np.uint8)
    assert seq1.shape == seq2.shape and seq1.ndim == 2
    n, L = seq1.shape
    out = np.empty(n, dtype=np.int32)
    m = _mat(score_matrix)
    port().swo_score_batch_mt(_p(seq1, _u8p), _p(seq2, _u8p), L, _p(m, _i8p), int(gap), _p(out, _i32p), n, threads)
    return out


def ref_score_repeat(variant: int, seq1: np.ndarray, seq2: np.ndarray, score_matrix, gap: int, calls: int) -> float:
    """The reference's own way of timing (SpeedTest, source.cpp:3047-3055): ONE pair scored `calls` times by variant
    0 = scalar, 1..9 = simd..simd9.  Returns seconds per call."""
    import time
    a = np.ascontiguousarray(seq1, dtype=np.uint8).reshape(128)
    b = np.ascontiguousarray(seq2, dtype=np.uint8).reshape(128)
    m = _mat(score_matrix)
    lib = ref()
    lib.swref_score_repeat(variant, _p(a, _u8p), _p(b, _u8p), _p(m, _i8p), int(gap), 1000)
    t = time.perf_counter()
    lib.swref_score_repeat(variant, _p(a, _u8p), _p(b, _u8p), _p(m, _i8p), int(gap), int(calls))
    return (time.perf_counter() - t) / calls


def ref_score_batch(variant: int, seq1: np.ndarray, seq2: np.ndarray, score_matrix, gap: int, threads: int = 1) -> np.ndarray:
    """The reference itself. variant 0 = scalar, 1..9 = SmithWaterman_simd..simd9. L must be 128."""
    seq1 = np.ascontiguousarray(seq1, dtype=np.uint8)
    seq2 = np.ascontiguousarray(seq2, dtype=np.uint8)
    assert seq1.shape == seq2.shape and seq1.ndim == 2 and seq1.shape[1] == 128
    n = seq1.shape[0]
    out = np.empty(n, dtype=np.int32)
    m = _mat(score_matrix)
    rc = ref().swref_score_batch(variant, _p(seq1, _u8p), _p(seq2, _u8p), _p(m, _i8p), int(gap), _p(out, _i32p), n, threads)
    if rc != 0:
        raise ValueError(f"unknown reference variant {variant}")
    return out


def ref_x32(mark: int, seq1_32: np.ndarray, seq2: np.ndarray) -> np.ndarray:
    """The reference's SmithWaterman_8b111x32mark{1,2,3}: 32 queries [32][128] vs one target [128], fixed 1/1/1."""
    a = np.ascontiguousarray(seq1_32, dtype=np.uint8).reshape(32 * 128)
    b = np.ascontiguousarray(seq2, dtype=np.uint8).reshape(128)
    out = np.empty(32, dtype=np.int32)
    rc = ref().swref_x32(mark, _p(a, _u8p), _p(b, _u8p), _p(out, _i32p))
    assert rc == out[0]          # the reference returns dest[0] (source.cpp:1233, 1296)
    return out


def ref_111(seq1: np.ndarray, seq2: np.ndarray) -> int:
    a = np.ascontiguousarray(seq1, dtype=np.uint8).reshape(128)
    b = np.ascontiguousarray(seq2, dtype=np.uint8).reshape(128)
    return int(ref().swref_111(_p(a, _u8p), _p(b, _u8p)))


def ref_8bit111(seq1: np.ndarray, seq2: np.ndarray) -> int:
    """The reference's SmithWaterman_8bit111simd (source.cpp:1105-1225)."""
    a = np.ascontiguousarray(seq1, dtype=np.uint8).reshape(128)
    b = np.ascontiguousarray(seq2, dtype=np.uint8).reshape(128)
    return int(ref().swref_8bit111(_p(a, _u8p), _p(b, _u8p)))


def x32_stream(iterations: int, seed: int = 10000):
    """Inputs of TestSimdSmithWaterman111x32 (source.cpp:3004-3013): per iteration 4096 draws for the 32
    queries, then 128 draws for the target."""
    n = iterations * (4096 + 128)
    flat1 = np.empty((n // 2 // 128 + 1, 128), dtype=np.uint8)
    flat2 = np.empty_like(flat1)
    port().swo_reference_stream(seed, flat1.shape[0], _p(flat1, _u8p), _p(flat2, _u8p))
    draws = np.stack([flat1, flat2], axis=2).reshape(-1)[:n]     # undo the a/b interleaving: draw order
    draws = draws.reshape(iterations, 4096 + 128)
    return np.ascontiguousarray(draws[:, :4096].reshape(iterations, 32, 128)), np.ascontiguousarray(draws[:, 4096:])


def reference_stream(n: int, seed: int = 10000, use_ref: bool = False):
    """Pairs [0,n) of the reference's test stream (source.cpp:2944-2953)."""
    a = np.empty((n, 128), dtype=np.uint8)
    b = np.empty((n, 128), dtype=np.uint8)
    if use_ref:
        ref().swref_reference_stream(seed, n, _p(a, _u8p), _p(b, _u8p))
    else:
        port().swo_reference_stream(seed, n, _p(a, _u8p), _p(b, _u8p))
    return a, b


def fnv1a64(scores: np.ndarray) -> int:
    scores = np.ascontiguousarray(scores, dtype=np.int32)
    return int(port().swo_fnv1a64_scores(_p(scores, _i32p), scores.size))


def pack2bit(codes: np.ndarray) -> np.ndarray:
    codes = np.ascontiguousarray(codes, dtype=np.uint8)
    n = codes.shape[0]
    assert codes.shape[1] == 128
    out = np.empty((n, 32), dtype=np.uint8)
    port().swo_pack2bit(_p(codes, _u8p), _p(out, _u8p), n)
    return out


def unpack2bit(packed: np.ndarray) -> np.ndarray:
    packed = np.ascontiguousarray(packed, dtype=np.uint8)
    n = packed.shape[0]
    assert packed.shape[1] == 32
    out = np.empty((n, 128), dtype=np.uint8)
    port().swo_unpack2bit(_p(packed, _u8p), _p(out, _u8p), n)
    return out


MATRIX_SPEEDTEST = [10, -30, -30, -30, -30, 10, -30, -30, -30, -30, 10, -30, -30, -30, -30, 10]  # source.cpp:3041-3045
GAP_SPEEDTEST = 15                                                                           # source.cpp:3046
MATRIX_111 = [1, -1, -1, -1, -1, 1, -1, -1, -1, -1, 1, -1, -1, -1, -1, 1]                    # source.cpp:3202-3206
GAP_111 = 1                                                                                  # source.cpp:3207


# ------------------------------------------------------------------ semi-global X-drop aligner
SG_LEN = 16384          # source.cpp:1837-1838
SG_OPS = ("diagonal", "down", "right")


def semiglobal_xdrop(seq1: np.ndarray, seq2: np.ndarray):
    """oracle/sg_oracle.c on one pair of equal length.  Returns (score, end_y, end_x, ops uint8[n]) with
    ops in forward order from (0,0): 0 = diagonal, 1 = down (y+1), 2 = right (x+1)."""
    a = np.ascontiguousarray(seq1, dtype=np.uint8).reshape(-1)
    b = np.ascontiguousarray(seq2, dtype=np.uint8).reshape(-1)
    assert a.size == b.size and a.size >= 1
    ops = np.empty(2 * a.size, np.uint8)
    score, ey, ex, n = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    rc = port().swo_semiglobal_xdrop(_p(a, _u8p), _p(b, _u8p), a.size, C.byref(score), C.byref(ey), C.byref(ex), _p(ops, _u8p), C.byref(n))
    if rc != 0:
        raise RuntimeError(f"swo_semiglobal_xdrop rc={rc}")
    return int(score.value), int(ey.value), int(ex.value), ops[:n.value].copy()


def ops_to_traceback(ops: np.ndarray) -> np.ndarray:
    """The reference's traceback vector ((0,0) ... best cell) as int32 [n+1][2] from the op string."""
    ops = np.asarray(ops)
    dy = np.cumsum((ops != 2).astype(np.int32))
    dx = np.cumsum((ops != 1).astype(np.int32))
    tb = np.zeros((ops.size + 1, 2), np.int32)
    tb[1:, 0] = dy
    tb[1:, 1] = dx
    return tb


def ref_semiglobal(variant: int, seq1: np.ndarray, seq2: np.ndarray):
    """The reference's SemiGlobal_AdaptiveBanded_XDrop_111_32_70 (variant 0, source.cpp:1836-1976) or its AVX2
    forms _simd / _simd_mark2 / _mark3 / _mark4 (variants 1-4).  Returns (score, traceback int32 [n][2])."""
    a = np.ascontiguousarray(seq1, dtype=np.uint8).reshape(SG_LEN)
    b = np.ascontiguousarray(seq2, dtype=np.uint8).reshape(SG_LEN)
    cap = 2 * SG_LEN + 2
    tb = np.empty((cap, 2), np.int32)
    score, n = C.c_int32(), C.c_int32()
    rc = ref().swref_semiglobal(variant, _p(a, _u8p), _p(b, _u8p), C.byref(score), _p(tb, _i32p), cap, C.byref(n))
    assert rc == 0
    return int(score.value), tb[:n.value].copy()


def ref_semiglobal_batch(variant: int, seq1: np.ndarray, seq2: np.ndarray, threads: int = 1) -> np.ndarray:
    a = np.ascontiguousarray(seq1, dtype=np.uint8).reshape(-1, SG_LEN)
    b = np.ascontiguousarray(seq2, dtype=np.uint8).reshape(-1, SG_LEN)
    out = np.empty(a.shape[0], np.int32)
    rc = ref().swref_semiglobal_batch(variant, _p(a, _u8p), _p(b, _u8p), a.shape[0], _p(out, _i32p), threads)
    assert rc == 0
    return out


def ref_semiglobal_test_inputs(n: int, seed: int = 10000):
    """Inputs of the reference's TestSemiGlobal (source.cpp:2734-2771), generated by the reference's own code path."""
    a = np.empty((n, SG_LEN), np.uint8)
    b = np.empty((n, SG_LEN), np.uint8)
    ref().swref_semiglobal_test_inputs(seed, n, _p(a, _u8p), _p(b, _u8p))
    return a, b


def ref_semiglobal_speedtest_input(seed: int = 10000):
    a = np.empty(SG_LEN, np.uint8)
    b = np.empty(SG_LEN, np.uint8)
    ref().swref_semiglobal_speedtest_input(seed, _p(a, _u8p), _p(b, _u8p))
    return a, b
