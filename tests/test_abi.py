"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol
include/swb200.h declares; host-side argument handling fails loudly.  No compute call is
made here (there is no GPU and no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "swb200.h")


@pytest.fixture(scope="module")
def lib(swb):
    import build as swb_build   # smith-waterman-simd_b200/build.py
    swb_build.build()
    return swb.load_library()


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(swb200_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(swb):
    assert declared_symbols() == sorted(swb.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert getattr(lib, name) is not None, name
    out = subprocess.run(["nm", "-D", "--defined-only", lib._name], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (swb200_[a-z0-9_]+)", out))
    assert exported == set(declared_symbols())


def test_library_is_sm100a_only(lib):
    out = subprocess.run(["cuobjdump", "-lelf", lib._name], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_kernel_uses_packed_int16_instructions(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib._name], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    sass = out.stdout
    for mnemonic in ("VIADDMNMX.S16x2", "VIMNMX3.S16x2", "PRMT"):
        assert mnemonic in sass, mnemonic
    assert "HMMA" not in sass and "UTC" not in sass   # no tensor cores: this is not a contraction


def test_strerror_and_no_device_is_loud(lib, swb):
    assert lib.swb200_strerror(0) == b"ok"
    assert b"domain" in lib.swb200_strerror(swb.ERR_DOMAIN)
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device error path is for the CPU container")
    with pytest.raises(swb.SwbError) as e:
        swb.Context(n_devices=1)
    assert e.value.code == swb.ERR_NO_DEVICE
    with pytest.raises(swb.SwbError):
        swb.SmithWaterman_b200(np.zeros(128, np.uint8), np.zeros(128, np.uint8), swb.MATRIX_SPEEDTEST, 15)


def test_host_argument_checks(swb):
    with pytest.raises(ValueError):
        swb._matrix([1, 2, 3])
    with pytest.raises(ValueError):
        swb._matrix([300] * 16)
    with pytest.raises(swb.SwbError):
        swb._gap(200)


def test_counter_pairs_are_index_addressable(swb):
    a, b = swb.counter_pairs(0, 1000)
    an, bn = swb.counter_pairs_numpy(0, 1000)          # independent restatement of the C++ generator
    assert np.array_equal(a, an) and np.array_equal(b, bn)
    a2, b2 = swb.counter_pairs(250, 500)
    assert np.array_equal(a[250:750], a2) and np.array_equal(b[250:750], b2)
    assert a.max() == 3 and a.min() == 0
    hist = np.bincount(np.concatenate([a.ravel(), b.ravel()]), minlength=4) / (a.size + b.size)
    assert np.all(np.abs(hist - 0.25) < 0.01)
    assert not np.array_equal(a, b)
    pa, pb = swb.counter_pairs(250, 500, packed=True)  # 2-bit layout of source.cpp:1580-1583
    unpacked = ((pa[:, :, None] >> (2 * np.arange(4, dtype=np.uint8))[None, None, :]) & 3).reshape(500, 128)
    assert np.array_equal(unpacked, a2)


def test_reference_stream_generator_matches_oracle(swb, oracle):
    a, b = swb.reference_stream(3000)
    ao, bo = oracle.reference_stream(3000)
    assert np.array_equal(a, ao) and np.array_equal(b, bo)
    s = oracle.score_batch(a[:16], b[:16], oracle.MATRIX_SPEEDTEST, 15)
    assert swb.fnv1a64(s) == oracle.fnv1a64(s)


def test_params_derivation_matches_documented_domain():
    # sw_params.h through the host emulator build: rc = 1 fast, 0 general, -1 refused
    lib_path = os.path.join(ROOT, "tests", "emu", "libswemu.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", f"-I{os.path.join(ROOT, 'smith-waterman-simd_b200', 'csrc')}",
                    "-o", lib_path, os.path.join(ROOT, "tests", "emu", "emu_main.cpp")], check=True)
    emu = C.CDLL(lib_path)
    emu.swemu_score_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_int]
    z = np.zeros((2, 128), np.uint8)
    out = np.zeros(2, np.int32)

    def path(sm, g):
        m = np.asarray(sm, dtype=np.int8)
        return emu.swemu_score_batch(z.ctypes.data, z.ctypes.data, m.ctypes.data, g, out.ctypes.data, 2, 0)
    mm = lambda m, x: [m if i == j else x for i in range(4) for j in range(4)]
    assert path(mm(10, -30), 15) == 1 and path(mm(1, -1), 1) == 1
    assert path(mm(127, -127), 127) == 0
    assert path(mm(10, -128), 15) == -1      # -128 is outside the reference's domain (source.cpp:492)
    assert path(mm(10, -30), -1) == -1


def test_host_packer_is_the_inverse_of_the_reference_unpack(swb, oracle):
    # hostpack.cpp (AVX2 and the SWAR tail) against the oracle's restatement of source.cpp:1580-1583
    rng = np.random.default_rng(11)
    for n_codes in (8, 24, 64, 128, 136, 128 * 257):
        codes = rng.integers(0, 4, n_codes, dtype=np.uint8)
        packed = swb.pack2bit(codes)
        assert packed.shape == (n_codes // 4,)
        # source.cpp:1580-1583: dest[i*4+j] = (src[i] >> 2j) & 3
        assert np.array_equal(((packed[:, None] >> (2 * np.arange(4, dtype=np.uint8))[None, :]) & 3).reshape(-1), codes)
    codes = rng.integers(0, 4, (300, 128), dtype=np.uint8)
    assert np.array_equal(swb.pack2bit(codes), oracle.pack2bit(codes))
    # codes above 3 are masked to two bits, never smeared into a neighbour
    dirty = codes | rng.integers(0, 64, codes.shape, dtype=np.uint8) << 2
    assert np.array_equal(swb.pack2bit(dirty.astype(np.uint8)), oracle.pack2bit(codes))
    with pytest.raises(ValueError):
        swb.pack2bit(np.zeros(12, np.uint8))


def test_host_packer_alignments_tails_and_the_streaming_store_path(swb):
    # hostpack.cpp takes non-temporal stores when the destination is 32-byte aligned and the buffer is large, plain
    # stores otherwise, 128 codes per iteration with a SWAR tail: every combination must give the same bytes and must
    # not write outside its output.
    import ctypes as C
    lib = swb.load_library()
    rng = np.random.default_rng(12)
    for n_codes in (8, 120, 128, 136, 65536, 65536 + 8, 65536 + 128 + 40, 1 << 19):
        src_buf = rng.integers(0, 256, n_codes + 64, dtype=np.uint8)          # any byte: the packer masks to two bits
        for off_in in (0, 1, 13):
            for off_out in (0, 1, 32):
                out_buf = np.full(n_codes // 4 + 128, 0xAA, np.uint8)
                base = out_buf.ctypes.data
                j = (-base) % 32 + off_out
                codes = src_buf[off_in:off_in + n_codes]
                rc = lib.swb200_pack2bit_host(C.c_void_p(codes.ctypes.data), C.c_void_p(base + j), C.c_uint64(n_codes))
                assert rc == 0
                c = (codes & 3).reshape(-1, 4).astype(np.uint8)
                want = c[:, 0] | (c[:, 1] << 2) | (c[:, 2] << 4) | (c[:, 3] << 6)
                assert np.array_equal(out_buf[j:j + n_codes // 4], want), (n_codes, off_in, off_out)
                assert (out_buf[:j] == 0xAA).all() and (out_buf[j + n_codes // 4:] == 0xAA).all(), (n_codes, off_in, off_out)


def test_host_unpacker_is_the_references_unpack(swb):
    # swb200_unpack2bit_host = `unpack` (source.cpp:1580-1583): dest[4i + j] = (src[i] >> 2j) & 3, for any count of codes,
    # any alignment (non-temporal stores when the destination is 32-byte aligned and long enough), never a byte outside
    # its output; and it is the inverse of the host packer.
    import ctypes as C
    lib = swb.load_library()
    rng = np.random.default_rng(14)
    for n_codes in (1, 3, 4, 5, 127, 128, 129, 4096, 4096 + 128, 4096 + 131, 32768, 100003):
        packed = rng.integers(0, 256, (n_codes + 3) // 4 + 8, dtype=np.uint8)
        want = ((packed[:, None] >> (2 * np.arange(4, dtype=np.uint8))[None, :]) & 3).reshape(-1)[:n_codes]
        for off_in in (0, 1):
            for off_out in (0, 1, 32):
                out_buf = np.full(n_codes + 160, 0xAA, np.uint8)
                base = out_buf.ctypes.data
                j = (-base) % 32 + off_out
                src = np.ascontiguousarray(np.concatenate([np.zeros(off_in, np.uint8), packed]))[off_in:]
                rc = lib.swb200_unpack2bit_host(C.c_void_p(src.ctypes.data), C.c_void_p(base + j), C.c_uint64(n_codes))
                assert rc == 0
                assert np.array_equal(out_buf[j:j + n_codes], want), (n_codes, off_in, off_out)
                assert (out_buf[:j] == 0xAA).all() and (out_buf[j + n_codes:] == 0xAA).all(), (n_codes, off_in, off_out)
    codes = rng.integers(0, 4, (64, 16384), dtype=np.uint8)
    assert np.array_equal(swb.unpack2bit(swb.pack2bit(codes)), codes)
    assert np.array_equal(swb.unpack2bit(swb.pack2bit(codes).reshape(-1), 1001), codes.reshape(-1)[:1001])


def test_the_product_never_touches_the_oracle_and_bench_only_in_its_baseline_legs():
    # The oracle is test infrastructure: nothing under smith-waterman-simd_b200/ or include/ may import, link or load it,
    # libswb200.so must not depend on it, and bench.py may reach it only inside the cpu_baseline / --impl reference legs.
    import ast
    import re
    import subprocess
    pkg = os.path.join(ROOT, "smith-waterman-simd_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", ".inc")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"(import\s+oracle|from\s+oracle|libsworacle|libswref|oracle/)", text), os.path.join(dirpath, f)
    needed = subprocess.run(["readelf", "-d", os.path.join(pkg, "libswb200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in needed and "swref" not in needed
    # bench.py: every `from oracle import ...` sits in a function of the baseline / reference legs or under `if not args.no_cpu_baseline`
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    allowed_functions = {"cpu_reference_run", "run_cpu_table", "sg_cpu_reference", "run_reference_arm"}   # the reference arm loads NOTHING of the product
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        for node in ast.walk(fn):
            if isinstance(node, ast.ImportFrom) and node.module == "oracle":
                if fn.name in allowed_functions:
                    continue
                guarded = any(isinstance(g, ast.If) and "no_cpu_baseline" in ast.unparse(g.test) and node in list(ast.walk(g))
                              for g in ast.walk(fn))
                assert guarded, (fn.name, node.lineno)
