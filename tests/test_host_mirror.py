"""The C++ host mirror (host/smith_waterman_b200.hpp) compiled the way the reference's own
test would use it (tests/emu/host_mirror_example.cpp; cf. source.cpp:2943-2982)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "smith-waterman-simd_b200")
EXE = os.path.join(ROOT, "tests", "emu", "host_mirror_example")


def build_example():
    import build as swb_build
    swb_build.build()
    subprocess.run(["g++", "-std=c++17", "-O2", "-o", EXE, os.path.join(ROOT, "tests", "emu", "host_mirror_example.cpp"),
                    "-L", PKG, "-lswb200", f"-Wl,-rpath,{PKG}"], check=True)


def test_host_mirror_compiles_and_fails_loudly_without_gpu(swb):
    import torch
    build_example()
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu-marked test")
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 2 and "no CUDA device" in r.stdout   # an exception with the library's message, not a CPU answer


@pytest.mark.gpu
def test_host_mirror_matches_known_answers_on_gpu(swb):
    build_example()
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "16/16 scores equal" in r.stdout
