import json
import os
import sys

import numpy as np
import pytest

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # before any CUDA context exists: see smith-waterman-simd_b200/swb200.py

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "smith-waterman-simd_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        meta = json.load(f)
    meta["structured_npz"] = np.load(os.path.join(GOLDEN, "structured.npz"))
    return meta


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def swb():
    import swb200
    return swb200


@pytest.fixture(scope="session")
def ctx(swb):
    """One library context on cuda:0 for the whole GPU session.  Fails loudly (no skip, no
    CPU substitute) when the extension or the GPU is missing."""
    c = swb.Context(n_devices=1)
    yield c
    c.close()
