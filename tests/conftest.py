import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle.get()


@pytest.fixture(scope="session")
def kat():
    import json
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def aaa_luma():
    return np.load(os.path.join(GOLDEN, "aaa_orig_luma512.npz"))["luma"].astype(np.float32)


@pytest.fixture(scope="session")
def cuda():
    """The product library through its C ABI (ctypes).  Fails loudly without a GPU."""
    import yume_pdq_b200 as yp
    return yp


def hamming(a, b) -> int:
    a = np.frombuffer(bytes(a), dtype=np.uint8)
    b = np.frombuffer(bytes(b), dtype=np.uint8)
    return int(np.unpackbits(a ^ b).sum())
