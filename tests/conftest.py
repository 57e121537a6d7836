import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "optical-flow-optimal-transport_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def relerr(a, ref):
    """max-abs difference over max-abs reference: the 'relative per flow component' of the
    parity contract (BASELINE.json north_star: <= 1e-9)."""
    a = np.asarray(a); ref = np.asarray(ref)
    return float(np.max(np.abs(a - ref)) / max(float(np.max(np.abs(ref))), 1e-300))


def epe_max(u, v, ur, vr):
    return float(np.max(np.sqrt((np.asarray(u) - ur) ** 2 + (np.asarray(v) - vr) ** 2)))


def gpu_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    o.build()
    return o
