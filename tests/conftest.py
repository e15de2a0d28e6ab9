import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def rel_err(a, b):
    """max|a-b| / max|b| — the norm-relative parity criterion (individual gamma components cancel
    to ~0, so an element-wise relative error is meaningless)."""
    a = np.asarray(a)
    b = np.asarray(b)
    scale = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (scale if scale > 0 else 1.0))


TOL_F64 = 1e-12   # BASELINE.json north_star: relative 1e-12 in FP64
TOL_F32 = 2e-5    # FP32 instantiations: ~100 ulp of the buffer scale


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.lib()
    return orc
