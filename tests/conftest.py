"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` covers the oracle against the committed reference goldens, the host logic and
the C-ABI surface; `-m gpu` are the parity tests proper (CUDA path vs oracle / goldens).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as d:
        return {k: d[k] for k in d.files}


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session", params=["scenarios", "config4_sample", "config5_sample"])
def golden(request):
    return request.param, load_golden(request.param)
