import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Native artefacts (CUDA library, host layer, oracle) built in-tree; no GPU needed to build."""
    import __graft_entry__ as g

    from sift_b200 import capi

    need = [capi.LIB_PATH, os.path.join(ROOT, "oracle", "_build", "liboracle.so"), os.path.join(ROOT, "sift_b200", "sift")]
    if not all(os.path.exists(p) for p in need):
        g.build()
    return True


@pytest.fixture(scope="session")
def parrot():
    import numpy as np

    return np.load(os.path.join(ROOT, "tests", "golden", "parrot_r.npy")).astype(np.float32)
