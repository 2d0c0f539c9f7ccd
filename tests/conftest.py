import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib():
    """libb2c.so through the product's own loader (built in-tree by __graft_entry__.build())."""
    from clip_assisted_data_labeling_b200 import _build, _lib
    if not os.path.exists(_lib.LIB_PATH):
        _build.build_library()
    return _lib.load()


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return load


def reference_present():
    return os.path.isfile("/root/reference/utils/embedder.py")
