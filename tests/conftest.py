import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): built on demand with gcc."""
    from oracle import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def capi():
    """The product library.  Built in-tree by nvcc when missing (cross-compiles without a GPU)."""
    from btb_phylo_b200 import build
    build.build()
    from btb_phylo_b200 import _capi
    _capi.lib()
    return _capi


def golden(name):
    with open(os.path.join(GOLDEN, name), "rb") as f:
        return f.read()
