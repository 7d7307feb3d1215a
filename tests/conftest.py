import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """CPU oracle A (C, oracle/kmer_oracle.c).  Test infrastructure only."""
    from oracle import oracle_a

    oracle_a.build()
    return oracle_a


@pytest.fixture(scope="session")
def apgk_lib():
    """libapgk.so via ctypes; built on demand (nvcc cross-compiles without a GPU)."""
    from allpathslg_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    return _lib.lib()
