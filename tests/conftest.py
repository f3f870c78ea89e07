import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (C++ restatement, ctypes).  Built on demand with oracle/Makefile."""
    from oracle import oracle_api

    oracle_api.build()
    oracle_api.load()
    oracle_api.set_constants()
    return oracle_api


@pytest.fixture(scope="session")
def rbis_lib():
    """librbis_b200.so through ctypes; built on demand (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g

    g.build_cuda()
    from pronto_b200 import capi

    return capi.load()
