import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def product():
    """The CUDA product through its C ABI. Fails (does not skip) if the extension is missing."""
    import ogmaneo2_b200 as og
    from ogmaneo2_b200 import build
    if build.stale() and os.path.exists(build.NVCC):
        build.build()
    return og


@pytest.fixture(scope="session")
def oracle():
    """(FlatLib, kind) of the strongest CPU checker available: the compiled reference, else the port."""
    from oracle import loader
    return loader.best()
