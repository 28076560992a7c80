import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def engine_lib():
    """The product library, built in-tree if stale (nvcc cross-compiles without a GPU)."""
    from ismcsolver_b200.build import build_library
    from ismcsolver_b200 import capi
    build_library()
    return capi.load_library()


@pytest.fixture(scope="session")
def engine(engine_lib):
    """A search context on cuda:0 — GPU tests only; fails loudly without a device."""
    from ismcsolver_b200 import capi
    eng = capi.Engine(0)
    yield eng
    eng.close()
