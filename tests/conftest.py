import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import mzoracle

    mzoracle.build()
    return mzoracle


@pytest.fixture(scope="session")
def sm():
    """The product package (simd-minimizers_b200); builds libmzb200.so if it is missing."""
    lib = os.path.join(ROOT, "simd-minimizers_b200", "libmzb200.so")
    if not os.path.exists(lib):
        import __graft_entry__ as g

        g.build()
    return importlib.import_module("simd-minimizers_b200")
