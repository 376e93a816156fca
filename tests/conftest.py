import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def engine():
    """one engine (wd_handle) on cuda:0 for the GPU parity tests; the CUDA library must be the thing that runs."""
    import wignerd.jl_b200 as w
    eng = w.Engine(0)
    yield eng
    eng.close()
