import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def compile_ctx():
    """Compile-only context: NVRTC + metadata, no device."""
    import lensed_b200 as L
    ctx = L.Context(device=-1)
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def gpu_ctx():
    import lensed_b200 as L
    ctx = L.Context(device=0)
    yield ctx
    ctx.close()
