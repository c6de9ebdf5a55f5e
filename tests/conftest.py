import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# the plugin directory of every test: the reference's own objects/*.cl, verbatim
# (tests/golden/objects/README.md); the library ships no object files
OBJECTS_DIR = os.path.join(ROOT, "tests", "golden", "objects")


def _ensure_built():
    """Built artefacts are git-ignored: on a fresh checkout compile the C-ABI
    library and the oracle before anything imports them (same recipes as
    __graft_entry__.build)."""
    import subprocess
    if not os.path.exists(os.path.join(ROOT, "lensed_b200", "liblensed_cuda.so")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "lensed_b200", "csrc")], check=True)
    if not all(os.path.exists(os.path.join(ROOT, "oracle", n)) for n in ("liboracle.so", "liboracle_f64.so", "liboracle_fast.so")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True)
    ref = os.environ.get("LENSED_REFERENCE", "/root/reference")
    if os.path.isdir(ref) and not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "liblensed_ref.so")):
        subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "build_ref.py")], check=False)


_ensure_built()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def compile_ctx():
    """Compile-only context: NVRTC + metadata, no device."""
    import lensed_b200 as L
    ctx = L.Context(device=-1, objects_dir=OBJECTS_DIR)
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def gpu_ctx():
    import lensed_b200 as L
    ctx = L.Context(device=0, objects_dir=OBJECTS_DIR)
    yield ctx
    ctx.close()
