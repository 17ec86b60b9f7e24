import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# turbopy_b200 is a plugin for turboPy: put the UNMODIFIED reference on sys.path before anything imports the
# package (build container: /root/reference; GPU box: the staged copy under oracle/_ref), so that the tools
# register into the reference's own registry and the tests drive them through it.
from oracle import refload  # noqa: E402

if refload.available():
    refload.load()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session", autouse=True)
def built_library():
    """The C-ABI library is built in-tree (python -m turbopy_b200.build); make sure
    it exists and matches the sources before any test loads it (no-op when fresh)."""
    from turbopy_b200 import build
    try:
        build.build()
    except RuntimeError as exc:                      # no nvcc: tests that need the library will say so
        print(f"[conftest] could not build libturbopy_b200.so: {exc}")
    yield
