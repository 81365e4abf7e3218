import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "box2d-mt_b200", "python"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"),
          ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def gpu():
    """The product library with a live device; GPU tests fail loudly (never skip) when it is missing."""
    import b2cuda
    n = b2cuda.device_count()
    assert n > 0, "no CUDA device visible: the -m gpu tests must run on the GPU box (there is no CPU fallback)"
    return b2cuda
