import os
import sys

import pytest

# Shards that share ONE device (tests/test_sharding.py on a one-GPU box) wait for each other inside their solver kernels.
# With CUDA's default lazy module loading the first launch of any not-yet-loaded kernel synchronises the whole context --
# behind the neighbour's spinning kernel: a dead-lock until the bounded waits give up.  Eager loading (set before CUDA
# initialises) loads every kernel up front.  One GPU per shard, the deployment case, does not need this.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

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
