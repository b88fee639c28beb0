import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib

    return oracle_lib.oracle()


@pytest.fixture(scope="session")
def ref():
    import oracle_lib

    r = oracle_lib.ref()
    if r is None:
        pytest.skip("oracle/_ref/libalpaka_ref.so not available (needs /root/reference to build)")
    return r


@pytest.fixture(scope="session")
def gpu():
    """Platform/device/queue on cuda:0 through the C ABI. Fails (not skips) if the CUDA library is missing."""
    import alpaka_b200 as ab

    platform = ab.Platform()
    if platform.get_dev_count() < 1:
        pytest.fail("-m gpu tests need a CUDA device; there is no CPU fallback")
    dev = platform.get_dev_by_idx(0)
    queue = ab.Queue(dev, blocking=False)
    yield ab, dev, queue
    queue.wait()
