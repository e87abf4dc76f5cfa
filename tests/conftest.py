import os
import sys
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def oracle():
    import orc
    orc.lib()
    return orc


@pytest.fixture(scope='session')
def cp():
    """The product's Python mirror.  New contexts of the test session default to the STRICT math mode (every elementary function of the
    shading stages correctly rounded), in which the device replays the oracle bit for bit -- the path-replay and golden-vector tests
    depend on it.  The product's own default, the fast mode (include/cudapath.h: cudapath_set_math_mode), is selected explicitly by the
    tests that cover it: test_fast_math_*, the converged 4096-spp images, the multi-GPU film test, and __graft_entry__.smoke()."""
    os.environ.setdefault('CUDAPATH_MATH', 'strict')
    import cudapath
    cudapath.lib()
    return cudapath
