import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu')


@pytest.fixture(scope='session')
def native_so():
    from ptina_b200 import build
    return build.build()


@pytest.fixture(scope='session')
def gpu(native_so):
    """Process-wide native context behind the PTina-style facade (one per test session, like the reference's singletons)."""
    import torch
    if not torch.cuda.is_available():
        pytest.fail('a -m gpu test ran without a CUDA device; ptina_b200 has no CPU fallback')
    from ptina_b200 import worker, _native
    worker.init()
    yield _native.context()
    from ptina_b200 import things
    things.shutdown()
