import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session', autouse=True)
def _oracle_built():
    """The C part of the oracle is tiny; build it on demand so that a fresh checkout can run the CPU suite."""
    so = os.path.join(ROOT, 'oracle', 'liboccu_oracle.so')
    if not os.path.exists(so):
        import subprocess
        subprocess.check_call(['make', '-C', os.path.join(ROOT, 'oracle')])
    yield


def golden(name):
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))
