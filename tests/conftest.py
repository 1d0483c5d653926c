import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def nn_golden():
    import numpy as np
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'nn_golden.npz'))


@pytest.fixture(scope='session')
def host_golden():
    import numpy as np
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'host_golden.npz'))


@pytest.fixture(scope='session')
def ext_golden():
    import numpy as np
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'ext_golden.npz'))
