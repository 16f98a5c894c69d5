import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')
GOLDEN_3D = ['g3_a', 'g3_b', 'g3_c', 'g3_d', 'g3_e', 'g3_f', 'g3_g']
GOLDEN_2D = ['g2_a', 'g2_b', 'g2_c', 'g2_d', 'g2_e']


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    d = {k: z[k] for k in z.files}
    sig = d['sig']
    # integer-valued sigmas are passed as ints, exactly as the generating script did
    d['sigmas'] = tuple(int(s) if float(s).is_integer() else float(s) for s in sig)
    return d


@pytest.fixture(scope='session')
def golden():
    return load_golden
