import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    """Fixture written by oracle/pin/make_golden.py from the unmodified reference."""
    z = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    meta = json.loads(str(z['meta']))
    inputs = {k[3:]: z[k] for k in z.files if k.startswith('in_')}
    outputs = {k[4:]: z[k] for k in z.files if k.startswith('out_')}
    return meta, inputs, outputs


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope='session')
def cuda_device():
    from audio_source_separation_b200 import _lib
    lib = _lib.load()   # raises when the library is not built: GPU tests must never pass without it
    return 0
