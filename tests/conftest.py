import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + '.npz')) as f:
        return {k: f[k] for k in f.files}


def unpack_mask(packed, shape):
    shape = tuple(int(s) for s in shape)
    n = int(np.prod(shape))
    return np.unpackbits(packed)[:n].reshape(shape).astype(bool)


@pytest.fixture(scope='session')
def golden():
    return load_golden


def tol_report(got, ref, rtol=1e-5):
    """The parity rule of SURVEY.md §8d: |d| <= rtol * max(|ref|, rms(ref))."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    rms = float(np.sqrt(np.mean(ref ** 2))) if ref.size else 0.0
    bound = rtol * np.maximum(np.abs(ref), rms)
    err = np.abs(got - ref)
    return dict(ok=bool(np.all(err <= bound)), max_abs=float(err.max()) if err.size else 0.0,
                rms=rms, worst=float((err / np.maximum(bound, 1e-300)).max()) if err.size else 0.0)
