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
    # never test a stale library: rebuild in-tree when a source changed since the last build
    try:
        from vittles_b200 import build as vb
        if not vb.is_current():
            vb.build(verbose=False)
    except Exception as exc:      # no nvcc: the tests that need the library will say so
        sys.stderr.write('vittles_b200: could not (re)build the CUDA library: {}\n'.format(exc))


@pytest.fixture(scope='session')
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + '.npz'))
    return load


def assert_close(actual, desired, rtol=1e-8, atol_scale=1e-12, what=''):
    """The parity bar of BASELINE.json's north_star: rtol 1e-8 in float64,
    elementwise, with an absolute floor of atol_scale * max|desired| for entries
    that are small only through cancellation."""
    import torch
    if isinstance(actual, torch.Tensor):
        actual = actual.detach().cpu().numpy()
    if isinstance(desired, torch.Tensor):
        desired = desired.detach().cpu().numpy()
    actual, desired = np.asarray(actual), np.asarray(desired)
    assert actual.shape == desired.shape, '{}: shape {} vs {}'.format(what, actual.shape, desired.shape)
    scale = float(np.max(np.abs(desired))) if desired.size else 0.0
    np.testing.assert_allclose(actual, desired, rtol=rtol, atol=atol_scale * scale, err_msg=what)
