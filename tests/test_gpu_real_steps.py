"""The drop-in hook on the reference's OWN step layer (see tests/real_steps_harness.py): the reference's
``Step.__call__`` / ``DataObj`` / ``store_cube`` machinery driving

* ``patch_steps(fused=False)``: the reference's own ``run`` bodies (steps.py:420-489, 681-704, 756-802, 851-892)
  calling the B200 functions rebound in the module's namespace, and
* ``patch_steps(fused=True)``: the fused ``run`` methods with their lazy, device-backed products,

on the cube of the reference-generated chain fixture (``tests/golden/chain.npz``).  Both must reproduce the
fixture (step04 runs with a threshold nothing exceeds, as the fixture has no PCA step) and each other."""

import numpy as np
import pytest

from oracle import ref_loader
from real_steps_harness import check_against_fixture, close_thr, run_pipeline

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(ref_loader.find_reference_file('muse_origin/steps.py') is None,
                                 reason='reference steps.py not present (oracle/_ref)')]


@pytest.mark.parametrize('mode', ['rebound', 'fused'])
def test_reference_step_classes_with_the_b200_path(monkeypatch, mode):
    g, out = run_pipeline(monkeypatch, mode)
    check_against_fixture(g, out)


def test_fused_and_rebound_steps_agree(monkeypatch):
    _, a = run_pipeline(monkeypatch, 'rebound')
    _, b = run_pipeline(monkeypatch, 'fused')
    for key in ('cube_std', 'cube_correl', 'maxmap', 'minmap'):
        scale = np.abs(a[key]).max()
        assert np.abs(np.asarray(a[key], dtype=np.float64) - b[key]).max() <= 2e-5 * scale, key
    assert np.mean(a['profile'] == b['profile']) > 0.9999
    assert close_thr(a['threshold'], b['threshold']) and np.abs(a['det_M'] - b['det_M']).max() <= 2
    np.testing.assert_allclose(a['pca_fit'], b['pca_fit'], rtol=1e-4)     # step03: O2 map of step01 vs the cube reduced again
