"""The step04 / step08 fixtures (``tests/golden/pca.npz``, ``lines.npz``) are what the UNMODIFIED reference computes:
when the reference is present (``oracle/_ref`` or ``/root/reference``) its ``Compute_GreedyPCA_area`` and
``GridAnalysis`` are run again on the regenerated inputs and compared with the committed files (scipy's ``svds`` starts
ARPACK from a random vector, so equality is to solver precision, not bitwise)."""

import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, load_golden
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason='reference module not present (oracle/_ref)')


@pytest.fixture(scope='module')
def regenerated():
    sys.path.insert(0, GOLDEN)
    try:
        import make_golden
    finally:
        sys.path.remove(GOLDEN)
    return make_golden.next_rows_outputs(ref_loader.load_lib_origin())


def test_pca_fixture_is_the_reference_output(regenerated):
    pca_in, pca, _, _ = regenerated
    g = load_golden('pca')
    np.testing.assert_array_equal(g['cube'], pca_in['cube'])
    np.testing.assert_array_equal(g['areamap'], pca_in['areamap'])
    np.testing.assert_array_equal(g['map_o2'], pca['map_o2'])
    np.testing.assert_array_equal(g['map_o2_itermax3'], pca['map_o2_itermax3'])
    assert int(g['nstop']) == pca['nstop'] == 0 and int(g['nstop_itermax3']) == pca['nstop_itermax3'] >= 1
    assert g['map_o2'].max() >= 5 and (g['test0'] == 0).sum() == 3
    for key in ('faint', 'faint_itermax3'):
        assert np.abs(g[key] - pca[key]).max() <= 1e-10 * np.abs(g[key]).max(), key


def test_lines_fixture_is_the_reference_output(regenerated):
    _, _, lines_in, lines = regenerated
    g = load_golden('lines')
    for key in ('raw', 'var', 'fsf', 'wght', 'dets'):
        np.testing.assert_array_equal(g[key], lines_in[key])
    for tag in ('single_g1_flux', 'single_g1_mse_pcals', 'mosaic_g0_flux', 'mosaic_g1_flux'):
        for key in ('y', 'x', 'z'):
            np.testing.assert_array_equal(g[tag + '_' + key], lines[tag + '_' + key])
        assert np.isfinite(g[tag + '_line']).all()
        assert np.abs(g[tag + '_line'] - lines[tag + '_line']).max() <= 1e-9 * np.abs(g[tag + '_line']).max(), tag
        np.testing.assert_allclose(g[tag + '_lvar'], lines[tag + '_lvar'], rtol=1e-9)
        np.testing.assert_allclose(g[tag + '_flux'], lines[tag + '_flux'], rtol=1e-9)
    # the mosaic cases differ from each other and from the single field: the weights matter
    assert np.abs(g['mosaic_g1_flux_flux'] - g['mosaic_g0_flux_flux']).max() > 1.0
    assert os.path.getsize(os.path.join(GOLDEN, 'lines.npz')) < 2 << 20
