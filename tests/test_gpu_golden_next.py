"""step04 greedy PCA and step08 line estimation on the device against the committed golden fixtures
(``tests/golden/pca.npz``, ``lines.npz``: outputs of the unmodified reference's ``Compute_GreedyPCA_area`` and
``GridAnalysis``, written by ``tests/golden/make_golden.py next``).  Unlike ``test_gpu_pca.py`` / ``test_gpu_lines.py``
these need no copy of the reference at test time."""

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('variant', ['default', 'itermax3'])
def test_greedy_pca_golden(variant):
    from origin_b200 import lib_origin
    g = load_golden('pca')
    cube, areamap = g['cube'].astype(np.float64), g['areamap'].astype(int)
    test = [g['test0'], g['test1']]
    if variant == 'default':
        faint, map_o2, nstop = lib_origin.Compute_GreedyPCA_area(2, cube, areamap, 50, list(g['thr']), 100, test)
        rfaint, rmap, rstop = g['faint'], g['map_o2'], int(g['nstop'])
    else:       # the O2 test computed on the device, a smaller background population, the iteration limit hit
        faint, map_o2, nstop = lib_origin.Compute_GreedyPCA_area(2, cube, areamap, 20, list(g['thr']), 3, None)
        rfaint, rmap, rstop = g['faint_itermax3'], g['map_o2_itermax3'], int(g['nstop_itermax3'])
    assert faint.dtype == np.float64 and nstop == rstop
    np.testing.assert_array_equal(map_o2, rmap)
    assert np.abs(faint - rfaint).max() <= 1e-9 * np.abs(rfaint).max()
    assert np.abs(faint - cube).max() > 1e-2                      # something was projected out


def test_greedy_pca_golden_float32_on_the_device():
    import torch
    from origin_b200 import lib_origin
    g = load_golden('pca')
    faint, map_o2, nstop = lib_origin.Compute_GreedyPCA_area(2, torch.from_numpy(g['cube']).cuda(), g['areamap'].astype(int), 50,
                                                             list(g['thr']), 100, None)
    assert faint.is_cuda and faint.dtype == torch.float32 and nstop == int(g['nstop'])
    np.testing.assert_array_equal(map_o2, g['map_o2'])        # the float32 cube is exact in float64: same decisions
    assert np.abs(faint.cpu().numpy() - g['faint']).max() <= 2e-6 * np.abs(g['faint']).max()


@pytest.mark.parametrize('tag,mosaic,size_grid,criteria,order_dct', [
    ('single_g1_flux', False, 1, 'flux', 20), ('single_g1_mse_pcals', False, 1, 'mse', None),
    ('mosaic_g0_flux', True, 0, 'flux', 20), ('mosaic_g1_flux', True, 1, 'flux', 20)])
def test_estimation_line_golden(tag, mosaic, size_grid, criteria, order_dct):
    from origin_b200 import lib_origin
    g = load_golden('lines')
    raw, var = g['raw'].astype(np.float64), g['var'].astype(np.float64)
    cat = dict(z0=g['dets'][:, 0], y0=g['dets'][:, 1], x0=g['dets'][:, 2])
    psf, wght = (list(g['fsf']), list(g['wght'])) if mosaic else (g['fsf'][0], None)
    cat2, lin_est, var_est = lib_origin.estimation_line(cat, raw, var, psf, wght, None, None, size_grid=size_grid,
                                                        criteria=criteria, order_dct=order_dct, horiz_psf=1, horiz=5)
    for d in range(len(g['dets'])):
        assert (cat2['y'][d], cat2['x'][d], cat2['z'][d]) == (g[tag + '_y'][d], g[tag + '_x'][d], g[tag + '_z'][d]), d
        line = g[tag + '_line'][d]
        assert np.abs(lin_est[d] - line).max() <= 1e-8 * np.abs(line).max(), d
        np.testing.assert_allclose(var_est[d], g[tag + '_lvar'][d], rtol=1e-8)
        assert cat2['flux'][d] == pytest.approx(g[tag + '_flux'][d], rel=1e-8)
        assert cat2['residual'][d] == pytest.approx(g[tag + '_mse'][d], rel=1e-7)
