"""step08 line estimation on the device against the reference's own ``GridAnalysis``
(``lib_origin.py:1620-1790``: ``method_PCA_wgt`` with scipy's ARPACK ``svds``, ``LS_deconv_wgt``, the flux / mse
criteria), run unmodified through ``oracle/ref_loader`` on the padded minicubes ``estimation_line`` builds
(``:1886-1897``; its ``overlap_slices`` comes from astropy, so the padding is restated in the test)."""

import warnings

import numpy as np
import pytest

from oracle import ref_loader
from origin_b200 import synthetic

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_loader.available(), reason='reference module not present (oracle/_ref)')]


@pytest.fixture(scope='module')
def scene():
    shape = (300, 40, 44)
    nz, ny, nx = shape
    fsf = synthetic.moffat_fsf(nz)
    raw, var, mask = synthetic.raw_cube(shape, fsf, n_cont=3, n_src=0, seed=31, mask=np.zeros(shape, dtype=bool))
    rng = np.random.default_rng(4)
    dets = [(150, 20, 22), (60, 3, 40), (240, 38, 2), (100, 12, 30), (200, 25, 10)]     # interior, two corners, ...
    for z0, y0, x0 in dets:                                                              # a line at (and near) each detection
        sig = rng.uniform(1.0, 3.0)
        zz = np.arange(max(0, z0 - 12), min(nz, z0 + 13))
        line = 40.0 * np.exp(-0.5 * ((zz - z0 - rng.integers(-1, 2)) / sig) ** 2)
        ya, yb, xa, xb = max(0, y0 - 12), min(ny, y0 + 13), max(0, x0 - 12), min(nx, x0 + 13)
        raw[zz[0]:zz[-1] + 1, ya:yb, xa:xb] += line[:, None, None] * fsf[z0, ya - y0 + 12:yb - y0 + 12, xa - x0 + 12:xb - x0 + 12] * 20
    var[:, 30:33, 5:8] = np.inf                                                          # a masked patch (origin.py:262-274)
    raw[:, 30:33, 5:8] = 0.0
    cat = dict(z0=np.array([d[0] for d in dets]), y0=np.array([d[1] for d in dets]), x0=np.array([d[2] for d in dets]))
    return raw, var, fsf, cat


def _reference_grid(raw, var, psf, cat, size_grid, criteria, order_dct, horiz_psf, horiz):
    lib = ref_loader.load_lib_origin()
    nz, ny, nx = raw.shape
    P = psf.shape[1]
    side = P + 2 * size_grid
    half = side // 2
    out = []
    for z, y, x in zip(cat['z0'], cat['y0'], cat['x0']):
        red_dat = np.zeros((nz, side, side))
        red_var = np.full((nz, side, side), np.inf)
        ya, yb, xa, xb = max(0, y - half), min(ny, y + half + 1), max(0, x - half), min(nx, x + half + 1)
        red_dat[:, ya - (y - half):yb - (y - half), xa - (x - half):xb - (x - half)] = raw[:, ya:yb, xa:xb]
        red_var[:, ya - (y - half):yb - (y - half), xa - (x - half):xb - (x - half)] = var[:, ya:yb, xa:xb]
        with warnings.catch_warnings(), np.errstate(all='ignore'):
            warnings.simplefilter('ignore')
            out.append(lib.GridAnalysis(red_dat, red_var, psf, None, horiz, size_grid, int(y), int(x), int(z), ny, nx,
                                        horiz_psf, criteria, order_dct))
    return out


@pytest.mark.parametrize('criteria,order_dct', [('flux', 30), ('mse', 30), ('flux', None)])
def test_estimation_line_matches_grid_analysis(scene, criteria, order_dct):
    from origin_b200 import lib_origin
    raw, var, psf, cat = scene
    ref = _reference_grid(raw, var, psf, cat, 1, criteria, order_dct, 1, 5)
    cat2, lin_est, var_est = lib_origin.estimation_line(cat, raw, var, psf, None, None, None, size_grid=1, criteria=criteria,
                                                        order_dct=order_dct, horiz_psf=1, horiz=5)
    assert list(cat2['num_line']) == [1, 2, 3, 4, 5]
    for d, r in enumerate(ref):
        flux, mse5, line, lvar, y, x, z = r
        assert (cat2['y'][d], cat2['x'][d], cat2['z'][d]) == (y, x, z), d
        scale = np.abs(line).max()
        assert np.abs(lin_est[d] - line).max() <= 1e-8 * scale, d
        np.testing.assert_allclose(var_est[d], lvar, rtol=1e-9)
        assert cat2['flux'][d] == pytest.approx(flux, rel=1e-8)
        assert cat2['residual'][d] == pytest.approx(mse5, rel=1e-7)


def test_line_estimates_float32_device_cubes(scene):
    """float32 cubes that already live on the device (the step01 inputs): same estimates to float32 input precision."""
    import torch
    from origin_b200 import lib_origin
    raw, var, psf, cat = scene
    cen = np.stack([cat['y0'], cat['x0']], axis=1)
    a, va = lib_origin.line_estimates(raw, var, psf, cen, 30)
    b, vb = lib_origin.line_estimates(torch.from_numpy(raw.astype(np.float32)).cuda(),
                                      torch.from_numpy(var.astype(np.float32)).cuda(), psf, cen, 30)
    assert np.abs(a - b).max() <= 1e-5 * np.abs(a).max()
    np.testing.assert_allclose(va, vb, rtol=1e-5)


@pytest.mark.parametrize('size_grid', [0, 1])
def test_estimation_line_weighted_mosaic_matches_grid_analysis(size_grid):
    """Weighted mosaic (``wght`` not None, lib_origin.py:1899-1906, :1713-1717): three fields with different FSFs,
    an uncovered strip, a field that only touches one corner; the device combines the fields' FSFs per window
    (``ogn_line_estimates_fields``).  ``size_grid = 0`` is the step's default (``grid_dxy``, steps.py:1082)."""
    import test_lines_host as th
    from origin_b200 import lib_origin
    raw, var, psf, wght, cat = th.make_scene(True)
    ref = th.reference_grid(raw, var, psf, wght, cat, size_grid, 'flux', 20)
    cat2, lin_est, var_est = lib_origin.estimation_line(cat, raw, var, psf, wght, None, None, size_grid=size_grid,
                                                        criteria='flux', order_dct=20, horiz_psf=1, horiz=5)
    th.check(cat2, lin_est, var_est, ref, 1e-8)


def test_line_estimates_fields_equals_single_field_with_unit_weights():
    """One field with weight 1 everywhere through the weighted entry point = the single-field entry point."""
    import test_lines_host as th
    from origin_b200 import lib_origin
    raw, var, psf, _, cat = th.make_scene(False)
    cen = np.stack([cat['y0'], cat['x0']], axis=1)
    a, va = lib_origin.line_estimates(raw, var, psf, cen, 20)
    b, vb = lib_origin.line_estimates(raw, var, psf[None], cen, 20, coef=np.ones((len(cen), 1) + psf.shape[1:]))
    assert np.abs(a - b).max() <= 1e-11 * np.abs(a).max()
    np.testing.assert_allclose(va, vb, rtol=1e-11)
