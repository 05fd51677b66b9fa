"""step04 greedy PCA on the device against the reference's own ``Compute_GreedyPCA_area`` /
``Compute_GreedyPCA`` (``lib_origin.py:769-954``), run unmodified through ``oracle/ref_loader`` (numpy + scipy's
ARPACK ``svds``).  The iteration is steered by threshold decisions, so the comparison is on everything it
produces: the deflated cube (1e-9 of its scale: two eigen-solvers converged to round-off), the per-spaxel
iteration counts (identical) and the number of areas that hit the iteration limit (identical)."""

import warnings

import numpy as np
import pytest

from oracle import origin_oracle as orc
from oracle import ref_loader
from origin_b200 import synthetic

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_loader.available(), reason='reference module not present (oracle/_ref)')]


@pytest.fixture(scope='module')
def lo():
    from origin_b200 import lib_origin
    return lib_origin


@pytest.fixture(scope='module')
def field():
    shape = (400, 48, 56)
    fsf = synthetic.moffat_fsf(shape[0])
    raw, var, mask = synthetic.raw_cube(shape, fsf, n_cont=4, n_src=10, seed=7)
    with np.errstate(all='ignore'):
        cube_std = orc.preprocessing(raw, var, mask, 10, False, 3)['cube_std']
    areamap = np.ones(shape[1:], dtype=int)
    areamap[:, 28:] = 2
    areamap[30:, 10:30] = 3
    lib = ref_loader.load_lib_origin()
    test = [lib.O2test(cube_std[:, areamap == a]) for a in (1, 2, 3)]
    thr = [float(np.percentile(t[t > 0], 85)) for t in test]
    return cube_std, areamap, test, thr


def _reference(cube_std, areamap, test, thr, noise_pop=50, itermax=100):
    lib = ref_loader.load_lib_origin()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        return lib.Compute_GreedyPCA_area(len(thr), cube_std, areamap, noise_pop, thr, itermax, test)


def test_greedy_pca_areas_match_the_reference(lo, field):
    cube_std, areamap, test, thr = field
    rfaint, rmap, rstop = _reference(cube_std, areamap, test, thr)
    assert rmap.max() >= 5 and (test[0] == 0).any()          # several iterations; masked spaxels exercise the index quirk
    faint, map_o2, nstop = lo.Compute_GreedyPCA_area(3, cube_std, areamap, 50, thr, 100, test)
    assert faint.dtype == np.float64 and nstop == rstop
    np.testing.assert_array_equal(map_o2, rmap)
    scale = np.abs(rfaint).max()
    assert np.abs(faint - rfaint).max() <= 1e-9 * scale
    assert np.abs(faint - cube_std).max() > 1e-3 * scale        # something was projected out


def test_greedy_pca_iteration_limit_and_computed_test(lo, field):
    cube_std, areamap, test, thr = field
    rfaint, rmap, rstop = _reference(cube_std, areamap, test, thr, noise_pop=20, itermax=3)
    assert rstop >= 1
    # testO2 = None: the O2 test of every area is computed on the device (Compute_PCA_threshold's first line)
    faint, map_o2, nstop = lo.Compute_GreedyPCA_area(3, cube_std, areamap, 20, thr, 3, None)
    assert nstop == rstop
    np.testing.assert_array_equal(map_o2, rmap)
    assert np.abs(faint - rfaint).max() <= 1e-9 * np.abs(rfaint).max()


def test_greedy_pca_block_and_single_nuisance_spaxel(lo, field):
    cube_std, areamap, test, thr = field
    lib = ref_loader.load_lib_origin()
    block = np.ascontiguousarray(cube_std[:, areamap == 2])
    t = lib.O2test(block)
    for thres in (float(np.percentile(t[t > 0], 90)), float(np.sort(t)[-2] * 0.5 + np.sort(t)[-1] * 0.5)):   # many / one spaxel above
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            rf, rm, rs = lib.Compute_GreedyPCA(block, t, thres, 50, 100)
        f, m, s = lo.Compute_GreedyPCA(block, t, thres, 50, 100)
        assert s == rs
        np.testing.assert_array_equal(m, rm)
        assert np.abs(f - rf).max() <= 1e-9 * np.abs(rf).max()


def test_greedy_pca_device_float32_cube_stays_on_the_device(lo, field):
    import torch
    cube_std, areamap, test, thr = field
    c32 = cube_std.astype(np.float32)
    rfaint, rmap, rstop = _reference(c32.astype(np.float64), areamap, None if False else
                                     [ref_loader.load_lib_origin().O2test(c32.astype(np.float64)[:, areamap == a]) for a in (1, 2, 3)],
                                     thr)
    faint, map_o2, nstop = lo.Compute_GreedyPCA_area(3, torch.from_numpy(c32).cuda(), areamap, 50, thr, 100, None)
    assert torch.is_tensor(faint) and faint.is_cuda and faint.dtype == torch.float32
    np.testing.assert_array_equal(map_o2, rmap)
    got = faint.cpu().numpy()
    assert np.abs(got - rfaint).max() <= 2e-6 * np.abs(rfaint).max()     # float32 storage of the result
