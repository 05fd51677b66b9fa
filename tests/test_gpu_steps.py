"""The drop-in hook: ``patch_steps`` on a stand-in for ``muse_origin.steps`` (mpdaf / astropy are
not installed here), driving the fused ``run`` methods with stub ORIGIN / Step objects and checking the
stored products against the reference-generated chain fixture."""

import types

import numpy as np
import pytest
from scipy import ndimage as ndi

from conftest import load_golden, unpack_mask
from origin_b200 import dictionaries

pytestmark = pytest.mark.gpu


class FakeData:
    def __init__(self, data):
        self._data = data


class FakeStep:
    def __init__(self, orig):
        self.orig = orig
        self.log = []

    def _loginfo(self, *a):
        self.log.append(a)

    def _logwarning(self, *a):
        self.log.append(a)

    # like the reference's Step.store_cube / store_image (steps.py:284-299): the product becomes an attribute of
    # the STEP; the ORIGIN object finds it there (origin.py:246-253)
    def store_cube(self, name, data, **kw):
        setattr(self, name, FakeData(data))

    def store_image(self, name, data, **kw):
        setattr(self, name, FakeData(data))


class FakeOrigin:
    """The part of ``ORIGIN`` the fused steps touch: inputs as attributes, products looked up on the steps."""

    def __init__(self, **inputs):
        self.__dict__.update(inputs)
        self.param, self.steps = {}, {}

    def __getattr__(self, name):
        for step in self.__dict__.get('steps', {}).values():
            if name in step.__dict__:
                return getattr(step, name)
        raise AttributeError(name)


def fake_steps_module():
    mod = types.ModuleType('fake_muse_origin_steps')
    mod.ndi = ndi
    mod.compute_segmap_gauss = lambda img, pfa, fwhm, bins='fd': (0.0, (img > np.percentile(img, 97)).astype(int))
    for name in ('dct_residual', 'compute_local_max', 'Correlation_GLR_test', 'Compute_threshold_purity', 'O2test'):
        setattr(mod, name, None)
    for cls in ('Preprocessing', 'ComputeGreedyPCA', 'ComputeTGLR', 'ComputePurityThreshold'):
        setattr(mod, cls, type(cls, (FakeStep,), {'run': lambda self, orig: None}))
    return mod


def test_patch_steps_runs_the_fused_steps():
    from origin_b200 import lib_origin, steps
    g = load_golden('chain')
    shape = tuple(int(s) for s in g['shape'])
    mask = unpack_mask(g['mask'], shape)
    mod = fake_steps_module()
    steps.patch_steps(mod, fused=True)
    try:
        assert mod.Correlation_GLR_test is lib_origin.Correlation_GLR_test
        orig = FakeOrigin(cube_raw=g['raw'].astype(np.float64), var=g['var'].astype(np.float64), mask=mask,
                          PSF=g['fsf'], wfields=None, profiles=dictionaries.dico_3fwhm()[0], FWHM_PSF=[3.3])
        pre, tglr, pur = mod.Preprocessing(orig), mod.ComputeTGLR(orig), mod.ComputePurityThreshold(orig)
        orig.steps = {'preprocessing': pre, 'compute_TGLR': tglr, 'purity': pur}
        pre.run(orig)
        np.testing.assert_allclose(orig.cube_std._data[100], g['cube_std_plane'], rtol=2e-4, atol=2e-4)
        orig.cube_faint = FakeData(orig.cube_std._data)          # steps 02-04 (PCA) are out of scope
        tglr.run(orig, pcut=1e-8)
        np.testing.assert_allclose(orig.cube_correl._data[100], g['correl_plane'], rtol=2e-4, atol=2e-4)
        np.testing.assert_allclose(orig.maxmap._data, g['maxmap'], rtol=2e-4, atol=2e-4)
        # the dense extremum cubes are placeholders until somebody reads them, then real step products
        assert isinstance(tglr.__dict__['cube_local_max'], steps.LazyProduct)
        assert tglr.__dict__['cube_local_max'].shape == shape
        assert abs(np.count_nonzero(orig.cube_local_max._data) - int(g['n_local_max'])) <= 3
        assert isinstance(tglr.__dict__['cube_local_max'], FakeData)
        assert isinstance(tglr.__dict__['cube_local_min'], steps.LazyProduct)       # nobody asked for this one
        # step06 with the fixture's segmap instead of the gaussian-fit one
        orig.segmap_merged = FakeData(g['segmap'])
        mod.compute_segmap_gauss = lambda img, pfa, fwhm, bins='fd': (0.0, np.zeros_like(g['segmap']))
        pur.to_astropy = None
        lib_origin.PurityTable.to_astropy = lambda self: self
        pur.run(orig, purity=0.8)
        assert np.abs(np.asarray(pur.Pval['Det_M']) - g['tab_Det_M']).max() <= 1
        ref_thr = float(g['thr'])
        assert (np.isinf(ref_thr) and np.isinf(orig.param['threshold'])) or \
            abs(orig.param['threshold'] - ref_thr) <= 1e-3 * abs(ref_thr)
        # step07 rows
        cat0 = steps.detection_cat0(tglr._ogn_extrema, orig.cube_profile._data, float(g['use_thr']),
                                    pre._ogn_extrema_std, float(g['use_std']))
        n = len(g['cat_z'])
        np.testing.assert_array_equal(cat0['z0'][:n], g['cat_z'])
        np.testing.assert_array_equal(cat0['x0'][:n], g['cat_x'])
        np.testing.assert_array_equal(cat0['profile'][:n], g['cat_profile'])
        np.testing.assert_array_equal(cat0['z0'][n:], g['std_z'])
        assert np.all(cat0['comp'][:n] == 0) and np.all(cat0['comp'][n:] == 1)
    finally:
        steps.unpatch_steps()


def test_fused_tglr_keeps_unread_products_on_the_device():
    """float32 host cube, width a multiple of 8: the fused step uploads the mask bit-packed, returns correl,
    the maps and the lists, and leaves correl_min / profile on the GPU behind LazyProduct placeholders that
    materialise (float64 / uint8, the reference's container types) on first access and equal the eager run."""
    import torch
    from origin_b200 import lib_origin, steps, synthetic
    shape = (120, 128, 64)
    fsf = synthetic.moffat_fsf(shape[0])
    cube, _ = synthetic.faint_cube(shape, fsf, n_src=6, seed=5)
    mask = synthetic.footprint_mask(shape, seed=5)
    profs = dictionaries.dico_3fwhm()[0]
    mod = fake_steps_module()
    steps.patch_steps(mod, fused=True)
    try:
        orig = FakeOrigin(mask=mask, PSF=fsf, wfields=None, profiles=profs, cube_faint=FakeData(cube))
        tglr = mod.ComputeTGLR(orig)
        orig.steps = {'compute_TGLR': tglr}
        assert steps.pack_mask(orig).size == (cube.size + 7) // 8       # a session that re-runs the step packs once
        tglr.run(orig, pcut=1e-8)
        assert tglr._ogn_used_mask_bits
        lazy = tglr.__dict__['cube_correl_min']
        assert isinstance(lazy, steps.LazyProduct) and torch.is_tensor(tglr._ogn_profile) and tglr._ogn_profile.is_cuda
        ref = lib_origin.step05(cube, fsf, None, profs, mask, 3, 1e-8, True)
        np.testing.assert_array_equal(orig.cube_correl._data, ref['correl'])
        got = orig.cube_correl_min._data                       # materialises
        assert got.dtype == np.float64 and isinstance(tglr.__dict__['cube_correl_min'], FakeData)
        np.testing.assert_array_equal(got.astype(np.float32), ref['correl_min'])
        np.testing.assert_array_equal(orig.cube_profile._data, ref['profile'])
        np.testing.assert_array_equal(tglr._ogn_extrema.max_index, ref['extrema'].max_index)
        np.testing.assert_array_equal(orig.maxmap._data, ref['maxmap'])
    finally:
        steps.unpatch_steps()


def test_fused_greedy_pca_hands_cube_faint_to_step05_on_the_device():
    """Fused step04 -> step05 (steps.py:681-704, :756-802): ``cube_faint`` is a LazyProduct backed by a CUDA
    tensor, the fused ComputeTGLR reads it there (nothing is uploaded, the placeholder survives), and the
    products equal the host route (``Compute_GreedyPCA_area`` on numpy arrays, then ``step05`` on its result)."""
    import torch
    from origin_b200 import lib_origin, steps, synthetic
    shape = (200, 48, 64)
    fsf = synthetic.moffat_fsf(shape[0])
    cube_std, _ = synthetic.faint_cube(shape, fsf, n_src=8, seed=11)
    cube_std = cube_std.astype(np.float32)
    rng = np.random.default_rng(3)
    spec = np.sin(np.arange(shape[0]) / 9.0)[:, None, None]
    cube_std[:, 5:12, 8:20] += 3.0 * spec * rng.uniform(0.5, 1.5, size=(1, 7, 12)).astype(np.float32)   # a continuum residual
    mask = synthetic.footprint_mask(shape, seed=2)
    profs = dictionaries.dico_3fwhm()[0]
    areamap = np.ones(shape[1:], dtype=int)
    areamap[:, 32:] = 2
    o2 = [np.mean(cube_std[:, areamap == a].astype(np.float64) ** 2, axis=0) for a in (1, 2)]
    thr = [float(np.percentile(t, 90)) for t in o2]
    mod = fake_steps_module()
    steps.patch_steps(mod, fused=True)
    try:
        assert mod.Compute_GreedyPCA_area is lib_origin.Compute_GreedyPCA_area
        orig = FakeOrigin(mask=mask, PSF=fsf, wfields=None, profiles=profs, cube_std=FakeData(cube_std),
                          areamap=FakeData(areamap), nbAreas=2, thresO2=thr, testO2=None)
        pca, tglr = mod.ComputeGreedyPCA(orig), mod.ComputeTGLR(orig)
        orig.steps = {'compute_greedy_PCA': pca, 'compute_TGLR': tglr}
        pca.run(orig)
        lazy = pca.__dict__['cube_faint']
        assert isinstance(lazy, steps.LazyProduct) and lazy.on_device().is_cuda and orig.param['threshold_list'] == thr
        rfaint, rmap, rstop = lib_origin.Compute_GreedyPCA_area(2, cube_std, areamap, 50, thr, 100, None)
        assert rmap.max() >= 1
        np.testing.assert_array_equal(pca.mapO2._data, rmap)
        dev_faint = lazy.on_device().cpu().numpy()
        assert dev_faint.dtype == np.float32 and np.abs(dev_faint - rfaint).max() <= 1e-5 * np.abs(rfaint).max()
        tglr.run(orig, pcut=1e-8)
        assert isinstance(pca.__dict__['cube_faint'], steps.LazyProduct)         # step05 read the device tensor
        ref = lib_origin.step05(dev_faint, fsf, None, profs, mask, 3, 1e-8, True)
        assert isinstance(orig.cube_correl._data, np.ndarray)
        np.testing.assert_allclose(orig.cube_correl._data, ref['correl'], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(orig.maxmap._data, ref['maxmap'], rtol=1e-6, atol=1e-6)
        ext = tglr._ogn_extrema
        np.testing.assert_array_equal(ext._host(ext.max_index), ref['extrema'].max_index)
        np.testing.assert_array_equal(ext._host(ext.min_value), ref['extrema'].min_value)
        assert isinstance(tglr.__dict__['cube_profile'], steps.LazyProduct)
        np.testing.assert_array_equal(orig.cube_profile._data, ref['profile'])
        got = orig.cube_faint._data                                                # materialises: float64, as the reference
        assert got.dtype == np.float64 and isinstance(pca.__dict__['cube_faint'], FakeData)
        np.testing.assert_array_equal(got.astype(np.float32), dev_faint)
    finally:
        steps.unpatch_steps()
        assert not hasattr(mod, 'estimation_line')
