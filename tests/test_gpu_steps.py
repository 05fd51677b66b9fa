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

    def store_cube(self, name, data, **kw):
        setattr(self, name, FakeData(data))
        setattr(self.orig, name, FakeData(data))

    def store_image(self, name, data, **kw):
        setattr(self, name, FakeData(data))
        setattr(self.orig, name, FakeData(data))


def fake_steps_module():
    mod = types.ModuleType('fake_muse_origin_steps')
    mod.ndi = ndi
    mod.compute_segmap_gauss = lambda img, pfa, fwhm, bins='fd': (0.0, (img > np.percentile(img, 97)).astype(int))
    for name in ('dct_residual', 'compute_local_max', 'Correlation_GLR_test', 'Compute_threshold_purity', 'O2test'):
        setattr(mod, name, None)
    for cls in ('Preprocessing', 'ComputeTGLR', 'ComputePurityThreshold'):
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
        orig = types.SimpleNamespace(cube_raw=g['raw'].astype(np.float64), var=g['var'].astype(np.float64), mask=mask,
                                     PSF=g['fsf'], wfields=None, profiles=dictionaries.dico_3fwhm()[0],
                                     FWHM_PSF=[3.3], param={}, steps={})
        pre, tglr, pur = mod.Preprocessing(orig), mod.ComputeTGLR(orig), mod.ComputePurityThreshold(orig)
        orig.steps = {'preprocessing': pre, 'compute_TGLR': tglr}
        pre.run(orig)
        np.testing.assert_allclose(orig.cube_std._data[100], g['cube_std_plane'], rtol=2e-4, atol=2e-4)
        orig.cube_faint = FakeData(orig.cube_std._data)          # steps 02-04 (PCA) are out of scope
        tglr.run(orig, pcut=1e-8)
        np.testing.assert_allclose(orig.cube_correl._data[100], g['correl_plane'], rtol=2e-4, atol=2e-4)
        np.testing.assert_allclose(orig.maxmap._data, g['maxmap'], rtol=2e-4, atol=2e-4)
        assert abs(np.count_nonzero(orig.cube_local_max._data) - int(g['n_local_max'])) <= 3
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
