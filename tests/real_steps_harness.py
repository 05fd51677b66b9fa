"""Harness shared by tests/test_gpu_real_steps.py (GPU) and tests/test_real_steps_reference.py (CPU): the reference's
own ``muse_origin/steps.py``, executed unmodified through ``oracle/ref_loader.load_steps`` (mpdaf / astropy are stubs,
``Cube`` / ``Image`` are replaced by plain containers), runs steps 01 -> 04 -> 05 -> 06 through ``Step.__call__`` on
the cube of the reference-generated chain fixture."""

from collections import OrderedDict

import numpy as np

from conftest import load_golden, unpack_mask
from oracle import ref_loader
from origin_b200 import dictionaries


class Holder:
    """Stand-in for ``mpdaf.obj.Cube`` / ``Image``: keeps what ``store_cube`` / ``store_image`` hand over."""

    def __init__(self, data=None, **kw):
        self._data = self.data = data
        self.kw = kw

    @property
    def shape(self):
        return self._data.shape


class Orig:
    """The part of ``ORIGIN`` the steps touch: inputs as attributes; the steps' products are looked up like
    ``ORIGIN.__getattr__`` does (origin.py:246-253)."""

    def __init__(self, **inputs):
        self.__dict__.update(inputs)
        self.param, self.steps, self._dataobjs = {}, OrderedDict(), {}
        self.wave = self.wcs = None

    def add(self, step):
        self.steps[step.name] = step
        for name, _ in step._dataobjs:
            self._dataobjs[name] = step
        return step

    def __getattr__(self, name):
        objs = self.__dict__.get('_dataobjs', {})
        if name in objs:
            return getattr(objs[name], name)
        raise AttributeError('unknown attribute %s' % name)


def run_pipeline(monkeypatch, mode):
    """``mode``: 'reference' (nothing patched: the reference's own functions on the CPU), 'rebound'
    (``patch_steps(fused=False)``) or 'fused' (``patch_steps(fused=True)``)."""
    from origin_b200 import lib_origin, steps
    fused = mode == 'fused'
    rsteps = ref_loader.load_steps()
    g = load_golden('chain')
    shape = tuple(int(s) for s in g['shape'])
    mask = unpack_mask(g['mask'], shape)
    monkeypatch.setattr(rsteps, 'Cube', Holder)
    monkeypatch.setattr(rsteps, 'Image', Holder)
    monkeypatch.setattr(rsteps, 'compute_segmap_gauss',
                        lambda img, pfa, fwhm, bins='fd': (0.0, (img > np.percentile(img, 97)).astype(int)))
    monkeypatch.setattr(lib_origin.PurityTable, 'to_astropy', lambda self: self)
    if mode != 'reference':
        steps.patch_steps(rsteps, fused=fused)
    try:
        orig = Orig(cube_raw=g['raw'].astype(np.float64), var=g['var'].astype(np.float64), mask=mask, PSF=g['fsf'],
                    wfields=None, profiles=dictionaries.dico_3fwhm()[0], FWHM_PSF=[3.3], nbAreas=1,
                    areamap=Holder(np.ones(shape[1:], dtype=int)), thresO2=[1e9], testO2=None)      # testO2: set after step01
        param = {}
        pre = orig.add(rsteps.Preprocessing(orig, 1, param))
        orig.add(rsteps.CreateAreas(orig, 2, param)).status = rsteps.Status.RUN           # not run: areamap is given above
        thr_step = orig.add(rsteps.ComputePCAThreshold(orig, 3, param))
        pca = orig.add(rsteps.ComputeGreedyPCA(orig, 4, param))
        tglr = orig.add(rsteps.ComputeTGLR(orig, 5, param))
        pur = orig.add(rsteps.ComputePurityThreshold(orig, 6, param))
        pre()                                                     # Step.__call__: parameters, requirements, status
        assert pre.status is rsteps.Status.RUN and pre.param['dct_order'] == 10
        std64 = np.asarray(orig.cube_std._data, dtype=np.float64)
        o2 = np.mean(std64.reshape(shape[0], -1) ** 2, axis=0)
        if mode == 'reference':                                   # its compute_thresh_gaussfit needs astropy: not run
            thr_step.status = rsteps.Status.RUN
            orig.testO2 = [o2]                                    # what ComputePCAThreshold.run leaves (:617-633)
            fit = None
        else:
            thr_step()                                            # step03: O2 test per area + Gaussian fit of its distribution
            np.testing.assert_allclose(orig.testO2[0], o2, rtol=1e-6)
            fit = (thr_step.thresO2[0], thr_step.meaO2[0], thr_step.stdO2[0])
            assert np.isfinite(fit).all() and fit[0] > fit[1] > 0 and len(orig.histO2[0]) + 1 == len(orig.binO2[0])
        pca()
        if fused:
            assert isinstance(pca.__dict__['cube_faint'], steps.LazyProduct) and pca.__dict__['cube_faint'].on_device().is_cuda
        tglr(pcut=1e-8)
        assert tglr.param['pcut'] == 1e-8 and tglr.status is rsteps.Status.RUN
        if fused:                                                  # step05 read cube_faint where it was
            assert isinstance(pca.__dict__['cube_faint'], steps.LazyProduct)
            assert isinstance(tglr.__dict__['cube_local_min'], steps.LazyProduct)
        pre.segmap_merged = Holder(g['segmap'])                   # step06 with the fixture's segmap
        monkeypatch.setattr(rsteps, 'compute_segmap_gauss', lambda img, pfa, fwhm, bins='fd': (0.0, np.zeros_like(g['segmap'])))
        pur(purity=0.8)
        out = dict(cube_std=np.asarray(orig.cube_std._data), cube_faint=np.asarray(orig.cube_faint._data),
                   cube_correl=np.asarray(orig.cube_correl._data), maxmap=np.asarray(orig.maxmap._data),
                   minmap=np.asarray(orig.minmap._data), profile=np.asarray(orig.cube_profile._data),
                   local_max=np.asarray(orig.cube_local_max._data), local_min=np.asarray(orig.cube_local_min._data),
                   std_local_max=np.asarray(orig.cube_std_local_max._data), mapO2=np.asarray(orig.mapO2._data),
                   threshold=orig.param['threshold'], threshold_std=orig.param['threshold_std'],
                   det_M=np.asarray(pur.Pval['Det_M']), std_det_M=np.asarray(pur.Pval_comp['Det_M']), pca_fit=fit)
        return g, out
    finally:
        if mode != 'reference':
            steps.unpatch_steps()


def close_thr(a, b):
    return (np.isinf(a) and np.isinf(b)) or abs(a - b) <= 1e-3 * abs(b)




def check_against_fixture(g, out):
    z, y, x = g['cat_z'], g['cat_y'], g['cat_x']
    np.testing.assert_allclose(out['cube_std'][100], g['cube_std_plane'], rtol=2e-4, atol=2e-4)
    np.testing.assert_array_equal(np.asarray(out['cube_faint'], dtype=np.float64), np.asarray(out['cube_std'], dtype=np.float64))
    assert not out['mapO2'].any()                                                         # threshold 1e9: nothing projected
    np.testing.assert_allclose(out['cube_correl'][100], g['correl_plane'], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(out['maxmap'], g['maxmap'], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(out['minmap'], g['minmap'], rtol=2e-4, atol=2e-4)
    assert abs(np.count_nonzero(out['local_max']) - int(g['n_local_max'])) <= 3
    assert abs(np.count_nonzero(out['local_min']) - int(g['n_local_min'])) <= 3
    np.testing.assert_allclose(out['local_max'][z, y, x], g['cat_tglr'], rtol=2e-4)      # the reference's detections
    np.testing.assert_array_equal(out['profile'][z, y, x], g['cat_profile'])
    np.testing.assert_allclose(out['std_local_max'][g['std_z'], g['std_y'], g['std_x']], g['std_val'], rtol=2e-4)
    assert np.abs(out['det_M'] - g['tab_Det_M']).max() <= 1 and np.abs(out['std_det_M'] - g['tabstd_Det_M']).max() <= 1
    assert close_thr(out['threshold'], float(g['thr'])) and close_thr(out['threshold_std'], float(g['thr_std']))
