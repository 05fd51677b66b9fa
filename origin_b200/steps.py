"""Step-level mirror of the reference's hot-path steps and the drop-in hook.

The reference's step layer (``muse_origin/steps.py``) imports the numerical
functions *by name* (``steps.py:19-41``) and calls them from four ``run``
methods: ``Preprocessing.run`` (:420-489), ``ComputeTGLR.run`` (:756-802),
``ComputePurityThreshold.run`` (:851-892) and ``Detection.run`` (:941-974 for
the thresholding block).  Two levels of drop-in are offered:

``patch_steps(fused=False)``
    rebinds ``dct_residual``, ``compute_local_max``, ``Correlation_GLR_test``,
    ``Compute_threshold_purity`` and ``O2test`` in the ``muse_origin.steps``
    namespace to the B200 implementations of :mod:`origin_b200.lib_origin`;
    the reference's ``run`` methods stay untouched.

``patch_steps(fused=True)`` (default)
    additionally replaces the ``run`` methods of steps 01, 05 and 06 by the
    fused versions below, which keep intermediates on the device, carry the
    local extrema as compact lists (:class:`~origin_b200.lib_origin.LocalExtrema`)
    and materialise the dense cubes the step API promises
    (``cube_local_max`` ...) from them.  ``Detection.run`` is left as it is
    (its two ``np.where`` blocks then read those dense cubes); INTEGRATION.md
    shows the six-line change that makes it read the lists through
    :func:`detection_cat0` instead.

The array parts are also available as plain functions (``preprocessing``,
``compute_tglr``, ``compute_purity_threshold``, ``detection_cat0``) taking
numpy arrays; they are what the parity tests and the multi-GPU driver call.
Everything outside the hot path (segmentation maps, merging, WCS, tables)
remains the reference's code and is only *called* from here.
"""

import numpy as np

from . import lib_origin as lo

__all__ = ['preprocessing', 'compute_tglr', 'compute_purity_threshold', 'detection_cat0', 'patch_steps',
           'LazyDense']


# --------------------------------------------------------------------------
# array-level step functions
# --------------------------------------------------------------------------

def preprocessing(cube_raw, var, mask, dct_order=10, dct_approx=False, local_max_size=3, allreduce=None,
                  owned=None, ctx=None):
    """Array part of ``Preprocessing.run`` (reference steps.py:430-465 plus the
    two segmentation inputs of :472 and :480).

    Returns ``cube_std, cont_dct`` (float32 cubes), ``ima_std, ima_dct``
    (images), ``extrema_std`` (local maxima of ``cube_std`` / ``-cube_std``),
    ``cont_sumsq`` (argument of the log10 at :472) and ``o2map`` (:480).
    """
    out = lo.preprocess(cube_raw, var, mask, dct_order, dct_approx, allreduce=allreduce, owned=owned, ctx=ctx)
    ext, _, _ = lo.local_extrema(out['cube_std'], out['cube_std'], mask, local_max_size, ctx=ctx)
    out['extrema_std'] = ext
    return out


def compute_tglr(cube_faint, fsf, wfields, profiles, mask, size=3, pcut=1e-8, pmeansub=True, ctx=None, **kw):
    """Array part of ``ComputeTGLR.run`` (reference steps.py:768-802):
    ``cube_correl, cube_correl_min, cube_profile, maxmap, minmap`` and the
    local extrema as ``extrema``."""
    res = lo.step05(cube_faint, fsf, wfields, profiles, mask, size, pcut, pmeansub, ctx=ctx, **kw)
    return dict(cube_correl=res['correl'], cube_correl_min=res['correl_min'], cube_profile=res['profile'],
                maxmap=res['maxmap'], minmap=res['minmap'], extrema=res['extrema'])


def compute_purity_threshold(extrema, extrema_std, segmap, purity=0.9, purity_std=None, threshlist=None,
                             allreduce=None, ctx=None):
    """Array part of ``ComputePurityThreshold.run`` (reference steps.py:860-892)
    given the purity segmap: ``(threshold, Pval, threshold_std, Pval_comp)``."""
    if purity_std is None:
        purity_std = purity
    thr, pval = lo.Compute_threshold_purity(purity, extrema, None, segmap, threshlist, allreduce=allreduce, ctx=ctx)
    thr_std, pval_comp = lo.Compute_threshold_purity(purity_std, extrema_std, None, None, threshlist,
                                                     allreduce=allreduce, ctx=ctx)
    return thr, pval, thr_std, pval_comp


def detection_cat0(extrema, cube_profile, threshold, extrema_std, threshold_std, ctx=None):
    """Rows of ``Cat0`` before formatting (reference steps.py:956-974): the
    correl detections (``comp`` 0) followed by the std-cube detections
    (``comp`` 1), each in C order.  Returns a dict of equal-length columns
    ``x0, y0, z0, comp, STD, T_GLR, profile``."""
    a = lo.threshold_rows(extrema, threshold, cube_profile, 'max', ctx)
    b = lo.threshold_rows(extrema_std, threshold_std, None, 'max', ctx)
    na, nb = len(a['x0']), len(b['x0'])
    return dict(
        x0=np.concatenate([a['x0'], b['x0']]), y0=np.concatenate([a['y0'], b['y0']]),
        z0=np.concatenate([a['z0'], b['z0']]),
        comp=np.concatenate([np.zeros(na, int), np.ones(nb, int)]),
        STD=np.concatenate([np.full(na, np.nan), b['value'].astype(np.float64)]),
        T_GLR=np.concatenate([a['value'].astype(np.float64), np.full(nb, np.nan)]),
        profile=np.concatenate([np.asarray(a['profile'], dtype=np.int64), np.zeros(nb, np.int64)]),
    )


# --------------------------------------------------------------------------
# drop-in for the reference's step objects
# --------------------------------------------------------------------------

class LazyDense:
    """Dense ``cube_local_max`` / ``cube_local_min`` materialised on first use.

    ``mpdaf.obj.Cube(data=LazyDense(...))`` is not possible (mpdaf copies into a
    masked array), so the fused steps store real dense arrays when the step API
    is asked for them; this class is the container used by the array-level API
    and by ``Detection`` when only the lists are needed."""

    def __init__(self, extrema, which):
        self.extrema, self.which, self._dense = extrema, which, None
        self.shape = extrema.shape

    def __array__(self, dtype=None, copy=None):
        if self._dense is None:
            self._dense = self.extrema.dense(self.which)
        return self._dense if dtype is None else self._dense.astype(dtype)


def _run_preprocessing(self, orig, dct_order=10, dct_approx=False, pfasegcont=0.01, pfasegres=0.01,
                       local_max_size=3, bins='fd'):
    """Fused ``Preprocessing.run`` (reference steps.py:420-489)."""
    mod = _STEPS_MODULE
    self._loginfo('DCT computation (B200)')
    out = preprocessing(orig.cube_raw, orig.var, orig.mask, dct_order, dct_approx, local_max_size)
    self._loginfo('Std signal saved in self.cube_std and self.ima_std')
    self.store_cube('cube_std', out['cube_std'])
    self.store_image('ima_std', out['ima_std'])
    ext = out['extrema_std']
    self._ogn_extrema_std = ext
    self.store_cube('cube_std_local_max', ext.dense('max'))
    self.store_cube('cube_std_local_min', ext.dense('min'))
    self._loginfo('DCT continuum saved in self.cont_dct and self.ima_dct')
    self.store_cube('cont_dct', out['cont_dct'])
    self.store_image('ima_dct', out['ima_dct'])
    # segmentation stays the reference's code (steps.py:467-489); only its two
    # 2-D inputs come from the device pass
    mean_fwhm = int(np.ceil(np.mean(self.orig.FWHM_PSF)))
    self._loginfo('Segmentation based on the continuum')
    map1 = np.log10(out['cont_sumsq'])
    thresh, map_cont = mod.compute_segmap_gauss(map1, pfasegcont, mean_fwhm, bins=bins)
    self.store_image('segmap_cont', map_cont)
    self._loginfo('Segmentation based on the residual')
    thresh, map_res = mod.compute_segmap_gauss(out['o2map'], pfasegres, mean_fwhm, bins=bins)
    segmap, nlabels = mod.ndi.label((map_cont > 0) | (map_res > 0))
    self.store_image('segmap_merged', segmap)


def _run_compute_tglr(self, orig, size=3, ncpu=1, pcut=1e-8, pmeansub=True):
    """Fused ``ComputeTGLR.run`` (reference steps.py:756-802); ``ncpu`` is ignored."""
    self._loginfo('Correlation (B200)')
    out = compute_tglr(orig.cube_faint._data, orig.PSF, orig.wfields, orig.profiles, orig.mask, size, pcut, pmeansub)
    self.store_cube('cube_correl', out['cube_correl'])
    self.store_cube('cube_correl_min', out['cube_correl_min'])
    self.store_cube('cube_profile', out['cube_profile'])
    self.store_image('maxmap', out['maxmap'])
    self.store_image('minmap', out['minmap'])
    ext = out['extrema']
    self._ogn_extrema = ext
    self.store_cube('cube_local_max', ext.dense('max'))
    self.store_cube('cube_local_min', ext.dense('min'))


def _run_purity(self, orig, purity=0.9, purity_std=None, threshlist=None, pfasegfinal=1e-5, bins='fd'):
    """Fused ``ComputePurityThreshold.run`` (reference steps.py:851-892)."""
    mod = _STEPS_MODULE
    if purity_std is None:
        purity_std = purity
    orig.param.update(dict(purity=purity, purity_std=purity_std))
    thresh, map_res = mod.compute_segmap_gauss(self.orig.maxmap._data, pfasegfinal, 0, bins=bins)
    segmap, nlabels = mod.ndi.label((map_res > 0) | (orig.segmap_merged._data > 0))
    self.store_image('segmap_purity', segmap)
    tglr_step = orig.steps['compute_TGLR']
    prep_step = orig.steps['preprocessing']
    ext = getattr(tglr_step, '_ogn_extrema', None)
    ext_std = getattr(prep_step, '_ogn_extrema_std', None)
    if ext is None:        # session reloaded from disk: rebuild the lists from the dense cubes
        ext = lo._as_extrema(orig.cube_local_max._data, orig.cube_local_min._data)
    if ext_std is None:
        ext_std = lo._as_extrema(orig.cube_std_local_max._data, orig.cube_std_local_min._data)
    thr, pval, thr_std, pval_comp = compute_purity_threshold(ext, ext_std, segmap, purity, purity_std, threshlist)
    self.Pval = pval.to_astropy()
    orig.param['threshold'] = thr
    self._loginfo('Threshold: %.2f ', thr)
    self.Pval_comp = pval_comp.to_astropy()
    orig.param['threshold_std'] = thr_std
    self._loginfo('Threshold: %.2f ', thr_std)


_STEPS_MODULE = None
_ORIGINALS = {}


def patch_steps(steps_module=None, fused=True):
    """Route the reference's hot path through libogn.  ``steps_module`` is
    ``muse_origin.steps`` (imported here when omitted).  Returns the dict of the
    replaced attributes so that :func:`unpatch_steps` can restore them."""
    global _STEPS_MODULE
    if steps_module is None:
        import muse_origin.steps as steps_module
    _STEPS_MODULE = steps_module
    names = {
        'dct_residual': lo.dct_residual,
        'compute_local_max': lo.compute_local_max,
        'Correlation_GLR_test': lo.Correlation_GLR_test,
        'Compute_threshold_purity': _threshold_purity_astropy,
        'O2test': lo.O2test,
    }
    for name, fn in names.items():
        _ORIGINALS.setdefault(name, getattr(steps_module, name, None))
        setattr(steps_module, name, fn)
    if fused:
        for cls_name, run in (('Preprocessing', _run_preprocessing), ('ComputeTGLR', _run_compute_tglr),
                              ('ComputePurityThreshold', _run_purity)):
            cls = getattr(steps_module, cls_name)
            _ORIGINALS.setdefault(cls_name + '.run', cls.run)
            cls.run = run
    return dict(_ORIGINALS)


def unpatch_steps():
    mod = _STEPS_MODULE
    if mod is None:
        return
    for key, val in _ORIGINALS.items():
        if '.' in key:
            cls_name, attr = key.split('.')
            setattr(getattr(mod, cls_name), attr, val)
        elif val is not None:
            setattr(mod, key, val)
    _ORIGINALS.clear()


def _threshold_purity_astropy(purity, cube_local_max, cube_local_min, segmap=None, threshlist=None):
    """``Compute_threshold_purity`` with the reference's return types
    (float, astropy Table)."""
    thr, tab = lo.Compute_threshold_purity(purity, cube_local_max, cube_local_min, segmap, threshlist)
    try:
        return thr, tab.to_astropy()
    except ImportError:
        return thr, tab
