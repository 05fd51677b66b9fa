"""Host-side logic that needs no GPU: the step06 tail against the reference's own function on the edge
cases it handles specially, and the signatures of the drop-in layer against the reference's step layer."""

import inspect

import numpy as np
import pytest

from oracle import ref_loader
from test_tiles_dist import NumpyCounter, _fake_extrema

needs_ref = pytest.mark.skipif(not ref_loader.available(), reason='reference module not present (oracle/_ref)')


def _lists(lmax, lmin):
    from origin_b200 import lib_origin
    mi, ni = np.flatnonzero(lmax), np.flatnonzero(lmin)
    return lib_origin.LocalExtrema(lmax.shape, mi, lmax.reshape(-1)[mi], ni, lmin.reshape(-1)[ni])


@needs_ref
@pytest.mark.parametrize('case', ['default', 'segmap', 'negative', 'negative_segmap', 'decreasing_default', 'unsorted'])
def test_threshold_purity_tail_matches_the_reference(case):
    """``Compute_threshold_purity`` (lib_origin.py:1391-1479) on dense cubes vs our list-based tail, including
    thresholds below zero (zeros of the dense cubes count, background-only with a segmap, :1429) and a
    decreasing default list (threshmax < threshmin: the '> threshmin' prefilter of :1443 decides)."""
    import warnings
    from origin_b200 import lib_origin
    lib = ref_loader.load_lib_origin()
    lmax, lmin, seg = _fake_extrema(11, (10, 14, 12))
    lmax, lmin = lmax.astype(np.float64), lmin.astype(np.float64)
    segmap, thr = None, None
    if case in ('segmap', 'negative_segmap'):
        segmap = seg
    if case.startswith('negative'):
        thr = np.linspace(-2.0, 6.0, 17)
    if case == 'unsorted':
        thr = np.array([5.0, 1.0, 3.0, 2.0, 8.0])
    if case == 'decreasing_default':
        lmin = lmin * 0.05                       # max(lmin) far below 1.1 * median(max_z lmax)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        rthr, rtab = lib.Compute_threshold_purity(0.6, lmax.copy(), lmin.copy(), segmap, thr)
        thr_, tab = lib_origin.Compute_threshold_purity(0.6, _lists(lmax.astype(np.float32), lmin.astype(np.float32)),
                                                        None, segmap, thr, _backend=NumpyCounter())
    np.testing.assert_allclose(tab['Tval_r'], np.asarray(rtab['Tval_r']), rtol=1e-6)
    np.testing.assert_array_equal(tab['Det_M'], np.asarray(rtab['Det_M']))
    np.testing.assert_array_equal(tab['Det_m'], np.asarray(rtab['Det_m']))
    assert (np.isinf(thr_) and np.isinf(rthr)) or thr_ == pytest.approx(rthr, rel=1e-6)


def test_check_counts_flags_the_overflow_mark():
    from origin_b200 import lib_origin
    lib_origin.check_counts(np.array([3, 0]), np.array([1, 0]))
    with pytest.raises(OverflowError):
        lib_origin.check_counts(np.array([3, 0]), np.array([-(1 << 56) + 5, 2]))


@pytest.mark.skipif(ref_loader.find_reference_file('muse_origin/steps.py') is None,
                    reason='reference steps.py only exists in the build container')
def test_drop_in_signatures_equal_the_reference():
    """The fused ``run`` methods patched into the reference's step classes and the function mirror keep the
    reference's signatures (steps.py:420-429, 756, 851-859; lib_origin.py:150, 1070, 1220, 1391)."""
    from origin_b200 import lib_origin, steps as osteps
    rsteps = ref_loader.load_steps()
    lib = ref_loader.load_lib_origin()

    def params(fn):
        return [(p.name, p.default) for p in inspect.signature(fn).parameters.values()]

    assert params(osteps._run_preprocessing) == params(rsteps.Preprocessing.run)
    assert params(osteps._run_compute_tglr) == params(rsteps.ComputeTGLR.run)
    assert params(osteps._run_purity) == params(rsteps.ComputePurityThreshold.run)
    for name in ('dct_residual', 'Correlation_GLR_test', 'compute_local_max', 'Compute_threshold_purity', 'O2test',
                 'DCTMAT'):
        ref = params(getattr(lib, name))
        ours = params(getattr(lib_origin, name))
        assert ours[:len(ref)] == ref, name           # ours may append keyword-only extras (ctx, out_dtype)
    # what the step layer imports by name (steps.py:19-41) is what patch_steps rebinds
    for name in ('dct_residual', 'compute_local_max', 'Correlation_GLR_test', 'Compute_threshold_purity', 'O2test'):
        assert hasattr(rsteps, name)
