"""numpy restatement of the Gaussian-fit segmentation (origin_b200/segmap.py).  astropy is not in
this image, so the checks are closed-form: a Gaussian sample returns its own parameters, sigma clipping follows
astropy's documented rule, the morphology equals direct scipy calls.  Runs on the CPU."""

import numpy as np
import pytest
from scipy import ndimage as ndi
from scipy import stats

from origin_b200 import segmap


def test_sigma_clip_follows_the_documented_rule():
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.normal(5.0, 1.0, 5000), [40.0, 55.0, -30.0]])
    kept = segmap.sigma_clip(x, 4.0)
    assert kept.size == 5000 and kept.max() < 11 and kept.min() > -1
    # nothing to reject: one pass, everything kept
    y = rng.normal(0, 1, 100)
    assert segmap.sigma_clip(y, 10.0).size == 100


def test_gaussian_fit_recovers_the_parameters():
    x = np.linspace(-3, 6, 80)
    y = 2.5 * np.exp(-0.5 * ((x - 1.2) / 0.7) ** 2)
    a, m, s = segmap.fit_gaussian(x, y, 2.0, 1.0, 1.0)
    assert (a, m, s) == pytest.approx((2.5, 1.2, 0.7), rel=1e-6)


@pytest.mark.parametrize('pfa', [0.01, 1e-5])
def test_threshold_of_a_gaussian_statistic(pfa):
    rng = np.random.default_rng(3)
    data = rng.normal(10.0, 0.5, (300, 300))
    hist, edges, thr, mea, std = segmap.compute_thresh_gaussfit(data, pfa)
    assert mea == pytest.approx(10.0, abs=0.02) and std == pytest.approx(0.5, rel=0.05)
    assert thr == pytest.approx(10.0 - 0.5 * stats.norm.ppf(pfa), rel=0.02)
    assert hist.size + 1 == edges.size and np.isclose((hist * np.diff(edges)).sum(), 1.0)


def test_segmap_finds_the_sources_and_drops_single_pixels():
    rng = np.random.default_rng(5)
    img = rng.normal(1.0, 0.05, (120, 140))
    img[30:40, 50:62] += 1.0                      # a source
    img[90:97, 20:26] += 0.8                      # another
    img[70, 100] += 5.0                           # an isolated hot pixel: removed by the erosion
    gamma, lab = segmap.compute_segmap_gauss(img, 0.001)
    assert lab.max() == 2 and lab[34, 55] > 0 and lab[93, 22] > 0 and lab[70, 100] == 0
    # same morphology as the reference's two scipy calls
    m = ndi.binary_dilation(ndi.binary_erosion(img > gamma, border_value=1, iterations=1), iterations=1)
    assert np.array_equal(lab > 0, m)
    # with the FSF disc the sources grow by its radius
    _, lab2 = segmap.compute_segmap_gauss(img, 0.001, fwhm_fsf=5)
    assert (lab2 > 0).sum() > (lab > 0).sum() and lab2[29, 55] > 0 and lab[29, 55] == 0


# ---- the reference's own function bodies around three astropy shims ---------------------------------------------

class _Param(float):
    """What the reference touches of an astropy model parameter: arithmetic and ``.value``."""
    value = property(lambda self: float(self))


class _Gaussian1D:
    def __init__(self, amplitude, mean, stddev):
        self.amplitude, self.mean, self.stddev = _Param(amplitude), _Param(mean), _Param(stddev)


class _Fitter:
    def __call__(self, model, x, y):
        a, m, s = segmap.fit_gaussian(x, y, model.amplitude, model.mean, model.stddev)
        return _Gaussian1D(a, m, s)


class _Clipped:
    def __init__(self, values):
        self._values = values

    def compressed(self):
        return self._values


@pytest.fixture
def reference_with_shims(monkeypatch):
    """The UNMODIFIED ``compute_thresh_gaussfit`` / ``compute_segmap_gauss`` of the reference (lib_origin.py:977-1024,
    :243-280) with only astropy's three objects replaced: ``sigma_clip`` / ``Gaussian1D`` / ``LevMarLSQFitter`` by the
    shims above, which delegate to ``segmap.sigma_clip`` / ``segmap.fit_gaussian``.  What stays unpinned is exactly
    the behaviour of those three astropy objects; everything the reference itself writes around them (positive
    values only, histogram, mode and half-maximum estimate, the cut at mean + FWHM / 2, the threshold formula, the
    erosion / dilation / disc convolution / labelling) runs as the reference's own code."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip('reference module not present (oracle/_ref)')
    lib = ref_loader.load_lib_origin()
    monkeypatch.setattr(lib, 'sigma_clip', lambda data, sigma: _Clipped(segmap.sigma_clip(data, sigma)))
    monkeypatch.setattr(lib, 'Gaussian1D', _Gaussian1D)
    monkeypatch.setattr(lib, 'LevMarLSQFitter', _Fitter)
    monkeypatch.setattr(lib, 'gaussian_sigma_to_fwhm', segmap.GAUSSIAN_SIGMA_TO_FWHM)
    return lib


@pytest.mark.parametrize('pfa,bins', [(0.01, 'fd'), (1e-5, 'fd'), (0.01, 40)])
def test_restatement_equals_the_reference_body_around_the_astropy_shims(reference_with_shims, pfa, bins):
    lib = reference_with_shims
    rng = np.random.default_rng(8)
    data = rng.normal(1.0, 0.08, (90, 110))
    data[20:30, 40:52] += 0.9
    data[60:66, 15:21] += 0.6
    data[5, 5] = -0.2                                # non-positive values are dropped (:1000)
    data[45, 70] += 30.0                             # far outlier: sigma clipping at 10 sigma removes it
    rh, re, rthr, rmea, rstd = lib.compute_thresh_gaussfit(data, pfa, bins=bins)
    h, e, thr, mea, std = segmap.compute_thresh_gaussfit(data, pfa, bins=bins)
    np.testing.assert_array_equal(h, rh)
    np.testing.assert_array_equal(e, re)
    assert (thr, mea, std) == pytest.approx((rthr, rmea, rstd), rel=1e-12) and isinstance(rthr, float)
    for fwhm in (0, 5):
        rgamma, rlab = lib.compute_segmap_gauss(data, pfa, fwhm, bins=bins)
        gamma, lab = segmap.compute_segmap_gauss(data, pfa, fwhm, bins=bins)
        assert gamma == pytest.approx(rgamma, rel=1e-12)
        np.testing.assert_array_equal(lab, rlab)
        assert lab.max() >= 2
