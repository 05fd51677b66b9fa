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
