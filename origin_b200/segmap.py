"""Gaussian-statistics segmentation maps: numpy / scipy restatement of the reference's
``compute_thresh_gaussfit`` (``lib_origin.py:977-1024``) and ``compute_segmap_gauss`` (``:243-280``).

The reference builds these 2-D maps on the host from three images the device pass already emits —
``log10(sum_z cont_dct^2)`` (``steps.py:472``), ``O2test(cube_std)`` (``:480``) and ``maxmap`` (``:866``) — but
through ``astropy.stats.sigma_clip`` and ``astropy.modeling`` (``Gaussian1D`` + ``LevMarLSQFitter``), which are
not a dependency of the ported path.  This module restates those two building blocks from their published
definitions so that step01's ``segmap_merged`` and step06's ``segmap_purity`` can be produced without astropy:

* ``sigma_clip``: astropy's defaults — centre = median, spread = standard deviation, at most 5 iterations,
  values with ``|x - median| > sigma * std`` rejected, stop when nothing is rejected;
* the Gaussian fit: Levenberg-Marquardt (MINPACK ``lmder`` through ``scipy.optimize.leastsq``, which is what
  ``LevMarLSQFitter`` calls) on ``amplitude * exp(-(x - mean)^2 / (2 stddev^2))`` with the analytic Jacobian,
  ``xtol = 1e-7`` and at most 100 function evaluations (astropy's ``acc`` / ``maxiter`` defaults; it leaves ``ftol``
  and ``gtol`` at scipy's own defaults, and so does this).

Parity: PARTLY PINNED.  Everything the reference itself writes in those two functions (positive values only, the
histogram, the mode / half-maximum start values, the cut at mean + FWHM / 2, the threshold formula, erosion,
dilation, disc convolution, labelling) is pinned: ``tests/test_segmap.py`` runs the reference's UNMODIFIED function
bodies with only astropy's three objects (``sigma_clip``, ``Gaussian1D``, ``LevMarLSQFitter``) replaced by shims
that delegate to the two building blocks above, and the results are identical.  What stays UNPINNED — astropy is
absent from this image — is exactly the behaviour of those three objects (clipping rule and defaults, the
Levenberg-Marquardt settings); they are checked against closed-form expectations only (a Gaussian sample gives
back its mean / sigma / threshold).  The step mirror prefers the reference's functions whenever ``muse_origin``
really imports.
"""

import numpy as np
from scipy import ndimage as ndi
from scipy import optimize, signal, stats

__all__ = ['sigma_clip', 'fit_gaussian', 'compute_thresh_gaussfit', 'compute_segmap_gauss']

GAUSSIAN_SIGMA_TO_FWHM = 2.0 * np.sqrt(2.0 * np.log(2.0))


def sigma_clip(data, sigma=3.0, maxiters=5):
    """Values of ``data`` surviving astropy's default sigma clipping, as a 1-D array
    (``sigma_clip(data, sigma).compressed()`` in the reference, ``lib_origin.py:1001-1002``)."""
    x = np.asarray(data, dtype=np.float64).ravel()
    x = x[np.isfinite(x)]
    for _ in range(maxiters):
        if x.size == 0:
            break
        centre, spread = np.median(x), np.std(x)
        keep = np.abs(x - centre) <= sigma * spread
        if keep.all():
            break
        x = x[keep]
    return x


def _gauss(p, x):
    return p[0] * np.exp(-0.5 * ((x - p[1]) / p[2]) ** 2)


def fit_gaussian(x, y, amplitude, mean, stddev, acc=1e-7, maxiter=100):
    """Least-squares ``(amplitude, mean, stddev)`` of a 1-D Gaussian, started from the given values."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)

    def resid(p):
        return _gauss(p, x) - y

    def jac(p):                           # rows = parameters (col_deriv=True), as astropy's fit_deriv
        e = np.exp(-0.5 * ((x - p[1]) / p[2]) ** 2)
        return np.array([e, p[0] * e * (x - p[1]) / p[2] ** 2, p[0] * e * (x - p[1]) ** 2 / p[2] ** 3])

    if x.size < 3:
        return float(amplitude), float(mean), float(stddev)
    p, _ = optimize.leastsq(resid, [amplitude, mean, stddev], Dfun=jac, col_deriv=True, xtol=acc, maxfev=maxiter)
    return float(p[0]), float(p[1]), float(abs(p[2]))


def compute_thresh_gaussfit(data, pfa, bins='fd', sigclip=10):
    """Threshold of a test statistic from a Gaussian fit of its distribution (``lib_origin.py:977-1024``).
    Returns ``(histO2, frecO2, thresO2, mea, std)`` like the reference."""
    data = np.asarray(data, dtype=np.float64)
    data = data[data > 0]                                            # :1000
    data = sigma_clip(data, sigclip)                                 # :1001-1002
    hist, edges = np.histogram(data, bins=bins, density=True)        # :1003
    ind = int(np.argmax(hist))
    mod = edges[ind]
    ind2 = int(np.argmin((hist[ind] / 2 - hist[:ind]) ** 2)) if ind > 0 else 0   # :1006 (argmin of an empty slice raises there)
    fwhm = mod - edges[ind2]
    sigma = fwhm / np.sqrt(2 * np.log(2))                            # :1008
    coef = stats.norm.ppf(pfa)
    x = (edges[1:] + edges[:-1]) / 2                                 # :1014
    xcut = mod + GAUSSIAN_SIGMA_TO_FWHM * sigma / 2                  # :1017
    ksel = x < xcut
    _, mea, std = fit_gaussian(x[ksel], hist[ksel], hist.max(), mod, sigma if sigma > 0 else np.std(data))
    return hist, edges, float(mea - std * coef), mea, std            # :1022


def compute_segmap_gauss(data, pfa, fwhm_fsf=0, bins='fd'):
    """Segmentation map of an image from Gaussian statistics (``lib_origin.py:243-280``): threshold, erosion and
    dilation to drop isolated pixels, optional convolution with a disc of the FSF's half-width, labelling.
    Returns ``(gamma, labelled image)``."""
    data = np.asarray(data, dtype=np.float64)
    _, _, gamma, _, _ = compute_thresh_gaussfit(data, pfa, bins=bins)
    mask = data > gamma
    mask = ndi.binary_erosion(mask, border_value=1, iterations=1)    # :267
    mask = ndi.binary_dilation(mask, iterations=1)                   # :268
    if fwhm_fsf > 0:
        half = int(fwhm_fsf) // 2
        size = half * 2 + 1
        disc = np.hypot(*list(np.mgrid[:size, :size] - half)) < half
        mask = signal.fftconvolve(mask, disc, mode='same') > 1e-9    # :275-276
    return gamma, ndi.label(mask)[0]
