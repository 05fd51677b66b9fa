"""Spectral line-profile dictionaries.

The reference ships two FITS dictionaries next to its code and reads them in
``ORIGIN.profiles`` / ``ORIGIN.FWHM_profiles`` (``muse_origin/origin.py:515-533``):
``Dico_FWHM_2_12.fits`` (20 profiles) and ``Dico_3FWHM.fits`` (3 profiles, the
default, ``origin.py:222-223``).  Every profile is a 201-sample L2-normalised
Gaussian, ``g(x) = exp(-x^2 / 2 sigma^2) / ||g||``, ``x = -100..100``,
``sigma = FWHM / (2 sqrt(2 ln 2))`` with ``FWHM_k = 2 + 10 k / 19`` pixels;
``Dico_3FWHM`` holds entries 0, 9 and 19.  They are regenerated analytically
here (``tests/golden/dictionaries.npz``, dumped from the shipped FITS files, pins
this to 1e-15) and user dictionaries are read with the numpy-only FITS parser.
"""

import numpy as np

from .fitsmini import read_hdus

_FWHM_TO_SIGMA = 2.0 * np.sqrt(2.0 * np.log(2.0))
PROFILE_SIZE = 201


def gaussian_profile(fwhm, size=PROFILE_SIZE):
    x = np.arange(size, dtype=np.float64) - size // 2
    sigma = fwhm / _FWHM_TO_SIGMA
    g = np.exp(-x * x / (2.0 * sigma * sigma))
    return g / np.linalg.norm(g)


def dico_fwhm_2_12():
    """The 20 profiles of ``Dico_FWHM_2_12.fits`` and their FWHMs."""
    fwhm = np.linspace(2.0, 12.0, 20)
    return [gaussian_profile(f) for f in fwhm], list(fwhm)


def dico_3fwhm():
    """The 3 profiles of ``Dico_3FWHM.fits`` (entries 0, 9, 19 of 2_12)."""
    profs, fwhm = dico_fwhm_2_12()
    keep = (0, 9, 19)
    return [profs[k] for k in keep], [fwhm[k] for k in keep]


def get_dictionary(name):
    name = str(name)
    if name in ('Dico_3FWHM', 'Dico_3FWHM.fits', '3FWHM'):
        return dico_3fwhm()
    if name in ('Dico_FWHM_2_12', 'Dico_FWHM_2_12.fits', 'FWHM_2_12', '2_12'):
        return dico_fwhm_2_12()
    return read_dictionary(name)


def read_dictionary(path):
    """Read a dictionary FITS file the way ``origin.py:520-533`` does."""
    hdus = read_hdus(path)[1:]
    profiles = [np.asarray(d, dtype=np.float64).ravel() for _, d in hdus]
    if len({p.shape[0] for p in profiles}) != 1:
        raise ValueError('The profiles must have the same size')
    fwhm = [h.get('FWHM') for h, _ in hdus]
    return profiles, fwhm
