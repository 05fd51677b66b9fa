"""Seeded synthetic MUSE-shaped inputs.

There is no network and the reference's only test cube (``tests/minicube.fits``)
is absent, so every parity test and benchmark runs on cubes made here:

* ``moffat_fsf``     circular Moffat FSF per wavelength, beta 2.8, FWHM linear
                     3.6 -> 2.9 px (the model of ``docs/usage.rst:103-117`` at
                     0.2"/px), P x P, each plane normalised to sum 1.
* ``faint_cube``     step05 input: N(0,1) noise (post-PCA cubes are
                     standardised) plus injected emitters (Moffat x Gaussian
                     line, FWHM 2-12 px).
* ``raw_cube``       step01 input: smooth DCT-representable continuum sources +
                     emitters + noise scaled by a variance cube with sky-line
                     spikes and an exposure map; NaN convention of
                     ``origin.py:262-274`` (raw 0 / var +inf under the mask).
* ``footprint_mask`` all-lambda mask outside a rotated-square footprint plus
                     0.1 % random voxels.
"""

import numpy as np


def moffat_fsf(nz, size=25, beta=2.8, fwhm0=3.6, fwhm1=2.9, dtype=np.float64):
    c = size // 2
    yy, xx = np.mgrid[:size, :size]
    r2 = ((yy - c) ** 2 + (xx - c) ** 2).astype(np.float64)
    fwhm = np.linspace(fwhm0, fwhm1, nz)
    alpha = fwhm / (2.0 * np.sqrt(2.0 ** (1.0 / beta) - 1.0))
    psf = (1.0 + r2[None] / alpha[:, None, None] ** 2) ** (-beta)
    psf /= psf.sum(axis=(1, 2), keepdims=True)
    return psf.astype(dtype)


def _inject_emitters(cube, fsf, n_src, rng, amp_lo, amp_hi):
    nz, ny, nx = cube.shape
    p = fsf.shape[-1]
    c = p // 2
    cat = []
    for _ in range(n_src):
        z0 = int(rng.integers(0, nz))
        y0 = int(rng.integers(0, ny))
        x0 = int(rng.integers(0, nx))
        fw = rng.uniform(2.0, 12.0)
        amp = rng.uniform(amp_lo, amp_hi)
        sig = fw / 2.3548200450309493
        half = int(np.ceil(4 * sig))
        zz = np.arange(max(0, z0 - half), min(nz, z0 + half + 1))
        line = np.exp(-0.5 * ((zz - z0) / sig) ** 2)
        ya, yb = max(0, y0 - c), min(ny, y0 + c + 1)
        xa, xb = max(0, x0 - c), min(nx, x0 + c + 1)
        spat = fsf[z0, ya - y0 + c:yb - y0 + c, xa - x0 + c:xb - x0 + c]
        # amplitude normalised so that the matched filter output is ~amp sigma
        norm = np.sqrt((spat ** 2).sum() * (line ** 2).sum())
        cube[zz[0]:zz[-1] + 1, ya:yb, xa:xb] += (amp / norm) * line[:, None, None] * spat[None]
        cat.append((z0, y0, x0, fw, amp))
    return np.array(cat)


def faint_cube(shape, fsf=None, n_src=None, seed=0, dtype=np.float32, amp=(5.0, 30.0)):
    """Standardised noise cube with injected line emitters (step05 input)."""
    nz, ny, nx = shape
    rng = np.random.default_rng(seed)
    cube = rng.standard_normal(shape, dtype=np.float32)
    if fsf is None:
        fsf = moffat_fsf(nz)
    if n_src is None:
        n_src = max(1, int(round(200 * nz * ny * nx / (3681 * 320 * 320))))
    cat = _inject_emitters(cube, np.asarray(fsf, dtype=np.float64), n_src, rng, *amp)
    return cube.astype(dtype, copy=False), cat


def footprint_mask(shape, seed=0, frac_random=1e-3, border=True):
    """Boolean mask: all-lambda masked spaxels outside a slightly rotated
    square footprint (~5 % of the field) and ``frac_random`` random voxels."""
    nz, ny, nx = shape
    rng = np.random.default_rng(seed + 7919)
    mask = np.zeros(shape, dtype=bool)
    if border:
        yy, xx = np.mgrid[:ny, :nx]
        cy, cx = (ny - 1) / 2, (nx - 1) / 2
        th = np.deg2rad(3.0)
        u = (xx - cx) * np.cos(th) + (yy - cy) * np.sin(th)
        v = -(xx - cx) * np.sin(th) + (yy - cy) * np.cos(th)
        out = (np.abs(u) > 0.49 * nx) | (np.abs(v) > 0.49 * ny)
        mask |= out[None]
    nrand = int(frac_random * nz * ny * nx)
    if nrand:
        idx = rng.integers(0, nz * ny * nx, size=nrand)
        mask.reshape(-1)[idx] = True
    return mask


def raw_cube(shape, fsf=None, n_cont=None, n_src=None, seed=1, mask=None):
    """step01 input ``(cube_raw, var, mask)`` in float64 with the NaN
    convention of ``origin.py:262-274`` already applied."""
    nz, ny, nx = shape
    rng = np.random.default_rng(seed)
    if fsf is None:
        fsf = moffat_fsf(nz)
    if mask is None:
        mask = footprint_mask(shape, seed)
    lam = np.linspace(0.0, 1.0, nz)
    sky = 1.0 + 0.3 * np.sin(6.0 * lam) ** 2
    spikes = rng.choice(nz, size=max(1, nz // 60), replace=False)
    sky[spikes] *= rng.uniform(3.0, 20.0, size=spikes.size)
    yy, xx = np.mgrid[:ny, :nx]
    expo = 1.0 + 0.25 * np.cos(2 * np.pi * yy / max(ny, 2)) * np.cos(2 * np.pi * xx / max(nx, 2))
    var = (sky[:, None, None] / expo[None]).astype(np.float64)
    cube = rng.standard_normal(shape) * np.sqrt(var)
    if n_cont is None:
        n_cont = max(1, int(round(30 * ny * nx / (320 * 320))))
    c = fsf.shape[-1] // 2
    for _ in range(n_cont):
        y0 = int(rng.integers(0, ny))
        x0 = int(rng.integers(0, nx))
        peak = 10 ** rng.uniform(2.0, 3.0)
        a, b, ph = rng.uniform(0.3, 1.0), rng.uniform(-0.5, 0.5), rng.uniform(0, np.pi)
        spec = peak * (a + b * lam + 0.2 * np.cos(3 * np.pi * lam + ph))
        ya, yb = max(0, y0 - c), min(ny, y0 + c + 1)
        xa, xb = max(0, x0 - c), min(nx, x0 + c + 1)
        spat = fsf[:, ya - y0 + c:yb - y0 + c, xa - x0 + c:xb - x0 + c]
        cube[:, ya:yb, xa:xb] += spec[:, None, None] * spat / spat.max()
    if n_src is None:
        n_src = max(1, int(round(200 * nz * ny * nx / (3681 * 320 * 320))))
    _inject_emitters(cube, np.asarray(fsf, dtype=np.float64), n_src, rng, 8.0, 40.0)
    cube[mask] = 0.0
    var[mask] = np.inf
    return cube, var, mask


def field_weights(ny, nx, nfields=2):
    """Overlapping smooth weight maps for the multi-field (mosaic) case
    (``origin.py:600-609``): they sum to 1 where covered, 0 in an uncovered
    corner strip."""
    xx = np.linspace(0.0, 1.0, nx)[None, :] * np.ones((ny, 1))
    w0 = np.clip(1.5 - 2.0 * xx, 0.0, 1.0)
    maps = [w0, 1.0 - w0]
    if nfields > 2:
        raise ValueError('synthetic generator supports 1 or 2 fields')
    return maps[:nfields]
