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


# --------------------------------------------------------------------------
# Counter-based generator: any window of a large cube, on the host or on the device
# --------------------------------------------------------------------------
# The benchmark cube (3681 x 320 x 320, or the 3681 x 900 x 900 mosaic) is generated on the GPU by
# every rank, while the CPU reference arm and the spot oracle need only a tile of the SAME cube on the
# host.  Each voxel's value is therefore a pure function of (seed, global linear index): a 64-bit
# integer mix (wrapping int64 arithmetic, identical in numpy and torch) feeding a float64 Box-Muller
# transform.  The integer stage is bit-identical everywhere; the float64 log / cos may differ in the
# last place between libm and CUDA, which survives the rounding to float32 for about one voxel in 1e9.

_M1 = -7046029254386353131      # 0x9E3779B97F4A7C15 as a signed 64-bit integer
_M2 = -4658895280553007687      # 0xBF58476D1CE4E5B9
_M3 = -7723592293110705685      # 0x94D049BB133111EB


def _mix64(h, xp):
    """splitmix64 finaliser on int64 arrays / tensors with logical shifts emulated by masks."""
    h = h ^ ((h >> 30) & ((1 << 34) - 1))
    h = h * _M2
    h = h ^ ((h >> 27) & ((1 << 37) - 1))
    h = h * _M3
    h = h ^ ((h >> 31) & ((1 << 33) - 1))
    return h


def _window_index(window, gshape, xp, device=None):
    (z0, z1), (y0, y1), (x0, x1) = window
    _, gny, gnx = gshape
    if xp is np:
        z = np.arange(z0, z1, dtype=np.int64)[:, None, None]
        y = np.arange(y0, y1, dtype=np.int64)[None, :, None]
        x = np.arange(x0, x1, dtype=np.int64)[None, None, :]
    else:
        z = xp.arange(z0, z1, dtype=xp.int64, device=device)[:, None, None]
        y = xp.arange(y0, y1, dtype=xp.int64, device=device)[None, :, None]
        x = xp.arange(x0, x1, dtype=xp.int64, device=device)[None, None, :]
    return (z * gny + y) * gnx + x


def hashed_uniform(seed, window, gshape, xp=np, device=None):
    """float64 uniform(0, 1) per voxel of ``window = ((z0,z1),(y0,y1),(x0,x1))`` of a ``gshape`` cube."""
    idx = _window_index(window, gshape, xp, device)
    with np.errstate(over='ignore'):
        h = _mix64(idx * _M1 + (int(seed) * 2 + 1) * 0x632BE5AB, xp)
    u = ((h >> 11) & ((1 << 52) - 1))
    u = u.astype(np.float64) if xp is np else u.to(xp.float64)
    return (u + 0.5) * (1.0 / (1 << 52))


def hashed_normal(seed, window, gshape, xp=np, device=None):
    """float32 N(0, 1) per voxel of a window of a ``gshape`` cube (Box-Muller on two hashed uniforms)."""
    u1 = hashed_uniform(2 * int(seed), window, gshape, xp, device)
    u2 = hashed_uniform(2 * int(seed) + 1, window, gshape, xp, device)
    if xp is np:
        return (np.sqrt(-2.0 * np.log(u1)) * np.cos((2.0 * np.pi) * u2)).astype(np.float32)
    return (xp.sqrt(-2.0 * xp.log(u1)) * xp.cos((2.0 * np.pi) * u2)).to(xp.float32)


def emitter_catalogue(gshape, n_src=None, seed=0):
    """Injected line emitters of the benchmark cube: rows ``(z0, y0, x0, sigma_z, amplitude)``."""
    nz, ny, nx = gshape
    if n_src is None:
        n_src = max(1, int(round(200 * nz * ny * nx / (3681 * 320 * 320))))
    rng = np.random.default_rng(seed)
    cat = np.empty((n_src, 5))
    cat[:, 0] = rng.integers(20, max(21, nz - 20), n_src)
    cat[:, 1] = rng.integers(13, max(14, ny - 13), n_src)
    cat[:, 2] = rng.integers(13, max(14, nx - 13), n_src)
    cat[:, 3] = rng.uniform(2.0, 12.0, n_src) / 2.3548200450309493
    cat[:, 4] = rng.uniform(5.0, 30.0, n_src)
    return cat


def _emitter_patches(cat, fsf, window, gshape):
    """Yield ``(zslice, yslice, xslice, patch)`` (window coordinates, float32) of the emitters touching ``window``."""
    nz = gshape[0]
    (z0w, z1w), (y0w, y1w), (x0w, x1w) = window
    c = fsf.shape[-1] // 2
    for z0, y0, x0, sig, amp in cat:
        z0, y0, x0 = int(z0), int(y0), int(x0)
        hw = int(np.ceil(4 * sig))
        za, zb = max(0, z0 - hw), min(nz, z0 + hw + 1)
        if zb <= z0w or za >= z1w or y0 + c + 1 <= y0w or y0 - c >= y1w or x0 + c + 1 <= x0w or x0 - c >= x1w:
            continue
        zz = np.arange(za, zb)
        line = np.exp(-0.5 * ((zz - z0) / sig) ** 2)
        spat = np.asarray(fsf[z0], dtype=np.float64)
        patch = (amp / np.sqrt((spat ** 2).sum() * (line ** 2).sum())) * line[:, None, None] * spat[None]
        # clip the (zz, y0-c.., x0-c..) box to the window
        a = [max(za, z0w), max(y0 - c, y0w), max(x0 - c, x0w)]
        b = [min(zb, z1w), min(y0 + c + 1, y1w), min(x0 + c + 1, x1w)]
        sub = patch[a[0] - za:b[0] - za, a[1] - (y0 - c):b[1] - (y0 - c), a[2] - (x0 - c):b[2] - (x0 - c)]
        yield (slice(a[0] - z0w, b[0] - z0w), slice(a[1] - y0w, b[1] - y0w), slice(a[2] - x0w, b[2] - x0w),
               sub.astype(np.float32))


def _footprint(gshape, window):
    _, ny, nx = gshape
    (_, _), (y0, y1), (x0, x1) = window
    yy, xx = np.mgrid[y0:y1, x0:x1]
    cy, cx, th = (ny - 1) / 2, (nx - 1) / 2, np.deg2rad(3.0)
    u = (xx - cx) * np.cos(th) + (yy - cy) * np.sin(th)
    v = -(xx - cx) * np.sin(th) + (yy - cy) * np.cos(th)
    return (np.abs(u) > 0.49 * nx) | (np.abs(v) > 0.49 * ny)


def bench_window(gshape, window=None, fsf=None, seed=0, xp=np, device=None, zchunk=128, frac_random=1e-3):
    """``(cube float32, mask uint8)`` of a window of the benchmark's step05 input: hashed N(0,1) noise
    plus the emitters of :func:`emitter_catalogue`; mask = rotated-square footprint (~5 % of the spaxels)
    plus ``frac_random`` hashed voxels.  ``xp=np`` builds it on the host, ``xp=torch`` on ``device``
    (in chunks of ``zchunk`` planes); both give the same cube."""
    nz, ny, nx = gshape
    if window is None:
        window = ((0, nz), (0, ny), (0, nx))
    (z0, z1), (y0, y1), (x0, x1) = window
    if fsf is None:
        fsf = moffat_fsf(nz)
    shape = (z1 - z0, y1 - y0, x1 - x0)
    foot = _footprint(gshape, window)
    if xp is np:
        cube = np.empty(shape, dtype=np.float32)
        mask = np.empty(shape, dtype=np.uint8)
    else:
        cube = xp.empty(shape, dtype=xp.float32, device=device)
        mask = xp.empty(shape, dtype=xp.uint8, device=device)
        foot = xp.from_numpy(foot).to(device)
    for za in range(z0, z1, zchunk):
        zb = min(z1, za + zchunk)
        sub = ((za, zb), (y0, y1), (x0, x1))
        cube[za - z0:zb - z0] = hashed_normal(seed, sub, gshape, xp, device)
        rnd = hashed_uniform(1000003 + int(seed), sub, gshape, xp, device) < frac_random
        m = rnd | foot[None]
        mask[za - z0:zb - z0] = m.astype(np.uint8) if xp is np else m.to(xp.uint8)
    for zs, ys, xs, patch in _emitter_patches(emitter_catalogue(gshape, None, seed), fsf, window, gshape):
        if xp is np:
            cube[zs, ys, xs] += patch
        else:
            cube[zs, ys, xs] += xp.from_numpy(patch).to(device)
    return cube, mask
