"""CPU float64 oracle for ORIGIN's detection hot path.  TEST INFRASTRUCTURE ONLY.

This module is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  Nothing under ``origin_b200/`` does.

It restates, in plain numpy/scipy float64, the algorithm of the reference
(``musevlt/origin``, mounted read-only at ``/root/reference`` in the build
container) for the path

    step01 dct_residual + standardisation  ->  step05 Correlation_GLR_test
    -> compute_local_max -> step06 Compute_threshold_purity counts
    -> step07 thresholding into Cat0 rows.

Each function cites the reference lines it follows.  The reference itself is
pure Python on top of numpy / scipy (pocketfft, ndimage, LAPACK — third-party,
unpinned in the reference's ``setup.cfg:29-38``; numpy 2.3.5 / scipy 1.18.1 in
this image), so the restatement uses the same published building blocks
(zero-padded FFT linear convolution, separable maximum filter) but is written
independently, batched instead of looped, and has direct-space twins
(``*_direct``) that share no code with the FFT route.

Parity pin: PINNED.  The reference's own tests hold no array-level golden
vectors for this path (only summary pins on the missing
``tests/minicube.fits``).  The oracle is therefore pinned against outputs of the
*unmodified* reference functions run in the build container
(``tests/golden/make_golden.py`` loads ``/root/reference/muse_origin/lib_origin.py``
by path and dumps its inputs/outputs to ``tests/golden/*.npz``);
``tests/test_oracle_golden.py`` checks every function here against those
fixtures.
"""

import numpy as np
from scipy import fft as sfft

__all__ = [
    'dctmat', 'dct_residual', 'preprocessing', 'prepare_profiles',
    'fsf_correlate', 'fsf_correlate_direct', 'spectral_glr', 'spectral_glr_direct',
    'correlation_glr_test', 'compute_local_max', 'tglr_step',
    'threshold_purity', 'detection_rows', 'o2test',
]


# --------------------------------------------------------------------------
# step01: DCT continuum
# --------------------------------------------------------------------------

def dctmat(nl, order):
    """DCT-II synthesis matrix, ``nl x (order+1)`` (lib_origin.py:127-146).

    ``D0[z, j] = sqrt(2/nl) cos((z + 1/2) pi j / nl)``, column 0 divided by
    sqrt(2); columns are orthonormal.
    """
    z = np.arange(nl, dtype=np.float64)[:, None]
    j = np.arange(order + 1, dtype=np.float64)[None, :]
    d0 = np.sqrt(2 / nl) * np.cos((z + 0.5) * (np.pi / nl) * j)
    d0[:, 0] *= 1 / np.sqrt(2)
    return d0


def dct_residual(w_raw, order, var, approx, mask, chunk=4096):
    """Continuum estimated on ``order+1`` DCT atoms (lib_origin.py:150-240).

    approx (``:176-194``): ``cont = D0 D0^T s`` for every spaxel.
    default (``:203-238``): spaxels without any masked voxel
    (``valid = ~any(mask, axis=0)``, ``:226``) get the weighted LS projection
    ``D0 (D0^T W D0)^-1 D0^T W s`` with ``W = diag(1/var)`` (``:233-235``), the
    others the unweighted one (``:237``).  Returns the continuum only.
    """
    w_raw = np.asarray(w_raw, dtype=np.float64)
    nl = w_raw.shape[0]
    d0 = dctmat(nl, order)
    spec = w_raw.reshape(nl, -1)
    nspec = spec.shape[1]
    cont = np.empty_like(spec)
    if approx:
        valid = np.zeros(nspec, dtype=bool)
    else:
        valid = ~np.any(np.asarray(mask).reshape(nl, -1), axis=0)
        var2 = np.asarray(var, dtype=np.float64).reshape(nl, -1)
    plain = np.flatnonzero(~valid)
    for i in range(0, plain.size, chunk):
        sel = plain[i:i + chunk]
        cont[:, sel] = d0 @ (d0.T @ spec[:, sel])
    wls = np.flatnonzero(valid)
    for i in range(0, wls.size, chunk):
        sel = wls[i:i + chunk]
        w = 1.0 / var2[:, sel]                                   # (nl, n)
        gram = np.einsum('zi,zn,zj->nij', d0, w, d0, optimize=True)
        rhs = np.einsum('zi,zn->ni', d0, spec[:, sel] * w)
        coef = np.linalg.solve(gram, rhs[:, :, None])[:, :, 0]   # (n, m)
        cont[:, sel] = d0 @ coef.T
    return cont.reshape(w_raw.shape)


def o2test(arr):
    """Second-order test per spaxel, ``mean_z arr^2`` (lib_origin.py:957-974)."""
    return np.mean(np.asarray(arr, dtype=np.float64) ** 2, axis=0)


def preprocessing(cube_raw, var, mask, dct_order=10, dct_approx=False,
                  local_max_size=3):
    """The array part of ``Preprocessing.run`` (steps.py:430-465, 472, 480).

    ``cube_raw`` has NaN replaced by 0 and ``var`` NaN replaced by +inf
    (origin.py:262-274).  Returns a dict with ``cube_std, cont_dct (float32),
    ima_std, ima_dct, cube_std_local_max, cube_std_local_min, cont_sumsq
    (= sum_z cont_dct^2, the argument of the log10 at :472), o2map (:480)``.
    """
    cube_raw = np.asarray(cube_raw, dtype=np.float64)
    var = np.asarray(var, dtype=np.float64)
    mask = np.asarray(mask, dtype=bool)
    cont = dct_residual(cube_raw, dct_order, var, dct_approx, mask)
    data = cube_raw - cont
    data[mask] = np.nan
    std = np.sqrt(var)
    cont = cont / std
    with np.errstate(invalid='ignore'):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore', RuntimeWarning)
            mean = np.nanmean(data, axis=(1, 2))
    data -= mean[:, None, None]
    data /= std
    data[mask] = 0
    lmax, lmin = compute_local_max(data, data, mask, local_max_size)
    cont32 = cont.astype(np.float32)
    return dict(
        cube_std=data,
        ima_std=data.mean(axis=0),
        cube_std_local_max=lmax,
        cube_std_local_min=lmin,
        cont_dct=cont32,
        ima_dct=cont32.mean(axis=0),
        cont_sumsq=np.sum(cont32 ** 2, axis=0),
        o2map=o2test(data),
        mean_lambda=mean,
    )


# --------------------------------------------------------------------------
# step05: TGLR matched filter
# --------------------------------------------------------------------------

def prepare_profiles(profiles, pcut=None, pmeansub=True):
    """Cut, L2-normalise and de-mean the profiles (lib_origin.py:1155-1165)."""
    out = []
    for prof in profiles:
        prof = np.array(prof, dtype=np.float64)
        if pcut is not None:
            peak = int(prof.argmax())
            above = np.flatnonzero(prof >= pcut)
            half = int(max(abs(above[0] - peak), abs(above[-1] - peak)))
            prof = prof[peak - half:peak + half + 1]
        prof = prof / np.linalg.norm(prof)
        if pmeansub:
            prof = prof - prof.mean()
        out.append(prof)
    return out


def _as_fields(fsf, weights):
    if weights is None:                       # lib_origin.py:1112-1114
        return [np.asarray(fsf)], [None]
    return [np.asarray(f) for f in fsf], list(weights)


def fsf_correlate(cube, fsf, weights=None, workers=1, zchunk=64):
    """Spatial stage (lib_origin.py:1027-1043 per plane, summed over fields
    at :1130-1147), as a zero-padded FFT linear convolution.

    ``K = flip(psf) - mean(psf)``; ``cube_fsf = conv2_same(cube * w, K)``;
    ``norm_fsf = conv2_same(w or 1, K^2)``.
    """
    cube = np.asarray(cube, dtype=np.float64)
    nz, ny, nx = cube.shape
    fsfs, wts = _as_fields(fsf, weights)
    p = fsfs[0].shape[-1]
    c = (p - 1) // 2
    fy = sfft.next_fast_len(ny + p - 1, real=True)
    fx = sfft.next_fast_len(nx + p - 1, real=True)
    cube_fsf = np.zeros_like(cube)
    norm_fsf = np.zeros_like(cube)
    for psf, w in zip(fsfs, wts):
        ones = np.ones((ny, nx)) if w is None else np.asarray(w, dtype=np.float64)
        ones_f = sfft.rfft2(ones, (fy, fx))
        for z0 in range(0, nz, zchunk):
            z1 = min(nz, z0 + zchunk)
            ker = np.array(psf[z0:z1, ::-1, ::-1], dtype=np.float64)
            ker -= ker.mean(axis=(1, 2), keepdims=True)
            data = cube[z0:z1] if w is None else cube[z0:z1] * w
            ker_f = sfft.rfft2(ker, (fy, fx), workers=workers)
            full = sfft.irfft2(sfft.rfft2(data, (fy, fx), workers=workers) * ker_f,
                               (fy, fx), workers=workers)
            cube_fsf[z0:z1] += full[:, c:c + ny, c:c + nx]
            ker_f = sfft.rfft2(ker * ker, (fy, fx), workers=workers)
            full = sfft.irfft2(ones_f[None] * ker_f, (fy, fx), workers=workers)
            norm_fsf[z0:z1] += full[:, c:c + ny, c:c + nx]
    return cube_fsf, norm_fsf


def fsf_correlate_direct(cube, fsf, weights=None):
    """Direct-space twin of :func:`fsf_correlate`:
    ``cube_fsf[y,x] = sum_{dy,dx} (w cube)[y+dy, x+dx] (psf[c+dy, c+dx] - mean)``
    with zeros outside the image."""
    cube = np.asarray(cube, dtype=np.float64)
    nz, ny, nx = cube.shape
    fsfs, wts = _as_fields(fsf, weights)
    p = fsfs[0].shape[-1]
    c = (p - 1) // 2
    cube_fsf = np.zeros_like(cube)
    norm_fsf = np.zeros_like(cube)
    for psf, w in zip(fsfs, wts):
        ker = np.asarray(psf, dtype=np.float64)
        ker = ker - ker.mean(axis=(1, 2), keepdims=True)
        wmap = np.ones((ny, nx)) if w is None else np.asarray(w, dtype=np.float64)
        data = np.zeros((nz, ny + 2 * c, nx + 2 * c))
        data[:, c:c + ny, c:c + nx] = cube * wmap
        wpad = np.zeros((ny + 2 * c, nx + 2 * c))
        wpad[c:c + ny, c:c + nx] = wmap
        for dy in range(p):
            for dx in range(p):
                k = ker[:, dy, dx][:, None, None]
                cube_fsf += k * data[:, dy:dy + ny, dx:dx + nx]
                norm_fsf += (k * k) * wpad[None, dy:dy + ny, dx:dx + nx]
    return cube_fsf, norm_fsf


def spectral_glr(cube_fsf, norm_fsf, prof_cut, workers=1, colchunk=8192):
    """Spectral stage and reduction over profiles (lib_origin.py:1046-1066,
    1170-1217).

    Linear convolution along lambda of every spectrum with every profile,
    'same' window starting at ``(L_k - 1)//2`` (``:1179-1181``), normalised by
    ``sqrt(conv(norm_fsf, d_k^2))`` with non-positive norms sent to +inf
    (``:1057-1059``); running strict-``>`` argmax, max and min (``:1210-1212``).
    """
    nz = cube_fsf.shape[0]
    shape = cube_fsf.shape
    num_in = cube_fsf.reshape(nz, -1)
    den_in = norm_fsf.reshape(nz, -1)
    ncol = num_in.shape[1]
    lmax = max(d.shape[0] for d in prof_cut)
    flen = sfft.next_fast_len(nz + lmax - 1, real=True)
    taps_f = [sfft.rfft(d, flen) for d in prof_cut]
    taps2_f = [sfft.rfft(d * d, flen) for d in prof_cut]
    correl = np.full((nz, ncol), -np.inf)
    correl_min = np.full((nz, ncol), np.inf)
    profile = np.zeros((nz, ncol), dtype=np.uint8)
    for c0 in range(0, ncol, colchunk):
        sl = slice(c0, min(ncol, c0 + colchunk))
        num_f = sfft.rfft(num_in[:, sl], flen, axis=0, workers=workers)
        den_f = sfft.rfft(den_in[:, sl], flen, axis=0, workers=workers)
        best = correl[:, sl]
        worst = correl_min[:, sl]
        arg = profile[:, sl]
        for k, d in enumerate(prof_cut):
            start = (d.shape[0] - 1) // 2
            num = sfft.irfft(num_f * taps_f[k][:, None], flen, axis=0,
                             workers=workers)[start:start + nz]
            den = sfft.irfft(den_f * taps2_f[k][:, None], flen, axis=0,
                             workers=workers)[start:start + nz]
            den[den <= 0] = np.inf
            np.sqrt(den, out=den)
            num /= den
            arg[num > best] = k
            np.maximum(best, num, out=best)
            np.minimum(worst, num, out=worst)
    return correl.reshape(shape), profile.reshape(shape), correl_min.reshape(shape)


def spectral_glr_direct(cube_fsf, norm_fsf, prof_cut):
    """Direct-space twin of :func:`spectral_glr`:
    ``num_k[z] = sum_j d_k[j] cube_fsf[z + c_k - j]``, ``c_k = (L_k - 1)//2``,
    terms outside [0, Nz) dropped; ``den_k`` likewise with ``d_k^2`` and
    ``norm_fsf``; ``T_k = num_k / sqrt(den_k)``, 0 where ``den_k <= 0``.
    Returns ``(correl, profile, correl_min, T)`` with ``T`` the (K, ...) stack.
    """
    nz = cube_fsf.shape[0]
    tk = []
    for d in prof_cut:
        ck = (d.shape[0] - 1) // 2
        num = np.zeros_like(cube_fsf)
        den = np.zeros_like(norm_fsf)
        for j, dj in enumerate(d):
            shift = ck - j                      # source index = z + shift
            lo, hi = max(0, -shift), min(nz, nz - shift)
            if hi <= lo:
                continue
            num[lo:hi] += dj * cube_fsf[lo + shift:hi + shift]
            den[lo:hi] += dj * dj * norm_fsf[lo + shift:hi + shift]
        with np.errstate(divide='ignore', invalid='ignore'):
            t = np.where(den > 0, num / np.sqrt(np.where(den > 0, den, 1.0)), 0.0)
        tk.append(t)
    tk = np.stack(tk)
    correl = tk.max(axis=0)
    profile = tk.argmax(axis=0).astype(np.uint8)      # first maximum wins
    correl_min = tk.min(axis=0)
    return correl, profile, correl_min, tk


def correlation_glr_test(cube, fsf, weights, profiles, nthreads=1, pcut=None,
                         pmeansub=True):
    """``Correlation_GLR_test`` (lib_origin.py:1070-1217); same arguments and
    return order ``(correl, profile, correl_min)``."""
    cube_fsf, norm_fsf = fsf_correlate(cube, fsf, weights, workers=nthreads)
    prof_cut = prepare_profiles(profiles, pcut, pmeansub)
    return spectral_glr(cube_fsf, norm_fsf, prof_cut, workers=nthreads)


# --------------------------------------------------------------------------
# 3-D local extrema
# --------------------------------------------------------------------------

def _max_filter(arr, size):
    """Separable maximum filter, odd window, scipy's default ``mode='reflect'``
    (= numpy ``symmetric`` padding) as used at lib_origin.py:1244,1251."""
    out = arr
    for axis, s in enumerate(size):
        if s <= 1:
            continue
        if s % 2 == 0:
            raise ValueError('only odd window sizes are supported')
        r = s // 2
        pad = [(0, 0)] * arr.ndim
        pad[axis] = (r, r)
        padded = np.pad(out, pad, mode='symmetric')
        n = arr.shape[axis]
        res = None
        for o in range(s):
            idx = [slice(None)] * arr.ndim
            idx[axis] = slice(o, o + n)
            piece = padded[tuple(idx)]
            res = piece.copy() if res is None else np.maximum(res, piece, out=res)
        out = res
    return out


def compute_local_max(correl, correl_min, mask, size=3):
    """``compute_local_max`` (lib_origin.py:1220-1256): values of ``correl`` at
    its (size^3) local maxima outside the mask, 0 elsewhere; same for
    ``-correl_min``."""
    if np.isscalar(size):
        size = (size, size, size)
    mask = np.asarray(mask, dtype=bool)
    out = []
    for arr in (np.asarray(correl, dtype=np.float64),
                -np.asarray(correl_min, dtype=np.float64)):
        peak = _max_filter(arr, size)
        keep = (arr == peak) & ~mask
        out.append(peak * keep)
    return out[0], out[1]


def tglr_step(cube, fsf, weights, profiles, mask, size=3, nthreads=1, pcut=1e-8,
              pmeansub=True):
    """The array part of ``ComputeTGLR.run`` (steps.py:768-802)."""
    mask = np.asarray(mask, dtype=bool)
    correl, profile, correl_min = correlation_glr_test(
        cube, fsf, weights, profiles, nthreads, pcut, pmeansub)
    correl[mask] = 0                                   # steps.py:781
    profile[mask] = 0                                  # steps.py:788
    lmax, lmin = compute_local_max(correl, correl_min, mask, size)
    return dict(
        cube_correl=correl, cube_correl_min=correl_min, cube_profile=profile,
        maxmap=np.amax(correl, axis=0), minmap=np.amin(correl_min, axis=0),
        cube_local_max=lmax, cube_local_min=lmin,
    )


# --------------------------------------------------------------------------
# step06 / step07
# --------------------------------------------------------------------------

def threshold_purity(purity, cube_local_max, cube_local_min, segmap=None,
                     threshlist=None):
    """``Compute_threshold_purity`` (lib_origin.py:1391-1479).

    Returns ``(threshold, table)`` where ``table`` is a dict of the four
    columns ``Tval_r, Pval_r, Det_m, Det_M`` sorted by ``Tval_r`` (the astropy
    Table is only a container in the reference).
    """
    lmax = np.asarray(cube_local_max)
    lmin = np.asarray(cube_local_min)
    l1 = int(np.prod(lmin.shape[1:]))
    if segmap is not None:
        segmask = np.asarray(segmap) == 0
        lmin = lmin * segmask
        l0 = int(np.count_nonzero(segmask))
    else:
        l0 = l1
    if threshlist is None:
        threshmax = min(lmin.max(), lmax.max())
        threshmin = np.median(np.amax(lmax, axis=0)) * 1.1
        threshlist = np.linspace(threshmin, threshmax, 50)
    else:
        threshlist = np.asarray(threshlist, dtype=np.float64)
        threshmin = np.min(threshlist)
    loc_max = lmax[lmax > threshmin]
    loc_min = lmin[lmin > threshmin]
    n1 = np.array([np.count_nonzero(loc_max > t) for t in threshlist])
    n0 = np.array([np.count_nonzero(loc_min > t) for t in threshlist]) * (l1 / l0)
    with np.errstate(divide='ignore', invalid='ignore'):
        est = 1 - n0 / n1
    order = np.argsort(threshlist, kind='stable')
    table = dict(Tval_r=np.asarray(threshlist, dtype=np.float64)[order],
                 Pval_r=est[order], Det_m=n0.astype(int)[order], Det_M=n1[order])
    if est[-1] < purity:
        threshold = np.inf
    else:
        threshold = np.interp(purity, table['Pval_r'], table['Tval_r'])
    return float(threshold), table


def detection_rows(cube_local_max, cube_profile, threshold):
    """Raw detection rows of ``Detection.run`` (steps.py:956-964): C-order
    ``np.where(local_max > threshold)`` -> ``x0, y0, z0, T_GLR, profile``."""
    z, y, x = np.where(cube_local_max > threshold)
    value = cube_local_max[z, y, x]
    prof = cube_profile[z, y, x] if cube_profile is not None else np.zeros(len(z), np.uint8)
    return dict(x0=x, y0=y, z0=z, value=value, profile=prof)


def spot_tglr(cube, fsf, prof_cut, points):
    """Direct-space float64 ``T_k`` at isolated voxels of a single-field cube
    (SURVEY.md appendix A.2 evaluated pointwise): for each ``(z, y, x)`` in
    ``points`` returns the K normalised correlations.  Size-independent check
    for cubes too large for the full oracle."""
    cube = np.asarray(cube)
    nz, ny, nx = cube.shape
    fsf = np.asarray(fsf, dtype=np.float64)
    p = fsf.shape[-1]
    c = p // 2
    out = np.zeros((len(points), len(prof_cut)))
    half = max((len(d) - 1) // 2 for d in prof_cut) + 1
    for n, (z, y, x) in enumerate(points):
        z0, z1 = max(0, z - half), min(nz, z + half + 1)
        ya, yb = max(0, y - c), min(ny, y + c + 1)
        xa, xb = max(0, x - c), min(nx, x + c + 1)
        ker = fsf[z0:z1] - fsf[z0:z1].mean(axis=(1, 2), keepdims=True)
        ker = ker[:, ya - y + c:yb - y + c, xa - x + c:xb - x + c]
        patch = np.asarray(cube[z0:z1, ya:yb, xa:xb], dtype=np.float64)
        num_z = (patch * ker).sum(axis=(1, 2))           # cube_fsf[z0:z1, y, x]
        den_z = (ker * ker).sum(axis=(1, 2))             # norm_fsf[z0:z1, y, x]
        for k, d in enumerate(prof_cut):
            ck = (len(d) - 1) // 2
            num = den = 0.0
            for j, dj in enumerate(d):
                zz = z + ck - j
                if 0 <= zz < nz:
                    num += dj * num_z[zz - z0]
                    den += dj * dj * den_z[zz - z0]
            out[n, k] = num / np.sqrt(den) if den > 0 else 0.0
    return out


def spot_box(win, fsf, prof_cut, origin, gshape, centre):
    """Direct-space float64 ``T_k`` on the 3x3x3 neighbourhood of ``centre = (z, y, x)`` of a single-field
    ``gshape`` cube (SURVEY.md appendix A.2 evaluated pointwise), from a window ``win`` of that cube whose
    first voxel is ``origin`` and which covers the neighbourhood grown by the FSF half-size and the
    longest profile.  Returns ``dict(tk=[3][3][3][K] (NaN outside the cube), valid=[3][3][3] bool)``:
    enough to check correl / correl_min / argmax at the centre AND whether the centre is a 3x3x3 local
    extremum (lib_origin.py:1244-1253), on cubes far too large for the full oracle."""
    nz, ny, nx = gshape
    z, y, x = centre
    z0, y0, x0 = origin
    fsf = np.asarray(fsf, dtype=np.float64)
    p = fsf.shape[-1]
    c = p // 2
    reach = max(len(d) for d in prof_cut)
    za, zb = max(0, z - 1 - reach), min(nz, z + 2 + reach)
    za, zb = max(za, z0), min(zb, z0 + win.shape[0])
    ker_all = fsf[za:zb] - fsf[za:zb].mean(axis=(1, 2), keepdims=True)
    nk = len(prof_cut)
    tk = np.full((3, 3, 3, nk), np.nan)
    valid = np.zeros((3, 3, 3), dtype=bool)
    for dy in range(3):
        for dx in range(3):
            yy, xx = y - 1 + dy, x - 1 + dx
            if not (0 <= yy < ny and 0 <= xx < nx):
                continue
            ya, yb = max(0, yy - c), min(ny, yy + c + 1)
            xa, xb = max(0, xx - c), min(nx, xx + c + 1)
            if ya < y0 or xa < x0 or yb > y0 + win.shape[1] or xb > x0 + win.shape[2]:
                raise ValueError('window does not cover the FSF footprint of (%d, %d)' % (yy, xx))
            ker = ker_all[:, ya - yy + c:yb - yy + c, xa - xx + c:xb - xx + c]
            patch = np.asarray(win[za - z0:zb - z0, ya - y0:yb - y0, xa - x0:xb - x0], dtype=np.float64)
            num_z = (patch * ker).sum(axis=(1, 2))          # cube_fsf[za:zb, yy, xx]
            den_z = (ker * ker).sum(axis=(1, 2))            # norm_fsf[za:zb, yy, xx]
            for dz in range(3):
                zz = z - 1 + dz
                if not 0 <= zz < nz:
                    continue
                valid[dz, dy, dx] = True
                for k, d in enumerate(prof_cut):
                    ck = (len(d) - 1) // 2
                    src = zz + ck - np.arange(len(d))
                    ok = (src >= 0) & (src < nz)
                    if np.any(ok & ((src < za) | (src >= zb))):
                        raise ValueError('window does not cover the profile reach at z = %d' % zz)
                    num = float(np.dot(d[ok], num_z[src[ok] - za]))
                    den = float(np.dot(d[ok] ** 2, den_z[src[ok] - za]))
                    tk[dz, dy, dx, k] = num / np.sqrt(den) if den > 0 else 0.0
    return dict(tk=tk, valid=valid)
