"""Host-side mirror of the reference's numerical library for the hot path.

Same function names, argument order and error behaviour as
``muse_origin/lib_origin.py`` for the functions the step layer imports by name
(``steps.py:19-41``): ``DCTMAT``, ``dct_residual``, ``O2test``,
``Correlation_GLR_test``, ``compute_local_max``, ``Compute_threshold_purity``.
All arithmetic runs in ``libogn.so`` (hand-written CUDA for sm_100a); without
it, or without a B200, every call raises — there is no CPU fallback.

Inputs may be numpy arrays (host) or torch CUDA tensors (device; used in
place).  Results come back as numpy arrays for numpy inputs and as torch
tensors on the same device for tensor inputs.  Floating-point cubes are
float32 (the kernels compute in FP32, the DCT fit in FP64); pass
``out_dtype=np.float64`` where the reference's float64 container is needed.
"""

import numpy as np

from . import _lib
from ._lib import OGN_F32, OGN_F64, OgnError, default_context, ptr

__all__ = ['DCTMAT', 'dct_residual', 'O2test', 'Compute_PCA_threshold', 'Compute_GreedyPCA', 'Compute_GreedyPCA_area',
           'Correlation_GLR_test', 'compute_local_max',
           'Compute_threshold_purity', 'prepare_profiles', 'tglr', 'local_extrema', 'LocalExtrema', 'DeviceExtrema',
           'purity_counts', 'check_counts', 'threshold_rows', 'preprocess', 'PurityTable', 'step05', 'fsf_stage',
           'line_estimates', 'estimation_line', 'peakdet']


# --------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------

def _is_torch(x):
    return _lib._is_torch(x)


def _torch():
    import torch
    return torch


def _dtype_code(x):
    if _is_torch(x):
        torch = _torch()
        if x.dtype == torch.float32:
            return OGN_F32
        if x.dtype == torch.float64:
            return OGN_F64
        raise TypeError('cube tensors must be float32 or float64')
    if x.dtype == np.float32:
        return OGN_F32
    if x.dtype == np.float64:
        return OGN_F64
    raise TypeError('cube arrays must be float32 or float64')


def _as_float_cube(x):
    """C-contiguous float32/float64 view of a cube (numpy or torch)."""
    if _is_torch(x):
        torch = _torch()
        if x.dtype not in (torch.float32, torch.float64):
            x = x.to(torch.float32)
        return x.contiguous()
    x = np.asarray(x)
    if x.dtype not in (np.float32, np.float64):
        x = x.astype(np.float64)
    return np.ascontiguousarray(x)


def _as_u8(x, like=None):
    if x is None:
        return None
    if _is_torch(x):
        torch = _torch()
        return (x if x.dtype == torch.uint8 else x.to(torch.uint8)).contiguous()
    x = np.asarray(x)
    if x.dtype == np.bool_:
        return np.ascontiguousarray(x).view(np.uint8)
    if x.dtype == np.uint8:                      # any non-zero byte means "masked"; no copy
        return np.ascontiguousarray(x)
    return np.ascontiguousarray(x != 0).view(np.uint8)


def _empty_like_kind(ref, shape, dtype):
    """Uninitialised output living where ``ref`` lives."""
    if _is_torch(ref) and ref.is_cuda:
        torch = _torch()
        tdt = {np.float32: torch.float32, np.float64: torch.float64, np.uint8: torch.uint8,
               np.int64: torch.int64}[np.dtype(dtype).type]
        return torch.empty(tuple(shape), dtype=tdt, device=ref.device)
    return np.empty(shape, dtype=dtype)


def _ctx_for(x, ctx):
    if ctx is not None:
        return ctx
    if _is_torch(x) and x.is_cuda:
        return default_context(x.device.index)
    return default_context()


def _f64_host(x):
    if _is_torch(x):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def _f64_any(x):
    """float64 C-contiguous array on the host, or tensor left on its CUDA device."""
    if _is_torch(x) and x.is_cuda:
        return x.double().contiguous()
    return _f64_host(x)


class _PtrArray:
    """ctypes array of raw addresses (``const double *const *``), keeping the
    referenced arrays alive."""

    def __init__(self, arrays):
        import ctypes
        self.keep = list(arrays)
        self.arr = (ctypes.c_void_p * len(self.keep))(*[ptr(a) for a in self.keep])

    @property
    def address(self):
        import ctypes
        return ctypes.addressof(self.arr)


# --------------------------------------------------------------------------
# step01
# --------------------------------------------------------------------------

def DCTMAT(nl, order):
    """DCT synthesis matrix, ``nl x (order+1)`` (reference lib_origin.py:127-146).
    Host numpy: it is a 3681 x 11 table; the kernels rebuild it on their side."""
    yy, xx = np.mgrid[:nl, :order + 1]
    d0 = np.sqrt(2 / nl) * np.cos((yy + 0.5) * (np.pi / nl) * xx)
    d0[:, 0] *= 1 / np.sqrt(2)
    return d0


def dct_residual(w_raw, order, var, approx, mask, out_dtype=np.float64, ctx=None):
    """Continuum estimated by a (variance-weighted) projection on the first
    ``order+1`` DCT atoms (reference lib_origin.py:150-240).  Returns the
    continuum, like the reference."""
    raw = _as_float_cube(w_raw)
    if raw.ndim != 3:
        raise ValueError('w_raw must be a (nz, ny, nx) cube')
    code = _dtype_code(raw)
    v = None
    if var is not None:
        v = _as_float_cube(var)
        if _dtype_code(v) != code:
            v = v.to(raw.dtype) if _is_torch(v) else v.astype(raw.dtype)
    m = _as_u8(mask)
    ctx = _ctx_for(raw, ctx)
    nz, ny, nx = raw.shape
    cont = _empty_like_kind(raw, raw.shape, out_dtype)
    ctx.check(ctx.lib.ogn_dct_residual(ctx.handle, ptr(raw), ptr(v), code, ptr(m), nz, ny, nx, int(order),
                                       int(bool(approx)), ptr(cont), OGN_F64 if np.dtype(out_dtype) == np.float64 else OGN_F32))
    return cont


def O2test(arr):
    """Second-order test per spaxel, ``mean_z arr^2`` (reference
    lib_origin.py:957-974).  The fused step01 path (:func:`preprocess`) returns
    this map as a by-product; this stand-alone form is a 2-D reduction of an
    array the caller already holds."""
    if _is_torch(arr):
        return (arr.double() ** 2).mean(dim=0)
    return np.mean(np.asarray(arr, dtype=np.float64) ** 2, axis=0)


def preprocess(cube_raw, var, mask, dct_order=10, dct_approx=False, allreduce=None, owned=None, ctx=None):
    """Array part of ``Preprocessing.run`` (reference steps.py:431-465, :472,
    :480) in two device phases.

    Multi-GPU: ``owned = (y0, y1, x0, x1)`` restricts the per-wavelength sums
    (``np.nanmean`` over all spaxels, steps.py:442) to the spaxels this tile
    owns, and ``allreduce(sum, cnt)`` combines the partial sums across ranks in
    place (float64 arrays of length nz) before the mean is taken."""
    raw = _as_float_cube(cube_raw)
    code = _dtype_code(raw)
    v = _as_float_cube(var)
    if _dtype_code(v) != code:
        v = v.to(raw.dtype) if _is_torch(v) else v.astype(raw.dtype)
    m = _as_u8(mask)
    ctx = _ctx_for(raw, ctx)
    nz, ny, nx = raw.shape
    out = dict(
        cube_std=_empty_like_kind(raw, raw.shape, np.float32),
        cont_dct=_empty_like_kind(raw, raw.shape, np.float32),
        ima_std=_empty_like_kind(raw, (ny, nx), np.float64), ima_dct=_empty_like_kind(raw, (ny, nx), np.float64),
        cont_sumsq=_empty_like_kind(raw, (ny, nx), np.float64), o2map=_empty_like_kind(raw, (ny, nx), np.float64),
    )
    if allreduce is None and owned is None:
        # one device: both phases in one call, the per-wavelength mean stays on the device (device inputs:
        # nothing synchronises, the images come back as device tensors too)
        mean = _empty_like_kind(raw, (nz,), np.float64)
        ctx.check(ctx.lib.ogn_preprocess(ctx.handle, ptr(raw), ptr(v), code, ptr(m), nz, ny, nx, int(dct_order),
                                         int(bool(dct_approx)), ptr(mean), ptr(out['cube_std']), ptr(out['cont_dct']),
                                         ptr(out['ima_std']), ptr(out['ima_dct']), ptr(out['cont_sumsq']),
                                         ptr(out['o2map'])))
        out['mean_lambda'] = mean
        return out
    lsum = np.zeros(nz)
    lcnt = np.zeros(nz)
    win = None if owned is None else np.ascontiguousarray(owned, dtype=np.int32)
    ctx.check(ctx.lib.ogn_preprocess_begin(ctx.handle, ptr(raw), ptr(v), code, ptr(m), nz, ny, nx, int(dct_order),
                                           int(bool(dct_approx)), ptr(win), ptr(lsum), ptr(lcnt)))
    if allreduce is not None:
        allreduce(lsum, lcnt)
    with np.errstate(invalid='ignore', divide='ignore'):
        mean = lsum / lcnt                      # NaN for fully masked planes, as np.nanmean
    ctx.check(ctx.lib.ogn_preprocess_finish(ctx.handle, ptr(mean), ptr(out['cube_std']), ptr(out['cont_dct']),
                                            ptr(out['ima_std']), ptr(out['ima_dct']), ptr(out['cont_sumsq']),
                                            ptr(out['o2map'])))
    out['mean_lambda'] = mean
    return out


# --------------------------------------------------------------------------
# step03 / step04: PCA threshold and greedy PCA
# --------------------------------------------------------------------------

def Compute_PCA_threshold(faint, pfa):
    """O2 test of a block of spectra and its threshold from a Gaussian fit of the test's distribution
    (reference lib_origin.py:821-842): ``(test, histO2, frecO2, thresO2, mea, std)``.  The fit is the numpy
    restatement of :mod:`origin_b200.segmap` (the reference's needs astropy)."""
    from . import segmap
    test = O2test(faint)
    test_host = test.detach().cpu().numpy() if _is_torch(test) else test
    hist, edges, thres, mea, std = segmap.compute_thresh_gaussfit(test_host, pfa)
    return test_host, hist, edges, thres, mea, std


def _greedy_pca_block(cube, cols, test, thres, noise_population, itermax, faint, ctx):
    """One ``ogn_greedy_pca`` call: the spaxels ``cols`` of the ``(nz, ld)`` matrix view of ``cube``."""
    nz = cube.shape[0]
    ld = int(np.prod(cube.shape[1:]))
    n = ld if cols is None else len(cols)
    cols = None if cols is None else np.ascontiguousarray(cols, dtype=np.int64)
    t0 = None if test is None else np.ascontiguousarray(test, dtype=np.float64)
    if t0 is not None and t0.size != n:
        raise ValueError('test must hold one value per spaxel of the block')
    map_o2 = np.zeros(n, dtype=np.float64)
    info = np.zeros(3, dtype=np.int32)
    ctx.check(ctx.lib.ogn_greedy_pca(ctx.handle, ptr(cube), _dtype_code(cube), nz, ld, ptr(cols), n, ptr(t0), float(thres),
                                     float(noise_population), int(itermax), ptr(faint), _dtype_code(faint), ptr(map_o2),
                                     ptr(info)))
    return map_o2, int(info[0]), int(info[1]), int(info[2])


def Compute_GreedyPCA(cube_in, test, thresO2, Noise_population, itermax, ctx=None):
    """Greedy PCA of a block of spectra ``(nz, npix)`` (reference lib_origin.py:858-954): returns
    ``(faint, mapO2, nstop)``.  ``test`` is the block's O2 test (or None to compute it)."""
    cube = _as_float_cube(cube_in)
    if cube.ndim != 2:
        raise ValueError('cube_in must be (nz, npix)')
    ctx = _ctx_for(cube, ctx)
    faint = cube.clone() if _is_torch(cube) else cube.copy()
    map_o2, nstop, _, _ = _greedy_pca_block(cube, None, test, thresO2, Noise_population, itermax, faint, ctx)
    return faint, map_o2, nstop


def Compute_GreedyPCA_area(NbArea, cube_std, areamap, Noise_population, threshold_test, itermax, testO2, ctx=None):
    """Greedy PCA on each area of the field (reference lib_origin.py:769-818): returns
    ``(cube_faint, mapO2, nstop)``.  ``cube_std`` may be a numpy cube (the result is a numpy cube of the same
    dtype) or a CUDA tensor (float32 from the fused step01: ``cube_faint`` then stays on the device for step05).
    A numpy cube crosses PCIe twice — up once, all areas run on the device copy, the result comes down once —
    not once per area."""
    cube = _as_float_cube(cube_std)
    if cube.ndim != 3:
        raise ValueError('cube_std must be (nz, ny, nx)')
    ctx = _ctx_for(cube, ctx)
    areamap = areamap.detach().cpu().numpy() if _is_torch(areamap) else np.asarray(areamap)
    host_in = not _is_torch(cube)
    if host_in:
        torch = _torch()
        cube_faint = torch.from_numpy(cube).to(torch.device('cuda', ctx.device))    # :797 (the areas only write their own
        cube = cube_faint                                                            # columns: input and output can share)
    else:
        cube_faint = cube.clone()
    map_o2 = np.zeros(cube.shape[1:], dtype=np.float64)
    nstop = 0
    for area_ind in range(1, NbArea + 1):
        ksel = areamap == area_ind                                               # :806
        cols = np.flatnonzero(ksel.reshape(-1))                                  # C order = boolean-mask order of cube[:, ksel]
        if cols.size == 0:
            continue
        test = None if testO2 is None else testO2[area_ind - 1]
        m, k, _, _ = _greedy_pca_block(cube, cols, test, threshold_test[area_ind - 1], Noise_population, itermax,
                                       cube_faint, ctx)
        map_o2[ksel] = m                                                         # :815
        nstop += k
    if host_in:
        cube_faint = cube_faint.cpu().numpy()
    return cube_faint, map_o2, nstop


# --------------------------------------------------------------------------
# step05
# --------------------------------------------------------------------------

def prepare_profiles(profiles, pcut=None, pmeansub=True):
    """Cut at ``pcut``, L2-normalise, subtract the mean (reference
    lib_origin.py:1155-1165); float64 on the host, K short vectors."""
    out = []
    for prof in profiles:
        prof = np.array(prof, dtype=np.float64).ravel()
        if pcut is not None:
            lpeak = int(prof.argmax())
            sel = np.where(prof >= pcut)[0]
            lw = int(np.max(np.abs(sel[[0, -1]] - lpeak)))
            prof = prof[lpeak - lw:lpeak + lw + 1]
        prof = prof / np.linalg.norm(prof)
        if pmeansub:
            prof = prof - prof.mean()
        out.append(prof)
    return out


def _pack_profiles(prof_cut):
    offs = np.zeros(len(prof_cut) + 1, dtype=np.int32)
    offs[1:] = np.cumsum([len(p) for p in prof_cut])
    return np.ascontiguousarray(np.concatenate(prof_cut)), offs


def tglr(cube, fsf, weights, profiles, mask=None, pcut=None, pmeansub=True, want=('correl', 'profile', 'correl_min',
                                                                                  'maxmap', 'minmap'), ctx=None):
    """``Correlation_GLR_test`` fused with the masking / maxmap / minmap glue
    of ``ComputeTGLR.run`` (reference steps.py:781-793).  Returns a dict with
    the requested products."""
    cube, fsfs, fsf_ptrs, w_ptrs, taps, offs, nprof, m = _tglr_args(cube, fsf, weights, profiles, mask, pcut, pmeansub)
    nz, ny, nx = cube.shape
    psize = fsfs[0].shape[1]
    ctx = _ctx_for(cube, ctx)
    out = {}
    if 'correl' in want:
        out['correl'] = _empty_like_kind(cube, cube.shape, np.float32)
    if 'correl_min' in want:
        out['correl_min'] = _empty_like_kind(cube, cube.shape, np.float32)
    if 'profile' in want:
        out['profile'] = _empty_like_kind(cube, cube.shape, np.uint8)
    if 'maxmap' in want:
        out['maxmap'] = _empty_like_kind(cube, (ny, nx), np.float32)
    if 'minmap' in want:
        out['minmap'] = _empty_like_kind(cube, (ny, nx), np.float32)
    ctx.check(ctx.lib.ogn_tglr(
        ctx.handle, ptr(cube), _dtype_code(cube), nz, ny, nx, len(fsfs), fsf_ptrs.address, psize,
        w_ptrs.address if w_ptrs else None, ptr(taps), ptr(offs), nprof, ptr(m),
        ptr(out.get('correl')), ptr(out.get('correl_min')), ptr(out.get('profile')),
        ptr(out.get('maxmap')), ptr(out.get('minmap'))))
    return out


def _tglr_args(cube, fsf, weights, profiles, mask, pcut, pmeansub):
    """Validated, C-ready arguments shared by :func:`tglr` and :func:`step05`."""
    cube = _as_float_cube(cube)
    if cube.ndim != 3:
        raise ValueError('cube must be (nz, ny, nx)')
    nz, ny, nx = cube.shape
    if weights is None:                         # one FSF (lib_origin.py:1112-1114)
        fsfs, wmaps = [fsf], None
    else:
        fsfs, wmaps = list(fsf), list(weights)
        if len(fsfs) != len(wmaps):
            raise ValueError('fsf and weights must have the same length')
    fsfs = [_f64_any(f) for f in fsfs]
    for f in fsfs:
        if f.ndim != 3 or f.shape[0] != nz or f.shape[1] != f.shape[2]:
            raise ValueError('each FSF must be (nz, P, P), got %r' % (tuple(f.shape),))
    fsf_ptrs = _PtrArray(fsfs)
    w_ptrs = None
    if wmaps is not None:
        wmaps = [_f64_any(w) for w in wmaps]
        for w in wmaps:
            if tuple(w.shape) != (ny, nx):
                raise ValueError('weight maps must be (ny, nx)')
        w_ptrs = _PtrArray(wmaps)
    prof_cut = prepare_profiles(profiles, pcut, pmeansub)
    taps, offs = _pack_profiles(prof_cut)
    return cube, fsfs, fsf_ptrs, w_ptrs, taps, offs, len(prof_cut), _as_u8(mask)


def step05(cube, fsf, weights, profiles, mask, size=3, pcut=1e-8, pmeansub=True, out=None, dense=False,
           capacity=None, want=('correl', 'profile', 'correl_min', 'maxmap', 'minmap'), tile=None, ctx=None,
           sync=True, mask_bits=None, on_device=()):
    """The array part of ``ComputeTGLR.run`` (reference steps.py:768-802) in one
    device pass: TGLR, masking, maxmap / minmap and the local extrema.

    ``out`` may hold preallocated arrays (e.g. pinned host buffers from
    :func:`origin_b200._lib.pinned_empty`) for ``correl, correl_min, profile,
    maxmap, minmap, max_index, max_value, min_index, min_value``.
    Returns a dict with the products in ``want`` plus ``extrema``
    (:class:`LocalExtrema`) and, when ``dense``, ``cube_local_max`` /
    ``cube_local_min``.

    Multi-GPU: ``tile`` is an :class:`origin_b200.tiles.Tile` plus the field size,
    ``(tile, (gny, gnx))``; ``cube`` / ``mask`` are then the tile's padded
    sub-cube, only the owned window (grown by the extremum radius) is computed,
    the products are sub-cube shaped, and ``extrema`` holds the owned voxels with
    linear indices of the whole ``(nz, gny, gnx)`` field.

    Host cubes (float32, C-contiguous; pinned memory is fastest): the field crosses PCIe in slabs of image
    rows that overlap with the kernels and with the download of the products.  ``mask_bits`` — the mask as
    ``np.packbits(mask)`` (flattened ``[nz][ny][nx]``), to be packed once per session — replaces ``mask``
    on the link (1/8 of the bytes; ``nx`` a multiple of 8).  Products named in ``on_device`` (e.g.
    ``('correl_min', 'profile')``) are left on the GPU as torch tensors instead of being copied back: the
    step mirror hands them out lazily (:class:`origin_b200.steps.LazyProduct`).

    ``sync=False`` (CUDA tensors in and out only): the call returns with the kernels in flight and
    ``extrema`` is a :class:`DeviceExtrema` — the list lengths stay on the device until somebody reads
    them, so a pipeline of steps never stalls the GPU on the host.
    """
    cube, fsfs, fsf_ptrs, w_ptrs, taps, offs, nprof, m = _tglr_args(cube, fsf, weights, profiles, mask, pcut, pmeansub)
    if np.isscalar(size):
        size = (size, size, size)
    size = tuple(int(s) for s in size)
    if any(s < 1 or s % 2 == 0 for s in size):
        raise ValueError('only odd window sizes are supported, got %r' % (size,))
    nz, ny, nx = cube.shape
    vol = nz * ny * nx
    ctx = _ctx_for(cube, ctx)
    out = dict(out or {})
    res = {}
    shapes = dict(correl=(cube.shape, np.float32), correl_min=(cube.shape, np.float32), profile=(cube.shape, np.uint8),
                  maxmap=((ny, nx), np.float32), minmap=((ny, nx), np.float32))
    for key, (shape, dt) in shapes.items():
        if key in want:
            res[key] = out.get(key)
            if res[key] is None:
                if key in on_device and not _is_torch(cube):
                    torch = _torch()
                    res[key] = torch.empty(tuple(shape), dtype={np.float32: torch.float32, np.uint8: torch.uint8}[dt],
                                           device=torch.device('cuda', ctx.device))
                else:
                    res[key] = _empty_like_kind(cube, shape, dt)
    if dense:
        res['cube_local_max'] = _empty_like_kind(cube, cube.shape, np.float32)
        res['cube_local_min'] = _empty_like_kind(cube, cube.shape, np.float32)
    if capacity is None:
        capacity = len(out['max_index']) if out.get('max_index') is not None else max(4096, vol // 40)
    counts = np.zeros(2, dtype=np.int64)
    if not sync:
        if not (_is_torch(cube) and cube.is_cuda) or dense:
            raise ValueError('sync=False needs CUDA tensors and dense=False')
        counts = _torch().zeros(2, dtype=_torch().int64, device=cube.device)
    tdesc, ext_shape = None, cube.shape
    if tile is not None:
        if dense:
            raise ValueError('dense extremum cubes are not produced in tile mode')
        t, (gny, gnx) = tile
        tdesc = np.array([gny, gnx, t.py0, t.px0, t.y0 - t.py0, t.y1 - t.py0, t.x0 - t.px0, t.x1 - t.px0],
                         dtype=np.int32)
        if (t.py1 - t.py0, t.px1 - t.px0) != (ny, nx):
            raise ValueError('cube does not have the shape of the padded tile')
        psize = fsfs[0].shape[1]
        need = (psize // 2 + size[1] // 2, psize // 2 + size[2] // 2)
        for have, want_, at_edge, side in ((t.y0 - t.py0, need[0], t.py0 == 0, 'top'),
                                           (t.py1 - t.y1, need[0], t.py1 == gny, 'bottom'),
                                           (t.x0 - t.px0, need[1], t.px0 == 0, 'left'),
                                           (t.px1 - t.x1, need[1], t.px1 == gnx, 'right')):
            if have < want_ and not at_edge:
                raise ValueError('tile halo on the %s side is %d pixels; the FSF (P = %d) and the %r extremum window '
                                 'need %d where the field continues' % (side, have, psize, size, want_))
        ext_shape = (nz, gny, gnx)
    while True:
        lists = {}
        for key, dt in (('max_index', np.int64), ('max_value', np.float32), ('min_index', np.int64),
                        ('min_value', np.float32)):
            buf = out.get(key)
            lists[key] = buf if buf is not None and len(buf) >= capacity else _empty_like_kind(cube, (capacity,), dt)
        if mask_bits is not None:
            if tdesc is not None or dense or _is_torch(cube):
                raise ValueError('mask_bits goes with a host cube, no tile and dense=False')
            bits = np.ascontiguousarray(mask_bits, dtype=np.uint8)
            if bits.size != (vol + 7) // 8:
                raise ValueError('mask_bits must be np.packbits of the flattened (nz, ny, nx) mask')
            rc = ctx.lib.ogn_step05_bits(
                ctx.handle, ptr(cube), _dtype_code(cube), nz, ny, nx, len(fsfs), fsf_ptrs.address, fsfs[0].shape[1],
                w_ptrs.address if w_ptrs else None, ptr(taps), ptr(offs), nprof, ptr(bits), size[0], size[1], size[2],
                ptr(res.get('correl')), ptr(res.get('correl_min')), ptr(res.get('profile')), ptr(res.get('maxmap')),
                ptr(res.get('minmap')), ptr(lists['max_index']), ptr(lists['max_value']), ptr(lists['min_index']),
                ptr(lists['min_value']), capacity, ptr(counts))
        elif tdesc is None:
            rc = ctx.lib.ogn_step05(
                ctx.handle, ptr(cube), _dtype_code(cube), nz, ny, nx, len(fsfs), fsf_ptrs.address, fsfs[0].shape[1],
                w_ptrs.address if w_ptrs else None, ptr(taps), ptr(offs), nprof, ptr(m), size[0], size[1], size[2],
                ptr(res.get('correl')), ptr(res.get('correl_min')), ptr(res.get('profile')), ptr(res.get('maxmap')),
                ptr(res.get('minmap')), ptr(res.get('cube_local_max')), ptr(res.get('cube_local_min')),
                ptr(lists['max_index']), ptr(lists['max_value']), ptr(lists['min_index']), ptr(lists['min_value']),
                capacity, ptr(counts))
        else:
            rc = ctx.lib.ogn_step05_tile(
                ctx.handle, ptr(cube), _dtype_code(cube), nz, ny, nx, ptr(tdesc), len(fsfs), fsf_ptrs.address,
                fsfs[0].shape[1], w_ptrs.address if w_ptrs else None, ptr(taps), ptr(offs), nprof, ptr(m), size[0],
                size[1], size[2], ptr(res.get('correl')), ptr(res.get('correl_min')), ptr(res.get('profile')),
                ptr(res.get('maxmap')), ptr(res.get('minmap')), ptr(lists['max_index']), ptr(lists['max_value']),
                ptr(lists['min_index']), ptr(lists['min_value']), capacity, ptr(counts))
        rc = ctx.check(rc, allow_overflow=True)
        if not sync:
            res['extrema'] = DeviceExtrema(ext_shape, lists, counts, capacity)
            return res
        if rc == 0:
            break
        capacity = int(counts.max())
        out = {k: v for k, v in out.items() if k not in lists}
        if (tdesc is None and not dense and mask_bits is None and res.get('correl') is not None
                and res.get('correl_min') is not None):
            # a list did not fit: every other product is complete, so only the extremum pass is repeated on the
            # correl / correl_min just computed (not FSF + spectral stage + PCIe again)
            ext, _, _ = local_extrema(res['correl'], res['correl_min'], m, size, capacity=capacity, ctx=ctx)
            res['extrema'] = ext
            return res
    n1, n0 = int(counts[0]), int(counts[1])
    res['extrema'] = LocalExtrema(ext_shape, lists['max_index'][:n1], lists['max_value'][:n1],
                                  lists['min_index'][:n0], lists['min_value'][:n0])
    return res


def Correlation_GLR_test(cube, fsf, weights, profiles, nthreads=1, pcut=None, pmeansub=True, out_dtype=np.float64,
                         ctx=None):
    """GLR test cubes for the given FSF(s) and profile dictionary (reference
    lib_origin.py:1070-1217): returns ``(correl, profile, correl_min)``.
    ``nthreads`` is accepted for signature compatibility and ignored (the
    reference only uses it to chunk its FFTs; results do not depend on it).
    ``correl`` / ``correl_min`` come back as float64 like the reference's (callers
    mutate and keep them, steps.py:781); the kernels compute in FP32 — pass
    ``out_dtype=np.float32`` to skip the widening."""
    res = tglr(cube, fsf, weights, profiles, None, pcut, pmeansub, want=('correl', 'profile', 'correl_min'), ctx=ctx)
    correl, correl_min = res['correl'], res['correl_min']
    if np.dtype(out_dtype) == np.float64:
        correl = correl.double() if _is_torch(correl) else correl.astype(np.float64)
        correl_min = correl_min.double() if _is_torch(correl_min) else correl_min.astype(np.float64)
    return correl, res['profile'], correl_min


def fsf_stage(cube, fsf, weights, ctx=None):
    """The two intermediates of the spatial stage (reference ``_convolve_fsf``,
    lib_origin.py:1027-1043, summed over fields): ``(cube_fsf, norm_fsf)``."""
    cube = _as_float_cube(cube)
    nz, ny, nx = cube.shape
    if weights is None:
        fsfs, wmaps = [fsf], None
    else:
        fsfs, wmaps = list(fsf), list(weights)
    fsfs = [_f64_host(f) for f in fsfs]
    fsf_ptrs = _PtrArray(fsfs)
    w_ptrs = _PtrArray([_f64_host(w) for w in wmaps]) if wmaps is not None else None
    ctx = _ctx_for(cube, ctx)
    a = _empty_like_kind(cube, cube.shape, np.float32)
    b = _empty_like_kind(cube, cube.shape, np.float32)
    ctx.check(ctx.lib.ogn_fsf_stage(ctx.handle, ptr(cube), _dtype_code(cube), nz, ny, nx, len(fsfs),
                                    fsf_ptrs.address, fsfs[0].shape[1], w_ptrs.address if w_ptrs else None,
                                    ptr(a), ptr(b)))
    return a, b


# --------------------------------------------------------------------------
# local extrema
# --------------------------------------------------------------------------

class LocalExtrema:
    """Compact form of the two cubes ``compute_local_max`` returns: the kept
    voxels as ``(linear index, value)`` lists sorted in C order (the order of
    ``np.where``), for the maxima of ``correl`` and of ``-correl_min``."""

    def __init__(self, shape, max_index, max_value, min_index, min_value):
        self.shape = tuple(int(s) for s in shape)
        self.max_index, self.max_value = max_index, max_value
        self.min_index, self.min_value = min_index, min_value

    @property
    def counts(self):
        return len(self.max_index), len(self.min_index)

    def _host(self, x):
        return x.detach().cpu().numpy() if _is_torch(x) else x

    def dense(self, which='max', dtype=np.float32):
        """Materialise ``cube_local_max`` / ``cube_local_min`` on the host."""
        idx = self._host(self.max_index if which == 'max' else self.min_index)
        val = self._host(self.max_value if which == 'max' else self.min_value)
        out = np.zeros(int(np.prod(self.shape)), dtype=dtype)
        out[idx] = val
        return out.reshape(self.shape)

    def coords(self, which='max'):
        idx = self._host(self.max_index if which == 'max' else self.min_index)
        return np.unravel_index(idx, self.shape)


class DeviceExtrema(LocalExtrema):
    """:class:`LocalExtrema` whose lists are still being written by the GPU: the capacity-sized device buffers
    of an asynchronous :func:`step05` call plus the device array holding the two list lengths.  Nothing
    here touches the host until ``counts`` / ``max_index`` / ... are read (that synchronises and checks
    for overflow); :func:`purity_counts` consumes it without synchronising."""

    def __init__(self, shape, bufs, counts_dev, capacity):
        self.shape = tuple(int(s) for s in shape)
        self._bufs, self.counts_dev, self.capacity = bufs, counts_dev, int(capacity)
        self._counts = None

    @property
    def counts(self):
        if self._counts is None:
            n1, n0 = (int(v) for v in self.counts_dev.cpu())
            if max(n1, n0) > self.capacity:
                raise OverflowError('extremum lists need %d / %d entries, capacity was %d: call step05 again with '
                                    'capacity >= %d' % (n1, n0, self.capacity, max(n1, n0)))
            self._counts = (n1, n0)
        return self._counts

    max_index = property(lambda self: self._bufs['max_index'][:self.counts[0]])
    max_value = property(lambda self: self._bufs['max_value'][:self.counts[0]])
    min_index = property(lambda self: self._bufs['min_index'][:self.counts[1]])
    min_value = property(lambda self: self._bufs['min_value'][:self.counts[1]])


def local_extrema(correl, correl_min, mask, size=3, dense=False, capacity=None, ctx=None):
    """3-D local maxima of ``correl`` and of ``-correl_min`` outside the mask
    (reference lib_origin.py:1220-1256).  Returns ``(LocalExtrema, dense_max,
    dense_min)``; the dense cubes are None unless ``dense``."""
    if np.isscalar(size):
        size = (size, size, size)
    size = tuple(int(s) for s in size)
    if any(s < 1 or s % 2 == 0 for s in size):
        raise ValueError('only odd window sizes are supported, got %r' % (size,))
    a = _as_float_cube(correl)
    if _dtype_code(a) != OGN_F32:
        a = a.float() if _is_torch(a) else a.astype(np.float32)
    if correl_min is correl:
        b = a
    else:
        b = _as_float_cube(correl_min)
        if _dtype_code(b) != OGN_F32:
            b = b.float() if _is_torch(b) else b.astype(np.float32)
    if a.shape != b.shape or a.ndim != 3:
        raise ValueError('correl and correl_min must be cubes of the same shape')
    m = _as_u8(mask)
    ctx = _ctx_for(a, ctx)
    nz, ny, nx = a.shape
    vol = nz * ny * nx
    dmax = _empty_like_kind(a, a.shape, np.float32) if dense else None
    dmin = _empty_like_kind(a, a.shape, np.float32) if dense else None
    if capacity is None:
        # 3x3x3 maxima of white noise are 1/27 = 3.7 % of the voxels (the step01 call on cube_std, steps.py:453)
        capacity = max(4096, vol // 20)
    counts = np.zeros(2, dtype=np.int64)
    while True:
        mi = _empty_like_kind(a, (capacity,), np.int64)
        mv = _empty_like_kind(a, (capacity,), np.float32)
        ni = _empty_like_kind(a, (capacity,), np.int64)
        nv = _empty_like_kind(a, (capacity,), np.float32)
        rc = ctx.check(ctx.lib.ogn_local_extrema(ctx.handle, ptr(a), ptr(b), ptr(m), nz, ny, nx, size[0], size[1],
                                                 size[2], ptr(dmax), ptr(dmin), ptr(mi), ptr(mv), ptr(ni), ptr(nv),
                                                 capacity, ptr(counts)), allow_overflow=True)
        if rc == 0:
            break
        capacity = int(counts.max())
    n1, n0 = int(counts[0]), int(counts[1])
    ext = LocalExtrema(a.shape, mi[:n1], mv[:n1], ni[:n0], nv[:n0])
    return ext, dmax, dmin


def compute_local_max(correl, correl_min, mask, size=3, ctx=None):
    """Drop-in for the reference's ``compute_local_max`` (lib_origin.py:1220-1256):
    returns the dense ``(local_max, local_min)`` cubes (float32)."""
    _, dmax, dmin = local_extrema(correl, correl_min, mask, size, dense=True, ctx=ctx)
    return dmax, dmin


# --------------------------------------------------------------------------
# step06 / step07
# --------------------------------------------------------------------------

class PurityTable(dict):
    """Columns ``Tval_r, Pval_r, Det_m, Det_M`` of the purity table (the
    reference returns an astropy Table, lib_origin.py:1454-1460; astropy is not
    a dependency of the hot path, :meth:`to_astropy` converts when present)."""

    colnames = ('Tval_r', 'Pval_r', 'Det_m', 'Det_M')

    def to_astropy(self):
        from astropy.table import Table
        tab = Table([self[c] for c in self.colnames], names=self.colnames)
        tab['Tval_r'].format = '.2f'
        tab['Pval_r'].format = '.2f'
        return tab

    def __len__(self):
        return len(self['Tval_r'])


def _lists_from_dense(cube):
    flat = np.asarray(cube).reshape(-1)
    idx = np.flatnonzero(flat)
    return idx.astype(np.int64), flat[idx].astype(np.float32)


def _as_extrema(cube_local_max, cube_local_min):
    if isinstance(cube_local_max, LocalExtrema):
        return cube_local_max
    if _is_torch(cube_local_max):
        cube_local_max = cube_local_max.detach().cpu().numpy()
        cube_local_min = cube_local_min.detach().cpu().numpy()
    mi, mv = _lists_from_dense(cube_local_max)
    ni, nv = _lists_from_dense(cube_local_min)
    return LocalExtrema(np.shape(cube_local_max), mi, mv, ni, nv)


def purity_stats(ext, segmask, ctx=None):
    """``(max of maxima, max of background minima, per-spaxel max map)``
    from the compact lists (reference lib_origin.py:1437-1438 inputs)."""
    ctx = _ctx_for(ext.max_value, ctx)
    nz, ny, nx = ext.shape
    stats = np.zeros(2)
    spmax = np.empty((ny, nx), dtype=np.float32)
    seg = _as_u8(segmask)
    n1, n0 = ext.counts
    ctx.check(ctx.lib.ogn_purity_stats(ctx.handle, ptr(ext.max_index), ptr(ext.max_value), n1, ptr(ext.min_index),
                                       ptr(ext.min_value), n0, ptr(seg), ny, nx, ptr(stats), ptr(spmax)))
    return float(stats[0]), float(stats[1]), spmax


def _check_device_counts_args(thresholds, out):
    torch = _torch()
    if thresholds.dtype != torch.float64 or not thresholds.is_contiguous():
        raise TypeError('device thresholds must be a contiguous float64 tensor')
    nt = thresholds.numel()
    if out is None:
        out = torch.empty(2 * nt, dtype=torch.int64, device=thresholds.device)
    if not (_is_torch(out) and out.is_cuda and out.dtype == torch.int64 and out.numel() == 2 * nt and out.is_contiguous()):
        raise TypeError('out must be a contiguous int64 CUDA tensor of 2 * len(thresholds) entries')
    return nt, out


def check_counts(n1, n0):
    """Raise OverflowError when per-threshold counts carry the overflow mark of the device-only counting
    path (negative values: an extremum list of an asynchronous step did not fit its capacity)."""
    for n in (n1, n0):
        n = n.detach().cpu().numpy() if _is_torch(n) else np.asarray(n)
        if n.size and n.min() < 0:
            raise OverflowError('an extremum list overflowed its capacity in an asynchronous step05: the purity '
                                'counts are invalid; repeat the step with a larger capacity')


def purity_counts(ext, segmask, thresholds, ctx=None, out=None):
    """``n1[t] = #{maxima > t}``, ``n0[t] = #{background minima > t}`` (int64;
    the loop at reference lib_origin.py:1443-1449).

    Host thresholds give numpy counts (the call synchronises).  When ``thresholds`` is a float64 CUDA
    tensor the counts are written to ``out`` (an int64 CUDA tensor of ``2 * len(thresholds)`` entries,
    ``n1`` then ``n0``; allocated when None) and the call returns without synchronising, so a
    multi-GPU caller can hand ``out`` straight to an NCCL allreduce."""
    nz, ny, nx = ext.shape
    seg = _as_u8(segmask)
    if isinstance(ext, DeviceExtrema) and ext._counts is None and _is_torch(thresholds) and thresholds.is_cuda:
        # lists still in flight: device-only entry point, lengths read on the device
        torch = _torch()
        b = ext._bufs
        ctx = _ctx_for(b['max_value'], ctx)
        nt, out = _check_device_counts_args(thresholds, out)
        if seg is not None and not _is_torch(seg):
            seg = torch.from_numpy(seg).to(thresholds.device)
        # an overflowed list cannot raise here (nothing is synchronised): the kernel then makes every count
        # of that list negative; check_counts() / Compute_threshold_purity test for it once they are read
        ctx.check(ctx.lib.ogn_purity_counts_dev(ctx.handle, ptr(b['max_index']), ptr(b['max_value']), ptr(b['min_index']),
                                                ptr(b['min_value']), ext.capacity, ext.counts_dev.data_ptr(), ptr(seg), ny, nx,
                                                ptr(thresholds), nt, out.data_ptr(), out.data_ptr() + 8 * nt))
        return out[:nt], out[nt:]
    ctx = _ctx_for(ext.max_value, ctx)
    c1, c0 = ext.counts
    if _is_torch(thresholds) and thresholds.is_cuda:
        torch = _torch()
        nt, out = _check_device_counts_args(thresholds, out)
        if seg is not None and not _is_torch(seg):
            seg = torch.from_numpy(seg).to(thresholds.device)
        ctx.check(ctx.lib.ogn_purity_counts(ctx.handle, ptr(ext.max_index), ptr(ext.max_value), c1, ptr(ext.min_index),
                                            ptr(ext.min_value), c0, ptr(seg), ny, nx, ptr(thresholds), nt,
                                            out.data_ptr(), out.data_ptr() + 8 * nt))
        return out[:nt], out[nt:]
    thr = np.ascontiguousarray(thresholds, dtype=np.float64)
    n1 = np.zeros(len(thr), dtype=np.int64)
    n0 = np.zeros(len(thr), dtype=np.int64)
    ctx.check(ctx.lib.ogn_purity_counts(ctx.handle, ptr(ext.max_index), ptr(ext.max_value), c1, ptr(ext.min_index),
                                        ptr(ext.min_value), c0, ptr(seg), ny, nx, ptr(thr), len(thr), ptr(n1),
                                        ptr(n0)))
    return n1, n0


class _DeviceCounter:
    """Default backend of :func:`Compute_threshold_purity`: the K4 kernels."""

    def __init__(self, ctx):
        self.ctx = ctx

    def stats(self, ext, segmask):
        return purity_stats(ext, segmask, self.ctx)

    def counts(self, ext, segmask, thresholds):
        return purity_counts(ext, segmask, thresholds, self.ctx)


def Compute_threshold_purity(purity, cube_local_max, cube_local_min, segmap=None, threshlist=None, allreduce=None,
                             ctx=None, _backend=None):
    """Threshold for a target purity (reference lib_origin.py:1391-1479).

    ``cube_local_max`` may be the dense cube, as in the reference, or a
    :class:`LocalExtrema` (then ``cube_local_min`` is ignored).  Returns
    ``(threshold, table)``.

    Multi-GPU: every rank passes the extrema of the voxels it owns (global
    shape, global linear indices) and the global ``segmap``; ``allreduce`` is
    an object with ``sum(int64 array)``, ``max(float64 array)`` and
    ``max_image(float32 image)`` (see :mod:`origin_b200.distributed`) that
    combines the tile-local statistics and the per-threshold counts — the
    purity "histograms" — across ranks, so every rank returns the same result.
    """
    ext = _as_extrema(cube_local_max, cube_local_min)
    backend = _backend or _DeviceCounter(ctx)
    nz, ny, nx = ext.shape
    vol = nz * ny * nx
    l1 = ny * nx                                           # lib_origin.py:1424
    segmask = None
    if segmap is not None:
        segmap = segmap.detach().cpu().numpy() if _is_torch(segmap) else np.asarray(segmap)
        segmask = segmap != 0                              # complement of :1428
        l0 = int(np.count_nonzero(~segmask))               # :1431
    else:
        l0 = l1
    n_max, n_min = ext.counts
    if allreduce is not None:
        n_max, n_min = (int(v) for v in allreduce.sum(np.array([n_max, n_min], dtype=np.int64)))
    if threshlist is None:
        mx_max, mx_min, spmax = backend.stats(ext, segmask)
        if allreduce is not None:
            mx_max, mx_min = (float(v) for v in allreduce.max(np.array([mx_max, mx_min], dtype=np.float64)))
            spmax = allreduce.max_image(spmax)
        # the dense cubes hold 0 wherever a voxel is not an extremum (:1247, :1254)
        if n_max < vol:
            mx_max = max(mx_max, 0.0)
        if n_min < vol or segmask is not None:
            mx_min = max(mx_min, 0.0)
        threshmax = min(mx_min, mx_max)                    # :1437
        threshmin = float(np.median(np.asarray(spmax, dtype=np.float64))) * 1.1   # :1438
        threshlist = np.linspace(threshmin, threshmax, 50)  # :1439
    else:
        threshlist = np.asarray(threshlist, dtype=np.float64)
        threshmin = float(np.min(threshlist)) if threshlist.size else 0.0   # :1441
    # the reference prefilters both cubes with "> threshmin" (:1443-1444), so a threshold below threshmin
    # (only possible for a decreasing default list, threshmax < threshmin) counts as threshmin
    tcount = np.maximum(threshlist, threshmin)
    neg = tcount < 0
    if neg.any() and segmask is not None:
        # background-only minima (cube_local_min * segmask zeroes the others, :1429): one extra threshold
        tcount_ext = np.concatenate([tcount, [-np.inf]])
    else:
        tcount_ext = tcount
    n1, n0 = backend.counts(ext, segmask, tcount_ext)
    check_counts(n1, n0)
    if allreduce is not None:
        both = allreduce.sum(np.concatenate([n1, n0]).astype(np.int64))
        n1, n0 = both[:len(n1)], both[len(n1):]
    if len(tcount_ext) != len(tcount):
        n_min = int(n0[-1])                                # minima lying in the background
        n1, n0 = n1[:-1], n0[:-1]
    # zeros of the dense cubes count for negative thresholds
    if neg.any():
        n1 = n1 + neg * (vol - n_max)
        n0 = n0 + neg * (vol - n_min)
    n0 = n0 * (l1 / l0)                                    # :1451
    with np.errstate(divide='ignore', invalid='ignore'):
        est_purity = 1 - n0 / n1                           # :1453
    order = np.argsort(threshlist, kind='stable')          # res.sort('Tval_r'), :1460
    table = PurityTable(Tval_r=np.asarray(threshlist, dtype=np.float64)[order], Pval_r=est_purity[order],
                        Det_m=n0.astype(int)[order], Det_M=n1[order])
    if est_purity[-1] < purity:                            # :1463
        threshold = np.inf
    else:
        threshold = np.interp(purity, table['Pval_r'], table['Tval_r'])   # :1469
    return float(threshold), table


def threshold_rows(ext, threshold, profile=None, which='max', ctx=None):
    """Rows of the raw detection catalogue (reference steps.py:956-964 for
    ``which='max'``, :966-974 on the std cube, :935-939 for ``'min'``): entries
    above ``threshold`` in C order -> dict ``x0, y0, z0, value[, profile]``."""
    idx = ext.max_index if which == 'max' else ext.min_index
    val = ext.max_value if which == 'max' else ext.min_value
    n = len(idx)
    if profile is not None and tuple(profile.shape) != tuple(ext.shape):
        # e.g. a tile's sub-cube profile with the whole-field indices of tile mode: the lookup would run off the end
        raise ValueError('profile has shape %r but the extrema index a %r cube' % (tuple(profile.shape), tuple(ext.shape)))
    ctx = _ctx_for(val, ctx)
    dev_profile = profile if (profile is not None and _is_torch(profile) and profile.is_cuda) else None
    cap = max(1024, n // 64)
    count = np.zeros(1, dtype=np.int64)
    while True:
        oi = np.empty(cap, dtype=np.int64)
        ov = np.empty(cap, dtype=np.float32)
        op = np.empty(cap, dtype=np.uint8) if dev_profile is not None else None
        rc = ctx.check(ctx.lib.ogn_threshold_extract(ctx.handle, ptr(idx), ptr(val), n, float(threshold),
                                                     ptr(dev_profile), ptr(oi), ptr(ov), ptr(op), cap, ptr(count)),
                       allow_overflow=True)
        if rc == 0:
            break
        cap = int(count[0])
    m = int(count[0])
    oi, ov = oi[:m], ov[:m]
    z, y, x = np.unravel_index(oi, ext.shape)
    rows = dict(x0=x, y0=y, z0=z, value=ov)
    if profile is not None:
        if dev_profile is not None:
            rows['profile'] = op[:m]
        else:
            rows['profile'] = np.asarray(profile)[z, y, x]
    return rows


# --------------------------------------------------------------------------
# step08: line estimation
# --------------------------------------------------------------------------

def line_estimates(raw, var, psf, centres, order_dct=30, ctx=None, coef=None):
    """``method_PCA_wgt`` (reference lib_origin.py:1535-1617) on the ``P x P x nz`` windows of ``raw`` / ``var``
    centred at ``centres`` (``(npos, 2)`` array of ``(y, x)``): returns ``(line, linevar)``, float64 arrays
    ``(npos, nz)``.  One batched device call for all windows.

    Weighted mosaics: ``psf`` is ``(nf, nz, P, P)`` (the fields' FSFs) and ``coef`` ``(npos, nf, P, P)`` holds,
    per window, the factors that combine them, ``sum_f coef[p, f] * psf[f]`` — the weight maps cut to the window,
    as ``GridAnalysis`` forms them (:1713-1717); :func:`estimation_line` builds it."""
    raw = _as_float_cube(raw)
    v = _as_float_cube(var)
    if _dtype_code(v) != _dtype_code(raw):
        v = v.to(raw.dtype) if _is_torch(v) else v.astype(raw.dtype)
    nz, ny, nx = raw.shape
    psf = _f64_any(psf)
    cen = np.ascontiguousarray(centres, dtype=np.int32).reshape(-1, 2)
    if coef is None:
        if psf.ndim != 3 or psf.shape[0] != nz or psf.shape[1] != psf.shape[2]:
            raise ValueError('psf must be (nz, P, P) for one field; pass (nf, nz, P, P) with coef for a weighted mosaic')
    else:
        if psf.ndim != 4 or psf.shape[1] != nz or psf.shape[2] != psf.shape[3]:
            raise ValueError('psf must be (nf, nz, P, P) when coef is given')
        coef = _f64_any(coef)
        if tuple(coef.shape) != (len(cen), psf.shape[0], psf.shape[2], psf.shape[3]):
            raise ValueError('coef must be (npos, nf, P, P)')
    ctx = _ctx_for(raw, ctx)
    line = np.empty((len(cen), nz), dtype=np.float64)
    lvar = np.empty((len(cen), nz), dtype=np.float64)
    info = np.zeros(2, dtype=np.int32)
    order = -1 if order_dct is None else int(order_dct)
    if coef is None:
        ctx.check(ctx.lib.ogn_line_estimates(ctx.handle, ptr(raw), ptr(v), _dtype_code(raw), nz, ny, nx, ptr(psf), psf.shape[1],
                                             ptr(cen), len(cen), order, ptr(line), ptr(lvar), ptr(info)))
    else:
        ctx.check(ctx.lib.ogn_line_estimates_fields(ctx.handle, ptr(raw), ptr(v), _dtype_code(raw), nz, ny, nx, ptr(psf),
                                                    psf.shape[0], psf.shape[2], ptr(coef), ptr(cen), len(cen), order,
                                                    ptr(line), ptr(lvar), ptr(info)))
    return line, lvar


def peakdet(v):
    """Local maximum closest to the centre of ``v`` (reference lib_origin.py:1793-1801)."""
    ind = np.where((v[1:-1] > v[:-2]) & (v[1:-1] > v[2:]))[0] + 1
    imax = v.size // 2
    if len(ind) > 0:
        imax = ind[np.argmin((ind - imax) ** 2)]
    return imax


def _grid_criteria(lines, lvars, offsets, cut, psf_core, nl, y0, x0, z0, size_grid, horiz, horiz_psf, criteria):
    """The scalar part of ``GridAnalysis`` (reference lib_origin.py:1680-1790) for one detection, given the line
    estimates of its grid offsets.  ``cut(zsel, dy, dx)`` returns the ``(len(zsel), 2 horiz_psf + 1, 2 horiz_psf + 1)``
    values of the raw cube around the centre of the window at offset ``(dy, dx)``; ``psf_core(k, zsel)`` the same
    cut of the FSF the ``k``-th offset was estimated with (one FSF for a single field, the weighted combination of
    that window for a mosaic)."""
    if criteria not in ('flux', 'mse'):
        raise ValueError('Bad criteria: (flux) or (mse)')
    shape = (1 + 2 * size_grid, 1 + 2 * size_grid)
    zest = np.zeros(shape)
    fest_00 = np.zeros(shape)
    mse = np.full(shape, np.inf)
    fest_05 = np.zeros(shape)
    mse_5 = np.full(shape, np.inf)
    ind_max = slice(max(0, z0 - 5), min(nl, z0 + 6))
    kept = {}
    zidx = np.arange(nl)
    skip_col = set()
    for k, (dy, dx) in enumerate(offsets):          # reference order: dx outer, dy inner (:1701-1702)
        if dx in skip_col:
            continue
        deconv_met, varest_met = lines[k], lvars[k]
        z_est = peakdet(deconv_met[ind_max])
        if z_est == 0:                               # `break` leaves the dy loop of this dx (:1716-1717)
            skip_col.add(dx)
            continue
        maxz = z0 - 5 + z_est
        zest[dy, dx] = maxz
        kept[(dy, dx)] = k
        ind_hrz = zidx[slice(maxz - horiz, maxz + horiz + 1)]      # python slice semantics, as the reference
        if criteria == 'mse':
            lc = psf_core(k, ind_hrz) * deconv_met[ind_hrz][:, None, None]
            r1 = cut(ind_hrz, dy, dx)
            mse[dy, dx] = np.sum((r1 - lc) ** 2) / np.sum(r1 ** 2)
        ind_z5 = np.arange(max(0, maxz - 5), min(maxz + 6, nl))
        lc = psf_core(k, ind_z5) * deconv_met[ind_z5][:, None, None]
        r1 = cut(ind_z5, dy, dx)
        with np.errstate(divide='ignore', invalid='ignore'):
            mse_5[dy, dx] = np.sum((r1 - lc) ** 2) / np.sum(r1 ** 2)
        if criteria == 'flux':
            fest_00[dy, dx] = np.sum(deconv_met[ind_hrz])
        fest_05[dy, dx] = np.sum(deconv_met[ind_z5])
    if criteria == 'flux':
        wy, wx = np.where(fest_00 == fest_00.max())
    else:
        wy, wx = np.where(mse == mse.min())
    if len(wx) == 0 or len(wy) == 0:
        return 0.0, 1.0e6, np.array([0]), np.array([0]), y0, x0, z0
    wy, wx = int(wy[0]), int(wx[0])
    k = kept.get((wy, wx))
    est = lines[k] if k is not None else np.zeros(nl)
    evar = lvars[k] if k is not None else np.zeros(nl)
    return (float(fest_05[wy, wx]), float(mse_5[wy, wx]), est, evar, int(y0 - size_grid + wy), int(x0 - size_grid + wx),
            int(zest[wy, wx]))


def _window_weights(wght, y0, x0, P, size_grid, offs, ny, nx):
    """Per-offset factors of the fields' FSFs for one detection of a weighted mosaic, ``(len(offs), nf, P, P)``,
    restating what the reference's loops do with ``wght`` (lib_origin.py:1899-1906 and :1713-1717):

    * ``estimation_line`` cuts every weight map to the detection's padded minicube (zeros outside the image) and
      keeps the fields whose cut is not empty; a dropped field has factor 0 here;
    * at the FIRST grid offset ``GridAnalysis`` forms ``psf = sum_f wgt_f * psf_f`` with the maps cut to that
      window — and assigns it to the variable the next offsets read, so from the second offset on it computes
      ``sum_f wgt_f * psf`` with the PREVIOUS combination: the factors of offset k are those of the first one
      times ``sum_f wgt_f`` of every later offset up to k.  (The ``break`` at :1716 cannot change which offsets
      take part: it fires only when the +-5 wavelength window around z0 holds at most one sample, which is the
      same for all offsets of a detection.)  With the step's default ``grid_dxy = 0`` there is one offset."""
    side = P + 2 * size_grid
    half = side // 2
    ya, yb, xa, xb = max(0, y0 - half), min(ny, y0 + half + 1), max(0, x0 - half), min(nx, x0 + half + 1)
    red = np.zeros((len(wght), side, side))
    for f, w in enumerate(wght):
        w = np.asarray(w, dtype=np.float64)
        if ya < yb and xa < xb and np.sum(w[ya:yb, xa:xb]) > 0:                           # :1901
            red[f, ya - (y0 - half):yb - (y0 - half), xa - (x0 - half):xb - (x0 - half)] = w[ya:yb, xa:xb]
    coef = np.zeros((len(offs), len(wght), P, P))
    for k, (dy, dx) in enumerate(offs):
        wk = red[:, dy:dy + P, dx:dx + P]
        coef[k] = wk if k == 0 else coef[k - 1] * wk.sum(axis=0)[None]
    return coef


def estimation_line(Cat1, raw, var, psf, wght=None, wcs=None, wave=None, size_grid=1, criteria='flux', order_dct=30,
                    horiz_psf=1, horiz=5, ctx=None, _backend=None):
    """Estimated emission line and re-estimated position of every detection (reference ``estimation_line``,
    lib_origin.py:1805-1938, which calls ``GridAnalysis`` :1620-1790 per detection).  All the windows of all
    detections go through ONE batched device call (:func:`line_estimates`); the scalar criteria run on the host.

    ``Cat1`` needs columns ``x0, y0, z0`` (an astropy Table or a dict of arrays).  ``psf`` / ``wght`` as in the
    reference: one ``(nz, P, P)`` FSF and None, or a list of FSFs with the list of the fields' weight maps
    (:func:`_window_weights`).  Returns ``(cat2, lin_est, var_est)`` where ``cat2`` is a dict with the reference's
    added columns ``x, y, z, residual, flux, num_line`` (plus ``ra, dec, lbda`` when ``wcs`` / ``wave`` are given)
    next to the input columns.  ``_backend`` replaces :func:`line_estimates` (CPU tests of this host logic)."""
    nz, ny, nx = raw.shape
    zs, ys, xs = (np.asarray(Cat1[k], dtype=int) for k in ('z0', 'y0', 'x0'))
    to_host = lambda a: a.detach().cpu().numpy().astype(np.float64) if _is_torch(a) else np.asarray(a, dtype=np.float64)
    if wght is None:
        psf_h = to_host(psf)                                                            # (nz, P, P)
        P = psf_h.shape[1]
    else:
        psf_h = np.stack([to_host(f) for f in psf])                                     # (nf, nz, P, P)
        if len(wght) != len(psf_h):
            raise ValueError('psf and wght must have the same length')
        P = psf_h.shape[2]
    centres, owner, coefs = [], [], []
    for d, (y0, x0) in enumerate(zip(ys, xs)):
        dxl = [dx for dx in range(1 + 2 * size_grid) if 0 <= x0 + dx - size_grid < nx]      # :1696-1699
        dyl = [dy for dy in range(1 + 2 * size_grid) if 0 <= y0 + dy - size_grid < ny]
        offs = [(dy, dx) for dx in dxl for dy in dyl]
        owner.append((len(centres), offs))
        centres += [(y0 + dy - size_grid, x0 + dx - size_grid) for dy, dx in offs]
        if wght is not None and offs:
            coefs.append(_window_weights(wght, int(y0), int(x0), P, size_grid, offs, ny, nx))
    coef = np.concatenate(coefs) if coefs else None
    if centres:
        lines, lvars = (_backend or line_estimates)(raw, var, psf_h, np.array(centres, dtype=np.int32), order_dct, ctx, coef)
    else:
        lines = lvars = np.zeros((0, nz))
    hp = horiz_psf
    inds = slice(P // 2 - hp, P // 2 + 1 + hp)

    def make_cut(y0, x0):
        def cut(zsel, dy, dx):
            cy, cx = y0 + dy - size_grid, x0 + dx - size_grid
            out = np.zeros((len(zsel), 2 * hp + 1, 2 * hp + 1))
            ya, yb, xa, xb = max(0, cy - hp), min(ny, cy + hp + 1), max(0, cx - hp), min(nx, cx + hp + 1)
            if ya < yb and xa < xb and len(zsel):
                blk = raw[zsel[0]:zsel[-1] + 1, ya:yb, xa:xb] if np.all(np.diff(zsel) == 1) else raw[zsel][:, ya:yb, xa:xb]
                blk = blk.detach().cpu().numpy() if _is_torch(blk) else np.asarray(blk)
                out[:, ya - (cy - hp):yb - (cy - hp), xa - (cx - hp):xb - (cx - hp)] = blk
            return out
        return cut

    def make_psf_core(first):
        if coef is None:
            return lambda k, zsel: psf_h[zsel][:, inds, inds]
        return lambda k, zsel: np.sum(coef[first + k][:, None, inds, inds] * psf_h[:, zsel][:, :, inds, inds], axis=0)

    res = []
    for d, (first, offs) in enumerate(owner):
        sl = slice(first, first + len(offs))
        res.append(_grid_criteria(lines[sl], lvars[sl], offs, make_cut(int(ys[d]), int(xs[d])), make_psf_core(first), nz,
                                  int(ys[d]), int(xs[d]), int(zs[d]), size_grid, horiz, horiz_psf, criteria))
    cat2 = {k: np.asarray(Cat1[k]) for k in (Cat1.colnames if hasattr(Cat1, 'colnames') else Cat1.keys())}
    if res:
        flux5, res5, lin_est, var_est, yg, xg, zg = zip(*res)
    else:
        flux5 = res5 = yg = xg = zg = ()
        lin_est = var_est = []
    cat2.update(x=np.array(xg, dtype=int), y=np.array(yg, dtype=int), z=np.array(zg, dtype=int),
                residual=np.array(res5, dtype=float), flux=np.array(flux5, dtype=float),
                num_line=np.arange(1, len(res) + 1))
    if wcs is not None and len(res):
        dec, ra = wcs.pix2sky(np.stack((yg, xg)).T).T
        cat2['ra'], cat2['dec'] = ra, dec
    if wave is not None and len(res):
        cat2['lbda'] = wave.coord(np.array(zg))
    return cat2, list(lin_est), list(var_est)
