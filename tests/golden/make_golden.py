"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs ``/root/reference``):

    python tests/golden/make_golden.py

Every ``*.npz`` written here holds seeded inputs and the outputs of the
reference's own functions (``lib_origin.dct_residual``, ``Correlation_GLR_test``,
``compute_local_max``, ``Compute_threshold_purity``, ``DCTMAT``, ``O2test``, and for
``pca.npz`` / ``lines.npz`` — ``python tests/golden/make_golden.py next`` writes only
those two — ``Compute_GreedyPCA_area`` and ``GridAnalysis``)
loaded by ``reference_loader.load_lib_origin``.  The step glue that lives in
``steps.py`` (which cannot be imported without mpdaf) is replayed by
``_step01_glue`` / ``_step05_glue`` below, line-referenced.
The fixtures pin ``oracle/origin_oracle.py`` (tests/test_oracle_golden.py) and,
on the GPU box, the CUDA path (tests/test_gpu_golden.py).
"""

import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from origin_b200 import synthetic  # noqa: E402
from origin_b200.fitsmini import read_hdus  # noqa: E402
from reference_loader import REFERENCE_ROOT, load_lib_origin  # noqa: E402


def save(name, **arrays):
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **arrays)
    print('%-24s %8.1f KB' % (name, os.path.getsize(path) / 1024))


def table_cols(tab):
    return {k: np.asarray(tab[k].data) for k in ('Tval_r', 'Pval_r', 'Det_m', 'Det_M')}


def _step01_glue(lib, raw, var, mask, order, approx, size):
    """steps.py:431-465 with the reference functions doing the work."""
    cont = lib.dct_residual(raw, order, var, approx, mask)          # :431
    data = raw - cont                                               # :434
    data[mask] = np.nan                                             # :435
    std = np.sqrt(var)                                              # :439
    cont /= std                                                     # :440
    with warnings.catch_warnings():
        warnings.simplefilter('ignore', RuntimeWarning)
        mean = np.nanmean(data, axis=(1, 2))                        # :442
    data -= mean[:, None, None]                                     # :444
    data /= std                                                     # :445
    data[mask] = 0                                                  # :446
    lmax, lmin = lib.compute_local_max(data, data, mask, size)      # :453
    cont32 = cont.astype(np.float32)                                # :463
    return dict(cube_std=data, ima_std=data.mean(axis=0), cube_std_local_max=lmax,
                cube_std_local_min=lmin, cont_dct=cont32, ima_dct=cont32.mean(axis=0),
                cont_sumsq=np.sum(cont32 ** 2, axis=0),            # :472 (inside log10)
                o2map=lib.O2test(data))                             # :480


def _step05_glue(lib, cube, fsf, weights, profiles, mask, size, pcut, pmeansub):
    """steps.py:770-802 with the reference functions doing the work."""
    correl, profile, correl_min = lib.Correlation_GLR_test(
        cube, fsf, weights, profiles, nthreads=1, pcut=pcut, pmeansub=pmeansub)  # :770
    correl[mask] = 0                                                # :781
    profile[mask] = 0                                               # :788
    maxmap = np.amax(correl, axis=0)                                # :792
    minmap = np.amin(correl_min, axis=0)                            # :793
    lmax, lmin = lib.compute_local_max(correl, correl_min, mask, size)  # :796
    return dict(cube_correl=correl, cube_correl_min=correl_min, cube_profile=profile,
                maxmap=maxmap, minmap=minmap, cube_local_max=lmax, cube_local_min=lmin)


def main():
    lib = load_lib_origin()
    warnings.simplefilter('ignore', DeprecationWarning)

    # ---- dictionaries and the segmap shipped with the reference ----------
    d212 = read_hdus(os.path.join(REFERENCE_ROOT, 'muse_origin', 'Dico_FWHM_2_12.fits'))[1:]
    d3 = read_hdus(os.path.join(REFERENCE_ROOT, 'muse_origin', 'Dico_3FWHM.fits'))[1:]
    seg = read_hdus(os.path.join(REFERENCE_ROOT, 'tests', 'segmap.fits'))
    segmap = next(d for _, d in seg if d is not None)
    save('dictionaries',
         dico_2_12=np.stack([d for _, d in d212]), fwhm_2_12=np.array([h['FWHM'] for h, _ in d212]),
         dico_3=np.stack([d for _, d in d3]), fwhm_3=np.array([h['FWHM'] for h, _ in d3]),
         segmap=segmap.astype(np.int16))
    prof3 = [np.array(d) for _, d in d3]
    prof20 = [np.array(d) for _, d in d212]

    # ---- DCTMAT ----------------------------------------------------------
    save('dctmat', d0_3681_10=lib.DCTMAT(3681, 10), d0_150_4=lib.DCTMAT(150, 4))

    # ---- step05, single field, Dico_3FWHM, pcut=1e-8 ---------------------
    shape = (96, 34, 40)
    fsf = synthetic.moffat_fsf(shape[0])
    cube, _ = synthetic.faint_cube(shape, fsf, n_src=6, seed=0)
    mask = synthetic.footprint_mask(shape, seed=0)
    out = _step05_glue(lib, cube, fsf, None, prof3, mask, 3, 1e-8, True)
    correl_u, profile_u, correl_min_u = lib.Correlation_GLR_test(
        cube, fsf, None, prof3, nthreads=1, pcut=1e-8, pmeansub=True)
    save('tglr_single', cube=cube, fsf=fsf, mask=np.packbits(mask), shape=np.array(shape),
         correl_unmasked=correl_u, profile_unmasked=profile_u, **out)

    # ---- step06 on those extrema ------------------------------------------
    seg_small = (segmap[:shape[1], :shape[2]] > 0).astype(np.int16) * 3
    thr_a, tab_a = lib.Compute_threshold_purity(0.8, out['cube_local_max'], out['cube_local_min'], seg_small)
    thr_b, tab_b = lib.Compute_threshold_purity(0.9, out['cube_local_max'], out['cube_local_min'])
    tl = np.linspace(2.0, 7.0, 11)
    thr_c, tab_c = lib.Compute_threshold_purity(0.5, out['cube_local_max'], out['cube_local_min'],
                                                seg_small, threshlist=tl)
    save('purity', segmap=seg_small, threshlist_c=tl,
         thr_a=thr_a, thr_b=thr_b, thr_c=thr_c,
         **{'a_' + k: v for k, v in table_cols(tab_a).items()},
         **{'b_' + k: v for k, v in table_cols(tab_b).items()},
         **{'c_' + k: v for k, v in table_cols(tab_c).items()})

    # ---- step05, full dictionary ------------------------------------------
    shape = (80, 30, 32)
    fsf = synthetic.moffat_fsf(shape[0])
    cube, _ = synthetic.faint_cube(shape, fsf, n_src=5, seed=3)
    c, p, cm = lib.Correlation_GLR_test(cube, fsf, None, prof20, nthreads=1, pcut=1e-8, pmeansub=True)
    save('tglr_2_12', cube=cube, fsf=fsf, correl=c, profile=p, correl_min=cm)

    # ---- step05, no cut (201-sample profiles longer than the cube), no mean sub
    shape = (64, 28, 29)
    fsf = synthetic.moffat_fsf(shape[0])
    cube, _ = synthetic.faint_cube(shape, fsf, n_src=3, seed=4)
    c, p, cm = lib.Correlation_GLR_test(cube, fsf, None, prof3, nthreads=1, pcut=None, pmeansub=False)
    save('tglr_nocut', cube=cube, fsf=fsf, correl=c, profile=p, correl_min=cm)

    # ---- step05, image smaller than the FSF --------------------------------
    shape = (40, 9, 17)
    fsf = synthetic.moffat_fsf(shape[0])
    cube, _ = synthetic.faint_cube(shape, fsf, n_src=2, seed=5)
    c, p, cm = lib.Correlation_GLR_test(cube, fsf, None, prof3, nthreads=1, pcut=1e-8, pmeansub=True)
    save('tglr_tiny', cube=cube, fsf=fsf, correl=c, profile=p, correl_min=cm)

    # ---- step05, two fields with weight maps (mosaic) ----------------------
    shape = (72, 30, 40)
    fsf0 = synthetic.moffat_fsf(shape[0])
    fsf1 = synthetic.moffat_fsf(shape[0], fwhm0=4.2, fwhm1=3.1)
    cube, _ = synthetic.faint_cube(shape, fsf0, n_src=5, seed=6)
    w = synthetic.field_weights(shape[1], shape[2], 2)
    w[0][:, :6] = 0.0
    w[1][:, :6] = 0.0                                   # an uncovered strip (total weight 0)
    c, p, cm = lib.Correlation_GLR_test(cube, [fsf0, fsf1], w, prof3, nthreads=1, pcut=1e-8, pmeansub=True)
    save('tglr_multifield', cube=cube, fsf0=fsf0, fsf1=fsf1, w0=w[0], w1=w[1],
         correl=c, profile=p, correl_min=cm)

    # ---- step01: dct_residual (both branches) + glue -------------------------
    shape = (150, 12, 14)
    fsf = synthetic.moffat_fsf(shape[0], size=9)
    raw, var, mask = synthetic.raw_cube(shape, fsf, n_cont=4, n_src=3, seed=1)
    cont_w = lib.dct_residual(raw, 10, var, False, mask)
    cont_a = lib.dct_residual(raw, 10, var, True, mask)
    cont_4 = lib.dct_residual(raw, 4, var, False, mask)
    g = _step01_glue(lib, raw, var, mask, 10, False, 3)
    save('dct', raw=raw, var=var, mask=np.packbits(mask), shape=np.array(shape),
         cont_weighted=cont_w, cont_approx=cont_a, cont_order4=cont_4, **g)

    # ---- chain: step01 -> step05 -> step06 -> step07 on one cube ------------
    shape = (220, 30, 36)
    fsf = synthetic.moffat_fsf(shape[0])
    mask = synthetic.footprint_mask(shape, seed=2)
    raw, var, mask = synthetic.raw_cube(shape, fsf, n_cont=3, n_src=10, seed=2, mask=mask)
    g1 = _step01_glue(lib, raw, var, mask, 10, False, 3)
    g5 = _step05_glue(lib, g1['cube_std'], fsf, None, prof3, mask, 3, 1e-8, True)
    seg_small = (segmap[:shape[1], :shape[2]] > 0).astype(np.int16)
    thr, tab = lib.Compute_threshold_purity(0.8, g5['cube_local_max'], g5['cube_local_min'], seg_small)
    thr_std, tab_std = lib.Compute_threshold_purity(0.8, g1['cube_std_local_max'], g1['cube_std_local_min'])
    use_thr = thr if np.isfinite(thr) else float(tab['Tval_r'].data[len(tab['Tval_r'].data) // 2])
    use_std = thr_std if np.isfinite(thr_std) else float(tab_std['Tval_r'].data[len(tab_std['Tval_r'].data) // 2])
    z, y, x = np.where(g5['cube_local_max'] > use_thr)                      # steps.py:958
    zs, ys, xs = np.where(g1['cube_std_local_max'] > use_std)              # steps.py:968
    zm, ym, xm = np.where(g5['cube_local_min'] > use_thr)                  # steps.py:938
    save('chain', raw=raw.astype(np.float32), var=var.astype(np.float32), mask=np.packbits(mask),
         shape=np.array(shape), fsf=fsf, segmap=seg_small,
         thr=thr, thr_std=thr_std, use_thr=use_thr, use_std=use_std,
         **{'tab_' + k: v for k, v in table_cols(tab).items()},
         **{'tabstd_' + k: v for k, v in table_cols(tab_std).items()},
         cat_z=z, cat_y=y, cat_x=x, cat_tglr=g5['cube_local_max'][z, y, x],
         cat_profile=g5['cube_profile'][z, y, x],
         std_z=zs, std_y=ys, std_x=xs, std_val=g1['cube_std_local_max'][zs, ys, xs],
         min_z=zm, min_y=ym, min_x=xm,
         maxmap=g5['maxmap'], minmap=g5['minmap'], ima_std=g1['ima_std'], ima_dct=g1['ima_dct'],
         n_local_max=np.count_nonzero(g5['cube_local_max']),
         n_local_min=np.count_nonzero(g5['cube_local_min']),
         correl_sum=g5['cube_correl'].sum(), correl_min_sum=g5['cube_correl_min'].sum(),
         correl_plane=g5['cube_correl'][100], correl_min_plane=g5['cube_correl_min'][100],
         cube_std_plane=g1['cube_std'][100])


def next_rows_inputs():
    """Seeded inputs of the step04 / step08 fixtures, stored as float32 in the fixtures and
    fed to the reference as the float64 values of those float32 numbers."""
    # step04: a standardised cube with continuum residuals left in it, two areas
    rng = np.random.default_rng(21)
    shape = (100, 20, 24)
    cube = rng.standard_normal(shape).astype(np.float32)
    lam = np.linspace(0.0, 1.0, shape[0])
    for y0, x0, amp, k in ((5, 6, 3.0, 2.0), (14, 17, 4.0, 3.5), (10, 11, 2.5, 5.0)):
        yy, xx = np.mgrid[:shape[1], :shape[2]]
        foot = np.exp(-0.5 * ((yy - y0) ** 2 + (xx - x0) ** 2) / 2.0 ** 2)
        cube += (amp * (0.6 + 0.4 * np.cos(k * lam))[:, None, None] * foot[None]).astype(np.float32)
    cube[:, 0, :3] = 0.0                                   # fully masked spaxels: test = 0 (the index quirk of :895-903)
    areamap = np.ones(shape[1:], dtype=np.int16)
    areamap[:, 12:] = 2
    # step08: a three-field mosaic with an uncovered strip and a field that only touches one corner
    rng = np.random.default_rng(7)
    shape8 = (90, 34, 38)
    nz, ny, nx = shape8
    fsf_a = synthetic.moffat_fsf(nz)
    fsf_b = synthetic.moffat_fsf(nz, fwhm0=4.4, fwhm1=3.5)
    raw = rng.normal(size=shape8)
    var = rng.uniform(0.5, 2.0, size=shape8)
    dets = [(45, 17, 19), (20, 2, 35), (70, 32, 1), (30, 10, 8)]
    for z0, y0, x0 in dets:
        zz = np.arange(max(0, z0 - 8), min(nz, z0 + 9))
        line = 30.0 * np.exp(-0.5 * ((zz - z0) / 1.8) ** 2)
        ya, yb, xa, xb = max(0, y0 - 12), min(ny, y0 + 13), max(0, x0 - 12), min(nx, x0 + 13)
        raw[zz[0]:zz[-1] + 1, ya:yb, xa:xb] += line[:, None, None] * fsf_a[z0, ya - y0 + 12:yb - y0 + 12, xa - x0 + 12:xb - x0 + 12] * 25
    raw = raw.astype(np.float32)
    var = var.astype(np.float32)
    var[:, 20:22, 30:33] = np.inf                          # masked voxels (origin.py:262-274)
    raw[:, 20:22, 30:33] = 0.0
    yy, xx = np.mgrid[:ny, :nx]
    w1 = np.clip((xx - 4.0) / (nx - 10.0), 0.0, 1.0)
    w0 = 1.0 - w1
    w0[:3] = 0.0
    w1[:3] = 0.0
    w2 = np.zeros((ny, nx))
    w2[28:, 30:] = 0.3
    return dict(cube=cube, areamap=areamap), dict(raw=raw, var=var, fsf=np.stack([fsf_a, fsf_b, fsf_a * 0.9 + fsf_b * 0.1]),
                                                  wght=np.stack([w0, w1, w2]), dets=np.array(dets))


def grid_analysis(lib, raw, var, psf, wght, dets, size_grid, criteria, order_dct, horiz_psf=1, horiz=5):
    """``GridAnalysis`` (lib_origin.py:1620-1790) on the padded minicubes ``estimation_line`` builds per detection
    (:1886-1906; ``overlap_slices`` is astropy's, the padding is restated here)."""
    nz, ny, nx = raw.shape
    P = psf.shape[1] if wght is None else psf[0].shape[1]
    side = P + 2 * size_grid
    half = side // 2
    res = []
    for z, y, x in dets:
        z, y, x = int(z), int(y), int(x)
        red_dat = np.zeros((nz, side, side))
        red_var = np.full((nz, side, side), np.inf)
        ya, yb, xa, xb = max(0, y - half), min(ny, y + half + 1), max(0, x - half), min(nx, x + half + 1)
        dst = (slice(ya - (y - half), yb - (y - half)), slice(xa - (x - half), xb - (x - half)))
        red_dat[(slice(None),) + dst] = raw[:, ya:yb, xa:xb]
        red_var[(slice(None),) + dst] = var[:, ya:yb, xa:xb]
        red_psf, red_wgt = psf, None
        if wght is not None:
            red_psf, red_wgt = [], []
            for n, w in enumerate(wght):
                if np.sum(w[ya:yb, xa:xb]) > 0:                                  # :1901
                    tmp = np.zeros((side, side))
                    tmp[dst] = w[ya:yb, xa:xb]
                    red_wgt.append(tmp)
                    red_psf.append(psf[n])
        with warnings.catch_warnings(), np.errstate(all='ignore'):
            warnings.simplefilter('ignore')
            res.append(lib.GridAnalysis(red_dat, red_var, red_psf, red_wgt, horiz, size_grid, y, x, z, ny, nx, horiz_psf,
                                        criteria, order_dct))
    flux, mse, line, lvar, yy, xx, zz = zip(*res)
    return dict(flux=np.array(flux), mse=np.array(mse), line=np.array(line), lvar=np.array(lvar), y=np.array(yy),
                x=np.array(xx), z=np.array(zz))


def next_rows_outputs(lib):
    """What the unmodified reference computes on those inputs: ``Compute_GreedyPCA_area`` (:769-818) and
    ``GridAnalysis`` for a single field and for the weighted mosaic."""
    pca_in, lines_in = next_rows_inputs()
    cube = pca_in['cube'].astype(np.float64)
    areamap = pca_in['areamap'].astype(int)
    test = [lib.O2test(cube[:, areamap == a]) for a in (1, 2)]
    thr = [float(np.percentile(t[t > 0], 80)) for t in test]
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        faint, map_o2, nstop = lib.Compute_GreedyPCA_area(2, cube, areamap, 50, thr, 100, test)
        faint3, map3, nstop3 = lib.Compute_GreedyPCA_area(2, cube, areamap, 20, thr, 3, test)
    pca = dict(thr=np.array(thr), test0=test[0], test1=test[1], faint=faint, map_o2=map_o2, nstop=nstop,
               faint_itermax3=faint3, map_o2_itermax3=map3, nstop_itermax3=nstop3)
    raw, var = lines_in['raw'].astype(np.float64), lines_in['var'].astype(np.float64)
    fsf, wght, dets = lines_in['fsf'], lines_in['wght'], lines_in['dets']
    lines = {}
    for tag, args in (('single_g1_flux', (fsf[0], None, 1, 'flux', 20)), ('single_g1_mse_pcals', (fsf[0], None, 1, 'mse', None)),
                      ('mosaic_g0_flux', (list(fsf), list(wght), 0, 'flux', 20)),
                      ('mosaic_g1_flux', (list(fsf), list(wght), 1, 'flux', 20))):
        out = grid_analysis(lib, raw, var, args[0], args[1], dets, args[2], args[3], args[4])
        lines.update({tag + '_' + k: v for k, v in out.items()})
    return pca_in, pca, lines_in, lines


def main_next():
    lib = load_lib_origin()
    warnings.simplefilter('ignore', DeprecationWarning)
    pca_in, pca, lines_in, lines = next_rows_outputs(lib)
    save('pca', **pca_in, **pca)
    save('lines', **lines_in, **lines)


if __name__ == '__main__':
    if sys.argv[1:] != ['next']:
        main()
    main_next()
