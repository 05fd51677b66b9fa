"""Parity of the CUDA path (through the C ABI / host mirror) against the
reference-generated golden fixtures and the float64 oracle.

Tolerances (BASELINE.json north_star): ``correl`` and
``correl_min`` within ``1e-5 * max(|ref|, rms(ref))``; argmax profile, extremum
lists and catalogue rows identical except where competing values differ by
less than that tolerance (the number of such ties is asserted to be small and
each one is verified to be a genuine near-tie); integer / index work bit-exact.
"""

import numpy as np
import pytest

from conftest import load_golden, tol_report, unpack_mask
from oracle import origin_oracle as orc
from origin_b200 import dictionaries, synthetic

pytestmark = pytest.mark.gpu

RTOL = 1e-5


@pytest.fixture(scope='module')
def lo():
    from origin_b200 import lib_origin
    return lib_origin


def assert_close(got, ref, what, rtol=RTOL):
    rep = tol_report(got, ref, rtol)
    assert rep['ok'], '%s: max_abs=%.3g worst=%.3g x bound (rms %.3g)' % (what, rep['max_abs'], rep['worst'], rep['rms'])
    return rep


def check_profile(got, ref_profile, tk, what, max_frac=1e-3):
    """argmax must match except at genuine near-ties of the reference's T_k."""
    mism = np.argwhere(got != ref_profile)
    if len(mism) == 0:
        return 0
    rms = float(np.sqrt(np.mean(tk.max(axis=0) ** 2)))
    for idx in mism:
        col = tk[(slice(None),) + tuple(idx)]
        top = np.sort(col)[::-1]
        gap = top[0] - col[got[tuple(idx)]]
        assert gap <= RTOL * max(abs(top[0]), rms), '%s: profile mismatch at %s is not a tie (gap %.3g)' % (what, idx, gap)
    assert len(mism) <= max(2, max_frac * got.size), '%s: %d argmax ties' % (what, len(mism))
    return len(mism)


def oracle_tk(cube, fsf, weights, profiles, pcut, pmeansub):
    cf, nf = orc.fsf_correlate(cube, fsf, weights)
    prof = orc.prepare_profiles(profiles, pcut, pmeansub)
    return orc.spectral_glr_direct(cf, nf, prof)[3]


# --------------------------------------------------------------------------
# step05
# --------------------------------------------------------------------------

def test_fsf_stage_matches_oracle(lo):
    g = load_golden('tglr_single')
    cf, nf = lo.fsf_stage(g['cube'], g['fsf'], None)
    rcf, rnf = orc.fsf_correlate_direct(g['cube'], g['fsf'])
    assert_close(cf, rcf, 'cube_fsf')
    assert_close(nf, rnf, 'norm_fsf')


def test_tglr_single_field_golden(lo):
    g = load_golden('tglr_single')
    mask = unpack_mask(g['mask'], g['shape'])
    profs = dictionaries.dico_3fwhm()[0]
    correl, profile, correl_min = lo.Correlation_GLR_test(g['cube'], g['fsf'], None, profs, pcut=1e-8)
    # the reference's container types (lib_origin.py:1197-1217): float64 cubes, uint8 profile
    assert correl.dtype == np.float64 and correl_min.dtype == np.float64 and profile.dtype == np.uint8
    c32, _, _ = lo.Correlation_GLR_test(g['cube'], g['fsf'], None, profs, pcut=1e-8, out_dtype=np.float32)
    assert c32.dtype == np.float32 and np.array_equal(c32.astype(np.float64), correl)
    assert_close(correl, g['correl_unmasked'], 'correl')
    assert_close(correl_min, g['cube_correl_min'], 'correl_min')
    tk = oracle_tk(g['cube'], g['fsf'], None, profs, 1e-8, True)
    check_profile(profile, g['profile_unmasked'], tk, 'profile')
    # fused step glue: mask, maxmap, minmap (steps.py:781-793)
    out = lo.tglr(g['cube'], g['fsf'], None, profs, mask=mask, pcut=1e-8)
    assert_close(out['correl'], g['cube_correl'], 'masked correl')
    assert np.all(out['correl'][mask] == 0) and np.all(out['profile'][mask] == 0)
    assert_close(out['correl_min'], g['cube_correl_min'], 'correl_min (unmasked by design)')
    assert_close(out['maxmap'], g['maxmap'], 'maxmap')
    assert_close(out['minmap'], g['minmap'], 'minmap')
    np.testing.assert_array_equal(out['maxmap'], out['correl'].max(axis=0))
    np.testing.assert_array_equal(out['minmap'], out['correl_min'].min(axis=0))


@pytest.mark.parametrize('name,pcut,pmeansub,full', [
    ('tglr_2_12', 1e-8, True, True), ('tglr_nocut', None, False, False),
    ('tglr_tiny', 1e-8, True, False)])
def test_tglr_variants_golden(lo, name, pcut, pmeansub, full):
    g = load_golden(name)
    profs = dictionaries.dico_fwhm_2_12()[0] if full else dictionaries.dico_3fwhm()[0]
    correl, profile, correl_min = lo.Correlation_GLR_test(g['cube'], g['fsf'], None, profs, pcut=pcut,
                                                          pmeansub=pmeansub)
    assert_close(correl, g['correl'], name + ' correl')
    assert_close(correl_min, g['correl_min'], name + ' correl_min')
    tk = oracle_tk(g['cube'], g['fsf'], None, profs, pcut, pmeansub)
    check_profile(profile, g['profile'], tk, name + ' profile', max_frac=3e-3)


def test_tglr_multifield_golden(lo):
    g = load_golden('tglr_multifield')
    profs = dictionaries.dico_3fwhm()[0]
    fsf, w = [g['fsf0'], g['fsf1']], [g['w0'], g['w1']]
    correl, profile, correl_min = lo.Correlation_GLR_test(g['cube'], fsf, w, profs, pcut=1e-8)
    # SURVEY.md note N1: where no field covers the footprint the reference holds FFT round-off;
    # compare where the 25x25 footprint sees data (x >= 6 + 12)
    sel = np.zeros(g['cube'].shape[1:], dtype=bool)
    sel[:, 18:] = True
    assert_close(correl[:, sel], g['correl'][:, sel], 'multifield correl')
    assert_close(correl_min[:, sel], g['correl_min'][:, sel], 'multifield correl_min')
    tk = oracle_tk(g['cube'], fsf, w, profs, 1e-8, True)
    check_profile(profile[:, sel], g['profile'][:, sel], tk[:, :, sel], 'multifield profile')


def test_tglr_float64_input_and_odd_width(lo):
    # nx not a multiple of 4 (TMA needs a padded copy) and a float64 cube (reference dtype)
    shape = (50, 31, 37)
    fsf = synthetic.moffat_fsf(shape[0])
    cube, _ = synthetic.faint_cube(shape, fsf, n_src=3, seed=11)
    profs = dictionaries.dico_3fwhm()[0]
    ref = orc.correlation_glr_test(cube, fsf, None, profs, pcut=1e-8)
    for arr in (cube, cube.astype(np.float64)):
        correl, profile, correl_min = lo.Correlation_GLR_test(arr, fsf, None, profs, pcut=1e-8)
        assert_close(correl, ref[0], 'correl')
        assert_close(correl_min, ref[2], 'correl_min')
        assert np.mean(profile == ref[1]) > 0.999


@pytest.mark.parametrize('kind', ['asymmetric', 'one_plane_asymmetric', 'symmetric'])
def test_tglr_fsf_symmetry_paths(lo, kind):
    # K1 folds FSF rows dy and P-1-dy when every plane's float32 weights are mirror-symmetric in y
    # (a Moffat FSF is); an FSF that is not — a sheared / random one, or a single odd plane — must take
    # the general path.  Both are checked against the float64 oracle, which knows nothing of symmetry.
    shape = (48, 70, 72)   # more than one 64x64 K1 tile in both directions
    rng = np.random.default_rng(21)
    fsf = synthetic.moffat_fsf(shape[0])
    if kind == 'asymmetric':
        fsf = fsf * (1.0 + 0.3 * rng.random(fsf.shape))
        fsf /= fsf.sum(axis=(1, 2), keepdims=True)
    elif kind == 'one_plane_asymmetric':
        fsf = fsf.copy()
        fsf[17, 3, 5] *= 1.5
    cube, _ = synthetic.faint_cube(shape, synthetic.moffat_fsf(shape[0]), n_src=3, seed=22)
    cf, nf = lo.fsf_stage(cube, fsf, None)
    rcf, rnf = orc.fsf_correlate_direct(cube, fsf)
    assert_close(cf, rcf, kind + ' cube_fsf')
    assert_close(nf, rnf, kind + ' norm_fsf')
    assert lo.default_context().fsf_folded == (kind == 'symmetric')
    profs = dictionaries.dico_3fwhm()[0]
    ref = orc.correlation_glr_test(cube, fsf, None, profs, pcut=1e-8)
    correl, profile, correl_min = lo.Correlation_GLR_test(cube, fsf, None, profs, pcut=1e-8)
    assert_close(correl, ref[0], kind + ' correl')
    assert_close(correl_min, ref[2], kind + ' correl_min')
    assert np.mean(profile == ref[1]) > 0.999


def test_tglr_small_fsf_fallback(lo):
    # FSF size without a TMA-tiled instantiation (9x9) goes through the generic kernel
    shape = (40, 20, 24)
    fsf = synthetic.moffat_fsf(shape[0], size=9)
    cube, _ = synthetic.faint_cube(shape, fsf, n_src=2, seed=12)
    profs = dictionaries.dico_3fwhm()[0]
    ref = orc.correlation_glr_test(cube, fsf, None, profs, pcut=1e-8)
    correl, profile, correl_min = lo.Correlation_GLR_test(cube, fsf, None, profs, pcut=1e-8)
    assert_close(correl, ref[0], 'correl')
    assert_close(correl_min, ref[2], 'correl_min')


def test_tglr_device_tensors(lo):
    import torch
    g = load_golden('tglr_single')
    profs = dictionaries.dico_3fwhm()[0]
    cube = torch.from_numpy(g['cube']).cuda()
    out = lo.tglr(cube, g['fsf'], None, profs, pcut=1e-8)
    torch.cuda.synchronize()
    assert out['correl'].is_cuda
    assert_close(out['correl'].cpu().numpy(), g['correl_unmasked'], 'device correl')


def test_tglr_errors(lo):
    from origin_b200._lib import OgnError
    profs = dictionaries.dico_3fwhm()[0]
    cube = np.zeros((8, 6, 6), np.float32)
    with pytest.raises(ValueError):
        lo.Correlation_GLR_test(cube, np.zeros((7, 5, 5)), None, profs)          # wrong nz
    with pytest.raises(OgnError):
        lo.Correlation_GLR_test(cube, np.zeros((8, 4, 4)), None, profs)          # even FSF
    with pytest.raises(OgnError):
        lo.Correlation_GLR_test(cube, np.ones((8, 5, 5)), None, profs * 100)     # > 255 profiles


# --------------------------------------------------------------------------
# local extrema, purity, thresholding
# --------------------------------------------------------------------------

def test_local_extrema_bit_exact_on_reference_inputs(lo):
    g = load_golden('tglr_single')
    mask = unpack_mask(g['mask'], g['shape'])
    a = g['cube_correl'].astype(np.float32)
    b = g['cube_correl_min'].astype(np.float32)
    rmax, rmin = orc.compute_local_max(a, b, mask, 3)
    dmax, dmin = lo.compute_local_max(a, b, mask, 3)
    np.testing.assert_array_equal(dmax, rmax.astype(np.float32))
    np.testing.assert_array_equal(dmin, rmin.astype(np.float32))
    ext, _, _ = lo.local_extrema(a, b, mask, 3)
    np.testing.assert_array_equal(ext.max_index, np.flatnonzero(rmax))
    np.testing.assert_array_equal(ext.min_index, np.flatnonzero(rmin))
    assert np.all(np.diff(ext.max_index) > 0) and np.all(np.diff(ext.min_index) > 0)
    np.testing.assert_array_equal(ext.dense('max'), dmax)
    np.testing.assert_array_equal(ext.dense('min'), dmin)
    # overflow path: tiny capacity is grown transparently
    ext2, _, _ = lo.local_extrema(a, b, mask, 3, capacity=8)
    np.testing.assert_array_equal(ext2.max_index, ext.max_index)
    np.testing.assert_array_equal(ext2.min_value, ext.min_value)


@pytest.mark.parametrize('size', [(1, 3, 3), 5, (3, 1, 5)])
def test_local_extrema_other_windows(lo, size):
    rng = np.random.default_rng(5)
    a = rng.standard_normal((20, 17, 45)).astype(np.float32)
    b = rng.standard_normal((20, 17, 45)).astype(np.float32)
    a[3:6, 2:5, 10:20] = 1.5                                   # plateau: every tied voxel is kept
    mask = rng.random(a.shape) < 0.05
    rmax, rmin = orc.compute_local_max(a, b, mask, size)
    dmax, dmin = lo.compute_local_max(a, b, mask, size)
    np.testing.assert_array_equal(dmax, rmax.astype(np.float32))
    np.testing.assert_array_equal(dmin, rmin.astype(np.float32))


def test_local_extrema_same_array_and_ragged_width(lo):
    rng = np.random.default_rng(6)
    a = rng.standard_normal((9, 5, 33)).astype(np.float32)    # nx = 33: two ballot words per row
    mask = np.zeros(a.shape, bool)
    rmax, rmin = orc.compute_local_max(a, a, mask, 3)
    dmax, dmin = lo.compute_local_max(a, a, mask, 3)
    np.testing.assert_array_equal(dmax, rmax.astype(np.float32))
    np.testing.assert_array_equal(dmin, rmin.astype(np.float32))


def test_purity_threshold_golden(lo):
    g = load_golden('purity')
    t = load_golden('tglr_single')
    lmax, lmin = t['cube_local_max'], t['cube_local_min']
    for tag, purity, seg, tl in (('a', 0.8, g['segmap'], None), ('b', 0.9, None, None),
                                 ('c', 0.5, g['segmap'], g['threshlist_c'])):
        thr, tab = lo.Compute_threshold_purity(purity, lmax, lmin, seg, tl)
        # the lists are float32-rounded values of the float64 cubes: thresholds agree to 1e-6
        np.testing.assert_allclose(tab['Tval_r'], g[tag + '_Tval_r'], rtol=2e-6)
        dm = np.abs(tab['Det_M'] - g[tag + '_Det_M'])
        dn = np.abs(tab['Det_m'] - g[tag + '_Det_m'])
        assert dm.max() <= 1 and dn.max() <= 1, (tag, dm.max(), dn.max())
        ref_thr = float(g['thr_' + tag])
        if np.isfinite(ref_thr):
            assert abs(thr - ref_thr) <= 2e-3 * abs(ref_thr), (tag, thr, ref_thr)
        else:
            assert np.isinf(thr)


def test_purity_counts_exact(lo):
    rng = np.random.default_rng(7)
    shape = (30, 12, 14)
    lmax = np.where(rng.random(shape) < 0.02, rng.gamma(2.0, 2.0, shape), 0).astype(np.float32)
    lmin = np.where(rng.random(shape) < 0.02, rng.gamma(2.0, 1.5, shape), 0).astype(np.float32)
    seg = (rng.random(shape[1:]) < 0.3).astype(np.int16) * 4
    thr, tab = lo.Compute_threshold_purity(0.7, lmax, lmin, seg)
    # the oracle gets float64 cubes like the reference does (float32 inputs would make numpy do
    # the 1.1 * median arithmetic in float32)
    rthr, rtab = orc.threshold_purity(0.7, lmax.astype(np.float64), lmin.astype(np.float64), seg)
    np.testing.assert_allclose(tab['Tval_r'], rtab['Tval_r'], rtol=1e-12)
    np.testing.assert_array_equal(tab['Det_M'], rtab['Det_M'])
    np.testing.assert_array_equal(tab['Det_m'], rtab['Det_m'])
    assert thr == pytest.approx(rthr, rel=1e-12) or (np.isinf(thr) and np.isinf(rthr))
    # explicit list with negative thresholds (zeros of the dense cube count)
    tl = np.array([-1.0, 0.5, 3.0])
    thr, tab = lo.Compute_threshold_purity(0.1, lmax, lmin, None, tl)
    rthr, rtab = orc.threshold_purity(0.1, lmax, lmin, None, tl)
    np.testing.assert_array_equal(tab['Det_M'], rtab['Det_M'])
    np.testing.assert_array_equal(tab['Det_m'], rtab['Det_m'])


def test_threshold_rows_order_and_values(lo):
    rng = np.random.default_rng(8)
    shape = (25, 11, 40)
    a = rng.standard_normal(shape).astype(np.float32)
    prof = rng.integers(0, 3, shape).astype(np.uint8)
    mask = np.zeros(shape, bool)
    ext, dmax, _ = lo.local_extrema(a, a, mask, 3, dense=True)
    rows = lo.threshold_rows(ext, 1.2, prof)
    ref = orc.detection_rows(dmax, prof, 1.2)
    for k in ('x0', 'y0', 'z0', 'profile'):
        np.testing.assert_array_equal(rows[k], ref[k])
    np.testing.assert_array_equal(rows['value'], ref['value'])
    rows = lo.threshold_rows(ext, 1e9, prof)
    assert len(rows['x0']) == 0


# --------------------------------------------------------------------------
# step01
# --------------------------------------------------------------------------

def test_dct_residual_golden(lo):
    g = load_golden('dct')
    mask = unpack_mask(g['mask'], g['shape'])
    scale = np.abs(g['cont_weighted']).max()
    for key, order, approx in (('cont_weighted', 10, False), ('cont_approx', 10, True), ('cont_order4', 4, False)):
        cont = lo.dct_residual(g['raw'], order, g['var'], approx, mask)
        assert cont.dtype == np.float64
        assert np.abs(cont - g[key]).max() <= 1e-9 * scale, key


def test_preprocess_golden(lo):
    g = load_golden('dct')
    mask = unpack_mask(g['mask'], g['shape'])
    out = lo.preprocess(g['raw'], g['var'], mask, 10, False)
    assert_close(out['cube_std'], g['cube_std'], 'cube_std')
    assert_close(out['cont_dct'], g['cont_dct'], 'cont_dct')
    for key in ('ima_std', 'ima_dct', 'cont_sumsq', 'o2map'):
        assert_close(out[key], g[key], key)
    ext, _, _ = lo.local_extrema(out['cube_std'], out['cube_std'], mask, 3)
    ref_idx = np.flatnonzero(g['cube_std_local_max'])
    only = np.setxor1d(ext.max_index, ref_idx)
    assert len(only) <= 2, 'std local maxima differ at %d voxels' % len(only)


# --------------------------------------------------------------------------
# chain: step01 -> step05 -> step06 -> step07 on one cube, vs the reference
# --------------------------------------------------------------------------

def test_chain_catalogue_matches_reference(lo):
    g = load_golden('chain')
    shape = tuple(int(s) for s in g['shape'])
    mask = unpack_mask(g['mask'], shape)
    profs = dictionaries.dico_3fwhm()[0]
    s1 = lo.preprocess(g['raw'], g['var'], mask, 10, False)
    assert_close(s1['cube_std'][100], g['cube_std_plane'], 'cube_std plane', rtol=2e-5)
    out = lo.tglr(s1['cube_std'], g['fsf'], None, profs, mask=mask, pcut=1e-8)
    assert_close(out['correl'][100], g['correl_plane'], 'correl plane', rtol=2e-5)
    assert_close(out['maxmap'], g['maxmap'], 'maxmap', rtol=2e-5)
    ext, _, _ = lo.local_extrema(out['correl'], out['correl_min'], mask, 3)
    n1, n0 = ext.counts
    assert abs(n1 - int(g['n_local_max'])) <= 3 and abs(n0 - int(g['n_local_min'])) <= 3
    thr, tab = lo.Compute_threshold_purity(0.8, ext, None, g['segmap'])
    np.testing.assert_allclose(tab['Tval_r'], g['tab_Tval_r'], rtol=1e-4)
    assert np.abs(tab['Det_M'] - g['tab_Det_M']).max() <= 1
    rows = lo.threshold_rows(ext, float(g['use_thr']), out['profile'])
    np.testing.assert_array_equal(rows['z0'], g['cat_z'])
    np.testing.assert_array_equal(rows['y0'], g['cat_y'])
    np.testing.assert_array_equal(rows['x0'], g['cat_x'])
    np.testing.assert_array_equal(rows['profile'], g['cat_profile'])
    np.testing.assert_allclose(rows['value'], g['cat_tglr'], rtol=1e-4)
    ext_std, _, _ = lo.local_extrema(s1['cube_std'], s1['cube_std'], mask, 3)
    rows = lo.threshold_rows(ext_std, float(g['use_std']))
    np.testing.assert_array_equal(rows['z0'], g['std_z'])
    np.testing.assert_array_equal(rows['x0'], g['std_x'])


# --------------------------------------------------------------------------
# larger cubes
# --------------------------------------------------------------------------

def test_tglr_full_lambda_axis_vs_oracle(lo):
    # the BASELINE wavelength axis on a small field; the oracle needs ~10 s
    shape = (3681, 40, 48)
    fsf = synthetic.moffat_fsf(shape[0])
    cube, cat = synthetic.faint_cube(shape, fsf, n_src=8, seed=21)
    profs = dictionaries.dico_3fwhm()[0]
    ref = orc.correlation_glr_test(cube, fsf, None, profs, nthreads=4, pcut=1e-8)
    correl, profile, correl_min = lo.Correlation_GLR_test(cube, fsf, None, profs, pcut=1e-8)
    assert_close(correl, ref[0], 'correl')
    assert_close(correl_min, ref[2], 'correl_min')
    ties = int(np.count_nonzero(profile != ref[1]))
    assert ties <= 1e-4 * profile.size, ties
    # injected emitters are recovered as the global maxima of their neighbourhood
    z, y, x = (int(v) for v in cat[np.argmax(cat[:, 4])][:3])
    zz = slice(max(0, z - 3), z + 4)
    assert correl[zz, max(0, y - 2):y + 3, max(0, x - 2):x + 3].max() > 8


def test_tglr_tile_consistency_and_spot_oracle(lo):
    """Size-independent properties on a bigger field: (1) a sub-tile cut with
    a 12-pixel halo reproduces the full-cube result bit for bit in its
    interior (the multi-GPU decomposition relies on it); (2) sampled voxels
    match the direct-space float64 oracle."""
    shape = (600, 96, 160)
    fsf = synthetic.moffat_fsf(shape[0])
    cube, _ = synthetic.faint_cube(shape, fsf, n_src=10, seed=22)
    profs = dictionaries.dico_fwhm_2_12()[0]
    full = lo.tglr(cube, fsf, None, profs, pcut=1e-8, want=('correl', 'correl_min', 'profile'))
    y0, y1, x0, x1 = 20, 70, 40, 120
    sub = np.ascontiguousarray(cube[:, y0 - 12:y1 + 12, x0 - 12:x1 + 12])
    part = lo.tglr(sub, fsf, None, profs, pcut=1e-8, want=('correl', 'correl_min', 'profile'))
    for k in ('correl', 'correl_min', 'profile'):
        np.testing.assert_array_equal(part[k][:, 12:-12, 12:-12], full[k][:, y0:y1, x0:x1], err_msg=k)
    rng = np.random.default_rng(3)
    pts = [(int(rng.integers(0, shape[0])), int(rng.integers(0, shape[1])), int(rng.integers(0, shape[2])))
           for _ in range(40)] + [(0, 0, 0), (599, 95, 159), (300, 0, 80), (10, 50, 159)]
    prof_cut = orc.prepare_profiles(profs, 1e-8, True)
    tk = orc.spot_tglr(cube, fsf, prof_cut, pts)
    rms = float(np.sqrt(np.mean(full['correl'].astype(np.float64) ** 2)))
    for (z, y, x), t in zip(pts, tk):
        assert abs(full['correl'][z, y, x] - t.max()) <= RTOL * max(abs(t.max()), rms)
        assert abs(full['correl_min'][z, y, x] - t.min()) <= RTOL * max(abs(t.min()), rms)


# --------------------------------------------------------------------------
# tiled and streamed execution
# --------------------------------------------------------------------------

@pytest.mark.parametrize('shape,ntiles,dico', [((300, 100, 150), 4, '3FWHM'), ((200, 120, 176), 2, '3FWHM'),
                                               ((150, 96, 320), 8, '3FWHM'), ((120, 100, 150), 4, '2_12')])
def test_step05_tiles_reproduce_the_full_cube(lo, shape, ntiles, dico):
    """The multi-GPU decomposition run serially on one GPU: tiles with >= 13-pixel halos give
    bit-identical products on their owned windows and the same global extremum lists.  The second case
    splits columns so that a tile's window starts at a column that is not a multiple of 4 (the library
    has to widen it: TMA boxes start on 16-byte boundaries)."""
    from origin_b200 import tiles
    nz, ny, nx = shape
    fsf = synthetic.moffat_fsf(nz)
    cube, _ = synthetic.faint_cube(shape, fsf, n_src=8, seed=31)
    mask = synthetic.footprint_mask(shape, seed=31)
    profs = dictionaries.dico_3fwhm()[0] if dico == '3FWHM' else dictionaries.dico_fwhm_2_12()[0]   # K2 / K2f
    full = lo.step05(cube, fsf, None, profs, mask, 3, 1e-8, True)
    got_max, got_min, val_max = [], [], []
    for t in tiles.plan_tiles(ny, nx, ntiles, 13):
        sl = (slice(None),) + t.padded
        part = lo.step05(np.ascontiguousarray(cube[sl]), fsf, None, profs, np.ascontiguousarray(mask[sl]), 3, 1e-8,
                         True, tile=(t, (ny, nx)))
        ys, xs = t.owned
        gy, gx = t.global_owned
        for k in ('correl', 'correl_min', 'profile'):
            np.testing.assert_array_equal(part[k][:, ys, xs], full[k][:, gy, gx], err_msg=k)
        np.testing.assert_array_equal(part['maxmap'][ys, xs], full['maxmap'][gy, gx])
        np.testing.assert_array_equal(part['minmap'][ys, xs], full['minmap'][gy, gx])
        assert part['extrema'].shape == shape
        got_max.append(part['extrema'].max_index)
        val_max.append(part['extrema'].max_value)
        got_min.append(part['extrema'].min_index)
    order = np.argsort(np.concatenate(got_max))
    np.testing.assert_array_equal(np.concatenate(got_max)[order], full['extrema'].max_index)
    np.testing.assert_array_equal(np.concatenate(val_max)[order], full['extrema'].max_value)
    np.testing.assert_array_equal(np.sort(np.concatenate(got_min)), full['extrema'].min_index)


def test_step05_streamed_host_path_matches_device_path(lo):
    """Host buffers with ny >= 128 take the slab-pipelined path (upload / K1+K2 / download on
    three streams); it must give exactly what the monolithic device path gives."""
    import torch
    shape = (200, 160, 96)
    fsf = synthetic.moffat_fsf(shape[0])
    cube, _ = synthetic.faint_cube(shape, fsf, n_src=6, seed=32)
    mask = synthetic.footprint_mask(shape, seed=32)
    profs = dictionaries.dico_3fwhm()[0]
    host = lo.step05(cube, fsf, None, profs, mask, 3, 1e-8, True)                      # streamed
    dev = lo.step05(torch.from_numpy(cube).cuda(), fsf, None, profs, torch.from_numpy(mask.view(np.uint8)).cuda(),
                    3, 1e-8, True)                                                      # monolithic
    torch.cuda.synchronize()
    for k in ('correl', 'correl_min', 'profile', 'maxmap', 'minmap'):
        np.testing.assert_array_equal(host[k], dev[k].cpu().numpy(), err_msg=k)
    np.testing.assert_array_equal(host['extrema'].max_index, dev['extrema'].max_index.cpu().numpy())
    np.testing.assert_array_equal(host['extrema'].min_value, dev['extrema'].min_value.cpu().numpy())
    # and both agree with the oracle
    ref = orc.tglr_step(cube, fsf, None, profs, mask, 3, 4, 1e-8, True)
    assert_close(host['correl'], ref['cube_correl'], 'streamed correl')


def test_purity_counts_device_mode_matches_host_mode(lo):
    """Device thresholds -> device counts, no host synchronisation (what the multi-GPU step hands to the
    NCCL allreduce); must equal the host-mode counts, also with a segmentation mask."""
    import torch
    rng = np.random.default_rng(41)
    shape = (40, 24, 32)
    a = rng.standard_normal(shape).astype(np.float32)
    ext, _, _ = lo.local_extrema(torch.from_numpy(a).cuda(), torch.from_numpy(-a).cuda(), None, 3)
    seg = (rng.random(shape[1:]) < 0.25).astype(np.uint8)
    thr = np.linspace(0.5, 3.0, 50)
    host = lo.purity_counts(lo.LocalExtrema(shape, ext.max_index.cpu().numpy(), ext.max_value.cpu().numpy(),
                                            ext.min_index.cpu().numpy(), ext.min_value.cpu().numpy()), seg, thr)
    out = torch.full((100,), -7, dtype=torch.int64, device='cuda')
    n1, n0 = lo.purity_counts(ext, seg, torch.from_numpy(thr).cuda(), out=out)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(n1.cpu().numpy(), host[0])
    np.testing.assert_array_equal(n0.cpu().numpy(), host[1])
    assert host[0][0] > 0 and host[1][0] > 0


def test_scatter_tile_assembles_the_field(lo):
    """ogn_scatter_tile with a local destination (what rank 0 does with its own tile; the peer mapping
    itself needs two processes: tools/check_sharded.py): four tiles copied into one cube reproduce it,
    halos never leak, vector and scalar paths agree."""
    import torch
    from origin_b200 import tiles
    from origin_b200._lib import ptr
    ctx = lo.default_context()
    for (ny, nx) in ((64, 96), (37, 51)):          # float4-aligned and ragged
        nz = 9
        full = torch.arange(nz * ny * nx, dtype=torch.float32, device='cuda').reshape(nz, ny, nx)
        dst = torch.full_like(full, -1.0)
        for t in tiles.plan_tiles(ny, nx, 4, 13):
            sub = (full[(slice(None),) + t.padded] + 0.0).contiguous()
            desc = np.array([ny, nx, t.py0, t.px0, t.y0 - t.py0, t.y1 - t.py0, t.x0 - t.px0, t.x1 - t.px0], dtype=np.int32)
            ctx.check(ctx.lib.ogn_scatter_tile(ctx.handle, ptr(sub), nz, sub.shape[1], sub.shape[2], ptr(desc),
                                               dst.data_ptr()))
            ctx.check(ctx.lib.ogn_peer_sync(ctx.handle))   # `sub` dies at the end of the iteration
        ctx.check(ctx.lib.ogn_peer_join(ctx.handle))
        torch.cuda.synchronize()
        assert torch.equal(dst, full)


def test_step05_async_matches_sync(lo):
    """sync=False: no host synchronisation inside the call, list lengths stay on the device
    (DeviceExtrema), the device-only purity counts read them there; everything must equal the
    synchronous call."""
    import torch
    shape = (120, 48, 64)
    fsf = synthetic.moffat_fsf(shape[0])
    cube_h, _ = synthetic.faint_cube(shape, fsf, n_src=5, seed=51)
    mask_h = synthetic.footprint_mask(shape, seed=51)
    profs = dictionaries.dico_3fwhm()[0]
    cube, mask = torch.from_numpy(cube_h).cuda(), torch.from_numpy(mask_h.view(np.uint8)).cuda()
    ref = lo.step05(cube, fsf, None, profs, mask, 3, 1e-8, True)
    got = lo.step05(cube, fsf, None, profs, mask, 3, 1e-8, True, sync=False)
    ext = got['extrema']
    assert isinstance(ext, lo.DeviceExtrema) and ext._counts is None
    thr = torch.linspace(1.0, 6.0, 50, dtype=torch.float64, device='cuda')
    n1, n0 = lo.purity_counts(ext, None, thr)                 # device-only entry point, still no sync
    assert ext._counts is None
    r1, r0 = lo.purity_counts(ref['extrema'], None, thr)
    torch.cuda.synchronize()
    assert torch.equal(n1, r1) and torch.equal(n0, r0) and int(r1[0]) > 0
    assert ext.counts == ref['extrema'].counts
    for k in ('max_index', 'max_value', 'min_index', 'min_value'):
        assert torch.equal(getattr(ext, k), getattr(ref['extrema'], k)), k
    for k in ('correl', 'correl_min', 'profile', 'maxmap', 'minmap'):
        assert torch.equal(got[k], ref[k]), k
    # a capacity that is too small is detected when the lengths are finally read
    small = lo.step05(cube, fsf, None, profs, mask, 3, 1e-8, True, sync=False, capacity=16)
    with pytest.raises(OverflowError):
        small['extrema'].counts


@pytest.mark.parametrize('case', ['seven_sorted', 'twelve_sorted_short_cube', 'unsorted_falls_back', 'asymmetric_falls_back'])
def test_tglr_folded_spectral_kernel_cases(lo, case):
    """K2f (folded spectral kernel) takes symmetric, width-sorted dictionaries with more than 3 profiles:
    partially filled groups, two groups, a cube shorter than the widest profile's reach; dictionaries it
    must refuse (widths not sorted: the argmax is first-wins in dictionary order; an asymmetric profile)
    go through K2.  All against the float64 oracle."""
    full = dictionaries.dico_fwhm_2_12()[0]
    rng = np.random.default_rng(61)
    shape = (140, 20, 45)
    if case == 'seven_sorted':
        profs = [full[i] for i in (0, 3, 5, 8, 12, 15, 19)]
    elif case == 'twelve_sorted_short_cube':
        profs, shape = list(full[4:16]), (40, 18, 37)
    elif case == 'unsorted_falls_back':
        profs = [full[i] for i in (10, 2, 19, 7, 14)]
    else:
        profs = [np.array(p, dtype=np.float64) for p in full[:6]]
        profs[3] = profs[3] * (1.0 + 0.2 * np.linspace(-1, 1, len(profs[3])))     # skewed line
    fsf = synthetic.moffat_fsf(shape[0])
    cube, _ = synthetic.faint_cube(shape, fsf, n_src=4, seed=62)
    mask = rng.random(shape) < 0.02
    ref = orc.tglr_step(cube, fsf, None, profs, mask, 3, 1, 1e-8, True)
    out = lo.tglr(cube, fsf, None, profs, mask=mask, pcut=1e-8)
    assert_close(out['correl'], ref['cube_correl'], case + ' correl')
    assert_close(out['correl_min'], ref['cube_correl_min'], case + ' correl_min')
    tk = oracle_tk(cube, fsf, None, profs, 1e-8, True)
    check_profile(np.where(mask, 0, out['profile']), np.where(mask, 0, ref['cube_profile']), tk, case + ' profile',
                  max_frac=3e-3)


def test_threshold_rows_rejects_a_profile_of_another_shape(lo):
    ext = lo.LocalExtrema((4, 6, 8), np.array([5, 100], dtype=np.int64), np.array([9.0, 7.0], dtype=np.float32),
                          np.zeros(0, np.int64), np.zeros(0, np.float32))
    with pytest.raises(ValueError):
        lo.threshold_rows(ext, 1.0, np.zeros((4, 3, 8), np.uint8))


def test_step05_host_tile_on_the_streamed_path(lo):
    """A HOST sub-cube in tile mode takes the slab-pipelined path (upload / kernels / download overlapped): same
    products on the owned window and same whole-field lists as the device-resident tile call."""
    import torch
    from origin_b200 import tiles
    shape = (96, 160, 192)
    nz, ny, nx = shape
    fsf = synthetic.moffat_fsf(nz)
    cube, _ = synthetic.faint_cube(shape, fsf, n_src=8, seed=13)
    mask = synthetic.footprint_mask(shape, seed=13)
    profs = dictionaries.dico_3fwhm()[0]
    for t in tiles.plan_tiles(ny, nx, 2, 13):
        sl = (slice(None),) + t.padded
        sub, msub = np.ascontiguousarray(cube[sl]), np.ascontiguousarray(mask[sl])
        host = lo.step05(sub, fsf, None, profs, msub, 3, 1e-8, True, tile=(t, (ny, nx)))
        dev = lo.step05(torch.from_numpy(sub).cuda(), fsf, None, profs, torch.from_numpy(msub.view(np.uint8)).cuda(), 3,
                        1e-8, True, tile=(t, (ny, nx)))
        ys, xs = t.owned
        for key in ('correl', 'correl_min', 'profile'):
            np.testing.assert_array_equal(host[key][:, ys, xs], dev[key].cpu().numpy()[:, ys, xs], err_msg=key)
        np.testing.assert_array_equal(host['extrema'].max_index, dev['extrema'].max_index.cpu().numpy())
        np.testing.assert_array_equal(host['extrema'].min_value, dev['extrema'].min_value.cpu().numpy())
        np.testing.assert_array_equal(host['maxmap'][ys, xs], dev['maxmap'].cpu().numpy()[ys, xs])


def test_two_contexts_on_one_device_do_not_mix_their_dictionaries(lo):
    """The taps of a TGLR call live in __constant__ memory shared by the contexts of a device: the library orders
    the calls of different contexts (ogn_tglr_guard), so interleaved asynchronous calls with different
    dictionaries on two streams still give each context its own result."""
    import torch
    from origin_b200 import _lib
    shape = (200, 64, 64)
    fsf = torch.from_numpy(synthetic.moffat_fsf(shape[0])).cuda()
    cube = torch.from_numpy(synthetic.faint_cube(shape, None, n_src=4, seed=17)[0]).cuda()
    mask = torch.zeros(shape, dtype=torch.uint8, device='cuda')
    p3, p20 = dictionaries.dico_3fwhm()[0], dictionaries.dico_fwhm_2_12()[0]
    ref3 = lo.step05(cube, fsf, None, p3, mask, 3, 1e-8, True)
    ref20 = lo.step05(cube, fsf, None, p20, mask, 3, 1e-8, True)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    c1, c2 = _lib.Context(0, s1.cuda_stream), _lib.Context(0, s2.cuda_stream)
    outs = []
    for _ in range(6):        # nothing synchronises in between: the streams would overlap without the guard
        a = lo.step05(cube, fsf, None, p3, mask, 3, 1e-8, True, ctx=c1, sync=False)
        b = lo.step05(cube, fsf, None, p20, mask, 3, 1e-8, True, ctx=c2, sync=False)
        outs.append((a, b))
    torch.cuda.synchronize()
    for a, b in outs:
        assert torch.equal(a['correl'], ref3['correl']) and torch.equal(a['profile'], ref3['profile'])
        assert torch.equal(b['correl'], ref20['correl']) and torch.equal(b['profile'], ref20['profile'])
    c1.close()
    c2.close()


def test_default_kernel_variants(lo):
    """The code paths the library picks when no OGN_* diagnostic switch is set (``ogn_variants``): TMA spatial
    kernel, constant-tap ring kernel for Dico_3FWHM / folded FFMA2 kernel for Dico_FWHM_2_12, TMA 3x3x3
    extremum kernel, slab-pipelined step05 for host cubes, TMA-streamed step01 for float32 cubes."""
    import os
    import torch
    from origin_b200._lib import default_context
    switches = [k for k in os.environ if k.startswith('OGN_') and k not in ('OGN_BENCH_STAGGER_US',)]
    if switches:
        pytest.skip('diagnostic switches set: %s' % switches)
    ctx = default_context()
    shape = (160, 64, 96)
    fsf = synthetic.moffat_fsf(shape[0])
    cube, _ = synthetic.faint_cube(shape, fsf, n_src=3, seed=5)
    mask = synthetic.footprint_mask(shape, seed=5)
    lo.step05(cube, fsf, None, dictionaries.dico_3fwhm()[0], mask, 3, 1e-8, True)
    v = ctx.variants()
    assert v['k1'] == 'tma25' and v['k2'] == 'ring:0' and v['k3'] == 'tma3x3x3' and v['step05'] == 'slab-pipelined', v
    assert ctx.fsf_folded                                   # the Moffat FSF is mirror-symmetric: row-folded K1
    dcube = torch.from_numpy(cube).cuda()
    dmask = torch.from_numpy(mask.view(np.uint8)).cuda()
    lo.step05(dcube, fsf, None, dictionaries.dico_fwhm_2_12()[0], dmask, 3, 1e-8, True)
    v = ctx.variants()
    assert v['k2'] == 'folded:g10:ffma2' and v['step05'] == 'resident', v
    raw = (cube * 2 + 1).astype(np.float32)
    var = np.full(shape, 4.0, np.float32)
    lo.preprocess(raw, var, mask, dct_order=10)
    assert ctx.variants()['step01'] == 'tma-stream', ctx.variants()
    lo.preprocess(raw.astype(np.float64), var.astype(np.float64), mask, dct_order=10)
    assert ctx.variants()['step01'] == 'column', ctx.variants()


def test_mosaic_fields_restricted_to_their_footprints(lo):
    """Two fields whose weight maps cover different parts of the image: each field's spatial passes run only on
    the rectangle its weights can reach (``ogn_variants``: k1 = tma25:footprint) and the sum must still be what
    ``_convolve_fsf`` gives over the whole image (lib_origin.py:1027-1043)."""
    from origin_b200._lib import default_context
    shape = (120, 96, 224)
    nz, ny, nx = shape
    fsf0 = synthetic.moffat_fsf(nz, fwhm0=3.6, fwhm1=2.9)
    fsf1 = synthetic.moffat_fsf(nz, fwhm0=4.2, fwhm1=3.1)
    cube, _ = synthetic.faint_cube(shape, fsf0, n_src=8, seed=41)
    xx = np.arange(nx)[None, :] * np.ones((ny, 1))
    w0 = np.clip((100 - xx) / 16.0, 0.0, 1.0)            # field 0: x < 100, full weight below 84
    w1 = 1.0 - w0                                        # field 1: x > 84 ...
    w1[:30] = 0.0                                        # ... and only rows >= 30
    w0[:30, 100:] = 0.0
    cf, nf = lo.fsf_stage(cube, [fsf0, fsf1], [w0, w1])
    assert default_context().variants()['k1'] == 'tma25:footprint'
    rcf, rnf = orc.fsf_correlate(cube, [fsf0, fsf1], [w0, w1])
    assert_close(cf, rcf, 'mosaic cube_fsf')
    assert_close(nf, rnf, 'mosaic norm_fsf')
    # nothing reaches rows < 30 - 12 right of x = 100 + 12: exact zeros there, not stale scratch
    assert not cf[:, :18, 112:].any() and not nf[:, :18, 112:].any()
    profs = dictionaries.dico_3fwhm()[0]
    ref = orc.correlation_glr_test(cube, [fsf0, fsf1], [w0, w1], profs, pcut=1e-8)
    correl, profile, correl_min = lo.Correlation_GLR_test(cube, [fsf0, fsf1], [w0, w1], profs, pcut=1e-8)
    sel = np.ones((ny, nx), dtype=bool)
    sel[:30 + 12, 100 - 12:] = False                     # the footprint sees uncovered voxels there (note N1)
    assert_close(correl[:, sel], ref[0][:, sel], 'mosaic correl')
    assert_close(correl_min[:, sel], ref[2][:, sel], 'mosaic correl_min')
    assert np.mean(profile[:, sel] == ref[1][:, sel]) > 0.999
