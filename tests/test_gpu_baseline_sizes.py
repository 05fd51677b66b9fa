"""Parity at (or towards) the BASELINE sizes: the cases round 1's tests stopped short of.

* step01 on the full 3681-plane wavelength axis (16 wavelength segments per spaxel) with a fully masked
  plane (``np.nanmean`` of an empty plane: NaN mean, ``lib_origin.py:196-197`` / ``steps.py:442``) and
  spaxels with a single masked voxel (they take the UNWEIGHTED branch, ``lib_origin.py:226-237``);
* the weighted two-field path on a 400 x 96 x 128 mosaic with an uncovered strip;
* the 3681 x 320 x 320 benchmark cube itself for both dictionaries, checked pointwise with the float64
  direct-space oracle on seeded voxels and their 3x3x3 neighbourhoods (64 x 64 K1 tiles, wavelength
  waves, the 25 x 25 edge-class table and the K3 grid all differ from the small cubes);
* the multi-GPU stitched-parity program, run under torchrun when the box has >= 2 GPUs.
"""

import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, tol_report
from oracle import origin_oracle as orc
from origin_b200 import dictionaries, synthetic

pytestmark = pytest.mark.gpu

RTOL = 1e-5


@pytest.fixture(scope='module')
def lo():
    from origin_b200 import lib_origin
    return lib_origin


def _close(got, ref, what, rtol=RTOL):
    rep = tol_report(got, ref, rtol)
    assert rep['ok'], '%s: max_abs=%.3g worst=%.3g x bound (rms %.3g)' % (what, rep['max_abs'], rep['worst'], rep['rms'])


def test_step01_full_wavelength_axis_with_masked_plane(lo):
    shape = (3681, 32, 48)
    fsf = synthetic.moffat_fsf(shape[0])
    raw, var, mask = synthetic.raw_cube(shape, fsf, n_cont=3, n_src=6, seed=21)
    mask = mask.copy()
    mask[1234] = True                       # a fully masked wavelength plane: nanmean over no voxel
    mask[77, 5, 9] = True                   # spaxels with exactly one masked voxel: unweighted branch
    mask[3000, 20, 40] = True
    raw[mask] = 0.0
    var[mask] = np.inf
    with np.errstate(all='ignore'):
        ref = orc.preprocessing(raw, var, mask, 10, False, 3)
    out = lo.preprocess(raw, var, mask, 10, False)
    assert np.isnan(out['mean_lambda'][1234]) and np.isfinite(np.delete(out['mean_lambda'], 1234)).all()
    # the masked plane is zeroed by data[mask] = 0 (steps.py:446) whatever the NaN mean did to it
    assert np.all(out['cube_std'][1234] == 0)
    _close(out['cube_std'], ref['cube_std'], 'cube_std', rtol=2e-5)
    _close(out['cont_dct'], ref['cont_dct'], 'cont_dct', rtol=2e-5)
    for key in ('ima_std', 'ima_dct', 'cont_sumsq', 'o2map'):
        _close(out[key], ref[key], key, rtol=2e-5)
    # the masked plane puts EVERY spaxel on the unweighted branch (valid = ~any(mask, axis=0), lib_origin.py:226);
    # without it both branches of dct_residual occur in one cube
    mask2 = mask.copy()
    mask2[1234] = False
    mask2[1234, 3:9, 4:11] = True
    raw2, var2 = raw.copy(), var.copy()
    var2[1234] = var[1233]
    raw2[mask2] = 0.0
    var2[mask2] = np.inf
    valid = ~mask2.any(axis=0)
    assert valid.any() and (~valid).any()
    cont = lo.dct_residual(raw2, 10, var2, False, mask2)
    rcont = orc.dct_residual(raw2, 10, var2, False, mask2)
    assert np.abs(cont - rcont).max() <= 1e-9 * np.abs(rcont).max()


def test_step01_is_reproducible_run_to_run(lo):
    """The per-wavelength sums are reduced in a fixed order: two runs give bit-identical cube_std."""
    shape = (700, 40, 48)
    fsf = synthetic.moffat_fsf(shape[0])
    raw, var, mask = synthetic.raw_cube(shape, fsf, n_cont=3, n_src=4, seed=22)
    raw, var = raw.astype(np.float32), var.astype(np.float32)      # MUSE cubes are float32: the streamed kernels
    a = lo.preprocess(raw, var, mask, 10, False)
    b = lo.preprocess(raw, var, mask, 10, False)
    assert np.array_equal(a['mean_lambda'], b['mean_lambda'], equal_nan=True)
    assert np.array_equal(a['cube_std'], b['cube_std'])


def test_two_field_mosaic_with_uncovered_strip(lo):
    shape = (400, 96, 128)
    nz, ny, nx = shape
    fsf0 = synthetic.moffat_fsf(nz, fwhm0=3.6, fwhm1=2.9)
    fsf1 = synthetic.moffat_fsf(nz, fwhm0=4.2, fwhm1=3.1)
    cube, _ = synthetic.faint_cube(shape, fsf0, n_src=10, seed=31)
    w0, w1 = synthetic.field_weights(ny, nx, 2)
    w0, w1 = w0.copy(), w1.copy()
    w0[:, :20] = 0.0                        # a strip no field covers
    w1[:, :20] = 0.0
    profs = dictionaries.dico_fwhm_2_12()[0]
    ref = orc.correlation_glr_test(cube, [fsf0, fsf1], [w0, w1], profs, pcut=1e-8)
    correl, profile, correl_min = lo.Correlation_GLR_test(cube, [fsf0, fsf1], [w0, w1], profs, pcut=1e-8)
    sel = np.zeros((ny, nx), dtype=bool)
    sel[:, 20 + 12:] = True                 # where the 25 x 25 footprint sees covered data only (note N1)
    _close(correl[:, sel], ref[0][:, sel], 'two-field correl')
    _close(correl_min[:, sel], ref[2][:, sel], 'two-field correl_min')
    assert np.mean(profile[:, sel] == ref[1][:, sel]) > 0.999
    # uncovered strip: the reference holds FFT round-off there, we hold exact zeros
    assert np.abs(correl[:, :, :8]).max() <= 1e-5


@pytest.mark.parametrize('dico', ['3FWHM', '2_12'])
def test_benchmark_cube_spot_parity(dico):
    """The 3681 x 320 x 320 cube every headline number is measured on."""
    import torch
    import bench
    from origin_b200 import _lib
    env = bench.Env()
    sys_argv, sys.argv = sys.argv, ['bench.py']
    try:
        env.args = bench.parse_args()
    finally:
        sys.argv = sys_argv
    env.world, env.rank, env.local_rank = 1, 0, 0
    env.dev = torch.device('cuda', 0)
    env.ctx = _lib.default_context(0)
    job = bench.Job(env, bench.SHAPE, dico)
    res = job.lo.step05(job.cube, job.fsf, None, job.profs, job.mask, 3, 1e-8, True, ctx=env.ctx)
    spot = job.parity_spot(res['correl'], res['correl_min'], res['profile'], job.host_lists(res['extrema']), job.cube)
    assert spot['n'] >= 64
    assert spot['worst_over_bound'] <= 1.0, spot
    assert spot['argmax_mismatches'] == 0 and spot['extremum_mismatches'] == 0, spot
    assert spot['argmax_ties'] + spot['extremum_ties'] <= 3, spot
    assert spot['maxima_checked'] >= 8 and spot['minima_checked'] >= 4
    assert spot['input_mismatch_voxels'] <= 2
    # size-independent properties at full size: masked voxels hold exact zeros, maxmap is the plane-wise maximum
    m = job.mask.bool()
    assert float(res['correl'][m].abs().max().item()) == 0.0 and int(res['profile'][m].max().item()) == 0
    assert torch.equal(res['maxmap'], res['correl'].amax(dim=0))
    assert torch.equal(res['minmap'], res['correl_min'].amin(dim=0))
    idx = res['extrema'].max_index
    assert bool((idx[1:] > idx[:-1]).all())                       # C order = np.where order (steps.py:958)
    assert torch.equal(res['correl'].reshape(-1)[idx], res['extrema'].max_value)
    job.close()
    del res, job
    torch.cuda.empty_cache()


def test_sharded_pipeline_matches_one_gpu():
    """tools/check_sharded.py under torchrun on 2 GPUs (when the box has them): stitched cube_std / correl,
    extremum lists, purity table and the peer-memory gather against the single-GPU run."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs on the box')
    n = 2
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(n),
                          '--master-addr', '127.0.0.1', '--master-port', '29533',
                          os.path.join(ROOT, 'tools', 'check_sharded.py')], capture_output=True, text=True, timeout=900)
    log = out.stdout + out.stderr
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'check_sharded_n%d.log' % n), 'w') as f:
        f.write(log)
    assert out.returncode == 0 and 'CHECK_SHARDED PASS' in log, log[-2000:]
