"""Wall-clock probes of the widened path at the benchmark size (development aid, run on a B200):

* fused step04 -> step05 through ``patch_steps`` on a 3681x320x320 field with four areas: ``cube_faint`` handed over
  on the device (LazyProduct) against the host route (numpy in / numpy out of ``Compute_GreedyPCA_area``, then the
  fused step05 on the host cube), pageable numpy arrays like the reference's step objects hold;
* step08 at catalogue scale: 1000 detections (9000 windows of 3681x25x25) on the 3681x96x96 cube of
  ``tools/pca_lines_probe.py``.

usage: python tools/chain_probe.py [chain] [lines]   -> one JSON line"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import torch  # noqa: E402

from origin_b200 import dictionaries, lib_origin as lo, steps, synthetic  # noqa: E402

what = sys.argv[1:] or ['chain', 'lines']
out = {}


def wall(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = fn()
    torch.cuda.synchronize()
    return res, time.perf_counter() - t0


if 'chain' in what:
    import test_gpu_steps as tg
    nz, ny, nx = 3681, 320, 320
    fsf = synthetic.moffat_fsf(nz)
    g = torch.Generator(device='cuda').manual_seed(5)
    cube = torch.randn((nz, ny, nx), generator=g, device='cuda', dtype=torch.float32)
    rng = np.random.default_rng(11)
    lam = torch.linspace(0, 1, nz, device='cuda')
    prof = torch.from_numpy((fsf / fsf.max(axis=(1, 2), keepdims=True)).astype(np.float32)).cuda()
    for _ in range(40):                                   # continuum sources left in the standardised cube
        y0, x0 = int(rng.integers(12, ny - 12)), int(rng.integers(12, nx - 12))
        spec = float(rng.uniform(2, 8)) * (0.6 + 0.4 * torch.cos(float(rng.uniform(1, 6)) * lam + float(rng.uniform(0, 3))))
        cube[:, y0 - 12:y0 + 13, x0 - 12:x0 + 13] += spec[:, None, None] * prof
    areamap = np.ones((ny, nx), dtype=int)
    areamap[:160, 160:] = 2
    areamap[160:, :160] = 3
    areamap[160:, 160:] = 4
    thr, tests = [], []
    for a in range(1, 5):
        sel = torch.from_numpy(areamap == a).cuda()
        t, _, _, th, _, _ = lo.Compute_PCA_threshold(cube[:, sel], 0.01)
        thr.append(th)
        tests.append(t)
    cube_std = cube.cpu().numpy()                         # pageable host array, as Preprocessing stores it
    del cube, prof
    torch.cuda.empty_cache()
    mask = synthetic.footprint_mask((nz, ny, nx), seed=5)
    profs = dictionaries.dico_3fwhm()[0]
    mod = tg.fake_steps_module()
    steps.patch_steps(mod, fused=True)
    res = dict(shape=[nz, ny, nx], areas=4, thresholds=thr)
    for rep in range(2):                                  # second pass: scratch and caches warm
        orig = tg.FakeOrigin(mask=mask, PSF=fsf, wfields=None, profiles=profs, cube_std=tg.FakeData(cube_std),
                             areamap=tg.FakeData(areamap), nbAreas=4, thresO2=thr, testO2=tests)
        pca, tglr = mod.ComputeGreedyPCA(orig), mod.ComputeTGLR(orig)
        orig.steps = {'compute_greedy_PCA': pca, 'compute_TGLR': tglr}
        _, t04 = wall(lambda: pca.run(orig))
        _, t05 = wall(lambda: tglr.run(orig, pcut=1e-8))
        res['device_handoff'] = dict(step04_s=t04, step05_s=t05, total_s=t04 + t05, iterations=float(pca.mapO2._data.max()),
                                     faint_still_on_device=isinstance(pca.__dict__['cube_faint'], steps.LazyProduct))
        correl_dev = orig.cube_correl._data
        # host route: numpy cube_faint comes back from step04 and goes up again in step05
        (faint, map_o2, nstop), h04 = wall(lambda: lo.Compute_GreedyPCA_area(4, cube_std, areamap, 50, thr, 100, tests))
        orig2 = tg.FakeOrigin(mask=mask, PSF=fsf, wfields=None, profiles=profs, cube_faint=tg.FakeData(faint))
        tglr2 = mod.ComputeTGLR(orig2)
        orig2.steps = {'compute_TGLR': tglr2}
        _, h05 = wall(lambda: tglr2.run(orig2, pcut=1e-8))
        res['host_route'] = dict(step04_s=h04, step05_s=h05, total_s=h04 + h05)
        res['correl_max_abs_diff'] = float(np.abs(orig2.cube_correl._data - correl_dev).max())
        del orig, orig2, pca, tglr, tglr2, faint, correl_dev
    steps.unpatch_steps()
    out['chain'] = res
    del cube_std, mask
    torch.cuda.empty_cache()

if 'lines' in what:
    nz, ny, nx = 3681, 96, 96
    rng = np.random.default_rng(11)
    fsf = synthetic.moffat_fsf(nz)
    cube = rng.standard_normal((nz, ny, nx)).astype(np.float32)
    lam = np.linspace(0, 1, nz)
    ndet = 1000
    dets = dict(z0=rng.integers(50, nz - 50, ndet), y0=rng.integers(0, ny, ndet), x0=rng.integers(0, nx, ndet))
    for z0, y0, x0 in zip(dets['z0'][:200], dets['y0'][:200], dets['x0'][:200]):      # a line under a fifth of them
        ya, yb, xa, xb = max(0, y0 - 12), min(ny, y0 + 13), max(0, x0 - 12), min(nx, x0 + 13)
        zz = np.arange(z0 - 10, z0 + 11)
        cube[z0 - 10:z0 + 11, ya:yb, xa:xb] += (25 * np.exp(-0.5 * ((zz - z0) / 2.0) ** 2)[:, None, None]
                                                * fsf[z0, ya - y0 + 12:yb - y0 + 12, xa - x0 + 12:xb - x0 + 12] * 20).astype(np.float32)
    var = (1.0 + 0.3 * np.sin(6 * lam) ** 2)[:, None, None] * np.ones((1, ny, nx))
    r = torch.from_numpy((cube * np.sqrt(var)).astype(np.float32)).cuda()
    v = torch.from_numpy(var.astype(np.float32)).cuda()
    (cat2, le, ve), dt = wall(lambda: lo.estimation_line(dets, r, v, fsf, size_grid=1))
    (cat0, le0, _), dt0 = wall(lambda: lo.estimation_line(dets, r, v, fsf, size_grid=0))
    out['lines'] = dict(detections=ndet, seconds_grid1=dt, ms_per_detection_grid1=dt * 1e3 / ndet,
                        seconds_grid0=dt0, ms_per_detection_grid0=dt0 * 1e3 / ndet,
                        moved=int(np.sum((cat2['y'] != dets['y0']) | (cat2['x'] != dets['x0']))),
                        finite=int(sum(np.isfinite(l).all() for l in le)))
print(json.dumps(out))
