"""Timings of the step04 greedy PCA and of the step08 line estimation at MUSE wavelength depth (development aid).
usage: python tools/pca_lines_probe.py [gpu|cpu]   (cpu: the unmodified reference from oracle/_ref on the same inputs)"""
import json, os, sys, time, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from origin_b200 import synthetic

mode = sys.argv[1] if len(sys.argv) > 1 else 'gpu'
nz, ny, nx = 3681, 96, 96
rng = np.random.default_rng(11)
fsf = synthetic.moffat_fsf(nz)
# a standardised cube with a few bright continuum sources left in it: what step04 exists to remove
cube = rng.standard_normal((nz, ny, nx)).astype(np.float32)
lam = np.linspace(0, 1, nz)
for _ in range(12):
    y0, x0 = int(rng.integers(12, ny - 12)), int(rng.integers(12, nx - 12))
    spec = rng.uniform(2, 8) * (0.6 + 0.4 * np.cos(rng.uniform(1, 6) * lam + rng.uniform(0, 3)))
    cube[:, y0 - 12:y0 + 13, x0 - 12:x0 + 13] += (spec[:, None, None] * fsf / fsf.max(axis=(1, 2), keepdims=True)).astype(np.float32)
areamap = np.ones((ny, nx), dtype=int)
out = dict(shape=[nz, ny, nx], mode=mode)
dets = dict(z0=rng.integers(50, nz - 50, 32), y0=rng.integers(0, ny, 32), x0=rng.integers(0, nx, 32))
var = (1.0 + 0.3 * np.sin(6 * lam) ** 2)[:, None, None] * np.ones((1, ny, nx))
raw = (cube * np.sqrt(var)).astype(np.float32)
var = var.astype(np.float32)
if mode == 'gpu':
    import torch
    from origin_b200 import lib_origin
    c = torch.from_numpy(cube).cuda()
    test, _, _, thr, mea, std = lib_origin.Compute_PCA_threshold(c.reshape(nz, -1), 0.01)
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        faint, mapo2, nstop = lib_origin.Compute_GreedyPCA_area(1, c, areamap, 50, [thr], 100, [test])
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    out['pca'] = dict(seconds=dt, threshold=thr, iterations=float(mapo2.max()), nstop=nstop, nuisance_spaxels=int((mapo2 > 0).sum()))
    r, v = torch.from_numpy(raw).cuda(), torch.from_numpy(var).cuda()
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        cat2, le, ve = lib_origin.estimation_line(dets, r, v, fsf)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    out['lines'] = dict(seconds=dt, detections=32, windows=int(9 * 32), ms_per_detection=dt * 1e3 / 32)
else:
    from oracle import ref_loader
    from origin_b200 import segmap
    lib = ref_loader.load_lib_origin()
    c64 = cube.astype(np.float64)
    test = lib.O2test(c64.reshape(nz, -1))
    thr = segmap.compute_thresh_gaussfit(test, 0.01)[2]
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        t0 = time.perf_counter()
        faint, mapo2, nstop = lib.Compute_GreedyPCA_area(1, c64, areamap, 50, [thr], 100, [test])
        dt = time.perf_counter() - t0
    out['pca'] = dict(seconds=dt, threshold=thr, iterations=float(mapo2.max()), nstop=nstop, nuisance_spaxels=int((mapo2 > 0).sum()))
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
    import test_gpu_lines as tl
    sub = {k: v_[:4] for k, v_ in dets.items()}
    t0 = time.perf_counter()
    tl._reference_grid(raw.astype(np.float64), var.astype(np.float64), fsf, sub, 1, 'flux', 30, 1, 5)
    dt = time.perf_counter() - t0
    out['lines'] = dict(seconds=dt, detections=4, ms_per_detection=dt * 1e3 / 4)
print(json.dumps(out))
