"""Host-side wall-clock breakdown of one device-resident step (development aid)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from origin_b200 import dictionaries, lib_origin, synthetic

shape = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (3681, 320, 320)
fsf = torch.from_numpy(synthetic.moffat_fsf(shape[0])).cuda()
profs = dictionaries.dico_3fwhm()[0]
g = torch.Generator(device='cuda').manual_seed(0)
cube = torch.randn(shape, device='cuda', generator=g)
mask = (torch.rand(shape, device='cuda', generator=g) < 0.01).to(torch.uint8)
ctx = lib_origin.default_context()
cap = cube.numel() // 40
out = dict(correl=torch.empty_like(cube), correl_min=torch.empty_like(cube), profile=torch.empty_like(mask),
           maxmap=torch.empty(shape[1:], device='cuda'), minmap=torch.empty(shape[1:], device='cuda'),
           max_index=torch.empty(cap, dtype=torch.int64, device='cuda'), max_value=torch.empty(cap, device='cuda'),
           min_index=torch.empty(cap, dtype=torch.int64, device='cuda'), min_value=torch.empty(cap, device='cuda'))
thr = np.linspace(4.0, 12.0, 50)
for rep in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = lib_origin.step05(cube, fsf, None, profs, mask, 3, 1e-8, True, out=out, ctx=ctx)
    t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    n1, n0 = lib_origin.purity_counts(res['extrema'], None, thr, ctx)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print('step05 call %.3f ms (+%.3f to drain)  purity_counts %.3f ms' % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3), flush=True)
ctx.timing(True)
res = lib_origin.step05(cube, fsf, None, profs, mask, 3, 1e-8, True, out=out, ctx=ctx)
print(' '.join('%s=%.3f' % kv for kv in ctx.timing_report()))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    res = lib_origin.step05(cube, fsf, None, profs, mask, 3, 1e-8, True, out=out, ctx=ctx)
    lib_origin.purity_counts(res['extrema'], None, thr, ctx)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
