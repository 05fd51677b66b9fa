"""Host-side cost of one device-resident asynchronous step05 call on a small tile (development aid).
usage: OGN_HOST_TRACE=1 python tools/host_trace_probe.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from origin_b200 import dictionaries, lib_origin, synthetic, tiles
import cProfile, pstats
shape = (3681, 106, 192)
profs = dictionaries.dico_3fwhm()[0]
fsf = torch.from_numpy(synthetic.moffat_fsf(shape[0])).cuda()
cube = torch.randn(shape, device='cuda')
mask = (torch.rand(shape, device='cuda') < 0.01).to(torch.uint8)
ctx = lib_origin.default_context()
t = tiles.plan_tiles(320, 320, 8, 13)[0]
cube = cube[:, :t.shape[0], :t.shape[1]].contiguous(); mask = mask[:, :t.shape[0], :t.shape[1]].contiguous()
out = None
def step():
    return lib_origin.step05(cube, fsf, None, profs, mask, 3, 1e-8, True, ctx=ctx, tile=(t, (320, 320)), sync=False)
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
print('host ms per call: %.3f' % ((t1 - t0) * 1e3 / 50))
pr = cProfile.Profile(); pr.enable()
for _ in range(50): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
