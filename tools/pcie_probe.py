"""PCIe probe: pinned H2D / D2H bandwidth, alone and concurrently (development aid)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from origin_b200 import _lib

n = 3681 * 320 * 320
h_t = torch.empty(n, dtype=torch.float32, pin_memory=True)
h_o = _lib.pinned_empty((n,), np.float32)
h_o_t = torch.from_numpy(h_o)
print('ogn pinned is_pinned:', h_o_t.is_pinned())
d = torch.empty(n, dtype=torch.float32, device='cuda')
d2 = torch.empty(n, dtype=torch.float32, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def timed(fn, reps=3):
    torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best

gb = n * 4 / 1e9
print('H2D torch-pinned  %.1f GB/s' % (gb / timed(lambda: d.copy_(h_t, non_blocking=True))))
print('D2H torch-pinned  %.1f GB/s' % (gb / timed(lambda: h_t.copy_(d, non_blocking=True))))
print('H2D ogn-pinned    %.1f GB/s' % (gb / timed(lambda: d.copy_(h_o_t, non_blocking=True))))
print('D2H ogn-pinned    %.1f GB/s' % (gb / timed(lambda: h_o_t.copy_(d, non_blocking=True))))
def both():
    with torch.cuda.stream(s1): d.copy_(h_t, non_blocking=True)
    with torch.cuda.stream(s2): h_o_t.copy_(d2, non_blocking=True)
t = timed(both)
print('H2D+D2H concurrent: %.1f GB/s each direction' % (gb / t))
x = np.empty(n, dtype=np.float32)
xt = torch.from_numpy(x)
print('H2D pageable      %.1f GB/s' % (gb / timed(lambda: d.copy_(xt))))
print('D2H pageable      %.1f GB/s' % (gb / timed(lambda: xt.copy_(d))))
