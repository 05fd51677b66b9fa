"""Device-resident timing of step01 (DCT continuum + standardisation) on a BASELINE-shaped cube (development aid)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from origin_b200 import lib_origin

shape = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (3681, 320, 320)
g = torch.Generator(device='cuda').manual_seed(1)
raw = torch.randn(shape, device='cuda', generator=g) + 5.0
var = torch.rand(shape, device='cuda', generator=g) + 0.5
mask = (torch.rand(shape, device='cuda', generator=g) < 0.002).to(torch.uint8)
mask[:, :8, :] = 1
ctx = lib_origin.default_context()
ctx.timing(True)
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = lib_origin.preprocess(raw, var, mask, 10, False)
    t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    ext, _, _ = lib_origin.local_extrema(out['cube_std'], out['cube_std'], mask, 3, capacity=shape[0] * shape[1] * shape[2] // 16)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    print('preprocess: host %.2f ms, done %.2f ms; extrema %.2f ms  %s' % ((t1 - t0) * 1e3, (t2 - t0) * 1e3, (t3 - t2) * 1e3,
          ' '.join('%s=%.3f' % kv for kv in ctx.timing_report())), flush=True)
