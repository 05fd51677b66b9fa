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
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = lib_origin.preprocess(raw, var, mask, 10, False)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print('step01 %.2f ms  (%.1f Mspaxel/s)  %s' % (dt * 1e3, shape[1] * shape[2] / dt / 1e6,
          ' '.join('%s=%.3f' % kv for kv in ctx.timing_report())), flush=True)
