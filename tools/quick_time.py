"""Device-resident timing of step05 on a BASELINE-shaped cube (development aid)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from origin_b200 import dictionaries, lib_origin, synthetic  # noqa: E402

shape = tuple(int(v) for v in (sys.argv[1:4] if len(sys.argv) >= 4 else (3681, 320, 320)))
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
fsf = synthetic.moffat_fsf(shape[0])
g = torch.Generator(device='cuda').manual_seed(0)
cube = torch.randn(shape, device='cuda', dtype=torch.float32, generator=g)
mask = (torch.rand(shape, device='cuda', generator=g) < 0.01).to(torch.uint8)
ctx = lib_origin.default_context()
ctx.timing(True)
for name, profs in (('3FWHM', dictionaries.dico_3fwhm()[0]), ('2_12', dictionaries.dico_fwhm_2_12()[0])):
    for r in range(reps):
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        out = lib_origin.tglr(cube, fsf, None, profs, mask=mask, pcut=1e-8)
        e1.record()
        ext, _, _ = lib_origin.local_extrema(out['correl'], out['correl_min'], mask, 3)
        e2.record()
        torch.cuda.synchronize()
        v = shape[0] * shape[1] * shape[2] * len(profs)
        print('%s rep %d: tglr %.2f ms  extrema %.2f ms  -> %.1f Gvoxel.profiles/s (tglr only)  lists %s' % (
            name, r, e0.elapsed_time(e1), e1.elapsed_time(e2), v / e0.elapsed_time(e1) / 1e6, ext.counts), flush=True)
        print('   stages:', ' '.join('%s=%.3f' % kv for kv in ctx.timing_report()), flush=True)
