"""Device-resident timing of the weighted / multi-field TGLR path (development aid)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from origin_b200 import dictionaries, lib_origin, synthetic

shape = tuple(int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (3681, 320, 320)
nf = int(sys.argv[4]) if len(sys.argv) > 4 else 2
nz, ny, nx = shape
fsf0 = synthetic.moffat_fsf(nz)
fsfs = [torch.from_numpy(fsf0 * (1.0 + 0.0 * f)).cuda() for f in range(nf)]
xx = np.linspace(0, 1, nx)[None, :] * np.ones((ny, 1))
ws = [torch.from_numpy(np.clip(1.0 - np.abs(xx - (f + 0.5) / nf) * nf, 0, 1)).cuda() for f in range(nf)]
g = torch.Generator(device='cuda').manual_seed(0)
cube = torch.randn(shape, device='cuda', generator=g)
ctx = lib_origin.default_context()
ctx.timing(True)
print('variants after the runs are printed last; OGN_K1_NO_FOOTPRINT=%s' % os.environ.get('OGN_K1_NO_FOOTPRINT'))
for name, profs in (('3FWHM', dictionaries.dico_3fwhm()[0]), ('2_12', dictionaries.dico_fwhm_2_12()[0])):
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        out = lib_origin.tglr(cube, fsfs, ws, profs, mask=None, pcut=1e-8)
        e1.record(); torch.cuda.synchronize()
        print('%s, %d fields: tglr %.2f ms  %s' % (name, nf, e0.elapsed_time(e1),
              ' '.join('%s=%.3f' % kv for kv in ctx.timing_report())), flush=True)
print(ctx.variants())
