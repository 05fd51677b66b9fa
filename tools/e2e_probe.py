"""Wall-clock breakdown of the host-buffer step05 path (development aid)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from origin_b200 import _lib, dictionaries, lib_origin, synthetic

shape = (3681, 320, 320)
fsf = synthetic.moffat_fsf(shape[0])
profs = dictionaries.dico_3fwhm()[0]
cube = _lib.pinned_empty(shape, np.float32); cube[...] = np.random.default_rng(0).standard_normal(shape, dtype=np.float32)
mask = _lib.pinned_empty(shape, np.uint8); mask[...] = 0
cap = cube.size // 40
out = dict(correl=_lib.pinned_empty(shape, np.float32), correl_min=_lib.pinned_empty(shape, np.float32),
           profile=_lib.pinned_empty(shape, np.uint8), maxmap=_lib.pinned_empty(shape[1:], np.float32),
           minmap=_lib.pinned_empty(shape[1:], np.float32), max_index=_lib.pinned_empty((cap,), np.int64),
           max_value=_lib.pinned_empty((cap,), np.float32), min_index=_lib.pinned_empty((cap,), np.int64),
           min_value=_lib.pinned_empty((cap,), np.float32))
ctx = lib_origin.default_context()
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = lib_origin.step05(cube, fsf, None, profs, mask, 3, 1e-8, True, out=out, ctx=ctx)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print('step05 host buffers: %.1f ms' % ((t1 - t0) * 1e3), flush=True)
t0 = time.perf_counter()
args = lib_origin._tglr_args(cube, fsf, None, profs, mask, 1e-8, True)
print('python arg prep: %.1f ms' % ((time.perf_counter() - t0) * 1e3))
