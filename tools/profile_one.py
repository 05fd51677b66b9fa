"""One step05 + extrema on a BASELINE-shaped cube, for ncu captures (development aid).
usage: profile_one.py [3FWHM|2_12] [nz ny nx]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from origin_b200 import dictionaries, lib_origin, synthetic
name = sys.argv[1] if len(sys.argv) > 1 else '3FWHM'
shape = tuple(int(v) for v in sys.argv[2:5]) if len(sys.argv) >= 5 else (3681, 320, 320)
profs = dictionaries.dico_3fwhm()[0] if name == '3FWHM' else dictionaries.dico_fwhm_2_12()[0]
fsf = torch.from_numpy(synthetic.moffat_fsf(shape[0])).cuda()
g = torch.Generator(device='cuda').manual_seed(0)
cube = torch.randn(shape, device='cuda', generator=g)
mask = (torch.rand(shape, device='cuda', generator=g) < 0.01).to(torch.uint8)
for _ in range(2):
    res = lib_origin.step05(cube, fsf, None, profs, mask, 3, 1e-8, True)
torch.cuda.synchronize()
print('ok', res['extrema'].counts)
