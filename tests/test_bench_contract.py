"""The JSON contract of ``bench.py --impl reference`` (the CPU arm; runs without a GPU) and the static
pieces of the GPU arm's contract."""

import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--shape', '96', '64', '96',
                          '--steps', '1', '--warmup', '0'], capture_output=True, text=True, timeout=600, check=True).stdout
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'Gvoxel.profiles/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['ms_per_step'] > 0 and d['n_gpus'] == 1
    cb = d['cpu_baseline']
    from oracle import ref_loader
    assert cb['kind'] == ('reference' if ref_loader.available() else 'port')
    assert cb['cores'] >= 1 and cb['value'] == d['value'] and 'tiles of' in cb['sample']
    assert d['steps'] == 1 and d['warmup'] == 0           # the arm honours --steps / --warmup
    assert d['e2e'] == dict(value=d['value'], unit=d['unit'], h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert 'workload' in d['config'] and 'model' not in d['config']


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2'], env=env,
                         capture_output=True, text=True, timeout=120, check=True).stdout
    assert out.strip() == ''


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1', '--warmup', '0'], capture_output=True,
                       text=True, timeout=300)
    assert r.returncode != 0 and 'no CPU fallback' in (r.stderr + r.stdout)


def test_both_arms_print_the_same_config():
    import bench
    cfg = bench.workload_config((3681, 320, 320), '3FWHM', 3, 8)
    assert set(cfg) == {'workload', 'psf_size', 'parallelism', 'l2'} and '4x2' in cfg['parallelism']


def test_executed_slot_models():
    """The FP32 issue-slot counts the roofline divides by (mirrors ogn_k2f_prepare / K1's row folding)."""
    import bench
    pc = bench.cut_profiles(bench.dictionary('2_12'))
    fma, add = bench.k2f_layout(pc)
    assert (fma, add) == (10 * 17 + 10 * 33, 16 + 32)
    m = bench.kernel_models(pc, True, True)
    assert m['k1_fsf_correlate']['algorithmic'] == 1250 and abs(m['k1_fsf_correlate']['executed'] - 692) < 1
    assert m['k2_spectral_glr']['algorithmic'] == 2 * 704
