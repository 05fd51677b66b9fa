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
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and 'tile' in cb['sample']
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
