"""The CPU-runnable half of bench.py's contract: `--impl reference` (the oracle port of the reference's CPU path on the
host cores) prints ONE JSON line with the agreed keys; the GPU arm refuses to run without a device instead of falling
back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + list(args), capture_output=True, text=True, timeout=600, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    res = _run('--impl', 'reference', '--steps', '1', '--warmup', '0')
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    b = json.loads(lines[0])
    assert b['impl'] == 'reference' and b['unit'] == 'voxel-samples/s' and b['higher_is_better'] is True
    for key in ('metric', 'value', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'e2e',
                'cpu_baseline'):
        assert key in b, key
    assert b['vs_baseline'] is None and b['steps'] == 1 and b['value'] > 0 and 'workload' in b['config']
    assert b['e2e']['value'] == b['value'] and b['e2e']['h2d_bytes_per_step'] == 0 and b['e2e']['d2h_bytes_per_step'] == 0
    cb = b['cpu_baseline']
    assert cb['kind'] in ('port', 'reference') and cb['cores'] >= 1 and cb['value'] == b['value'] and cb['sample']


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return
    res = _run('--steps', '1', '--warmup', '1', '--no-cpu-baseline')
    assert res.returncode != 0 and 'no CUDA device' in (res.stderr + res.stdout)
    assert not [l for l in res.stdout.splitlines() if l.startswith('{')]
