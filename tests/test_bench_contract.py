"""The bench.py contract the driver relies on, checked on the CPU with the reference arm (no GPU needed): exactly ONE
JSON line on stdout carrying the required keys, whatever libraries print."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REQUIRED = ['metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
            'vs_baseline', 'dtype', 'data', 'config', 'e2e', 'gpu_launches', 'cpu_baseline', 'impl']


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS='4')
    cmd = [sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
           '--cpu-sample', '1', '--nobs', '256']
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, k
    assert d['impl'] == 'reference' and d['unit'] == 'evals/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['e2e']['value'] == d['value']
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert d['cpu_baseline']['kind'] in ('port', 'reference') and d['cpu_baseline']['cores'] >= 1
    assert 'workload' in d['config']


def test_reference_arm_under_torchrun_env_prints_only_on_rank0():
    """N > 1: rank 0 alone runs and prints the line, the other ranks exit 0 without work."""
    env = dict(os.environ, OMP_NUM_THREADS='4', RANK='1', LOCAL_RANK='1', WORLD_SIZE='2')
    cmd = [sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1',
           '--warmup', '0', '--cpu-sample', '1', '--nobs', '256']
    out = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=ROOT, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip() == ''
