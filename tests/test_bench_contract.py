"""bench.py's contract where it can be checked without a GPU: the reference arm prints exactly one JSON line
with the agreed keys (it times the CPU oracle = the reference algorithm), and the CUDA arm refuses to run —
loudly, no CPU fallback — when there is no device."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ, **(env or {}))
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), *args], cwd=ROOT, env=e,
                          capture_output=True, text=True, timeout=300)


@pytest.mark.timeout(300)
def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run('--impl', 'reference', '--steps', '2', '--warmup', '0', '--ref-seconds', '0.2')
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'vs_baseline', 'dtype', 'data', 'config', 'impl', 'cpu_baseline', 'e2e'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['steps'] == 2 and d['higher_is_better'] is True and d['vs_baseline'] is None
    assert d['unit'] == 'columns/s' and d['value'] > 0 and 'workload' in d['config']
    assert d['cpu_baseline']['kind'] in ('port', 'reference') and d['cpu_baseline']['cores'] >= 1
    assert d['cpu_baseline']['value'] == d['value'] == d['e2e']['value']
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0


@pytest.mark.timeout(300)
def test_reference_arm_under_torchrun_only_rank0_works():
    env = {'RANK': '1', 'LOCAL_RANK': '1', 'WORLD_SIZE': '2'}
    r = _run('--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0', '--ref-seconds', '0.1', env=env)
    assert r.returncode == 0 and r.stdout.strip() == ''


@pytest.mark.timeout(300)
def test_cuda_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    r = _run('--steps', '1', '--warmup', '0', '--no-cpu-baseline', '--cols', '4096')
    assert r.returncode != 0 and r.stdout.strip() == ''
    assert 'no CPU fallback' in r.stderr or 'no CUDA device' in r.stderr
