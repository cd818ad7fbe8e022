"""CPU: the bench.py contract that can be checked without a GPU.

* `--impl reference` times the oracle port of the reference's path on the host cores and prints ONE JSON line with the
  keys the driver reads (tier framing (4)).
* our arm has no CPU fallback: without a CUDA device it must fail loudly, not print a number.
"""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ['--leaves', '300', '--sites', '300', '--queries-per-gpu', '64', '--steps', '1', '--warmup', '1']


def _run(extra, env_extra=None):
    env = dict(os.environ)
    env.pop('RANK', None)
    env.pop('WORLD_SIZE', None)
    if env_extra:
        env.update(env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + SMALL + extra, capture_output=True,
                          text=True, cwd=ROOT, env=env, timeout=600)


def test_reference_arm_prints_one_json_line():
    p = _run(['--impl', 'reference', '--cpu-sample', '8'])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference'
    assert d['metric'] == 'queries placed/sec' and d['unit'] == 'queries/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['steps'] == 1 and d['warmup'] == 1 and d['n_gpus'] == 1
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    assert d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'queries/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config']


def test_reference_arm_other_ranks_exit_quietly():
    p = _run(['--impl', 'reference', '--cpu-sample', '8'], {'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert p.returncode == 0, p.stderr[-2000:]
    assert p.stdout.strip() == ''


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the behaviour on a box without a CUDA device')
def test_our_arm_fails_loudly_without_cuda():
    p = _run(['--no-cpu-baseline', '--no-e2e'])
    assert p.returncode != 0
    assert p.stdout.strip() == ''
    assert 'CUDA' in p.stderr
