"""CPU: the parts of bench.py's contract that do not need a GPU -- the reference arm prints one JSON line with the keys the
driver reads, and the product arm refuses to run without a CUDA device (there is no CPU path)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), *args], capture_output=True, text=True, timeout=timeout,
                          cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    if not os.path.isfile(os.path.join(ROOT, 'baseline', '_ref', 'train_parts', 'generators.py')):
        pytest.skip('baseline/_ref snapshot missing')
    r = _run('--impl', 'reference', '--steps', '1', '--warmup', '0', '--workload', 'sg2ada64')
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
    assert len(lines) == 1, r.stderr[-2000:]
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'img/s' and d['higher_is_better'] is True and d['value'] > 0
    for k in ('metric', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'scaling', 'vs_baseline', 'dtype', 'data', 'config'):
        assert k in d, k
    assert d['cpu_baseline']['kind'] in ('reference', 'port') and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['sample']
    assert d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == dict(value=d['value'], unit='img/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0)


def test_reference_arm_only_rank_zero_prints():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and not [ln for ln in r.stdout.splitlines() if ln.startswith('{')]


def test_product_arm_needs_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a CUDA device is present')
    r = _run('--lean', '--steps', '1', timeout=300)
    assert r.returncode != 0 and 'no CUDA device' in (r.stderr + r.stdout)
