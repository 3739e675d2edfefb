"""bench.py contract on CPU: the reference arm prints ONE JSON line with the
keys the driver reads, and the b200 arm fails loudly without a GPU (no CPU
fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] +
                          list(args), capture_output=True, text=True,
                          timeout=600, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    out = _run('--impl', 'reference', '--steps', '2', '--warmup', '1',
               '--cpu-bodies', '48', '--cpu-settle', '2')
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d['impl'] == 'reference'
    assert d['metric'] == 'particle-updates/sec' and d['value'] > 0
    assert d['steps'] == 2 and d['warmup'] == 1 and d['n_gpus'] == 1
    assert d['higher_is_better'] is True and d['dtype'] == 'f64'
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['sample']
    assert cb['value'] == d['value']
    e = d['e2e']
    assert e['value'] == d['value'] and e['unit'] == d['unit']
    assert e['h2d_bytes_per_step'] == 0 and e['d2h_bytes_per_step'] == 0
    assert 'workload' in d['config'] and 'model' not in d['config']


@pytest.mark.skipif(torch.cuda.is_available(), reason='CPU-only check')
def test_b200_arm_needs_a_gpu():
    out = _run('--steps', '1', '--warmup', '1', '--bodies', '8', '--settle',
               '0', '--no-cpu')
    assert out.returncode != 0
    assert not [ln for ln in out.stdout.splitlines() if ln.startswith('{')]


def test_stdout_carries_only_the_json_line():
    """Native libraries print to file descriptor 1 (the NCCL version banner
    at N > 1): bench.py moves descriptor 1 to stderr and writes its one JSON
    line to the real stdout."""
    code = ('import os, sys; sys.path.insert(0, %r); import bench; '
            'fd = bench.claim_stdout(); os.write(1, b"NCCL version x\\n"); '
            'print("python noise"); sys.stdout.flush(); '
            'bench.emit(fd, \'{"a": 1}\')' % ROOT)
    r = subprocess.run([sys.executable, '-c', code], capture_output=True,
                       text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout == '{"a": 1}\n'
    assert 'NCCL version x' in r.stderr and 'python noise' in r.stderr
