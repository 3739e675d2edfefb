"""2-rank slab decomposition on real GPUs (NCCL): skipped on a 1-GPU box."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_run_equals_single_rank():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
           '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29517', os.path.join(ROOT, 'tests',
                                                  'mgpu_equiv.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert 'MGPU_EQUIV_OK' in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
