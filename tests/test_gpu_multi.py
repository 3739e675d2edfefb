"""2-rank slab decomposition against the single-rank run: over NCCL on a
multi-GPU box, over gloo with both ranks on GPU 0 otherwise (so the 1-GPU
test box runs it too), with and without bodies that change owner."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(port, **env):
    e = dict(os.environ)
    e.update(dict((k, str(v)) for k, v in env.items()))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
           '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(port), os.path.join(ROOT, 'tests',
                                                    'mgpu_equiv.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900,
                         env=e)
    assert 'MGPU_EQUIV_OK' in out.stdout, \
        out.stdout[-3000:] + out.stderr[-3000:]
    return out.stdout


def test_two_rank_run_equals_single_rank_and_oracle():
    """... and, over the first 300 steps (contact from step ~250 on), the CPU
    oracle run on the whole scene (1e-8 on xcm and R, frictionless)."""
    out = _run(29517, MGPU_BODIES=600, MGPU_STEPS=700, MGPU_ORACLE=300)
    assert 'oracle_xcm' in out


def test_two_rank_run_with_migrating_bodies_equals_single_rank():
    """Every body starts with vcm_x = 3 m/s: a column of bodies crosses the
    cut, ownership migrates (state, body-frame vectors, contact history), and
    the run still equals the single-rank one (frictionless, 1e-9)."""
    out = _run(29519, MGPU_BODIES=600, MGPU_STEPS=700, MGPU_LATERAL=3.0,
               MGPU_MIGRATE=50)
    assert 'migrated=0 ' not in out.splitlines()[-2]
