"""C oracle of the Canelas Hertz equations against the reference's own
RigidBodyCanelasRigidRigid / RigidBodyCanelasRigidWall .loop
(rigid_body_common.py:244-628; fixture: oracle/make_golden.py canelas2d)."""
import numpy as np

from oracle import rbo
from tests.util import assert_close, load_case


def canelas_case():
    """Scene = the state after stages 1 + 2 of one GTVF step (the particles
    carry their rigid-body velocities), just before the force evaluation."""
    return load_case('canelas2d')


def test_canelas_oracle_matches_reference():
    arrays, ref, meta = canelas_case()
    body = arrays[0]
    p = rbo.make_params(meta['dim'], meta['dt'], gx=meta['gx'], gy=meta['gy'],
                        gz=meta['gz'])
    rbo.canelas(arrays, meta['rigid'], p, Cn=meta['Cn'])
    pre = 'ref/1/body/'
    f = np.sqrt(ref[pre + 'fx']**2 + ref[pre + 'fy']**2).sum()
    assert np.abs(ref[pre + 'fy']).max() > 1e3 * body.m[0] * 9.81   # contacts
    for n in ('fx', 'fy', 'fz', 'force'):
        assert_close(getattr(body, n), ref[pre + n], 1e-10, n, f)
    assert_close(body.torque, ref[pre + 'torque'], 1e-10, 'torque',
                 f * 4 * 0.025)
