"""CUDA Canelas Hertz contact (rbx_contact_canelas, through the C ABI) against
the reference's own RigidBodyCanelasRigidRigid / RigidBodyCanelasRigidWall
.loop (fixture canelas2d) and against the C oracle on a 3-D scene."""
import numpy as np
import pytest

from oracle import rbo
from tests.util import assert_close, load_case

pytestmark = pytest.mark.gpu


def _scene(arrays, meta):
    from rigid_body_2d_3d_pysph_b200.device import DeviceScene
    return DeviceScene(arrays, meta['rigid'], meta['boundaries'],
                       dim=meta['dim'], kr=meta['kr'], kf=meta['kf'],
                       fric_coeff=meta['fric_coeff'], gx=meta['gx'],
                       gy=meta['gy'], gz=meta['gz'])


def test_canelas_matches_reference():
    arrays, ref, meta = load_case('canelas2d')
    body = arrays[0]
    sc = _scene(arrays, meta)
    sc.contact_canelas(meta['dt'], Cn=meta['Cn'])
    sc.reduce_bodies()
    sc.check_status()
    pre = 'ref/1/body/'
    f = np.sqrt(ref[pre + 'fx']**2 + ref[pre + 'fy']**2).sum()
    assert np.abs(ref[pre + 'fy']).max() > 1e3 * body.m[0] * 9.81
    for n in ('fx', 'fy', 'fz', 'force'):
        assert_close(getattr(body, n), ref[pre + n], 1e-10, n, f)
    assert_close(body.torque, ref[pre + 'torque'], 1e-10, 'torque',
                 f * 4 * 0.025)


def test_canelas_3d_matches_oracle():
    """Two 3^3 cubes on a floor (the cubes3d scene), pressed 4 % of a radius
    into each other and into the floor."""
    arrays, _, meta = load_case('cubes3d')
    oarrays, _, _ = load_case('cubes3d')
    for arrs in (arrays, oarrays):
        b = arrs[0]
        sel = b.body_id == 1
        b.y[sel] -= 0.05 * 0.05
        b.y[:] -= 0.03 * 0.05
        b.u[:] = 0.2
        b.v[sel] = -0.3
    sc = _scene(arrays, meta)
    sc.contact_canelas(meta['dt'])
    sc.reduce_bodies()
    p = rbo.make_params(meta['dim'], meta['dt'], gx=meta['gx'], gy=meta['gy'],
                        gz=meta['gz'])
    rbo.canelas(oarrays, meta['rigid'], p)
    g, o = arrays[0], oarrays[0]
    f = np.sqrt(o.fx**2 + o.fy**2 + o.fz**2).sum()
    assert np.abs(o.fy).max() > 100 * o.m[0] * 9.81
    for n in ('fx', 'fy', 'fz', 'force'):
        assert_close(getattr(g, n), getattr(o, n), 1e-10, n, f)
    assert_close(g.torque, o.torque, 1e-10, 'torque', f * 0.2)
