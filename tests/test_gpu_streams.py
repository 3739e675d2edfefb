"""Host-driven boundaries through the streamed API (DeviceScene.static_update,
BoundaryStream, BodyStateStream) against the plain one (the host writes
``pa.x``): the same moving floor, the same results.

An Application's post_step may move a boundary (the reference does it once,
/root/reference/code/stack_of_cylinders.py:438-445).  Writing ``pa.x`` uploads
synchronously and rebuilds the neighbour lists unconditionally; the streamed
path uploads on a copy stream and rebuilds only when a particle has moved
more than half the skin since the lists were built -- and the pair set, hence
the forces, must not depend on when lists are rebuilt."""
import numpy as np
import pytest

from tests.util import assert_close, load_case

pytestmark = pytest.mark.gpu


def _scene(arrays, meta, **kw):
    from rigid_body_2d_3d_pysph_b200.device import DeviceScene
    return DeviceScene(arrays, meta['rigid'], meta['boundaries'],
                       dim=meta['dim'], kr=meta['kr'], kf=meta['kf'],
                       fric_coeff=meta['fric_coeff'], gx=meta['gx'],
                       gy=meta['gy'], gz=meta['gz'], **kw)


@pytest.mark.parametrize('amp', [2e-5, 4e-4])
def test_moving_floor_streamed_equals_host_writes(amp):
    """cubes3d with the tank shaken vertically: amp = 2e-5 per step stays
    inside the skin for many steps (the streamed path reuses its lists), 4e-4
    leaves it every few steps."""
    import torch
    from rigid_body_2d_3d_pysph_b200.device import (BodyStateStream,
                                                   BoundaryStream)
    nsteps = 40
    a1, _, meta = load_case('cubes3d')
    a2, _, _ = load_case('cubes3d')
    s1, s2 = _scene(a1, meta), _scene(a2, meta)
    tank1, tank2 = a1[1], a2[1]
    y0 = tank1.y.copy()
    feeder = BoundaryStream(s2, 'tank', ('x', 'y', 'z', 'u', 'v', 'w'))
    reader = BodyStateStream(s2, ('xcm', 'force'))
    host = [dict((n, torch.from_numpy(np.ascontiguousarray(
        tank2.properties[n])).clone().pin_memory())
        for n in ('x', 'y', 'z', 'u', 'v', 'w')) for _ in range(2)]
    last = -1
    for k in range(nsteps):
        dy = amp * np.sin(0.7 * k) * (k + 1)
        vy = amp * 0.7 * np.cos(0.7 * k) * (k + 1) / meta['dt']
        # plain path: the host writes the array
        tank1.y[:] = y0 + dy
        tank1.v[:] = vy
        tank1.touch('y', 'v')
        s1.gtvf_step(meta['dt'], 1)
        # streamed path
        h = host[k & 1]
        feeder.ev_up[k & 1].synchronize()
        h['y'].copy_(torch.from_numpy(y0 + dy))
        h['v'].fill_(vy)
        feeder.submit(h)
        feeder.apply()
        s2.gtvf_step(meta['dt'], 1)
        j = reader.snapshot()
        if last >= 0:
            reader.wait(last)
        last = j
    got = reader.wait(last)
    s1.check_status()
    s2.check_status()
    b1, b2 = a1[0], a2[0]
    f = np.sqrt(b1.fx**2 + b1.fy**2 + b1.fz**2).sum()
    assert f > 3 * np.abs(b1.m * 9.81).sum(), "the floor is not felt"
    for n in ('fx', 'fy', 'fz', 'force'):
        assert_close(getattr(b2, n), getattr(b1, n), 1e-11, n, f)
    for n in ('xcm', 'R', 'vcm', 'omega', 'x', 'y', 'u', 'v'):
        assert_close(getattr(b2, n), getattr(b1, n), 1e-11, n)
    # the streamed results are what the scene holds
    assert np.array_equal(got['xcm'].numpy(), b2.xcm)
    assert np.array_equal(got['force'].numpy(), b2.force)
    # and the streamed path did reuse its lists where the plain one rebuilt
    e1 = s1.read_counters()['list_entries']
    e2 = s2.read_counters()['list_entries']
    assert e2 <= e1
    if amp < 1e-4:
        assert e2 < 0.5 * e1, (e1, e2)
