"""Neighbour pair sets of the PRODUCTION contact path, bit-exact.

north_star: "neighbour pair sets are bit-exact (compared as sorted index
pairs)".  The pairs compared here are the ones the contact kernel itself acts
on: every entry of the neighbour lists built by k_neighbours (an FP32 superset
with a skin, possibly several steps old) that passes the gate of
/root/reference/code/rigid_body_common.py:678-679 and the exact FP64 predicate
in k_slots -- dumped by that kernel (RbxDiag.pairs) -- against the oracle's
NNPS pairs filtered by the same gate.
"""
import numpy as np
import pytest

from oracle import rbo
from tests.util import load_case, load_config

pytestmark = pytest.mark.gpu


def _scene(arrays, meta, **kw):
    from rigid_body_2d_3d_pysph_b200.device import DeviceScene
    return DeviceScene(arrays, meta['rigid'], meta['boundaries'],
                       dim=meta['dim'], kr=meta['kr'], kf=meta['kf'],
                       fric_coeff=meta['fric_coeff'], gx=meta['gx'],
                       gy=meta['gy'], gz=meta['gz'],
                       planar=(meta.get('stepper') == 'gtvf2d'), **kw)


def oracle_gated_pairs(sc):
    """Sorted global (d, s) pairs from the oracle's NNPS on the CURRENT host
    state of the scene's arrays: s flagged contact_force_is_boundary == 1,
    dem_id different (rigid_body_common.py:678-679)."""
    sc.sync_to_host()
    arrays = sc.arrays
    names = [a.name for a in arrays]
    out = []
    for d in sc.rigid:
        di = names.index(d.name)
        for si, s in enumerate(arrays):
            if 'contact_force_is_boundary' not in s.properties:
                continue
            off, idx = rbo.nnps_pairs(arrays, di, si)
            i = np.repeat(np.arange(len(off) - 1), np.diff(off))
            j = idx.astype(np.int64)
            keep = (s.contact_force_is_boundary[j] == 1.) & \
                (d.dem_id[i] != s.dem_id[j])
            out.append(np.stack([i[keep] + sc.p_off[d.name],
                                 j[keep] + sc.p_off[s.name]], 1))
    out = np.concatenate(out).astype(np.int32) if out else \
        np.zeros((0, 2), np.int32)
    return out[np.lexsort((out[:, 1], out[:, 0]))]


def _check(sc, dt, what):
    got = sc.contact_pairs(dt)
    want = oracle_gated_pairs(sc)
    assert got.shape == want.shape and np.array_equal(got, want), (
        what, got.shape, want.shape)
    return want.shape[0]


@pytest.mark.parametrize('skin', [0.0, 0.05, 0.25])
@pytest.mark.parametrize('name', ['cubes3d', 'collide2d', 'wall2d'])
def test_production_pairs_golden(name, skin):
    arrays, ref, meta = load_case(name)
    sc = _scene(arrays, meta, skin_factor=skin, list_cap=192)
    total = _check(sc, meta['dt'], '%s skin %g initial' % (name, skin))
    for k in range(4):                   # lists reused (stale) in between
        sc.gtvf_step(meta['dt'], max(meta['nsteps'] // 4, 1))
        total += _check(sc, meta['dt'], '%s skin %g step %d' % (
            name, skin, sc.steps_done))
    assert total > 0


@pytest.mark.parametrize('skin', [0.0, 0.05, 0.25])
@pytest.mark.parametrize('name,steps', [('benchmark_5_3d', (1, 60, 60, 30)),
                                        ('stack_of_cylinders', (1, 50, 100))])
def test_production_pairs_configs(name, steps, skin):
    arrays, meta = load_config(name)
    sc = _scene(arrays, meta, skin_factor=skin, list_cap=192)
    for n in steps:
        sc.gtvf_step(meta['dt'], n)
        assert _check(sc, meta['dt'], '%s skin %g step %d' % (
            name, skin, sc.steps_done)) > 0
    if skin > 0:
        # the lists were in fact reused: fewer entries written than steps x
        # one build
        cnt = sc.read_counters()
        assert cnt['list_entries'] > 0


@pytest.mark.parametrize('skin', [0.05, 0.25])
def test_production_pairs_pile(skin):
    from rigid_body_2d_3d_pysph_b200.device import DeviceScene
    from rigid_body_2d_3d_pysph_b200.scenes import synthetic_pile
    arrays, scheme, info = synthetic_pile(300, seed=0)
    sc = DeviceScene(arrays, ['body'], ['wall'], dim=3, gy=-9.81,
                     eta_uniform=info['eta_uniform'], skin_factor=skin,
                     list_cap=160)
    for n in (1, 700, 1500):
        sc.gtvf_step(1e-4, n, graph=True)
        sc.check_status()
        assert _check(sc, 1e-4, 'pile skin %g step %d' % (
            skin, sc.steps_done)) > 300000
