"""Per-step force parity INSIDE contact on the named configurations.

tests/test_gpu_configs.py compares free-running runs, which the chaotic
contact model (quirk Q1) lets drift apart a few steps after first contact.
Here the CUDA path runs into contact, its complete state and contact history
are handed to the oracle (the reference's dense tnb-strided slot layout), and
both advance ONE step from that common state: particle forces, per-body
force/torque within 1e-10 (scale: the terms of the particle's force), the new
history compared slot by slot.  Force law:
/root/reference/code/rigid_body_common.py:839-1032, reduction :128-175.
"""
import numpy as np
import pytest

from oracle import rbo
from tests.util import (assert_close, assert_step_matches, load_config,
                        oracle_params, oracle_twin)

pytestmark = pytest.mark.gpu


def _scene(arrays, meta, **kw):
    from rigid_body_2d_3d_pysph_b200.device import DeviceScene
    return DeviceScene(arrays, meta['rigid'], meta['boundaries'],
                       dim=meta['dim'], kr=meta['kr'], kf=meta['kf'],
                       fric_coeff=meta['fric_coeff'], gx=meta['gx'],
                       gy=meta['gy'], gz=meta['gz'],
                       planar=(meta['stepper'] == 'gtvf2d'), **kw)


def _dense_history_matches(what, sc, oarr, rigid):
    """fn of every active slot against the oracle's dense fn_* arrays."""
    hkey, hdlt, hfn = sc.history()
    nact = 0
    for o in oarr:
        if o.name not in rigid:
            continue
        off = sc.p_off[o.name]
        n = o.get_number_of_particles()
        tnb = int(o.total_no_bodies[0])
        ii = np.arange(n)
        active = o.overlap.reshape(n, tnb) > 0.
        got_active = np.zeros((n, tnb), dtype=bool)
        fscale = max(np.abs(o.fn_x).max(), np.abs(o.fn_y).max(),
                     np.abs(o.fn_z).max(), 1e-300)
        for s in range(sc.ks):
            key = hkey[s, off:off + n]
            m = key >= 0
            got_active[ii[m], key[m]] = True
            t2 = tnb * ii[m] + key[m]
            for c, ax in enumerate('xyz'):
                assert_close(hfn[c, s, off:off + n][m],
                             o.properties['fn_' + ax][t2], 1e-10,
                             '%s hist fn_%s' % (what, ax), fscale)
        assert np.array_equal(active, got_active), what + ': active slot set'
        nact += int(active.sum())
    return nact


@pytest.mark.parametrize('name,steps', [
    # cubes reach the floor near step 105
    ('benchmark_5_3d', (120, 80, 200)),
    ('stack_of_cylinders', (20, 180, 400)),
    # first contact between steps 2000 and 3000
    ('benchmark_3', (2600, 400, 600)),
    ('benchmark_4', (2450, 100, 150)),   # in contact 2380..2760, then rebounds
    ('benchmark_5_2d', (200, 300)),
    ('benchmark_2', (1100, 200)),
])
def test_single_step_in_contact(name, steps):
    garr, meta = load_config(name)
    planar = meta['stepper'] == 'gtvf2d'
    sc = _scene(garr, meta)
    p = oracle_params(meta)
    total = 0
    for n in steps:
        sc.gtvf_step(meta['dt'], n, graph=True)
        sc.check_status()
        for rep in range(3):
            oarr = oracle_twin(sc, ks=0)
            rbo.gtvf_step(oarr, meta['rigid'], p, planar=planar, nsteps=1)
            sc.gtvf_step(meta['dt'], 1)
            sc.sync_to_host()
            sc.check_status()
            what = '%s step %d' % (name, sc.steps_done)
            st = assert_step_matches(what, sc, garr, oarr, meta['rigid'])
            nact = _dense_history_matches(what, sc, oarr, meta['rigid'])
            assert nact == st['active'], (what, nact, st)
            total += nact
            for g, o in zip(garr, oarr):
                if g.name in meta['rigid']:
                    for nme in ('xcm', 'R', 'vcm', 'omega'):
                        assert_close(getattr(g, nme), getattr(o, nme), 1e-9,
                                     what + ' ' + nme,
                                     max(np.abs(getattr(o, nme)).max(), 1e-2))
    assert total > 0, name + ': never in contact'
