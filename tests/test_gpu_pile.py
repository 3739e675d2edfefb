"""BASELINE.json config 5 (the benchmarked configuration) on the CUDA path
against the oracle: small instances of scenes.synthetic_pile -- sparse contact
history, one uniform damping coefficient (eta_mode 2), hundreds of bodies in
one array, many simultaneous contacts.

The pile is settled on the GPU (cheap), then at several depths of settlement
its complete state (bodies, particles, sparse history) is handed to the
oracle's sparse-history mode and BOTH advance one step from that common
state: particle forces, per-body force/torque and the new history are held to
1e-10 (force law: /root/reference/code/rigid_body_common.py:839-1032,
reduction :128-175).  Free-running trajectories are held to 1e-6 over a
stated horizon of 40 steps from the settled state.
"""
import numpy as np
import pytest

from oracle import rbo
from tests.util import assert_close, assert_step_matches, oracle_twin

pytestmark = pytest.mark.gpu

DT = 1e-4


def _pile(nb, ks, **kw):
    from rigid_body_2d_3d_pysph_b200.device import DeviceScene
    from rigid_body_2d_3d_pysph_b200.scenes import synthetic_pile
    arrays, scheme, info = synthetic_pile(nb, seed=0)
    sc = DeviceScene(arrays, ['body'], ['wall'], dim=3, kr=1e5, kf=1e3,
                     fric_coeff=0.5, gy=-9.81, ks=ks,
                     eta_uniform=info['eta_uniform'], **kw)
    p = rbo.make_params(3, DT, 1e5, 1e3, 0.5, 0., -9.81, 0.,
                        eta_uniform=info['eta_uniform'])
    return arrays, sc, p, info


def _history_matches(what, sc, obody, ks):
    """Keys identical; fn of every active slot to 1e-10 of the largest
    normal force; delta_lt (a unit vector after quirk Q1) where the slot's
    friction direction is well conditioned."""
    n = obody.get_number_of_particles()
    hkey, hdlt, hfn = sc.history()
    okey = obody.sp_key.reshape(n, ks)
    ofn = obody.sp_fn.reshape(n, ks, 3)
    assert np.array_equal(hkey[:, :n].T, okey[:, :sc.ks]), what + ' keys'
    fscale = max(np.abs(ofn).max(), 1e-300)
    used = okey[:, :sc.ks] >= 0       # the device leaves unused entries stale
    for c in range(3):
        assert_close(hfn[c, :, :n].T[used], ofn[:, :sc.ks, c][used], 1e-10,
                     what + ' hist fn', fscale)
    return int((okey >= 0).sum())


@pytest.mark.parametrize('nb,ks', [(300, 8), (300, 4), (1000, 8)])
def test_pile_single_steps_match_oracle(nb, ks):
    arrays, sc, p, info = _pile(nb, ks)
    total_active = 0
    for k, settle in enumerate((1500, 1500, 2000, 3000)):
        sc.gtvf_step(DT, settle, graph=True)
        sc.check_status()
        for rep in range(2):              # two consecutive re-synchronised steps
            oarr = oracle_twin(sc, ks=ks)
            rbo.gtvf_step(oarr, ['body'], p, ks=ks, nsteps=1)
            sc.read_counters(reset=True)
            sc.gtvf_step(DT, 1)
            sc.sync_to_host()
            sc.check_status()
            what = 'pile nb=%d ks=%d after %d steps' % (nb, ks, sc.steps_done)
            st = assert_step_matches(what, sc, arrays, oarr, ['body'])
            nact = _history_matches(what, sc, oarr[0], ks)
            assert nact == st['active'], (what, nact, st)
            assert sc.read_counters(reset=True)['active_slots'] == nact, what
            total_active += nact
            # positions / orientation after the step
            for n in ('xcm', 'R', 'vcm', 'omega'):
                assert_close(getattr(arrays[0], n), getattr(oarr[0], n), 1e-9,
                             what + ' ' + n,
                             max(np.abs(getattr(oarr[0], n)).max(), 1e-2))
        # well-conditioned contacts must dominate, or the test has no teeth
        assert st['loose'] <= 0.5 * max(st['active'], 1) + 10, (what, st)
    # the settled pile really is in contact
    assert total_active > 20 * nb, total_active


def test_pile_trajectory_horizon():
    """Free-running: 40 steps from a settled state, xcm / R / particle
    positions within 1e-6 relative (north_star trajectory gate; horizon
    stated here)."""
    nb, ks = 300, 8
    arrays, sc, p, info = _pile(nb, ks)
    sc.gtvf_step(DT, 4000, graph=True)
    oarr = oracle_twin(sc, ks=ks)
    rbo.gtvf_step(oarr, ['body'], p, ks=ks, nsteps=40)
    sc.gtvf_step(DT, 40)
    sc.sync_to_host()
    sc.check_status()
    g, o = arrays[0], oarr[0]
    for n in ('xcm', 'R', 'x', 'y', 'z'):
        assert_close(getattr(g, n), getattr(o, n), 1e-6, 'pile horizon ' + n,
                     max(np.abs(getattr(o, n)).max(), 1e-2))
    assert sc.read_counters()['active_slots'] > 40 * 100


def test_pile_first_steps_from_rest():
    """The unsettled pile from its initial condition: the first contacts
    (a handful of blocks start inside the contact zone of the floor)."""
    nb, ks = 300, 8
    arrays, sc, p, info = _pile(nb, ks)
    oarr = oracle_twin(sc, ks=ks)
    done = 0
    for upto in (1, 10, 100):
        sc.gtvf_step(DT, upto - done)
        rbo.gtvf_step(oarr, ['body'], p, ks=ks, nsteps=upto - done)
        done = upto
        sc.sync_to_host()
        sc.check_status()
        assert_step_matches('pile from rest, step %d' % upto, sc, arrays,
                            oarr, ['body'], min_active=1)
