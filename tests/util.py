"""Shared helpers for the parity tests: golden fixtures and comparisons."""
import json
import os

import numpy as np

from rigid_body_2d_3d_pysph_b200.compat.output import load_scene

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

CASES = ['free2d_gtvf2d', 'free2d_gtvf3d', 'wall2d', 'wall2d_planar',
         'wall2d_rest', 'wall2d_normal', 'collide2d', 'cubes3d', 'rk2_3d',
         'rk2_3d_nb2']

# dense slot arrays compared against the reference (ti_* excluded: quirk Q5)
SLOT_PROPS = ['contact_force_normal_x', 'contact_force_normal_y',
              'contact_force_normal_z', 'contact_force_dist', 'overlap',
              'ft_x', 'ft_y', 'ft_z', 'fn_x', 'fn_y', 'fn_z',
              'delta_lt_x', 'delta_lt_y', 'delta_lt_z',
              'vx_source', 'vy_source', 'vz_source',
              'x_source', 'y_source', 'z_source',
              'closest_point_dist_to_source']
STATE = ['x', 'y', 'z', 'u', 'v', 'w', 'fx', 'fy', 'fz', 'force', 'torque',
         'xcm', 'vcm', 'omega', 'ang_mom', 'R']


def load_case(name):
    arrays, _ = load_scene(os.path.join(GOLDEN, name + '_scene.npz'))
    ref = np.load(os.path.join(GOLDEN, name + '_ref.npz'))
    meta = json.loads(str(ref['__meta__']))
    return arrays, ref, meta


def assert_close(got, want, rtol, what, scale=None):
    """|got - want| <= rtol * scale, scale = max|want| unless given (sums
    that cancel are compared against the size of their terms); NaNs must
    coincide (quirk Q2 puts NaN into delta_lt)."""
    got = np.asarray(got, dtype=float)
    want = np.asarray(want, dtype=float)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    assert np.array_equal(nan_g, nan_w), '%s: NaN pattern differs' % what
    ok = ~nan_w
    if not ok.any():
        return
    if scale is None:
        scale = max(np.max(np.abs(want[ok])), 1e-300)
    err = np.max(np.abs(got[ok] - want[ok]))
    assert err <= rtol * scale, '%s: err %.3e > %.1e * %.3e' % (
        what, err, rtol, scale)


DENSE_SLOTS = ['contact_force_normal_x', 'contact_force_normal_y',
               'contact_force_normal_z', 'contact_force_normal_wij',
               'contact_force_normal_tmp_x', 'contact_force_normal_tmp_y',
               'contact_force_normal_tmp_z', 'contact_force_dist_tmp',
               'contact_force_dist', 'overlap', 'ft_x', 'ft_y', 'ft_z',
               'fn_x', 'fn_y', 'fn_z', 'delta_lt_x', 'delta_lt_y',
               'delta_lt_z', 'vx_source', 'vy_source', 'vz_source',
               'x_source', 'y_source', 'z_source', 'ti_x', 'ti_y', 'ti_z',
               'closest_point_dist_to_source']


def load_config(name):
    """A BASELINE.json config scene frozen by tests/make_config_fixtures.py:
    (arrays, meta).  Rigid arrays get the reference's dense slot arrays so
    that the oracle can run them."""
    arrays, meta = load_scene(os.path.join(GOLDEN, 'cfg_%s.npz' % name))
    for pa in arrays:
        if pa.name not in meta['rigid']:
            continue
        if 'spacing0' not in pa.constants:      # divergence D5
            pa.add_constant('spacing0', pa.constants['initial_spacing0'])
        tnb = int(pa.total_no_bodies[0])
        for n in DENSE_SLOTS:
            if n not in pa.properties:
                pa.add_property(n, stride=tnb)
        if 'dem_id_source' not in pa.properties:
            pa.add_property('dem_id_source', type='int', stride=tnb)
    return arrays, meta


def oracle_params(meta, **kw):
    from oracle import rbo
    return rbo.make_params(meta['dim'], meta['dt'], meta['kr'], meta['kf'],
                           meta['fric_coeff'], meta['gx'], meta['gy'],
                           meta['gz'], **kw)


# ---------------------------------------------------------------------------
# re-synchronised single-step comparisons (the contact model is chaotic:
# quirk Q1 caps friction along the direction of a tangential velocity that
# can be rounding noise, so free-running trajectories of two correct
# implementations separate; one step from a COMMON state does not)
# ---------------------------------------------------------------------------
def clone_arrays(arrays):
    """Deep host copy of a list of ParticleArrays (no device binding)."""
    from rigid_body_2d_3d_pysph_b200.compat.particle_array import \
        get_particle_array
    out = []
    for pa in arrays:
        q = get_particle_array(name=pa.name)
        q.__dict__['_n'] = pa.get_number_of_particles()
        q.__dict__['num_real_particles'] = pa.get_number_of_particles()
        for n, v in pa.properties.items():
            q.add_property(n, type=pa.property_types[n], data=v,
                           stride=pa.stride[n])
        for n, v in pa.constants.items():
            q.add_constant(n, v)
        out.append(q)
    return out


def oracle_twin(sc, ks=0):
    """Host copy of the CURRENT state of DeviceScene ``sc`` for the oracle:
    every array cloned after a device->host sync, the contact history of the
    rigid arrays converted to the oracle's layout -- sparse (ks > 0:
    sp_key/sp_delta_lt/sp_fn) or the reference's dense tnb-strided slot
    arrays (ks == 0) -- plus the test-aid property ft_cond."""
    from oracle import rbo
    sc.sync_to_host()
    oarr = clone_arrays(sc.arrays)
    hkey, hdlt, hfn = sc.history()
    for pa in oarr:
        if pa.name not in [r.name for r in sc.rigid]:
            continue
        o = sc.p_off[pa.name]
        n = pa.get_number_of_particles()
        key = hkey[:, o:o + n]
        if 'ft_cond' not in pa.properties:
            pa.add_property('ft_cond')
        if ks > 0:
            assert sc.ks <= ks
            rbo.add_sparse_history(pa, ks)
            k = pa.properties['sp_key'].reshape(n, ks)
            d = pa.properties['sp_delta_lt'].reshape(n, ks, 3)
            f = pa.properties['sp_fn'].reshape(n, ks, 3)
            k[:, :sc.ks] = key.T
            for c in range(3):
                d[:, :sc.ks, c] = hdlt[c, :, o:o + n].T
                f[:, :sc.ks, c] = hfn[c, :, o:o + n].T
            used = k >= 0
            d[~used] = 0.
            f[~used] = 0.
        else:
            if 'spacing0' not in pa.constants:      # divergence D5
                pa.add_constant('spacing0', pa.constants['initial_spacing0'])
            tnb = int(pa.total_no_bodies[0])
            for nme in DENSE_SLOTS:
                if nme not in pa.properties:
                    pa.add_property(nme, stride=tnb)
                else:
                    pa.properties[nme][:] = 0.
            if 'dem_id_source' not in pa.properties:
                pa.add_property('dem_id_source', type='int', stride=tnb)
            ii = np.arange(n)
            for s in range(sc.ks):
                m = key[s] >= 0
                t2 = tnb * ii[m] + key[s][m]
                for c, ax in enumerate('xyz'):
                    pa.properties['delta_lt_' + ax][t2] = hdlt[c, s, o:o + n][m]
                    pa.properties['fn_' + ax][t2] = hfn[c, s, o:o + n][m]
    return oarr


def assert_step_matches(what, sc, garr, oarr, rigid, rtol=1e-10,
                        min_active=0):
    """After ONE step from a common state: particle forces and per-body
    force / torque of the CUDA path against the oracle at ``rtol`` relative
    to the sum of the magnitudes of the terms that make up the force of that
    particle (body): |m g|, and per active slot the two terms of
    kr * (spacing0 - dist) (rigid_body_common.py:906, 921) and the force
    itself -- widened by the oracle's friction sensitivity ft_cond
    (:961-1007: the friction force has magnitude <= mu |fn| along t / |t|, so
    rounding noise in t moves it by that over |t|; 0 for well-conditioned
    contacts).  Returns statistics."""
    stats = {'particles': 0, 'loose': 0, 'active': 0, 'max_err': 0.}
    g3 = np.sqrt(sc.g[0]**2 + sc.g[1]**2 + sc.g[2]**2)
    for g, o in zip(garr, oarr):
        if g.name not in rigid:
            continue
        n = g.get_number_of_particles()
        if 'sp_key' in o.properties:
            nact = (o.sp_key.reshape(n, -1) >= 0).sum(1)
        else:
            nact = (o.overlap.reshape(n, -1) > 0.).sum(1)
        f = np.sqrt(o.fx**2 + o.fy**2 + o.fz**2)
        terms = o.m * g3 + f + nact * 2. * sc.kr * float(o.spacing0[0])
        cond = o.ft_cond
        tol = rtol * terms + cond
        err = np.zeros(n)
        for nme in ('fx', 'fy', 'fz'):
            gv, ov = getattr(g, nme), getattr(o, nme)
            assert not np.isnan(gv).any(), '%s %s has NaN' % (what, nme)
            err = np.maximum(err, np.abs(gv - ov))
        bad = np.nonzero(err > tol)[0]
        if bad.size:
            w = bad[np.argmax(err[bad] / tol[bad])]
            raise AssertionError(
                '%s: %d particle forces off; particle %d: err %.3e > tol %.3e '
                '(terms %.3e, cond %.3e, %d active slots)' % (
                    what, bad.size, w, err[w], tol[w], terms[w], cond[w],
                    nact[w]))
        stats['particles'] += n
        stats['loose'] += int((cond > rtol * terms).sum())
        stats['max_err'] = max(stats['max_err'],
                               float((err / np.maximum(terms, 1e-300)).max()))
        bid = g.body_id
        nb = int(g.nb[0])
        tb = np.bincount(bid, weights=tol, minlength=nb)
        lever = np.sqrt(o.dx0**2 + o.dy0**2 + o.dz0**2).max()
        ferr = np.abs(g.force - o.force).reshape(nb, 3).max(1)
        terr = np.abs(g.torque - o.torque).reshape(nb, 3).max(1)
        assert (ferr <= tb).all(), '%s: body force err %.3e over tol' % (
            what, (ferr - tb).max())
        assert (terr <= 2. * lever * tb).all(), \
            '%s: torque err %.3e over tol' % (what,
                                              (terr - 2. * lever * tb).max())
        stats['active'] += int(nact.sum())
    assert stats['active'] >= min_active, (what, stats)
    return stats
