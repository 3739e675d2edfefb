"""Shared helpers for the parity tests: golden fixtures and comparisons."""
import json
import os

import numpy as np

from rigid_body_2d_3d_pysph_b200.compat.output import load_scene

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

CASES = ['free2d_gtvf2d', 'free2d_gtvf3d', 'wall2d', 'wall2d_planar',
         'wall2d_rest', 'wall2d_normal', 'collide2d', 'cubes3d', 'rk2_3d']

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
