"""Two-precision contact evaluation (k_filter + k_slots<COMPACT>) against the
one-pass FP64 evaluation (RBX_PARAM_EXACT): identical results, bit for bit.

The FP32 first pass may only drop a (particle, source body) slot when the
contact condition of /root/reference/code/rigid_body_common.py:906-907 fails
for every value its sums can take inside the running error bound; whatever it
keeps goes through the same FP64 code as the one-pass evaluation.  So fx, fy,
fz, the history and the trajectories must be EQUAL, not close -- on every
scene, at every distance from the origin, and for gaps that straddle
spacing0."""
import numpy as np
import pytest

from rigid_body_2d_3d_pysph_b200.compat.particle_array import \
    get_particle_array
from tests.util import CASES, load_case, load_config

pytestmark = pytest.mark.gpu


def _scene(arrays, meta, **kw):
    from rigid_body_2d_3d_pysph_b200.device import DeviceScene
    return DeviceScene(arrays, meta['rigid'], meta['boundaries'],
                       dim=meta['dim'], kr=meta['kr'], kf=meta['kf'],
                       fric_coeff=meta['fric_coeff'], gx=meta['gx'],
                       gy=meta['gy'], gz=meta['gz'],
                       planar=(meta.get('stepper') == 'gtvf2d'), **kw)


def _equal(what, a, b, sa, sb):
    for pa, pb in zip(a, b):
        if pa.name not in [r.name for r in sa.rigid]:
            continue
        for n in ('fx', 'fy', 'fz', 'x', 'y', 'z', 'u', 'v', 'w', 'xcm', 'R',
                  'vcm', 'omega', 'force', 'torque'):
            assert np.array_equal(getattr(pa, n), getattr(pb, n)), \
                '%s: %s.%s differs' % (what, pa.name, n)
    for ha, hb in zip(sa.history(), sb.history()):
        assert np.array_equal(ha, hb, equal_nan=True), what + ': history'
    ca, cb = sa.read_counters(), sb.read_counters()
    assert ca['active_slots'] == cb['active_slots'], (what, ca, cb)


@pytest.mark.parametrize('name', [c for c in CASES if c != 'rk2_3d'])
def test_fast_equals_exact_golden(name):
    a, _, meta = load_case(name)
    b, _, _ = load_case(name)
    sa, sb = _scene(a, meta), _scene(b, meta, exact=True)
    for k in range(4):
        n = max(meta['nsteps'] // 4, 1)
        sa.gtvf_step(meta['dt'], n)
        sb.gtvf_step(meta['dt'], n)
        sa.check_status()
        sb.check_status()
        _equal('%s step %d' % (name, sa.steps_done), a, b, sa, sb)


@pytest.mark.parametrize('name,steps', [
    ('benchmark_2', (1040, 100, 160)),
    ('benchmark_5_3d', (100, 20, 80, 200)),
    ('stack_of_cylinders', (10, 190, 300)),
    ('benchmark_4', (100, 200)),
])
def test_fast_equals_exact_configs(name, steps):
    a, meta = load_config(name)
    b, _ = load_config(name)
    sa, sb = _scene(a, meta), _scene(b, meta, exact=True)
    for n in steps:
        sa.gtvf_step(meta['dt'], n, graph=True)
        sb.gtvf_step(meta['dt'], n, graph=True)
        sa.check_status()
        sb.check_status()
        _equal('%s step %d' % (name, sa.steps_done), a, b, sa, sb)


def _pile_pair(nb, **kw):
    from rigid_body_2d_3d_pysph_b200.device import DeviceScene
    from rigid_body_2d_3d_pysph_b200.scenes import synthetic_pile
    out = []
    for exact in (False, True):
        arrays, scheme, info = synthetic_pile(nb, seed=0)
        if kw.get('shift'):
            for pa in arrays:
                for k, n in enumerate('xyz'):
                    pa.properties[n] += kw['shift'][k]
            arrays[0].xcm.reshape(-1, 3)[:] += np.asarray(kw['shift'])
        sc = DeviceScene(arrays, ['body'], ['wall'], dim=3, gy=-9.81,
                         eta_uniform=info['eta_uniform'], exact=exact)
        out.append((arrays, sc))
    return out


@pytest.mark.parametrize('shift', [None, (1000., -300., 5000.),
                                   (3e5, 0., -7e5)])
def test_fast_equals_exact_pile(shift):
    """The benchmarked scene, also far from the origin of the FP32
    coordinates' frame -- the frame is the scene's own centre, so the shift
    must not matter -- settled in steps."""
    (a, sa), (b, sb) = _pile_pair(400, shift=shift)
    for n in (1, 1500, 2500, 2000):
        sa.gtvf_step(1e-4, n, graph=True)
        sb.gtvf_step(1e-4, n, graph=True)
        sa.check_status()
        sb.check_status()
        _equal('pile shift %s step %d' % (shift, sa.steps_done), a, b, sa, sb)
    assert sa.read_counters()['active_slots'] > 0
    # and the first pass does exclude most of the work
    assert int(sa.counters[6].item()) < 0.9 * sa.n_rigid


def test_filter_off_centre_frame():
    """A scene whose FP32 frame is far from its particles (origin forced
    10 km away): the error bound grows, more slots go to the exact pass, the
    results stay identical."""
    (a, sa), (b, sb) = _pile_pair(200)
    sa.origin = [1e4, -1e4, 1e4]
    sa._refresh_structs()
    for n in (800, 2500):
        sa.gtvf_step(1e-4, n)
        sb.gtvf_step(1e-4, n)
        _equal('far frame step %d' % sa.steps_done, a, b, sa, sb)
    assert int(sa.counters[6].item()) > 0


def _cube_over_wall(gap, dx=0.05, n=4, origin_shift=0.):
    """n^3 cube whose lowest layer sits `gap` above a one-layer wall."""
    from rigid_body_2d_3d_pysph_b200.rigid_body_3d import RigidBody3DScheme
    i, j, k = np.meshgrid(np.arange(n), np.arange(n), np.arange(n),
                          indexing='ij')
    x = i.ravel() * dx + origin_shift
    y = j.ravel() * dx + gap
    z = k.ravel() * dx
    body = get_particle_array(name='body', x=x, y=y, z=z, h=dx,
                              m=2000. * dx**3, rho=2000.,
                              constants={'spacing0': dx})
    body.add_property('body_id', type='int', data=0)
    body.add_property('dem_id', type='int', data=0)
    body.add_constant('total_no_bodies', [2])
    g = (np.arange(-6, n + 6)) * dx
    wx, wz = np.meshgrid(g, g, indexing='ij')
    wall = get_particle_array(name='wall', x=wx.ravel() + origin_shift,
                              y=np.zeros(wx.size), z=wz.ravel(), h=dx,
                              m=2000. * dx**3, rho=2000.)
    wall.add_property('dem_id', type='int', data=1)
    s = RigidBody3DScheme(['body'], ['wall'], dim=3, gy=-9.81)
    s.kf = 1e3
    s.setup_properties([body, wall])
    for pa in (body, wall):
        pa.add_property('contact_force_is_boundary')
    body.contact_force_is_boundary[:] = body.is_boundary[:]
    wall.contact_force_is_boundary[:] = 1.
    return [body, wall]


@pytest.mark.parametrize('shift', [0., 777.])
def test_filter_keeps_every_contact_near_spacing0(shift):
    """Gaps that straddle the contact threshold dist == spacing0 by 1e-3 ...
    1e-12 (relative), on either side: the first pass must keep every slot the
    FP64 evaluation finds in contact."""
    from rigid_body_2d_3d_pysph_b200.device import DeviceScene
    dx = 0.05
    rels = [0.] + [s * 10.**(-e) for e in range(3, 13) for s in (-1., 1.)]
    nact = []
    for rel in rels:
        res = []
        for exact in (False, True):
            arrays = _cube_over_wall(dx * (1. + rel), origin_shift=shift)
            sc = DeviceScene(arrays, ['body'], ['wall'], dim=3, gy=-9.81,
                             exact=exact)
            sc.gtvf_step(1e-5, 3)
            sc.check_status()
            res.append((arrays, sc))
        (a, sa), (b, sb) = res
        _equal('gap %+.0e' % rel, a, b, sa, sb)
        nact.append(sb.read_counters()['active_slots'])
    assert max(nact) > 0 and min(nact) == 0, nact   # both sides were sampled
