"""TEST INFRASTRUCTURE -- generate tests/golden/*.npz from the reference itself.

Runs ONLY in the build container: it imports the reference's own modules from
/root/reference/code under the PySPH stub (oracle/ref_harness) and executes
their equation / stepper methods with the interpreter.  Each fixture holds

  * the full initial scene (every property and constant after the
    reference's ``setup_properties``), so tests can rebuild it anywhere;
  * ``ref/<step>/<array>/<name>``: the reference's state after <step> steps
    (forces, per-body force/torque, trajectories, dense slot arrays);
  * ``pairs/<dst>/<src>``: the neighbour pairs of the LAST force evaluation.

Usage:  python -m oracle.make_golden [scene ...]
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

from oracle.ref_harness import pysph_stub, interp  # noqa: E402

pysph_stub.install()
import rigid_body_2d  # noqa: E402  (reference module)
import rigid_body_3d  # noqa: E402  (reference module)
import rigid_body_common  # noqa: E402  (reference module)
import dem  # noqa: E402  (reference module)
from pysph.base.kernels import QuinticSpline  # noqa: E402
from pysph.base.utils import get_particle_array  # noqa: E402
from pysph.tools.geometry import get_2d_block, get_3d_block  # noqa: E402

from rigid_body_2d_3d_pysph_b200.compat.output import dump  # noqa: E402

SLOT_PROPS = ['contact_force_normal_x', 'contact_force_normal_y',
              'contact_force_normal_z', 'contact_force_dist', 'overlap',
              'ft_x', 'ft_y', 'ft_z', 'fn_x', 'fn_y', 'fn_z',
              'delta_lt_x', 'delta_lt_y', 'delta_lt_z',
              'vx_source', 'vy_source', 'vz_source',
              'x_source', 'y_source', 'z_source',
              'closest_point_dist_to_source', 'contact_force_normal_wij',
              'contact_force_dist_tmp']
STATE = ['x', 'y', 'z', 'u', 'v', 'w', 'fx', 'fy', 'fz', 'force', 'torque',
         'xcm', 'vcm', 'omega', 'ang_mom', 'R',
         'inertia_tensor_inverse_global_frame', 'normal']


def _body(name, x, y, z, dx, h, rho, dim, body_id, dem_id, tnb):
    m = rho * dx**dim
    pa = get_particle_array(name=name, x=x, y=y, z=z, h=h, m=m, rho=rho,
                            rad_s=dx / 2.,
                            constants={'E': 69e9, 'poisson_ratio': 0.3,
                                       'spacing0': dx})
    pa.add_property('body_id', type='int', data=body_id)
    pa.add_property('dem_id', type='int', data=dem_id)
    pa.add_constant('total_no_bodies', [tnb])
    return pa


def _wall(name, x, y, z, dx, h, rho, dim, dem_id):
    m = rho * dx**dim
    pa = get_particle_array(name=name, x=x, y=y, z=z, h=h, m=m, rho=rho,
                            rad_s=dx / 2.,
                            constants={'E': 69e9, 'poisson_ratio': 0.3})
    pa.add_property('dem_id', type='int', data=dem_id)
    return pa


def _finish(scheme, bodies, walls):
    scheme.setup_properties(bodies + walls)
    for pa in bodies + walls:
        pa.add_property('contact_force_is_boundary')
        pa.contact_force_is_boundary[:] = pa.is_boundary[:]


class Case(object):
    def __init__(self, name, scheme, arrays, dt, stepper='gtvf3d',
                 save_steps=(1,), nsteps=1, notes=''):
        self.name = name
        self.scheme = scheme
        self.arrays = arrays
        self.dt = dt
        self.stepper = stepper
        self.save_steps = set(save_steps)
        self.nsteps = nsteps
        self.notes = notes


def case_free2d(stepper):
    dx = 0.1
    x, y = get_2d_block(dx, 1., 1.)
    z = np.zeros_like(x)
    n = len(x)
    body = _body('body', x, y, z, dx, dx, 10.0, 2, np.zeros(n, int),
                 np.zeros(n, int), 1)
    cls = rigid_body_2d.RigidBody2DScheme if stepper == 'gtvf2d' else \
        rigid_body_3d.RigidBody3DScheme
    s = cls(rigid_bodies=['body'], boundaries=None, dim=2)
    _finish(s, [body], [])
    s.set_linear_velocity(body, np.array([0.5, 0.5, 0.]))
    s.set_angular_velocity(body, np.array([0., 0., 1.]))
    return Case('free2d_' + stepper, s, [body], 1e-3, stepper,
                save_steps=(1, 10, 50), nsteps=50,
                notes='benchmark_1 scene (11x11, v=(.5,.5), omega_z=1)')


def case_wall2d(name, u, g, stepper='gtvf3d', nsteps=40):
    dx = 0.05
    xb, yb = get_2d_block(dx, 2 * dx, 2 * dx)
    n = len(xb)
    body = _body('body', xb, yb, np.zeros(n), dx, dx, 2000., 2,
                 np.zeros(n, int), np.zeros(n, int), 2)
    xw = (np.arange(9) - 4) * dx
    yw = np.full(9, min(yb) - 0.98 * dx)
    wall = _wall('wall', xw, yw, np.zeros(9), dx, dx, 2000., 2, 1)
    if stepper == 'gtvf2d':
        s = rigid_body_2d.RigidBody2DScheme(['body'], ['wall'], dim=2,
                                            gy=g)
    else:
        s = rigid_body_3d.RigidBody3DScheme(['body'], ['wall'], dim=2, gy=g)
    s.kf = 1e3   # effective CLI default (divergence D6)
    _finish(s, [body], [wall])
    wall.contact_force_is_boundary[:] = 1.
    s.set_linear_velocity(body, np.array([u[0], u[1], 0.]))
    return Case(name, s, [body, wall], 1e-4, stepper,
                save_steps=(1, 2, 3, 5, 10, 20, nsteps), nsteps=nsteps,
                notes='3x3 body 0.98dx above a 9-particle wall row')


def case_collide2d():
    dx = 0.025
    h = 1.3 * dx
    xb, yb = get_2d_block(dx, 4 * dx, 4 * dx)
    n = len(xb)
    b1 = _body('body1', xb, yb, np.zeros(n), dx, h, 2000., 2,
               np.zeros(n, int), np.zeros(n, int), 2)
    b2 = _body('body2', xb + 4 * dx + 1.2 * dx, yb + 0.4 * dx, np.zeros(n),
               dx, h, 2000., 2, np.zeros(n, int), np.ones(n, int), 2)
    s = rigid_body_3d.RigidBody3DScheme(['body1', 'body2'], None, dim=2)
    s.kf = 1e3
    _finish(s, [b1, b2], [])
    s.set_linear_velocity(b1, np.array([0.5, 0., 0.]))
    s.set_linear_velocity(b2, np.array([-0.5, 0., 0.]))
    s.set_angular_velocity(b2, np.array([0., 0., 3.]))
    return Case('collide2d', s, [b1, b2], 1e-4, 'gtvf3d',
                save_steps=(1, 40, 60, 80, 100, 120), nsteps=120,
                notes='benchmark_2-like: two arrays, dem_id 0/1, h=1.3dx')


def case_cubes3d():
    dx = 0.05
    xb, yb, zb = get_3d_block(dx, 2 * dx, 2 * dx, 2 * dx)
    n = len(xb)
    x = np.concatenate([xb, xb + 0.3 * dx])
    y = np.concatenate([yb, yb + 3 * dx - 0.03 * dx])
    z = np.concatenate([zb, zb - 0.2 * dx])
    bid = np.concatenate([np.zeros(n, int), np.ones(n, int)])
    body = _body('body', x, y, z, dx, dx, 2000., 3, bid, bid.copy(), 3)
    xt, yt, zt = get_3d_block(dx, 6 * dx, dx, 6 * dx)
    yt += min(y) - max(yt) - 0.97 * dx
    tank = _wall('tank', xt, yt, zt, dx, dx, 2000., 3, 2)
    s = rigid_body_3d.RigidBody3DScheme(['body'], ['tank'], dim=3, gy=-9.81)
    s.kf = 1e3
    _finish(s, [body], [tank])
    tank.contact_force_is_boundary[:] = 1.
    cor = np.ones(body.nb[0] * body.total_no_bodies[0]) * 0.6
    body.add_constant('coeff_of_rest', cor)
    rigid_body_common.setup_damping_coefficient(body, [body],
                                                boundaries=[tank])
    s.set_linear_velocity(body, np.array([0.1, 0., 0.05, -0.2, -0.1, 0.]))
    s.set_angular_velocity(body, np.array([0., 0., 0., 0.5, 1.0, -0.7]))
    return Case('cubes3d', s, [body, tank], 1e-4, 'gtvf3d',
                save_steps=(1, 2, 5, 10, 20, 40), nsteps=40,
                notes='two 3^3 cubes in one array on a 7x2x7 floor, e=0.6')


def case_rk2_3d():
    dx = 0.05
    xb, yb, zb = get_3d_block(dx, 2 * dx, 2 * dx, 2 * dx)
    n = len(xb)
    body = _body('body', xb, yb, zb, dx, dx, 2000., 3, np.zeros(n, int),
                 np.zeros(n, int), 2)
    xt, yt, zt = get_3d_block(dx, 4 * dx, dx, 4 * dx)
    yt += min(yb) - max(yt) - 0.97 * dx
    tank = _wall('tank', xt, yt, zt, dx, dx, 2000., 3, 1)
    s = rigid_body_3d.RigidBody3DScheme(['body'], ['tank'], dim=3, gy=-9.81)
    s.kf = 1e3
    _finish(s, [body], [tank])
    tank.contact_force_is_boundary[:] = 1.
    s.set_linear_velocity(body, np.array([0.1, -0.1, 0.05]))
    s.set_angular_velocity(body, np.array([0.4, 0.2, -0.3]))
    return Case('rk2_3d', s, [body, tank], 1e-4, 'rk2',
                save_steps=(1, 2, 5, 10, 20), nsteps=20,
                notes='RK2RigidBody3DStep under EPEC sequencing, nb=1')


def case_rk2_3d_nb2():
    """Two bodies in ONE array under the RK2 stepper: py_initialize saves the
    angular momentum of body 0 only (rigid_body_3d.py:415, quirk Q7), so body
    1 integrates from whatever ang_mom0 held at setup."""
    dx = 0.05
    xb, yb, zb = get_3d_block(dx, 2 * dx, 2 * dx, 2 * dx)
    n = len(xb)
    x = np.concatenate([xb, xb + 0.3 * dx])
    y = np.concatenate([yb, yb + 3 * dx - 0.03 * dx])
    z = np.concatenate([zb, zb - 0.2 * dx])
    bid = np.concatenate([np.zeros(n, int), np.ones(n, int)])
    body = _body('body', x, y, z, dx, dx, 2000., 3, bid, bid.copy(), 3)
    xt, yt, zt = get_3d_block(dx, 6 * dx, dx, 6 * dx)
    yt += min(y) - max(yt) - 0.97 * dx
    tank = _wall('tank', xt, yt, zt, dx, dx, 2000., 3, 2)
    s = rigid_body_3d.RigidBody3DScheme(['body'], ['tank'], dim=3, gy=-9.81)
    s.kf = 1e3
    _finish(s, [body], [tank])
    tank.contact_force_is_boundary[:] = 1.
    s.set_linear_velocity(body, np.array([0.1, 0., 0.05, -0.2, -0.1, 0.]))
    s.set_angular_velocity(body, np.array([0.3, -0.2, 0.4, 0.5, 1.0, -0.7]))
    return Case('rk2_3d_nb2', s, [body, tank], 1e-4, 'rk2',
                save_steps=(1, 2, 5, 10, 20), nsteps=20,
                notes='RK2RigidBody3DStep under EPEC sequencing, nb=2 in one '
                      'array (quirk Q7: ang_mom0 of body 0 only)')


def run_case(case):
    kernel = QuinticSpline(dim=case.scheme.dim)
    eqs = case.scheme.get_equations()
    rigid = list(case.scheme.rigid_bodies)
    bounds = list(case.scheme.boundaries)
    if case.stepper == 'gtvf2d':
        st = dict((n, rigid_body_2d.GTVFRigidBody2DStep()) for n in rigid)
    elif case.stepper == 'gtvf3d':
        st = dict((n, rigid_body_3d.GTVFRigidBody3DStep()) for n in rigid)
    else:
        st = dict((n, rigid_body_3d.RK2RigidBody3DStep()) for n in rigid)
    meta = {'name': case.name, 'rigid': rigid, 'boundaries': bounds,
            'dim': case.scheme.dim, 'dt': case.dt, 'stepper': case.stepper,
            'kr': case.scheme.kr, 'kf': case.scheme.kf,
            'fric_coeff': case.scheme.fric_coeff,
            'gx': case.scheme.gx, 'gy': case.scheme.gy, 'gz': case.scheme.gz,
            'nsteps': case.nsteps, 'save_steps': sorted(case.save_steps),
            'notes': case.notes}
    fname = os.path.join(GOLDEN, case.name + '_scene.npz')
    dump(fname, case.arrays, {'t': 0.0}, detailed_output=True, compress=True)
    out = {}
    nbr_log = {}
    t0 = time.time()
    for step in range(1, case.nsteps + 1):
        t = (step - 1) * case.dt
        if case.stepper == 'rk2':
            interp.epec_step(st, eqs.groups[1], case.arrays, kernel, t,
                             case.dt)
        else:
            interp.gtvf_step(st, eqs, case.arrays, kernel, t, case.dt,
                             nbr_log if step == case.nsteps else None)
        if step in case.save_steps:
            for pa in case.arrays:
                if pa.name not in rigid:
                    continue
                for n in STATE + SLOT_PROPS + ['dem_id_source']:
                    if n in pa.properties:
                        v = pa.properties[n]
                    elif n in pa.constants:
                        v = pa.constants[n]
                    else:
                        continue
                    out['ref/%d/%s/%s' % (step, pa.name, n)] = v.copy()
    for (d, s), nb in nbr_log.items():
        pairs = np.array([(i, j) for i, row in enumerate(nb) for j in row],
                         dtype=np.int32).reshape(-1, 2)
        out['pairs/%s/%s' % (d, s)] = pairs
    meta['seconds'] = time.time() - t0
    out['__meta__'] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(GOLDEN, case.name + '_ref.npz'), **out)
    print('%-16s %4d steps %6.1fs' % (case.name, case.nsteps,
                                       meta['seconds']))


DEM_STATE = ['x', 'y', 'z', 'u', 'v', 'w', 'wx', 'wy', 'wz', 'fx', 'fy', 'fz',
             'torx', 'tory', 'torz', 'tng_idx', 'tng_idx_dem_id', 'tng_x',
             'tng_y', 'tng_z', 'total_tng_contacts']


def run_dem_case(dim=2):
    """DEMScheme (dem.py:628-828): LVCDisplacement + tangential-contact
    bookkeeping + DEMStep under GTVF sequencing.  No script of the reference
    instantiates the scheme; the constants it needs but does not create
    (kn, kt, alpha, mu by source dem_id, max_tng_contacts_limit, moi) are
    added here.  dim = 3: a 4 x 3 x 3 block of spheres on a
    floor of spheres, velocities and spins about all three axes."""
    from pysph.base.kernels import CubicSpline
    rng = np.random.default_rng(3 if dim == 2 else 5)
    rad = 0.01
    name = 'dem%dd' % dim
    if dim == 2:
        nx, ny = 6, 5
        gx, gy = np.meshgrid(np.arange(nx), np.arange(ny), indexing='ij')
        n = gx.size
        x = gx.ravel() * 1.98 * rad + rng.uniform(-0.02, 0.02, n) * rad
        y = gy.ravel() * 1.97 * rad + rad * 0.99 + \
            rng.uniform(-0.02, 0.02, n) * rad
        z = np.zeros(n)
        m = 2500. * np.pi * rad**2
        moi = 0.5 * m * rad**2
    else:
        gx, gy, gz = np.meshgrid(np.arange(4), np.arange(3), np.arange(3),
                                 indexing='ij')
        n = gx.size
        x = gx.ravel() * 1.98 * rad + rng.uniform(-0.02, 0.02, n) * rad
        y = gy.ravel() * 1.97 * rad + rad * 0.99 + \
            rng.uniform(-0.02, 0.02, n) * rad
        z = gz.ravel() * 1.98 * rad + rng.uniform(-0.02, 0.02, n) * rad
        m = 2500. * 4. / 3. * np.pi * rad**3
        moi = 0.4 * m * rad**2
    sand = get_particle_array(name='sand', x=x, y=y, z=z, h=1.2 * rad, m=m,
                              rho=2500., rad_s=rad,
                              u=rng.uniform(-0.05, 0.05, n),
                              v=rng.uniform(-0.05, 0.05, n))
    if dim == 3:
        sand.w[:] = rng.uniform(-0.05, 0.05, n)
    sand.add_property('dem_id', type='int', data=0)
    sand.add_property('moi', data=moi)
    # quirk Q13 (dem.py:239-242): initialize_pair indexes the SOURCE array's
    # dem_id with an index recorded against another array; keep every array
    # at least as long as the longest one so that read stays in bounds
    if dim == 2:
        xw = (np.arange(n + 6) - 14) * 2 * rad
        yw = np.zeros_like(xw) - rad
        zw = np.zeros_like(xw)
    else:
        fi, fk = np.meshgrid(np.arange(-2, 6), np.arange(-2, 5),
                             indexing='ij')
        xw = fi.ravel() * 2. * rad
        zw = fk.ravel() * 2. * rad
        yw = np.zeros_like(xw) - rad
        # ... and the floor particles that can touch the block come first, so
        # that their indices also index the (shorter) sand array
        near = np.argsort((xw - x.mean())**2 + (zw - z.mean())**2,
                          kind='stable')
        xw, zw = xw[near], zw[near]
        assert xw.size >= n
    wall = get_particle_array(name='wall', x=xw, y=yw, z=zw, h=1.2 * rad,
                              m=m, rho=2500., rad_s=rad)
    wall.add_property('dem_id', type='int', data=1)
    for pa in (sand, wall):
        for p in ('wx', 'wy', 'wz'):
            pa.add_property(p)
    sand.wz[:] = rng.uniform(-5, 5, n)
    if dim == 3:
        sand.wx[:] = rng.uniform(-5, 5, n)
        sand.wy[:] = rng.uniform(-5, 5, n)
    limit = 8 if dim == 2 else 12
    sand.add_constant('max_tng_contacts_limit', limit)
    sand.add_constant('kn', [1e5, 2e5])
    sand.add_constant('kt', [2. / 7. * 1e5, 2. / 7. * 2e5])
    sand.add_constant('alpha', [40., 60.])
    sand.add_constant('mu', [0.5, 0.3])
    s = dem.DEMScheme(['sand'], ['wall'], dim=dim, gy=-9.81)
    s.setup_properties([sand, wall])
    eqs = s.get_equations()
    st = {'sand': dem.DEMStep()}
    kernel = CubicSpline(dim=dim)
    arrays = [sand, wall]
    dt = 2e-5
    nsteps, save = (60, (1, 2, 5, 10, 30, 60)) if dim == 2 else \
        (30, (1, 2, 5, 10, 30))
    meta = {'name': name, 'granular': ['sand'], 'boundaries': ['wall'],
            'dim': dim, 'dt': dt, 'gx': 0., 'gy': -9.81, 'gz': 0.,
            'nsteps': nsteps, 'save_steps': list(save), 'limit': limit,
            'radius_scale': kernel.radius_scale}
    dump(os.path.join(GOLDEN, name + '_scene.npz'), arrays, {'t': 0.0},
         detailed_output=True, compress=True)
    out = {}
    t0 = time.time()
    for step in range(1, nsteps + 1):
        interp.gtvf_step(st, eqs, arrays, kernel, (step - 1) * dt, dt)
        if step in save:
            for nme in DEM_STATE:
                out['ref/%d/sand/%s' % (step, nme)] = \
                    sand.properties[nme].copy()
    meta['seconds'] = time.time() - t0
    out['__meta__'] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(GOLDEN, name + '_ref.npz'), **out)
    print('%-16s %4d steps %6.1fs  contacts %d' % (
        name, nsteps, meta['seconds'], sand.total_tng_contacts.sum()))


def run_canelas_case():
    """RigidBodyCanelasRigidRigid / RigidBodyCanelasRigidWall
    (rigid_body_common.py:244-628).  No scheme of the reference wires them;
    the group below is the obvious one: BodyForce + the two Hertz loops, then
    SumUpExternalForces.  The tangential-history properties their signatures
    name (and never touch, the tangential part is commented out) are added."""
    from pysph.sph.equation import Group
    dx = 0.025
    h = 1.3 * dx
    xb, yb = get_2d_block(dx, 3 * dx, 3 * dx)
    n = len(xb)
    x = np.concatenate([xb, xb + 4 * dx - 0.04 * dx])
    y = np.concatenate([yb, yb + 0.12 * dx])
    bid = np.concatenate([np.zeros(n, int), np.ones(n, int)])
    body = _body('body', x, y, np.zeros(2 * n), dx, h, 2000., 2, bid, bid, 3)
    xw = (np.arange(16) - 5.5) * dx + 0.07 * dx
    wall = _wall('wall', xw, np.full(16, yb.min() - 0.97 * dx), np.zeros(16),
                 dx, h, 2000., 2, 2)
    s = rigid_body_3d.RigidBody3DScheme(['body'], ['wall'], dim=2, gy=-9.81)
    _finish(s, [body], [wall])
    s.set_linear_velocity(body, np.array([0.3, -0.2, 0., -0.4, 0.1, 0.]))
    s.set_angular_velocity(body, np.array([0., 0., 2., 0., 0., -3.]))
    limit = 6
    body.add_constant('max_tng_contacts_limit', [limit])
    for nme in ('tng_idx', 'tng_idx_dem_id'):
        body.add_property(nme, type='int', stride=limit)
        body.properties[nme][:] = -1
    for nme in ('tng_fx', 'tng_fy', 'tng_fz'):
        body.add_property(nme, stride=limit)
    body.add_property('total_tng_contacts', type='int')
    arrays = [body, wall]
    st = {'body': rigid_body_3d.GTVFRigidBody3DStep()}
    kernel = QuinticSpline(dim=2)
    dt = 1e-6
    # stage 1 + 2 of one GTVF step give the particles their rigid-body
    # velocities and positions; then one force evaluation
    interp.run_stage(st['body'], 'stage1', body, 0., dt)
    interp.run_stage(st['body'], 'stage2', body, 0., dt)
    Cn = 1.4e-5
    groups = [Group(equations=[
        rigid_body_common.BodyForce(dest='body', sources=None, gx=0.,
                                    gy=-9.81, gz=0.),
        rigid_body_common.RigidBodyCanelasRigidRigid(
            dest='body', sources=['body'], Cn=Cn),
        rigid_body_common.RigidBodyCanelasRigidWall(
            dest='body', sources=['wall'], Cn=Cn)]),
        Group(equations=[rigid_body_common.SumUpExternalForces(
            dest='body', sources=None)])]
    meta = {'name': 'canelas2d', 'rigid': ['body'], 'boundaries': ['wall'],
            'dim': 2, 'dt': dt, 'gx': 0., 'gy': -9.81, 'gz': 0., 'Cn': Cn,
            'kr': s.kr, 'kf': s.kf, 'fric_coeff': s.fric_coeff,
            'stepper': 'gtvf3d'}
    dump(os.path.join(GOLDEN, 'canelas2d_scene.npz'), arrays, {'t': 0.0},
         detailed_output=True, compress=True)
    interp.run_groups(groups, arrays, kernel, 0., dt)
    out = {}
    for nme in ('fx', 'fy', 'fz'):
        out['ref/1/body/' + nme] = body.properties[nme].copy()
    for nme in ('force', 'torque'):
        out['ref/1/body/' + nme] = np.asarray(body.constants[nme]).copy()
    out['__meta__'] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(GOLDEN, 'canelas2d_ref.npz'), **out)
    f = np.abs(body.fy).max()
    print('%-16s |fy|max %.4g  (m g = %.4g)' % ('canelas2d', f,
                                               body.m[0] * 9.81))


def known_answers():
    """Reference-owned numeric pins."""
    dx = 0.1
    n = 4
    pa = get_particle_array(name='b', x=np.arange(n) * dx, h=dx, m=1., rho=1.)
    pa.add_property('body_id', type='int', data=np.zeros(n, int))
    pa.add_property('dem_id', type='int', data=np.zeros(n, int))
    pa.add_constant('total_no_bodies', [2])
    pa.add_constant('total_mass', [4.0])
    pa.add_constant('min_dem_id', 0)
    pa.add_constant('max_dem_id', 0)
    pa.add_constant('eta', np.zeros(2))
    out = {}
    for e in (0.8, 0.6, 1.0, 0.2):
        pa.add_constant('coeff_of_rest', np.ones(2) * e)
        rigid_body_common.setup_damping_coefficient(pa, [pa], boundaries=[])
        out['eta/%g' % e] = pa.eta[0]
    np.savez(os.path.join(GOLDEN, 'known_answers.npz'), **out)
    print('known answers', out)


CASES = {
    'free2d_gtvf2d': lambda: case_free2d('gtvf2d'),
    'free2d_gtvf3d': lambda: case_free2d('gtvf3d'),
    'wall2d': lambda: case_wall2d('wall2d', (0.2, -0.1), -9.81),
    'wall2d_planar': lambda: case_wall2d('wall2d_planar', (0.2, -0.1), -9.81,
                                         'gtvf2d'),
    'wall2d_rest': lambda: case_wall2d('wall2d_rest', (0., 0.), 0., nsteps=3),
    'wall2d_normal': lambda: case_wall2d('wall2d_normal', (0., -0.1), 0.,
                                         nsteps=10),
    'collide2d': case_collide2d,
    'cubes3d': case_cubes3d,
    'rk2_3d': case_rk2_3d,
    'rk2_3d_nb2': case_rk2_3d_nb2,
}

if __name__ == '__main__':
    os.makedirs(GOLDEN, exist_ok=True)
    which = sys.argv[1:] or list(CASES) + ['known', 'dem2d', 'dem3d',
                                           'canelas2d']
    for name in which:
        if name == 'known':
            known_answers()
        elif name == 'dem2d':
            run_dem_case(2)
        elif name == 'dem3d':
            run_dem_case(3)
        elif name == 'canelas2d':
            run_canelas_case()
        else:
            run_case(CASES[name]())
