"""Scene builders: the synthetic pile of BASELINE.json config 5 and small
helpers shared by the benchmark applications.

Host NumPy only; everything here runs once, before the device scene exists.
"""
import numpy as np

from .compat.particle_array import get_particle_array
from .rigid_body_3d import RigidBody3DScheme


def _template(shape, dx, rho, dim=3):
    """One lattice block, its boundary flags and normals (computed once with
    the scheme's own boundary identification, then tiled)."""
    nx, ny, nz = shape
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz),
                          indexing='ij')
    x = (i.ravel() - (nx - 1) / 2.) * dx
    y = (j.ravel() - (ny - 1) / 2.) * dx
    z = (k.ravel() - (nz - 1) / 2.) * dx
    pa = get_particle_array(name='tmpl', x=x, y=y, z=z, h=dx,
                            m=rho * dx**dim, rho=rho)
    RigidBody3DScheme(['tmpl'], None, dim=dim)._identify_boundary(pa)
    return x, y, z, pa.is_boundary.copy(), pa.normal.copy()


def pile_dims(n_bodies):
    """Bodies per side and layers: 50 x 50 x 40 at 100 000 bodies."""
    layers = max(1, int(round(0.8 * n_bodies ** (1. / 3.) * 40. / 37.13)))
    side = int(np.ceil(np.sqrt(n_bodies / float(layers))))
    while side * side * layers < n_bodies:
        layers += 1
    return side, layers


def synthetic_pile(n_bodies=100000, block=(5, 5, 4), dx=0.05, rho=2000.,
                   seed=0, e=0.6, wall_layers=5, slab=None):
    """BASELINE.json config 5: a 3-D pile of lattice blocks dropped on a wall.

    Bodies of ``block`` lattice points (dx = h) sit on a jittered grid of
    pitch 6 dx (jitter U(-0.5, 0.5) dx per axis), each with a random
    axis-aligned orientation, ``side x side`` bodies per layer, over a
    ``wall_layers``-thick floor with side walls; g_y = -9.81.
    ``slab = (k, n)``: keep only the bodies of x-slab k of n (multi-GPU weak
    scaling builds one pile per rank).  Returns (body, wall, info).
    """
    rng = np.random.default_rng(seed)
    side, layers = pile_dims(n_bodies)
    pitch = 6. * dx
    nper = block[0] * block[1] * block[2]
    perms = [(0, 1, 2), (2, 0, 1), (1, 2, 0)]   # short axis z, x, y
    tmpl = []
    for pm in perms:
        shape = tuple(block[a] for a in pm)
        tmpl.append(_template(shape, dx, rho))
    b = np.arange(n_bodies)
    ix = b % side
    iz = (b // side) % side
    iy = b // (side * side)
    jit = rng.uniform(-0.5, 0.5, size=(n_bodies, 3)) * dx
    orient = rng.integers(0, 3, size=n_bodies)
    cx = ix * pitch + jit[:, 0]
    cz = iz * pitch + jit[:, 2]
    cy = iy * pitch + jit[:, 1] + 3.5 * dx
    x = np.empty(n_bodies * nper)
    y = np.empty_like(x)
    z = np.empty_like(x)
    isb = np.empty(n_bodies * nper, dtype=np.int32)
    nrm = np.empty(3 * n_bodies * nper)
    xv, yv, zv = (a.reshape(n_bodies, nper) for a in (x, y, z))
    bv = isb.reshape(n_bodies, nper)
    nv = nrm.reshape(n_bodies, 3 * nper)
    for o in range(3):
        m = orient == o
        tx, ty, tz, tb, tn = tmpl[o]
        xv[m] = cx[m, None] + tx[None, :]
        yv[m] = cy[m, None] + ty[None, :]
        zv[m] = cz[m, None] + tz[None, :]
        bv[m] = tb[None, :]
        nv[m] = tn[None, :]
    body_id = np.repeat(np.arange(n_bodies, dtype=np.int32), nper)
    body = get_particle_array(name='body', x=x, y=y, z=z, h=dx,
                              m=rho * dx**3, rho=rho, rad_s=dx / 2.,
                              constants={'spacing0': dx})
    body.add_property('body_id', type='int', data=body_id)
    body.add_property('dem_id', type='int', data=body_id)
    body.add_constant('total_no_bodies', [n_bodies + 1])
    # ---- floor + four side walls, innermost layer is the contact surface --
    L = side * pitch
    H = layers * pitch + 6 * dx
    lo = -pitch / 2. - dx
    n_in = int(round((L + 2 * dx) / dx)) + 1
    g = lo + np.arange(-wall_layers + 1, n_in + wall_layers - 1) * dx
    wx, wy, wz, wf = [], [], [], []
    gx, gz = np.meshgrid(g, g, indexing='ij')
    for k in range(wall_layers):
        wx.append(gx.ravel())
        wz.append(gz.ravel())
        wy.append(np.full(gx.size, -k * dx))
        wf.append(np.full(gx.size, 1. if k == 0 else 0.))
    hy = np.arange(1, int(H / dx) + 1) * dx
    gi = lo + np.arange(0, n_in) * dx
    for k in range(wall_layers):
        for sgn in (0, 1):
            c = (lo - k * dx) if sgn == 0 else (lo + (n_in - 1) * dx + k * dx)
            a, bb = np.meshgrid(gi, hy, indexing='ij')
            flag = np.full(a.size, 1. if k == 0 else 0.)
            # walls normal to x
            wx.append(np.full(a.size, c)); wy.append(bb.ravel())
            wz.append(a.ravel()); wf.append(flag)
            # walls normal to z
            wx.append(a.ravel()); wy.append(bb.ravel())
            wz.append(np.full(a.size, c)); wf.append(flag)
    wx, wy, wz, wf = (np.concatenate(a) for a in (wx, wy, wz, wf))
    wall = get_particle_array(name='wall', x=wx, y=wy, z=wz, h=dx,
                              m=rho * dx**3, rho=rho, rad_s=dx / 2.)
    wall.add_property('dem_id', type='int', data=n_bodies)
    wall.add_property('contact_force_is_boundary', data=wf)
    scheme = RigidBody3DScheme(['body'], ['wall'], dim=3, gy=-9.81)
    scheme.kf = 1e3
    scheme.setup_rigid_array(body, is_boundary=isb, normal=nrm)
    body.add_property('contact_force_is_boundary',
                      data=isb.astype(np.float64))
    from math import log, pi
    t1 = log(e)
    eta = -2. * t1 * (1. / (t1**2. + pi**2.))**0.5
    info = {'n_bodies': n_bodies, 'n_body_particles': int(x.size),
            'n_wall_particles': int(wx.size),
            'n_wall_sources': int((wf == 1.).sum()),
            'side': side, 'layers': layers, 'eta_uniform': eta,
            'dx': dx, 'block': list(block), 'seed': seed}
    return body, wall, scheme, info
