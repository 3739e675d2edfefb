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
                   seed=0, e=0.6, wall_layers=5, slab=None, halo_cap=0,
                   span=1):
    """BASELINE.json config 5: a 3-D pile of lattice blocks dropped on a wall.

    Bodies of ``block`` lattice points (dx = h) sit on a jittered grid of
    pitch 6 dx (jitter U(-0.5, 0.5) dx per axis), each with a random
    axis-aligned orientation, ``side x side`` bodies per layer, over a
    ``wall_layers``-thick floor with side walls; g_y = -9.81.

    ``slab = (k, n)``: the scene is ``n`` such piles side by side along x
    (n * n_bodies bodies, one walled box) and only slab k is built: its
    bodies (global body index = dem_id, so every rank numbers bodies alike),
    the wall particles within reach of the slab, and -- if ``halo_cap`` > 0 --
    an empty 'halo' array that receives the neighbouring ranks' source
    particles each step (parallel.py).  ``span`` > 1 builds slabs k..k+span-1
    together (span = n: the whole multi-slab scene on one rank, the
    single-GPU twin of an n-rank run).  Returns (arrays, scheme, info).
    """
    k_slab, n_slab = slab if slab is not None else (0, 1)
    side, layers = pile_dims(n_bodies)
    side_x = side * n_slab
    n_total = n_bodies * n_slab
    rng = np.random.default_rng(seed)
    pitch = 6. * dx
    nper = block[0] * block[1] * block[2]
    perms = [(0, 1, 2), (2, 0, 1), (1, 2, 0)]   # short axis z, x, y
    tmpl = []
    for pm in perms:
        shape = tuple(block[a] for a in pm)
        tmpl.append(_template(shape, dx, rho))
    # global body table (cheap: 3 doubles + 1 int per body), then the slab
    per_layer = side_x * side
    n_grid = per_layer * layers
    jit_all = rng.uniform(-0.5, 0.5, size=(n_grid, 3)) * dx
    orient_all = rng.integers(0, 3, size=n_grid)
    g = np.arange(n_grid)
    gix = g % side_x
    giz = (g // side_x) % side
    giy = g // per_layer
    # number bodies slab by slab so that each slab holds n_bodies of them
    slab_of = gix // side
    order = np.lexsort((g, slab_of))
    rank_in_slab = np.empty(n_grid, dtype=np.int64)
    for s in range(n_slab):
        sel = order[slab_of[order] == s]
        rank_in_slab[sel] = np.arange(sel.size)
    keep = rank_in_slab < n_bodies
    gid = slab_of * n_bodies + rank_in_slab          # global body index
    mine = keep & (slab_of >= k_slab) & (slab_of < k_slab + span)
    idx = np.nonzero(mine)[0]
    idx = idx[np.argsort(gid[idx])]
    nb = idx.size
    ix, iy, iz = gix[idx], giy[idx], giz[idx]
    jit, orient = jit_all[idx], orient_all[idx]
    cx = ix * pitch + jit[:, 0]
    cz = iz * pitch + jit[:, 2]
    cy = iy * pitch + jit[:, 1] + 3.5 * dx
    x = np.empty(nb * nper)
    y = np.empty_like(x)
    z = np.empty_like(x)
    isb = np.empty(nb * nper, dtype=np.int32)
    nrm = np.empty(3 * nb * nper)
    xv, yv, zv = (a.reshape(nb, nper) for a in (x, y, z))
    bv = isb.reshape(nb, nper)
    nv = nrm.reshape(nb, 3 * nper)
    for o in range(3):
        m = orient == o
        tx, ty, tz, tb, tn = tmpl[o]
        xv[m] = cx[m, None] + tx[None, :]
        yv[m] = cy[m, None] + ty[None, :]
        zv[m] = cz[m, None] + tz[None, :]
        bv[m] = tb[None, :]
        nv[m] = tn[None, :]
    body_id = np.repeat(np.arange(nb, dtype=np.int32), nper)
    dem_id = np.repeat(gid[idx].astype(np.int32), nper)
    body = get_particle_array(name='body', x=x, y=y, z=z, h=dx,
                              m=rho * dx**3, rho=rho, rad_s=dx / 2.,
                              constants={'spacing0': dx})
    body.add_property('body_id', type='int', data=body_id)
    body.add_property('dem_id', type='int', data=dem_id)
    body.add_constant('total_no_bodies', [n_total + 1])
    # ---- floor + four side walls, innermost layer is the contact surface --
    Lx = side_x * pitch
    Lz = side * pitch
    H = layers * pitch + 6 * dx
    lo = -pitch / 2. - dx
    nx_in = int(round((Lx + 2 * dx) / dx)) + 1
    nz_in = int(round((Lz + 2 * dx) / dx)) + 1
    # x-range of wall particles this slab needs (its bodies +- a margin)
    margin = 4 * pitch
    x_lo = k_slab * side * pitch - pitch - margin if k_slab > 0 else -1e30
    x_hi = (k_slab + span) * side * pitch + margin \
        if k_slab + span < n_slab else 1e30
    gxs = lo + np.arange(-wall_layers + 1, nx_in + wall_layers - 1) * dx
    gxs = gxs[(gxs >= x_lo) & (gxs <= x_hi)]
    gzs = lo + np.arange(-wall_layers + 1, nz_in + wall_layers - 1) * dx
    wx, wy, wz, wf = [], [], [], []
    gx, gz = np.meshgrid(gxs, gzs, indexing='ij')
    for k in range(wall_layers):
        wx.append(gx.ravel())
        wz.append(gz.ravel())
        wy.append(np.full(gx.size, -k * dx))
        wf.append(np.full(gx.size, 1. if k == 0 else 0.))
    hy = np.arange(1, int(H / dx) + 1) * dx
    gix_in = lo + np.arange(0, nx_in) * dx
    gix_in = gix_in[(gix_in >= x_lo) & (gix_in <= x_hi)]
    giz_in = lo + np.arange(0, nz_in) * dx
    for k in range(wall_layers):
        flag1 = 1. if k == 0 else 0.
        for sgn in (0, 1):
            # walls normal to z (span x)
            c = (lo - k * dx) if sgn == 0 else (lo + (nz_in - 1) * dx + k * dx)
            a, bb = np.meshgrid(gix_in, hy, indexing='ij')
            wx.append(a.ravel()); wy.append(bb.ravel())
            wz.append(np.full(a.size, c)); wf.append(np.full(a.size, flag1))
            # walls normal to x (span z): only the end slabs have them
            c = (lo - k * dx) if sgn == 0 else (lo + (nx_in - 1) * dx + k * dx)
            if x_lo <= c <= x_hi:
                a, bb = np.meshgrid(giz_in, hy, indexing='ij')
                wx.append(np.full(a.size, c)); wy.append(bb.ravel())
                wz.append(a.ravel()); wf.append(np.full(a.size, flag1))
    wx, wy, wz, wf = (np.concatenate(a) for a in (wx, wy, wz, wf))
    wall = get_particle_array(name='wall', x=wx, y=wy, z=wz, h=dx,
                              m=rho * dx**3, rho=rho, rad_s=dx / 2.)
    wall.add_property('dem_id', type='int', data=n_total)
    wall.add_property('contact_force_is_boundary', data=wf)
    bounds = ['wall']
    arrays = [body, wall]
    if halo_cap > 0:
        # parked inside the slab (never binned while unused: the source
        # count passed to rbx_cells_build excludes the empty tail)
        halo = get_particle_array(name='halo',
                                  x=np.full(halo_cap, float(np.mean(x))),
                                  y=np.full(halo_cap, float(np.mean(y))),
                                  z=np.full(halo_cap, float(np.mean(z))),
                                  h=dx,
                                  m=rho * dx**3, rho=rho)
        halo.add_property('dem_id', type='int', data=n_total)
        halo.add_property('contact_force_is_boundary',
                          data=np.ones(halo_cap))
        bounds.append('halo')
        arrays.append(halo)
    scheme = RigidBody3DScheme(['body'], bounds, dim=3, gy=-9.81)
    scheme.kf = 1e3
    scheme.setup_rigid_array(body, is_boundary=isb, normal=nrm)
    body.add_property('contact_force_is_boundary',
                      data=isb.astype(np.float64))
    from math import log, pi
    t1 = log(e)
    eta = -2. * t1 * (1. / (t1**2. + pi**2.))**0.5
    info = {'n_bodies': nb, 'n_bodies_total': n_total,
            'n_body_particles': int(x.size),
            'n_wall_particles': int(wx.size),
            'n_wall_sources': int((wf == 1.).sum()),
            'side': side, 'side_x': side_x, 'layers': layers,
            'eta_uniform': eta, 'dx': dx, 'block': list(block), 'seed': seed,
            'slab': [k_slab, n_slab]}
    return arrays, scheme, info
