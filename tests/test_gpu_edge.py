"""Edge cases of the CUDA path against the oracle: non-uniform h (the other
template instantiation), bodies larger than one 128-particle chunk, ragged
bodies, capacity overflows reported through the status word, empty pieces."""
import numpy as np
import pytest

from oracle import rbo
from rigid_body_2d_3d_pysph_b200.compat.particle_array import \
    get_particle_array
from tests.util import DENSE_SLOTS, assert_close

pytestmark = pytest.mark.gpu


def _make(dim, bodies, wall, h_body, h_wall, rho=2000., spacing=0.05):
    """bodies: list of (x, y, z) arrays, one rigid array holding all."""
    from rigid_body_2d_3d_pysph_b200.rigid_body_3d import RigidBody3DScheme
    x = np.concatenate([b[0] for b in bodies])
    y = np.concatenate([b[1] for b in bodies])
    z = np.concatenate([b[2] for b in bodies])
    bid = np.concatenate([np.full(len(b[0]), k) for k, b in enumerate(bodies)])
    nb = len(bodies)
    body = get_particle_array(name='body', x=x, y=y, z=z, h=h_body,
                              m=rho * spacing**dim, rho=rho,
                              constants={'spacing0': spacing})
    body.add_property('body_id', type='int', data=bid)
    body.add_property('dem_id', type='int', data=bid)
    body.add_constant('total_no_bodies', [nb + 1])
    arrays = [body]
    bounds = []
    if wall is not None:
        w = get_particle_array(name='wall', x=wall[0], y=wall[1], z=wall[2],
                               h=h_wall, m=rho * spacing**dim, rho=rho)
        w.add_property('dem_id', type='int', data=nb)
        arrays.append(w)
        bounds = ['wall']
    s = RigidBody3DScheme(['body'], bounds, dim=dim, gy=-9.81)
    s.kf = 1e3
    s.setup_properties(arrays)
    for pa in arrays:
        pa.add_property('contact_force_is_boundary')
        pa.contact_force_is_boundary[:] = pa.is_boundary[:]
    if wall is not None:
        arrays[1].contact_force_is_boundary[:] = 1.
    return arrays, s


def _clone(arrays):
    out = []
    for pa in arrays:
        q = get_particle_array(name=pa.name)
        q.__dict__['_n'] = pa.get_number_of_particles()
        for n, v in pa.properties.items():
            q.add_property(n, type=pa.property_types[n], data=v,
                           stride=pa.stride[n])
        for n, v in pa.constants.items():
            q.add_constant(n, v)
        out.append(q)
    return out


def _run_both(arrays, s, dim, dt, nsteps, **kw):
    from rigid_body_2d_3d_pysph_b200.device import DeviceScene
    oarr = _clone(arrays)
    for pa in oarr[:1]:
        tnb = int(pa.total_no_bodies[0])
        for n in DENSE_SLOTS:
            if n not in pa.properties:
                pa.add_property(n, stride=tnb)
        if 'dem_id_source' not in pa.properties:
            pa.add_property('dem_id_source', type='int', stride=tnb)
    sc = DeviceScene(arrays, ['body'], [a.name for a in arrays[1:]], dim=dim,
                     kr=s.kr, kf=s.kf, fric_coeff=s.fric_coeff, gx=s.gx,
                     gy=s.gy, gz=s.gz, **kw)
    p = rbo.make_params(dim, dt, s.kr, s.kf, s.fric_coeff, s.gx, s.gy, s.gz)
    sc.gtvf_step(dt, nsteps)
    rbo.gtvf_step(oarr, ['body'], p, nsteps=nsteps)
    sc.check_status()
    g, o = arrays[0], oarr[0]
    f = np.sqrt(o.fx**2 + o.fy**2 + o.fz**2).sum()
    for n in ('fx', 'fy', 'fz', 'force'):
        assert_close(getattr(g, n), getattr(o, n), 1e-10, n, max(f, 1e-300))
    for n in ('xcm', 'R', 'vcm', 'omega', 'x', 'y'):
        assert_close(getattr(g, n), getattr(o, n), 1e-9, n,
                     max(np.abs(getattr(o, n)).max(), 1e-2))
    return sc, g, o


def _block2d(nx, ny, dx, x0, y0):
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing='ij')
    return x0 + i.ravel() * dx, y0 + j.ravel() * dx, np.zeros(nx * ny)


def test_nonuniform_h_and_multi_chunk_body():
    """A 20x20 body (400 particles = 4 chunks) and a ragged 3x7 one on a wall
    whose particles carry a different h (non-uniform-h kernel variants; the
    neighbour predicate then really uses max(h_i, h_j))."""
    dx = 0.05
    b1 = _block2d(20, 20, dx, 0.0, 0.0)
    b2 = _block2d(3, 7, dx, 20 * dx + 0.9 * dx, 0.0)
    xw = (np.arange(60) - 10) * dx
    wall = (xw, np.full(60, -0.97 * dx), np.zeros(60))
    arrays, s = _make(2, [b1, b2], wall, h_body=dx, h_wall=1.2 * dx)
    sc, g, o = _run_both(arrays, s, 2, 1e-4, 30)
    assert sc.h_uniform == 0.0 and sc.n_chunks == 5
    assert np.abs(o.fy).max() > 1.0          # the wall is felt


def test_no_boundaries_no_neighbours():
    """Two bodies far apart, no boundary arrays at all: empty source sets for
    every slot, free fall."""
    dx = 0.05
    arrays, s = _make(2, [_block2d(4, 4, dx, 0., 0.),
                          _block2d(4, 4, dx, 5., 0.)], None, dx, dx)
    sc, g, o = _run_both(arrays, s, 2, 1e-4, 20)
    assert sc.read_counters()['active_slots'] == 0
    # GTVF: no initial acceleration, so the first half kick uses F = 0
    assert np.allclose(g.vcm.reshape(-1, 3)[:, 1], -9.81 * 19.5e-4)


def test_neighbour_list_overflow_is_reported():
    from rigid_body_2d_3d_pysph_b200 import _lib
    from rigid_body_2d_3d_pysph_b200.device import DeviceScene
    dx = 0.05
    b1 = _block2d(6, 6, dx, 0.0, 0.0)
    xw = (np.arange(30) - 10) * dx
    wall = (xw, np.full(30, -0.97 * dx), np.zeros(30))
    arrays, s = _make(2, [b1], wall, dx, dx)
    sc = DeviceScene(arrays, ['body'], ['wall'], dim=2, gy=-9.81, list_cap=2)
    sc.gtvf_step(1e-4, 1)
    with pytest.raises(_lib.RbxError, match='neighbour list overflow'):
        sc.check_status()


def test_history_overflow_is_reported():
    """ks = 1 but a corner particle is in contact with the floor and a second
    body at once."""
    from rigid_body_2d_3d_pysph_b200 import _lib
    from rigid_body_2d_3d_pysph_b200.device import DeviceScene
    dx = 0.05
    b1 = _block2d(4, 4, dx, 0.0, 0.0)
    b2 = _block2d(4, 4, dx, 4 * dx - 0.35 * dx, 0.0)
    xw = (np.arange(30) - 10) * dx
    wall = (xw, np.full(30, -0.9 * dx), np.zeros(30))
    arrays, s = _make(2, [b1, b2], wall, dx, dx)
    sc = DeviceScene(arrays, ['body'], ['wall'], dim=2, gy=-9.81, ks=1)
    sc.gtvf_step(1e-4, 2)
    with pytest.raises(_lib.RbxError, match='simultaneous contacts'):
        sc.check_status()


def test_more_than_four_source_bodies_on_one_particle():
    """A small body surrounded by five others and the wall: its particles see
    six source bodies, more than the kAcc = 4 slot accumulators of one pass
    (extra rounds in k_slots)."""
    dx = 0.05
    g = 0.96 * dx
    c = _block2d(2, 2, dx, 0.0, 0.0)
    around = [(-2 * dx - g + dx, 0.0), (2 * dx + g - dx, 0.0),
              (-2 * dx - g + dx, 2 * dx + g - dx), (2 * dx + g - dx, 2 * dx + g - dx),
              (0.0, 2 * dx + g - dx)]
    bodies = [c] + [_block2d(2, 2, dx, ax, ay) for ax, ay in around]
    xw = (np.arange(30) - 15) * dx
    wall = (xw, np.full(30, -0.97 * dx), np.zeros(30))
    arrays, s = _make(2, bodies, wall, dx, dx)
    sc, gg, o = _run_both(arrays, s, 2, 5e-5, 12)
    tnb = int(o.total_no_bodies[0])
    seen = (o.contact_force_normal_wij.reshape(-1, tnb) > 0).sum(1)
    assert seen[:4].max() >= 5, seen[:4]


def test_split_source_bodies_and_long_lists():
    """h = 6 dx: one chunk sees more gated sources (18 wall rows + a second
    body) than a shared-memory tile of k_neighbours holds, so a source body
    is spread over several runs of the neighbour lists (bit 30 of nbr_cnt)
    and k_slots has to add the partial slots up; the lists are also far
    longer than one strip of the transposition."""
    dx = 0.05
    b1 = _block2d(8, 8, dx, 0.0, 0.0)
    b2 = _block2d(5, 5, dx, 8 * dx + 0.95 * dx, 0.0)
    i, j = np.meshgrid(np.arange(-20, 34), np.arange(18), indexing='ij')
    wall = (i.ravel() * dx, -0.97 * dx - j.ravel() * dx, np.zeros(i.size))
    # spacing0 = 8 dx: the kernel-weighted distance to an 18-row wall is
    # several dx, the contact law must still see an overlap
    arrays, s = _make(2, [b1, b2], wall, h_body=6 * dx, h_wall=6 * dx,
                      spacing=8 * dx)
    sc, g, o = _run_both(arrays, s, 2, 1e-5, 12, list_cap=1200)
    cnt = sc.T['nbr_cnt'].cpu().numpy()
    assert ((cnt >> 30) & 1).any(), 'no chunk needed a second tile'
    assert (cnt & 0x3fffffff).max() > 500
    assert np.abs(o.fy).max() > 3 * o.m[0] * 9.81    # in contact
    assert sc.read_counters()['active_slots'] > 0


def test_more_source_bodies_than_hash_slots_in_one_tile():
    """A hollow frame (124 particles = one chunk) around a raft of 2x2 bodies:
    the chunk's box holds more than 128 distinct source bodies, i.e. more
    than k_neighbours' tile hash has slots (bodies share a group and come out
    as interleaved runs).  The chunk must be flagged as split (bit 30 of
    nbr_cnt) so that no partial run is prefiltered on its own, and the
    forces must still be the oracle's; the outer ring of small bodies
    touches the frame."""
    dx = 0.05
    n = 32
    k = np.arange(n)
    fx = np.r_[k, k, np.zeros(n - 2), np.full(n - 2, n - 1)] * dx
    fy = np.r_[np.zeros(n), np.full(n, n - 1), k[1:-1], k[1:-1]] * dx
    frame = (fx, fy, np.zeros(fx.size))
    # 2x2 bodies on a 13 x 13 lattice inside the frame: 0.55 dx from it on
    # the left and at the bottom (kernel-weighted distance 0.66 dx: in
    # contact), 2.2 dx apart from one another
    pitch = 2.2 * dx
    bodies = [frame]
    for a in range(13):
        for b in range(13):
            bodies.append(_block2d(2, 2, dx, 0.55 * dx + a * pitch,
                                   0.55 * dx + b * pitch))
    arrays, s = _make(2, bodies, None, dx, dx)
    # (the SPH boundary identification finds no surface on a 2x2 block)
    arrays[0].contact_force_is_boundary[:] = 1.
    # the frame moves: with no relative velocity the contact law returns the
    # stale fn = 0 (quirk Q3, rigid_body_common.py:945-953)
    arrays[0].vcm[0:3] = [0.05, 0.02, 0.]
    sc, g, o = _run_both(arrays, s, 2, 2e-5, 6)
    cnt = sc.T['nbr_cnt'].cpu().numpy()
    assert ((cnt[:fx.size] >> 30) & 1).all(), 'frame chunk not flagged split'
    assert sc.read_counters()['active_slots'] > 0
    assert np.abs(o.fx[:fx.size]).max() > 0


def test_device_boundary_identification_equals_host():
    """the device setup path gives the host evaluator's
    is_boundary (exactly) and normals (to rounding) on a 3-D body + tank and
    on a 2-D ring-shaped body."""
    from rigid_body_2d_3d_pysph_b200.rigid_body_3d import RigidBody3DScheme
    from rigid_body_2d_3d_pysph_b200.setup_device import identify_boundary
    from tests.util import load_config
    for name, dim in (('benchmark_5_3d', 3), ('stack_of_cylinders', 2)):
        arrays, meta = load_config(name)
        for pa in arrays[:2]:
            host = get_particle_array(name=pa.name, x=pa.x, y=pa.y, z=pa.z,
                                      h=pa.h, m=pa.m, rho=pa.rho)
            dev = get_particle_array(name=pa.name, x=pa.x, y=pa.y, z=pa.z,
                                     h=pa.h, m=pa.m, rho=pa.rho)
            RigidBody3DScheme([pa.name], None, dim=dim)._identify_boundary(host)
            identify_boundary(dev, dim)
            assert np.array_equal(host.is_boundary, dev.is_boundary), \
                (name, pa.name)
            assert host.is_boundary.sum() > 0
            assert np.abs(host.normal - dev.normal).max() < 1e-10


def test_device_body_setup_equals_host():
    """total mass, centre of mass, izz, inertia tensor and
    inverse, body-frame vectors from the device (rbx_setup_bodies) equal the
    host helpers (rigid_body_common.py:21-107) to rounding, on the six cubes
    of benchmark_5_3d, the 33 cylinders of stack_of_cylinders and a 2000-body
    pile."""
    from rigid_body_2d_3d_pysph_b200 import rigid_body_common as rc
    from rigid_body_2d_3d_pysph_b200.scenes import synthetic_pile
    from rigid_body_2d_3d_pysph_b200.setup_device import setup_rigid_bodies
    from tests.util import load_config
    cases = [load_config('benchmark_5_3d')[0][0],
             load_config('stack_of_cylinders')[0][0],
             synthetic_pile(2000)[0][0]]
    names = ['total_mass', 'xcm', 'izz', 'inertia_tensor_body_frame',
             'inertia_tensor_inverse_body_frame',
             'inertia_tensor_global_frame',
             'inertia_tensor_inverse_global_frame']
    for k, pa in enumerate(cases):
        planar = k == 1                  # cylinders: singular 3-D tensor
        host, dev = _clone([pa])[0], _clone([pa])[0]
        for q in (host, dev):
            for n in names:
                q.constants[n][:] = 0.
            q.dx0[:] = 0.
        rc.set_total_mass(host)
        rc.set_center_of_mass(host)
        rc.set_moment_of_inertia_izz(host)
        if not planar:
            rc.set_moment_of_inertia_and_its_inverse(host)
        rc.set_body_frame_position_vectors(host)
        setup_rigid_bodies(dev, tensor=not planar)
        for n in names:
            a, b = host.constants[n], dev.constants[n]
            assert np.abs(a - b).max() <= 1e-11 * max(np.abs(a).max(), 1e-300), \
                (k, n, np.abs(a - b).max())
        # (the centre of mass rounds at 1e-16 of the coordinates)
        scale = max(np.abs(pa.x).max(), np.abs(pa.y).max(),
                    np.abs(pa.z).max(), 1.)
        for n in ('dx0', 'dy0', 'dz0'):
            assert np.abs(host.properties[n] -
                          dev.properties[n]).max() < 1e-14 * scale
        assert host.total_mass.min() > 0 and np.abs(host.xcm).max() > 0


def test_neighbour_list_structures():
    """The data structures between the three contact kernels, on configs 3
    and 4: nbr_pos lists grouped by source body with the first entry of a
    body marked; nbr_order a permutation of every 256-particle window in
    descending list length; nbr_srt the same lists, transposed into that
    order, with the LAST entry of a body marked; and the lists a superset of
    the exact neighbour sets restricted to gated sources."""
    from rigid_body_2d_3d_pysph_b200.device import DeviceScene
    from tests.util import load_config
    for name in ('benchmark_5_3d', 'stack_of_cylinders'):
        arrays, meta = load_config(name)
        sc = DeviceScene(arrays, meta['rigid'], meta['boundaries'],
                         dim=meta['dim'], kr=meta['kr'], kf=meta['kf'],
                         fric_coeff=meta['fric_coeff'], gx=meta['gx'],
                         gy=meta['gy'], gz=meta['gz'],
                         planar=(meta['stepper'] == 'gtvf2d'))
        sc.gtvf_step(meta['dt'], 1)
        sc.check_status()
        n, cap = sc.n_rigid, sc.list_cap
        raw = sc.T['nbr_pos'].view(cap, n).cpu().numpy()
        cnt = sc.T['nbr_cnt'].cpu().numpy()
        srt = sc.T['nbr_srt'].view(cap, n).cpu().numpy()
        order = sc.T['nbr_order'].cpu().numpy()
        cnt_s = sc.T['nbr_cnt_srt'].cpu().numpy()
        dem = sc.P['dem_id'].cpu().numpy()
        x, y, z, h = (sc.P[k].cpu().numpy() for k in 'xyzh')
        src = np.zeros(sc.n_total, bool)
        src[sc.T['src_index'].cpu().numpy()] = True
        lens = cnt & 0x3fffffff
        assert lens.max() > 0, name
        # windows of 256: a permutation, longest lists first
        for w0 in range(0, n, 256):
            o = order[w0:w0 + 256]
            assert np.array_equal(np.sort(o), np.arange(w0, min(w0 + 256, n)))
            assert np.all(np.diff(lens[o]) <= 0), (name, w0)
        assert np.array_equal(cnt_s, cnt[order])
        reach2 = (sc.reach + sc.skin)**2
        for t in range(n):
            p, ln = order[t], lens[order[t]]
            a = raw[:ln, p].astype(np.int64)
            b = srt[:ln, t].astype(np.int64)
            qa, qb = a & 0x7fffffff, b & 0x7fffffff
            assert np.array_equal(qa, qb), (name, p)
            assert len(set(qa.tolist())) == ln            # no duplicates
            assert src[qa].all() and (dem[qa] != dem[p]).all()
            d = dem[qa]
            first = np.r_[True, d[1:] != d[:-1]] if ln else np.zeros(0, bool)
            last = np.r_[d[1:] != d[:-1], True] if ln else np.zeros(0, bool)
            assert np.array_equal(a < 0, first), (name, p)  # bit 31 (int32 sign)
            assert np.array_equal(b < 0, last), (name, p)
            if not (cnt[p] >> 30) & 1:                     # one run per body
                assert len(set(d[first].tolist())) == int(first.sum())
            # superset of the gated sources within the reach (lists were
            # built from the positions of this very step)
            r2 = (x - x[p])**2 + (y - y[p])**2 + (z - z[p])**2
            want = np.nonzero(src & (dem != dem[p]) &
                              (r2 < 0.999 * (sc.radius_scale *
                                             max(h[p], h.max()))**2))[0]
            assert np.isin(want, qa).all(), (name, p)
            assert (r2[qa] < 1.01 * reach2).all(), (name, p)
