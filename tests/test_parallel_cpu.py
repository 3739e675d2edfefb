"""Multi-rank host logic on CPU (gloo, world_size 2): slab scenes number
bodies consistently, and the halo exchange delivers exactly the foreign
source particles inside each rank's region of interest."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rigid_body_2d_3d_pysph_b200.scenes import synthetic_pile

NB = 60        # bodies per slab


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _cols(body):
    src = np.nonzero(body.contact_force_is_boundary == 1.)[0]
    c = np.stack([body.x[src], body.y[src], body.z[src], body.u[src],
                  body.v[src], body.w[src], body.h[src],
                  body.dem_id[src].astype(np.float64)], 1)
    return torch.from_numpy(c)


def _worker(rank, world, port, q):
    from rigid_body_2d_3d_pysph_b200 import parallel
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        arrays, _, info = synthetic_pile(NB, slab=(rank, world), halo_cap=4000)
        body = arrays[0]
        reach = 3.0 * info['dx']
        cols = _cols(body)
        iv = parallel.interest_interval(torch.from_numpy(body.x), reach)
        ivs = [torch.empty_like(iv) for _ in range(world)]
        dist.all_gather(ivs, iv)
        ivs = torch.stack(ivs)
        rows = parallel.select_halo(cols, ivs, rank)
        got, ns, nr, _ = parallel.exchange_rows(cols, rows, rank, world)
        # expectation from the other slab, built locally
        other, _, _ = synthetic_pile(NB, slab=(1 - rank, world))
        oc = _cols(other[0]).numpy()
        lo, hi = float(iv[0]), float(iv[1])
        want = oc[(oc[:, 0] >= lo) & (oc[:, 0] <= hi)]
        ok = got.shape[0] == want.shape[0] and \
            np.array_equal(got.numpy(), want) and nr == want.shape[0]
        q.put((rank, bool(ok), int(nr), int(ns)))
    finally:
        dist.destroy_process_group()


def test_halo_exchange_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q))
             for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] and res[1][1], res
    # what rank 0 sent is what rank 1 received, and vice versa
    assert res[0][3] == res[1][2] and res[1][3] == res[0][2]
    assert res[0][2] > 0 and res[1][2] > 0


def test_slab_scenes_tile_the_full_scene():
    full, _, fi = synthetic_pile(NB, slab=(0, 2), span=2)
    a0, _, i0 = synthetic_pile(NB, slab=(0, 2))
    a1, _, i1 = synthetic_pile(NB, slab=(1, 2))
    assert fi['n_bodies'] == 2 * NB and i0['n_bodies'] == i1['n_bodies'] == NB
    fb = full[0]
    for n in ('x', 'y', 'z', 'dem_id', 'is_boundary', 'dx0'):
        both = np.concatenate([getattr(a0[0], n), getattr(a1[0], n)])
        assert np.array_equal(getattr(fb, n), both), n
    assert np.array_equal(np.unique(fb.dem_id), np.arange(2 * NB))
    assert int(fb.total_no_bodies[0]) == 2 * NB + 1
    # slab walls cover what the slab's bodies can reach
    for a in (a0, a1):
        b, w = a[0], a[1]
        assert w.x.min() <= b.x.min() - 0.15 and w.x.max() >= b.x.max() + 0.15


def test_body_level_selection_equals_full_scan():
    """The rebuild-step halo selection narrows to the bodies whose bounding
    sphere reaches the neighbour's interval before it tests particles; it has
    to pick exactly the rows, in the order, of a scan over every own source
    (parallel.select_halo)."""
    from rigid_body_2d_3d_pysph_b200 import parallel
    arrays, _, info = synthetic_pile(NB, slab=(0, 2))
    body = arrays[0]
    nb = int(body.nb[0])
    x = torch.from_numpy(body.x)
    own_src = torch.from_numpy(
        np.nonzero(body.contact_force_is_boundary == 1.)[0])
    bid = torch.from_numpy(body.body_id.astype(np.int64))
    cnt = torch.bincount(bid[own_src], minlength=nb)
    start = torch.cumsum(cnt, 0) - cnt
    xcm = torch.from_numpy(np.asarray(body.xcm).reshape(-1, 3)[:, 0].copy())
    r0 = np.sqrt(body.dx0**2 + body.dy0**2 + body.dz0**2)
    rmax = torch.from_numpy(np.array([r0[body.body_id == b].max()
                                      for b in range(nb)]))
    xs = x[own_src]
    lo0, hi0 = float(x.min()), float(x.max())
    span = hi0 - lo0
    for lo, hi in [(hi0 - 0.2 * span, hi0 + 1.0), (lo0 - 1.0, lo0 + 0.05),
                   (lo0 + 0.4 * span, lo0 + 0.45 * span),
                   (hi0 + 0.5, hi0 + 1.0), (lo0 - 1.0, hi0 + 1.0)]:
        want = torch.nonzero((xs >= lo) & (xs <= hi)).flatten()
        got = parallel.select_sources_by_body(x, own_src, start, cnt, xcm,
                                              rmax, lo, hi)
        assert torch.equal(got, want), (lo, hi, got.numel(), want.numel())


def test_body_parcels_round_trip_and_balanced_cuts():
    """Host side of SlabScene.migrate: bodies taken out of a rigid array as
    parcels (properties, per-body constants, contact history) and merged back
    in another grouping give the original array, ordered by dem_id; cuts
    placed by balanced_cuts give every rank the same number of bodies."""
    from rigid_body_2d_3d_pysph_b200 import parallel
    (body, wall), _, info = synthetic_pile(12, seed=3)
    n = body.get_number_of_particles()
    nb = int(body.nb[0])
    rng = np.random.default_rng(0)
    body.vcm[:] = rng.normal(size=3 * nb)
    body.R[:] = rng.normal(size=9 * nb)
    ks = 4
    key = rng.integers(-1, 5, size=(ks, n)).astype(np.int32)
    dlt = rng.normal(size=(3, ks, n))
    fn = rng.normal(size=(3, ks, n))
    a = parallel.take_bodies(body, [0, 3, 4, 11], (key, dlt, fn))
    b = parallel.take_bodies(body, [1, 2, 5, 6, 7, 8, 9, 10], (key, dlt, fn))
    assert a['nb'] == 4 and b['nb'] == 8
    merged, hist = parallel.merge_bodies(body, [b, a])
    assert merged.get_number_of_particles() == n
    assert int(merged.nb[0]) == nb
    for name in body.properties:
        assert np.array_equal(merged.properties[name], body.properties[name]), name
    for name in parallel.body_strides(body):
        assert np.array_equal(merged.constants[name], body.constants[name]), name
    assert np.array_equal(hist[0], key) and np.array_equal(hist[1], dlt)
    assert np.array_equal(hist[2], fn)
    # a subset: bodies renumbered 0.., dem_id kept
    sub, _ = parallel.merge_bodies(body, [a])
    assert int(sub.nb[0]) == 4
    assert np.array_equal(np.unique(sub.body_id), np.arange(4))
    assert np.array_equal(np.unique(sub.dem_id), [0, 3, 4, 11])
    assert np.array_equal(sub.xcm, body.xcm.reshape(-1, 3)[[0, 3, 4, 11]].ravel())
    # cuts
    x = rng.uniform(0, 10, 1001)
    cuts = parallel.balanced_cuts(x, 4)
    owner = np.searchsorted(cuts[1:-1], x, side='right')
    cnt = np.bincount(owner, minlength=4)
    assert cnt.max() - cnt.min() <= 1 and cuts[0] == -np.inf
