"""Multi-GPU: x-slab decomposition of a large multi-body scene (SURVEY.md
section 8e), one process per GPU, torch.distributed (NCCL over NVLink on the
GPU box, gloo in the CPU tests) for the plumbing.

Ownership is by *body*: every rank owns whole bodies (those whose centre of
mass started in its slab).  All destination work of a body -- contact forces
on its particles, the force/torque reduction, the integrator -- is therefore
local, and the "straddler" all-reduce of a particle-wise partition never
arises.  What crosses ranks each step is the *halo*: the source particles
(contact_force_is_boundary == 1) of a rank's bodies that lie inside another
rank's region of interest [x_min - reach, x_max + reach] of its own particles:
position, velocity, h and dem_id, 64 B per particle, sent after the drift /
re-pose and before the cell-list build.  Received particles land in the
'halo' array of the local scene, which is an ordinary static-boundary source
array to the kernels (scenes.synthetic_pile(halo_cap=...)).

The region of interest is recomputed from the live positions every step, so
the exchange stays correct even when bodies drift across the initial cuts.
Ownership follows them: ``SlabScene.migrate`` hands every body to the rank
whose interval [cut_k, cut_k+1) holds its centre of mass -- per-body state,
body-frame vectors and the contact history of its particles travel with it --
and can move the cuts so that every rank holds the same number of bodies
again.  It is a rare, host-mediated event (the scene is
re-created from the merged arrays); ``gtvf_step(..., migrate_every=M)`` checks
every M steps whether any body has left its slab.
Static wall particles are replicated per slab when the scene is built.

Backends: NCCL on a multi-GPU box (device tensors go straight into the
collectives); with gloo -- the CPU tests, and two ranks SHARING one GPU, which
NCCL refuses -- the payloads are staged through host memory.

Neighbour lists are reused across steps (device.py, skin), so the halo must
keep its identity between list rebuilds: a *full* exchange (interest intervals
padded by the skin, fresh selection, counts) happens only on the steps where
the lists are rebuilt -- a global decision, one all_reduce(MAX) of the
device-side rebuild flag per step -- and the steps in between only refresh the
positions and velocities of the same particles in the same halo slots
(point-to-point payload between slab neighbours, no selection, no counts).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _lib

HALO_COLS = 8     # x y z u v w h dem_id


def interest_interval(x_own, reach):
    """[x_min - reach, x_max + reach] of the rank's own particles."""
    if x_own.numel() == 0:
        return torch.tensor([1e300, -1e300], dtype=torch.float64,
                            device=x_own.device)
    return torch.stack([x_own.min() - reach, x_own.max() + reach])


def select_halo(cols, intervals, rank):
    """Rows of `cols` ([n, HALO_COLS], column 0 = x) wanted by each rank.

    intervals: [world, 2] tensor.  Returns a list of row-index tensors, empty
    for `rank` itself.  Order inside a message = ascending local index, so
    the exchange is deterministic."""
    out = []
    x = cols[:, 0]
    for q in range(intervals.shape[0]):
        if q == rank:
            out.append(torch.zeros(0, dtype=torch.int64, device=cols.device))
            continue
        m = (x >= intervals[q, 0]) & (x <= intervals[q, 1])
        out.append(torch.nonzero(m).flatten())
    return out


def select_sources_by_body(x, own_src, src_start, src_count, xcm_x, rmax,
                           lo, hi):
    """Positions k in `own_src` with lo <= x[own_src[k]] <= hi, ascending --
    the rows select_halo would pick -- found by testing only the sources of
    the bodies whose bounding sphere [xcm - rmax, xcm + rmax] reaches the
    interval.  own_src is ascending and grouped by body: the sources of body
    b are own_src[src_start[b] : src_start[b] + src_count[b]]."""
    dev = own_src.device
    bsel = torch.nonzero((xcm_x + rmax >= lo) & (xcm_x - rmax <= hi)).flatten()
    cnt = src_count[bsel]
    first = src_start[bsel]
    rep = torch.repeat_interleave(torch.arange(bsel.numel(), device=dev), cnt)
    off = torch.cumsum(cnt, 0) - cnt
    pos = first[rep] + (torch.arange(rep.numel(), device=dev) - off[rep])
    xs = x[own_src[pos]]
    return pos[torch.nonzero((xs >= lo) & (xs <= hi)).flatten()]


def exchange_rows(cols, rows, rank, world, group=None):
    """Send cols[rows[q]] to rank q, receive what the others send here.
    Returns the received rows concatenated in rank order ([m, HALO_COLS])."""
    dev = cols.device
    counts = torch.tensor([r.numel() for r in rows], dtype=torch.int64,
                          device=dev)
    table = [torch.zeros(world, dtype=torch.int64, device=dev)
             for _ in range(world)]
    dist.all_gather(table, counts, group=group)
    table = torch.stack(table).cpu()          # [src, dst] (host sync point)
    recv_counts = [int(table[q, rank]) for q in range(world)]
    recv = [torch.empty(recv_counts[q], cols.shape[1], dtype=cols.dtype,
                        device=dev) for q in range(world)]
    send = [cols.index_select(0, rows[q]).contiguous() for q in range(world)]
    ops = []
    for q in range(world):
        if q == rank:
            continue
        if int(table[rank, q]) > 0:
            ops.append(dist.P2POp(dist.isend, send[q], q, group))
        if recv_counts[q] > 0:
            ops.append(dist.P2POp(dist.irecv, recv[q], q, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    got = torch.cat(recv, 0) if recv else cols[:0]
    return got, int(table[rank].sum()), int(sum(recv_counts)), table


def _staged(group=None):
    """True if device tensors have to go through host memory (gloo)."""
    return dist.get_backend(group) != 'nccl'


# per-body constants of a rigid array (rigid_body_3d.py:781-831) -> stride
BODY_CONSTS = {'total_mass': 1, 'izz': 1, 'xcm': 3, 'xcm0': 3, 'vcm': 3,
               'vcm0': 3, 'ang_mom': 3, 'ang_mom0': 3, 'omega': 3,
               'omega0': 3, 'force': 3, 'torque': 3, 'R': 9, 'R0': 9,
               'inertia_tensor_body_frame': 9,
               'inertia_tensor_inverse_body_frame': 9,
               'inertia_tensor_global_frame': 9,
               'inertia_tensor_inverse_global_frame': 9}


def body_strides(pa):
    nb = int(pa.constants['nb'][0])
    tnb = int(pa.constants['total_no_bodies'][0])
    st = dict((k, v) for k, v in BODY_CONSTS.items() if k in pa.constants)
    for k in ('eta', 'coeff_of_rest'):
        if k in pa.constants and pa.constants[k].size == nb * tnb:
            st[k] = tnb
    return st


def take_bodies(pa, bodies, hist=None):
    """The bodies (local indices, ascending) of rigid array ``pa`` as a
    picklable payload: per-particle properties, per-body constants, and --
    hist = (key [ks, n], dlt [3, ks, n], fn [3, ks, n]) -- the contact
    history of their particles."""
    bodies = np.asarray(bodies, dtype=np.int64)
    bid = pa.properties['body_id']
    sel = np.nonzero(np.isin(bid, bodies))[0]
    out = {'props': {}, 'consts': {}, 'nb': int(bodies.size)}
    for name, arr in pa.properties.items():
        s = pa.stride[name]
        out['props'][name] = arr.reshape(-1, s)[sel].ravel().copy()
    for name, s in body_strides(pa).items():
        out['consts'][name] = \
            pa.constants[name].reshape(-1, s)[bodies].ravel().copy()
    # bodies are identified by their dem_id across ranks
    first = np.searchsorted(bid, bodies)
    out['dem'] = pa.properties['dem_id'][first].astype(np.int64)
    out['count'] = np.bincount(np.searchsorted(bodies, bid[sel]),
                               minlength=bodies.size).astype(np.int64)
    if hist is not None:
        key, dlt, fn = hist
        out['hist'] = (key[:, sel].copy(), dlt[:, :, sel].copy(),
                       fn[:, :, sel].copy())
    return out


def merge_bodies(template, parts):
    """A rigid ParticleArray like ``template`` holding the bodies of the
    payloads ``parts`` in ascending dem_id order, and their history."""
    from .compat.particle_array import ParticleArray
    parts = [p for p in parts if p['nb'] > 0]
    dem = np.concatenate([p['dem'] for p in parts]) if parts else \
        np.zeros(0, np.int64)
    cnt = np.concatenate([p['count'] for p in parts]) if parts else \
        np.zeros(0, np.int64)
    order = np.argsort(dem, kind='stable')
    nb = int(dem.size)
    # particle permutation: bodies in `order`, particles of a body as they were
    start = np.cumsum(cnt) - cnt
    pidx = np.concatenate([np.arange(start[b], start[b] + cnt[b])
                           for b in order]) if nb else np.zeros(0, np.int64)
    new = ParticleArray(name=template.name)
    new.__dict__['_n'] = int(pidx.size)
    for name in template.properties:
        s = template.stride[name]
        cat = np.concatenate([p['props'][name] for p in parts]) if parts \
            else np.zeros(0, template.properties[name].dtype)
        new.add_property(name, type=template.property_types[name],
                         data=cat.reshape(-1, s)[pidx].ravel(), stride=s)
    new.properties['body_id'][:] = np.repeat(
        np.arange(nb, dtype=np.int32), cnt[order])
    st = body_strides(template)
    for name, val in template.constants.items():
        if name in st:
            s = st[name]
            cat = np.concatenate([p['consts'][name] for p in parts]) \
                if parts else np.zeros(0)
            new.add_constant(name, cat.reshape(-1, s)[order].ravel())
        elif name == 'nb':
            new.add_constant('nb', nb)
        elif name in ('min_dem_id', 'max_dem_id') and nb:
            new.add_constant(name, int(dem.min() if name == 'min_dem_id'
                                       else dem.max()))
        else:
            new.add_constant(name, val)
    new.set_output_arrays(list(template.output_property_arrays))
    hist = None
    if parts and all('hist' in p for p in parts):
        key = np.concatenate([p['hist'][0] for p in parts], 1)[:, pidx]
        dlt = np.concatenate([p['hist'][1] for p in parts], 2)[:, :, pidx]
        fn = np.concatenate([p['hist'][2] for p in parts], 2)[:, :, pidx]
        hist = (key, dlt, fn)
    return new, hist


def balanced_cuts(xcm_all, world):
    """Cut planes that give every rank the same number of bodies (+-1):
    midpoints between the neighbours in the sorted centre-of-mass
    coordinates; -inf / +inf at the ends."""
    xs = np.sort(np.asarray(xcm_all, dtype=np.float64))
    cuts = [-np.inf]
    for k in range(1, world):
        i = (k * xs.size) // world
        cuts.append(0.5 * (xs[i - 1] + xs[i]) if 0 < i < xs.size else
                    (xs[0] if xs.size else 0.))
    cuts.append(np.inf)
    return np.array(cuts)


class SlabScene(object):
    """A DeviceScene that is one x-slab of a larger scene."""

    def __init__(self, scene, rank, world, group=None, halo_name='halo',
                 cuts=None):
        self.rank, self.world, self.group = rank, world, group
        self.halo_name = halo_name
        self.staged = world > 1 and _staged(group)
        self.use_graphs = True
        # ownership intervals [cuts[k], cuts[k+1]) along x; None until the
        # first migrate() (ownership = where the scene builder put the body)
        self.cuts = None if cuts is None else np.asarray(cuts, np.float64)
        self.bytes_sent = 0
        self.bytes_recv = 0
        self.migrations = 0
        self.bodies_moved = 0
        self._flag_host = self._flag_event = None
        self._bind(scene)

    def _bind(self, scene):
        self.sc = scene
        self._graphs = {}
        halo_name = self.halo_name
        pas = dict((a.name, a) for a in scene.arrays)
        self.halo_off = scene.p_off[halo_name]
        self.halo_cap = pas[halo_name].get_number_of_particles()
        src = scene.T['src_index'].long()
        self.n_src_static = int((src < self.halo_off).sum().item())
        # the halo array is the last one and fully flagged
        assert int(src.numel()) == self.n_src_static + self.halo_cap, \
            'halo array must be last and fully source-flagged'
        self.own_src = src[src < scene.n_rigid]
        # own sources per body: particles are grouped by body and own_src is
        # ascending, so a body's sources are one contiguous range of own_src
        nb = scene.n_bodies
        cnt = torch.bincount(scene.P['body'][self.own_src].long(),
                             minlength=max(nb, 1))
        self.src_count = cnt
        self.src_start = torch.cumsum(cnt, 0) - cnt
        self.n_halo = 0
        self._send_idx = None
        self._send_counts = self._recv_counts = None
        self._send_buf = self._recv_buf = None
        self._set_source_count(0)

    # -- collectives, staged through host memory under gloo -------------------
    def _all_reduce_max(self, t):
        if self.staged:
            h = t.cpu()
            dist.all_reduce(h, op=dist.ReduceOp.MAX, group=self.group)
            t.copy_(h)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)

    def _all_gather(self, t):
        src = t.cpu() if self.staged else t
        out = [torch.empty_like(src) for _ in range(self.world)]
        dist.all_gather(out, src, group=self.group)
        return torch.stack(out)

    def _p2p(self, send, recv, send_counts, recv_counts):
        """send[q] -> rank q, recv[q] <- rank q (lists of [m, HALO_COLS]
        device tensors, empty ones skipped)."""
        if self.staged:
            hs = [t.cpu() for t in send]
            hr = [torch.empty(t.shape, dtype=t.dtype) for t in recv]
        else:
            hs, hr = send, recv
        ops = []
        for q in range(self.world):
            if q == self.rank:
                continue
            if send_counts[q] > 0:
                ops.append(dist.P2POp(dist.isend, hs[q], q, self.group))
            if recv_counts[q] > 0:
                ops.append(dist.P2POp(dist.irecv, hr[q], q, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        if self.staged:
            for q in range(self.world):
                if q != self.rank and recv_counts[q] > 0:
                    recv[q].copy_(hr[q])

    def _set_source_count(self, n_halo):
        self.n_halo = n_halo
        self.sc._src.n = self.n_src_static + n_halo

    def lists_need_rebuild(self):
        """Global decision (all ranks rebuild together): has any body on any
        rank moved more than half the skin since the lists were built?"""
        flag = self.sc.rebuild
        if self.world > 1:
            self._all_reduce_max(flag)
        return bool(flag.item() != 0)          # host sync

    def exchange_halo(self, full=True):
        """full: choose afresh which source particles every rank needs
        (interest intervals padded by the skin) and remember the choice;
        otherwise only refresh positions and velocities of the particles
        chosen at the last full exchange -- same particles, same halo slots,
        so neighbour lists that name them stay valid."""
        sc = self.sc
        P = sc.P
        if not full and self._send_idx is not None:
            self._refresh_halo()
            return
        iv = interest_interval(P['x'][:sc.n_rigid], sc.reach + sc.skin)
        if self.world > 1:
            ivs = self._all_gather(iv)
        else:
            ivs = iv[None, :]
        ivh = ivs.cpu()                      # host sync (2 doubles / rank)
        # my sources lie inside my own interval shrunk by its padding; only
        # ranks whose interval overlaps that can want any of them (the two
        # slab neighbours unless bodies have wandered)
        pad = sc.reach + sc.skin
        lo_own = float(ivh[self.rank, 0]) + pad
        hi_own = float(ivh[self.rank, 1]) - pad
        idx = self.own_src
        rows = []
        dev = idx.device
        empty = torch.zeros(0, dtype=torch.int64, device=dev)
        nb = sc.n_bodies
        xc = sc.B['xcm'][0:3 * nb:3]
        rb = sc.B['rmax'][:nb]
        for q in range(self.world):
            lo_q, hi_q = float(ivh[q, 0]), float(ivh[q, 1])
            if q == self.rank or hi_q < lo_own or lo_q > hi_own:
                rows.append(empty)
                continue
            # bodies that can reach into q's interval, then only their
            # sources are tested (a few per cent of the slab): the same set,
            # in the same ascending order, as testing every own source
            rows.append(select_sources_by_body(
                P['x'], idx, self.src_start, self.src_count, xc, rb,
                lo_q, hi_q))
        gsel = idx[torch.cat(rows)]
        cols = self._pack(gsel)
        # rows of `cols` per destination are consecutive blocks
        offs = np.cumsum([0] + [r.numel() for r in rows])
        rel = [torch.arange(int(offs[q]), int(offs[q + 1]), device=dev)
               for q in range(self.world)]
        if self.world > 1:
            got, ns, nr, table = self._exchange_rows(cols, rel)
        else:
            got, ns, nr, table = cols[:0], 0, 0, None
        if nr > self.halo_cap:
            raise _lib.RbxError('halo capacity %d < %d received particles' %
                                (self.halo_cap, nr))
        self._send_idx = gsel
        self._send_counts = [int(offs[q + 1] - offs[q])
                             for q in range(self.world)]
        self._recv_counts = [int(table[q, self.rank]) if table is not None
                             else 0 for q in range(self.world)]
        self._unpack(got, nr)
        self._set_source_count(nr)
        self.bytes_sent += ns * HALO_COLS * 8
        self.bytes_recv += nr * HALO_COLS * 8

    def _exchange_rows(self, cols, rows):
        """exchange_rows through this scene's backend."""
        dev = cols.device
        counts = torch.tensor([r.numel() for r in rows], dtype=torch.int64,
                              device=dev)
        table = self._all_gather(counts).cpu()     # [src, dst] (host sync)
        recv_counts = [int(table[q, self.rank]) for q in range(self.world)]
        send_counts = [int(table[self.rank, q]) for q in range(self.world)]
        recv = [torch.empty(recv_counts[q], cols.shape[1], dtype=cols.dtype,
                            device=dev) for q in range(self.world)]
        send = [cols.index_select(0, rows[q]).contiguous()
                for q in range(self.world)]
        self._p2p(send, recv, send_counts, recv_counts)
        got = torch.cat(recv, 0) if recv else cols[:0]
        return got, int(table[self.rank].sum()), int(sum(recv_counts)), table

    def _pack(self, gsel):
        """rows {x, y, z, u, v, w, h, dem_id} of the particles gsel; the
        velocities are the stage-1 velocities formed from the body state
        (u, v, w are not kept current inside a step)."""
        import ctypes
        sc = self.sc
        rows = torch.empty(gsel.numel(), HALO_COLS, dtype=torch.float64,
                           device=sc.device)
        if gsel.numel():
            _lib.check(sc.lib.rbx_halo_pack(
                ctypes.byref(sc.scene), gsel.data_ptr(), int(gsel.numel()),
                rows.data_ptr(), 1, sc.stream), 'rbx_halo_pack')
        return rows

    def _unpack(self, got, nr):
        P = self.sc.P
        o = self.halo_off
        for c, n in enumerate(['x', 'y', 'z', 'u', 'v', 'w', 'h']):
            P[n][o:o + nr] = got[:, c]
        P['dem_id'][o:o + nr] = got[:, 7].to(torch.int32)

    def _refresh_halo(self):
        """Same particles, same halo slots: one gather kernel, point-to-point
        payload between slab neighbours, one scatter kernel."""
        import ctypes
        sc = self.sc
        ns, nr = sum(self._send_counts), sum(self._recv_counts)
        dev = sc.device
        if self._send_buf is None or self._send_buf.shape[0] != ns:
            self._send_buf = torch.empty(ns, HALO_COLS, dtype=torch.float64,
                                         device=dev)
        if self._recv_buf is None or self._recv_buf.shape[0] != nr:
            self._recv_buf = torch.empty(nr, HALO_COLS, dtype=torch.float64,
                                         device=dev)
        if ns:
            _lib.check(sc.lib.rbx_halo_pack(
                ctypes.byref(sc.scene), self._send_idx.data_ptr(), ns,
                self._send_buf.data_ptr(), 1, sc.stream), 'rbx_halo_pack')
        send = list(torch.split(self._send_buf, self._send_counts, 0))
        recv = list(torch.split(self._recv_buf, self._recv_counts, 0))
        self._p2p(send, recv, self._send_counts, self._recv_counts)
        if nr:
            _lib.check(sc.lib.rbx_halo_unpack(
                ctypes.byref(sc.scene), self.halo_off, nr,
                self._recv_buf.data_ptr(), 1e300, sc.stream),
                'rbx_halo_unpack')
            # (no receiver-side displacement check, skin = inf: the rebuild
            # decision is global and the OWNER of these particles raises
            # the flag -- its per-body bound is never smaller than the
            # displacement of a particle -- before the all-reduce)
        self.bytes_sent += ns * HALO_COLS * 8
        self.bytes_recv += nr * HALO_COLS * 8

    # -- ownership ------------------------------------------------------------
    def bodies_outside(self):
        """How many of this rank's bodies have their centre of mass outside
        its ownership interval (0 before the first migrate(): no cuts yet)."""
        if self.cuts is None:
            return 0
        nb = self.sc.n_bodies
        x = self.sc.B['xcm'][0:3 * nb:3]
        lo, hi = float(self.cuts[self.rank]), float(self.cuts[self.rank + 1])
        return int(((x < lo) | (x >= hi)).sum().item())

    def migrate(self, rebalance=False, cuts=None):
        """Hand every body to the rank whose interval holds its centre of mass
        (SURVEY 8e step 5): per-body state, body-frame vectors, and the contact
        history of its particles move; the device scene is re-created from
        the merged arrays and the next step does a full halo exchange.

        cuts: world + 1 ascending cut planes (first -inf, last +inf);
        rebalance=True: place them so that every rank holds the same number
        of bodies (balanced_cuts); neither: keep the current ones (the first
        call then balances).  Collective.  Returns the number of bodies this
        rank received."""
        from .device import DeviceScene
        sc = self.sc
        if len(sc.rigid) != 1:
            raise NotImplementedError('migration handles one rigid array')
        pa = sc.rigid[0]
        sc.sync_to_host()
        hist = sc.history()
        nb = int(pa.constants['nb'][0])
        xcm_x = np.asarray(pa.constants['xcm'])[0:3 * nb:3].copy()
        if cuts is not None:
            self.cuts = np.asarray(cuts, np.float64)
        elif rebalance or self.cuts is None:
            allx = [None] * self.world
            dist.all_gather_object(allx, xcm_x, group=self.group)
            self.cuts = balanced_cuts(np.concatenate(allx), self.world)
        owner = np.searchsorted(self.cuts[1:-1], xcm_x, side='right')
        out = []
        for q in range(self.world):
            b = np.nonzero(owner == q)[0]
            out.append(take_bodies(pa, b, hist))
        # everybody sees everybody's parcels and takes its own (the parcels
        # are a few bodies; a rare event)
        mine = out[self.rank]
        parcels = [out[q] if q != self.rank else None
                   for q in range(self.world)]
        gathered = [None] * self.world
        dist.all_gather_object(gathered, parcels, group=self.group)
        incoming = [gathered[src][self.rank] for src in range(self.world)
                    if src != self.rank]
        received = sum(p['nb'] for p in incoming)
        moved = nb - mine['nb']
        tot = torch.tensor([received + moved], dtype=torch.int64)
        if dist.get_backend(self.group) == 'nccl':
            tot = tot.to(self.sc.device)
        dist.all_reduce(tot, group=self.group)
        self.migrations += 1
        self.bodies_moved += moved
        if int(tot.item()) == 0:
            return 0                       # nobody moved: keep the scene
        new_pa, new_hist = merge_bodies(pa, [mine] + incoming)
        arrays = [new_pa] + [a for a in sc.arrays if a is not pa]
        for a in arrays[1:]:
            a.__dict__['_device'] = None
        kw = dict(sc._ctor)
        rigid_names, boundary_names = kw.pop('rigid_names'), \
            kw.pop('boundary_names')
        steps_done = sc.steps_done
        del hist
        self.sc = None
        sc = None
        torch.cuda.empty_cache()
        new_sc = DeviceScene(arrays, rigid_names, boundary_names, **kw)
        new_sc.steps_done = steps_done
        if new_hist is not None and new_pa.get_number_of_particles():
            new_sc.set_history(new_pa.name, *new_hist)
        self._bind(new_sc)
        return received

    def _half(self, p, flags):
        """One half of the fused step (rbx_gtvf_step flag bit 1: up to the
        positions; bit 2: from the cell list on).  Each half is a fixed
        sequence of launches with no host decision inside, so it is captured
        once per history parity into a CUDA graph and replayed: one launch
        instead of 2 / 20, and the list rebuild of the second half becomes
        the body of a conditional node (see rbx_gtvf_step).  The halo
        exchange and the host's rebuild decision stay between the two."""
        sc = self.sc
        second = bool(flags & 4)
        if not self.use_graphs or sc._dense_pending > 0:
            sc._gtvf_step_call(p, flags=flags, evaluated=second)
            return
        key = (flags, sc.parity, p.dt, p.flags)
        g = self._graphs.get(key)
        if g is None:
            # (the kernels have all run eagerly by now: the dense evaluations)
            parity, pending = sc.parity, sc._dense_pending
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream(sc.device)
            side.wait_stream(torch.cuda.current_stream(sc.device))
            with torch.cuda.stream(side):
                with torch.cuda.graph(g, stream=side):
                    sc._gtvf_step_call(p, flags=flags, evaluated=False)
            torch.cuda.current_stream(sc.device).wait_stream(side)
            sc.parity, sc._dense_pending = parity, pending
            self._graphs[key] = g
        g.replay()
        if second:
            sc._evaluated()

    def gtvf_step(self, dt, nsteps=1, migrate_every=0):
        """GTVFIntegrator.one_timestep with the halo exchange between the
        re-pose (stage 2) and the force evaluation.  migrate_every = M > 0:
        every M steps, if a body of any rank has left its slab, ownership
        migrates (cuts are kept; the first time they are count-balanced)."""
        if migrate_every and self.world > 1:
            done = 0
            while done < nsteps:
                n = min(migrate_every - self.sc.steps_done % migrate_every,
                        nsteps - done)
                self.gtvf_step(dt, n)
                done += n
                if self.sc.steps_done % migrate_every == 0:
                    away = torch.tensor([self.bodies_outside() if self.cuts
                                         is not None else 1])
                    if dist.get_backend(self.group) == 'nccl':
                        away = away.to(self.sc.device)
                    dist.all_reduce(away, group=self.group)
                    if int(away.item()) > 0:
                        self.migrate()
            return
        sc = self.sc
        sc.push_touched()
        for k in range(nsteps):
            # the fused step in two halves (two calls into the library), the
            # halo exchange in between.  Particle velocities are formed where
            # they are needed (RBX_PARAM_BODY_VEL; rbx_halo_pack does the
            # same for the payload); u, v, w and the boundary normals are
            # formed when somebody asks for them, as in DeviceScene.gtvf_step
            p = sc.params(dt)
            self._half(p, 2)
            # The rebuild decision is global (all_reduce MAX of the device
            # flag) and the host has to know it: it picks between the halo
            # refresh and a full exchange.  The refresh is right nine times
            # out of ten, so it is issued BEFORE the host reads the flag --
            # the device packs, sends and unpacks while the flag travels --
            # and a full exchange follows only if the flag says so (it
            # replaces what the refresh wrote).
            if self.world > 1:
                self._all_reduce_max(sc.rebuild)
            if self._flag_host is None:
                self._flag_host = torch.zeros(1, dtype=sc.rebuild.dtype
                                              ).pin_memory()
                self._flag_event = torch.cuda.Event()
            self._flag_host.copy_(sc.rebuild, non_blocking=True)
            self._flag_event.record(torch.cuda.current_stream(sc.device))
            if self._send_idx is not None:
                self._refresh_halo()
            self._flag_event.synchronize()
            if self._send_idx is None or int(self._flag_host[0]) != 0:
                self.exchange_halo(full=True)
            self._half(p, 4 | 1)
        if nsteps > 0:
            sc._particles_stale = True     # see DeviceScene.finalize_particles
        sc.steps_done += nsteps
        sc.mark_device_newer()
