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
the exchange stays correct even when bodies drift across the initial cuts
(ownership does not migrate; only the efficiency of the slabs would degrade).
Static wall particles are replicated per slab when the scene is built.

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


class SlabScene(object):
    """A DeviceScene that is one x-slab of a larger scene."""

    def __init__(self, scene, rank, world, group=None, halo_name='halo'):
        self.sc = scene
        self.rank, self.world, self.group = rank, world, group
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
        self.bytes_sent = 0
        self.bytes_recv = 0
        self._set_source_count(0)

    def _set_source_count(self, n_halo):
        self.n_halo = n_halo
        self.sc._src.n = self.n_src_static + n_halo

    def lists_need_rebuild(self):
        """Global decision (all ranks rebuild together): has any body on any
        rank moved more than half the skin since the lists were built?"""
        flag = self.sc.rebuild
        if self.world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
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
            ivs = [torch.empty_like(iv) for _ in range(self.world)]
            dist.all_gather(ivs, iv, group=self.group)
            ivs = torch.stack(ivs)
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
            got, ns, nr, table = exchange_rows(cols, rel, self.rank,
                                               self.world, self.group)
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

    def _pack(self, gsel):
        P = self.sc.P
        return torch.stack([P['x'][gsel], P['y'][gsel], P['z'][gsel],
                            P['u'][gsel], P['v'][gsel], P['w'][gsel],
                            P['h'][gsel], P['dem_id'][gsel].double()], 1)

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
                self._send_buf.data_ptr(), sc.stream), 'rbx_halo_pack')
        send = torch.split(self._send_buf, self._send_counts, 0)
        recv = torch.split(self._recv_buf, self._recv_counts, 0)
        ops = []
        for q in range(self.world):
            if q == self.rank:
                continue
            if self._send_counts[q] > 0:
                ops.append(dist.P2POp(dist.isend, send[q], q, self.group))
            if self._recv_counts[q] > 0:
                ops.append(dist.P2POp(dist.irecv, recv[q], q, self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        if nr:
            _lib.check(sc.lib.rbx_halo_unpack(
                ctypes.byref(sc.scene), self.halo_off, nr,
                self._recv_buf.data_ptr(), sc.skin, sc.stream),
                'rbx_halo_unpack')
        self.bytes_sent += ns * HALO_COLS * 8
        self.bytes_recv += nr * HALO_COLS * 8

    def gtvf_step(self, dt, nsteps=1):
        """GTVFIntegrator.one_timestep with the halo exchange between the
        re-pose (stage 2) and the force evaluation."""
        sc = self.sc
        sc.push_touched()
        for k in range(nsteps):
            sc.gtvf_kick(dt)
            sc.gtvf_drift(dt)
            sc.pose(_lib.POSE_POS | _lib.POSE_VEL | _lib.POSE_VEL_PREV |
                    _lib.POSE_NORMALS)
            self.exchange_halo(full=self.lists_need_rebuild())
            sc.cells_build()
            sc.contact(dt)
            sc.reduce_bodies()
            sc.gtvf_kick(dt)
            # the stage-3 particle velocities of a step are overwritten by
            # stage 1 of the next one before anything reads them (the halo
            # payload is packed after stage 1): only the last step of a
            # batch writes them, as in DeviceScene.gtvf_step
            if k == nsteps - 1:
                sc.pose(_lib.POSE_VEL)
        sc.steps_done += nsteps
        sc.mark_device_newer()
