"""Device-resident scene and the host side of the rigid-body step.

``DeviceScene`` owns the HBM layout of one simulation (torch CUDA tensors,
FP64 / int32 SoA) and drives the C ABI of librbx.so (include/rbx.h).  It is
what the PySPH-shaped ``Integrator`` replacement calls (SURVEY.md section 8b
"Callers"); nothing in here does arithmetic of the hot path on the CPU.

HBM layout
  * global particle order: rigid-body arrays first (in scheme order, each
    array's particles grouped by body), static boundary arrays after;
  * per particle (n_total): x y z u v w h m rho (f64), dem_id (i32);
  * per rigid particle (n_rigid): fx fy fz dx0 dy0 dz0 (f64), body (i32,
    global body index), is_boundary (i32), normal0/normal (f64 x3);
  * per body: the reference's own strided constants (xcm[3b+j], R[9b+j], ...)
    concatenated over arrays, so every ``pa.xcm`` is one contiguous slice;
  * sparse contact history [ks][n_rigid] (key = source dem_id) x 2 buffers;
  * the cell list over the source particles
    (contact_force_is_boundary == 1) with sorted SoA copies.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import (RbxCells, RbxDiag, RbxParams, RbxPoints, RbxScene)

CHUNK = 128

_PARTICLE_F64 = ['x', 'y', 'z', 'u', 'v', 'w', 'h', 'm', 'rho']
_RIGID_F64 = ['fx', 'fy', 'fz', 'dx0', 'dy0', 'dz0']
# per-body constants: name -> stride
_BODY_F64 = {'total_mass': 1, 'izz': 1, 'xcm': 3, 'vcm': 3, 'ang_mom': 3,
             'omega': 3, 'force': 3, 'torque': 3, 'R': 9,
             'inertia_tensor_inverse_body_frame': 9,
             'inertia_tensor_inverse_global_frame': 9,
             'xcm0': 3, 'vcm0': 3, 'ang_mom0': 3, 'R0': 9}
# names the device overwrites during a step (host copies become stale)
_MUTATED_PARTICLE = ['x', 'y', 'z', 'u', 'v', 'w', 'fx', 'fy', 'fz', 'normal']
_MUTATED_BODY = ['xcm', 'vcm', 'ang_mom', 'omega', 'force', 'torque', 'R',
                 'inertia_tensor_inverse_global_frame', 'xcm0', 'vcm0',
                 'ang_mom0', 'R0']
# names whose change invalidates the static tables (source list, h_max, ...)
_STATIC = ['h', 'm', 'rho', 'dem_id', 'body_id', 'contact_force_is_boundary',
           'dx0', 'dy0', 'dz0', 'is_boundary', 'normal0', 'total_mass', 'izz',
           'inertia_tensor_inverse_body_frame', 'eta', 'spacing0']


def _ptr(t):
    return None if t is None else t.data_ptr()


class DeviceScene(object):
    def __init__(self, arrays, rigid_names, boundary_names=(), dim=3,
                 kr=1e5, kf=1e3, fric_coeff=0.5, gx=0., gy=0., gz=0.,
                 planar=False, ks=8, radius_scale=3.0, eta_uniform=None,
                 cap_cells=None, list_cap=96, skin_factor=0.075, device=None,
                 exact=False):
        if not torch.cuda.is_available():
            raise _lib.RbxError('DeviceScene needs a CUDA device; the '
                                'rigid-body path has no CPU fallback')
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else
                                   'cuda:%d' % torch.cuda.current_device())
        # everything needed to build a scene like this one over other arrays
        # (parallel.py re-creates the scene when bodies change owner)
        self._ctor = dict(rigid_names=list(rigid_names),
                          boundary_names=list(boundary_names), dim=dim, kr=kr,
                          kf=kf, fric_coeff=fric_coeff, gx=gx, gy=gy, gz=gz,
                          planar=planar, ks=ks, radius_scale=radius_scale,
                          eta_uniform=eta_uniform, cap_cells=cap_cells,
                          list_cap=list_cap, skin_factor=skin_factor,
                          device=self.device, exact=exact)
        pas = dict((a.name, a) for a in arrays)
        self.rigid = [pas[n] for n in rigid_names]
        self.bounds = [pas[n] for n in boundary_names]
        self.arrays = self.rigid + self.bounds
        self.dim = int(dim)
        self.planar = bool(planar)
        self.ks = int(ks)
        self.list_cap = int(list_cap)
        # neighbour lists are built with reach * (1 + skin_factor) and reused
        # until a body has moved more than half the skin; 0 rebuilds them at
        # every force evaluation (the reference's NNPS.update every step)
        self.skin_factor = float(skin_factor)
        self.radius_scale = float(radius_scale)
        self.kr, self.kf, self.fric_coeff = float(kr), float(kf), \
            float(fric_coeff)
        self.g = (float(gx), float(gy), float(gz))
        self.eta_uniform = eta_uniform
        # exact=True: every (particle, source body) slot is evaluated in FP64
        # in one pass (RBX_PARAM_EXACT); default: FP32 first pass with an
        # error bound, FP64 for what it cannot exclude -- identical results
        self.exact = bool(exact)
        self._cap_cells_req = cap_cells
        self.parity = 0          # history ping-pong
        self.steps_done = 0
        self._graph = None
        self._build_static()
        for pa in self.arrays:
            pa.__dict__['_device'] = self
            pa.__dict__['_host_touched'].clear()
            pa.__dict__['_host_read'].clear()
            pa.__dict__['_device_newer'].clear()

    # ------------------------------------------------------------------
    def _t(self, arr, dtype):
        return torch.as_tensor(np.ascontiguousarray(arr), dtype=dtype) \
            .to(self.device)

    def _build_static(self):
        dev = self.device
        f64, i32 = torch.float64, torch.int32
        # ---- offsets -------------------------------------------------
        self.p_off, self.b_off = {}, {}
        off = 0
        boff = 0
        for pa in self.rigid:
            self.p_off[pa.name] = off
            self.b_off[pa.name] = boff
            off += pa.get_number_of_particles()
            boff += int(pa.constants['nb'][0])
        self.n_rigid = off
        self.n_bodies = boff
        for pa in self.bounds:
            self.p_off[pa.name] = off
            off += pa.get_number_of_particles()
        self.n_total = off
        # ---- particles -----------------------------------------------
        self.P = {}
        for n in _PARTICLE_F64:
            self.P[n] = self._t(np.concatenate(
                [pa.properties[n] for pa in self.arrays]), f64)
        self.P['dem_id'] = self._t(np.concatenate(
            [pa.properties['dem_id'] for pa in self.arrays]), i32)
        if self.n_total and int(self.P['dem_id'].min().item()) < 0:
            raise ValueError('dem_id must be >= 0 (-1 is the empty key of the '
                             'sparse contact history)')
        body = []
        for pa in self.rigid:
            bid = pa.properties['body_id'].astype(np.int64)
            dem = pa.properties['dem_id']
            if bid.size and np.any(np.diff(bid) < 0):
                raise ValueError(
                    "array '%s': particles must be grouped by ascending "
                    'body_id' % pa.name)
            if bid.size:
                first = np.searchsorted(bid, np.arange(bid[-1] + 1))
                first = np.minimum(first, bid.size - 1)
                if np.any(dem != dem[first[bid]]):
                    raise ValueError("array '%s': a body carries more than "
                                     'one dem_id' % pa.name)
            body.append(bid + self.b_off[pa.name])
        body = np.concatenate(body) if body else np.zeros(0, np.int64)
        self.P['body'] = self._t(body, i32)
        for n in _RIGID_F64:
            self.P[n] = self._t(np.concatenate(
                [pa.properties[n] for pa in self.rigid]), f64)
        have_normals = all('normal' in pa.properties and 'normal0' in
                           pa.properties and 'is_boundary' in pa.properties
                           for pa in self.rigid)
        self.have_normals = have_normals
        if have_normals:
            self.P['normal'] = self._t(np.concatenate(
                [pa.properties['normal'] for pa in self.rigid]), f64)
            self.P['normal0'] = self._t(np.concatenate(
                [pa.properties['normal0'] for pa in self.rigid]), f64)
            self.P['is_boundary'] = self._t(np.concatenate(
                [pa.properties['is_boundary'] for pa in self.rigid]), i32)
        # ---- bodies ----------------------------------------------------
        self.B = {}
        for n, s in _BODY_F64.items():
            parts = []
            for pa in self.rigid:
                nb = int(pa.constants['nb'][0])
                if n in pa.constants:
                    parts.append(np.asarray(pa.constants[n], dtype=np.float64))
                else:
                    parts.append(np.zeros(nb * s))
            self.B[n] = self._t(np.concatenate(parts) if parts
                                else np.zeros(0), f64)
        self.B['R_prev'] = self.B['R'].clone()
        sp = []
        for pa in self.rigid:
            nb = int(pa.constants['nb'][0])
            if 'spacing0' not in pa.constants:
                # divergence D5: stack_of_cylinders.py:146 passes
                # `initial_spacing0` but the equations read `spacing0`
                # (rigid_body_common.py:752, 898); under PySPH the script
                # cannot run as shipped.  Use the value it meant.
                if 'initial_spacing0' not in pa.constants:
                    raise KeyError("array '%s' has no constant spacing0" %
                                   pa.name)
                pa.add_constant('spacing0', pa.constants['initial_spacing0'])
            sp.append(np.full(nb, float(pa.constants['spacing0'][0])))
        self.B['spacing0'] = self._t(np.concatenate(sp) if sp
                                     else np.zeros(0), f64)
        # list reuse: where every body was at the last build, and how far a
        # particle of it can be from the centre of mass
        self.B['xcm_ref'] = self.B['xcm'].clone()
        self.B['R_ref'] = self.B['R'].clone()
        r0 = torch.sqrt(self.P['dx0']**2 + self.P['dy0']**2 +
                        self.P['dz0']**2)
        rmax = torch.zeros(max(self.n_bodies, 1), dtype=f64, device=dev)
        if self.n_rigid:
            rmax.scatter_reduce_(0, self.P['body'].long(), r0, 'amax')
        self.B['rmax'] = rmax
        # first body of the array a body belongs to (quirk Q7 of the RK2 step)
        bf = [np.full(int(pa.constants['nb'][0]), self.b_off[pa.name])
              for pa in self.rigid]
        self.B['body_first'] = self._t(np.concatenate(bf) if bf
                                       else np.zeros(0), i32)
        self.rebuild = torch.ones(1, dtype=i32, device=dev)
        # ---- chunks ----------------------------------------------------
        counts = np.bincount(body, minlength=self.n_bodies) \
            if body.size else np.zeros(self.n_bodies, np.int64)
        bstart = np.concatenate([[0], np.cumsum(counts)])
        nch = (counts + CHUNK - 1) // CHUNK
        bc = np.concatenate([[0], np.cumsum(nch)]).astype(np.int64)
        cb = np.repeat(np.arange(self.n_bodies, dtype=np.int64), nch)
        within = np.arange(int(bc[-1]), dtype=np.int64) - bc[cb]
        cs = np.concatenate([bstart[cb] + CHUNK * within, [self.n_rigid]])
        self.n_chunks = int(bc[-1])
        self.T = {'chunk_start': self._t(cs, i32),
                  'chunk_body': self._t(cb, i32),
                  'body_chunk': self._t(bc, i32)}
        # neighbour lists [list_cap][n_rigid] (scratch of the contact op)
        nr_ = max(self.n_rigid, 1)
        self.T['nbr_pos'] = torch.empty(self.list_cap * nr_, dtype=i32,
                                        device=dev)
        self.T['nbr_cnt'] = torch.zeros(nr_, dtype=i32, device=dev)
        # the same lists as the pair kernel reads them (length-ordered work
        # items, entries grouped by source body), made on every rebuild
        self.T['nbr_srt'] = torch.empty(self.list_cap * nr_, dtype=i32,
                                        device=dev)
        self.T['nbr_order'] = torch.arange(nr_, dtype=i32, device=dev)
        self.T['nbr_cnt_srt'] = torch.zeros(nr_, dtype=i32, device=dev)
        # two-precision contact evaluation: FP32 positions relative to the
        # centre of the scene and the compact list of the exact pass
        self.T['pos32'] = torch.zeros(4 * max(self.n_total, 1),
                                      dtype=torch.float32, device=dev)
        self.T['clist'] = torch.zeros(4 * nr_, dtype=i32, device=dev)
        self.origin = [0., 0., 0.]
        if self.n_total:
            for k, n in enumerate('xyz'):
                self.origin[k] = 0.5 * float((self.P[n].min() +
                                              self.P[n].max()).item())
        # ---- damping table ---------------------------------------------
        self.eta_mode = 0
        self.T['eta'] = None
        self.T['eta_row'] = None
        if self.eta_uniform is not None:
            self.eta_mode = 2
            self.T['eta'] = self._t(np.array([self.eta_uniform]), f64)
        else:
            etas, rows, eo = [], [], 0
            for pa in self.rigid:
                nb = int(pa.constants['nb'][0])
                tnb = int(pa.constants['total_no_bodies'][0])
                # no 'eta' constant = no damping (eta_mode 0): nothing is
                # materialised (nb * tnb doubles is 80 GB at 100 000 bodies)
                if 'eta' not in pa.constants:
                    rows.append(np.full(nb, -1, dtype=np.int64))
                    continue
                e = np.asarray(pa.constants['eta'], dtype=np.float64)
                etas.append(e)
                rows.append(eo + np.arange(nb, dtype=np.int64) * tnb)
                eo += e.size
            eta = np.concatenate(etas) if etas else np.zeros(0)
            if eta.size and np.any(eta != 0.):
                self.eta_mode = 1
                self.T['eta'] = self._t(eta, f64)
                self.T['eta_row'] = self._t(np.concatenate(rows),
                                            torch.int64)
        # ---- sources ---------------------------------------------------
        cfb = np.concatenate([
            pa.properties['contact_force_is_boundary']
            if 'contact_force_is_boundary' in pa.properties
            else np.zeros(pa.get_number_of_particles())
            for pa in self.arrays]) if self.arrays else np.zeros(0)
        src = np.nonzero(cfb == 1.)[0]
        self.n_src = int(src.size)
        self.T['src_index'] = self._t(src, i32)
        h = self.P['h']
        self.hmax = float(h.max().item()) if h.numel() else 1.0
        hmin = float(h.min().item()) if h.numel() else 1.0
        self.h_uniform = self.hmax if hmin == self.hmax else 0.0
        self.reach = self.radius_scale * self.hmax
        self.skin = self.skin_factor * self.reach
        # ---- history, status, counters ---------------------------------
        nr = max(self.n_rigid, 1)
        self.H = []
        for _ in range(2):
            self.H.append({
                'key': torch.full((self.ks * nr,), -1, dtype=i32, device=dev),
                'dlt': torch.zeros(3 * self.ks * nr, dtype=f64, device=dev),
                'fn': torch.zeros(3 * self.ks * nr, dtype=f64, device=dev)})
        self.status = torch.zeros(1, dtype=i32, device=dev)
        self.counters = torch.zeros(8, dtype=torch.int64, device=dev)
        # sparse outputs (RbxScene.alist_out): the particles in contact after
        # the last two evaluations, one list per history buffer; tags of the
        # bodies they belong to; {m / rho, spacing0} as floats for k_filter
        self.A = [torch.zeros(nr, dtype=i32, device=dev) for _ in range(2)]
        self.acount = torch.zeros(2, dtype=i32, device=dev)
        self.T['body_tag'] = torch.zeros(max(self.n_bodies, 1), dtype=i32,
                                         device=dev)
        aux = torch.zeros(2 * nr, dtype=torch.float32, device=dev)
        if self.n_rigid:
            aux[0:2 * self.n_rigid:2] = (self.P['m'][:self.n_rigid] /
                                         self.P['rho'][:self.n_rigid]).float()
            aux[1:2 * self.n_rigid:2] = \
                self.B['spacing0'][self.P['body'].long()].float()
        self.T['aux32'] = aux
        # where the static particles (walls, halo) were at the last list build
        self.T['static_ref'] = torch.zeros(
            3 * max(self.n_total - self.n_rigid, 1), dtype=f64, device=dev)
        # evaluations that still have to write every particle (fx = m g and
        # an empty history where nothing is in contact) before the sparse
        # writes can rely on what is in place
        self._dense_pending = 2
        # ---- cell list over the sources ----------------------------------
        self._alloc_cells(max(self.n_src, 1), self.T['src_index'])
        self._refresh_structs()

    def _alloc_cells(self, cap_points, index):
        dev = self.device
        f64, i32 = torch.float64, torch.int32
        if self._cap_cells_req is not None:
            cap_cells = int(self._cap_cells_req)
        else:
            ncell = 1
            if index is not None and index.numel():
                idx = index.long()
                for n in 'xyz':
                    v = self.P[n][idx]
                    ext = float((v.max() - v.min()).item())
                    ncell *= int(ext / self.reach) + 2
            # room for the scene to spread by 2x per axis before coarsening
            cap_cells = int(min(max(8 * ncell, 4096), 1 << 26))
        self.cap_cells = cap_cells
        self.C = {'info': torch.zeros(64, dtype=torch.uint8, device=dev),
                  'cell_start': torch.zeros(cap_cells + 1, dtype=i32,
                                            device=dev),
                  'cell_of': torch.zeros(cap_points, dtype=i32, device=dev),
                  'rank': torch.zeros(cap_points, dtype=i32, device=dev),
                  'gidx': torch.zeros(cap_points, dtype=i32, device=dev),
                  'sdem': torch.zeros(cap_points, dtype=i32, device=dev)}
        for n in ('sx', 'sy', 'sz', 'sh'):
            self.C[n] = torch.zeros(cap_points, dtype=f64, device=dev)
        self.cap_points = cap_points
        nbytes = self.lib.rbx_cells_workspace_bytes(cap_cells, cap_points)
        self.workspace = torch.zeros(nbytes, dtype=torch.uint8, device=dev)

    # ------------------------------------------------------------------
    def _refresh_structs(self):
        P, B, T = self.P, self.B, self.T
        s = RbxScene()
        s.n_total, s.n_rigid, s.n_bodies = self.n_total, self.n_rigid, \
            self.n_bodies
        s.n_chunks, s.dim, s.ks = self.n_chunks, self.dim, self.ks
        s.eta_mode, s.planar = self.eta_mode, int(self.planar)
        for n in ['x', 'y', 'z', 'u', 'v', 'w', 'h', 'm', 'rho', 'dem_id',
                  'fx', 'fy', 'fz', 'dx0', 'dy0', 'dz0', 'body']:
            setattr(s, n, _ptr(P[n]))
        if self.have_normals:
            s.is_boundary = _ptr(P['is_boundary'])
            s.normal0 = _ptr(P['normal0'])
            s.normal = _ptr(P['normal'])
        s.list_cap = self.list_cap
        for n in ['chunk_start', 'chunk_body', 'body_chunk', 'nbr_pos',
                  'nbr_cnt', 'nbr_srt', 'nbr_order', 'nbr_cnt_srt',
                  'eta', 'eta_row', 'pos32', 'clist']:
            setattr(s, n, _ptr(T[n]))
        for k in range(3):
            s.origin[k] = self.origin[k]
            s.gravity[k] = self.g[k]
        s.h_uniform = self.h_uniform
        s.body_tag, s.aux32 = _ptr(T['body_tag']), _ptr(T['aux32'])
        s.static_ref = _ptr(T['static_ref'])
        # for the conditional list rebuild inside captured graphs
        if getattr(self, '_aux_stream', None) is None:
            self._aux_stream = torch.cuda.Stream(self.device)
        s.aux_stream = self._aux_stream.cuda_stream
        for n in ['total_mass', 'izz', 'spacing0', 'xcm', 'vcm', 'ang_mom',
                  'omega', 'force', 'torque', 'R', 'R_prev', 'xcm0', 'vcm0',
                  'ang_mom0', 'R0']:
            setattr(s, n, _ptr(B[n]))
        s.iinv_b = _ptr(B['inertia_tensor_inverse_body_frame'])
        s.iinv_g = _ptr(B['inertia_tensor_inverse_global_frame'])
        s.status = _ptr(self.status)
        s.counters = _ptr(self.counters)
        s.rebuild = _ptr(self.rebuild)
        s.xcm_ref, s.R_ref = _ptr(B['xcm_ref']), _ptr(B['R_ref'])
        s.rmax = _ptr(B['rmax'])
        s.body_first = _ptr(B['body_first'])
        self._scene = [None, None]
        for par in (0, 1):
            c = RbxScene.from_buffer_copy(s)
            hin, hout = self.H[par], self.H[1 - par]
            c.hist_key_in, c.hist_dlt_in, c.hist_fn_in = \
                _ptr(hin['key']), _ptr(hin['dlt']), _ptr(hin['fn'])
            c.hist_key_out, c.hist_dlt_out, c.hist_fn_out = \
                _ptr(hout['key']), _ptr(hout['dlt']), _ptr(hout['fn'])
            # A[k] / acount[k] go with history buffer k
            c.alist_out, c.alist_prev = _ptr(self.A[1 - par]), _ptr(self.A[par])
            c.acount_out = self.acount.data_ptr() + 4 * (1 - par)
            c.acount_prev = self.acount.data_ptr() + 4 * par
            self._scene[par] = c
        self._src = self.points(self.T['src_index'])
        c = RbxCells()
        c.cap_cells, c.cap_points = self.cap_cells, self.cap_points
        for n in ['info', 'cell_start', 'cell_of', 'rank', 'gidx', 'sx', 'sy',
                  'sz', 'sh', 'sdem']:
            setattr(c, n, _ptr(self.C[n]))
        c.cond = _ptr(self.rebuild)     # skip the build while lists are valid
        self._cells = c
        self._graph = None
        self._canelas = None
        self.rebuild.fill_(1)
        self.pos32_refresh()

    def pos32_refresh(self):
        """FP32 positions of every particle from x, y, z, h (after the host
        wrote them; the step kernels keep them current themselves)."""
        if self.n_total:
            _lib.check(self.lib.rbx_pos32_refresh(
                ctypes.byref(self._scene[0]), 0, self.n_total, self.stream),
                'rbx_pos32_refresh')

    def static_update(self, name, state, first=0):
        """New state of (a leading part of) boundary array ``name`` from DEVICE
        tensors: state = {'x': t, 'y': t, ...} over x, y, z, u, v, w, all of
        one length.  Asynchronous on the current stream, no host round trip:
        the neighbour lists are rebuilt only if a particle has moved more
        than half the skin since they were built (rbx_static_update)."""
        n = None
        ptr = {}
        for k in ('x', 'y', 'z', 'u', 'v', 'w'):
            t = state.get(k)
            if t is None:
                ptr[k] = None
                continue
            if t.dtype != torch.float64 or not t.is_cuda or \
                    not t.is_contiguous():
                raise ValueError('static_update wants contiguous float64 '
                                 'CUDA tensors')
            if n is None:
                n = int(t.numel())
            elif n != int(t.numel()):
                raise ValueError('static_update: tensors differ in length')
            ptr[k] = t.data_ptr()
        if not n:
            return
        pas = dict((a.name, a) for a in self.bounds)
        if first < 0 or first + n > pas[name].get_number_of_particles():
            raise ValueError('static_update: range outside array %s' % name)
        _lib.check(self.lib.rbx_static_update(
            ctypes.byref(self.scene), self.p_off[name] + first, n, ptr['x'],
            ptr['y'], ptr['z'], ptr['u'], ptr['v'], ptr['w'], self.skin,
            self.stream), 'rbx_static_update')
        dn = pas[name].__dict__['_device_newer']
        for k in state:
            dn.add(k)

    def points(self, index=None, n=None):
        p = RbxPoints()
        if index is not None:
            p.n = int(index.numel())
            p.index = _ptr(index)
        else:
            p.n = int(self.n_total if n is None else n)
            p.index = None
        p.x, p.y, p.z, p.h = (_ptr(self.P[k]) for k in 'xyzh')
        p.dem_id = _ptr(self.P['dem_id'])
        return p

    def params(self, dt):
        flags = _lib.PARAM_EXACT if self.exact else 0
        if self._dense_pending > 0:
            flags |= _lib.PARAM_DENSE_OUT
        return RbxParams(self.radius_scale, self.kr, self.kf,
                         self.fric_coeff, self.g[0], self.g[1], self.g[2],
                         float(dt), self.reach, self.h_uniform, self.skin,
                         flags, 0)

    def _evaluated(self):
        """One contact evaluation has been issued: the history buffers (and
        the lists that go with them) swap roles."""
        self.parity ^= 1
        self._reduce_dense = False
        if self._dense_pending > 0:
            self._dense_pending -= 1

    def invalidate_outputs(self):
        """The host changed forces, masses or the history behind the back of
        the sparse writes: the next two evaluations (one per history buffer)
        write every particle."""
        self._dense_pending = 2
        self.acount.zero_()

    def force_rebuild(self):
        """Neighbour lists must be rebuilt at the next force evaluation."""
        self.rebuild.fill_(1)

    @property
    def stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    @property
    def scene(self):
        return self._scene[self.parity]

    # ------------------------------------------------------------------
    # host <-> device coherence
    # ------------------------------------------------------------------
    def _slice(self, pa, name):
        """Device tensor slice that mirrors pa.<name>, or None."""
        if name in _PARTICLE_F64 or name == 'dem_id':
            o = self.p_off[pa.name]
            return self.P[name][o:o + pa.get_number_of_particles()]
        if pa in self.rigid:
            o = self.p_off[pa.name]
            n = pa.get_number_of_particles()
            if name in _RIGID_F64 or (name == 'is_boundary' and
                                      self.have_normals):
                return self.P[name][o:o + n]
            if name in ('normal', 'normal0') and self.have_normals:
                return self.P[name][3 * o:3 * (o + n)]
            if name in _BODY_F64:
                s = _BODY_F64[name]
                bo = self.b_off[pa.name]
                nb = int(pa.constants['nb'][0])
                return self.B[name][s * bo:s * (bo + nb)]
        return None

    def pull(self, pa, name):
        if name in ('u', 'v', 'w', 'normal'):
            self.finalize_particles()
        t = self._slice(pa, name)
        if t is None:
            return
        host = pa.properties[name] if name in pa.properties else \
            pa.constants[name]
        host[:] = t.cpu().numpy()

    def push_touched(self):
        """Upload whatever the host code touched since the last step."""
        rebuild = False
        refresh32 = False
        for pa in self.arrays:
            touched = pa.__dict__['_host_touched']
            # reads hand out live arrays: what changed under a read counts
            touched |= pa.modified_since_read()
            if not touched:
                continue
            if touched & set(_STATIC):
                rebuild = True
            for name in list(touched):
                t = self._slice(pa, name)
                if t is None:
                    continue
                host = pa.properties[name] if name in pa.properties else \
                    pa.constants[name]
                t.copy_(torch.as_tensor(host, dtype=t.dtype))
                if name in ('x', 'y', 'z', 'xcm', 'R'):
                    self.rebuild.fill_(1)      # positions changed under us
                if name in ('x', 'y', 'z', 'h'):
                    refresh32 = True
            touched.clear()
            self.invalidate_outputs()
        if refresh32 and not rebuild:
            self.pos32_refresh()
        if rebuild:
            for pa in self.arrays:
                for n in list(pa.__dict__['_device_newer']):
                    self.pull(pa, n)
                pa.__dict__['_device_newer'].clear()
            hist, parity = self.H, self.parity
            shape = (self.n_rigid, self.ks)
            self._build_static()
            # the contact history survives a change of the static tables, but
            # not one of the particle set (remove_particles / add_particles):
            # its rows are per particle
            if shape == (self.n_rigid, self.ks):
                self.H, self.parity = hist, parity
                self._refresh_structs()

    def mark_device_newer(self):
        for pa in self.rigid:
            dn = pa.__dict__['_device_newer']
            for n in _MUTATED_PARTICLE:
                if n in pa.properties:
                    dn.add(n)
            for n in _MUTATED_BODY:
                if n in pa.constants:
                    dn.add(n)

    def sync_to_host(self):
        """Bring every stale host array up to date (before output dumps)."""
        for pa in self.arrays:
            for n in list(pa.__dict__['_device_newer']):
                self.pull(pa, n)
            pa.__dict__['_device_newer'].clear()

    # ------------------------------------------------------------------
    # operations (thin wrappers over the C ABI)
    # ------------------------------------------------------------------
    def cells_build(self, points=None, cells=None):
        pts = points if points is not None else self._src
        c = cells if cells is not None else self._cells
        _lib.check(self.lib.rbx_cells_build(
            ctypes.byref(pts), ctypes.byref(c), self.reach,
            _ptr(self.status), _ptr(self.workspace), self.workspace.numel(),
            self.stream), 'rbx_cells_build')

    def contact(self, dt, diag=None):
        self.finalize_particles()      # the unfused op reads u, v, w
        p = self.params(dt)
        _lib.check(self.lib.rbx_contact_mofidi(
            ctypes.byref(self.scene), ctypes.byref(self._cells),
            ctypes.byref(p), ctypes.byref(diag) if diag is not None else None,
            self.stream), 'rbx_contact_mofidi')
        self._evaluated()

    def contact_canelas(self, dt, Cn=1.4e-5):
        """BodyForce + RigidBodyCanelasRigidRigid / RigidBodyCanelasRigidWall
        (rigid_body_common.py:115-125, 244-628): fx, fy, fz of every rigid
        particle from the Hertz contact of the ``rad_s`` spheres of ALL
        particles of the scene (the equations have no boundary gate).  Every
        array needs the property ``rad_s`` and the constants ``E`` and
        ``poisson_ratio``.  Follow with ``reduce_bodies()``."""
        self.push_touched()
        self.finalize_particles()
        if getattr(self, '_canelas', None) is None:
            f64, i32 = torch.float64, torch.int32
            for pa in self.arrays:
                for n in ('E', 'poisson_ratio'):
                    if n not in pa.constants:
                        raise KeyError("array '%s' has no constant %s" %
                                       (pa.name, n))
                if 'rad_s' not in pa.properties:
                    raise KeyError("array '%s' has no property rad_s" %
                                   pa.name)
            cat = np.concatenate
            t = {'rad_s': self._t(cat([pa.properties['rad_s']
                                       for pa in self.arrays]), f64),
                 'E': self._t(cat([np.full(pa.get_number_of_particles(),
                                           float(pa.constants['E'][0]))
                                   for pa in self.arrays]), f64),
                 'nu': self._t(cat([np.full(
                     pa.get_number_of_particles(),
                     float(pa.constants['poisson_ratio'][0]))
                     for pa in self.arrays]), f64)}
            sb = torch.full((max(self.n_total, 1),), -1, dtype=i32,
                            device=self.device)
            sb[:self.n_rigid] = self.P['body']
            t['src_body'] = sb
            # a cell list of its own over all particles
            saved = (self.C, self.cap_cells, self.cap_points, self.workspace)
            self._alloc_cells(max(self.n_total, 1), torch.arange(
                self.n_total, dtype=i32, device=self.device))
            c = RbxCells()
            c.cap_cells, c.cap_points = self.cap_cells, self.cap_points
            for n in ['info', 'cell_start', 'cell_of', 'rank', 'gidx', 'sx',
                      'sy', 'sz', 'sh', 'sdem']:
                setattr(c, n, _ptr(self.C[n]))
            c.cond = None
            t['C'], t['workspace'], t['cells'] = self.C, self.workspace, c
            (self.C, self.cap_cells, self.cap_points, self.workspace) = saved
            k = _lib.RbxCanelas()
            k.rad_s, k.E, k.nu = _ptr(t['rad_s']), _ptr(t['E']), _ptr(t['nu'])
            k.src_body = _ptr(t['src_body'])
            t['struct'] = k
            self._canelas = t
        t = self._canelas
        t['struct'].Cn = float(Cn)
        _lib.check(self.lib.rbx_cells_build(
            ctypes.byref(self.points()), ctypes.byref(t['cells']), self.reach,
            _ptr(self.status), _ptr(t['workspace']), t['workspace'].numel(),
            self.stream), 'rbx_cells_build')
        p = self.params(dt)
        _lib.check(self.lib.rbx_contact_canelas(
            ctypes.byref(self.scene), ctypes.byref(t['cells']),
            ctypes.byref(p), ctypes.byref(t['struct']), self.stream),
            'rbx_contact_canelas')
        self._reduce_dense = True
        self.invalidate_outputs()       # fx was written densely, untagged
        self.mark_device_newer()

    def reduce_bodies(self):
        sc = self.scene
        if getattr(self, '_reduce_dense', False):
            # forces written by something that does not tag the bodies in
            # contact (the Canelas kernel): sum every body's particles
            sc = RbxScene.from_buffer_copy(sc)
            sc.body_tag = None
        _lib.check(self.lib.rbx_reduce_bodies(ctypes.byref(sc), self.stream),
                   'reduce')

    def gtvf_kick(self, dt):
        _lib.check(self.lib.rbx_gtvf_kick(ctypes.byref(self.scene), float(dt),
                                          self.stream), 'kick')

    def gtvf_drift(self, dt):
        _lib.check(self.lib.rbx_gtvf_drift(ctypes.byref(self.scene),
                                           float(dt), self.skin, self.stream),
                   'drift')

    def pose(self, flags):
        if flags & _lib.POSE_VEL:
            self._particles_stale = False
        _lib.check(self.lib.rbx_pose_particles(ctypes.byref(self.scene),
                                               int(flags), self.stream),
                   'pose')

    def rk2_stage(self, stage, dt, fix_q7=False):
        _lib.check(self.lib.rbx_rk2_stage(ctypes.byref(self.scene),
                                          int(stage), float(dt), int(fix_q7),
                                          self.skin, self.stream), 'rk2')

    def _gtvf_step_call(self, p, flags=0, evaluated=True):
        _lib.check(self.lib.rbx_gtvf_step(
            ctypes.byref(self.scene), ctypes.byref(self._src),
            ctypes.byref(self._cells), ctypes.byref(p), _ptr(self.workspace),
            self.workspace.numel(), int(flags), self.stream), 'rbx_gtvf_step')
        if evaluated:           # (the first half of a split step is not one)
            self._evaluated()

    def gtvf_step(self, dt, nsteps=1, graph=False):
        """nsteps x GTVFIntegrator.one_timestep on the device."""
        self.push_touched()
        # The stage-3 particle velocities of a step are overwritten by stage 1
        # of the next one before anything reads them, and the boundary
        # normals are a function of R alone: inside a batch only the last
        # step writes them (flag bit 0 of rbx_gtvf_step).
        # They are not even written by the last step: whoever looks at u, v,
        # w or the normals first -- a pull to the host, an unfused op -- gets
        # them from finalize_particles().  An application that steps one
        # step at a time and reads per-body results only never pays for them.
        left = nsteps
        while left and self._dense_pending:     # never captured in a graph
            self._gtvf_step_call(self.params(dt), flags=1)
            left -= 1
        p = self.params(dt)
        if graph and left >= 4:
            self._run_graph(p, left)
        else:
            for k in range(left):
                self._gtvf_step_call(p, flags=1)
        if nsteps > 0:
            self._particles_stale = True
        self.steps_done += nsteps
        self.mark_device_newer()

    def finalize_particles(self):
        """Stage-3 particle velocities and the rotated boundary normals of
        the last fused step, if they have not been formed yet."""
        if getattr(self, '_particles_stale', False):
            self._particles_stale = False
            self.pose(_lib.POSE_VEL | _lib.POSE_NORMALS)

    def _run_graph(self, p, nsteps):
        """Two steps (one history ping-pong period) captured as a CUDA graph:
        small scenes are launch-bound (about a dozen launches per step).  A
        graph is bound to the history buffer its first step reads, so both
        are captured at the first use -- capturing executes nothing -- and a
        later call that starts on the other buffer (an odd number of steps
        in between) replays at once instead of capturing in the middle of
        somebody's timed loop."""
        if self._graph is None or self._graph[0] != p.dt:
            start = self.parity
            cur = torch.cuda.current_stream(self.device)
            s = torch.cuda.Stream(self.device)
            s.wait_stream(cur)
            graphs = {}
            with torch.cuda.stream(s):
                self._gtvf_step_call(p, 1)     # warm-up outside capture
                self._gtvf_step_call(p, 1)
                s.synchronize()
                for q in (start, start ^ 1):
                    self.parity = q
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=s):
                        self._gtvf_step_call(p, 1)
                        self._gtvf_step_call(p, 1)
                    assert self.parity == q
                    graphs[q] = g
                self.parity = start
            cur.wait_stream(s)
            self._graph = (p.dt, graphs)
            nsteps -= 2
        g = self._graph[1][self.parity]
        while nsteps >= 2:
            g.replay()
            nsteps -= 2
        for _ in range(nsteps):
            self._gtvf_step_call(p, 1)

    def rk2_step(self, dt, nsteps=1, fix_q7=False):
        """EPEC sequencing with the RK2 stepper."""
        self.push_touched()
        for _ in range(nsteps):
            self.rk2_stage(0, dt, fix_q7)
            self.cells_build()
            self.contact(dt)
            self.reduce_bodies()
            self.rk2_stage(1, dt, fix_q7)
            self.pose(_lib.POSE_POS | _lib.POSE_VEL)
            self.cells_build()
            self.contact(dt)
            self.reduce_bodies()
            self.rk2_stage(2, dt, fix_q7)
            self.pose(_lib.POSE_POS | _lib.POSE_VEL)
        self.steps_done += nsteps
        self.mark_device_newer()

    # ------------------------------------------------------------------
    def check_status(self, raise_on_error=True):
        st = int(self.status.item()) & 0xffffffff
        msgs = []
        if st & _lib.STATUS_SLOT_OVERFLOW:
            msgs.append('a particle touches more than %d bodies' %
                        _lib.RBX_MAX_KEYS)
        if st & _lib.STATUS_HIST_OVERFLOW:
            msgs.append('more than ks=%d simultaneous contacts on a particle '
                        '(pass a larger ks)' % self.ks)
        if st & _lib.STATUS_LIST_OVERFLOW:
            msgs.append('per-particle neighbour list overflow (list_cap=%d; '
                        'pass a larger list_cap or a smaller skin_factor)'
                        % self.list_cap)
        if msgs and raise_on_error:
            raise _lib.RbxError('device status 0x%x: %s' %
                                (st, '; '.join(msgs)))
        return st

    def read_counters(self, reset=False):
        c = self.counters.cpu().numpy().copy()
        if reset:
            self.counters[:4].zero_()
        return {'gated_pairs': int(c[0]), 'active_slots': int(c[1]),
                'candidates': int(c[2]), 'list_entries': int(c[3])}

    def grid_info(self):
        raw = self.C['info'].cpu().numpy().tobytes()
        return _lib.RbxGridInfo.from_buffer_copy(raw)

    # ------------------------------------------------------------------
    # parity helpers
    # ------------------------------------------------------------------
    def make_diag(self, pair_cap=0):
        """Per-slot diagnostics of a contact evaluation; pair_cap > 0 also
        collects the neighbour pairs the contact kernel accepts."""
        nr = max(self.n_rigid, 1)
        K = _lib.RBX_MAX_KEYS
        dev = self.device
        t = {'key': torch.full((K * nr,), -1, dtype=torch.int32, device=dev),
             'closest': torch.full((K * nr,), -1, dtype=torch.int32,
                                   device=dev)}
        for n in ['nx', 'ny', 'nz', 'dist', 'overlap', 'ftx', 'fty', 'ftz']:
            t[n] = torch.zeros(K * nr, dtype=torch.float64, device=dev)
        d = RbxDiag()
        for n, v in t.items():
            setattr(d, n, _ptr(v))
        if pair_cap > 0:
            t['pairs'] = torch.zeros(2 * pair_cap, dtype=torch.int32,
                                     device=dev)
            t['pair_count'] = torch.zeros(1, dtype=torch.int64, device=dev)
            d.pairs, d.pair_count = _ptr(t['pairs']), _ptr(t['pair_count'])
            d.pair_cap = int(pair_cap)
        return d, t

    def contact_pairs(self, dt, pair_cap=None):
        """The (destination, source) pairs one contact evaluation acts on --
        every neighbour-list entry that passes the gate of
        rigid_body_common.py:678-679 and the exact neighbour predicate with the
        current positions -- as a sorted int array [npairs, 2] of GLOBAL
        particle indices.  Uses the neighbour lists as they are (rebuilt only
        if the rebuild flag is up) and leaves the history untouched."""
        if pair_cap is None:
            pair_cap = max(self.n_rigid, 1) * self.list_cap
        d, t = self.make_diag(pair_cap)
        d.key = None               # pairs only
        p = self.params(dt)
        self.push_touched()
        self.cells_build()
        _lib.check(self.lib.rbx_contact_mofidi(
            ctypes.byref(self.scene), ctypes.byref(self._cells),
            ctypes.byref(p), ctypes.byref(d), self.stream), 'pairs')
        n = int(t['pair_count'].item())
        if n > pair_cap:
            raise _lib.RbxError('pair buffer too small: %d > %d' %
                                (n, pair_cap))
        out = t['pairs'][:2 * n].view(n, 2).cpu().numpy()
        order = np.lexsort((out[:, 1], out[:, 0]))
        return out[order]

    def history(self):
        """Current history as host arrays: key [ks,n], dlt/fn [3,ks,n].  The
        kernels end a particle's entries with one key of -1 and leave what
        follows stale; here everything past the terminator reads as unused
        (key -1, zeros)."""
        h = self.H[self.parity]
        nr = max(self.n_rigid, 1)
        key = h['key'].view(self.ks, nr).cpu().numpy().copy()
        dlt = h['dlt'].view(3, self.ks, nr).cpu().numpy().copy()
        fn = h['fn'].view(3, self.ks, nr).cpu().numpy().copy()
        used = np.logical_and.accumulate(key >= 0, axis=0)
        key[~used] = -1
        dlt[:, ~used] = 0.
        fn[:, ~used] = 0.
        return key, dlt, fn

    def set_history(self, name, key, dlt, fn):
        """Load the contact history of rigid array ``name`` (restart from a
        saved state): key [k, n] source dem_id per slot (-1 = unused, the used
        ones first), dlt / fn [3, k, n] = delta_lt and fn of those slots
        (rigid_body_common.py:1005-1012), k <= ks."""
        pas = dict((a.name, a) for a in self.rigid)
        o = self.p_off[name]
        n = pas[name].get_number_of_particles()
        key = np.asarray(key, dtype=np.int32)
        k = key.shape[0]
        if k > self.ks or key.shape[1] != n:
            raise ValueError('history of %d slots x %d particles does not '
                             'fit ks=%d, n=%d' % (k, key.shape[1], self.ks, n))
        nr = max(self.n_rigid, 1)
        h = self.H[self.parity]
        hk = h['key'].view(self.ks, nr)
        hd = h['dlt'].view(3, self.ks, nr)
        hf = h['fn'].view(3, self.ks, nr)
        hk[:, o:o + n] = -1
        hd[:, :, o:o + n] = 0.
        hf[:, :, o:o + n] = 0.
        self.invalidate_outputs()
        hk[:k, o:o + n] = self._t(key, torch.int32)
        hd[:, :k, o:o + n] = self._t(np.asarray(dlt, dtype=np.float64),
                                     torch.float64)
        hf[:, :k, o:o + n] = self._t(np.asarray(fn, dtype=np.float64),
                                     torch.float64)

    def pairs(self, dst_name, src_name):
        """NNPS neighbour pairs (i, j) of array dst among array src, as a
        sorted int array [npairs, 2] of array-local indices (parity mode)."""
        pas = dict((a.name, a) for a in self.arrays)
        dev = self.device
        so, sn = self.p_off[src_name], pas[src_name].get_number_of_particles()
        do, dn = self.p_off[dst_name], pas[dst_name].get_number_of_particles()
        sidx = torch.arange(so, so + sn, dtype=torch.int32, device=dev)
        didx = torch.arange(do, do + dn, dtype=torch.int32, device=dev)
        saved = (self.C, self.cap_cells, self.cap_points, self.workspace,
                 self._cells)
        self._alloc_cells(max(sn, 1), sidx)
        c = RbxCells()
        c.cap_cells, c.cap_points = self.cap_cells, self.cap_points
        for n in ['info', 'cell_start', 'cell_of', 'rank', 'gidx', 'sx', 'sy',
                  'sz', 'sh', 'sdem']:
            setattr(c, n, _ptr(self.C[n]))
        c.cond = None
        self.cells_build(self.points(sidx), c)
        dpts = self.points(didx)
        counts = torch.zeros(max(dn, 1), dtype=torch.int32, device=dev)
        _lib.check(self.lib.rbx_pairs_dump(
            ctypes.byref(dpts), ctypes.byref(c), self.radius_scale,
            _ptr(counts), None, None, self.stream), 'pairs count')
        offs = torch.zeros(dn + 1, dtype=torch.int64, device=dev)
        offs[1:] = torch.cumsum(counts[:dn].long(), 0)
        tot = int(offs[-1].item())
        idx = torch.zeros(max(tot, 1), dtype=torch.int32, device=dev)
        _lib.check(self.lib.rbx_pairs_dump(
            ctypes.byref(dpts), ctypes.byref(c), self.radius_scale,
            _ptr(counts), _ptr(offs), _ptr(idx), self.stream), 'pairs dump')
        torch.cuda.synchronize(dev)
        (self.C, self.cap_cells, self.cap_points, self.workspace,
         self._cells) = saved
        i = torch.repeat_interleave(torch.arange(dn, device=dev),
                                    counts[:dn].long())
        j = idx[:tot].long() - so
        out = torch.stack([i, j], 1).cpu().numpy().astype(np.int32)
        order = np.lexsort((out[:, 1], out[:, 0]))
        return out[order]


class BoundaryStream(object):
    """Host-driven boundary, streamed: the state of boundary array ``name``
    for step k + 1 travels host -> device (pinned memory, a copy stream, a
    staging buffer) while step k computes; ``apply`` hands the staged state to
    the scene on the compute stream (DeviceScene.static_update).  This is how
    an Application whose post_step moves a wall every step
    (stack_of_cylinders.py:438-445 does it once) feeds the device path without
    stalling it; ``pa.x[:] = ...`` works too, but uploads synchronously from
    pageable memory and rebuilds the neighbour lists unconditionally."""

    def __init__(self, scene, name, props=('x', 'y', 'z', 'u', 'v', 'w')):
        self.sc, self.name, self.props = scene, name, tuple(props)
        pas = dict((a.name, a) for a in scene.bounds)
        self.n = pas[name].get_number_of_particles()
        dev = scene.device
        self.copy_stream = torch.cuda.Stream(dev)
        self.stage = [dict((p, torch.empty(self.n, dtype=torch.float64,
                                           device=dev)) for p in self.props)
                      for _ in range(2)]
        self.ev_up = [torch.cuda.Event() for _ in range(2)]
        self.ev_used = [torch.cuda.Event() for _ in range(2)]
        cur = torch.cuda.current_stream(dev)
        for e in self.ev_used:
            e.record(cur)
        self.k_sub = self.k_app = 0
        self.bytes_per_submit = 8 * self.n * len(self.props)

    def submit(self, host):
        """host: {prop: pinned float64 CPU tensor of the array's length}."""
        k = self.k_sub & 1
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.ev_used[k])  # staging is free
            for p in self.props:
                self.stage[k][p].copy_(host[p], non_blocking=True)
            self.ev_up[k].record(self.copy_stream)
        self.k_sub += 1

    def apply(self):
        if self.k_app >= self.k_sub:
            raise RuntimeError('BoundaryStream.apply without a submit')
        k = self.k_app & 1
        cur = torch.cuda.current_stream(self.sc.device)
        cur.wait_event(self.ev_up[k])
        self.sc.static_update(self.name, self.stage[k])
        self.ev_used[k].record(cur)
        self.k_app += 1


class BodyStateStream(object):
    """Per-body results, streamed the other way: ``snapshot`` copies the named
    per-body arrays on the compute stream and sends the copy to pinned host
    memory on a copy stream; ``wait(k)`` returns the host tensors of snapshot
    k once they have arrived.  Two snapshots can be in flight."""

    def __init__(self, scene, names=('xcm', 'vcm', 'omega', 'R', 'force',
                                     'torque')):
        self.sc, self.names = scene, tuple(names)
        dev = scene.device
        self.copy_stream = torch.cuda.Stream(dev)
        self.snap = [dict((n, torch.empty_like(scene.B[n])) for n in names)
                     for _ in range(2)]
        self.host = [dict((n, torch.empty_like(scene.B[n],
                                               device='cpu').pin_memory())
                          for n in names) for _ in range(2)]
        self.ev_snap = [torch.cuda.Event() for _ in range(2)]
        self.ev_out = [torch.cuda.Event() for _ in range(2)]
        cur = torch.cuda.current_stream(dev)
        for e in self.ev_out:
            e.record(cur)
        self.k = 0
        self.bytes_per_snapshot = sum(8 * scene.B[n].numel() for n in names)

    def snapshot(self):
        k = self.k & 1
        cur = torch.cuda.current_stream(self.sc.device)
        cur.wait_event(self.ev_out[k])          # snapshot buffer is free
        for n in self.names:
            self.snap[k][n].copy_(self.sc.B[n])
        self.ev_snap[k].record(cur)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.ev_snap[k])
            for n in self.names:
                self.host[k][n].copy_(self.snap[k][n], non_blocking=True)
            self.ev_out[k].record(self.copy_stream)
        self.k += 1
        return self.k - 1

    def wait(self, k):
        self.ev_out[k & 1].synchronize()
        return self.host[k & 1]
