"""``DEMScheme`` and its equations / stepper -- same surface as the reference's
``code/dem.py`` (linear visco-elastic contact with Coulomb friction, per-pair
tangential displacement history; Luding 2008), B200-native underneath
(csrc/rbx_lvc.cu through ``DemDeviceScene``).

Only the LVCDisplacement model is reachable in the reference: both branches of
``DEMScheme._get_gtvf_equations`` test ``== "LVCDisplacement"`` (dem.py:722,
729, 746, 750), and the ``--contact-model`` flag only accepts 'LVC', which
selects no model at all (dem.py:681-687).  The same behaviour is kept.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import RbxCells, RbxDemScene, RbxParams, RbxPoints
from .compat.equation import Equation, Group, MultiStageEquations
from .compat.integrator import GTVFIntegrator, IntegratorStep
from .compat.kernels import CubicSpline
from .compat.scheme import Scheme
from .rigid_body_common import BodyForce


class LVCDisplacement(Equation):
    """dem.py:35-205 -> rbx_contact_lvc (k_dem_force)."""


class UpdateTangentialContactsLVCDisplacement(Equation):
    """dem.py:208-293 -> rbx_contact_lvc (k_dem_update)."""


class DEMStep(IntegratorStep):
    """dem.py:595-625 -> rbx_dem_step."""
    kind = 'dem'


class DEMScheme(Scheme):
    def __init__(self, granular_particles, boundaries, kn=1e5, en=0.5,
                 integrator="gtvf", dim=2, gx=0.0, gy=0.0, gz=0.0,
                 kernel_choice="1", kernel_factor=3,
                 contact_model="LVCDisplacement"):
        self.granular_particles = granular_particles
        self.boundaries = [] if boundaries is None else boundaries
        self.dim = dim
        self.kernel = CubicSpline
        self.integrator = integrator
        self.gx, self.gy, self.gz = gx, gy, gz
        self.kn = kn
        self.en = en
        self.contact_model = contact_model
        self.solver = None

    def add_user_options(self, group):
        choices = ['LVC']
        group.add_argument("--contact-model", action="store",
                           dest='contact_model', default="LVCDisplacement",
                           choices=choices,
                           help="Specify what contact model to use %s" %
                           choices)

    def consume_user_options(self, options):
        _vars = ['contact_model']
        data = dict((var, self._smart_getattr(options, var)) for var in _vars)
        self.configure(**data)

    def get_equations(self):
        return self._get_gtvf_equations()

    def _get_gtvf_equations(self):
        """dem.py:697-756.  The reference builds the source list with
        ``list(set(...))`` (hash order); here it is granular arrays first,
        boundaries after -- only the order of history entries depends on it."""
        all = []
        for n in self.granular_particles + self.boundaries:
            if n not in all:
                all.append(n)
        stage1, stage2 = [], []
        if self.contact_model == "LVCDisplacement":
            stage2.append(Group(equations=[
                UpdateTangentialContactsLVCDisplacement(dest=g, sources=all)
                for g in self.granular_particles], real=False))
        g2 = [BodyForce(dest=g, sources=None, gx=self.gx, gy=self.gy,
                        gz=self.gz) for g in self.granular_particles]
        if self.contact_model == "LVCDisplacement":
            g2 += [LVCDisplacement(dest=g, sources=all)
                   for g in self.granular_particles]
        stage2.append(Group(equations=g2, real=False))
        return MultiStageEquations([stage1, stage2])

    def configure_solver(self, kernel=None, integrator_cls=None,
                         extra_steppers=None, **kw):
        from .compat.solver import Solver
        if kernel is None:
            kernel = CubicSpline(dim=self.dim)
        steppers = {}
        if extra_steppers is not None:
            steppers.update(extra_steppers)
        for g in self.granular_particles:
            if g not in steppers:
                steppers[g] = DEMStep()
        cls = integrator_cls if integrator_cls is not None else GTVFIntegrator
        self.solver = Solver(dim=self.dim, integrator=cls(**steppers),
                             kernel=kernel, **kw)

    def setup_properties(self, particles, clean=True):
        """dem.py:785-825"""
        pas = dict([(p.name, p) for p in particles])
        for name in self.granular_particles:
            pa = pas[name]
            for p in ('fx', 'fy', 'fz', 'torx', 'tory', 'torz', 'wx', 'wy',
                      'wz'):
                pa.add_property(p)
            limit = int(pa.max_tng_contacts_limit[0])
            pa.add_property('tng_idx', stride=limit, type="int")
            pa.tng_idx[:] = -1
            pa.add_property('tng_idx_dem_id', stride=limit, type="int")
            pa.tng_idx_dem_id[:] = -1
            if self.contact_model == "LVCDisplacement":
                for p in ('tng_x', 'tng_y', 'tng_z'):
                    pa.add_property(p, stride=limit)
            if self.contact_model == "LVCForce":
                for p in ('tng_fx', 'tng_fy', 'tng_fz'):
                    pa.add_property(p, stride=limit)
            pa.add_property('total_tng_contacts', type="int")
            pa.total_tng_contacts[:] = 0
            pa.set_output_arrays(['x', 'y', 'z', 'u', 'v', 'w', 'fx', 'fy',
                                  'fz', 'm', 'moi'])

    def get_solver(self):
        return self.solver


# ----------------------------------------------------------------------
def _ptr(t):
    return None if t is None else t.data_ptr()


_ALL_F64 = ['x', 'y', 'z', 'u', 'v', 'w', 'wx', 'wy', 'wz', 'h', 'm',
            'rad_s']
_DEST_F64 = ['moi', 'fx', 'fy', 'fz', 'torx', 'tory', 'torz']
_MUTATED = ['x', 'y', 'z', 'u', 'v', 'w', 'wx', 'wy', 'wz', 'fx', 'fy', 'fz',
            'torx', 'tory', 'torz', 'tng_idx', 'tng_idx_dem_id', 'tng_x',
            'tng_y', 'tng_z', 'total_tng_contacts']


class DemDeviceScene(object):
    """HBM layout and host driver of the DEMScheme step (cf. device.py)."""

    def __init__(self, arrays, granular_names, boundary_names=(), dim=2,
                 gx=0., gy=0., gz=0., radius_scale=2.0, device=None):
        if not torch.cuda.is_available():
            raise _lib.RbxError('DemDeviceScene needs a CUDA device; there '
                                'is no CPU fallback')
        self.lib = _lib.load()
        self.device = torch.device(device if device is not None else
                                   'cuda:%d' % torch.cuda.current_device())
        pas = dict((a.name, a) for a in arrays)
        self.dest = [pas[n] for n in granular_names]
        self.bounds = [pas[n] for n in boundary_names]
        self.arrays = self.dest + self.bounds
        self.dim = dim
        self.g = (float(gx), float(gy), float(gz))
        self.radius_scale = float(radius_scale)
        dev = self.device
        f64, i32 = torch.float64, torch.int32

        def t(a, dt):
            return torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(dev)
        self.p_off = {}
        off = 0
        for pa in self.arrays:
            self.p_off[pa.name] = off
            off += pa.get_number_of_particles()
        self.n_total = off
        self.n_dest = sum(pa.get_number_of_particles() for pa in self.dest)
        limits = set(int(pa.max_tng_contacts_limit[0]) for pa in self.dest)
        if len(limits) != 1:
            raise ValueError('granular arrays must share '
                             'max_tng_contacts_limit')
        self.limit = limits.pop()
        self.P = {}
        for n in _ALL_F64:
            self.P[n] = t(np.concatenate([
                pa.properties[n] if n in pa.properties
                else np.zeros(pa.get_number_of_particles())
                for pa in self.arrays]), f64)
        self.P['dem_id'] = t(np.concatenate(
            [pa.properties['dem_id'] for pa in self.arrays]), i32)
        for n in _DEST_F64:
            self.P[n] = t(np.concatenate([pa.properties[n]
                                          for pa in self.dest]), f64)
        self.P['tng_idx'] = t(np.concatenate(
            [pa.properties['tng_idx'] for pa in self.dest]), i32)
        self.P['tng_idx_dem_id'] = t(np.concatenate(
            [pa.properties['tng_idx_dem_id'] for pa in self.dest]), i32)
        for n in ('tng_x', 'tng_y', 'tng_z'):
            self.P[n] = t(np.concatenate([pa.properties[n]
                                          for pa in self.dest]), f64)
        self.P['total_tng_contacts'] = t(np.concatenate(
            [pa.properties['total_tng_contacts'] for pa in self.dest]), i32)
        starts = [self.p_off[pa.name] for pa in self.arrays] + [self.n_total]
        arr_off = np.concatenate([
            np.full(pa.get_number_of_particles(), self.p_off[pa.name])
            for pa in self.arrays])
        self.T = {'arr_start': t(np.array(starts), i32),
                  'arr_off': t(arr_off, i32)}
        ndem = int(max(int(pa.properties['dem_id'].max())
                       for pa in self.arrays)) + 1
        rows, tabs = [], dict((k, []) for k in ('kn', 'kt', 'alpha', 'mu'))
        for a, pa in enumerate(self.dest):
            rows.append(np.full(pa.get_number_of_particles(), a * ndem))
            for k in tabs:
                v = np.zeros(ndem)
                c = np.asarray(pa.constants[k], dtype=np.float64)
                v[:min(ndem, c.size)] = c[:ndem]
                tabs[k].append(v)
        self.T['tbl_row'] = t(np.concatenate(rows), i32)
        for k in tabs:
            self.T[k] = t(np.concatenate(tabs[k]), f64)
        self.status = torch.zeros(1, dtype=i32, device=dev)
        self.hmax = float(self.P['h'].max().item())
        self.reach = self.radius_scale * self.hmax
        # cell list over ALL particles
        ncell = 1
        for n in 'xyz':
            ext = float((self.P[n].max() - self.P[n].min()).item())
            ncell *= int(ext / self.reach) + 2
        self.cap_cells = int(min(max(8 * ncell, 4096), 1 << 26))
        n = max(self.n_total, 1)
        self.C = {'info': torch.zeros(64, dtype=torch.uint8, device=dev),
                  'cell_start': torch.zeros(self.cap_cells + 1, dtype=i32,
                                            device=dev)}
        for k in ('cell_of', 'rank', 'gidx', 'sdem'):
            self.C[k] = torch.zeros(n, dtype=i32, device=dev)
        for k in ('sx', 'sy', 'sz', 'sh'):
            self.C[k] = torch.zeros(n, dtype=f64, device=dev)
        self.workspace = torch.zeros(
            self.lib.rbx_cells_workspace_bytes(self.cap_cells, n),
            dtype=torch.uint8, device=dev)
        s = RbxDemScene()
        s.n_total, s.n_dest = self.n_total, self.n_dest
        s.n_arrays, s.limit = len(self.arrays), self.limit
        for k in _ALL_F64 + _DEST_F64 + ['dem_id', 'tng_idx', 'tng_x',
                                        'tng_y', 'tng_z']:
            setattr(s, k, _ptr(self.P[k]))
        s.tng_dem = _ptr(self.P['tng_idx_dem_id'])
        s.total_tng = _ptr(self.P['total_tng_contacts'])
        for k in ('arr_off', 'arr_start', 'tbl_row', 'kn', 'kt', 'alpha',
                  'mu'):
            setattr(s, k, _ptr(self.T[k]))
        s.status = _ptr(self.status)
        self._scene = s
        c = RbxCells()
        c.cap_cells, c.cap_points = self.cap_cells, n
        for k in ('info', 'cell_start', 'cell_of', 'rank', 'gidx', 'sx', 'sy',
                  'sz', 'sh', 'sdem'):
            setattr(c, k, _ptr(self.C[k]))
        self._cells = c
        p = RbxPoints()
        p.n, p.index = self.n_total, None
        p.x, p.y, p.z, p.h = (_ptr(self.P[k]) for k in 'xyzh')
        p.dem_id = _ptr(self.P['dem_id'])
        self._pts = p
        self.steps_done = 0
        for pa in self.arrays:
            pa.__dict__['_device'] = self
            pa.__dict__['_host_touched'].clear()
            pa.__dict__['_device_newer'].clear()

    @property
    def stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _slice(self, pa, name):
        n = pa.get_number_of_particles()
        o = self.p_off[pa.name]
        if name in _ALL_F64 or name == 'dem_id':
            if name in pa.properties:
                return self.P[name][o:o + n]
            return None
        if pa in self.dest:
            if name in _DEST_F64 or name == 'total_tng_contacts':
                return self.P[name][o:o + n]
            if name in ('tng_idx', 'tng_idx_dem_id', 'tng_x', 'tng_y',
                        'tng_z'):
                return self.P[name][o * self.limit:(o + n) * self.limit]
        return None

    def pull(self, pa, name):
        t = self._slice(pa, name)
        if t is not None:
            pa.properties[name][:] = t.cpu().numpy()

    def push_touched(self):
        for pa in self.arrays:
            touched = pa.__dict__['_host_touched']
            for name in list(touched):
                t = self._slice(pa, name)
                if t is not None and name in pa.properties:
                    t.copy_(torch.as_tensor(pa.properties[name],
                                            dtype=t.dtype))
            touched.clear()

    def mark_device_newer(self):
        for pa in self.dest:
            dn = pa.__dict__['_device_newer']
            for n in _MUTATED:
                if n in pa.properties:
                    dn.add(n)

    def sync_to_host(self):
        for pa in self.arrays:
            for n in list(pa.__dict__['_device_newer']):
                self.pull(pa, n)
            pa.__dict__['_device_newer'].clear()

    def params(self, dt):
        return RbxParams(self.radius_scale, 0., 0., 0., self.g[0], self.g[1],
                         self.g[2], float(dt), self.reach, 0.)

    def gtvf_step(self, dt, nsteps=1, graph=False):
        """GTVFIntegrator.one_timestep with DEMStep (dem.py:595-625)."""
        self.push_touched()
        L, S, st = self.lib, ctypes.byref(self._scene), None
        p = self.params(dt)
        for _ in range(nsteps):
            st = self.stream
            _lib.check(L.rbx_dem_step(S, 1, float(dt), st), 'dem stage1')
            _lib.check(L.rbx_dem_step(S, 2, float(dt), st), 'dem stage2')
            _lib.check(L.rbx_cells_build(
                ctypes.byref(self._pts), ctypes.byref(self._cells),
                self.reach, _ptr(self.status), _ptr(self.workspace),
                self.workspace.numel(), st), 'cells_build')
            _lib.check(L.rbx_contact_lvc(S, ctypes.byref(self._cells),
                                         ctypes.byref(p), st), 'contact_lvc')
            _lib.check(L.rbx_dem_step(S, 3, float(dt), st), 'dem stage3')
        self.steps_done += nsteps
        self.mark_device_newer()

    def check_status(self, raise_on_error=True):
        st = int(self.status.item()) & 0xffffffff
        if (st & _lib.STATUS_LVC_OVERFLOW) and raise_on_error:
            raise _lib.RbxError(
                'a particle has more than max_tng_contacts_limit=%d '
                'simultaneous contacts (the reference would write past its '
                'list, dem.py:144-148)' % self.limit)
        return st
