"""Setup path on the device.

``identify_boundary(pa, dim)`` fills ``normal_tmp``, ``normal``,
``normal_norm`` and ``is_boundary`` of one particle array exactly like
``SPHEvaluator(ComputeNormals -> SmoothNormals ->
IdentifyBoundaryParticleCosAngle)`` does on the host
(compat/sph_evaluator.py), but with the CUDA cell list and three
thread-per-particle kernels (csrc/rbx_setup.cu): seconds instead of minutes
at 10^7 particles.  ``setup_rigid_bodies(pa)`` does the per-body part of the
setup -- total mass, centre of mass, izz, inertia tensor and inverse,
body-frame position vectors (rigid_body_common.py:21-107) -- with one warp
per body.  ``RigidBody3DScheme.setup_properties`` stays on the host by default
(it has to work where the scene is merely *built*); ``scheme.device_setup =
True`` routes both parts here.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import RbxCells, RbxPoints
from .boundary_particles import add_boundary_identification_properties


def identify_boundary(pa, dim, radius_scale=3.0, device=None):
    if not torch.cuda.is_available():
        raise _lib.RbxError('setup_device.identify_boundary needs a CUDA '
                            'device (the host path is '
                            'compat.sph_evaluator.SPHEvaluator)')
    L = _lib.load()
    dev = torch.device(device if device is not None else
                       'cuda:%d' % torch.cuda.current_device())
    add_boundary_identification_properties(pa)
    n = pa.get_number_of_particles()
    f64, i32 = torch.float64, torch.int32

    def t(a, dt=f64):
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(dev)
    x, y, z, h = t(pa.x), t(pa.y), t(pa.z), t(pa.h)
    m, rho = t(pa.m), t(pa.rho)
    dem = torch.zeros(max(n, 1), dtype=i32, device=dev)
    reach = radius_scale * float(h.max().item()) if n else 1.0
    ncell = 1
    for v in (x, y, z):
        ncell *= int(float((v.max() - v.min()).item()) / reach) + 2 if n else 1
    cap_cells = int(min(max(2 * ncell, 4096), 1 << 26))
    C = {'info': torch.zeros(64, dtype=torch.uint8, device=dev),
         'cell_start': torch.zeros(cap_cells + 1, dtype=i32, device=dev)}
    for k in ('cell_of', 'rank', 'gidx', 'sdem'):
        C[k] = torch.zeros(max(n, 1), dtype=i32, device=dev)
    for k in ('sx', 'sy', 'sz', 'sh'):
        C[k] = torch.zeros(max(n, 1), dtype=f64, device=dev)
    ws = torch.zeros(L.rbx_cells_workspace_bytes(cap_cells, max(n, 1)),
                     dtype=torch.uint8, device=dev)
    status = torch.zeros(1, dtype=i32, device=dev)
    pts = RbxPoints()
    pts.n, pts.index = n, None
    pts.x, pts.y, pts.z, pts.h = (v.data_ptr() for v in (x, y, z, h))
    pts.dem_id = dem.data_ptr()
    c = RbxCells()
    c.cap_cells, c.cap_points = cap_cells, max(n, 1)
    for k in ('info', 'cell_start', 'cell_of', 'rank', 'gidx', 'sx', 'sy',
              'sz', 'sh', 'sdem'):
        setattr(c, k, C[k].data_ptr())
    c.cond = None
    stream = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(L.rbx_cells_build(ctypes.byref(pts), ctypes.byref(c), reach,
                                 status.data_ptr(), ws.data_ptr(), ws.numel(),
                                 stream), 'rbx_cells_build')
    ntmp = torch.zeros(3 * max(n, 1), dtype=f64, device=dev)
    nrm = torch.zeros(3 * max(n, 1), dtype=f64, device=dev)
    isb = torch.zeros(max(n, 1), dtype=i32, device=dev)
    _lib.check(L.rbx_boundary_identify(
        ctypes.byref(pts), ctypes.byref(c), int(dim), float(radius_scale),
        m.data_ptr(), rho.data_ptr(), ntmp.data_ptr(), nrm.data_ptr(),
        isb.data_ptr(), stream), 'rbx_boundary_identify')
    torch.cuda.synchronize(dev)
    pa.normal_tmp[:] = ntmp[:3 * n].cpu().numpy()
    pa.normal[:] = nrm[:3 * n].cpu().numpy()
    nn = pa.normal.reshape(n, 3)
    pa.normal_norm[:] = nn[:, 0]**2. + nn[:, 1]**2. + nn[:, 2]**2.
    pa.is_boundary[:] = isb[:n].cpu().numpy()
    return pa


def setup_rigid_bodies(pa, device=None, tensor=True):
    """set_total_mass + set_center_of_mass + set_moment_of_inertia_izz +
    set_moment_of_inertia_and_its_inverse + set_body_frame_position_vectors
    of rigid array ``pa`` (particles grouped by ascending body_id) on the
    device; the results land in the array's constants / properties exactly
    where the host helpers put them.  tensor=False (the 2-D scheme, which
    only uses izz: rigid_body_2d.py) leaves the inertia tensors alone."""
    if not torch.cuda.is_available():
        raise _lib.RbxError('setup_device.setup_rigid_bodies needs a CUDA '
                            'device (the host path is rigid_body_common)')
    L = _lib.load()
    dev = torch.device(device if device is not None else
                       'cuda:%d' % torch.cuda.current_device())
    bid = np.asarray(pa.body_id, dtype=np.int64)
    if bid.size and np.any(np.diff(bid) < 0):
        raise ValueError('particles must be grouped by ascending body_id')
    nb = int(bid.max()) + 1 if bid.size else 0
    start = np.searchsorted(bid, np.arange(nb + 1)).astype(np.int32)
    f64 = torch.float64

    def t(a, dt=f64):
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(dev)
    n = pa.get_number_of_particles()
    x, y, z, m = t(pa.x), t(pa.y), t(pa.z), t(pa.m)
    st = t(start, torch.int32)
    out = dict((k, torch.zeros(max(s * nb, 1), dtype=f64, device=dev))
               for k, s in (('total_mass', 1), ('xcm', 3), ('izz', 1),
                            ('I', 9), ('Iinv', 9)))
    d0 = [torch.zeros(max(n, 1), dtype=f64, device=dev) for _ in range(3)]
    _lib.check(L.rbx_setup_bodies(
        nb, st.data_ptr(), x.data_ptr(), y.data_ptr(), z.data_ptr(),
        m.data_ptr(), out['total_mass'].data_ptr(), out['xcm'].data_ptr(),
        out['izz'].data_ptr(), out['I'].data_ptr(), out['Iinv'].data_ptr(),
        d0[0].data_ptr(), d0[1].data_ptr(), d0[2].data_ptr(),
        torch.cuda.current_stream(dev).cuda_stream), 'rbx_setup_bodies')
    torch.cuda.synchronize(dev)
    pa.total_mass[:nb] = out['total_mass'][:nb].cpu().numpy()
    pa.xcm[:3 * nb] = out['xcm'][:3 * nb].cpu().numpy()
    if 'izz' in pa.constants:
        pa.izz[:nb] = out['izz'][:nb].cpu().numpy()
    if tensor:
        I = out['I'][:9 * nb].cpu().numpy()
        Iinv = out['Iinv'][:9 * nb].cpu().numpy()
        pa.inertia_tensor_body_frame[:9 * nb] = I
        pa.inertia_tensor_inverse_body_frame[:9 * nb] = Iinv
        pa.inertia_tensor_global_frame[:9 * nb] = I
        pa.inertia_tensor_inverse_global_frame[:9 * nb] = Iinv
    for k, name in enumerate(('dx0', 'dy0', 'dz0')):
        getattr(pa, name)[:] = d0[k][:n].cpu().numpy()
    return pa
