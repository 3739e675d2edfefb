"""Setup path on the device.

``identify_boundary(pa, dim)`` fills ``normal_tmp``, ``normal``,
``normal_norm`` and ``is_boundary`` of one particle array exactly like
``SPHEvaluator(ComputeNormals -> SmoothNormals ->
IdentifyBoundaryParticleCosAngle)`` does on the host
(compat/sph_evaluator.py), but with the CUDA cell list and three
thread-per-particle kernels (csrc/rbx_setup.cu): seconds instead of minutes
at 10^7 particles.  ``RigidBody3DScheme.setup_properties`` stays on the host
(it has to work where the scene is merely *built*); pass
``device_setup=True`` to the scheme to route its boundary identification here.
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
