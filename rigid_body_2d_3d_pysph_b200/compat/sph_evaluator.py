"""``pysph.tools.sph_evaluator.SPHEvaluator`` for the setup-time equations.

The reference evaluates three equation groups once per array at setup
(/root/reference/code/rigid_body_3d.py:862-871): ComputeNormals ->
SmoothNormals -> IdentifyBoundaryParticleCosAngle.  This evaluator knows
exactly those (by class name) and runs them as vectorised NumPy over
neighbour pairs found with a KD-tree and filtered by the NNPS predicate
(SURVEY.md App. C-1).  Any other equation is an error: the hot-path
equations run on the GPU through the integrator, never here.
"""
import numpy as np
from scipy.spatial import cKDTree

from .kernels import QuinticSpline


def _pairs(dst, src, radius_scale):
    """All (i, j) with r2 < (k h_i)^2 or r2 < (k h_j)^2, sorted by (i, j)."""
    n = dst.get_number_of_particles()
    m = src.get_number_of_particles()
    if n == 0 or m == 0:
        return (np.zeros(0, np.int64),) * 2
    a = np.c_[dst.x, dst.y, dst.z]
    b = np.c_[src.x, src.y, src.z]
    hmax = max(dst.h.max(), src.h.max())
    tree_b = cKDTree(b)
    tree_a = cKDTree(a)
    sp = tree_a.sparse_distance_matrix(tree_b, radius_scale * hmax * 1.0001,
                                       output_type='coo_matrix')
    i = sp.row.astype(np.int64)
    j = sp.col.astype(np.int64)
    dx = dst.x[i] - src.x[j]
    dy = dst.y[i] - src.y[j]
    dz = dst.z[i] - src.z[j]
    r2 = dx * dx + dy * dy + dz * dz
    rs2 = radius_scale * radius_scale
    keep = (r2 < rs2 * dst.h[i] * dst.h[i]) | (r2 < rs2 * src.h[j] * src.h[j])
    i, j = i[keep], j[keep]
    order = np.lexsort((j, i))
    return i[order], j[order]


class SPHEvaluator(object):
    def __init__(self, arrays, equations, dim, kernel=None, domain_manager=None,
                 backend=None, nnps_factory=None, **kw):
        self.arrays = arrays
        self.equations = equations
        self.dim = dim
        self.kernel = kernel if kernel is not None else QuinticSpline(dim=dim)

    def evaluate(self, t=0.0, dt=0.1):
        pas = dict((a.name, a) for a in self.arrays)
        for group in self.equations:
            for eq in group.equations:
                fn = getattr(self, '_' + eq.__class__.__name__, None)
                if fn is None:
                    raise NotImplementedError(
                        'SPHEvaluator (setup path) does not implement %s' %
                        eq.__class__.__name__)
                fn(pas[eq.dest], [pas[s] for s in (eq.sources or [])])

    # -- ComputeNormals: boundary_particles.py:71-112 (renamed props) -------
    def _ComputeNormals(self, dst, srcs):
        n = dst.get_number_of_particles()
        tmp = np.zeros((n, 3))
        dst.normal[:] = 0.0
        k = self.kernel
        for src in srcs:
            i, j = _pairs(dst, src, k.radius_scale)
            xij = np.c_[dst.x[i] - src.x[j], dst.y[i] - src.y[j],
                        dst.z[i] - src.z[j]]
            rij = np.sqrt(xij[:, 0] * xij[:, 0] + xij[:, 1] * xij[:, 1] +
                          xij[:, 2] * xij[:, 2])
            hij = 0.5 * (dst.h[i] + src.h[j])
            wdash = k.dwdq_np(rij, hij)
            with np.errstate(all='ignore'):
                g = np.where(rij > 1e-12, wdash * (1. / hij) / rij, 0.0)
            fac = -src.m[j] / src.rho[j]
            for c in range(3):
                tmp[:, c] += np.bincount(i, weights=fac * (g * xij[:, c]),
                                         minlength=n)
        mag = np.sqrt(tmp[:, 0]**2 + tmp[:, 1]**2 + tmp[:, 2]**2)
        ok = mag > 0.25 / dst.h
        with np.errstate(all='ignore'):
            tmp = np.where(ok[:, None], tmp / mag[:, None], 0.0)
        dst.normal_tmp[:] = tmp.ravel()

    # -- SmoothNormals: boundary_particles.py:114-135 -----------------------
    def _SmoothNormals(self, dst, srcs):
        n = dst.get_number_of_particles()
        nrm = dst.normal.reshape(n, 3).copy()
        k = self.kernel
        for src in srcs:
            i, j = _pairs(dst, src, k.radius_scale)
            dx = dst.x[i] - src.x[j]
            dy = dst.y[i] - src.y[j]
            dz = dst.z[i] - src.z[j]
            rij = np.sqrt(dx * dx + dy * dy + dz * dz)
            hij = 0.5 * (dst.h[i] + src.h[j])
            fac = src.m[j] / src.rho[j] * k.kernel_np(rij, hij)
            st = src.normal_tmp.reshape(-1, 3)
            for c in range(3):
                nrm[:, c] += np.bincount(i, weights=fac * st[j, c],
                                         minlength=n)
        mag = np.sqrt(nrm[:, 0]**2 + nrm[:, 1]**2 + nrm[:, 2]**2)
        ok = mag > 1e-3
        with np.errstate(all='ignore'):
            nrm = np.where(ok[:, None], nrm / mag[:, None], 0.0)
        dst.normal[:] = nrm.ravel()

    # -- IdentifyBoundaryParticleCosAngle: boundary_particles.py:22-68 ------
    def _IdentifyBoundaryParticleCosAngle(self, dst, srcs):
        n = dst.get_number_of_particles()
        nrm = dst.normal.reshape(n, 3)
        norm = nrm[:, 0]**2. + nrm[:, 1]**2. + nrm[:, 2]**2.
        dst.normal_norm[:] = norm
        isb = (norm > 1e-6).astype(np.int32)
        k = self.kernel
        for src in srcs:
            i, j = _pairs(dst, src, k.radius_scale)
            dx = dst.x[i] - src.x[j]
            dy = dst.y[i] - src.y[j]
            dz = dst.z[i] - src.z[j]
            rij = np.sqrt(dx**2. + dy**2. + dz**2.)
            hi = dst.h[i]
            dot = -(nrm[i, 0] * dx + nrm[i, 1] * dy + nrm[i, 2] * dz)
            with np.errstate(all='ignore'):
                fac = dot / rij
            hit = (rij > 1e-9 * hi) & (rij < 2. * hi) & (fac > 0.5)
            kill = np.bincount(i[hit], minlength=n) > 0
            isb[kill] = 0
        dst.is_boundary[:] = isb
