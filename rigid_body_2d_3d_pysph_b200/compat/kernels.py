"""SPH smoothing kernels with the ``pysph.base.kernels`` call surface.

Only ``QuinticSpline`` is on the rigid-body path (it weights contact passes
1-2, /root/reference/code/rigid_body_3d.py:705-708, and its gradient feeds the
setup-time surface normals, /root/reference/code/boundary_particles.py:71-135).
The formulas restate SURVEY.md App. C-3; ``CubicSpline`` is kept because
``DEMScheme.configure_solver`` names it (/root/reference/code/dem.py:763-767),
where it only sets the search radius (2h).

Host scalar/NumPy code, used at setup time and by tests.  The device copy of
the quintic lives in csrc/rbx_math.cuh.
"""
from math import pi
import numpy as np

M_1_PI = 1.0 / pi


class QuinticSpline(object):
    def __init__(self, dim=2):
        self.radius_scale = 3.0
        self.dim = dim
        if dim == 1:
            self.fac = 1.0 / 120.0
        elif dim == 2:
            self.fac = M_1_PI * 7.0 / 478.0
        else:
            self.fac = M_1_PI / 120.0

    def _norm(self, h1):
        if self.dim == 1:
            return self.fac * h1
        elif self.dim == 2:
            return self.fac * h1 * h1
        return self.fac * h1 * h1 * h1

    def kernel(self, xij=(0., 0., 0.), rij=1.0, h=1.0):
        h1 = 1. / h
        q = rij * h1
        fac = self._norm(h1)
        tmp3 = 3. - q
        tmp2 = 2. - q
        tmp1 = 1. - q
        if q > 3.0:
            val = 0.0
        elif q > 2.0:
            val = tmp3 * tmp3 * tmp3 * tmp3 * tmp3
        elif q > 1.0:
            val = tmp3 * tmp3 * tmp3 * tmp3 * tmp3
            val -= 6.0 * tmp2 * tmp2 * tmp2 * tmp2 * tmp2
        else:
            val = tmp3 * tmp3 * tmp3 * tmp3 * tmp3
            val -= 6.0 * tmp2 * tmp2 * tmp2 * tmp2 * tmp2
            val += 15. * tmp1 * tmp1 * tmp1 * tmp1 * tmp1
        return val * fac

    def dwdq(self, rij=1.0, h=1.0):
        h1 = 1. / h
        q = rij * h1
        fac = self._norm(h1)
        tmp3 = 3. - q
        tmp2 = 2. - q
        tmp1 = 1. - q
        if rij > 1e-12:
            if q > 3.0:
                val = 0.0
            elif q > 2.0:
                val = -5.0 * tmp3 * tmp3 * tmp3 * tmp3
            elif q > 1.0:
                val = -5.0 * tmp3 * tmp3 * tmp3 * tmp3
                val += 30.0 * tmp2 * tmp2 * tmp2 * tmp2
            else:
                val = -5.0 * tmp3 * tmp3 * tmp3 * tmp3
                val += 30.0 * tmp2 * tmp2 * tmp2 * tmp2
                val -= 75.0 * tmp1 * tmp1 * tmp1 * tmp1
        else:
            val = 0.0
        return val * fac

    def gradient(self, xij=(0., 0., 0.), rij=1.0, h=1.0, grad=None):
        if grad is None:
            grad = [0.0, 0.0, 0.0]
        h1 = 1. / h
        if rij > 1e-12:
            wdash = self.dwdq(rij, h)
            tmp = wdash * h1 / rij
        else:
            tmp = 0.0
        grad[0] = tmp * xij[0]
        grad[1] = tmp * xij[1]
        grad[2] = tmp * xij[2]
        return grad

    # vectorised forms used by the host-side setup evaluator -------------
    def kernel_np(self, rij, h):
        h1 = 1. / h
        q = rij * h1
        fac = self._norm(h1)
        t3 = 3. - q
        t2 = 2. - q
        t1 = 1. - q
        v3 = t3 * t3 * t3 * t3 * t3
        v2 = 6.0 * t2 * t2 * t2 * t2 * t2
        v1 = 15. * t1 * t1 * t1 * t1 * t1
        val = np.where(q > 3.0, 0.0,
                       np.where(q > 2.0, v3,
                                np.where(q > 1.0, v3 - v2, v3 - v2 + v1)))
        return val * fac

    def dwdq_np(self, rij, h):
        h1 = 1. / h
        q = rij * h1
        fac = self._norm(h1)
        t3 = 3. - q
        t2 = 2. - q
        t1 = 1. - q
        v3 = -5.0 * t3 * t3 * t3 * t3
        v2 = 30.0 * t2 * t2 * t2 * t2
        v1 = 75.0 * t1 * t1 * t1 * t1
        val = np.where(q > 3.0, 0.0,
                       np.where(q > 2.0, v3,
                                np.where(q > 1.0, v3 + v2, v3 + v2 - v1)))
        val = np.where(rij > 1e-12, val, 0.0)
        return val * fac


class CubicSpline(object):
    def __init__(self, dim=1):
        self.radius_scale = 2.0
        self.dim = dim
        if dim == 3:
            self.fac = M_1_PI
        elif dim == 2:
            self.fac = 10 * M_1_PI / 7.0
        else:
            self.fac = 2.0 / 3.0

    def kernel(self, xij=(0., 0., 0.), rij=1.0, h=1.0):
        h1 = 1. / h
        q = rij * h1
        fac = self.fac * h1 ** self.dim
        tmp2 = 2. - q
        if q > 2.0:
            val = 0.0
        elif q > 1.0:
            val = 0.25 * tmp2 * tmp2 * tmp2
        else:
            val = 1 - 1.5 * q * q * (1 - 0.5 * q)
        return val * fac


class _RadiusOnly(object):
    """Kernels the reference imports but never evaluates on this path."""
    radius_scale = 2.0

    def __init__(self, dim=1):
        self.dim = dim


class WendlandQuintic(_RadiusOnly):
    radius_scale = 2.0


class WendlandQuinticC4(_RadiusOnly):
    radius_scale = 2.0


class Gaussian(_RadiusOnly):
    radius_scale = 3.0


class SuperGaussian(_RadiusOnly):
    radius_scale = 3.0
