"""Rigid-body equations and setup helpers -- same names as the reference's
``code/rigid_body_common.py``, B200-native underneath.

* The ``Equation`` classes are descriptors with the reference's constructor
  signatures.  They carry no per-particle Python: the integrator maps the
  group list onto the CUDA contact kernel (csrc/rbx_contact.cu) and the body
  reduction (csrc/rbx_bodies.cu).
* The setup helpers (one-off, host side) are vectorised
  NumPy with the reference's results: the reference's versions are Python
  double loops (rigid_body_common.py:46-107) that take minutes at 10^7
  particles.
"""
from math import log, pi

import numpy as np

from .compat.equation import Equation

M_PI = pi


def add_properties_stride(pa, stride=1, *props):
    """rigid_body_common.py:16-18"""
    for prop in props:
        pa.add_property(name=prop, stride=stride)


def _nb(pa):
    return int(max(pa.body_id)) + 1


def set_total_mass(pa):
    """rigid_body_common.py:21-26"""
    nb = _nb(pa)
    pa.total_mass[:nb] = np.bincount(pa.body_id, weights=pa.m, minlength=nb)
    assert np.all(pa.total_mass[:nb] > 0.), \
        "Total mass has to be greater than zero"


def set_center_of_mass(pa):
    """rigid_body_common.py:29-35"""
    nb = _nb(pa)
    bid = pa.body_id
    xcm = pa.xcm.reshape(-1, 3)
    for k, c in enumerate((pa.x, pa.y, pa.z)):
        xcm[:nb, k] = np.bincount(bid, weights=pa.m * c, minlength=nb) / \
            pa.total_mass[:nb]


def set_moment_of_inertia_izz(pa):
    """rigid_body_common.py:38-43"""
    nb = _nb(pa)
    bid = pa.body_id
    xcm = pa.xcm.reshape(-1, 3)
    dx = pa.x - xcm[bid, 0]
    dy = pa.y - xcm[bid, 1]
    pa.izz[:nb] = np.bincount(bid, weights=pa.m * (dx**2. + dy**2.),
                              minlength=nb)


def set_moment_of_inertia_and_its_inverse(pa):
    """rigid_body_common.py:46-94: inertia tensor about the centre of mass,
    its inverse (np.linalg.inv per body), body frame = global frame at t=0."""
    nb = int(pa.nb[0])
    bid = pa.body_id
    xcm = pa.xcm.reshape(-1, 3)
    dx = pa.x - xcm[bid, 0]
    dy = pa.y - xcm[bid, 1]
    dz = pa.z - xcm[bid, 2]
    m = pa.m

    def acc(w):
        return np.bincount(bid, weights=w, minlength=nb)
    I = np.zeros((nb, 9))
    I[:, 0] = acc(m * (dy**2. + dz**2.))
    I[:, 4] = acc(m * (dx**2. + dz**2.))
    I[:, 8] = acc(m * (dx**2. + dy**2.))
    I[:, 1] = -acc(m * dx * dy)
    I[:, 2] = -acc(m * dx * dz)
    I[:, 5] = -acc(m * dy * dz)
    I[:, 3] = I[:, 1]
    I[:, 6] = I[:, 2]
    I[:, 7] = I[:, 5]
    pa.inertia_tensor_body_frame[:] = I.ravel()
    with np.errstate(all='ignore'):
        try:
            Iinv = np.linalg.inv(I.reshape(nb, 3, 3)).reshape(nb, 9)
        except np.linalg.LinAlgError:
            # planar bodies have a singular tensor: invert body by body so
            # one singular body fails the way the reference does
            Iinv = np.stack([np.linalg.inv(b.reshape(3, 3)).ravel()
                             for b in I])
    pa.inertia_tensor_inverse_body_frame[:] = Iinv.ravel()
    pa.inertia_tensor_global_frame[:] = I.ravel()
    pa.inertia_tensor_inverse_global_frame[:] = Iinv.ravel()


def set_body_frame_position_vectors(pa):
    """rigid_body_common.py:97-107"""
    bid = pa.body_id
    xcm = pa.xcm.reshape(-1, 3)
    pa.dx0[:] = pa.x - xcm[bid, 0]
    pa.dy0[:] = pa.y - xcm[bid, 1]
    pa.dz0[:] = pa.z - xcm[bid, 2]


def set_body_frame_normal_vectors(pa):
    """rigid_body_common.py:110-112"""
    pa.normal0[:] = pa.normal[:]


def normalize_R_orientation(orien):
    """rigid_body_common.py:178-203 (classical Gram-Schmidt on columns)."""
    a1 = np.array([orien[0], orien[3], orien[6]])
    a2 = np.array([orien[1], orien[4], orien[7]])
    a3 = np.array([orien[2], orien[5], orien[8]])
    b1 = a1 / np.linalg.norm(a1)
    b2 = a2 - np.dot(b1, a2) * b1
    b2 = b2 / np.linalg.norm(b2)
    b3 = a3 - np.dot(b1, a3) * b1 - np.dot(b2, a3) * b2
    b3 = b3 / np.linalg.norm(b3)
    orien[0], orien[3], orien[6] = b1
    orien[1], orien[4], orien[7] = b2
    orien[2], orien[5], orien[8] = b3


def setup_damping_coefficient(body, rigid_bodies, boundaries=[]):
    """rigid_body_common.py:206-241: eta[i*tnb + k] = -2 ln e / sqrt(ln^2 e +
    pi^2) for every (body i, source dem_id k); the mass factor lives in the
    contact kernel (quirk Q10)."""
    no_bodies_dest = int(max(body.body_id)) + 1
    tnb = int(body.total_no_bodies[0])

    def eta_of(e):
        t1 = log(e)
        t2 = t1**2. + M_PI**2.
        return -2. * t1 * (1. / t2)**0.5
    for i in range(no_bodies_dest):
        idx = i * tnb
        for src in rigid_bodies:
            l1 = int(src.min_dem_id[0])
            l2 = int(src.max_dem_id[0]) + 1
            for j, k in zip(range(int(max(src.body_id)) + 1), range(l1, l2)):
                body.eta[idx + k] = eta_of(body.coeff_of_rest[idx + k])
        for src in boundaries:
            dem_id = int(src.dem_id[0])
            body.eta[idx + dem_id] = eta_of(body.coeff_of_rest[idx + dem_id])
    body._touch('eta')


# ----------------------------------------------------------------------
# equation descriptors (constructor signatures of the reference)
# ----------------------------------------------------------------------
class BodyForce(Equation):
    """rigid_body_common.py:115-125 -> fused into the contact kernel."""

    def __init__(self, dest, sources, gx=0.0, gy=0.0, gz=0.0):
        self.gx = gx
        self.gy = gy
        self.gz = gz
        super(BodyForce, self).__init__(dest, sources)


class SumUpExternalForces(Equation):
    """rigid_body_common.py:128-175 -> chunk partials + rbx_reduce_bodies."""


class ComputeContactForceNormals(Equation):
    """rigid_body_common.py:631-723 -> rbx_contact_mofidi (pass 1)."""


class ComputeContactForceDistanceAndClosestPoint(Equation):
    """rigid_body_common.py:726-836 -> rbx_contact_mofidi (pass 2)."""


class ComputeContactForce(Equation):
    """rigid_body_common.py:839-1032 -> rbx_contact_mofidi (force law)."""

    def __init__(self, dest, sources, fric_coeff=0.5, kr=1e5, kf=1e3):
        self.kr = kr
        self.kf = kf
        self.fric_coeff = fric_coeff
        super(ComputeContactForce, self).__init__(dest, sources)


class RigidBodyCanelasRigidRigid(Equation):
    """rigid_body_common.py:244-442 -> rbx_contact_canelas
    (``DeviceScene.contact_canelas``).  No scheme of the reference wires the
    Canelas equations; the planner does not accept them in a group list."""

    def __init__(self, dest, sources, Cn=1.4 * 1e-5):
        self.Cn = Cn
        super(RigidBodyCanelasRigidRigid, self).__init__(dest, sources)


class RigidBodyCanelasRigidWall(Equation):
    """rigid_body_common.py:445-628 -> rbx_contact_canelas."""

    def __init__(self, dest, sources, Cn=1.4 * 1e-5):
        self.Cn = Cn
        super(RigidBodyCanelasRigidWall, self).__init__(dest, sources)
