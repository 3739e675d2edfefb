"""``RigidBody2DScheme`` -- same surface as the reference's
``code/rigid_body_2d.py``.  It differs from the 3-D scheme in the stepper
(planar: only x, y of vcm/xcm, omega_z from izz, rigid_body_2d.py:40-205) and
in the setup (``set_moment_of_inertia_izz`` instead of the tensor,
rigid_body_2d.py:486-506; the inertia tensors stay zero, so
``set_angular_velocity`` leaves ``ang_mom`` at zero -- quirk Q12).
"""
from .compat.integrator import IntegratorStep
from .rigid_body_3d import RigidBody3DScheme
from .rigid_body_common import set_moment_of_inertia_izz


class GTVFRigidBody2DStep(IntegratorStep):
    """rigid_body_2d.py:40-205"""
    kind = 'gtvf2d'


class RigidBody2DScheme(RigidBody3DScheme):
    _stepper_cls = GTVFRigidBody2DStep

    def __init__(self, rigid_bodies, boundaries, dim, kr=1e5, kf=1e5, en=0.5,
                 fric_coeff=0.5, gx=0.0, gy=0.0, gz=0.0):
        super(RigidBody2DScheme, self).__init__(
            rigid_bodies, boundaries, dim, kr=kr, kf=kf, en=en,
            fric_coeff=fric_coeff, gx=gx, gy=gy, gz=gz)
        if self.dim != 2:
            print("#============The current scheme cannot be used to "
                  "================#")
            print("#============simulate problems other than 2 dimensions"
                  "============#")

    _inertia_tensor = False

    def _set_inertia(self, pa):
        set_moment_of_inertia_izz(pa)
