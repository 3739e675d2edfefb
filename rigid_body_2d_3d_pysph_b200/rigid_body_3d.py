"""``RigidBody3DScheme`` and its steppers -- same surface as the reference's
``code/rigid_body_3d.py`` (constructor, CLI flags, get_equations,
configure_solver, setup_properties, set_linear_velocity,
set_angular_velocity), B200-native underneath.

The steppers are descriptors: their arithmetic lives in csrc/rbx_bodies.cu
(rbx_gtvf_kick / rbx_gtvf_drift / rbx_pose_particles / rbx_rk2_stage).
"""
import numpy as np

from .boundary_particles import (add_boundary_identification_properties,
                                 get_boundary_identification_etvf_equations)
from .compat.equation import Group, MultiStageEquations
from .compat.integrator import (EPECIntegrator, GTVFIntegrator,
                                IntegratorStep)
from .compat.kernels import QuinticSpline
from .compat.scheme import Scheme
from .compat.sph_evaluator import SPHEvaluator
from .rigid_body_common import (BodyForce, ComputeContactForce,
                                ComputeContactForceDistanceAndClosestPoint,
                                ComputeContactForceNormals,
                                SumUpExternalForces, add_properties_stride,
                                set_body_frame_normal_vectors,
                                set_body_frame_position_vectors,
                                set_center_of_mass,
                                set_moment_of_inertia_and_its_inverse,
                                set_total_mass)


class GTVFRigidBody3DStep(IntegratorStep):
    """rigid_body_3d.py:40-225 (kick-drift-kick, one force evaluation)."""
    kind = 'gtvf3d'


class RK2RigidBody3DStep(IntegratorStep):
    """rigid_body_3d.py:406-575, run under EPEC sequencing.  The reference
    saves only the angular momentum of the first body of an array
    (``ang_mom0[j] = ang_mom[j]``, :415, quirk Q7) and that is the default
    here too (as in ``DeviceScene.rk2_step``); ``fix_q7=True`` saves every
    body's."""
    kind = 'rk2'

    def __init__(self, fix_q7=False):
        self.fix_q7 = fix_q7


class LeapFrogRigidBody3DStep(RK2RigidBody3DStep):
    """rigid_body_3d.py:228-403: byte-identical to the RK2 stepper in the
    reference ("FIXME: not implemented yet")."""


# dense (particle, body) slot arrays of the reference (rigid_body_3d.py:
# 739-771).  They exist on the host array for scripts/viewers that name
# them; the device keeps the sparse equivalent (DESIGN.md "sparse slots").
SLOT_PROPS = ('contact_force_normal_x', 'contact_force_normal_y',
              'contact_force_normal_z', 'contact_force_normal_wij',
              'contact_force_normal_tmp_x', 'contact_force_normal_tmp_y',
              'contact_force_normal_tmp_z', 'contact_force_dist_tmp',
              'contact_force_dist', 'overlap', 'ft_x', 'ft_y', 'ft_z',
              'fn_x', 'fn_y', 'fn_z', 'delta_lt_x', 'delta_lt_y',
              'delta_lt_z', 'vx_source', 'vy_source', 'vz_source',
              'x_source', 'y_source', 'z_source', 'ti_x', 'ti_y', 'ti_z',
              'closest_point_dist_to_source')

# above this many dense slot doubles per array the 30 strided arrays are not
# materialised on the host (10 M particles x 100 k bodies would be 2.4e14 B)
DENSE_SLOT_LIMIT = 1 << 24


class RigidBody3DScheme(Scheme):
    _stepper_cls = GTVFRigidBody3DStep
    _planar_inertia = False

    def __init__(self, rigid_bodies, boundaries, dim, kr=1e5, kf=1e5, en=0.5,
                 fric_coeff=0.5, gx=0.0, gy=0.0, gz=0.0):
        self.boundaries = [] if boundaries is None else boundaries
        self.rigid_bodies = [] if rigid_bodies is None else rigid_bodies
        self.dim = dim
        self.kernel = QuinticSpline
        self.integrator = "gtvf"
        self.gx = gx
        self.gy = gy
        self.gz = gz
        self.kr = kr
        self.kf = kf
        self.fric_coeff = fric_coeff
        self.solver = None

    def add_user_options(self, group):
        group.add_argument("--kr-stiffness", action="store", dest="kr",
                           default=1e5, type=float,
                           help="Repulsive spring stiffness")
        group.add_argument("--kf-stiffness", action="store", dest="kf",
                           default=1e3, type=float,
                           help="Tangential spring stiffness")
        group.add_argument("--fric-coeff", action="store", dest="fric_coeff",
                           default=0.5, type=float,
                           help="Friction coefficient")

    def consume_user_options(self, options):
        _vars = ['kr', 'kf', 'fric_coeff']
        data = dict((var, self._smart_getattr(options, var)) for var in _vars)
        self.configure(**data)

    def get_equations(self):
        return self._get_gtvf_equations()

    def _get_gtvf_equations(self):
        """Same groups, same order as rigid_body_3d.py:641-698."""
        stage1 = []
        stage2 = []
        if len(self.rigid_bodies) > 0:
            srcs = self.rigid_bodies + self.boundaries
            stage2.append(Group(equations=[
                ComputeContactForceNormals(dest=name, sources=srcs)
                for name in self.rigid_bodies], real=False))
            stage2.append(Group(equations=[
                ComputeContactForceDistanceAndClosestPoint(dest=name,
                                                           sources=srcs)
                for name in self.rigid_bodies], real=False))
            stage2.append(Group(equations=[
                BodyForce(dest=name, sources=None, gx=self.gx, gy=self.gy,
                          gz=self.gz)
                for name in self.rigid_bodies], real=False))
            stage2.append(Group(equations=[
                ComputeContactForce(dest=name, sources=None, kr=self.kr,
                                    kf=self.kf, fric_coeff=self.fric_coeff)
                for name in self.rigid_bodies], real=False))
            stage2.append(Group(equations=[
                SumUpExternalForces(dest=name, sources=None)
                for name in self.rigid_bodies], real=False))
        return MultiStageEquations([stage1, stage2])

    def configure_solver(self, kernel=None, integrator_cls=None,
                         extra_steppers=None, **kw):
        from .compat.solver import Solver
        if kernel is None:
            kernel = QuinticSpline(dim=self.dim)
        steppers = {}
        if extra_steppers is not None:
            steppers.update(extra_steppers)
        # the reference hard-wires GTVF (rigid_body_3d.py:714-715); an
        # explicit EPECIntegrator selects the RK2 stepper instead
        if integrator_cls is EPECIntegrator:
            bodystep = RK2RigidBody3DStep()
        else:
            bodystep = self._stepper_cls()
            integrator_cls = GTVFIntegrator
        for body in self.rigid_bodies:
            if body not in steppers:
                steppers[body] = bodystep
        integrator = integrator_cls(**steppers)
        if integrator_cls is EPECIntegrator:
            integrator.fix_q7 = bodystep.fix_q7
        self.solver = Solver(dim=self.dim, integrator=integrator,
                             kernel=kernel, **kw)

    def setup_properties(self, particles, clean=True):
        """rigid_body_3d.py:729-903"""
        pas = dict([(p.name, p) for p in particles])
        for rigid_body in self.rigid_bodies:
            self.setup_rigid_array(pas[rigid_body])
        for boundary in self.boundaries:
            self._identify_boundary(pas[boundary])

    def setup_rigid_array(self, pa, is_boundary=None, normal=None):
        """Per-array part of setup_properties (rigid_body_3d.py:734-886).
        ``is_boundary`` / ``normal`` given: skip the SPH boundary
        identification (scene generators that tile a template body)."""
        tnb = int(pa.total_no_bodies[0])
        n = pa.get_number_of_particles()
        if n * tnb <= DENSE_SLOT_LIMIT:
            add_properties_stride(pa, tnb, *SLOT_PROPS)
            pa.add_property(name='dem_id_source', stride=tnb, type='int')
        for prop in ('fx', 'fy', 'fz', 'dx0', 'dy0', 'dz0', 'rho_fsi',
                     'm_fsi', 'p_fsi'):
            pa.add_property(name=prop)
        nb = int(np.max(pa.body_id) + 1)
        eye = np.tile([1., 0., 0., 0., 1., 0., 0., 0., 1.], nb)
        consts = {
            'total_mass': np.zeros(nb), 'xcm': np.zeros(3 * nb),
            'xcm0': np.zeros(3 * nb), 'R': eye, 'R0': eye,
            'izz': np.zeros(nb),
            'inertia_tensor_body_frame': np.zeros(9 * nb),
            'inertia_tensor_inverse_body_frame': np.zeros(9 * nb),
            'inertia_tensor_global_frame': np.zeros(9 * nb),
            'inertia_tensor_inverse_global_frame': np.zeros(9 * nb),
            'force': np.zeros(3 * nb), 'torque': np.zeros(3 * nb),
            'vcm': np.zeros(3 * nb), 'vcm0': np.zeros(3 * nb),
            'ang_mom': np.zeros(3 * nb), 'ang_mom0': np.zeros(3 * nb),
            'omega': np.zeros(3 * nb), 'omega0': np.zeros(3 * nb),
            'nb': nb}
        for key, elem in consts.items():
            pa.add_constant(key, elem)
        pa.add_constant('min_dem_id', int(np.min(pa.dem_id)))
        pa.add_constant('max_dem_id', int(np.max(pa.dem_id)))
        if nb * tnb <= DENSE_SLOT_LIMIT:
            pa.add_constant('eta', np.zeros(nb * tnb))
        if self.device_setup:
            from .setup_device import setup_rigid_bodies
            setup_rigid_bodies(pa, tensor=self._inertia_tensor)
        else:
            set_total_mass(pa)
            set_center_of_mass(pa)
            self._set_inertia(pa)
            set_body_frame_position_vectors(pa)
        if is_boundary is None:
            self._identify_boundary(pa)
        else:
            add_boundary_identification_properties(pa)
            pa.is_boundary[:] = is_boundary
            if normal is not None:
                pa.normal[:] = normal
        set_body_frame_normal_vectors(pa)
        pa.set_output_arrays(['x', 'y', 'z', 'u', 'v', 'w', 'fx', 'fy',
                              'normal', 'is_boundary', 'fz', 'm',
                              'body_id', 'h'])

    _inertia_tensor = True      # (the 2-D scheme sets izz only)

    def _set_inertia(self, pa):
        set_moment_of_inertia_and_its_inverse(pa)

    # device_setup = True routes the per-body setup (mass, centre of mass,
    # inertia, body-frame vectors) and the boundary identification through
    # setup_device.py (CUDA) instead of the host helpers
    device_setup = False

    def _identify_boundary(self, pa):
        if self.device_setup:
            from .setup_device import identify_boundary
            identify_boundary(pa, self.dim)
            return
        add_boundary_identification_properties(pa)
        equations = get_boundary_identification_etvf_equations([pa.name],
                                                               [pa.name])
        sph_eval = SPHEvaluator(arrays=[pa], equations=equations,
                                dim=self.dim,
                                kernel=QuinticSpline(dim=self.dim))
        sph_eval.evaluate(dt=0.1)

    def _set_particle_velocities(self, pa):
        """rigid_body_3d.py:905-926"""
        bid = pa.body_id
        R = pa.R.reshape(-1, 9)[bid]
        om = pa.omega.reshape(-1, 3)[bid]
        vcm = pa.vcm.reshape(-1, 3)[bid]
        dx = (R[:, 0] * pa.dx0 + R[:, 1] * pa.dy0 + R[:, 2] * pa.dz0)
        dy = (R[:, 3] * pa.dx0 + R[:, 4] * pa.dy0 + R[:, 5] * pa.dz0)
        dz = (R[:, 6] * pa.dx0 + R[:, 7] * pa.dy0 + R[:, 8] * pa.dz0)
        pa.u[:] = vcm[:, 0] + (om[:, 1] * dz - om[:, 2] * dy)
        pa.v[:] = vcm[:, 1] + (om[:, 2] * dx - om[:, 0] * dz)
        pa.w[:] = vcm[:, 2] + (om[:, 0] * dy - om[:, 1] * dx)

    def set_linear_velocity(self, pa, linear_vel):
        """rigid_body_3d.py:928-931 (3*nb values; quirk Q11)."""
        pa.vcm[:] = linear_vel
        self._set_particle_velocities(pa)

    def set_angular_velocity(self, pa, angular_vel):
        """rigid_body_3d.py:933-946: omega, then ang_mom = I_g omega."""
        pa.omega[:] = angular_vel[:]
        nb = int(max(pa.body_id)) + 1
        I = pa.inertia_tensor_global_frame.reshape(-1, 3, 3)[:nb]
        om = pa.omega.reshape(-1, 3)[:nb]
        pa.ang_mom[:3 * nb] = np.einsum('bij,bj->bi', I, om).ravel()
        self._set_particle_velocities(pa)

    def get_solver(self):
        return self.solver
