"""``pysph.solver.solver.Solver`` surface driving the device scene.

[upstream, restated] the loop ``while t < tf: integrator.step(t, dt); t += dt;
post-step callbacks; dump every pfreq``.  Between
observation points (callbacks, dumps) the steps run back to back on the GPU,
small scenes through a CUDA graph; host arrays are refreshed lazily.
"""
import os

from . import output as _output
from .integrator import plan_from_equations


class Solver(object):
    def __init__(self, dim=2, integrator=None, kernel=None, n_damp=0, tf=1.0,
                 dt=1e-3, adaptive_timestep=False, cfl=0.3, output_at_times=(),
                 fixed_h=False, pfreq=100, **kwargs):
        self.dim = dim
        self.integrator = integrator
        self.kernel = kernel
        self.tf = tf
        self.dt = dt
        self.pfreq = pfreq
        self.t = 0.0
        self.count = 0
        self.max_steps = None
        self.disable_output = False
        self.output_directory = '.'
        self.fname = 'output'
        self.detailed_output = False
        self.post_step_callbacks = []
        self.pre_step_callbacks = []
        self.particles = None
        self.scene = None
        self.output_files = []
        # capacities / tuning of the device scene (device.py)
        self.ks = kwargs.pop('ks', 8)
        self.list_cap = kwargs.pop('list_cap', 96)
        self.skin_factor = kwargs.pop('skin_factor', 0.075)
        self.use_graph = kwargs.pop('use_graph', True)
        # one damping coefficient for every (body, body) pair instead of the
        # reference's dense nb x tnb `eta` table (large scenes)
        self.eta_uniform = kwargs.pop('eta_uniform', None)
        self.extra = kwargs

    # -- configuration hooks used by Application ---------------------------
    def set_final_time(self, tf):
        self.tf = tf

    def set_time_step(self, dt):
        self.dt = dt

    def set_print_freq(self, n):
        self.pfreq = n

    def set_max_steps(self, n):
        self.max_steps = n

    def set_disable_output(self, value=True):
        self.disable_output = value

    def set_output_directory(self, path):
        self.output_directory = path

    def set_output_fname(self, fname):
        self.fname = fname

    def add_post_step_callback(self, cb):
        self.post_step_callbacks.append(cb)

    def add_pre_step_callback(self, cb):
        self.pre_step_callbacks.append(cb)

    # ----------------------------------------------------------------------
    def setup(self, particles, equations, nnps=None, kernel=None,
              fixed_h=False):
        from ..device import DeviceScene
        self.particles = particles
        plan = plan_from_equations(equations, self.integrator)
        radius_scale = getattr(self.kernel, 'radius_scale', 3.0)
        self.plan = plan
        if plan.kind != 'dem':
            # The contact passes weight with WIJ (rigid_body_common.py:664,
            # 793) and the device path evaluates the quintic spline only;
            # any other kernel would silently run as a truncated quintic.
            from .kernels import QuinticSpline
            if self.kernel is not None and not (
                    isinstance(self.kernel, QuinticSpline) and
                    radius_scale == 3.0):
                raise NotImplementedError(
                    'the rigid-body contact path evaluates QuinticSpline '
                    '(radius_scale 3) on the device; kernel %s is not '
                    'mapped' % type(self.kernel).__name__)
        if plan.kind == 'dem':
            from ..dem import DemDeviceScene
            self.scene = DemDeviceScene(
                particles, plan.rigid, plan.boundaries, dim=self.dim,
                gx=plan.gx, gy=plan.gy, gz=plan.gz, radius_scale=radius_scale)
            self.integrator.set_scene(self.scene)
            return
        self.scene = DeviceScene(
            particles, plan.rigid, plan.boundaries, dim=self.dim, kr=plan.kr,
            kf=plan.kf, fric_coeff=plan.fric_coeff, gx=plan.gx, gy=plan.gy,
            gz=plan.gz, planar=plan.planar, ks=self.ks,
            list_cap=self.list_cap, skin_factor=self.skin_factor,
            radius_scale=radius_scale, eta_uniform=self.eta_uniform)
        self.integrator.set_scene(self.scene)
        self.plan = plan

    def dump_output(self):
        if self.disable_output:
            return
        self.scene.sync_to_host()
        os.makedirs(self.output_directory, exist_ok=True)
        fname = os.path.join(self.output_directory,
                             '%s_%d' % (self.fname, self.count))
        f = _output.dump(fname, self.particles,
                         {'t': self.t, 'dt': self.dt, 'count': self.count},
                         detailed_output=self.detailed_output)
        self.output_files.append(f)

    def _steps_left(self):
        # number of whole dt steps to tf (the last one may be shortened)
        return max(0, int((self.tf - self.t) / self.dt + 1e-9))

    def solve(self, show_progress=False):
        if self.count == 0:
            self.dump_output()
        observed = bool(self.post_step_callbacks or self.pre_step_callbacks)
        eps = 1e-12 * max(abs(self.tf), 1.0)
        while self.t < self.tf - eps:
            if self.max_steps is not None and self.count >= self.max_steps:
                break
            if observed:
                batch = 1
            else:
                to_dump = self.pfreq - (self.count % self.pfreq)
                batch = max(1, min(to_dump, self._steps_left()))
                if self.max_steps is not None:
                    batch = min(batch, self.max_steps - self.count)
            dt = self.dt
            if batch == 1 and self.t + dt > self.tf:
                dt = self.tf - self.t
            for cb in self.pre_step_callbacks:
                cb(self)
            self.integrator.step(self.t, dt, batch,
                                 graph=self.use_graph and batch >= 8)
            self.t += dt * batch
            self.count += batch
            for cb in self.post_step_callbacks:
                cb(self)
            if self.count % self.pfreq == 0:
                self.scene.check_status()
                self.dump_output()
        self.scene.check_status()
        if self.count % self.pfreq != 0:
            self.dump_output()
