"""``pysph.sph.integrator`` / ``integrator_step`` surface and the planner that
maps a scheme's equations + steppers onto the CUDA step.

[upstream, restated] SURVEY.md App. C-5/6:
  GTVFIntegrator.one_timestep = stage1, eval(0) (empty), stage2, eval(1),
  stage3 with no initial acceleration; EPECIntegrator.one_timestep =
  initialize, eval, stage1, eval, stage2.
Both are executed by DeviceScene (device.py) as fused launches; this module
only decides WHICH device program a (steppers, equations) pair means, and
refuses anything it does not know (no interpreted fallback).
"""


class IntegratorStep(object):
    """Base of the stepper descriptors (rigid_body_3d.GTVFRigidBody3DStep,
    rigid_body_2d.GTVFRigidBody2DStep, rigid_body_3d.RK2RigidBody3DStep)."""
    kind = None


class Integrator(object):
    sequencing = None

    def __init__(self, **kw):
        self.steppers = kw
        self.scene = None
        # reference behaviour (quirk Q7, rigid_body_3d.py:415) unless the
        # stepper asks for the repaired save: one default on every entry point
        self.fix_q7 = False

    def set_scene(self, scene):
        self.scene = scene

    def step(self, t, dt, nsteps=1, graph=False):
        raise NotImplementedError()


class GTVFIntegrator(Integrator):
    sequencing = 'gtvf'

    def step(self, t, dt, nsteps=1, graph=False):
        self.scene.gtvf_step(dt, nsteps, graph=graph)


class EPECIntegrator(Integrator):
    sequencing = 'epec'

    def step(self, t, dt, nsteps=1, graph=False):
        self.scene.rk2_step(dt, nsteps, fix_q7=self.fix_q7)


_KNOWN = ('ComputeContactForceNormals',
          'ComputeContactForceDistanceAndClosestPoint', 'BodyForce',
          'ComputeContactForce', 'SumUpExternalForces')


class StepPlan(object):
    def __init__(self):
        self.rigid = []
        self.boundaries = []
        self.kr, self.kf, self.fric_coeff = 1e5, 1e3, 0.5
        self.gx = self.gy = self.gz = 0.0
        self.planar = False
        self.sequencing = 'gtvf'


_KNOWN_DEM = ('UpdateTangentialContactsLVCDisplacement', 'BodyForce',
              'LVCDisplacement')


def _plan_dem(seen, integrator, plan):
    """DEMScheme's equation list (dem.py:697-756)."""
    for need in _KNOWN_DEM:
        if need not in seen:
            raise NotImplementedError('DEM path needs %s' % need)
    plan.kind = 'dem'
    plan.rigid = [eq.dest for eq in seen['LVCDisplacement']]
    src = []
    for eq in seen['LVCDisplacement']:
        for s in (eq.sources or []):
            if s not in src:
                src.append(s)
    plan.boundaries = [s for s in src if s not in plan.rigid]
    bf = seen['BodyForce'][0]
    plan.gx, plan.gy, plan.gz = bf.gx, bf.gy, bf.gz
    for name in plan.rigid:
        st = integrator.steppers.get(name)
        if st is None or st.kind != 'dem':
            raise NotImplementedError('array %s needs a DEMStep' % name)
    if integrator.sequencing != 'gtvf':
        raise NotImplementedError('DEMStep runs under GTVF sequencing')
    return plan


def plan_from_equations(equations, integrator):
    """Validate the equation list against the programs the device runs
    (rigid_body_3d.py:641-698, dem.py:697-756) and extract the parameters."""
    from .equation import MultiStageEquations
    plan = StepPlan()
    plan.kind = 'rigid'
    if isinstance(equations, MultiStageEquations):
        groups = []
        for stage in equations.groups:
            groups.extend(stage)
    else:
        groups = list(equations)
    names = set(eq.__class__.__name__ for g in groups for eq in g.equations)
    if 'LVCDisplacement' in names or \
            'UpdateTangentialContactsLVCDisplacement' in names:
        seen = {}
        for g in groups:
            for eq in g.equations:
                name = eq.__class__.__name__
                if name not in _KNOWN_DEM:
                    raise NotImplementedError(
                        'equation %s has no CUDA implementation on the DEM '
                        'path' % name)
                seen.setdefault(name, []).append(eq)
        return _plan_dem(seen, integrator, plan)
    seen = {}
    for g in groups:
        for eq in g.equations:
            name = eq.__class__.__name__
            if name not in _KNOWN:
                raise NotImplementedError(
                    'equation %s has no CUDA implementation on this path '
                    '(known: %s)' % (name, ', '.join(_KNOWN)))
            seen.setdefault(name, []).append(eq)
    if 'ComputeContactForce' not in seen:
        raise NotImplementedError('equation list has no ComputeContactForce')
    for need in _KNOWN:
        if need not in seen:
            raise NotImplementedError(
                'the fused contact kernel needs the full group list of '
                'rigid_body_3d.py:641-698; %s is missing' % need)
    plan.rigid = [eq.dest for eq in seen['ComputeContactForce']]
    src = []
    for eq in seen['ComputeContactForceNormals']:
        for s in (eq.sources or []):
            if s not in src:
                src.append(s)
    plan.boundaries = [s for s in src if s not in plan.rigid]
    cf = seen['ComputeContactForce'][0]
    plan.kr, plan.kf, plan.fric_coeff = cf.kr, cf.kf, cf.fric_coeff
    bf = seen['BodyForce'][0]
    plan.gx, plan.gy, plan.gz = bf.gx, bf.gy, bf.gz
    kinds = set()
    for name in plan.rigid:
        st = integrator.steppers.get(name)
        if st is None or st.kind is None:
            raise NotImplementedError('no CUDA stepper for array %s' % name)
        kinds.add(st.kind)
    if len(kinds) != 1:
        raise NotImplementedError('mixed steppers: %s' % sorted(kinds))
    kind = kinds.pop()
    plan.planar = (kind == 'gtvf2d')
    plan.sequencing = integrator.sequencing
    if (kind == 'rk2') != (plan.sequencing == 'epec'):
        raise NotImplementedError(
            'stepper %s cannot run under %s sequencing' %
            (kind, plan.sequencing))
    return plan
