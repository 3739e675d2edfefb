"""TEST INFRASTRUCTURE -- reference-as-oracle harness, part 1: PySPH stub.

PySPH is not installed and cannot be (no network), but the reference's
equations and steppers are valid Python.  This module registers just enough
of the ``pysph`` / ``compyle`` import surface in ``sys.modules`` that
``/root/reference/code/{rigid_body_common,rigid_body_3d,rigid_body_2d,dem,
boundary_particles}.py`` import unmodified, so that
``interp.py`` can execute their ``initialize/loop/post_loop/reduce/py_stage*/
stage*`` methods directly.  It only runs in the build container (the GPU box
has no /root/reference); its outputs are frozen as fixtures in tests/golden/
by ``oracle/make_golden.py``.

Nothing under oracle/ is imported by the product package.
"""
import builtins
import dis
import sys
import types

import numpy as np

from rigid_body_2d_3d_pysph_b200.compat import (equation, geometry, kernels,
                                                particle_array, scheme)

REFERENCE_CODE = '/root/reference/code'


# ----------------------------------------------------------------------
# compyle.api.declare with call-site arity (quirk Q8:
# ``t1, t2 = declare('int', 3)`` at rigid_body_common.py:676)
# ----------------------------------------------------------------------
_arity_cache = {}


def _unpack_arity(frame):
    key = (frame.f_code, frame.f_lasti)
    if key in _arity_cache:
        return _arity_cache[key]
    arity = None
    for ins in dis.get_instructions(frame.f_code):
        if ins.offset <= frame.f_lasti:
            continue
        if ins.opname == 'CACHE':
            continue
        if ins.opname == 'UNPACK_SEQUENCE':
            arity = ins.argval
        break
    _arity_cache[key] = arity
    return arity


def declare(type, num=1):
    if type.startswith('matrix'):
        shape = eval(type[len('matrix'):])
        one = lambda: np.zeros(shape)
    elif type in ('int', 'long', 'unsigned int'):
        one = lambda: 0
    elif type in ('double', 'float'):
        one = lambda: 0.0
    else:
        one = lambda: None
    arity = _unpack_arity(sys._getframe(1))
    if arity is None:
        if num == 1:
            return one()
        return tuple(one() for _ in range(num))
    return tuple(one() for _ in range(arity))


# ----------------------------------------------------------------------
class IntegratorStep(object):
    pass


class Integrator(object):
    def __init__(self, **kw):
        self.steppers = kw


class EPECIntegrator(Integrator):
    pass


class GTVFIntegrator(Integrator):
    pass


class _Solver(object):
    def __init__(self, dim=2, integrator=None, kernel=None, **kw):
        self.dim = dim
        self.integrator = integrator
        self.kernel = kernel
        self.dt = kw.get('dt')
        self.tf = kw.get('tf')
        self.pfreq = kw.get('pfreq', 100)
        self.t = 0.0


# upstream pysph.sph.isph.wall_normal restated as interpretable equations
# (text: /root/reference/code/boundary_particles.py:71-135 with the upstream
# property names normal_tmp / normal)
class ComputeNormals(equation.Equation):
    def initialize(self, d_idx, d_normal_tmp, d_normal):
        idx = 3 * d_idx
        d_normal_tmp[idx] = 0.0
        d_normal_tmp[idx + 1] = 0.0
        d_normal_tmp[idx + 2] = 0.0
        d_normal[idx] = 0.0
        d_normal[idx + 1] = 0.0
        d_normal[idx + 2] = 0.0

    def loop(self, d_idx, d_normal_tmp, s_idx, s_m, s_rho, DWIJ):
        idx = 3 * d_idx
        fac = -s_m[s_idx] / s_rho[s_idx]
        d_normal_tmp[idx] += fac * DWIJ[0]
        d_normal_tmp[idx + 1] += fac * DWIJ[1]
        d_normal_tmp[idx + 2] += fac * DWIJ[2]

    def post_loop(self, d_idx, d_normal_tmp, d_h):
        idx = 3 * d_idx
        mag = np.sqrt(d_normal_tmp[idx]**2 + d_normal_tmp[idx + 1]**2 +
                      d_normal_tmp[idx + 2]**2)
        if mag > 0.25 / d_h[d_idx]:
            d_normal_tmp[idx] /= mag
            d_normal_tmp[idx + 1] /= mag
            d_normal_tmp[idx + 2] /= mag
        else:
            d_normal_tmp[idx] = 0.0
            d_normal_tmp[idx + 1] = 0.0
            d_normal_tmp[idx + 2] = 0.0


class SmoothNormals(equation.Equation):
    def loop(self, d_idx, d_normal, s_normal_tmp, s_idx, s_m, s_rho, WIJ):
        idx = 3 * d_idx
        fac = s_m[s_idx] / s_rho[s_idx] * WIJ
        d_normal[idx] += fac * s_normal_tmp[3 * s_idx]
        d_normal[idx + 1] += fac * s_normal_tmp[3 * s_idx + 1]
        d_normal[idx + 2] += fac * s_normal_tmp[3 * s_idx + 2]

    def post_loop(self, d_idx, d_normal, d_h):
        idx = 3 * d_idx
        mag = np.sqrt(d_normal[idx]**2 + d_normal[idx + 1]**2 +
                      d_normal[idx + 2]**2)
        if mag > 1e-3:
            d_normal[idx] /= mag
            d_normal[idx + 1] /= mag
            d_normal[idx + 2] /= mag
        else:
            d_normal[idx] = 0.0
            d_normal[idx + 1] = 0.0
            d_normal[idx + 2] = 0.0


def add_properties(pa, *props):
    for prop in props:
        pa.add_property(name=prop)


class SPHEvaluator(object):
    """Interpreting evaluator: runs the given groups once (setup time)."""

    def __init__(self, arrays, equations, dim, kernel=None, **kw):
        self.arrays = arrays
        self.equations = equations
        self.dim = dim
        self.kernel = kernel

    def evaluate(self, t=0.0, dt=0.1):
        from . import interp
        interp.run_groups(self.equations, self.arrays, self.kernel, t, dt)


def _module(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


_REF_MODULES = ['rigid_body_common', 'rigid_body_3d', 'rigid_body_2d', 'dem',
                'boundary_particles', 'geometry', 'rigid_fluid_coupling']


def install():
    """Register the stub and put the reference sources on sys.path."""
    for name in list(sys.modules):
        if name == 'pysph' or name.startswith('pysph.') or \
                name.startswith('compyle') or name in _REF_MODULES:
            del sys.modules[name]
    builtins.declare = declare     # rigid_body_common.py:130 uses it bare
    pkgs = ['pysph', 'pysph.sph', 'pysph.base', 'pysph.tools', 'pysph.solver',
            'pysph.sph.wc', 'pysph.sph.isph', 'pysph.examples',
            'pysph.sph.solid_mech',
            'pysph.examples.solid_mech', 'pysph.examples.rigid_body',
            'compyle']
    for p in pkgs:
        m = _module(p)
        m.__path__ = []
    _module('compyle.api', declare=declare)
    _module('pysph.sph.equation', Equation=equation.Equation,
            Group=equation.Group,
            MultiStageEquations=equation.MultiStageEquations)
    _module('pysph.sph.scheme', Scheme=scheme.Scheme,
            SchemeChooser=scheme.SchemeChooser,
            add_bool_argument=scheme.add_bool_argument)
    _module('pysph.sph.integrator_step', IntegratorStep=IntegratorStep)
    _module('pysph.sph.integrator', EPECIntegrator=EPECIntegrator,
            Integrator=Integrator)
    _module('pysph.sph.wc.gtvf', GTVFIntegrator=GTVFIntegrator)
    _module('pysph.tools.sph_evaluator', SPHEvaluator=SPHEvaluator)
    _module('pysph.base.kernels', **dict(
        (k, getattr(kernels, k)) for k in
        ['CubicSpline', 'WendlandQuintic', 'QuinticSpline',
         'WendlandQuinticC4', 'Gaussian', 'SuperGaussian']))
    _module('pysph.base.utils',
            get_particle_array=particle_array.get_particle_array)
    _module('pysph.sph.isph.wall_normal', ComputeNormals=ComputeNormals,
            SmoothNormals=SmoothNormals)
    _module('pysph.examples.solid_mech.impact', add_properties=add_properties)
    _module('pysph.examples.rigid_body.sphere_in_vessel_akinci',
            create_boundary=None, create_fluid=None, create_sphere=None)
    _module('pysph.tools.geometry', get_2d_block=geometry.get_2d_block,
            get_3d_block=geometry.get_3d_block,
            get_2d_tank=geometry.get_2d_tank)
    _module('pysph.solver.solver', Solver=_Solver)
    # imported (unused) inside DEMScheme._get_gtvf_equations, dem.py:699-705
    _module('pysph.sph.basic_equations', ContinuityEquation=None,
            MonaghanArtificialViscosity=None, VelocityGradient3D=None,
            VelocityGradient2D=None)
    _module('pysph.sph.solid_mech.basic', IsothermalEOS=None,
            HookesDeviatoricStressRate=None, MonaghanArtificialStress=None)
    _module('pysph.solver.application', Application=object)
    _module('matplotlib', pyplot=None)
    _module('matplotlib.pyplot')
    if REFERENCE_CODE not in sys.path:
        sys.path.insert(0, REFERENCE_CODE)


def uninstall():
    for name in list(sys.modules):
        if name == 'pysph' or name.startswith('pysph.') or \
                name.startswith('compyle') or name in _REF_MODULES or \
                name.startswith('matplotlib'):
            del sys.modules[name]
    if REFERENCE_CODE in sys.path:
        sys.path.remove(REFERENCE_CODE)
    if hasattr(builtins, 'declare'):
        del builtins.declare
