"""Make ``import pysph...`` / ``import rigid_body_3d`` resolve to this package.

The reference's Application scripts import PySPH and their sibling modules by
bare name (e.g. /root/reference/code/benchmark_5_steady_cubes_on_a_wall_3d.py:
6-20).  ``install()`` registers the compat layer under those names in
``sys.modules`` so that such a script runs unmodified with the B200 scheme
swapped in:

    python -m rigid_body_2d_3d_pysph_b200.run path/to/benchmark_2_....py [args]

A real PySPH installation, if importable, is never shadowed unless
``force=True``.
"""
import importlib
import sys
import types


def _module(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def install(force=False):
    if not force:
        try:
            importlib.import_module('pysph')
            if not getattr(sys.modules['pysph'], '_rbx_shim', False):
                raise RuntimeError('a real pysph is importable; pass '
                                   'force=True to shadow it')
        except ImportError:
            pass
    from . import (application, equation, geometry, integrator, kernels,
                   output, particle_array, scheme, solver, sph_evaluator,
                   wall_normal)
    from .. import (boundary_particles, rigid_body_2d, rigid_body_3d,
                    rigid_body_common)
    import builtins

    def declare(type, num=1):
        return None

    for p in ['pysph', 'pysph.sph', 'pysph.base', 'pysph.tools',
              'pysph.solver', 'pysph.sph.wc', 'pysph.sph.isph',
              'pysph.examples', 'pysph.examples.solid_mech',
              'pysph.examples.rigid_body', 'compyle']:
        m = _module(p)
        m.__path__ = []
        m._rbx_shim = True
    _module('compyle.api', declare=declare)
    builtins.declare = declare
    _module('pysph.sph.equation', Equation=equation.Equation,
            Group=equation.Group,
            MultiStageEquations=equation.MultiStageEquations)
    _module('pysph.sph.scheme', Scheme=scheme.Scheme,
            SchemeChooser=scheme.SchemeChooser,
            add_bool_argument=scheme.add_bool_argument)
    _module('pysph.sph.integrator_step',
            IntegratorStep=integrator.IntegratorStep)
    _module('pysph.sph.integrator', Integrator=integrator.Integrator,
            EPECIntegrator=integrator.EPECIntegrator)
    _module('pysph.sph.wc.gtvf', GTVFIntegrator=integrator.GTVFIntegrator)
    _module('pysph.tools.sph_evaluator',
            SPHEvaluator=sph_evaluator.SPHEvaluator)
    _module('pysph.base.kernels', **dict(
        (k, getattr(kernels, k)) for k in
        ['CubicSpline', 'WendlandQuintic', 'QuinticSpline',
         'WendlandQuinticC4', 'Gaussian', 'SuperGaussian']))
    _module('pysph.base.utils',
            get_particle_array=particle_array.get_particle_array)
    _module('pysph.base.particle_array',
            ParticleArray=particle_array.ParticleArray)
    _module('pysph.sph.isph.wall_normal',
            ComputeNormals=wall_normal.ComputeNormals,
            SmoothNormals=wall_normal.SmoothNormals)

    def add_properties(pa, *props):
        for prop in props:
            pa.add_property(name=prop)
    _module('pysph.examples.solid_mech.impact', add_properties=add_properties)
    _module('pysph.examples.rigid_body.sphere_in_vessel_akinci',
            create_boundary=None, create_fluid=None, create_sphere=None)
    _module('pysph.tools.geometry', get_2d_block=geometry.get_2d_block,
            get_3d_block=geometry.get_3d_block,
            get_2d_tank=geometry.get_2d_tank,
            remove_overlap_particles=geometry.remove_overlap_particles)
    _module('pysph.solver.solver', Solver=solver.Solver)
    _module('pysph.solver.application', Application=application.Application)
    _module('pysph.solver.utils', iter_output=output.iter_output,
            load=output.load, get_files=output.get_files, dump=output.dump)
    # the reference's own sibling modules, by bare name
    sys.modules['rigid_body_common'] = rigid_body_common
    sys.modules['rigid_body_3d'] = rigid_body_3d
    sys.modules['rigid_body_2d'] = rigid_body_2d
    sys.modules['boundary_particles'] = boundary_particles
    # (`geometry`, the scripts' scene helper, is the script's own sibling: it
    # only needs pysph.tools.geometry, registered above; run.py puts the
    # script's directory on sys.path)
    try:
        importlib.import_module('matplotlib')
    except ImportError:
        class _Plt(types.ModuleType):
            def __getattr__(self, name):
                if name.startswith('__'):
                    raise AttributeError(name)
                return lambda *a, **k: _Plt('x')
        mpl = _module('matplotlib')
        mpl.__path__ = []
        mpl.use = lambda *a, **k: None
        plt = _Plt('matplotlib.pyplot')
        sys.modules['matplotlib.pyplot'] = plt
        mpl.pyplot = plt
