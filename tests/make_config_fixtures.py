"""Freeze the BASELINE.json config scenes as fixtures (build container only).

Runs the REFERENCE's own Application scripts from /root/reference/code,
unmodified, with this package's PySPH-compatible layer installed
(compat.install), up to and including ``create_particles()`` -- i.e. the
scene and every property/constant the reference's setup produces -- and stores
the inputs of the hot path in tests/golden/cfg_<name>.npz.  The GPU box has no
/root/reference, so the config-level parity tests start from these files.

    python -m tests.make_config_fixtures
"""
import os
import runpy
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rigid_body_2d_3d_pysph_b200.compat.install import install  # noqa: E402
from rigid_body_2d_3d_pysph_b200.compat.output import dump  # noqa: E402
from rigid_body_2d_3d_pysph_b200.compat.particle_array import \
    ParticleArray  # noqa: E402

REF = '/root/reference/code/'
CONFIGS = {
    'benchmark_1': ('benchmark_1_rigid_body_rotating_and_traslating_freely.py',
                    'Case0', []),
    'benchmark_1_rb3d': (
        'benchmark_1_rigid_body_rotating_and_traslating_freely.py', 'Case0',
        ['--scheme', 'rb3d']),
    'benchmark_2': (
        'benchmark_2_multiple_rigid_bodies_colliding_same_particle_array.py',
        'RigidFluidCoupling', []),
    'benchmark_5_3d': ('benchmark_5_steady_cubes_on_a_wall_3d.py',
                       'Dinesh2022SteadyCubesOnAWall3D', ['--pyramid-cubes']),
    'stack_of_cylinders': ('stack_of_cylinders.py', 'ZhangStackOfCylinders',
                           []),
    # next rows of SURVEY section 8f-3 (not named by BASELINE.json)
    'benchmark_3': (
        'benchmark_3_multiple_rigid_bodies_colliding_same_particle_array.py',
        'RigidFluidCoupling', []),
    'benchmark_4': ('benchmark_4_rigid_cube_bouncing_on_a_wall.py',
                    'RigidFluidCoupling', ['--coeff-of-restitution', '0.6']),
    'benchmark_5_2d': ('benchmark_5_steady_cubes_on_a_wall_2d.py',
                       'Dinesh2022SteadyCubesOnAWall2D', ['--pyramid-cubes']),
}
KEEP = ['x', 'y', 'z', 'u', 'v', 'w', 'h', 'm', 'rho', 'dem_id', 'body_id',
        'contact_force_is_boundary', 'is_boundary', 'normal', 'normal0',
        'dx0', 'dy0', 'dz0', 'fx', 'fy', 'fz']
DROP_CONST = set()


def main():
    install()
    sys.path.insert(0, REF)     # the scripts' own sibling geometry.py
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    for name, (script, cls, argv) in CONFIGS.items():
        ns = runpy.run_path(REF + script, run_name='fixture')
        app = ns[cls]()
        app._parse(argv)
        app.consume_user_options()
        app.scheme.consume_user_options(app.options)
        app.configure_scheme()
        particles = app.create_particles()
        sch = app.scheme.scheme
        slim = []
        for pa in particles:
            q = ParticleArray(name=pa.name)
            q.__dict__['_n'] = pa.get_number_of_particles()
            for n in KEEP:
                if n in pa.properties:
                    q.add_property(n, type=pa.property_types[n],
                                   data=pa.properties[n],
                                   stride=pa.stride[n])
            for n, v in pa.constants.items():
                q.add_constant(n, v)
            slim.append(q)
        meta = {'rigid': list(sch.rigid_bodies),
                'boundaries': list(sch.boundaries), 'dim': sch.dim,
                'kr': sch.kr, 'kf': sch.kf, 'fric_coeff': sch.fric_coeff,
                'gx': sch.gx, 'gy': sch.gy, 'gz': sch.gz,
                'dt': sch.solver.dt, 'tf': sch.solver.tf,
                'stepper': list(sch.solver.integrator.steppers.values())[0]
                .kind, 'script': script, 'argv': argv}
        f = dump(os.path.join(out_dir, 'cfg_' + name), slim, meta,
                 detailed_output=True, compress=True)
        print(name, [(p.name, p.get_number_of_particles()) for p in slim],
              os.path.getsize(f) // 1024, 'KiB')


if __name__ == '__main__':
    main()
