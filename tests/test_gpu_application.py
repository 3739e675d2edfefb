"""The PySPH-shaped Application/Scheme/Solver surface end to end on the GPU:
hooks in upstream order, CLI flags, post_step with a host write, output dumps
and iter_output -- driven by the frozen stack_of_cylinders scene."""
import numpy as np
import pytest

from oracle import rbo
from tests.util import load_config, oracle_params

pytestmark = pytest.mark.gpu


def _make_app(tmp_path, calls):
    from rigid_body_2d_3d_pysph_b200.compat.application import Application
    from rigid_body_2d_3d_pysph_b200.compat.scheme import SchemeChooser
    from rigid_body_2d_3d_pysph_b200.rigid_body_3d import RigidBody3DScheme

    class Cylinders(Application):
        def initialize(self):
            calls.append('initialize')
            self.wall_time = 15.75 * 5e-5

        def create_scheme(self):
            calls.append('create_scheme')
            rb3d = RigidBody3DScheme(rigid_bodies=['cylinders'],
                                     boundaries=['dam', 'wall'], gx=0.,
                                     gy=-9.81, gz=0., dim=2, fric_coeff=0.45)
            return SchemeChooser(default='rb3d', rb3d=rb3d)

        def configure_scheme(self):
            calls.append('configure_scheme')
            self.scheme.configure_solver(dt=5e-5, tf=40 * 5e-5, pfreq=10)

        def create_particles(self):
            calls.append('create_particles')
            arrays, _ = load_config('stack_of_cylinders')
            return arrays

        def post_step(self, solver):
            t, dt, T = solver.t, solver.dt, self.wall_time
            if (T - dt / 2) < t < (T + dt / 2):
                calls.append('wall moved at step %d' % solver.count)
                for pa in self.particles:
                    if pa.name == 'wall':
                        pa.x += 0.25
    return Cylinders(fname='cyl', output_dir=str(tmp_path))


def test_application_run(tmp_path):
    from rigid_body_2d_3d_pysph_b200.compat.output import iter_output
    calls = []
    app = _make_app(tmp_path, calls)
    app.run(['-d', str(tmp_path), '-q'])
    assert calls[:4] == ['initialize', 'create_scheme', 'configure_scheme',
                         'create_particles']
    assert 'wall moved at step 16' in calls
    # divergence D6: fric_coeff=0.45 from the constructor is overridden by
    # the CLI default 0.5; kf becomes 1e3
    assert app.scheme.scheme.fric_coeff == 0.5 and app.scheme.scheme.kf == 1e3
    files = app.output_files
    assert len(files) == 5                       # steps 0, 10, 20, 30, 40
    ts, xs = [], []
    for sd, cyl in iter_output(files, 'cylinders'):
        ts.append(sd['t'])
        assert int(cyl.nb[0]) == 33
        xs.append(np.mean(cyl.xcm.reshape(-1, 3)[:, 0]))
        for n in ('x', 'y', 'u', 'v', 'fx', 'fy', 'body_id', 'is_boundary'):
            assert n in cyl.properties
    assert np.allclose(ts, np.arange(5) * 10 * 5e-5)

    # same run on the oracle
    oarr, meta = load_config('stack_of_cylinders')
    p = oracle_params(meta)
    rbo.gtvf_step(oarr, meta['rigid'], p, nsteps=16)
    [pa for pa in oarr if pa.name == 'wall'][0].x += 0.25
    rbo.gtvf_step(oarr, meta['rigid'], p, nsteps=24)
    got = [pa for pa in app.particles if pa.name == 'cylinders'][0]
    want = oarr[0]
    assert np.allclose(got.xcm, want.xcm, rtol=0, atol=1e-9)
    assert np.allclose(got.R, want.R, rtol=0, atol=1e-8)
    assert abs(xs[-1] - np.mean(want.xcm.reshape(-1, 3)[:, 0])) < 1e-9


def test_cli_overrides(tmp_path):
    calls = []
    app = _make_app(tmp_path, calls)
    app.run(['-d', str(tmp_path), '-q', '--kr-stiffness', '2e5',
             '--fric-coeff', '0.3', '--max-steps', '7', '--pfreq', '5',
             '--openmp'])
    s = app.scheme.scheme
    assert s.kr == 2e5 and s.fric_coeff == 0.3
    assert app.solver.count == 7
    assert app.solver.scene.kr == 2e5 and app.solver.scene.fric_coeff == 0.3
    assert len(app.output_files) == 3            # steps 0, 5 and the last (7)
