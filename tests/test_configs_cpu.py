"""BASELINE.json configs on the CPU side: the frozen scenes have the sizes the
reference's scripts produce, the oracle reproduces the reference-owned known
answers on them, and the product's host-side setup equals the reference's."""
import numpy as np
import pytest

from oracle import rbo
from tests.util import load_case, load_config, oracle_params


def test_config_sizes():
    # SURVEY.md App. D
    a, m = load_config('benchmark_1')
    assert [p.get_number_of_particles() for p in a] == [121]
    a, m = load_config('benchmark_2')
    assert [p.get_number_of_particles() for p in a] == [81, 81]
    assert abs(m['dt'] - 1.6675589e-4) < 1e-10
    a, m = load_config('benchmark_5_3d')
    assert [p.get_number_of_particles() for p in a] == [750, 30225]
    assert int(a[0].nb[0]) == 6 and int(a[0].total_no_bodies[0]) == 7
    a, m = load_config('stack_of_cylinders')
    assert a[0].get_number_of_particles() == 2541
    assert int(a[0].nb[0]) == 33 and int(a[0].total_no_bodies[0]) == 35
    a, m = load_config('benchmark_4')
    assert [p.get_number_of_particles() for p in a] == [162, 855]
    assert m['stepper'] == 'gtvf2d'           # SchemeChooser default rb2d
    assert np.count_nonzero(a[0].eta) == 6    # --coeff-of-restitution 0.6
    a, m = load_config('benchmark_5_2d')
    assert int(a[0].nb[0]) == 6
    # divergence D6: the command-line defaults win over constructor arguments
    assert m['kf'] == 1e3 and m['fric_coeff'] == 0.5
    # e = 0.6 for every pair (stack_of_cylinders.py:231-234)
    assert np.allclose(a[0].eta[a[0].eta != 0], 0.3209860933285677)


@pytest.mark.parametrize('name,planar', [('benchmark_1', True),
                                         ('benchmark_1_rb3d', False)])
def test_benchmark_1_known_answers(name, planar):
    """Free rigid motion (benchmark_1...py:30-31, 106-107): after 1000 steps
    xcm = (0.5, 0.5), omega_z = 1, theta = 1000 atan(1e-3), energy constant,
    M = 12.1, I_zz = 2.42."""
    arrays, meta = load_config(name)
    body = arrays[0]
    assert abs(body.total_mass[0] - 12.1) < 1e-12
    izz = body.izz[0] if planar else body.inertia_tensor_body_frame[8]
    assert abs(izz - 2.42) < 1e-12
    e0 = 0.5 * np.sum(body.m * (body.u**2 + body.v**2))
    rbo.gtvf_step(arrays, meta['rigid'], oracle_params(meta), planar=planar,
                  nsteps=1000)
    assert np.allclose(body.xcm[:2], [0.5, 0.5], atol=1e-12)
    assert abs(body.omega[2] - 1.0) < 1e-12
    theta = np.arctan2(body.R[3], body.R[0])
    assert abs(theta - 0.9999996666668673) < 1e-12
    e1 = 0.5 * np.sum(body.m * (body.u**2 + body.v**2))
    assert abs(e1 - e0) < 1e-12 * e0


def test_product_setup_equals_reference_setup():
    """setup_properties of the product (vectorised NumPy + KD-tree evaluator)
    against the reference's own setup executed by the harness (fixture)."""
    from rigid_body_2d_3d_pysph_b200.compat.particle_array import \
        get_particle_array
    from rigid_body_2d_3d_pysph_b200.rigid_body_3d import RigidBody3DScheme
    arrays, ref, meta = load_case('cubes3d')
    body, tank = arrays
    b2 = get_particle_array(name='body', x=body.x, y=body.y, z=body.z,
                            h=body.h, m=body.m, rho=body.rho,
                            constants={'spacing0': body.spacing0[0]})
    b2.add_property('body_id', type='int', data=body.body_id)
    b2.add_property('dem_id', type='int', data=body.dem_id)
    b2.add_constant('total_no_bodies', [3])
    t2 = get_particle_array(name='tank', x=tank.x, y=tank.y, z=tank.z,
                            h=tank.h, m=tank.m, rho=tank.rho)
    t2.add_property('dem_id', type='int', data=2)
    RigidBody3DScheme(['body'], ['tank'], dim=3).setup_properties([b2, t2])
    for n in ['total_mass', 'xcm', 'inertia_tensor_inverse_body_frame',
              'inertia_tensor_body_frame', 'dx0', 'dy0', 'dz0', 'normal',
              'normal0', 'R']:
        assert np.allclose(getattr(b2, n), getattr(body, n), rtol=1e-12,
                           atol=1e-13), n
    assert np.array_equal(b2.is_boundary, body.is_boundary)
    assert np.array_equal(t2.is_boundary, tank.is_boundary)


def test_equation_planner_rejects_unknown():
    from rigid_body_2d_3d_pysph_b200.compat.equation import (Equation, Group,
                                                             MultiStageEquations)
    from rigid_body_2d_3d_pysph_b200.compat.integrator import (
        GTVFIntegrator, plan_from_equations)
    from rigid_body_2d_3d_pysph_b200.rigid_body_3d import (
        GTVFRigidBody3DStep, RigidBody3DScheme)

    class Mystery(Equation):
        pass
    s = RigidBody3DScheme(['body'], ['tank'], dim=3, gy=-9.81)
    integ = GTVFIntegrator(body=GTVFRigidBody3DStep())
    plan = plan_from_equations(s.get_equations(), integ)
    assert plan.rigid == ['body'] and plan.boundaries == ['tank']
    assert plan.gy == -9.81 and not plan.planar
    bad = MultiStageEquations([[], [Group([Mystery('body', None)])]])
    with pytest.raises(NotImplementedError):
        plan_from_equations(bad, integ)
