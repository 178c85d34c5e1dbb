"""PrescribedMotion (schemes/boundary/prescribed_motion.jl) on the oracle side: the properties a moving wall of
dummy particles must have -- CPU only (`-m "not gpu"`).  The GPU parity tests are in test_gpu_prescribed_motion.py."""
import numpy as np
import pytest

import trixiparticles.jl_b200 as tp
from trixiparticles.jl_b200 import examples
from oracle import adapter


def rel_inf(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def test_movement_function_derivatives():
    """prescribed_motion.jl:95-121: position, d/dt and d2/dt2 of the movement function.  The finite-difference
    fallback against the exact derivatives of the doc-string example (circular motion, :24-26)."""
    f = lambda x, t: x + np.stack([np.cos(2 * np.pi * t) * np.ones(len(x)), np.sin(2 * np.pi * t) * np.ones(len(x))], 1)
    pm = tp.PrescribedMotion(f, lambda t: t < 1.5).initialize(4)
    x0 = np.arange(8.0).reshape(4, 2)
    pos, vel, acc = pm(x0, 0.3)
    w = 2 * np.pi
    assert np.allclose(pos, x0 + [np.cos(w * 0.3), np.sin(w * 0.3)], rtol=0, atol=1e-15)
    assert np.allclose(vel, [-w * np.sin(w * 0.3), w * np.cos(w * 0.3)], rtol=0, atol=1e-8)
    assert np.allclose(acc, [-w * w * np.cos(w * 0.3), -w * w * np.sin(w * 0.3)], rtol=0, atol=1e-6)
    assert pm.is_moving(1.0) and not pm.is_moving(2.0)
    assert list(pm.moving_particles) == [0, 1, 2, 3]
    # the rotation example (:29-33) keeps the distance from the origin; OscillatingMotion2D at t = 0 is the identity
    osc = tp.OscillatingMotion2D(frequency=1.0, translation_vector=(1.0, 0.0), rotation_angle=np.pi / 2,
                                 rotation_center=(0.0, 0.0)).initialize(4)
    assert np.allclose(osc(x0, 0.0)[0], x0, atol=1e-15)
    assert np.allclose(osc(x0, 0.25)[0], np.stack([-x0[:, 1], x0[:, 0]], 1) + [1.0, 0.0], atol=1e-12)


def test_moving_particles_subset_and_rest():
    """moving_wall_2d.jl: only `tank.face_indices[2]` moves; after is_moving(t) turns false the particles rest at
    their last position with zero velocity (wall_boundary/system.jl:109-127)."""
    fluid, wall, tank = examples.moving_wall_2d(0.1)
    x0 = wall.initial_condition.coordinates.copy()
    mov = tank.face_indices[1]
    assert np.allclose(x0[mov, 0].min(), tank.fluid_size[0] + 0.05)        # reset_wall!: next to the column
    assert wall.apply_prescribed_motion(1.0)
    rest = np.setdiff1d(np.arange(wall.nparticles), mov)
    assert np.array_equal(wall.clamped_coordinates[rest], x0[rest])
    assert np.allclose(wall.clamped_coordinates[mov], x0[mov] + [0.5, 0.0])
    assert np.allclose(wall.clamped_velocity[mov], [1.0, 0.0]) and np.all(wall.clamped_velocity[rest] == 0)
    assert not wall.apply_prescribed_motion(1.6) and not wall.ismoving
    assert np.allclose(wall.clamped_coordinates[mov], x0[mov] + [0.5, 0.0])    # stays where it was


def test_accelerated_tank_equals_hydrostatic_tank():
    """accelerated_tank_2d.jl: no gravity, the tank accelerated upwards with g.  In the tank's frame this IS the
    hydrostatic water column: with positions shifted by g t^2 / 2 and velocities by g t, every pair term is the
    same (the Adami hydrostatic term uses acceleration_source - current_acceleration(wall), dummy_particles.jl:
    652-654; the continuity equation the velocity difference to the moving wall particle), so
    dv_accelerated = dv_hydrostatic + g e_y for the momentum rows and equal density rates."""
    g, t = 9.81, 0.37
    fluid_h, wall_h, _ = examples.hydrostatic_water_column_2d(0.1, eltype=np.float64, coordinates_eltype=np.float64)
    fluid_a, wall_a, _ = examples.accelerated_tank_2d(0.1)
    u, v = examples.perturbed_state(fluid_h)
    ref = adapter.kick(fluid_h, wall_h, u, v)
    assert wall_a.apply_prescribed_motion(t)
    shift, vel = np.array([0.0, 0.5 * g * t * t]), np.array([0.0, g * t])
    assert np.allclose(wall_a.clamped_coordinates, wall_h.coordinates + shift)
    u_a, v_a = u + shift, v.copy()
    v_a[:, :2] += vel
    out = adapter.kick_moving_wall(fluid_a, wall_a, u_a, v_a)
    want = ref["dv"].copy()
    want[:, 1] += g
    assert rel_inf(out["dv"][:, :2], want[:, :2]) <= 1e-10
    assert rel_inf(out["dv"][:, 2], want[:, 2]) <= 1e-10
    assert rel_inf(out["pressure_wall"], ref["wall_pressure"]) <= 1e-10
    # and it matters: with the wall at rest (is_moving false) the result differs
    wall_a.ismoving = False
    still = adapter.kick_moving_wall(fluid_a, wall_a, u_a, v_a)
    assert np.abs(still["dv"] - want).max() > 1.0


def test_bernoulli_term_only_while_moving():
    """BernoulliPressureExtrapolation (dummy_particles.jl:674-694): + factor rho_f (v_rel . n)^2 / 2 per fluid
    neighbour while the wall moves, nothing when it rests; one wall particle, one fluid particle."""
    fluid, wall, _ = examples.hydrostatic_water_column_2d(0.1, eltype=np.float64, coordinates_eltype=np.float64)
    kernel, h = fluid.smoothing_kernel, fluid.smoothing_length
    se = fluid.state_equation
    ic_w = tp.InitialCondition(np.array([[0.0, 0.0]]), np.zeros((1, 2)), np.array([10.0]), np.array([1000.0]),
                               np.zeros(1), 0.1)
    ic_f = tp.InitialCondition(np.array([[0.06, 0.08]]), np.zeros((1, 2)), np.array([10.0]), np.array([1000.0]),
                               np.zeros(1), 0.1)
    fl = tp.WeaklyCompressibleSPHSystem(ic_f, smoothing_kernel=kernel, smoothing_length=h,
                                        density_calculator=tp.ContinuityDensity(), state_equation=se,
                                        acceleration=(0.0, 0.0))
    res = {}
    for name, calc in (("adami", tp.AdamiPressureExtrapolation()), ("bernoulli", tp.BernoulliPressureExtrapolation(factor=0.7))):
        model = tp.BoundaryModelDummyParticles(ic_w.density, ic_w.mass, calc, kernel, h, state_equation=se)
        motion = tp.PrescribedMotion(lambda x, t: x + np.array([[2.0 * t, 0.0]]), lambda t: t < 1.0,
                                     velocity_function=lambda x, t: np.array([[2.0, 0.0]]),
                                     acceleration_function=lambda x, t: np.zeros((1, 2)))
        w = tp.WallBoundarySystem(ic_w, model, prescribed_motion=motion)
        u = ic_f.coordinates.copy()
        v = np.array([[0.5, -0.25, 1010.0]])
        for t in (0.0, 2.0):
            w.apply_prescribed_motion(t)
            w.clamped_coordinates[:] = 0.0      # same geometry in both states
            res[name, t] = adapter.kick_moving_wall(fl, w, u, v)["pressure_wall"][0]
    vn = ((2.0 - 0.5) * (0.0 - 0.06) + (0.0 + 0.25) * (0.0 - 0.08)) / 0.1
    assert res["adami", 0.0] == res["adami", 2.0] == res["bernoulli", 2.0]
    assert np.isclose(res["bernoulli", 0.0] - res["adami", 0.0], 0.7 * 1010.0 * vn * vn / 2, rtol=1e-12)
