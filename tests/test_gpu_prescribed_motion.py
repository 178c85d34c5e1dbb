"""PrescribedMotion on the GPU: moving walls of dummy particles (examples/fluid/moving_wall_2d.jl,
accelerated_tank_2d.jl) and moving clamped particles of a TotalLagrangianSPHSystem, through the C ABI
(tpb_set_clamped_motion + tpb_kick) against the CPU oracle.  `-m gpu` only."""
import numpy as np
import pytest

import trixiparticles.jl_b200 as tp
from trixiparticles.jl_b200 import examples
from oracle import adapter

pytestmark = pytest.mark.gpu


def rel_inf(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def gpu_kick(semi, ode, u, v, t):
    dv = np.full(v.size, np.nan, dtype=v.dtype)
    tp.kick_(dv, np.ascontiguousarray(v).reshape(-1).copy(), np.ascontiguousarray(u).reshape(-1), ode.p, t)
    assert np.isfinite(dv).all()
    return dv.reshape(v.shape)


BOX = tp.GridNeighborhoodSearch(2, cell_list=tp.FullGridCellList((-1.0, -1.0), (5.0, 3.0)))


@pytest.mark.parametrize("eltype,tol", [(np.float64, 1e-11), (np.float32, 3e-5)])
def test_accelerated_tank_kick(eltype, tol):
    """accelerated_tank_2d.jl at t = 0.37: GPU == oracle, and == the hydrostatic tank seen from the tank's frame
    (dv + g e_y), the property test_oracle_prescribed_motion.py establishes for the oracle."""
    g, t = 9.81, 0.37
    fluid, wall, _ = examples.accelerated_tank_2d(0.05, eltype=eltype, coordinates_eltype=eltype)
    fluid_h, wall_h, _ = examples.hydrostatic_water_column_2d(0.05, eltype=eltype, coordinates_eltype=eltype)
    u0, v0 = examples.perturbed_state(fluid_h)
    shift, vel = np.array([0.0, 0.5 * g * t * t]), np.array([0.0, g * t])
    u, v = (u0 + shift).astype(eltype), v0.copy()
    v[:, :2] += vel.astype(eltype)
    semi = tp.Semidiscretization(fluid, wall, neighborhood_search=BOX, parallelization_backend=tp.B200Backend())
    ode = tp.semidiscretize(semi, (0.0, 1.0))
    dv = gpu_kick(semi, ode, u, v, t)
    ref = adapter.kick_moving_wall(fluid, wall, u, v)      # the wall object is in the state of the kick at t
    assert np.allclose(wall.coordinates, wall_h.coordinates + shift, atol=1e-6)
    assert rel_inf(dv[:, :2], ref["dv"][:, :2]) <= tol and rel_inf(dv[:, 2], ref["dv"][:, 2]) <= tol
    for name, key in (("pressure", "pressure_wall"), ("density", "density_wall")):
        assert rel_inf(semi.system_field(wall, name), ref[key]) <= 10 * tol, name
    semi.close()
    if eltype == np.float64:
        semi_h = tp.Semidiscretization(fluid_h, wall_h, parallelization_backend=tp.B200Backend())
        ode_h = tp.semidiscretize(semi_h, (0.0, 1.0))
        want = gpu_kick(semi_h, ode_h, u0, v0, 0.0)
        want[:, 1] += g
        assert rel_inf(dv, want) <= 1e-9
        semi_h.close()


@pytest.mark.parametrize("eltype,tol", [(np.float64, 1e-11), (np.float32, 3e-5)])
@pytest.mark.parametrize("extrapolation", ["adami", "bernoulli"])
def test_moving_wall_kick(eltype, tol, extrapolation):
    """moving_wall_2d.jl: only the right wall's face block moves (x + t^2/2 while t < 1.5).  Kicks while it moves
    (t = 0.3: velocity 0.3, acceleration 1 enter the continuity equation and the Adami / Bernoulli extrapolation),
    and after it stopped (t = 1.6 following a kick at t = 1.4: rests at its last position, velocity 0)."""
    fluid, wall, tank = examples.moving_wall_2d(0.05, eltype=eltype, coordinates_eltype=eltype)
    if extrapolation == "bernoulli":
        m = wall.boundary_model
        wall = tp.WallBoundarySystem(wall.initial_condition, tp.BoundaryModelDummyParticles(
            m.initial_density, m.hydrodynamic_mass, tp.BernoulliPressureExtrapolation(factor=0.8), m.smoothing_kernel,
            m.smoothing_length, state_equation=m.state_equation), prescribed_motion=wall.prescribed_motion)
    u, v = examples.perturbed_state(fluid)
    semi = tp.Semidiscretization(fluid, wall, neighborhood_search=BOX, parallelization_backend=tp.B200Backend())
    ode = tp.semidiscretize(semi, (0.0, 2.0))
    static = None
    for t in (0.3, 1.4, 1.6):
        # the fluid follows the wall so that the two stay in touch
        shift = np.array([0.5 * min(t, 1.4) ** 2, 0.0])
        u_t = (u + 0.995 * shift).astype(eltype)
        dv = gpu_kick(semi, ode, u_t, v, t)
        ref = adapter.kick_moving_wall(fluid, wall, u_t, v)
        assert wall.ismoving == (t < 1.5)
        assert np.allclose(wall.coordinates[tank.face_indices[1]],
                           wall.initial_condition.coordinates[tank.face_indices[1]] + shift, atol=1e-6)
        assert rel_inf(dv[:, :2], ref["dv"][:, :2]) <= tol and rel_inf(dv[:, 2], ref["dv"][:, 2]) <= tol, t
        assert rel_inf(semi.system_field(wall, "pressure"), ref["pressure_wall"]) <= 10 * tol, t
        if t == 1.4:
            static = dv
        if t == 1.6:
            # same geometry as at t = 1.4 (u_t equal), but the wall is at rest now
            assert np.abs(dv - static).max() > 1e-3 * np.abs(static).max()
    semi.close()


@pytest.mark.parametrize("boundary_model", ["monaghan_kajtar", "dummy_particles", "dummy_bernoulli"])
def test_structure_with_moving_clamped_particles(boundary_model):
    """`TotalLagrangianSPHSystem(...; clamped_particles_motion)` (system.jl:108-184, :403-447): the clamped base of
    the plate of dam_break_plate_2d.jl is shaken sideways; the plate's stress, the fluid's continuity equation and
    (Bernoulli) the extrapolated pressure see the prescribed positions and velocities."""
    from test_gpu_fsi import fsi_state
    motion = tp.PrescribedMotion(lambda x, t: x + np.array([[0.004 * np.sin(40.0 * t), 0.0]]), lambda t: True)
    fluid, wall, structure, _ = examples.dam_break_plate_2d(
        0.01, eltype=np.float64, coordinates_eltype=np.float64, initial_fluid_size=(0.15, 0.29),
        plate_position=(0.165, 0.0), clamped_particles_motion=motion,
        structure_boundary_model="monaghan_kajtar" if boundary_model == "monaghan_kajtar" else "dummy_particles",
        structure_pressure_extrapolation=tp.BernoulliPressureExtrapolation() if boundary_model == "dummy_bernoulli" else None)
    assert structure.n_clamped_particles > 0 and structure.prescribed_motion is motion
    u, v = fsi_state(fluid, structure)
    semi = tp.Semidiscretization(fluid, wall, structure, parallelization_backend=tp.B200Backend())
    ode = tp.semidiscretize(semi, (0.0, 1.0))
    results = []
    for t in (0.0, 0.03):
        dv = np.full_like(v, np.nan)
        ode.f1(dv, v, u, ode.p, t)
        ref = adapter.kick_fsi(fluid, wall, structure, u, v)
        n_f = fluid.nparticles
        for name, a, b in (("fluid", dv[: 3 * n_f], ref["dv"][: 3 * n_f]), ("structure", dv[3 * n_f:], ref["dv"][3 * n_f:])):
            assert rel_inf(a, b) <= 1e-11, (name, t, rel_inf(a, b))
        for name, key in (("deformation_grad", "F"), ("pk1_rho2", "pk1_rho2")):
            assert rel_inf(semi.system_field(structure, name), ref[key]) <= 1e-11, name
        results.append(dv.copy())
    n_f = fluid.nparticles
    assert np.abs(results[0][3 * n_f:] - results[1][3 * n_f:]).max() > 1.0    # the shaken base strains the plate
    assert np.abs(results[0][: 3 * n_f] - results[1][: 3 * n_f]).max() > 1e-6
    semi.close()


def test_accelerated_tank_time_loop():
    """accelerated_tank_2d.jl integrated in time: in the tank's frame the water column stays the hydrostatic one --
    the same run as hydrostatic_water_column_2d.jl up to rounding."""
    from trixiparticles.jl_b200.time_integration import CarpenterKennedy2N54, solve
    g = 9.81
    fluid_a, wall_a, _ = examples.accelerated_tank_2d(0.05)
    fluid_h, wall_h, _ = examples.hydrostatic_water_column_2d(0.05, eltype=np.float64, coordinates_eltype=np.float64)
    out = {}
    for name, fluid, wall, nhs in (("acc", fluid_a, wall_a, BOX), ("hyd", fluid_h, wall_h, None)):
        semi = tp.Semidiscretization(fluid, wall, neighborhood_search=nhs,
                                     parallelization_backend=tp.B200Backend(ode_memory="device"))
        ode = tp.semidiscretize(semi, (0.0, 0.05))
        sol = solve(ode, CarpenterKennedy2N54(), dt=2.5e-4, cuda_graph=True)   # (a moving wall switches the graph off)
        out[name] = (sol.u.cpu().numpy().reshape(-1, 2), sol.v.cpu().numpy().reshape(-1, 3), sol.t)
        semi.close()
    (u_a, v_a, t_a), (u_h, v_h, t_h) = out["acc"], out["hyd"]
    assert t_a == t_h
    assert np.abs(u_a - [0.0, 0.5 * g * t_a ** 2] - u_h).max() <= 1e-9
    assert np.abs(v_a[:, :2] - [0.0, g * t_a] - v_h[:, :2]).max() <= 1e-7
    assert rel_inf(v_a[:, 2], v_h[:, 2]) <= 1e-10


def test_moving_wall_2d_time_loop():
    """examples/fluid/moving_wall_2d.jl:71-77 as the reference's example test runs it (test/examples/
    examples_fluid.jl:690-700: retcode Success, no NaN), shortened to t = 0.8: `solve(ode, RDPK3SpFSAL35(),
    abstol=1e-6, reltol=1e-4, dtmax=1e-2)`.  The right wall recedes with x = 1 + t^2/2, the column collapses behind
    it: the front follows the wall, nothing passes through it, the mass stays in the tank."""
    from trixiparticles.jl_b200.time_integration import RDPK3SpFSAL35, solve
    fluid, wall, tank = examples.moving_wall_2d(0.05)
    semi = tp.Semidiscretization(fluid, wall, parallelization_backend=tp.B200Backend(ode_memory="device"))
    ode = tp.semidiscretize(semi, (0.0, 0.8))
    sol = solve(ode, RDPK3SpFSAL35(), abstol=1e-6, reltol=1e-4, dtmax=1e-2)
    assert sol.retcode == "Success" and sol.t == 0.8
    u = sol.u.cpu().numpy().reshape(-1, 2)
    v = sol.v.cpu().numpy().reshape(-1, 3)
    assert np.isfinite(u).all() and np.isfinite(v).all()
    x_wall = tank.fluid_size[0] + 0.5 * 0.8 ** 2            # inner surface of the moving wall
    mov = tank.face_indices[1]
    assert np.allclose(wall.coordinates[mov, 0].min(), x_wall + 0.025, atol=1e-9)
    assert u[:, 0].max() < x_wall and u[:, 0].min() > 0.0 and u[:, 1].min() > 0.0
    assert u[:, 0].max() > 1.1                              # the front has followed the wall
    assert u[:, 1].max() < 0.8 + 1e-3                       # and the column has not grown
    assert 950.0 < v[:, 2].min() and v[:, 2].max() < 1100.0
    semi.close()


def test_dam_break_gate_kick_is_the_superposition_of_its_parts():
    """examples/fsi/dam_break_gate_2d.jl: fluid + tank + moving gate + elastic plate.  Inside the library the gate's
    dummy particles share the plate's slot (clamped particles with a prescribed motion).  Independent check by
    superposition of pair sums: the fluid's dv with gate AND plate = (with the plate) + (with the gate) - (with
    neither), each from its own oracle path; the plate's dv and both Adami pressure fields as in the separate cases."""
    from test_gpu_fsi import fsi_state
    fluid, tank_w, gate_w, plate, _ = examples.dam_break_gate_2d(0.02)
    u, v = fsi_state(fluid, plate, seed=11)
    n_f, n_int = fluid.nparticles, plate.n_integrated_particles
    u_f = u[: 2 * n_f].reshape(n_f, 2)
    u_f[: n_f // 2, 0] += 0.33                   # half of the water already at the plate
    v_f = v[: 3 * n_f].reshape(n_f, 3)
    t = 0.05                                     # the gate is on its way up: velocity 5.2, acceleration 59
    semi = tp.Semidiscretization(fluid, tank_w, gate_w, plate, parallelization_backend=tp.B200Backend())
    ode = tp.semidiscretize(semi, (0.0, 1.0))
    assert semi.lib_index(gate_w) == semi.lib_index(plate) == 2 and semi.ranges_u[-1][1] == u.size
    dv = np.full_like(v, np.nan)
    ode.f1(dv, v, u, ode.p, t)
    assert np.isfinite(dv).all() and gate_w.ismoving
    lift = -285.115 * t ** 3 + 72.305 * t ** 2 + 0.1463 * t
    assert np.allclose(gate_w.coordinates, gate_w.initial_condition.coordinates + [0.0, lift])
    with_plate = adapter.kick_fsi(fluid, tank_w, plate, u, v)
    with_gate = adapter.kick_moving_wall(fluid, gate_w, u_f, v_f, static_wall=tank_w)
    neither = adapter.kick(fluid, tank_w, u_f, v_f)
    want_f = with_plate["dv"][: 3 * n_f].reshape(n_f, 3) + with_gate["dv"] - neither["dv"]
    got_f, got_s = dv[: 3 * n_f].reshape(n_f, 3), dv[3 * n_f:]
    # both couplings are active
    assert np.abs(with_gate["dv"] - neither["dv"]).max() > 1.0
    assert np.abs(with_plate["dv"][: 3 * n_f].reshape(n_f, 3) - neither["dv"]).max() > 1.0
    assert rel_inf(got_f[:, :2], want_f[:, :2]) <= 1e-11 and rel_inf(got_f[:, 2], want_f[:, 2]) <= 1e-11
    assert rel_inf(got_s, with_plate["dv"][3 * n_f:]) <= 1e-11
    assert rel_inf(semi.system_field(gate_w, "pressure"), with_gate["pressure_wall"]) <= 1e-10
    assert rel_inf(semi.system_field(plate, "pressure"), with_plate["structure_pressure"]) <= 1e-10
    assert rel_inf(semi.system_field(plate, "deformation_grad"), with_plate["F"]) <= 1e-11
    assert semi.system_field(gate_w, "density").shape == (gate_w.nparticles,)
    semi.close()


def test_dam_break_gate_2d_time_loop():
    """The example run to t = 0.25 (CarpenterKennedy2N54; the movement function runs on the host per kick): the gate
    stops at its prescribed height, the released water reaches the plate and bends it downstream."""
    from trixiparticles.jl_b200.time_integration import CarpenterKennedy2N54, solve
    fluid, tank_w, gate_w, plate, _ = examples.dam_break_gate_2d(0.02)
    semi = tp.Semidiscretization(fluid, tank_w, gate_w, plate, parallelization_backend=tp.B200Backend(ode_memory="device"))
    ode = tp.semidiscretize(semi, (0.0, 0.25))
    sol = solve(ode, CarpenterKennedy2N54(), dt=5e-5)
    assert sol.retcode == "Success"
    u = sol.u.cpu().numpy()
    assert np.isfinite(u).all()
    n_f, n_int = fluid.nparticles, plate.n_integrated_particles
    lift = -285.115 * 0.1 ** 3 + 72.305 * 0.1 ** 2 + 0.1463 * 0.1
    assert not gate_w.ismoving
    assert np.allclose(gate_w.coordinates[:, 1] - gate_w.initial_condition.coordinates[:, 1], lift, atol=2e-3)
    x_f = u[: 2 * n_f].reshape(n_f, 2)
    assert x_f[:, 0].max() > 0.5                          # the front has passed under the gate and reached the plate
    x_s = u[2 * n_f:].reshape(n_int, 2)
    x0 = plate.initial_coordinates[:n_int]
    top = x0[:, 1] > x0[:, 1].max() - 1e-9
    assert (x_s[top, 0] - x0[top, 0]).mean() > 1e-4       # bent downstream
    semi.close()
