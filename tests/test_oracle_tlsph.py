"""The TLSPH / FSI part of the CPU oracle against the known answers the reference's own tests hold
(SURVEY.md section 8(c) / 8(f) rank 3).  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O

ROT = np.array([[np.cos(0.3), -np.sin(0.3)], [np.sin(0.3), np.cos(0.3)]])
STRETCH = np.array([[2.0, 0.0], [0.0, 3.0]])


def grid_9x9(x_fastest=True):
    r = np.arange(1, 10) * 0.1
    if x_fastest:    # Iterators.product(range, range): the first index runs fastest
        return np.array([[x, y] for y in r for x in r])
    return np.array([[x, y] for x in r for y in r])    # particle = (x - 1) * 9 + y


def params(h, E=1.0, nu=1.0, kernel=O.KERNEL_SCHOENBERG_CUBIC):
    sp = O.TlsphParams()
    sp.ndims, sp.kernel, sp.smoothing_length = 2, kernel, h
    sp.young_modulus, sp.poisson_ratio = E, nu
    return sp


@pytest.mark.parametrize("name,deform,expected", [
    ("Stretch x", lambda x: np.array([[2.0, 0.0], [0.0, 1.0]]) @ x, np.array([[2.0, 0.0], [0.0, 1.0]])),
    ("Stretch Both", lambda x: STRETCH @ x, STRETCH),
    ("Rotation", lambda x: ROT @ x, ROT),
    ("Nonlinear Stretching", lambda x: np.array([x[0] ** 2, x[1]]), np.eye(2)),
])
def test_deformation_gradient_of_deformation_functions(name, deform, expected):
    """test/systems/tlsph_system.jl:194-247: 9 x 9 grid, spacing 0.1, cubic spline h = 0.12; the
    deformation gradient of the middle particle (41) equals the deformation matrix."""
    x0 = grid_9x9()
    mass, rho = np.full(81, 10.0), np.full(81, 1000.0)
    sp = params(0.12)
    L = O.tlsph_correction_matrix(sp, x0, mass, rho, np.float64)
    cur = np.array([deform(c) for c in x0])
    F, _ = O.tlsph_update(sp, x0, cur, mass, rho, L, np.float64)
    assert np.allclose(F[40].T, expected, rtol=1.5e-8, atol=1e-14)


@pytest.mark.parametrize("J,pk2,pk1", [
    (ROT, np.zeros((2, 2)), np.zeros((2, 2))),
    (STRETCH, np.array([[8.5, 0.0], [0.0, 13.5]]), np.array([[17.0, 0.0], [0.0, 40.5]])),
    (ROT @ STRETCH, np.array([[8.5, 0.0], [0.0, 13.5]]), ROT @ STRETCH @ np.array([[8.5, 0.0], [0.0, 13.5]])),
])
def test_stress_tensors(J, pk2, pk1):
    """test/systems/tlsph_system.jl:250-303: lambda = mu = 1 (E = 2.5, nu = 0.25); PK1 = F S.  The
    deformation is applied to the grid, where the corrected gradient reproduces F = J exactly; the
    oracle returns PK1 L / rho^2, so PK1 = rho^2 (PK1 L / rho^2) L^-1."""
    x0 = grid_9x9()
    mass, rho = np.full(81, 10.0), np.full(81, 1000.0)
    sp = params(0.12, E=2.5, nu=0.25)
    L = O.tlsph_correction_matrix(sp, x0, mass, rho, np.float64)
    cur = x0 @ J.T
    F, P = O.tlsph_update(sp, x0, cur, mass, rho, L, np.float64)
    assert np.allclose(F[40].T, J, atol=1e-13)
    got = 1000.0 ** 2 * P[40].T @ np.linalg.inv(L[40].T)
    assert np.allclose(got, pk1, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("deform,expected", [
    (lambda x: ROT @ x, [0.0, 0.0]),
    (lambda x: STRETCH @ x, [0.0, 0.0]),
    (lambda x: ROT @ STRETCH @ x, [0.0, 0.0]),
    (lambda x: np.array([x[0] ** 2, x[1]]), [10 / 1000 ** 2 * 1.5400218087591082 * 324.67072684047224 * 1.224, 0.0]),
])
def test_structure_interact_with_deformation_function(deform, expected):
    """test/schemes/structure/total_lagrangian_sph/rhs.jl:118-203: `kick!` of a 9 x 9 TLSPH grid
    (cubic spline h = 0.07, E = 2.5, nu = 0.25, no penalty force, no gravity): acceleration of the
    middle particle; tolerance sqrt(eps) as in the reference."""
    x0 = grid_9x9(x_fastest=False)
    mass, rho = np.full(81, 10.0), np.full(81, 1000.0)
    sp = params(0.07, E=2.5, nu=0.25)
    L = O.tlsph_correction_matrix(sp, x0, mass, rho, np.float64)
    cur = np.array([deform(c) for c in x0])
    F, P = O.tlsph_update(sp, x0, cur, mass, rho, L, np.float64)
    dv = O.tlsph_interact(sp, 81, x0, cur, mass, rho, F, P, np.float64)
    tol = np.sqrt(np.finfo(np.float64).eps)
    assert np.allclose(dv[40], expected, rtol=tol, atol=tol)


def test_monaghan_kajtar_rhs():
    """test/schemes/boundary/monaghan_kajtar/monaghan_kajtar.jl:11-67: 3 x 3 fluid particles left of a
    1 x 3 wall of repulsive particles (here: three clamped structure particles carrying the same
    BoundaryModelMonaghanKajtar(K = 1, beta = 0.5, spacing 0.2)); the reference's numbers for the
    middle column, zero for the left one, < -300 for the right one."""
    dx = 0.1
    xs = np.array([-0.1, 0.0, 0.1]) - 1.5 * dx
    ys = np.array([-0.1, 0.0, 0.1])
    fluid = np.array([[x, y] for y in ys for x in xs])
    wall = np.array([[0.0, y] for y in (-0.2, 0.0, 0.2)])
    fp = O.FluidParams()
    fp.ndims, fp.kernel, fp.density_calculator = 2, O.KERNEL_SCHOENBERG_CUBIC, O.DENSITY_CONTINUITY
    fp.smoothing_length, fp.sound_speed, fp.exponent, fp.reference_density = 1.2 * dx, 0.0, 1.0, 1000.0
    sp = params(1.2 * dx)
    sp.boundary_model, sp.mk_K, sp.mk_beta, sp.mk_spacing = O.BOUNDARY_MONAGHAN_KAJTAR, 1.0, 0.5, 2 * dx
    mass_f = np.full(9, 1000.0 * dx ** 2)
    mass_s = np.full(3, 1000.0 * (2 * dx) ** 2)
    v = np.concatenate([np.zeros((9, 2)), np.full((9, 1), 1000.0)], axis=1)
    L = np.tile(np.eye(2), (3, 1, 1))
    out = O.kick_fsi(fp, None, sp, mass_f, None, None, 0, wall, mass_s, np.full(3, 1000.0), mass_s, L,
                     v.reshape(-1), fluid.reshape(-1), np.float64)
    dv = out["dv"].reshape(9, 3)
    assert np.abs(dv[:, 1]).max() <= 1e-14
    assert np.all(dv[[0, 3, 6], :2] == 0)
    assert np.all(dv[[2, 5, 8], 0] < -300)
    assert np.allclose(dv[[1, 4, 7], 0], [-26.052449, -95.162888, -26.052449], rtol=1.5e-8)


def _dummy_structure_case(seed=3):
    """A block of structure particles under a block of fluid, the structure carrying
    BoundaryModelDummyParticles{AdamiPressureExtrapolation} (examples/fsi/hydrostatic_water_column_2d.jl:109-124)."""
    import trixiparticles.jl_b200 as tp
    rng = np.random.default_rng(seed)
    dx = 0.05
    se = tp.StateEquationCole(sound_speed=10.0, reference_density=1000.0, exponent=7)
    kernel, h = tp.WendlandC2Kernel(2), np.sqrt(2) * dx
    f_ic = tp.RectangularShape(dx, (8, 6), (0.0, 0.0), density=1000.0)
    s_ic = tp.RectangularShape(dx, (8, 3), (0.0, -3 * dx), density=2700.0)
    fluid = tp.WeaklyCompressibleSPHSystem(f_ic, smoothing_kernel=kernel, smoothing_length=h,
                                           density_calculator=tp.ContinuityDensity(), state_equation=se,
                                           acceleration=(0.0, 0.0))
    hyd_rho = np.full(s_ic.nparticles, 1000.0)
    model = tp.BoundaryModelDummyParticles(hyd_rho, hyd_rho * dx ** 2, tp.AdamiPressureExtrapolation(), kernel, h,
                                           state_equation=se)
    st = tp.TotalLagrangianSPHSystem(s_ic, smoothing_kernel=kernel, smoothing_length=h, young_modulus=1e6,
                                     poisson_ratio=0.3, boundary_model=model, acceleration=(0.0, 0.0))
    u_f = f_ic.coordinates + rng.uniform(-0.1 * dx, 0.1 * dx, f_ic.coordinates.shape)
    v_f = np.concatenate([rng.uniform(-0.5, 0.5, (f_ic.nparticles, 2)),
                          1000.0 * (1 + rng.uniform(-0.01, 0.01, (f_ic.nparticles, 1)))], axis=1)
    return tp, fluid, st, model, u_f, v_f


def test_dummy_particle_structure_at_rest_acts_like_a_wall():
    """A structure at rest with BoundaryModelDummyParticles gives the fluid exactly what a WallBoundarySystem of
    the same particles with the same boundary model gives it (wcsph/rhs.jl treats both alike; the Adami pass
    of dummy_particles.jl:489-672 does not care which system carries the model)."""
    from oracle import adapter
    tp, fluid, st, model, u_f, v_f = _dummy_structure_case()
    wall = tp.WallBoundarySystem(st.initial_condition, model)
    ref = adapter.kick(fluid, wall, u_f, v_f, use_grid=False)
    u_ode = np.concatenate([u_f.reshape(-1), st.initial_coordinates.reshape(-1)])
    v_ode = np.concatenate([v_f.reshape(-1), np.zeros(st.nparticles * 2)])
    out = adapter.kick_fsi(fluid, None, st, u_ode, v_ode)
    dv_f = out["dv"][: v_f.size].reshape(v_f.shape)
    np.testing.assert_allclose(dv_f, ref["dv"], rtol=1e-12, atol=1e-12 * np.abs(ref["dv"]).max())
    np.testing.assert_allclose(out["structure_pressure"], ref["wall_pressure"], rtol=1e-13, atol=1e-9)
    np.testing.assert_allclose(out["structure_density"], ref["wall_density"], rtol=1e-13)


def test_dummy_particle_structure_conserves_momentum_with_the_fluid():
    """structure.jl:60-95: the structure receives "the exact same pair force" the fluid feels, with flipped
    sign -- without gravity and with an undeformed structure the total momentum of fluid + structure is
    conserved: sum_f m_f dv_f + sum_s m_s dv_s = 0 (material masses on the structure side)."""
    from oracle import adapter
    tp, fluid, st, model, u_f, v_f = _dummy_structure_case(seed=5)
    rng = np.random.default_rng(9)
    v_s = rng.uniform(-0.2, 0.2, (st.nparticles, 2))        # a moving (still undeformed) structure
    u_ode = np.concatenate([u_f.reshape(-1), st.initial_coordinates.reshape(-1)])
    v_ode = np.concatenate([v_f.reshape(-1), v_s.reshape(-1)])
    out = adapter.kick_fsi(fluid, None, st, u_ode, v_ode)
    dv_f = out["dv"][: v_f.size].reshape(v_f.shape)[:, :2]
    dv_s = out["dv"][v_f.size:].reshape(-1, 2)
    assert np.abs(dv_s).max() > 0
    mom = (fluid.mass[:, None] * dv_f).sum(axis=0) + (st.mass[:, None] * dv_s).sum(axis=0)
    scale = np.abs(st.mass[:, None] * dv_s).sum()
    assert np.abs(mom).max() <= 1e-12 * scale
    # the structure's velocity enters the fluid's continuity equation (a wall's would be zero)
    still = adapter.kick_fsi(fluid, None, st, u_ode, np.concatenate([v_f.reshape(-1), 0 * v_s.reshape(-1)]))
    drho = out["dv"][: v_f.size].reshape(v_f.shape)[:, 2]
    assert np.abs(drho - still["dv"][: v_f.size].reshape(v_f.shape)[:, 2]).max() > 0


# ---------------------------------------------------------------------------------------------- three dimensions
def _rot3(ax, ay, az):
    cx, sx, cy, sy, cz, sz = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay), np.cos(az), np.sin(az)
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return rz @ ry @ rx


def _grid_3d(n=7, dx=0.1):
    r = np.arange(1, n + 1) * dx
    return np.array([[x, y, z] for z in r for y in r for x in r])


@pytest.mark.parametrize("name,J", [("rotation", _rot3(0.3, -0.2, 0.5)), ("stretch", np.diag([2.0, 3.0, 0.5])),
                                    ("rotation * stretch", _rot3(0.1, 0.7, -0.4) @ np.diag([1.5, 0.8, 1.2]))])
def test_deformation_gradient_and_stress_3d(name, J):
    """The 3-D counterpart of test/systems/tlsph_system.jl:194-303 (the reference tests these in 2-D only): on a
    7^3 grid the corrected gradient reproduces an affine deformation exactly, F = J in the interior, and
    PK1 = F (lambda tr(E) I + 2 mu E) with E = (F^T F - I) / 2; a rigid rotation is stress free."""
    x0 = _grid_3d()
    n = len(x0)
    mass, rho = np.full(n, 1.0), np.full(n, 1000.0)
    sp = params(0.12, E=2.5, nu=0.25)     # lambda = mu = 1
    sp.ndims = 3
    L = O.tlsph_correction_matrix(sp, x0, mass, rho, np.float64)
    cur = x0 @ J.T
    F, P = O.tlsph_update(sp, x0, cur, mass, rho, L, np.float64)
    mid = n // 2
    assert np.allclose(F[mid].T, J, atol=1e-12)
    E = 0.5 * (J.T @ J - np.eye(3))
    pk1 = J @ (np.trace(E) * np.eye(3) + 2 * E)
    got = 1000.0 ** 2 * P[mid].T @ np.linalg.inv(L[mid].T)
    assert np.allclose(got, pk1, atol=1e-9)
    if name == "rotation":
        dv = O.tlsph_interact(sp, n, x0, cur, mass, rho, F, P, np.float64)
        assert np.abs(dv).max() <= 1e-9          # no stress, no penalty force: hourglass-free rigid motion


def test_fsi_3d_momentum_balance():
    """3-D plate next to the water column (examples.dam_break_plate_3d): with gravity switched off the forces between
    the fluid and the plate's dummy particles are equal and opposite -- sum m_a dv_a over the fluid's share of the
    coupling + sum m_s dv_s over the plate's vanishes (the Adami coupling uses the same pair term on both sides,
    structure.jl:20-102)."""
    from trixiparticles.jl_b200 import examples
    from oracle import adapter
    import trixiparticles.jl_b200 as tp
    fluid, wall, st0, _ = examples.dam_break_plate_3d(0.03, structure_boundary_model="dummy_particles",
                                                      plate_position=(0.19, 0.0, 0.0075))
    # every plate particle integrated, so that the whole reaction shows up in dv_s
    st = tp.TotalLagrangianSPHSystem(st0.initial_condition, smoothing_kernel=st0.smoothing_kernel,
                                     smoothing_length=st0.smoothing_length, young_modulus=st0.young_modulus,
                                     poisson_ratio=st0.poisson_ratio, boundary_model=st0.boundary_model,
                                     acceleration=st0.acceleration, penalty_force=st0.penalty_force)
    n_f, n_int = fluid.nparticles, st.n_integrated_particles
    assert n_int == st.nparticles
    rng = np.random.default_rng(3)
    u_f = fluid.initial_condition.coordinates + rng.uniform(-0.002, 0.002, (n_f, 3))
    v_f = np.concatenate([rng.uniform(-0.1, 0.1, (n_f, 3)), fluid.initial_condition.density[:, None] * 1.001], axis=1)
    u = np.concatenate([u_f.reshape(-1), st.initial_coordinates.reshape(-1)])
    v = np.concatenate([v_f.reshape(-1), np.zeros(3 * n_int)])
    with_plate = adapter.kick_fsi(fluid, wall, st, u, v)["dv"]
    # the same fluid without the plate: the difference is the plate's force on the fluid
    no_plate = adapter.kick(fluid, wall, u_f, v_f)["dv"]
    f_fluid = ((with_plate[: 4 * n_f].reshape(n_f, 4) - no_plate)[:, :3] * fluid.mass[:, None]).sum(axis=0)
    # the plate at rest in its initial configuration: no stress, so dv_s = coupling + gravity
    dv_s = with_plate[4 * n_f:].reshape(n_int, 3) - st.acceleration[None, :]
    f_plate = (dv_s * st.mass[:, None]).sum(axis=0)
    assert np.linalg.norm(f_fluid) > 1e-3
    assert np.linalg.norm(f_fluid + f_plate) <= 1e-10 * np.linalg.norm(f_fluid)
