"""Pins the CPU oracle against the known answers held by the reference's own tests
(SURVEY.md section 8(c)).  CPU only.  Paths cited are relative to /root/reference."""
import numpy as np
import pytest

import trixiparticles.jl_b200 as tp
from oracle import adapter

WC2, CUBIC, WC4, WC6, QUARTIC, QUINTIC = 0, 1, 2, 3, 4, 5
SUPPORT = {WC2: 2.0, CUBIC: 2.0, WC4: 2.0, WC6: 2.0, QUARTIC: 2.5, QUINTIC: 3.0}   # compact_support / h


# test/schemes/fluid/viscosity.jl:2-42
def test_monaghan_viscosity_known_answer(oracle):
    particle_spacing = 0.2
    h = 1.2 * particle_spacing
    c = 10 * np.sqrt(9.81 * 0.9)
    dv = oracle.viscosity_pair(CUBIC, 2, h, 0.02, 0.0, 0.01, c, 1.0, 1000.0, 1000.0,
                               [0.3, -1.0], [-0.25 * h, 0.375 * h])
    assert dv[0] == pytest.approx(-0.02049217623299368, abs=6e-15)
    assert dv[1] == pytest.approx(0.03073826434949052, abs=6e-15)


# test/schemes/fluid/viscosity.jl:43-105: the same pair through ViscosityMorris / ViscosityAdami
@pytest.mark.parametrize("model,expected", [
    (2, (-1.0895602048035404e-5, 3.631867349345135e-5)),      # ViscosityMorris(nu=7e-3)
    (3, (-1.089560204803541e-5, 3.6318673493451364e-5)),      # ViscosityAdami(nu=7e-3)
])
def test_morris_adami_viscosity_known_answers(oracle, model, expected):
    particle_spacing = 0.2
    h = 1.2 * particle_spacing
    dv = oracle.viscosity_pair_nu(model, CUBIC, 2, h, 7e-3, 0.01, 0.01, 0.01, 1000.0, 1000.0,
                                  [0.3, -1.0], [-0.25 * h, 0.375 * h])
    assert dv[0] == pytest.approx(expected[0], abs=6e-15)
    assert dv[1] == pytest.approx(expected[1], abs=6e-15)


# test/schemes/fluid/weakly_compressible_sph/state_equation.jl:24-38, :85-98 (exact `==`)
@pytest.mark.parametrize("gamma,expected", [
    (7.15, [998.34, 1002.8323123356663, 1019.8235062499685, 1038.8747989986027,
            1099.0607413267035, 1210.4689472510186]),
    (1.0, [998.34, 1002.8949541016121, 1021.298809057621, 1044.3036277526319,
           1136.3229025326757, 1412.380726872807]),
])
def test_cole_inverse_known_answers(oracle, gamma, expected):
    ATM = 101_325.0
    for mult, exp in zip([1, 100, 500, 1000, 3000, 9000], expected):
        got = oracle.inverse_eos(1484.0, gamma, 998.34, ATM, mult * ATM)
        assert got == exp  # bit-exact, as in the reference test


# state_equation.jl:41-61
def test_cole_background_pressure_and_clipping(oracle):
    for pb in [0.0, 10_000.0, 100_000.0, 200_000.0]:
        assert oracle.eos(10.0, 7, 1000.0, pb, 0, 1000.0) == pb
        assert oracle.eos(10.0, 7, 1000.0, pb, 0, 1001.0) > pb + 10
        assert oracle.eos(10.0, 7, 1000.0, pb, 0, 999.0) < pb - 10
    assert oracle.eos(10.0, 7, 1000.0, 0.0, 1, 999.0) == 0.0
    assert oracle.eos(10.0, 7, 1000.0, 0.0, 1, 900.0) == 0.0


# state_equation.jl:103-141
def test_cole_inverse_roundtrip(oracle):
    eqs = [(1484.0, 7.15, 998.34, 101_325.0), (10.0, 7, 1000.0, 10_000.0), (10.0, 7, 1000.0, 0.0),
           (10.0, 7, 1000.0, -100_000.0), (1484.0, 1, 998.34, 101_325.0),
           (10.0, 1, 1000.0, 100_000.0), (10.0, 1, 1000.0, 90_000.0), (10.0, 1, 1000.0, 0.0),
           (10.0, 1, 1000.0, -100_000.0)]
    for c, g, r0, pb in eqs:
        for density in [100.0, 500.0, 900.0, 990.0, 1000.0, 1005.0, 1100.0, 1600.0]:
            p = oracle.eos(c, g, r0, pb, 0, density)
            assert oracle.inverse_eos(c, g, r0, pb, p) == pytest.approx(density, rel=1.5e-8)
        for p in [-100.0, 0.0, 100.0, 10_000.0, 100_000.0, 100_000_000.0]:
            rho = oracle.inverse_eos(c, g, r0, pb, p)
            if not np.isfinite(rho):
                continue  # negative base of a fractional power, NaN in the reference too
            assert oracle.eos(c, g, r0, pb, 0, rho) == pytest.approx(p, abs=2e-7, rel=1e-10)


# test/general/smoothing_kernels.jl:61-73 (normalisation) and :99-132 (derivative)
@pytest.mark.parametrize("kernel", [WC2, CUBIC, WC4, WC6, QUARTIC, QUINTIC])
@pytest.mark.parametrize("nd", [2, 3])
def test_kernel_normalisation_and_derivative(oracle, kernel, nd):
    from scipy.integrate import quad
    sup = SUPPORT[kernel]
    for h in [0.1, 1.0, 1.7]:
        surf = (lambda r: 2 * np.pi * r) if nd == 2 else (lambda r: 4 * np.pi * r * r)
        integral, _ = quad(lambda r: oracle.kernel(kernel, nd, r, h) * surf(r), 0, sup * h,
                           points=[0.5 * h, h, 1.5 * h, 2 * h][: 4 if sup > 2 else 2], epsabs=1e-13, epsrel=1e-12)
        assert integral == pytest.approx(1.0, abs=1e-12)
        for r in np.linspace(0.05, sup - 0.05, 11) * h:
            d = 1e-6 * h
            fd = (oracle.kernel(kernel, nd, r + d, h) - oracle.kernel(kernel, nd, r - d, h)) / (2 * d)
            assert oracle.kernel_deriv_div_r(kernel, nd, r, h) * r == pytest.approx(fd, rel=2e-6, abs=1e-9)
        # compact support: strict `<` (smoothing_kernels.jl:30-34)
        assert oracle.kernel(kernel, nd, sup * h, h) == 0.0
        assert oracle.kernel(kernel, nd, np.nextafter(sup * h, 0), h) >= 0.0


def test_wendland_c4_c6_closed_forms(oracle):
    """smoothing_kernels.jl:491-514, :550-574 evaluated directly (independent of the C code)."""
    for nd in (2, 3):
        for h in (0.3, 1.1):
            for r in (0.0, 0.4 * h, 1.3 * h, 1.99 * h):
                q = r / h
                s4 = {2: 9 / (4 * np.pi), 3: 495 / (256 * np.pi)}[nd] / h ** nd
                s6 = {2: 39 / (14 * np.pi), 3: 1365 / (512 * np.pi)}[nd] / h ** nd
                w4 = s4 * (1 - q / 2) ** 6 * (35 * q * q / 12 + 3 * q + 1)
                w6 = s6 * (1 - q / 2) ** 8 * (4 * q ** 3 + 25 * q * q / 4 + 4 * q + 1)
                assert oracle.kernel(WC4, nd, r, h) == pytest.approx(w4, rel=1e-14, abs=1e-300)
                assert oracle.kernel(WC6, nd, r, h) == pytest.approx(w6, rel=1e-14, abs=1e-300)
                if r > 0:
                    d4 = s4 * (-7 / 3) * (2 + 5 * q) * (1 - q / 2) ** 5 / h ** 2
                    d6 = s6 * (-11 / 4) * (8 * q * q + 7 * q + 2) * (1 - q / 2) ** 7 / h ** 2
                    assert oracle.kernel_deriv_div_r(WC4, nd, r, h) == pytest.approx(d4, rel=1e-13)
                    assert oracle.kernel_deriv_div_r(WC6, nd, r, h) == pytest.approx(d6, rel=1e-13)


# test/general/smoothing_kernels.jl:135-176: Float32 evaluation stays close to Float64
@pytest.mark.parametrize("kernel", [WC2, CUBIC, WC4, WC6, QUARTIC, QUINTIC])
def test_kernel_float32(oracle, kernel):
    for nd in (2, 3):
        for r in [0.1, 0.5, 1.3]:
            w64 = oracle.kernel(kernel, nd, r, 0.8, np.float64)
            w32 = oracle.kernel(kernel, nd, r, 0.8, np.float32)
            assert w32 == pytest.approx(w64, rel=2e-6)
            assert np.float32(w32) == w32  # value is representable in Float32


def _adami_wall_velocity_fixture(eltype=np.float64):
    # test/schemes/boundary/dummy_particles/dummy_particles.jl:104-140
    dx = 0.1
    b1 = tp.RectangularShape(dx, (10, 1), (0.0, 0.2), density=257.0, eltype=eltype)
    b2 = tp.RectangularShape(dx, (10, 1), (0.0, 0.1), density=257.0, eltype=eltype)
    b3 = tp.RectangularShape(dx, (10, 1), (0.0, 0.0), density=257.0, eltype=eltype)
    boundary = tp.union(b1, b2, b3)
    fluid = tp.RectangularShape(dx, (16, 5), (-0.3, 0.3), density=257.0, loop_order="x_first",
                                eltype=eltype)
    return boundary, fluid


# dummy_particles.jl:197-233 + :283-297: the kernel-weighted average that defines the Adami
# extrapolation, probed with a linear EOS so that p_f plays the role of the test's velocity.
@pytest.mark.parametrize("scale", [1.0, 0.5, 0.7, 1.8, 67.5])
def test_adami_weights_known_answers(oracle, scale):
    boundary, fluid_ic = _adami_wall_velocity_fixture()
    h = 1.2 * 0.1
    c, rho0 = 10.0, 257.0
    B = rho0 * c * c  # exponent 1
    se = tp.StateEquationCole(sound_speed=c, reference_density=rho0, exponent=1)
    kernel = tp.SchoenbergCubicSplineKernel(2)
    # staggered profile: 1-based odd particles carry `scale` (dummy_particles.jl:252-257)
    p_target = np.where(np.arange(1, fluid_ic.nparticles + 1) % 2 == 1, scale, 0.0)
    rho = rho0 * (p_target / B + 1.0)
    fluid = tp.WeaklyCompressibleSPHSystem(fluid_ic, smoothing_kernel=kernel, smoothing_length=h,
                                           density_calculator=tp.ContinuityDensity(),
                                           state_equation=se)  # zero acceleration
    model = tp.BoundaryModelDummyParticles(boundary.density, boundary.mass,
                                           tp.AdamiPressureExtrapolation(), kernel, h,
                                           state_equation=se)
    wall = tp.WallBoundarySystem(boundary, model)
    u = fluid_ic.coordinates
    v = np.concatenate([np.zeros((fluid_ic.nparticles, 2)), rho[:, None]], axis=1)
    for use_grid in (False, True):
        out = adapter.kick(fluid, wall, u, v, use_grid=use_grid)
        pw = out["wall_pressure"]
        expected = np.zeros(30)
        for i in range(1, 11):
            expected[i - 1] = (0.42040669416720744 if i % 2 == 1 else 0.5795933058327924) * scale
        for i in range(11, 21):
            expected[i - 1] = (0.12101100073462243 if i % 2 == 1 else 0.8789889992653775) * scale
        np.testing.assert_allclose(pw, expected, rtol=1e-11, atol=1e-13)
        # third row is outside the compact support: volume 0, pressure stays 0
        assert np.all(out["wall_volume"][20:] == 0.0)
        assert np.all(out["wall_volume"][:20] > 0.0)


def _no_slip_fixture():
    # test/schemes/boundary/dummy_particles/dummy_particles.jl:104-178: ViscosityAdami(nu=1e-6) wall
    boundary, fluid_ic = _adami_wall_velocity_fixture()
    h = 1.2 * 0.1
    se = tp.StateEquationCole(sound_speed=10.0, reference_density=257.0, exponent=7)
    kernel = tp.SchoenbergCubicSplineKernel(2)
    fluid = tp.WeaklyCompressibleSPHSystem(fluid_ic, smoothing_kernel=kernel, smoothing_length=h,
                                           density_calculator=tp.ContinuityDensity(),
                                           state_equation=se, viscosity=tp.ViscosityAdami(nu=1e-6))
    model = tp.BoundaryModelDummyParticles(boundary.density, boundary.mass,
                                           tp.AdamiPressureExtrapolation(), kernel, h,
                                           state_equation=se, viscosity=tp.ViscosityAdami(nu=1e-6))
    return fluid_ic, fluid, tp.WallBoundarySystem(boundary, model)


# dummy_particles.jl:180-233: constant fluid velocity => v_wall = -v_fluid inside the compact support
@pytest.mark.parametrize("v_fluid", [(0.0, -1.0), (1.0, 1.0), (-1.0, 0.0), (0.7, 0.2), (0.3, 0.8)])
def test_no_slip_wall_velocity_constant_profile(oracle, v_fluid):
    fluid_ic, fluid, wall = _no_slip_fixture()
    n = fluid_ic.nparticles
    v = np.concatenate([np.tile(np.array(v_fluid), (n, 1)), np.full((n, 1), 257.0)], axis=1)
    for use_grid in (False, True):
        out = adapter.kick(fluid, wall, fluid_ic.coordinates, v, use_grid=use_grid)
        expected = np.zeros((30, 2))
        expected[:20] = -np.array(v_fluid)
        np.testing.assert_allclose(out["wall_velocity"], expected, rtol=1e-8, atol=1e-14)  # isapprox


# dummy_particles.jl:235-303: staggered profile, the reference's explicit numbers
@pytest.mark.parametrize("scale", [1.0, 0.5, 0.7, 1.8, 67.5])
def test_no_slip_wall_velocity_staggered_known_answers(oracle, scale):
    fluid_ic, fluid, wall = _no_slip_fixture()
    n = fluid_ic.nparticles
    vel = np.where((np.arange(1, n + 1) % 2 == 1)[:, None], scale, 0.0) * np.ones((n, 2))
    v = np.concatenate([vel, np.full((n, 1), 257.0)], axis=1)
    out = adapter.kick(fluid, wall, fluid_ic.coordinates, v)
    expected = np.zeros((30, 2))
    for i in range(1, 11):
        expected[i - 1] = -(0.42040669416720744 if i % 2 == 1 else 0.5795933058327924) * scale
    for i in range(11, 21):
        expected[i - 1] = -(0.12101100073462243 if i % 2 == 1 else 0.8789889992653775) * scale
    np.testing.assert_allclose(out["wall_velocity"], expected, rtol=1e-11, atol=1e-13)


def test_no_slip_wall_viscous_term_matches_pair_functions(oracle):
    """The fluid<-wall viscous term of a no-slip wall is the wall model's pair formula (pinned above to
    test/schemes/fluid/viscosity.jl) evaluated with v_b = wall_velocity: difference of two oracle
    kicks (no-slip minus free-slip) against the sum over the wall neighbours."""
    fluid_ic, fluid, wall = _no_slip_fixture()
    n = fluid_ic.nparticles
    rng = np.random.default_rng(5)
    v = np.concatenate([rng.normal(size=(n, 2)), 257.0 * (1 + 0.01 * rng.normal(size=(n, 1)))], axis=1)
    u = fluid_ic.coordinates
    for visc in (tp.ViscosityAdami(nu=1e-3), tp.ViscosityMorris(nu=2e-3),
                 tp.ArtificialViscosityMonaghan(alpha=0.05, beta=0.1)):
        wall.boundary_model.viscosity = visc
        ns = adapter.kick(fluid, wall, u, v)
        wall.boundary_model.viscosity = None
        fs = adapter.kick(fluid, wall, u, v)
        diff = ns["dv"][:, :2] - fs["dv"][:, :2]
        h, c = 0.12, 10.0
        expected = np.zeros((n, 2))
        xw, mw = wall.coordinates, wall.boundary_model.hydrodynamic_mass
        for a in range(n):
            for b in range(xw.shape[0]):
                pd = u[a] - xw[b]
                if pd @ pd > (2 * h) ** 2:
                    continue
                vd = v[a, :2] - ns["wall_velocity"][b]
                if isinstance(visc, tp.ArtificialViscosityMonaghan):
                    expected[a] += oracle.viscosity_pair(CUBIC, 2, h, visc.alpha, visc.beta, visc.epsilon, c,
                                                         mw[b], v[a, 2], ns["wall_density"][b], vd, pd)
                else:
                    # nu_a (fluid: ViscosityAdami(1e-6)) != nu_b (wall): checked in numpy below
                    nu_a, nu_b = 1e-6, visc.nu
                    r2 = pd @ pd
                    r = np.sqrt(r2)
                    grad = oracle.kernel_deriv_div_r(CUBIC, 2, r, h) * pd
                    rho_a, rho_b, m_a, m_b = v[a, 2], ns["wall_density"][b], fluid.mass[a], mw[b]
                    d2e = r2 + visc.epsilon * h * h
                    if isinstance(visc, tp.ViscosityMorris):
                        coef = m_b * (nu_a * rho_a + nu_b * rho_b) * (pd @ grad) / (rho_a * rho_b * d2e)
                    else:
                        eta_a, eta_b = nu_a * rho_a, nu_b * rho_b
                        tmp = 2 * eta_a * eta_b / ((eta_a + eta_b) * d2e * m_a)
                        coef = ((m_a / rho_a) ** 2 + (m_b / rho_b) ** 2) * (grad @ pd) * tmp
                    expected[a] += coef * vd
        scale = np.abs(fs["dv"][:, :2]).max()
        assert np.abs(expected).max() > 0
        np.testing.assert_allclose(diff, expected, rtol=0, atol=1e-12 * scale)


def _adami_tank(density, acceleration=None, eltype=np.float64):
    # dummy_particles.jl:305-328
    dx, n, n_layers = 0.1, 10, 2
    width = height = dx * n
    se = tp.StateEquationCole(sound_speed=10, reference_density=257, exponent=7)
    tank = tp.RectangularTank(dx, (width, height), (width, height), density, n_layers=n_layers,
                              faces=(True, True, True, False), acceleration=acceleration,
                              state_equation=se if acceleration is not None else None, eltype=eltype)
    kernel = tp.SchoenbergCubicSplineKernel(2)
    h = 1.5 * dx
    model = tp.BoundaryModelDummyParticles(tank.boundary.density, tank.boundary.mass,
                                           tp.AdamiPressureExtrapolation(), kernel, h,
                                           state_equation=se)
    wall = tp.WallBoundarySystem(tank.boundary, model)
    fluid = tp.WeaklyCompressibleSPHSystem(tank.fluid, smoothing_kernel=kernel, smoothing_length=h,
                                           density_calculator=tp.ContinuityDensity(),
                                           state_equation=se,
                                           acceleration=acceleration or (0.0, 0.0))
    return tank, fluid, wall


def _state(tank):
    n = tank.fluid.nparticles
    v = np.concatenate([np.zeros((n, 2)), tank.fluid.density[:, None]], axis=1)
    return tank.fluid.coordinates, v


# dummy_particles.jl:337-366
def test_adami_constant_zero_pressure(oracle):
    tank, fluid, wall = _adami_tank(257)
    out = adapter.kick(fluid, wall, *_state(tank))
    assert np.all(out["wall_pressure"] == 0.0)
    assert np.all(out["pressure"] == 0.0)


# dummy_particles.jl:371-410
def test_adami_constant_nonzero_pressure(oracle):
    tank, fluid, wall = _adami_tank(260)
    out = adapter.kick(fluid, wall, *_state(tank))
    assert np.allclose(out["pressure"], out["pressure"][0], rtol=1.5e-8)
    # only wall particles with fluid in their support carry the extrapolated value
    has_fluid = out["wall_volume"] > np.finfo(float).eps
    assert has_fluid.all()
    np.testing.assert_allclose(out["wall_pressure"], out["pressure"][0], atol=1e-12, rtol=0)


# dummy_particles.jl:611-697: hydrostatic gradient vs a bigger reference tank, atol = 4.0
def test_adami_hydrostatic_gradient(oracle):
    dx, n, n_layers = 0.1, 10, 2
    tank, fluid, wall = _adami_tank(257, acceleration=(0.0, -9.81))
    out = adapter.kick(fluid, wall, *_state(tank))
    se = fluid.state_equation
    ref = tp.RectangularTank(dx, (dx * (n + 2 * n_layers), dx * (n + n_layers)),
                             (dx * (n + 2 * n_layers), dx * (n + n_layers)), 257,
                             acceleration=(0.0, -9.81), state_equation=se, n_layers=0,
                             faces=(True, True, True, False))

    def to_matrix(coords, values, offset):
        m = np.zeros((n + 2 * n_layers, n + n_layers))
        idx = np.rint(coords / dx + offset).astype(int) - 1
        m[idx[:, 0], idx[:, 1]] = values
        return m

    pressure = to_matrix(wall.coordinates, out["wall_pressure"], n_layers + 0.5)
    pressure += to_matrix(tank.fluid.coordinates, out["pressure"], n_layers + 0.5)
    pressure_ref = to_matrix(ref.fluid.coordinates, ref.fluid.pressure, 0.5)
    np.testing.assert_allclose(pressure, pressure_ref, atol=4.0, rtol=0)


def _patch(rng, nd=2, dx=0.3, perturb=True):
    # test/rectangular_patch.jl:2-38 with our own RNG (the properties are seed-independent)
    size = (3,) * nd
    ic = tp.RectangularShape(dx, size, tuple(-dx / 2 * s for s in size), density=1000.0)
    if perturb:
        ic.coordinates += rng.uniform(-0.5 * dx, 0.5 * dx, ic.coordinates.shape)
        ic.coordinates[int(np.ceil(ic.nparticles / 2)) - 1] = 0.0
        ic.mass += rng.uniform(-0.1, 0.1, ic.nparticles) * ic.mass[0]
        ic.density += rng.uniform(-0.1, 0.1, ic.nparticles) * 1000.0
        ic.velocity += rng.uniform(-0.5 * dx, 0.5 * dx, ic.velocity.shape)
    return ic


# test/schemes/fluid/rhs.jl:4-107: m_a dv_ab = -m_b dv_ba for the pressure + viscous terms
@pytest.mark.parametrize("density_calculator", [tp.ContinuityDensity(), tp.SummationDensity()])
def test_pair_antisymmetry(oracle, density_calculator):
    rng = np.random.default_rng(7)
    ic = _patch(rng, perturb=False)
    se = tp.StateEquationCole(sound_speed=10.0, reference_density=1000.0, exponent=7)
    fluid = tp.WeaklyCompressibleSPHSystem(ic, smoothing_kernel=tp.SchoenbergCubicSplineKernel(2),
                                           smoothing_length=0.36, density_calculator=density_calculator,
                                           state_equation=se,
                                           viscosity=tp.ArtificialViscosityMonaghan(alpha=0.02, beta=0.1))
    fp = adapter.fluid_params(fluid)
    for _ in range(200):
        m_a, m_b = rng.uniform(0.5, 50, 2)
        rho_a, rho_b = rng.uniform(900, 1100, 2)
        p_a, p_b = rng.uniform(-1e4, 1e5, 2)
        v_a, v_b = rng.uniform(-1, 1, 2), rng.uniform(-1, 1, 2)
        pd = rng.uniform(-0.3, 0.3, 2)
        dv1, _ = oracle.interact_pair(fp, 0, m_b, rho_a, rho_b, p_a, p_b, v_a, v_b, pd)
        dv2, _ = oracle.interact_pair(fp, 0, m_a, rho_b, rho_a, p_b, p_a, v_b, v_a, -pd)
        np.testing.assert_allclose(m_a * dv1, -m_b * dv2, rtol=4 * np.finfo(float).eps, atol=1e-13)


# test/schemes/fluid/rhs.jl:115-263: momentum / energy conservation of `interact!` run on a
# perturbed 3x3 patch through the trivial (all-pairs) neighbourhood search
@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("density_calculator", [tp.ContinuityDensity(), tp.SummationDensity()])
def test_interact_conservation(oracle, seed, density_calculator):
    rng = np.random.default_rng(seed)
    ic = _patch(rng)
    se = tp.StateEquationCole(sound_speed=10.0, reference_density=1000.0, exponent=7)
    cont = isinstance(density_calculator, tp.ContinuityDensity)
    fluid = tp.WeaklyCompressibleSPHSystem(ic, smoothing_kernel=tp.SchoenbergCubicSplineKernel(2),
                                           smoothing_length=0.36, density_calculator=density_calculator,
                                           state_equation=se)
    n = ic.nparticles
    v = np.concatenate([ic.velocity, ic.density[:, None]], axis=1) if cont else ic.velocity.copy()
    out = adapter.kick(fluid, None, ic.coordinates, v, use_grid=False)
    dv = out["dv"][:, :2]
    m = ic.mass
    # linear momentum: sum m dv = 0
    np.testing.assert_allclose((m[:, None] * dv).sum(axis=0), 0.0, atol=5e-14 * np.abs(m[:, None] * dv).max())
    # angular momentum: sum m (x cross dv) = 0
    x = ic.coordinates
    ang = (m * (x[:, 0] * dv[:, 1] - x[:, 1] * dv[:, 0]))
    assert abs(ang.sum()) <= 1e-13 * np.abs(ang).max()
    if cont:
        # total energy: sum m v.dv - sum p/rho^2 m drho = 0 (rhs.jl:220-262)
        rho, p, drho = out["density"], out["pressure"], out["dv"][:, 2]
        de_kin = (m * (ic.velocity * dv).sum(axis=1))
        de_int = p / rho**2 * drho * m
        assert abs(de_kin.sum() + de_int.sum()) <= 1e-13 * max(np.abs(de_kin).max(), np.abs(de_int).max())
    # grid search gives the same result up to summation order
    out_grid = adapter.kick(fluid, None, ic.coordinates, v, use_grid=True)
    np.testing.assert_allclose(out_grid["dv"], out["dv"], rtol=1e-12, atol=1e-12 * np.abs(out["dv"]).max())


def _continuity_wall_case(rng, density_calculator, eltype=np.float64):
    """A perturbed 5 x 5 patch split at the centre particle into a fluid and a dummy-particle wall with
    ContinuityDensity, as test/schemes/boundary/dummy_particles/rhs.jl:126-150 splits its 3 x 3 patch."""
    dx = 0.1
    ic = tp.RectangularShape(dx, (5, 5), (-2.5 * dx, -2.5 * dx), density=1000.0)
    ic.coordinates += rng.uniform(-0.3 * dx, 0.3 * dx, ic.coordinates.shape)
    ic.mass += rng.uniform(-0.1, 0.1, ic.nparticles) * ic.mass[0]
    ic.density += rng.uniform(-0.1, 0.1, ic.nparticles) * 1000.0
    ic.velocity += rng.uniform(-1.0, 1.0, ic.velocity.shape)
    c = int(np.ceil(ic.nparticles / 2))
    part = lambda sl: tp.InitialCondition(coordinates=ic.coordinates[sl].astype(eltype),
                                          velocity=ic.velocity[sl].astype(eltype), mass=ic.mass[sl].astype(eltype),
                                          density=ic.density[sl].astype(eltype),
                                          pressure=np.zeros(len(ic.mass[sl]), dtype=eltype), particle_spacing=dx)
    f_ic, w_ic = part(slice(0, c)), part(slice(c, None))
    se = tp.StateEquationCole(sound_speed=10.0, reference_density=1000.0, exponent=7)
    kernel, h = tp.SchoenbergCubicSplineKernel(2), 1.2 * dx
    fluid = tp.WeaklyCompressibleSPHSystem(f_ic, smoothing_kernel=kernel, smoothing_length=h,
                                           density_calculator=density_calculator, state_equation=se)
    model = tp.BoundaryModelDummyParticles(w_ic.density, w_ic.mass, tp.ContinuityDensity(), kernel, h,
                                           state_equation=se)
    wall = tp.WallBoundarySystem(w_ic, model)
    cont = isinstance(density_calculator, tp.ContinuityDensity)
    v_f = np.concatenate([f_ic.velocity, f_ic.density[:, None]], axis=1) if cont else f_ic.velocity.copy()
    rho_w = (w_ic.density * (1 + rng.uniform(-0.02, 0.02, w_ic.nparticles))).astype(eltype)   # the integrated density has moved
    return fluid, wall, f_ic, w_ic, v_f, rho_w


# test/schemes/boundary/dummy_particles/rhs.jl:13-308, "Fluid-BoundaryDummyContinuityDensity": with the
# continuity equation solved for the dummy particles too, fluid + wall conserve the total energy
#   sum_a m_a (v_a . dv_a + p_a / rho_a^2 drho_a) = 0
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_continuity_density_wall_conserves_total_energy(oracle, seed):
    rng = np.random.default_rng(seed)
    fluid, wall, f_ic, w_ic, v_f, rho_w = _continuity_wall_case(rng, tp.ContinuityDensity())
    v_ode = np.concatenate([v_f.reshape(-1), rho_w])
    out = adapter.kick(fluid, wall, f_ic.coordinates, v_ode, use_grid=False)
    assert wall.n_integrated_particles == wall.nparticles and wall.v_nvariables == 1 and wall.u_nvariables == 0
    np.testing.assert_array_equal(out["wall_density"], rho_w)          # current_density = the wall's rows of v
    se = wall.boundary_model.state_equation
    np.testing.assert_allclose(out["wall_pressure"], [se(r) for r in rho_w], rtol=1e-14)
    dv, drho_f, drho_w = out["dv"][:, :2], out["dv"][:, 2], out["dv_wall"]
    assert np.abs(drho_w).max() > 0
    e_f = f_ic.mass * ((f_ic.velocity * dv).sum(axis=1) + out["pressure"] / out["density"] ** 2 * drho_f)
    e_w = w_ic.mass * (out["wall_pressure"] / rho_w ** 2 * drho_w)      # v_wall = 0: no kinetic part
    scale = max(np.abs(e_f).max(), np.abs(e_w).max())
    assert abs(e_f.sum() + e_w.sum()) <= 1e-13 * scale
    # the cell-list search gives the same sums up to their order
    out_grid = adapter.kick(fluid, wall, f_ic.coordinates, v_ode, use_grid=True)
    np.testing.assert_allclose(out_grid["dv_wall"], drho_w, rtol=1e-12, atol=1e-12 * np.abs(drho_w).max())


@pytest.mark.parametrize("density_calculator", [tp.ContinuityDensity(), tp.SummationDensity()])
def test_continuity_density_wall_rhs_matches_pair_formula(oracle, density_calculator):
    """wall_boundary/rhs.jl:11-79 restated pair by pair in numpy: drho_a = sum_b [rho_a / rho_b] m_b
    (0 - v_b) . grad W(x_a - x_b) with the boundary model's kernel; the bracket only for a fluid with
    ContinuityDensity (rhs.jl:63-79)."""
    rng = np.random.default_rng(11)
    fluid, wall, f_ic, w_ic, v_f, rho_w = _continuity_wall_case(rng, density_calculator)
    cont = isinstance(density_calculator, tp.ContinuityDensity)
    v_ode = np.concatenate([v_f.reshape(-1), rho_w])
    out = adapter.kick(fluid, wall, f_ic.coordinates, v_ode, use_grid=False)
    h = float(wall.boundary_model.smoothing_length)
    R = 2 * h
    expected = np.zeros(w_ic.nparticles)
    for a in range(w_ic.nparticles):
        for b in range(f_ic.nparticles):
            pd = w_ic.coordinates[a] - f_ic.coordinates[b]
            r = np.sqrt(pd @ pd)
            if pd @ pd > R * R or r < np.sqrt(np.spacing(h * h)):
                continue
            grad = oracle.kernel_deriv_div_r(1, 2, r, h) * pd
            term = f_ic.mass[b] * ((0.0 - f_ic.velocity[b]) @ grad)
            expected[a] += (rho_w[a] / out["density"][b]) * term if cont else term
    np.testing.assert_allclose(out["dv_wall"], expected, rtol=1e-12, atol=1e-12 * np.abs(expected).max())


# test/general/density_calculator.jl:30-31: lone particle, rho === m W(0)
def test_summation_density_lone_particle(oracle):
    ic = tp.RectangularShape(0.1, (1, 1), (0.0, 0.0), density=1000.0)
    se = tp.StateEquationCole(sound_speed=10.0, reference_density=1000.0, exponent=7)
    fluid = tp.WeaklyCompressibleSPHSystem(ic, smoothing_kernel=tp.SchoenbergCubicSplineKernel(2),
                                           smoothing_length=0.12, density_calculator=tp.SummationDensity(),
                                           state_equation=se)
    out = adapter.kick(fluid, None, ic.coordinates, np.zeros((1, 2)))
    assert out["density"][0] == ic.mass[0] * oracle.kernel(CUBIC, 2, 0.0, 0.12)
    assert np.all(out["dv"] == 0.0)  # the self pair is skipped (rhs.jl:52)


# PointNeighbors semantics: inclusive d^2 <= R^2, self pair delivered, set independent of the
# cell list; lattice inputs have many exact-distance ties.
@pytest.mark.parametrize("dtype,cdtype", [(np.float64, np.float64), (np.float32, np.float32),
                                          (np.float32, np.float64)])
def test_neighbor_sets_grid_equals_bruteforce(oracle, dtype, cdtype):
    for nd, n_per_dim, factor in [(2, (12, 9), 4.0), (3, (7, 6, 5), 3.0), (2, (10, 10), 2.4)]:
        dx = 0.05
        x = tp.setups.rectangular_shape_coords(dx, n_per_dim, (0.0,) * nd, coordinates_eltype=cdtype)
        radius = float(np.dtype(dtype).type(factor) * np.dtype(dtype).type(dx))
        bi, bj = oracle.neighbor_pairs(x, x, radius, dtype=dtype, grid=False)
        gi, gj = oracle.neighbor_pairs(x, x, radius, dtype=dtype, grid=True)
        assert np.array_equal(bi, gi) and np.array_equal(bj, gj)
        assert np.all(np.isin(np.arange(x.shape[0]), bi[bi == bj]))  # self pairs present
        rng = np.random.default_rng(5)
        xj = (x + rng.uniform(-0.3 * dx, 0.3 * dx, x.shape)).astype(cdtype)
        bi, bj = oracle.neighbor_pairs(xj, x, radius, dtype=dtype, grid=False)
        gi, gj = oracle.neighbor_pairs(xj, x, radius, dtype=dtype, grid=True)
        assert np.array_equal(bi, gi) and np.array_equal(bj, gj)


def test_neighbor_counts_match_survey(oracle):
    """Interior lattice neighbour counts quoted in SURVEY.md section 8: 49 (config 1),
    21 (config 2), 123 (3-D, R = 3 dx), self included."""
    for nd, factor, expected in [(2, 4.0, 49), (2, 2.4, 21), (3, 3.0, 123)]:
        n = 11
        x = tp.setups.rectangular_shape_coords(1.0, (n,) * nd, (0.0,) * nd)
        centre = np.argmin(np.abs(x - x.mean(axis=0)).sum(axis=1))
        i, j = oracle.neighbor_pairs(x[centre:centre + 1], x, factor)
        assert len(j) == expected


def test_density_reinitialisation_properties():
    """DensityReinitializationCallback -> reinit_density! (callbacks/density_reinit.jl:83-121, wcsph/system.jl:398-415).
    The reference has no numerical test for it; the restatement is pinned by what the formula implies:
    without a wall the result depends on positions and masses only -- the integrated density is forgotten --, deep
    inside a uniform lattice the Shepard coefficient is exactly 1 and the density is the plain kernel sum, and at a free
    surface the correction raises the deficient kernel sum."""
    import numpy as np
    import trixiparticles.jl_b200 as tp
    from oracle import adapter
    n, dx = 21, 0.05
    r = (np.arange(n) + 0.5) * dx
    x = np.array([[a, b] for b in r for a in r])
    ic = tp.InitialCondition(x, np.zeros_like(x), np.full(len(x), 1000.0 * dx * dx), np.full(len(x), 1000.0),
                             np.zeros(len(x)), dx)
    fluid = tp.WeaklyCompressibleSPHSystem(
        ic, smoothing_kernel=tp.WendlandC2Kernel(2), smoothing_length=2 * dx, density_calculator=tp.ContinuityDensity(),
        state_equation=tp.StateEquationCole(sound_speed=20.0, reference_density=1000.0, exponent=7))
    rng = np.random.default_rng(0)
    v1 = np.concatenate([np.zeros_like(x), np.full((len(x), 1), 1000.0)], axis=1)
    v2 = v1.copy()
    v2[:, 2] *= 1 + 0.05 * rng.uniform(-1, 1, len(x))
    a, b = adapter.reinit_density(fluid, None, x, v1), adapter.reinit_density(fluid, None, x, v2)
    assert np.array_equal(a, b)                                   # the integrated density is forgotten
    from oracle import oracle as O
    mid = (n // 2) * n + n // 2
    w = sum(1000.0 * dx * dx * O.kernel(O.KERNEL_WENDLAND_C2, 2, float(np.linalg.norm(x[mid] - xb)), 2 * dx)
            for xb in x if np.linalg.norm(x[mid] - xb) < 4 * dx)
    assert abs(a[mid] - w) <= 1e-9 * w and abs(a[mid] - 1000.0) < 5.0     # c = 1: the plain kernel sum
    corner = 0
    plain = sum(1000.0 * dx * dx * O.kernel(O.KERNEL_WENDLAND_C2, 2, float(np.linalg.norm(x[corner] - xb)), 2 * dx)
                for xb in x if np.linalg.norm(x[corner] - xb) < 4 * dx)
    assert plain < 500.0 and plain + 100.0 < a[corner] < 1000.0   # a quarter of the support is filled; corrected upwards
