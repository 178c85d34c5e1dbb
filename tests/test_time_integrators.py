"""Host-side checks of the time integrators behind `solve` (no GPU): the verified coefficients of
RDPK3SpFSAL35, its controller on a toy problem, SymplecticPositionVerlet's density update, the
gravity-free `calculate_dt`, the FSI layout and the TLSPH step size."""
import math
from types import SimpleNamespace

import numpy as np
import pytest

import trixiparticles.jl_b200 as tp
from trixiparticles.jl_b200 import examples
from trixiparticles.jl_b200.semidiscretization import DynamicalODEProblem
from trixiparticles.jl_b200.time_integration import (RDPK3SpFSAL35, StepsizeCallback, SymplecticPositionVerlet,
                                                     calculate_dt, calculate_dt_structure, solve)


def test_rdpk3spfsal35_order_conditions():
    """The low-storage coefficients reproduce their own abscissae and satisfy the four conditions
    of a third-order Runge-Kutta method; the embedded weights are second-order consistent."""
    A, b = RDPK3SpFSAL35.butcher()
    c = A.sum(axis=1)
    assert np.allclose(c, RDPK3SpFSAL35.c, rtol=0, atol=2e-16)
    assert np.all(np.triu(A) == 0)                     # explicit
    for got, want in ((b.sum(), 1.0), (b @ c, 0.5), (b @ c ** 2, 1 / 3), (b @ (A @ c), 1 / 6)):
        assert abs(got - want) < 5e-16
    e = np.array(RDPK3SpFSAL35.error_weights())
    c_all = np.append(c, 1.0)
    assert abs(e.sum()) < 1e-15 and abs(e @ c_all) < 1e-15
    assert abs(e @ c_all ** 2) > 1e-3                  # ... and not third-order: a usable error estimate


class _ToySemi:
    """Stands in for a Semidiscretization with host ODE vectors (numpy): one 'system' of n particles,
    u = position (1 entry), v = (velocity, density)."""
    def __init__(self, n, with_density):
        self.parallelization_backend = SimpleNamespace(ode_memory="host")
        nv = 2 if with_density else 1
        self.system = SimpleNamespace(ndims=1, v_nvariables=nv)
        self.systems = (self.system,)
        self.ranges_v = ((0, nv * n),)
        self.ranges_u = ((0, n),)

    def system_index(self, s):
        return 0


def oscillator(n=3, with_density=False):
    """u'' = -omega^2 u per particle; optional density row with drho/dt = -rho * a (exact: rho0 exp(-a t))."""
    omega = np.linspace(1.0, 2.0, n)
    rate = 0.3
    nv = 2 if with_density else 1
    semi = _ToySemi(n, with_density)

    def f1(dv, v, u, p, t):
        dv.reshape(n, nv)[:, 0] = -omega ** 2 * u
        if with_density:
            dv.reshape(n, nv)[:, 1] = -rate * v.reshape(n, nv)[:, 1]
        return dv

    def f2(du, v, u, p, t):
        du[:] = v.reshape(n, nv)[:, 0]
        return du

    v0 = np.zeros(n * nv)
    if with_density:
        v0.reshape(n, nv)[:, 1] = 1000.0
    u0 = np.ones(n)
    return DynamicalODEProblem(f1, f2, v0, u0, (0.0, 2.0), SimpleNamespace(semi=semi)), omega, rate


def test_rdpk3spfsal35_adaptive_solve_on_oscillator():
    errs, steps = [], []
    for tol in (1e-4, 1e-6, 1e-8):
        ode, omega, _ = oscillator()
        sol = solve(ode, RDPK3SpFSAL35(), abstol=tol, reltol=tol)
        assert sol.retcode == "Success" and sol.t == 2.0
        errs.append(np.abs(sol.u - np.cos(omega * 2.0)).max())
        steps.append(sol.nsteps)
    assert errs[0] < 2e-3 and errs[1] < errs[0] / 20 and errs[2] < errs[1] / 20    # third order: 100x tol -> ~21x dt
    assert steps[0] < steps[1] < steps[2]
    # a fixed initial step far too large is rejected, not accepted
    ode, omega, _ = oscillator()
    sol = solve(ode, RDPK3SpFSAL35(), abstol=1e-6, reltol=1e-6, dt=1.5)
    assert sol.nrejected >= 1 and np.abs(sol.u - np.cos(omega * 2.0)).max() < 1e-4
    # dtmax caps the step
    ode, _, _ = oscillator()
    sol = solve(ode, RDPK3SpFSAL35(), abstol=1e-2, reltol=1e-2, dtmax=0.01, save_everystep=True)
    assert max(sol.dts) <= 0.01 + 1e-15 and sol.nsteps >= 200
    ode, _, _ = oscillator()
    assert solve(ode, RDPK3SpFSAL35(), abstol=1e-8, reltol=1e-8, maxiters=5).retcode == "MaxIters"


def test_symplectic_position_verlet_on_oscillator():
    """Second order in dt for position / velocity; the density follows
    rho_prev (2 - eps) / (2 + eps) with eps = -(drho / rho_half) dt -- for drho = -a rho this is the
    (1,1) Pade step exp(-a dt) = (2 - a dt) / (2 + a dt)."""
    errs = []
    for dt in (0.02, 0.01):
        ode, omega, rate = oscillator(with_density=True)
        sol = solve(ode, SymplecticPositionVerlet(), dt=dt)
        assert sol.retcode == "Success" and sol.nsteps == round(2.0 / dt)
        errs.append(np.abs(sol.u - np.cos(omega * 2.0)).max())
        rho = sol.v.reshape(-1, 2)[:, 1]
        pade = 1000.0 * ((2 - rate * dt) / (2 + rate * dt)) ** sol.nsteps
        assert np.allclose(rho, pade, rtol=1e-12)
    assert 3.5 < errs[0] / errs[1] < 4.5


def test_calculate_dt_without_gravity():
    """fluid.jl:199-239: sqrt(h / 0) = Inf, the other limits decide (no ZeroDivisionError)."""
    ic = tp.RectangularShape(0.05, (4, 4), (0.0, 0.0), density=1000.0)
    se = tp.StateEquationCole(sound_speed=10.0, reference_density=1000.0, exponent=7)
    fluid = tp.WeaklyCompressibleSPHSystem(ic, smoothing_kernel=tp.SchoenbergCubicSplineKernel(2), smoothing_length=0.06,
                                           density_calculator=tp.ContinuityDensity(), state_equation=se)
    assert calculate_dt(fluid, 0.9) == pytest.approx(0.9 * 0.06 / 10.0)
    fluid_v = tp.WeaklyCompressibleSPHSystem(ic, smoothing_kernel=tp.SchoenbergCubicSplineKernel(2), smoothing_length=0.06,
                                             density_calculator=tp.ContinuityDensity(), state_equation=se,
                                             viscosity=tp.ArtificialViscosityMonaghan(alpha=100.0))
    nu = 100.0 * 0.06 * 10.0 / 8
    assert calculate_dt(fluid_v, 0.9) == pytest.approx(0.125 * 0.06 ** 2 / nu)


def test_fsi_example_layout_and_step_size():
    """examples/fsi/dam_break_plate_2d.jl with the dimensions of test/examples/gpu.jl:664-726."""
    fluid, wall, structure, _ = examples.dam_break_plate_2d(0.01, initial_fluid_size=(0.15, 0.29),
                                                            eltype=np.float32, coordinates_eltype=np.float32)
    assert (fluid.nparticles, structure.nparticles, structure.n_integrated_particles) == (435, 140, 135)
    # the clamped particles (first five of the union) sit at the end (system.jl:131-147)
    assert np.allclose(structure.initial_coordinates[-5:, 1], 0.0) and np.all(structure.initial_coordinates[:135, 1] > 0)
    assert np.allclose(structure.mass, 2500 * 0.003 ** 2)
    assert structure.boundary_model.hydrodynamic_mass.shape == (140,)
    semi = tp.Semidiscretization(fluid, wall, structure)
    assert semi.ranges_u == ((0, 870), (870, 870), (870, 870 + 270))
    assert semi.ranges_v == ((0, 1305), (1305, 1305), (1305, 1305 + 270))
    # system.jl:701-717: cfl h / sqrt(K / rho), K = E / (ND (1 - 2 nu))
    h = math.sqrt(2) * 0.003
    assert calculate_dt_structure(structure, 1.2) == pytest.approx(1.2 * h / math.sqrt(1e6 / 2 / 2500), rel=1e-6)
    assert StepsizeCallback(cfl=1.2).dt(semi) == pytest.approx(calculate_dt_structure(structure, 1.2))
    with pytest.raises(ValueError):
        tp.TotalLagrangianSPHSystem(structure.initial_condition, smoothing_kernel=tp.WendlandC2Kernel(3),
                                    smoothing_length=h, young_modulus=1e6, poisson_ratio=0.0)
