"""Host-side point / line interpolation (trixiparticles.jl_b200/interpolation.py) against the
known answers of the reference's own test, /root/reference/test/general/interpolation.jl:44-262
(2-D patch of 10 x 10 fluid particles, cubic spline, h = 0.12, "no boundary" cases)."""
import numpy as np
import pytest

import trixiparticles.jl_b200 as tp
from trixiparticles.jl_b200.interpolation import interpolate_line, interpolate_points


TOL = 5e-4   # `compare_interpolation_result(...; tolerance=5e-4)`, interpolation.jl:2-18; counts and coordinates exact


def rectangular_patch(dx, size, offset, set_values):
    """test/rectangular_patch.jl without perturbation: lattice centred at the origin, the
    particle number ceil(n / 2) moved to the origin, everything shifted by `offset`."""
    ic = tp.RectangularShape(dx, size, tuple(-dx / 2 * s for s in size), density=1000.0)
    x = ic.coordinates.copy()
    x[int(np.ceil(np.prod(size) / 2)) - 1] = 0.0
    x += np.asarray(offset)[None, :]
    m, rho, p, v = zip(*(set_values(c) for c in x))
    return tp.InitialCondition(coordinates=x, velocity=np.array(v), mass=np.array(m), density=np.array(rho),
                               pressure=np.array(p), particle_spacing=dx)


@pytest.fixture(scope="module")
def patch():
    dx, nx, ny = 0.2, 10, 10

    def set_values(c):
        return 1.0, 666.0, 2 * c[1], (5 + 3 * c[0], 0.1 * c[1] ** 2 + 0.1)

    ic = rectangular_patch(dx, (nx, ny), (0.0, ny * 0.5 * dx), set_values)
    eos = tp.StateEquationCole(sound_speed=10 * np.sqrt(9.81 * 0.9), reference_density=1000.0, exponent=7)
    fluid = tp.WeaklyCompressibleSPHSystem(ic, smoothing_kernel=tp.SchoenbergCubicSplineKernel(2),
                                           smoothing_length=1.2 * 0.5 * dx,
                                           density_calculator=tp.ContinuityDensity(), state_equation=eos,
                                           viscosity=tp.ArtificialViscosityMonaghan(alpha=0.02, beta=0.0),
                                           acceleration=(0.0, -9.81))
    semi = tp.Semidiscretization(fluid)
    u = ic.coordinates.reshape(-1).copy()
    v = np.concatenate([ic.velocity, ic.density[:, None]], axis=1).reshape(-1)
    return semi, fluid, v, u, ic.pressure.copy(), dx, ny


@pytest.mark.parametrize("cut_off_bnd", [True, False])
def test_points_match_reference_known_answers(patch, cut_off_bnd):
    semi, fluid, v, u, p, dx, ny = patch
    kw = dict(cut_off_bnd=cut_off_bnd, pressure=p)      # the reference test overwrites system.pressure
    at = lambda y: interpolate_points(np.array([[0.0], [y]]), semi, fluid, v, u, **kw)

    def binary_search_outside(inside, outside, tol=1e-5):
        start = inside
        while abs(outside - inside) > tol:
            mid = (inside + outside) / 2
            if at(mid)["neighbor_count"][0] == 0:
                outside = mid
            else:
                inside = mid
        return abs(outside - start)

    # interpolation.jl:147-150, :187-189
    assert binary_search_outside(ny * dx, (ny + 2) * dx) == pytest.approx(0.11817626953124982, abs=1e-14)
    assert binary_search_outside(0.0, -2 * dx) == pytest.approx(0.11817626953125, abs=1e-14)
    res = at(ny * dx + 0.11817626953124982)                # :152-155 nothing in reach: NaN, count 0
    assert res["neighbor_count"][0] == 0 and np.isnan(res["density"][0]) and np.isnan(res["pressure"][0])
    # :203-217 three points
    res = interpolate_points(np.array([[0.0, 0.0, 0.0], [0.0, 0.5, 1.0]]), semi, fluid, v, u, **kw)
    assert res["neighbor_count"].tolist() == [2, 6, 5]
    assert res["density"] == pytest.approx([666.0, 666.0000000000001, 666.0], abs=TOL)
    assert res["velocity"][0] == pytest.approx([5.0, 5.0, 5.0], abs=TOL)
    assert res["velocity"][1] == pytest.approx([0.101, 0.125, 0.20035665520692278], abs=TOL)
    assert res["pressure"] == pytest.approx([0.19999999999999996, 1.0000000000000002, 2.0], abs=TOL)


def test_line_matches_reference_known_answers(patch):
    semi, fluid, v, u, p, dx, ny = patch
    # interpolation.jl:221-256
    res = interpolate_line([1.0, -0.05], [1.0, 1.0], 5, semi, fluid, v, u, endpoint=True, pressure=p)
    assert res["neighbor_count"].tolist() == [1, 2, 2, 1, 1]
    assert res["point_coords"][1] == pytest.approx([-0.05, 0.2125, 0.475, 0.7375, 1.0], abs=1e-15)
    assert res["density"] == pytest.approx([666.0] * 5, abs=TOL)
    assert res["velocity"][0] == pytest.approx([7.7] * 5, abs=TOL)
    assert res["velocity"][1] == pytest.approx([0.10100000000000002, 0.10605429538320173, 0.12465095587703466,
                                                0.14900000000000002, 0.22100000000000006], abs=TOL)
    assert res["pressure"] == pytest.approx([0.19999999999999998, 0.4527147691600855, 0.9912738969258663,
                                             1.4000000000000001, 2.2], abs=TOL)
    inner = interpolate_line([1.0, -0.05], [1.0, 1.0], 5, semi, fluid, v, u, endpoint=False, pressure=p)
    assert inner["neighbor_count"].tolist() == [2, 2, 1]
    assert inner["pressure"] == pytest.approx([0.4527147691600855, 0.9912738969258665, 1.4000000000000001], abs=TOL)


def test_default_pressure_is_the_state_equation(patch):
    semi, fluid, v, u, p, dx, ny = patch
    res = interpolate_points(np.array([[0.0], [1.0]]), semi, fluid, v, u, cut_off_bnd=False)
    assert res["pressure"][0] == pytest.approx(float(fluid.state_equation(666.0)), rel=1e-12)
    clipped = interpolate_points(np.array([[0.0], [1.0]]), semi, fluid, v, u, cut_off_bnd=False,
                                 clip_negative_pressure=True)
    assert clipped["pressure"][0] == 0.0          # rho < rho_0: negative pressure, clipped per particle
