"""Point / line interpolation of fluid fields (post-processing, host side).

Mirror of `interpolate_points` / `interpolate_line` of the reference
(/root/reference/src/general/interpolation.jl:392-410 `interpolate_line`, :478-660
`interpolate_points`, :703-720 `interpolate_system!`) for the systems of the accelerated path:
a `WeaklyCompressibleSPHSystem` as reference system, optionally a `WallBoundarySystem` for the
`cut_off_bnd` test.  Every point gets the Shepard-normalised kernel average

    f(r_a) = sum_b f_b V_b W(r_a - r_b, h) / sum_b V_b W(r_a - r_b, h),     V_b = m_b / rho_b

over the fluid particles within the compact support of the kernel at the *given* smoothing length
(the validation sensors use twice the simulation's, validation/dam_break_2d/sensors.jl:11-18).
`W` is the safe kernel (`kernel(...)`: zero for r >= R, smoothing_kernels.jl:30-34).

This is a callback-time operation (a few points every 0.0025 s of simulated time), not part of
the right-hand side: it runs in numpy on host copies of the ODE vectors (a KD-tree query per
call), exactly like the reference initialises a fresh CPU neighbourhood search for it when the
smoothing length differs (interpolation.jl:512-546).
"""
from __future__ import annotations

import numpy as np

from .model import (SchoenbergCubicSplineKernel, SchoenbergQuarticSplineKernel,
                    SchoenbergQuinticSplineKernel, WallBoundarySystem, WeaklyCompressibleSPHSystem,
                    WendlandC2Kernel, WendlandC4Kernel, WendlandC6Kernel, compact_support)


def kernel(smoothing_kernel, r, h):
    """Safe kernel value W(r, h) (smoothing_kernels.jl:30-34, :193-214, :436-446), vectorised."""
    r = np.asarray(r, dtype=np.float64)
    nd = smoothing_kernel.ndims
    q = r / h
    if isinstance(smoothing_kernel, WendlandC2Kernel):
        sigma = {2: 7 / (4 * np.pi), 3: 21 / (16 * np.pi)}[nd]
        w = sigma / h ** nd * (1 - q / 2) ** 4 * (2 * q + 1)
    elif isinstance(smoothing_kernel, WendlandC4Kernel):
        sigma = {2: 9 / (4 * np.pi), 3: 495 / (256 * np.pi)}[nd]
        w = sigma / h ** nd * (1 - q / 2) ** 6 * (35 * q ** 2 / 12 + 3 * q + 1)
    elif isinstance(smoothing_kernel, WendlandC6Kernel):
        sigma = {2: 39 / (14 * np.pi), 3: 1365 / (512 * np.pi)}[nd]
        w = sigma / h ** nd * (1 - q / 2) ** 8 * (4 * q ** 3 + 25 * q ** 2 / 4 + 4 * q + 1)
    elif isinstance(smoothing_kernel, SchoenbergQuarticSplineKernel):
        sigma = {2: 96 / (1199 * np.pi), 3: 1 / (20 * np.pi)}[nd]
        w = sigma / h ** nd * ((2.5 - q) ** 4 - 5 * np.where(q < 1.5, (1.5 - q) ** 4, 0.0)
                               + 10 * np.where(q < 0.5, (0.5 - q) ** 4, 0.0))
    elif isinstance(smoothing_kernel, SchoenbergQuinticSplineKernel):
        sigma = {2: 7 / (478 * np.pi), 3: 1 / (120 * np.pi)}[nd]
        w = sigma / h ** nd * ((3 - q) ** 5 - 6 * np.where(q < 2, (2 - q) ** 5, 0.0)
                               + 15 * np.where(q < 1, (1 - q) ** 5, 0.0))
    elif isinstance(smoothing_kernel, SchoenbergCubicSplineKernel):
        sigma = {2: 10 / (7 * np.pi), 3: 1 / np.pi}[nd]
        w = sigma / h ** nd * ((2 - q) ** 3 / 4 - np.where(q < 1, (1 - q) ** 3, 0.0))
    else:
        raise ValueError(f"unsupported smoothing kernel {smoothing_kernel!r}")
    return np.where(r < compact_support(smoothing_kernel, h), w, 0.0)


def _host(x):
    return x if isinstance(x, np.ndarray) else x.detach().cpu().numpy()


def interpolate_points(point_coords, semi, ref_system, v_ode, u_ode, *, smoothing_length=None,
                       cut_off_bnd=True, clip_negative_pressure=False, pressure=None):
    """`point_coords`: (ND, n_points) as in the reference (column i = point i).  Returns a dict
    with `computed_density`, `neighbor_count`, `point_coords`, `velocity` (ND, n), `pressure`,
    `density` -- NaN where the point has no fluid (or, with `cut_off_bnd`, more wall than fluid).
    `pressure`: per-particle fluid pressure (the reference reads `system.pressure`, which
    `update_pressure!` fills with the state equation at the current density -- the default here)."""
    if not isinstance(ref_system, WeaklyCompressibleSPHSystem):
        raise ValueError("interpolation needs a WeaklyCompressibleSPHSystem as reference system")
    from scipy.spatial import cKDTree
    nd = ref_system.ndims
    pts = np.asarray(point_coords, dtype=np.float64).reshape(nd, -1).T
    n = len(pts)
    h = float(ref_system.smoothing_length if smoothing_length is None else smoothing_length)
    R = compact_support(ref_system.smoothing_kernel, h)
    u = np.asarray(semi.wrap_u(_host(u_ode), ref_system), dtype=np.float64).reshape(-1, nd)
    v = np.asarray(semi.wrap_v(_host(v_ode), ref_system), dtype=np.float64).reshape(u.shape[0], -1)
    if v.shape[1] > nd:                      # ContinuityDensity: rho is the last row of v
        rho = v[:, nd]
    else:                                    # SummationDensity: the last kick's cache.density
        rho = np.asarray(semi.system_field(ref_system, "density"), dtype=np.float64)
    if pressure is not None:
        p = np.asarray(pressure, dtype=np.float64).copy()
    else:
        eos = ref_system.state_equation
        p = np.array([float(eos(r)) for r in rho]) if len(rho) < 4096 else _eos_vec(eos, rho)
    if clip_negative_pressure:
        p = np.maximum(p, 0.0)
    mass = np.asarray(ref_system.mass, dtype=np.float64)

    computed_density = np.zeros(n)
    other_density = np.zeros(n)
    shepard = np.zeros(n)
    count = np.zeros(n, dtype=np.int64)
    vel = np.zeros((nd, n))
    pres = np.zeros(n)
    dens = np.zeros(n)
    tree = cKDTree(u)
    for a, nbrs in enumerate(tree.query_ball_point(pts, R)):
        if not nbrs:
            continue
        nbrs = np.sort(np.asarray(nbrs))
        d = np.sqrt(((pts[a] - u[nbrs]) ** 2).sum(axis=1))
        w = kernel(ref_system.smoothing_kernel, d, h)
        vol = mass[nbrs] / rho[nbrs]
        computed_density[a] = (mass[nbrs] * w).sum()
        shepard[a] = (vol * w).sum()
        vel[:, a] = (v[nbrs, :nd] * (vol * w)[:, None]).sum(axis=0)
        pres[a] = (p[nbrs] * vol * w).sum()
        dens[a] = (rho[nbrs] * vol * w).sum()
        count[a] = len(nbrs)
    if cut_off_bnd:
        for wall in semi.systems:
            if not isinstance(wall, WallBoundarySystem):
                continue
            xw = np.asarray(wall.coordinates, dtype=np.float64)
            mw = np.asarray(wall.boundary_model.hydrodynamic_mass, dtype=np.float64)
            wtree = cKDTree(xw)
            for a, nbrs in enumerate(wtree.query_ball_point(pts, R)):
                if nbrs:
                    nbrs = np.asarray(nbrs)
                    d = np.sqrt(((pts[a] - xw[nbrs]) ** 2).sum(axis=1))
                    other_density[a] += (mw[nbrs] * kernel(ref_system.smoothing_kernel, d, h)).sum()
                    count[a] += len(nbrs)
    cut = (computed_density < np.finfo(np.float64).eps) | (cut_off_bnd & (other_density > computed_density))
    with np.errstate(invalid="ignore", divide="ignore"):
        vel = np.where(cut, np.nan, vel / shepard)
        pres = np.where(cut, np.nan, pres / shepard)
        dens = np.where(cut, np.nan, dens / shepard)
    computed_density = np.where(cut, np.nan, computed_density)
    count = np.where(cut, 0, count)
    return dict(computed_density=computed_density, point_coords=pts.T.copy(), neighbor_count=count,
                velocity=vel, pressure=pres, density=dens)


def _eos_vec(eos, rho):
    B = eos.reference_density * eos.sound_speed ** 2 / eos.exponent
    p = B * ((rho / eos.reference_density) ** eos.exponent - 1.0) + eos.background_pressure
    return np.maximum(p, 0.0) if eos.clip_negative_pressure else p


def interpolate_line(start, end_, n_points, semi, ref_system, v_ode, u_ode, *, endpoint=True,
                     smoothing_length=None, cut_off_bnd=True, clip_negative_pressure=False, pressure=None):
    """`n_points` equidistant points from `start` to `end_` (interpolation.jl:392-410);
    `endpoint=False` drops the first and the last point."""
    start, end_ = np.asarray(start, dtype=np.float64), np.asarray(end_, dtype=np.float64)
    pts = np.linspace(start, end_, int(n_points))
    if not endpoint:
        pts = pts[1:-1]
    return interpolate_points(pts.T, semi, ref_system, v_ode, u_ode, smoothing_length=smoothing_length,
                              cut_off_bnd=cut_off_bnd, clip_negative_pressure=clip_negative_pressure,
                              pressure=pressure)


def interpolated_pressure(coord_top, coord_bottom):
    """The validation pressure sensor (validation/dam_break_2d/sensors.jl:7-18) as a
    `PostprocessCallback` function: mean over 10 points of the line, NaN counted as 0,
    smoothing length 2 h, negative pressures clipped, no boundary cut-off."""
    def probe(system, v_ode, u_ode, semi, t):
        n_interpolation_points = 10
        res = interpolate_line(coord_top, coord_bottom, n_interpolation_points, semi, system, v_ode, u_ode,
                               smoothing_length=2.0 * float(system.smoothing_length),
                               clip_negative_pressure=True, cut_off_bnd=False)
        return float(np.nansum(res["pressure"]) / n_interpolation_points)
    return probe
