"""Runs the CPU oracle on the package's system objects (TEST INFRASTRUCTURE ONLY).

Converts `WeaklyCompressibleSPHSystem` / `WallBoundarySystem` records into the oracle's
parameter structs; used by tests/, smoke() and bench.py's cpu_baseline leg."""
from __future__ import annotations

import numpy as np

from . import oracle as O


def fluid_params(fluid) -> O.FluidParams:
    se = fluid.state_equation
    t = np.dtype(fluid.eltype).type
    p = O.FluidParams()
    p.ndims = fluid.ndims
    p.kernel = fluid.smoothing_kernel.kernel_id
    p.density_calculator = fluid.density_calculator.density_id
    p.clip_negative_pressure = int(se.clip_negative_pressure)
    p.has_viscosity = 0 if fluid.viscosity is None else int(getattr(fluid.viscosity, "viscosity_id", 1))
    p.has_diffusion = int(fluid.density_diffusion is not None)
    p.smoothing_length = float(t(fluid.smoothing_length))
    p.sound_speed = float(t(se.sound_speed))
    p.exponent = float(t(se.exponent))
    p.reference_density = float(t(se.reference_density))
    p.background_pressure = float(t(se.background_pressure))
    if fluid.viscosity is not None:
        if p.has_viscosity == 1:
            p.alpha = float(t(fluid.viscosity.alpha))
            p.beta = float(t(fluid.viscosity.beta))
        else:
            p.alpha = float(t(fluid.viscosity.nu))
            p.beta = 0.0
        p.epsilon = float(t(fluid.viscosity.epsilon))
    if fluid.density_diffusion is not None:
        p.delta = float(t(fluid.density_diffusion.delta))
    for d in range(fluid.ndims):
        p.acceleration[d] = float(fluid.acceleration[d])
    st = getattr(fluid, "source_terms", None)
    if st is not None:
        p.damping_coefficient = float(t(st.damping_coefficient))
    return p


def wall_params(wall) -> O.WallParams:
    m = wall.boundary_model
    se = m.state_equation
    t = np.dtype(wall.eltype).type
    p = O.WallParams()
    p.kernel = m.smoothing_kernel.kernel_id
    p.clip_negative_pressure = int(m.clip_negative_pressure)
    p.smoothing_length = float(t(m.smoothing_length))
    p.sound_speed = float(t(se.sound_speed))
    p.exponent = float(t(se.exponent))
    p.reference_density = float(t(se.reference_density))
    p.background_pressure = float(t(se.background_pressure))
    p.eos_clip_negative_pressure = int(getattr(se, "clip_negative_pressure", False))
    if type(m.density_calculator).__name__ == "ContinuityDensity":
        p.density_calculator = O.WALL_CONTINUITY
    else:
        p.pressure_offset = float(t(m.density_calculator.pressure_offset))
    visc = getattr(m, "viscosity", None)
    if visc is not None:   # no-slip wall
        p.has_viscosity = int(getattr(visc, "viscosity_id", 1))
        if p.has_viscosity == 1:
            p.visc_alpha, p.visc_beta = float(t(visc.alpha)), float(t(visc.beta))
        else:
            p.visc_alpha = float(t(visc.nu))
        p.visc_epsilon = float(t(visc.epsilon))
    return p


def _f32_bits(x) -> int:
    return int(np.array([x], dtype=np.float32).view(np.int32)[0])


def kick(fluid, wall, u, v, use_grid=True, nthreads=0, fluid_wall_interaction=True):
    """Oracle `kick!` for Semidiscretization(fluid[, wall]); u (n, ND), v (n, nv)."""
    se = fluid.state_equation
    adaptive = hasattr(se, "update_speed_of_sound")
    if adaptive:
        # update_speed_of_sound! (wcsph/system.jl:307-321) comes first in the reference's kick!;
        # the C oracle then sees a Cole equation with that speed of sound
        se.update_speed_of_sound(np.asarray(v)[:, :fluid.ndims], fluid.eltype)
    fp = fluid_params(fluid)
    mixed = adaptive and se.param_eltype == np.float32 and np.dtype(fluid.eltype) != np.float32
    if mixed:   # B formed in Float32 by the state equation itself
        fp.reserved0, fp.reserved1 = _f32_bits(se._B()), 1
    if wall is not None and fluid_wall_interaction:
        wp = wall_params(wall)
        if mixed and wall.boundary_model.state_equation is se:
            wp.reserved0 = _f32_bits(se._B())
        cw, mw = wall.coordinates, wall.boundary_model.hydrodynamic_mass
    else:
        wp, cw, mw = None, None, None
    v_wall = None
    if wp is not None and wp.density_calculator == O.WALL_CONTINUITY:
        v = np.asarray(v)
        nv = fluid.v_nvariables
        if v.ndim == 1:      # the flat ODE vector [fluid | wall] (semidiscretization.jl:128-135)
            v_wall = v[fluid.nparticles * nv:]
            v = v[: fluid.nparticles * nv].reshape(-1, nv)
        else:                # fluid rows only: the wall at its initial density
            v_wall = np.asarray(wall.boundary_model.initial_density, dtype=fluid.eltype)
    return O.kick(fp, wp, fluid.mass, cw, mw, v, u, fluid.eltype, use_grid=use_grid,
                  nthreads=nthreads, v_wall=v_wall)


def reinit_density(fluid, wall, u, v):
    """Oracle `reinit_density!` (DensityReinitializationCallback) for Semidiscretization(fluid[, wall])."""
    fp = fluid_params(fluid)
    wp = wall_params(wall) if wall is not None else None
    return O.reinit_density(fp, wp, fluid.mass, wall.coordinates if wall is not None else None,
                            wall.boundary_model.hydrodynamic_mass if wall is not None else None, v, u, fluid.eltype)


def structure_params(st) -> O.TlsphParams:
    t = np.dtype(st.eltype).type
    p = O.TlsphParams()
    p.ndims = st.ndims
    p.kernel = st.smoothing_kernel.kernel_id
    p.smoothing_length = float(t(st.smoothing_length))
    p.young_modulus, p.poisson_ratio = float(t(st.young_modulus)), float(t(st.poisson_ratio))
    if st.penalty_force is not None:
        p.has_penalty, p.penalty_alpha = 1, float(t(st.penalty_force.alpha))
    for d in range(st.ndims):
        p.acceleration[d] = float(st.acceleration[d])
    m = st.boundary_model
    if m is not None and type(m).__name__ == "BoundaryModelDummyParticles":
        se = m.state_equation
        p.boundary_model = O.BOUNDARY_DUMMY_PARTICLES
        p.bm_kernel = m.smoothing_kernel.kernel_id
        p.bm_clip_negative_pressure = int(m.clip_negative_pressure)
        p.bm_smoothing_length = float(t(m.smoothing_length))
        p.bm_sound_speed, p.bm_exponent = float(t(se.sound_speed)), float(t(se.exponent))
        p.bm_reference_density = float(t(se.reference_density))
        p.bm_background_pressure = float(t(se.background_pressure))
        p.bm_pressure_offset = float(t(m.density_calculator.pressure_offset))
        p.bm_bernoulli_factor = float(t(getattr(m.density_calculator, "factor", 0.0)))
    elif m is not None:
        p.boundary_model = O.BOUNDARY_MONAGHAN_KAJTAR
        p.mk_K, p.mk_beta, p.mk_spacing = float(t(m.K)), float(t(m.beta)), float(t(m.boundary_particle_spacing))
    return p


def kick_fsi(fluid, wall, structure, u_ode, v_ode, nthreads=0):
    """Oracle `kick!` of Semidiscretization(fluid, wall, structure) on flat ODE vectors laid out
    [fluid | structure].  Returns dict(dv, F, pk1_rho2, L)."""
    fp = fluid_params(fluid)
    wp = wall_params(wall) if wall is not None else None
    sp = structure_params(structure)
    hyd = (structure.boundary_model.hydrodynamic_mass if structure.boundary_model is not None
           else np.zeros(structure.nparticles, dtype=structure.eltype))
    L = O.tlsph_correction_matrix(sp, structure.initial_coordinates, structure.mass, structure.material_density,
                                  structure.eltype)
    out = O.kick_fsi(fp, wp, sp, fluid.mass, wall.coordinates if wall is not None else None,
                     wall.boundary_model.hydrodynamic_mass if wall is not None else None,
                     structure.n_integrated_particles, structure.initial_coordinates, structure.mass,
                     structure.material_density, hyd, L, v_ode, u_ode, fluid.eltype, nthreads=nthreads,
                     **_clamped_state(structure))
    out["L"] = L
    return out


def _clamped_state(system):
    """The prescribed state of a system's clamped particles as left by its last apply_prescribed_motion(t)."""
    if getattr(system, "prescribed_motion", None) is None:
        return {}
    moving = bool(system.ismoving)
    return dict(clamped_coords=system.clamped_coordinates,
                clamped_velocity=system.clamped_velocity if moving else None,
                clamped_acceleration=system.clamped_acceleration if moving else None)


def kick_moving_wall(fluid, wall, u, v, static_wall=None, nthreads=0):
    """Oracle `kick!` of Semidiscretization(fluid, [static_wall,] wall) where `wall` is a
    `WallBoundarySystem(...; prescribed_motion)` in the state of its last apply_prescribed_motion(t): the FSI
    oracle with the wall as an all-clamped system of dummy particles with wall semantics (orc_kick_fsi3).
    Returns dict(dv (n_f, nv), pressure_wall, density_wall)."""
    t = np.dtype(fluid.eltype).type
    nd, n = fluid.ndims, wall.nparticles
    m, se = wall.boundary_model, wall.boundary_model.state_equation
    sp = O.TlsphParams()
    sp.ndims = nd
    sp.kernel = fluid.smoothing_kernel.kernel_id
    sp.smoothing_length = float(t(fluid.smoothing_length))
    sp.young_modulus, sp.poisson_ratio = 1.0, 0.0
    sp.boundary_model = O.BOUNDARY_DUMMY_PARTICLES
    sp.bm_kernel = m.smoothing_kernel.kernel_id
    sp.bm_clip_negative_pressure = int(m.clip_negative_pressure)
    sp.bm_smoothing_length = float(t(m.smoothing_length))
    sp.bm_sound_speed, sp.bm_exponent = float(t(se.sound_speed)), float(t(se.exponent))
    sp.bm_reference_density = float(t(se.reference_density))
    sp.bm_background_pressure = float(t(se.background_pressure))
    sp.bm_pressure_offset = float(t(m.density_calculator.pressure_offset))
    sp.bm_wall_semantics = 1
    sp.bm_bernoulli_factor = float(t(getattr(m.density_calculator, "factor", 0.0)))
    L = np.tile(np.eye(nd, dtype=fluid.eltype), (n, 1, 1))
    x0 = np.asarray(wall.initial_condition.coordinates)
    out = O.kick_fsi(fluid_params(fluid), wall_params(static_wall) if static_wall is not None else None, sp, fluid.mass,
                     static_wall.coordinates if static_wall is not None else None,
                     static_wall.boundary_model.hydrodynamic_mass if static_wall is not None else None,
                     0, x0, m.hydrodynamic_mass, np.ones(n, dtype=fluid.eltype), m.hydrodynamic_mass, L,
                     np.ascontiguousarray(v).reshape(-1), np.ascontiguousarray(u).reshape(-1), fluid.eltype,
                     nthreads=nthreads, **_clamped_state(wall))
    return dict(dv=out["dv"].reshape(np.asarray(v).shape), pressure_wall=out["structure_pressure"],
                density_wall=out["structure_density"])
