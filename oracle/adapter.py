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
    p.pressure_offset = float(t(m.density_calculator.pressure_offset))
    return p


def kick(fluid, wall, u, v, use_grid=True, nthreads=0, fluid_wall_interaction=True):
    """Oracle `kick!` for Semidiscretization(fluid[, wall]); u (n, ND), v (n, nv)."""
    fp = fluid_params(fluid)
    if wall is not None and fluid_wall_interaction:
        wp = wall_params(wall)
        cw, mw = wall.coordinates, wall.boundary_model.hydrodynamic_mass
    else:
        wp, cw, mw = None, None, None
    return O.kick(fp, wp, fluid.mass, cw, mw, v, u, fluid.eltype, use_grid=use_grid,
                  nthreads=nthreads)
