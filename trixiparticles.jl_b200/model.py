"""Model-parameter types mirroring the reference's constructors (same names, same keywords).

These are plain host-side records; all arithmetic on particles happens in the CUDA library.
Reference: smoothing_kernels.jl:191-227,400,434-455; state_equations.jl:100-139;
viscosity.jl:68-87; density_diffusion.jl:41-47; density_calculators.jl:1-24;
wcsph/system.jl:65-225; wall_boundary/system.jl:22-60; dummy_particles.jl:52-149.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from .setups import InitialCondition

# ids shared with include/tpb200.h
KERNEL_WENDLAND_C2 = 0
KERNEL_SCHOENBERG_CUBIC = 1
KERNEL_WENDLAND_C4 = 2
KERNEL_WENDLAND_C6 = 3
KERNEL_SCHOENBERG_QUARTIC = 4
KERNEL_SCHOENBERG_QUINTIC = 5
DENSITY_CONTINUITY = 0
DENSITY_SUMMATION = 1


@dataclass(frozen=True)
class WendlandC2Kernel:
    ndims: int
    kernel_id: int = KERNEL_WENDLAND_C2


@dataclass(frozen=True)
class SchoenbergCubicSplineKernel:
    ndims: int
    kernel_id: int = KERNEL_SCHOENBERG_CUBIC


@dataclass(frozen=True)
class WendlandC4Kernel:
    """smoothing_kernels.jl:489-514."""
    ndims: int
    kernel_id: int = KERNEL_WENDLAND_C4


@dataclass(frozen=True)
class WendlandC6Kernel:
    """smoothing_kernels.jl:548-574."""
    ndims: int
    kernel_id: int = KERNEL_WENDLAND_C6


@dataclass(frozen=True)
class SchoenbergQuarticSplineKernel:
    """smoothing_kernels.jl:264-322 (compact support 5/2 h)."""
    ndims: int
    kernel_id: int = KERNEL_SCHOENBERG_QUARTIC


@dataclass(frozen=True)
class SchoenbergQuinticSplineKernel:
    """smoothing_kernels.jl:357-395 (compact support 3 h)."""
    ndims: int
    kernel_id: int = KERNEL_SCHOENBERG_QUINTIC


def compact_support(kernel, h):
    """smoothing_kernels.jl:215, :400 (2h: cubic spline, Wendland), :310 (5/2 h), :383 (3h);
    evaluated in the type of `h`."""
    t = type(h) if isinstance(h, np.floating) else float
    if kernel.kernel_id == KERNEL_SCHOENBERG_QUARTIC:
        return t(5 / 2) * h
    if kernel.kernel_id == KERNEL_SCHOENBERG_QUINTIC:
        return 3 * h
    return 2 * h


class ContinuityDensity:
    density_id = DENSITY_CONTINUITY


class SummationDensity:
    density_id = DENSITY_SUMMATION


@dataclass(frozen=True)
class StateEquationCole:
    """state_equations.jl:100-139.  Fields are stored as ELTYPE = typeof(sound_speed)."""
    sound_speed: float
    reference_density: float
    exponent: float
    background_pressure: float = 0.0
    clip_negative_pressure: bool = False

    def __call__(self, density, dtype=np.float64):
        t = np.dtype(dtype).type
        c, g, r0, pb = t(self.sound_speed), t(self.exponent), t(self.reference_density), t(self.background_pressure)
        B = r0 * (c * c) / g
        x = t(density) / r0
        p = B * (t(np.float64(x) ** np.float64(g)) - t(1)) + pb
        return max(t(0), p) if self.clip_negative_pressure else p

    def inverse(self, pressure, dtype=np.float64):
        t = np.dtype(dtype).type
        c, g, r0, pb = t(self.sound_speed), t(self.exponent), t(self.reference_density), t(self.background_pressure)
        B = r0 * (c * c) / g
        tmp = (t(pressure) - pb) / B + t(1)
        return r0 * t(np.float64(tmp) ** np.float64(t(1) / g))


class StateEquationAdaptiveCole:
    """state_equations.jl:37-84: Cole's equation whose speed of sound follows the largest particle
    velocity, c = clamp(max|v| / mach_number_target, min_sound_speed, max_sound_speed), updated at
    the start of every right-hand-side evaluation (`update_speed_of_sound!`,
    wcsph/system.jl:307-321).  All fields are stored in ELTYPE = typeof(mach_number_target) --
    Float32 with the reference's default literals, whatever eltype(system) is; `sound_speed` starts
    at `min_sound_speed`.  One object may be shared by the fluid and the boundary model."""

    def __init__(self, *, reference_density, exponent, mach_number_target=np.float32(0.1),
                 min_sound_speed=np.float32(10.0), max_sound_speed=np.float32(100.0),
                 background_pressure=np.float32(0.0), clip_negative_pressure=False):
        t = type(mach_number_target) if isinstance(mach_number_target, np.floating) else np.float64
        self.param_eltype = np.dtype(t)
        self.mach_number_target = t(mach_number_target)
        self.min_sound_speed = t(min_sound_speed)
        self.max_sound_speed = t(max_sound_speed)
        self.exponent = t(exponent)
        self.reference_density = t(reference_density)
        self.background_pressure = t(background_pressure)
        self.clip_negative_pressure = bool(clip_negative_pressure)
        self.sound_speed = t(min_sound_speed)

    def update_speed_of_sound(self, velocity, eltype):
        """`velocity`: (n, ND) array of eltype(system).  Sets and returns `sound_speed`."""
        v = np.asarray(velocity, dtype=eltype)
        if len(v) == 0:
            return self.sound_speed
        s = v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]
        if v.shape[1] == 3:
            s = s + v[:, 2] * v[:, 2]
        v_max = np.sqrt(s.max())                       # eltype(system)
        q = v_max / self.mach_number_target            # promoted
        self.sound_speed = self.param_eltype.type(min(self.max_sound_speed, max(self.min_sound_speed, q)))
        return self.sound_speed

    def _B(self):
        return self.reference_density * self.sound_speed ** 2 / self.exponent   # in ELTYPE

    def __call__(self, density, dtype=np.float64):
        t = np.dtype(dtype).type
        x = t(density) / t(self.reference_density)
        p = t(self._B()) * (t(np.float64(x) ** np.float64(self.exponent)) - t(1)) + t(self.background_pressure)
        return max(t(0), p) if self.clip_negative_pressure else p

    def inverse(self, pressure, dtype=np.float64):
        t = np.dtype(dtype).type
        tmp = (t(pressure) - t(self.background_pressure)) / t(self._B()) + t(1)
        return t(self.reference_density) * t(np.float64(tmp) ** np.float64(t(1) / t(self.exponent)))


@dataclass(frozen=True)
class ArtificialViscosityMonaghan:
    alpha: float
    beta: float = 0.0
    epsilon: float = 0.01


@dataclass(frozen=True)
class ViscosityMorris:
    """viscosity.jl:134-205 (kinematic viscosity `nu`)."""
    nu: float
    epsilon: float = 0.01
    viscosity_id: int = 2


@dataclass(frozen=True)
class ViscosityAdami:
    """viscosity.jl:207-285 (kinematic viscosity `nu`)."""
    nu: float
    epsilon: float = 0.01
    viscosity_id: int = 3


@dataclass(frozen=True)
class DensityDiffusionMolteniColagrossi:
    delta: float


@dataclass(frozen=True)
class SourceTermDamping:
    damping_coefficient: float


@dataclass(frozen=True)
class AdamiPressureExtrapolation:
    pressure_offset: float = 0.0
    allow_loop_flipping: bool = True  # accepted for API parity; the GPU sweep is always wall-major


@dataclass(frozen=True)
class BernoulliPressureExtrapolation:
    """dummy_particles.jl:150-187.  Its dynamic-pressure term `factor * rho_f * (v_rel . n)^2 / 2` applies to a
    wall only while it moves (`if system.ismoving[]`, dummy_particles.jl:680-694) -- for a static wall it is exactly
    `AdamiPressureExtrapolation` with the same `pressure_offset` -- and always to a structure (:696-707)."""
    pressure_offset: float = 0.0
    factor: float = 1.0
    allow_loop_flipping: bool = True


class WeaklyCompressibleSPHSystem:
    """wcsph/system.jl:65-225 -- only the options on the accelerated path are accepted;
    anything else raises `ValueError` (the reference throws `ArgumentError`)."""

    def __init__(self, initial_condition: InitialCondition, *, smoothing_kernel,
                 smoothing_length, density_calculator, state_equation,
                 viscosity: Optional[ArtificialViscosityMonaghan] = None,
                 density_diffusion: Optional[DensityDiffusionMolteniColagrossi] = None,
                 acceleration: Optional[Sequence[float]] = None, source_terms=None,
                 correction=None, surface_tension=None, pressure_acceleration=None,
                 shifting_technique=None, buffer_size=None, reference_particle_spacing=0):
        nd = initial_condition.ndims
        if smoothing_kernel.ndims != nd:
            raise ValueError("smoothing kernel dimensionality doesn't match problem dimensionality")
        if acceleration is None:
            acceleration = (0.0,) * nd
        if len(acceleration) != nd:
            raise ValueError(f"`acceleration` must be of length {nd} for a {nd}D problem")
        for name, val in (("correction", correction), ("surface_tension", surface_tension),
                          ("pressure_acceleration", pressure_acceleration),
                          ("shifting_technique", shifting_technique), ("buffer_size", buffer_size)):
            if val is not None:
                raise ValueError(f"`{name}` is outside the accelerated hot path (see DESIGN.md)")
        if source_terms is not None and not isinstance(source_terms, SourceTermDamping):
            raise ValueError("only `SourceTermDamping` source terms are supported")
        if density_diffusion is not None and isinstance(density_calculator, SummationDensity):
            raise ValueError("`density_diffusion` is not used with `SummationDensity`")
        self.initial_condition = initial_condition
        self.smoothing_kernel = smoothing_kernel
        self.eltype = initial_condition.eltype
        self.coordinates_eltype = initial_condition.coordinates_eltype
        self.smoothing_length = self.eltype.type(smoothing_length)
        self.density_calculator = density_calculator
        self.state_equation = state_equation
        self.viscosity = viscosity
        self.density_diffusion = density_diffusion
        self.acceleration = np.asarray(acceleration, dtype=self.eltype)
        self.source_terms = source_terms
        self.mass = initial_condition.mass.copy()
        # filled by the semidiscretization after every kick (system.pressure / cache.density)
        self.pressure = np.zeros(initial_condition.nparticles, dtype=self.eltype)

    ndims = property(lambda self: self.initial_condition.ndims)
    nparticles = property(lambda self: self.initial_condition.nparticles)
    n_integrated_particles = property(lambda self: self.initial_condition.nparticles)
    u_nvariables = property(lambda self: self.ndims)

    @property
    def v_nvariables(self):  # wcsph/system.jl:232-242
        return self.ndims + (0 if isinstance(self.density_calculator, SummationDensity) else 1)


class BoundaryModelDummyParticles:
    """dummy_particles.jl:52-113 with `AdamiPressureExtrapolation` (`BernoulliPressureExtrapolation` is the
    same thing for a static wall) or `ContinuityDensity` (the wall density is integrated: one row of the
    ODE vector per wall particle, wall_boundary/rhs.jl:11-59, system.jl:78-90)."""

    def __init__(self, initial_density, hydrodynamic_mass, density_calculator, smoothing_kernel,
                 smoothing_length, *, viscosity=None, state_equation=None, correction=None,
                 clip_negative_pressure=False, reference_particle_spacing=0.0):
        if not isinstance(density_calculator, (AdamiPressureExtrapolation, BernoulliPressureExtrapolation,
                                               ContinuityDensity)):
            raise ValueError("only `AdamiPressureExtrapolation` / `BernoulliPressureExtrapolation` / "
                             "`ContinuityDensity` are on the accelerated path")
        if isinstance(density_calculator, ContinuityDensity) and viscosity is not None:
            raise ValueError("a no-slip wall with `ContinuityDensity` is outside the accelerated path")
        if correction is not None:
            raise ValueError("wall `correction` is outside the accelerated hot path")
        if viscosity is not None and not isinstance(viscosity, (ArtificialViscosityMonaghan, ViscosityMorris,
                                                                 ViscosityAdami)):
            raise ValueError("wall `viscosity` must be ArtificialViscosityMonaghan, ViscosityMorris or "
                             "ViscosityAdami on the accelerated path")
        if state_equation is None:
            raise ValueError("the boundary model needs a `state_equation`")
        self.initial_density = np.asarray(initial_density)
        self.hydrodynamic_mass = np.asarray(hydrodynamic_mass)
        assert self.initial_density.shape == self.hydrodynamic_mass.shape
        self.density_calculator = density_calculator
        self.smoothing_kernel = smoothing_kernel
        self.eltype = self.hydrodynamic_mass.dtype
        self.smoothing_length = self.eltype.type(smoothing_length)
        self.state_equation = state_equation
        self.viscosity = viscosity   # None: free-slip; a model: no-slip wall (dummy_particles.jl:299-307)
        self.clip_negative_pressure = bool(clip_negative_pressure)
        n = self.hydrodynamic_mass.shape[0]
        self.pressure = np.zeros(n, dtype=self.eltype)
        self.cache = dict(density=self.initial_density.copy(), volume=np.zeros(n, dtype=self.eltype))


class PrescribedMotion:
    """schemes/boundary/prescribed_motion.jl:1-121: `PrescribedMotion(movement_function, is_moving;
    moving_particles=nothing)`.

    `movement_function(x, t)`: here `x` is the (n, ND) float64 array of the INITIAL positions of the moving
    particles (the reference calls it per particle with an SVector; write it with `x[:, 0]`, `x[:, 1]`, ...) and
    the return value the (n, ND) array of their positions at time `t`.  `is_moving(t)` -> bool.
    `moving_particles`: 0-based indices into the system (default: every particle of a `WallBoundarySystem`, every
    clamped particle of a `TotalLagrangianSPHSystem`).

    The reference differentiates the movement function twice with ForwardDiff (:107-111).  Without automatic
    differentiation the caller can hand over the exact `velocity_function(x, t)` / `acceleration_function(x, t)`;
    otherwise fourth-order central differences in float64 with step `fd_step` are used (exact up to rounding for
    polynomials in t up to degree 4)."""

    def __init__(self, movement_function, is_moving, *, moving_particles=None, velocity_function=None,
                 acceleration_function=None, fd_step=1e-3):
        self.movement_function = movement_function
        self.is_moving = is_moving
        self.moving_particles = None if moving_particles is None else np.asarray(moving_particles, dtype=np.int64).ravel()
        self.velocity_function = velocity_function
        self.acceleration_function = acceleration_function
        self.fd_step = float(fd_step)

    def initialize(self, n_particles, n_clamped=None):
        """initialize_prescribed_motion! (:65-88): an empty `moving_particles` means all clamped particles (the
        last `n_clamped` of the system; a wall: all)."""
        n_clamped = n_particles if n_clamped is None else n_clamped
        if self.moving_particles is None or self.moving_particles.size == 0:
            self.moving_particles = np.arange(n_particles - n_clamped, n_particles, dtype=np.int64)
        if len(self.moving_particles) and (self.moving_particles.min() < n_particles - n_clamped
                                            or self.moving_particles.max() >= n_particles):
            raise ValueError("`moving_particles` must be clamped particles of the system")
        return self

    def __call__(self, x0, t):
        """(:95-121) positions, velocities and accelerations (float64, (n, ND)) of the moving particles whose
        initial positions are `x0`, at time `t`."""
        x0 = np.asarray(x0, dtype=np.float64)
        t = float(t)
        f = lambda tt: np.asarray(self.movement_function(x0, tt), dtype=np.float64).reshape(x0.shape)
        pos = f(t)
        h = self.fd_step
        if self.velocity_function is None or self.acceleration_function is None:
            fm2, fm1, fp1, fp2 = f(t - 2 * h), f(t - h), f(t + h), f(t + 2 * h)
        if self.velocity_function is not None:
            vel = np.asarray(self.velocity_function(x0, t), dtype=np.float64).reshape(x0.shape)
        else:
            vel = (fm2 - 8.0 * fm1 + 8.0 * fp1 - fp2) / (12.0 * h)
        if self.acceleration_function is not None:
            acc = np.asarray(self.acceleration_function(x0, t), dtype=np.float64).reshape(x0.shape)
        else:
            acc = (-fm2 + 16.0 * fm1 - 30.0 * pos + 16.0 * fp1 - fp2) / (12.0 * h * h)
        return pos, vel, acc


def OscillatingMotion2D(*, frequency, translation_vector, rotation_angle, rotation_center,
                        rotation_phase_offset=0, tspan=(-np.inf, np.inf), ramp_up_tspan=(0.0, 0.0),
                        moving_particles=None):
    """prescribed_motion.jl:123-203: translation + rotation about a centre with the same frequency, optional
    smoothstep ramp-up."""
    tv = np.asarray(translation_vector, dtype=np.float64).reshape(2)
    rc = np.asarray(rotation_center, dtype=np.float64).reshape(2)

    def movement_function(x, t):
        if np.isfinite(tspan[0]):
            t = t - tspan[0]
        sin_scaled = np.sin(frequency * 2 * np.pi * t)
        xc = x - rc[None, :]
        angle = rotation_angle * np.sin(2 * np.pi * (frequency * t - rotation_phase_offset))
        rotated = np.stack([xc[:, 0] * np.cos(angle) - xc[:, 1] * np.sin(angle),
                            xc[:, 0] * np.sin(angle) + xc[:, 1] * np.cos(angle)], axis=1)
        result = rotated + rc[None, :] + sin_scaled * tv[None, :]
        if ramp_up_tspan[1] > ramp_up_tspan[0] and ramp_up_tspan[0] <= t <= ramp_up_tspan[1]:
            t_rel = (t - ramp_up_tspan[0]) / (ramp_up_tspan[1] - ramp_up_tspan[0])
            ramp = 3 * t_rel ** 2 - 2 * t_rel ** 3
            return result * ramp + (1 - ramp) * x
        return result

    return PrescribedMotion(movement_function, lambda t: tspan[0] <= t <= tspan[1], moving_particles=moving_particles)


class _MovingParticles:
    """The host side of apply_prescribed_motion! (wall_boundary/system.jl:199-205, total_lagrangian_sph/
    system.jl:436-447) for the clamped tail of a system: where those particles are, how they move."""

    def _init_motion(self, motion, initial_coordinates, n_clamped):
        n = initial_coordinates.shape[0]
        self.prescribed_motion = motion
        self.ismoving = motion is not None
        if motion is None:
            return
        motion.initialize(n, n_clamped)
        self._x0_clamped = np.array(initial_coordinates[n - n_clamped:], dtype=np.float64)
        self._moving_local = motion.moving_particles - (n - n_clamped)
        self.clamped_coordinates = self._x0_clamped.copy()
        self.clamped_velocity = np.zeros_like(self._x0_clamped)
        self.clamped_acceleration = np.zeros_like(self._x0_clamped)

    def apply_prescribed_motion(self, t):
        """-> is_moving(t); updates clamped_coordinates / _velocity / _acceleration (float64)."""
        m = self.prescribed_motion
        self.ismoving = bool(m.is_moving(t))
        if not self.ismoving:
            return False
        idx = self._moving_local
        pos, vel, acc = m(self._x0_clamped[idx], t)
        self.clamped_coordinates[idx] = pos
        self.clamped_velocity[idx] = vel
        self.clamped_acceleration[idx] = acc
        return True


class WallBoundarySystem(_MovingParticles):
    """wall_boundary/system.jl:22-60.  With `prescribed_motion` the wall is registered with the library as a
    system of clamped, moving dummy particles (free-slip Adami / Bernoulli extrapolation)."""

    def __init__(self, initial_condition: InitialCondition, boundary_model, *,
                 prescribed_motion=None, adhesion_coefficient=0.0):
        if adhesion_coefficient != 0.0:
            raise ValueError("adhesion is outside the accelerated hot path")
        if prescribed_motion is not None:
            if not isinstance(prescribed_motion, PrescribedMotion):
                raise TypeError("`prescribed_motion` must be a PrescribedMotion")
            if (not isinstance(boundary_model, BoundaryModelDummyParticles)
                    or isinstance(boundary_model.density_calculator, ContinuityDensity)
                    or boundary_model.viscosity is not None):
                raise ValueError("a moving wall on the accelerated path: BoundaryModelDummyParticles with Adami / "
                                 "Bernoulli pressure extrapolation and no viscosity (free-slip)")
        self.initial_condition = initial_condition
        self.coordinates = initial_condition.coordinates
        self.boundary_model = boundary_model
        self.eltype = boundary_model.eltype
        self.coordinates_eltype = initial_condition.coordinates_eltype
        self._init_motion(prescribed_motion, initial_condition.coordinates, initial_condition.nparticles)

    ndims = property(lambda self: self.initial_condition.ndims)
    nparticles = property(lambda self: self.initial_condition.nparticles)
    # wall_boundary/system.jl:78-90: nothing is integrated, except the density of dummy particles with
    # `ContinuityDensity` (one v variable per particle, no u variable)
    integrates_density = property(lambda self: isinstance(getattr(self.boundary_model, "density_calculator", None),
                                                          ContinuityDensity))
    n_integrated_particles = property(lambda self: self.nparticles if self.integrates_density else 0)
    u_nvariables = property(lambda self: 0)
    v_nvariables = property(lambda self: 1)


@dataclass(frozen=True)
class PenaltyForceGanzenmueller:
    """structure/total_lagrangian_sph/penalty_force.jl:1-16."""
    alpha: float = 0.1


class BoundaryModelMonaghanKajtar:
    """wall_boundary/monaghan_kajtar.jl:1-34: repulsive boundary particles
    `BoundaryModelMonaghanKajtar(K, beta, boundary_particle_spacing, mass; viscosity=nothing)`."""

    def __init__(self, K, beta, boundary_particle_spacing, mass, *, viscosity=None):
        if viscosity is not None:
            raise ValueError("a viscous Monaghan-Kajtar boundary is outside the accelerated hot path")
        self.hydrodynamic_mass = np.asarray(mass)
        self.eltype = self.hydrodynamic_mass.dtype
        self.K = self.eltype.type(K)
        self.beta = self.eltype.type(beta)
        self.boundary_particle_spacing = self.eltype.type(boundary_particle_spacing)
        self.viscosity = None


class TotalLagrangianSPHSystem(_MovingParticles):
    """structure/total_lagrangian_sph/system.jl:76-184 on the accelerated path (BASELINE config 5):
    scalar `young_modulus` / `poisson_ratio`, clamped particles fixed or moved by
    `clamped_particles_motion=PrescribedMotion(...)` (whose `moving_particles` index the sorted system), optional
    `PenaltyForceGanzenmueller`, `boundary_model` = `BoundaryModelMonaghanKajtar`, `BoundaryModelDummyParticles`
    or None.
    As in the reference, the clamped particles are moved to the end of the particle list
    (`move_particles_to_end!`); `clamped_particles` are 0-based indices here."""

    def __init__(self, initial_condition: InitialCondition, *, smoothing_kernel, smoothing_length,
                 young_modulus, poisson_ratio, clamped_particles=(), clamped_particles_motion=None,
                 acceleration: Optional[Sequence[float]] = None, penalty_force=None, viscosity=None,
                 source_terms=None, boundary_model=None, self_interaction_nhs="default",
                 velocity_averaging=None):
        nd = initial_condition.ndims
        if smoothing_kernel.ndims != nd:
            raise ValueError(f"smoothing kernel dimensionality must be {nd} for a {nd}D problem")
        if acceleration is None:
            acceleration = (0.0,) * nd
        if len(acceleration) != nd:
            raise ValueError(f"`acceleration` must be of length {nd} for a {nd}D problem")
        if clamped_particles_motion is not None and not isinstance(clamped_particles_motion, PrescribedMotion):
            raise TypeError("`clamped_particles_motion` must be a PrescribedMotion")
        for name, val in (("viscosity", viscosity),
                          ("source_terms", source_terms), ("velocity_averaging", velocity_averaging)):
            if val is not None:
                raise ValueError(f"`{name}` is outside the accelerated hot path (see DESIGN.md)")
        # scalars or one value per particle (system.jl:108-161); vectors follow the particles to their sorted places
        per_particle = not (np.isscalar(young_modulus) and np.isscalar(poisson_ratio))
        if per_particle:
            n_ = initial_condition.nparticles
            young_modulus = np.broadcast_to(np.asarray(young_modulus, dtype=np.float64), (n_,)).copy()
            poisson_ratio = np.broadcast_to(np.asarray(poisson_ratio, dtype=np.float64), (n_,)).copy()
        if penalty_force is not None and not isinstance(penalty_force, PenaltyForceGanzenmueller):
            raise ValueError("`penalty_force` must be a PenaltyForceGanzenmueller")
        dummy = isinstance(boundary_model, BoundaryModelDummyParticles)
        if boundary_model is not None and not dummy and not isinstance(boundary_model, BoundaryModelMonaghanKajtar):
            raise ValueError("structure `boundary_model`: BoundaryModelMonaghanKajtar or BoundaryModelDummyParticles")
        if dummy and (isinstance(boundary_model.density_calculator, ContinuityDensity) or boundary_model.viscosity is not None):
            raise ValueError("structure `BoundaryModelDummyParticles`: Adami / Bernoulli pressure extrapolation without "
                             "viscosity are on the accelerated path")
        clamped = np.asarray(list(clamped_particles), dtype=np.int64)
        n = initial_condition.nparticles
        if len(np.unique(clamped)) != len(clamped):
            raise ValueError("`clamped_particles` contains duplicate particle indices")
        order = np.concatenate([np.setdiff1d(np.arange(n), clamped, assume_unique=False), clamped])
        ic = initial_condition
        self.initial_condition = InitialCondition(
            coordinates=np.ascontiguousarray(ic.coordinates[order]), velocity=np.ascontiguousarray(ic.velocity[order]),
            mass=np.ascontiguousarray(ic.mass[order]), density=np.ascontiguousarray(ic.density[order]),
            pressure=np.ascontiguousarray(ic.pressure[order]), particle_spacing=ic.particle_spacing)
        if dummy and len(clamped):
            # the constructor moves the clamped particles to the end (system.jl:131-147): so do the model's arrays
            boundary_model = BoundaryModelDummyParticles(
                boundary_model.initial_density[order], boundary_model.hydrodynamic_mass[order],
                boundary_model.density_calculator, boundary_model.smoothing_kernel, boundary_model.smoothing_length,
                state_equation=boundary_model.state_equation,
                clip_negative_pressure=boundary_model.clip_negative_pressure)
        elif boundary_model is not None and len(clamped):
            boundary_model = BoundaryModelMonaghanKajtar(boundary_model.K, boundary_model.beta,
                                                         boundary_model.boundary_particle_spacing,
                                                         boundary_model.hydrodynamic_mass[order])
        self.particle_order = order
        self.n_clamped_particles = len(clamped)
        self.smoothing_kernel = smoothing_kernel
        self.eltype = ic.eltype
        self.coordinates_eltype = ic.coordinates_eltype
        self.smoothing_length = self.eltype.type(smoothing_length)
        if per_particle:
            self.young_modulus = young_modulus[order].astype(self.eltype)
            self.poisson_ratio = poisson_ratio[order].astype(self.eltype)
        else:
            self.young_modulus = self.eltype.type(young_modulus)
            self.poisson_ratio = self.eltype.type(poisson_ratio)
        self.acceleration = np.asarray(acceleration, dtype=self.eltype)
        self.penalty_force = penalty_force
        self.boundary_model = boundary_model
        self.initial_coordinates = self.initial_condition.coordinates
        self.mass = self.initial_condition.mass
        self.material_density = self.initial_condition.density
        self._init_motion(clamped_particles_motion, self.initial_coordinates, self.n_clamped_particles)

    ndims = property(lambda self: self.initial_condition.ndims)
    nparticles = property(lambda self: self.initial_condition.nparticles)
    n_integrated_particles = property(lambda self: self.initial_condition.nparticles - self.n_clamped_particles)
    u_nvariables = property(lambda self: self.ndims)
    v_nvariables = property(lambda self: self.ndims)   # system.jl:277-279
