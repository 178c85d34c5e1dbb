"""The reference's example setups on the accelerated path, as functions.

Each mirrors one script under /root/reference/examples/fluid/ (same parameter values and
construction order) and returns `(fluid_system, boundary_system, tank)`.  `eltype` /
`coordinates_eltype` play the role of `trixi_include_changeprecision` (docs/src/gpu.md:107-133).
"""
from __future__ import annotations

import numpy as np

from .model import (AdamiPressureExtrapolation, ArtificialViscosityMonaghan,
                    BoundaryModelDummyParticles, ContinuityDensity,
                    DensityDiffusionMolteniColagrossi, SchoenbergCubicSplineKernel,
                    StateEquationAdaptiveCole, StateEquationCole, SummationDensity, WallBoundarySystem,
                    WeaklyCompressibleSPHSystem, WendlandC2Kernel, BoundaryModelMonaghanKajtar,
                    PenaltyForceGanzenmueller, TotalLagrangianSPHSystem)
from .setups import RectangularShape, RectangularTank, union


def dam_break_2d(particles_per_height=40, *, eltype=np.float64, coordinates_eltype=np.float64,
                 density_calculator=None, alpha=0.02, delta=0.1, sound_speed_factor=20.0,
                 boundary_density_calculator=None, boundary_model=None, boundary_layers=4, spacing_ratio=1):
    """examples/fluid/dam_break_2d.jl:19-97 (BASELINE config 1 at particles_per_height=40).
    `boundary_density_calculator=ContinuityDensity()` and `boundary_model="monaghan_kajtar"` (with
    `boundary_layers=1, spacing_ratio=3`) are the variants of the reference's GPU tests
    (test/examples/gpu.jl:198-253)."""
    t = np.dtype(eltype).type
    H = 0.6
    W = 2 * H
    dx = H / particles_per_height
    gravity = 9.81
    tank_size = (np.floor(5.366 * H / dx) * dx, 4.0)
    sound_speed = sound_speed_factor * np.sqrt(gravity * H)
    state_equation = StateEquationCole(sound_speed=float(t(sound_speed)), reference_density=1000.0,
                                       exponent=1, clip_negative_pressure=False)
    tank = RectangularTank(dx, (W, H), tank_size, 1000.0, n_layers=boundary_layers, spacing_ratio=spacing_ratio,
                           acceleration=(0.0, -gravity), state_equation=state_equation,
                           coordinates_eltype=coordinates_eltype, eltype=eltype)
    h = 2 * dx
    kernel = WendlandC2Kernel(2)
    dc = density_calculator or ContinuityDensity()
    fluid = WeaklyCompressibleSPHSystem(
        tank.fluid, smoothing_kernel=kernel, smoothing_length=h, density_calculator=dc,
        state_equation=state_equation, viscosity=ArtificialViscosityMonaghan(alpha=alpha, beta=0.0),
        density_diffusion=(None if isinstance(dc, SummationDensity)
                           else DensityDiffusionMolteniColagrossi(delta=delta)),
        acceleration=(0.0, -gravity))
    if boundary_model == "monaghan_kajtar":
        from .model import BoundaryModelMonaghanKajtar
        model = BoundaryModelMonaghanKajtar(0.5, spacing_ratio, dx / spacing_ratio, tank.boundary.mass)
    else:
        model = BoundaryModelDummyParticles(tank.boundary.density, tank.boundary.mass,
                                            boundary_density_calculator or AdamiPressureExtrapolation(), kernel, h,
                                            state_equation=state_equation, clip_negative_pressure=True)
    wall = WallBoundarySystem(tank.boundary, model)
    return fluid, wall, tank


def hydrostatic_water_column_2d(fluid_particle_spacing=0.05, *, eltype=np.float32,
                                coordinates_eltype=np.float32, density_calculator=None,
                                prescribed_motion=None, system_acceleration=None):
    """examples/fluid/hydrostatic_water_column_2d.jl:13-69 (BASELINE config 2); `prescribed_motion` /
    `system_acceleration` as overridden by examples/fluid/accelerated_tank_2d.jl."""
    gravity = 9.81
    dx = fluid_particle_spacing
    state_equation = StateEquationCole(sound_speed=10.0, reference_density=1000.0, exponent=7,
                                       clip_negative_pressure=False)
    tank = RectangularTank(dx, (1.0, 0.9), (1.0, 1.0), 1000.0, n_layers=3,
                           acceleration=(0.0, -gravity), state_equation=state_equation,
                           coordinates_eltype=coordinates_eltype, eltype=eltype)
    h = 1.2 * dx
    kernel = SchoenbergCubicSplineKernel(2)
    fluid = WeaklyCompressibleSPHSystem(
        tank.fluid, smoothing_kernel=kernel, smoothing_length=h,
        density_calculator=density_calculator or ContinuityDensity(),
        state_equation=state_equation, viscosity=ArtificialViscosityMonaghan(alpha=0.02, beta=0.0),
        acceleration=(0.0, -gravity) if system_acceleration is None else tuple(system_acceleration))
    model = BoundaryModelDummyParticles(tank.boundary.density, tank.boundary.mass,
                                        AdamiPressureExtrapolation(), kernel, h,
                                        state_equation=state_equation)
    wall = WallBoundarySystem(tank.boundary, model, prescribed_motion=prescribed_motion)
    return fluid, wall, tank


def hydrostatic_water_column_3d(fluid_particle_spacing=0.05, *, eltype=np.float32, coordinates_eltype=np.float32):
    """examples/fluid/hydrostatic_water_column_3d.jl: the 2-D example with `initial_fluid_size = (1, 1, 0.9)`,
    `tank_size = (1, 1, 1.2)`, gravity along z and the 3-D cubic spline."""
    gravity = 9.81
    dx = fluid_particle_spacing
    acc = (0.0, 0.0, -gravity)
    state_equation = StateEquationCole(sound_speed=10.0, reference_density=1000.0, exponent=7,
                                       clip_negative_pressure=False)
    tank = RectangularTank(dx, (1.0, 1.0, 0.9), (1.0, 1.0, 1.2), 1000.0, n_layers=3, acceleration=acc,
                           state_equation=state_equation, coordinates_eltype=coordinates_eltype, eltype=eltype)
    h = 1.2 * dx
    kernel = SchoenbergCubicSplineKernel(3)
    fluid = WeaklyCompressibleSPHSystem(
        tank.fluid, smoothing_kernel=kernel, smoothing_length=h, density_calculator=ContinuityDensity(),
        state_equation=state_equation, viscosity=ArtificialViscosityMonaghan(alpha=0.02, beta=0.0), acceleration=acc)
    model = BoundaryModelDummyParticles(tank.boundary.density, tank.boundary.mass, AdamiPressureExtrapolation(), kernel, h,
                                        state_equation=state_equation)
    return fluid, WallBoundarySystem(tank.boundary, model), tank


def falling_water_column_2d(fluid_particle_spacing=0.02, *, eltype=np.float64, coordinates_eltype=np.float64):
    """examples/fluid/falling_water_column_2d.jl:13-75: a 0.5 x 1 block of water released 0.2 above the floor in the
    middle of a 4 x 4 tank (no initial hydrostatic pressure: the block is in free fall)."""
    gravity = 9.81
    dx = fluid_particle_spacing
    initial_fluid_size, tank_size = (0.5, 1.0), (4.0, 4.0)
    sound_speed = 10 * np.sqrt(gravity * initial_fluid_size[1])
    state_equation = StateEquationCole(sound_speed=sound_speed, reference_density=1000.0, exponent=7)
    tank = RectangularTank(dx, initial_fluid_size, tank_size, 1000.0, n_layers=3, spacing_ratio=1,
                           coordinates_eltype=coordinates_eltype, eltype=eltype)
    shift = np.array([0.5 * tank_size[0] - 0.5 * initial_fluid_size[0], 0.2])
    tank.fluid.coordinates = (tank.fluid.coordinates + shift[None, :]).astype(coordinates_eltype)
    h = 1.2 * dx
    kernel = SchoenbergCubicSplineKernel(2)
    fluid = WeaklyCompressibleSPHSystem(
        tank.fluid, smoothing_kernel=kernel, smoothing_length=h, density_calculator=ContinuityDensity(),
        state_equation=state_equation, viscosity=ArtificialViscosityMonaghan(alpha=0.02, beta=0.0),
        acceleration=(0.0, -gravity))
    model = BoundaryModelDummyParticles(tank.boundary.density, tank.boundary.mass, AdamiPressureExtrapolation(), kernel, h,
                                        state_equation=state_equation, clip_negative_pressure=True)
    return fluid, WallBoundarySystem(tank.boundary, model), tank


def accelerated_tank_2d(fluid_particle_spacing=0.05, *, eltype=np.float64, coordinates_eltype=np.float64):
    """examples/fluid/accelerated_tank_2d.jl: the hydrostatic water column without gravity in a tank that is
    accelerated upwards with g -- in the tank's frame the same problem."""
    from .model import PrescribedMotion
    g = 9.81
    motion = PrescribedMotion(lambda x, t: x + np.array([0.0, 0.5 * g * t * t])[None, :], lambda t: True,
                              velocity_function=lambda x, t: np.broadcast_to(np.array([0.0, g * t]), x.shape),
                              acceleration_function=lambda x, t: np.broadcast_to(np.array([0.0, g]), x.shape))
    return hydrostatic_water_column_2d(fluid_particle_spacing, eltype=eltype, coordinates_eltype=coordinates_eltype,
                                       prescribed_motion=motion, system_acceleration=(0.0, 0.0))


def moving_wall_2d(fluid_particle_spacing=0.05, *, eltype=np.float64, coordinates_eltype=np.float64):
    """examples/fluid/moving_wall_2d.jl:13-77: a water column in a tank whose right wall, reset to the edge of
    the column, moves away with x + t^2 / 2 for t < 1.5.  Returns (fluid, wall, tank)."""
    from .model import PrescribedMotion
    from .setups import reset_wall_
    gravity = 9.81
    dx = fluid_particle_spacing
    initial_fluid_size, tank_size = (1.0, 0.8), (4.0, 1.0)
    sound_speed = 10 * np.sqrt(gravity * initial_fluid_size[1])
    state_equation = StateEquationCole(sound_speed=sound_speed, reference_density=1000.0, exponent=7)
    tank = RectangularTank(dx, initial_fluid_size, tank_size, 1000.0, n_layers=3, spacing_ratio=1.0,
                           acceleration=(0.0, -gravity), state_equation=state_equation,
                           coordinates_eltype=coordinates_eltype, eltype=eltype)
    reset_wall_(tank, (False, True, False, False), (0.0, tank.fluid_size[0], 0.0, 0.0))
    motion = PrescribedMotion(lambda x, t: x + np.array([0.5 * t * t, 0.0])[None, :], lambda t: t < 1.5,
                              moving_particles=tank.face_indices[1],
                              velocity_function=lambda x, t: np.broadcast_to(np.array([t, 0.0]), x.shape),
                              acceleration_function=lambda x, t: np.broadcast_to(np.array([1.0, 0.0]), x.shape))
    h = 1.2 * dx
    kernel = SchoenbergCubicSplineKernel(2)
    fluid = WeaklyCompressibleSPHSystem(
        tank.fluid, smoothing_kernel=kernel, smoothing_length=h, density_calculator=ContinuityDensity(),
        state_equation=state_equation, viscosity=ArtificialViscosityMonaghan(alpha=0.1, beta=0.0),
        acceleration=(0.0, -gravity))
    model = BoundaryModelDummyParticles(tank.boundary.density, tank.boundary.mass,
                                        AdamiPressureExtrapolation(), kernel, h, state_equation=state_equation)
    wall = WallBoundarySystem(tank.boundary, model, prescribed_motion=motion)
    return fluid, wall, tank


def dam_break_3d(fluid_particle_spacing=0.08, *, eltype=np.float32, coordinates_eltype=None,
                 sound_speed=None, fluid_size=(2.0, 1.0, 1.0), tank_size=None, adaptive_sound_speed=False,
                 x_window=None, boundary_x_window=None, density_calculator=None):
    """examples/fluid/dam_break_3d.jl:13-66 (BASELINE configs 3/4 at smaller spacings).

    As SURVEY.md section 8(d) M3 prescribes, the headline runs use a static
    `StateEquationCole(c = 20 sqrt(g * 1.0), exponent = 7)` instead of the script's
    `StateEquationAdaptiveCole` (which only adds one global max|v| reduction per RHS)."""
    gravity = 9.81
    dx = fluid_particle_spacing
    coordinates_eltype = coordinates_eltype or eltype
    if tank_size is None:
        tank_size = (np.floor(5.366 / dx) * dx, 4.0, 1.0)
    c = sound_speed if sound_speed is not None else 20 * np.sqrt(gravity * 1.0)
    state_equation = StateEquationCole(sound_speed=float(np.dtype(eltype).type(c)),
                                       reference_density=1000.0, exponent=7)
    if adaptive_sound_speed:
        # the script as shipped (dam_break_3d.jl:33-34): one object shared by fluid and boundary model
        state_equation = StateEquationAdaptiveCole(reference_density=1000.0, exponent=7)
    tank = RectangularTank(dx, fluid_size, tank_size, 1000.0, n_layers=4, spacing_ratio=1,
                           acceleration=(0.0, -gravity, 0.0), state_equation=state_equation,
                           coordinates_eltype=coordinates_eltype, eltype=eltype, x_window=x_window,
                           boundary_x_window=boundary_x_window)
    h = 1.5 * dx
    kernel = WendlandC2Kernel(3)
    dc = density_calculator or ContinuityDensity()
    fluid = WeaklyCompressibleSPHSystem(
        tank.fluid, smoothing_kernel=kernel, smoothing_length=h,
        density_calculator=dc, state_equation=state_equation,
        viscosity=ArtificialViscosityMonaghan(alpha=0.02, beta=0.0),
        density_diffusion=(None if isinstance(dc, SummationDensity)
                           else DensityDiffusionMolteniColagrossi(delta=0.1)),
        acceleration=(0.0, -gravity, 0.0))
    model = BoundaryModelDummyParticles(tank.boundary.density, tank.boundary.mass,
                                        AdamiPressureExtrapolation(), kernel, h,
                                        state_equation=state_equation, clip_negative_pressure=True)
    wall = WallBoundarySystem(tank.boundary, model)
    return fluid, wall, tank


def perturbed_state(fluid, seed=1234, position_jitter=0.1, velocity_scale=0.05,
                    density_jitter=0.01):
    """positions + U(-0.1dx, 0.1dx), velocities U(-0.05c, 0.05c),
    densities rho (1 +- 0.01), numpy default_rng(seed).  Removes exact-distance ties and
    exercises the `vr < 0` viscosity branch.  Returns (u, v) particle-major arrays."""
    rng = np.random.default_rng(seed)
    ic = fluid.initial_condition
    n, nd = ic.coordinates.shape
    dx = ic.particle_spacing
    c = float(fluid.state_equation.sound_speed)
    u = ic.coordinates.astype(np.float64) + rng.uniform(-position_jitter * dx, position_jitter * dx, (n, nd))
    vel = rng.uniform(-velocity_scale * c, velocity_scale * c, (n, nd))
    rho = ic.density.astype(np.float64) * (1 + rng.uniform(-density_jitter, density_jitter, n))
    u = u.astype(fluid.coordinates_eltype)
    if fluid.v_nvariables == nd + 1:
        v = np.concatenate([vel, rho[:, None]], axis=1).astype(fluid.eltype)
    else:
        v = vel.astype(fluid.eltype)
    return np.ascontiguousarray(u), np.ascontiguousarray(v)


def dam_break_plate_2d(fluid_particle_spacing=0.01, *, n_particles_x=5, eltype=np.float64,
                       coordinates_eltype=None, initial_fluid_size=(0.146, 2 * 0.146), plate_position=None,
                       E=1e6, nu=0.0, structure_boundary_model="monaghan_kajtar", clamped_particles_motion=None,
                       structure_pressure_extrapolation=None):
    """examples/fsi/dam_break_plate_2d.jl:19-134 (BASELINE config 5): dam break against an elastic
    plate clamped at its base; WCSPH fluid, dummy-particle tank, TLSPH plate with a
    BoundaryModelMonaghanKajtar towards the fluid and a PenaltyForceGanzenmueller.  The reference's
    GPU tests run it with `initial_fluid_size = (0.15, 0.29)` in Float32 (test/examples/gpu.jl:664-726).
    Returns (fluid_system, boundary_system, structure_system, tank)."""
    t = np.dtype(eltype).type
    coordinates_eltype = coordinates_eltype or np.float64
    gravity = 9.81
    dx = fluid_particle_spacing
    tank_size = tuple(4 * s for s in initial_fluid_size)
    fluid_density = 1000.0
    sound_speed = 20 * np.sqrt(gravity * initial_fluid_size[1])
    state_equation = StateEquationCole(sound_speed=float(t(sound_speed)), reference_density=fluid_density,
                                       exponent=1)
    tank = RectangularTank(dx, initial_fluid_size, tank_size, fluid_density, n_layers=4, spacing_ratio=1,
                           acceleration=(0.0, -gravity), state_equation=state_equation,
                           coordinates_eltype=coordinates_eltype, eltype=eltype)
    length_beam, thickness, structure_density = 0.08, 0.012, 2500
    ds = thickness / (n_particles_x - 1)
    n_particles_y = int(np.rint(length_beam / ds)) + 1
    if plate_position is None:
        plate_position = (2 * initial_fluid_size[0], 0.0)
    non_fixed_position = (plate_position[0], plate_position[1] + ds)
    plate = RectangularShape(ds, (n_particles_x, n_particles_y - 1), non_fixed_position, density=structure_density,
                             place_on_shell=True, coordinates_eltype=coordinates_eltype, eltype=eltype)
    clamped = RectangularShape(ds, (n_particles_x, 1), plate_position, density=structure_density,
                               place_on_shell=True, coordinates_eltype=coordinates_eltype, eltype=eltype)
    structure = union(clamped, plate)
    h = 1.75 * dx
    kernel = WendlandC2Kernel(2)
    fluid = WeaklyCompressibleSPHSystem(
        tank.fluid, smoothing_kernel=kernel, smoothing_length=h, density_calculator=ContinuityDensity(),
        state_equation=state_equation, viscosity=ArtificialViscosityMonaghan(alpha=0.02, beta=0.0),
        acceleration=(0.0, -gravity))
    model = BoundaryModelDummyParticles(tank.boundary.density, tank.boundary.mass, AdamiPressureExtrapolation(),
                                        kernel, h, state_equation=state_equation, clip_negative_pressure=True)
    wall = WallBoundarySystem(tank.boundary, model)
    hydrodynamic_densities = t(fluid_density) * np.ones(structure.nparticles, dtype=eltype)
    hydrodynamic_masses = (hydrodynamic_densities * t(ds) ** 2).astype(eltype)
    k_structure = gravity * initial_fluid_size[1]
    if structure_boundary_model == "dummy_particles":
        # the alternative the example keeps in a comment (dam_break_plate_2d.jl:120-133) and
        # examples/fsi/hydrostatic_water_column_2d.jl:109-124 uses
        model_structure = BoundaryModelDummyParticles(hydrodynamic_densities, hydrodynamic_masses,
                                                      structure_pressure_extrapolation or AdamiPressureExtrapolation(),
                                                      kernel, h, state_equation=state_equation)
    else:
        model_structure = BoundaryModelMonaghanKajtar(k_structure, dx / ds, ds, hydrodynamic_masses)
    structure_system = TotalLagrangianSPHSystem(
        structure, smoothing_kernel=WendlandC2Kernel(2), smoothing_length=np.sqrt(2) * ds, young_modulus=E,
        poisson_ratio=nu, boundary_model=model_structure, clamped_particles=range(clamped.nparticles),
        clamped_particles_motion=clamped_particles_motion,
        acceleration=(0.0, -gravity), penalty_force=PenaltyForceGanzenmueller(alpha=0.01))
    return fluid, wall, structure_system, tank


def dam_break_plate_3d(fluid_particle_spacing=0.02, *, n_particles_x=3, eltype=np.float64, coordinates_eltype=None,
                       initial_fluid_size=(0.16, 0.3, 0.12), plate_position=None, E=1e6, nu=0.3,
                       structure_boundary_model="monaghan_kajtar", clamped_particles_motion=None):
    """The set-up of examples/fsi/dam_break_plate_2d.jl extruded along z (BASELINE config 5 names "2D/3D"; the
    reference ships the plate example in 2-D only, its 3-D FSI example is falling_sphere_3d.jl): a water column in a
    closed tank next to an elastic plate that spans the tank's depth and is clamped along its base.  Same models as
    the 2-D example (WendlandC2, h = 1.75 dx fluid, sqrt(2) ds structure, Monaghan-Kajtar or dummy-particle coupling,
    PenaltyForceGanzenmueller).  Returns (fluid_system, boundary_system, structure_system, tank)."""
    t = np.dtype(eltype).type
    coordinates_eltype = coordinates_eltype or np.float64
    gravity = 9.81
    dx = fluid_particle_spacing
    tank_size = (4 * initial_fluid_size[0], 2 * initial_fluid_size[1], initial_fluid_size[2])
    fluid_density = 1000.0
    sound_speed = 20 * np.sqrt(gravity * initial_fluid_size[1])
    state_equation = StateEquationCole(sound_speed=float(t(sound_speed)), reference_density=fluid_density, exponent=1)
    acc = (0.0, -gravity, 0.0)
    tank = RectangularTank(dx, initial_fluid_size, tank_size, fluid_density, n_layers=3, spacing_ratio=1,
                           acceleration=acc, state_equation=state_equation,
                           coordinates_eltype=coordinates_eltype, eltype=eltype)
    length_beam, thickness, structure_density = 0.12, 0.03, 2500
    ds = thickness / (n_particles_x - 1)
    n_y = int(np.rint(length_beam / ds)) + 1
    n_z = int(np.rint(tank.tank_size[2] / ds))
    if plate_position is None:
        plate_position = (1.5 * initial_fluid_size[0], 0.0, 0.5 * (tank.tank_size[2] - (n_z - 1) * ds))
    px, py, pz = plate_position
    plate = RectangularShape(ds, (n_particles_x, n_y - 1, n_z), (px, py + ds, pz), density=structure_density,
                             place_on_shell=True, coordinates_eltype=coordinates_eltype, eltype=eltype)
    clamped = RectangularShape(ds, (n_particles_x, 1, n_z), (px, py, pz), density=structure_density,
                               place_on_shell=True, coordinates_eltype=coordinates_eltype, eltype=eltype)
    structure = union(clamped, plate)
    h = 1.75 * dx
    kernel = WendlandC2Kernel(3)
    fluid = WeaklyCompressibleSPHSystem(
        tank.fluid, smoothing_kernel=kernel, smoothing_length=h, density_calculator=ContinuityDensity(),
        state_equation=state_equation, viscosity=ArtificialViscosityMonaghan(alpha=0.02, beta=0.0), acceleration=acc)
    model = BoundaryModelDummyParticles(tank.boundary.density, tank.boundary.mass, AdamiPressureExtrapolation(),
                                        kernel, h, state_equation=state_equation, clip_negative_pressure=True)
    wall = WallBoundarySystem(tank.boundary, model)
    hyd_rho = t(fluid_density) * np.ones(structure.nparticles, dtype=eltype)
    hyd_mass = (hyd_rho * t(ds) ** 3).astype(eltype)
    if structure_boundary_model == "dummy_particles":
        model_structure = BoundaryModelDummyParticles(hyd_rho, hyd_mass, AdamiPressureExtrapolation(), kernel, h,
                                                      state_equation=state_equation)
    else:
        model_structure = BoundaryModelMonaghanKajtar(gravity * initial_fluid_size[1], dx / ds, ds, hyd_mass)
    structure_system = TotalLagrangianSPHSystem(
        structure, smoothing_kernel=WendlandC2Kernel(3), smoothing_length=np.sqrt(2) * ds, young_modulus=E,
        poisson_ratio=nu, boundary_model=model_structure, clamped_particles=range(clamped.nparticles),
        clamped_particles_motion=clamped_particles_motion, acceleration=acc,
        penalty_force=PenaltyForceGanzenmueller(alpha=0.01))
    return fluid, wall, structure_system, tank


def dam_break_gate_2d(fluid_particle_spacing=0.02, *, n_particles_x=4, eltype=np.float64, coordinates_eltype=np.float64):
    """examples/fsi/dam_break_gate_2d.jl:18-160: a water column behind a gate that is pulled up with
    y + (-285.115 t^3 + 72.305 t^2 + 0.1463 t) for t < 0.1, the released water hitting an elastic plate.  Four systems:
    fluid, tank (static dummy-particle wall), gate (moving dummy-particle wall), plate (TLSPH with the same dummy-particle
    boundary model).  Returns (fluid, tank_wall, gate_wall, plate, tank)."""
    from .model import PrescribedMotion
    t = np.dtype(eltype).type
    gravity = 9.81
    dx = fluid_particle_spacing
    boundary_layers, spacing_ratio = 3, 1
    bdx = dx / spacing_ratio
    initial_fluid_size, tank_size = (0.2, 0.4), (0.8, 0.8)
    fluid_density = 997.0
    sound_speed = 10 * np.sqrt(2 * gravity * initial_fluid_size[1])
    state_equation = StateEquationCole(sound_speed=float(t(sound_speed)), reference_density=fluid_density, exponent=7)
    tank = RectangularTank(dx, initial_fluid_size, tank_size, fluid_density, n_layers=boundary_layers,
                           spacing_ratio=spacing_ratio, acceleration=(0.0, -gravity), state_equation=state_equation,
                           coordinates_eltype=coordinates_eltype, eltype=eltype)
    gate_height = initial_fluid_size[1] + 4 * dx
    gate = RectangularShape(bdx, (boundary_layers, int(np.rint(gate_height / bdx))), (initial_fluid_size[0], 0.0),
                            density=fluid_density, coordinates_eltype=coordinates_eltype, eltype=eltype)
    lift = lambda tt: -285.115 * tt ** 3 + 72.305 * tt ** 2 + 0.1463 * tt
    motion = PrescribedMotion(
        lambda x, tt: x + np.array([0.0, lift(tt)])[None, :], lambda tt: tt < 0.1,
        velocity_function=lambda x, tt: np.broadcast_to(np.array([0.0, -855.345 * tt ** 2 + 144.61 * tt + 0.1463]), x.shape),
        acceleration_function=lambda x, tt: np.broadcast_to(np.array([0.0, -1710.69 * tt + 144.61]), x.shape))
    length_beam, thickness = 0.09, 0.004 * 10
    structure_density, E, nu = 1161.54, 3.5e6 / 10, 0.45
    ds = thickness / (n_particles_x - 1)
    n_particles_y = int(np.rint(length_beam / ds)) + 1
    plate_position = 0.6 - n_particles_x * ds
    plate = RectangularShape(ds, (n_particles_x, n_particles_y - 1), (plate_position, ds), density=structure_density,
                             place_on_shell=True, coordinates_eltype=coordinates_eltype, eltype=eltype)
    clamped = RectangularShape(ds, (n_particles_x, 1), (plate_position, 0.0), density=structure_density,
                               place_on_shell=True, coordinates_eltype=coordinates_eltype, eltype=eltype)
    structure = union(clamped, plate)
    h = 1.75 * dx
    kernel = WendlandC2Kernel(2)
    fluid = WeaklyCompressibleSPHSystem(
        tank.fluid, smoothing_kernel=kernel, smoothing_length=h, density_calculator=ContinuityDensity(),
        state_equation=state_equation, viscosity=ArtificialViscosityMonaghan(alpha=0.1, beta=0.0),
        acceleration=(0.0, -gravity))
    model = lambda dens, mass: BoundaryModelDummyParticles(dens, mass, AdamiPressureExtrapolation(), kernel, h,
                                                           state_equation=state_equation, clip_negative_pressure=True)
    tank_wall = WallBoundarySystem(tank.boundary, model(tank.boundary.density, tank.boundary.mass))
    gate_wall = WallBoundarySystem(gate, model(gate.density, gate.mass), prescribed_motion=motion)
    hyd_rho = t(fluid_density) * np.ones(structure.nparticles, dtype=eltype)
    hyd_mass = (hyd_rho * t(ds) ** 2).astype(eltype)
    plate_system = TotalLagrangianSPHSystem(
        structure, smoothing_kernel=WendlandC2Kernel(2), smoothing_length=np.sqrt(2) * ds, young_modulus=E,
        poisson_ratio=nu, boundary_model=model(hyd_rho, hyd_mass), clamped_particles=range(clamped.nparticles),
        acceleration=(0.0, -gravity))
    return fluid, tank_wall, gate_wall, plate_system, tank


def falling_sphere_2d(fluid_particle_spacing=0.02, *, eltype=np.float64, coordinates_eltype=np.float64,
                      initial_fluid_size=(1.0, 0.9), tank_size=(1.0, 1.0), sphere_center=(0.5, 1.6)):
    """examples/fsi/falling_sphere_2d.jl (= falling_spheres_2d.jl:18-140 with `structure_system_2 = nothing`): an
    elastic sphere (radius 0.3, density 500, E = 7e4, nu = 0) dropped into a tank without lid.  Tank and sphere carry
    dummy particles with BernoulliPressureExtrapolation: for the static tank that is Adami's extrapolation, for the
    sphere the dynamic pressure term always applies (dummy_particles.jl:696-707).
    Returns (fluid_system, boundary_system, structure_system, tank)."""
    from .model import BernoulliPressureExtrapolation
    from .setups import SphereShape
    t = np.dtype(eltype).type
    gravity = 9.81
    dx = ds = fluid_particle_spacing
    fluid_density = 1000.0
    sound_speed = 10 * np.sqrt(gravity * initial_fluid_size[1])
    state_equation = StateEquationCole(sound_speed=float(t(sound_speed)), reference_density=fluid_density, exponent=1)
    tank = RectangularTank(dx, initial_fluid_size, tank_size, fluid_density, n_layers=3, spacing_ratio=1,
                           faces=(True, True, True, False), acceleration=(0.0, -gravity), state_equation=state_equation,
                           coordinates_eltype=coordinates_eltype, eltype=eltype)
    sphere = SphereShape(ds, 0.3, sphere_center, 500.0, coordinates_eltype=coordinates_eltype, eltype=eltype)
    h = 1.5 * dx
    kernel = WendlandC2Kernel(2)
    fluid = WeaklyCompressibleSPHSystem(
        tank.fluid, smoothing_kernel=kernel, smoothing_length=h, density_calculator=ContinuityDensity(),
        state_equation=state_equation, viscosity=ArtificialViscosityMonaghan(alpha=0.02, beta=0.0),
        density_diffusion=DensityDiffusionMolteniColagrossi(delta=0.1), acceleration=(0.0, -gravity))
    model = lambda dens, mass: BoundaryModelDummyParticles(dens, mass, BernoulliPressureExtrapolation(), kernel, h,
                                                           state_equation=state_equation, clip_negative_pressure=True)
    wall = WallBoundarySystem(tank.boundary, model(tank.boundary.density, tank.boundary.mass))
    hyd_rho = t(fluid_density) * np.ones(sphere.nparticles, dtype=eltype)
    hyd_mass = (hyd_rho * t(ds) ** 2).astype(eltype)
    structure = TotalLagrangianSPHSystem(
        sphere, smoothing_kernel=WendlandC2Kernel(2), smoothing_length=np.sqrt(2) * ds, young_modulus=7e4,
        poisson_ratio=0.0, acceleration=(0.0, -gravity), boundary_model=model(hyd_rho, hyd_mass),
        penalty_force=PenaltyForceGanzenmueller(alpha=0.3))
    return fluid, wall, structure, tank


def falling_spheres_2d(fluid_particle_spacing=0.02, *, eltype=np.float64, coordinates_eltype=np.float64,
                       sphere1_center=(0.5, 1.6), sphere2_center=(1.5, 1.6)):
    """examples/fsi/falling_spheres_2d.jl:18-140: two elastic spheres of different size, density and stiffness
    (radius 0.3 / 0.2, density 500 / 1100, E = 7e4 / 1e5) dropped into a tank -- two `TotalLagrangianSPHSystem`s.
    Returns (fluid_system, boundary_system, structure_system_1, structure_system_2, tank)."""
    from .model import BernoulliPressureExtrapolation
    from .setups import SphereShape
    t = np.dtype(eltype).type
    fluid, wall, sphere1, tank = falling_sphere_2d(fluid_particle_spacing, eltype=eltype,
                                                   coordinates_eltype=coordinates_eltype, initial_fluid_size=(2.0, 0.9),
                                                   tank_size=(2.0, 1.0), sphere_center=sphere1_center)
    ds = fluid_particle_spacing
    ball = SphereShape(ds, 0.2, sphere2_center, 1100.0, coordinates_eltype=coordinates_eltype, eltype=eltype)
    m1 = sphere1.boundary_model
    hyd_rho = t(1000.0) * np.ones(ball.nparticles, dtype=eltype)
    model = BoundaryModelDummyParticles(hyd_rho, (hyd_rho * t(ds) ** 2).astype(eltype), BernoulliPressureExtrapolation(),
                                        m1.smoothing_kernel, m1.smoothing_length, state_equation=m1.state_equation,
                                        clip_negative_pressure=True)
    sphere2 = TotalLagrangianSPHSystem(
        ball, smoothing_kernel=WendlandC2Kernel(2), smoothing_length=np.sqrt(2) * ds, young_modulus=1e5,
        poisson_ratio=0.0, acceleration=(0.0, -9.81), boundary_model=model,
        penalty_force=PenaltyForceGanzenmueller(alpha=0.3))
    return fluid, wall, sphere1, sphere2, tank


def oscillating_beam_2d(n_particles_y=5, *, eltype=np.float64, coordinates_eltype=np.float64, penalty_force=None,
                        thickness=0.02, gravity=2.0, boundary_model_factory=None):
    """examples/structure/oscillating_beam_2d.jl:13-92: an elastic beam (0.35 x 0.02, E = 1.4e6, nu = 0.4) clamped in
    a disc of fixed particles, swinging under gravity 2.0 -- a structure-only semidiscretization.  The validation run
    (validation/oscillating_beam_2d/validation_oscillating_beam_2d.jl) adds PenaltyForceGanzenmueller(alpha=0.01) and
    records the deflection of the particle in the middle of the free end.
    Returns (structure_system, info) with info["mid_particle"] the 0-based index of that particle in the system."""
    from .setups import SphereShape
    length = 0.35
    density, E, nu = 1000.0, 1.4e6, 0.4
    clamp_radius = 0.05
    ds = thickness / (n_particles_y - 1)
    clamped = SphereShape(ds, clamp_radius + ds / 2, (0.0, thickness / 2), density, cutout_min=(0.0, 0.0),
                          cutout_max=(clamp_radius, thickness), place_on_shell=True,
                          coordinates_eltype=coordinates_eltype, eltype=eltype)
    n_clamp_x = int(np.rint(clamp_radius / ds))
    n_per_dim = (int(np.rint(length / ds)) + n_clamp_x + 1, n_particles_y)
    beam = RectangularShape(ds, n_per_dim, (0.0, 0.0), density=density, place_on_shell=True,
                            coordinates_eltype=coordinates_eltype, eltype=eltype)
    structure = union(clamped, beam)
    system = TotalLagrangianSPHSystem(
        structure, smoothing_kernel=WendlandC2Kernel(2), smoothing_length=np.sqrt(2) * ds, young_modulus=E,
        poisson_ratio=nu, clamped_particles=range(clamped.nparticles), acceleration=(0.0, -gravity),
        penalty_force=penalty_force,
        boundary_model=None if boundary_model_factory is None else boundary_model_factory(structure, ds))
    # middle_particle_id (1-based, in the beam = in the system, whose clamped particles sit behind the beam's)
    mid = n_per_dim[0] * (n_per_dim[1] + 1) // 2
    info = dict(mid_particle=mid - 1, start_position=beam.coordinates[mid - 1].astype(np.float64), particle_spacing=ds,
                n_particles_per_dimension=n_per_dim)
    return system, info


def falling_water_column_fsi_2d(n_particles_y=5, *, eltype=np.float64, coordinates_eltype=np.float64):
    """examples/fsi/falling_water_column_2d.jl:13-85: a block of water (0.525 x 1.0125, three structure spacings per
    fluid particle) dropped onto the clamped beam of oscillating_beam_2d.jl (thickness 0.05); no tank -- the water
    runs off into the void, so the caller supplies a bounding box.  Cubic-spline fluid kernel, Monaghan-Kajtar
    coupling.  Returns (fluid_system, structure_system, info)."""
    t = np.dtype(eltype).type
    gravity = 9.81
    initial_fluid_size = (0.525, 1.0125)
    fluid_density = 1000.0
    info_ds = 0.05 / (n_particles_y - 1)
    dx = 3 * info_ds
    sound_speed = 10 * np.sqrt(gravity * initial_fluid_size[1])
    state_equation = StateEquationCole(sound_speed=float(t(sound_speed)), reference_density=fluid_density, exponent=7)
    n_f = tuple(int(np.rint(sz / dx)) for sz in initial_fluid_size)
    water = RectangularShape(dx, n_f, (0.1, 0.2), density=fluid_density, coordinates_eltype=coordinates_eltype,
                             eltype=eltype)
    fluid = WeaklyCompressibleSPHSystem(
        water, smoothing_kernel=SchoenbergCubicSplineKernel(2), smoothing_length=1.2 * dx,
        density_calculator=ContinuityDensity(), state_equation=state_equation,
        viscosity=ArtificialViscosityMonaghan(alpha=0.02, beta=0.0), acceleration=(0.0, -gravity))
    k = gravity * initial_fluid_size[1]

    def model(structure, ds):
        hyd_mass = (t(fluid_density) * np.ones(structure.nparticles, dtype=eltype) * t(ds) ** 2).astype(eltype)
        return BoundaryModelMonaghanKajtar(k, dx / ds, ds, hyd_mass)

    beam, info = oscillating_beam_2d(n_particles_y, eltype=eltype, coordinates_eltype=coordinates_eltype,
                                     thickness=0.05, gravity=gravity, boundary_model_factory=model)
    return fluid, beam, info


def hydrostatic_water_column_fsi_2d(n_particles_plate_y=3, *, eltype=np.float64, coordinates_eltype=None,
                                    initial_fluid_size=(1.0, 2.0), plate_size=(1.0, 0.05), E=67.5e9, nu=0.3,
                                    sound_speed=50.0, damping_coefficient=0.05):
    """examples/fsi/hydrostatic_water_column_2d.jl:13-177 with `use_edac = false`: a water column resting on an
    elastic aluminium plate clamped at both ends.  WCSPH fluid (ContinuityDensity, Molteni-Colagrossi diffusion,
    SourceTermDamping), tank with side walls only, TLSPH plate whose particles are dummy particles with
    AdamiPressureExtrapolation towards the fluid.  The validation run of the reference
    (validation/hydrostatic_water_column_2d/validation.jl, test/validation/validation.jl:95-109) compares the
    mid-plate deflection with `analytical_value`.
    Returns (structure_system, fluid_system, boundary_system, info) -- the system order of the example."""
    from .model import DensityDiffusionMolteniColagrossi, SourceTermDamping
    t = np.dtype(eltype).type
    coordinates_eltype = coordinates_eltype or np.float64
    gravity, boundary_layers, spacing_ratio = 9.81, 3, 1
    fluid_density, structure_density = 1000.0, 2700.0
    ds = plate_size[1] / (n_particles_plate_y - 1)                    # structure = fluid particle spacing
    dx = ds
    D = E * plate_size[1] ** 3 / (12 * (1 - nu ** 2))
    analytical_value = -0.0026 * gravity * (fluid_density * initial_fluid_size[1] +
                                            structure_density * plate_size[1]) / D
    n_plate_x = int(np.rint(plate_size[0] / ds + 1))
    shape = lambda n, mc: RectangularShape(ds, n, mc, density=structure_density, place_on_shell=True,
                                           coordinates_eltype=coordinates_eltype, eltype=eltype)
    plate = shape((n_plate_x, n_particles_plate_y), (0.0, -plate_size[1]))
    left_wall = shape((3, n_particles_plate_y), (-3 * ds, -plate_size[1]))
    right_wall = shape((3, n_particles_plate_y), (plate_size[0] + ds, -plate_size[1]))
    fixed = union(left_wall, right_wall)
    geometry = union(fixed, plate)
    kernel = WendlandC2Kernel(2)
    h_s = h_f = np.sqrt(2) * ds
    state_equation = StateEquationCole(sound_speed=float(t(sound_speed)), reference_density=fluid_density, exponent=7,
                                       clip_negative_pressure=False)
    tank = RectangularTank(dx, initial_fluid_size, (plate_size[0], 3.0), fluid_density,
                           min_coordinates=(0.0, dx / 2), n_layers=boundary_layers, spacing_ratio=spacing_ratio,
                           faces=(True, True, False, False), acceleration=(0.0, -gravity),
                           state_equation=state_equation, coordinates_eltype=coordinates_eltype, eltype=eltype)
    fluid = WeaklyCompressibleSPHSystem(
        tank.fluid, smoothing_kernel=kernel, smoothing_length=h_f, density_calculator=ContinuityDensity(),
        state_equation=state_equation, density_diffusion=DensityDiffusionMolteniColagrossi(delta=0.1),
        acceleration=(0.0, -gravity), source_terms=SourceTermDamping(damping_coefficient=damping_coefficient))
    model = BoundaryModelDummyParticles(tank.boundary.density, tank.boundary.mass, AdamiPressureExtrapolation(),
                                        kernel, h_f, state_equation=state_equation)
    wall = WallBoundarySystem(tank.boundary, model)
    hyd_rho = t(fluid_density) * np.ones(geometry.nparticles, dtype=eltype)
    hyd_mass = (hyd_rho * t(ds) ** 2).astype(eltype)
    model_structure = BoundaryModelDummyParticles(hyd_rho, hyd_mass, AdamiPressureExtrapolation(), kernel, h_s,
                                                  state_equation=state_equation)
    structure = TotalLagrangianSPHSystem(
        geometry, smoothing_kernel=kernel, smoothing_length=h_s, young_modulus=E, poisson_ratio=nu,
        boundary_model=model_structure, clamped_particles=range(fixed.nparticles), acceleration=(0.0, -gravity))
    # validation.jl:15-21: the particle in the middle of the plate (1-based, in the system's own order: the
    # clamped particles are moved behind the plate's)
    mid = int(n_plate_x * (n_particles_plate_y + 1) / 2 - (n_plate_x + 1) / 2 + 1)
    info = dict(analytical_value=analytical_value, mid_particle=mid - 1, plate_size=plate_size, tank=tank,
                n_particles_plate=(n_plate_x, n_particles_plate_y))
    return structure, fluid, wall, info
