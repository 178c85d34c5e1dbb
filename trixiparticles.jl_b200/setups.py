"""Particle generators: `RectangularShape`, `RectangularTank`, `InitialCondition`.

Host-side restatement (numpy) of the reference's pre-processing that the hot path's
inputs come from; runs once per simulation and is not on the accelerated path.
Reference: src/setups/rectangular_shape.jl:79-267, src/setups/rectangular_tank.jl:99-200,
:406-1100, src/general/initial_condition.jl:190-215.

Array convention of this package: particle-major `(n, ND)` C-contiguous arrays, which is
bit-for-bit the memory layout of Julia's column-major `ND x n` matrices.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np


@dataclass
class InitialCondition:
    """initial_condition.jl: coordinates (n, ND) [coordinates_eltype], velocity (n, ND),
    mass (n,), density (n,), pressure (n,) [eltype]."""
    coordinates: np.ndarray
    velocity: np.ndarray
    mass: np.ndarray
    density: np.ndarray
    pressure: np.ndarray
    particle_spacing: float

    @property
    def ndims(self) -> int:
        return self.coordinates.shape[1]

    @property
    def nparticles(self) -> int:
        return self.coordinates.shape[0]

    @property
    def eltype(self):
        return self.mass.dtype

    @property
    def coordinates_eltype(self):
        return self.coordinates.dtype


def union(*ics: InitialCondition) -> InitialCondition:
    """`union(ic1, ic2, ...)` without duplicate removal (callers pass disjoint shapes)."""
    return InitialCondition(
        coordinates=np.concatenate([ic.coordinates for ic in ics]),
        velocity=np.concatenate([ic.velocity for ic in ics]),
        mass=np.concatenate([ic.mass for ic in ics]),
        density=np.concatenate([ic.density for ic in ics]),
        pressure=np.concatenate([ic.pressure for ic in ics]),
        particle_spacing=ics[0].particle_spacing)


def _loop_permutation(loop_order, ndims):
    # rectangular_shape.jl:134-187
    if ndims == 2:
        if loop_order in (None, "y_first"):
            return (0, 1)
        if loop_order == "x_first":
            return (1, 0)
    elif ndims == 3:
        if loop_order in (None, "z_first"):
            return (0, 1, 2)
        if loop_order == "y_first":
            return (1, 0, 2)
        if loop_order == "x_first":
            return (2, 1, 0)
    raise ValueError(f"{loop_order} is not a valid loop order")


def _permuted_indices(n_per_dim, loop_order, x_range=None):
    """1-based Cartesian index of every particle in storage order.

    `permutedims(CartesianIndices(n), perm)` iterated column-major: the dimension
    perm[0] runs fastest (rectangular_shape.jl:205-207).  `x_range = (i0, i1)`: only the particles
    whose 0-based index along the first coordinate lies in [i0, i1), in the same relative order
    (a slab of the lattice)."""
    ndims = len(n_per_dim)
    perm = _loop_permutation(loop_order, ndims)
    i0, i1 = (0, n_per_dim[0]) if x_range is None else x_range
    sizes = list(n_per_dim)
    sizes[0] = max(i1 - i0, 0)
    shape = [sizes[p] for p in perm]
    # column-major iteration over `shape`
    grids = np.meshgrid(*[np.arange(1, s + 1) for s in shape], indexing="ij")
    flat = [g.ravel(order="F") for g in grids]
    idx = np.empty((flat[0].size, ndims), dtype=np.int64)
    for k, p in enumerate(perm):
        idx[:, p] = flat[k] + (i0 if p == 0 else 0)
    return idx


def _x_index_range(particle_spacing, n0, min0, shift0, x_window, coordinates_eltype):
    """0-based index range [i0, i1) along x of the lattice columns whose final x coordinate
    (`coordinates_eltype`, after the tank's `min_coordinates` shift) lies in x_window = [lo, hi)."""
    ct = np.dtype(coordinates_eltype).type
    xs = np.float64(min0) + np.float64(ct(particle_spacing)) * (np.arange(1, n0 + 1, dtype=np.float64) - 0.5)
    xs = (xs.astype(ct) + ct(shift0)).astype(ct)
    keep = np.nonzero((xs >= x_window[0]) & (xs < x_window[1]))[0]
    if keep.size == 0:
        return 0, 0
    return int(keep[0]), int(keep[-1]) + 1


def rectangular_shape_coords(particle_spacing, n_particles_per_dimension, min_coordinates,
                             loop_order=None, coordinates_eltype=np.float64, x_range=None,
                             place_on_shell=False):
    """rectangular_shape.jl:189-224: min_coordinates + spacing * (index - 0.5); `place_on_shell`:
    the first particle sits AT min_coordinates (structures; rectangular_shape.jl:205-209)."""
    ct = np.dtype(coordinates_eltype).type
    idx = _permuted_indices(tuple(n_particles_per_dimension), loop_order, x_range)
    spacing = ct(particle_spacing)
    mins = np.asarray(min_coordinates, dtype=np.float64)
    if place_on_shell:
        mins = mins - 0.5 * np.float64(spacing)
    # Julia: min_coordinates (Float64 tuple) .+ particle_spacing .* (index .- 0.5)
    coords = mins[None, :] + np.float64(spacing) * (idx.astype(np.float64) - 0.5)
    if ct is np.float32:
        # Float32 spacing times Float64 (index - 0.5) promotes to Float64, then stored as cT
        return coords.astype(np.float32)
    return coords


def _initialize_pressure(n_per_dim, particle_spacing, acceleration, density_fun, loop_order, x_range=None):
    """rectangular_shape.jl:226-267: 1-D explicit-Euler hydrostatic column (Float64)."""
    eps = np.finfo(np.float64).eps
    acc = np.asarray(acceleration, dtype=np.float64)
    active = np.nonzero(np.abs(acc) > eps)[0]
    if active.size > 1:
        raise ValueError("hydrostatic pressure calculation is not supported with diagonal acceleration")
    if active.size == 0:
        n0 = n_per_dim[0] if x_range is None else max(x_range[1] - x_range[0], 0)
        return np.zeros(int(np.prod(n_per_dim[1:])) * n0)
    accel_dim = int(active[0])
    factor = float(particle_spacing) * abs(acc[accel_dim])
    n = n_per_dim[accel_dim]
    p1d = np.zeros(n)
    p1d[0] = 0.5 * factor * density_fun(0.0)
    for i in range(n - 1):
        p1d[i + 1] = p1d[i] + factor * density_fun(p1d[i])
    if acc[accel_dim] < 0:
        p1d = p1d[::-1].copy()
    idx = _permuted_indices(tuple(n_per_dim), loop_order, x_range)
    return p1d[idx[:, accel_dim] - 1]


def RectangularShape(particle_spacing, n_particles_per_dimension, min_coordinates, *,
                     velocity=None, mass=None, density=None, pressure=0.0, acceleration=None,
                     state_equation=None, coordinates_eltype=np.float64, loop_order=None,
                     eltype=np.float64, x_range=None, place_on_shell=False) -> InitialCondition:
    """rectangular_shape.jl:79-151.  `eltype` plays the role of `eltype(particle_spacing)`.
    `x_range`: only the lattice columns [i0, i1) along x (see `_permuted_indices`)."""
    t = np.dtype(eltype)
    n_per_dim = tuple(int(n) for n in n_particles_per_dimension)
    ndims = len(n_per_dim)
    coords = rectangular_shape_coords(particle_spacing, n_per_dim, min_coordinates,
                                      loop_order=loop_order, coordinates_eltype=coordinates_eltype,
                                      x_range=x_range, place_on_shell=place_on_shell)
    n = coords.shape[0]
    if acceleration is not None:
        if state_equation is None:
            if density is None:
                raise ValueError("`density` must be specified")
            dens = float(density)
            density_fun = lambda p: dens
        else:
            if density is not None:
                raise ValueError("`density` cannot be used together with `acceleration` and `state_equation`")
            density_fun = lambda p: state_equation.inverse(p, dtype=np.float64)
        press = _initialize_pressure(n_per_dim, particle_spacing, acceleration, density_fun, loop_order, x_range)
        press = press.astype(t)  # Vector{ELTYPE}
        if state_equation is not None:
            # one inverse per distinct pressure (one per lattice layer), not per particle
            levels, inv = np.unique(press, return_inverse=True)
            densities = np.array([state_equation.inverse(p, dtype=t) for p in levels], dtype=t)[inv]
        else:
            densities = np.full(n, density, dtype=t)
    else:
        if state_equation is not None:
            raise ValueError("`state_equation` must be used together with `acceleration`")
        if density is None:
            raise ValueError("`density` must be specified when not using `acceleration` and `state_equation`")
        press = np.full(n, pressure, dtype=t)
        densities = np.full(n, density, dtype=t) if np.isscalar(density) else np.asarray(density, dtype=t)
    vel = np.zeros((n, ndims), dtype=t)
    if velocity is not None:
        vel[:] = np.asarray(velocity, dtype=t)[None, :]
    if mass is None:
        # initial_condition.jl:204-205: particle_volume = particle_spacing^NDIMS; m = V * rho
        vol = t.type(particle_spacing) ** ndims
        masses = (vol * densities).astype(t)
    else:
        masses = np.full(n, mass, dtype=t) if np.isscalar(mass) else np.asarray(mass, dtype=t)
    return InitialCondition(coords, vel, masses, densities, press, float(particle_spacing))


def SphereShape(particle_spacing, radius, center_position, density, *, n_layers=-1, layer_outwards=False,
                cutout_min=None, cutout_max=None, place_on_shell=False, velocity=None, mass=None, pressure=0.0,
                coordinates_eltype=np.float64, eltype=np.float64) -> InitialCondition:
    """sphere_shape.jl:97-134 with `sphere_type = VoxelSphere()` (:183-237): the points center + spacing * (i, j[, k])
    with inner_radius + 10 eps < |x - center| <= outer_radius + 10 eps, the first index running fastest
    (CartesianIndices), minus those inside the closed box [cutout_min, cutout_max]."""
    t = np.dtype(eltype)
    ct = np.dtype(coordinates_eltype)
    center = np.asarray(center_position, dtype=np.float64)
    ndims = center.size
    dx = float(ct.type(particle_spacing))
    if n_layers > 0:
        if layer_outwards:
            inner, outer = radius, radius + n_layers * dx
            if not place_on_shell:
                inner, outer = inner + dx / 2, outer + dx / 2
        else:
            inner, outer = radius - n_layers * dx, radius
            if not place_on_shell:
                inner, outer = inner - dx / 2, outer - dx / 2
    else:
        inner, outer = -1.0, radius
        if not place_on_shell:
            outer -= dx / 2
    n_cube = int(np.rint(outer / dx))
    r = np.arange(-n_cube, n_cube + 1)
    idx = np.stack(np.meshgrid(*([r] * ndims), indexing="ij"), axis=-1).reshape(-1, ndims)
    # CartesianIndices: the first index fastest
    order = np.lexsort(tuple(idx[:, d] for d in range(ndims)))
    idx = idx[order]
    x = center[None, :] + dx * idx
    dist = np.linalg.norm(x - center[None, :], axis=1)
    eps = np.finfo(np.float64).eps
    keep = (inner + 10 * eps < dist) & (dist <= outer + 10 * eps)
    x = x[keep]
    if cutout_min is not None and cutout_max is not None:
        lo, hi = np.asarray(cutout_min, dtype=np.float64), np.asarray(cutout_max, dtype=np.float64)
        if np.linalg.norm(hi - lo) > eps:
            inside = np.all((lo[None, :] <= x) & (x <= hi[None, :]), axis=1)
            x = x[~inside]
    n = x.shape[0]
    dens = np.full(n, density, dtype=t)
    masses = (t.type(particle_spacing) ** ndims * dens).astype(t) if mass is None else np.full(n, mass, dtype=t)
    vel = np.zeros((n, ndims), dtype=t)
    if velocity is not None:
        vel[:] = np.asarray(velocity, dtype=t)[None, :]
    return InitialCondition(np.ascontiguousarray(x.astype(ct)), vel, masses, dens, np.full(n, pressure, dtype=t),
                            float(particle_spacing))


def _round_n_particles(size, spacing, t=np.float64):
    # rectangular_tank.jl:406-416 (Julia `round` = ties-to-even, as np.rint); evaluated in
    # ELTYPE like the reference (`size` and `spacing` are ELTYPE values there)
    t = np.dtype(t).type
    n = int(np.rint(t(size) / t(spacing)))
    return n, float(t(n) * t(spacing))


@dataclass
class RectangularTankResult:
    fluid: InitialCondition
    boundary: InitialCondition
    fluid_size: tuple
    tank_size: tuple
    n_layers: int
    n_particles_per_dimension: tuple
    face_ranges: dict = field(default_factory=dict)
    # `face_indices[f]`: 0-based indices (into `boundary`) of the particles of face f's own block (edges and
    # corners excluded), faces numbered left, right, bottom, top(, front, back) from 0 (rectangular_tank.jl:93,
    # :535-620, :701-832); empty when the tank was built with an x window
    face_indices: tuple = ()
    particle_spacing: float = 0.0
    spacing_ratio: float = 1.0


def _tank_boundary_blocks(ndims, spacing, tank_size, n_b, n_layers, faces):
    """Order of the boundary blocks: rectangular_tank.jl:530-690 (2D), :693-1100 (3D).
    Returns a list of (n_per_dim, min_coords, loop_order)."""
    L = n_layers
    off = -L * spacing
    tx, ty = tank_size[0], tank_size[1]
    blocks = []
    if ndims == 2:
        nx, ny = n_b
        left, right, bottom, top = faces
        if left: blocks.append(((L, ny), (off, 0.0), "x_first"))
        if right: blocks.append(((L, ny), (tx, 0.0), "x_first"))
        if bottom: blocks.append(((nx, L), (0.0, off), "y_first"))
        if top: blocks.append(((nx, L), (0.0, ty), "y_first"))
        if left and bottom: blocks.append(((L, L), (off, off), None))
        if left and top: blocks.append(((L, L), (off, ty), None))
        if right and bottom: blocks.append(((L, L), (tx, off), None))
        if right and top: blocks.append(((L, L), (tx, ty), None))
        return blocks
    nx, ny, nz = n_b
    tz = tank_size[2]
    left, right, bottom, top, front, back = faces
    # faces
    if left: blocks.append(((L, ny, nz), (off, 0.0, 0.0), "x_first"))
    if right: blocks.append(((L, ny, nz), (tx, 0.0, 0.0), "x_first"))
    if bottom: blocks.append(((nx, L, nz), (0.0, off, 0.0), "y_first"))
    if top: blocks.append(((nx, L, nz), (0.0, ty, 0.0), "y_first"))
    if front: blocks.append(((nx, ny, L), (0.0, 0.0, off), "z_first"))
    if back: blocks.append(((nx, ny, L), (0.0, 0.0, tz), "z_first"))
    # edges
    if left and bottom: blocks.append(((L, L, nz), (off, off, 0.0), None))
    if left and top: blocks.append(((L, L, nz), (off, ty, 0.0), None))
    if right and bottom: blocks.append(((L, L, nz), (tx, off, 0.0), None))
    if right and top: blocks.append(((L, L, nz), (tx, ty, 0.0), None))
    if front and bottom: blocks.append(((nx, L, L), (0.0, off, off), None))
    if front and top: blocks.append(((nx, L, L), (0.0, ty, off), None))
    if back and bottom: blocks.append(((nx, L, L), (0.0, off, tz), None))
    if back and top: blocks.append(((nx, L, L), (0.0, ty, tz), None))
    if left and front: blocks.append(((L, ny, L), (off, 0.0, off), None))
    if left and back: blocks.append(((L, ny, L), (off, 0.0, tz), None))
    if right and front: blocks.append(((L, ny, L), (tx, 0.0, off), None))
    if right and back: blocks.append(((L, ny, L), (tx, 0.0, tz), None))
    # corners
    for cond, mc in (
        (left and bottom and front, (off, off, off)), (left and top and front, (off, ty, off)),
        (left and bottom and back, (off, off, tz)), (left and top and back, (off, ty, tz)),
        (right and bottom and front, (tx, off, off)), (right and top and front, (tx, ty, off)),
        (right and bottom and back, (tx, off, tz)), (right and top and back, (tx, ty, tz)),
    ):
        if cond:
            blocks.append(((L, L, L), mc, None))
    return blocks


def RectangularTank(particle_spacing, fluid_size: Sequence[float], tank_size: Sequence[float],
                    fluid_density, *, velocity=None, pressure=0.0, acceleration=None,
                    state_equation=None, boundary_density=None, n_layers=1, spacing_ratio=1,
                    min_coordinates=None, faces=None, coordinates_eltype=np.float64,
                    eltype=np.float64, x_window=None, boundary_x_window=None) -> RectangularTankResult:
    """rectangular_tank.jl:99-200.  `x_window = (lo, hi)` (not in the reference): only the particles
    with lo <= x < hi are generated -- exactly the rows of the full tank with that property, in the
    same relative order -- so that a rank of a slab-decomposed run never builds the whole lattice
    (`boundary_x_window`: a different window for the boundary particles; default: `x_window`)."""
    t = np.dtype(eltype)
    ndims = len(fluid_size)
    if faces is None:
        faces = (True,) * (2 * ndims)
    if boundary_density is None:
        boundary_density = fluid_density
    if min_coordinates is None:
        min_coordinates = np.zeros(ndims)
    spacing = float(t.type(particle_spacing))
    fluid_size_ = [float(t.type(s)) for s in fluid_size]
    tank_size_ = [float(t.type(s)) for s in tank_size]

    n_f = []
    for d in range(ndims):
        n, new = _round_n_particles(fluid_size_[d], spacing, t)
        n_f.append(n)
        fluid_size_[d] = new
    for d in range(ndims):
        if np.isclose(fluid_size[d], tank_size[d]):
            tank_size_[d] = fluid_size_[d]
    n_b = []
    b_spacing = spacing / spacing_ratio
    for d in range(ndims):
        n, new = _round_n_particles(tank_size_[d], b_spacing, t)
        n_b.append(n)
        tank_size_[d] = new

    blocks = _tank_boundary_blocks(ndims, b_spacing, tank_size_, n_b, n_layers, faces)
    shift0 = float(np.asarray(min_coordinates, dtype=np.float64)[0])
    xr = lambda sp, n0, mc0, win=x_window: (None if win is None else
                                            _x_index_range(sp, n0, mc0, shift0, win, coordinates_eltype))
    b_win = x_window if boundary_x_window is None else boundary_x_window
    parts = [rectangular_shape_coords(b_spacing, npd, mc, loop_order=lo, coordinates_eltype=coordinates_eltype,
                                      x_range=xr(b_spacing, npd[0], mc[0], b_win))
             for npd, mc, lo in blocks if int(np.prod(npd)) > 0]
    face_indices = ()
    if b_win is None:
        # the face blocks come first, in face order (only the faces that exist)
        sizes = [int(np.prod(npd)) for npd, mc, lo in blocks if int(np.prod(npd)) > 0]
        starts = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        fi, k = [], 0
        for present in faces:
            if present and k < len(sizes):
                fi.append(np.arange(starts[k], starts[k + 1], dtype=np.int64))
                k += 1
            else:
                fi.append(np.zeros(0, dtype=np.int64))
        face_indices = tuple(fi)
    if parts:
        b_coords = np.concatenate(parts)
    else:
        b_coords = np.zeros((0, ndims), dtype=coordinates_eltype)
    nb = b_coords.shape[0]
    b_mass = (t.type(boundary_density) * t.type(b_spacing) ** ndims * np.ones(nb, dtype=t)).astype(t)
    b_dens = (t.type(boundary_density) * np.ones(nb, dtype=t)).astype(t)
    b_coords = (b_coords + np.asarray(min_coordinates, dtype=b_coords.dtype)[None, :]).astype(coordinates_eltype)
    boundary = InitialCondition(b_coords, np.zeros((nb, ndims), dtype=t), b_mass, b_dens,
                                np.zeros(nb, dtype=t), b_spacing)

    # check_tank_overlap (rectangular_tank.jl:474-527)
    for d in range(ndims):
        if tank_size_[d] < fluid_size_[d] - 1e-5 * spacing:
            n_f[d] -= 1
            fluid_size_[d] = n_f[d] * spacing

    if np.linalg.norm(fluid_size) > np.finfo(float).eps:
        if state_equation is not None:
            fluid = RectangularShape(spacing, n_f, np.zeros(ndims), velocity=velocity,
                                     pressure=pressure, acceleration=acceleration,
                                     state_equation=state_equation,
                                     coordinates_eltype=coordinates_eltype, eltype=t,
                                     x_range=xr(spacing, n_f[0], 0.0))
        else:
            fluid = RectangularShape(spacing, n_f, np.zeros(ndims), density=fluid_density,
                                     velocity=velocity, pressure=pressure,
                                     acceleration=acceleration,
                                     coordinates_eltype=coordinates_eltype, eltype=t,
                                     x_range=xr(spacing, n_f[0], 0.0))
        fluid.coordinates = (fluid.coordinates + np.asarray(min_coordinates, dtype=fluid.coordinates.dtype)[None, :]).astype(coordinates_eltype)
    else:
        fluid = InitialCondition(np.zeros((0, ndims), dtype=coordinates_eltype),
                                 np.zeros((0, ndims), dtype=t), np.zeros(0, dtype=t),
                                 np.zeros(0, dtype=t), np.zeros(0, dtype=t), spacing)
    return RectangularTankResult(fluid, boundary, tuple(fluid_size_), tuple(tank_size_),
                                 n_layers, tuple(n_f), face_indices=face_indices, particle_spacing=spacing,
                                 spacing_ratio=spacing_ratio)


def reset_wall_(tank: RectangularTankResult, reset_faces, positions):
    """`reset_wall!(tank, reset_faces, positions)` (rectangular_tank.jl:1160-1191): moves the particles of the
    given faces so that the face's inner surface is at `positions[face]`; layer l (1 = next to the fluid) lands at
    positions + (l - 1) dx + dx/2 for the even (right/top/back) faces, positions - (l - 1) dx - dx/2 for the odd."""
    dx = tank.particle_spacing / tank.spacing_ratio
    x = tank.boundary.coordinates
    for face, do in enumerate(reset_faces):
        if not do or face >= len(tank.face_indices) or tank.face_indices[face].size == 0:
            continue
        dim, idx = face // 2, tank.face_indices[face]
        c = x[idx, dim].astype(np.float64)
        if face % 2 == 1:   # right / top / back: layers count away from the tank's far side
            layer = np.rint((c - c.min()) / dx)
            x[idx, dim] = (positions[face] + layer * dx + 0.5 * dx).astype(x.dtype)
        else:
            layer = np.rint((c.max() - c) / dx)
            x[idx, dim] = (positions[face] - layer * dx - dx + 0.5 * dx).astype(x.dtype)
    return tank
