"""`Semidiscretization`, `semidiscretize`, `kick!`, `drift!` on the B200 library.

Python mirror of /root/reference/src/general/semidiscretization.jl (constructor :106-191,
ODE layout :128-135, `semidiscretize` :293-396, `drift!` :522-536, `kick!` :589-612) for the
accelerated path: one `WeaklyCompressibleSPHSystem` plus an optional `WallBoundarySystem`.
`kick!`/`drift!` are spelled `kick_`/`drift_` (Python has no `!`).  All particle arithmetic
runs in libtpb200.so through the C ABI of include/tpb200.h; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from types import SimpleNamespace
from typing import Optional, Sequence

import numpy as np

from . import _lib
from .model import (compact_support, AdamiPressureExtrapolation, ArtificialViscosityMonaghan, BoundaryModelDummyParticles,
                    BoundaryModelMonaghanKajtar,
                    ContinuityDensity,
                    DensityDiffusionMolteniColagrossi, SourceTermDamping, StateEquationAdaptiveCole,
                    SummationDensity, TotalLagrangianSPHSystem, WallBoundarySystem,
                    WeaklyCompressibleSPHSystem)


@dataclass
class FullGridCellList:
    """PointNeighbors `FullGridCellList(; min_corner, max_corner, max_points_per_cell=100)`."""
    min_corner: Sequence[float]
    max_corner: Sequence[float]
    max_points_per_cell: int = 100


@dataclass
class GridNeighborhoodSearch:
    """PointNeighbors `GridNeighborhoodSearch{NDIMS}(; cell_list, update_strategy)`."""
    ndims: int
    cell_list: Optional[FullGridCellList] = None
    update_strategy: object = None  # accepted for API parity; the GPU rebuild is always parallel


@dataclass
class B200Backend:
    """`parallelization_backend` selecting this library.

    ode_memory: "host"  -- kick_/drift_ take numpy arrays (copied per call, like a CPU backend);
                "device" -- they take torch CUDA tensors and are stream-ordered (GPU-resident).
    """
    device: int = 0
    ode_memory: str = "host"
    deterministic: bool = True
    interact_variant: int = 0
    ghost_capacity: int = 0  # extra fluid rows for slab ghosts (slabs.py); 0 on a single GPU


def _dtype_id(dt):
    dt = np.dtype(dt)
    if dt == np.float32:
        return _lib.F32
    if dt == np.float64:
        return _lib.F64
    raise ValueError(f"unsupported eltype {dt}")


def _same_boundary_model(a, b) -> bool:
    """Two structure boundary models that the library's one structure slot can serve together."""
    if a is None or b is None:
        return a is b
    if type(a) is not type(b):
        return False
    if isinstance(a, BoundaryModelMonaghanKajtar):
        return (float(a.K) == float(b.K) and float(a.beta) == float(b.beta)
                and float(a.boundary_particle_spacing) == float(b.boundary_particle_spacing))
    sa, sb = a.state_equation, b.state_equation
    return (a.smoothing_kernel.kernel_id == b.smoothing_kernel.kernel_id
            and float(a.smoothing_length) == float(b.smoothing_length)
            and bool(a.clip_negative_pressure) == bool(b.clip_negative_pressure)
            and type(a.density_calculator) is type(b.density_calculator)
            and float(a.density_calculator.pressure_offset) == float(b.density_calculator.pressure_offset)
            and getattr(a.density_calculator, "factor", 0.0) == getattr(b.density_calculator, "factor", 0.0)
            and a.viscosity is None and b.viscosity is None
            and all(float(getattr(sa, f)) == float(getattr(sb, f))
                    for f in ("sound_speed", "exponent", "reference_density", "background_pressure")))


class _MergedStructure:
    """Several `TotalLagrangianSPHSystem`s (examples/fsi/falling_spheres_2d.jl) in the library's one structure slot.
    Exact, because TLSPH pairs its particles once, in the initial configuration (system.jl:391-401): bodies that start
    further apart than the kernel support never see each other.  Same kernel, smoothing length, acceleration, penalty
    force and boundary model; Young's modulus and Poisson ratio go to the library per particle
    (tpb_set_structure_material).  Library order: the integrated particles of all parts, then the clamped ones --
    the ODE rows [part 1 | part 2 | ...] are the library's rows."""

    def __init__(self, parts):
        from scipy.spatial import cKDTree
        p0 = parts[0]
        for q in parts[1:]:
            if not (q.smoothing_kernel.kernel_id == p0.smoothing_kernel.kernel_id
                    and float(q.smoothing_length) == float(p0.smoothing_length)
                    and np.array_equal(q.acceleration, p0.acceleration)
                    and (q.penalty_force is None) == (p0.penalty_force is None)
                    and (q.penalty_force is None or q.penalty_force.alpha == p0.penalty_force.alpha)
                    and _same_boundary_model(q.boundary_model, p0.boundary_model)
                    and q.prescribed_motion is None and p0.prescribed_motion is None):
                raise ValueError("several structure systems share the library's structure slot only with the same "
                                 "kernel, smoothing length, acceleration, penalty force and boundary model, and without "
                                 "a PrescribedMotion")
        support = float(compact_support(p0.smoothing_kernel, p0.smoothing_length))
        for i, a in enumerate(parts):
            tree = cKDTree(np.asarray(a.initial_coordinates, dtype=np.float64))
            for b in parts[i + 1:]:
                d, _ = tree.query(np.asarray(b.initial_coordinates, dtype=np.float64))
                if d.min() <= support * (1 + 1e-6):
                    raise ValueError("structure systems that start within each other's kernel support cannot share "
                                     "the structure slot")
        self.parts = tuple(parts)
        n_int = [q.n_integrated_particles for q in parts]
        n_cl = [q.nparticles - q.n_integrated_particles for q in parts]
        io = np.concatenate([[0], np.cumsum(n_int)])
        co = io[-1] + np.concatenate([[0], np.cumsum(n_cl)])
        # library index of every particle of part k, in the part's own order (integrated first, then clamped)
        self.index_of = [np.concatenate([np.arange(io[k], io[k + 1]), np.arange(co[k], co[k + 1])]).astype(np.int64)
                         for k in range(len(parts))]
        self.nparticles = int(co[-1])
        self.n_integrated_particles = int(io[-1])
        gather = lambda get: np.concatenate([get(q)[:n_int[k]] for k, q in enumerate(parts)]
                                            + [get(q)[n_int[k]:] for k, q in enumerate(parts)])
        self.initial_coordinates = gather(lambda q: np.asarray(q.initial_coordinates))
        self.mass = gather(lambda q: np.asarray(q.mass))
        self.material_density = gather(lambda q: np.asarray(q.material_density))
        full = lambda q, v: np.broadcast_to(np.asarray(v, dtype=np.float64), (q.nparticles,))
        self.young_modulus = gather(lambda q: full(q, q.young_modulus))
        self.poisson_ratio = gather(lambda q: full(q, q.poisson_ratio))
        self.hydrodynamic_mass = (None if p0.boundary_model is None
                                  else gather(lambda q: np.asarray(q.boundary_model.hydrodynamic_mass)))


class Semidiscretization:
    def __init__(self, *systems, neighborhood_search: Optional[GridNeighborhoodSearch] = None,
                 parallelization_backend: Optional[B200Backend] = None, interaction_matrix=None):
        if not systems:
            raise ValueError("at least one system is required")
        fluids = [s for s in systems if isinstance(s, WeaklyCompressibleSPHSystem)]
        walls = [s for s in systems if isinstance(s, WallBoundarySystem)]
        structures = [s for s in systems if isinstance(s, TotalLagrangianSPHSystem)]
        if len(fluids) + len(walls) + len(structures) != len(systems):
            raise ValueError("only WeaklyCompressibleSPHSystem, WallBoundarySystem and TotalLagrangianSPHSystem "
                             "are on the accelerated path")
        if len(fluids) > 1 or (not fluids and not structures):
            raise ValueError("the accelerated path takes one fluid system (none for a structure-only set-up), any "
                             "number of wall systems with the same boundary model and structure systems that can "
                             "share one slot")
        self._merged = _MergedStructure(structures) if len(structures) > 1 else None
        if self._merged is not None:
            ks = [i for i, s_ in enumerate(systems) if isinstance(s_, TotalLagrangianSPHSystem)]
            between = [s_ for s_ in systems[ks[0]:ks[-1] + 1] if not isinstance(s_, TotalLagrangianSPHSystem)]
            if any(s_.n_integrated_particles > 0 for s_ in between):
                raise ValueError("structure systems that share the slot must follow each other in the ODE vectors")
        nd = {s.ndims for s in systems}
        if len(nd) != 1:
            raise ValueError("all systems must have the same number of dimensions")
        # semidiscretization.jl:303-317
        if len({np.dtype(s.eltype) for s in systems}) != 1:
            raise ValueError("all systems must have the same eltype")
        if len({np.dtype(s.coordinates_eltype) for s in systems}) != 1:
            raise ValueError("all systems must have the same eltype for their coordinates")
        self.systems = tuple(systems)
        self.ndims = nd.pop()
        self.eltype = np.dtype(systems[0].eltype)
        self.coordinates_eltype = np.dtype(systems[0].coordinates_eltype)
        if neighborhood_search is not None and neighborhood_search.ndims != self.ndims:
            raise ValueError("neighborhood search dimensionality mismatch")
        self.neighborhood_search = neighborhood_search
        self.parallelization_backend = parallelization_backend or B200Backend()
        n = len(systems)
        if interaction_matrix is None:
            interaction_matrix = np.ones((n, n), dtype=bool)
        self.interaction_matrix = np.asarray(interaction_matrix, dtype=bool)
        if self.interaction_matrix.shape != (n, n):
            raise ValueError(f"`interaction_matrix` must be of size ({n}, {n})")
        # ranges_u / ranges_v (0-based half-open), semidiscretization.jl:128-135
        sizes_u = [s.u_nvariables * s.n_integrated_particles for s in systems]
        sizes_v = [s.v_nvariables * s.n_integrated_particles for s in systems]
        self.ranges_u = tuple((sum(sizes_u[:i]), sum(sizes_u[:i + 1])) for i in range(n))
        self.ranges_v = tuple((sum(sizes_v[:i]), sum(sizes_v[:i + 1])) for i in range(n))
        self._handle = None
        # the system whose clamped particles follow a PrescribedMotion (a moving wall or a structure), if any
        movers = [s for s in systems if getattr(s, "prescribed_motion", None) is not None]
        if len(movers) > 1:
            raise ValueError("at most one system with a PrescribedMotion is on the accelerated path")
        self._motion_system = movers[0] if movers else None
        # inside the library a moving wall and a Monaghan-Kajtar wall occupy the (single) structure slot
        slot = [s for s in systems if (isinstance(s, TotalLagrangianSPHSystem) and s is structures[0])
                or (isinstance(s, WallBoundarySystem)
                    and (s.prescribed_motion is not None or isinstance(s.boundary_model, BoundaryModelMonaghanKajtar)))]
        if self._merged is not None and len(slot) > 1:
            raise ValueError("several structure systems and a moving / Monaghan-Kajtar wall in one semidiscretization "
                             "are outside the accelerated path")
        # A moving wall NEXT TO a structure (examples/fsi/dam_break_gate_2d.jl: the gate and the plate): when both
        # carry dummy particles with the same boundary model, the gate's particles join the structure's slot as
        # additional clamped particles.  Exact as long as no gate particle is within the structure's own kernel
        # support of a structure particle in the initial configuration -- TLSPH pairs its particles once, there
        # (total_lagrangian_sph/system.jl:391-401) -- which is checked here.
        self._gate = None
        if len(slot) == 2 and self._mergeable(slot):
            self._gate = next(s for s in slot if isinstance(s, WallBoundarySystem))
        elif len(slot) > 1:
            raise ValueError("a structure system, a moving wall and a Monaghan-Kajtar wall share one slot on the "
                             "accelerated path: one of them per semidiscretization, or a moving dummy-particle wall "
                             "together with a structure that carries the same BoundaryModelDummyParticles")
        # library system numbers (call order of tpb_add_*_system): the merged gate has the structure's
        self._lib_index, k = {}, 0
        for s_ in systems:
            if s_ is self._gate:
                continue
            if isinstance(s_, TotalLagrangianSPHSystem) and s_ is not structures[0]:
                self._lib_index[id(s_)] = self._lib_index[id(structures[0])]   # shares the first structure's slot
                continue
            self._lib_index[id(s_)] = k
            k += 1
        if self._gate is not None:
            self._lib_index[id(self._gate)] = self._lib_index[id(self.structure)]

    def _mergeable(self, slot) -> bool:
        gate = next((s for s in slot if isinstance(s, WallBoundarySystem) and s.prescribed_motion is not None), None)
        st = next((s for s in slot if isinstance(s, TotalLagrangianSPHSystem)), None)
        if gate is None or st is None or st.prescribed_motion is not None:
            return False
        a, b = gate.boundary_model, st.boundary_model
        if not (isinstance(a, BoundaryModelDummyParticles) and isinstance(b, BoundaryModelDummyParticles)):
            return False
        sa, sb = a.state_equation, b.state_equation
        same = (a.smoothing_kernel.kernel_id == b.smoothing_kernel.kernel_id
                and float(a.smoothing_length) == float(b.smoothing_length)
                and bool(a.clip_negative_pressure) == bool(b.clip_negative_pressure)
                and type(a.density_calculator) is type(b.density_calculator)
                and float(a.density_calculator.pressure_offset) == float(b.density_calculator.pressure_offset)
                and getattr(a.density_calculator, "factor", 0.0) == 0.0 == getattr(b.density_calculator, "factor", 0.0)
                and a.viscosity is None and b.viscosity is None
                and all(float(getattr(sa, f)) == float(getattr(sb, f))
                        for f in ("sound_speed", "exponent", "reference_density", "background_pressure")))
        if not same:
            return False
        from scipy.spatial import cKDTree
        support = float(compact_support(st.smoothing_kernel, st.smoothing_length))
        d, _ = cKDTree(np.asarray(st.initial_coordinates, dtype=np.float64)).query(
            np.asarray(gate.initial_condition.coordinates, dtype=np.float64))
        if d.min() <= support * (1 + 1e-6):
            raise ValueError("the moving wall starts within the structure's kernel support of the structure: the two "
                             "cannot share the structure slot (see DESIGN.md)")
        return True

    def lib_index(self, system) -> int:
        """The system's number inside the library (differs from `system_index` behind a merged moving wall)."""
        return self._lib_index[id(system)]

    # -- helpers ---------------------------------------------------------------------------
    @property
    def fluid(self) -> Optional[WeaklyCompressibleSPHSystem]:
        return next((s for s in self.systems if isinstance(s, WeaklyCompressibleSPHSystem)), None)

    def _placeholder_fluid(self) -> WeaklyCompressibleSPHSystem:
        """`Semidiscretization(structure_system)` (examples/structure/oscillating_beam_2d.jl): the library's grid
        belongs to a fluid system, so a structure-only set-up registers an EMPTY one behind the user's systems -- no
        particles, no rows in the ODE vectors; its kernel and smoothing length (the structure's) only size the cells."""
        from .model import StateEquationCole
        from .setups import InitialCondition
        st, nd, t = self.structure, self.ndims, self.eltype
        ic = InitialCondition(np.zeros((0, nd), dtype=self.coordinates_eltype), np.zeros((0, nd), dtype=t),
                              np.zeros(0, dtype=t), np.zeros(0, dtype=t), np.zeros(0, dtype=t),
                              st.initial_condition.particle_spacing)
        return WeaklyCompressibleSPHSystem(ic, smoothing_kernel=st.smoothing_kernel, smoothing_length=st.smoothing_length,
                                           density_calculator=ContinuityDensity(),
                                           state_equation=StateEquationCole(sound_speed=1.0, reference_density=1.0,
                                                                            exponent=1))

    @property
    def wall(self) -> Optional[WallBoundarySystem]:
        return next((s for s in self.systems if isinstance(s, WallBoundarySystem)), None)

    @property
    def structure(self) -> Optional[TotalLagrangianSPHSystem]:
        return next((s for s in self.systems if isinstance(s, TotalLagrangianSPHSystem)), None)

    def system_index(self, system) -> int:
        return next(i for i, s in enumerate(self.systems) if s is system)

    def wrap_u(self, u_ode, system):
        a, b = self.ranges_u[self.system_index(system)]
        return u_ode[a:b].reshape(system.n_integrated_particles, system.u_nvariables)

    def wrap_v(self, v_ode, system):
        a, b = self.ranges_v[self.system_index(system)]
        return v_ode[a:b].reshape(system.n_integrated_particles, system.v_nvariables)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def close(self):
        if self._handle is not None:
            _lib.load().tpb_destroy(self._handle)
            self._handle = None

    # -- library plumbing ------------------------------------------------------------------
    def _fluid_params(self, f: WeaklyCompressibleSPHSystem) -> _lib.FluidParams:
        se = f.state_equation
        p = _lib.FluidParams()
        p.struct_size = C.sizeof(_lib.FluidParams)
        p.kernel = f.smoothing_kernel.kernel_id
        p.density_calculator = f.density_calculator.density_id
        p.clip_negative_pressure = int(se.clip_negative_pressure)
        p.has_viscosity = 0 if f.viscosity is None else int(getattr(f.viscosity, "viscosity_id", 1))
        p.has_diffusion = int(f.density_diffusion is not None)
        t = self.eltype.type
        p.smoothing_length = float(t(f.smoothing_length))
        p.sound_speed = float(t(se.sound_speed))
        p.exponent = float(t(se.exponent))
        p.reference_density = float(t(se.reference_density))
        p.background_pressure = float(t(se.background_pressure))
        if f.viscosity is not None:
            if p.has_viscosity == 1:
                p.alpha, p.beta, p.epsilon = (float(t(f.viscosity.alpha)), float(t(f.viscosity.beta)),
                                              float(t(f.viscosity.epsilon)))
            else:  # ViscosityMorris / ViscosityAdami: the kinematic viscosity travels in `alpha`
                p.alpha, p.beta, p.epsilon = float(t(f.viscosity.nu)), 0.0, float(t(f.viscosity.epsilon))
        if f.density_diffusion is not None:
            p.delta = float(t(f.density_diffusion.delta))
        for d in range(self.ndims):
            p.acceleration[d] = float(f.acceleration[d])
        if isinstance(se, StateEquationAdaptiveCole):
            p.adaptive_sound_speed = 1
            p.adaptive_params_f32 = int(se.param_eltype == np.float32)
            if not (se.param_eltype == np.float32 or se.param_eltype == self.eltype):
                raise ValueError("StateEquationAdaptiveCole: parameters must be Float32 or eltype(system)")
            p.mach_number_target = float(se.mach_number_target)
            p.min_sound_speed = float(se.min_sound_speed)
            p.max_sound_speed = float(se.max_sound_speed)
        if isinstance(f.source_terms, SourceTermDamping):
            p.damping_coefficient = float(t(f.source_terms.damping_coefficient))
        return p

    def _wall_params(self, w: WallBoundarySystem) -> _lib.WallParams:
        m = w.boundary_model
        se = m.state_equation
        t = self.eltype.type
        p = _lib.WallParams()
        p.struct_size = C.sizeof(_lib.WallParams)
        p.kernel = m.smoothing_kernel.kernel_id
        p.clip_negative_pressure = int(m.clip_negative_pressure)
        p.smoothing_length = float(t(m.smoothing_length))
        p.sound_speed = float(t(se.sound_speed))
        p.exponent = float(t(se.exponent))
        p.reference_density = float(t(se.reference_density))
        p.background_pressure = float(t(se.background_pressure))
        if isinstance(m.density_calculator, ContinuityDensity):
            p.density_calculator = _lib.WALL_DENSITY_CONTINUITY
            p.eos_clip_negative_pressure = int(getattr(se, "clip_negative_pressure", False))
        else:
            p.pressure_offset = float(t(m.density_calculator.pressure_offset))
        if m.viscosity is not None:   # no-slip wall
            p.has_viscosity = int(getattr(m.viscosity, "viscosity_id", 1))
            if p.has_viscosity == 1:
                p.alpha, p.beta = float(t(m.viscosity.alpha)), float(t(m.viscosity.beta))
            else:
                p.alpha = float(t(m.viscosity.nu))
            p.epsilon = float(t(m.viscosity.epsilon))
            if p.has_viscosity >= 2 and self.fluid.viscosity is None:
                raise ValueError("a ViscosityMorris / ViscosityAdami wall needs a fluid viscosity model "
                                 "(kinematic_viscosity(fluid, nothing, ...) has no method in the reference)")
        if isinstance(se, StateEquationAdaptiveCole):
            if se is not self.fluid.state_equation:
                raise ValueError("a boundary model with its own StateEquationAdaptiveCole is outside the "
                                 "accelerated hot path (share the fluid's state equation object)")
            p.sound_speed_from_fluid = 1
        return p

    def _structure_params(self, st: TotalLagrangianSPHSystem) -> _lib.StructureParams:
        t = self.eltype.type
        p = _lib.StructureParams()
        p.struct_size = C.sizeof(_lib.StructureParams)
        p.kernel = st.smoothing_kernel.kernel_id
        p.smoothing_length = float(t(st.smoothing_length))
        # (vectors: the library gets them per particle after tpb_semidiscretize; the scalars only have to be valid)
        p.young_modulus = float(t(np.ravel(st.young_modulus)[0]))
        p.poisson_ratio = float(t(np.ravel(st.poisson_ratio)[0]))
        if st.penalty_force is not None:
            p.has_penalty_force = 1
            p.penalty_alpha = float(t(st.penalty_force.alpha))
        for d in range(self.ndims):
            p.acceleration[d] = float(st.acceleration[d])
        m = st.boundary_model
        if isinstance(m, BoundaryModelDummyParticles):
            se = m.state_equation
            p.boundary_model = _lib.BOUNDARY_DUMMY_PARTICLES
            p.bm_kernel = m.smoothing_kernel.kernel_id
            p.bm_clip_negative_pressure = int(m.clip_negative_pressure)
            p.bm_smoothing_length = float(t(m.smoothing_length))
            p.bm_sound_speed, p.bm_exponent = float(t(se.sound_speed)), float(t(se.exponent))
            p.bm_reference_density = float(t(se.reference_density))
            p.bm_background_pressure = float(t(se.background_pressure))
            p.bm_pressure_offset = float(t(m.density_calculator.pressure_offset))
            # BernoulliPressureExtrapolation on a structure: the dynamic term always applies (dummy_particles.jl:696-707)
            p.bm_bernoulli_factor = float(t(getattr(m.density_calculator, "factor", 0.0)))
        elif m is not None:
            p.boundary_model = _lib.BOUNDARY_MONAGHAN_KAJTAR
            p.mk_K, p.mk_beta, p.mk_spacing = float(t(m.K)), float(t(m.beta)), float(t(m.boundary_particle_spacing))
        return p

    def _create(self, u0_ode: np.ndarray):
        L = _lib.load()
        be = self.parallelization_backend
        cfg = _lib.Config()
        cfg.struct_size = C.sizeof(_lib.Config)
        cfg.ndims = self.ndims
        cfg.eltype = _dtype_id(self.eltype)
        cfg.coords_eltype = _dtype_id(self.coordinates_eltype)
        cfg.device = be.device
        cfg.ode_memory = _lib.MEM_DEVICE if be.ode_memory == "device" else _lib.MEM_HOST
        cfg.deterministic = int(be.deterministic)
        cfg.interact_variant = int(be.interact_variant)
        nhs = self.neighborhood_search
        if nhs is not None and nhs.cell_list is not None:
            cfg.has_bounds = 1
            cfg.max_points_per_cell = nhs.cell_list.max_points_per_cell
            for d in range(self.ndims):
                cfg.min_corner[d] = float(nhs.cell_list.min_corner[d])
                cfg.max_corner[d] = float(nhs.cell_list.max_corner[d])
        elif be.ghost_capacity:
            # without a box the library derives one from u0, which only holds the owned rows
            raise ValueError("B200Backend.ghost_capacity needs a FullGridCellList bounding box")
        h = C.c_void_p()
        _lib.check(None, L.tpb_create(C.byref(cfg), C.byref(h)))
        self._handle = h
        try:
            for s in self.systems:
                idx = C.c_int32(-1)
                if s is self._gate or (isinstance(s, TotalLagrangianSPHSystem) and s is not self.structure):
                    continue    # registered with the (first) structure
                if isinstance(s, WeaklyCompressibleSPHSystem):
                    mass = np.ascontiguousarray(s.mass, dtype=self.eltype)
                    if be.ghost_capacity:
                        mass = np.concatenate([mass, np.zeros(be.ghost_capacity, dtype=self.eltype)])
                    fp = self._fluid_params(s)
                    _lib.check(h, L.tpb_add_fluid_system(h, C.byref(fp), mass.size,
                                                         mass.ctypes.data, C.byref(idx)))
                elif isinstance(s, TotalLagrangianSPHSystem):
                    if be.ghost_capacity:
                        raise ValueError("slab ghosts are not combined with a structure system")
                    src = self._merged if self._merged is not None else s
                    x0 = np.ascontiguousarray(src.initial_coordinates, dtype=self.coordinates_eltype)
                    mass = np.ascontiguousarray(src.mass, dtype=self.eltype)
                    rho = np.ascontiguousarray(src.material_density, dtype=self.eltype)
                    hyd_src = (self._merged.hydrodynamic_mass if self._merged is not None
                               else (s.boundary_model.hydrodynamic_mass if s.boundary_model is not None else None))
                    hyd = np.ascontiguousarray(hyd_src, dtype=self.eltype) if hyd_src is not None else None
                    sp = self._structure_params(s)
                    n_total = src.nparticles
                    if self._gate is not None:
                        # the gate's dummy particles behind the structure's clamped ones; their "material" mass and
                        # density never enter (no structure particle has them as a neighbour)
                        g = self._gate
                        gm = np.ascontiguousarray(g.boundary_model.hydrodynamic_mass, dtype=self.eltype)
                        x0 = np.ascontiguousarray(np.concatenate(
                            [x0, np.asarray(g.initial_condition.coordinates, dtype=self.coordinates_eltype)]))
                        mass, hyd = np.concatenate([mass, gm]), np.concatenate([hyd, gm])
                        rho = np.concatenate([rho, np.asarray(g.boundary_model.initial_density, dtype=self.eltype)])
                        sp.bm_wall_semantics = 1   # the acceleration of the gate's particles enters their Adami sum;
                        n_total += g.nparticles    # the structure's own clamped particles have none
                    _lib.check(h, L.tpb_add_structure_system(h, C.byref(sp), n_total, src.n_integrated_particles,
                                                             x0.ctypes.data, mass.ctypes.data, rho.ctypes.data,
                                                             hyd.ctypes.data if hyd is not None else None,
                                                             C.byref(idx)))
                elif s.prescribed_motion is not None:
                    # a moving wall (wall_boundary/system.jl:22-60 with prescribed_motion): dummy particles whose
                    # positions / velocities / accelerations are prescribed -- inside the library a structure
                    # without integrated particles carrying the wall's boundary model (tpb200.h, bm_wall_semantics)
                    if be.ghost_capacity:
                        raise ValueError("slab ghosts are not combined with a moving wall")
                    m, t = s.boundary_model, self.eltype.type
                    se = m.state_equation
                    sp = _lib.StructureParams()
                    sp.struct_size = C.sizeof(_lib.StructureParams)
                    sp.kernel = self.fluid.smoothing_kernel.kernel_id
                    sp.smoothing_length = float(t(self.fluid.smoothing_length))
                    sp.young_modulus, sp.poisson_ratio = 1.0, 0.0
                    sp.boundary_model = _lib.BOUNDARY_DUMMY_PARTICLES
                    sp.bm_kernel = m.smoothing_kernel.kernel_id
                    sp.bm_clip_negative_pressure = int(m.clip_negative_pressure)
                    sp.bm_smoothing_length = float(t(m.smoothing_length))
                    sp.bm_sound_speed, sp.bm_exponent = float(t(se.sound_speed)), float(t(se.exponent))
                    sp.bm_reference_density = float(t(se.reference_density))
                    sp.bm_background_pressure = float(t(se.background_pressure))
                    sp.bm_pressure_offset = float(t(m.density_calculator.pressure_offset))
                    sp.bm_wall_semantics = 1
                    sp.bm_bernoulli_factor = float(t(getattr(m.density_calculator, "factor", 0.0)))
                    x0 = np.ascontiguousarray(s.coordinates, dtype=self.coordinates_eltype)
                    hyd = np.ascontiguousarray(m.hydrodynamic_mass, dtype=self.eltype)
                    rho = np.ascontiguousarray(m.initial_density, dtype=self.eltype)
                    _lib.check(h, L.tpb_add_structure_system(h, C.byref(sp), s.nparticles, 0, x0.ctypes.data,
                                                             hyd.ctypes.data, rho.ctypes.data, hyd.ctypes.data,
                                                             C.byref(idx)))
                elif isinstance(s.boundary_model, BoundaryModelMonaghanKajtar):
                    # a wall of repulsive particles (test/examples/gpu.jl:219-253): inside the library the
                    # Monaghan-Kajtar coupling belongs to the structure path, and a static wall is a structure
                    # without integrated particles (all clamped) -- no entries in the ODE vectors, like a wall
                    if self.structure is not None:
                        raise ValueError("a Monaghan-Kajtar wall and a structure system in one semidiscretization "
                                         "are outside the accelerated path")
                    m = s.boundary_model
                    t = self.eltype.type
                    sp = _lib.StructureParams()
                    sp.struct_size = C.sizeof(_lib.StructureParams)
                    sp.kernel = self.fluid.smoothing_kernel.kernel_id
                    sp.smoothing_length = float(t(self.fluid.smoothing_length))
                    sp.young_modulus, sp.poisson_ratio = 1.0, 0.0
                    sp.boundary_model = _lib.BOUNDARY_MONAGHAN_KAJTAR
                    sp.mk_K, sp.mk_beta, sp.mk_spacing = float(t(m.K)), float(t(m.beta)), float(t(m.boundary_particle_spacing))
                    x0 = np.ascontiguousarray(s.coordinates, dtype=self.coordinates_eltype)
                    hyd = np.ascontiguousarray(m.hydrodynamic_mass, dtype=self.eltype)
                    rho = np.ones(s.nparticles, dtype=self.eltype)
                    _lib.check(h, L.tpb_add_structure_system(h, C.byref(sp), s.nparticles, 0, x0.ctypes.data,
                                                             hyd.ctypes.data, rho.ctypes.data, hyd.ctypes.data,
                                                             C.byref(idx)))
                else:
                    coords = np.ascontiguousarray(s.coordinates, dtype=self.coordinates_eltype)
                    mass = np.ascontiguousarray(s.boundary_model.hydrodynamic_mass, dtype=self.eltype)
                    dens = np.ascontiguousarray(s.boundary_model.initial_density, dtype=self.eltype)
                    wp = self._wall_params(s)
                    _lib.check(h, L.tpb_add_wall_system(h, C.byref(wp), s.nparticles,
                                                        coords.ctypes.data, mass.ctypes.data,
                                                        dens.ctypes.data, C.byref(idx)))
                assert idx.value == self.lib_index(s)
            if self.fluid is None:
                ph, idx = self._placeholder_fluid(), C.c_int32(-1)
                fp = self._fluid_params(ph)
                _lib.check(h, L.tpb_add_fluid_system(h, C.byref(fp), 0, None, C.byref(idx)))
            n = len(self.systems)
            if self._gate is not None:
                gi, si = self.system_index(self._gate), self.system_index(self.structure)
                others = [k for k in range(n) if k not in (gi, si)]
                im = self.interaction_matrix
                if not (np.array_equal(im[gi, others], im[si, others]) and np.array_equal(im[others, gi], im[others, si])):
                    raise ValueError("a moving wall that shares the structure's slot needs the structure's "
                                     "interaction switches")
            for i in range(n):
                for j in range(n):
                    li, lj = self.lib_index(self.systems[i]), self.lib_index(self.systems[j])
                    if not self.interaction_matrix[i, j] and not (self._gate is not None and li == lj and i != j):
                        _lib.check(h, L.tpb_set_interaction(h, li, lj, 0))
            _lib.check(h, L.tpb_semidiscretize(h, u0_ode.ctypes.data))
            st_ = self._merged if self._merged is not None else self.structure
            if st_ is not None and np.ndim(st_.young_modulus) > 0:
                n_lib = st_.nparticles + (self._gate.nparticles if self._gate is not None else 0)
                pad = lambda a: np.ascontiguousarray(np.concatenate([np.asarray(a, dtype=np.float64),
                                                                     np.full(n_lib - st_.nparticles, np.ravel(a)[0])]),
                                                     dtype=self.eltype)
                E_, nu_ = pad(st_.young_modulus), pad(st_.poisson_ratio)
                _lib.check(h, L.tpb_set_structure_material(h, E_.ctypes.data, nu_.ctypes.data))
            if be.ghost_capacity:
                n_own = self.fluid.nparticles
                _lib.check(h, L.tpb_set_fluid_count(h, n_own, n_own))
            nu, nv = C.c_int64(), C.c_int64()
            _lib.check(h, L.tpb_ode_sizes(h, C.byref(nu), C.byref(nv)))
            assert nu.value == self.ranges_u[-1][1] and nv.value == self.ranges_v[-1][1]
        except Exception:
            self.close()
            raise

    def _ptr(self, arr, n_expected, dtype, name):
        """Raw pointer of an ODE vector, after checking type / size / residency."""
        be = self.parallelization_backend
        if be.ode_memory == "device":
            import torch
            if not (isinstance(arr, torch.Tensor) and arr.is_cuda):
                raise TypeError(f"{name}: B200Backend(ode_memory='device') expects torch CUDA tensors")
            if arr.device.index != be.device:
                raise ValueError(f"{name} lives on cuda:{arr.device.index}, backend on cuda:{be.device}")
            want = torch.float32 if np.dtype(dtype) == np.float32 else torch.float64
            if arr.dtype != want or not arr.is_contiguous() or arr.numel() != n_expected:
                raise ValueError(f"{name}: expected contiguous {want} tensor of length {n_expected}")
            return arr.data_ptr()
        if not isinstance(arr, np.ndarray):
            raise TypeError(f"{name}: B200Backend(ode_memory='host') expects numpy arrays")
        if arr.dtype != np.dtype(dtype) or not arr.flags.c_contiguous or arr.size != n_expected:
            raise ValueError(f"{name}: expected contiguous {np.dtype(dtype)} array of length {n_expected}")
        return arr.ctypes.data

    def _bind_stream(self):
        if self.parallelization_backend.ode_memory == "device":
            import torch
            stream = torch.cuda.current_stream(self.parallelization_backend.device).cuda_stream
            # torch's default stream has the handle 0, which tpb_set_stream reads as "use the
            # handle's own stream": pass cudaStreamLegacy (0x1) for it instead
            _lib.check(self._handle, _lib.load().tpb_set_stream(self._handle, C.c_void_p(stream if stream else 1)))

    def synchronize(self):
        # the handle may still be bound to another stream (e.g. the side stream of a CUDA-graph
        # capture): wait on torch's current stream, where the caller's work was queued
        self._bind_stream()
        _lib.check(self._handle, _lib.load().tpb_synchronize(self._handle))

    def sound_speed(self) -> float:
        """`system_sound_speed(fluid)` as of the last kick (StateEquationAdaptiveCole), else the constant."""
        out = C.c_double(0.0)
        _lib.check(self._handle, _lib.load().tpb_get_sound_speed(self._handle, C.byref(out)))
        return out.value

    def apply_prescribed_motion(self, t):
        """update_positions! -> apply_prescribed_motion! (wall_boundary/system.jl:192-205, total_lagrangian_sph/
        system.jl:403-447) for the system that has a PrescribedMotion: the movement function runs on the host, the
        library gets the clamped particles' positions, velocities and accelerations for the next kick."""
        s = self._motion_system
        if s is None:
            return
        L = _lib.load()
        if s.apply_prescribed_motion(t):
            xs, vs, as_ = s.clamped_coordinates, s.clamped_velocity, s.clamped_acceleration
            if s is self._gate:
                # the structure's own clamped particles come first in the slot: fixed
                st = self.structure
                fixed = np.asarray(st.initial_coordinates[st.n_integrated_particles:], dtype=np.float64)
                xs = np.concatenate([fixed, xs])
                vs = np.concatenate([np.zeros_like(fixed), vs])
                as_ = np.concatenate([np.zeros_like(fixed), as_])
            x = np.ascontiguousarray(xs, dtype=self.coordinates_eltype)
            v = np.ascontiguousarray(vs, dtype=self.eltype)
            a = np.ascontiguousarray(as_, dtype=self.eltype)
            _lib.check(self._handle, L.tpb_set_clamped_motion(self._handle, x.ctypes.data, v.ctypes.data,
                                                              a.ctypes.data, 1))
            if isinstance(s, WallBoundarySystem):
                # current_coordinates(u, wall) = system.coordinates
                s.coordinates = np.ascontiguousarray(s.clamped_coordinates, dtype=self.coordinates_eltype)
        else:
            _lib.check(self._handle, L.tpb_set_clamped_motion(self._handle, None, None, None, 0))

    def reinit_density(self, v_ode, u_ode):
        """`reinit_density!` (DensityReinitializationCallback; wcsph/system.jl:398-415): the fluid's density rows of
        `v_ode` become the Shepard-corrected summation density, in place (tpb_reinit_density)."""
        nu, nv = self.ranges_u[-1][1], self.ranges_v[-1][1]
        pv = self._ptr(v_ode, nv, self.eltype, "v_ode")
        pu = self._ptr(u_ode, nu, self.coordinates_eltype, "u_ode")
        self._bind_stream()
        _lib.check(self._handle, _lib.load().tpb_reinit_density(self._handle, pv, pu))

    def sort_particles(self, v_ode, u_ode):
        """`sort_particles!` of every fluid system (callbacks/sorting.jl:123-157): the rows of (v_ode, u_ode) in
        place in grid-cell order (tpb_sort_system); later kicks give bit-identical results, row for row."""
        nu, nv = self.ranges_u[-1][1], self.ranges_v[-1][1]
        pv = self._ptr(v_ode, nv, self.eltype, "v_ode")
        pu = self._ptr(u_ode, nu, self.coordinates_eltype, "u_ode")
        self._bind_stream()
        for s in self.systems:
            if isinstance(s, WeaklyCompressibleSPHSystem):
                _lib.check(self._handle, _lib.load().tpb_sort_system(self._handle, self.lib_index(s), pv, pu))

    def set_integrate_structure(self, enabled: bool):
        """`semi.integrate_tlsph[] = enabled` (semidiscretization.jl:149): with a SplitIntegrationCallback kick! /
        drift! leave the structure's rows zero."""
        _lib.check(self._handle, _lib.load().tpb_set_integrate_structure(self._handle, int(bool(enabled))))
        self.integrate_tlsph = bool(enabled)

    def adaptive_sound_speed_on_device(self) -> bool:
        """StateEquationAdaptiveCole without a host round trip (tile sweeps, ContinuityDensity, no structure
        system; `TPB_ADAPTIVE_HOST` selects the host-scalar path): such a kick is stream-ordered and can
        be captured in a CUDA graph."""
        import os
        from .model import ContinuityDensity, TotalLagrangianSPHSystem
        fluids = [s_ for s_ in self.systems if isinstance(s_, WeaklyCompressibleSPHSystem)]
        return ("TPB_ADAPTIVE_HOST" not in os.environ and self.parallelization_backend.interact_variant in (0, 2)
                and all(isinstance(f.density_calculator, ContinuityDensity) for f in fluids)
                and not any(isinstance(s_, TotalLagrangianSPHSystem) for s_ in self.systems))

    def stats(self) -> _lib.Stats:
        st = _lib.Stats()
        _lib.check(self._handle, _lib.load().tpb_get_stats(self._handle, C.byref(st)))
        return st

    def set_profiling(self, max_kicks: int):
        """Arm CUDA-event phase timing for the next `max_kicks` kicks (0 disables)."""
        _lib.check(self._handle, _lib.load().tpb_set_profiling(self._handle, int(max_kicks)))

    def phase_times(self) -> dict:
        """Mean duration in ms of each kick phase since profiling was armed (synchronises)."""
        ms = (C.c_double * len(_lib.PHASES))()
        n = C.c_int32(0)
        _lib.check(self._handle, _lib.load().tpb_get_phase_times(self._handle, ms, C.byref(n)))
        out = {name: ms[i] for i, name in enumerate(_lib.PHASES)}
        out["n_kicks"] = n.value
        return out

    def system_field(self, system, field: str) -> np.ndarray:
        """`system.pressure`, `cache.density`, `boundary_model.pressure/cache.density/cache.volume`
        after the last kick (fluid.jl:312-326, wall_boundary/system.jl:339-350)."""
        fid = {"pressure": _lib.FIELD_PRESSURE, "density": _lib.FIELD_DENSITY,
               "volume": _lib.FIELD_VOLUME, "wall_velocity": _lib.FIELD_WALL_VELOCITY,
               "deformation_grad": _lib.FIELD_DEFORMATION_GRADIENT, "pk1_rho2": _lib.FIELD_PK1_RHO2,
               "correction_matrix": _lib.FIELD_CORRECTION_MATRIX}[field]
        li = self.lib_index(system)
        # a moving wall merged into the structure's slot: the library's system holds [structure | gate]
        n_lib, lo = system.nparticles, 0
        if self._gate is not None and system in (self._gate, self.structure):
            n_lib = self.structure.nparticles + self._gate.nparticles
            lo = self.structure.nparticles if system is self._gate else 0
        pick = None
        if self._merged is not None and isinstance(system, TotalLagrangianSPHSystem):
            n_lib = self._merged.nparticles
            pick = self._merged.index_of[self._merged.parts.index(system)]
        if fid >= _lib.FIELD_DEFORMATION_GRADIENT:
            # structure: (n, ND, ND) with [p, j, i] = M[i, j, p] (the reference's ND x ND x n memory layout)
            out = np.zeros((n_lib, self.ndims, self.ndims), dtype=self.eltype)
            _lib.check(self._handle, _lib.load().tpb_get_system_field(self._handle, li, fid, out.ctypes.data, n_lib))
            return out[pick] if pick is not None else out[lo:lo + system.nparticles]
        if field == "wall_velocity":   # boundary_model.cache.wall_velocity, (n, ND)
            out = np.zeros((n_lib, self.ndims), dtype=self.eltype)
            _lib.check(self._handle, _lib.load().tpb_get_system_field(self._handle, li, fid, out.ctypes.data, n_lib))
            return out[lo:lo + system.nparticles]
        out = np.zeros(n_lib, dtype=self.eltype)
        _lib.check(self._handle, _lib.load().tpb_get_system_field(self._handle, li, fid, out.ctypes.data, out.size))
        return out[pick] if pick is not None else out[lo:lo + system.nparticles]

    def count_neighbor_pairs(self, system, neighbor, u_ode) -> int:
        """Number of ordered neighbour pairs of (system, neighbor) for coordinates `u_ode`."""
        L = _lib.load()
        ptr = self._ptr(u_ode, self.ranges_u[-1][1], self.coordinates_eltype, "u_ode")
        self._bind_stream()
        cnt = C.c_int64(0)
        rc = L.tpb_neighbor_pairs(self._handle, self.lib_index(system),
                                  self.lib_index(neighbor), ptr, 0, None, None, C.byref(cnt))
        if rc not in (0, 6):  # TPB_ERR_CAPACITY is expected: only the count is wanted
            _lib.check(self._handle, rc)
        return int(cnt.value)

    def neighbor_pairs(self, system, neighbor, u_ode):
        """Sorted (i, j) neighbour pairs of the ordered system pair (test hook)."""
        L = _lib.load()
        nu = self.ranges_u[-1][1]
        ptr = self._ptr(u_ode, nu, self.coordinates_eltype, "u_ode")
        self._bind_stream()
        cap = max(1024, 80 * max(system.nparticles, 1))
        while True:
            oi = np.empty(cap, dtype=np.int32)
            oj = np.empty(cap, dtype=np.int32)
            cnt = C.c_int64(0)
            rc = L.tpb_neighbor_pairs(self._handle, self.lib_index(system),
                                      self.lib_index(neighbor), ptr, cap, oi.ctypes.data,
                                      oj.ctypes.data, C.byref(cnt))
            if rc == 6:  # TPB_ERR_CAPACITY
                cap = int(cnt.value)
                continue
            _lib.check(self._handle, rc)
            n = cnt.value
            order = np.lexsort((oj[:n], oi[:n]))
            return oi[:n][order], oj[:n][order]


@dataclass
class DynamicalODEProblem:
    """What `semidiscretize` returns: `DynamicalODEProblem(kick!, drift!, v0, u0, tspan, p)`
    (semidiscretization.jl:391-395)."""
    f1: object
    f2: object
    v0: object
    u0: object
    tspan: tuple
    p: SimpleNamespace


def semidiscretize(semi: Semidiscretization, tspan) -> DynamicalODEProblem:
    """semidiscretization.jl:293-396: builds u0/v0 (`write_u0!`/`write_v0!`, fluid.jl:79-97,
    wcsph/system.jl:440-445), initialises the neighbourhood search and moves everything to the
    device."""
    nu, nv = semi.ranges_u[-1][1], semi.ranges_v[-1][1]
    u0 = np.zeros(nu, dtype=semi.coordinates_eltype)
    v0 = np.zeros(nv, dtype=semi.eltype)
    for s in semi.systems:
        if s.n_integrated_particles == 0:
            continue
        u = semi.wrap_u(u0, s)
        v = semi.wrap_v(v0, s)
        if isinstance(s, WallBoundarySystem):
            # write_v0! of dummy particles with ContinuityDensity (wall_boundary/system.jl:243-252)
            v[:, 0] = s.boundary_model.initial_density
            continue
        if isinstance(s, TotalLagrangianSPHSystem):
            # write_u0! / write_v0! (total_lagrangian_sph/system.jl:588-612): integrated particles only
            n_int = s.n_integrated_particles
            u[:] = s.initial_condition.coordinates[:n_int]
            v[:] = s.initial_condition.velocity[:n_int]
            continue
        u[:] = s.initial_condition.coordinates
        v[:, : s.ndims] = s.initial_condition.velocity
        if not isinstance(s.density_calculator, SummationDensity):
            v[:, s.ndims] = s.initial_condition.density
    semi._create(u0)
    if semi.parallelization_backend.ode_memory == "device":
        import torch
        dev = torch.device("cuda", semi.parallelization_backend.device)
        u0 = torch.from_numpy(u0).to(dev)
        v0 = torch.from_numpy(v0).to(dev)
    return DynamicalODEProblem(kick_, drift_, v0, u0, tuple(tspan), SimpleNamespace(semi=semi))


def kick_(dv_ode, v_ode, u_ode, p, t):
    """`kick!(dv_ode, v_ode, u_ode, p, t)` (semidiscretization.jl:589-612)."""
    semi: Semidiscretization = p.semi
    nu, nv = semi.ranges_u[-1][1], semi.ranges_v[-1][1]
    pdv = semi._ptr(dv_ode, nv, semi.eltype, "dv_ode")
    pv = semi._ptr(v_ode, nv, semi.eltype, "v_ode")
    pu = semi._ptr(u_ode, nu, semi.coordinates_eltype, "u_ode")
    semi._bind_stream()
    semi.apply_prescribed_motion(float(t))
    _lib.check(semi._handle, _lib.load().tpb_kick(semi._handle, pdv, pv, pu, float(t)))
    return dv_ode


def drift_(du_ode, v_ode, u_ode, p, t):
    """`drift!(du_ode, v_ode, u_ode, p, t)` (semidiscretization.jl:522-536)."""
    semi: Semidiscretization = p.semi
    nu, nv = semi.ranges_u[-1][1], semi.ranges_v[-1][1]
    pdu = semi._ptr(du_ode, nu, semi.coordinates_eltype, "du_ode")
    pv = semi._ptr(v_ode, nv, semi.eltype, "v_ode")
    pu = semi._ptr(u_ode, nu, semi.coordinates_eltype, "u_ode")
    semi._bind_stream()
    _lib.check(semi._handle, _lib.load().tpb_drift(semi._handle, pdu, pv, pu, float(t)))
    return du_ode
