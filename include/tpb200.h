/*
 * tpb200.h -- C ABI of libtpb200.so: the B200-native (sm_100a) WCSPH right-hand side that
 * drops in under TrixiParticles.jl's `Semidiscretization` / `semidiscretize` / `kick!` /
 * `drift!` interface.  Host code in any language binds exactly these entry points
 * (Julia: `ccall`; Python: ctypes, see trixiparticles.jl_b200/_lib.py).
 *
 * Conventions
 *   - every function returns an int32 status (TPB_OK == 0); nothing throws across the ABI;
 *     `tpb_last_error` gives the message for the last non-zero status.
 *   - plain pointers and sizes only.  Arrays use the reference's memory layout: Julia
 *     column-major `nvars x nparticles` matrices == particle-major C arrays.
 *   - `eltype` (T) is the system element type, `coords_eltype` (cT) the coordinate type
 *     (/root/reference/src/general/semidiscretization.jl:326-332).  Real parameters cross
 *     the ABI as double; callers compute them in T first, so the conversion back is exact.
 *   - one call at a time per handle from one host thread.  All device work of a handle is
 *     issued on one CUDA stream (`tpb_set_stream`).  With TPB_MEM_DEVICE the calls are
 *     stream-ordered and return without synchronising; with TPB_MEM_HOST they return after
 *     the results are in the caller's buffers.  No allocation happens inside `tpb_kick` /
 *     `tpb_drift` (reference: zero allocations per RHS, test/count_allocations.jl:111-121).
 *
 * Paths cited below are relative to /root/reference.
 */
#ifndef TPB200_H
#define TPB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TPB_VERSION_MAJOR 0
#define TPB_VERSION_MINOR 1

/* status codes */
#define TPB_OK 0
#define TPB_ERR_INVALID_ARGUMENT 1 /* reference: ArgumentError at construction */
#define TPB_ERR_UNSUPPORTED 2      /* option outside the accelerated hot path */
#define TPB_ERR_CUDA 3             /* CUDA runtime failure (message has the cudaError) */
#define TPB_ERR_OUT_OF_BOUNDS 4    /* a particle left the FullGridCellList bounding box */
#define TPB_ERR_STATE 5            /* call order violated (e.g. kick before semidiscretize) */
#define TPB_ERR_CAPACITY 6         /* output buffer too small */

/* element types */
#define TPB_F32 0
#define TPB_F64 1

/* where the ODE vectors handed to kick/drift live */
#define TPB_MEM_HOST 0
#define TPB_MEM_DEVICE 1

/* smoothing kernels (src/general/smoothing_kernels.jl:191-227, :434-455) */
#define TPB_VISCOSITY_NONE 0
#define TPB_VISCOSITY_MONAGHAN 1
#define TPB_VISCOSITY_MORRIS 2 /* viscosity.jl:134-205 */
#define TPB_VISCOSITY_ADAMI 3  /* viscosity.jl:207-285 */
#define TPB_KERNEL_WENDLAND_C2 0
#define TPB_KERNEL_SCHOENBERG_CUBIC 1
#define TPB_KERNEL_WENDLAND_C4 2 /* smoothing_kernels.jl:489-514 */
#define TPB_KERNEL_WENDLAND_C6 3 /* smoothing_kernels.jl:548-574 */
#define TPB_KERNEL_SCHOENBERG_QUARTIC 4 /* smoothing_kernels.jl:264-322, compact support 5/2 h */
#define TPB_KERNEL_SCHOENBERG_QUINTIC 5 /* smoothing_kernels.jl:357-395, compact support 3 h */

/* density calculators (src/general/density_calculators.jl) */
#define TPB_DENSITY_CONTINUITY 0
#define TPB_DENSITY_SUMMATION 1

/* fields for tpb_get_system_field (fluid.jl:312-326, wall_boundary/system.jl:339-350) */
#define TPB_FIELD_PRESSURE 0
#define TPB_FIELD_DENSITY 1
#define TPB_FIELD_VOLUME 2 /* wall: boundary_model.cache.volume */
#define TPB_FIELD_WALL_VELOCITY 3 /* no-slip wall: boundary_model.cache.wall_velocity, ND x n values
                                   * (dummy_particles.jl:299-307) */

/* structure system (n = particle count; ND x ND matrices in the reference's memory layout,
 * entry (i, j) of particle p at [i + ND * j + ND * ND * p]: `out` holds ND * ND * n values) */
#define TPB_FIELD_DEFORMATION_GRADIENT 4 /* system.deformation_grad after the last kick */
#define TPB_FIELD_PK1_RHO2 5             /* system.pk1_rho2 (corrected PK1 / rho^2) */
#define TPB_FIELD_CORRECTION_MATRIX 6    /* system.correction_matrix (initialize!) */

typedef struct tpb_semi_s *tpb_semi_t; /* opaque; mirrors `Semidiscretization` */

/* `Semidiscretization(systems...; neighborhood_search=GridNeighborhoodSearch{ND}(cell_list=
 * FullGridCellList(; min_corner, max_corner)), parallelization_backend)`
 * (semidiscretization.jl:106-110; examples/fluid/dam_break_2d_gpu.jl:30-33). */
typedef struct {
    int32_t struct_size;   /* = sizeof(tpb_config), for forward compatibility */
    int32_t ndims;         /* 2 or 3 */
    int32_t eltype;        /* TPB_F32 | TPB_F64 */
    int32_t coords_eltype; /* TPB_F32 | TPB_F64, >= eltype */
    int32_t device;        /* CUDA device ordinal */
    int32_t ode_memory;    /* TPB_MEM_HOST | TPB_MEM_DEVICE */
    int32_t has_bounds;    /* 0: bounding box = extent of all initial coordinates */
    int32_t max_points_per_cell; /* accepted for compatibility; the counting sort has no cap */
    int32_t deterministic; /* 1: neighbours of a cell are ordered by particle index (default) */
    int32_t interact_variant; /* 0 = auto; 1 = per-particle sweep; 2 = cell-tile sweep */
    double min_corner[3];
    double max_corner[3];
} tpb_config;

/* `WeaklyCompressibleSPHSystem` fields read by the RHS (wcsph/system.jl:65-86) */
typedef struct {
    int32_t struct_size;
    int32_t kernel;                 /* TPB_KERNEL_* */
    int32_t density_calculator;     /* TPB_DENSITY_* */
    int32_t clip_negative_pressure; /* StateEquationCole CLIP */
    int32_t has_viscosity;          /* TPB_VISCOSITY_*: nothing | ArtificialViscosityMonaghan |
                                     * ViscosityMorris | ViscosityAdami (viscosity.jl:68-279) */
    int32_t has_diffusion;          /* DensityDiffusionMolteniColagrossi | nothing */
    double smoothing_length;
    double sound_speed, exponent, reference_density, background_pressure; /* StateEquationCole */
    double alpha, beta, epsilon;    /* Monaghan: alpha, beta, epsilon (viscosity.jl:68-76);
                                     * Morris / Adami: alpha = kinematic viscosity nu, epsilon */
    double delta;                   /* density_diffusion.jl:41-47 */
    double acceleration[3];         /* system.acceleration */
    double damping_coefficient;     /* SourceTermDamping (semidiscretization.jl:795-807); 0 = none */
    /* StateEquationAdaptiveCole (state_equations.jl:37-84; update_speed_of_sound!,
     * wcsph/system.jl:307-321): every kick sets sound_speed = clamp(max|v| / mach_number_target,
     * min_sound_speed, max_sound_speed) before the pressure update; `sound_speed` above is the
     * initial value (min_sound_speed).  adaptive_params_f32: the state equation's fields are
     * Float32 (its default literals) although eltype(system) may be Float64. */
    int32_t adaptive_sound_speed;
    int32_t adaptive_params_f32;
    double mach_number_target, min_sound_speed, max_sound_speed;
} tpb_fluid_params;

/* `WallBoundarySystem(ic, BoundaryModelDummyParticles(density, mass,
 * AdamiPressureExtrapolation(pressure_offset), kernel, h; state_equation,
 * clip_negative_pressure))` (wall_boundary/system.jl:22-43, dummy_particles.jl:52-77,142-149) */
typedef struct {
    int32_t struct_size;
    int32_t kernel;
    int32_t clip_negative_pressure; /* boundary model flag */
    int32_t sound_speed_from_fluid; /* 1: the boundary model shares the fluid's StateEquationAdaptiveCole */
    double smoothing_length;
    double sound_speed, exponent, reference_density, background_pressure;
    double pressure_offset;
    /* `BoundaryModelDummyParticles(...; viscosity)` (dummy_particles.jl:52-77): TPB_VISCOSITY_NONE =
     * free-slip wall; any model = no-slip wall: the wall velocity v_w = -sum_f v_f W / sum_f W
     * (compute_wall_velocity!, dummy_particles.jl:710-758) enters the model's viscous term of the
     * fluid (viscous_velocity, wall_boundary/system.jl:148-163).  Monaghan: alpha, beta, epsilon;
     * Morris / Adami: alpha = kinematic viscosity nu, epsilon. */
    int32_t has_viscosity;
    /* the boundary model's density calculator: TPB_WALL_DENSITY_ADAMI = AdamiPressureExtrapolation (and
     * BernoulliPressureExtrapolation, the same thing for a static wall); TPB_WALL_DENSITY_CONTINUITY =
     * ContinuityDensity (wall_boundary/rhs.jl:11-59, system.jl:78-90, :243-252; dummy_particles.jl:364-368,
     * :458-478): the wall density is integrated -- the system contributes ONE entry per wall particle to
     * v_ode / dv_ode (none to u_ode), pressure = state_equation(density) with the state equation's own
     * clip flag (eos_clip_negative_pressure) and then the boundary model's, and kick! fills the wall rows
     * of dv_ode with the continuity equation over the fluid neighbours.  Free-slip only; not combined
     * with slab ghosts. */
    int32_t density_calculator;
    double alpha, beta, epsilon;
    int32_t eos_clip_negative_pressure; /* StateEquationCole{..., CLIP} of the boundary model (ContinuityDensity) */
    int32_t reserved;
} tpb_wall_params;
#define TPB_WALL_DENSITY_ADAMI 0
#define TPB_WALL_DENSITY_CONTINUITY 1

/* `TotalLagrangianSPHSystem(ic; smoothing_kernel, smoothing_length, young_modulus, poisson_ratio,
 * clamped_particles, acceleration, penalty_force=PenaltyForceGanzenmueller(alpha), boundary_model=
 * BoundaryModelMonaghanKajtar(K, beta, boundary_particle_spacing, hydrodynamic_mass))`
 * (structure/total_lagrangian_sph/system.jl:76-184, penalty_force.jl:11-16,
 * wall_boundary/monaghan_kajtar.jl:18-34) -- BASELINE config 5 (examples/fsi/dam_break_plate_2d.jl).
 * Scalar material constants; the clamped particles are the LAST n - n_integrated particles (the
 * reference's constructor moves them there, system.jl:131-147); they rest at their initial positions unless
 * tpb_set_clamped_motion prescribes otherwise (clamped_particles_motion).  Scalar material constants here;
 * per particle: tpb_set_structure_material. */
#define TPB_BOUNDARY_NONE 0            /* boundary_model = nothing: no coupling with the fluid */
#define TPB_BOUNDARY_MONAGHAN_KAJTAR 1
#define TPB_BOUNDARY_DUMMY_PARTICLES 2 /* BoundaryModelDummyParticles{AdamiPressureExtrapolation} on the structure
                                        * (examples/fsi/hydrostatic_water_column_2d.jl:109-124): bm_* below */
typedef struct {
    int32_t struct_size;
    int32_t kernel;            /* TPB_KERNEL_* */
    int32_t has_penalty_force; /* PenaltyForceGanzenmueller | nothing */
    int32_t boundary_model;    /* TPB_BOUNDARY_* */
    double smoothing_length;
    double young_modulus, poisson_ratio;
    double penalty_alpha;
    double acceleration[3];
    double mk_K, mk_beta, mk_spacing; /* BoundaryModelMonaghanKajtar */
    /* TPB_BOUNDARY_DUMMY_PARTICLES: the boundary model's kernel, smoothing length, state equation, pressure
     * offset and clip flag (dummy_particles.jl:52-77); free-slip (no viscosity) */
    int32_t bm_kernel, bm_clip_negative_pressure;
    double bm_smoothing_length;
    double bm_sound_speed, bm_exponent, bm_reference_density, bm_background_pressure;
    double bm_pressure_offset;
    /* bm_wall_semantics != 0: the system is a moving `WallBoundarySystem(ic, model; prescribed_motion)` registered
     * with n_integrated = 0 -- the Adami hydrostatic term subtracts each particle's prescribed acceleration
     * (current_acceleration, wall_boundary/system.jl:129-142; dummy_particles.jl:652-654) and the Bernoulli term
     * exists only while the wall moves (dummy_particles.jl:680-694).  0: TotalLagrangianSPHSystem
     * (current_acceleration = 0, general/abstract_system.jl:121; Bernoulli term always, :696-707).
     * bm_bernoulli_factor: BernoulliPressureExtrapolation(factor); 0 = AdamiPressureExtrapolation */
    int32_t bm_wall_semantics, bm_reserved;
    double bm_bernoulli_factor;
} tpb_structure_params;

/* launch/traffic accounting of the last kick (what bench.py reports as gpu_launches) */
typedef struct {
    int64_t kernel_launches_total; /* since create */
    int64_t kicks, drifts;
    int64_t n_cells;
    int32_t launches_last_kick, launches_last_drift;
    int32_t interact_variant_used;
    int32_t reserved;
} tpb_stats;

const char *tpb_version(void);

/* message for the last failure on this handle (or of tpb_create when handle == NULL) */
const char *tpb_last_error(tpb_semi_t semi);

/* ---- construction: mirrors the `Semidiscretization` constructor ------------------------ */
int32_t tpb_create(const tpb_config *config, tpb_semi_t *out);
int32_t tpb_destroy(tpb_semi_t semi);

/* Systems are numbered in call order == position in the ODE vectors
 * (semidiscretization.jl:128-135).  `mass`: T[n].  Arrays are host pointers, copied. */
int32_t tpb_add_fluid_system(tpb_semi_t semi, const tpb_fluid_params *params, int64_t n,
                             const void *mass, int32_t *system_index);
/* `coords`: cT[ND x n]; `hydrodynamic_mass`, `initial_density`: T[n].  Several wall systems may be added
 * (e.g. tank and obstacle as separate `WallBoundarySystem`s, semidiscretization.jl:813-829 loops over all
 * ordered pairs) as long as their boundary models are the same (all tpb_wall_params fields equal; not with
 * ContinuityDensity): inside the library they form one static wall set, every system keeps its own index
 * for tpb_get_system_field / tpb_system_range / tpb_set_interaction (whose switches towards the fluid must
 * agree between the walls).  Different boundary models: TPB_ERR_UNSUPPORTED. */
int32_t tpb_add_wall_system(tpb_semi_t semi, const tpb_wall_params *params, int64_t n,
                            const void *coords, const void *hydrodynamic_mass,
                            const void *initial_density, int32_t *system_index);
/* `initial_coords`: cT[ND x n] (integrated particles first); `mass`, `material_density`: T[n];
 * `hydrodynamic_mass`: T[n] (boundary_model.hydrodynamic_mass; may be NULL with TPB_BOUNDARY_NONE).
 * The system contributes ND x n_integrated entries to u_ode and to v_ode
 * (system.jl:277-289).  One structure system per handle -- which may hold several bodies (several
 * `TotalLagrangianSPHSystem`s, a moving wall's dummy particles as extra clamped particles): TLSPH pairs its particles
 * once, in the initial configuration, so bodies that start further apart than the kernel support never interact
 * elastically.  Not combined with slab ghosts. */
int32_t tpb_add_structure_system(tpb_semi_t semi, const tpb_structure_params *params, int64_t n,
                                 int64_t n_integrated, const void *initial_coords, const void *mass,
                                 const void *material_density, const void *hydrodynamic_mass,
                                 int32_t *system_index);
/* `interaction_matrix[system, neighbor]` (semidiscretization.jl:157-187); default all true */
int32_t tpb_set_interaction(tpb_semi_t semi, int32_t system, int32_t neighbor, int32_t enabled);

/* `semidiscretize(semi, tspan)` (semidiscretization.jl:293-396): sizes the cell grid, sorts
 * the static wall particles, allocates every device buffer.  `u0_ode`: host cT[n_u] initial
 * coordinates (used for the automatic bounding box; may be NULL when has_bounds != 0). */
int32_t tpb_semidiscretize(tpb_semi_t semi, const void *u0_ode);

/* lengths of u_ode / v_ode and the 0-based range of one system inside them
 * (`ranges_u` / `ranges_v`, semidiscretization.jl:128-135) */
int32_t tpb_ode_sizes(tpb_semi_t semi, int64_t *n_u, int64_t *n_v);
int32_t tpb_system_range(tpb_semi_t semi, int32_t system, int64_t *u_first, int64_t *u_len,
                         int64_t *v_first, int64_t *v_len);

/* ---- slab decomposition (one handle per GPU) -------------------------
 * The fluid system was added with a capacity of n particles.  A rank keeps its owned particles
 * in rows [0, n_targets) of the ODE vectors and appends the ghost particles received from its
 * slab neighbours in rows [n_targets, n_active): all n_active particles are binned and act as
 * neighbours, only the first n_targets get a dv.  Ghost masses travel with the particles. */
int32_t tpb_set_fluid_count(tpb_semi_t semi, int64_t n_active, int64_t n_targets);
/* overwrite mass[first, first + count): host or device pointer according to ode_memory */
int32_t tpb_set_fluid_mass(tpb_semi_t semi, int64_t first, int64_t count, const void *mass);

/* ---- the hot path ----------------------------------------------------------------------- */
/* `kick!(dv_ode, v_ode, u_ode, p, t)` (semidiscretization.jl:589-612): dv_ode is fully
 * overwritten.  Host pointers are not retained after return. */
int32_t tpb_kick(tpb_semi_t semi, void *dv_ode, const void *v_ode, const void *u_ode, double t);
/* `drift!(du_ode, v_ode, u_ode, p, t)` (semidiscretization.jl:522-536) */
int32_t tpb_drift(tpb_semi_t semi, void *du_ode, const void *v_ode, const void *u_ode, double t);

/* ---- outputs / diagnostics ----------------------------------------------------------------- */
/* system data after the last kick, in the system's own particle order; `out`: host T[n]
 * (TPB_FIELD_WALL_VELOCITY: host T[ndims * n], particle-major; n stays the particle count) */
int32_t tpb_get_system_field(tpb_semi_t semi, int32_t system, int32_t field, void *out, int64_t n);
/* Test hook: rebuilds the grids for `u_ode` and writes every ordered neighbour pair
 * (i in `system`, j in `neighbor`, both 0-based) with |x_i - x_j|^2 <= R^2 for the radius of
 * that ordered pair (neighborhood_search.jl:73-77,134-140), in unspecified order.
 * `*count` receives the true number; TPB_ERR_CAPACITY if it exceeds `capacity`. */
int32_t tpb_neighbor_pairs(tpb_semi_t semi, int32_t system, int32_t neighbor, const void *u_ode,
                           int64_t capacity, int32_t *out_i, int32_t *out_j, int64_t *count);
/* blocks until all queued work of the handle is done; returns a deferred device-side error
 * (TPB_ERR_OUT_OF_BOUNDS) if one was raised by an asynchronous kick */
int32_t tpb_synchronize(tpb_semi_t semi);
/* `stream` is a cudaStream_t; NULL selects the handle's own stream */
int32_t tpb_set_stream(tpb_semi_t semi, void *stream);
int32_t tpb_get_stats(tpb_semi_t semi, tpb_stats *out);
/* `system_sound_speed(fluid)` as of the last kick (StateEquationAdaptiveCole; else the constant) */
int32_t tpb_get_sound_speed(tpb_semi_t semi, double *out);
/* StateEquationAdaptiveCole on several slabs (update_speed_of_sound!, wcsph/system.jl:307-321, takes the
 * maximum over ALL fluid particles): `tpb_max_speed2` reduces max |v|^2 over the handle's owned rows of
 * the device vector `v_ode` into the 8-byte device word `out_bits` (the IEEE bit pattern of the value,
 * zero-extended: the maximum of non-negative numbers is the maximum of their bit patterns, so an
 * integer MAX all-reduce over the ranks finishes the job); `tpb_set_max_speed2` hands the reduced word
 * (device memory) to the next tpb_kick, which then skips its own reduction.  Both are stream-ordered;
 * a kick of a handle with ghost particles and an adaptive state equation fails with TPB_ERR_STATE
 * unless tpb_set_max_speed2 preceded it. */
int32_t tpb_max_speed2(tpb_semi_t semi, const void *v_ode, void *out_bits);
/* ---- SplitIntegrationCallback (callbacks/split_integration.jl): the structure is integrated by its own
 * sub-integrator with smaller steps.  `tpb_set_integrate_structure(h, 0)` = `semi.integrate_tlsph[] = false`
 * (semidiscretization.jl:149, :868-880, :545-552): tpb_kick / tpb_drift leave the structure's rows of dv / du
 * zero, the fluid still feels the structure.  `tpb_structure_fluid_force` = update_systems_and_nhs +
 * other_interaction_split! (:232-234, :444-470): the force of the fluid on the integrated structure particles
 * for the state (v_ode, u_ode), ND x n_integrated values.  `tpb_kick_structure` = kick_split! (:352-371) on the
 * structure's own vectors (ND x n_integrated each): structure <- structure + `dv_const` (may be NULL) +
 * gravity.  Device vectors only; stream-ordered. */
int32_t tpb_set_integrate_structure(tpb_semi_t semi, int32_t enabled);
int32_t tpb_structure_fluid_force(tpb_semi_t semi, void *dv_split, const void *v_ode, const void *u_ode);
int32_t tpb_kick_structure(tpb_semi_t semi, void *dv_split, const void *v_split, const void *u_split,
                           const void *dv_const);
int32_t tpb_set_max_speed2(tpb_semi_t semi, const void *bits);
/* `young_modulus` / `poisson_ratio` of the structure given per particle (total_lagrangian_sph/system.jl:108-161: scalars
 * or vectors; penalty_force.jl:42-53 uses E of both particles of a pair): T[n] each in the system's particle order
 * (integrated particles first), HOST pointers, copied.  Overrides the scalars of tpb_structure_params from the next
 * kick on.  Also what lets several `TotalLagrangianSPHSystem`s with different materials share the one structure slot
 * (TLSPH pairs its particles in the initial configuration, so bodies that start apart never interact elastically). */
int32_t tpb_set_structure_material(tpb_semi_t semi, const void *young_modulus, const void *poisson_ratio);
/* ---- SortingCallback (callbacks/sorting.jl:100-157: sort_particles! / sort_system!): reorders the rows of
 * `system` inside the caller's (v_ode, u_ode) -- in place -- by the grid cell of their current coordinates (linear
 * cell index, x fastest; inside a cell the previous order is kept, so every later kick sums its neighbours in the
 * same order and gives bit-identical results, row for row), together with the library's per-particle masses (the
 * reference leaves those as a TODO and assumes uniform particles).  Only fluid systems are sorted
 * (RequiresSortingSystem, sorting.jl:5); for any other system the call returns TPB_OK and does nothing.  The
 * library counting-sorts its own records at every kick whatever the order of the ODE vectors is; sorted vectors
 * only make that gather (and the scatter of dv) contiguous: 1 M particles in random order cost 5.5 % of a step,
 * 10 M 8.8 %.  Host or device pointers according to ode_memory; stream-ordered on device
 * vectors.  Not with slab ghosts. */
int32_t tpb_sort_system(tpb_semi_t semi, int32_t system, void *v_ode, void *u_ode);
/* ---- DensityReinitializationCallback (callbacks/density_reinit.jl:83-121 -> reinit_density!, schemes/fluid/
 * weakly_compressible_sph/system.jl:398-415; Panizzo 2007): the fluid's density rows of `v_ode` are replaced, in place,
 * by the Shepard-corrected summation density
 *     rho~_a = sum_b m_b W_ab (fluid + wall),   c_a = sum_b (m_b / rho_b) W_ab,   rho_a = rho~_a / c_a,
 * where a fluid neighbour enters c with rho~_b (the reference overwrites v before it computes the coefficient; its
 * `v[end, :]` is a view for the CPU's PtrArray) and a wall neighbour with its boundary-model density, which is updated
 * here for the given state first (initialize_reinit_cb!, :100-112).  ContinuityDensity fluids only (SummationDensity:
 * no-op); fluid + Adami walls.  The reference has no numerical test for this callback: parity rests on the
 * restatement in oracle/ and on "a uniform field stays uniform". */
int32_t tpb_reinit_density(tpb_semi_t semi, void *v_ode, const void *u_ode);
/* ---- PrescribedMotion (schemes/boundary/prescribed_motion.jl:95-121; apply_prescribed_motion!,
 * wall_boundary/system.jl:199-205 and total_lagrangian_sph/system.jl:436-447).  The movement function is the
 * caller's: before a kick it hands over where the clamped particles of the structure system (all particles of
 * a moving wall registered with n_integrated = 0) are now and how they move.  `coords`: cT[ND x n_clamped],
 * `velocity`, `acceleration`: T[ND x n_clamped], HOST pointers, copied; `coords` may be NULL (particles stay
 * where the last call put them).  `is_moving` = the motion's is_moving(t): 0 = velocity and acceleration count
 * as zero (and may be NULL), the particles rest at their last position.  Takes effect with the next tpb_kick;
 * a captured CUDA graph of kicks does not see it. */
int32_t tpb_set_clamped_motion(tpb_semi_t semi, const void *coords, const void *velocity,
                               const void *acceleration, int32_t is_moving);

/* ---- device ODE-vector algebra ----------------------------------------------------------------
 * What a GPU-resident ODE-vector type binds for the integrator's broadcasts (the reference's
 * CPU counterpart is `ThreadedBroadcastArray`, util.jl:183-303).  TPB_MEM_DEVICE handles only;
 * pointers are device pointers, work is queued on the handle's stream. */
/* y = a * x + b * y */
int32_t tpb_vec_axpby(tpb_semi_t semi, int64_t n, int32_t eltype, double a, const void *x, double b, void *y);
/* one 2N-storage Runge-Kutta stage: tmp = A * tmp + dt * rhs; state += B * tmp */
int32_t tpb_vec_rk2n_stage(tpb_semi_t semi, int64_t n, int32_t eltype, double A, double B, double dt,
                           const void *rhs, void *tmp, void *state);
int32_t tpb_vec_fill(tpb_semi_t semi, int64_t n, int32_t eltype, double value, void *x);
/* y = a0 x0 + a1 x1 + a2 x2 + a3 x3; a NULL x skips its term, any x may alias y.  The register
 * updates of a 3S*+ low-storage stage (RDPK3SpFSAL35, the integrator of examples/fluid/dam_break_3d.jl:85:
 * u = gamma1 u + gamma2 tmp + gamma3 uprev + beta dt k) in one pass. */
int32_t tpb_vec_lincomb4(tpb_semi_t semi, int64_t n, int32_t eltype, double a0, const void *x0, double a1,
                         const void *x1, double a2, const void *x2, double a3, const void *x3, void *y);
/* `SymplecticPositionVerlet` velocity / density update of one WCSPH system
 * (ext/TrixiParticlesOrdinaryDiffEqSymplecticRKExt.jl:137-171) on its rows of v_ode: `nvars` entries
 * per particle (ndims velocities [+ density]); `du` holds the half-step state on entry. */
int32_t tpb_vec_verlet_update(tpb_semi_t semi, int64_t n_particles, int32_t ndims, int32_t nvars, int32_t eltype,
                              double dt, const void *kdu, const void *duprev, void *du);
/* test hook: out[i] = div_fast(x, y[i]) with the library's fast division (util.jl:3-5,
 * ext/TrixiParticlesCUDAExt.jl:12-33; reference test: test/examples/gpu.jl:30-78) */
int32_t tpb_vec_div_fast(tpb_semi_t semi, int64_t n, int32_t eltype, double x, const void *y, void *out);
/* max_k x[offset + k * stride], k < count, to the host (synchronises); `max_x_coord` of
 * general/custom_quantities.jl is (count = nparticles, stride = ND, offset = 0) on u_ode */
int32_t tpb_vec_strided_max(tpb_semi_t semi, int64_t count, int32_t eltype, int32_t stride, int32_t offset,
                            const void *x, double *out_host);

/* sqrt(mean_i (err_i / (abstol + reltol * max(|u_prev_i|, |u_i|)))^2) to the host (synchronises):
 * the residual norm an adaptive integrator (RDPK3SpFSAL35 in the reference's examples) accepts or
 * rejects a step on -- OrdinaryDiffEq `calculate_residuals` + `ODE_DEFAULT_NORM` */
int32_t tpb_vec_wrms_norm(tpb_semi_t semi, int64_t n, int32_t eltype, const void *err, const void *u_prev,
                          const void *u, double abstol, double reltol, double *out_host);

/* Phase timing of `tpb_kick` with CUDA events on the handle's stream (the reference's
 * TimerOutputs sections "update systems and nhs" / "system interaction",
 * semidiscretization.jl:596-606).  `tpb_set_profiling(max_kicks)` arms recording for the next
 * `max_kicks` kicks (0 disables); `tpb_get_phase_times` synchronises, writes the mean
 * duration in ms of each phase into ms_mean[TPB_N_PHASES - 1] and re-arms. */
#define TPB_PHASE_REBUILD 0  /* NHS rebuild: cell keys, counting sort, reorder + EOS */
#define TPB_PHASE_DENSITY 1  /* summation density sweep (SummationDensity only) */
#define TPB_PHASE_BOUNDARY 2 /* Adami wall pressure extrapolation */
#define TPB_PHASE_INTERACT 3 /* fluid-fluid + fluid-wall interact!, source terms */
#define TPB_PHASE_END 4
#define TPB_N_PHASES 5
int32_t tpb_set_profiling(tpb_semi_t semi, int32_t max_kicks);
int32_t tpb_get_phase_times(tpb_semi_t semi, double *ms_mean, int32_t *n_kicks);

/* ---- ghost exchange between the slabs of one box over peer memory ------------------------------
 * The reference has no distributed path; SURVEY.md section 8(e) describes the decomposition.
 * One process per GPU.  `tpb_peer_alloc` returns zeroed device memory that `tpb_peer_export` can
 * describe in 64 bytes (CUDA IPC); the neighbour process maps it with `tpb_peer_import`.
 * `tpb_halo_pack` (sender): one kernel gathers the rows (x[ND], v[NV]) of `candidates` from the
 * ODE vectors, marks rows whose x is not on the live side of `threshold` (side < 0: x < threshold,
 * side > 0: x >= threshold) with x = NaN, stores them into the neighbour's receive area
 * (`peer_u`, `peer_v`) over NVLink and, from its last block, publishes `epoch` in `peer_flag`.
 * `done_counter` is a 64-bit device counter private to this (sender, neighbour) pair;
 * `done_target_blocks` is the number of blocks all earlier calls on it have launched, `*blocks_out`
 * the number this call launched.  `tpb_halo_install` (receiver): one kernel waits until the flags
 * (NULL: no neighbour on that side) hold at least `epoch`, then copies `n_rows` rows from the
 * receive area behind the owned particles (`u_ghost`, `v_ghost`).  A neighbour that does not
 * deliver within `timeout_s` sets bit 1 of `*timed_out_flag` (device int; NULL: the handle's own
 * status word, reported by the next `tpb_synchronize` as TPB_ERR_STATE) instead of hanging the GPU.
 * All work is queued on the handle's stream; nothing synchronises with the host. */
int32_t tpb_peer_alloc(int64_t bytes, void **out);
int32_t tpb_peer_free(void *ptr);
int32_t tpb_peer_export(void *ptr, void *handle64);
int32_t tpb_peer_import(const void *handle64, void **out);
int32_t tpb_peer_close(void *ptr);
int32_t tpb_halo_pack(tpb_semi_t semi, int32_t side, double threshold, const void *u, const void *v,
                      const int64_t *candidates, int64_t n_candidates, void *peer_u, void *peer_v,
                      void *done_counter, int64_t done_target_blocks, void *peer_flag, uint32_t epoch,
                      int32_t *blocks_out);
int32_t tpb_halo_install(tpb_semi_t semi, int64_t n_rows, const void *stage_u, const void *stage_v,
                         void *u_ghost, void *v_ghost, const void *flag_left, const void *flag_right,
                         uint32_t epoch, double timeout_s, void *timed_out_flag);

/* page-lock / unlock caller memory so TPB_MEM_HOST transfers run at full PCIe speed */
int32_t tpb_host_register(void *ptr, int64_t bytes);
int32_t tpb_host_unregister(void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* TPB200_H */
