/*
 * sph_oracle.h -- CPU oracle for the WCSPH right-hand side of TrixiParticles.jl.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's algorithm
 * (each function cites the reference file:line it follows).  It is the checker for the
 * CUDA path in trixiparticles.jl_b200/; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product path never
 * links or calls anything in oracle/.
 *
 * Pinning: the oracle reproduces the reference's own known-answer tests
 * (tests/test_oracle_golden.py): Monaghan / Morris / Adami viscosity pairs, Adami kernel weights, the
 * no-slip wall velocities (constant and staggered profiles, dummy_particles.jl test :104-303),
 * Cole EOS inverse values, kernel normalisation, conservation properties, and the first
 * samples of the dam-break surge-front trace.  The neighbour search itself lives in the
 * third-party PointNeighbors.jl (compat 0.6.6, not under /root/reference); its published
 * semantics are restated here (predicate d^2 <= R^2, pos_diff = x_i - y_j, self pair
 * included) and anchored on the reference's call sites.
 *
 * Three precision combinations are compiled from one implementation file:
 *   suffix _f64    : T = double, cT = double
 *   suffix _f32    : T = float,  cT = float
 *   suffix _f32c64 : T = float,  cT = double   (Float32 system with Float64 coordinates)
 * where T = eltype(system) and cT = coordinates_eltype (semidiscretization.jl:326-332).
 *
 * All real-valued parameters cross this interface as double; callers compute them in the
 * target precision first so the conversion back to T is exact.
 */
#ifndef SPH_ORACLE_H
#define SPH_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* `has_viscosity` selects the model; Morris / Adami carry their kinematic viscosity nu in `alpha` */
enum { ORC_VISCOSITY_NONE = 0, ORC_VISCOSITY_MONAGHAN = 1, ORC_VISCOSITY_MORRIS = 2, ORC_VISCOSITY_ADAMI = 3 };
enum { ORC_KERNEL_WENDLAND_C2 = 0, ORC_KERNEL_SCHOENBERG_CUBIC = 1, ORC_KERNEL_WENDLAND_C4 = 2,
       ORC_KERNEL_WENDLAND_C6 = 3, ORC_KERNEL_SCHOENBERG_QUARTIC = 4, ORC_KERNEL_SCHOENBERG_QUINTIC = 5 };
enum { ORC_DENSITY_CONTINUITY = 0, ORC_DENSITY_SUMMATION = 1 };

/* WeaklyCompressibleSPHSystem fields (wcsph/system.jl:65-86) that the RHS reads. */
typedef struct {
    int32_t ndims;                  /* 2 or 3 */
    int32_t kernel;                 /* ORC_KERNEL_* */
    int32_t density_calculator;     /* ORC_DENSITY_* */
    int32_t clip_negative_pressure; /* StateEquationCole{..., CLIP} */
    int32_t has_viscosity;          /* ArtificialViscosityMonaghan or nothing */
    int32_t has_diffusion;          /* DensityDiffusionMolteniColagrossi or nothing */
    int32_t reserved0, reserved1;   /* reserved1 != 0: reserved0 = bits of a Float32 B (see eos_B) */
    double smoothing_length;
    double sound_speed, exponent, reference_density, background_pressure;
    double alpha, beta, epsilon;    /* viscosity.jl:68-76 */
    double delta;                   /* density_diffusion.jl:41-47 */
    double acceleration[3];
    double damping_coefficient;     /* SourceTermDamping, 0 = none */
} orc_fluid_params;

/* WallBoundarySystem + BoundaryModelDummyParticles{AdamiPressureExtrapolation}
 * (wall_boundary/system.jl:22-43, dummy_particles.jl:52-77,142-149). */
typedef struct {
    int32_t kernel;
    int32_t clip_negative_pressure; /* boundary model's own clip flag */
    int32_t eos_clip_negative_pressure;
    int32_t reserved0;
    double smoothing_length;
    double sound_speed, exponent, reference_density, background_pressure;
    double pressure_offset;
    /* boundary_model.viscosity (dummy_particles.jl:52-77): nothing = free-slip; any model = no-slip
     * wall (compute_wall_velocity!, :710-758).  Morris / Adami carry nu in `visc_alpha`. */
    int32_t has_viscosity; /* ORC_VISCOSITY_* */
    /* boundary model's density calculator: 0 = AdamiPressureExtrapolation; ORC_WALL_CONTINUITY =
     * ContinuityDensity (wall_boundary/rhs.jl:11-59): the wall density is integrated, v_ode / dv_ode
     * carry one row per wall particle BEHIND the fluid's rows (systems in the order fluid, wall),
     * pressure = state_equation(density) (dummy_particles.jl:458-478) */
    int32_t density_calculator;
    double visc_alpha, visc_beta, visc_epsilon;
} orc_wall_params;
enum { ORC_WALL_ADAMI = 0, ORC_WALL_CONTINUITY = 1 };

/* TotalLagrangianSPHSystem (structure/total_lagrangian_sph/system.jl:76-106) with scalar material
 * constants, PenaltyForceGanzenmueller (penalty_force.jl) and, for the coupling with a fluid, a
 * BoundaryModelMonaghanKajtar (wall_boundary/monaghan_kajtar.jl:18-34). */
enum { ORC_BOUNDARY_NONE = 0, ORC_BOUNDARY_MONAGHAN_KAJTAR = 1, ORC_BOUNDARY_DUMMY_PARTICLES = 2 };
typedef struct {
    int32_t ndims, kernel;
    int32_t has_penalty;      /* PenaltyForceGanzenmueller or nothing */
    int32_t boundary_model;   /* ORC_BOUNDARY_* */
    double smoothing_length;
    double young_modulus, poisson_ratio;
    double penalty_alpha;
    double acceleration[3];
    double mk_K, mk_beta, mk_spacing; /* BoundaryModelMonaghanKajtar(K, beta, boundary_particle_spacing, ...) */
    /* ORC_BOUNDARY_DUMMY_PARTICLES: BoundaryModelDummyParticles(hydrodynamic_density, hydrodynamic_mass,
     * AdamiPressureExtrapolation(pressure_offset), kernel, smoothing_length; state_equation) on the structure
     * (examples/fsi/hydrostatic_water_column_2d.jl:109-124) */
    int32_t bm_kernel, bm_clip_negative_pressure;
    double bm_smoothing_length;
    double bm_sound_speed, bm_exponent, bm_reference_density, bm_background_pressure;
    double bm_pressure_offset;
    /* prescribed motion of the clamped particles (boundary/prescribed_motion.jl:95-121).  bm_wall_semantics:
     * the particles are a moving WallBoundarySystem -- the Adami hydrostatic term subtracts the particle's
     * prescribed acceleration (wall_boundary/system.jl:133-142, dummy_particles.jl:652-654) and the Bernoulli
     * term exists only while the wall moves (dummy_particles.jl:680-694); 0 = TotalLagrangianSPHSystem
     * (current_acceleration = 0, abstract_system.jl:121; Bernoulli term always, dummy_particles.jl:696-707).
     * bm_bernoulli_factor: BernoulliPressureExtrapolation's factor, 0 = AdamiPressureExtrapolation */
    int32_t bm_wall_semantics, bm_reserved;
    double bm_bernoulli_factor;
} orc_tlsph_params;

#define ORC_DECLARE(SUF, T, CT)                                                              \
    double orc_kernel_##SUF(int kernel, int ndims, double r, double h);                      \
    double orc_kernel_unsafe_##SUF(int kernel, int ndims, double r, double h);               \
    double orc_kernel_deriv_div_r_##SUF(int kernel, int ndims, double r, double h);          \
    double orc_eos_##SUF(double c, double gamma, double rho0, double p_bg, int clip,         \
                         double density);                                                    \
    double orc_inverse_eos_##SUF(double c, double gamma, double rho0, double p_bg,           \
                                 double pressure);                                           \
    void orc_viscosity_pair_##SUF(int kernel, int ndims, double h, double alpha,             \
                                  double beta, double epsilon, double sound_speed,           \
                                  double m_b, double rho_a, double rho_b,                    \
                                  const double *v_diff, const double *pos_diff,              \
                                  double *dv_out);                                           \
    void orc_viscosity_pair_nu_##SUF(int model, int kernel, int ndims, double h, double nu,  \
                                     double epsilon, double m_a, double m_b, double rho_a,   \
                                     double rho_b, const double *v_diff,                     \
                                     const double *pos_diff, double *dv_out);                \
    void orc_interact_pair_##SUF(const orc_fluid_params *fp, int neighbor_is_wall,           \
                                 double m_b, double rho_a, double rho_b, double p_a,         \
                                 double p_b, const double *v_a, const double *v_b,           \
                                 const double *pos_diff, double *dv_out, double *drho_out);  \
    int64_t orc_pairs_bruteforce_##SUF(int ndims, int64_t nx, const CT *x, int64_t ny,       \
                                       const CT *y, double radius, int64_t capacity,         \
                                       int32_t *out_i, int32_t *out_j);                      \
    int64_t orc_pairs_grid_##SUF(int ndims, int64_t nx, const CT *x, int64_t ny,             \
                                 const CT *y, double radius, int64_t capacity,               \
                                 int32_t *out_i, int32_t *out_j);                            \
    int orc_kick_##SUF(const orc_fluid_params *fp, const orc_wall_params *wp, int64_t n_f,   \
                       const T *mass_f, int64_t n_w, const CT *coords_w, const T *mass_w,    \
                       const T *v_ode, const CT *u_ode, T *dv_ode, T *pressure_f,            \
                       T *density_f, T *pressure_w, T *density_w, T *volume_w,               \
                       int use_grid, int nthreads);                                          \
    /* as orc_kick; wall_velocity_w (ND x n_w, may be NULL): boundary_model.cache.wall_velocity */ \
    int orc_kick_noslip_##SUF(const orc_fluid_params *fp, const orc_wall_params *wp, int64_t n_f, \
                       const T *mass_f, int64_t n_w, const CT *coords_w, const T *mass_w,    \
                       const T *v_ode, const CT *u_ode, T *dv_ode, T *pressure_f,            \
                       T *density_f, T *pressure_w, T *density_w, T *volume_w,               \
                       T *wall_velocity_w, int use_grid, int nthreads);                      \
    void orc_drift_##SUF(int ndims, int nvars_v, int64_t n_f, const T *v_ode, CT *du_ode);    \
    /* DensityReinitializationCallback: Shepard-corrected summation density (see the .inc) */ \
    int orc_reinit_density_##SUF(const orc_fluid_params *fp, const orc_wall_params *wp,      \
                                 int64_t n_f, const T *mass_f, int64_t n_w,                   \
                                 const CT *coords_w, const T *mass_w, const T *v_ode,         \
                                 const CT *u_ode, T *out_density);                            \
    int orc_tlsph_correction_matrix_##SUF(const orc_tlsph_params *sp, int64_t n, const CT *x0, \
                                          const T *mass, const T *rho, T *L);                \
    int orc_tlsph_update_##SUF(const orc_tlsph_params *sp, int64_t n, const CT *x0,          \
                               const CT *x_cur, const T *mass, const T *rho, const T *L,     \
                               T *F_out, T *pk1_rho2);                                       \
    int orc_tlsph_interact_##SUF(const orc_tlsph_params *sp, int64_t n, int64_t n_int,       \
                                 const CT *x0, const CT *x_cur, const T *mass, const T *rho, \
                                 const T *F, const T *pk1_rho2, T *dv);                      \
    int orc_kick_fsi_##SUF(const orc_fluid_params *fp, const orc_wall_params *wp,            \
                           const orc_tlsph_params *sp, int64_t n_f, const T *mass_f,         \
                           int64_t n_w, const CT *coords_w, const T *mass_w, int64_t n_s,    \
                           int64_t n_s_int, const CT *x0_s, const T *mass_s, const T *rho_s, \
                           const T *hydro_mass_s, const T *L, const T *v_ode,                \
                           const CT *u_ode, T *dv_ode, T *F_out, T *pk1_out, int nthreads);  \
    /* as orc_kick_fsi; pressure_s / density_s (n_s, may be NULL): the structure's boundary-model    \
     * pressure and density (dummy particles: Adami extrapolation) */                                \
    int orc_kick_fsi2_##SUF(const orc_fluid_params *fp, const orc_wall_params *wp,           \
                           const orc_tlsph_params *sp, int64_t n_f, const T *mass_f,         \
                           int64_t n_w, const CT *coords_w, const T *mass_w, int64_t n_s,    \
                           int64_t n_s_int, const CT *x0_s, const T *mass_s, const T *rho_s, \
                           const T *hydro_mass_s, const T *L, const T *v_ode,                \
                           const CT *u_ode, T *dv_ode, T *F_out, T *pk1_out, T *pressure_s,  \
                           T *density_s, int nthreads);                                      \
    /* as orc_kick_fsi2 with prescribed motion of the clamped particles: x_clamped (current positions), \
     * v_clamped, a_clamped: (n_s - n_s_int) x ND, each may be NULL (positions: the initial ones;        \
     * velocity NULL = not moving now, velocity and acceleration count as zero) */                       \
    int orc_kick_fsi3_##SUF(const orc_fluid_params *fp, const orc_wall_params *wp,           \
                           const orc_tlsph_params *sp, int64_t n_f, const T *mass_f,         \
                           int64_t n_w, const CT *coords_w, const T *mass_w, int64_t n_s,    \
                           int64_t n_s_int, const CT *x0_s, const T *mass_s, const T *rho_s, \
                           const T *hydro_mass_s, const T *L, const T *v_ode,                \
                           const CT *u_ode, T *dv_ode, T *F_out, T *pk1_out, T *pressure_s,  \
                           T *density_s, const CT *x_clamped, const T *v_clamped,            \
                           const T *a_clamped, int nthreads);

ORC_DECLARE(f64, double, double)
ORC_DECLARE(f32, float, float)
ORC_DECLARE(f32c64, float, double)

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
