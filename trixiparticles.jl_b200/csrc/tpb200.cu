// tpb200.cu -- libtpb200.so: C ABI (include/tpb200.h) + host orchestration of the sm_100a
// kernels in tpb_nhs.cuh / tpb_sweeps.cuh / tpb_tiles.cuh.
//
// Drop-in boundary: /root/reference/src/general/semidiscretization.jl `kick!` (:589-612) and
// `drift!` (:522-536).  The library owns all device memory; nothing is allocated per call.
#include <cuda_runtime.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../include/tpb200.h"
#include "tpb_device.cuh"
#include "tpb_nhs.cuh"
#include "tpb_sweeps.cuh"
#include "tpb_tiles.cuh"
#include "tpb_structure.cuh"
#include "tpb_vec.cuh"
#include "tpb_halo.cuh"

using namespace tpb;

// The library is built from several translation units of THIS file (build.py): one per
// (ndims, eltype, coordinates eltype) combination, compiled with -DTPB_TU_ND / _T / _CT / _TAG and
// holding the kernels of that combination behind five entry functions, plus the main one with the
// C ABI.  Everything shared lives in this named namespace with internal linkage per unit.
namespace tpbhost {

static thread_local std::string g_create_error;

struct Semi {
    tpb_config cfg{};
    std::string err;
    // systems (one fluid, at most one wall; numbered in call order)
    int n_systems = 0;
    int fluid_index = -1, wall_index = -1;
    // several WallBoundarySystems with the same boundary model are ONE static wall set inside the library
    // (the fluid sums over all wall particles alike; wall <-> wall never interact): `wall_index` is the
    // first one, every system keeps its own index, particle range and interaction switches
    struct WallPart {
        int index;
        int64_t first, n;
        int fluid_from_wall = 1, wall_from_fluid = 1;
    };
    std::vector<WallPart> wall_parts;
    const WallPart *wall_part(int system) const
    {
        for (const WallPart &w : wall_parts)
            if (w.index == system) return &w;
        return nullptr;
    }
    tpb_fluid_params fp{};
    tpb_wall_params wp{};
    int64_t n_f = 0, n_w = 0;   // n_f: capacity of the fluid system (particles allocated)
    int64_t n_act = 0, n_tgt = 0;  // particles binned as neighbours / particles that get a dv (slab ghosts: n_tgt < n_act)
    std::vector<unsigned char> h_mass_f, h_coords_w, h_mass_w, h_dens_w;
    int interaction[2][2] = {{1, 1}, {1, 1}};
    bool ready = false;

    // structure (TotalLagrangianSPHSystem), at most one; clamped particles are the last n_s - n_s_int
    int struct_index = -1;
    tpb_structure_params sp{};
    int64_t n_s = 0, n_s_int = 0;
    int struct_fluid[2] = {1, 1};  // interaction_matrix[structure, fluid], [fluid, structure]
    int struct_self = 1;           // interaction_matrix[structure, structure]
    int integrate_structure = 1;   // semi.integrate_tlsph[] (semidiscretization.jl:149): 0 with a SplitIntegrationCallback
    std::vector<unsigned char> h_x0_s, h_mass_s, h_rho_s, h_hydro_s;
    void *d_x0_s = nullptr, *d_xcur_s = nullptr, *d_mass_s = nullptr, *d_rho_s = nullptr, *d_hydro_s = nullptr;
    void *d_L_s = nullptr, *d_F_s = nullptr, *d_pk1_s = nullptr, *d_As = nullptr, *d_Bs = nullptr;
    int *d_nbr_start = nullptr, *d_nbr = nullptr, *d_scell_start = nullptr;
    // structure with BoundaryModelDummyParticles: sorted -> own index, Adami pressure (sorted / own order), density
    int *d_sperm = nullptr;
    void *d_Ps = nullptr, *d_p_s = nullptr, *d_rhoh_s = nullptr;
    // prescribed motion of the clamped particles (tpb_set_clamped_motion): d_xcl_s = where every particle is when
    // u_ode does not say (n_s x ND, the clamped tail is what counts; starts as the initial coordinates);
    // velocity / acceleration of the clamped particles ((n_s - n_s_int) x ND), used while clamped_moving
    void *d_xcl_s = nullptr, *d_vcl_s = nullptr, *d_acl_s = nullptr;
    int clamped_moving = 0;
    void *d_mat_s = nullptr;  // per-particle (lambda, mu, E) of the structure (tpb_set_structure_material), else scalars
    // tpb_sort_system on device vectors: gather targets (allocated at the first call; host mode uses d_du / d_dv)
    void *d_sort_u = nullptr, *d_sort_v = nullptr;

    // geometry of the shared cell grid (double; typed copies are built per call)
    double cell_size = 0, origin[3] = {0, 0, 0}, lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    int ncell[3] = {1, 1, 1};
    int row_axis = 0;  // coordinate axis of the cell rows; origin/lo/hi/ncell hold entries 0 and row_axis swapped
    int64_t ncells = 0;
    int xsplit = 1;       // cells are cell_size / xsplit wide in x
    double gap_tol = 0;   // rounding bound of cell-boundary coordinates in cT

    // device
    cudaStream_t own_stream = nullptr, stream = nullptr;
    void *d_mass_f = nullptr;
    void *d_u = nullptr, *d_v = nullptr, *d_dv = nullptr, *d_du = nullptr;  // host-mode staging
    int *d_key = nullptr, *d_slot = nullptr, *d_tmp_perm = nullptr, *d_perm_f = nullptr;
    int *d_count = nullptr, *d_fcell_start = nullptr, *d_wcell_start = nullptr;
    int *d_block_sums = nullptr, *d_flags = nullptr;
    // one-pass cell scan + tile table (k_scan_cells_tiles): look-back status words, block tickets
    unsigned long long *d_scan_status = nullptr;
    unsigned long long *d_scan_ticket = nullptr;
    int scan_blocks = 0, scan_rows_per_block = 0;  // 0 blocks: grid shape not supported, three-kernel scan instead
    int subkey_mode = 1;           // order inside a cell: 0 previous index, 1 position key (tpb_nhs.cuh, TPB_SUBKEY)
    bool count_clean = false;      // the cell histogram is all zero (left so by k_scan_cells_tiles)
    bool wall_prep_done = false;   // this kick's rebuild launch has sorted the wall tiles into empty / active
    int *h_flags = nullptr;  // pinned
    void *d_A = nullptr, *d_B = nullptr, *d_P = nullptr;
    void *d_Aw = nullptr, *d_Ww = nullptr, *d_volw = nullptr;
    void *d_Pw = nullptr;  // no-slip wall: p_w of the sorted wall particles as a scalar array
    void *d_Vw = nullptr;  // no-slip wall: (cache.wall_velocity, rho_w) of the sorted wall particles (V4<T>)
    bool wall_viscous_done = false;   // this kick's interact! sweep carried the wall's viscous term
    bool wall_velocity_done = false;  // this kick's Adami sweep has written d_Vw already (fused tile version)
    int *d_perm_w = nullptr;
    void *d_scratch = nullptr;  // max(n_f, n_w) * sizeof(double): field unsort
    unsigned long long *d_vmax2 = nullptr, *h_vmax2 = nullptr;  // StateEquationAdaptiveCole: max |v|^2 (bits)
    void *d_adapt = nullptr;       // AdaptConsts<T> + the speed of sound as a double behind it (k_adaptive_consts)
    const void *ad_kick = nullptr; // this kick's kernels read their sound-speed constants from d_adapt
    bool vmax2_ready = false;      // d_vmax2 already holds this kick's maximum (tpb_set_max_speed2: all slabs)
    bool c_on_device = false;      // the current speed of sound lives in d_adapt, fp.sound_speed is stale
    void *d_Ff = nullptr, *d_Fw = nullptr;  // Float32 filter copies of the sorted positions (f32 / f64 coords)
    double filter_ref[3] = {0, 0, 0};       // reference point of the copies (coordinate order)
    float filter_pad = 0;                   // see FilterRef
    TileState tiles;

    tpb_stats stats{};
    unsigned smem_opt_in = 0;  // dynamic shared-memory opt-ins done for this handle (kernel templates are fixed per handle)
    int launches_this_call = 0;
    int deferred_status = TPB_OK;
    bool host_zero_copy = true;  // TPB_MEM_HOST: use mapped page-locked ODE vectors in place (TPB_HOST_ZEROCOPY=0 disables)

    // phase profiling (tpb_set_profiling): one row of TPB_N_PHASES + 1 events per recorded kick
    std::vector<cudaEvent_t> prof_events;
    int prof_capacity = 0, prof_kicks = 0;
};

// event marking the start of phase `ph` of the kick being recorded (ph == TPB_N_PHASES - 1: end)
inline void prof_mark(Semi &s, int ph)
{
    if (s.prof_capacity > 0 && s.prof_kicks < s.prof_capacity)
        cudaEventRecord(s.prof_events[(size_t)s.prof_kicks * TPB_N_PHASES + ph], s.stream);
}

inline size_t tsize(int eltype) { return eltype == TPB_F64 ? 8 : 4; }

// ODE layout (`ranges_u` / `ranges_v`, semidiscretization.jl:128-135): systems in call order; the
// fluid holds ND (u) and NV (v) entries per active particle, the structure ND each per integrated
// particle, the wall none -- except dummy particles with ContinuityDensity, whose density is integrated
// (one v entry per wall particle, wall_boundary/system.jl:78-90).  Offsets / totals in elements.
struct OdeLayout {
    int64_t off_u_f = 0, off_v_f = 0, off_u_s = 0, off_v_s = 0, off_v_w = 0, len_v_w = 0, tot_u = 0, tot_v = 0;
};
inline bool wall_integrates_density(const Semi &s)
{
    return s.wall_index >= 0 && s.wp.density_calculator == TPB_WALL_DENSITY_CONTINUITY;
}
inline OdeLayout ode_layout(const Semi &s)
{
    const int nd = s.cfg.ndims;
    const int nvars = s.fp.density_calculator == TPB_DENSITY_SUMMATION ? nd : nd + 1;
    OdeLayout L;
    L.len_v_w = wall_integrates_density(s) ? s.n_w : 0;
    for (int sys = 0; sys < s.n_systems; ++sys) {
        if (sys == s.fluid_index) {
            L.off_u_f = L.tot_u, L.off_v_f = L.tot_v;
            L.tot_u += s.n_act * nd, L.tot_v += s.n_act * nvars;
        } else if (sys == s.struct_index) {
            L.off_u_s = L.tot_u, L.off_v_s = L.tot_v;
            L.tot_u += s.n_s_int * nd, L.tot_v += s.n_s_int * nd;
        } else if (sys == s.wall_index) {
            L.off_v_w = L.tot_v;
            L.tot_v += L.len_v_w;
        }
    }
    return L;
}

// Device-side alias of a page-locked, mapped host pointer (cudaHostAlloc / cudaHostRegister,
// e.g. through tpb_host_register); nullptr for pageable memory.
inline void *mapped_host_alias(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

static int fail(Semi *s, int code, const std::string &msg)
{
    if (s) s->err = msg; else g_create_error = msg;
    return code;
}

#define CUDA_TRY(S, EXPR)                                                                  \
    do {                                                                                   \
        cudaError_t e_ = (EXPR);                                                           \
        if (e_ != cudaSuccess)                                                             \
            return fail((S), TPB_ERR_CUDA,                                                 \
                        std::string(#EXPR) + ": " + cudaGetErrorString(e_));               \
    } while (0)

#define LAUNCH(S, KERNEL, GRID, BLOCK, SMEM, ...)                                          \
    do {                                                                                   \
        KERNEL<<<(GRID), (BLOCK), (SMEM), (S).stream>>>(__VA_ARGS__);                      \
        (S).launches_this_call++;                                                          \
        (S).stats.kernel_launches_total++;                                                 \
    } while (0)

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------
// host-side constants, computed in T with the reference's operation order
template <typename T>
T eps_of(T x)
{
    x = std::fabs(x);
    return std::nextafter(x, std::numeric_limits<T>::infinity()) - x;
}

template <typename T>
KernelConst<T> make_kernel_const(int kernel, int nd, double h_)
{
    const double pi = 3.14159265358979323846;
    KernelConst<T> k;
    k.h = (T)h_;
    k.h_inv = (T)1 / k.h;
    T sigma;
    k.c_d = (T)0;
    k.order = 0;
    if (kernel == TPB_KERNEL_WENDLAND_C2) {
        sigma = nd == 2 ? (T)(7.0 / (4.0 * pi)) : (T)(21.0 / (16.0 * pi));
    } else if (kernel == TPB_KERNEL_WENDLAND_C4) {
        sigma = nd == 2 ? (T)(9.0 / (pi * 4.0)) : (T)(495.0 / (pi * 256.0));
        k.c_d = (T)(-7.0 / 3.0);
        k.order = 4;
    } else if (kernel == TPB_KERNEL_WENDLAND_C6) {
        sigma = nd == 2 ? (T)(39.0 / (pi * 14.0)) : (T)(1365.0 / (pi * 512.0));
        k.c_d = (T)(-11.0 / 4.0);
        k.order = 6;
    } else if (kernel == TPB_KERNEL_SCHOENBERG_QUARTIC) {
        sigma = nd == 2 ? (T)(96.0 / (pi * 1199.0)) : (T)(1.0 / (pi * 20.0));
        k.order = 4;
    } else if (kernel == TPB_KERNEL_SCHOENBERG_QUINTIC) {
        sigma = nd == 2 ? (T)(7.0 / (pi * 478.0)) : (T)(1.0 / (pi * 120.0));
        k.order = 5;
    } else {
        sigma = nd == 2 ? (T)(10.0 / (pi * 7.0)) : (T)(1.0 / pi);
    }
    k.nf = nd == 2 ? sigma * (k.h_inv * k.h_inv) : sigma * (k.h_inv * k.h_inv * k.h_inv);
    k.m5nf = (T)(-5) * k.nf;
    k.h_inv2 = k.h_inv * k.h_inv;
    // compact_support (smoothing_kernels.jl:215, :310, :383, :400)
    k.support = kernel == TPB_KERNEL_SCHOENBERG_QUARTIC   ? (T)(5.0 / 2.0) * k.h
                : kernel == TPB_KERNEL_SCHOENBERG_QUINTIC ? (T)3 * k.h
                                                          : (T)2 * k.h;
    return k;
}

// params_f32: the state equation's own fields are Float32 (StateEquationAdaptiveCole with its
// default literals): B = reference_density * sound_speed^2 / exponent is then formed in Float32.
template <typename T>
EosConst<T> make_eos_const(double c_, double gamma_, double rho0_, double pbg_, int clip, int params_f32 = 0)
{
    EosConst<T> e;
    T c = (T)c_, gamma = (T)gamma_, rho0 = (T)rho0_;
    e.B = rho0 * (c * c) / gamma;
    if (params_f32) e.B = (T)((float)rho0_ * ((float)c_ * (float)c_) / (float)gamma_);
    e.gamma = gamma;
    e.inv_gamma = (T)1 / gamma;
    e.rho0 = rho0;
    e.p_bg = (T)pbg_;
    e.clip = clip;
    return e;
}

template <typename T>
PairConst<T> make_pair_const(const tpb_fluid_params &fp, int nd)
{
    PairConst<T> k;
    k.kern = make_kernel_const<T>(fp.kernel, nd, fp.smoothing_length);
    T h = k.kern.h;
    k.c = (T)fp.sound_speed;
    k.alpha = (T)fp.alpha;
    k.beta = (T)fp.beta;
    k.eps = (T)fp.epsilon;
    k.eps_h2 = k.eps * (h * h);
    T h_avg = (h + h) / (T)2;
    k.delta_h_c = (T)fp.delta * h_avg * k.c;
    T R = k.kern.support;
    k.radius2 = R * R;
    k.almostzero = std::sqrt(eps_of<T>(R * R));
    k.has_viscosity = fp.has_viscosity;
    k.has_diffusion = fp.has_diffusion;
    return k;
}

template <typename CT>
GridConst<CT> make_grid_const(const Semi &s)
{
    GridConst<CT> g;
    for (int d = 0; d < 3; ++d) {
        g.origin[d] = (CT)s.origin[d];
        g.n[d] = s.ncell[d];
        // outward-rounded bounding box in cT
        CT lo = (CT)s.lo[d], hi = (CT)s.hi[d];
        if ((double)lo > s.lo[d]) lo = std::nextafter(lo, -std::numeric_limits<CT>::infinity());
        if ((double)hi < s.hi[d]) hi = std::nextafter(hi, std::numeric_limits<CT>::infinity());
        g.lo[d] = lo;
        g.hi[d] = hi;
    }
    g.inv_cell = (CT)(1.0 / s.cell_size);
    g.inv_cell_x = (CT)((double)s.xsplit / s.cell_size);
    g.cell = (CT)s.cell_size;
    g.gap_tol = (CT)s.gap_tol;
    g.ncells = (int)s.ncells;
    g.sx = s.xsplit;
    g.ax = s.row_axis;
    for (int d = 0; d < 3; ++d) g.fref.ref[d] = (CT)s.filter_ref[d];
    g.fref.pad = s.filter_pad;
    return g;
}

// ---------------------------------------------------------------------------------------
static int exclusive_scan(Semi &s, const int *d_in, int n, int *d_out)
{
    if (n <= SCAN_SINGLE_MAX) {
        LAUNCH(s, k_scan_single, 1, SCAN_THREADS, 0, d_in, n, d_out);
        return TPB_OK;
    }
    int nblocks = cdiv(n, SCAN_TILE);
    if (nblocks > SCAN_TILE) return fail(&s, TPB_ERR_UNSUPPORTED, "cell grid too large for the scan");
    LAUNCH(s, k_scan_block_sums, nblocks, SCAN_THREADS, 0, d_in, n, s.d_block_sums);
    LAUNCH(s, k_scan_top, 1, SCAN_THREADS, 0, s.d_block_sums, nblocks);
    LAUNCH(s, k_scan_final, nblocks, SCAN_THREADS, 0, d_in, n, s.d_block_sums, d_out);
    return TPB_OK;
}

template <int ND, typename T, typename CT>
struct Ops {
    // threads per target particle in the tile sweeps (tpb_tiles.cuh): the Float32 kernels (with
    // Float32 or Float64 coordinates) fit 2 x 384 threads into the register file; the Float64
    // ones (106 registers) two
    static constexpr int KS = std::is_same<T, float>::value ? TPB_SPLIT : 2;
    static constexpr int nv(const Semi &s) { return s.fp.density_calculator == TPB_DENSITY_SUMMATION ? ND : ND + 1; }

    // ---- counting sort of one point set into the shared grid: key/slot/count/scan/scatter
    // bits of a tmp_perm entry that hold the particle index (the position key sits above them)
    static int perm_bits(const Semi &s, int64_t n)
    {
        if (!s.subkey_mode) return 31;
        if (s.subkey_mode > 1) return std::min(std::max(s.subkey_mode, PERM_IDX_BITS), 27);  // TPB_SUBKEY=26 / 27: test hook
        for (int b = PERM_IDX_BITS; b <= 27; ++b)
            if (n < ((int64_t)1 << b)) return b;
        return 31;
    }

    static int bin_points(Semi &s, const CT *d_coords, int n, int n_targets, int *d_cell_start,
                          const CT *d_tail_coords = nullptr, int n_head = 0, CT *d_out_coords = nullptr)
    {
        GridConst<CT> g = make_grid_const<CT>(s);
        const int pb = perm_bits(s, n);
        CUDA_TRY(&s, cudaMemsetAsync(s.d_count, 0, sizeof(int) * (size_t)s.ncells, s.stream));
        s.count_clean = false;
        if (n > 0)
            LAUNCH(s, (k_cell_count<ND, CT>), cdiv(n, 256), 256, 0, d_coords, n, n_targets, g, s.d_key,
                   s.d_slot, s.d_count, s.d_flags, pb, d_tail_coords, n_head, d_out_coords);
        int rc = exclusive_scan(s, s.d_count, (int)s.ncells, d_cell_start);
        if (rc) return rc;
        if (n > 0)
            LAUNCH(s, k_scatter, cdiv(n, 256), 256, 0, s.d_key, s.d_slot, d_cell_start, n,
                   s.d_tmp_perm, pb);
        return TPB_OK;
    }

    // ---- NHS rebuild of the fluid + EOS (update_nhs!, update_pressure!)
    // wall tiles: empty / active (k_wall_tile_prep); the no-slip records ride along when the Adami sweep
    // carries the wall velocity (launch_adami)
    static bool adami_fused_noslip(const Semi &s)
    {
        constexpr bool SMALL = KS > 1 && sizeof(T) == 4;
        const int list_len = SMALL ? s.tiles.adami_list_len : s.tiles.list(KS);
        return s.wp.has_viscosity && list_len >= ADAMI_NOSLIP_RED * (int)sizeof(T) / 2;
    }
    static WallPrepArgs<T, CT> wall_prep_args(Semi &s)
    {
        const EosConst<T> weos = make_eos_const<T>(s.wp.sound_speed, s.wp.exponent, s.wp.reference_density,
                                                   s.wp.background_pressure, 0,
                                                   s.wp.sound_speed_from_fluid && s.fp.adaptive_params_f32);
        const bool fused = adami_fused_noslip(s);
        WallPrepArgs<T, CT> a;
        a.n_tiles = s.tiles.d_wrow_tile_start + s.tiles.nrows;
        a.tile_desc = s.tiles.d_wtile_desc;
        a.Aw = (const V4<CT> *)s.d_Aw;
        const T p_empty = (T)0;  // empty sum, clipped or not: p = 0
        a.rho_empty = weos.rho0 * (T)std::pow((double)((p_empty - weos.p_bg) / weos.B + (T)1), (double)weos.inv_gamma);
        a.W = (V2<T> *)s.d_Ww;
        a.volume = (T *)s.d_volw;
        a.active = s.tiles.d_wactive;
        a.n_active = s.tiles.d_n_wactive;
        a.rng = s.tiles.d_wtile_rng;
        a.ext = s.tiles.d_wtile_ext;
        a.Vw = fused ? (V4<T> *)s.d_Vw : (V4<T> *)nullptr;
        a.Pw = (T *)s.d_Pw;
        a.state = s.tiles.d_wtile_state;
        a.ad = s.ad_kick && s.wp.sound_speed_from_fluid ? (const AdaptConsts<T> *)s.ad_kick : nullptr;
        if (a.ad) {
            // the value lives on the device: it only moves with the speed of sound when p_background != 0
            a.rewrite = s.wp.background_pressure != 0 || s.tiles.wall_rho_empty != -2;
            s.tiles.wall_rho_empty = -2;
        } else {
            a.rewrite = (double)a.rho_empty != s.tiles.wall_rho_empty;
            s.tiles.wall_rho_empty = (double)a.rho_empty;
        }
        return a;
    }

    // ---- the per-kick rebuild in four launches: histogram -> one-pass scan + tile table -> scatter +
    // candidate ranges + wall-tile sorting -> gather into sorted records (+ EOS)
    static int rebuild_fluid_fused(Semi &s, const CT *d_u, const T *d_v)
    {
        const int n = (int)s.n_act;
        const GridConst<CT> g = make_grid_const<CT>(s);
        if (!s.count_clean) {
            CUDA_TRY(&s, cudaMemsetAsync(s.d_count, 0, sizeof(int) * (size_t)s.ncells, s.stream));
            s.count_clean = true;
        }
        if (n > 0)
            LAUNCH(s, (k_cell_count<ND, CT>), cdiv(n, 256), 256, 0, d_u, n, (int)s.n_tgt, g, s.d_key, s.d_slot,
                   s.d_count, s.d_flags, perm_bits(s, n));
        const bool has_wall = s.n_w > 0;                           // a second neighbour set for the fluid tiles
        const bool wall = has_wall && !wall_integrates_density(s);  // Adami walls: sort the wall tiles as well
        LAUNCH(s, k_scan_cells_tiles, s.scan_blocks, CSCAN_THREADS, 0, s.d_count, s.ncell[0], s.tiles.nrows,
               s.scan_rows_per_block, s.d_scan_ticket, s.d_scan_status, s.d_fcell_start,
               s.tiles.d_frow_tile_start, s.tiles.d_ftile_desc, wall ? s.tiles.d_n_wactive : (int *)nullptr);
        const int nb_scatter = cdiv(n, 256), nb_ranges = cdiv((int64_t)s.tiles.max_ftiles * 32, 256);
        const int nb_wprep = wall ? cdiv((int64_t)s.tiles.max_wtiles * 32, 256) : 0;
        WallPrepArgs<T, CT> wpa{};
        if (wall) wpa = wall_prep_args(s);
        LAUNCH(s, (k_post_scan<ND, T, CT>), nb_scatter + nb_ranges + nb_wprep, 256, 0, nb_scatter, nb_ranges, s.d_key,
               s.d_slot, s.d_fcell_start, n, s.d_tmp_perm, g, s.tiles.d_frow_tile_start + s.tiles.nrows,
               s.tiles.d_ftile_desc, has_wall ? s.d_wcell_start : (const int *)nullptr, s.tiles.d_ftile_rng,
               s.tiles.d_ftile_ext, wpa, perm_bits(s, n));
        s.wall_prep_done = wall;
        return TPB_OK;
    }

    static int rebuild_fluid(Semi &s, const CT *d_u, const T *d_v)
    {
        int n = (int)s.n_act;
        const bool fused = use_tiles(s) && s.scan_blocks > 0;
        s.wall_prep_done = false;
        int rc = fused ? rebuild_fluid_fused(s, d_u, d_v) : bin_points(s, d_u, n, (int)s.n_tgt, s.d_fcell_start);
        if (rc) return rc;
        EosConst<T> eos = make_eos_const<T>(s.fp.sound_speed, s.fp.exponent, s.fp.reference_density,
                                            s.fp.background_pressure, s.fp.clip_negative_pressure,
                                            s.fp.adaptive_sound_speed && s.fp.adaptive_params_f32);
        const GridConst<CT> g = make_grid_const<CT>(s);
        if (n > 0) {
            if (s.fp.density_calculator == TPB_DENSITY_CONTINUITY)
                LAUNCH(s, (k_reorder_fluid<ND, T, CT, 0>), cdiv(n, 256), 256, 0, d_u, d_v,
                       (const T *)s.d_mass_f, s.d_key, s.d_fcell_start, s.d_tmp_perm, n,
                       s.d_fcell_start + s.ncells, s.cfg.deterministic, eos, (V4<CT> *)s.d_A, (V4<T> *)s.d_B, (T *)s.d_P,
                       s.d_perm_f, g.fref, (V4<float> *)s.d_Ff, (const AdaptConsts<T> *)s.ad_kick, perm_bits(s, n));
            else
                LAUNCH(s, (k_reorder_fluid<ND, T, CT, 1>), cdiv(n, 256), 256, 0, d_u, d_v,
                       (const T *)s.d_mass_f, s.d_key, s.d_fcell_start, s.d_tmp_perm, n,
                       s.d_fcell_start + s.ncells, s.cfg.deterministic, eos, (V4<CT> *)s.d_A, (V4<T> *)s.d_B, (T *)s.d_P,
                       s.d_perm_f, g.fref, (V4<float> *)s.d_Ff, (const AdaptConsts<T> *)nullptr, perm_bits(s, n));
        }
        if (use_tiles(s) && !fused) {
            rc = build_tile_table(s, s.d_fcell_start, s.tiles.d_frow_tile_start, s.tiles.d_ftile_desc);
            if (rc) return rc;
            LAUNCH(s, (k_tile_ranges<ND>), cdiv((int64_t)s.tiles.max_ftiles * 32, 256), 256, 0, s.ncell[0],
                   s.ncell[1], s.xsplit, s.tiles.d_frow_tile_start + s.tiles.nrows, s.tiles.d_ftile_desc,
                   s.d_fcell_start, s.d_fcell_start, s.n_w > 0 ? s.d_wcell_start : (const int *)nullptr,
                   s.tiles.d_ftile_rng, 18, s.tiles.d_ftile_ext);
        }
        return TPB_OK;
    }

    // ---- static wall: sorted once (initialize_neighborhood_searches!, neighborhood_search.jl:439-464)
    static int init_wall(Semi &s)
    {
        int n = (int)s.n_w;
        CT *d_coords = nullptr;
        T *d_mass = nullptr, *d_dens = nullptr;
        CUDA_TRY(&s, cudaMalloc(&d_coords, sizeof(CT) * ND * (size_t)std::max(n, 1)));
        CUDA_TRY(&s, cudaMalloc(&d_mass, sizeof(T) * (size_t)std::max(n, 1)));
        CUDA_TRY(&s, cudaMalloc(&d_dens, sizeof(T) * (size_t)std::max(n, 1)));
        CUDA_TRY(&s, cudaMemcpyAsync(d_coords, s.h_coords_w.data(), sizeof(CT) * ND * (size_t)n,
                                     cudaMemcpyHostToDevice, s.stream));
        CUDA_TRY(&s, cudaMemcpyAsync(d_mass, s.h_mass_w.data(), sizeof(T) * (size_t)n,
                                     cudaMemcpyHostToDevice, s.stream));
        CUDA_TRY(&s, cudaMemcpyAsync(d_dens, s.h_dens_w.data(), sizeof(T) * (size_t)n,
                                     cudaMemcpyHostToDevice, s.stream));
        int rc = bin_points(s, d_coords, n, n, s.d_wcell_start);
        if (rc) return rc;
        if (n > 0)
            LAUNCH(s, (k_reorder_wall<ND, T, CT>), cdiv(n, 256), 256, 0, d_coords, d_mass, d_dens,
                   s.d_key, s.d_wcell_start, s.d_tmp_perm, n, (V4<CT> *)s.d_Aw, (V2<T> *)s.d_Ww,
                   s.d_perm_w, make_grid_const<CT>(s).fref, (V4<float> *)s.d_Fw, perm_bits(s, n));
        CUDA_TRY(&s, cudaMemsetAsync(s.d_volw, 0, sizeof(T) * (size_t)std::max(n, 1), s.stream));
        rc = build_tile_table(s, s.d_wcell_start, s.tiles.d_wrow_tile_start, s.tiles.d_wtile_desc, true);
        if (rc) return rc;
        CUDA_TRY(&s, cudaStreamSynchronize(s.stream));
        cudaFree(d_coords);
        cudaFree(d_mass);
        cudaFree(d_dens);
        return TPB_OK;
    }

    // ---- TotalLagrangianSPHSystem -------------------------------------------------------------
    static StructConst<T> make_struct_const(const Semi &s)
    {
        StructConst<T> k;
        k.kern = make_kernel_const<T>(s.sp.kernel, ND, s.sp.smoothing_length);
        const T h = k.kern.h;
        k.almostzero = std::sqrt(eps_of<T>(h * h));
        const T E = (T)s.sp.young_modulus, nu = (T)s.sp.poisson_ratio;
        // system.jl:158-161
        k.lambda = E * nu / (((T)1 + nu) * ((T)1 - (T)2 * nu));
        k.mu = (E / (T)2) / ((T)1 + nu);
        k.young = E;
        k.half_alpha = (T)s.sp.penalty_alpha / (T)2;
        k.has_penalty = s.sp.has_penalty_force;
        for (int d = 0; d < 3; ++d) k.acc[d] = (T)s.sp.acceleration[d];
        return k;
    }
    static MKConst<T> make_mk_const(const Semi &s, const PairConst<T> &pc)
    {
        MKConst<T> k;
        const T K = (T)s.sp.mk_K, beta = (T)s.sp.mk_beta, spacing = (T)s.sp.mk_spacing;
        k.K_bpow = K / (ND == 2 ? beta : beta * beta);
        k.spacing = spacing;
        k.min_dfs = spacing / (T)100;
        k.h_fluid = pc.kern.h;
        k.vol = ND == 2 ? spacing * spacing : spacing * spacing * spacing;
        k.radius2 = pc.radius2;
        k.almostzero_fs = pc.almostzero;
        k.almostzero_sf = std::sqrt(eps_of<T>(pc.kern.h * pc.kern.h));
        return k;
    }
    static DummyConst<T> make_dummy_const(const Semi &s, const PairConst<T> &pc)
    {
        DummyConst<T> k;
        k.kern = make_kernel_const<T>(s.sp.bm_kernel, ND, s.sp.bm_smoothing_length);
        k.eos = make_eos_const<T>(s.sp.bm_sound_speed, s.sp.bm_exponent, s.sp.bm_reference_density,
                                  s.sp.bm_background_pressure, 0);
        const T Rb = k.kern.support;
        k.radius2_b = Rb * Rb;
        k.radius2_f = pc.radius2;
        for (int d = 0; d < 3; ++d) k.acc[d] = (T)s.fp.acceleration[d];
        k.p_off = (T)s.sp.bm_pressure_offset;
        k.clip = s.sp.bm_clip_negative_pressure;
        k.almostzero_fs = pc.almostzero;
        k.almostzero_sf = std::sqrt(eps_of<T>(pc.kern.h * pc.kern.h));
        // BernoulliPressureExtrapolation (dummy_particles.jl:674-707): a wall has the term only while it moves
        k.bernoulli = s.sp.bm_bernoulli_factor != 0.0 && (!s.sp.bm_wall_semantics || s.clamped_moving)
                          ? (T)s.sp.bm_bernoulli_factor : (T)0;
        return k;
    }
    static int kernel_template_id(int kernel) { return kernel <= 1 ? kernel : kernel <= TPB_KERNEL_WENDLAND_C6 ? 2 : 3; }
    static int struct_kernel_id(const Semi &s)
    {
        const int kernel = s.sp.kernel;
        return kernel <= 1 ? kernel : kernel <= TPB_KERNEL_WENDLAND_C6 ? 2 : 3;
    }

    // once: neighbour list over the initial configuration (exact predicate, ascending neighbour
    // index) + gradient correction matrices (initialize!, system.jl:391-401)
    static int init_structure(Semi &s)
    {
        const int n = (int)s.n_s;
        const CT *x0 = (const CT *)s.h_x0_s.data();
        const T R = make_kernel_const<T>(s.sp.kernel, ND, s.sp.smoothing_length).support;
        const T R2 = R * R;
        std::vector<int> start((size_t)n + 1, 0), nbr;
        {
            // cells of size R over the structure's own bounding box
            double lo[3] = {0, 0, 0}, cell = (double)R * (1.0 + 1e-6);
            for (int d = 0; d < ND; ++d) {
                lo[d] = (double)x0[d];
                for (int i = 1; i < n; ++i) lo[d] = std::min(lo[d], (double)x0[(size_t)i * ND + d]);
            }
            auto cell_of = [&](int i, int d) { return (int64_t)std::floor(((double)x0[(size_t)i * ND + d] - lo[d]) / cell); };
            std::vector<std::pair<std::array<int64_t, 3>, int>> keyed((size_t)n);
            for (int i = 0; i < n; ++i) {
                std::array<int64_t, 3> c = {0, 0, 0};
                for (int d = 0; d < ND; ++d) c[d] = cell_of(i, d);
                keyed[(size_t)i] = {c, i};
            }
            std::sort(keyed.begin(), keyed.end());
            auto lower = [&](const std::array<int64_t, 3> &c) {
                return std::lower_bound(keyed.begin(), keyed.end(), std::make_pair(c, -1)) - keyed.begin();
            };
            std::vector<int> cand;
            for (int a = 0; a < n; ++a) {
                cand.clear();
                std::array<int64_t, 3> ca = {0, 0, 0};
                for (int d = 0; d < ND; ++d) ca[d] = cell_of(a, d);
                for (int64_t dz = (ND == 3 ? -1 : 0); dz <= (ND == 3 ? 1 : 0); ++dz)
                    for (int64_t dy = -1; dy <= 1; ++dy)
                        for (int64_t dx = -1; dx <= 1; ++dx) {
                            const std::array<int64_t, 3> c = {ca[0] + dx, ca[1] + dy, ca[2] + dz};
                            for (size_t q = (size_t)lower(c); q < keyed.size() && keyed[q].first == c; ++q) {
                                const int b = keyed[q].second;
                                // pos_diff = convert.(T, x_a - x_b); d2 = dot (left to right); d2 <= R^2
                                T d2 = 0;
                                for (int d = 0; d < ND; ++d) {
                                    const T pd = (T)(x0[(size_t)a * ND + d] - x0[(size_t)b * ND + d]);
                                    const T sq = pd * pd;
                                    d2 = d == 0 ? sq : d2 + sq;
                                }
                                if (d2 <= R2) cand.push_back(b);
                            }
                        }
                std::sort(cand.begin(), cand.end());
                start[(size_t)a + 1] = start[(size_t)a] + (int)cand.size();
                nbr.insert(nbr.end(), cand.begin(), cand.end());
            }
        }
        const size_t nn = std::max<size_t>(nbr.size(), 1);
        CUDA_TRY(&s, cudaMalloc(&s.d_nbr_start, sizeof(int) * ((size_t)n + 1)));
        CUDA_TRY(&s, cudaMalloc(&s.d_nbr, sizeof(int) * nn));
        CUDA_TRY(&s, cudaMemcpy(s.d_nbr_start, start.data(), sizeof(int) * ((size_t)n + 1), cudaMemcpyHostToDevice));
        if (!nbr.empty())
            CUDA_TRY(&s, cudaMemcpy(s.d_nbr, nbr.data(), sizeof(int) * nbr.size(), cudaMemcpyHostToDevice));
        const size_t nz = (size_t)std::max(n, 1);
        CUDA_TRY(&s, cudaMemcpy(s.d_x0_s, s.h_x0_s.data(), sizeof(CT) * ND * (size_t)n, cudaMemcpyHostToDevice));
        CUDA_TRY(&s, cudaMemcpy(s.d_xcur_s, s.h_x0_s.data(), sizeof(CT) * ND * (size_t)n, cudaMemcpyHostToDevice));
        CUDA_TRY(&s, cudaMemcpy(s.d_xcl_s, s.h_x0_s.data(), sizeof(CT) * ND * (size_t)n, cudaMemcpyHostToDevice));
        CUDA_TRY(&s, cudaMemcpy(s.d_mass_s, s.h_mass_s.data(), sizeof(T) * (size_t)n, cudaMemcpyHostToDevice));
        CUDA_TRY(&s, cudaMemcpy(s.d_rho_s, s.h_rho_s.data(), sizeof(T) * (size_t)n, cudaMemcpyHostToDevice));
        if (!s.h_hydro_s.empty())
            CUDA_TRY(&s, cudaMemcpy(s.d_hydro_s, s.h_hydro_s.data(), sizeof(T) * (size_t)n, cudaMemcpyHostToDevice));
        else
            CUDA_TRY(&s, cudaMemset(s.d_hydro_s, 0, sizeof(T) * nz));
        if (n > 0) {
            const StructConst<T> k = make_struct_const(s);
            const T eps_h2 = eps_of<T>(k.kern.h * k.kern.h);
            switch (struct_kernel_id(s)) {
#define TPB_CORR(KID)                                                                                              \
    case KID:                                                                                                      \
        LAUNCH(s, (k_struct_correction_matrix<ND, T, CT, KID>), cdiv(n, 128), 128, 0, n, k, eps_h2, s.d_nbr_start,  \
               s.d_nbr, (const CT *)s.d_x0_s, (const T *)s.d_mass_s, (const T *)s.d_rho_s, (T *)s.d_L_s);         \
        break;
                TPB_CORR(0) TPB_CORR(1) TPB_CORR(2) TPB_CORR(3)
#undef TPB_CORR
            }
            CUDA_TRY(&s, cudaStreamSynchronize(s.stream));
        }
        return TPB_OK;
    }

    // per kick, after the fluid rebuild: current coordinates, the structure binned into the shared
    // grid (neighbour of the fluid), deformation gradient + PK1
    static int update_structure(Semi &s, const GridConst<CT> &g, const PairConst<T> &pc, const CT *d_u_s,
                                const T *d_v_s)
    {
        const int n = (int)s.n_s, n_int = (int)s.n_s_int;
        if (n == 0) return TPB_OK;
        if (s.sp.boundary_model == TPB_BOUNDARY_NONE)
            LAUNCH(s, (k_struct_positions<ND, CT>), cdiv((int64_t)n * ND, 256), 256, 0, n, n_int, d_u_s,
                   (const CT *)s.d_xcl_s, (CT *)s.d_xcur_s);
        if (s.sp.boundary_model != TPB_BOUNDARY_NONE) {
            // current coordinates (integrated particles from u_ode, clamped ones from d_xcl_s) written by the
            // cell count itself
            int rc = bin_points(s, d_u_s, n, n, s.d_scell_start, (const CT *)s.d_xcl_s, n_int, (CT *)s.d_xcur_s);
            if (rc) return rc;
            if (s.sp.boundary_model == TPB_BOUNDARY_DUMMY_PARTICLES) {
                // dummy particles: sorted records first, then the Adami pass over the fluid's sorted records
                // (rebuild_fluid has run: positions, density and pressure of the fluid are in place)
                LAUNCH(s, (k_reorder_struct<ND, T, CT>), cdiv(n, 256), 256, 0, (const CT *)s.d_xcur_s, d_v_s,
                       (const T *)s.d_hydro_s, s.d_key, s.d_scell_start, s.d_tmp_perm, n, n_int, (T)1,
                       (V4<CT> *)s.d_As, (V4<T> *)s.d_Bs, s.d_sperm,
                       s.clamped_moving ? (const T *)s.d_vcl_s : (const T *)nullptr, perm_bits(s, n));
                const DummyConst<T> dk = make_dummy_const(s, pc);
                const int enabled = s.struct_fluid[0] && s.n_act > 0;
                switch (kernel_template_id(s.sp.bm_kernel)) {
#define TPB_SADAMI(KID)                                                                                           \
    case KID:                                                                                                     \
        LAUNCH(s, (k_struct_adami<ND, T, CT, KID>), cdiv(n, 128), 128, 0, n, g, (const V4<CT> *)s.d_As, s.d_sperm, \
               s.d_fcell_start, (const V4<CT> *)s.d_A, (const V4<T> *)s.d_B, (const T *)s.d_P, enabled, dk,        \
               (V4<T> *)s.d_Bs, (T *)s.d_Ps, (T *)s.d_p_s, (T *)s.d_rhoh_s, s.d_flags,                             \
               s.clamped_moving && s.sp.bm_wall_semantics ? (const T *)s.d_acl_s : (const T *)nullptr, n_int);     \
        break;
                    TPB_SADAMI(0) TPB_SADAMI(1) TPB_SADAMI(2) TPB_SADAMI(3)
#undef TPB_SADAMI
                }
            } else {
                const MKConst<T> mk = make_mk_const(s, pc);
                LAUNCH(s, (k_reorder_struct<ND, T, CT>), cdiv(n, 256), 256, 0, (const CT *)s.d_xcur_s, d_v_s,
                       (const T *)s.d_hydro_s, s.d_key, s.d_scell_start, s.d_tmp_perm, n, n_int, mk.vol,
                       (V4<CT> *)s.d_As, (V4<T> *)s.d_Bs, (int *)nullptr,
                       s.clamped_moving ? (const T *)s.d_vcl_s : (const T *)nullptr, perm_bits(s, n));
            }
        }
        return structure_deformation(s);
    }

    // deformation gradient + PK1 of every structure particle from the current positions
    static int structure_deformation(Semi &s)
    {
        const int n = (int)s.n_s;
        const StructConst<T> k = make_struct_const(s);
        switch (struct_kernel_id(s)) {
#define TPB_DEFGRAD(KID)                                                                                          \
    case KID:                                                                                                     \
        LAUNCH(s, (k_struct_defgrad_pk1<ND, T, CT, KID>), cdiv(n, 128), 128, 0, n, k, s.d_nbr_start, s.d_nbr,      \
               (const CT *)s.d_x0_s, (const CT *)s.d_xcur_s, (const T *)s.d_mass_s, (const T *)s.d_rho_s,         \
               (const T *)s.d_L_s, (T *)s.d_F_s, (T *)s.d_pk1_s, (const T *)s.d_mat_s);                           \
        break;
            TPB_DEFGRAD(0) TPB_DEFGRAD(1) TPB_DEFGRAD(2) TPB_DEFGRAD(3)
#undef TPB_DEFGRAD
        }
        return TPB_OK;
    }

    // per kick, after interact!: fluid <- structure, structure <- fluid, structure <- structure + gravity
    // mode 0: everything; 1: the structure is not integrated by this kick (split integration,
    // apply_system_interaction!, semidiscretization.jl:868-880): fluid <- structure only, dv_s = 0;
    // 2: only structure <- fluid into d_dv_s (other_interaction_split!, split_integration.jl:444-470)
    template <int FK, int DENS>
    static int interact_structure(Semi &s, const GridConst<CT> &g, const PairConst<T> &pc, T *d_dv_f, T *d_dv_s,
                                  int mode = 0)
    {
        constexpr int NV = DENS == 0 ? ND + 1 : ND;
        const int n = (int)s.n_s, n_int = (int)s.n_s_int;
        if (n == 0) return TPB_OK;
        const bool fluid_side = mode != 2, structure_side = mode != 1;
        if (!structure_side && n_int > 0)
            CUDA_TRY(&s, cudaMemsetAsync(d_dv_s, 0, sizeof(T) * ND * (size_t)n_int, s.stream));
        if (s.sp.boundary_model == TPB_BOUNDARY_DUMMY_PARTICLES) {
            const DummyConst<T> dk = make_dummy_const(s, pc);
            if (fluid_side && s.struct_fluid[1] && s.n_act > 0)
                LAUNCH(s, (k_fluid_from_struct_dummy<ND, T, CT, FK, DENS>), cdiv(s.n_act, 128), 128, 0, (int)s.n_act, g,
                       s.d_fcell_start, (const V4<CT> *)s.d_A, (const V4<T> *)s.d_B, (const T *)s.d_P, s.d_perm_f,
                       s.d_scell_start, (const V4<CT> *)s.d_As, (const V4<T> *)s.d_Bs, (const T *)s.d_Ps, pc.kern, dk,
                       d_dv_f, (int)s.n_tgt);
            if (n_int == 0 || !structure_side) return TPB_OK;
            LAUNCH(s, (k_struct_from_fluid_dummy<ND, T, CT, FK, DENS>), cdiv(n_int, 128), 128, 0, n_int, g,
                   (const CT *)s.d_xcur_s, (const T *)s.d_mass_s, (const T *)s.d_hydro_s, (const T *)s.d_p_s,
                   (const T *)s.d_rhoh_s, s.d_fcell_start, (const V4<CT> *)s.d_A, (const V4<T> *)s.d_B,
                   (const T *)s.d_P, (int)(s.struct_fluid[0] && s.n_act > 0), pc.kern, dk, d_dv_s, s.d_flags);
            return mode == 2 ? TPB_OK : interact_structure_self(s, d_dv_s);
        }
        const bool coupled = s.sp.boundary_model == TPB_BOUNDARY_MONAGHAN_KAJTAR;
        const MKConst<T> mk = make_mk_const(s, pc);
        if (fluid_side && coupled && s.struct_fluid[1] && s.n_act > 0)
            LAUNCH(s, (k_fluid_from_struct<ND, T, CT, FK, NV>), cdiv(s.n_act, 128), 128, 0, (int)s.n_act, g,
                   s.d_fcell_start, (const V4<CT> *)s.d_A, (const V4<T> *)s.d_B, s.d_perm_f, s.d_scell_start,
                   (const V4<CT> *)s.d_As, (const V4<T> *)s.d_Bs, pc.kern, mk, d_dv_f, (int)s.n_tgt);
        if (n_int == 0 || !structure_side) return TPB_OK;
        if (coupled && s.struct_fluid[0] && s.n_act > 0)
            LAUNCH(s, (k_struct_from_fluid<ND, T, CT>), cdiv(n_int, 128), 128, 0, n_int, g, (const CT *)s.d_xcur_s,
                   (const T *)s.d_mass_s, s.d_fcell_start, (const V4<CT> *)s.d_A, 1, mk, d_dv_s, s.d_flags);
        else  // nothing to feel (no boundary model, no fluid particles): the structure need not stay inside the grid
            CUDA_TRY(&s, cudaMemsetAsync(d_dv_s, 0, sizeof(T) * ND * (size_t)n_int, s.stream));
        return mode == 2 ? TPB_OK : interact_structure_self(s, d_dv_s);
    }

    // ---- SplitIntegrationCallback (callbacks/split_integration.jl): the force of the fluid on the structure for
    // the state (v_ode, u_ode), kept constant during one sub-integration (other_interaction_split!, :444-470,
    // after update_systems_and_nhs, :232-234): rebuild, structure kinematics + binning + boundary model, then
    // structure <- fluid alone into `d_out` (ND x n_integrated).  Device vectors.
    static int structure_fluid_force(Semi &s, void *out, const void *v_ode, const void *u_ode)
    {
        if (s.struct_index < 0) return fail(&s, TPB_ERR_STATE, "no structure system");
        const OdeLayout lay = ode_layout(s);
        const T *d_v_ode = (const T *)v_ode;
        const CT *d_u_ode = (const CT *)u_ode;
        s.ad_kick = nullptr;
        int rc = rebuild_fluid(s, d_u_ode + lay.off_u_f, d_v_ode + lay.off_v_f);
        if (rc) return rc;
        const GridConst<CT> g = make_grid_const<CT>(s);
        const PairConst<T> pc = make_pair_const<T>(s.fp, ND);
        rc = update_structure(s, g, pc, d_u_ode + lay.off_u_s, d_v_ode + lay.off_v_s);
        if (rc) return rc;
        auto tk = [](int kernel) { return kernel <= 1 ? kernel : kernel <= TPB_KERNEL_WENDLAND_C6 ? 2 : 3; };
        const int fk = tk(s.fp.kernel);
        const bool summ = s.fp.density_calculator == TPB_DENSITY_SUMMATION;
        if (summ) return fail(&s, TPB_ERR_UNSUPPORTED, "split integration: ContinuityDensity fluid only");
        T *d_out = (T *)out;
        if (fk == 0) rc = interact_structure<0, 0>(s, g, pc, nullptr, d_out, 2);
        else if (fk == 1) rc = interact_structure<1, 0>(s, g, pc, nullptr, d_out, 2);
        else if (fk == 2) rc = interact_structure<2, 0>(s, g, pc, nullptr, d_out, 2);
        else rc = interact_structure<3, 0>(s, g, pc, nullptr, d_out, 2);
        if (rc) return rc;
        CUDA_TRY(&s, cudaGetLastError());
        return TPB_OK;
    }

    // kick_split! (split_integration.jl:352-371): the structure alone -- positions from u_split, deformation
    // gradient and PK1, structure <- structure, + the constant force of the other systems, + gravity
    static int kick_structure(Semi &s, void *dv_split, const void *v_split, const void *u_split, const void *dv_const)
    {
        (void)v_split;  // (no velocity-dependent term: penalty force and stress depend on positions only)
        if (s.struct_index < 0) return fail(&s, TPB_ERR_STATE, "no structure system");
        const int n = (int)s.n_s, n_int = (int)s.n_s_int;
        if (n_int == 0) return TPB_OK;
        LAUNCH(s, (k_struct_positions<ND, CT>), cdiv((int64_t)n * ND, 256), 256, 0, n, n_int, (const CT *)u_split,
               (const CT *)s.d_xcl_s, (CT *)s.d_xcur_s);
        int rc = structure_deformation(s);
        if (rc) return rc;
        if (dv_const)
            CUDA_TRY(&s, cudaMemcpyAsync(dv_split, dv_const, sizeof(T) * ND * (size_t)n_int, cudaMemcpyDeviceToDevice, s.stream));
        else
            CUDA_TRY(&s, cudaMemsetAsync(dv_split, 0, sizeof(T) * ND * (size_t)n_int, s.stream));
        rc = interact_structure_self(s, (T *)dv_split);
        if (rc) return rc;
        CUDA_TRY(&s, cudaGetLastError());
        return TPB_OK;
    }

    // structure <- structure + gravity (adds to what the fluid has left in dv_s)
    static int interact_structure_self(Semi &s, T *d_dv_s)
    {
        const int n_int = (int)s.n_s_int;
        StructConst<T> k = make_struct_const(s);
        switch (struct_kernel_id(s)) {
#define TPB_SINTERACT(KID)                                                                                       \
    case KID:                                                                                                    \
        LAUNCH(s, (k_struct_interact<ND, T, CT, KID>), cdiv(n_int, 128), 128, 0, n_int, k,                        \
               s.d_nbr_start, s.d_nbr,                                      (const CT *)s.d_x0_s, (const CT *)s.d_xcur_s,   \
               (const T *)s.d_mass_s, (const T *)s.d_rho_s, (const T *)s.d_F_s, (const T *)s.d_pk1_s, d_dv_s,    \
               (const T *)s.d_mat_s);                                                                            \
        break;
            TPB_SINTERACT(0) TPB_SINTERACT(1) TPB_SINTERACT(2) TPB_SINTERACT(3)
#undef TPB_SINTERACT
        }
        return TPB_OK;
    }

    template <int KERNEL>
    static void launch_summation(Semi &s, const GridConst<CT> &g, const PairConst<T> &pc,
                                 const EosConst<T> &eos)
    {
        int n = (int)s.n_act;
        if (use_tiles(s) && !getenv("TPB_SUMMATION_PP")) {
            // the tile sweep of interact! with a W-only body (the fluid's tile table is ready)
            const int list_len = s.tiles.list(KS);
            const int cap = tile_capacity<T, CT>(s.tiles.smem_budget, list_len, KS);
            const size_t smem = tile_smem_bytes<T, CT>(cap, list_len, KS);
            if (!(s.smem_opt_in & 4)) {
                if (set_smem(s, k_summation_tiles<KS, ND, T, CT, KERNEL>, 227 * 1024)) return;
                s.smem_opt_in |= 4;
            }
            LAUNCH(s, (k_summation_tiles<KS, ND, T, CT, KERNEL>), s.tiles.max_ftiles, KS * TILE_TB, smem, g,
                   s.tiles.d_frow_tile_start + s.tiles.nrows, s.tiles.d_ftile_desc, s.tiles.d_ftile_ext,
                   s.tiles.d_ftile_rng, s.d_fcell_start, (const V4<CT> *)s.d_A, s.d_perm_f,
                   (int)(s.n_w > 0 && s.interaction[0][1]), s.d_wcell_start, (const V4<CT> *)s.d_Aw, s.d_perm_w,
                   pc.kern, pc.radius2, eos, (V4<T> *)s.d_B, (T *)s.d_P, cap, list_len,
                   (const V4<float> *)s.d_Ff, (const V4<float> *)s.d_Fw);
            return;
        }
        LAUNCH(s, (k_summation_density<ND, T, CT, KERNEL>), cdiv(n, 128), 128, 0, n, g,
               s.d_fcell_start, (const V4<CT> *)s.d_A, (int)(s.n_w > 0), s.d_wcell_start,
               (const V4<CT> *)s.d_Aw, s.interaction[0][1], pc.kern, pc.radius2, eos,
               (V4<T> *)s.d_B, (T *)s.d_P);
    }

    // ---- tile table of one sorted point set: tiles per cell row + exclusive scan
    static int build_tile_table(Semi &s, const int *d_cell_start, int *d_row_tile_start, int4 *d_desc,
                                bool cut_at_gaps = false)
    {
        const int nrows = s.tiles.nrows;
        if (cut_at_gaps) {
            const int gap = 2 * s.xsplit;
            LAUNCH(s, k_row_tiles_gaps<false>, cdiv(nrows, 128), 128, 0, d_cell_start, s.ncell[0], nrows, gap,
                   s.tiles.d_row_tiles, (const int *)nullptr, (int4 *)nullptr);
            int rc = exclusive_scan(s, s.tiles.d_row_tiles, nrows, d_row_tile_start);
            if (rc) return rc;
            // static point set, built once: size the tile arrays to the exact count
            int n_tiles = 0;
            CUDA_TRY(&s, cudaMemcpyAsync(&n_tiles, d_row_tile_start + nrows, sizeof(int), cudaMemcpyDeviceToHost,
                                         s.stream));
            CUDA_TRY(&s, cudaStreamSynchronize(s.stream));
            if (tiles_reserve_wall(s.tiles, n_tiles)) return fail(&s, TPB_ERR_CUDA, "out of device memory (wall tiles)");
            d_desc = s.tiles.d_wtile_desc;
            LAUNCH(s, k_row_tiles_gaps<true>, cdiv(nrows, 128), 128, 0, d_cell_start, s.ncell[0], nrows, gap,
                   (int *)nullptr, (const int *)d_row_tile_start, d_desc);
            return TPB_OK;
        }
        if (nrows <= TILE_TABLE_MAX_ROWS) {
            LAUNCH(s, k_row_tile_table, 1, SCAN_THREADS, 0, d_cell_start, s.ncell[0], nrows, d_row_tile_start,
                   d_desc);
            return TPB_OK;
        }
        LAUNCH(s, k_row_tiles, cdiv(nrows, 256), 256, 0, d_cell_start, s.ncell[0], nrows,
               s.tiles.d_row_tiles);
        int rc = exclusive_scan(s, s.tiles.d_row_tiles, nrows, d_row_tile_start);
        if (rc) return rc;
        LAUNCH(s, k_fill_tiles, cdiv(nrows, 256), 256, 0, d_cell_start, s.ncell[0], nrows,
               d_row_tile_start, d_desc);
        return TPB_OK;
    }

    static int use_tiles(Semi &s)
    {
        int variant = s.cfg.interact_variant;
        return variant == 0 || variant == 2;
    }

    template <typename K>
    static int set_smem(Semi &s, K kernel, size_t bytes)
    {
        CUDA_TRY(&s, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        return TPB_OK;
    }

    template <int KERNEL>
    static int launch_adami(Semi &s, const GridConst<CT> &g)
    {
        int n = (int)s.n_w;
        AdamiConst<T> k;
        k.kern = make_kernel_const<T>(s.wp.kernel, ND, s.wp.smoothing_length);
        k.eos = make_eos_const<T>(s.wp.sound_speed, s.wp.exponent, s.wp.reference_density,
                                  s.wp.background_pressure, 0,
                                  s.wp.sound_speed_from_fluid && s.fp.adaptive_params_f32);
        T R = k.kern.support;
        k.radius2 = R * R;
        for (int d = 0; d < 3; ++d) k.acc[d] = (T)s.fp.acceleration[d];
        k.p_off = (T)s.wp.pressure_offset;
        k.clip = s.wp.clip_negative_pressure;
        if (use_tiles(s)) {
            constexpr bool SMALL = KS > 1 && sizeof(T) == 4;  // three blocks per SM (see k_adami_tiles)
            const int list_len = SMALL ? s.tiles.adami_list_len : s.tiles.list(KS);
            const int budget = SMALL ? std::min(s.tiles.adami_smem_budget, s.tiles.smem_budget) : s.tiles.smem_budget;
            const int cap = tile_capacity<T, CT>(budget, list_len, KS);
            const size_t smem = tile_smem_bytes<T, CT>(cap, list_len, KS);
            int grid = s.tiles.max_wtiles;  // one block per tile slot; surplus blocks exit at once
            if (const char *e = getenv("TPB_ADAMI_GRID"))  // tuning: fewer blocks, each walking several tiles
                if (atoi(e) > 0) grid = std::min(grid, atoi(e));
            // the opt-in is per device (a handle on another device of the same process needs its own)
            if (!(s.smem_opt_in & 1)) {
                int rc = set_smem(s, k_adami_tiles<KS, ND, T, CT, KERNEL>, 227 * 1024);
                if (rc) return rc;
                rc = set_smem(s, k_adami_tiles<KS, ND, T, CT, KERNEL, true>, 227 * 1024);
                if (rc) return rc;
                s.smem_opt_in |= 1;
            }
            // no-slip wall: the wall velocity rides along in the same sweep when the lists leave room
            // for the five reduction slots; otherwise k_wall_velocity follows (kick_device)
            const bool fused = adami_fused_noslip(s);
            s.wall_velocity_done = fused;
            const AdaptConsts<T> *wall_ad =
                s.ad_kick && s.wp.sound_speed_from_fluid ? (const AdaptConsts<T> *)s.ad_kick : nullptr;
            if (!s.wall_prep_done) {  // (the fused rebuild launch has done it already)
                CUDA_TRY(&s, cudaMemsetAsync(s.tiles.d_n_wactive, 0, sizeof(int), s.stream));
                const WallPrepArgs<T, CT> a = wall_prep_args(s);
                LAUNCH(s, (k_wall_tile_prep<ND, T, CT>), cdiv((int64_t)s.tiles.max_wtiles * 32, 256), 256, 0, g,
                       a.n_tiles, a.tile_desc, a.Aw, s.d_fcell_start, a.rho_empty, a.W, a.volume, a.active,
                       a.n_active, a.rng, a.ext, a.Vw, a.Pw, a.state, a.rewrite);
            }
            if (fused) {
                LAUNCH(s, (k_adami_tiles<KS, ND, T, CT, KERNEL, true>), grid, KS * TILE_TB, smem, g,
                       s.tiles.d_n_wactive, s.tiles.d_wactive, s.tiles.d_wtile_desc, s.tiles.d_wtile_ext,
                       s.tiles.d_wtile_rng, s.d_wcell_start, (const V4<CT> *)s.d_Aw,
                       s.d_fcell_start, (const V4<CT> *)s.d_A, (const V4<T> *)s.d_B, (const T *)s.d_P,
                       s.interaction[1][0], k, (V2<T> *)s.d_Ww, (T *)s.d_volw, cap, list_len,
                       (const V4<float> *)s.d_Ff, (V4<T> *)s.d_Vw, (T *)s.d_Pw, wall_ad);
                return TPB_OK;
            }
            LAUNCH(s, (k_adami_tiles<KS, ND, T, CT, KERNEL>), grid, KS * TILE_TB, smem, g,
                   s.tiles.d_n_wactive, s.tiles.d_wactive, s.tiles.d_wtile_desc, s.tiles.d_wtile_ext,
                   s.tiles.d_wtile_rng, s.d_wcell_start, (const V4<CT> *)s.d_Aw,
                   s.d_fcell_start, (const V4<CT> *)s.d_A, (const V4<T> *)s.d_B, (const T *)s.d_P,
                   s.interaction[1][0], k, (V2<T> *)s.d_Ww, (T *)s.d_volw, cap, list_len,
                   (const V4<float> *)s.d_Ff, (V4<T> *)nullptr, (T *)nullptr, wall_ad);
            return TPB_OK;
        }
        s.wall_velocity_done = false;
        LAUNCH(s, (k_adami<ND, T, CT, KERNEL>), cdiv(n, 128), 128, 0, n, g,
               (const V4<CT> *)s.d_Aw, s.d_fcell_start, (const V4<CT> *)s.d_A,
               (const V4<T> *)s.d_B, (const T *)s.d_P, s.interaction[1][0], k,
               (V2<T> *)s.d_Ww, (T *)s.d_volw);
        return TPB_OK;
    }

    // ---- no-slip wall (boundary_model.viscosity !== nothing): wall velocity after the Adami pass
    template <int KERNEL>
    static int launch_wall_velocity(Semi &s, const GridConst<CT> &g)
    {
        int n = (int)s.n_w;
        KernelConst<T> kern = make_kernel_const<T>(s.wp.kernel, ND, s.wp.smoothing_length);
        T R = kern.support;
        LAUNCH(s, (k_wall_velocity<ND, T, CT, KERNEL>), cdiv(n, 128), 128, 0, n, g, (const V4<CT> *)s.d_Aw,
               s.d_fcell_start, (const V4<CT> *)s.d_A, (const V4<T> *)s.d_B, s.interaction[1][0], kern,
               (T)(R * R), (const V2<T> *)s.d_Ww, (V4<T> *)s.d_Vw, (T *)s.d_Pw);
        return TPB_OK;
    }

    // ---- ... and the wall model's viscous term of the fluid after interact!
    static WallViscConst<T> make_wall_visc_const(const Semi &s, const PairConst<T> &pc)
    {
        WallViscConst<T> k;
        k.kern = pc.kern;
        k.model = s.wp.has_viscosity;
        k.alpha = (T)s.wp.alpha;
        k.beta = (T)s.wp.beta;
        const T h_f = pc.kern.h, h_w = (T)s.wp.smoothing_length;
        // kinematic_viscosity (viscosity.jl:82-87, :154-157, :282-285)
        auto kin = [&](int model, T alpha_or_nu, T h) {
            return model == TPB_VISCOSITY_MONAGHAN ? alpha_or_nu * h * pc.c / (T)(2 * ND + 4) : alpha_or_nu;
        };
        k.nu_a = kin(s.fp.has_viscosity, (T)s.fp.alpha, h_f);
        k.nu_b = kin(s.wp.has_viscosity, (T)s.wp.alpha, h_w);
        k.h = (h_f + h_w) / (T)2;
        k.eps_h2 = (T)s.wp.epsilon * (k.h * k.h);
        k.c = pc.c;
        k.radius2 = pc.radius2;
        k.almostzero = pc.almostzero;
        return k;
    }

    // (per-particle variant; the tile sweep has the term inside k_interact_tiles<NOSLIP>)
    template <int KERNEL, int DENS>
    static int launch_wall_viscous(Semi &s, const GridConst<CT> &g, const PairConst<T> &pc, T *d_dv)
    {
        constexpr int NV = DENS == 0 ? ND + 1 : ND;
        int n = (int)s.n_act;
        const WallViscConst<T> k = make_wall_visc_const(s, pc);
        LAUNCH(s, (k_wall_viscous<ND, T, CT, KERNEL, NV>), cdiv(n, 128), 128, 0, n, g, s.d_fcell_start,
               (const V4<CT> *)s.d_A, (const V4<T> *)s.d_B, s.d_perm_f, s.d_wcell_start,
               (const V4<CT> *)s.d_Aw, (const V2<T> *)s.d_Ww, (const V4<T> *)s.d_Vw, k, d_dv, (int)s.n_tgt);
        return TPB_OK;
    }

    template <int KERNEL, int DENS>
    static int launch_interact(Semi &s, const GridConst<CT> &g, const PairConst<T> &pc, T *d_dv)
    {
        int n = (int)s.n_act;
        SourceConst<T> src;
        src.any = s.fp.damping_coefficient != 0.0;
        for (int d = 0; d < 3; ++d) {
            src.acc[d] = (T)s.fp.acceleration[d];
            src.any |= s.fp.acceleration[d] != 0.0;
        }
        src.damping = (T)s.fp.damping_coefficient;
        int has_wall = s.n_w > 0 && s.interaction[0][1];
        s.wall_viscous_done = false;
        s.stats.interact_variant_used = use_tiles(s) ? 2 : 1;
        if (use_tiles(s)) {
            const int list_len = s.tiles.list(KS);
            const int cap = tile_capacity<T, CT>(s.tiles.smem_budget, list_len, KS);
            const size_t smem = tile_smem_bytes<T, CT>(cap, list_len, KS);
            if (!(s.smem_opt_in & 2)) {
                int rc = set_smem(s, k_interact_tiles<KS, ND, T, CT, KERNEL, DENS>, 227 * 1024);
                if (rc) return rc;
                rc = set_smem(s, k_interact_tiles<KS, ND, T, CT, KERNEL, DENS, true>, 227 * 1024);
                if (rc) return rc;
                s.smem_opt_in |= 2;
            }
            if (has_wall && s.wp.has_viscosity) {
                // no-slip wall: pressure and viscous term of the wall in one sweep
                s.wall_viscous_done = true;
                LAUNCH(s, (k_interact_tiles<KS, ND, T, CT, KERNEL, DENS, true>), s.tiles.max_ftiles, KS * TILE_TB,
                       smem, g, s.tiles.d_frow_tile_start + s.tiles.nrows, s.tiles.d_ftile_desc,
                       s.tiles.d_ftile_ext, s.tiles.d_ftile_rng, s.d_fcell_start, (const V4<CT> *)s.d_A,
                       (const V4<T> *)s.d_B, (const T *)s.d_P, s.d_perm_f, s.interaction[0][0], has_wall,
                       s.d_wcell_start, (const V4<CT> *)s.d_Aw, (const V2<T> *)s.d_Ww, pc, src, d_dv,
                       (int)s.n_tgt, cap, list_len, (const V4<float> *)s.d_Ff, (const V4<float> *)s.d_Fw,
                       (const V4<T> *)s.d_Vw, (const T *)s.d_Pw, make_wall_visc_const(s, pc),
                       (const AdaptConsts<T> *)s.ad_kick);
                return TPB_OK;
            }
            LAUNCH(s, (k_interact_tiles<KS, ND, T, CT, KERNEL, DENS>), s.tiles.max_ftiles, KS * TILE_TB, smem, g,
                   s.tiles.d_frow_tile_start + s.tiles.nrows, s.tiles.d_ftile_desc, s.tiles.d_ftile_ext,
                   s.tiles.d_ftile_rng, s.d_fcell_start, (const V4<CT> *)s.d_A,
                   (const V4<T> *)s.d_B, (const T *)s.d_P, s.d_perm_f, s.interaction[0][0], has_wall,
                   s.d_wcell_start, (const V4<CT> *)s.d_Aw, (const V2<T> *)s.d_Ww, pc, src, d_dv,
                   (int)s.n_tgt, cap, list_len, (const V4<float> *)s.d_Ff, (const V4<float> *)s.d_Fw,
                   (const V4<T> *)nullptr, (const T *)nullptr, WallViscConst<T>(), (const AdaptConsts<T> *)s.ad_kick);
            return TPB_OK;
        }
        LAUNCH(s, (k_interact_pp<ND, T, CT, KERNEL, DENS>), cdiv(n, 128), 128, 0, n, g,
               s.d_fcell_start, (const V4<CT> *)s.d_A, (const V4<T> *)s.d_B, (const T *)s.d_P,
               s.d_perm_f, s.interaction[0][0], has_wall, s.d_wcell_start,
               (const V4<CT> *)s.d_Aw, (const V2<T> *)s.d_Ww, pc, src, d_dv, (int)s.n_tgt);
        return TPB_OK;
    }

    // ---- update_speed_of_sound! (wcsph/system.jl:307-321, StateEquationAdaptiveCole): one max
    // reduction over the fluid velocities; like the reference (`maximum` returns a host scalar)
    // the new sound speed is a host value, so this kick waits for the reduction.
    static int update_sound_speed(Semi &s, const T *d_v)
    {
        const int64_t n = s.n_act;
        CUDA_TRY(&s, cudaMemsetAsync(s.d_vmax2, 0, sizeof(unsigned long long), s.stream));
        LAUNCH(s, (k_max_speed2<ND, T>), (int)std::min<int64_t>(cdiv(n, 256), 148 * 8), 256, 0, n, nv(s), d_v,
               s.d_vmax2);
        CUDA_TRY(&s, cudaMemcpyAsync(s.h_vmax2, s.d_vmax2, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                     s.stream));
        CUDA_TRY(&s, cudaStreamSynchronize(s.stream));
        T v_max2;
        if (sizeof(T) == 4) {
            const uint32_t bits = (uint32_t)*s.h_vmax2;
            std::memcpy(&v_max2, &bits, sizeof(T));
        } else {
            std::memcpy(&v_max2, s.h_vmax2, sizeof(T));
        }
        const T v_max = std::sqrt(v_max2);
        double c;
        if (s.fp.adaptive_params_f32 && sizeof(T) == 8) {
            // Float64 velocities, Float32 parameters: the quotient and the clamps are evaluated in
            // Float64 (promotion), the result is stored in the state equation's Float32 Ref
            const double q = (double)v_max / (double)(float)s.fp.mach_number_target;
            c = (double)(float)std::min((double)(float)s.fp.max_sound_speed,
                                        std::max((double)(float)s.fp.min_sound_speed, q));
        } else {
            const T q = v_max / (T)s.fp.mach_number_target;
            c = (double)std::min((T)s.fp.max_sound_speed, std::max((T)s.fp.min_sound_speed, q));
        }
        s.fp.sound_speed = c;
        if (s.wp.sound_speed_from_fluid) s.wp.sound_speed = c;
        s.c_on_device = false;
        return TPB_OK;
    }

    // ---- the same without the host: k_max_speed2 -> k_adaptive_consts -> the kick's kernels read
    // AdaptConsts.  Stream-ordered (no synchronisation, CUDA-graph capturable); the maximum may come from
    // outside (tpb_set_max_speed2: reduced over all slabs).  Tile path with ContinuityDensity only.
    static bool adaptive_on_device(const Semi &s)
    {
        const bool off = getenv("TPB_ADAPTIVE_HOST") != nullptr;
        return !off && (s.cfg.interact_variant == 0 || s.cfg.interact_variant == 2) &&
               s.fp.density_calculator == TPB_DENSITY_CONTINUITY && s.struct_index < 0;
    }
    static int max_speed2(Semi &s, const T *d_v, int64_t n, unsigned long long *d_out)
    {
        CUDA_TRY(&s, cudaMemsetAsync(d_out, 0, sizeof(unsigned long long), s.stream));
        if (n > 0)
            LAUNCH(s, (k_max_speed2<ND, T>), (int)std::min<int64_t>(cdiv(n, 256), 148 * 8), 256, 0, n, nv(s), d_v,
                   d_out);
        return TPB_OK;
    }
    static int update_sound_speed_device(Semi &s, const T *d_v)
    {
        if (!s.d_adapt) CUDA_TRY(&s, cudaMalloc(&s.d_adapt, sizeof(AdaptConsts<T>) + 16));
        if (!s.vmax2_ready) {
            int rc = max_speed2(s, d_v, s.n_act, s.d_vmax2);
            if (rc) return rc;
        }
        s.vmax2_ready = false;
        AdaptParams<T> p{};
        p.mach = (T)s.fp.mach_number_target, p.c_min = (T)s.fp.min_sound_speed, p.c_max = (T)s.fp.max_sound_speed;
        p.mach_f = (float)s.fp.mach_number_target, p.c_min_f = (float)s.fp.min_sound_speed;
        p.c_max_f = (float)s.fp.max_sound_speed;
        p.params_f32 = s.fp.adaptive_params_f32;
        p.gamma_f = (T)s.fp.exponent, p.rho0_f = (T)s.fp.reference_density;
        p.gamma_f32 = (float)s.fp.exponent, p.rho0_f32 = (float)s.fp.reference_density;
        p.wall_follows = s.n_w > 0 && s.wp.sound_speed_from_fluid;
        p.gamma_w = (T)s.wp.exponent, p.rho0_w = (T)s.wp.reference_density, p.pbg_w = (T)s.wp.background_pressure;
        p.inv_gamma_w = (T)1 / p.gamma_w;
        p.gamma_w32 = (float)s.wp.exponent, p.rho0_w32 = (float)s.wp.reference_density;
        const KernelConst<T> kern = make_kernel_const<T>(s.fp.kernel, ND, s.fp.smoothing_length);
        p.delta_h = (T)s.fp.delta * ((kern.h + kern.h) / (T)2);
        p.visc_f = s.fp.has_viscosity, p.visc_w = s.wp.has_viscosity, p.nd = ND;
        p.alpha_f = (T)s.fp.alpha, p.alpha_w = (T)s.wp.alpha;
        p.h_f = kern.h, p.h_w = (T)s.wp.smoothing_length;
        LAUNCH(s, (k_adaptive_consts<T>), 1, 32, 0, s.d_vmax2, p, (AdaptConsts<T> *)s.d_adapt,
               (double *)((unsigned char *)s.d_adapt + sizeof(AdaptConsts<T>)));
        s.ad_kick = s.d_adapt;
        s.c_on_device = true;
        return TPB_OK;
    }

    // ---- kick! on device pointers
    static int kick_device(Semi &s, T *d_dv_ode, const T *d_v_ode, const CT *d_u_ode)
    {
        const OdeLayout lay = ode_layout(s);
        T *d_dv = d_dv_ode + lay.off_v_f;
        const T *d_v = d_v_ode + lay.off_v_f;
        const CT *d_u = d_u_ode + lay.off_u_f;
        if (s.n_act == 0 && s.n_s == 0) return TPB_OK;
        s.ad_kick = nullptr;
        if (s.fp.adaptive_sound_speed) {
            if (!adaptive_on_device(s) && s.n_tgt < s.n_act)
                return fail(&s, TPB_ERR_UNSUPPORTED,
                            "StateEquationAdaptiveCole across slabs needs the device path (tile sweeps, ContinuityDensity)");
            if (adaptive_on_device(s) && s.n_tgt < s.n_act && !s.vmax2_ready)
                return fail(&s, TPB_ERR_STATE,
                            "StateEquationAdaptiveCole across slabs: call tpb_set_max_speed2 (maximum over all slabs) before every kick");
            int rc_c = adaptive_on_device(s) ? update_sound_speed_device(s, d_v) : update_sound_speed(s, d_v);
            if (rc_c) return rc_c;
        }
        prof_mark(s, TPB_PHASE_REBUILD);
        int rc = rebuild_fluid(s, d_u, d_v);
        if (rc) return rc;
        GridConst<CT> g = make_grid_const<CT>(s);
        PairConst<T> pc = make_pair_const<T>(s.fp, ND);
        EosConst<T> eos = make_eos_const<T>(s.fp.sound_speed, s.fp.exponent, s.fp.reference_density,
                                            s.fp.background_pressure, s.fp.clip_negative_pressure,
                                            s.fp.adaptive_sound_speed && s.fp.adaptive_params_f32);
        const bool summ = s.fp.density_calculator == TPB_DENSITY_SUMMATION;
        if (s.struct_index >= 0) {
            rc = update_structure(s, g, pc, d_u_ode + lay.off_u_s, d_v_ode + lay.off_v_s);
            if (rc) return rc;
        }
        prof_mark(s, TPB_PHASE_DENSITY);
        // kernel template value: 0 Wendland C2, 1 cubic spline, 2 Wendland C4 / C6
        auto tk = [](int kernel) { return kernel <= 1 ? kernel : kernel <= TPB_KERNEL_WENDLAND_C6 ? 2 : 3; };
        const int fk = tk(s.fp.kernel), wk = tk(s.wp.kernel);
        if (summ) {
            if (fk == 0) launch_summation<0>(s, g, pc, eos);
            else if (fk == 1) launch_summation<1>(s, g, pc, eos);
            else if (fk == 2) launch_summation<2>(s, g, pc, eos);
            else launch_summation<3>(s, g, pc, eos);
        }
        prof_mark(s, TPB_PHASE_BOUNDARY);
        if (s.n_w > 0 && wall_integrates_density(s)) {
            // BoundaryModelDummyParticles{ContinuityDensity}: density = the wall's rows of v_ode,
            // pressure = state_equation(density) (dummy_particles.jl:364-368, :458-478)
            EosConst<T> weos = make_eos_const<T>(s.wp.sound_speed, s.wp.exponent, s.wp.reference_density,
                                                 s.wp.background_pressure, s.wp.eos_clip_negative_pressure,
                                                 s.wp.sound_speed_from_fluid && s.fp.adaptive_params_f32);
            LAUNCH(s, (k_wall_density_eos<T>), cdiv(s.n_w, 256), 256, 0, (int)s.n_w, d_v_ode + lay.off_v_w, s.d_perm_w,
                   weos, s.wp.clip_negative_pressure, (V2<T> *)s.d_Ww, (T *)s.d_volw,
                   s.wp.sound_speed_from_fluid ? (const AdaptConsts<T> *)s.ad_kick : (const AdaptConsts<T> *)nullptr);
        } else if (s.n_w > 0) {
            rc = wk == 0   ? launch_adami<0>(s, g)
                 : wk == 1 ? launch_adami<1>(s, g)
                 : wk == 2 ? launch_adami<2>(s, g)
                           : launch_adami<3>(s, g);
            if (rc) return rc;
            if (s.wp.has_viscosity && !s.wall_velocity_done) {
                rc = wk == 0   ? launch_wall_velocity<0>(s, g)
                     : wk == 1 ? launch_wall_velocity<1>(s, g)
                     : wk == 2 ? launch_wall_velocity<2>(s, g)
                               : launch_wall_velocity<3>(s, g);
                if (rc) return rc;
            }
        }
        prof_mark(s, TPB_PHASE_INTERACT);
        if (fk == 0)
            rc = summ ? launch_interact<0, 1>(s, g, pc, d_dv) : launch_interact<0, 0>(s, g, pc, d_dv);
        else if (fk == 1)
            rc = summ ? launch_interact<1, 1>(s, g, pc, d_dv) : launch_interact<1, 0>(s, g, pc, d_dv);
        else if (fk == 2)
            rc = summ ? launch_interact<2, 1>(s, g, pc, d_dv) : launch_interact<2, 0>(s, g, pc, d_dv);
        else
            rc = summ ? launch_interact<3, 1>(s, g, pc, d_dv) : launch_interact<3, 0>(s, g, pc, d_dv);
        if (rc) return rc;
        if (s.n_w > 0 && s.wp.has_viscosity && s.interaction[0][1] && !s.wall_viscous_done) {
            if (fk == 0)
                rc = summ ? launch_wall_viscous<0, 1>(s, g, pc, d_dv) : launch_wall_viscous<0, 0>(s, g, pc, d_dv);
            else if (fk == 1)
                rc = summ ? launch_wall_viscous<1, 1>(s, g, pc, d_dv) : launch_wall_viscous<1, 0>(s, g, pc, d_dv);
            else if (fk == 2)
                rc = summ ? launch_wall_viscous<2, 1>(s, g, pc, d_dv) : launch_wall_viscous<2, 0>(s, g, pc, d_dv);
            else
                rc = summ ? launch_wall_viscous<3, 1>(s, g, pc, d_dv) : launch_wall_viscous<3, 0>(s, g, pc, d_dv);
            if (rc) return rc;
        }
        if (s.struct_index >= 0) {
            T *d_dv_s = d_dv_ode + lay.off_v_s;
            const int sm = s.integrate_structure ? 0 : 1;
            if (fk == 0)
                rc = summ ? interact_structure<0, 1>(s, g, pc, d_dv, d_dv_s, sm) : interact_structure<0, 0>(s, g, pc, d_dv, d_dv_s, sm);
            else if (fk == 1)
                rc = summ ? interact_structure<1, 1>(s, g, pc, d_dv, d_dv_s, sm) : interact_structure<1, 0>(s, g, pc, d_dv, d_dv_s, sm);
            else if (fk == 2)
                rc = summ ? interact_structure<2, 1>(s, g, pc, d_dv, d_dv_s, sm) : interact_structure<2, 0>(s, g, pc, d_dv, d_dv_s, sm);
            else
                rc = summ ? interact_structure<3, 1>(s, g, pc, d_dv, d_dv_s, sm) : interact_structure<3, 0>(s, g, pc, d_dv, d_dv_s, sm);
            if (rc) return rc;
        }
        if (s.n_w > 0 && wall_integrates_density(s)) {
            // interact!(wall, fluid): the continuity equation of the dummy particles (wall_boundary/rhs.jl:11-59)
            KernelConst<T> wkern = make_kernel_const<T>(s.wp.kernel, ND, s.wp.smoothing_length);
            const T Rw = wkern.support;
            const T az = (T)std::sqrt(eps_of<T>(wkern.h * wkern.h));
            T *d_dv_w = d_dv_ode + lay.off_v_w;
            auto go = [&](auto kernel_tag) {
                constexpr int K = decltype(kernel_tag)::value;
                LAUNCH(s, (k_wall_continuity<ND, T, CT, K>), cdiv(s.n_w, 128), 128, 0, (int)s.n_w, g,
                       (const V4<CT> *)s.d_Aw, (const V2<T> *)s.d_Ww, s.d_perm_w, s.d_fcell_start,
                       (const V4<CT> *)s.d_A, (const V4<T> *)s.d_B, s.interaction[1][0], wkern, (T)(Rw * Rw), az,
                       (int)summ, d_dv_w);
            };
            if (wk == 0) go(std::integral_constant<int, 0>());
            else if (wk == 1) go(std::integral_constant<int, 1>());
            else if (wk == 2) go(std::integral_constant<int, 2>());
            else go(std::integral_constant<int, 3>());
        }
        prof_mark(s, TPB_PHASE_END);
        if (s.prof_capacity > 0 && s.prof_kicks < s.prof_capacity) s.prof_kicks++;
        CUDA_TRY(&s, cudaGetLastError());
        return TPB_OK;
    }

    static int kick(Semi &s, void *dv, const void *v, const void *u)
    {
        const OdeLayout lay = ode_layout(s);
        const size_t nu = sizeof(CT) * (size_t)lay.tot_u, nvb = sizeof(T) * (size_t)lay.tot_v;
        s.launches_this_call = 0;
        int rc;
        if (s.cfg.ode_memory == TPB_MEM_HOST) {
            CUDA_TRY(&s, cudaMemcpyAsync(s.d_u, u, nu, cudaMemcpyHostToDevice, s.stream));
            CUDA_TRY(&s, cudaMemcpyAsync(s.d_v, v, nvb, cudaMemcpyHostToDevice, s.stream));
            // Page-locked (pinned / tpb_host_register'ed) dv: the interact! kernel stores every
            // particle's dv straight into the mapped host buffer, so the device->host transfer
            // rides along with the kernel instead of following it.
            void *dv_alias = s.host_zero_copy ? mapped_host_alias(dv) : nullptr;
            rc = kick_device(s, (T *)(dv_alias ? dv_alias : s.d_dv), (const T *)s.d_v, (const CT *)s.d_u);
            if (rc) return rc;
            if (!dv_alias)
                CUDA_TRY(&s, cudaMemcpyAsync(dv, s.d_dv, nvb, cudaMemcpyDeviceToHost, s.stream));
            CUDA_TRY(&s, cudaMemcpyAsync(s.h_flags, s.d_flags, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
            CUDA_TRY(&s, cudaStreamSynchronize(s.stream));
            if (*s.h_flags & 1)
                return fail(&s, TPB_ERR_OUT_OF_BOUNDS,
                            "particle coordinates are NaN or outside the FullGridCellList bounding box");
        } else {
            rc = kick_device(s, (T *)dv, (const T *)v, (const CT *)u);
            if (rc) return rc;
            CUDA_TRY(&s, cudaMemcpyAsync(s.h_flags, s.d_flags, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
        }
        s.stats.kicks++;
        s.stats.launches_last_kick = s.launches_this_call;
        return TPB_OK;
    }

    static int drift(Semi &s, void *du, const void *v, const void *u)
    {
        (void)u;
        // (slab handles drift their owned rows only: n_tgt, not n_act)
        OdeLayout lay = ode_layout(s);
        const int64_t total_f = s.n_tgt * ND, total_s = s.struct_index >= 0 ? s.n_s_int * ND : 0;
        if (s.struct_index < 0 && lay.len_v_w == 0) lay.tot_u = total_f, lay.tot_v = s.n_tgt * nv(s);
        const size_t nu = sizeof(CT) * (size_t)lay.tot_u, nvb = sizeof(T) * (size_t)lay.tot_v;
        s.launches_this_call = 0;
        if (total_f + total_s == 0) return TPB_OK;
        auto launch = [&](const T *v_base, CT *du_base) {
            // fluid and (integrated) structure in one launch; structure: v holds ND entries per particle, du = v
            const int64_t n_s = s.integrate_structure ? total_s : 0;
            if (total_f + n_s > 0)
                LAUNCH(s, (k_drift2<ND, T, CT>), cdiv(total_f + n_s, 256), 256, 0, total_f, nv(s), v_base + lay.off_v_f,
                       du_base + lay.off_u_f, n_s, v_base + lay.off_v_s, du_base + lay.off_u_s);
            if (total_s > 0 && !s.integrate_structure)  // set_velocity! without integrate_tlsph: du = 0
                cudaMemsetAsync(du_base + lay.off_u_s, 0, sizeof(CT) * (size_t)total_s, s.stream);
        };
        if (s.cfg.ode_memory == TPB_MEM_HOST) {
            // Page-locked v and du: one kernel streams v in and du out over PCIe at the same
            // time (full duplex) instead of copy -> kernel -> copy.
            const void *v_alias = s.host_zero_copy ? mapped_host_alias(v) : nullptr;
            void *du_alias = s.host_zero_copy ? mapped_host_alias(du) : nullptr;
            if (v_alias && du_alias) {
                launch((const T *)v_alias, (CT *)du_alias);
            } else {
                CUDA_TRY(&s, cudaMemcpyAsync(s.d_v, v, nvb, cudaMemcpyHostToDevice, s.stream));
                launch((const T *)s.d_v, (CT *)s.d_du);
                CUDA_TRY(&s, cudaMemcpyAsync(du, s.d_du, nu, cudaMemcpyDeviceToHost, s.stream));
            }
            CUDA_TRY(&s, cudaStreamSynchronize(s.stream));
        } else {
            launch((const T *)v, (CT *)du);
            CUDA_TRY(&s, cudaGetLastError());
        }
        s.stats.drifts++;
        s.stats.launches_last_drift = s.launches_this_call;
        return TPB_OK;
    }

    // `DensityReinitializationCallback` (callbacks/density_reinit.jl:83-121; reinit_density!, wcsph/system.jl:398-415):
    // the density rows of the fluid in v_ode are replaced by the Shepard-corrected summation density.  Systems are
    // updated for the given state first (sorted records, wall densities), as initialize_reinit_cb! does.
    static int reinit_density(Semi &s, void *v_ode, const void *u_ode)
    {
        if (s.fp.density_calculator != TPB_DENSITY_CONTINUITY) return TPB_OK;  // reinit_density!(..., ::SummationDensity) = nothing
        if (s.struct_index >= 0 || s.n_tgt != s.n_act || s.fp.adaptive_sound_speed || wall_integrates_density(s))
            return fail(&s, TPB_ERR_UNSUPPORTED, "tpb_reinit_density: fluid + Adami walls with a fixed speed of sound only "
                                                 "(no structure, slab ghosts, ContinuityDensity wall, adaptive state equation)");
        const int n = (int)s.n_act;
        if (n == 0) return TPB_OK;
        const OdeLayout lay = ode_layout(s);
        const int NV = nv(s);
        const bool host = s.cfg.ode_memory == TPB_MEM_HOST;
        const size_t ub = sizeof(CT) * (size_t)lay.tot_u, vb = sizeof(T) * (size_t)lay.tot_v;
        const CT *u = host ? (const CT *)s.d_u : (const CT *)u_ode;
        T *v = host ? (T *)s.d_v : (T *)v_ode;
        if (host) {
            CUDA_TRY(&s, cudaMemcpyAsync(s.d_u, u_ode, ub, cudaMemcpyHostToDevice, s.stream));
            CUDA_TRY(&s, cudaMemcpyAsync(s.d_v, v_ode, vb, cudaMemcpyHostToDevice, s.stream));
        }
        s.ad_kick = nullptr;
        int rc = rebuild_fluid(s, u + lay.off_u_f, v + lay.off_v_f);
        if (rc) return rc;
        GridConst<CT> g = make_grid_const<CT>(s);
        PairConst<T> pc = make_pair_const<T>(s.fp, ND);
        EosConst<T> eos = make_eos_const<T>(s.fp.sound_speed, s.fp.exponent, s.fp.reference_density,
                                            s.fp.background_pressure, s.fp.clip_negative_pressure, false);
        auto tk = [](int kernel) { return kernel <= 1 ? kernel : kernel <= TPB_KERNEL_WENDLAND_C6 ? 2 : 3; };
        const int fk = tk(s.fp.kernel), wk = tk(s.wp.kernel);
        const bool has_wall = s.n_w > 0 && s.interaction[0][1];
        if (s.n_w > 0) {  // wall densities of the state before the reinitialisation
            rc = wk == 0 ? launch_adami<0>(s, g) : wk == 1 ? launch_adami<1>(s, g) : wk == 2 ? launch_adami<2>(s, g)
                                                                                              : launch_adami<3>(s, g);
            if (rc) return rc;
        }
        switch (fk) {  // first pass: rho~ = sum m W into the sorted records
#define TPB_REINIT(K)                                                                                              \
    case K:                                                                                                        \
        launch_summation<K>(s, g, pc, eos);                                                                        \
        LAUNCH(s, (k_shepard_reinit<ND, T, CT, K>), cdiv(n, 128), 128, 0, n, g, s.d_fcell_start,                   \
               (const V4<CT> *)s.d_A, (const V4<T> *)s.d_B, s.d_perm_f, (int)has_wall, s.d_wcell_start,            \
               (const V4<CT> *)s.d_Aw, (const V2<T> *)s.d_Ww, pc.kern, pc.radius2, NV, (int)s.n_tgt,               \
               v + lay.off_v_f);                                                                                   \
        break;
            TPB_REINIT(0) TPB_REINIT(1) TPB_REINIT(2) TPB_REINIT(3)
#undef TPB_REINIT
        }
        if (host) {
            CUDA_TRY(&s, cudaMemcpyAsync(v_ode, s.d_v, vb, cudaMemcpyDeviceToHost, s.stream));
            CUDA_TRY(&s, cudaMemcpyAsync(s.h_flags, s.d_flags, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
            CUDA_TRY(&s, cudaStreamSynchronize(s.stream));
            if (*s.h_flags & 1)
                return fail(&s, TPB_ERR_OUT_OF_BOUNDS,
                            "particle coordinates are NaN or outside the FullGridCellList bounding box");
        }
        CUDA_TRY(&s, cudaGetLastError());
        return TPB_OK;
    }

    // `SortingCallback` (callbacks/sorting.jl:100-157): the fluid's rows of (v_ode, u_ode) -- and the library's
    // per-particle masses -- reordered by the cell of their current coordinates.  Scratch: the sorted pressure
    // array (rewritten by the next kick) and the handle's dv / du staging buffers (host mode) or two buffers of
    // their own (device mode).
    static int sort_system(Semi &s, void *v_ode, void *u_ode)
    {
        const int n = (int)s.n_act;
        if (s.n_tgt != s.n_act)
            return fail(&s, TPB_ERR_UNSUPPORTED, "tpb_sort_system: not with slab ghosts (the exchange owns the row order)");
        if (n == 0) return TPB_OK;
        const OdeLayout lay = ode_layout(s);
        const int NV = nv(s);
        const bool host = s.cfg.ode_memory == TPB_MEM_HOST;
        const size_t ub = sizeof(CT) * ND * (size_t)n, vb = sizeof(T) * NV * (size_t)n;
        CT *u = host ? (CT *)s.d_u : (CT *)u_ode + lay.off_u_f;
        T *v = host ? (T *)s.d_v : (T *)v_ode + lay.off_v_f;
        if (host) {
            CUDA_TRY(&s, cudaMemcpyAsync(u, (const CT *)u_ode + lay.off_u_f, ub, cudaMemcpyHostToDevice, s.stream));
            CUDA_TRY(&s, cudaMemcpyAsync(v, (const T *)v_ode + lay.off_v_f, vb, cudaMemcpyHostToDevice, s.stream));
        }
        if (!host && !s.d_sort_u) {
            CUDA_TRY(&s, cudaMalloc(&s.d_sort_u, sizeof(CT) * ND * (size_t)s.n_f));
            CUDA_TRY(&s, cudaMalloc(&s.d_sort_v, sizeof(T) * NV * (size_t)s.n_f));
        }
        CT *out_u = host ? (CT *)s.d_du : (CT *)s.d_sort_u;
        T *out_v = host ? (T *)s.d_dv : (T *)s.d_sort_v;
        int rc = bin_points(s, u, n, n, s.d_fcell_start);
        if (rc) return rc;
        LAUNCH(s, (k_sort_gather<ND, T, CT>), cdiv(n, 256), 256, 0, (const CT *)u, (const T *)v, (const T *)s.d_mass_f,
               s.d_key, s.d_fcell_start, s.d_tmp_perm, n, NV, out_u, out_v, (T *)s.d_P, perm_bits(s, n));
        CUDA_TRY(&s, cudaMemcpyAsync(s.d_mass_f, s.d_P, sizeof(T) * (size_t)n, cudaMemcpyDeviceToDevice, s.stream));
        // the fused rebuild (and a CUDA graph that captured it) counts into a histogram its scan left at zero
        CUDA_TRY(&s, cudaMemsetAsync(s.d_count, 0, sizeof(int) * (size_t)s.ncells, s.stream));
        s.count_clean = true;
        if (host) {
            CUDA_TRY(&s, cudaMemcpyAsync((CT *)u_ode + lay.off_u_f, out_u, ub, cudaMemcpyDeviceToHost, s.stream));
            CUDA_TRY(&s, cudaMemcpyAsync((T *)v_ode + lay.off_v_f, out_v, vb, cudaMemcpyDeviceToHost, s.stream));
            CUDA_TRY(&s, cudaMemcpyAsync(s.h_flags, s.d_flags, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
            CUDA_TRY(&s, cudaStreamSynchronize(s.stream));
            if (*s.h_flags & 1)
                return fail(&s, TPB_ERR_OUT_OF_BOUNDS,
                            "particle coordinates are NaN or outside the FullGridCellList bounding box");
        } else {
            CUDA_TRY(&s, cudaMemcpyAsync(u, out_u, ub, cudaMemcpyDeviceToDevice, s.stream));
            CUDA_TRY(&s, cudaMemcpyAsync(v, out_v, vb, cudaMemcpyDeviceToDevice, s.stream));
        }
        return TPB_OK;
    }

    static int get_field(Semi &s, int system, int field, void *out, int64_t n)
    {
        if (field == TPB_FIELD_WALL_VELOCITY) {
            // ND x n_w values; not on the hot path: a buffer of its own
            const Semi::WallPart *part = s.wall_part(system);
            if (!part || !s.d_Vw)
                return fail(&s, TPB_ERR_INVALID_ARGUMENT, "wall_velocity: the system is not a wall with a viscosity model");
            if (n != part->n) return fail(&s, TPB_ERR_INVALID_ARGUMENT, "field length mismatch");
            if (n == 0) return TPB_OK;
            T *d_tmp = nullptr;
            const size_t bytes = sizeof(T) * ND * (size_t)s.n_w;
            CUDA_TRY(&s, cudaMalloc(&d_tmp, bytes));
            cudaMemsetAsync(d_tmp, 0, bytes, s.stream);
            LAUNCH(s, (k_unsort_vector<T>), cdiv(s.n_w, 256), 256, 0, (int)s.n_w, s.d_wcell_start + s.ncells, s.d_perm_w,
                   (const V4<T> *)s.d_Vw, ND, d_tmp);
            cudaError_t e = cudaMemcpyAsync(out, d_tmp + ND * part->first, sizeof(T) * ND * (size_t)n,
                                            cudaMemcpyDeviceToHost, s.stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
            cudaFree(d_tmp);
            CUDA_TRY(&s, e);
            return TPB_OK;
        }
        if (field == TPB_FIELD_DEFORMATION_GRADIENT || field == TPB_FIELD_PK1_RHO2 || field == TPB_FIELD_CORRECTION_MATRIX) {
            if (system != s.struct_index) return fail(&s, TPB_ERR_INVALID_ARGUMENT, "the system is not the structure system");
            if (n != s.n_s) return fail(&s, TPB_ERR_INVALID_ARGUMENT, "field length mismatch");
            if (n == 0) return TPB_OK;
            const void *src = field == TPB_FIELD_DEFORMATION_GRADIENT ? s.d_F_s
                              : field == TPB_FIELD_PK1_RHO2           ? s.d_pk1_s
                                                                      : s.d_L_s;
            CUDA_TRY(&s, cudaMemcpyAsync(out, src, sizeof(T) * ND * ND * (size_t)n, cudaMemcpyDeviceToHost, s.stream));
            CUDA_TRY(&s, cudaStreamSynchronize(s.stream));
            return TPB_OK;
        }
        if (system == s.struct_index && s.struct_index >= 0) {
            // boundary-model pressure / density of a structure with dummy particles (own particle order)
            if (s.sp.boundary_model != TPB_BOUNDARY_DUMMY_PARTICLES || (field != TPB_FIELD_PRESSURE && field != TPB_FIELD_DENSITY))
                return fail(&s, TPB_ERR_INVALID_ARGUMENT, "structure field: pressure / density exist with BoundaryModelDummyParticles only");
            if (n != s.n_s) return fail(&s, TPB_ERR_INVALID_ARGUMENT, "field length mismatch");
            if (n == 0) return TPB_OK;
            CUDA_TRY(&s, cudaMemcpyAsync(out, field == TPB_FIELD_PRESSURE ? s.d_p_s : s.d_rhoh_s, sizeof(T) * (size_t)n,
                                         cudaMemcpyDeviceToHost, s.stream));
            CUDA_TRY(&s, cudaStreamSynchronize(s.stream));
            return TPB_OK;
        }
        T *scratch = (T *)s.d_scratch;
        CUDA_TRY(&s, cudaMemsetAsync(scratch, 0, sizeof(T) * (size_t)std::max<int64_t>(n, 1), s.stream));
        if (system == s.fluid_index) {
            if (n != s.n_act) return fail(&s, TPB_ERR_INVALID_ARGUMENT, "field length mismatch");
            if (n == 0) return TPB_OK;
            if (field == TPB_FIELD_PRESSURE)
                LAUNCH(s, (k_unsort_scalar<T>), cdiv(n, 256), 256, 0, (int)n, s.d_fcell_start + s.ncells, s.d_perm_f, (const T *)s.d_P, 1, 0, scratch);
            else if (field == TPB_FIELD_DENSITY)
                LAUNCH(s, (k_unsort_scalar<T>), cdiv(n, 256), 256, 0, (int)n, s.d_fcell_start + s.ncells, s.d_perm_f, (const T *)s.d_B, 4, 3, scratch);
            else
                return fail(&s, TPB_ERR_INVALID_ARGUMENT, "unknown fluid field");
        } else if (const Semi::WallPart *part = s.wall_part(system)) {
            // the wall set is unsorted as a whole, the system's own particles are a slice of it
            if (n != part->n) return fail(&s, TPB_ERR_INVALID_ARGUMENT, "field length mismatch");
            if (n == 0) return TPB_OK;
            const int nw = (int)s.n_w;
            CUDA_TRY(&s, cudaMemsetAsync(scratch, 0, sizeof(T) * (size_t)nw, s.stream));
            if (field == TPB_FIELD_PRESSURE)
                LAUNCH(s, (k_unsort_scalar<T>), cdiv(nw, 256), 256, 0, nw, s.d_wcell_start + s.ncells, s.d_perm_w, (const T *)s.d_Ww, 2, 0, scratch);
            else if (field == TPB_FIELD_DENSITY)
                LAUNCH(s, (k_unsort_scalar<T>), cdiv(nw, 256), 256, 0, nw, s.d_wcell_start + s.ncells, s.d_perm_w, (const T *)s.d_Ww, 2, 1, scratch);
            else if (field == TPB_FIELD_VOLUME)
                LAUNCH(s, (k_unsort_scalar<T>), cdiv(nw, 256), 256, 0, nw, s.d_wcell_start + s.ncells, s.d_perm_w, (const T *)s.d_volw, 1, 0, scratch);
            else
                return fail(&s, TPB_ERR_INVALID_ARGUMENT, "unknown wall field");
            scratch += part->first;
        } else {
            return fail(&s, TPB_ERR_INVALID_ARGUMENT, "unknown system index");
        }
        CUDA_TRY(&s, cudaMemcpyAsync(out, scratch, sizeof(T) * (size_t)n, cudaMemcpyDeviceToHost, s.stream));
        CUDA_TRY(&s, cudaStreamSynchronize(s.stream));
        return TPB_OK;
    }

    static int neighbor_pairs(Semi &s, int system, int neighbor, const void *u_ode, int64_t capacity,
                              int32_t *out_i, int32_t *out_j, int64_t *count)
    {
        // rebuild the fluid grid for the given coordinates (velocities are irrelevant here)
        const size_t nu = sizeof(CT) * ND * (size_t)s.n_act, nvb = sizeof(T) * nv(s) * (size_t)s.n_act;
        u_ode = (const CT *)u_ode + ode_layout(s).off_u_f;  // the fluid's part of u_ode
        const CT *d_u = (const CT *)u_ode;
        T *d_vzero = nullptr;
        CUDA_TRY(&s, cudaMalloc(&d_vzero, std::max(nvb, (size_t)16)));
        CUDA_TRY(&s, cudaMemsetAsync(d_vzero, 0, std::max(nvb, (size_t)16), s.stream));
        CT *d_utmp = nullptr;
        if (s.cfg.ode_memory == TPB_MEM_HOST) {
            CUDA_TRY(&s, cudaMalloc(&d_utmp, std::max(nu, (size_t)16)));
            CUDA_TRY(&s, cudaMemcpyAsync(d_utmp, u_ode, nu, cudaMemcpyHostToDevice, s.stream));
            d_u = d_utmp;
        }
        int rc = rebuild_fluid(s, d_u, d_vzero);
        if (rc) return rc;
        GridConst<CT> g = make_grid_const<CT>(s);
        const bool x_fluid = system == s.fluid_index, y_fluid = neighbor == s.fluid_index;
        if (s.wall_parts.size() > 1 && (!x_fluid || !y_fluid))
            return fail(&s, TPB_ERR_UNSUPPORTED, "tpb_neighbor_pairs with several wall systems: fluid-fluid pairs only");
        if ((!x_fluid && system != s.wall_index) || (!y_fluid && neighbor != s.wall_index))
            return fail(&s, TPB_ERR_INVALID_ARGUMENT, "unknown system index");
        // radius of the ordered pair: compact_support(system, neighbor)
        T R = make_kernel_const<T>(x_fluid ? s.fp.kernel : s.wp.kernel, ND,
                                   x_fluid ? s.fp.smoothing_length : s.wp.smoothing_length).support;
        T r2 = R * R;
        int n_x = (int)(x_fluid ? s.n_act : s.n_w);
        int *d_oi = nullptr, *d_oj = nullptr;
        unsigned long long *d_counter = nullptr;
        CUDA_TRY(&s, cudaMalloc(&d_oi, sizeof(int) * (size_t)std::max<int64_t>(capacity, 1)));
        CUDA_TRY(&s, cudaMalloc(&d_oj, sizeof(int) * (size_t)std::max<int64_t>(capacity, 1)));
        CUDA_TRY(&s, cudaMalloc(&d_counter, sizeof(unsigned long long)));
        CUDA_TRY(&s, cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), s.stream));
        if (n_x > 0 && use_tiles(s)) {
            const int list_len = s.tiles.list(KS);
            const int cap = tile_capacity<T, CT>(s.tiles.smem_budget, list_len, KS);
            const size_t smem = tile_smem_bytes<T, CT>(cap, list_len, KS);
            int rc2 = set_smem(s, k_pairs_tiles<KS, ND, T, CT>, 227 * 1024);
            if (rc2) return rc2;
            const int max_tiles = x_fluid ? s.tiles.max_ftiles : s.tiles.max_wtiles;
            const int *n_tiles = (x_fluid ? s.tiles.d_frow_tile_start : s.tiles.d_wrow_tile_start) + s.tiles.nrows;
            LAUNCH(s, (k_tile_ranges<ND>), cdiv((int64_t)max_tiles * 32, 256), 256, 0, s.ncell[0], s.ncell[1],
                   s.xsplit, n_tiles, x_fluid ? s.tiles.d_ftile_desc : s.tiles.d_wtile_desc,
                   x_fluid ? s.d_fcell_start : s.d_wcell_start, y_fluid ? s.d_fcell_start : s.d_wcell_start,
                   (const int *)nullptr, s.tiles.d_ptile_rng, 9, s.tiles.d_ptile_ext);
            LAUNCH(s, (k_pairs_tiles<KS, ND, T, CT>), max_tiles, KS * TILE_TB,
                   smem, g, n_tiles,
                   x_fluid ? s.tiles.d_ftile_desc : s.tiles.d_wtile_desc, s.tiles.d_ptile_ext, s.tiles.d_ptile_rng,
                   (const V4<CT> *)(x_fluid ? s.d_A : s.d_Aw), x_fluid ? s.d_perm_f : s.d_perm_w,
                   y_fluid ? s.d_fcell_start : s.d_wcell_start,
                   (const V4<CT> *)(y_fluid ? s.d_A : s.d_Aw), y_fluid ? s.d_perm_f : s.d_perm_w, r2,
                   (long long)capacity, d_oi, d_oj, d_counter, cap, list_len,
                   (const V4<float> *)(y_fluid ? s.d_Ff : s.d_Fw));
        } else if (n_x > 0)
            LAUNCH(s, (k_pairs<ND, T, CT>), cdiv(n_x, 128), 128, 0, n_x,
                   (x_fluid ? s.d_fcell_start : s.d_wcell_start) + s.ncells, g,
                   (const V4<CT> *)(x_fluid ? s.d_A : s.d_Aw), x_fluid ? s.d_perm_f : s.d_perm_w,
                   y_fluid ? s.d_fcell_start : s.d_wcell_start,
                   (const V4<CT> *)(y_fluid ? s.d_A : s.d_Aw), y_fluid ? s.d_perm_f : s.d_perm_w, r2,
                   (long long)capacity, d_oi, d_oj, d_counter);
        unsigned long long h_count = 0;
        CUDA_TRY(&s, cudaMemcpyAsync(&h_count, d_counter, sizeof(h_count), cudaMemcpyDeviceToHost, s.stream));
        CUDA_TRY(&s, cudaStreamSynchronize(s.stream));
        *count = (int64_t)h_count;
        int64_t ncopy = std::min<int64_t>((int64_t)h_count, capacity);
        if (ncopy > 0) {
            CUDA_TRY(&s, cudaMemcpy(out_i, d_oi, sizeof(int) * (size_t)ncopy, cudaMemcpyDeviceToHost));
            CUDA_TRY(&s, cudaMemcpy(out_j, d_oj, sizeof(int) * (size_t)ncopy, cudaMemcpyDeviceToHost));
        }
        cudaFree(d_oi); cudaFree(d_oj); cudaFree(d_counter); cudaFree(d_vzero);
        if (d_utmp) cudaFree(d_utmp);
        CUDA_TRY(&s, cudaMemcpy(s.h_flags, s.d_flags, sizeof(int), cudaMemcpyDeviceToHost));
        if (*s.h_flags & 1)
            return fail(&s, TPB_ERR_OUT_OF_BOUNDS, "particle coordinates are NaN or outside the bounding box");
        if ((int64_t)h_count > capacity) return fail(&s, TPB_ERR_CAPACITY, "pair buffer too small");
        return TPB_OK;
    }
};

// ---- entry points of one (ND, T, CT) combination (external linkage: defined in that unit)
#define TPB_CAT_(a, b) a##b
#define TPB_CAT(a, b) TPB_CAT_(a, b)
#define TPB_DECLARE_ENTRIES(TAG)                                                                         \
    int TPB_CAT(init_wall_, TAG)(Semi &s);                                                               \
    int TPB_CAT(init_structure_, TAG)(Semi &s);                                                          \
    int TPB_CAT(kick_, TAG)(Semi &s, void *dv, const void *v, const void *u);                            \
    int TPB_CAT(drift_, TAG)(Semi &s, void *du, const void *v, const void *u);                           \
    int TPB_CAT(get_field_, TAG)(Semi &s, int sys, int field, void *out, int64_t n);                     \
    int TPB_CAT(pairs_, TAG)(Semi &s, int sys, int nb, const void *u, int64_t cap, int32_t *oi, int32_t *oj, \
                             int64_t *cnt);                                                              \
    int TPB_CAT(max_speed2_, TAG)(Semi &s, const void *v, void *out_bits);                               \
    int TPB_CAT(struct_force_, TAG)(Semi &s, void *out, const void *v, const void *u);                   \
    int TPB_CAT(kick_struct_, TAG)(Semi &s, void *dv, const void *v, const void *u, const void *dv_const); \
    int TPB_CAT(sort_system_, TAG)(Semi &s, void *v, void *u);                                           \
    int TPB_CAT(reinit_density_, TAG)(Semi &s, void *v, const void *u);
#define TPB_DEFINE_ENTRIES(TAG, ND, T, CT)                                                               \
    int TPB_CAT(init_wall_, TAG)(Semi &s) { return Ops<ND, T, CT>::init_wall(s); }                       \
    int TPB_CAT(init_structure_, TAG)(Semi &s) { return Ops<ND, T, CT>::init_structure(s); }             \
    int TPB_CAT(kick_, TAG)(Semi &s, void *dv, const void *v, const void *u)                             \
    {                                                                                                    \
        return Ops<ND, T, CT>::kick(s, dv, v, u);                                                        \
    }                                                                                                    \
    int TPB_CAT(drift_, TAG)(Semi &s, void *du, const void *v, const void *u)                            \
    {                                                                                                    \
        return Ops<ND, T, CT>::drift(s, du, v, u);                                                       \
    }                                                                                                    \
    int TPB_CAT(get_field_, TAG)(Semi &s, int sys, int field, void *out, int64_t n)                      \
    {                                                                                                    \
        return Ops<ND, T, CT>::get_field(s, sys, field, out, n);                                         \
    }                                                                                                    \
    int TPB_CAT(pairs_, TAG)(Semi &s, int sys, int nb, const void *u, int64_t cap, int32_t *oi, int32_t *oj, \
                             int64_t *cnt)                                                               \
    {                                                                                                    \
        return Ops<ND, T, CT>::neighbor_pairs(s, sys, nb, u, cap, oi, oj, cnt);                          \
    }                                                                                                    \
    int TPB_CAT(max_speed2_, TAG)(Semi &s, const void *v, void *out_bits)                                \
    {                                                                                                    \
        return Ops<ND, T, CT>::max_speed2(s, (const T *)v + ode_layout(s).off_v_f, s.n_tgt,              \
                                          (unsigned long long *)out_bits);                               \
    }                                                                                                    \
    int TPB_CAT(struct_force_, TAG)(Semi &s, void *out, const void *v, const void *u)                    \
    {                                                                                                    \
        return Ops<ND, T, CT>::structure_fluid_force(s, out, v, u);                                      \
    }                                                                                                    \
    int TPB_CAT(kick_struct_, TAG)(Semi &s, void *dv, const void *v, const void *u, const void *dv_const) \
    {                                                                                                    \
        return Ops<ND, T, CT>::kick_structure(s, dv, v, u, dv_const);                                    \
    }                                                                                                    \
    int TPB_CAT(sort_system_, TAG)(Semi &s, void *v, void *u) { return Ops<ND, T, CT>::sort_system(s, v, u); } \
    int TPB_CAT(reinit_density_, TAG)(Semi &s, void *v, const void *u) { return Ops<ND, T, CT>::reinit_density(s, v, u); }
TPB_DECLARE_ENTRIES(2ff)
TPB_DECLARE_ENTRIES(3ff)
TPB_DECLARE_ENTRIES(2fd)
TPB_DECLARE_ENTRIES(3fd)
TPB_DECLARE_ENTRIES(2dd)
TPB_DECLARE_ENTRIES(3dd)

#if defined(TPB_TU_TAG)  // a kernel unit: -DTPB_TU_TAG=3ff -DTPB_TU_ND=3 -DTPB_TU_T=float -DTPB_TU_CT=float
TPB_DEFINE_ENTRIES(TPB_TU_TAG, TPB_TU_ND, TPB_TU_T, TPB_TU_CT)
#elif defined(TPB_SINGLE_TU)  // everything in one unit (slow to compile)
TPB_DEFINE_ENTRIES(2ff, 2, float, float)
TPB_DEFINE_ENTRIES(3ff, 3, float, float)
TPB_DEFINE_ENTRIES(2fd, 2, float, double)
TPB_DEFINE_ENTRIES(3fd, 3, float, double)
TPB_DEFINE_ENTRIES(2dd, 2, double, double)
TPB_DEFINE_ENTRIES(3dd, 3, double, double)
#endif

#ifndef TPB_TU_TAG  // the main unit: dispatch + C ABI
#define DISPATCH(S, NAME, ...)                                                                \
    do {                                                                                      \
        const int nd_ = (S).cfg.ndims, t_ = (S).cfg.eltype, ct_ = (S).cfg.coords_eltype;      \
        if (nd_ == 2 && t_ == TPB_F32 && ct_ == TPB_F32) return NAME##2ff(__VA_ARGS__);       \
        if (nd_ == 3 && t_ == TPB_F32 && ct_ == TPB_F32) return NAME##3ff(__VA_ARGS__);       \
        if (nd_ == 2 && t_ == TPB_F32 && ct_ == TPB_F64) return NAME##2fd(__VA_ARGS__);       \
        if (nd_ == 3 && t_ == TPB_F32 && ct_ == TPB_F64) return NAME##3fd(__VA_ARGS__);       \
        if (nd_ == 2 && t_ == TPB_F64 && ct_ == TPB_F64) return NAME##2dd(__VA_ARGS__);       \
        if (nd_ == 3 && t_ == TPB_F64 && ct_ == TPB_F64) return NAME##3dd(__VA_ARGS__);       \
        return fail(&(S), TPB_ERR_UNSUPPORTED, "unsupported ndims / eltype combination");     \
    } while (0)

static int dispatch_init_wall(Semi &s) { DISPATCH(s, init_wall_, s); }
static int dispatch_init_structure(Semi &s) { DISPATCH(s, init_structure_, s); }
static int dispatch_kick(Semi &s, void *dv, const void *v, const void *u) { DISPATCH(s, kick_, s, dv, v, u); }
static int dispatch_drift(Semi &s, void *du, const void *v, const void *u) { DISPATCH(s, drift_, s, du, v, u); }
static int dispatch_get_field(Semi &s, int sys, int field, void *out, int64_t n)
{
    DISPATCH(s, get_field_, s, sys, field, out, n);
}
static int dispatch_pairs(Semi &s, int sys, int nb, const void *u, int64_t cap, int32_t *oi, int32_t *oj,
                          int64_t *cnt)
{
    DISPATCH(s, pairs_, s, sys, nb, u, cap, oi, oj, cnt);
}
static int dispatch_max_speed2(Semi &s, const void *v, void *out_bits) { DISPATCH(s, max_speed2_, s, v, out_bits); }
static int dispatch_struct_force(Semi &s, void *out, const void *v, const void *u) { DISPATCH(s, struct_force_, s, out, v, u); }
static int dispatch_kick_struct(Semi &s, void *dv, const void *v, const void *u, const void *c)
{
    DISPATCH(s, kick_struct_, s, dv, v, u, c);
}

static int dispatch_sort_system(Semi &s, void *v, void *u) { DISPATCH(s, sort_system_, s, v, u); }
static int dispatch_reinit_density(Semi &s, void *v, const void *u) { DISPATCH(s, reinit_density_, s, v, u); }

static void free_device(Semi &s)
{
    void *ptrs[] = {s.d_mass_f, s.d_u, s.d_v, s.d_dv, s.d_du, s.d_key, s.d_slot, s.d_tmp_perm,
                    s.d_perm_f, s.d_count, s.d_fcell_start, s.d_wcell_start, s.d_block_sums, s.d_scan_status,
                    s.d_scan_ticket,
                    s.d_flags, s.d_A, s.d_B, s.d_P, s.d_Aw, s.d_Ww, s.d_volw, s.d_Vw, s.d_Pw, s.d_perm_w,
                    s.d_scratch, s.d_Ff, s.d_Fw, s.d_vmax2, s.d_adapt, s.d_x0_s, s.d_xcur_s, s.d_mass_s, s.d_rho_s, s.d_hydro_s,
                    s.d_L_s, s.d_F_s, s.d_pk1_s, s.d_As, s.d_Bs, s.d_nbr_start, s.d_nbr, s.d_scell_start, s.d_sperm,
                    s.d_Ps, s.d_p_s, s.d_rhoh_s, s.d_xcl_s, s.d_vcl_s, s.d_acl_s, s.d_sort_u, s.d_sort_v, s.d_mat_s};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    tiles_free(s.tiles);
    for (cudaEvent_t e : s.prof_events) cudaEventDestroy(e);
    s.prof_events.clear();
    if (s.h_flags) cudaFreeHost(s.h_flags);
    if (s.h_vmax2) cudaFreeHost(s.h_vmax2);
    if (s.own_stream) cudaStreamDestroy(s.own_stream);
}

static double as_double(const unsigned char *p, int eltype, size_t i)
{
    return eltype == TPB_F64 ? ((const double *)p)[i] : (double)((const float *)p)[i];
}
#endif  // !TPB_TU_TAG

}  // namespace tpbhost

#ifndef TPB_TU_TAG
using namespace tpbhost;

// =========================================================================================
// C ABI
// =========================================================================================
extern "C" {

const char *tpb_version(void) { return "tpb200 0.1 (sm_100a)"; }

const char *tpb_last_error(tpb_semi_t semi)
{
    return semi ? ((Semi *)semi)->err.c_str() : g_create_error.c_str();
}

int32_t tpb_create(const tpb_config *config, tpb_semi_t *out)
{
    if (!config || !out) return fail(nullptr, TPB_ERR_INVALID_ARGUMENT, "null argument");
    if (config->struct_size != (int32_t)sizeof(tpb_config))
        return fail(nullptr, TPB_ERR_INVALID_ARGUMENT, "tpb_config.struct_size mismatch");
    if (config->ndims != 2 && config->ndims != 3)
        return fail(nullptr, TPB_ERR_INVALID_ARGUMENT, "ndims must be 2 or 3");
    if ((config->eltype != TPB_F32 && config->eltype != TPB_F64) ||
        (config->coords_eltype != TPB_F32 && config->coords_eltype != TPB_F64))
        return fail(nullptr, TPB_ERR_INVALID_ARGUMENT, "eltype must be TPB_F32 or TPB_F64");
    if (config->eltype == TPB_F64 && config->coords_eltype == TPB_F32)
        return fail(nullptr, TPB_ERR_INVALID_ARGUMENT,
                    "coordinates eltype must not be narrower than the system eltype");
    if (config->ode_memory != TPB_MEM_HOST && config->ode_memory != TPB_MEM_DEVICE)
        return fail(nullptr, TPB_ERR_INVALID_ARGUMENT, "ode_memory must be TPB_MEM_HOST or TPB_MEM_DEVICE");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, TPB_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (config->device < 0 || config->device >= ndev)
        return fail(nullptr, TPB_ERR_INVALID_ARGUMENT, "device ordinal out of range");
    e = cudaSetDevice(config->device);
    if (e != cudaSuccess) return fail(nullptr, TPB_ERR_CUDA, cudaGetErrorString(e));
    Semi *s = new Semi();
    s->cfg = *config;
    e = cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete s;
        return fail(nullptr, TPB_ERR_CUDA, cudaGetErrorString(e));
    }
    s->stream = s->own_stream;
    if (const char *z = getenv("TPB_HOST_ZEROCOPY")) s->host_zero_copy = atoi(z) != 0;
    *out = (tpb_semi_t)s;
    return TPB_OK;
}

int32_t tpb_destroy(tpb_semi_t semi)
{
    if (!semi) return TPB_OK;
    Semi *s = (Semi *)semi;
    cudaSetDevice(s->cfg.device);
    cudaStreamSynchronize(s->stream);
    free_device(*s);
    delete s;
    return TPB_OK;
}

int32_t tpb_add_fluid_system(tpb_semi_t semi, const tpb_fluid_params *p, int64_t n, const void *mass,
                             int32_t *system_index)
{
    Semi *s = (Semi *)semi;
    if (!s || !p || (n > 0 && !mass)) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null argument");
    if (s->ready) return fail(s, TPB_ERR_STATE, "systems must be added before tpb_semidiscretize");
    if (p->struct_size != (int32_t)sizeof(tpb_fluid_params))
        return fail(s, TPB_ERR_INVALID_ARGUMENT, "tpb_fluid_params.struct_size mismatch");
    if (s->fluid_index >= 0)
        return fail(s, TPB_ERR_UNSUPPORTED, "only one fluid system per semidiscretization is supported");
    if (p->kernel < TPB_KERNEL_WENDLAND_C2 || p->kernel > TPB_KERNEL_SCHOENBERG_QUINTIC)
        return fail(s, TPB_ERR_INVALID_ARGUMENT, "unknown smoothing kernel");
    if (p->density_calculator != TPB_DENSITY_CONTINUITY && p->density_calculator != TPB_DENSITY_SUMMATION)
        return fail(s, TPB_ERR_INVALID_ARGUMENT, "unknown density calculator");
    if (!(p->smoothing_length > 0) || !(p->sound_speed > 0) || !(p->reference_density > 0) || p->exponent == 0)
        return fail(s, TPB_ERR_INVALID_ARGUMENT, "smoothing_length, sound_speed, reference_density must be positive");
    if (p->has_diffusion && p->density_calculator == TPB_DENSITY_SUMMATION)
        return fail(s, TPB_ERR_INVALID_ARGUMENT, "density diffusion requires ContinuityDensity");
    if (p->adaptive_sound_speed &&
        (!(p->mach_number_target > 0) || !(p->min_sound_speed > 0) || !(p->max_sound_speed >= p->min_sound_speed)))
        return fail(s, TPB_ERR_INVALID_ARGUMENT, "StateEquationAdaptiveCole: mach_number_target > 0 and "
                                                 "0 < min_sound_speed <= max_sound_speed are required");
    if (n < 0 || n > 0x7fffffff / 4) return fail(s, TPB_ERR_INVALID_ARGUMENT, "particle count out of range");
    s->fp = *p;
    s->n_f = n; s->n_act = n; s->n_tgt = n;
    s->h_mass_f.assign((const unsigned char *)mass, (const unsigned char *)mass + tsize(s->cfg.eltype) * (size_t)n);
    s->fluid_index = s->n_systems++;
    if (system_index) *system_index = s->fluid_index;
    return TPB_OK;
}

int32_t tpb_add_wall_system(tpb_semi_t semi, const tpb_wall_params *p, int64_t n, const void *coords,
                            const void *hydrodynamic_mass, const void *initial_density,
                            int32_t *system_index)
{
    Semi *s = (Semi *)semi;
    if (!s || !p || (n > 0 && (!coords || !hydrodynamic_mass || !initial_density)))
        return fail(s, TPB_ERR_INVALID_ARGUMENT, "null argument");
    if (s->ready) return fail(s, TPB_ERR_STATE, "systems must be added before tpb_semidiscretize");
    if (p->struct_size != (int32_t)sizeof(tpb_wall_params))
        return fail(s, TPB_ERR_INVALID_ARGUMENT, "tpb_wall_params.struct_size mismatch");
    if (s->wall_index >= 0) {
        // a further WallBoundarySystem joins the wall set if its boundary model is the same
        const tpb_wall_params &a = s->wp, &b = *p;
        const bool same = a.kernel == b.kernel && a.clip_negative_pressure == b.clip_negative_pressure &&
                          a.sound_speed_from_fluid == b.sound_speed_from_fluid && a.smoothing_length == b.smoothing_length &&
                          a.sound_speed == b.sound_speed && a.exponent == b.exponent &&
                          a.reference_density == b.reference_density && a.background_pressure == b.background_pressure &&
                          a.pressure_offset == b.pressure_offset && a.has_viscosity == b.has_viscosity &&
                          a.density_calculator == b.density_calculator && a.alpha == b.alpha && a.beta == b.beta &&
                          a.epsilon == b.epsilon && a.eos_clip_negative_pressure == b.eos_clip_negative_pressure;
        if (!same)
            return fail(s, TPB_ERR_UNSUPPORTED, "wall systems with different boundary models are not supported "
                                                "(kernel, smoothing length, state equation, viscosity, density calculator)");
        if (a.density_calculator == TPB_WALL_DENSITY_CONTINUITY)
            return fail(s, TPB_ERR_UNSUPPORTED, "only one wall system with ContinuityDensity is supported");
    }
    if (p->kernel < TPB_KERNEL_WENDLAND_C2 || p->kernel > TPB_KERNEL_SCHOENBERG_QUINTIC)
        return fail(s, TPB_ERR_INVALID_ARGUMENT, "unknown smoothing kernel");
    if (!(p->smoothing_length > 0) || !(p->sound_speed > 0) || !(p->reference_density > 0) || p->exponent == 0)
        return fail(s, TPB_ERR_INVALID_ARGUMENT, "smoothing_length, sound_speed, reference_density must be positive");
    if (n < 0 || n > 0x7fffffff / 4) return fail(s, TPB_ERR_INVALID_ARGUMENT, "particle count out of range");
    if (p->has_viscosity < TPB_VISCOSITY_NONE || p->has_viscosity > TPB_VISCOSITY_ADAMI)
        return fail(s, TPB_ERR_INVALID_ARGUMENT, "unknown wall viscosity model");
    if (p->density_calculator != TPB_WALL_DENSITY_ADAMI && p->density_calculator != TPB_WALL_DENSITY_CONTINUITY)
        return fail(s, TPB_ERR_INVALID_ARGUMENT, "unknown wall density calculator");
    if (p->density_calculator == TPB_WALL_DENSITY_CONTINUITY && p->has_viscosity != TPB_VISCOSITY_NONE)
        return fail(s, TPB_ERR_UNSUPPORTED, "a no-slip wall with ContinuityDensity is not supported");
    if (s->n_w + n > 0x7fffffff / 4) return fail(s, TPB_ERR_INVALID_ARGUMENT, "particle count out of range");
    s->wp = *p;
    const size_t ts = tsize(s->cfg.eltype), cs = tsize(s->cfg.coords_eltype);
    auto append = [](std::vector<unsigned char> &v, const void *src, size_t bytes) {
        v.insert(v.end(), (const unsigned char *)src, (const unsigned char *)src + bytes);
    };
    append(s->h_coords_w, coords, cs * s->cfg.ndims * (size_t)n);
    append(s->h_mass_w, hydrodynamic_mass, ts * (size_t)n);
    append(s->h_dens_w, initial_density, ts * (size_t)n);
    Semi::WallPart part;
    part.index = s->n_systems++;
    part.first = s->n_w;
    part.n = n;
    s->wall_parts.push_back(part);
    s->n_w += n;
    if (s->wall_index < 0) s->wall_index = part.index;
    if (system_index) *system_index = part.index;
    return TPB_OK;
}

int32_t tpb_add_structure_system(tpb_semi_t semi, const tpb_structure_params *p, int64_t n, int64_t n_integrated,
                                 const void *initial_coords, const void *mass, const void *material_density,
                                 const void *hydrodynamic_mass, int32_t *system_index)
{
    Semi *s = (Semi *)semi;
    if (!s || !p || (n > 0 && (!initial_coords || !mass || !material_density)))
        return fail(s, TPB_ERR_INVALID_ARGUMENT, "null argument");
    if (s->ready) return fail(s, TPB_ERR_STATE, "systems must be added before tpb_semidiscretize");
    if (p->struct_size != (int32_t)sizeof(tpb_structure_params))
        return fail(s, TPB_ERR_INVALID_ARGUMENT, "tpb_structure_params.struct_size mismatch");
    if (s->struct_index >= 0)
        return fail(s, TPB_ERR_UNSUPPORTED, "only one structure system per semidiscretization is supported");
    if (p->kernel < TPB_KERNEL_WENDLAND_C2 || p->kernel > TPB_KERNEL_SCHOENBERG_QUINTIC)
        return fail(s, TPB_ERR_INVALID_ARGUMENT, "unknown smoothing kernel");
    if (!(p->smoothing_length > 0) || !(p->young_modulus > 0))
        return fail(s, TPB_ERR_INVALID_ARGUMENT, "smoothing_length and young_modulus must be positive");
    if (p->boundary_model != TPB_BOUNDARY_NONE && p->boundary_model != TPB_BOUNDARY_MONAGHAN_KAJTAR &&
        p->boundary_model != TPB_BOUNDARY_DUMMY_PARTICLES)
        return fail(s, TPB_ERR_UNSUPPORTED, "structure boundary model: BoundaryModelMonaghanKajtar, "
                                            "BoundaryModelDummyParticles (Adami) or none");
    if (p->boundary_model == TPB_BOUNDARY_DUMMY_PARTICLES &&
        (!hydrodynamic_mass || !(p->bm_smoothing_length > 0) || !(p->bm_sound_speed > 0) ||
         !(p->bm_reference_density > 0) || p->bm_exponent == 0 || p->bm_kernel < TPB_KERNEL_WENDLAND_C2 ||
         p->bm_kernel > TPB_KERNEL_SCHOENBERG_QUINTIC))
        return fail(s, TPB_ERR_INVALID_ARGUMENT, "BoundaryModelDummyParticles on the structure needs kernel, smoothing "
                                                 "length, state equation and the hydrodynamic masses");
    if (p->boundary_model == TPB_BOUNDARY_MONAGHAN_KAJTAR &&
        (!hydrodynamic_mass || !(p->mk_K > 0) || !(p->mk_beta > 0) || !(p->mk_spacing > 0)))
        return fail(s, TPB_ERR_INVALID_ARGUMENT, "BoundaryModelMonaghanKajtar needs K, beta, spacing > 0 and the hydrodynamic masses");
    if (n < 0 || n > 0x7fffffff / 16 || n_integrated < 0 || n_integrated > n)
        return fail(s, TPB_ERR_INVALID_ARGUMENT, "particle count out of range");
    s->sp = *p;
    s->n_s = n;
    s->n_s_int = n_integrated;
    const size_t ts = tsize(s->cfg.eltype), cs = tsize(s->cfg.coords_eltype);
    auto bytes = [](const void *q, size_t len) { return std::vector<unsigned char>((const unsigned char *)q, (const unsigned char *)q + len); };
    s->h_x0_s = bytes(initial_coords, cs * s->cfg.ndims * (size_t)n);
    s->h_mass_s = bytes(mass, ts * (size_t)n);
    s->h_rho_s = bytes(material_density, ts * (size_t)n);
    if (hydrodynamic_mass) s->h_hydro_s = bytes(hydrodynamic_mass, ts * (size_t)n);
    s->struct_index = s->n_systems++;
    if (system_index) *system_index = s->struct_index;
    return TPB_OK;
}

int32_t tpb_set_interaction(tpb_semi_t semi, int32_t system, int32_t neighbor, int32_t enabled)
{
    Semi *s = (Semi *)semi;
    if (!s) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null handle");
    if (system < 0 || system >= s->n_systems || neighbor < 0 || neighbor >= s->n_systems)
        return fail(s, TPB_ERR_INVALID_ARGUMENT, "system index out of range");
    if (system == s->struct_index || neighbor == s->struct_index) {
        // structure <-> wall pairs never interact (rhs.jl:111-119); the rest is switchable
        if (system == s->struct_index && neighbor == s->struct_index) {
            if (!enabled) return fail(s, TPB_ERR_UNSUPPORTED, "the structure's self-interaction cannot be switched off");
        } else if (system == s->struct_index && neighbor == s->fluid_index) s->struct_fluid[0] = enabled ? 1 : 0;
        else if (system == s->fluid_index && neighbor == s->struct_index) s->struct_fluid[1] = enabled ? 1 : 0;
        return TPB_OK;
    }
    // internal matrix is indexed [fluid=0|wall=1]; the walls of the set must agree (tpb_semidiscretize)
    int a = system == s->fluid_index ? 0 : 1, b = neighbor == s->fluid_index ? 0 : 1;
    if (a == 1 && b == 1) return TPB_OK;  // wall <-> wall: never interact (wall_boundary/rhs.jl:2-8)
    for (Semi::WallPart &w : s->wall_parts) {
        if (a == 0 && b == 1 && w.index == neighbor) w.fluid_from_wall = enabled ? 1 : 0;
        if (a == 1 && b == 0 && w.index == system) w.wall_from_fluid = enabled ? 1 : 0;
    }
    if (a == 0 && b == 0) s->interaction[0][0] = enabled ? 1 : 0;
    return TPB_OK;
}

int32_t tpb_semidiscretize(tpb_semi_t semi, const void *u0_ode)
{
    Semi *s = (Semi *)semi;
    if (!s) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null handle");
    if (s->ready) return fail(s, TPB_ERR_STATE, "tpb_semidiscretize was already called");
    if (s->fluid_index < 0) return fail(s, TPB_ERR_INVALID_ARGUMENT, "a fluid system is required");
    for (const Semi::WallPart &w : s->wall_parts) {
        const Semi::WallPart &w0 = s->wall_parts.front();
        if (w.fluid_from_wall != w0.fluid_from_wall || w.wall_from_fluid != w0.wall_from_fluid)
            return fail(s, TPB_ERR_UNSUPPORTED, "the wall systems of one semidiscretization must share their "
                                                "interaction switches with the fluid");
        s->interaction[0][1] = w0.fluid_from_wall;
        s->interaction[1][0] = w0.wall_from_fluid;
    }
    if (s->wall_index >= 0 && s->wp.has_viscosity >= TPB_VISCOSITY_MORRIS && s->fp.has_viscosity == TPB_VISCOSITY_NONE)
        return fail(s, TPB_ERR_INVALID_ARGUMENT,
                    "a ViscosityMorris / ViscosityAdami wall needs a fluid viscosity model: "
                    "kinematic_viscosity(fluid, nothing, ...) has no method in the reference");
    CUDA_TRY(s, cudaSetDevice(s->cfg.device));
    const int nd = s->cfg.ndims;
    const size_t ts = tsize(s->cfg.eltype), cs = tsize(s->cfg.coords_eltype);

    // search radii in T: compact_support (smoothing_kernels.jl:215, :310, :383, :400)
    auto radius = [&](int kernel, double h) {
        return s->cfg.eltype == TPB_F64 ? (double)make_kernel_const<double>(kernel, nd, h).support
                                        : (double)make_kernel_const<float>(kernel, nd, h).support;
    };
    double R = radius(s->fp.kernel, s->fp.smoothing_length);
    if (s->wall_index >= 0) R = std::max(R, radius(s->wp.kernel, s->wp.smoothing_length));
    if (s->struct_index >= 0 && s->sp.boundary_model == TPB_BOUNDARY_DUMMY_PARTICLES)
        R = std::max(R, radius(s->sp.bm_kernel, s->sp.bm_smoothing_length));

    // bounding box
    double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    if (s->cfg.has_bounds) {
        for (int d = 0; d < nd; ++d) {
            lo[d] = s->cfg.min_corner[d];
            hi[d] = s->cfg.max_corner[d];
            if (!(hi[d] > lo[d])) return fail(s, TPB_ERR_INVALID_ARGUMENT, "max_corner must exceed min_corner");
        }
    } else {
        bool first = true;
        auto extend = [&](const unsigned char *p, int64_t n) {
            for (int64_t i = 0; i < n; ++i)
                for (int d = 0; d < nd; ++d) {
                    double v = as_double(p, s->cfg.coords_eltype, (size_t)i * nd + d);
                    if (first && d == 0) { for (int e = 0; e < nd; ++e) lo[e] = hi[e] = as_double(p, s->cfg.coords_eltype, (size_t)i * nd + e); first = false; }
                    lo[d] = std::min(lo[d], v);
                    hi[d] = std::max(hi[d], v);
                }
        };
        if (u0_ode && s->n_f > 0)
            extend((const unsigned char *)u0_ode + cs * (size_t)ode_layout(*s).off_u_f, s->n_f);
        if (s->n_w > 0) extend(s->h_coords_w.data(), s->n_w);
        if (s->n_s > 0) extend(s->h_x0_s.data(), s->n_s);
        if (first) return fail(s, TPB_ERR_INVALID_ARGUMENT, "no coordinates to derive the bounding box from");
        for (int d = 0; d < nd; ++d) { lo[d] -= 2 * R; hi[d] += 2 * R; }
    }
    // cell size = R (1 + margin): the margin absorbs the rounding of the cT cell-coordinate
    // computation so that |x_i - x_j| <= R always implies adjacent cells
    double eps_ct = s->cfg.coords_eltype == TPB_F64 ? 2.220446049250313e-16 : 1.1920929e-07;
    double max_cells = 4;
    for (int d = 0; d < nd; ++d) max_cells = std::max(max_cells, (hi[d] - lo[d]) / R + 4);
    double max_abs = 0;
    for (int d = 0; d < nd; ++d) max_abs = std::max(max_abs, std::max(std::fabs(lo[d]), std::fabs(hi[d])) / R + 4);
    double margin = std::max(1e-6, 16 * eps_ct * std::max(max_cells, max_abs));
    s->cell_size = R * (1 + margin);
    s->ncells = 1;
    for (int d = 0; d < 3; ++d) {
        if (d < nd) {
            s->lo[d] = lo[d];
            s->hi[d] = hi[d];
            s->origin[d] = lo[d] - 1.001 * s->cell_size;
            s->ncell[d] = (int)std::floor((hi[d] - s->origin[d]) / s->cell_size) + 3;
        } else {
            s->lo[d] = s->hi[d] = s->origin[d] = 0;
            s->ncell[d] = 1;
        }
        s->ncells *= s->ncell[d];
    }
    // Float32 filter copy of Float64 positions (FilterRef): reference point = lower corner of
    // the bounding box, pad = 2 x the worst-case distance error 2 sqrt(3) 2^-24 L
    {
        double extent = 0;
        for (int d = 0; d < nd; ++d) {
            s->filter_ref[d] = lo[d];
            extent = std::max(extent, hi[d] - lo[d]);
        }
        s->filter_pad = (float)(4.0 * std::sqrt(3.0) * std::ldexp(1.0, -24) * extent * 1.01);
    }
    // Rows (the fastest cell index; one row = one contiguous run of sorted records, the unit the
    // tile sweeps stage and split into tiles) run along the longest side of the bounding box: long
    // rows give full tiles and little staging overhead, and a slab of a decomposed domain is
    // thin along the decomposition axis.  Ties keep x.
    {
        int ax = 0;
        for (int d = 1; d < nd; ++d)
            if (s->hi[d] - s->lo[d] > (s->hi[ax] - s->lo[ax]) * (1 + 1e-9)) ax = d;
        if (const char *e = getenv("TPB_ROW_AXIS")) ax = std::min(std::max(atoi(e), 0), nd - 1);
        s->row_axis = ax;
        std::swap(s->lo[0], s->lo[ax]);
        std::swap(s->hi[0], s->hi[ax]);
        std::swap(s->origin[0], s->origin[ax]);
        std::swap(s->ncell[0], s->ncell[ax]);
    }
    // x-split: finer cells along x (the fastest index) let the tile sweep clip each row to the
    // chord of the search sphere; bounded so that the cell table stays scannable
    {
        int want = 1;  // lock-step sweeps gain nothing from finer x cells (experiments)
        if (const char *e = getenv("TPB_XSPLIT")) want = std::max(1, atoi(e));
        const int64_t limit = (int64_t)SCAN_TILE * SCAN_TILE - 8;
        while (want > 1 && s->ncells * want > limit) --want;
        s->xsplit = want;
        s->ncell[0] *= want;
        s->ncells *= want;
        double max_coord = 0;
        for (int d = 0; d < nd; ++d)
            max_coord = std::max(max_coord, std::max(std::fabs(s->origin[d]), std::fabs(hi[d]) + 2 * s->cell_size));
        s->gap_tol = 16 * eps_ct * max_coord;
    }
    if (s->ncells + 4 > (int64_t)SCAN_TILE * SCAN_TILE || s->ncells > 0x7ffffff0)
        return fail(s, TPB_ERR_UNSUPPORTED, "bounding box / search radius gives too many cells");

    // device buffers
    const size_t nf = (size_t)std::max<int64_t>(s->n_f, 1), nw = (size_t)std::max<int64_t>(s->n_w, 1);
    const size_t ns = (size_t)std::max<int64_t>(s->n_s, 1);
    const size_t nmax = std::max(std::max(nf, nw), ns);
    const int nvars = s->fp.density_calculator == TPB_DENSITY_SUMMATION ? nd : nd + 1;
    CUDA_TRY(s, cudaMalloc(&s->d_mass_f, ts * nf));
    CUDA_TRY(s, cudaMemcpy(s->d_mass_f, s->h_mass_f.data(), ts * (size_t)s->n_f, cudaMemcpyHostToDevice));
    if (s->cfg.ode_memory == TPB_MEM_HOST) {
        // staging copies of the whole ODE vectors (fluid + structure entries)
        const size_t extra = s->struct_index >= 0 ? (size_t)nd * (size_t)s->n_s_int : 0;
        const size_t extra_w = wall_integrates_density(*s) ? (size_t)s->n_w : 0;  // the wall's density rows
        CUDA_TRY(s, cudaMalloc(&s->d_u, cs * (nd * nf + extra)));
        CUDA_TRY(s, cudaMalloc(&s->d_v, ts * (nvars * nf + extra + extra_w)));
        CUDA_TRY(s, cudaMalloc(&s->d_dv, ts * (nvars * nf + extra + extra_w)));
        CUDA_TRY(s, cudaMalloc(&s->d_du, cs * (nd * nf + extra)));
    }
    if (s->struct_index >= 0) {
        CUDA_TRY(s, cudaMalloc(&s->d_x0_s, cs * nd * ns));
        CUDA_TRY(s, cudaMalloc(&s->d_xcur_s, cs * nd * ns));
        CUDA_TRY(s, cudaMalloc(&s->d_xcl_s, cs * nd * ns));
        {
            const size_t ncl = (size_t)(s->n_s - s->n_s_int) + 8;
            CUDA_TRY(s, cudaMalloc(&s->d_vcl_s, ts * nd * ncl));
            CUDA_TRY(s, cudaMalloc(&s->d_acl_s, ts * nd * ncl));
            CUDA_TRY(s, cudaMemset(s->d_vcl_s, 0, ts * nd * ncl));
            CUDA_TRY(s, cudaMemset(s->d_acl_s, 0, ts * nd * ncl));
        }
        CUDA_TRY(s, cudaMalloc(&s->d_mass_s, ts * ns));
        CUDA_TRY(s, cudaMalloc(&s->d_rho_s, ts * ns));
        CUDA_TRY(s, cudaMalloc(&s->d_hydro_s, ts * ns));
        CUDA_TRY(s, cudaMalloc(&s->d_L_s, ts * nd * nd * ns));
        CUDA_TRY(s, cudaMalloc(&s->d_F_s, ts * nd * nd * ns));
        CUDA_TRY(s, cudaMalloc(&s->d_pk1_s, ts * nd * nd * ns));
        CUDA_TRY(s, cudaMemset(s->d_F_s, 0, ts * nd * nd * ns));
        CUDA_TRY(s, cudaMemset(s->d_pk1_s, 0, ts * nd * nd * ns));
        CUDA_TRY(s, cudaMalloc(&s->d_As, 4 * cs * (ns + 8)));
        CUDA_TRY(s, cudaMalloc(&s->d_Bs, 4 * ts * (ns + 8)));
        if (s->sp.boundary_model == TPB_BOUNDARY_DUMMY_PARTICLES) {
            CUDA_TRY(s, cudaMalloc(&s->d_sperm, sizeof(int) * (ns + 8)));
            CUDA_TRY(s, cudaMalloc(&s->d_Ps, ts * (ns + 8)));
            CUDA_TRY(s, cudaMalloc(&s->d_p_s, ts * (ns + 8)));
            CUDA_TRY(s, cudaMalloc(&s->d_rhoh_s, ts * (ns + 8)));
            CUDA_TRY(s, cudaMemset(s->d_p_s, 0, ts * (ns + 8)));
            CUDA_TRY(s, cudaMemset(s->d_rhoh_s, 0, ts * (ns + 8)));
        }
        CUDA_TRY(s, cudaMalloc(&s->d_scell_start, sizeof(int) * (size_t)(s->ncells + 4)));
        CUDA_TRY(s, cudaMemset(s->d_scell_start, 0, sizeof(int) * (size_t)(s->ncells + 4)));
    }
    CUDA_TRY(s, cudaMalloc(&s->d_key, sizeof(int) * nmax));
    CUDA_TRY(s, cudaMalloc(&s->d_slot, sizeof(int) * nmax));
    CUDA_TRY(s, cudaMalloc(&s->d_tmp_perm, sizeof(int) * nmax));
    CUDA_TRY(s, cudaMalloc(&s->d_perm_f, sizeof(int) * nf));
    CUDA_TRY(s, cudaMalloc(&s->d_perm_w, sizeof(int) * nw));
    CUDA_TRY(s, cudaMalloc(&s->d_count, sizeof(int) * (size_t)(s->ncells + 4)));
    CUDA_TRY(s, cudaMalloc(&s->d_fcell_start, sizeof(int) * (size_t)(s->ncells + 4)));
    CUDA_TRY(s, cudaMalloc(&s->d_wcell_start, sizeof(int) * (size_t)(s->ncells + 4)));
    CUDA_TRY(s, cudaMemset(s->d_wcell_start, 0, sizeof(int) * (size_t)(s->ncells + 4)));
    CUDA_TRY(s, cudaMalloc(&s->d_block_sums, sizeof(int) * (size_t)SCAN_TILE));
    if (const char *e = getenv("TPB_SUBKEY")) s->subkey_mode = atoi(e);  // tuning: order inside a cell (tpb_nhs.cuh)
    {
        // k_scan_cells_tiles: whole cell rows per block
        const int n0 = s->ncell[0], nrows = s->ncell[1] * s->ncell[2];
        if (n0 <= CSCAN_TILE && !getenv("TPB_SCAN3")) {
            s->scan_rows_per_block = std::max(1, std::min(CSCAN_TILE / n0, CSCAN_MAX_ROWS));
            if (s->scan_rows_per_block >= 4) s->scan_rows_per_block &= ~3;  // 16-byte aligned block starts
            s->scan_blocks = cdiv(nrows, s->scan_rows_per_block);
            // (both running totals share one status word: 29 bits of particles, 24 bits of tiles)
            if ((int64_t)(s->n_f + 127) / 128 + nrows >= (1 << 24)) s->scan_blocks = 0;
            // Large, mostly empty grids (the end slab of the 8-GPU config-4 run: 29 M cells for 12.5 M
            // particles) are faster through the separate kernels: there the one-pass kernel takes 0.39 ms
            // against 0.20 ms for the three-kernel scan (measured, profiles/r2_p_end_slab.txt), while at
            // 4.5 M cells (10 M particles) it wins.
            const char *lim = getenv("TPB_SCAN_FUSED_MAX_CELLS");
            if (s->ncells > (lim ? atoll(lim) : 12000000ll)) s->scan_blocks = 0;
        }
        if (s->scan_blocks > 0) {
            CUDA_TRY(s, cudaMalloc(&s->d_scan_status, sizeof(unsigned long long) * 2 * (size_t)s->scan_blocks));
            CUDA_TRY(s, cudaMemset(s->d_scan_status, 0, sizeof(unsigned long long) * 2 * (size_t)s->scan_blocks));
            CUDA_TRY(s, cudaMalloc(&s->d_scan_ticket, sizeof(unsigned long long) * 2));
            CUDA_TRY(s, cudaMemset(s->d_scan_ticket, 0, sizeof(unsigned long long) * 2));
        }
    }
    CUDA_TRY(s, cudaMalloc(&s->d_flags, sizeof(int) * 4));
    CUDA_TRY(s, cudaMemset(s->d_flags, 0, sizeof(int) * 4));
    CUDA_TRY(s, cudaHostAlloc(&s->h_flags, sizeof(int) * 4, cudaHostAllocDefault));
    s->h_flags[0] = 0;
    // +1 record of slack so vector loads at the end of the last run stay inside the buffer
    CUDA_TRY(s, cudaMalloc(&s->d_A, 4 * cs * (nf + 8)));
    CUDA_TRY(s, cudaMalloc(&s->d_B, 4 * ts * (nf + 8)));
    CUDA_TRY(s, cudaMalloc(&s->d_P, ts * (nf + 8)));
    CUDA_TRY(s, cudaMalloc(&s->d_Aw, 4 * cs * (nw + 8)));
    CUDA_TRY(s, cudaMalloc(&s->d_Ww, 2 * ts * (nw + 8)));
    CUDA_TRY(s, cudaMalloc(&s->d_volw, ts * (nw + 8)));
    if (s->wp.has_viscosity) {
        CUDA_TRY(s, cudaMalloc(&s->d_Vw, 4 * ts * (nw + 8)));
        CUDA_TRY(s, cudaMemset(s->d_Vw, 0, 4 * ts * (nw + 8)));
        CUDA_TRY(s, cudaMalloc(&s->d_Pw, ts * (nw + 8)));
        CUDA_TRY(s, cudaMemset(s->d_Pw, 0, ts * (nw + 8)));
    }
    CUDA_TRY(s, cudaMalloc(&s->d_scratch, sizeof(double) * nmax));
    CUDA_TRY(s, cudaMalloc(&s->d_vmax2, sizeof(unsigned long long)));
    CUDA_TRY(s, cudaHostAlloc(&s->h_vmax2, sizeof(unsigned long long), cudaHostAllocDefault));
    if (s->cfg.eltype == TPB_F32 && s->cfg.coords_eltype == TPB_F64) {
        // (+8 records of slack like the other sorted arrays: the bulk copies of a tile round their runs up
        // to four records -- compute-sanitizer found the last run of a small system one record over)
        CUDA_TRY(s, cudaMalloc(&s->d_Ff, sizeof(float) * 4 * (nf + 8)));
        CUDA_TRY(s, cudaMalloc(&s->d_Fw, sizeof(float) * 4 * (nw + 8)));
        CUDA_TRY(s, cudaMemset(s->d_Ff, 0, sizeof(float) * 4 * (nf + 8)));
        CUDA_TRY(s, cudaMemset(s->d_Fw, 0, sizeof(float) * 4 * (nw + 8)));
    }
    CUDA_TRY(s, cudaMemset(s->d_A, 0, 4 * cs * (nf + 8)));
    CUDA_TRY(s, cudaMemset(s->d_B, 0, 4 * ts * (nf + 8)));
    CUDA_TRY(s, cudaMemset(s->d_P, 0, ts * (nf + 8)));
    CUDA_TRY(s, cudaMemset(s->d_Aw, 0, 4 * cs * (nw + 8)));
    CUDA_TRY(s, cudaMemset(s->d_Ww, 0, 2 * ts * (nw + 8)));
    int rc = tiles_alloc(s->tiles, s->ncell[1] * s->ncell[2], s->n_f, s->n_w);
    if (rc) return fail(s, TPB_ERR_CUDA, "tile scheduler allocation failed");

    s->stats.n_cells = s->ncells;
    if (s->wall_index >= 0) {
        rc = dispatch_init_wall(*s);
        if (rc) return rc;
        CUDA_TRY(s, cudaMemcpy(s->h_flags, s->d_flags, sizeof(int), cudaMemcpyDeviceToHost));
        if (*s->h_flags & 1)
            return fail(s, TPB_ERR_OUT_OF_BOUNDS, "wall particles outside the bounding box");
    }
    if (s->struct_index >= 0) {
        rc = dispatch_init_structure(*s);
        if (rc) return rc;
    }
    s->h_mass_f.clear(); s->h_mass_f.shrink_to_fit();
    s->h_mass_w.clear(); s->h_mass_w.shrink_to_fit();
    s->h_dens_w.clear(); s->h_dens_w.shrink_to_fit();
    s->ready = true;
    return TPB_OK;
}

int32_t tpb_ode_sizes(tpb_semi_t semi, int64_t *n_u, int64_t *n_v)
{
    Semi *s = (Semi *)semi;
    if (!s || s->fluid_index < 0) return fail(s, TPB_ERR_STATE, "no fluid system");
    const OdeLayout lay = ode_layout(*s);
    if (n_u) *n_u = lay.tot_u;
    if (n_v) *n_v = lay.tot_v;
    return TPB_OK;
}

int32_t tpb_system_range(tpb_semi_t semi, int32_t system, int64_t *u_first, int64_t *u_len,
                         int64_t *v_first, int64_t *v_len)
{
    Semi *s = (Semi *)semi;
    if (!s || system < 0 || system >= s->n_systems) return fail(s, TPB_ERR_INVALID_ARGUMENT, "system index out of range");
    const OdeLayout lay = ode_layout(*s);
    const int nd = s->cfg.ndims;
    const int nvars = s->fp.density_calculator == TPB_DENSITY_SUMMATION ? nd : nd + 1;
    int64_t uf = 0, ul = 0, vf = 0, vl = 0;
    if (system == s->fluid_index) {
        uf = lay.off_u_f, ul = s->n_act * nd, vf = lay.off_v_f, vl = s->n_act * nvars;
    } else if (system == s->struct_index) {
        uf = lay.off_u_s, ul = s->n_s_int * nd, vf = lay.off_v_s, vl = s->n_s_int * nd;
    } else {
        // the wall: no u entries (a zero-length range where they would start); v entries only with
        // ContinuityDensity dummy particles
        for (int other = 0; other < system; ++other) {
            if (other == s->fluid_index) uf += s->n_act * nd, vf += s->n_act * nvars;
            if (other == s->struct_index) uf += s->n_s_int * nd, vf += s->n_s_int * nd;
        }
        if (lay.len_v_w > 0) vf = lay.off_v_w, vl = lay.len_v_w;
    }
    if (u_first) *u_first = uf;
    if (u_len) *u_len = ul;
    if (v_first) *v_first = vf;
    if (v_len) *v_len = vl;
    return TPB_OK;
}

int32_t tpb_kick(tpb_semi_t semi, void *dv_ode, const void *v_ode, const void *u_ode, double t)
{
    (void)t;
    Semi *s = (Semi *)semi;
    if (!s) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null handle");
    if (!s->ready) return fail(s, TPB_ERR_STATE, "tpb_kick before tpb_semidiscretize");
    if (s->n_f > 0 && (!dv_ode || !v_ode || !u_ode)) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null ODE vector");
    CUDA_TRY(s, cudaSetDevice(s->cfg.device));
    return dispatch_kick(*s, dv_ode, v_ode, u_ode);
}

int32_t tpb_drift(tpb_semi_t semi, void *du_ode, const void *v_ode, const void *u_ode, double t)
{
    (void)t;
    Semi *s = (Semi *)semi;
    if (!s) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null handle");
    if (!s->ready) return fail(s, TPB_ERR_STATE, "tpb_drift before tpb_semidiscretize");
    if (s->n_f > 0 && (!du_ode || !v_ode)) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null ODE vector");
    CUDA_TRY(s, cudaSetDevice(s->cfg.device));
    return dispatch_drift(*s, du_ode, v_ode, u_ode);
}

int32_t tpb_get_system_field(tpb_semi_t semi, int32_t system, int32_t field, void *out, int64_t n)
{
    Semi *s = (Semi *)semi;
    if (!s || !out) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null argument");
    if (!s->ready) return fail(s, TPB_ERR_STATE, "tpb_get_system_field before tpb_semidiscretize");
    CUDA_TRY(s, cudaSetDevice(s->cfg.device));
    return dispatch_get_field(*s, system, field, out, n);
}

int32_t tpb_neighbor_pairs(tpb_semi_t semi, int32_t system, int32_t neighbor, const void *u_ode,
                           int64_t capacity, int32_t *out_i, int32_t *out_j, int64_t *count)
{
    Semi *s = (Semi *)semi;
    if (!s || !count || (capacity > 0 && (!out_i || !out_j))) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null argument");
    if (!s->ready) return fail(s, TPB_ERR_STATE, "tpb_neighbor_pairs before tpb_semidiscretize");
    CUDA_TRY(s, cudaSetDevice(s->cfg.device));
    return dispatch_pairs(*s, system, neighbor, u_ode, capacity, out_i, out_j, count);
}

int32_t tpb_synchronize(tpb_semi_t semi)
{
    Semi *s = (Semi *)semi;
    if (!s) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null handle");
    CUDA_TRY(s, cudaSetDevice(s->cfg.device));
    CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    if (s->h_flags && (*s->h_flags & 1))
        return fail(s, TPB_ERR_OUT_OF_BOUNDS,
                    "particle coordinates are NaN or outside the FullGridCellList bounding box");
    if (s->h_flags && (*s->h_flags & 2))
        return fail(s, TPB_ERR_STATE, "ghost exchange: a neighbour rank did not deliver its rows in time");
    return TPB_OK;
}

int32_t tpb_set_stream(tpb_semi_t semi, void *stream)
{
    Semi *s = (Semi *)semi;
    if (!s) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null handle");
    s->stream = stream ? (cudaStream_t)stream : s->own_stream;
    return TPB_OK;
}

int32_t tpb_get_sound_speed(tpb_semi_t semi, double *out)
{
    Semi *s = (Semi *)semi;
    if (!s || !out) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null argument");
    if (s->fluid_index < 0) return fail(s, TPB_ERR_STATE, "no fluid system");
    if (s->c_on_device && s->d_adapt) {
        // the kick left the value on the device (k_adaptive_consts): fetch it, stream-ordered
        const size_t off = s->cfg.eltype == TPB_F64 ? sizeof(tpb::AdaptConsts<double>) : sizeof(tpb::AdaptConsts<float>);
        double c = 0;
        CUDA_TRY(s, cudaMemcpyAsync(&c, (const unsigned char *)s->d_adapt + off, sizeof(double), cudaMemcpyDeviceToHost,
                                    s->stream));
        CUDA_TRY(s, cudaStreamSynchronize(s->stream));
        s->fp.sound_speed = c;
        if (s->wp.sound_speed_from_fluid) s->wp.sound_speed = c;
        s->c_on_device = false;
    }
    *out = s->fp.sound_speed;
    return TPB_OK;
}

int32_t tpb_max_speed2(tpb_semi_t semi, const void *v_ode, void *out_bits)
{
    Semi *s = (Semi *)semi;
    if (!s || !v_ode || !out_bits) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null argument");
    if (!s->ready || s->fluid_index < 0) return fail(s, TPB_ERR_STATE, "call tpb_semidiscretize first");
    if (s->cfg.ode_memory != TPB_MEM_DEVICE) return fail(s, TPB_ERR_UNSUPPORTED, "device ODE vectors only");
    CUDA_TRY(s, cudaSetDevice(s->cfg.device));
    return dispatch_max_speed2(*s, v_ode, out_bits);
}

int32_t tpb_set_integrate_structure(tpb_semi_t semi, int32_t enabled)
{
    Semi *s = (Semi *)semi;
    if (!s) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null handle");
    if (s->struct_index < 0) return fail(s, TPB_ERR_STATE, "no structure system");
    s->integrate_structure = enabled ? 1 : 0;
    return TPB_OK;
}

int32_t tpb_reinit_density(tpb_semi_t semi, void *v_ode, const void *u_ode)
{
    Semi *s = (Semi *)semi;
    if (!s) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null handle");
    if (!s->ready) return fail(s, TPB_ERR_STATE, "tpb_reinit_density before tpb_semidiscretize");
    if (s->n_f > 0 && (!v_ode || !u_ode)) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null ODE vector");
    CUDA_TRY(s, cudaSetDevice(s->cfg.device));
    return dispatch_reinit_density(*s, v_ode, u_ode);
}

int32_t tpb_set_structure_material(tpb_semi_t semi, const void *young_modulus, const void *poisson_ratio)
{
    Semi *s = (Semi *)semi;
    if (!s || !young_modulus || !poisson_ratio) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null argument");
    if (!s->ready || s->struct_index < 0) return fail(s, TPB_ERR_STATE, "needs a semidiscretized handle with a structure system");
    const size_t n = (size_t)s->n_s;
    CUDA_TRY(s, cudaSetDevice(s->cfg.device));
    // (lambda, mu, E) per particle in the system's eltype, with the reference's roundings (system.jl:158-161)
    auto fill = [&](auto tag) {
        using T = decltype(tag);
        const T *E = (const T *)young_modulus, *nu = (const T *)poisson_ratio;
        std::vector<T> m(3 * n);
        for (size_t a = 0; a < n; ++a) {
            if (!(E[a] > (T)0)) return false;
            m[3 * a] = E[a] * nu[a] / (((T)1 + nu[a]) * ((T)1 - (T)2 * nu[a]));
            m[3 * a + 1] = (E[a] / (T)2) / ((T)1 + nu[a]);
            m[3 * a + 2] = E[a];
        }
        if (!s->d_mat_s && cudaMalloc(&s->d_mat_s, sizeof(T) * 3 * (n + 8)) != cudaSuccess) return false;
        cudaStreamSynchronize(s->stream);
        return cudaMemcpy(s->d_mat_s, m.data(), sizeof(T) * 3 * n, cudaMemcpyHostToDevice) == cudaSuccess;
    };
    const bool ok = s->cfg.eltype == TPB_F64 ? fill(double()) : fill(float());
    if (!ok) return fail(s, TPB_ERR_INVALID_ARGUMENT, "young_modulus must be positive (or the device copy failed)");
    return TPB_OK;
}

int32_t tpb_sort_system(tpb_semi_t semi, int32_t system, void *v_ode, void *u_ode)
{
    Semi *s = (Semi *)semi;
    if (!s) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null handle");
    if (!s->ready) return fail(s, TPB_ERR_STATE, "tpb_sort_system before tpb_semidiscretize");
    if (system < 0 || system >= s->n_systems) return fail(s, TPB_ERR_INVALID_ARGUMENT, "system index out of range");
    if (system != s->fluid_index) return TPB_OK;  // sort_particles!(system, v, u, semi) = system for everything else
    if (s->n_f > 0 && (!v_ode || !u_ode)) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null ODE vector");
    CUDA_TRY(s, cudaSetDevice(s->cfg.device));
    return dispatch_sort_system(*s, v_ode, u_ode);
}

int32_t tpb_set_clamped_motion(tpb_semi_t semi, const void *coords, const void *velocity, const void *acceleration,
                               int32_t is_moving)
{
    Semi *s = (Semi *)semi;
    if (!s) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null handle");
    if (!s->ready || s->struct_index < 0)
        return fail(s, TPB_ERR_STATE, "needs a semidiscretized handle with a structure (or moving wall) system");
    if (is_moving && (!velocity || !acceleration))
        return fail(s, TPB_ERR_INVALID_ARGUMENT, "a moving system needs velocity and acceleration");
    const size_t ncl = (size_t)(s->n_s - s->n_s_int), nd = (size_t)s->cfg.ndims;
    const size_t ts = tsize(s->cfg.eltype), cs = tsize(s->cfg.coords_eltype);
    CUDA_TRY(s, cudaSetDevice(s->cfg.device));
    // synchronous copies from pageable host memory: the caller's arrays are free on return
    CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    if (ncl > 0 && coords)
        CUDA_TRY(s, cudaMemcpy((char *)s->d_xcl_s + cs * nd * (size_t)s->n_s_int, coords, cs * nd * ncl,
                               cudaMemcpyHostToDevice));
    if (ncl > 0 && is_moving) {
        CUDA_TRY(s, cudaMemcpy(s->d_vcl_s, velocity, ts * nd * ncl, cudaMemcpyHostToDevice));
        CUDA_TRY(s, cudaMemcpy(s->d_acl_s, acceleration, ts * nd * ncl, cudaMemcpyHostToDevice));
    }
    s->clamped_moving = is_moving ? 1 : 0;
    return TPB_OK;
}

static int split_ready(Semi *s)
{
    if (!s) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null handle");
    if (!s->ready || s->struct_index < 0) return fail(s, TPB_ERR_STATE, "needs a semidiscretized handle with a structure system");
    if (s->cfg.ode_memory != TPB_MEM_DEVICE) return fail(s, TPB_ERR_UNSUPPORTED, "split integration: device ODE vectors only");
    return TPB_OK;
}

int32_t tpb_structure_fluid_force(tpb_semi_t semi, void *dv_split, const void *v_ode, const void *u_ode)
{
    Semi *s = (Semi *)semi;
    if (int rc = split_ready(s)) return rc;
    if (!dv_split || !v_ode || !u_ode) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null argument");
    CUDA_TRY(s, cudaSetDevice(s->cfg.device));
    return dispatch_struct_force(*s, dv_split, v_ode, u_ode);
}

int32_t tpb_kick_structure(tpb_semi_t semi, void *dv_split, const void *v_split, const void *u_split,
                           const void *dv_const)
{
    Semi *s = (Semi *)semi;
    if (int rc = split_ready(s)) return rc;
    if (!dv_split || !u_split) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null argument");
    CUDA_TRY(s, cudaSetDevice(s->cfg.device));
    return dispatch_kick_struct(*s, dv_split, v_split, u_split, dv_const);
}

int32_t tpb_set_max_speed2(tpb_semi_t semi, const void *bits)
{
    Semi *s = (Semi *)semi;
    if (!s || !bits) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null argument");
    if (!s->fp.adaptive_sound_speed) return fail(s, TPB_ERR_STATE, "the fluid has no StateEquationAdaptiveCole");
    CUDA_TRY(s, cudaSetDevice(s->cfg.device));
    CUDA_TRY(s, cudaMemcpyAsync(s->d_vmax2, bits, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s->stream));
    s->vmax2_ready = true;
    return TPB_OK;
}

int32_t tpb_get_stats(tpb_semi_t semi, tpb_stats *out)
{
    Semi *s = (Semi *)semi;
    if (!s || !out) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null argument");
    *out = s->stats;
    return TPB_OK;
}

int32_t tpb_set_fluid_count(tpb_semi_t semi, int64_t n_active, int64_t n_targets)
{
    Semi *s = (Semi *)semi;
    if (!s) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null handle");
    if (n_active < 0 || n_active > s->n_f || n_targets < 0 || n_targets > n_active)
        return fail(s, TPB_ERR_INVALID_ARGUMENT, "need 0 <= n_targets <= n_active <= capacity of the fluid system");
    if (n_targets < n_active && s->fp.density_calculator == TPB_DENSITY_SUMMATION)
        return fail(s, TPB_ERR_UNSUPPORTED, "ghost particles need ContinuityDensity (their density travels with them)");
    if (n_targets < n_active && wall_integrates_density(*s))
        return fail(s, TPB_ERR_UNSUPPORTED, "slab ghosts are not combined with a ContinuityDensity wall");
    if (s->struct_index >= 0 && (n_active != s->n_f || n_targets != s->n_f))
        return fail(s, TPB_ERR_UNSUPPORTED, "slab ghosts are not combined with a structure system");
    s->n_act = n_active;
    s->n_tgt = n_targets;
    return TPB_OK;
}

int32_t tpb_set_fluid_mass(tpb_semi_t semi, int64_t first, int64_t count, const void *mass)
{
    Semi *s = (Semi *)semi;
    if (!s || (count > 0 && !mass)) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null argument");
    if (!s->ready) return fail(s, TPB_ERR_STATE, "tpb_set_fluid_mass before tpb_semidiscretize");
    if (first < 0 || count < 0 || first + count > s->n_f)
        return fail(s, TPB_ERR_INVALID_ARGUMENT, "mass range outside the fluid system");
    CUDA_TRY(s, cudaSetDevice(s->cfg.device));
    const size_t ts = tsize(s->cfg.eltype);
    CUDA_TRY(s, cudaMemcpyAsync((unsigned char *)s->d_mass_f + ts * (size_t)first, mass, ts * (size_t)count,
                                s->cfg.ode_memory == TPB_MEM_DEVICE ? cudaMemcpyDeviceToDevice
                                                                    : cudaMemcpyHostToDevice,
                                s->stream));
    if (s->cfg.ode_memory == TPB_MEM_HOST) CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    return TPB_OK;
}

// ---- device ODE-vector algebra (TPB_MEM_DEVICE handles; stream-ordered) -------------------
namespace {
int vec_check(Semi *s, int64_t n, int32_t eltype)
{
    if (!s) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null handle");
    if (s->cfg.ode_memory != TPB_MEM_DEVICE)
        return fail(s, TPB_ERR_STATE, "tpb_vec_* need a TPB_MEM_DEVICE handle (device pointers)");
    if (n < 0 || (eltype != TPB_F32 && eltype != TPB_F64)) return fail(s, TPB_ERR_INVALID_ARGUMENT, "bad vector length or eltype");
    cudaError_t e = cudaSetDevice(s->cfg.device);
    if (e != cudaSuccess) return fail(s, TPB_ERR_CUDA, cudaGetErrorString(e));
    return TPB_OK;
}
inline int vec_grid(int64_t n) { return (int)std::min<int64_t>(std::max<int64_t>((n + 255) / 256, 1), 148 * 16); }
}  // namespace

int32_t tpb_vec_axpby(tpb_semi_t semi, int64_t n, int32_t eltype, double a, const void *x, double b, void *y)
{
    Semi *s = (Semi *)semi;
    int rc = vec_check(s, n, eltype);
    if (rc || n == 0) return rc;
    if (eltype == TPB_F32) LAUNCH(*s, k_vec_axpby<float>, vec_grid(n), 256, 0, n, a, (const float *)x, b, (float *)y);
    else LAUNCH(*s, k_vec_axpby<double>, vec_grid(n), 256, 0, n, a, (const double *)x, b, (double *)y);
    CUDA_TRY(s, cudaGetLastError());
    return TPB_OK;
}

int32_t tpb_vec_rk2n_stage(tpb_semi_t semi, int64_t n, int32_t eltype, double A, double B, double dt,
                           const void *rhs, void *tmp, void *state)
{
    Semi *s = (Semi *)semi;
    int rc = vec_check(s, n, eltype);
    if (rc || n == 0) return rc;
    if (eltype == TPB_F32)
        LAUNCH(*s, k_vec_rk2n_stage<float>, vec_grid(n), 256, 0, n, A, B, dt, (const float *)rhs, (float *)tmp, (float *)state);
    else
        LAUNCH(*s, k_vec_rk2n_stage<double>, vec_grid(n), 256, 0, n, A, B, dt, (const double *)rhs, (double *)tmp, (double *)state);
    CUDA_TRY(s, cudaGetLastError());
    return TPB_OK;
}

int32_t tpb_vec_fill(tpb_semi_t semi, int64_t n, int32_t eltype, double value, void *x)
{
    Semi *s = (Semi *)semi;
    int rc = vec_check(s, n, eltype);
    if (rc || n == 0) return rc;
    if (eltype == TPB_F32) LAUNCH(*s, k_vec_fill<float>, vec_grid(n), 256, 0, n, value, (float *)x);
    else LAUNCH(*s, k_vec_fill<double>, vec_grid(n), 256, 0, n, value, (double *)x);
    CUDA_TRY(s, cudaGetLastError());
    return TPB_OK;
}

int32_t tpb_vec_lincomb4(tpb_semi_t semi, int64_t n, int32_t eltype, double a0, const void *x0, double a1,
                         const void *x1, double a2, const void *x2, double a3, const void *x3, void *y)
{
    Semi *s = (Semi *)semi;
    int rc = vec_check(s, n, eltype);
    if (rc || n == 0) return rc;
    if (!y) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null argument");
    if (eltype == TPB_F32)
        LAUNCH(*s, k_vec_lincomb4<float>, vec_grid(n), 256, 0, n, a0, (const float *)x0, a1, (const float *)x1, a2,
               (const float *)x2, a3, (const float *)x3, (float *)y);
    else
        LAUNCH(*s, k_vec_lincomb4<double>, vec_grid(n), 256, 0, n, a0, (const double *)x0, a1, (const double *)x1,
               a2, (const double *)x2, a3, (const double *)x3, (double *)y);
    CUDA_TRY(s, cudaGetLastError());
    return TPB_OK;
}

int32_t tpb_vec_verlet_update(tpb_semi_t semi, int64_t n_particles, int32_t ndims, int32_t nvars, int32_t eltype,
                              double dt, const void *kdu, const void *duprev, void *du)
{
    Semi *s = (Semi *)semi;
    int rc = vec_check(s, n_particles, eltype);
    if (rc || n_particles == 0) return rc;
    if (!kdu || !duprev || !du || ndims < 1 || nvars < ndims || nvars > ndims + 1)
        return fail(s, TPB_ERR_INVALID_ARGUMENT, "invalid argument");
    if (eltype == TPB_F32)
        LAUNCH(*s, k_vec_verlet_update<float>, vec_grid(n_particles), 256, 0, n_particles, ndims, nvars, (float)dt,
               (const float *)kdu, (const float *)duprev, (float *)du);
    else
        LAUNCH(*s, k_vec_verlet_update<double>, vec_grid(n_particles), 256, 0, n_particles, ndims, nvars, dt,
               (const double *)kdu, (const double *)duprev, (double *)du);
    CUDA_TRY(s, cudaGetLastError());
    return TPB_OK;
}

int32_t tpb_vec_div_fast(tpb_semi_t semi, int64_t n, int32_t eltype, double x, const void *y, void *out)
{
    Semi *s = (Semi *)semi;
    int rc = vec_check(s, n, eltype);
    if (rc || n == 0) return rc;
    if (!y || !out) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null argument");
    if (eltype == TPB_F32)
        LAUNCH(*s, k_div_fast<float>, cdiv(n, 256), 256, 0, n, (float)x, (const float *)y, (float *)out);
    else
        LAUNCH(*s, k_div_fast<double>, cdiv(n, 256), 256, 0, n, x, (const double *)y, (double *)out);
    CUDA_TRY(s, cudaGetLastError());
    return TPB_OK;
}

int32_t tpb_vec_strided_max(tpb_semi_t semi, int64_t count, int32_t eltype, int32_t stride, int32_t offset,
                            const void *x, double *out_host)
{
    Semi *s = (Semi *)semi;
    int rc = vec_check(s, count, eltype);
    if (rc) return rc;
    if (!out_host || stride <= 0 || offset < 0) return fail(s, TPB_ERR_INVALID_ARGUMENT, "bad argument");
    double *d_out = (double *)s->d_scratch;  // >= 8 bytes
    const double ninf = -std::numeric_limits<double>::infinity();
    CUDA_TRY(s, cudaMemcpyAsync(d_out, &ninf, sizeof(double), cudaMemcpyHostToDevice, s->stream));
    if (count > 0) {
        if (eltype == TPB_F32) LAUNCH(*s, k_vec_strided_max<float>, vec_grid(count), 256, 0, count, stride, offset, (const float *)x, d_out);
        else LAUNCH(*s, k_vec_strided_max<double>, vec_grid(count), 256, 0, count, stride, offset, (const double *)x, d_out);
    }
    CUDA_TRY(s, cudaMemcpyAsync(out_host, d_out, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    return TPB_OK;
}

int32_t tpb_vec_wrms_norm(tpb_semi_t semi, int64_t n, int32_t eltype, const void *err, const void *u_prev,
                          const void *u, double abstol, double reltol, double *out_host)
{
    Semi *s = (Semi *)semi;
    int rc = vec_check(s, n, eltype);
    if (rc) return rc;
    if (!out_host) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null argument");
    *out_host = 0.0;
    if (n == 0) return TPB_OK;
    const int blocks = (int)std::min<int64_t>(vec_grid(n), 1024);
    double *d_part = (double *)s->d_scratch;
    if ((int64_t)blocks > std::max<int64_t>(std::max(s->n_f, s->n_w), 1))  // scratch holds max(n_f, n_w) doubles
        return fail(s, TPB_ERR_INVALID_ARGUMENT, "vector too short for the reduction scratch space");
    if (eltype == TPB_F32)
        LAUNCH(*s, k_vec_wrms_partial<float>, blocks, 256, 0, n, (const float *)err, (const float *)u_prev,
               (const float *)u, abstol, reltol, d_part);
    else
        LAUNCH(*s, k_vec_wrms_partial<double>, blocks, 256, 0, n, (const double *)err, (const double *)u_prev,
               (const double *)u, abstol, reltol, d_part);
    std::vector<double> part((size_t)blocks);
    CUDA_TRY(s, cudaMemcpyAsync(part.data(), d_part, sizeof(double) * (size_t)blocks, cudaMemcpyDeviceToHost,
                                s->stream));
    CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    double sum = 0.0;
    for (double p : part) sum += p;
    *out_host = std::sqrt(sum / (double)n);
    return TPB_OK;
}

int32_t tpb_set_profiling(tpb_semi_t semi, int32_t max_kicks)
{
    Semi *s = (Semi *)semi;
    if (!s || max_kicks < 0) return fail(s, TPB_ERR_INVALID_ARGUMENT, "invalid argument");
    CUDA_TRY(s, cudaSetDevice(s->cfg.device));
    for (cudaEvent_t e : s->prof_events) cudaEventDestroy(e);
    s->prof_events.assign((size_t)max_kicks * TPB_N_PHASES, nullptr);
    for (auto &e : s->prof_events) CUDA_TRY(s, cudaEventCreate(&e));
    s->prof_capacity = max_kicks;
    s->prof_kicks = 0;
    return TPB_OK;
}

int32_t tpb_get_phase_times(tpb_semi_t semi, double *ms_mean, int32_t *n_kicks)
{
    Semi *s = (Semi *)semi;
    if (!s || !ms_mean) return fail(s, TPB_ERR_INVALID_ARGUMENT, "null argument");
    CUDA_TRY(s, cudaSetDevice(s->cfg.device));
    CUDA_TRY(s, cudaStreamSynchronize(s->stream));
    for (int ph = 0; ph < TPB_N_PHASES - 1; ++ph) ms_mean[ph] = 0.0;
    for (int k = 0; k < s->prof_kicks; ++k)
        for (int ph = 0; ph < TPB_N_PHASES - 1; ++ph) {
            float ms = 0.f;
            CUDA_TRY(s, cudaEventElapsedTime(&ms, s->prof_events[(size_t)k * TPB_N_PHASES + ph],
                                             s->prof_events[(size_t)k * TPB_N_PHASES + ph + 1]));
            ms_mean[ph] += ms;
        }
    if (s->prof_kicks > 0)
        for (int ph = 0; ph < TPB_N_PHASES - 1; ++ph) ms_mean[ph] /= s->prof_kicks;
    if (n_kicks) *n_kicks = s->prof_kicks;
    s->prof_kicks = 0;
    return TPB_OK;
}

// ---- ghost exchange over peer memory (tpb_halo.cuh)
int32_t tpb_peer_alloc(int64_t bytes, void **out)
{
    if (!out || bytes <= 0) return TPB_ERR_INVALID_ARGUMENT;
    void *p = nullptr;
    if (cudaMalloc(&p, (size_t)bytes) != cudaSuccess) return TPB_ERR_CUDA;
    if (cudaMemset(p, 0, (size_t)bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
        cudaFree(p);
        return TPB_ERR_CUDA;
    }
    *out = p;
    return TPB_OK;
}

int32_t tpb_peer_free(void *ptr)
{
    return cudaFree(ptr) == cudaSuccess ? TPB_OK : TPB_ERR_CUDA;
}

int32_t tpb_peer_export(void *ptr, void *handle64)
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, ptr) != cudaSuccess) return TPB_ERR_CUDA;
    std::memcpy(handle64, &h, sizeof(h));
    return TPB_OK;
}

int32_t tpb_peer_import(const void *handle64, void **out)
{
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, sizeof(h));
    void *p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError();
        return TPB_ERR_CUDA;
    }
    *out = p;
    return TPB_OK;
}

int32_t tpb_peer_close(void *ptr)
{
    return cudaIpcCloseMemHandle(ptr) == cudaSuccess ? TPB_OK : TPB_ERR_CUDA;
}

int32_t tpb_halo_pack(tpb_semi_t semi, int32_t side, double threshold, const void *u, const void *v,
                      const int64_t *candidates, int64_t n_candidates, void *peer_u, void *peer_v,
                      void *done_counter, int64_t done_target_blocks, void *peer_flag, uint32_t epoch,
                      int32_t *blocks_out)
{
    Semi *s = (Semi *)semi;
    if (!s || !s->ready) return fail(s, TPB_ERR_STATE, "tpb_halo_pack: handle not semidiscretized");
    const int nd = s->cfg.ndims;
    const int nv = s->fp.density_calculator == TPB_DENSITY_SUMMATION ? nd : nd + 1;
    const int blocks = (int)std::min<int64_t>(std::max<int64_t>((n_candidates + 255) / 256, 1), 148 * 4);
    if (blocks_out) *blocks_out = blocks;
    const unsigned long long target = (unsigned long long)(done_target_blocks + blocks);
    const bool f32 = s->cfg.eltype == TPB_F32, c32 = s->cfg.coords_eltype == TPB_F32;
#define TPB_PACK(T, CT)                                                                                  \
    LAUNCH(*s, (k_halo_pack<T, CT>), blocks, 256, 0, nd, nv, (const CT *)u, (const T *)v, candidates,    \
           n_candidates, (CT)threshold, (int)side, (CT *)peer_u, (T *)peer_v,                            \
           (unsigned long long *)done_counter, target, (uint32_t *)peer_flag, epoch)
    if (f32 && c32) TPB_PACK(float, float);
    else if (f32) TPB_PACK(float, double);
    else if (c32) return fail(s, TPB_ERR_INVALID_ARGUMENT, "Float64 systems need Float64 coordinates");
    else TPB_PACK(double, double);
#undef TPB_PACK
    CUDA_TRY(s, cudaGetLastError());
    return TPB_OK;
}

int32_t tpb_halo_install(tpb_semi_t semi, int64_t n_rows, const void *stage_u, const void *stage_v,
                         void *u_ghost, void *v_ghost, const void *flag_left, const void *flag_right,
                         uint32_t epoch, double timeout_s, void *timed_out_flag)
{
    Semi *s = (Semi *)semi;
    if (!s || !s->ready) return fail(s, TPB_ERR_STATE, "tpb_halo_install: handle not semidiscretized");
    const int nd = s->cfg.ndims;
    const int nv = s->fp.density_calculator == TPB_DENSITY_SUMMATION ? nd : nd + 1;
    const int64_t n_u = n_rows * nd, n_v = n_rows * nv;
    const int blocks = (int)std::min<int64_t>(std::max<int64_t>((n_v + 255) / 256, 1), 148 * 2);
    const unsigned long long timeout_ns = (unsigned long long)(timeout_s * 1e9);
    const bool f32 = s->cfg.eltype == TPB_F32, c32 = s->cfg.coords_eltype == TPB_F32;
#define TPB_INSTALL(T, CT)                                                                               \
    LAUNCH(*s, (k_halo_install<T, CT>), blocks, 256, 0, n_u, n_v, (const CT *)stage_u, (const T *)stage_v, \
           (CT *)u_ghost, (T *)v_ghost, (const uint32_t *)flag_left, (const uint32_t *)flag_right, epoch, \
           timeout_ns, timed_out_flag ? (int *)timed_out_flag : s->d_flags)
    if (f32 && c32) TPB_INSTALL(float, float);
    else if (f32) TPB_INSTALL(float, double);
    else if (c32) return fail(s, TPB_ERR_INVALID_ARGUMENT, "Float64 systems need Float64 coordinates");
    else TPB_INSTALL(double, double);
#undef TPB_INSTALL
    CUDA_TRY(s, cudaGetLastError());
    return TPB_OK;
}

int32_t tpb_host_register(void *ptr, int64_t bytes)
{
    return cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault) == cudaSuccess ? TPB_OK : TPB_ERR_CUDA;
}

int32_t tpb_host_unregister(void *ptr)
{
    return cudaHostUnregister(ptr) == cudaSuccess ? TPB_OK : TPB_ERR_CUDA;
}

}  // extern "C"
#endif  // !TPB_TU_TAG (main unit)
