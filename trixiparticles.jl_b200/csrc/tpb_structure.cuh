// tpb_structure.cuh -- TotalLagrangianSPHSystem on the device and its coupling to the fluid
// (BASELINE config 5, examples/fsi/dam_break_plate_2d.jl).  B200 counterpart of
//   update_tlsph_positions!, calc_deformation_grad!, compute_pk1_corrected!
//        /root/reference/src/schemes/structure/total_lagrangian_sph/system.jl:403-424, :468-586
//   interact_structure_structure! + PenaltyForceGanzenmueller
//        /root/reference/src/schemes/structure/total_lagrangian_sph/rhs.jl:15-98, penalty_force.jl:24-66
//   interact_structure_fluid!      /root/reference/src/schemes/structure/structure.jl:20-102
//   fluid <- structure `interact!` /root/reference/src/schemes/fluid/weakly_compressible_sph/rhs.jl:5-127
//   BoundaryModelMonaghanKajtar    /root/reference/src/schemes/boundary/wall_boundary/monaghan_kajtar.jl:48-111
//
// The structure interacts with itself over the INITIAL configuration, so its neighbour list is
// built once (`tpb_semidiscretize`, the reference's frozen PrecomputedNeighborhoodSearch,
// system.jl:186-247) and stored in CSR form; every kick walks it twice (deformation gradient, then
// forces -- the second sweep needs the first one's result of every neighbour).  The structure of an
// FSI run is small next to the fluid: these are plain one-thread-per-particle kernels.  Towards the
// fluid the structure particles are binned into the shared cell grid every kick (they move) and
// swept by the fluid particles after `k_interact_tiles`; the structure reads the fluid's sorted
// records through the fluid's own cell list.
// Matrices use the reference's memory layout of an ND x ND x n array: (i, j) of particle p at
// [i + ND * j + ND * ND * p].
#pragma once
#include "tpb_device.cuh"
#include "tpb_sweeps.cuh"

namespace tpb {

#define TPB_MAT(M, i, j) (M)[(i) + ND * (j)]

template <typename T>
struct StructConst {
    KernelConst<T> kern;  // structure's smoothing kernel / length
    T almostzero;         // sqrt(eps(h^2))
    T lambda, mu, young;  // Lame constants, Young's modulus (scalars)
    T half_alpha;         // PenaltyForceGanzenmueller: alpha / 2
    int has_penalty;
    T acc[3];             // system.acceleration
};

// BoundaryModelMonaghanKajtar(K, beta, boundary_particle_spacing, hydrodynamic_mass) seen from the fluid
template <typename T>
struct MKConst {
    T K_bpow;     // K / beta^(ND - 1)
    T spacing;    // boundary_particle_spacing
    T min_dfs;    // spacing / 100
    T h_fluid;    // smoothing length of the fluid particle (boundary_kernel argument)
    T vol;        // spacing^ND: current_density = hydrodynamic_mass / spacing^ND
    T radius2;    // compact_support(fluid, structure)^2 = the fluid's
    T almostzero_fs;  // fluid <- structure: sqrt(eps(compact_support^2)) (wcsph/rhs.jl:27-28)
    T almostzero_sf;  // structure <- fluid: sqrt(eps(h_fluid^2))          (structure.jl:36-37)
};

// boundary_kernel (monaghan_kajtar.jl:90-100)
template <typename T>
__device__ __forceinline__ T mk_boundary_kernel(T r, T h)
{
    const T q = r / h;
    if (q >= (T)2) return (T)0;
    const T x = (T)2 - q;
    const T x2 = x * x;
    return (T)(177.0 / 3200.0) * ((T)1 + (T)2.5 * q + (T)2 * (q * q)) * ((x2 * x2) * x);
}

// pressure_acceleration against a Monaghan-Kajtar neighbour (monaghan_kajtar.jl:48-73)
template <int ND, typename T>
__device__ __forceinline__ void mk_pressure_acceleration(const MKConst<T> &k, const T (&pd)[3], T dist, T (&out)[3])
{
    T dfs = dist - k.spacing;
    dfs = k.min_dfs > dfs ? k.min_dfs : dfs;
    const T bk = mk_boundary_kernel(dist, k.h_fluid);
    const T den = dist * dfs;
#pragma unroll
    for (int d = 0; d < ND; ++d) out[d] = (k.K_bpow * pd[d]) / den * bk;
}

// current_coordinates <- u (integrated particles); clamped particles keep their initial position
template <int ND, typename CT>
__global__ void __launch_bounds__(256)
k_struct_positions(int n, int n_int, const CT *__restrict__ u_s, const CT *__restrict__ x_clamped,
                   CT *__restrict__ x_cur)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n * ND) return;
    x_cur[q] = q < n_int * ND ? u_s[q] : x_clamped[q];
}

// sorted records of the structure as a neighbour of the fluid: A = (x, hydrodynamic mass),
// B = (velocity (clamped particles: the prescribed one while they move, else 0; system.jl:302-321), m / spacing^ND)
template <int ND, typename T, typename CT>
__global__ void __launch_bounds__(256)
k_reorder_struct(const CT *__restrict__ x_cur, const T *__restrict__ v_s, const T *__restrict__ hydro_mass,
                 const int *__restrict__ key, const int *__restrict__ cell_start,
                 const int *__restrict__ tmp_perm, int n, int n_int, T mk_vol, V4<CT> *__restrict__ A,
                 V4<T> *__restrict__ B, int *__restrict__ sperm = nullptr, const T *__restrict__ v_clamped = nullptr,
                 int pbits = 31)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int e = tmp_perm[s];
    const int i = perm_index(e, pbits);
    const int c = key[i];
    const int a = cell_start[c], b = cell_start[c + 1];
    const int dst = a + rank_in_cell(tmp_perm, a, b, e);
    V4<CT> ra;
    ra.x = x_cur[(int64_t)i * ND + 0];
    ra.y = x_cur[(int64_t)i * ND + 1];
    ra.z = ND == 3 ? x_cur[(int64_t)i * ND + 2] : (CT)0;
    ra.w = (CT)hydro_mass[i];
    V4<T> rb;
    const bool moving = i < n_int || v_clamped != nullptr;
    const T *vp = i < n_int ? v_s + (int64_t)i * ND : v_clamped + (int64_t)(i - n_int) * ND;
    rb.x = moving ? vp[0] : (T)0;
    rb.y = moving ? vp[1] : (T)0;
    rb.z = ND == 3 && moving ? vp[ND - 1] : (T)0;
    rb.w = hydro_mass[i] / mk_vol;
    A[dst] = ra;
    B[dst] = rb;
    if (sperm) sperm[dst] = i;
}

template <int ND, typename T, typename CT>
__device__ __forceinline__ T struct_pos_diff(const CT *__restrict__ x, int a, int b, T (&pd)[3])
{
    V4<CT> xa, xb;
    xa.x = x[(int64_t)a * ND], xa.y = x[(int64_t)a * ND + 1], xa.z = ND == 3 ? x[(int64_t)a * ND + 2] : (CT)0;
    xb.x = x[(int64_t)b * ND], xb.y = x[(int64_t)b * ND + 1], xb.z = ND == 3 ? x[(int64_t)b * ND + 2] : (CT)0;
    xa.w = xb.w = (CT)0;
    return pos_diff_d2<ND, T, CT>(xa, xb, pd);
}

// initialize! -> compute_gradient_correction_matrix! (general/corrections.jl:329-355, :425-456), once:
// L_a = (-sum_b V_b grad W_ab (x) x_ab)^-1 over the initial configuration; identity when |det| < 1f-9
template <int ND, typename T, typename CT, int KERNEL>
__global__ void __launch_bounds__(128)
k_struct_correction_matrix(int n, StructConst<T> k, T eps_h2, const int *__restrict__ nbr_start,
                           const int *__restrict__ nbr, const CT *__restrict__ x0, const T *__restrict__ mass,
                           const T *__restrict__ rho, T *__restrict__ L)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    T C[ND * ND];
#pragma unroll
    for (int q = 0; q < ND * ND; ++q) C[q] = (T)0;
    for (int e = nbr_start[a]; e < nbr_start[a + 1]; ++e) {
        const int b = nbr[e];
        T pd[3];
        const T dist = sqrt_rn(struct_pos_diff<ND, T, CT>(x0, a, b, pd));
        if (dist >= k.kern.support || dist * dist < eps_h2) continue;  // the safe gradient is zero
        const T wdr = SmoothingKernel<KERNEL, T>::dw_div_r(k.kern, dist);
        const T volume = mass[b] / rho[b];
#pragma unroll
        for (int j = 0; j < ND; ++j)
#pragma unroll
            for (int i = 0; i < ND; ++i) TPB_MAT(C, i, j) -= (volume * (wdr * pd[i])) * pd[j];
    }
    T inv[ND * ND], det;
    if (ND == 2) {
        const T a_ = TPB_MAT(C, 0, 0), b_ = TPB_MAT(C, 0, 1), c_ = TPB_MAT(C, 1, 0), d_ = TPB_MAT(C, 1, 1);
        det = a_ * d_ - b_ * c_;
        const T idet = (T)1 / det;
        TPB_MAT(inv, 0, 0) = d_ * idet;
        TPB_MAT(inv, 0, 1) = -b_ * idet;
        TPB_MAT(inv, 1, 0) = -c_ * idet;
        TPB_MAT(inv, 1, 1) = a_ * idet;
    } else {
        const T a_ = TPB_MAT(C, 0, 0), b_ = TPB_MAT(C, 0, 1), c_ = TPB_MAT(C, 0, ND - 1);
        const T d_ = TPB_MAT(C, 1, 0), e_ = TPB_MAT(C, 1, 1), f_ = TPB_MAT(C, 1, ND - 1);
        const T g_ = TPB_MAT(C, ND - 1, 0), h_ = TPB_MAT(C, ND - 1, 1), i_ = TPB_MAT(C, ND - 1, ND - 1);
        const T c00 = e_ * i_ - f_ * h_, c01 = f_ * g_ - d_ * i_, c02 = d_ * h_ - e_ * g_;
        det = a_ * c00 + b_ * c01 + c_ * c02;
        const T idet = (T)1 / det;
        TPB_MAT(inv, 0, 0) = c00 * idet;
        TPB_MAT(inv, 0, 1) = (c_ * h_ - b_ * i_) * idet;
        TPB_MAT(inv, 0, ND - 1) = (b_ * f_ - c_ * e_) * idet;
        TPB_MAT(inv, 1, 0) = c01 * idet;
        TPB_MAT(inv, 1, 1) = (a_ * i_ - c_ * g_) * idet;
        TPB_MAT(inv, 1, ND - 1) = (c_ * d_ - a_ * f_) * idet;
        TPB_MAT(inv, ND - 1, 0) = c02 * idet;
        TPB_MAT(inv, ND - 1, 1) = (b_ * g_ - a_ * h_) * idet;
        TPB_MAT(inv, ND - 1, ND - 1) = (a_ * e_ - b_ * d_) * idet;
    }
    const bool singular = fabs((double)det) < (double)1.0e-9f;
#pragma unroll
    for (int j = 0; j < ND; ++j)
#pragma unroll
        for (int i = 0; i < ND; ++i) L[(int64_t)a * ND * ND + i + ND * j] = singular ? (T)(i == j) : TPB_MAT(inv, i, j);
}

// calc_deformation_grad! + compute_pk1_corrected! for every particle
template <int ND, typename T, typename CT, int KERNEL>
__global__ void __launch_bounds__(128)
k_struct_defgrad_pk1(int n, StructConst<T> k, const int *__restrict__ nbr_start, const int *__restrict__ nbr,
                     const CT *__restrict__ x0, const CT *__restrict__ x_cur, const T *__restrict__ mass,
                     const T *__restrict__ rho, const T *__restrict__ L, T *__restrict__ F_out,
                     T *__restrict__ pk1_rho2, const T *__restrict__ material = nullptr)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    // per-particle material constants (young_modulus / poisson_ratio given as vectors, system.jl:108-161):
    // (lambda, mu, E) per particle
    if (material != nullptr) k.lambda = material[3 * (int64_t)a], k.mu = material[3 * (int64_t)a + 1];
    T La[ND * ND], F[ND * ND];
#pragma unroll
    for (int q = 0; q < ND * ND; ++q) {
        La[q] = L[(int64_t)a * ND * ND + q];
        F[q] = (T)0;
    }
    for (int e = nbr_start[a]; e < nbr_start[a + 1]; ++e) {
        const int b = nbr[e];
        T pd0[3], pd[3];
        const T dist0 = sqrt_rn(struct_pos_diff<ND, T, CT>(x0, a, b, pd0));
        if (dist0 < k.almostzero) continue;
        const T wdr = SmoothingKernel<KERNEL, T>::dw_div_r(k.kern, dist0);
        const T volume = mass[b] / rho[b];
        struct_pos_diff<ND, T, CT>(x_cur, a, b, pd);
        T w[ND];
#pragma unroll
        for (int i = 0; i < ND; ++i) {
            T s = (-volume * TPB_MAT(La, i, 0)) * (wdr * pd0[0]);
#pragma unroll
            for (int q = 1; q < ND; ++q) s = s + (-volume * TPB_MAT(La, i, q)) * (wdr * pd0[q]);
            w[i] = s;
        }
#pragma unroll
        for (int j = 0; j < ND; ++j)
#pragma unroll
            for (int i = 0; i < ND; ++i) TPB_MAT(F, i, j) += w[j] * pd[i];
    }
    // E = (F'F - I) / 2;  S = lambda tr(E) I + 2 mu E;  PK1 = F S;  PK1 L / rho^2
    T E[ND * ND], S[ND * ND], P[ND * ND];
    T trE = (T)0;
#pragma unroll
    for (int i = 0; i < ND; ++i)
#pragma unroll
        for (int j = 0; j < ND; ++j) {
            T s = TPB_MAT(F, 0, i) * TPB_MAT(F, 0, j);
#pragma unroll
            for (int q = 1; q < ND; ++q) s = s + TPB_MAT(F, q, i) * TPB_MAT(F, q, j);
            TPB_MAT(E, i, j) = (s - (T)(i == j)) / (T)2;
        }
#pragma unroll
    for (int i = 0; i < ND; ++i) trE = i == 0 ? TPB_MAT(E, 0, 0) : trE + TPB_MAT(E, i, i);
#pragma unroll
    for (int i = 0; i < ND; ++i)
#pragma unroll
        for (int j = 0; j < ND; ++j)
            TPB_MAT(S, i, j) = (i == j ? k.lambda * trE : (T)0) + (T)2 * k.mu * TPB_MAT(E, i, j);
#pragma unroll
    for (int i = 0; i < ND; ++i)
#pragma unroll
        for (int j = 0; j < ND; ++j) {
            T s = TPB_MAT(F, i, 0) * TPB_MAT(S, 0, j);
#pragma unroll
            for (int q = 1; q < ND; ++q) s = s + TPB_MAT(F, i, q) * TPB_MAT(S, q, j);
            TPB_MAT(P, i, j) = s;
        }
    const T rho_a = rho[a];
    const T rho2_inv = (T)1 / (rho_a * rho_a);
#pragma unroll
    for (int i = 0; i < ND; ++i)
#pragma unroll
        for (int j = 0; j < ND; ++j) {
            T s = TPB_MAT(P, i, 0) * TPB_MAT(La, 0, j);
#pragma unroll
            for (int q = 1; q < ND; ++q) s = s + TPB_MAT(P, i, q) * TPB_MAT(La, q, j);
            pk1_rho2[(int64_t)a * ND * ND + i + ND * j] = s * rho2_inv;
        }
#pragma unroll
    for (int q = 0; q < ND * ND; ++q) F_out[(int64_t)a * ND * ND + q] = F[q];
}

// interact_structure_structure! for the integrated particles; dv_s += S_ss, then the source term
template <int ND, typename T, typename CT, int KERNEL>
__global__ void __launch_bounds__(128)
k_struct_interact(int n_int, StructConst<T> k, const int *__restrict__ nbr_start, const int *__restrict__ nbr,
                  const CT *__restrict__ x0, const CT *__restrict__ x_cur, const T *__restrict__ mass,
                  const T *__restrict__ rho, const T *__restrict__ F, const T *__restrict__ pk1_rho2,
                  T *__restrict__ dv_s, const T *__restrict__ material = nullptr)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n_int) return;
    const T m_a = mass[a], rho_a = rho[a];
    const T young_a = material != nullptr ? material[3 * (int64_t)a + 2] : k.young;
    T Pa[ND * ND], Fa[ND * ND];
#pragma unroll
    for (int q = 0; q < ND * ND; ++q) {
        Pa[q] = pk1_rho2[(int64_t)a * ND * ND + q];
        Fa[q] = F[(int64_t)a * ND * ND + q];
    }
    T acc[3] = {0, 0, 0};
    for (int e = nbr_start[a]; e < nbr_start[a + 1]; ++e) {
        const int b = nbr[e];
        T pd0[3], cpd[3];
        const T dist0 = sqrt_rn(struct_pos_diff<ND, T, CT>(x0, a, b, pd0));
        if (dist0 < k.almostzero) continue;
        const T wdr = SmoothingKernel<KERNEL, T>::dw_div_r(k.kern, dist0);
        const T m_b = mass[b], rho_b = rho[b];
        const T cd2 = struct_pos_diff<ND, T, CT>(x_cur, a, b, cpd);
        const T *Pb = pk1_rho2 + (int64_t)b * ND * ND;
#pragma unroll
        for (int i = 0; i < ND; ++i) {
            T s = (m_b * (TPB_MAT(Pa, i, 0) + TPB_MAT(Pb, i, 0))) * (wdr * pd0[0]);
#pragma unroll
            for (int q = 1; q < ND; ++q) s = s + (m_b * (TPB_MAT(Pa, i, q) + TPB_MAT(Pb, i, q))) * (wdr * pd0[q]);
            acc[i] += s;
        }
        if (k.has_penalty) {
            const T *Fb = F + (int64_t)b * ND * ND;
            const T volume_a = div_fast(m_a, rho_a), volume_b = div_fast(m_b, rho_b);
            const T kw = SmoothingKernel<KERNEL, T>::w_unsafe(k.kern, dist0);
            T da = (T)0, db = (T)0;
#pragma unroll
            for (int i = 0; i < ND; ++i) {
                T sa = TPB_MAT(Fa, i, 0) * pd0[0], sb = TPB_MAT(Fb, i, 0) * pd0[0];
#pragma unroll
                for (int q = 1; q < ND; ++q) {
                    sa = sa + TPB_MAT(Fa, i, q) * pd0[q];
                    sb = sb + TPB_MAT(Fb, i, q) * pd0[q];
                }
                da = i == 0 ? (sa - cpd[0]) * cpd[0] : da + (sa - cpd[i]) * cpd[i];
                db = i == 0 ? (sb - cpd[0]) * cpd[0] : db + (sb - cpd[i]) * cpd[i];
            }
            // E_a dot(eps_a, r) + E_b dot(eps_b, r) (penalty_force.jl:42-53)
            const T young_b = material != nullptr ? material[3 * (int64_t)b + 2] : k.young;
            const T delta_sum = young_a * da + young_b * db;
            // current_distance^2 = (sqrt(d2))^2 in the reference
            const T cdist = sqrt_rn(cd2);
            const T f = div_fast(k.half_alpha * volume_a * volume_b * kw * delta_sum,
                                 (dist0 * dist0) * (cdist * cdist) * m_a);
#pragma unroll
            for (int i = 0; i < ND; ++i) acc[i] += f * cpd[i];
        }
    }
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        const int64_t o = (int64_t)a * ND + d;
        dv_s[o] = (dv_s[o] + acc[d]) + k.acc[d];
    }
}

// structure <- fluid (interact_structure_fluid!, Monaghan-Kajtar): one thread per integrated
// structure particle walks the fluid's sorted records; dv_s = S_sf (overwrites)
template <int ND, typename T, typename CT>
__global__ void __launch_bounds__(128)
k_struct_from_fluid(int n_int, GridConst<CT> g, const CT *__restrict__ x_cur, const T *__restrict__ mass_s,
                    const int *__restrict__ fcell_start, const V4<CT> *__restrict__ A, int enabled, MKConst<T> k,
                    T *__restrict__ dv_s, int *__restrict__ flags)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n_int) return;
    V4<CT> xi;
    xi.x = x_cur[(int64_t)a * ND], xi.y = x_cur[(int64_t)a * ND + 1];
    xi.z = ND == 3 ? x_cur[(int64_t)a * ND + 2] : (CT)0;
    xi.w = (CT)0;
    T acc[3] = {0, 0, 0};
    int cx, cy, cz;
    if (!cell_coords<ND, CT>(g, xi.x, xi.y, xi.z, cx, cy, cz)) atomicOr(flags, 1);
    else if (enabled) {
        const T m_a = mass_s[a];
        for_neighbor_rows<ND, CT>(g, fcell_start, cx, cy, cz, [&](int j0, int j1) {
            for (int j = j0; j < j1; ++j) {
                const V4<CT> xj = A[j];
                T pd[3];
                const T d2 = pos_diff_d2<ND, T, CT>(xi, xj, pd);
                if (d2 <= k.radius2) {
                    const T dist = sqrt_rn(d2);
                    if (dist >= k.almostzero_sf) {
                        T dvp[3];
                        mk_pressure_acceleration<ND, T>(k, pd, dist, dvp);
#pragma unroll
                        for (int d = 0; d < ND; ++d) acc[d] += dvp[d] * (T)xj.w / m_a;
                    }
                }
            }
        });
    }
#pragma unroll
    for (int d = 0; d < ND; ++d) dv_s[(int64_t)a * ND + d] = acc[d];
}

// fluid <- structure (wcsph/rhs.jl with the Monaghan-Kajtar pressure acceleration, no viscous term,
// continuity with the structure's velocity): one thread per sorted fluid particle; dv_f += S_fs
template <int ND, typename T, typename CT, int KERNEL, int NV>
__global__ void __launch_bounds__(128)
k_fluid_from_struct(int n_f, GridConst<CT> g, const int *__restrict__ fcell_start, const V4<CT> *__restrict__ A,
                    const V4<T> *__restrict__ B, const int *__restrict__ perm, const int *__restrict__ scell_start,
                    const V4<CT> *__restrict__ As, const V4<T> *__restrict__ Bs, KernelConst<T> kern, MKConst<T> k,
                    T *__restrict__ dv, int n_targets)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_f || s >= fcell_start[g.ncells]) return;
    const int orig = perm[s];
    if (orig >= n_targets) return;
    const V4<CT> xi = A[s];
    int cx, cy, cz;
    cell_coords<ND, CT>(g, xi.x, xi.y, xi.z, cx, cy, cz);
    // most fluid particles are nowhere near the structure: look at the row ranges first
    bool any = false;
    for_neighbor_rows<ND, CT>(g, scell_start, cx, cy, cz, [&](int j0, int j1) { any = any || j1 > j0; });
    if (!any) return;
    const V4<T> bi = B[s];
    T acc[3] = {0, 0, 0}, drho = (T)0;
    for_neighbor_rows<ND, CT>(g, scell_start, cx, cy, cz, [&](int j0, int j1) {
        for (int j = j0; j < j1; ++j) {
            const V4<CT> xj = As[j];
            T pd[3];
            const T d2 = pos_diff_d2<ND, T, CT>(xi, xj, pd);
            if (d2 <= k.radius2) {
                const T dist = sqrt_rn(d2);
                if (dist >= k.almostzero_fs) {
                    const V4<T> bj = Bs[j];
                    const T wdr = SmoothingKernel<KERNEL, T>::dw_div_r(kern, dist);
                    T dvp[3];
                    mk_pressure_acceleration<ND, T>(k, pd, dist, dvp);
                    T vg = (bi.x - bj.x) * (wdr * pd[0]) + (bi.y - bj.y) * (wdr * pd[1]);
                    if (ND == 3) vg += (bi.z - bj.z) * (wdr * pd[2]);
#pragma unroll
                    for (int d = 0; d < ND; ++d) acc[d] += dvp[d];
                    drho += div_fast(bi.w, bj.w) * (T)xj.w * vg;
                }
            }
        }
    });
    const int64_t o = (int64_t)orig * NV;
#pragma unroll
    for (int d = 0; d < ND; ++d) dv[o + d] += acc[d];
    if (NV == ND + 1) dv[o + ND] += drho;
}

// ------------------------------------------------------------------ BoundaryModelDummyParticles on the structure
// examples/fsi/hydrostatic_water_column_2d.jl:109-124 (and the alternative dam_break_plate_2d.jl:120-133 keeps in
// a comment): the structure particles are dummy particles with AdamiPressureExtrapolation.
template <typename T>
struct DummyConst {
    KernelConst<T> kern;   // boundary model's kernel / smoothing length
    EosConst<T> eos;       // boundary model's state equation
    T radius2_b;           // compact_support(structure, fluid)^2 = the MODEL's (neighborhood_search.jl:134-140)
    T radius2_f;           // compact_support(fluid, structure)^2 = the fluid's
    T acc[3];              // acceleration_source(fluid) - current_acceleration(structure) (= 0, abstract_system.jl:121)
    T p_off;
    int clip;
    T almostzero_fs;       // fluid <- structure: sqrt(eps(compact_support_fluid^2))
    T almostzero_sf;       // structure <- fluid: sqrt(eps(h_fluid^2))
    T bernoulli;           // BernoulliPressureExtrapolation's factor when the dynamic term applies, else 0
};

// update_boundary_interpolation! for the structure's particles at their CURRENT positions (dummy_particles.jl:
// 489-569, :637-672): one thread per sorted structure particle walks the fluid's sorted records.  Writes the
// density into the sorted record (Bs.w) and pressure / density in the structure's own particle order.
template <int ND, typename T, typename CT, int KERNEL>
__global__ void __launch_bounds__(128)
k_struct_adami(int n_s, GridConst<CT> g, const V4<CT> *__restrict__ As, const int *__restrict__ sperm,
               const int *__restrict__ fcell_start, const V4<CT> *__restrict__ A, const V4<T> *__restrict__ B,
               const T *__restrict__ P, int enabled, DummyConst<T> k, V4<T> *__restrict__ Bs, T *__restrict__ Ps,
               T *__restrict__ p_orig, T *__restrict__ rho_orig, int *__restrict__ flags,
               const T *__restrict__ a_clamped = nullptr, int n_int = 0)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_s) return;
    const V4<CT> xi = As[w];
    const int i = sperm[w];
    int cx, cy, cz;
    T p = (T)0, vol = (T)0;
    // resulting_acceleration = acceleration_source(fluid) - current_acceleration(system, particle): the prescribed
    // acceleration of a moving WALL particle (wall_boundary/system.jl:133-142, dummy_particles.jl:652-654)
    T acc[3] = {k.acc[0], k.acc[1], k.acc[2]};
    if (a_clamped != nullptr && i >= n_int) {
#pragma unroll
        for (int d = 0; d < ND; ++d) acc[d] -= a_clamped[(int64_t)(i - n_int) * ND + d];
    }
    const V4<T> vi = Bs[w];  // (velocity, .) of this particle, written by k_reorder_struct
    if (!cell_coords<ND, CT>(g, xi.x, xi.y, xi.z, cx, cy, cz)) atomicOr(flags, 1);
    else if (enabled) {
        for_neighbor_rows<ND, CT>(g, fcell_start, cx, cy, cz, [&](int j0, int j1) {
            for (int j = j0; j < j1; ++j) {
                const V4<CT> xj = A[j];
                T pd[3];
                const T d2 = pos_diff_d2<ND, T, CT>(xi, xj, pd);
                if (d2 <= k.radius2_b) {
                    const T dist = sqrt_rn(d2);
                    const V4<T> bj = B[j];
                    const T rho_f = bj.w;
                    T hyd = acc[0] * (rho_f * pd[0]) + acc[1] * (rho_f * pd[1]);
                    if (ND == 3) hyd += acc[2] * (rho_f * pd[2]);
                    T dyn = (T)0;
                    if (k.bernoulli != (T)0) {  // dummy_particles.jl:680-707
                        T vn = (vi.x - bj.x) * pd[0] + (vi.y - bj.y) * pd[1];
                        if (ND == 3) vn += (vi.z - bj.z) * pd[2];
                        vn = vn / dist;
                        dyn = k.bernoulli * rho_f * (vn * vn) / (T)2;
                    }
                    const T sum_p = k.p_off + P[j] + dyn + hyd;
                    const T kw = kernel_safe<KERNEL, T>(k.kern, dist);
                    p += sum_p * kw;
                    vol += kw;
                }
            }
        });
    }
    if ((double)vol > 2.220446049250313e-16) p = p / vol;  // `volume > eps()`: eps(Float64)
    if (k.clip) p = p > (T)0 ? p : (T)0;
    const T rho = eos_inverse(k.eos, p);
    Bs[w].w = rho;
    Ps[w] = p;
    p_orig[i] = p;
    rho_orig[i] = rho;
}

// structure <- fluid with dummy particles (interact_structure_fluid!, structure.jl:20-102): pairs within the
// model's compact support, the FLUID's kernel gradient and pressure-acceleration formulation with the roles
// switched: dv_boundary = -m_a^hydro (p_b + p_a) / (rho_b rho_a) grad W;  dv_s = sum dv_boundary m_b / m_a^material
template <int ND, typename T, typename CT, int FKERNEL, int DENS>
__global__ void __launch_bounds__(128)
k_struct_from_fluid_dummy(int n_int, GridConst<CT> g, const CT *__restrict__ x_cur, const T *__restrict__ mass_s,
                          const T *__restrict__ hydro_s, const T *__restrict__ p_s, const T *__restrict__ rho_s,
                          const int *__restrict__ fcell_start, const V4<CT> *__restrict__ A,
                          const V4<T> *__restrict__ B, const T *__restrict__ P, int enabled, KernelConst<T> fkern,
                          DummyConst<T> k, T *__restrict__ dv_s, int *__restrict__ flags)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n_int) return;
    V4<CT> xi;
    xi.x = x_cur[(int64_t)a * ND], xi.y = x_cur[(int64_t)a * ND + 1];
    xi.z = ND == 3 ? x_cur[(int64_t)a * ND + 2] : (CT)0;
    xi.w = (CT)0;
    T acc[3] = {0, 0, 0};
    int cx, cy, cz;
    if (!cell_coords<ND, CT>(g, xi.x, xi.y, xi.z, cx, cy, cz)) atomicOr(flags, 1);
    else if (enabled) {
        const T m_a = mass_s[a], mh_a = hydro_s[a], p_a = p_s[a], rho_a = rho_s[a];
        for_neighbor_rows<ND, CT>(g, fcell_start, cx, cy, cz, [&](int j0, int j1) {
            for (int j = j0; j < j1; ++j) {
                const V4<CT> xj = A[j];
                T pd[3];
                const T d2 = pos_diff_d2<ND, T, CT>(xi, xj, pd);
                if (d2 <= k.radius2_b) {
                    const T dist = sqrt_rn(d2);
                    if (dist >= k.almostzero_sf) {
                        const T wdr = SmoothingKernel<FKERNEL, T>::dw_div_r(fkern, dist);
                        const T rho_b = B[j].w, p_b = P[j], m_b = (T)xj.w;
                        T f;
                        if (DENS == 0)
                            f = -mh_a * div_fast(p_b + p_a, rho_b * rho_a);
                        else
                            f = -mh_a * (div_fast(p_b, rho_b * rho_b) + div_fast(p_a, rho_a * rho_a));
#pragma unroll
                        for (int d = 0; d < ND; ++d) acc[d] += (f * (wdr * pd[d])) * m_b / m_a;
                    }
                }
            }
        });
    }
#pragma unroll
    for (int d = 0; d < ND; ++d) dv_s[(int64_t)a * ND + d] = acc[d];
}

// fluid <- structure with dummy particles (wcsph/rhs.jl:5-127): like a wall particle, but with the structure's
// velocity in the continuity equation; no viscous term, no density diffusion.  dv_f += S_fs
template <int ND, typename T, typename CT, int KERNEL, int DENS>
__global__ void __launch_bounds__(128)
k_fluid_from_struct_dummy(int n_f, GridConst<CT> g, const int *__restrict__ fcell_start, const V4<CT> *__restrict__ A,
                          const V4<T> *__restrict__ B, const T *__restrict__ P, const int *__restrict__ perm,
                          const int *__restrict__ scell_start, const V4<CT> *__restrict__ As,
                          const V4<T> *__restrict__ Bs, const T *__restrict__ Ps, KernelConst<T> kern,
                          DummyConst<T> k, T *__restrict__ dv, int n_targets)
{
    constexpr int NV = DENS == 0 ? ND + 1 : ND;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_f || s >= fcell_start[g.ncells]) return;
    const int orig = perm[s];
    if (orig >= n_targets) return;
    const V4<CT> xi = A[s];
    int cx, cy, cz;
    cell_coords<ND, CT>(g, xi.x, xi.y, xi.z, cx, cy, cz);
    bool any = false;
    for_neighbor_rows<ND, CT>(g, scell_start, cx, cy, cz, [&](int j0, int j1) { any = any || j1 > j0; });
    if (!any) return;
    const V4<T> bi = B[s];
    const T p_a = P[s], rho_a = bi.w;
    T acc[3] = {0, 0, 0}, drho = (T)0;
    for_neighbor_rows<ND, CT>(g, scell_start, cx, cy, cz, [&](int j0, int j1) {
        for (int j = j0; j < j1; ++j) {
            const V4<CT> xj = As[j];
            T pd[3];
            const T d2 = pos_diff_d2<ND, T, CT>(xi, xj, pd);
            if (d2 <= k.radius2_f) {
                const T dist = sqrt_rn(d2);
                if (dist >= k.almostzero_fs) {
                    const V4<T> bj = Bs[j];
                    const T p_b = Ps[j], rho_b = bj.w, m_b = (T)xj.w;
                    const T wdr = SmoothingKernel<KERNEL, T>::dw_div_r(kern, dist);
                    T f;
                    if (DENS == 0)
                        f = -m_b * div_fast(p_a + p_b, rho_a * rho_b);
                    else
                        f = -m_b * (div_fast(p_a, rho_a * rho_a) + div_fast(p_b, rho_b * rho_b));
                    T vg = (bi.x - bj.x) * (wdr * pd[0]) + (bi.y - bj.y) * (wdr * pd[1]);
                    if (ND == 3) vg += (bi.z - bj.z) * (wdr * pd[2]);
#pragma unroll
                    for (int d = 0; d < ND; ++d) acc[d] += f * (wdr * pd[d]);
                    if (DENS == 0) drho += div_fast(rho_a, rho_b) * m_b * vg;
                }
            }
        }
    });
    const int64_t o = (int64_t)orig * NV;
#pragma unroll
    for (int d = 0; d < ND; ++d) dv[o + d] += acc[d];
    if (NV == ND + 1) dv[o + ND] += drho;
}

#undef TPB_MAT

}  // namespace tpb
