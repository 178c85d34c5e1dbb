// tpb_sweeps.cuh -- neighbour sweeps, variant 1 ("per-particle"): one thread per sorted
// particle walks the 3^(ND-1) contiguous neighbour-cell rows of the sorted records.
// This is the general-purpose path (every ND / precision / kernel / density combination)
// and the correctness anchor for the tiled variant in tpb_tiles.cuh.
//
// B200 counterpart of PointNeighbors `foreach_point_neighbor` + the loop bodies of
//   interact!                 /root/reference/src/schemes/fluid/weakly_compressible_sph/rhs.jl:5-127
//   summation_density!        /root/reference/src/general/density_calculators.jl:26-50
//   boundary_pressure_extrapolation! + compute_adami_density!
//                             /root/reference/src/schemes/boundary/wall_boundary/dummy_particles.jl:489-672
//   add_source_terms!         /root/reference/src/general/semidiscretization.jl:668-731
//   drift!                    /root/reference/src/general/semidiscretization.jl:522-571
#pragma once
#include "tpb_device.cuh"

namespace tpb {

// Visit the rows of the neighbourhood of cell (cx, cy, cz): row(j0, j1) is called with the
// sorted-index range [j0, j1) of the cells {cx - sx .. cx + sx} of each of the 3^(ND-1) rows.
template <int ND, typename CT, typename F>
__device__ __forceinline__ void for_neighbor_rows(const GridConst<CT> &g,
                                                  const int *__restrict__ cell_start, int cx,
                                                  int cy, int cz, F &&row)
{
#pragma unroll
    for (int dz = (ND == 3 ? -1 : 0); dz <= (ND == 3 ? 1 : 0); ++dz) {
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
            int c0 = cell_linear(g, cx - g.sx, cy + dy, cz + dz);
            row(cell_start[c0], cell_start[c0 + 2 * g.sx + 1]);
        }
    }
}

// ------------------------------------------------------------------ summation density
template <int ND, typename T, typename CT, int KERNEL>
__global__ void __launch_bounds__(128)
k_summation_density(int n_f_cap, GridConst<CT> g, const int *__restrict__ fcell_start,
                    const V4<CT> *__restrict__ A, int has_wall,
                    const int *__restrict__ wcell_start, const V4<CT> *__restrict__ Aw,
                    int wall_enabled, KernelConst<T> kern, T radius2, EosConst<T> eos,
                    V4<T> *__restrict__ B, T *__restrict__ P)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_f_cap || s >= fcell_start[g.ncells]) return;
    const V4<CT> xi = A[s];
    int cx, cy, cz;
    cell_coords<ND, CT>(g, xi.x, xi.y, xi.z, cx, cy, cz);
    T rho = (T)0;
    for_neighbor_rows<ND, CT>(g, fcell_start, cx, cy, cz, [&](int j0, int j1) {
        for (int j = j0; j < j1; ++j) {
            const V4<CT> xj = A[j];
            T pd[3];
            T d2 = pos_diff_d2<ND, T, CT>(xi, xj, pd);
            if (d2 <= radius2) rho += (T)xj.w * kernel_safe<KERNEL, T>(kern, sqrt_rn(d2));
        }
    });
    if (has_wall && wall_enabled) {
        for_neighbor_rows<ND, CT>(g, wcell_start, cx, cy, cz, [&](int j0, int j1) {
            for (int j = j0; j < j1; ++j) {
                const V4<CT> xj = Aw[j];
                T pd[3];
                T d2 = pos_diff_d2<ND, T, CT>(xi, xj, pd);
                if (d2 <= radius2) rho += (T)xj.w * kernel_safe<KERNEL, T>(kern, sqrt_rn(d2));
            }
        });
    }
    V4<T> b = B[s];
    b.w = rho;
    B[s] = b;
    P[s] = eos_pressure(eos, rho);
}

// ------------------------------------------------------------------ density reinitialisation
// Second pass of reinit_density! (wcsph/system.jl:398-415; callbacks/density_reinit.jl:83-121): the summation
// density rho~ is in the sorted records (B.w, first pass = the SummationDensity sweep); here the Shepard
// coefficient c_a = sum_b (m_b / rho_b) W_ab (general/corrections.jl:138-170) over the fluid (rho~_b) and the wall
// (boundary-model density), and rho_a = rho~_a / c_a goes into the density row of the caller's v_ode.
template <int ND, typename T, typename CT, int KERNEL>
__global__ void __launch_bounds__(128)
k_shepard_reinit(int n_f_cap, GridConst<CT> g, const int *__restrict__ fcell_start, const V4<CT> *__restrict__ A,
                 const V4<T> *__restrict__ B, const int *__restrict__ perm, int has_wall,
                 const int *__restrict__ wcell_start, const V4<CT> *__restrict__ Aw, const V2<T> *__restrict__ Ww,
                 KernelConst<T> kern, T radius2, int nv, int n_targets, T *__restrict__ v_out)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_f_cap || s >= fcell_start[g.ncells]) return;
    const int orig = perm[s];
    if (orig >= n_targets) return;
    const V4<CT> xi = A[s];
    int cx, cy, cz;
    cell_coords<ND, CT>(g, xi.x, xi.y, xi.z, cx, cy, cz);
    T c = (T)0;
    for_neighbor_rows<ND, CT>(g, fcell_start, cx, cy, cz, [&](int j0, int j1) {
        for (int j = j0; j < j1; ++j) {
            const V4<CT> xj = A[j];
            T pd[3];
            const T d2 = pos_diff_d2<ND, T, CT>(xi, xj, pd);
            if (d2 <= radius2) c += ((T)xj.w / B[j].w) * kernel_safe<KERNEL, T>(kern, sqrt_rn(d2));
        }
    });
    if (has_wall) {
        for_neighbor_rows<ND, CT>(g, wcell_start, cx, cy, cz, [&](int j0, int j1) {
            for (int j = j0; j < j1; ++j) {
                const V4<CT> xj = Aw[j];
                T pd[3];
                const T d2 = pos_diff_d2<ND, T, CT>(xi, xj, pd);
                if (d2 <= radius2) c += ((T)xj.w / Ww[j].y) * kernel_safe<KERNEL, T>(kern, sqrt_rn(d2));
            }
        });
    }
    v_out[(int64_t)orig * nv + ND] = B[s].w / c;
}

// ------------------------------------------------------------------ Adami extrapolation
// One thread per sorted wall particle; fluid neighbours from the fluid grid.
//   p_w = sum_f (p_off + p_f + rho_f (g - a_w).r_wf) W(r_wf) / sum_f W(r_wf)   if sum W > eps()
//   clip; rho_w = EOS^-1(p_w)
template <typename T>
struct AdamiConst {
    KernelConst<T> kern;   // boundary model's kernel / smoothing length
    EosConst<T> eos;       // boundary model's state equation
    T radius2;             // compact_support(wall, fluid)^2 (neighborhood_search.jl:134-140)
    T acc[3];              // acceleration_source(fluid) - current_acceleration(wall) (= 0)
    T p_off;
    int clip;
};

template <int ND, typename T, typename CT, int KERNEL>
__global__ void __launch_bounds__(128)
k_adami(int n_w, GridConst<CT> g, const V4<CT> *__restrict__ Aw,
        const int *__restrict__ fcell_start, const V4<CT> *__restrict__ A,
        const V4<T> *__restrict__ B, const T *__restrict__ P, int interaction_enabled,
        AdamiConst<T> k, V2<T> *__restrict__ W, T *__restrict__ volume)
{
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_w) return;
    const V4<CT> xi = Aw[w];
    int cx, cy, cz;
    cell_coords<ND, CT>(g, xi.x, xi.y, xi.z, cx, cy, cz);
    T p = (T)0, vol = (T)0;
    if (interaction_enabled) {
        for_neighbor_rows<ND, CT>(g, fcell_start, cx, cy, cz, [&](int j0, int j1) {
            for (int j = j0; j < j1; ++j) {
                const V4<CT> xj = A[j];
                T pd[3];
                T d2 = pos_diff_d2<ND, T, CT>(xi, xj, pd);
                if (d2 <= k.radius2) {
                    T dist = sqrt_rn(d2);
                    T rho_f = B[j].w;
                    T hyd = k.acc[0] * (rho_f * pd[0]) + k.acc[1] * (rho_f * pd[1]);
                    if (ND == 3) hyd += k.acc[2] * (rho_f * pd[2]);
                    T sum_p = k.p_off + P[j] + hyd;
                    T kw = kernel_safe<KERNEL, T>(k.kern, dist);
                    p += sum_p * kw;
                    vol += kw;
                }
            }
        });
    }
    if ((double)vol > 2.220446049250313e-16) p = p / vol;  // `volume > eps()`: eps(Float64)
    if (k.clip) p = p > (T)0 ? p : (T)0;
    V2<T> out;
    out.x = p;
    out.y = eos_inverse(k.eos, p);
    W[w] = out;
    volume[w] = vol;
}

// ------------------------------------------------------------------ dummy particles with ContinuityDensity
// The wall density is integrated (wall_boundary/system.jl:78-90): it arrives in the wall's rows of v_ode
// (ODE order), pressure = state_equation(density), clipped by the boundary model's own flag
// (compute_pressure!, apply_state_equation!, dummy_particles.jl:458-478).  Fills the sorted wall records
// (p, rho) the fluid's sweep reads.
template <typename T>
__global__ void __launch_bounds__(256)
k_wall_density_eos(int n_w, const T *__restrict__ v_wall, const int *__restrict__ perm_w, EosConst<T> eos,
                   int clip, V2<T> *__restrict__ W, T *__restrict__ volume, const AdaptConsts<T> *__restrict__ ad)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_w) return;
    if (ad) eos.B = ad->B_w;
    const T rho = v_wall[perm_w[w]];
    T p = eos_pressure(eos, rho);
    if (clip) p = p > (T)0 ? p : (T)0;
    V2<T> out;
    out.x = p;
    out.y = rho;
    W[w] = out;
    volume[w] = (T)0;
}

// interact!(wall, fluid) of such a wall (wall_boundary/rhs.jl:11-59): the continuity equation of every wall
// particle over its fluid neighbours with the WALL's kernel and smoothing length, v_a = 0 (static wall),
// almostzero = sqrt(eps(h_wall^2)); the fluid's density calculator picks the form (rhs.jl:63-79).  One
// thread per sorted wall particle; dv_wall is in ODE order and written exactly once (set_zero! included).
template <int ND, typename T, typename CT, int KERNEL>
__global__ void __launch_bounds__(128)
k_wall_continuity(int n_w, GridConst<CT> g, const V4<CT> *__restrict__ Aw, const V2<T> *__restrict__ Ww,
                  const int *__restrict__ perm_w, const int *__restrict__ fcell_start,
                  const V4<CT> *__restrict__ A, const V4<T> *__restrict__ B, int interaction_enabled,
                  KernelConst<T> kern, T radius2, T almostzero, int fluid_summation, T *__restrict__ dv_wall)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_w) return;
    const V4<CT> xi = Aw[w];
    const T rho_a = Ww[w].y;
    int cx, cy, cz;
    cell_coords<ND, CT>(g, xi.x, xi.y, xi.z, cx, cy, cz);
    T drho = (T)0;
    if (interaction_enabled) {
        for_neighbor_rows<ND, CT>(g, fcell_start, cx, cy, cz, [&](int j0, int j1) {
            for (int j = j0; j < j1; ++j) {
                const V4<CT> xj = A[j];
                T pd[3];
                const T d2 = pos_diff_d2<ND, T, CT>(xi, xj, pd);
                if (d2 > radius2) continue;
                const T dist = sqrt_rn(d2);
                if (dist < almostzero) continue;
                const T wdr = SmoothingKernel<KERNEL, T>::dw_div_r(kern, dist);
                const V4<T> bj = B[j];
                // dot(v_a - v_b, grad_kernel), v_a = 0, grad_kernel = wdr * pos_diff
                T vg = ((T)0 - bj.x) * (wdr * pd[0]) + ((T)0 - bj.y) * (wdr * pd[1]);
                if (ND == 3) vg += ((T)0 - bj.z) * (wdr * pd[2]);
                const T m_b = (T)xj.w;
                drho += fluid_summation ? m_b * vg : rho_a / bj.w * m_b * vg;
            }
        });
    }
    dv_wall[perm_w[w]] = drho;
}

// ------------------------------------------------------------------ interact! (variant 1)
template <typename T>
struct SourceConst {
    T acc[3];
    T damping;
    int any;
};

template <int ND, typename T, typename CT, int KERNEL, int DENS>
__global__ void __launch_bounds__(128)
k_interact_pp(int n_f, GridConst<CT> g, const int *__restrict__ fcell_start,
              const V4<CT> *__restrict__ A, const V4<T> *__restrict__ B,
              const T *__restrict__ P, const int *__restrict__ perm, int ff_enabled,
              int has_wall, const int *__restrict__ wcell_start,
              const V4<CT> *__restrict__ Aw, const V2<T> *__restrict__ Ww, PairConst<T> k,
              SourceConst<T> src, T *__restrict__ dv /* NV x n_f, ODE order */, int n_targets)
{
    constexpr int NV = DENS == 0 ? ND + 1 : ND;
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_f || s >= fcell_start[g.ncells]) return;
    if (perm[s] >= n_targets) return;  // slab ghost: neighbour only
    const V4<CT> xi = A[s];
    const V4<T> bi = B[s];
    const T p_a = P[s];
    const T rho_a = bi.w;
    const T v_a[3] = {bi.x, bi.y, bi.z};
    int cx, cy, cz;
    cell_coords<ND, CT>(g, xi.x, xi.y, xi.z, cx, cy, cz);

    T dv_ff[3] = {0, 0, 0}, drho_ff = 0;
    if (ff_enabled) {
        for_neighbor_rows<ND, CT>(g, fcell_start, cx, cy, cz, [&](int j0, int j1) {
            for (int j = j0; j < j1; ++j) {
                const V4<CT> xj = A[j];
                T pd[3];
                T d2 = pos_diff_d2<ND, T, CT>(xi, xj, pd);
                if (d2 <= k.radius2) {
                    T dist = sqrt_rn(d2);
                    if (dist >= k.almostzero) {
                        const V4<T> bj = B[j];
                        const T v_b[3] = {bj.x, bj.y, bj.z};
                        interact_pair<ND, T, KERNEL, DENS, true>(k, (T)xj.w, rho_a, bj.w, p_a,
                                                                P[j], v_a, v_b, pd, dist, dv_ff,
                                                                drho_ff, (T)xi.w);
                    }
                }
            }
        });
    }
    T dv_fw[3] = {0, 0, 0}, drho_fw = 0;
    if (has_wall) {
        const T zero3[3] = {0, 0, 0};
        for_neighbor_rows<ND, CT>(g, wcell_start, cx, cy, cz, [&](int j0, int j1) {
            for (int j = j0; j < j1; ++j) {
                const V4<CT> xj = Aw[j];
                T pd[3];
                T d2 = pos_diff_d2<ND, T, CT>(xi, xj, pd);
                if (d2 <= k.radius2) {
                    T dist = sqrt_rn(d2);
                    if (dist >= k.almostzero) {
                        const V2<T> wj = Ww[j];
                        interact_pair<ND, T, KERNEL, DENS, false>(k, (T)xj.w, rho_a, wj.y, p_a,
                                                                 wj.x, v_a, zero3, pd, dist,
                                                                 dv_fw, drho_fw);
                    }
                }
            }
        });
    }
    // dv = ((0 + S_ff) + S_fw) + g [+ source]: same association as set_zero!, the two
    // interact! calls and add_source_terms! (semidiscretization.jl:600, :809-829, :668-731)
    const int64_t o = (int64_t)perm[s] * NV;
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        T val = dv_ff[d] + dv_fw[d];
        if (src.any) {
            val += src.acc[d];
            if (src.damping != (T)0) val += -src.damping * v_a[d];
        }
        dv[o + d] = val;
    }
    if (DENS == 0) dv[o + ND] = drho_ff + drho_fw;
}

// ------------------------------------------------------------------ no-slip wall
// `BoundaryModelDummyParticles(...; viscosity=model)`: the wall carries a velocity
//   v_w = 2 v_boundary - sum_f v_f W(r_wf) / sum_f W(r_wf)        (v_boundary = 0: static wall)
// (interpolate_fluid_velocity! / compute_wall_velocity!, dummy_particles.jl:710-758) and the fluid
// feels the wall through the wall's viscosity model with v_b = v_w (dv_viscosity!, viscosity.jl:9-40;
// viscous_velocity, wall_boundary/system.jl:148-163).  Two sweeps of their own, launched after the
// Adami pass and after interact!, so that the register budgets of k_adami_tiles / k_interact_tiles
// (the free-slip hot path) stay untouched.
template <int ND, typename T, typename CT, int KERNEL>
__global__ void __launch_bounds__(128)
k_wall_velocity(int n_w, GridConst<CT> g, const V4<CT> *__restrict__ Aw,
                const int *__restrict__ fcell_start, const V4<CT> *__restrict__ A,
                const V4<T> *__restrict__ B, int interaction_enabled, KernelConst<T> kern, T radius2,
                const V2<T> *__restrict__ Ww, V4<T> *__restrict__ Vw /* (v_w, rho_w) */,
                T *__restrict__ Pw /* p_w as a scalar array */)
{
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_w) return;
    const V4<CT> xi = Aw[w];
    int cx, cy, cz;
    cell_coords<ND, CT>(g, xi.x, xi.y, xi.z, cx, cy, cz);
    T wv[3] = {0, 0, 0}, vol = (T)0;
    if (interaction_enabled) {
        for_neighbor_rows<ND, CT>(g, fcell_start, cx, cy, cz, [&](int j0, int j1) {
            for (int j = j0; j < j1; ++j) {
                const V4<CT> xj = A[j];
                T pd[3];
                T d2 = pos_diff_d2<ND, T, CT>(xi, xj, pd);
                if (d2 <= radius2) {
                    T kw = kernel_safe<KERNEL, T>(kern, sqrt_rn(d2));
                    const V4<T> bj = B[j];
                    wv[0] += kw * bj.x;
                    wv[1] += kw * bj.y;
                    if (ND == 3) wv[2] += kw * bj.z;
                    vol += kw;
                }
            }
        });
    }
    if ((double)vol > 2.220446049250313e-16) {  // `volume > eps()` (dummy_particles.jl:554)
#pragma unroll
        for (int d = 0; d < ND; ++d) wv[d] = (T)0 - wv[d] / vol;
    }
    V4<T> out;
    out.x = wv[0];
    out.y = wv[1];
    out.z = ND == 3 ? wv[2] : (T)0;
    const V2<T> pr = Ww[w];
    out.w = pr.y;
    Vw[w] = out;
    Pw[w] = pr.x;
}

template <typename T>
struct WallViscConst {
    KernelConst<T> kern;  // the fluid's kernel (gradient)
    int model;            // TPB_VISCOSITY_* of the boundary model
    T alpha, beta;        // ArtificialViscosityMonaghan of the wall
    T nu_a, nu_b;         // Morris / Adami: kinematic_viscosity of the fluid / of the wall
    T h;                  // (h_fluid + h_wall) / 2
    T eps_h2;             // epsilon_wall * h^2
    T c;                  // system_sound_speed(fluid)
    T radius2, almostzero;
};

// The wall model's viscous term of one accepted pair (fluid particle a, wall particle b with the
// record wj = (v_w, rho_w)); pd = x_a - x_b, dist >= almostzero.
template <int ND, typename T, int KERNEL>
__device__ __forceinline__ void wall_viscous_term(const WallViscConst<T> &k, T m_a, T rho_a, const T (&v_a)[3],
                                                  T m_b, const V4<T> &wj, const T (&pd)[3], T dist,
                                                  T (&acc)[3])
{
    const T wdr = SmoothingKernel<KERNEL, T>::dw_div_r(k.kern, dist);
    const T rho_b = wj.w;
    const T vd[3] = {v_a[0] - wj.x, v_a[1] - wj.y, ND == 3 ? v_a[2] - wj.z : (T)0};
    T grad[3] = {0, 0, 0};
#pragma unroll
    for (int d = 0; d < ND; ++d) grad[d] = wdr * pd[d];
    const T d2e = dist * dist + k.eps_h2;
    if (k.model == 1) {
        // ArtificialViscosityMonaghan (viscosity.jl:89-132)
        T vr = vd[0] * pd[0] + vd[1] * pd[1];
        if (ND == 3) vr += vd[2] * pd[2];
        if (vr < (T)0) {
            const T rho_mean = (rho_a + rho_b) / (T)2;
            const T mu = div_fast(k.h * vr, d2e);
            const T dvv = div_fast(m_b * k.alpha * k.c * mu + m_b * k.beta * (mu * mu), rho_mean);
#pragma unroll
            for (int d = 0; d < ND; ++d) acc[d] += dvv * grad[d];
        }
    } else {
        T pg = pd[0] * grad[0] + pd[1] * grad[1];
        if (ND == 3) pg += pd[2] * grad[2];
        T coef;
        if (k.model == 2) {
            // ViscosityMorris (viscosity.jl:163-205)
            const T mu_a = k.nu_a * rho_a, mu_b = k.nu_b * rho_b;
            coef = div_fast(m_b * (mu_a + mu_b) * pg, rho_a * rho_b * d2e);
        } else {
            // ViscosityAdami (viscosity.jl:222-279)
            const T eta_a = k.nu_a * rho_a, eta_b = k.nu_b * rho_b;
            const T volume_a = div_fast(m_a, rho_a), volume_b = div_fast(m_b, rho_b);
            const T tmp = div_fast((T)2 * eta_a * eta_b, (eta_a + eta_b) * d2e * m_a);
            coef = (volume_a * volume_a + volume_b * volume_b) * pg * tmp;
        }
#pragma unroll
        for (int d = 0; d < ND; ++d) acc[d] += coef * vd[d];
    }
}

// One thread per sorted fluid particle: dv[1:ND, a] += sum_w viscous term (wall's model).
template <int ND, typename T, typename CT, int KERNEL, int NV>
__global__ void __launch_bounds__(128)
k_wall_viscous(int n_f, GridConst<CT> g, const int *__restrict__ fcell_start,
               const V4<CT> *__restrict__ A, const V4<T> *__restrict__ B,
               const int *__restrict__ perm, const int *__restrict__ wcell_start,
               const V4<CT> *__restrict__ Aw, const V2<T> *__restrict__ Ww,
               const V4<T> *__restrict__ Vw, WallViscConst<T> k, T *__restrict__ dv, int n_targets)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_f || s >= fcell_start[g.ncells]) return;
    if (perm[s] >= n_targets) return;  // slab ghost: neighbour only
    const V4<CT> xi = A[s];
    const V4<T> bi = B[s];
    const T rho_a = bi.w, m_a = (T)xi.w;
    const T v_a[3] = {bi.x, bi.y, bi.z};
    int cx, cy, cz;
    cell_coords<ND, CT>(g, xi.x, xi.y, xi.z, cx, cy, cz);
    T acc[3] = {0, 0, 0};
    for_neighbor_rows<ND, CT>(g, wcell_start, cx, cy, cz, [&](int j0, int j1) {
        for (int j = j0; j < j1; ++j) {
            const V4<CT> xj = Aw[j];
            T pd[3];
            T d2 = pos_diff_d2<ND, T, CT>(xi, xj, pd);
            if (d2 > k.radius2) continue;
            const T dist = sqrt_rn(d2);
            if (dist < k.almostzero) continue;
            wall_viscous_term<ND, T, KERNEL>(k, m_a, rho_a, v_a, (T)xj.w, Vw[j], pd, dist, acc);
        }
    });
    const int64_t o = (int64_t)perm[s] * NV;
#pragma unroll
    for (int d = 0; d < ND; ++d) dv[o + d] += acc[d];
}

// scatter a sorted per-particle vector field (V4 records) back to the system's own particle order
template <typename T>
__global__ void __launch_bounds__(256)
k_unsort_vector(int n, const int *__restrict__ n_sorted, const int *__restrict__ perm,
                const V4<T> *__restrict__ sorted, int nd, T *__restrict__ out)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n || s >= *n_sorted) return;
    const V4<T> r = sorted[s];
    T *o = out + (int64_t)perm[s] * nd;
    o[0] = r.x;
    o[1] = r.y;
    if (nd == 3) o[2] = r.z;
}

// ------------------------------------------------------------------ drift!
// du[1:ND, a] = v[1:ND, a]  (strided copy: v has NV rows, du has ND; T -> cT conversion)
template <int ND, typename T, typename CT>
__global__ void __launch_bounds__(256)
k_drift(int64_t n_total /* n_f * ND */, int nv, const T *__restrict__ v, CT *__restrict__ du)
{
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_total) return;
    int64_t a = q / ND;
    int d = (int)(q - a * ND);
    du[q] = (CT)v[a * nv + d];
}

// drift! of two systems in one launch: the fluid (nv entries per particle in v) and the structure (ND entries)
template <int ND, typename T, typename CT>
__global__ void __launch_bounds__(256)
k_drift2(int64_t n_f /* n_fluid * ND */, int nv, const T *__restrict__ v_f, CT *__restrict__ du_f,
         int64_t n_s /* n_integrated * ND */, const T *__restrict__ v_s, CT *__restrict__ du_s)
{
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n_f) {
        const int64_t a = q / ND;
        const int d = (int)(q - a * ND);
        du_f[q] = (CT)v_f[a * nv + d];
    } else if (q - n_f < n_s) {
        du_s[q - n_f] = (CT)v_s[q - n_f];
    }
}

// test hook behind tpb_vec_div_fast: out[i] = div_fast(x, y[i]) (test/examples/gpu.jl:30-78)
template <typename T>
__global__ void __launch_bounds__(256)
k_div_fast(int64_t n, T x, const T *__restrict__ y, T *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = div_fast(x, y[i]);
}

// ------------------------------------------------------------------ neighbour pair dump
// Test hook behind tpb_neighbor_pairs: appends (orig_i, orig_j) for every accepted pair.
template <int ND, typename T, typename CT>
__global__ void __launch_bounds__(128)
k_pairs(int n_x, const int *__restrict__ n_sorted_x, GridConst<CT> g, const V4<CT> *__restrict__ X,
        const int *__restrict__ perm_x,
        const int *__restrict__ ycell_start, const V4<CT> *__restrict__ Y,
        const int *__restrict__ perm_y, T radius2, long long capacity, int *__restrict__ out_i,
        int *__restrict__ out_j, unsigned long long *__restrict__ counter)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_x || s >= *n_sorted_x) return;
    const V4<CT> xi = X[s];
    int cx, cy, cz;
    cell_coords<ND, CT>(g, xi.x, xi.y, xi.z, cx, cy, cz);
    for_neighbor_rows<ND, CT>(g, ycell_start, cx, cy, cz, [&](int j0, int j1) {
        for (int j = j0; j < j1; ++j) {
            T pd[3];
            T d2 = pos_diff_d2<ND, T, CT>(xi, Y[j], pd);
            if (d2 <= radius2) {
                unsigned long long at = atomicAdd(counter, 1ull);
                if ((long long)at < capacity) {
                    out_i[at] = perm_x[s];
                    out_j[at] = perm_y[j];
                }
            }
        }
    });
}

// scatter a sorted per-particle field back to the system's own particle order
template <typename T>
__global__ void __launch_bounds__(256)
k_unsort_scalar(int n, const int *__restrict__ n_sorted, const int *__restrict__ perm,
                const T *__restrict__ sorted, int stride, int offset, T *__restrict__ out)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n || s >= *n_sorted) return;
    out[perm[s]] = sorted[(int64_t)s * stride + offset];
}

}  // namespace tpb
