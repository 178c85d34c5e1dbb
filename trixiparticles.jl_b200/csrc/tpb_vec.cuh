// tpb_vec.cuh -- fused algebra on device-resident ODE vectors, so that the integrator's
// broadcasts between RHS evaluations never leave the GPU.  Counterpart of the broadcast
// overloads of `ThreadedBroadcastArray` (/root/reference/src/util.jl:183-303) and of the
// 2N-storage Runge-Kutta stage update OrdinaryDiffEq performs on `ArrayPartition(v_ode, u_ode)`.
// Pure streaming kernels (HBM-bound): 16-byte vector accesses, grid-stride, arithmetic in
// double and rounded on store (Julia promotes Float64 dt * Float32 entries the same way).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "tpb_device.cuh"

namespace tpb {

template <typename T>
__global__ void __launch_bounds__(256)
k_vec_axpby(int64_t n, double a, const T *__restrict__ x, double b, T *__restrict__ y)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        y[i] = (T)(a * (double)x[i] + (b == 0.0 ? 0.0 : b * (double)y[i]));
}

// tmp = A * tmp + dt * rhs;  state += B * tmp      (Williamson 2N-storage stage)
template <typename T>
__global__ void __launch_bounds__(256)
k_vec_rk2n_stage(int64_t n, double A, double B, double dt, const T *__restrict__ rhs,
                 T *__restrict__ tmp, T *__restrict__ state)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const T t = (T)((A == 0.0 ? 0.0 : A * (double)tmp[i]) + dt * (double)rhs[i]);
        tmp[i] = t;
        state[i] = (T)((double)state[i] + B * (double)t);
    }
}

// y = a0 x0 + a1 x1 + a2 x2 + a3 x3 (terms with a null pointer are skipped; any x may alias y): the
// register updates of a low-storage 3S*+ Runge-Kutta stage (RDPK3SpFSAL35) in one pass
template <typename T>
__global__ void __launch_bounds__(256)
k_vec_lincomb4(int64_t n, double a0, const T *x0, double a1, const T *x1, double a2, const T *x2, double a3,
               const T *x3, T *y)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double s = 0.0;
        if (x0) s += a0 * (double)x0[i];
        if (x1) s += a1 * (double)x1[i];
        if (x2) s += a2 * (double)x2[i];
        if (x3) s += a3 * (double)x3[i];
        y[i] = (T)s;
    }
}

// SymplecticPositionVerlet, velocity / density update of one WCSPH system
// (ext/TrixiParticlesOrdinaryDiffEqSymplecticRKExt.jl:137-171): rows of NV entries per particle;
// du[0:ND] = duprev + dt * kdu; with the density as last entry (NV == ND + 1):
//   epsilon = -kdu[end] / du[end] * dt;  du[end] = duprev[end] * (2 - epsilon) / (2 + epsilon)
// (du holds the half-step state on entry).  Arithmetic in T, as the reference's broadcast.
template <typename T>
__global__ void __launch_bounds__(256)
k_vec_verlet_update(int64_t n_particles, int nd, int nv, T dt, const T *__restrict__ kdu,
                    const T *__restrict__ duprev, T *__restrict__ du)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_particles; p += stride) {
        const int64_t o = p * nv;
        for (int d = 0; d < nd; ++d) du[o + d] = duprev[o + d] + dt * kdu[o + d];
        if (nv > nd) {
            const T density_prev = duprev[o + nd], density_half = du[o + nd];
            const T epsilon = -kdu[o + nd] / density_half * dt;
            du[o + nd] = density_prev * ((T)2 - epsilon) / ((T)2 + epsilon);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
k_vec_fill(int64_t n, double value, T *__restrict__ x)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] = (T)value;
}

// max over x[offset + k * stride], k < count  (e.g. `max_x_coord`: custom_quantities.jl)
// out must hold the identity (-inf as ordered int bits) before the launch.
__device__ __forceinline__ void atomic_max_double(double *addr, double v)
{
    unsigned long long *p = (unsigned long long *)addr;
    unsigned long long old = *p, assumed;
    do {
        assumed = old;
        if (__longlong_as_double((long long)assumed) >= v) break;
        old = atomicCAS(p, assumed, (unsigned long long)__double_as_longlong(v));
    } while (assumed != old);
}

template <typename T>
__global__ void __launch_bounds__(256)
k_vec_strided_max(int64_t count, int stride, int offset, const T *__restrict__ x, double *__restrict__ out)
{
    double m = -1.0 / 0.0;
    const int64_t gs = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gs)
        m = fmax(m, (double)x[k * stride + offset]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ double wm[8];
    if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, wm[w]);
        atomic_max_double(out, m);  // order-independent: max is exact
    }
}

// partial[b] = sum over block b of (e_i / (abstol + reltol * max(|a_i|, |b_i|)))^2 in double: the
// residual norm of an adaptive integrator (OrdinaryDiffEq `calculate_residuals` +
// `ODE_DEFAULT_NORM`); the host adds the per-block partial sums in block order (deterministic).
template <typename T>
__global__ void __launch_bounds__(256)
k_vec_wrms_partial(int64_t n, const T *__restrict__ e, const T *__restrict__ a, const T *__restrict__ b,
                   double abstol, double reltol, double *__restrict__ partial)
{
    double s = 0.0;
    const int64_t gs = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gs) {
        const double r = (double)e[i] / (abstol + reltol * fmax(fabs((double)a[i]), fabs((double)b[i])));
        s += r * r;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ double ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) s += ws[w];
        partial[blockIdx.x] = s;
    }
}

// max_i |v_i|^2 over the first ND entries of every row of v (update_speed_of_sound!,
// wcsph/system.jl:307-315): products and sums separately rounded, left to right; the maximum of
// non-negative IEEE numbers is the maximum of their bit patterns.
template <int ND, typename T>
__global__ void __launch_bounds__(256)
k_max_speed2(int64_t n, int nv, const T *__restrict__ v, unsigned long long *__restrict__ out_bits)
{
    unsigned long long best = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const T *r = v + i * nv;
        T s;
        if constexpr (sizeof(T) == 4) {
            s = __fadd_rn(__fmul_rn(r[0], r[0]), __fmul_rn(r[1], r[1]));
            if (ND == 3) s = __fadd_rn(s, __fmul_rn(r[2], r[2]));
            best = max(best, (unsigned long long)__float_as_uint(s));
        } else {
            s = __dadd_rn(__dmul_rn(r[0], r[0]), __dmul_rn(r[1], r[1]));
            if (ND == 3) s = __dadd_rn(s, __dmul_rn(r[2], r[2]));
            best = max(best, (unsigned long long)__double_as_longlong(s));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0 && best) atomicMax(out_bits, best);
}

// ------------------------------------------------------------------ StateEquationAdaptiveCole on the device
// update_speed_of_sound! (wcsph/system.jl:307-321, state_equations.jl:37-84) without a host round trip:
// one thread turns max|v|^2 into the new speed of sound and into every constant of the kick that follows
// it, with the operations (and roundings) of the host path (update_sound_speed / make_eos_const /
// make_pair_const / make_wall_visc_const in tpb200.cu).  The tile kernels read the struct when they are
// handed its address.
template <typename T>
struct AdaptParams {
    T mach, c_min, c_max;
    float mach_f, c_min_f, c_max_f;  // the state equation's own Float32 fields (params_f32)
    int params_f32;
    T gamma_f, rho0_f;
    float gamma_f32, rho0_f32;
    int wall_follows;  // the boundary model shares the fluid's state equation
    T gamma_w, rho0_w, pbg_w, inv_gamma_w;
    float gamma_w32, rho0_w32;
    T delta_h;         // delta * (h + h) / 2
    int visc_f, visc_w, nd;
    T alpha_f, alpha_w, h_f, h_w;
};
template <typename T>
__global__ void k_adaptive_consts(const unsigned long long *__restrict__ vmax2_bits, AdaptParams<T> p,
                                  AdaptConsts<T> *__restrict__ out, double *__restrict__ c_out)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    T v_max2;
    if constexpr (sizeof(T) == 4)
        v_max2 = __uint_as_float((unsigned)*vmax2_bits);
    else
        v_max2 = __longlong_as_double((long long)*vmax2_bits);
    double c;  // the value tpb_get_sound_speed reports (the state equation's Ref)
    if constexpr (sizeof(T) == 4) {
        const float q = __fdiv_rn(__fsqrt_rn(v_max2), p.mach);
        c = (double)fminf(p.c_max, fmaxf(p.c_min, q));
    } else if (p.params_f32) {
        const double q = __ddiv_rn(__dsqrt_rn(v_max2), (double)p.mach_f);
        c = (double)(float)fmin((double)p.c_max_f, fmax((double)p.c_min_f, q));
    } else {
        const double q = __ddiv_rn(__dsqrt_rn(v_max2), p.mach);
        c = fmin(p.c_max, fmax(p.c_min, q));
    }
    const T ct = (T)c;
    auto bulk = [&](T rho0, T gamma, float rho0_32, float gamma_32) -> T {
        if (p.params_f32) return (T)__fdiv_rn(__fmul_rn(rho0_32, __fmul_rn((float)c, (float)c)), gamma_32);
        if constexpr (sizeof(T) == 4)
            return __fdiv_rn(__fmul_rn(rho0, __fmul_rn(ct, ct)), gamma);
        else
            return __ddiv_rn(__dmul_rn(rho0, __dmul_rn(ct, ct)), gamma);
    };
    AdaptConsts<T> a;
    a.c = ct;
    a.B_f = bulk(p.rho0_f, p.gamma_f, p.rho0_f32, p.gamma_f32);
    a.B_w = p.wall_follows ? bulk(p.rho0_w, p.gamma_w, p.rho0_w32, p.gamma_w32) : (T)0;
    a.delta_h_c = p.delta_h * ct;
    // kinematic_viscosity of ArtificialViscosityMonaghan: alpha h c / (2 ND + 4) (viscosity.jl:82-87)
    auto kin = [&](int model, T alpha_or_nu, T h) {
        return model == 1 ? alpha_or_nu * h * ct / (T)(2 * p.nd + 4) : alpha_or_nu;
    };
    a.nu_a = kin(p.visc_f, p.alpha_f, p.h_f);
    a.nu_b = kin(p.visc_w, p.alpha_w, p.h_w);
    // density of a wall particle without fluid in reach: inverse state equation of p = 0
    a.rho_empty_w = p.wall_follows
                        ? p.rho0_w * (T)pow((double)(((T)0 - p.pbg_w) / a.B_w + (T)1), (double)p.inv_gamma_w)
                        : (T)0;
    a.pad = (T)0;
    *out = a;
    *c_out = c;
}

}  // namespace tpb
