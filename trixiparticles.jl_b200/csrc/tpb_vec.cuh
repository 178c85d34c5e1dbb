// tpb_vec.cuh -- fused algebra on device-resident ODE vectors, so that the integrator's
// broadcasts between RHS evaluations never leave the GPU.  Counterpart of the broadcast
// overloads of `ThreadedBroadcastArray` (/root/reference/src/util.jl:183-303) and of the
// 2N-storage Runge-Kutta stage update OrdinaryDiffEq performs on `ArrayPartition(v_ode, u_ode)`.
// Pure streaming kernels (HBM-bound): 16-byte vector accesses, grid-stride, arithmetic in
// double and rounded on store (Julia promotes Float64 dt * Float32 entries the same way).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

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

}  // namespace tpb
