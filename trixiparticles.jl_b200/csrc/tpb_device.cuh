// tpb_device.cuh -- device-side building blocks of the WCSPH right-hand side:
// vector records, exactly-rounded neighbour predicate, smoothing kernels, Cole equation of
// state and the per-pair physics.  Hand-written for sm_100a; no library calls.
//
// Paths cited are relative to /root/reference.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tpb {

// ---------------------------------------------------------------------------------------
// Sorted particle records (HBM layout, see DESIGN.md "Data layout"):
//   A[j] = (x, y, z, m)      in cT   -- position + hydrodynamic mass
//   B[j] = (vx, vy, vz, rho) in T    -- velocity + density          (fluid only)
//   P[j] = p                 in T    -- pressure                    (fluid)
//   W[j] = (p, rho)          in T    -- Adami pressure / density    (wall)
// 16-byte records make every neighbour-cell run a 16-byte aligned, contiguous stream.
template <typename T>
struct alignas(16) V4 {
    T x, y, z, w;
};
template <typename T>
struct alignas(2 * sizeof(T)) V2 {
    T x, y;
};

// ---------------------------------------------------------------------------------------
// Exactly rounded primitives.  The neighbour predicate must reproduce
//   pos_diff = convert.(T, x_i - y_j); d2 = dot(pos_diff, pos_diff); d2 <= R^2
// bit for bit (PointNeighbors.jl foreach_neighbor; StaticArrays `dot` is a left-to-right
// sum of products, Julia never contracts to fma), so nvcc must not fuse these.
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float sqrt_rn(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double sqrt_rn(double a) { return __dsqrt_rn(a); }

// div_fast (util.jl:3-5).  Float32: the reference's GPU path uses LLVM fast division, mirrored
// here by the approximate (<= 2 ulp) hardware divide.  Float64: the reference's CUDA extension
// (ext/TrixiParticlesCUDAExt.jl:12-33) overrides it with x * (rcp.approx.ftz.f64 refined by one
// cubic iteration), relative error < 1e-15 -- the same here; the IEEE division it replaces costs
// about three times as many FP64 instructions.
__device__ __forceinline__ float div_fast(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ double div_fast(double a, double b)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    double e = fma(r, -b, 1.0);
    e = fma(e, e, e);
    r = fma(e, r, r);
    return a * r;
}

template <int ND, typename T, typename CT>
__device__ __forceinline__ T pos_diff_d2(const V4<CT> &xi, const V4<CT> &xj, T (&pd)[3])
{
    pd[0] = (T)sub_rn(xi.x, xj.x);
    pd[1] = (T)sub_rn(xi.y, xj.y);
    T d2 = add_rn(mul_rn(pd[0], pd[0]), mul_rn(pd[1], pd[1]));
    if (ND == 3) {
        pd[2] = (T)sub_rn(xi.z, xj.z);
        d2 = add_rn(d2, mul_rn(pd[2], pd[2]));
    } else {
        pd[2] = (T)0;
    }
    return d2;
}

// ---------------------------------------------------------------------------------------
// Smoothing kernels.  Constants are prepared on the host in T with the reference's own
// operation order (smoothing_kernels.jl:193-227, :436-455).
template <typename T>
struct KernelConst {
    T h;        // smoothing length
    T h_inv;    // 1 / h
    T nf;       // normalization_factor(kernel, h_inv) = sigma * h_inv^ND
    T m5nf;     // -5 * nf                (Wendland C2 derivative)
    T h_inv2;   // h_inv * h_inv
    T support;  // compact_support: 2h; quartic spline 5/2 h; quintic spline 3h
    T c_d;      // Wendland C4 / C6: derivative prefactor (-7/3 or -11/4) in T
    int order;  // Wendland C4 / C6 (SmoothingKernel<2>): 4 or 6; Schoenberg quartic / quintic (<3>): 4 or 5
};

template <int KERNEL, typename T>
struct SmoothingKernel;

// WendlandC2Kernel: W = nf (1-q/2)^4 (2q+1);  (dW/dr)/r = -5 nf (1-q/2)^3 h^-2
template <typename T>
struct SmoothingKernel<0, T> {
    static __device__ __forceinline__ T w_unsafe(const KernelConst<T> &k, T r)
    {
        T q = r * k.h_inv;
        T t = (T)1 - q / (T)2;
        T t2 = t * t;
        return k.nf * (t2 * t2) * ((T)2 * q + (T)1);
    }
    static __device__ __forceinline__ T dw_div_r(const KernelConst<T> &k, T r)
    {
        T q = r * k.h_inv;
        T t = (T)1 - q / (T)2;
        return k.m5nf * (t * t * t) * k.h_inv2;
    }
};

// SchoenbergCubicSplineKernel: W = nf [(2-q)^3/4 - (q<1)(1-q)^3];
// (dW/dr)/r = div_fast(nf [-3(2-q)^2/4 + 3(q<1)(1-q)^2] h^-1, r)
template <typename T>
struct SmoothingKernel<1, T> {
    static __device__ __forceinline__ T w_unsafe(const KernelConst<T> &k, T r)
    {
        T q = r * k.h_inv;
        T a = (T)2 - q, b = (T)1 - q;
        T lt = q < (T)1 ? (T)1 : (T)0;
        return k.nf * ((a * a * a) / (T)4 - lt * (b * b * b));
    }
    static __device__ __forceinline__ T dw_div_r(const KernelConst<T> &k, T r)
    {
        T q = r * k.h_inv;
        T a = (T)2 - q, b = (T)1 - q;
        T three_lt = q < (T)1 ? (T)3 : (T)0;
        T result = (T)(-3) * (a * a) / (T)4 + three_lt * (b * b);
        return div_fast(k.nf * result * k.h_inv, r);
    }
};

// WendlandC4Kernel / WendlandC6Kernel (smoothing_kernels.jl:489-514, :548-574), one template
// value, the order is a warp-uniform run-time switch:
//   C4: W = nf (1-q/2)^6 (35 q^2/12 + 3q + 1);        (dW/dr)/r = nf (-7/3)(2 + 5q)(1-q/2)^5 h^-2
//   C6: W = nf (1-q/2)^8 (4q^3 + 25 q^2/4 + 4q + 1);  (dW/dr)/r = nf (-11/4)(8q^2 + 7q + 2)(1-q/2)^7 h^-2
template <typename T>
struct SmoothingKernel<2, T> {
    static __device__ __forceinline__ T w_unsafe(const KernelConst<T> &k, T r)
    {
        T q = r * k.h_inv;
        T t = (T)1 - q / (T)2;
        T t2 = t * t;
        if (k.order == 4) {
            T t6 = t2 * t2 * t2;
            return k.nf * (t6 * ((T)35 * (q * q) / (T)12 + (T)3 * q + (T)1));
        }
        T t4 = t2 * t2;
        T t8 = t4 * t4;
        return k.nf * (t8 * ((T)4 * (q * q * q) + (T)25 * (q * q) / (T)4 + (T)4 * q + (T)1));
    }
    static __device__ __forceinline__ T dw_div_r(const KernelConst<T> &k, T r)
    {
        T q = r * k.h_inv;
        T t = (T)1 - q / (T)2;
        T t2 = t * t;
        if (k.order == 4) {
            T t5 = t2 * t2 * t;
            return k.nf * (k.c_d * ((T)2 + (T)5 * q) * t5 * k.h_inv2);
        }
        T t7 = t2 * t2 * t2 * t;
        return k.nf * (k.c_d * ((T)8 * (q * q) + (T)7 * q + (T)2) * t7 * k.h_inv2);
    }
};

// SchoenbergQuarticSplineKernel / SchoenbergQuinticSplineKernel (smoothing_kernels.jl:264-395), one
// template value, the order is a warp-uniform run-time switch:
//   quartic: W = nf [(5/2-q)^4 - 5 (q<3/2)(3/2-q)^4 + 10 (q<1/2)(1/2-q)^4]
//   quintic: W = nf [(3-q)^5 - 6 (q<2)(2-q)^5 + 15 (q<1)(1-q)^5];   (dW/dr)/r = div_fast(nf W'(q) h^-1, r)
template <typename T>
struct SmoothingKernel<3, T> {
    static __device__ __forceinline__ T w_unsafe(const KernelConst<T> &k, T r)
    {
        T q = r * k.h_inv;
        if (k.order == 4) {
            T a = (T)2.5 - q, b = (T)1.5 - q, c = (T)0.5 - q;
            T a2 = a * a, b2 = b * b, c2 = c * c;
            T lt15 = q < (T)1.5 ? (T)1 : (T)0, lt05 = q < (T)0.5 ? (T)1 : (T)0;
            return k.nf * (a2 * a2 - (T)5 * lt15 * (b2 * b2) + (T)10 * lt05 * (c2 * c2));
        }
        T a = (T)3 - q, b = (T)2 - q, c = (T)1 - q;
        T a2 = a * a, b2 = b * b, c2 = c * c;
        T lt2 = q < (T)2 ? (T)1 : (T)0, lt1 = q < (T)1 ? (T)1 : (T)0;
        return k.nf * (a2 * a2 * a - (T)6 * lt2 * (b2 * b2 * b) + (T)15 * lt1 * (c2 * c2 * c));
    }
    static __device__ __forceinline__ T dw_div_r(const KernelConst<T> &k, T r)
    {
        T q = r * k.h_inv;
        T result;
        if (k.order == 4) {
            T a = (T)2.5 - q, b = (T)1.5 - q, c = (T)0.5 - q;
            T lt15 = q < (T)1.5 ? (T)1 : (T)0, lt05 = q < (T)0.5 ? (T)1 : (T)0;
            result = (T)(-4) * (a * a * a) + (T)20 * lt15 * (b * b * b) - (T)40 * lt05 * (c * c * c);
        } else {
            T a = (T)3 - q, b = (T)2 - q, c = (T)1 - q;
            T a2 = a * a, b2 = b * b, c2 = c * c;
            T lt2 = q < (T)2 ? (T)1 : (T)0, lt1 = q < (T)1 ? (T)1 : (T)0;
            result = (T)(-5) * (a2 * a2) + (T)30 * lt2 * (b2 * b2) - (T)75 * lt1 * (c2 * c2);
        }
        return div_fast(k.nf * result * k.h_inv, r);
    }
};

// `kernel(k, r, h)`: strict r < compact_support (smoothing_kernels.jl:30-34)
template <int KERNEL, typename T>
__device__ __forceinline__ T kernel_safe(const KernelConst<T> &k, T r)
{
    return r < k.support ? SmoothingKernel<KERNEL, T>::w_unsafe(k, r) : (T)0;
}

// ---------------------------------------------------------------------------------------
// StateEquationCole (state_equations.jl:117-139).  The power is evaluated in double and
// rounded to T, which is how Julia evaluates Float32^Float32 and within 1 ulp of its
// Float64 pow.
template <typename T>
struct EosConst {
    T B;        // rho0 * c^2 / gamma
    T gamma, inv_gamma, rho0, p_bg;
    int clip;
};

// StateEquationAdaptiveCole: what follows the speed of sound, written on the device by k_adaptive_consts
// (tpb_vec.cuh) before the kick's kernels run -- the kick does not wait for the host
template <typename T>
struct AdaptConsts {
    T c;            // system_sound_speed(fluid)
    T B_f, B_w;     // rho0 c^2 / gamma of the fluid's / the boundary model's state equation
    T delta_h_c;    // density diffusion
    T nu_a, nu_b;   // kinematic viscosities seen by the wall's viscous term
    T rho_empty_w;  // wall density of an empty Adami sum
    T pad;
};

template <typename T>
__device__ __forceinline__ T eos_pressure(const EosConst<T> &e, T density)
{
    T x = density / e.rho0;  // IEEE division
    T xp = (T)pow((double)x, (double)e.gamma);
    T p = e.B * (xp - (T)1) + e.p_bg;
    return e.clip ? (p > (T)0 ? p : (T)0) : p;
}

template <typename T>
__device__ __forceinline__ T eos_inverse(const EosConst<T> &e, T pressure)
{
    T tmp = (pressure - e.p_bg) / e.B + (T)1;
    return e.rho0 * (T)pow((double)tmp, (double)e.inv_gamma);
}

// ---------------------------------------------------------------------------------------
// Per-pair physics of `interact!` (wcsph/rhs.jl:47-118).
template <typename T>
struct PairConst {
    KernelConst<T> kern;
    T c;                  // sound speed
    T alpha, beta, eps;   // ArtificialViscosityMonaghan; ViscosityMorris / ViscosityAdami: alpha = nu
    T eps_h2;             // epsilon * h^2
    T delta_h_c;          // delta * h_avg * c      (density_diffusion.jl:236-239)
    T almostzero;         // sqrt(eps(compact_support^2)) (rhs.jl:27-28)
    T radius2;            // search_radius^2
    int has_viscosity, has_diffusion;  // has_viscosity: TPB_VISCOSITY_* (0 none, 1 Monaghan, 2 Morris, 3 Adami)
};

// SAME: particle_system === neighbor_system (fluid-fluid); otherwise the neighbour is a
// static free-slip wall: v_b = 0, no viscous term, no density diffusion.
template <int ND, typename T, int KERNEL, int DENS, bool SAME>
__device__ __forceinline__ void interact_pair(const PairConst<T> &k, T m_b, T rho_a, T rho_b,
                                              T p_a, T p_b, const T (&v_a)[3],
                                              const T (&v_b)[3], const T (&pd)[3], T dist,
                                              T (&dv)[3], T &drho, T m_a = (T)0)
{
    const T wdr = SmoothingKernel<KERNEL, T>::dw_div_r(k.kern, dist);
    T grad[3];
#pragma unroll
    for (int d = 0; d < ND; ++d) grad[d] = wdr * pd[d];

    // pressure_acceleration.jl:8-15 / :34-41
    T f;
    if (DENS == 0)
        f = -m_b * div_fast(p_a + p_b, rho_a * rho_b);
    else
        f = -m_b * (div_fast(p_a, rho_a * rho_a) + div_fast(p_b, rho_b * rho_b));
#pragma unroll
    for (int d = 0; d < ND; ++d) dv[d] += f * grad[d];

    T vd[3];
#pragma unroll
    for (int d = 0; d < ND; ++d) vd[d] = SAME ? v_a[d] - v_b[d] : v_a[d];

    // ArtificialViscosityMonaghan (viscosity.jl:89-132)
    if (SAME && k.has_viscosity >= 2) {
        // ViscosityMorris (viscosity.jl:163-205) / ViscosityAdami (:222-279); nu_a = nu_b = nu
        const T nu = k.alpha;
        const T d2e = dist * dist + k.eps_h2;
        T coef;
        if (k.has_viscosity == 2) {
            T pg = pd[0] * grad[0] + pd[1] * grad[1];
            if (ND == 3) pg += pd[2] * grad[2];
            const T mu_a = nu * rho_a, mu_b = nu * rho_b;
            coef = div_fast(m_b * (mu_a + mu_b) * pg, rho_a * rho_b * d2e);
        } else {
            T gp = grad[0] * pd[0] + grad[1] * pd[1];
            if (ND == 3) gp += grad[2] * pd[2];
            const T eta_a = nu * rho_a, eta_b = nu * rho_b;
            const T volume_a = div_fast(m_a, rho_a), volume_b = div_fast(m_b, rho_b);
            const T tmp = div_fast((T)2 * eta_a * eta_b, (eta_a + eta_b) * d2e * m_a);
            coef = (volume_a * volume_a + volume_b * volume_b) * gp * tmp;
        }
#pragma unroll
        for (int d = 0; d < ND; ++d) dv[d] += coef * vd[d];
    }
    if (SAME && k.has_viscosity == 1) {
        T vr = vd[0] * pd[0] + vd[1] * pd[1];
        if (ND == 3) vr += vd[2] * pd[2];
        if (vr < (T)0) {
            T rho_mean = (rho_a + rho_b) / (T)2;
            T mu = div_fast(k.kern.h * vr, dist * dist + k.eps_h2);
            T dvv = div_fast(m_b * k.alpha * k.c * mu + m_b * k.beta * (mu * mu), rho_mean);
#pragma unroll
            for (int d = 0; d < ND; ++d) dv[d] += dvv * grad[d];
        }
    }

    // continuity_equation! (fluid.jl:160-186) + density_diffusion! (density_diffusion.jl:213-240)
    if (DENS == 0) {
        T vg = vd[0] * grad[0] + vd[1] * grad[1];
        if (ND == 3) vg += vd[2] * grad[2];
        drho += div_fast(rho_a, rho_b) * m_b * vg;
        if (SAME && k.has_diffusion) {
            T volume_b = div_fast(m_b, rho_b);
            T s = div_fast((T)2 * (rho_a - rho_b), dist * dist);
            T pg = (s * pd[0]) * grad[0] + (s * pd[1]) * grad[1];
            if (ND == 3) pg += (s * pd[2]) * grad[2];
            drho += k.delta_h_c * (volume_b * pg);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Float32 fast path of the same pair physics for the cell-tile sweep (tpb_tiles.cuh).
// Same formulas as `interact_pair`, regrouped so that one pair costs ~60 issue slots:
//   * the hardware approximations rcp.approx.ftz / rsqrt.approx.ftz (<= 2 ulp) stand in for
//     `div_fast` and sqrt -- the reference's own GPU path divides with LLVM fast division
//     (util.jl:3-5), and its Float32 accuracy bar is 1e-5;
//   * grad W = wdr * pos_diff is never formed: dv += [m_b (f + visc) wdr] * pos_diff,
//     v_ab . grad W = wdr (v_ab . pos_diff), and the Molteni-Colagrossi term collapses to
//     psi . grad W = 2 (rho_a - rho_b) wdr because pos_diff . pos_diff / r^2 = 1;
//   * the `vr < 0` viscosity switch is min(vr, 0); a rejected pair gets m_b = 0.
__device__ __forceinline__ float rcp_approx(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rsqrt_approx(float x)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

struct FastConst {
    float r2, az2;         // search radius^2, almostzero^2
    float h_inv, nfh;      // 1/h; cubic spline: nf * h_inv
    float c_t, c_w;        // Wendland C2: t = 1 + c_t r, wdr = c_w t^3
    float c_d;             // Wendland C4 / C6: nf * (-7/3 or -11/4) * h^-2
    int order;             // Wendland C4 / C6
    float h, eps_h2;       // viscosity
    float ac2, b2;         // 2 alpha c, 2 beta
    float nu;              // ViscosityMorris / ViscosityAdami
    int visc;              // TPB_VISCOSITY_*
    float dhc2;            // 2 delta h c (0 without density diffusion)
};

__host__ __device__ inline FastConst make_fast_const(const PairConst<float> &k)
{
    FastConst f;
    f.r2 = k.radius2;
    f.az2 = k.almostzero * k.almostzero;
    f.h_inv = k.kern.h_inv;
    f.nfh = k.kern.nf * k.kern.h_inv;
    f.c_t = -0.5f * k.kern.h_inv;
    f.c_w = k.kern.m5nf * k.kern.h_inv2;
    f.c_d = k.kern.nf * k.kern.c_d * k.kern.h_inv2;
    f.order = k.kern.order;
    f.h = k.kern.h;
    f.eps_h2 = k.eps_h2;
    f.ac2 = k.has_viscosity == 1 ? 2.0f * k.alpha * k.c : 0.0f;
    f.b2 = k.has_viscosity == 1 ? 2.0f * k.beta : 0.0f;
    f.nu = k.has_viscosity >= 2 ? k.alpha : 0.0f;
    f.visc = k.has_viscosity;
    f.dhc2 = k.has_diffusion ? 2.0f * k.delta_h_c : 0.0f;
    return f;
}

// (dW/dr)/r from d2 (already clamped to a finite non-zero value); rs = 1/r
template <int KERNEL>
__device__ __forceinline__ float fast_wdr(const FastConst &c, float dist, float rs)
{
    if (KERNEL == 0) {
        const float t = fmaf(dist, c.c_t, 1.0f);
        return c.c_w * (t * t * t);
    } else if (KERNEL == 2) {
        const float q = dist * c.h_inv;
        const float t = fmaf(dist, c.c_t, 1.0f);
        const float t2 = t * t;
        if (c.order == 4) return c.c_d * fmaf(5.0f, q, 2.0f) * (t2 * t2 * t);
        return c.c_d * fmaf(fmaf(8.0f, q, 7.0f), q, 2.0f) * (t2 * t2 * t2 * t);
    } else if (KERNEL == 3) {
        const float q = dist * c.h_inv;
        float result;
        if (c.order == 4) {
            const float a = 2.5f - q, b = 1.5f - q, d = 0.5f - q;
            result = -4.0f * (a * a * a) + (q < 1.5f ? 20.0f * (b * b * b) : 0.0f) -
                     (q < 0.5f ? 40.0f * (d * d * d) : 0.0f);
        } else {
            const float a = 3.0f - q, b = 2.0f - q, d = 1.0f - q;
            const float a2 = a * a, b2 = b * b, d2_ = d * d;
            result = -5.0f * (a2 * a2) + (q < 2.0f ? 30.0f * (b2 * b2) : 0.0f) -
                     (q < 1.0f ? 75.0f * (d2_ * d2_) : 0.0f);
        }
        return c.nfh * result * rs;
    } else {
        const float q = dist * c.h_inv;
        const float a = 2.0f - q, b = 1.0f - q;
        const float inner = q < 1.0f ? 3.0f * (b * b) : 0.0f;
        return c.nfh * fmaf(-0.75f * a, a, inner) * rs;
    }
}

// Exact reference predicate + fast physics.  pa_term: DENS == 1 ? p_a / rho_a^2 : unused.
// CT = double: the reference's default single-precision set-up (Float32 fields, Float64
// coordinates, docs/src/gpu.md "Single precision simulations"): the difference is formed in
// double and rounded to float exactly as `convert(T, x_i - y_j)` does.
template <int ND, int KERNEL, int DENS, bool SAME, typename CT>
__device__ __forceinline__ void interact_pair_fast(const FastConst &c, const V4<CT> &xi,
                                                   const V4<CT> &xj, float rho_a, float p_a,
                                                   float pa_term, const float (&v_a)[3], float vbx,
                                                   float vby, float vbz, float rho_b, float p_b,
                                                   float (&dv)[3], float &drho)
{
    // (the target's mass is xi.w)
    float pd[3];
    float d2 = pos_diff_d2<ND, float, CT>(xi, xj, pd);  // exactly rounded, as the reference
    const bool ok = d2 <= c.r2 && d2 >= c.az2;
    d2 = ok ? d2 : c.r2;
    const float mb = ok ? (float)xj.w : 0.0f;
    const float rs = rsqrt_approx(d2);
    const float dist = d2 * rs;
    const float wdr = fast_wdr<KERNEL>(c, dist, rs);
    const float mw = mb * wdr;
    const float rb = rcp_approx(rho_b);
    float f;
    if (DENS == 0)
        f = -(p_a + p_b) * (rb * rcp_approx(rho_a));
    else
        f = -(pa_term + p_b * (rb * rb));
    const float vdx = SAME ? v_a[0] - vbx : v_a[0], vdy = SAME ? v_a[1] - vby : v_a[1];
    const float vdz = ND == 3 ? (SAME ? v_a[2] - vbz : v_a[2]) : 0.0f;
    float vr = fmaf(vdy, pd[1], vdx * pd[0]);
    if (ND == 3) vr = fmaf(vdz, pd[2], vr);
    if (SAME && (c.ac2 != 0.0f || c.b2 != 0.0f)) {
        const float mu = (c.h * fminf(vr, 0.0f)) * rcp_approx(d2 + c.eps_h2);
        f = fmaf(fmaf(c.b2, mu, c.ac2) * mu, rcp_approx(rho_a + rho_b), f);
    }
    if (SAME && c.visc >= 2) {
        // Morris: m_b nu (rho_a + rho_b) (pos_diff . grad W) / (rho_a rho_b (r^2 + eps h^2)),
        // Adami:  (V_a^2 + V_b^2) (grad W . pos_diff) 2 nu rho_a rho_b / ((rho_a + rho_b)(r^2 + eps h^2) m_a);
        // pos_diff . grad W = wdr r^2; a rejected pair has m_b = 0 (Morris) or wdr masked (Adami)
        const float ra = rcp_approx(rho_a);
        const float d2e = d2 + c.eps_h2;
        float coef;
        if (c.visc == 2) {
            coef = (mw * d2) * (c.nu * (rho_a + rho_b)) * ((ra * rb) * rcp_approx(d2e));
        } else {
            const float m_a = (float)xi.w;
            const float va = m_a * ra, vb = (float)xj.w * rb;
            const float wm = ok ? wdr : 0.0f;
            coef = fmaf(va, va, vb * vb) * (wm * d2) *
                   ((2.0f * c.nu * rho_a * rho_b) * rcp_approx((rho_a + rho_b) * d2e * m_a));
        }
        dv[0] = fmaf(coef, vdx, dv[0]);
        dv[1] = fmaf(coef, vdy, dv[1]);
        if (ND == 3) dv[2] = fmaf(coef, vdz, dv[2]);
    }
    const float s = mw * f;
    dv[0] = fmaf(s, pd[0], dv[0]);
    dv[1] = fmaf(s, pd[1], dv[1]);
    if (ND == 3) dv[2] = fmaf(s, pd[2], dv[2]);
    if (DENS == 0) {
        float t1 = rho_a * vr;
        if (SAME) t1 = fmaf(c.dhc2, rho_a - rho_b, t1);
        drho = fmaf(mw * rb, t1, drho);
    }
}

// ---------------------------------------------------------------------------------------
// Uniform cell grid shared by every point set (FullGridCellList semantics: padded bounding
// box, particles may only live in interior cells, so the neighbourhood never leaves the grid).
// Cells are `cell` wide in y and z and `cell / sx` wide in x ("x-split"): the cell index is
// linear with x fastest, so the 2 sx + 1 cells {cx - sx .. cx + sx} of one (y, z) row are one
// contiguous run of sorted particles, and the finer x resolution lets a sweep clip that run to
// the chord of the search sphere in that row (tpb_tiles.cuh).
// Float32 fields with Float64 coordinates: the phase-1 filter of the tile sweeps works on a
// Float32 copy of the positions relative to a fixed reference point (the lower corner of the
// bounding box), F = fl32(x - ref).  |F - (x - ref)| <= 2^-24 L per coordinate (L = largest
// extent of the box), so a distance computed from two F's is within 2 sqrt(3) 2^-24 L of the
// true one; `pad` is twice that.  The exact predicate and the pair physics still use the
// Float64 difference (phase 2).
template <typename CT>
struct FilterRef {
    CT ref[3];   // reference point, coordinate order
    float pad;   // added to the search radius of the filter
};
template <typename CT>
__device__ __forceinline__ V4<float> filter_position(const FilterRef<CT> &r, const V4<CT> &x)
{
    V4<float> f;
    f.x = (float)(x.x - r.ref[0]);
    f.y = (float)(x.y - r.ref[1]);
    f.z = (float)(x.z - r.ref[2]);
    f.w = 0.0f;
    return f;
}

template <typename CT>
struct GridConst {
    CT origin[3];
    CT inv_cell;    // 1 / cell  (y, z)
    CT inv_cell_x;  // sx / cell (x)
    CT cell;        // cell size in y, z
    CT gap_tol;     // bound on the rounding of cell-boundary positions (window clipping)
    CT lo[3], hi[3];  // valid coordinate range (bounding box), for the bounds check
    int n[3];
    int ncells;
    int sx;         // x-split factor
    FilterRef<CT> fref;  // Float32 filter copy of Float64 positions (tile sweeps)
    int ax;         // coordinate axis of the rows (fastest cell index); every per-axis field above is
                    // stored with the entries 0 and ax swapped ("cell-axis order")
};

template <int ND, typename CT>
__device__ __forceinline__ bool cell_coords(const GridConst<CT> &g, CT x, CT y, CT z, int &cx,
                                            int &cy, int &cz)
{
    // rows run along coordinate axis g.ax (the longest side of the bounding box): cell axis 0
    if (g.ax == 1) {
        const CT t = x;
        x = y;
        y = t;
    } else if (ND == 3 && g.ax == 2) {
        const CT t = x;
        x = z;
        z = t;
    }
    // written so that NaN coordinates fail the test
    bool ok = (x >= g.lo[0]) && (x <= g.hi[0]) && (y >= g.lo[1]) && (y <= g.hi[1]);
    if (ND == 3) ok = ok && (z >= g.lo[2]) && (z <= g.hi[2]);
    if (!ok) {
        cx = g.sx;
        cy = 1;
        cz = ND == 3 ? 1 : 0;
        return false;
    }
    cx = (int)floor((x - g.origin[0]) * g.inv_cell_x);
    cy = (int)floor((y - g.origin[1]) * g.inv_cell);
    cz = ND == 3 ? (int)floor((z - g.origin[2]) * g.inv_cell) : 0;
    cx = min(max(cx, g.sx), g.n[0] - g.sx - 1);
    cy = min(max(cy, 1), g.n[1] - 2);
    if (ND == 3) cz = min(max(cz, 1), g.n[2] - 2);
    return true;
}

template <typename CT>
__device__ __forceinline__ int cell_linear(const GridConst<CT> &g, int cx, int cy, int cz)
{
    return cx + g.n[0] * (cy + g.n[1] * cz);
}

}  // namespace tpb
