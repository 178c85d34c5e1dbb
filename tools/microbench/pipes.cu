// Issue-rate microbenchmark for sm_100a: FFMA / FFMA2 / FADD / FADD2 / FMUL2 / MUFU / mixed,
// and shared-memory LDS.128 broadcast vs. random. Prints warp-instructions per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define N_ITER 4096
#define UNROLL 8
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, float s0, float s1, long long *cyc)
{
    float a[UNROLL], b[UNROLL];
    u64 p[UNROLL];
    for (int i = 0; i < UNROLL; ++i) { a[i] = s0 + i + threadIdx.x; b[i] = s1 * i; p[i] = ((u64)__float_as_uint(a[i]) << 32) | __float_as_uint(b[i]); }
    u64 ps = ((u64)__float_as_uint(s0) << 32) | __float_as_uint(s1);
    long long t0 = clock64();
    for (int it = 0; it < N_ITER; ++it) {
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) {
            if (MODE == 0) a[i] = fmaf(a[i], s0, s1);
            if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(ps));
            if (MODE == 2) a[i] = __fadd_rn(a[i], s1);
            if (MODE == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ps));
            if (MODE == 4) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(ps));
            if (MODE == 5) a[i] = __fmul_rn(a[i], s0);
            if (MODE == 6) { a[i] = fmaf(a[i], s0, s1); b[i] = __fadd_rn(b[i], s1); }            // FFMA + FADD
            if (MODE == 7) { a[i] = fmaf(a[i], s0, s1); p[i] = p[i] + (u64)it; }                    // FFMA + IADD64
            if (MODE == 8) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (MODE == 9) { a[i] = fmaf(a[i], s0, s1); int q = __float_as_int(b[i]); q = (q ^ it) + i; b[i] = __int_as_float(q); } // FFMA + LOP/IADD
        }
    }
    long long t1 = clock64();
    float acc = 0;
    for (int i = 0; i < UNROLL; ++i) acc += a[i] + b[i] + __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)p[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// shared-memory LDS.128: MODE 0 broadcast, 1 linear (conflict-free), 2 random 16-B records
template <int MODE>
__global__ void __launch_bounds__(256) ks(float *out, const int *idx, long long *cyc)
{
    __shared__ float4 tile[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) tile[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    int my[UNROLL];
    for (int i = 0; i < UNROLL; ++i) my[i] = MODE == 0 ? (i * 37) & 2047 : MODE == 1 ? (threadIdx.x + i * 256) & 2047 : idx[(threadIdx.x * UNROLL + i) & 4095] & 2047;
    float4 acc = make_float4(0, 0, 0, 0);
    long long t0 = clock64();
    for (int it = 0; it < N_ITER / 4; ++it) {
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) {
            float4 v = tile[(my[i] + it * (MODE == 2 ? 8 : 1)) & 2047];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int instr_per_slot, float *out, long long *cyc)
{
    int blocks = 148 * 2;  // 2 x 256 threads per SM = 16 warps / SM = 4 per SMSP
    k<MODE><<<blocks, 256>>>(out, 1.0001f, 0.5f, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, 1.0001f, 0.5f, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[1]; cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    double warp_instr_per_sm = 16.0 * N_ITER * UNROLL * instr_per_slot;
    printf("%-28s cycles %8lld  warp-instr/clk/SM %.3f  (per SMSP %.3f)  %.3f ms\n", name, h[0], warp_instr_per_sm / h[0], warp_instr_per_sm / h[0] / 4, ms);
}
template <int MODE>
void runs(const char *name, float *out, const int *idx, long long *cyc)
{
    int blocks = 148 * 2;
    ks<MODE><<<blocks, 256>>>(out, idx, cyc); cudaDeviceSynchronize();
    ks<MODE><<<blocks, 256>>>(out, idx, cyc); cudaDeviceSynchronize();
    long long h[1]; cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
    double lds_per_sm = 16.0 * (N_ITER / 4) * UNROLL;
    printf("%-28s cycles %8lld  cycles per warp-LDS.128 per SM %.2f\n", name, h[0], h[0] / lds_per_sm);
}
int main()
{
    float *out; long long *cyc; int *idx;
    cudaMalloc(&out, 148 * 2 * 256 * 4); cudaMalloc(&cyc, 148 * 2 * 8); cudaMalloc(&idx, 4096 * 4);
    int h[4096]; unsigned s = 12345; for (int i = 0; i < 4096; ++i) { s = s * 1664525u + 1013904223u; h[i] = (s >> 8) & 2047; }
    cudaMemcpy(idx, h, sizeof(h), cudaMemcpyHostToDevice);
    run<0>("FFMA", 1, out, cyc);
    run<1>("FFMA2", 1, out, cyc);
    run<2>("FADD", 1, out, cyc);
    run<3>("FADD2", 1, out, cyc);
    run<4>("FMUL2", 1, out, cyc);
    run<5>("FMUL", 1, out, cyc);
    run<6>("FFMA+FADD", 2, out, cyc);
    run<7>("FFMA+IADD64", 2, out, cyc);
    run<8>("MUFU.RSQ", 1, out, cyc);
    run<9>("FFMA+LOP+IADD", 3, out, cyc);
    runs<0>("LDS.128 broadcast", out, idx, cyc);
    runs<1>("LDS.128 linear", out, idx, cyc);
    runs<2>("LDS.128 random", out, idx, cyc);
    return 0;
}
