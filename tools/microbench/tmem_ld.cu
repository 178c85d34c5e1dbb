// tmem_ld.cu -- TMEM read throughput on sm_100a: how many bytes per clock per SM can tcgen05.ld deliver
// when all warps of 1 / 2 resident blocks (12 warps each, 3 per TMEM lane quadrant) read 32x32b.x32 tiles?
// Decides whether a tensor-core distance filter (d^2 matrix in TMEM, one 32-column read per 32 pair
// tests and thread) can beat the scalar filter of k_interact_tiles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld tmem_ld.cu && ./tmem_ld
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

template <int NCOL, bool REDUCE>
__global__ void __launch_bounds__(384, 2) k_ld(uint32_t *out, int iters, long long *cyc)
{
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_slot)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_slot;
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 32;
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t v[32];
        if (NCOL == 32) {
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                         "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                           "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                           "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(taddr));
        } else {
#pragma unroll
            for (int h = 0; h < 2; ++h)
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                             : "=r"(v[16 * h + 0]), "=r"(v[16 * h + 1]), "=r"(v[16 * h + 2]), "=r"(v[16 * h + 3]), "=r"(v[16 * h + 4]),
                               "=r"(v[16 * h + 5]), "=r"(v[16 * h + 6]), "=r"(v[16 * h + 7]), "=r"(v[16 * h + 8]), "=r"(v[16 * h + 9]),
                               "=r"(v[16 * h + 10]), "=r"(v[16 * h + 11]), "=r"(v[16 * h + 12]), "=r"(v[16 * h + 13]),
                               "=r"(v[16 * h + 14]), "=r"(v[16 * h + 15])
                             : "r"(taddr + 16 * h));
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (REDUCE) {
            // the epilogue of the filter: gather the 32 sign bits into one mask word (one SHF per column)
            uint32_t m = 0;
#pragma unroll
            for (int q = 0; q < 32; ++q) m = __funnelshift_l(v[q], m, 1);
            acc ^= m;
        } else {
            acc ^= v[it & 31];
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256));
}

template <int NCOL, bool REDUCE>
int run(const char *name, int blocks_per_sm)
{
    uint32_t *out; long long *cyc;
    const int grid = 148 * blocks_per_sm, iters = 2000;
    CK(cudaMalloc(&out, sizeof(uint32_t) * grid * 384));
    CK(cudaMalloc(&cyc, sizeof(long long) * grid));
    k_ld<NCOL, REDUCE><<<grid, 384>>>(out, 10, cyc);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_ld<NCOL, REDUCE><<<grid, 384>>>(out, iters, cyc);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c0; CK(cudaMemcpy(&c0, cyc, sizeof(c0), cudaMemcpyDeviceToHost));
    const double bytes_per_block = (double)iters * 12 * 32 * 32 * 4;
    printf("%-28s %d block(s)/SM: %8.0f cycles (block 0), %.3f ms -> %.1f B/clk/SM, %.2f us per 128x1824 tile per SM\n", name, blocks_per_sm,
           (double)c0, ms, bytes_per_block * blocks_per_sm / c0, 128.0 * 1824 * 4 / (bytes_per_block * blocks_per_sm / c0) / 1965.0);
    cudaFree(out); cudaFree(cyc);
    return 0;
}
int main()
{
    run<32, false>("ld.x32 only", 1);
    run<32, false>("ld.x32 only", 2);
    run<16, false>("2 x ld.x16 only", 2);
    run<32, true>("ld.x32 + 32 SHF (mask)", 1);
    run<32, true>("ld.x32 + 32 SHF (mask)", 2);
    return 0;
}
