// tc_filter.cu -- bring-up / micro-benchmark of the tensor-core distance filter (sm_100a).
//
// d^2(i, j) = |x_i - y_j|^2 for 128 targets x N candidates as ONE tcgen05.mma (kind::tf32, K = 8):
//   A_i = [ x~i, y~i, z~i, s_h, s_l, 1, 1, 0 ]          s = |x~_i|^2 = s_h + s_l
//   B_j = [ -2x~j, -2y~j, -2z~j, 1, 1, t_h, t_l, 0 ]    t = |x~_j|^2 = t_h + t_l
// with coordinates relative to a tile origin and rounded to tf32 (cvt.rna), so every product is
// exact in the fp32 accumulator and the only error of the filter is the rounding of the coordinates.
// Accumulators live in TMEM; every thread reads the row of its own target with tcgen05.ld and turns
// 32 columns into one 32-bit accept mask.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_filter tc_filter.cu && ./tc_filter
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float tf32_rna(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// K-major, no swizzle: ((8, n), 2) : ((16 B, SBO), LBO)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
    return d;
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

constexpr int M = 128;        // targets per tile
constexpr int NCH = 96;       // candidates per MMA chunk (3 blocks of 32 columns, one per split thread)
constexpr int TMEM_COLS = 256;

struct P3 { float x, y, z, w; };

// operand row r (of `rows`): first K half at base + (r / 8) * 128 + (r % 8) * 16, second at + lbo
__device__ __forceinline__ void put_row(unsigned char *base, uint32_t lbo, int r, float4 k03, float4 k47)
{
    unsigned char *p = base + (r >> 3) * 128 + (r & 7) * 16;
    *reinterpret_cast<float4 *>(p) = k03;
    *reinterpret_cast<float4 *>(p + lbo) = k47;
}

template <bool STORE_D>
__global__ void __launch_bounds__(384, 2)
k_tc_filter(const P3 *__restrict__ X, const P3 *__restrict__ Y, int n_cand, float ox, float oy, float oz, float r2,
            float *__restrict__ D_out, uint32_t *__restrict__ mask_out, int reps)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bar = (uint64_t *)smem;                 // [2]
    uint32_t *tmem_slot = (uint32_t *)(smem + 16);
    unsigned char *sA = smem + 128;                   // 128 rows x 32 B = 4096
    unsigned char *sB = sA + 4096;                    // 2 x (NCH x 32 B)
    constexpr uint32_t A_LBO = M / 8 * 128, B_LBO = NCH / 8 * 128, B_BYTES = NCH * 32;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 32) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // A operand: one target per row
    if (tid < M) {
        const P3 p = X[tid];
        const float x = tf32_rna(p.x - ox), y = tf32_rna(p.y - oy), z = tf32_rna(p.z - oz);
        const float s = fmaf(z, z, fmaf(y, y, x * x));
        const float sh = tf32_rna(s), sl = tf32_rna(s - sh);
        put_row(sA, A_LBO, tid, make_float4(x, y, z, sh), make_float4(sl, 1.f, 1.f, 0.f));
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = *tmem_slot;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NCH >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const int n_chunks = (n_cand + NCH - 1) / NCH;
    uint32_t phase[2] = {0, 0};
    uint32_t acc_mask = 0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int c = 0; c < n_chunks; ++c) {
            const int buf = c & 1;
            // convert NCH candidates into the B operand (threads 0..NCH-1)
            if (tid < NCH) {
                const int j = c * NCH + tid;
                P3 p = j < n_cand ? Y[j] : P3{1e18f, 1e18f, 1e18f, 0.f};
                const float x = tf32_rna(p.x - ox), y = tf32_rna(p.y - oy), z = tf32_rna(p.z - oz);
                const float t = fmaf(z, z, fmaf(y, y, x * x));
                const float th = tf32_rna(t), tl = tf32_rna(t - th);
                put_row(sB + buf * B_BYTES, B_LBO, tid, make_float4(-2.f * x, -2.f * y, -2.f * z, 1.f),
                        make_float4(1.f, th, tl, 0.f));
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> tensor core reads
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncthreads();   // (also: every warp is done reading the TMEM columns of this buffer, two chunks ago)
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;");
                const uint64_t da = make_desc(smem_u32(sA), A_LBO, 128);
                const uint64_t db = make_desc(smem_u32(sB + buf * B_BYTES), B_LBO, 128);
                asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                             "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem + buf * 128),
                             "l"(da), "l"(db), "r"(idesc), "r"(0));
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[buf])) : "memory");
            }
            mbar_wait(&bar[buf], phase[buf]);
            phase[buf] ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;");
            // epilogue: warp w reads lanes 32 (w % 4) .., column block w / 4 of this chunk
            const int kg = warp >> 2;
            const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + buf * 128 + kg * 32;
            uint32_t v[32];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                         "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                           "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                           "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            uint32_t m = 0;
#pragma unroll
            for (int q = 0; q < 32; ++q) m |= (__uint_as_float(v[q]) <= r2) ? (1u << q) : 0u;
            acc_mask ^= m;
            if (STORE_D) {
                const int row = (warp & 3) * 32 + lane;
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    const int j = c * NCH + kg * 32 + q;
                    if (j < n_cand) D_out[(size_t)row * n_cand + j] = __uint_as_float(v[q]);
                }
                mask_out[((size_t)c * 3 + kg) * M + row] = m;
            }
        }
    }
    if (!STORE_D) mask_out[(size_t)blockIdx.x * 384 + tid] = acc_mask;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
}

int main()
{
    const float R = 0.0378f;                     // search radius of the 1 M dam break
    const int n_cand = 1800;
    std::vector<P3> X(M), Y(n_cand);
    srand(1234);
    auto rnd = [] { return (float)rand() / RAND_MAX; };
    // tile-like geometry at an offset typical for the tank (absolute coordinates up to a few metres)
    const float bx = 1.7f, by = 0.6f, bz = 0.35f;
    for (auto &p : X) p = {bx + R + 4.7f * R * rnd(), by + R + R * rnd(), bz + R + R * rnd(), 0.f};
    for (auto &p : Y) p = {bx + 6.7f * R * rnd(), by + 3 * R * rnd(), bz + 3 * R * rnd(), 0.f};
    const float ox = bx + 3.35f * R, oy = by + 1.5f * R, oz = bz + 1.5f * R;
    P3 *dX, *dY;
    float *dD;
    uint32_t *dM;
    const int n_chunks = (n_cand + NCH - 1) / NCH;
    CK(cudaMalloc(&dX, sizeof(P3) * M));
    CK(cudaMalloc(&dY, sizeof(P3) * n_cand));
    CK(cudaMalloc(&dD, sizeof(float) * M * n_cand));
    CK(cudaMalloc(&dM, sizeof(uint32_t) * std::max(n_chunks * 3 * M, 148 * 4 * 384)));
    CK(cudaMemcpy(dX, X.data(), sizeof(P3) * M, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dY, Y.data(), sizeof(P3) * n_cand, cudaMemcpyHostToDevice));
    const size_t smem = 128 + 4096 + 2 * NCH * 32;
    CK(cudaFuncSetAttribute(k_tc_filter<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_tc_filter<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const float r2 = R * R;
    k_tc_filter<true><<<1, 384, smem>>>(dX, dY, n_cand, ox, oy, oz, r2, dD, dM, 1);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> D((size_t)M * n_cand);
    std::vector<uint32_t> Mk((size_t)n_chunks * 3 * M);
    CK(cudaMemcpy(D.data(), dD, sizeof(float) * D.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(Mk.data(), dM, sizeof(uint32_t) * Mk.size(), cudaMemcpyDeviceToHost));
    double max_err = 0, max_rel_r2 = 0;
    long inside = 0, mask_bits = 0, missed = 0;
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < n_cand; ++j) {
            const double dx = (double)X[i].x - Y[j].x, dy = (double)X[i].y - Y[j].y, dz = (double)X[i].z - Y[j].z;
            const double d2 = dx * dx + dy * dy + dz * dz;
            const double err = std::fabs(D[(size_t)i * n_cand + j] - d2);
            max_err = std::max(max_err, err);
            const int c = j / NCH, kg = (j % NCH) / 32, q = j % 32;
            const bool bit = (Mk[((size_t)c * 3 + kg) * M + i] >> q) & 1;
            mask_bits += bit;
            if (d2 <= r2) { ++inside; if (!bit && d2 <= r2 * 0.98) ++missed; }
        }
    max_rel_r2 = max_err / r2;
    printf("tc_filter: max |d2_mma - d2| = %.3e = %.3e R^2 (sqrt: %.4f R); pairs inside R %ld, mask bits %ld, missed well inside %ld\n",
           max_err, max_rel_r2, std::sqrt(max_rel_r2), inside, mask_bits, missed);
    // timing: every SM busy with two blocks, each doing `reps` tiles of n_cand candidates
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int reps = 40, grid = 148 * 2;
    k_tc_filter<false><<<grid, 384, smem>>>(dX, dY, n_cand, ox, oy, oz, r2, dD, dM, 2);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    k_tc_filter<false><<<grid, 384, smem>>>(dX, dY, n_cand, ox, oy, oz, r2, dD, dM, reps);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double tiles = (double)grid * reps;
    printf("tc_filter: %.3f ms for %.0f tiles of %d x %d pair tests -> %.3f us per tile per SM-slot; 11893 tiles would take %.3f ms\n",
           ms, tiles, M, n_cand, 1e3 * ms / reps, ms / tiles * 11893);
    return 0;
}
