// tpb_halo.cuh -- ghost-particle exchange between the slabs of one box over peer memory.
//
// The reference has no distributed path; this is the B200 design.
// One process per GPU.  Every rank owns a receive area (cudaMalloc'ed by tpb_peer_alloc, mapped
// into the neighbours' address spaces with CUDA IPC): two parities of `stage_u [slots][ND]`,
// `stage_v [slots][NV]` and one 32-bit flag per neighbour.  Per kick and neighbour:
//
//   k_halo_pack     (sender)    gathers the rows (x, v, rho) of its candidate particles, marks the
//                               ones currently outside the ghost layer with x = NaN and stores them
//                               straight into the neighbour's receive area over NVLink -- the pack
//                               and the transfer are one kernel; the last block publishes the
//                               kick's epoch in the neighbour's flag (st.release.sys);
//   k_halo_install  (receiver)  spins on its own flags (ld.acquire.sys) until both neighbours
//                               have delivered this epoch, then copies the rows behind the owned
//                               particles of the extended ODE vectors.
//
// Alternating parities make the protocol safe without a second handshake: a sender can only reach
// epoch n + 2 (same parity as n) after it has seen the receiver's epoch n + 1, which the receiver
// publishes after it has consumed epoch n (stream order).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace tpb {

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// side < 0: the neighbour is on the left, a row is live while x < threshold (= lo + halo);
// side > 0: the neighbour is on the right, live while x >= threshold (= hi - halo).
template <typename T, typename CT>
__global__ void __launch_bounds__(256)
k_halo_pack(int nd, int nv, const CT *__restrict__ u, const T *__restrict__ v,
            const int64_t *__restrict__ cand, int64_t n_cand, CT threshold, int side,
            CT *__restrict__ peer_u, T *__restrict__ peer_v, unsigned long long *__restrict__ done,
            unsigned long long done_target, uint32_t *__restrict__ peer_flag, uint32_t epoch)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_cand; r += stride) {
        const int64_t i = cand[r];
        const CT x = u[i * nd];
        const bool live = side < 0 ? x < threshold : x >= threshold;
        peer_u[r * nd] = live ? x : (CT)CUDART_NAN;
        for (int d = 1; d < nd; ++d) peer_u[r * nd + d] = u[i * nd + d];
        for (int d = 0; d < nv; ++d) peer_v[r * nv + d] = v[i * nv + d];
    }
    // last block: every row of every block is visible system-wide before the flag is
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long arrived = atomicAdd(done, 1ull) + 1ull;
        if (arrived == done_target) {
            __threadfence_system();
            st_release_sys(peer_flag, epoch);
        }
    }
}

__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// false: the neighbour did not deliver within `timeout_ns` (a dead peer must not hang the GPU)
__device__ __forceinline__ bool wait_epoch(const uint32_t *flag, uint32_t epoch, unsigned long long timeout_ns)
{
    const unsigned long long t0 = global_timer_ns();
    // epochs only grow; the (int32) difference tolerates wrap-around
    while ((int32_t)(ld_acquire_sys(flag) - epoch) < 0) {
        __nanosleep(64);
        if (global_timer_ns() - t0 > timeout_ns) return false;
    }
    return true;
}

template <typename T, typename CT>
__global__ void __launch_bounds__(256)
k_halo_install(int64_t n_u, int64_t n_v, const CT *__restrict__ stage_u, const T *__restrict__ stage_v,
               CT *__restrict__ u_ghost, T *__restrict__ v_ghost, const uint32_t *__restrict__ flag_left,
               const uint32_t *__restrict__ flag_right, uint32_t epoch, unsigned long long timeout_ns,
               int *__restrict__ timed_out)
{
    if (threadIdx.x == 0) {
        bool ok = true;
        if (flag_left) ok = wait_epoch(flag_left, epoch, timeout_ns);
        if (ok && flag_right) ok = wait_epoch(flag_right, epoch, timeout_ns);
        if (!ok) atomicOr(timed_out, 2);
    }
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // the rows were written by a peer GPU: read them past L1
    for (int64_t k = first; k < n_u; k += stride) u_ghost[k] = __ldcg(stage_u + k);
    for (int64_t k = first; k < n_v; k += stride) v_ghost[k] = __ldcg(stage_v + k);
}

}  // namespace tpb
