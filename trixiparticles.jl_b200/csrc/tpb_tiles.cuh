// tpb_tiles.cuh -- neighbour sweeps, variant 2 ("cell tiles"): the production path.
//
// B200 counterpart of PointNeighbors `foreach_point_neighbor` + the loop bodies of
//   interact!                 /root/reference/src/schemes/fluid/weakly_compressible_sph/rhs.jl:5-127
//   boundary_pressure_extrapolation! + compute_adami_density!
//                             /root/reference/src/schemes/boundary/wall_boundary/dummy_particles.jl:489-672
//
// Work decomposition.  Sorted particles of one cell row (fixed cy, cz; x fastest) are a
// contiguous range; it is cut into balanced "tiles" of <= TILE_TB target particles, one
// thread block each, one thread per target.  The 3^(ND-1) neighbour rows of a tile are 3 / 9
// contiguous runs of sorted records: one elected thread stages them into shared memory with
// 1-D TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx) -- UBLKCP in SASS.
//
// Per thread, two alternating phases remove the SIMT divergence of the naive loop
// (only 17 % of the 729 candidates of a 3-D particle are neighbours):
//   phase 1  scan the candidates of the own 3^ND cells (every lane of a warp reads the same
//            record: shared-memory broadcast), apply a cheap conservative distance filter and
//            append accepted tile indices to a private 16-bit list in shared memory;
//   phase 2  walk the list densely (all lanes busy): exact reference predicate
//            d^2 <= R^2 (bit-exact arithmetic, tpb_device.cuh) and the pair physics.
// Accumulators live in registers; each particle's dv is written once; no atomics.
//
// KS threads per target ("split", KS = 3 for Float32): thread k of a target scans every KS-th group
// of four candidates of each neighbour row, so a block of KS * 128 threads shares ONE staged tile.
// Shared memory limits an SM to two tiles; the split raises the resident warps from 8 to 24, which
// is what hides the shared-memory and ALU latencies of both phases.  The KS partial sums of a
// target are combined through shared memory in a fixed order (tile_reduce).
#pragma once
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "tpb_device.cuh"
#include "tpb_nhs.cuh"
#include "tpb_sweeps.cuh"
#include "tpb_sweeps.cuh"

// tuning knobs (compile time)
#ifndef TPB_P2_UNROLL
#define TPB_P2_UNROLL 4  // pairs in flight in phase 2
#endif
#ifndef TPB_SPLIT
#define TPB_SPLIT 3  // threads per target particle in the Float32 sweeps
#endif
#ifndef TPB_INTERLEAVE
#define TPB_INTERLEAVE 1  // 1: the KS threads of a target take every KS-th candidate; 0: every KS-th group of four
#endif
#ifndef TPB_LIST_MASKS
// 1: list entries are (first candidate, 8-bit accept mask) per group of eight candidates -- phase 1 shrinks from 91
// to 64 instructions per eight candidates (one funnel shift per verdict, one store per group), but phase 2 then has
// to find its pairs with CLZ in a per-lane loop: measured 0.880 ms (one pair at a time) / 0.908 ms (four decoded
// ahead) against 0.852 ms for 0: one 16-bit index per accepted candidate.  Kept as an experiment.
#define TPB_LIST_MASKS 0
#endif

namespace tpb {

#ifndef TPB_TILE_TB
#define TPB_TILE_TB 128
#endif
constexpr int TILE_TB = TPB_TILE_TB;  // target particles per tile; a block has KS * TILE_TB threads
constexpr int TILE_MAXSEG = 12;   // segments per staged chunk (>= 3^(ND-1))
constexpr int TILE_TABW = 16;     // cell_start entries per neighbour row kept in shared memory

// ------------------------------------------------------------------ PTX helpers (sm_100a)
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on `bar`.
// dst, src and bytes must be multiples of 16.
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TPB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TPB_DONE_%=;\n"
        "bra TPB_WAIT_%=;\n"
        "TPB_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------ tile table
// row_tiles[r] = number of tiles of cell row r (r = cy + n1 * cz)
static __global__ void __launch_bounds__(256)
k_row_tiles(const int *__restrict__ cell_start, int n0, int nrows, int *__restrict__ row_tiles)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    int cnt = cell_start[(int64_t)(r + 1) * n0] - cell_start[(int64_t)r * n0];
    row_tiles[r] = (cnt + TILE_TB - 1) / TILE_TB;
}

// tile descriptors: (first target, one past the last target, cell row, unused)
static __global__ void __launch_bounds__(256)
k_fill_tiles(const int *__restrict__ cell_start, int n0, int nrows,
             const int *__restrict__ row_tile_start, int4 *__restrict__ desc)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    const int first = row_tile_start[r], nt = row_tile_start[r + 1] - first;
    if (nt == 0) return;
    const int rs = cell_start[(int64_t)r * n0], re = cell_start[(int64_t)(r + 1) * n0];
    const long long cnt = re - rs;
    for (int k = 0; k < nt; ++k)
        desc[first + k] = make_int4(rs + (int)(cnt * k / nt), rs + (int)(cnt * (k + 1) / nt), r, 0);
}

// Both steps and the scan between them in ONE block, for grids of up to TILE_TABLE_MAX_ROWS cell
// rows (one launch instead of five on the per-kick path).
constexpr int TILE_TABLE_MAX_ROWS = 16384;
static __global__ void __launch_bounds__(SCAN_THREADS)
k_row_tile_table(const int *__restrict__ cell_start, int n0, int nrows, int *__restrict__ row_tile_start,
                 int4 *__restrict__ desc)
{
    int carry = 0;
    for (int base = 0; base < nrows; base += SCAN_THREADS) {
        const int r = base + threadIdx.x;
        int rs = 0, cnt = 0;
        if (r < nrows) {
            rs = cell_start[(int64_t)r * n0];
            cnt = cell_start[(int64_t)(r + 1) * n0] - rs;
        }
        const int nt = (cnt + TILE_TB - 1) / TILE_TB;
        int total;
        const int first = carry + block_inclusive_scan(nt, total) - nt;
        if (r < nrows) {
            row_tile_start[r] = first;
            for (int k = 0; k < nt; ++k)
                desc[first + k] = make_int4(rs + (int)((long long)cnt * k / nt),
                                            rs + (int)((long long)cnt * (k + 1) / nt), r, 0);
        }
        carry += total;
    }
    if (threadIdx.x == 0) row_tile_start[nrows] = carry;
}

// ------------------------------------------------------------------ one-pass cell scan + tile table
// cell_start = exclusive_scan(count) AND the tile table of the sorted point set in ONE launch (the
// per-kick path; replaces memset + three scan kernels + k_row_tile_table): every block owns a whole
// number of cell rows, scans its cells locally, and chains both running totals -- particles and
// tiles -- through a decoupled look-back (status words carry the number of the launch, so nothing has
// to be reset between launches; blocks take a ticket, so a block only ever waits for blocks that
// have already started).  The histogram is cleared on the way out for the next rebuild.
//   status[b]: bits 63..55 launch number, 54..53 flag (1: aggregate of block b, 2: inclusive prefix),
//              52..24 particles, 23..0 tiles
#ifndef TPB_CSCAN_THREADS
#define TPB_CSCAN_THREADS 1024  // measured: 1024 > 512 > 256 (fewer, larger blocks)
#endif
constexpr int CSCAN_THREADS = TPB_CSCAN_THREADS;
constexpr int CSCAN_TILE = CSCAN_THREADS * SCAN_ITEMS;
constexpr int CSCAN_MAX_ROWS = 2 * CSCAN_THREADS;  // rows per block (two per thread)
// The status word IS the message (no other data is handed over through it), so relaxed device-scope
// accesses are enough: `ld.acquire.gpu` in the polling loop costs a CCTL.IVALL (L1 invalidation) per
// poll -- 31 % of the kernel's stall samples in the first version -- and `st.release.gpu` a MEMBAR.
__device__ __forceinline__ void st_status_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_status_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
static __global__ void __launch_bounds__(CSCAN_THREADS)
k_scan_cells_tiles(int *__restrict__ count, int n0, int nrows, int rows_per_block,
                   unsigned long long *__restrict__ ticket, unsigned long long *__restrict__ status,
                   int *__restrict__ cell_start, int *__restrict__ row_tile_start, int4 *__restrict__ desc,
                   int *__restrict__ zero_word)
{
    __shared__ int s_bid, s_off[2];
    __shared__ unsigned s_epoch;
    __shared__ int row_pref[CSCAN_MAX_ROWS + 1];
    if (threadIdx.x == 0) {
        // launch number and block number from one monotonic 64-bit ticket: nothing depends on host-side
        // counters, so the launch can be replayed from a CUDA graph
        const unsigned long long t = atomicAdd(ticket, 1ull);
        const unsigned long long launch = t / gridDim.x;
        s_bid = (int)(t - launch * gridDim.x);
        s_epoch = 1u + (unsigned)(launch % 510ull);  // 9 bits; a stale word always carries the previous launch
    }
    __syncthreads();
    const int bid = s_bid;
    const unsigned epoch = s_epoch;
    const int r0 = bid * rows_per_block;
    const int rows_here = min(rows_per_block, nrows - r0);
    const int cells_here = rows_here * n0;
    int *const cnt_blk = count + (int64_t)r0 * n0;
    const int base = threadIdx.x * SCAN_ITEMS;
    static_assert(SCAN_ITEMS == 8, "two int4 per thread");
    // 16-byte accesses where the block's first cell allows it (rows_per_block is a multiple of four whenever
    // it can be): a warp then touches whole sectors -- with 4-byte accesses at a stride of 32 bytes every
    // store is a partial-sector write, and a 29 M-cell grid (the end slab of the 8-GPU run) took 0.39 ms
    const bool vec = ((((int64_t)r0 * n0) & 3) == 0) && base + SCAN_ITEMS <= cells_here;
    int v[SCAN_ITEMS];
    int sum = 0;
    if (vec) {
        const int4 a = reinterpret_cast<const int4 *>(cnt_blk + base)[0];
        const int4 b = reinterpret_cast<const int4 *>(cnt_blk + base)[1];
        v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) sum += v[k];
        reinterpret_cast<int4 *>(cnt_blk + base)[0] = make_int4(0, 0, 0, 0);
        reinterpret_cast<int4 *>(cnt_blk + base)[1] = make_int4(0, 0, 0, 0);
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            v[k] = base + k < cells_here ? cnt_blk[base + k] : 0;
            sum += v[k];
        }
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k)
            if (base + k < cells_here) cnt_blk[base + k] = 0;
    }
    int total_c;
    const int run0 = block_inclusive_scan(sum, total_c) - sum;  // local exclusive prefix of cell `base`
    {
        int r = (base + n0 - 1) / n0;  // first row starting at or after cell `base`
        int run = run0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            if (base + k < cells_here && base + k == r * n0) row_pref[r++] = run;
            run += v[k];
        }
        if (threadIdx.x == 0) row_pref[rows_here] = total_c;
    }
    __syncthreads();
    int nt[2], cnt[2], tsum = 0;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int r = threadIdx.x * 2 + e;
        cnt[e] = r < rows_here ? row_pref[r + 1] - row_pref[r] : 0;
        nt[e] = (cnt[e] + TILE_TB - 1) / TILE_TB;
        tsum += nt[e];
    }
    int total_t;
    const int trun0 = block_inclusive_scan(tsum, total_t) - tsum;
    {
        // Publish the block's aggregate, then look back with the WHOLE block: thread t reads the status of
        // predecessor bid - 1 - t, so the window covers every block that can be resident at once and the
        // look-back is one round trip (with one warp looking back 32 at a time a block spent most of its
        // life here: 0.39 ms for the 29 M cells of the end slab of the 8-GPU run).  Both totals travel in
        // one word: [63:55] launch number, [54:53] flag, [52:24] particles, [23:0] tiles.
        __shared__ unsigned long long w_sum[32];
        __shared__ int w_incl[32];
        const unsigned long long FIELD = (1ull << 53) - 1;
        const unsigned long long ep = (unsigned long long)epoch << 55;
        const unsigned long long mine = ((unsigned long long)(unsigned)total_c << 24) | (unsigned)total_t;
        if (threadIdx.x == 0 && bid > 0) st_status_u64(&status[bid], ep | (1ull << 53) | mine);
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        unsigned long long off = 0;
        bool found = false;
        for (int p0 = bid - 1; !found; p0 -= CSCAN_THREADS) {
            const int p = p0 - (int)threadIdx.x;
            unsigned long long sw = 2ull << 53;  // before the first block: an empty inclusive prefix
            if (p >= 0) {
                do {
                    sw = ld_status_u64(&status[p]);
                } while ((unsigned)(sw >> 55) != epoch || ((sw >> 53) & 3ull) == 0);
            }
            const unsigned incl = __ballot_sync(0xffffffffu, ((sw >> 53) & 3ull) == 2);
            const int stop = incl ? __ffs(incl) - 1 : 31;
            unsigned long long val = lane <= stop ? (sw & FIELD) : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
            if (lane == 0) {
                w_sum[wid] = val;
                w_incl[wid] = incl != 0;
            }
            __syncthreads();
            // every thread walks the 32 warp results (nearest predecessors first) -- uniform, no broadcast
            for (int w = 0; w < CSCAN_THREADS / 32 && !found; ++w) {
                off += w_sum[w];
                found = w_incl[w] != 0;
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            st_status_u64(&status[bid], ep | (2ull << 53) | (off + mine));
            s_off[0] = (int)(off >> 24);
            s_off[1] = (int)(off & 0xFFFFFFull);
            if (bid == 0 && zero_word) *zero_word = 0;
        }
    }
    __syncthreads();
    const int off_c = s_off[0], off_t = s_off[1];
    {
        int *const cs_blk = cell_start + (int64_t)r0 * n0;
        int run = off_c + run0;
        if (vec) {
            int e[SCAN_ITEMS];
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k) {
                e[k] = run;
                run += v[k];
            }
            reinterpret_cast<int4 *>(cs_blk + base)[0] = make_int4(e[0], e[1], e[2], e[3]);
            reinterpret_cast<int4 *>(cs_blk + base)[1] = make_int4(e[4], e[5], e[6], e[7]);
        } else {
#pragma unroll
            for (int k = 0; k < SCAN_ITEMS; ++k) {
                if (base + k < cells_here) cs_blk[base + k] = run;
                run += v[k];
            }
        }
    }
    const bool last_block = r0 + rows_here == nrows;
    if (last_block && threadIdx.x == 0) {
        cell_start[(int64_t)nrows * n0] = off_c + total_c;
        row_tile_start[nrows] = off_t + total_t;
    }
    int first = off_t + trun0;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int r = threadIdx.x * 2 + e;
        if (r < rows_here) {
            row_tile_start[r0 + r] = first;
            const int rs = off_c + row_pref[r];
            for (int k = 0; k < nt[e]; ++k)
                desc[first + k] = make_int4(rs + (int)((long long)cnt[e] * k / nt[e]),
                                            rs + (int)((long long)cnt[e] * (k + 1) / nt[e]), r0 + r, 0);
        }
        first += nt[e];
    }
}

// The same two steps for a static point set (the wall, built once): a row is first cut into
// segments wherever `gap` or more consecutive cells are empty, then every segment is split into
// balanced tiles.  Without the cut, the one tile of a tank-wall row that holds both its left and
// its right wall particles would span the whole tank and stage every fluid particle in between.
template <bool FILL>
__global__ void __launch_bounds__(128)
k_row_tiles_gaps(const int *__restrict__ cell_start, int n0, int nrows, int gap,
                 int *__restrict__ row_tiles, const int *__restrict__ row_tile_start, int4 *__restrict__ desc)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    const int *cs = cell_start + (int64_t)r * n0;
    int nt = 0, seg_first = 0, seg_cnt = 0, empty_run = 0;
    int out = FILL ? row_tile_start[r] : 0;
    auto close_segment = [&]() {
        const int k_tiles = (seg_cnt + TILE_TB - 1) / TILE_TB;
        if (FILL)
            for (int k = 0; k < k_tiles; ++k)
                desc[out++] = make_int4(seg_first + (int)((long long)seg_cnt * k / k_tiles),
                                        seg_first + (int)((long long)seg_cnt * (k + 1) / k_tiles), r, 0);
        nt += k_tiles;
        seg_cnt = 0;
    };
    int prev = cs[0];
    for (int c = 0; c < n0; ++c) {
        const int next = cs[c + 1];
        const int cnt = next - prev;
        if (cnt == 0) {
            if (++empty_run == gap && seg_cnt > 0) close_segment();
        } else {
            if (seg_cnt == 0) seg_first = prev;
            seg_cnt += cnt;
            empty_run = 0;
        }
        prev = next;
    }
    if (seg_cnt > 0) close_segment();
    if (!FILL) row_tiles[r] = nt;
}

// Candidate ranges of every tile, one warp per tile: for each of the 3^(ND-1) neighbour rows the
// run of sorted records [g0, g1) covering the cells {cxmin - sx .. cxmax + sx} of that row, in up
// to two neighbour sets (lanes 0..8: set 0, lanes 9..17: set 1).  Computed once per rebuild so
// that the sweeps start their TMA copies without a dependent chain of global loads.
//   rng[tile * stride + 9 * set + q] = (g0, g1);   ext[tile] = (cxmin, cxmax, total set 0, total set 1)
template <int ND>
__device__ __forceinline__ void
tile_ranges_body(int vblock, int n0, int n1, int sx, const int *__restrict__ n_tiles, const int4 *__restrict__ desc,
                 const int *__restrict__ xcell_start, const int *__restrict__ nb0_cell_start,
                 const int *__restrict__ nb1_cell_start, int2 *__restrict__ rng, int stride,
                 int4 *__restrict__ ext)
{
    constexpr int NROWS = ND == 3 ? 9 : 3;
    const int tile = (vblock * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (tile >= *n_tiles) return;
    const int4 d = desc[tile];
    const int row0 = d.z * n0;
    // cell of the first / last target: last cell of the row whose start is <= the sorted index
    int cx[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int target = e == 0 ? d.x : d.y - 1;
        int lo = row0, hi = row0 + n0;  // xcell_start[lo] <= target < xcell_start[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (xcell_start[mid] <= target) lo = mid; else hi = mid;
        }
        cx[e] = lo - row0;
    }
    const int set = lane / 9, q = lane - 9 * set;
    const int *nb = set == 0 ? nb0_cell_start : nb1_cell_start;
    int g0 = 0, g1 = 0;
    if (lane < 18 && q < NROWS && nb != nullptr) {
        const int dy = q % 3 - 1, dz = ND == 3 ? q / 3 - 1 : 0;
        const int cy = d.z % n1, cz = d.z / n1;
        const int c_lo = (cx[0] - sx) + n0 * ((cy + dy) + n1 * (cz + dz));
        g0 = nb[c_lo];
        g1 = nb[c_lo + (cx[1] - cx[0]) + 2 * sx + 1];
        rng[(int64_t)tile * stride + lane] = make_int2(g0, g1);
    }
    int cnt0 = set == 0 ? g1 - g0 : 0, cnt1 = set == 1 ? g1 - g0 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cnt0 += __shfl_xor_sync(0xffffffffu, cnt0, o);
        cnt1 += __shfl_xor_sync(0xffffffffu, cnt1, o);
    }
    if (lane == 0) ext[tile] = make_int4(cx[0], cx[1], cnt0, cnt1);
}
template <int ND>
__global__ void __launch_bounds__(256)
k_tile_ranges(int n0, int n1, int sx, const int *__restrict__ n_tiles, const int4 *__restrict__ desc,
              const int *__restrict__ xcell_start, const int *__restrict__ nb0_cell_start,
              const int *__restrict__ nb1_cell_start, int2 *__restrict__ rng, int stride,
              int4 *__restrict__ ext)
{
    tile_ranges_body<ND>(blockIdx.x, n0, n1, sx, n_tiles, desc, xcell_start, nb0_cell_start, nb1_cell_start, rng,
                         stride, ext);
}

struct TileHdr {
    int p0, p1, cy, cz, cxmin, cxmax;
    int tot0, tot1;    // candidates in neighbour set 0 / 1 (k_tile_ranges)
    int g0[9], g1[9];  // candidate range of every neighbour row (current neighbour set)
    int q, gpos;       // staging cursor: next row, next record in it
    int nseg, last;    // segments of the staged chunk; last: nothing left to stage after it
    int seg_row[TILE_MAXSEG], seg_begin[TILE_MAXSEG], seg_end[TILE_MAXSEG], seg_base[TILE_MAXSEG];
    // cell_start of the current neighbour set over the cells {cxmin - sx .. cxmax + sx + 1} of
    // every neighbour row (tab_w entries per row; 0: tile too long, read global memory instead)
    int tab_w;
    int tab[9 * TILE_TABW];
};
constexpr int TILE_HDR_BYTES = 16 + ((sizeof(TileHdr) + 15) / 16) * 16;  // mbarrier + header
constexpr int TILE_RED_VALS = 4;  // values per thread tile_reduce can combine (needs 2x list entries)

// shared-memory carve-up of one block
template <typename T, typename CT>
struct TileSmem {
    uint64_t *bar;
    TileHdr *hdr;
    unsigned short *list;  // [list_len][nt], nt = threads per block
    V4<CT> *tA;            // [cap]
    unsigned char *tB;     // [cap] of V4<T> (fluid neighbours) or V2<T> (wall neighbours)
    T *tP;                 // [cap]
    V4<float> *tF;         // [cap] Float32 filter positions (tile_has_filter_copy)
    int cap, list_len;
    __device__ TileSmem(unsigned char *base, int cap_, int list_len_, int nt) : cap(cap_), list_len(list_len_)
    {
        bar = (uint64_t *)base;
        hdr = (TileHdr *)(base + 16);
        list = (unsigned short *)(base + TILE_HDR_BYTES);
        unsigned char *p = (unsigned char *)list + (size_t)list_len * nt * sizeof(unsigned short);
        tA = (V4<CT> *)p;
        p += (size_t)cap * sizeof(V4<CT>);
        tB = p;
        p += (size_t)cap * sizeof(V4<T>);
        tP = (T *)p;
        p += (size_t)cap * sizeof(T);
        tF = (V4<float> *)p;  // cap is a multiple of 4: 16-byte aligned
    }
};
// Float32 fields with Float64 coordinates: a Float32 copy of the positions for the phase-1 filter
template <typename T, typename CT>
__host__ __device__ constexpr bool tile_has_filter_copy()
{
    return std::is_same<T, float>::value && !std::is_same<CT, float>::value;
}
// bytes per staged record
template <typename T, typename CT>
__host__ __device__ constexpr size_t tile_record_bytes()
{
    return sizeof(V4<CT>) + sizeof(V4<T>) + sizeof(T) + (tile_has_filter_copy<T, CT>() ? sizeof(V4<float>) : 0);
}
template <typename T, typename CT>
inline size_t tile_smem_bytes(int cap, int list_len, int ks = 1)
{
    // + 256: the last partial group of a candidate window reads up to 3 * KS records past it
    return TILE_HDR_BYTES + (size_t)list_len * ks * TILE_TB * sizeof(unsigned short) +
           (size_t)cap * tile_record_bytes<T, CT>() + 256;
}

// Load the tile descriptor of this block into hdr (thread 0).
__device__ __forceinline__ void tile_locate(TileHdr *hdr, int n1, const int4 &d, const int4 &e)
{
    hdr->p0 = d.x;
    hdr->p1 = d.y;
    hdr->cy = d.z % n1;
    hdr->cz = d.z / n1;
    hdr->cxmin = e.x;
    hdr->cxmax = e.y;
    hdr->tot0 = e.z;
    hdr->tot1 = e.w;
}

// Conservative phase-1 filter.  Float coordinates: fused arithmetic with a small margin on the
// radius (never rejects a pair the exact predicate accepts); otherwise the exact predicate.
template <int ND, typename T, typename CT>
struct Filter {
    T r2;
    __device__ __forceinline__ Filter(T radius2) : r2(radius2) {}
    __device__ __forceinline__ bool operator()(const V4<CT> &xi, const V4<CT> &xj) const
    {
        T pd[3];
        return pos_diff_d2<ND, T, CT>(xi, xj, pd) <= r2;
    }
};
// Blackwell packed FP32 (FADD2 / FMUL2): x and y of one record are an aligned register pair
// straight out of LDS.128, so (dx, dy) and (dx^2, dy^2) cost one issue slot each.
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
template <int ND>
struct Filter<ND, float, float> {
    float r2;
    __device__ __forceinline__ Filter(float radius2) : r2(radius2 * (1.0f + 8.0f * 1.1920929e-07f)) {}
    __device__ __forceinline__ bool operator()(const V4<float> &xi, const V4<float> &xj) const
    {
        unsigned long long dxy, sq;
        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dxy) : "l"(pack_f32x2(xi.x, xi.y)), "l"(pack_f32x2(xj.x, xj.y)));
        asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(sq) : "l"(dxy));
        float sx, sy;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(sx), "=f"(sy) : "l"(sq));
        float d2 = sx + sy;
        if (ND == 3) {
            float dz = xi.z - xj.z;
            d2 = fmaf(dz, dz, d2);
        }
        return d2 <= r2;
    }
    // d^2 - r2 with -r2 folded into the packed square (FFMA2): the SIGN BIT is the filter's verdict, so that a
    // funnel shift can collect it (the radius carries the margin: a true neighbour is strictly inside)
    __device__ __forceinline__ float margin(const V4<float> &xi, const V4<float> &xj) const
    {
        unsigned long long dxy, sq;
        asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(dxy) : "l"(pack_f32x2(xi.x, xi.y)), "l"(pack_f32x2(xj.x, xj.y)));
        asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(sq) : "l"(dxy), "l"(pack_f32x2(-r2, 0.0f)));
        float sx, sy;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(sx), "=f"(sy) : "l"(sq));
        float s = sx + sy;
        if (ND == 3) {
            float dz = xi.z - xj.z;
            s = fmaf(dz, dz, s);
        }
        return s;
    }
};
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory");
}

// Appending to the private list.  The append position is never updated in place: a predicated
// `@p add lpa, lpa, 256` right after `@p st [lpa]` has to wait until the store has read its
// address register (a write-after-read stall of ~20 cycles per candidate, serialised over the
// whole unrolled loop); `add` into a fresh register + `selp` keeps the chain at ALU latency.
template <int ESTEP>
__device__ __forceinline__ uint32_t list_append(uint32_t lpa, uint32_t idx, bool pass)
{
    uint32_t next;
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 ".reg .b32 t;\n"
                 ".reg .b16 h;\n"
                 "setp.ne.u32 p, %2, 0;\n"
                 "cvt.u16.u32 h, %3;\n"
                 "@p st.shared.u16 [%1], h;\n"
                 "add.u32 t, %1, %4;\n"
                 "selp.u32 %0, t, %1, p;\n"
                 "}"
                 : "=r"(next)
                 : "r"(lpa), "r"((uint32_t)pass), "r"(idx), "n"(ESTEP)
                 : "memory");
    return next;
}
// The same for a 32-bit entry (TPB_LIST_MASKS): (first candidate of a group << 8) | accept mask of the group.
template <int ESTEP>
__device__ __forceinline__ uint32_t entry_append(uint32_t lpa, uint32_t entry, bool nonempty)
{
    uint32_t next;
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 ".reg .b32 t;\n"
                 "setp.ne.u32 p, %2, 0;\n"
                 "@p st.shared.u32 [%1], %3;\n"
                 "add.u32 t, %1, %4;\n"
                 "selp.u32 %0, t, %1, p;\n"
                 "}"
                 : "=r"(next)
                 : "r"(lpa), "r"((uint32_t)nonempty), "r"(entry), "n"(ESTEP)
                 : "memory");
    return next;
}
// One neighbour set as the sweep sees it.  R1 = second record type (V4<T> fluid: v, rho;
// V2<T> wall: p, rho); HAS_P: a third scalar array (fluid pressure).
template <typename T, typename CT, typename R1_, bool HAS_P_>
struct NbSet {
    using R1 = R1_;
    static constexpr bool HAS_P = HAS_P_;
    const int *__restrict__ cell_start;
    const V4<CT> *__restrict__ A;
    const R1 *__restrict__ B;
    const T *__restrict__ P;
    const V4<float> *__restrict__ F;  // filter copy of A (tile_has_filter_copy), else unused
};

// Stage the neighbour rows of a sweep (warp 0 only, all 32 lanes; the staging area must be
// free).  `rng[q]` = candidate range of neighbour row q (k_tile_ranges).  Usual case: all rows
// fit into the staging area, then lane q issues the TMA copies of row q itself -- no serial
// loop; otherwise the chunked path of tile_sweep_staged takes over.
template <int ND, typename T, typename CT, typename NB>
__device__ __forceinline__ void tile_stage(TileSmem<T, CT> &sm, const NB &nb, const int2 *__restrict__ rng,
                                           int cxmin, int cxmax, int cy, int cz, int sx, int n0, int n1)
{
    using R1 = typename NB::R1;
    constexpr int NROWS = ND == 3 ? 9 : 3;
    constexpr bool HAS_F = tile_has_filter_copy<T, CT>();
    constexpr uint32_t REC_BYTES = sizeof(V4<CT>) + sizeof(R1) + (NB::HAS_P ? sizeof(T) : 0) +
                                   (HAS_F ? sizeof(V4<float>) : 0);
    const int tid = threadIdx.x;
    TileHdr *hdr = sm.hdr;
    R1 *tB = (R1 *)sm.tB;
    // all global loads first (row ranges and the per-lane window table are independent), so
    // that the TMA copies start after ONE memory latency; the table is written out afterwards
    int g0 = 0, g1 = 0;
    if (tid < NROWS) {
        const int2 r = rng[tid];
        g0 = r.x;
        g1 = r.y;
    }
    // per-lane candidate windows come from this table instead of global memory
    const int tab_w = cxmax - cxmin + 2 * sx + 2;
    const bool tab_fits = tab_w <= TILE_TABW;
    constexpr int NTV = (NROWS * TILE_TABW + 31) / 32;
    int tv[NTV];
#pragma unroll
    for (int i = 0; i < NTV; ++i) {
        const int k = tid + 32 * i;
        tv[i] = 0;
        if (tab_fits && k < NROWS * tab_w) {
            const int q = k / tab_w, c = k - q * tab_w;
            const int dy = q % 3 - 1, dz = ND == 3 ? q / 3 - 1 : 0;
            tv[i] = nb.cell_start[(cxmin - sx + c) + n0 * ((cy + dy) + n1 * (cz + dz))];
        }
    }
    if (tid < NROWS) {
        hdr->g0[tid] = g0;
        hdr->g1[tid] = g1;
    }
    const int a = g0 & ~3;  // 4 records: every array stays 16-byte aligned
    const int len = g1 > g0 ? ((g1 + 3) & ~3) - a : 0;
    int incl = len;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, o);
        if (tid >= o) incl += up;
    }
    const int total = __shfl_sync(0xffffffffu, incl, NROWS - 1);
    const unsigned nonempty = __ballot_sync(0xffffffffu, len > 0);
    if (total > 0 && total <= sm.cap) {
        if (tid == 0) {
            hdr->nseg = __popc(nonempty);
            hdr->last = 1;
            hdr->q = NROWS;
            mbar_expect_tx(sm.bar, (uint32_t)total * REC_BYTES);
        }
        __syncwarp();
        if (len > 0) {
            const int used = incl - len;
            const int k = __popc(nonempty & ((1u << tid) - 1u));
            hdr->seg_row[k] = tid;
            hdr->seg_begin[k] = g0;
            hdr->seg_end[k] = g1;
            hdr->seg_base[k] = used - a;  // tile index of record j = seg_base + j
            fence_proxy_async();
            bulk_g2s(sm.tA + used, nb.A + a, (uint32_t)(len * sizeof(V4<CT>)), sm.bar);
            bulk_g2s(tB + used, nb.B + a, (uint32_t)(len * sizeof(R1)), sm.bar);
            if constexpr (NB::HAS_P) bulk_g2s(sm.tP + used, nb.P + a, (uint32_t)(len * sizeof(T)), sm.bar);
            if constexpr (HAS_F) bulk_g2s(sm.tF + used, nb.F + a, (uint32_t)(len * sizeof(V4<float>)), sm.bar);
        }
    } else if (tid == 0) {
        hdr->nseg = 0;
        hdr->last = total == 0;  // no candidates at all (e.g. fluid far from any wall)
        hdr->q = 0;
        hdr->gpos = g0;
    }
    if (tid == 0) hdr->tab_w = tab_fits ? tab_w : 0;
#pragma unroll
    for (int i = 0; i < NTV; ++i) {
        const int k = tid + 32 * i;
        if (tab_fits && k < NROWS * tab_w) hdr->tab[k] = tv[i];
    }
}

// Sweep all neighbours (of one set) of the tile's targets after tile_stage and a block barrier.
// `body(xj, bj, pj)` is called for every candidate that passed the filter; it applies the exact
// predicate itself.  Must be called by all KS * TILE_TB threads of the block; thread
// `threadIdx.x` works for target `threadIdx.x % TILE_TB` and takes the groups of four candidates
// number kg, kg + KS, ... (kg = threadIdx.x / TILE_TB) of every neighbour row.
template <int KS, int ND, typename T, typename CT, typename NB, typename BODY>
__device__ __forceinline__ void tile_sweep_staged(TileSmem<T, CT> &sm, const GridConst<CT> &g, const NB &nb,
                                                  bool valid, int cx, const V4<CT> &xi, T radius2,
                                                  uint32_t &parity, BODY &&body)
{
    using R1 = typename NB::R1;
    constexpr int NROWS = ND == 3 ? 9 : 3;
    constexpr int NT = KS * TILE_TB;  // threads per block
    constexpr bool HAS_F = tile_has_filter_copy<T, CT>();
    constexpr uint32_t REC_BYTES = sizeof(V4<CT>) + sizeof(R1) + (NB::HAS_P ? sizeof(T) : 0) +
                                   (HAS_F ? sizeof(V4<float>) : 0);
    const int tid = threadIdx.x;
    const int kg = KS > 1 ? tid / TILE_TB : 0;  // warp-uniform
    TileHdr *hdr = sm.hdr;
    R1 *tB = (R1 *)sm.tB;
    const Filter<ND, T, CT> filter(radius2);
    // phase-1 records: the Float32 filter copy (padded radius) or the positions themselves
    using FRec = typename std::conditional<HAS_F, V4<float>, V4<CT>>::type;
    const FRec *const tFilt = HAS_F ? (const FRec *)sm.tF : (const FRec *)sm.tA;
    const float rf = HAS_F ? sqrtf((float)radius2) + g.fref.pad : 0.0f;
    const Filter<ND, float, float> ffilter(rf * rf);
    V4<float> xf = {};
    if constexpr (HAS_F) xf = filter_position<CT>(g.fref, xi);
    auto pass = [&](const FRec &c) {
        if constexpr (HAS_F)
            return ffilter(xf, c);
        else
            return filter(xi, c);
    };
    bool staged = hdr->nseg > 0;
    if (!staged && hdr->last) return;
    // chunked mode: thread 0 is about to rewrite the header every thread has just read
    if (!staged) __syncthreads();

    auto visit = [&](int idx) {
        if constexpr (NB::HAS_P)
            body(sm.tA[idx], tB[idx], sm.tP[idx]);
        else
            body(sm.tA[idx], tB[idx], (T)0);
    };
    constexpr int CS = TPB_INTERLEAVE ? KS : 1;   // distance between the candidates of one group
    constexpr int STEP = 4 * KS;                  // distance between two groups of one thread
#if TPB_LIST_MASKS
    // private list of 32-bit entries, entry e of thread t at list32[e * NT + t] (conflict-free): a group of
    // eight candidates (t + u CS, u = 0..3, and t + STEP + u CS) and the filter's verdicts as a bit mask,
    // candidate u at bit 7 - u.  Against one 16-bit index per accepted candidate this costs one funnel shift
    // per candidate instead of a compare, a predicated store and two address updates, and one store per
    // group: 91 -> 64 instructions per eight candidates in phase 1; phase 2 finds its pairs with CLZ.
    uint32_t *const my_list = reinterpret_cast<uint32_t *>(sm.list) + tid;
    const uint32_t list_a = smem_u32(my_list);
    const uint32_t list_end_a = list_a + (uint32_t)(sm.list_len / 2) * NT * 4u;
    constexpr uint32_t ESTEP = NT * 4u;  // bytes between consecutive entries of one thread
    uint32_t lpa = list_a;               // append position (shared-memory address)
    auto room = [&](int n) { return lpa + n * ESTEP <= list_end_a; };
    auto push_verdict = [&](uint32_t m, const FRec &c) -> uint32_t {
        if constexpr (HAS_F)
            return __funnelshift_l(__float_as_uint(ffilter.margin(xf, c)), m, 1);
        else if constexpr (std::is_same<T, float>::value && std::is_same<CT, float>::value)
            return __funnelshift_l(__float_as_uint(filter.margin(xi, c)), m, 1);
        else
            return (m << 1) | (uint32_t)pass(c);
    };
    auto flush = [&]() {
        const uint32_t *e = my_list;
        const uint32_t *const lp = my_list + (size_t)((lpa - list_a) / ESTEP) * NT;
        uint32_t cur = 0;
        constexpr int PF = TPB_P2_UNROLL;  // pairs in flight: decode PF candidates first, then run their bodies
        for (;;) {
            int idx[PF];
#pragma unroll
            for (int q = 0; q < PF; ++q) {
                if ((cur & 0xFFu) == 0 && e < lp) {  // (stored entries never have an empty mask)
                    cur = *e;
                    e += NT;
                }
                const uint32_t m = cur & 0xFFu;
                const int b = 31 - __clz((int)m);      // m == 0: b = -1
                const int u = 7 - b;
                idx[q] = m ? (int)(cur >> 8) + (u < 4 ? u * CS : STEP + (u - 4) * CS) : -1;
                cur &= ~((m ? 1u : 0u) << (b & 31));
            }
            if (idx[0] < 0) break;  // the list is consumed in order: no first pair, nothing left
#pragma unroll
            for (int q = 0; q < PF; ++q)
                if (idx[q] >= 0) visit(idx[q]);
        }
        lpa = list_a;
    };
#else
    // private list: entry e of thread t lives at list[e * NT + t] (conflict-free)
    unsigned short *const my_list = sm.list + tid;
    const uint32_t list_a = smem_u32(my_list);
    const uint32_t list_end_a = list_a + (uint32_t)sm.list_len * NT * 2u;
    constexpr uint32_t ESTEP = NT * 2u;  // bytes between consecutive entries of one thread
    uint32_t lpa = list_a;               // append position (shared-memory address)
    auto room = [&](int n) { return lpa + n * ESTEP <= list_end_a; };
    auto flush = [&]() {
        const unsigned short *e = my_list;
        const unsigned short *const lp = my_list + (lpa - list_a) / 2u;
        constexpr int PF = TPB_P2_UNROLL;  // pairs in flight for ILP
        for (; e + PF * NT <= lp; e += PF * NT) {
            int idx[PF];
#pragma unroll
            for (int u = 0; u < PF; ++u) idx[u] = e[u * NT];
#pragma unroll
            for (int u = 0; u < PF; ++u) visit(idx[u]);
        }
        for (; e < lp; e += NT) visit(e[0]);
        lpa = list_a;
    };
#endif

    while (true) {
        // ---- oversized neighbourhood: thread 0 stages the next chunk of whole rows (or one
        // piece of an oversized row)
        if (!staged) {
            if (tid == 0) {
                int q = hdr->q, gpos = hdr->gpos, used = 0, nseg = 0;
                uint32_t bytes = 0;
                while (q < NROWS && nseg < TILE_MAXSEG) {
                    const int gend_row = hdr->g1[q];
                    if (gpos >= gend_row) {
                        ++q;
                        if (q < NROWS) gpos = hdr->g0[q];
                        continue;
                    }
                    const int a = gpos & ~3;
                    const int need = ((gend_row + 3) & ~3) - a;
                    int gend;
                    if (need <= sm.cap - used) {
                        gend = gend_row;
                    } else if (nseg == 0) {
                        gend = a + sm.cap;  // piece of an oversized row (cap is a multiple of 4)
                    } else {
                        break;
                    }
                    const int len = ((gend + 3) & ~3) - a;
                    hdr->seg_row[nseg] = q;
                    hdr->seg_begin[nseg] = gpos;
                    hdr->seg_end[nseg] = gend;
                    hdr->seg_base[nseg] = used - a;
                    if (nseg == 0) fence_proxy_async();
                    bulk_g2s(sm.tA + used, nb.A + a, (uint32_t)(len * sizeof(V4<CT>)), sm.bar);
                    bulk_g2s(tB + used, nb.B + a, (uint32_t)(len * sizeof(R1)), sm.bar);
                    if constexpr (NB::HAS_P) bulk_g2s(sm.tP + used, nb.P + a, (uint32_t)(len * sizeof(T)), sm.bar);
                    if constexpr (HAS_F)
                        bulk_g2s(sm.tF + used, nb.F + a, (uint32_t)(len * sizeof(V4<float>)), sm.bar);
                    bytes += (uint32_t)len * REC_BYTES;
                    used += len;
                    ++nseg;
                    gpos = gend;
                    if (gend < gend_row) break;  // oversized row: continue with it next chunk
                }
                while (q < NROWS && gpos >= hdr->g1[q]) {  // skip exhausted / empty rows
                    ++q;
                    if (q < NROWS) gpos = hdr->g0[q];
                }
                hdr->q = q;
                hdr->gpos = gpos;
                hdr->nseg = nseg;
                hdr->last = q >= NROWS;
                if (nseg > 0) mbar_expect_tx(sm.bar, bytes);
            }
            __syncthreads();
        }
        staged = false;
        const int nseg = hdr->nseg;
        const bool last = hdr->last != 0;
        const int tab_w = hdr->tab_w;
        if (nseg == 0) break;

        mbar_wait(sm.bar, parity);
        parity ^= 1u;

        // ---- lock-step scan: all lanes of a warp walk the candidate run of their own cell
        // neighbourhood in step (lanes of the same cell read the same record: broadcast)
        for (int si = 0; si < nseg; ++si) {
            const int q = hdr->seg_row[si];
            const int dy = q % 3 - 1, dz = ND == 3 ? q / 3 - 1 : 0;
            int j = 0, j1 = 0;
            if (valid) {
                int lo, hi;
                if (tab_w > 0) {
                    const int *row = hdr->tab + q * tab_w + (cx - hdr->cxmin);
                    lo = row[0];
                    hi = row[2 * g.sx + 1];
                } else {
                    const int c0 = cell_linear(g, cx - g.sx, hdr->cy + dy, hdr->cz + dz);
                    lo = nb.cell_start[c0];
                    hi = nb.cell_start[c0 + 2 * g.sx + 1];
                }
                j = max(lo, hdr->seg_begin[si]);
                j1 = min(hi, hdr->seg_end[si]);
            }
            const int base = hdr->seg_base[si];
            const int t0 = base + j;      // tile index of the first candidate
            const int t1 = base + j1;     // one past the last
            // The KS threads of a target share its candidates.  TPB_INTERLEAVE: thread kg takes the
            // candidates kg, kg + KS, ... counted from the lane's own first candidate, so the accepted
            // pairs -- concentrated in the middle of every row -- are dealt out evenly and the three
            // lists of a target (and with them the lanes of a warp in phase 2) end up equally long;
            // otherwise groups of four consecutive candidates are dealt out.
            int t = t0 + (TPB_INTERLEAVE ? kg : 4 * kg);
            while (true) {
                // phase 1: filter candidates into the private list
#if TPB_LIST_MASKS
                while (t + STEP + 3 * CS < t1 && room(1)) {
                    FRec xc[8];
#pragma unroll
                    for (int u = 0; u < 4; ++u) xc[u] = tFilt[t + u * CS];
#pragma unroll
                    for (int u = 0; u < 4; ++u) xc[4 + u] = tFilt[t + STEP + u * CS];
                    uint32_t m = 0;
#pragma unroll
                    for (int u = 0; u < 8; ++u) m = push_verdict(m, xc[u]);
                    m &= 0xFFu;
                    lpa = entry_append<ESTEP>(lpa, ((uint32_t)t << 8) | m, m != 0);
                    t += 2 * STEP;
                }
                while (t < t1 && room(1)) {
                    // one group of four, the last one possibly partial (records past t1 are staged
                    // neighbours of other lanes or padding: read, never accepted)
                    FRec xc[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) xc[u] = tFilt[t + u * CS];
                    uint32_t m = 0;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        m = push_verdict(m, xc[u]);
                        if (t + u * CS >= t1) m &= ~1u;
                    }
                    m = (m & 0xFu) << 4;  // candidates 0..3 of the group: bits 7..4
                    lpa = entry_append<ESTEP>(lpa, ((uint32_t)t << 8) | m, m != 0);
                    t += STEP;
                }
#else
                while (t + STEP + 3 * CS < t1 && room(8)) {
                    FRec xc[8];
#pragma unroll
                    for (int u = 0; u < 4; ++u) xc[u] = tFilt[t + u * CS];
#pragma unroll
                    for (int u = 0; u < 4; ++u) xc[4 + u] = tFilt[t + STEP + u * CS];
#pragma unroll
                    for (int u = 0; u < 4; ++u) lpa = list_append<ESTEP>(lpa, (uint32_t)(t + u * CS), pass(xc[u]));
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        lpa = list_append<ESTEP>(lpa, (uint32_t)(t + STEP + u * CS), pass(xc[4 + u]));
                    t += 2 * STEP;
                }
                while (t < t1 && room(4)) {
                    // one group, the last one possibly partial (records past t1 are staged
                    // neighbours of other lanes or padding: read, never appended)
                    FRec xc[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) xc[u] = tFilt[t + u * CS];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        lpa = list_append<ESTEP>(lpa, (uint32_t)(t + u * CS), t + u * CS < t1 && pass(xc[u]));
                    t += STEP;
                }
#endif
                if (!__any_sync(0xffffffffu, t < t1)) break;
                flush();  // phase 2 (some list of the warp is full)
            }
        }
        flush();  // list entries point into this chunk: drain before it is replaced
        if (last) break;  // the next sweep's leading barrier protects the staged tile
        __syncthreads();
    }
}

// stage + sweep; must be called by all threads of the block
template <int KS, int ND, typename T, typename CT, typename NB, typename BODY>
__device__ __forceinline__ void tile_sweep(TileSmem<T, CT> &sm, const GridConst<CT> &g, const NB &nb,
                                           const int2 *__restrict__ rng, bool valid, int cx,
                                           const V4<CT> &xi, T radius2, uint32_t &parity, BODY &&body)
{
    __syncthreads();  // the previous sweep is done with hdr and the staged tile
    if (threadIdx.x < 32)
        tile_stage<ND, T, CT>(sm, nb, rng, sm.hdr->cxmin, sm.hdr->cxmax, sm.hdr->cy, sm.hdr->cz, g.sx, g.n[0],
                              g.n[1]);
    __syncthreads();
    tile_sweep_staged<KS, ND, T, CT>(sm, g, nb, valid, cx, xi, radius2, parity, body);
}

// bar.sync / bar.arrive on barrier `id` (1..4, warp-uniform) for COUNT threads; the id is an
// immediate so that the kernel reserves five hardware barriers, not all sixteen
template <bool WAIT, int COUNT>
__device__ __forceinline__ void named_barrier(int id)
{
#define TPB_BAR(ID)                                                                   \
    if (WAIT)                                                                         \
        asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory");            \
    else                                                                              \
        asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(COUNT) : "memory");
    switch (id) {
    case 1: TPB_BAR(1) break;
    case 2: TPB_BAR(2) break;
    case 3: TPB_BAR(3) break;
    default: TPB_BAR(4) break;
    }
#undef TPB_BAR
}

// Combine the KS partial results of every target: afterwards the threads with
// threadIdx.x < TILE_TB hold val = ((val_0 + val_1) + val_2) ...; the other threads are done.
// Only the KS warps that share 32 targets meet (named barrier 1 + warp % 4): the writers
// arrive and leave, the reader waits for them -- nobody waits for the slowest warp of the block.
// Scratch space without extra shared memory: the 16-bit list slots of entry e of W = sizeof(T) / 2
// neighbouring threads (a pair for Float32, four for Float64) form one aligned word of T; thread
// j of such a group parks its values in the words of the entries j * 4 .. j * 4 + 3 -- slots
// only this group of lanes ever touches, so warps that are still sweeping are not disturbed.
// RED: slots per thread (the list must hold RED * sizeof(T) / 2 entries).
template <int KS, int NVAL, typename T, typename CT, int RED = TILE_RED_VALS>
__device__ __forceinline__ void tile_reduce(TileSmem<T, CT> &sm, T (&val)[NVAL])
{
    static_assert(NVAL <= RED, "tile_reduce: scratch space too small");
    if constexpr (KS > 1) {
        constexpr int NT = KS * TILE_TB;
        constexpr int W = (int)sizeof(T) / 2;  // 16-bit slots per value
        const int ti = threadIdx.x % TILE_TB, kg = threadIdx.x / TILE_TB;
        const int bar_id = 1 + ti / 32;
#if TPB_LIST_MASKS
        // 32-bit list entries: entry e of thread t is the word list32[e * NT + t], so a thread parks value n
        // in its OWN entries n * W32 .. (private: warps that are still sweeping are not disturbed)
        constexpr int W32 = (int)sizeof(T) / 4;
        uint32_t *const l32 = reinterpret_cast<uint32_t *>(sm.list);
        auto put = [&](int tid, int n, T v) {
            if constexpr (W32 == 1) {
                l32[n * NT + tid] = __float_as_uint((float)v);
            } else {
                const unsigned long long b = (unsigned long long)__double_as_longlong((double)v);
                l32[(2 * n) * NT + tid] = (uint32_t)b;
                l32[(2 * n + 1) * NT + tid] = (uint32_t)(b >> 32);
            }
        };
        auto get = [&](int tid, int n) -> T {
            if constexpr (W32 == 1) {
                return (T)__uint_as_float(l32[n * NT + tid]);
            } else {
                const unsigned long long b = (unsigned long long)l32[(2 * n) * NT + tid] |
                                             ((unsigned long long)l32[(2 * n + 1) * NT + tid] << 32);
                return (T)__longlong_as_double((long long)b);
            }
        };
        (void)W;
        if (kg > 0) {
#pragma unroll
            for (int n = 0; n < NVAL; ++n) put(threadIdx.x, n, val[n]);
            __threadfence_block();
            named_barrier<false, 32 * KS>(bar_id);
        } else {
            named_barrier<true, 32 * KS>(bar_id);
#pragma unroll
            for (int k = 1; k < KS; ++k)
#pragma unroll
                for (int n = 0; n < NVAL; ++n) val[n] += get(k * TILE_TB + ti, n);
        }
#else
        auto slot = [&](int tid, int n) {
            return reinterpret_cast<T *>(sm.list + (n + RED * (tid % W)) * NT + (tid - tid % W));
        };
        __syncwarp();  // both lanes of a pair have drained their lists
        if (kg > 0) {
#pragma unroll
            for (int n = 0; n < NVAL; ++n) *slot(threadIdx.x, n) = val[n];
            __threadfence_block();
            named_barrier<false, 32 * KS>(bar_id);
        } else {
            named_barrier<true, 32 * KS>(bar_id);
#pragma unroll
            for (int k = 1; k < KS; ++k)
#pragma unroll
                for (int n = 0; n < NVAL; ++n) val[n] += *slot(k * TILE_TB + ti, n);
        }
#endif
    }
}

// ------------------------------------------------------------------ interact! (variant 2)
// NOSLIP (boundary_model.viscosity !== nothing): the wall records are (x, m | v_w, rho_w | p_w) -- the
// shape of the fluid's -- and the wall sweep adds the wall model's viscous term with v_b = v_w
// (dv_viscosity!, viscosity.jl:9-40; viscous_velocity, wall_boundary/system.jl:148-163) right after the
// pressure term, as rhs.jl:85-96 does.  The free-slip instantiation is unchanged.
template <int KS, int ND, typename T, typename CT, int KERNEL, int DENS, bool NOSLIP = false>
__global__ void __launch_bounds__(KS * TILE_TB, 2)
k_interact_tiles(GridConst<CT> g, const int *__restrict__ n_tiles, const int4 *__restrict__ tile_desc,
                 const int4 *__restrict__ tile_ext, const int2 *__restrict__ tile_rng,
                 const int *__restrict__ fcell_start, const V4<CT> *__restrict__ A,
                 const V4<T> *__restrict__ B, const T *__restrict__ P,
                 const int *__restrict__ perm, int ff_enabled, int has_wall,
                 const int *__restrict__ wcell_start, const V4<CT> *__restrict__ Aw,
                 const V2<T> *__restrict__ Ww, PairConst<T> k, SourceConst<T> src,
                 T *__restrict__ dv, int n_targets, int cap, int list_len,
                 const V4<float> *__restrict__ Ff, const V4<float> *__restrict__ Fw,
                 const V4<T> *__restrict__ Vw = nullptr, const T *__restrict__ Pw = nullptr,
                 WallViscConst<T> wk = WallViscConst<T>(), const AdaptConsts<T> *__restrict__ ad = nullptr)
{
    constexpr int NV = DENS == 0 ? ND + 1 : ND;
    extern __shared__ __align__(16) unsigned char tile_smem_raw[];
    const int tile = blockIdx.x;
    if (tile >= *n_tiles) return;
    if (ad) {  // StateEquationAdaptiveCole: this kick's speed of sound (k_adaptive_consts)
        k.c = ad->c;
        k.delta_h_c = ad->delta_h_c;
        if constexpr (NOSLIP) {
            wk.c = ad->c;
            wk.nu_a = ad->nu_a;
            wk.nu_b = ad->nu_b;
        }
    }
    TileSmem<T, CT> sm(tile_smem_raw, cap, list_len, KS * TILE_TB);
    const int2 *rng = tile_rng + (int64_t)tile * 18;
    const NbSet<T, CT, V4<T>, true> nb_f{fcell_start, A, B, P, Ff};
    // every thread reads the descriptor itself (one broadcast transaction) and starts loading its
    // own particle while warp 0 is staging; the barrier below also publishes the header
    const int4 desc = tile_desc[tile];
    const int4 ext = tile_ext[tile];
    if (threadIdx.x < 32) {
        // warp 0 starts the TMA copies of the fluid sweep before anything else happens
        if (threadIdx.x == 0) {
            tile_locate(sm.hdr, g.n[1], desc, ext);
            mbar_init(sm.bar, 1);
        }
        __syncwarp();
        if (ff_enabled)
            tile_stage<ND, T, CT>(sm, nb_f, rng, ext.x, ext.y, desc.z % g.n[1], desc.z / g.n[1], g.sx, g.n[0],
                                  g.n[1]);
    }
    const int s = desc.x + threadIdx.x % TILE_TB;
    const bool in_tile = s < desc.y;
    V4<CT> xi = {};
    V4<T> bi = {};
    T p_a = 0;
    int orig = n_targets;
    if (in_tile) {
        orig = perm[s];
        xi = A[s];
        bi = B[s];
        p_a = P[s];
    }
    // slab ghosts (original index >= n_targets) are neighbours only: no dv is computed for them
    const bool valid = in_tile && orig < n_targets;
    const bool any_fw = ext.w > 0;
    uint32_t parity = 0;
    if (!__syncthreads_or(valid)) {
        // ghost-only tile: let the copies in flight land before the shared memory is released
        if (ff_enabled && sm.hdr->nseg > 0) mbar_wait(sm.bar, parity);
        return;
    }
    int cx = ext.x, cy, cz;
    if (valid) cell_coords<ND, CT>(g, xi.x, xi.y, xi.z, cx, cy, cz);
    const T rho_a = bi.w;
    const T v_a[3] = {bi.x, bi.y, bi.z};

    constexpr bool FAST = std::is_same<T, float>::value;  // Float32 fields; coordinates float or double
    FastConst fc = {};
    float pa_term = 0.f;
    if constexpr (FAST) {
        fc = make_fast_const(k);
        pa_term = valid ? p_a / (rho_a * rho_a) : 0.f;
    }

    // acc[0..3]: fluid-fluid sums (dv, drho); acc[4..7]: fluid-wall sums
    T acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    T(&dv_ff)[3] = *reinterpret_cast<T(*)[3]>(acc);
    T &drho_ff = acc[3];
    if (ff_enabled) {
        tile_sweep_staged<KS, ND, T, CT>(sm, g, nb_f, valid, cx, xi, k.radius2, parity,
                              [&](const V4<CT> &xj, const V4<T> &bj, T pj) {
                                  if constexpr (FAST) {
                                      interact_pair_fast<ND, KERNEL, DENS, true>(
                                          fc, xi, xj, rho_a, p_a, pa_term, v_a, bj.x, bj.y, bj.z, bj.w, pj,
                                          dv_ff, drho_ff);
                                      return;
                                  }
                                  T pd[3];
                                  const T d2 = pos_diff_d2<ND, T, CT>(xi, xj, pd);
                                  if (d2 <= k.radius2) {
                                      const T dist = sqrt_rn(d2);
                                      if (dist >= k.almostzero) {
                                          const T v_b[3] = {bj.x, bj.y, bj.z};
                                          interact_pair<ND, T, KERNEL, DENS, true>(
                                              k, (T)xj.w, rho_a, bj.w, p_a, pj, v_a, v_b, pd, dist, dv_ff,
                                              drho_ff, (T)xi.w);
                                      }
                                  }
                              });
    }
    T(&dv_fw)[3] = *reinterpret_cast<T(*)[3]>(acc + 4);
    T &drho_fw = acc[7];
    if constexpr (NOSLIP) {
        if (has_wall && any_fw) {
            const T zero3[3] = {0, 0, 0};
            NbSet<T, CT, V4<T>, true> nb{wcell_start, Aw, Vw, Pw, Fw};
            tile_sweep<KS, ND, T, CT>(sm, g, nb, rng + 9, valid, cx, xi, k.radius2, parity,
                                  [&](const V4<CT> &xj, const V4<T> &wj, T pj) {
                                      if constexpr (FAST)
                                          interact_pair_fast<ND, KERNEL, DENS, false>(
                                              fc, xi, xj, rho_a, p_a, pa_term, v_a, 0.f, 0.f, 0.f, wj.w, pj,
                                              dv_fw, drho_fw);
                                      T pd[3];
                                      const T d2 = pos_diff_d2<ND, T, CT>(xi, xj, pd);
                                      if (d2 <= k.radius2) {
                                          const T dist = sqrt_rn(d2);
                                          if (dist >= k.almostzero) {
                                              if constexpr (!FAST)
                                                  interact_pair<ND, T, KERNEL, DENS, false>(
                                                      k, (T)xj.w, rho_a, wj.w, p_a, pj, v_a, zero3, pd, dist,
                                                      dv_fw, drho_fw);
                                              wall_viscous_term<ND, T, KERNEL>(wk, (T)xi.w, rho_a, v_a, (T)xj.w,
                                                                               wj, pd, dist, dv_fw);
                                          }
                                      }
                                  });
        }
    } else if (has_wall && any_fw) {
        const T zero3[3] = {0, 0, 0};
        NbSet<T, CT, V2<T>, false> nb{wcell_start, Aw, Ww, nullptr, Fw};
        tile_sweep<KS, ND, T, CT>(sm, g, nb, rng + 9, valid, cx, xi, k.radius2, parity,
                              [&](const V4<CT> &xj, const V2<T> &wj, T) {
                                  if constexpr (FAST) {
                                      interact_pair_fast<ND, KERNEL, DENS, false>(
                                          fc, xi, xj, rho_a, p_a, pa_term, v_a, 0.f, 0.f, 0.f, wj.y, wj.x,
                                          dv_fw, drho_fw);
                                      return;
                                  }
                                  T pd[3];
                                  const T d2 = pos_diff_d2<ND, T, CT>(xi, xj, pd);
                                  if (d2 <= k.radius2) {
                                      const T dist = sqrt_rn(d2);
                                      if (dist >= k.almostzero)
                                          interact_pair<ND, T, KERNEL, DENS, false>(
                                              k, (T)xj.w, rho_a, wj.y, p_a, wj.x, v_a, zero3, pd, dist,
                                              dv_fw, drho_fw);
                                  }
                              });
    }
    // the threads of a target combine (S_ff + S_fw); Float64 (one thread per target) keeps the
    // reference's order dv = ((0 + S_ff) + S_fw) + g
    T part[4] = {acc[0] + acc[4], acc[1] + acc[5], acc[2] + acc[6], acc[3] + acc[7]};
    tile_reduce<KS, 4>(sm, part);
    if (!valid || threadIdx.x >= TILE_TB) return;
    // dv = ((0 + S_ff) + S_fw) + g [+ source]  (semidiscretization.jl:600, :809-829, :668-731)
    const int64_t o = (int64_t)orig * NV;
    T out[4] = {0, 0, 0, 0};
#pragma unroll
    for (int d = 0; d < ND; ++d) {
        T val = part[d];
        if (src.any) {
            val += src.acc[d];
            if (src.damping != (T)0) val += -src.damping * v_a[d];
        }
        out[d] = val;
    }
    if (DENS == 0) out[ND] = part[3];
    // one 16-byte store per particle where the layout allows it (3-D Float32 with density):
    // dv may be a mapped host buffer, where every store instruction is a PCIe write
    if (NV == 4 && sizeof(T) == 4 && (reinterpret_cast<uintptr_t>(dv) & 15) == 0) {
        V4<T> o4;
        o4.x = out[0], o4.y = out[1], o4.z = out[2], o4.w = out[3];
        *reinterpret_cast<V4<T> *>(dv + o) = o4;
    } else {
#pragma unroll
        for (int d = 0; d < NV; ++d) dv[o + d] = out[d];
    }
}

// ------------------------------------------------------------------ Adami (variant 2)
// Most wall particles of a tank are far from any fluid.  One warp per wall tile checks whether
// the tile's neighbour rows hold any fluid particle: if not, it writes the result of an empty
// sum (p = 0, rho = EOS^-1(0), volume = 0) directly; otherwise the tile is appended to the
// active list that k_adami_tiles runs over.
template <typename T, typename CT>
struct WallPrepArgs {
    const int *n_tiles;
    const int4 *tile_desc;
    const V4<CT> *Aw;
    T rho_empty;
    V2<T> *W;
    T *volume;
    int *active, *n_active;
    int2 *rng;
    int4 *ext;
    V4<T> *Vw;  // no-slip wall: (v_w, rho_w) records; else nullptr
    T *Pw;      // no-slip wall: p_w as a scalar array
    unsigned char *state;  // per tile: 1 = the empty-sum values are in place since an earlier kick
    int rewrite;           // the empty-sum values have changed (adaptive sound speed): write them again
    const AdaptConsts<T> *ad;  // StateEquationAdaptiveCole shared with the fluid: rho_empty comes from there
};
template <int ND, typename T, typename CT>
__device__ __forceinline__ void
wall_tile_prep_body(int vblock, const GridConst<CT> &g, const int *__restrict__ n_tiles,
                    const int4 *__restrict__ tile_desc, const V4<CT> *__restrict__ Aw,
                    const int *__restrict__ fcell_start, T rho_empty, V2<T> *__restrict__ W,
                    T *__restrict__ volume, int *__restrict__ active, int *__restrict__ n_active,
                    int2 *__restrict__ rng, int4 *__restrict__ ext, V4<T> *__restrict__ Vw, T *__restrict__ Pw,
                    unsigned char *__restrict__ state = nullptr, int rewrite = 1)
{
    constexpr int NROWS = ND == 3 ? 9 : 3;
    const int tile = (vblock * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (tile >= *n_tiles) return;
    const int4 d = tile_desc[tile];
    int cxmin, cxmax, cy, cz;
    {
        const V4<CT> xa = Aw[d.x], xb = Aw[d.y - 1];
        cell_coords<ND, CT>(g, xa.x, xa.y, xa.z, cxmin, cy, cz);
        cell_coords<ND, CT>(g, xb.x, xb.y, xb.z, cxmax, cy, cz);
    }
    int cnt = 0;
    if (lane < NROWS) {
        const int dy = lane % 3 - 1, dz = ND == 3 ? lane / 3 - 1 : 0;
        const int c_lo = cell_linear(g, cxmin - g.sx, cy + dy, cz + dz);
        const int g0 = fcell_start[c_lo], g1 = fcell_start[c_lo + (cxmax - cxmin) + 2 * g.sx + 1];
        cnt = g1 - g0;
        rng[(int64_t)tile * 9 + lane] = make_int2(g0, g1);
    }
    int total = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    if (total > 0) {
        if (lane == 0) {
            ext[tile] = make_int4(cxmin, cxmax, total, 0);
            active[atomicAdd(n_active, 1)] = tile;
            if (state) state[tile] = 0;
        }
        return;
    }
    // Most wall tiles of a tank never see fluid: their empty-sum values are written once, not every kick
    // (at 10 M fluid particles the 7.3 M wall particles would cost 88 MB of stores per kick).
    if (state) {
        const bool in_place = state[tile] == 1 && !rewrite;
        __syncwarp();
        if (in_place) return;
        if (lane == 0) state[tile] = 1;
    }
    V2<T> empty;
    empty.x = (T)0;
    empty.y = rho_empty;
    V4<T> vempty;
    vempty.x = vempty.y = vempty.z = (T)0;
    vempty.w = rho_empty;
    for (int w = d.x + lane; w < d.y; w += 32) {
        W[w] = empty;
        volume[w] = (T)0;
        if (Vw) {
            Vw[w] = vempty;
            Pw[w] = (T)0;
        }
    }
}
template <int ND, typename T, typename CT>
__global__ void __launch_bounds__(256)
k_wall_tile_prep(GridConst<CT> g, const int *__restrict__ n_tiles, const int4 *__restrict__ tile_desc,
                 const V4<CT> *__restrict__ Aw, const int *__restrict__ fcell_start, T rho_empty,
                 V2<T> *__restrict__ W, T *__restrict__ volume, int *__restrict__ active,
                 int *__restrict__ n_active, int2 *__restrict__ rng, int4 *__restrict__ ext,
                 V4<T> *__restrict__ Vw /* no-slip wall: (v_w, rho_w) records; else nullptr */,
                 T *__restrict__ Pw /* no-slip wall: p_w as a scalar array */,
                 unsigned char *__restrict__ state, int rewrite)
{
    wall_tile_prep_body<ND, T, CT>(blockIdx.x, g, n_tiles, tile_desc, Aw, fcell_start, rho_empty, W, volume, active,
                                   n_active, rng, ext, Vw, Pw, state, rewrite);
}

// One launch for the three independent steps that follow the cell scan: blocks [0, nb_scatter) scatter the
// particle indices (k_scatter), the next nb_ranges blocks compute the candidate ranges of the fluid tiles
// (k_tile_ranges), the rest sort the wall tiles into empty / active (k_wall_tile_prep).  At a million
// particles each of them runs for a few microseconds: as separate launches their start-up and drain
// dominate.
template <int ND, typename T, typename CT>
__global__ void __launch_bounds__(256)
k_post_scan(int nb_scatter, int nb_ranges, const int *__restrict__ key, const int *__restrict__ slot,
            const int *__restrict__ fcell_start, int n, int *__restrict__ tmp_perm, GridConst<CT> g,
            const int *__restrict__ n_ftiles, const int4 *__restrict__ fdesc,
            const int *__restrict__ wcell_start, int2 *__restrict__ frng, int4 *__restrict__ fext,
            WallPrepArgs<T, CT> wp, int pbits = 31)
{
    int vb = blockIdx.x;
    if (vb < nb_scatter) {
        const int i = vb * blockDim.x + threadIdx.x;
        if (i < n && key[i] >= 0) {
            const int sl = slot[i];  // arrival number | position key << pbits (k_cell_count)
            if (pbits < 31)
                tmp_perm[fcell_start[key[i]] + (sl & ((1 << pbits) - 1))] = (sl & ~((1 << pbits) - 1)) | i;
            else
                tmp_perm[fcell_start[key[i]] + sl] = i;
        }
        return;
    }
    vb -= nb_scatter;
    if (vb < nb_ranges) {
        tile_ranges_body<ND>(vb, g.n[0], g.n[1], g.sx, n_ftiles, fdesc, fcell_start, fcell_start, wcell_start, frng,
                             18, fext);
        return;
    }
    vb -= nb_ranges;
    wall_tile_prep_body<ND, T, CT>(vb, g, wp.n_tiles, wp.tile_desc, wp.Aw, fcell_start,
                                   wp.ad ? wp.ad->rho_empty_w : wp.rho_empty, wp.W, wp.volume, wp.active,
                                   wp.n_active, wp.rng, wp.ext, wp.Vw, wp.Pw, wp.state, wp.rewrite);
}

// Targets: wall particles (active tiles over the wall's sorted order); neighbours: fluid.
// An active tile is little work (a wall particle has fluid on one side only, ~25 neighbours), so
// the kernel is bound by the latency of getting a tile started: warp 0 issues the TMA copies
// straight from the descriptors while the other warps load their particles, and the Float32
// version keeps three blocks per SM resident (launch bounds cap it at 56 registers, shared memory
// at 72 KB).  Blocks walk the active list with a grid stride; with the default grid (one block per
// tile slot) the hardware scheduler balances the tiles, which measured better than a persistent
// grid (0.100 vs 0.113 ms).
// NOSLIP (boundary_model.viscosity !== nothing): the same sweep also interpolates the fluid velocity,
// v_w = -sum_f v_f W / sum_f W (interpolate_fluid_velocity! / compute_wall_velocity!,
// dummy_particles.jl:710-758), and writes the records (v_w, rho_w | p_w) that k_interact_tiles<NOSLIP> stages;
// three more accumulators, so two blocks per SM.  The free-slip instantiation is unchanged.
constexpr int ADAMI_BLOCKS_PER_SM = 3;
constexpr int ADAMI_NOSLIP_RED = 6;  // tile_reduce slots of the no-slip version (5 values)
template <int KS, int ND, typename T, typename CT, int KERNEL, bool NOSLIP = false>
__global__ void __launch_bounds__(KS * TILE_TB, (KS > 1 && sizeof(T) == 4 && !NOSLIP) ? ADAMI_BLOCKS_PER_SM : 2)
k_adami_tiles(GridConst<CT> g, const int *__restrict__ n_active, const int *__restrict__ active,
              const int4 *__restrict__ tile_desc, const int4 *__restrict__ tile_ext,
              const int2 *__restrict__ tile_rng, const int *__restrict__ wcell_start, const V4<CT> *__restrict__ Aw,
              const int *__restrict__ fcell_start, const V4<CT> *__restrict__ A,
              const V4<T> *__restrict__ B, const T *__restrict__ P, int interaction_enabled,
              AdamiConst<T> k, V2<T> *__restrict__ W, T *__restrict__ volume, int cap, int list_len,
              const V4<float> *__restrict__ Ff, V4<T> *__restrict__ Vw = nullptr, T *__restrict__ Pw = nullptr,
              const AdaptConsts<T> *__restrict__ ad = nullptr)
{
    extern __shared__ __align__(16) unsigned char tile_smem_raw[];
    const int n_act = *n_active;
    if ((int)blockIdx.x >= n_act) return;
    if (ad) k.eos.B = ad->B_w;  // the boundary model shares the fluid's StateEquationAdaptiveCole
    TileSmem<T, CT> sm(tile_smem_raw, cap, list_len, KS * TILE_TB);
    if (threadIdx.x == 0) mbar_init(sm.bar, 1);
    uint32_t parity = 0;
    const NbSet<T, CT, V4<T>, true> nb{fcell_start, A, B, P, Ff};
    for (int it = blockIdx.x; it < n_act; it += gridDim.x) {
        // the previous tile is done with the staged records, the header and the reduction slots
        __syncthreads();
        const int tile = active[it];
        const int4 desc = tile_desc[tile];
        const int4 ext = tile_ext[tile];
        if (threadIdx.x < 32) {
            if (threadIdx.x == 0) tile_locate(sm.hdr, g.n[1], desc, ext);
            __syncwarp();
            if (interaction_enabled)
                tile_stage<ND, T, CT>(sm, nb, tile_rng + (int64_t)tile * 9, ext.x, ext.y, desc.z % g.n[1],
                                      desc.z / g.n[1], g.sx, g.n[0], g.n[1]);
        }
        const int w = desc.x + threadIdx.x % TILE_TB;
        const bool valid = w < desc.y;
        V4<CT> xi = {};
        int cx = ext.x, cy, cz;
        if (valid) {
            xi = Aw[w];
            cell_coords<ND, CT>(g, xi.x, xi.y, xi.z, cx, cy, cz);
        }
        __syncthreads();
        constexpr int NACC = NOSLIP ? 5 : 2;
        T acc[NACC] = {};
        T &p = acc[0], &vol = acc[1];
        if (interaction_enabled) {
            tile_sweep_staged<KS, ND, T, CT>(
                sm, g, nb, valid, cx, xi, k.radius2, parity, [&](const V4<CT> &xj, const V4<T> &bj, T pj) {
                    T pd[3];
                    const T d2 = pos_diff_d2<ND, T, CT>(xi, xj, pd);
                    if constexpr (std::is_same<T, float>::value) {
                        // branch-free Float32 form: a rejected pair gets weight 0
                        const bool ok = d2 <= k.radius2;
                        const float d2s = ok ? d2 : 1.0f;
                        const float dist = d2s * rsqrt_approx(fmaxf(d2s, 1e-30f));
                        float hyd = k.acc[0] * (bj.w * pd[0]) + k.acc[1] * (bj.w * pd[1]);
                        if (ND == 3) hyd += k.acc[2] * (bj.w * pd[2]);
                        const float kw = ok ? kernel_safe<KERNEL, float>(k.kern, dist) : 0.0f;
                        p = fmaf(k.p_off + pj + hyd, kw, p);
                        vol += kw;
                        if constexpr (NOSLIP) {
                            acc[2] = fmaf(kw, bj.x, acc[2]);
                            acc[3] = fmaf(kw, bj.y, acc[3]);
                            if (ND == 3) acc[4] = fmaf(kw, bj.z, acc[4]);
                        }
                        return;
                    }
                    if (d2 <= k.radius2) {
                        const T dist = sqrt_rn(d2);
                        const T rho_f = bj.w;
                        T hyd = k.acc[0] * (rho_f * pd[0]) + k.acc[1] * (rho_f * pd[1]);
                        if (ND == 3) hyd += k.acc[2] * (rho_f * pd[2]);
                        const T sum_p = k.p_off + pj + hyd;
                        const T kw = kernel_safe<KERNEL, T>(k.kern, dist);
                        p += sum_p * kw;
                        vol += kw;
                        if constexpr (NOSLIP) {
                            acc[2] += kw * bj.x;
                            acc[3] += kw * bj.y;
                            if (ND == 3) acc[4] += kw * bj.z;
                        }
                    }
                });
        }
        if constexpr (NOSLIP)
            tile_reduce<KS, NACC, T, CT, ADAMI_NOSLIP_RED>(sm, acc);
        else
            tile_reduce<KS, 2>(sm, acc);
        if (valid && threadIdx.x < TILE_TB) {
            const bool has_fluid = (double)vol > 2.220446049250313e-16;  // `volume > eps()`: eps(Float64)
            if (has_fluid) p = p / vol;
            if (k.clip) p = p > (T)0 ? p : (T)0;
            V2<T> out;
            out.x = p;
            out.y = eos_inverse(k.eos, p);
            W[w] = out;
            volume[w] = vol;
            if constexpr (NOSLIP) {
                V4<T> vw;
                vw.x = has_fluid ? (T)0 - acc[2] / vol : acc[2];
                vw.y = has_fluid ? (T)0 - acc[3] / vol : acc[3];
                vw.z = ND == 3 ? (has_fluid ? (T)0 - acc[4] / vol : acc[4]) : (T)0;
                vw.w = out.y;
                Vw[w] = vw;
                Pw[w] = p;
            }
        }
    }
}

// ------------------------------------------------------------------ summation density (variant 2)
// summation_density! (general/density_calculators.jl:26-50): rho_a = sum_b m_b W(r_ab) over the fluid
// itself and the wall, then the equation of state.  The same tile sweep as interact! with a W-only
// body: only the position / mass records are staged (the permutation rides in the second slot).
template <int KS, int ND, typename T, typename CT, int KERNEL>
__global__ void __launch_bounds__(KS * TILE_TB, 2)
k_summation_tiles(GridConst<CT> g, const int *__restrict__ n_tiles, const int4 *__restrict__ tile_desc,
                  const int4 *__restrict__ tile_ext, const int2 *__restrict__ tile_rng,
                  const int *__restrict__ fcell_start, const V4<CT> *__restrict__ A, const int *__restrict__ perm_f,
                  int has_wall, const int *__restrict__ wcell_start, const V4<CT> *__restrict__ Aw,
                  const int *__restrict__ perm_w, KernelConst<T> kern, T radius2, EosConst<T> eos,
                  V4<T> *__restrict__ B, T *__restrict__ P, int cap, int list_len,
                  const V4<float> *__restrict__ Ff, const V4<float> *__restrict__ Fw)
{
    extern __shared__ __align__(16) unsigned char tile_smem_raw[];
    const int tile = blockIdx.x;
    if (tile >= *n_tiles) return;
    TileSmem<T, CT> sm(tile_smem_raw, cap, list_len, KS * TILE_TB);
    const int2 *rng = tile_rng + (int64_t)tile * 18;
    const NbSet<T, CT, int, false> nb_f{fcell_start, A, perm_f, nullptr, Ff};
    const int4 desc = tile_desc[tile];
    const int4 ext = tile_ext[tile];
    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) {
            tile_locate(sm.hdr, g.n[1], desc, ext);
            mbar_init(sm.bar, 1);
        }
        __syncwarp();
        tile_stage<ND, T, CT>(sm, nb_f, rng, ext.x, ext.y, desc.z % g.n[1], desc.z / g.n[1], g.sx, g.n[0], g.n[1]);
    }
    const int s = desc.x + threadIdx.x % TILE_TB;
    const bool valid = s < desc.y;
    V4<CT> xi = {};
    int cx = ext.x, cy, cz;
    if (valid) {
        xi = A[s];
        cell_coords<ND, CT>(g, xi.x, xi.y, xi.z, cx, cy, cz);
    }
    __syncthreads();
    uint32_t parity = 0;
    T rho[1] = {(T)0};
    auto body = [&](const V4<CT> &xj, const int &, T) {
        T pd[3];
        const T d2 = pos_diff_d2<ND, T, CT>(xi, xj, pd);
        if constexpr (std::is_same<T, float>::value) {
            // exact square root: the equation of state amplifies a density error by gamma rho0 / (rho - rho0)
            const bool ok = d2 <= radius2;
            const float dist = sqrt_rn(ok ? d2 : 0.0f);
            rho[0] = fmaf(ok ? (float)xj.w : 0.0f, kernel_safe<KERNEL, float>(kern, dist), rho[0]);
        } else {
            if (d2 <= radius2) rho[0] += (T)xj.w * kernel_safe<KERNEL, T>(kern, sqrt_rn(d2));
        }
    };
    tile_sweep_staged<KS, ND, T, CT>(sm, g, nb_f, valid, cx, xi, radius2, parity, body);
    if (has_wall && ext.w > 0) {
        const NbSet<T, CT, int, false> nb_w{wcell_start, Aw, perm_w, nullptr, Fw};
        tile_sweep<KS, ND, T, CT>(sm, g, nb_w, rng + 9, valid, cx, xi, radius2, parity, body);
    }
    tile_reduce<KS, 1>(sm, rho);
    if (!valid || threadIdx.x >= TILE_TB) return;
    V4<T> b = B[s];
    b.w = rho[0];
    B[s] = b;
    P[s] = eos_pressure(eos, rho[0]);
}

// ------------------------------------------------------------------ neighbour pair dump
// Test hook behind tpb_neighbor_pairs (variant 2): the same tile sweep, filter and window
// clipping as interact!, with a body that records (orig_i, orig_j) of every pair accepted by
// the exact predicate.  The neighbour permutation rides in the second record slot.
template <int KS, int ND, typename T, typename CT>
__global__ void __launch_bounds__(KS * TILE_TB, 2)
k_pairs_tiles(GridConst<CT> g, const int *__restrict__ n_tiles, const int4 *__restrict__ tile_desc,
              const int4 *__restrict__ tile_ext, const int2 *__restrict__ tile_rng,
              const V4<CT> *__restrict__ X, const int *__restrict__ perm_x,
              const int *__restrict__ ycell_start, const V4<CT> *__restrict__ Y,
              const int *__restrict__ perm_y, T radius2, long long capacity, int *__restrict__ out_i,
              int *__restrict__ out_j, unsigned long long *__restrict__ counter, int cap, int list_len,
              const V4<float> *__restrict__ Fy)
{
    extern __shared__ __align__(16) unsigned char tile_smem_raw[];
    const int tile = blockIdx.x;
    if (tile >= *n_tiles) return;
    TileSmem<T, CT> sm(tile_smem_raw, cap, list_len, KS * TILE_TB);
    if (threadIdx.x == 0) {
        tile_locate(sm.hdr, g.n[1], tile_desc[tile], tile_ext[tile]);
        mbar_init(sm.bar, 1);
    }
    __syncthreads();
    const int s = sm.hdr->p0 + threadIdx.x % TILE_TB;
    const bool valid = s < sm.hdr->p1;
    V4<CT> xi = {};
    int cx = sm.hdr->cxmin, cy, cz, orig_i = 0;
    if (valid) {
        xi = X[s];
        orig_i = perm_x[s];
        cell_coords<ND, CT>(g, xi.x, xi.y, xi.z, cx, cy, cz);
    }
    uint32_t parity = 0;
    NbSet<T, CT, int, false> nb{ycell_start, Y, perm_y, nullptr, Fy};
    tile_sweep<KS, ND, T, CT>(sm, g, nb, tile_rng + (int64_t)tile * 9, valid, cx, xi, radius2, parity, [&](const V4<CT> &xj, const int &pj, T) {
        T pd[3];
        if (pos_diff_d2<ND, T, CT>(xi, xj, pd) <= radius2) {
            unsigned long long at = atomicAdd(counter, 1ull);
            if ((long long)at < capacity) {
                out_i[at] = orig_i;
                out_j[at] = pj;
            }
        }
    });
}

// ------------------------------------------------------------------ host side
struct TileState {
    int *d_row_tiles = nullptr;        // [nrows]
    int *d_frow_tile_start = nullptr;  // [nrows + 1] fluid tiles (rebuilt every kick)
    int *d_wrow_tile_start = nullptr;  // [nrows + 1] wall tiles (static)
    int4 *d_ftile_desc = nullptr;      // [max_ftiles]
    int4 *d_wtile_desc = nullptr;      // [max_wtiles]
    int4 *d_ftile_ext = nullptr;       // [max_ftiles] (cxmin, cxmax, fluid candidates, wall candidates)
    int4 *d_wtile_ext = nullptr;       // [max_wtiles] (cxmin, cxmax, fluid candidates, 0), active tiles only
    int2 *d_ftile_rng = nullptr;       // [max_ftiles][18] candidate ranges: 9 fluid rows, 9 wall rows
    int2 *d_wtile_rng = nullptr;       // [max_wtiles][9]  candidate ranges of the wall tiles in the fluid
    int4 *d_ptile_ext = nullptr;       // tpb_neighbor_pairs scratch
    int2 *d_ptile_rng = nullptr;
    int *d_wactive = nullptr;          // [max_wtiles] wall tiles with fluid in reach (per kick)
    int *d_n_wactive = nullptr;        // [1]
    unsigned char *d_wtile_state = nullptr;  // [max_wtiles] 1: tile without fluid in reach, empty-sum values in place
    double wall_rho_empty = -1;        // the value those tiles hold
    int nrows = 0;
    int max_ftiles = 0, max_wtiles = 0;
    int smem_budget = 112 * 1024;  // bytes per block: two blocks per SM
    int list_len = 160;            // private list entries per thread (one flush per sweep in 3-D)
    int list_len_split = 64;       // the same with TPB_SPLIT (three) threads per target
    int list_len_split2 = 96;      // two threads per target (Float64)
    int list(int ks) const { return ks > 2 ? list_len_split : ks == 2 ? list_len_split2 : list_len; }
    // Adami sweep with TPB_SPLIT threads per target: three blocks per SM, short lists
    int adami_smem_budget = 72 * 1024, adami_list_len = 32;  // 3 x (72 + 1) KB fit the 228 KB of an SM
};

inline int tiles_alloc(TileState &t, int nrows, int64_t n_f, int64_t n_w)
{
    t.nrows = nrows;
    // tuning overrides (bytes of shared memory per block, list entries per thread)
    if (const char *e = getenv("TPB_TILE_SMEM")) t.smem_budget = atoi(e);
    if (const char *e = getenv("TPB_TILE_LIST")) t.list_len = atoi(e);
    if (const char *e = getenv("TPB_TILE_LIST_SPLIT")) t.list_len_split = atoi(e);
    if (const char *e = getenv("TPB_ADAMI_SMEM")) t.adami_smem_budget = atoi(e);
    if (const char *e = getenv("TPB_ADAMI_LIST")) t.adami_list_len = atoi(e);
    t.list_len = std::max(t.list_len, 8);
    t.list_len_split = std::max(t.list_len_split, 8);
    if (const char *e = getenv("TPB_TILE_LIST_SPLIT2")) t.list_len_split2 = atoi(e);
    t.list_len_split2 = std::max(t.list_len_split2, 16);  // tile_reduce: 4 values x 4 threads per word
    t.adami_list_len = std::max(t.adami_list_len, 8);
    if (t.smem_budget > 227 * 1024) t.smem_budget = 227 * 1024;
    t.max_ftiles = (int)((n_f + TILE_TB - 1) / TILE_TB) + nrows;
    t.max_wtiles = (int)((n_w + TILE_TB - 1) / TILE_TB) + nrows;  // grown by tiles_reserve_wall if needed
    if (cudaMalloc(&t.d_row_tiles, sizeof(int) * (size_t)(nrows + 4)) != cudaSuccess) return 1;
    if (cudaMalloc(&t.d_frow_tile_start, sizeof(int) * (size_t)(nrows + 4)) != cudaSuccess) return 1;
    if (cudaMalloc(&t.d_wrow_tile_start, sizeof(int) * (size_t)(nrows + 4)) != cudaSuccess) return 1;
    if (cudaMalloc(&t.d_ftile_desc, sizeof(int4) * (size_t)t.max_ftiles) != cudaSuccess) return 1;
    if (cudaMalloc(&t.d_wtile_desc, sizeof(int4) * (size_t)t.max_wtiles) != cudaSuccess) return 1;
    const size_t max_tiles = (size_t)std::max(t.max_ftiles, t.max_wtiles);
    if (cudaMalloc(&t.d_ftile_ext, sizeof(int4) * (size_t)t.max_ftiles) != cudaSuccess) return 1;
    if (cudaMalloc(&t.d_wtile_ext, sizeof(int4) * (size_t)t.max_wtiles) != cudaSuccess) return 1;
    if (cudaMalloc(&t.d_ftile_rng, sizeof(int2) * 18 * (size_t)t.max_ftiles) != cudaSuccess) return 1;
    if (cudaMalloc(&t.d_wtile_rng, sizeof(int2) * 9 * (size_t)t.max_wtiles) != cudaSuccess) return 1;
    if (cudaMalloc(&t.d_ptile_ext, sizeof(int4) * max_tiles) != cudaSuccess) return 1;
    if (cudaMalloc(&t.d_ptile_rng, sizeof(int2) * 9 * max_tiles) != cudaSuccess) return 1;
    if (cudaMalloc(&t.d_wactive, sizeof(int) * (size_t)t.max_wtiles) != cudaSuccess) return 1;
    if (cudaMalloc(&t.d_n_wactive, sizeof(int) * 4) != cudaSuccess) return 1;
    if (cudaMalloc(&t.d_wtile_state, (size_t)t.max_wtiles + 4) != cudaSuccess) return 1;
    cudaMemset(t.d_wtile_state, 0, (size_t)t.max_wtiles + 4);
    cudaMemset(t.d_frow_tile_start, 0, sizeof(int) * (size_t)(nrows + 4));
    cudaMemset(t.d_wrow_tile_start, 0, sizeof(int) * (size_t)(nrows + 4));
    return 0;
}
// The wall's tile table is built once; rows cut at gaps may need more tiles than the first guess.
inline int tiles_reserve_wall(TileState &t, int n_tiles)
{
    if (n_tiles <= t.max_wtiles) return 0;
    t.max_wtiles = n_tiles;
    cudaFree(t.d_wtile_desc);
    cudaFree(t.d_wtile_ext);
    cudaFree(t.d_wtile_rng);
    cudaFree(t.d_wactive);
    cudaFree(t.d_wtile_state);
    t.d_wtile_desc = nullptr, t.d_wtile_ext = nullptr, t.d_wtile_rng = nullptr, t.d_wactive = nullptr;
    t.d_wtile_state = nullptr;
    if (cudaMalloc(&t.d_wtile_desc, sizeof(int4) * (size_t)t.max_wtiles) != cudaSuccess) return 1;
    if (cudaMalloc(&t.d_wtile_ext, sizeof(int4) * (size_t)t.max_wtiles) != cudaSuccess) return 1;
    if (cudaMalloc(&t.d_wtile_rng, sizeof(int2) * 9 * (size_t)t.max_wtiles) != cudaSuccess) return 1;
    if (cudaMalloc(&t.d_wactive, sizeof(int) * (size_t)t.max_wtiles) != cudaSuccess) return 1;
    if (cudaMalloc(&t.d_wtile_state, (size_t)t.max_wtiles + 4) != cudaSuccess) return 1;
    cudaMemset(t.d_wtile_state, 0, (size_t)t.max_wtiles + 4);
    if (t.max_wtiles > t.max_ftiles) {
        cudaFree(t.d_ptile_ext);
        cudaFree(t.d_ptile_rng);
        t.d_ptile_ext = nullptr, t.d_ptile_rng = nullptr;
        if (cudaMalloc(&t.d_ptile_ext, sizeof(int4) * (size_t)t.max_wtiles) != cudaSuccess) return 1;
        if (cudaMalloc(&t.d_ptile_rng, sizeof(int2) * 9 * (size_t)t.max_wtiles) != cudaSuccess) return 1;
    }
    return 0;
}
inline void tiles_free(TileState &t)
{
    if (t.d_row_tiles) cudaFree(t.d_row_tiles);
    if (t.d_frow_tile_start) cudaFree(t.d_frow_tile_start);
    if (t.d_wrow_tile_start) cudaFree(t.d_wrow_tile_start);
    if (t.d_ftile_desc) cudaFree(t.d_ftile_desc);
    if (t.d_wtile_desc) cudaFree(t.d_wtile_desc);
    if (t.d_ftile_ext) cudaFree(t.d_ftile_ext);
    if (t.d_wtile_ext) cudaFree(t.d_wtile_ext);
    if (t.d_ftile_rng) cudaFree(t.d_ftile_rng);
    if (t.d_wtile_rng) cudaFree(t.d_wtile_rng);
    if (t.d_ptile_ext) cudaFree(t.d_ptile_ext);
    if (t.d_ptile_rng) cudaFree(t.d_ptile_rng);
    if (t.d_wactive) cudaFree(t.d_wactive);
    if (t.d_n_wactive) cudaFree(t.d_n_wactive);
    if (t.d_wtile_state) cudaFree(t.d_wtile_state);
    t = TileState();
}

// records per staged chunk for a shared-memory budget (multiple of 4, 16-bit indexable)
template <typename T, typename CT>
inline int tile_capacity(int smem_budget, int list_len, int ks = 1)
{
    const size_t fixed = TILE_HDR_BYTES + (size_t)list_len * ks * TILE_TB * sizeof(unsigned short) + 256;
    const size_t rec = tile_record_bytes<T, CT>();
    int cap = (size_t)smem_budget > fixed ? (int)(((size_t)smem_budget - fixed) / rec) : 0;
    cap &= ~3;
    if (cap < 32) cap = 32;  // a budget below the fixed part is exceeded rather than refused
    return cap > 65532 ? 65532 : cap;
}

}  // namespace tpb
