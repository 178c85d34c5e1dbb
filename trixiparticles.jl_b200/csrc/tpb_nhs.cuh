// tpb_nhs.cuh -- neighbourhood-search rebuild: the B200 counterpart of
// PointNeighbors.update!(GridNeighborhoodSearch{FullGridCellList}) as called from
// /root/reference/src/general/neighborhood_search.jl:479-506, :804-808.
//
// Instead of per-cell index vectors filled with atomics and then chased through memory, the
// particles themselves are counting-sorted by linear cell index every rebuild, so that each
// neighbour-cell row is one contiguous stream of 16-byte records:
//   k_cell_count   key[i] = cell(u_i);  slot[i] = arrival number inside the cell
//   scan           cell_start = exclusive_scan(count)          (three small kernels)
//   k_scatter      tmp[cell_start[key[i]] + slot[i]] = i
//   k_reorder      final position inside the cell = number of smaller particle indices in
//                  that cell (deterministic, independent of atomic arrival order); gathers
//                  the AoS ODE vectors into the sorted SoA records and applies the EOS.
// All passes are pure streaming work (HBM-bound).
#pragma once
#include "tpb_device.cuh"

namespace tpb {

// ------------------------------------------------------------------ cell keys + histogram
// Order of the particles INSIDE a cell.  Consecutive sorted particles are consecutive lanes of the tile sweeps,
// and lanes whose targets sit next to each other accept nearly the same candidates: their shared-memory reads
// fall into the same 128-byte lines (the sweeps are bound by that pipe).  By previous index alone the order inside
// a cell is whatever the ODE vectors happen to hold -- lattice order at t = 0, arbitrary once the fluid has
// mixed.  So the entries of tmp_perm carry a 6-bit position key above the particle index,
//   entry = sub << pbits | i,   sub = the particle's place among 4 x 4 x 4 sub-boxes of its cell,
// and rank_in_cell orders by (sub, i): still a total order independent of the atomics' arrival order.
// pbits = bits left for the particle index: 25 (6-bit key, 4 x 4 x 4 sub-boxes) for sets of fewer than 2^25
// particles, 26 (5 bits: 4 x 4 x 2) below 2^26, 27 (4 bits: 4 x 4 x 1) below 2^27 -- the library's size limit --
// and 31 (no key) when the key is switched off.
// Measured at 1 M particles (interact! phase; profiles/r2_y_order_inside_cells.txt): lattice order in the ODE vectors
// 0.853 ms with or without the key; a random order 0.887 ms without, 0.854 ms with it.  Sub-boxes in lexicographic
// order with the row axis fastest are the best of those tried: the row axis slowest 0.860, Morton order 0.884 (no
// better than random), serpentine 0.867, 8 x 8 x 1 boxes 0.871, 2 x 2 x 16 0.880, 4 x 2 x 8 0.884, 3 x 3 x 7 0.855.
// TPB_SUBKEY=0 switches the key off.
constexpr int PERM_IDX_BITS = 25;   // smallest index width (largest key); slot[i] = arrival number | key << pbits
__host__ __device__ __forceinline__ int perm_index(int entry, int pbits) { return entry & (int)((1u << pbits) - 1u); }

template <int ND, typename CT>
__device__ __forceinline__ int cell_subkey(const GridConst<CT> &g, CT x, CT y, CT z, int cx, int cy, int cz, int pbits)
{
    if (g.ax == 1) {
        const CT t = x;
        x = y;
        y = t;
    } else if (ND == 3 && g.ax == 2) {
        const CT t = x;
        x = z;
        z = t;
    }
    const float fx = (float)((x - g.origin[0]) * g.inv_cell_x - (CT)cx);
    const float fy = (float)((y - g.origin[1]) * g.inv_cell - (CT)cy);
    const float fz = ND == 3 ? (float)((z - g.origin[2]) * g.inv_cell - (CT)cz) : 0.f;
    const int sx = min(max((int)(fx * 4.f), 0), 3), sy = min(max((int)(fy * 4.f), 0), 3);
    const int sz = min(max((int)(fz * 4.f), 0), 3);
    // lexicographic, the row axis fastest; fewer sub-boxes along the row when the index needs the bits
    if (pbits <= PERM_IDX_BITS) return (sz * 4 + sy) * 4 + sx;
    if (pbits == PERM_IDX_BITS + 1) return (sz * 4 + sy) * 2 + (sx >> 1);
    return sz * 4 + sy;
}

template <int ND, typename CT>
__global__ void __launch_bounds__(256)
k_cell_count(const CT *__restrict__ coords /* ND x n, AoS */, int n, int n_targets, GridConst<CT> g,
             int *__restrict__ key, int *__restrict__ slot, int *__restrict__ count,
             int *__restrict__ flags, int pbits = 31, const CT *__restrict__ tail_coords = nullptr,
             int n_head = 0, CT *__restrict__ out_coords = nullptr)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // structure: the integrated particles come from u_ode, the clamped tail from where it was put
    // (update_positions!, total_lagrangian_sph/system.jl:403-433); the merged coordinates are kept
    const CT *src = tail_coords != nullptr && i >= n_head ? tail_coords : coords;
    CT x = src[(int64_t)i * ND + 0];
    if (i >= n_targets && !(x == x)) {  // empty slab-ghost slot (NaN x): not binned, not an error
        key[i] = -1;
        slot[i] = 0;
        return;
    }
    CT y = src[(int64_t)i * ND + 1];
    CT z = ND == 3 ? src[(int64_t)i * ND + 2] : (CT)0;
    if (out_coords != nullptr) {
        out_coords[(int64_t)i * ND + 0] = x;
        out_coords[(int64_t)i * ND + 1] = y;
        if (ND == 3) out_coords[(int64_t)i * ND + 2] = z;
    }
    int cx, cy, cz;
    if (!cell_coords<ND, CT>(g, x, y, z, cx, cy, cz)) atomicOr(flags, 1);
    int c = cell_linear(g, cx, cy, cz);
    key[i] = c;
    const int sub = pbits < 31 ? cell_subkey<ND, CT>(g, x, y, z, cx, cy, cz, pbits) : 0;
    slot[i] = atomicAdd(&count[c], 1) | (pbits < 31 ? sub << pbits : 0);
}

// ------------------------------------------------------------------ exclusive scan (int32)
// Three-phase scan: per-block sums, scan of block sums (single block), per-block scan + offset.
// SCAN_ITEMS items per thread, 1024 threads per block.
constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int warp_inclusive_scan(int v)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) >= o) v += t;
    }
    return v;
}

// inclusive scan across the block of one value per thread; returns the inclusive value and
// the block total
__device__ __forceinline__ int block_inclusive_scan(int v, int &total)
{
    __shared__ int warp_sums[32];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = warp_inclusive_scan(v);
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int nw = (blockDim.x + 31) >> 5;
        int ws = lane < nw ? warp_sums[lane] : 0;
        ws = warp_inclusive_scan(ws);
        warp_sums[lane] = ws;
    }
    __syncthreads();
    int offset = wid > 0 ? warp_sums[wid - 1] : 0;
    total = warp_sums[((blockDim.x + 31) >> 5) - 1];
    __syncthreads();
    return inc + offset;
}

static __global__ void __launch_bounds__(SCAN_THREADS)
k_scan_block_sums(const int *__restrict__ in, int n, int *__restrict__ block_sums)
{
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) s += in[base + k];
    int total;
    block_inclusive_scan(s, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// single block: exclusive scan of up to SCAN_TILE block sums in place
static __global__ void __launch_bounds__(SCAN_THREADS)
k_scan_top(int *__restrict__ block_sums, int nblocks)
{
    int base = threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = base + k < nblocks ? block_sums[base + k] : 0;
        s += v[k];
    }
    int total;
    int inc = block_inclusive_scan(s, total);
    int run = inc - s;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < nblocks) block_sums[base + k] = run;
        run += v[k];
    }
}

// out[i] = exclusive prefix of in; out[n] = total
static __global__ void __launch_bounds__(SCAN_THREADS)
k_scan_final(const int *__restrict__ in, int n, const int *__restrict__ block_offsets,
             int *__restrict__ out)
{
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = base + k < n ? in[base + k] : 0;
        s += v[k];
    }
    int total;
    int inc = block_inclusive_scan(s, total);
    int run = inc - s + block_offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
        if (base + k == n - 1) out[n] = run;
    }
}

// small grids (up to SCAN_SINGLE_MAX cells: 2-D set-ups, the structure of an FSI case): one block walks the
// histogram tile by tile with a running carry -- one launch instead of three, which is what such a step consists of
constexpr int SCAN_SINGLE_MAX = 2 * SCAN_TILE;
static __global__ void __launch_bounds__(SCAN_THREADS)
k_scan_single(const int *__restrict__ in, int n, int *__restrict__ out)
{
    int carry = 0;
    for (int tile = 0; tile < n; tile += SCAN_TILE) {
        const int base = tile + (int)threadIdx.x * SCAN_ITEMS;
        int v[SCAN_ITEMS];
        int s = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            v[k] = base + k < n ? in[base + k] : 0;
            s += v[k];
        }
        int total;
        const int inc = block_inclusive_scan(s, total);
        int run = carry + inc - s;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            if (base + k < n) out[base + k] = run;
            run += v[k];
        }
        carry += total;
    }
    if (threadIdx.x == 0) out[n] = carry;
}

// ------------------------------------------------------------------ scatter of indices
static __global__ void __launch_bounds__(256)
k_scatter(const int *__restrict__ key, const int *__restrict__ slot,
          const int *__restrict__ cell_start, int n, int *__restrict__ tmp_perm, int pbits = 31)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (key[i] < 0) return;  // empty slab-ghost slot
    const int sl = slot[i];
    if (pbits < 31)  // slot = arrival number | key << pbits: the key moves over to the entry
        tmp_perm[cell_start[key[i]] + (sl & ((1 << pbits) - 1))] = (sl & ~((1 << pbits) - 1)) | i;
    else
        tmp_perm[cell_start[key[i]] + sl] = i;
}

// rank of the entry i (position key, particle index) among the entries of its cell [a, b) in tmp_perm
__device__ __forceinline__ int rank_in_cell(const int *__restrict__ tmp_perm, int a, int b, int i)
{
    int r = 0;
    for (int t = a; t < b; ++t) r += tmp_perm[t] < i;
    return r;
}

// ------------------------------------------------------------------ SortingCallback
// sort_system! (callbacks/sorting.jl:146-157): the rows of the caller's ODE vectors (and the masses, which the
// reference leaves as a TODO for non-uniform particles) in cell order; inside a cell by previous index, so the
// summation order of every later kick is unchanged.
template <int ND, typename T, typename CT>
__global__ void __launch_bounds__(256)
k_sort_gather(const CT *__restrict__ u, const T *__restrict__ v, const T *__restrict__ mass,
              const int *__restrict__ key, const int *__restrict__ cell_start, const int *__restrict__ tmp_perm,
              int n, int nv, CT *__restrict__ out_u, T *__restrict__ out_v, T *__restrict__ out_m, int pbits = 31)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int e = tmp_perm[s];
    const int i = perm_index(e, pbits);
    const int c = key[i];
    const int a = cell_start[c], b = cell_start[c + 1];
    const int dst = a + rank_in_cell(tmp_perm, a, b, e);
#pragma unroll
    for (int d = 0; d < ND; ++d) out_u[(int64_t)dst * ND + d] = u[(int64_t)i * ND + d];
    for (int q = 0; q < nv; ++q) out_v[(int64_t)dst * nv + q] = v[(int64_t)i * nv + q];
    out_m[dst] = mass[i];
}

// ------------------------------------------------------------------ fluid reorder + EOS
// Gathers the AoS ODE vectors (u: ND x n cT, v: NV x n T) into the sorted records.
// DENS == 0 (ContinuityDensity): rho = last row of v, pressure = EOS(rho) here.
// DENS == 1 (SummationDensity): rho/pressure are filled by the density sweep afterwards.
template <int ND, typename T, typename CT, int DENS>
__global__ void __launch_bounds__(256)
k_reorder_fluid(const CT *__restrict__ u, const T *__restrict__ v, const T *__restrict__ mass,
                const int *__restrict__ key, const int *__restrict__ cell_start,
                const int *__restrict__ tmp_perm, int n, const int *__restrict__ n_sorted,
                int deterministic, EosConst<T> eos, V4<CT> *__restrict__ A, V4<T> *__restrict__ B,
                T *__restrict__ P, int *__restrict__ perm, FilterRef<CT> fref, V4<float> *__restrict__ F,
                const AdaptConsts<T> *__restrict__ ad = nullptr, int pbits = 31)
{
    constexpr int NV = DENS == 0 ? ND + 1 : ND;
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n || s >= *n_sorted) return;  // n_sorted < n when ghost slots are empty
    if (ad) eos.B = ad->B_f;  // StateEquationAdaptiveCole: this kick's speed of sound (k_adaptive_consts)
    const int e = tmp_perm[s];
    const int i = perm_index(e, pbits);
    int dst = s;
    if (deterministic) {
        int c = key[i];
        int a = cell_start[c], b = cell_start[c + 1];
        dst = a + rank_in_cell(tmp_perm, a, b, e);
    }
    V4<CT> ra;
    ra.x = u[(int64_t)i * ND + 0];
    ra.y = u[(int64_t)i * ND + 1];
    ra.z = ND == 3 ? u[(int64_t)i * ND + 2] : (CT)0;
    ra.w = (CT)mass[i];
    V4<T> rb;
    rb.x = v[(int64_t)i * NV + 0];
    rb.y = v[(int64_t)i * NV + 1];
    rb.z = ND == 3 ? v[(int64_t)i * NV + 2] : (T)0;
    if (DENS == 0) {
        T rho = v[(int64_t)i * NV + ND];
        rb.w = rho;
        P[dst] = eos_pressure(eos, rho);
    } else {
        rb.w = (T)0;
    }
    A[dst] = ra;
    B[dst] = rb;
    perm[dst] = i;
    if (F) F[dst] = filter_position<CT>(fref, ra);
}

// ------------------------------------------------------------------ wall reorder (once)
template <int ND, typename T, typename CT>
__global__ void __launch_bounds__(256)
k_reorder_wall(const CT *__restrict__ coords, const T *__restrict__ mass,
               const T *__restrict__ density0, const int *__restrict__ key,
               const int *__restrict__ cell_start, const int *__restrict__ tmp_perm, int n,
               V4<CT> *__restrict__ A, V2<T> *__restrict__ W, int *__restrict__ perm, FilterRef<CT> fref,
               V4<float> *__restrict__ F, int pbits = 31)
{
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int e = tmp_perm[s];
    const int i = perm_index(e, pbits);
    int c = key[i];
    int a = cell_start[c], b = cell_start[c + 1];
    int dst = a + rank_in_cell(tmp_perm, a, b, e);
    V4<CT> ra;
    ra.x = coords[(int64_t)i * ND + 0];
    ra.y = coords[(int64_t)i * ND + 1];
    ra.z = ND == 3 ? coords[(int64_t)i * ND + 2] : (CT)0;
    ra.w = (CT)mass[i];
    A[dst] = ra;
    if (F) F[dst] = filter_position<CT>(fref, ra);
    V2<T> w;
    w.x = (T)0;          // pressure (initial_boundary_pressure: zero for Adami)
    w.y = density0[i];   // cache.density = copy(initial_density) (dummy_particles.jl:290-297)
    W[dst] = w;
    perm[dst] = i;
}

}  // namespace tpb
