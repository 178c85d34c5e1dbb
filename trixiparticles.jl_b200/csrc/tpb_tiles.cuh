// tpb_tiles.cuh -- neighbour sweeps, variant 2 ("cell tiles"): placeholder until the tiled
// kernel lands; reports itself unsupported so the per-particle sweep is used.
#pragma once
#include "tpb_device.cuh"
#include "tpb_sweeps.cuh"

namespace tpb {

struct TileState {
    int dummy = 0;
};

inline int tiles_alloc(TileState &, int64_t, int64_t) { return 0; }
inline void tiles_free(TileState &) {}

template <int ND, typename T, typename CT>
constexpr bool tiles_supported()
{
    return false;
}

template <int ND, typename T, typename CT, int KERNEL, int DENS>
int launch_interact_tiles(TileState &, cudaStream_t, int, const GridConst<CT> &, const int *,
                          const V4<CT> *, const V4<T> *, const T *, const int *, int, int,
                          const int *, const V4<CT> *, const V2<T> *, const PairConst<T> &,
                          const SourceConst<T> &, T *, int &, int64_t &)
{
    return 2;  // TPB_ERR_UNSUPPORTED
}

}  // namespace tpb
