// Row-tiled fast path for stereo disparity warps (placeholder until the kernels land).
#pragma once
#include "pd_warp_general.cuh"
namespace pd {
inline bool rows_path_supported(const WarpParams&) { return false; }
inline void launch_fwd_rows(const WarpParams&, cudaStream_t) {}
inline void launch_bwd_rows(const WarpParams&, cudaStream_t) {}
}  // namespace pd
