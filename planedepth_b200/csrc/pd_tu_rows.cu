// Translation unit: bit-faithful row-tiled kernels (pd_warp_rows.cuh; PD_FLAG_EXACT_COORDS).
#include "pd_warp_rows.cuh"

namespace pd {
namespace api {
bool rows_supported(const WarpParams& p) { return rows_path_supported(p); }
void rows_fwd(const WarpParams& p, cudaStream_t st) { launch_fwd_rows(p, st); }
void rows_bwd(const WarpParams& p, cudaStream_t st) { launch_bwd_rows(p, st); }
}  // namespace api
}  // namespace pd
