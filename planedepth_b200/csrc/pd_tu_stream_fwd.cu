// Translation unit: TMA-streamed forward row kernels (pd_warp_stream.cuh, rows_fwd_stream instantiations).
#define PD_TS_FWD_ONLY
#include "pd_warp_stream.cuh"

namespace pd {
namespace api {
bool stream_supported(const WarpParams& p) { return ts::stream_path_supported(p); }
bool stream_fwd_fits(const WarpParams& p) { return ts::launch_fwd_stream(p, nullptr, true); }  // dry run of the launcher
bool stream_fwd(const WarpParams& p, cudaStream_t st) { return ts::launch_fwd_stream(p, st); }
}  // namespace api
}  // namespace pd
