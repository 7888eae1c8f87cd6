// Translation unit: TMA-streamed backward row kernels (pd_warp_stream.cuh, rows_bwd_stream instantiations).
#define PD_TS_BWD_ONLY
#include "pd_warp_stream.cuh"

namespace pd {
namespace api {
// dry run of the launcher: true when stream_bwd() will find a configuration, so the caller may skip the zero-fill the
// scatter kernels need
bool stream_bwd_fits(const WarpParams& p) { return ts::launch_bwd_stream(p, nullptr, true); }
bool stream_bwd(const WarpParams& p, cudaStream_t st) { return ts::launch_bwd_stream(p, st); }
}  // namespace api
}  // namespace pd
