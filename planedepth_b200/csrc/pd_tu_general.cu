// Translation unit: reference-arithmetic kernels for any warp type / stride pattern (pd_warp_general.cuh).
#include "pd_warp_general.cuh"

namespace pd {
namespace {
template <int WARP, bool MIX>
void launch_fwd(const WarpParams& p, bool debug, cudaStream_t st) {
    const int64_t total = (int64_t)p.d.B * p.hw;
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (debug) warp_composite_fwd_general<WARP, MIX, true><<<grid, 256, 0, st>>>(p);
    else warp_composite_fwd_general<WARP, MIX, false><<<grid, 256, 0, st>>>(p);
}
template <int WARP, bool MIX>
void launch_bwd(const WarpParams& p, cudaStream_t st) {
    const int64_t total = (int64_t)p.d.B * p.hw;
    warp_composite_bwd_general<WARP, MIX><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(p);
}
}  // namespace

namespace api {
void general_fwd(const WarpParams& p, bool debug, cudaStream_t st) {
    const bool mix = p.d.mixture != 0;
    switch (p.d.warp_type) {
        case PD_WARP_DISP: mix ? launch_fwd<PD_WARP_DISP, true>(p, debug, st) : launch_fwd<PD_WARP_DISP, false>(p, debug, st); break;
        case PD_WARP_HOMOGRAPHY: mix ? launch_fwd<PD_WARP_HOMOGRAPHY, true>(p, debug, st) : launch_fwd<PD_WARP_HOMOGRAPHY, false>(p, debug, st); break;
        default: mix ? launch_fwd<PD_WARP_DEPTH, true>(p, debug, st) : launch_fwd<PD_WARP_DEPTH, false>(p, debug, st); break;
    }
}
void general_bwd(const WarpParams& p, cudaStream_t st) {
    const bool mix = p.d.mixture != 0;
    switch (p.d.warp_type) {
        case PD_WARP_DISP: mix ? launch_bwd<PD_WARP_DISP, true>(p, st) : launch_bwd<PD_WARP_DISP, false>(p, st); break;
        case PD_WARP_HOMOGRAPHY: mix ? launch_bwd<PD_WARP_HOMOGRAPHY, true>(p, st) : launch_bwd<PD_WARP_HOMOGRAPHY, false>(p, st); break;
        default: mix ? launch_bwd<PD_WARP_DEPTH, true>(p, st) : launch_bwd<PD_WARP_DEPTH, false>(p, st); break;
    }
}
}  // namespace api
}  // namespace pd
