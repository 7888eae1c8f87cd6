// Translation unit: homography-warp fast path (pd_warp_homo.cuh).
#include "pd_warp_homo.cuh"

namespace pd {
namespace api {
bool homo_supported(const WarpParams& p) { return hm::homo_path_supported(p); }
size_t homo_workspace_bytes(const pd_warp_desc* d) {
    if (d->warp_type != PD_WARP_HOMOGRAPHY || ((int64_t)d->H * d->W) % hm::HT != 0 || d->W % 32 != 0) return 0;
    return hm::homo_workspace_bytes(d);
}
int homo_fwd(const WarpParams& p, void* workspace, cudaStream_t st) {
    if (!workspace) return fail(PD_ERR_WORKSPACE, "homography warp needs the workspace of pd_warp_composite_workspace_bytes()");
    if (!(p.d.flags & PD_FLAG_WORKSPACE_READY)) {
        hm::homo_pack(p, (float4*)workspace, st);
        int rc = check_launch("pack_rgbx");
        if (rc) return rc;
    }
    hm::launch_homo_fwd(p, (const float4*)workspace, st);
    return check_launch("homo_fwd");
}
int homo_bwd(const WarpParams& p, void* workspace, cudaStream_t st) {
    if (!workspace) return fail(PD_ERR_WORKSPACE, "homography warp needs the workspace of pd_warp_composite_workspace_bytes()");
    if (!(p.d.flags & PD_FLAG_WORKSPACE_READY)) {
        hm::homo_pack(p, (float4*)workspace, st);
        int rc = check_launch("pack_rgbx");
        if (rc) return rc;
    }
    hm::launch_homo_bwd(p, (const float4*)workspace, st);
    return check_launch("homo_bwd");
}
}  // namespace api
}  // namespace pd
