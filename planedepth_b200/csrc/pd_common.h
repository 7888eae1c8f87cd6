// Host-side glue shared by the translation units of libplanedepth_b200.so.
//
// The library is built from several .cu files (one per kernel family, compiled in parallel by
// planedepth_b200/_lib.py); pd_abi.cu owns the error string, the launch counter and the tuning block,
// the family files reach them through the functions declared here.  Nothing here crosses the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/planedepth_b200.h"

namespace pd {

struct WarpParams;  // pd_warp_general.cuh

int fail(int code, const char* fmt, ...);  // records the thread-local message of pd_last_error(), returns code
int check_launch(const char* what);        // cudaGetLastError() -> pd_status; counts the launch
int check_device();                        // PD_ERR_ARCH unless the current device is sm_100 (cached per device)
int sm_count();                            // multiprocessors of the current device (cached per device)
const pd_tuning& tuning();                 // process-wide block behind pd_set_tuning / pd_get_tuning

// Opt a kernel in to > 48 KB of dynamic shared memory once per (kernel, size): steady-state launches issue no
// attribute call (those cannot be captured into a CUDA graph).
void smem_optin(const void* kernel, size_t smem);
// cudaOccupancyMaxActiveBlocksPerMultiprocessor, cached per (kernel, threads, smem)
int resident_ctas(const void* kernel, int threads, size_t smem);

namespace api {
// streamed row kernels (pd_tu_stream_fwd.cu / pd_tu_stream_bwd.cu)
bool stream_supported(const WarpParams& p);
bool stream_fwd_fits(const WarpParams& p);
bool stream_bwd_fits(const WarpParams& p);
bool stream_fwd(const WarpParams& p, cudaStream_t st);  // false = no launch configuration
bool stream_bwd(const WarpParams& p, cudaStream_t st);
// bit-faithful row kernels (pd_tu_rows.cu)
bool rows_supported(const WarpParams& p);
void rows_fwd(const WarpParams& p, cudaStream_t st);
void rows_bwd(const WarpParams& p, cudaStream_t st);
// homography fast path (pd_tu_homo.cu)
bool homo_supported(const WarpParams& p);
size_t homo_workspace_bytes(const pd_warp_desc* d);
int homo_fwd(const WarpParams& p, void* workspace, cudaStream_t st);  // pd_status
int homo_bwd(const WarpParams& p, void* workspace, cudaStream_t st);
// reference-arithmetic kernels, any warp type / strides (pd_tu_general.cu)
void general_fwd(const WarpParams& p, bool debug, cudaStream_t st);
void general_bwd(const WarpParams& p, cudaStream_t st);
}  // namespace api

}  // namespace pd
