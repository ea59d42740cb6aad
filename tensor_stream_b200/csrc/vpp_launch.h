// vpp_launch.h -- host-callable entry points of vpp_kernels.cu
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

#include "vpp_plan.h"

namespace tsvpp {

int fused_block_threads(const PlanDev &P);
size_t fused_smem_bytes(const PlanDev &P);
bool fused_supports_bw4(int rs);
int configure_fused(int rs, int ok, bool bulk, int bw, size_t smem_optin_max);
int max_ctas_per_sm(int rs, int ok, bool bulk, int bw, int threads, size_t smem_bytes);
int launch_fused(int rs, int ok, bool bulk, const PlanDev &P, const FrameDev *frames_dev, int nframes, int vec_ok,
                 int grid, size_t smem_bytes, cudaStream_t stream);
int launch_direct(int ok, const PlanDev &P, const FrameDev *frames_dev, int nframes, int vec_ok, cudaStream_t stream);
int launch_packed_yuv(int yuv444, int is_float, const FrameDev *frames_dev, int nframes, int w, int h, int pitch,
                      long y_bytes, long uv_bytes, int norm, int fast, cudaStream_t stream);

// AREA downscale whose tap window does not fit the fused kernel's stages: resize stage alone, into the tight
// NV12 intermediate (out = Y plane dw*dh, then the interleaved UV rows of dw bytes)
struct AreaWideParams {
    int dw, dh;           // output size
    int sw, sh;           // working (post-crop) source size
    int crop, x0, y0;     // crop origin in the source frame
    int pitch_y, pitch_uv;
    long plane_y, plane_uv;  // bytes addressable in each source plane
    float xr, yr;
    int KX, K, PX, PY;    // window and pattern sizes
    const float *patx, *paty;
};
int launch_area_wide(const AreaWideParams &P, const FrameDev *frames_dev, int nframes, cudaStream_t stream);
void launch_pow23(const double *w_dev, double *w2_dev, double *w3_dev, int n, cudaStream_t stream);
void launch_bicubic_weights(const double *w_dev, double *c_dev, int n, cudaStream_t stream);

}  // namespace tsvpp
