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
void launch_pow23(const double *w_dev, double *w2_dev, double *w3_dev, int n, cudaStream_t stream);
void launch_bicubic_weights(const double *w_dev, double *c_dev, int n, cudaStream_t stream);

}  // namespace tsvpp
