// Tensor-path PSNR + SSIM (metrics_mma.cu), called from run_psnr_ssim (metrics.cu).
#pragma once
#include <cuda_runtime.h>

namespace srb {

bool psnr_ssim_mma_eligible(const float* a, const float* b, int height, int width, int channels, float max_val);
// accumulates {squared error, SSIM-map sum} per image into acc[B][2] (zeroed by the caller)
int run_psnr_ssim_mma(const float* a, const float* b, int batch, int height, int width, int channels, float c1, float c2,
                      double* acc, cudaStream_t stream);

}  // namespace srb
