// sddmm_kernels.cuh -- launcher of the dense (tcgen05) and residual (CUDA-core) SDDMM kernels.
#pragma once
#include "common.cuh"

struct bsmr_layout;

namespace sb {
// one pass: dense blocks on `denseStream`, residual on `sparseStream` (may be the same stream)
enum { kLaunchDense = 1, kLaunchSparse = 2, kLaunchBoth = 3 };
void sddmm_launch(const bsmr_layout* L, u32 K, const float* dA, const float* dB, float* dP,
                  cudaStream_t denseStream, cudaStream_t sparseStream, int which = kLaunchBoth, u32 numBatch = 1);
}  // namespace sb
