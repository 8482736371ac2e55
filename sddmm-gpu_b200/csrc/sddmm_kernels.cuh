// sddmm_kernels.cuh -- plan resolution and launcher of the dense (tcgen05) and residual (CUDA-core) SDDMM kernels.
#pragma once
#include "common.cuh"

struct bsmr_layout;

namespace sb {
// defaults from the SDDMM_B200_* environment variables (AUTO where unset)
void plan_default(sddmm_plan* out);
// resolves every AUTO of `in` (nullptr = defaults) for this layout / K / batch count; throws on impossible choices
void plan_resolve(const bsmr_layout* L, u32 K, u32 numBatch, const sddmm_plan* in, sddmm_plan* out);
// builds the K-dependent private layouts / workspaces a RESOLVED plan needs (allocations, sorts, one sync)
void plan_prepare(const bsmr_layout* L, u32 K, u32 numBatch, const sddmm_plan& resolved, cudaStream_t s);
// one pass: dense blocks on `denseStream`, residual on `sparseStream` (may be the same stream)
enum { kLaunchDense = 1, kLaunchSparse = 2, kLaunchBoth = 3 };
void sddmm_launch(const bsmr_layout* L, u32 K, const float* dA, const float* dB, float* dP,
                  cudaStream_t denseStream, cudaStream_t sparseStream, int which = kLaunchBoth, u32 numBatch = 1,
                  const sddmm_plan* plan = nullptr);
}  // namespace sb
