// eval_stats.cuh -- device versions of the reference's reordering statistics (src/BSMR.cpp:826-994).
#pragma once
#include "common.cuh"

namespace sb {

void original_block_stats_dev(const u32* d_rowOff, const u32* d_colIdx, u32 M, u32 N, u32 nnz, float delta,
                              u32* numDenseBlocks, float* averageDensity, cudaStream_t s);
void layout_eval(const bsmr_layout* L, float delta, bsmr_eval* out, cudaStream_t s);

}  // namespace sb
