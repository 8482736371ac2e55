// reorder_rows.cuh -- internal interface of the row-reordering stage (see reorder_rows.cu).
#pragma once
#include "common.cuh"

namespace sb {

struct RowReorderStats {
  u32 nbpr = 0, blockDim = 0, keptMask = 0, zeroRows = 0, exactEvals = 0, encEntries = 0, clustersCreated = 0;
};

u32 calc_block_size(u32 M, u32 N, u64 freeMem);
u32 num_blocks_per_row(u32 N, u32 bs);
u32 cluster_blockdim(u32 nbpr);
u32 kept_warp_mask(u32 B);

void dispersion_dev(const u32* d_rowOff, const u32* d_colIdx, u32 M, u32 N, u32 nnz, u32 bs, u32* d_disp,
                    u32* nbprOut, cudaStream_t s);

void row_reorder_dev(const u32* d_rowOff, const u32* d_colIdx, u32 M, u32 N, u32 nnz, float alpha, u32 bs,
                     const bsmr_reorder_opts* opts, u32* d_reorderedRows, u32* numRows, int32_t* numClusters, RowReorderStats* st, cudaStream_t s);

}  // namespace sb
