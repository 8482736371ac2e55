// eval_stats.cu -- the block statistics of evaluationReordering (src/BSMR.cpp:826-925) and of
// calculateNumDenseBlocksAndAverageDensityInOriginalMatrix (src/BSMR.cpp:953-994) on the device.
// The reference walks every (row panel, column block) pair on the host, O(panels * N/16 * nnz_panel); here
// the original-matrix statistic is one key sort of (row/16, col/16) pairs plus a run-length pass, and the
// reordered-matrix statistic is a reduction over blockValues.
#include "common.cuh"
#include "eval_stats.cuh"
#include "layout.cuh"
#include "primitives.cuh"

namespace sb {

namespace {

// one warp per row: key = (row / 16) * numColBlocks + col / 16
__global__ void __launch_bounds__(256) k_block_keys(const u32* __restrict__ rowOff, const u32* __restrict__ colIdx,
                                                    u32 M, u32 numColBlocks, u64* __restrict__ keys) {
  const u32 lane = threadIdx.x & 31u;
  const u32 warpsPerGrid = (gridDim.x * blockDim.x) >> 5;
  for (u32 row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < M; row += warpsPerGrid) {
    const u32 b = rowOff[row], e = rowOff[row + 1];
    const u64 base = (u64)(row >> 4) * numColBlocks;
    for (u32 i = b + lane; i < e; i += 32u) keys[i] = base + (colIdx[i] >> 4);
  }
}

struct Accum {
  unsigned long long blocks;  // blocks with density >= delta
  double density;             // sum of their densities
  unsigned long long nonEmpty;
  double densityAll;          // sum over every non-empty block
};

// a run of equal sorted keys is one 16x16 block of the original matrix; runs are at most 256 long
__global__ void __launch_bounds__(256) k_block_runs(const u64* __restrict__ keys, u32 n, u32 M, u32 N,
                                                    u32 numColBlocks, float delta, Accum* __restrict__ acc) {
  unsigned long long nb = 0;
  double sum = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const u64 k = keys[i];
    if (i && keys[i - 1] == k) continue;
    u32 cnt = 1;
    while (i + cnt < n && keys[i + cnt] == k) ++cnt;
    const u32 rp = (u32)(k / numColBlocks), cb = (u32)(k - (u64)rp * numColBlocks);
    const u32 rows = min(16u, M - rp * 16u), cols = min(16u, N - cb * 16u);
    const float density = __fdiv_rn((float)cnt, (float)(rows * cols));  // BSMR.cpp:980
    if (density >= delta) {
      ++nb;
      sum += (double)density;
    }
  }
  // block reduction, then one atomic pair per CTA
  __shared__ unsigned long long sNb[8];
  __shared__ double sSum[8];
  for (int o = 16; o; o >>= 1) {
    nb += __shfl_down_sync(0xffffffffu, nb, o);
    sum += __shfl_down_sync(0xffffffffu, sum, o);
  }
  if ((threadIdx.x & 31u) == 0) { sNb[threadIdx.x >> 5] = nb; sSum[threadIdx.x >> 5] = sum; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { nb += sNb[w]; sum += sSum[w]; }
    if (nb) {
      atomicAdd(&acc->blocks, nb);
      atomicAdd(&acc->density, sum);
    }
  }
}

// one warp per dense block: stored entries = non-NULL slots of its 256 blockValues (BSMR.cpp:887-895, 904-915)
__global__ void __launch_bounds__(256) k_dense_block_density(const u32* __restrict__ blockValues, u32 numBlocks,
                                                             float delta, Accum* __restrict__ acc) {
  const u32 lane = threadIdx.x & 31u;
  const u32 warpsPerGrid = (gridDim.x * blockDim.x) >> 5;
  unsigned long long nb = 0, ne = 0;
  double sumAll = 0.0;
  for (u32 blk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; blk < numBlocks; blk += warpsPerGrid) {
    u32 cnt = 0;
#pragma unroll
    for (u32 j = 0; j < 8; ++j) cnt += blockValues[(size_t)blk * 256u + j * 32u + lane] != kNull;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0 && cnt) {
      const float density = __fdiv_rn((float)cnt, 256.0f);
      ++ne;
      sumAll += (double)density;
      if (density >= delta) ++nb;
    }
  }
  if (lane == 0 && ne) {
    atomicAdd(&acc->nonEmpty, ne);
    atomicAdd(&acc->densityAll, sumAll);
    if (nb) atomicAdd(&acc->blocks, nb);
  }
}

}  // namespace

void original_block_stats_dev(const u32* d_rowOff, const u32* d_colIdx, u32 M, u32 N, u32 nnz, float delta,
                              u32* numDenseBlocks, float* averageDensity, cudaStream_t s) {
  *numDenseBlocks = 0;
  *averageDensity = 0.f;
  if (!nnz || !M || !N) return;
  TempScope scope(s);
  const u32 nCB = ceil_div(N, 16u), nRP = ceil_div(M, 16u);
  DevBuf<u64> keysA(nnz), keysB(nnz);
  DevBuf<Accum> acc(1);
  SB_CUDA(cudaMemsetAsync(acc.get(), 0, sizeof(Accum), s));
  k_block_keys<<<grid_for((size_t)M * 32), 256, 0, s>>>(d_rowOff, d_colIdx, M, nCB, keysA.get());
  SB_LAUNCH_CHECK();
  const int which = radix_sort_pairs<u64>(keysA.get(), keysB.get(), nullptr, nullptr, nnz, 0,
                                          bits_for((u64)nRP * nCB - 1), s);
  k_block_runs<<<grid_for(nnz), 256, 0, s>>>(which ? keysB.get() : keysA.get(), nnz, M, N, nCB, delta, acc.get());
  SB_LAUNCH_CHECK();
  Accum h;
  SB_CUDA(cudaMemcpyAsync(&h, acc.get(), sizeof h, cudaMemcpyDeviceToHost, s));
  SB_CUDA(cudaStreamSynchronize(s));
  *numDenseBlocks = (u32)h.blocks;
  *averageDensity = h.blocks ? (float)(h.density / (double)h.blocks) : 0.f;  // BSMR.cpp:991
}

void layout_eval(const bsmr_layout* L, float delta, bsmr_eval* out, cudaStream_t s) {
  const bsmr_layout_info& I = L->info;
  memset(out, 0, sizeof *out);
  out->numDenseThreadBlocks = I.numDenseThreadBlocks;
  out->numSparseThreadBlocks = I.numSparseThreadBlocks;
  out->numSparseData = I.numSparseValues;
  out->numDenseData = I.nnz - I.numSparseValues;  // BSMR.cpp:924: everything that is not residual
  if (I.numDenseBlocks) {
    TempScope scope(s);
    DevBuf<Accum> acc(1);
    SB_CUDA(cudaMemsetAsync(acc.get(), 0, sizeof(Accum), s));
    k_dense_block_density<<<grid_for((size_t)I.numDenseBlocks * 32), 256, 0, s>>>(L->arr[RPHM_BLOCK_VALUES].get(),
                                                                                  I.numDenseBlocks, delta, acc.get());
    SB_LAUNCH_CHECK();
    Accum h;
    SB_CUDA(cudaMemcpyAsync(&h, acc.get(), sizeof h, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaStreamSynchronize(s));
    out->numDenseBlock = (u32)h.blocks;
    // the reference divides the density summed over ALL non-empty blocks by the number of blocks >= delta
    // (BSMR.cpp:917: `totalDensity / numDenseBlocks > 0 ? ... : 0`)
    const double q = h.blocks ? h.densityAll / (double)h.blocks : 0.0;
    out->averageDensity = q > 0.0 ? (float)q : 0.f;
  }
}

}  // namespace sb
