// primitives.cuh -- CUB-free device primitives used by the reordering / layout stages:
//   exclusive scan (u32), stable LSD radix sort (u32/u64 keys + u32 payload), small helpers.
// All are plain HBM-bound integer kernels: coalesced loads, shared-memory staging, warp
// match/ballot/popc ranking; grids sized from the data (tiles), not from tensor-core shapes.
#pragma once

#include "common.cuh"

namespace sb {

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void k_fill(T* p, size_t n, T v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
template <typename T>
__global__ void k_iota(T* p, size_t n, T start) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    p[i] = start + (T)i;
}

inline int grid_for(size_t n, int threads = 256, int maxBlocks = 148 * 16) {
  size_t b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > (size_t)maxBlocks) b = maxBlocks;
  return (int)b;
}

template <typename T>
inline void fill(T* p, size_t n, T v, cudaStream_t s) {
  if (!n) return;
  k_fill<T><<<grid_for(n), 256, 0, s>>>(p, n, v);
  SB_LAUNCH_CHECK();
}
template <typename T>
inline void iota(T* p, size_t n, T start, cudaStream_t s) {
  if (!n) return;
  k_iota<T><<<grid_for(n), 256, 0, s>>>(p, n, start);
  SB_LAUNCH_CHECK();
}

// ------------------------------------------------------------------------------------------
// exclusive scan, u32, n up to 2^32-1.  Three-phase (tile sums -> recursive scan -> apply).
// out may alias in.
// ------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ u32 warp_incl_scan(u32 v) {
  const unsigned lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    u32 t = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= (unsigned)d) v += t;
  }
  return v;
}

// block-wide exclusive scan of one value per thread (blockDim.x == kScanThreads); returns the
// exclusive prefix, *total gets the block sum.
__device__ __forceinline__ u32 block_excl_scan(u32 v, u32* total) {
  __shared__ u32 warpSums[kScanThreads / 32];
  __shared__ u32 blockTotal;
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 incl = warp_incl_scan(v);
  if (lane == 31) warpSums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    u32 w = lane < kScanThreads / 32 ? warpSums[lane] : 0;
    const u32 wi = warp_incl_scan(w);
    if (lane < kScanThreads / 32) warpSums[lane] = wi - w;
    if (lane == kScanThreads / 32 - 1) blockTotal = wi;
  }
  __syncthreads();
  const u32 r = warpSums[warp] + incl - v;
  *total = blockTotal;
  __syncthreads();
  return r;
}

static __global__ void __launch_bounds__(kScanThreads) k_scan_tile_sums(const u32* __restrict__ in, size_t n,
                                                                 u32* __restrict__ tileSums) {
  const size_t base = (size_t)blockIdx.x * kScanTile;
  u32 s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    const size_t idx = base + (size_t)i * kScanThreads + threadIdx.x;
    if (idx < n) s += in[idx];
  }
  u32 total;
  block_excl_scan(s, &total);
  if (threadIdx.x == 0) tileSums[blockIdx.x] = total;
}

static __global__ void __launch_bounds__(kScanThreads) k_scan_apply(const u32* in, u32* out, size_t n,
                                                             const u32* __restrict__ tileOffsets) {
  const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * kScanItems;
  u32 v[kScanItems];
  u32 s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = (base + i < n) ? in[base + i] : 0;
    s += v[i];
  }
  u32 total;
  u32 run = block_excl_scan(s, &total) + (tileOffsets ? tileOffsets[blockIdx.x] : 0u);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < n) out[base + i] = run;
    run += v[i];
  }
}

inline void exclusive_scan_u32(const u32* in, u32* out, size_t n, cudaStream_t s) {
  if (n == 0) return;
  const size_t tiles = (n + kScanTile - 1) / kScanTile;
  if (tiles == 1) {
    k_scan_apply<<<1, kScanThreads, 0, s>>>(in, out, n, nullptr);
    SB_LAUNCH_CHECK();
    return;
  }
  DevBuf<u32> sums(tiles);
  k_scan_tile_sums<<<(unsigned)tiles, kScanThreads, 0, s>>>(in, n, sums.get());
  SB_LAUNCH_CHECK();
  exclusive_scan_u32(sums.get(), sums.get(), tiles, s);
  k_scan_apply<<<(unsigned)tiles, kScanThreads, 0, s>>>(in, out, n, sums.get());
  SB_LAUNCH_CHECK();
  if (!g_temp.active) SB_CUDA(cudaStreamSynchronize(s));  // sums is freed on return (stream-ordered inside a TempScope)
}

// ------------------------------------------------------------------------------------------
// stable LSD radix sort, 8 bits per pass.
//   tile = 8 warps x 16 chunks x 32 keys; every warp walks its 512 consecutive keys in order and
//   ranks them with __match_any_sync / popc, which keeps equal digits in input order (stability
//   is what pins the reference's thrust::stable sorts).
// ------------------------------------------------------------------------------------------
constexpr int kSortWarps = 8;
constexpr int kSortChunks = 16;
constexpr int kSortThreads = kSortWarps * 32;
constexpr int kSortTile = kSortThreads * kSortChunks;

template <typename KeyT>
__global__ void __launch_bounds__(kSortThreads) k_radix_hist(const KeyT* __restrict__ keys, size_t n, int shift,
                                                             u32* __restrict__ blockHist, u32 numTiles) {
  __shared__ u32 hist[256];
  hist[threadIdx.x] = 0;
  __syncthreads();
  const size_t base = (size_t)blockIdx.x * kSortTile;
#pragma unroll 4
  for (int c = 0; c < kSortChunks; ++c) {
    const size_t idx = base + (size_t)c * kSortThreads + threadIdx.x;
    if (idx < n) atomicAdd(&hist[(u32)(keys[idx] >> shift) & 255u], 1u);
  }
  __syncthreads();
  blockHist[(size_t)threadIdx.x * numTiles + blockIdx.x] = hist[threadIdx.x];
}

template <typename KeyT, bool kHasVals>
__global__ void __launch_bounds__(kSortThreads) k_radix_scatter(const KeyT* __restrict__ keysIn,
                                                                const u32* __restrict__ valsIn,
                                                                KeyT* __restrict__ keysOut,
                                                                u32* __restrict__ valsOut, size_t n, int shift,
                                                                const u32* __restrict__ blockBase, u32 numTiles) {
  __shared__ u32 warpCnt[kSortWarps][256];
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int w = 0; w < kSortWarps; ++w) warpCnt[w][threadIdx.x] = 0;
  __syncthreads();
  const size_t wbase = (size_t)blockIdx.x * kSortTile + (size_t)warp * (kSortChunks * 32);
  KeyT k[kSortChunks];
  // phase 1: per-warp digit counts
#pragma unroll
  for (int c = 0; c < kSortChunks; ++c) {
    const size_t idx = wbase + (size_t)c * 32 + lane;
    const bool valid = idx < n;
    k[c] = valid ? keysIn[idx] : (KeyT)0;
    const u32 d = (u32)(k[c] >> shift) & 255u;
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      const unsigned peers = __match_any_sync(vmask, d);
      if ((peers & ((1u << lane) - 1u)) == 0) warpCnt[warp][d] += __popc(peers);
    }
    __syncwarp();
  }
  __syncthreads();
  // phase 2: digit-wise exclusive prefix over warps, seeded with this tile's global base
  {
    const u32 d = threadIdx.x;
    u32 run = blockBase[(size_t)d * numTiles + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const u32 t = warpCnt[w][d];
      warpCnt[w][d] = run;
      run += t;
    }
  }
  __syncthreads();
  // phase 3: ranked scatter, chunk by chunk in input order
#pragma unroll
  for (int c = 0; c < kSortChunks; ++c) {
    const size_t idx = wbase + (size_t)c * 32 + lane;
    const bool valid = idx < n;
    const u32 d = (u32)(k[c] >> shift) & 255u;
    const unsigned vmask = __ballot_sync(0xffffffffu, valid);
    if (valid) {
      const unsigned peers = __match_any_sync(vmask, d);
      const u32 rank = __popc(peers & ((1u << lane) - 1u));
      const u32 pos = warpCnt[warp][d] + rank;
      keysOut[pos] = k[c];
      if (kHasVals) valsOut[pos] = valsIn[idx];
    }
    __syncwarp();
    if (valid) {
      const unsigned peers = __match_any_sync(vmask, d);
      if ((peers & ((1u << lane) - 1u)) == 0) warpCnt[warp][d] += __popc(peers);
    }
    __syncwarp();
  }
}

// Sorts n (key, val) pairs by bits [beginBit, endBit) of the key, stable.  Buffers ping-pong;
// returns 0 if the result is in (keysA, valsA), 1 if in (keysB, valsB).  vals may be null.
template <typename KeyT>
inline int radix_sort_pairs(KeyT* keysA, KeyT* keysB, u32* valsA, u32* valsB, size_t n, int beginBit, int endBit,
                            cudaStream_t s) {
  if (n == 0 || endBit <= beginBit) return 0;
  if (n >= ((size_t)1 << 32)) fail(SDDMM_E_UNSUPPORTED, "radix sort: n >= 2^32");
  const u32 tiles = (u32)((n + kSortTile - 1) / kSortTile);
  DevBuf<u32> hist((size_t)256 * tiles);
  int cur = 0;
  for (int shift = beginBit; shift < endBit; shift += 8) {
    KeyT* kin = cur ? keysB : keysA;
    KeyT* kout = cur ? keysA : keysB;
    u32* vin = cur ? valsB : valsA;
    u32* vout = cur ? valsA : valsB;
    k_radix_hist<KeyT><<<tiles, kSortThreads, 0, s>>>(kin, n, shift, hist.get(), tiles);
    SB_LAUNCH_CHECK();
    exclusive_scan_u32(hist.get(), hist.get(), (size_t)256 * tiles, s);
    if (valsA)
      k_radix_scatter<KeyT, true><<<tiles, kSortThreads, 0, s>>>(kin, vin, kout, vout, n, shift, hist.get(), tiles);
    else
      k_radix_scatter<KeyT, false><<<tiles, kSortThreads, 0, s>>>(kin, nullptr, kout, nullptr, n, shift, hist.get(),
                                                                   tiles);
    SB_LAUNCH_CHECK();
    cur ^= 1;
  }
  if (!g_temp.active) SB_CUDA(cudaStreamSynchronize(s));  // hist is freed on return (stream-ordered inside a TempScope)
  return cur;
}

inline int bits_for(u64 maxValue) {
  int b = 0;
  while (maxValue) { ++b; maxValue >>= 1; }
  return b ? b : 1;
}

}  // namespace sb
