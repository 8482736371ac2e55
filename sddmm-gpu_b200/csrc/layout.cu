// layout.cu -- column reordering, dense/sparse split and RPHM construction ON DEVICE
// (a7 + a8 of SURVEY.md 8a; the reference does both on the host with OpenMP:
//  colReordering_cpu src/colReordering.cu:274-404, RPHM::RPHM src/BSMR.cpp:83-265).
//
// Formulation (all stages are stable radix sorts, scans and flat scatter kernels, so cost is
// O(nnz) HBM traffic per pass regardless of how skewed the row panels are):
//   1. every stored entry of the selected panels gets the key (panel | col | row-in-panel) and its
//      CSR index as payload; one stable sort groups entries by (panel, col), rows ascending;
//   2. run heads give the distinct (panel, col) "column groups" with their counts (1..16);
//   3. a stable sort of the groups by (panel, 16-count) yields, per panel, columns ordered by
//      count descending / column ascending  == thrust::stable_sort_by_key(greater) in the reference;
//   4. per panel: pad to x16 with sentinel N, blocks whose count-sum >= ceil(delta*256) are dense;
//   5. scans give every offset array; one scatter writes denseCols / sparseCols / blockValues and
//      the residual COO arrays in exactly the reference's order (appendix B of SURVEY.md).
#include <vector>

#include <algorithm>
#include <vector>

#include "layout.cuh"
#include "primitives.cuh"

namespace sb {

namespace {

__global__ void k_scatter_rowpos(const u32* __restrict__ R, u32 r0, u32 n, u32* __restrict__ rowLenOut,
                                 const u32* __restrict__ rowOff) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const u32 row = R[r0 + i];
    rowLenOut[i] = rowOff[row + 1] - rowOff[row];
  }
}

// one warp per reordered row: emit keys (panelLocal | col | r) + CSR index
__global__ void __launch_bounds__(256) k_make_entry_keys(const u32* __restrict__ R, u32 r0, u32 n,
                                                         const u32* __restrict__ rowOff,
                                                         const u32* __restrict__ colIdx,
                                                         const u32* __restrict__ eOff, int colBits,
                                                         u64* __restrict__ keys, u32* __restrict__ vals) {
  const u32 gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const u32 nw = (gridDim.x * blockDim.x) >> 5;
  for (u32 i = gw; i < n; i += nw) {
    const u32 row = R[r0 + i];
    const u32 b = rowOff[row], len = rowOff[row + 1] - b;
    const u32 o = eOff[i];
    const u64 hi = ((u64)(i >> 4) << (colBits + 4)) | (i & 15u);
    for (u32 j = lane; j < len; j += 32) {
      keys[o + j] = hi | ((u64)colIdx[b + j] << 4);
      vals[o + j] = b + j;
    }
  }
}

__global__ void k_group_heads(const u64* __restrict__ keys, size_t n, u32* __restrict__ head) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    head[i] = (i == 0 || (keys[i] >> 4) != (keys[i - 1] >> 4)) ? 1u : 0u;
}

// gid = exclusive scan of head; at heads gidIncl-1... we scan `head` exclusively into gidEx, so the group
// of element i is gidEx[i] + head[i] - 1.
__global__ void k_group_starts(const u32* __restrict__ head, const u32* __restrict__ gidEx, size_t n,
                               u32* __restrict__ gStart) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    if (head[i]) gStart[gidEx[i]] = (u32)i;
}

__global__ void k_group_keys(const u64* __restrict__ keys, const u32* __restrict__ gStart, u32 numGroups, u32 nSel,
                             int colBits, u32* __restrict__ sortKey, u32* __restrict__ gCount) {
  for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < numGroups; g += (size_t)gridDim.x * blockDim.x) {
    const u32 s = gStart[g];
    const u32 e = (g + 1 < numGroups) ? gStart[g + 1] : nSel;
    const u32 cnt = e - s;  // 1..16 (no duplicate (row,col): src/Matrix.cpp:447-461)
    const u32 panel = (u32)(keys[s] >> (colBits + 4));
    sortKey[g] = (panel << 5) | (16u - (cnt > 16u ? 16u : cnt));
    gCount[g] = cnt;
  }
}

// sorted group j -> panel boundaries, inverse permutation
__global__ void k_sorted_group_info(const u32* __restrict__ sortedKey, const u32* __restrict__ order, u32 numGroups,
                                    u32* __restrict__ pStart, u32* __restrict__ rank, u32 P) {
  for (size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x; j < numGroups; j += (size_t)gridDim.x * blockDim.x) {
    rank[order[j]] = (u32)j;
    const u32 p = sortedKey[j] >> 5;
    if (j == 0 || (sortedKey[j - 1] >> 5) != p) pStart[p] = (u32)j;
    if (j + 1 == numGroups) pStart[P] = numGroups;
  }
}

// per panel: nd (dense columns), padded column count, residual / dense entry counts.  One warp per
// panel, two 16-column blocks per step (half-warps).
__global__ void __launch_bounds__(256) k_panel_split(const u32* __restrict__ sortedKey, const u32* __restrict__ pStart,
                                                     u32 P, u32 T, u32* __restrict__ nd, u32* __restrict__ nsCols,
                                                     u32* __restrict__ nnzSparse, u32* __restrict__ nBlk,
                                                     u32* __restrict__ denseTB, u32* __restrict__ sparseTB,
                                                     u32* __restrict__ myDenseWork, u32* __restrict__ mySparseWork, u32 kSparseChunk,
                                                     u32* __restrict__ maxima /* [0]=maxDenseBlk [1]=maxSparseTB [2]=denseNnz */) {
  const u32 gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const u32 nw = (gridDim.x * blockDim.x) >> 5;
  for (u32 p = gw; p < P; p += nw) {
    const u32 s = pStart[p], nCols = pStart[p + 1] - s;
    const u32 padded = (nCols + 15u) & ~15u;
    u32 denseBlocks = 0, total = 0, denseNnz = 0;
    for (u32 base = 0; base < padded; base += 32) {
      const u32 slot = base + lane;
      const u32 cnt = slot < nCols ? 16u - (sortedKey[s + slot] & 31u) : 0u;
      u32 sum = cnt;
#pragma unroll
      for (int w = 1; w < 16; w <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, w);
      // sum = count-sum of this lane's 16-column block
      const bool blockExists = (base + (lane & 16u)) < padded;
      const bool dense = blockExists && sum >= T;
      const unsigned dm = __ballot_sync(0xffffffffu, dense && (lane & 15u) == 0);
      denseBlocks += __popc(dm);
      u32 t = cnt, dn = dense ? cnt : 0u;
#pragma unroll
      for (int w = 1; w < 32; w <<= 1) {
        t += __shfl_xor_sync(0xffffffffu, t, w);
        dn += __shfl_xor_sync(0xffffffffu, dn, w);
      }
      total += t;
      denseNnz += dn;
    }
    if (lane == 0) {
      // colReordering.cu:250-261 counts EVERY block whose sum reaches T; counts are non-increasing so
      // they form a prefix, and the first nd columns are taken as dense (:380-400)
      const u32 ndp = denseBlocks * 16u;
      // entries in the first ndp slots (prefix), which equals denseNnz because dense blocks are a prefix
      nd[p] = ndp;
      nsCols[p] = padded - ndp;
      nnzSparse[p] = total - denseNnz;
      nBlk[p] = denseBlocks;
      const u32 dtb = (denseBlocks + 3u) / 4u, stb = (total - denseNnz + 127u) / 128u;
      denseTB[p] = dtb;
      sparseTB[p] = stb;
      myDenseWork[p] = (denseBlocks + kDenseGroupBlocks - 1) / kDenseGroupBlocks;
      mySparseWork[p] = (total - denseNnz + kSparseChunk - 1) / kSparseChunk;
      atomicMax(maxima + 0, denseBlocks);
      atomicMax(maxima + 1, stb);
      atomicAdd(maxima + 2, denseNnz);
    }
  }
}

__global__ void k_sparse_group_counts(const u32* __restrict__ sortedKey, const u32* __restrict__ pStart,
                                      const u32* __restrict__ nd, u32 numGroups, u32* __restrict__ out) {
  for (size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x; j < numGroups; j += (size_t)gridDim.x * blockDim.x) {
    const u32 p = sortedKey[j] >> 5;
    const u32 slot = (u32)j - pStart[p];
    out[j] = slot >= nd[p] ? 16u - (sortedKey[j] & 31u) : 0u;
  }
}

// columns: sorted group j -> denseCols / sparseCols
__global__ void k_write_cols(const u32* __restrict__ sortedKey, const u32* __restrict__ order,
                             const u64* __restrict__ keys, const u32* __restrict__ gStart,
                             const u32* __restrict__ pStart, const u32* __restrict__ nd,
                             const u32* __restrict__ dOff, const u32* __restrict__ sOff, u32 numGroups, int colBits,
                             u32* __restrict__ denseCols, u32* __restrict__ sparseCols) {
  const u64 colMask = (((u64)1) << colBits) - 1;
  for (size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x; j < numGroups; j += (size_t)gridDim.x * blockDim.x) {
    const u32 p = sortedKey[j] >> 5;
    const u32 slot = (u32)j - pStart[p];
    const u32 col = (u32)((keys[gStart[order[j]]] >> 4) & colMask);
    const u32 ndp = nd[p];
    if (slot < ndp) denseCols[dOff[p] + slot] = col;
    else sparseCols[sOff[p] + (slot - ndp)] = col;
  }
}

// padding sentinels (colReordering.cu:338-343): slots [nCols, padded) hold N
__global__ void k_write_pad(const u32* __restrict__ pStart, const u32* __restrict__ nd, const u32* __restrict__ dOff,
                            const u32* __restrict__ sOff, u32 P, u32 N, u32* __restrict__ denseCols,
                            u32* __restrict__ sparseCols) {
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < (size_t)P * 16; t += (size_t)gridDim.x * blockDim.x) {
    const u32 p = (u32)(t >> 4), k = (u32)(t & 15);
    const u32 nCols = pStart[p + 1] - pStart[p];
    const u32 padded = (nCols + 15u) & ~15u;
    const u32 slot = nCols + k;
    if (slot >= padded) continue;
    const u32 ndp = nd[p];
    if (slot < ndp) denseCols[dOff[p] + slot] = N;
    else sparseCols[sOff[p] + (slot - ndp)] = N;
  }
}

// entries: sorted element i -> blockValues or the residual COO arrays
__global__ void k_write_entries(const u64* __restrict__ keys, const u32* __restrict__ vals,
                                const u32* __restrict__ head, const u32* __restrict__ gidEx,
                                const u32* __restrict__ gStart, const u32* __restrict__ rank,
                                const u32* __restrict__ pStart, const u32* __restrict__ nd,
                                const u32* __restrict__ bOff, const u32* __restrict__ gSparseOff, size_t n,
                                int colBits, u32* __restrict__ blockValues, u32* __restrict__ sparseValues,
                                u32* __restrict__ sparseRelRows, u32* __restrict__ sparseColIdx) {
  const u64 colMask = (((u64)1) << colBits) - 1;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const u64 k = keys[i];
    const u32 g = gidEx[i] + head[i] - 1u;
    const u32 j = rank[g];
    const u32 p = (u32)(k >> (colBits + 4));
    const u32 r = (u32)(k & 15u);
    const u32 slot = j - pStart[p];
    const u32 idx = vals[i];
    if (slot < nd[p]) {
      blockValues[((size_t)bOff[p] + (slot >> 4)) * 256u + r * 16u + (slot & 15u)] = idx;
    } else {
      const u32 pos = gSparseOff[j] + ((u32)i - gStart[g]);
      sparseValues[pos] = idx;
      sparseRelRows[pos] = r;
      sparseColIdx[pos] = (u32)((k >> 4) & colMask);
    }
  }
}

// reference-shaped work lists (BSMR.cpp:99-119, :221-246) and ours
__global__ void k_write_worklists(const u32* __restrict__ dOff, const u32* __restrict__ nBlk,
                                  const u32* __restrict__ nnzSparse, const u32* __restrict__ dtbOff,
                                  const u32* __restrict__ stbOff, const u32* __restrict__ myDOff,
                                  const u32* __restrict__ mySOff, u32 P, u32* __restrict__ dIds,
                                  u32* __restrict__ dIters, u32* __restrict__ sIds, u32* __restrict__ sIters,
                                  uint2* __restrict__ myDense, uint2* __restrict__ mySparse, u32 kSparseChunk) {
  const u32 gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const u32 nw = (gridDim.x * blockDim.x) >> 5;
  for (u32 p = gw; p < P; p += nw) {
    const u32 dtb = (nBlk[p] + 3u) / 4u, stb = (nnzSparse[p] + 127u) / 128u;
    for (u32 i = lane; i < dtb; i += 32) {
      dIds[dtbOff[p] + i] = p;
      dIters[dtbOff[p] + i] = dOff[p] / 16u + i * 4u;
    }
    for (u32 i = lane; i < stb; i += 32) {
      sIds[stbOff[p] + i] = p;
      sIters[stbOff[p] + i] = i * 128u;
    }
    const u32 md = (nBlk[p] + kDenseGroupBlocks - 1) / kDenseGroupBlocks;
    const u32 ms = (nnzSparse[p] + kSparseChunk - 1) / kSparseChunk;
    for (u32 i = lane; i < md; i += 32) myDense[myDOff[p] + i] = make_uint2(p, i * kDenseGroupBlocks);
    for (u32 i = lane; i < ms; i += 32) mySparse[mySOff[p] + i] = make_uint2(p, i * kSparseChunk);
  }
}

// chunk-major ordering of the residual work list: residual entries of a panel are ordered by
// (count desc, col asc), i.e. chunk c of every panel covers about the same column range, so running
// all panels' chunk c together keeps the active slice of B small enough to stay L2-resident.
__global__ void k_work_keys(const uint2* __restrict__ work, u32 n, u32 kSparseChunk, u32* __restrict__ keys) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    keys[i] = work[i].y / kSparseChunk;
}
__global__ void k_gather_work(const uint2* __restrict__ in, const u32* __restrict__ idx, u32 n,
                              uint2* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = in[idx[i]];
}


// ---- super-panel residual layout (K7b) -----------------------------------------------------------
// one warp per panel: key = (superPanel | col | rowInSuperPanel), payload = residual entry id
__global__ void __launch_bounds__(256) k_sp_keys(const u32* __restrict__ vOff, const u32* __restrict__ sCols,
                                                 const u32* __restrict__ sRows, u32 P, u32 G, int rowBits,
                                                 int colBits, u64* __restrict__ keys, u32* __restrict__ vals) {
  const u32 gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const u32 nw = (gridDim.x * blockDim.x) >> 5;
  for (u32 p = gw; p < P; p += nw) {
    const u32 b = vOff[p], e = vOff[p + 1];
    const u64 hi = (u64)(p / G) << (rowBits + colBits);
    const u32 rbase = (p % G) * 16u;
    for (u32 i = b + lane; i < e; i += 32) {
      keys[i] = hi | ((u64)sCols[i] << rowBits) | (rbase + sRows[i]);
      vals[i] = i;
    }
  }
}
__global__ void k_sp_unpack(const u64* __restrict__ keys, const u32* __restrict__ vals,
                            const u32* __restrict__ sVals, size_t n, int rowBits, int colBits,
                            u32* __restrict__ col, unsigned short* __restrict__ row, u32* __restrict__ idx,
                            u32* __restrict__ numRuns) {
  const u64 colMask = (((u64)1) << colBits) - 1, rowMask = (((u64)1) << rowBits) - 1;
  u32 heads = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const u64 k = keys[i];
    col[i] = (u32)((k >> rowBits) & colMask);
    row[i] = (unsigned short)(k & rowMask);
    idx[i] = sVals[vals[i]];
    heads += (i == 0 || (k >> rowBits) != (keys[i - 1] >> rowBits)) ? 1u : 0u;
  }
  heads = __reduce_add_sync(0xffffffffu, heads);
  if ((threadIdx.x & 31) == 0 && heads) atomicAdd(numRuns, heads);
}
__global__ void k_sp_offsets(const u32* __restrict__ vOff, u32 P, u32 G, u32 numSp, u32 segLen,
                             u32* __restrict__ off, u32* __restrict__ segCnt) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i <= numSp; i += (size_t)gridDim.x * blockDim.x) {
    const u32 p = (u32)i * G < P ? (u32)i * G : P;
    off[i] = vOff[p];
    if (i < numSp) {
      const u32 p1 = ((u32)i + 1) * G < P ? ((u32)i + 1) * G : P;
      segCnt[i] = (vOff[p1] - vOff[p] + segLen - 1) / segLen;
    }
  }
}
__global__ void k_sp_work(const u32* __restrict__ segOff, u32 numSp, u32 segLen, uint2* __restrict__ work,
                          u32* __restrict__ keys) {
  const u32 gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const u32 nw = (gridDim.x * blockDim.x) >> 5;
  for (u32 sp = gw; sp < numSp; sp += nw) {
    const u32 b = segOff[sp], n = segOff[sp + 1] - b;
    for (u32 j = lane; j < n; j += 32) {
      work[b + j] = make_uint2(sp, j * segLen);
      keys[b + j] = j;
    }
  }
}

// ---- full-tile layout (K8) -------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_tile_keys(const u32* __restrict__ R, u32 r0, u32 n,
                                                   const u32* __restrict__ rowOff, const u32* __restrict__ colIdx,
                                                   const u32* __restrict__ eOff, int ctBits, u64* __restrict__ keys,
                                                   u32* __restrict__ vals) {
  const u32 gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const u32 nw = (gridDim.x * blockDim.x) >> 5;
  for (u32 i = gw; i < n; i += nw) {
    const u32 row = R[r0 + i];
    const u32 b = rowOff[row], len = rowOff[row + 1] - b;
    const u32 o = eOff[i];
    const u64 hi = ((u64)(i >> 7) << (ctBits + 14)) | ((u64)(i & 127u) << 7);
    for (u32 j = lane; j < len; j += 32) {
      const u32 c = colIdx[b + j];
      keys[o + j] = hi | ((u64)(c >> 7) << 14) | (c & 127u);
      vals[o + j] = b + j;
    }
  }
}
__global__ void k_tile_heads(const u64* __restrict__ keys, size_t n, u32* __restrict__ head) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    head[i] = (i == 0 || (keys[i] >> 14) != (keys[i - 1] >> 14)) ? 1u : 0u;
}
__global__ void k_tile_records(const u64* __restrict__ keys, const u32* __restrict__ head,
                               const u32* __restrict__ tid, size_t n, int ctBits, uint4* __restrict__ tiles) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (!head[i]) continue;
    const u64 k = keys[i] >> 14;
    const u32 tc = (u32)(k & ((((u64)1) << ctBits) - 1));
    const u32 tr = (u32)(k >> ctBits);
    tiles[tid[i]] = make_uint4(tr, tc, (u32)i, 0u);
  }
}
// one CTA (128 threads) per tile: row masks + entry offsets
__global__ void __launch_bounds__(128) k_tile_meta(const u64* __restrict__ keys, uint4* __restrict__ tiles, u32 numTiles,
                                                   u32 nEntries, u32* __restrict__ rowMeta) {
  __shared__ u32 mask[128][4];
  __shared__ u32 warpTot[4];
  const u32 t = blockIdx.x;
  const u32 beg = tiles[t].z;
  const u32 end = (t + 1 < numTiles) ? tiles[t + 1].z : nEntries;
  for (u32 i = threadIdx.x; i < 512; i += 128) (&mask[0][0])[i] = 0u;
  __syncthreads();
  for (u32 e = beg + threadIdx.x; e < end; e += 128) {
    const u32 k = (u32)(keys[e] & 0x3FFFu);
    const u32 r = k >> 7, c = k & 127u;
    atomicOr(&mask[r][c >> 5], 1u << (c & 31u));
  }
  __syncthreads();
  const u32 r = threadIdx.x;
  const u32 cnt = __popc(mask[r][0]) + __popc(mask[r][1]) + __popc(mask[r][2]) + __popc(mask[r][3]);
  u32 incl = cnt;
  const u32 lane = r & 31u, warp = r >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const u32 v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= (u32)d) incl += v;
  }
  if (lane == 31) warpTot[warp] = incl;
  __syncthreads();
  u32 base = 0;
  for (u32 w = 0; w < warp; ++w) base += warpTot[w];
  u32* out = rowMeta + (size_t)t * 640u + r * 5u;
  out[0] = mask[r][0]; out[1] = mask[r][1]; out[2] = mask[r][2]; out[3] = mask[r][3];
  out[4] = beg + base + incl - cnt;
  if (threadIdx.x == 0) tiles[t].w = end - beg;
}

u32 read_u32(const u32* d, cudaStream_t s) {
  u32 h = 0;
  SB_CUDA(cudaMemcpyAsync(&h, d, 4, cudaMemcpyDeviceToHost, s));
  SB_CUDA(cudaStreamSynchronize(s));
  return h;
}

// exclusive scan of a P-array into a (P+1)-array (out[P] = total)
void scan_counts(const u32* cnt, u32* out, u32 P, cudaStream_t s) {
  SB_CUDA(cudaMemcpyAsync(out, cnt, (size_t)P * 4, cudaMemcpyDeviceToDevice, s));
  SB_CUDA(cudaMemsetAsync(out + P, 0, 4, s));
  exclusive_scan_u32(out, out, (size_t)P + 1, s);
}

}  // namespace

static void build_tiles(bsmr_layout* L, const u32* d_rowOff, const u32* d_colIdx, const u32* d_R, u32 r0, u32 nR,
                        const u32* eOff, u32 nSel, cudaStream_t s);

bsmr_layout* layout_build_dev(const u32* d_rowOff, const u32* d_colIdx, u32 M, u32 N, u32 nnz, const u32* d_R,
                              u32 numRows, float delta, u32 panelBegin, u32 panelEnd, u32 tileFlags, float* msCol,
                              float* msRphm, cudaStream_t s) {
  const u32 Pall = (numRows + kPanel - 1) / kPanel;  // BSMR.cpp:48
  if (panelEnd > Pall) panelEnd = Pall;
  if (panelBegin > panelEnd) panelBegin = panelEnd;
  const u32 P = panelEnd - panelBegin;
  const u32 r0 = panelBegin * kPanel;
  const u32 r1 = (panelEnd * kPanel < numRows) ? panelEnd * kPanel : numRows;
  const u32 nR = r1 > r0 ? r1 - r0 : 0;
  // colReordering.cu:246  static_cast<UIN>(std::ceil(delta * BLOCK_SIZE)), float arithmetic
  const u32 T = (u32)std::ceil(delta * (float)(kPanel * kBlockCols));

  auto* L = new bsmr_layout();
  TempScope tempScope(s);
  try {
    SB_CUDA(cudaGetDevice(&L->device));
    L->sparseChunk = kSparseChunkDefault;
    if (const char* e = getenv("SDDMM_B200_CHUNK")) {
      const long v = atol(e);
      if (v >= 32 && v <= 65536) L->sparseChunk = (u32)v;
    }
    bsmr_layout_info& I = L->info;
    I.M = M; I.N = N; I.nnz = nnz; I.numRows = nR; I.numRowPanels = P; I.panelBegin = panelBegin;
    auto A = [&](bsmr_array_id id) -> DevBuf<u32>& { return L->arr[id]; };
    A(BSMR_REORDERED_ROWS).alloc(nR ? nR : 1, true);
    if (nR) SB_CUDA(cudaMemcpyAsync(A(BSMR_REORDERED_ROWS).get(), d_R + r0, (size_t)nR * 4, cudaMemcpyDeviceToDevice, s));
    for (bsmr_array_id id : {BSMR_DENSE_COL_OFFSETS, BSMR_SPARSE_COL_OFFSETS, BSMR_SPARSE_VALUE_OFFSETS, RPHM_BLOCK_OFFSETS}) {
      A(id).alloc((size_t)P + 1, true);
      SB_CUDA(cudaMemsetAsync(A(id).get(), 0, ((size_t)P + 1) * 4, s));
    }
    Timer tCol(s);
    tCol.start();
    if (P == 0) {
      for (bsmr_array_id id : {BSMR_DENSE_COLS, BSMR_SPARSE_COLS, RPHM_BLOCK_VALUES, RPHM_SPARSE_VALUES,
                               RPHM_SPARSE_RELATIVE_ROWS, RPHM_SPARSE_COL_INDICES, RPHM_DENSE_ROW_PANEL_IDS,
                               RPHM_DENSE_COL_BLOCK_ITERS, RPHM_SPARSE_ROW_PANEL_IDS, RPHM_SPARSE_COL_BLOCK_ITERS})
        A(id).alloc(1, true), A(id).n = 0;
      if (msCol) *msCol = tCol.stop();
      if (msRphm) *msRphm = 0.f;
      return L;
    }

    // ---- 1. keys for every stored entry of the selected rows
    DevBuf<u32> eOff((size_t)nR + 1);
    k_scatter_rowpos<<<grid_for(nR), 256, 0, s>>>(d_R, r0, nR, eOff.get(), d_rowOff);
    SB_LAUNCH_CHECK();
    SB_CUDA(cudaMemsetAsync(eOff.get() + nR, 0, 4, s));
    exclusive_scan_u32(eOff.get(), eOff.get(), (size_t)nR + 1, s);
    const u32 nSel = read_u32(eOff.get() + nR, s);
    // dense enough for whole 128x128 tensor-core tiles to be worth considering? (>= ~1% of the slots)
    {
      const u64 slots = (u64)((nR + 127u) / 128u) * ((N + 127u) / 128u) * 16384ull;
      int plan = 1;  // 0 never, 1 auto, 2 always; the environment variable only fills in for AUTO
      if (tileFlags == BSMR_BUILD_TILES_ALWAYS) plan = 2;
      else if (tileFlags == BSMR_BUILD_TILES_NEVER) plan = 0;
      else if (const char* e = getenv("SDDMM_B200_PLAN")) plan = !strcmp(e, "full") ? 2 : !strcmp(e, "bsmr") ? 0 : 1;
      if (plan == 2 || (plan == 1 && (u64)nSel * 100ull >= slots)) build_tiles(L, d_rowOff, d_colIdx, d_R, r0, nR, eOff.get(), nSel, s);
    }
    const int colBits = bits_for(N);  // sentinel-free here: real columns are < N
    const int panelBits = bits_for(P);
    if (panelBits > 27) fail(SDDMM_E_UNSUPPORTED, "too many row panels (%u)", P);
    DevBuf<u64> keyA(nSel ? nSel : 1), keyB(nSel ? nSel : 1);
    DevBuf<u32> valA(nSel ? nSel : 1), valB(nSel ? nSel : 1);
    k_make_entry_keys<<<grid_for((size_t)nR * 32), 256, 0, s>>>(d_R, r0, nR, d_rowOff, d_colIdx, eOff.get(), colBits,
                                                              keyA.get(), valA.get());
    SB_LAUNCH_CHECK();
    const int w = radix_sort_pairs<u64>(keyA.get(), keyB.get(), valA.get(), valB.get(), nSel, 0,
                                        4 + colBits + panelBits, s);
    const u64* keys = w ? keyB.get() : keyA.get();
    const u32* vals = w ? valB.get() : valA.get();
    (w ? keyA : keyB).release();
    (w ? valA : valB).release();

    // ---- 2. (panel, col) groups
    DevBuf<u32> head(nSel ? nSel : 1), gidEx((size_t)nSel + 1);
    k_group_heads<<<grid_for(nSel), 256, 0, s>>>(keys, nSel, head.get());
    SB_LAUNCH_CHECK();
    SB_CUDA(cudaMemcpyAsync(gidEx.get(), head.get(), (size_t)nSel * 4, cudaMemcpyDeviceToDevice, s));
    SB_CUDA(cudaMemsetAsync(gidEx.get() + nSel, 0, 4, s));
    exclusive_scan_u32(gidEx.get(), gidEx.get(), (size_t)nSel + 1, s);
    const u32 numGroups = read_u32(gidEx.get() + nSel, s);
    DevBuf<u32> gStart(numGroups ? numGroups : 1), gCount(numGroups ? numGroups : 1);
    k_group_starts<<<grid_for(nSel), 256, 0, s>>>(head.get(), gidEx.get(), nSel, gStart.get());
    SB_LAUNCH_CHECK();

    // ---- 3. per panel: columns by (count desc, col asc)
    DevBuf<u32> skA(numGroups ? numGroups : 1), skB(numGroups ? numGroups : 1), ordA(numGroups ? numGroups : 1),
        ordB(numGroups ? numGroups : 1);
    k_group_keys<<<grid_for(numGroups), 256, 0, s>>>(keys, gStart.get(), numGroups, nSel, colBits, skA.get(),
                                                    gCount.get());
    SB_LAUNCH_CHECK();
    iota<u32>(ordA.get(), numGroups, 0u, s);
    const int w2 = radix_sort_pairs<u32>(skA.get(), skB.get(), ordA.get(), ordB.get(), numGroups, 0, 5 + panelBits, s);
    const u32* sortedKey = w2 ? skB.get() : skA.get();
    const u32* order = w2 ? ordB.get() : ordA.get();
    DevBuf<u32> pStart((size_t)P + 1), rank(numGroups ? numGroups : 1);
    SB_CUDA(cudaMemsetAsync(pStart.get(), 0, ((size_t)P + 1) * 4, s));
    k_sorted_group_info<<<grid_for(numGroups), 256, 0, s>>>(sortedKey, order, numGroups, pStart.get(), rank.get(), P);
    SB_LAUNCH_CHECK();

    // ---- 4. dense prefix per panel
    DevBuf<u32> nd(P), nsCols(P), nnzSparse(P), nBlk(P), denseTB(P), sparseTB(P), myDW(P), mySW(P), maxima(4);
    SB_CUDA(cudaMemsetAsync(maxima.get(), 0, 16, s));
    k_panel_split<<<grid_for((size_t)P * 32), 256, 0, s>>>(sortedKey, pStart.get(), P, T, nd.get(), nsCols.get(),
                                                          nnzSparse.get(), nBlk.get(), denseTB.get(), sparseTB.get(),
                                                          myDW.get(), mySW.get(), L->sparseChunk, maxima.get());
    SB_LAUNCH_CHECK();

    // ---- 5. offsets
    scan_counts(nd.get(), A(BSMR_DENSE_COL_OFFSETS).get(), P, s);
    scan_counts(nsCols.get(), A(BSMR_SPARSE_COL_OFFSETS).get(), P, s);
    scan_counts(nnzSparse.get(), A(BSMR_SPARSE_VALUE_OFFSETS).get(), P, s);
    const u32 dTot = read_u32(A(BSMR_DENSE_COL_OFFSETS).get() + P, s);
    const u32 sTot = read_u32(A(BSMR_SPARSE_COL_OFFSETS).get() + P, s);
    const u32 vTot = read_u32(A(BSMR_SPARSE_VALUE_OFFSETS).get() + P, s);
    A(BSMR_DENSE_COLS).alloc(dTot ? dTot : 1, true); A(BSMR_DENSE_COLS).n = dTot;
    A(BSMR_SPARSE_COLS).alloc(sTot ? sTot : 1, true); A(BSMR_SPARSE_COLS).n = sTot;
    k_write_cols<<<grid_for(numGroups), 256, 0, s>>>(sortedKey, order, keys, gStart.get(), pStart.get(), nd.get(),
                                                    A(BSMR_DENSE_COL_OFFSETS).get(), A(BSMR_SPARSE_COL_OFFSETS).get(),
                                                    numGroups, colBits, A(BSMR_DENSE_COLS).get(),
                                                    A(BSMR_SPARSE_COLS).get());
    SB_LAUNCH_CHECK();
    k_write_pad<<<grid_for((size_t)P * 16), 256, 0, s>>>(pStart.get(), nd.get(), A(BSMR_DENSE_COL_OFFSETS).get(),
                                                        A(BSMR_SPARSE_COL_OFFSETS).get(), P, N,
                                                        A(BSMR_DENSE_COLS).get(), A(BSMR_SPARSE_COLS).get());
    SB_LAUNCH_CHECK();
    const float colMs = tCol.stop();
    if (msCol) *msCol = colMs;

    // ---- RPHM (BSMR.cpp:83-265)
    Timer tR(s);
    tR.start();
    scan_counts(nBlk.get(), A(RPHM_BLOCK_OFFSETS).get(), P, s);
    const u32 numBlocks = read_u32(A(RPHM_BLOCK_OFFSETS).get() + P, s);
    const size_t nbv = (size_t)numBlocks * 256u;
    A(RPHM_BLOCK_VALUES).alloc(nbv ? nbv : 1, true); A(RPHM_BLOCK_VALUES).n = nbv;
    fill<u32>(A(RPHM_BLOCK_VALUES).get(), nbv, kNull, s);
    for (bsmr_array_id id : {RPHM_SPARSE_VALUES, RPHM_SPARSE_RELATIVE_ROWS, RPHM_SPARSE_COL_INDICES}) {
      A(id).alloc(vTot ? vTot : 1, true);
      A(id).n = vTot;
    }
    DevBuf<u32> gSparseOff((size_t)numGroups + 1);
    k_sparse_group_counts<<<grid_for(numGroups), 256, 0, s>>>(sortedKey, pStart.get(), nd.get(), numGroups,
                                                             gSparseOff.get());
    SB_LAUNCH_CHECK();
    SB_CUDA(cudaMemsetAsync(gSparseOff.get() + numGroups, 0, 4, s));
    exclusive_scan_u32(gSparseOff.get(), gSparseOff.get(), (size_t)numGroups + 1, s);
    k_write_entries<<<grid_for(nSel), 256, 0, s>>>(keys, vals, head.get(), gidEx.get(), gStart.get(), rank.get(),
                                                  pStart.get(), nd.get(), A(RPHM_BLOCK_OFFSETS).get(), gSparseOff.get(),
                                                  nSel, colBits, A(RPHM_BLOCK_VALUES).get(),
                                                  A(RPHM_SPARSE_VALUES).get(), A(RPHM_SPARSE_RELATIVE_ROWS).get(),
                                                  A(RPHM_SPARSE_COL_INDICES).get());
    SB_LAUNCH_CHECK();

    // work lists
    DevBuf<u32> dtbOff((size_t)P + 1), stbOff((size_t)P + 1), myDOff((size_t)P + 1), mySOff((size_t)P + 1);
    scan_counts(denseTB.get(), dtbOff.get(), P, s);
    scan_counts(sparseTB.get(), stbOff.get(), P, s);
    scan_counts(myDW.get(), myDOff.get(), P, s);
    scan_counts(mySW.get(), mySOff.get(), P, s);
    const u32 nDTB = read_u32(dtbOff.get() + P, s), nSTB = read_u32(stbOff.get() + P, s);
    L->numDenseWork = read_u32(myDOff.get() + P, s);
    L->numSparseWork = read_u32(mySOff.get() + P, s);
    A(RPHM_DENSE_ROW_PANEL_IDS).alloc(nDTB ? nDTB : 1, true); A(RPHM_DENSE_ROW_PANEL_IDS).n = nDTB;
    A(RPHM_DENSE_COL_BLOCK_ITERS).alloc(nDTB ? nDTB : 1, true); A(RPHM_DENSE_COL_BLOCK_ITERS).n = nDTB;
    A(RPHM_SPARSE_ROW_PANEL_IDS).alloc(nSTB ? nSTB : 1, true); A(RPHM_SPARSE_ROW_PANEL_IDS).n = nSTB;
    A(RPHM_SPARSE_COL_BLOCK_ITERS).alloc(nSTB ? nSTB : 1, true); A(RPHM_SPARSE_COL_BLOCK_ITERS).n = nSTB;
    L->denseWork.alloc(L->numDenseWork ? L->numDenseWork : 1, true);
    L->sparseWork.alloc(L->numSparseWork ? L->numSparseWork : 1, true);
    k_write_worklists<<<grid_for((size_t)P * 32), 256, 0, s>>>(
        A(BSMR_DENSE_COL_OFFSETS).get(), nBlk.get(), nnzSparse.get(), dtbOff.get(), stbOff.get(), myDOff.get(),
        mySOff.get(), P, A(RPHM_DENSE_ROW_PANEL_IDS).get(), A(RPHM_DENSE_COL_BLOCK_ITERS).get(),
        A(RPHM_SPARSE_ROW_PANEL_IDS).get(), A(RPHM_SPARSE_COL_BLOCK_ITERS).get(), L->denseWork.get(),
        L->sparseWork.get(), L->sparseChunk);
    SB_LAUNCH_CHECK();
    if (L->numSparseWork > 1) {
      const u32 nw = L->numSparseWork;
      DevBuf<u32> kA(nw), kB(nw), iA(nw), iB(nw);
      DevBuf<uint2> sorted;
      sorted.alloc(nw, true);  // becomes the layout's work list
      k_work_keys<<<grid_for(nw), 256, 0, s>>>(L->sparseWork.get(), nw, L->sparseChunk, kA.get());
      SB_LAUNCH_CHECK();
      iota<u32>(iA.get(), nw, 0u, s);
      const int ws = radix_sort_pairs<u32>(kA.get(), kB.get(), iA.get(), iB.get(), nw, 0, 24, s);
      k_gather_work<<<grid_for(nw), 256, 0, s>>>(L->sparseWork.get(), ws ? iB.get() : iA.get(), nw, sorted.get());
      SB_LAUNCH_CHECK();
      SB_CUDA(cudaStreamSynchronize(s));
      L->sparseWork = std::move(sorted);
    }
    u32 hmax[4];
    SB_CUDA(cudaMemcpyAsync(hmax, maxima.get(), 16, cudaMemcpyDeviceToHost, s));
    const float rMs = tR.stop();
    if (msRphm) *msRphm = rMs;

    A(BSMR_REORDERED_ROWS).n = nR;
    I.numDenseBlocks = numBlocks;
    I.numSparseValues = vTot;
    I.numDenseValues = hmax[2];
    I.maxNumDenseColBlocksInRowPanel = hmax[0];
    I.maxNumSparseColBlocksInRowPanel = hmax[1];
    I.numDenseThreadBlocks = nDTB;
    I.numSparseThreadBlocks = nSTB;
    return L;
  } catch (...) {
    delete L;
    throw;
  }
}

// builds L->tl (full-tile plan) from the selected rows
static void build_tiles(bsmr_layout* L, const u32* d_rowOff, const u32* d_colIdx, const u32* d_R, u32 r0, u32 nR,
                        const u32* eOff, u32 nSel, cudaStream_t s) {
  auto tl = std::make_unique<TileLayout>();
  const u32 N = L->info.N;
  tl->tileRows = (nR + 127u) / 128u;
  tl->tileCols = (N + 127u) / 128u;
  tl->numEntries = nSel;
  if (nSel == 0) { L->tl = std::move(tl); return; }
  const int ctBits = bits_for(tl->tileCols), trBits = bits_for(tl->tileRows);
  DevBuf<u64> kA(nSel), kB(nSel);
  DevBuf<u32> vA(nSel), vB(nSel);
  k_tile_keys<<<grid_for((size_t)nR * 32), 256, 0, s>>>(d_R, r0, nR, d_rowOff, d_colIdx, eOff, ctBits, kA.get(), vA.get());
  SB_LAUNCH_CHECK();
  const int w = radix_sort_pairs<u64>(kA.get(), kB.get(), vA.get(), vB.get(), nSel, 0, 14 + ctBits + trBits, s);
  const u64* keys = w ? kB.get() : kA.get();
  u32* vals = w ? vB.get() : vA.get();
  DevBuf<u32> head(nSel), tid((size_t)nSel + 1);
  k_tile_heads<<<grid_for(nSel), 256, 0, s>>>(keys, nSel, head.get());
  SB_LAUNCH_CHECK();
  SB_CUDA(cudaMemcpyAsync(tid.get(), head.get(), (size_t)nSel * 4, cudaMemcpyDeviceToDevice, s));
  SB_CUDA(cudaMemsetAsync(tid.get() + nSel, 0, 4, s));
  exclusive_scan_u32(tid.get(), tid.get(), (size_t)nSel + 1, s);
  tl->numTiles = read_u32(tid.get() + nSel, s);
  tl->tiles.alloc(tl->numTiles, true);
  tl->rowMeta.alloc((size_t)tl->numTiles * 640u, true);
  k_tile_records<<<grid_for(nSel), 256, 0, s>>>(keys, head.get(), tid.get(), nSel, ctBits, tl->tiles.get());
  SB_LAUNCH_CHECK();
  k_tile_meta<<<tl->numTiles, 128, 0, s>>>(keys, tl->tiles.get(), tl->numTiles, nSel, tl->rowMeta.get());
  SB_LAUNCH_CHECK();
  // the sorted payload IS the CSR index list; keep it
  tl->idx.alloc(nSel, true);
  SB_CUDA(cudaMemcpyAsync(tl->idx.get(), vals, (size_t)nSel * 4, cudaMemcpyDeviceToDevice, s));
  SB_CUDA(cudaStreamSynchronize(s));
  build_quads(*tl);
  L->tl = std::move(tl);
}

namespace {
__global__ void __launch_bounds__(128) k_tile_meta_t(const u32* __restrict__ rowMeta, const uint4* __restrict__ tiles,
                                                     uint2* __restrict__ out) {
  const u32 t = blockIdx.x, r = threadIdx.x;
  const u32* m = rowMeta + (size_t)t * 640u + r * 5u;
  u32 off = m[4] - tiles[t].z;
#pragma unroll
  for (u32 cq = 0; cq < 4; ++cq) {
    out[((size_t)t * 4u + cq) * 128u + r] = make_uint2(m[cq], off);
    off += __popc(m[cq]);
  }
}
}  // namespace

void build_quads(TileLayout& T) {
  T.numQuads = 0;
  if (!T.numTiles) return;
  T.rowMetaT.alloc((size_t)T.numTiles * 512u, true);
  k_tile_meta_t<<<T.numTiles, 128>>>(T.rowMeta.get(), T.tiles.get(), T.rowMetaT.get());
  SB_LAUNCH_CHECK();
  SB_CUDA(cudaDeviceSynchronize());
  std::vector<uint4> tiles(T.numTiles);
  SB_CUDA(cudaMemcpy(tiles.data(), T.tiles.get(), (size_t)T.numTiles * sizeof(uint4), cudaMemcpyDeviceToHost));
  // tiles are sorted by (row, col); quads in order of first appearance keyed by (row/2, col/2)
  std::vector<std::pair<u64, u32>> keyed(T.numTiles);
  for (u32 i = 0; i < T.numTiles; ++i) keyed[i] = {((u64)(tiles[i].x >> 1) << 32) | (tiles[i].y >> 1), i};
  std::sort(keyed.begin(), keyed.end());
  std::vector<uint2> quads;
  std::vector<u32> members;
  for (u32 i = 0; i < T.numTiles; ++i) {
    if (i == 0 || keyed[i].first != keyed[i - 1].first) {
      quads.push_back(make_uint2((u32)(keyed[i].first >> 32), (u32)(keyed[i].first & 0xFFFFFFFFu)));
      members.insert(members.end(), 4, 0xFFFFFFFFu);
    }
    const uint4& t = tiles[keyed[i].second];
    members[(quads.size() - 1) * 4 + (t.x & 1u) * 2u + (t.y & 1u)] = keyed[i].second;
  }
  T.numQuads = (u32)quads.size();
  T.quads.alloc(quads.size(), true);
  T.quadTiles.alloc(members.size(), true);
  SB_CUDA(cudaMemcpy(T.quads.get(), quads.data(), quads.size() * sizeof(uint2), cudaMemcpyHostToDevice));
  SB_CUDA(cudaMemcpy(T.quadTiles.get(), members.data(), members.size() * 4, cudaMemcpyHostToDevice));
}

namespace {
__global__ void k_col_degree(const u32* __restrict__ col, size_t n, u32* __restrict__ deg) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    atomicAdd(deg + col[i], 1u);
}
__global__ void k_deg_to_flag(const u32* __restrict__ deg, u32 n, u32* __restrict__ flag) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    flag[i] = deg[i] ? 1u : 0u;
}
__global__ void k_compact_flagged(const u32* __restrict__ flag, const u32* __restrict__ ex, u32 n, u32* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    if (flag[i]) out[ex[i]] = (u32)i;
}
__global__ void k_flag_tail_cols(u32* __restrict__ col, size_t n, const u32* __restrict__ deg, u32 minDegree) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const u32 c = col[i];
    if (deg[c] < minDegree) col[i] = c | 0x80000000u;
  }
}
}  // namespace

const SuperPanelLayout* ensure_superpanels(const bsmr_layout* L, u32 G, u32 hubBudget, cudaStream_t s) {
  if (L->info.N >= 0x80000000u) hubBudget = 0;  // bit 31 of a column index is not free
  const u64 key = (u64)G | ((u64)hubBudget << 32);
  {
    auto it = L->sp.find(key);
    if (it != L->sp.end()) return it->second.get();
  }
  TempScope tempScope(s);
  const bsmr_layout_info& I = L->info;
  auto sp = std::make_unique<SuperPanelLayout>();
  sp->G = G;
  sp->rows = G * kPanel;
  const u32 P = I.numRowPanels;
  sp->numSp = (P + G - 1) / G;
  sp->numEntries = I.numSparseValues;
  const u32 n = I.numSparseValues;
  if (n == 0 || P == 0) {
    sp->numWork = 0;
    return (L->sp[key] = std::move(sp)).get();
  }
  const int rowBits = bits_for(sp->rows - 1), colBits = bits_for(I.N), spBits = bits_for(sp->numSp);
  const u32* vOff = L->arr[BSMR_SPARSE_VALUE_OFFSETS].get();
  DevBuf<u64> kA(n), kB(n);
  DevBuf<u32> vA(n), vB(n);
  k_sp_keys<<<grid_for((size_t)P * 32), 256, 0, s>>>(vOff, L->arr[RPHM_SPARSE_COL_INDICES].get(),
                                                    L->arr[RPHM_SPARSE_RELATIVE_ROWS].get(), P, G, rowBits, colBits,
                                                    kA.get(), vA.get());
  SB_LAUNCH_CHECK();
  const int w = radix_sort_pairs<u64>(kA.get(), kB.get(), vA.get(), vB.get(), n, 0, rowBits + colBits + spBits, s);
  // 16 slack entries: the residual kernel reads metadata in aligned blocks of 8 that may straddle the end
  sp->col.alloc((size_t)n + 16, true);
  sp->idx.alloc((size_t)n + 16, true);
  sp->row.alloc((size_t)n + 16, true);
  SB_CUDA(cudaMemsetAsync(sp->col.get() + n, 0xFF, 16 * sizeof(u32), s));
  SB_CUDA(cudaMemsetAsync(sp->idx.get() + n, 0, 16 * sizeof(u32), s));
  SB_CUDA(cudaMemsetAsync(sp->row.get() + n, 0, 16 * sizeof(unsigned short), s));
  DevBuf<u32> nRuns(1);
  SB_CUDA(cudaMemsetAsync(nRuns.get(), 0, 4, s));
  k_sp_unpack<<<grid_for(n), 256, 0, s>>>(w ? kB.get() : kA.get(), w ? vB.get() : vA.get(),
                                         L->arr[RPHM_SPARSE_VALUES].get(), n, rowBits, colBits, sp->col.get(),
                                         sp->row.get(), sp->idx.get(), nRuns.get());
  SB_LAUNCH_CHECK();
  sp->numRuns = read_u32(nRuns.get(), s);
  // column degrees over the residual entries -> the list of referenced columns (and the residency classes below)
  DevBuf<u32> deg(I.N);
  SB_CUDA(cudaMemsetAsync(deg.get(), 0, (size_t)I.N * 4, s));
  k_col_degree<<<grid_for(n), 256, 0, s>>>(sp->col.get(), n, deg.get());
  SB_LAUNCH_CHECK();
  {
    DevBuf<u32> flag(I.N), ex((size_t)I.N + 1);
    k_deg_to_flag<<<grid_for(I.N), 256, 0, s>>>(deg.get(), I.N, flag.get());
    SB_LAUNCH_CHECK();
    scan_counts(flag.get(), ex.get(), I.N, s);
    sp->numUsedCols = read_u32(ex.get() + I.N, s);
    sp->usedCols.alloc(sp->numUsedCols ? sp->numUsedCols : 1u, true);
    k_compact_flagged<<<grid_for(I.N), 256, 0, s>>>(flag.get(), ex.get(), I.N, sp->usedCols.get());
    SB_LAUNCH_CHECK();
    SB_CUDA(cudaStreamSynchronize(s));
  }
  if (hubBudget) {
    // residency classes: the degree threshold that keeps at most hubBudget columns, then the flag (bit 31) on every
    // entry of a tail column
    std::vector<u32> h(I.N);
    SB_CUDA(cudaMemcpyAsync(h.data(), deg.get(), (size_t)I.N * 4, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaStreamSynchronize(s));
    constexpr u32 kCap = 1u << 20;
    std::vector<u64> cnt(kCap + 1, 0), ent(kCap + 1, 0);
    for (u32 d : h) {
      const u32 b = d < kCap ? d : kCap;
      cnt[b] += 1;
      ent[b] += d;
    }
    u64 cols = 0, entries = 0;
    u32 minDeg = kCap + 1;
    for (u32 d = kCap; d >= 2; --d) {  // a column referenced once has nothing to gain from residency
      if (cols + cnt[d] > hubBudget) break;
      cols += cnt[d];
      entries += ent[d];
      minDeg = d;
    }
    sp->hubCols = (u32)cols;
    sp->hubEntries = entries;
    sp->hubMinDegree = minDeg;
    sp->colMask = 0x7FFFFFFFu;
    k_flag_tail_cols<<<grid_for(n), 256, 0, s>>>(sp->col.get(), n, deg.get(), minDeg);
    SB_LAUNCH_CHECK();
    SB_CUDA(cudaStreamSynchronize(s));
  }
  // segment size: 32768 entries, shrunk for small problems so that there are several CTAs per SM
  {
    const u32 target = (u32)device_sm_count() * 4u;
    u32 sl = n / (target ? target : 1);
    if (sl > 32768u) sl = 32768u;
    if (sl < 2048u) sl = 2048u;
    sp->segLen = (sl + 1023u) & ~1023u;
  }
  sp->off.alloc((size_t)sp->numSp + 1, true);
  DevBuf<u32> segCnt(sp->numSp), segOff((size_t)sp->numSp + 1);
  k_sp_offsets<<<grid_for((size_t)sp->numSp + 1), 256, 0, s>>>(vOff, P, G, sp->numSp, sp->segLen, sp->off.get(),
                                                              segCnt.get());
  SB_LAUNCH_CHECK();
  scan_counts(segCnt.get(), segOff.get(), sp->numSp, s);
  sp->numWork = read_u32(segOff.get() + sp->numSp, s);
  if (sp->numWork) {
    const u32 nw = sp->numWork;
    DevBuf<uint2> work(nw), sorted;
    sorted.alloc(nw, true);  // becomes the layout's work list
    DevBuf<u32> k1(nw), k2(nw), i1(nw), i2(nw);
    k_sp_work<<<grid_for((size_t)sp->numSp * 32), 256, 0, s>>>(segOff.get(), sp->numSp, sp->segLen, work.get(), k1.get());
    SB_LAUNCH_CHECK();
    iota<u32>(i1.get(), nw, 0u, s);
    const int ws = radix_sort_pairs<u32>(k1.get(), k2.get(), i1.get(), i2.get(), nw, 0, 24, s);  // segment-major
    k_gather_work<<<grid_for(nw), 256, 0, s>>>(work.get(), ws ? i2.get() : i1.get(), nw, sorted.get());
    SB_LAUNCH_CHECK();
    SB_CUDA(cudaStreamSynchronize(s));
    sp->work = std::move(sorted);
  }
  SB_CUDA(cudaStreamSynchronize(s));
  return (L->sp[key] = std::move(sp)).get();
}

// ---- columns referenced by the layout (dense-block columns + residual columns), ascending
namespace {
__global__ void k_flag_cols(const u32* __restrict__ cols, size_t n, u32 N, u32* __restrict__ flag) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const u32 c = cols[i];
    if (c < N) flag[c] = 1u;  // denseCols pads with N
  }
}
}  // namespace

// the distinct values < N of up to two u32 lists, ascending, into `out`; returns their number
u32 distinct_values_dev(const u32* a, size_t na, const u32* b, size_t nb, u32 N, DevBuf<u32>& out, cudaStream_t s) {
  TempScope tempScope(s);
  const u32 n = N ? N : 1u;
  DevBuf<u32> flag(n), ex((size_t)n + 1);
  SB_CUDA(cudaMemsetAsync(flag.get(), 0, (size_t)n * 4, s));
  if (na) k_flag_cols<<<grid_for(na), 256, 0, s>>>(a, na, N, flag.get());
  if (nb) k_flag_cols<<<grid_for(nb), 256, 0, s>>>(b, nb, N, flag.get());
  SB_LAUNCH_CHECK();
  scan_counts(flag.get(), ex.get(), n, s);
  const u32 cnt = read_u32(ex.get() + n, s);
  out.alloc(cnt ? cnt : 1u, true);
  k_compact_flagged<<<grid_for(n), 256, 0, s>>>(flag.get(), ex.get(), n, out.get());
  SB_LAUNCH_CHECK();
  SB_CUDA(cudaStreamSynchronize(s));
  return cnt;
}

const bsmr_layout::HostRefs* ensure_host_refs(const bsmr_layout* L, cudaStream_t s) {
  if (L->hostRefs) return L->hostRefs.get();
  const bsmr_layout_info& I = L->info;
  auto h = std::make_unique<bsmr_layout::HostRefs>();
  h->numCols = distinct_values_dev(L->arr[BSMR_DENSE_COLS].get(), L->arr[BSMR_DENSE_COLS].size(),
                                   L->arr[RPHM_SPARSE_COL_INDICES].get(), I.numSparseValues, I.N, h->cols, s);
  return (L->hostRefs = std::move(h)).get();
}

// ---- row-stream residual layout (K7c): residual entries sorted by (reordered row position, column)
namespace {
__global__ void __launch_bounds__(256) k_st_keys(const u32* __restrict__ vOff, const u32* __restrict__ sCols,
                                                 const u32* __restrict__ sRows, u32 P, int colBits,
                                                 u64* __restrict__ keys, u32* __restrict__ vals) {
  const u32 gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const u32 nw = (gridDim.x * blockDim.x) >> 5;
  for (u32 p = gw; p < P; p += nw) {
    const u32 b = vOff[p], e = vOff[p + 1];
    for (u32 i = b + lane; i < e; i += 32) {
      keys[i] = ((u64)(p * 16u + sRows[i]) << colBits) | sCols[i];
      vals[i] = i;
    }
  }
}
__global__ void k_st_unpack(const u64* __restrict__ keys, const u32* __restrict__ vals, const u32* __restrict__ sVals,
                            const u32* __restrict__ R, size_t n, int colBits, u32* __restrict__ row,
                            u32* __restrict__ col, u32* __restrict__ idx) {
  const u64 colMask = (((u64)1) << colBits) - 1;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const u64 k = keys[i];
    row[i] = R[(u32)(k >> colBits)];
    col[i] = (u32)(k & colMask);
    idx[i] = sVals[vals[i]];
  }
}
}  // namespace

const StreamLayout* ensure_stream(const bsmr_layout* L, cudaStream_t s) {
  if (L->st) return L->st.get();
  TempScope tempScope(s);
  const bsmr_layout_info& I = L->info;
  auto st = std::make_unique<StreamLayout>();
  const u32 n = I.numSparseValues, P = I.numRowPanels;
  st->numEntries = n;
  // 32 slack entries: the kernel reads metadata in whole 32-entry batches
  st->row.alloc((size_t)n + 32, true);
  st->col.alloc((size_t)n + 32, true);
  st->idx.alloc((size_t)n + 32, true);
  SB_CUDA(cudaMemsetAsync(st->row.get() + n, 0, 32 * 4, s));
  SB_CUDA(cudaMemsetAsync(st->col.get() + n, 0, 32 * 4, s));
  SB_CUDA(cudaMemsetAsync(st->idx.get() + n, 0, 32 * 4, s));
  if (n && P) {
    const int colBits = bits_for(I.N), rowBits = bits_for((u64)P * 16u);
    DevBuf<u64> kA(n), kB(n);
    DevBuf<u32> vA(n), vB(n);
    k_st_keys<<<grid_for((size_t)P * 32), 256, 0, s>>>(L->arr[BSMR_SPARSE_VALUE_OFFSETS].get(),
                                                      L->arr[RPHM_SPARSE_COL_INDICES].get(),
                                                      L->arr[RPHM_SPARSE_RELATIVE_ROWS].get(), P, colBits, kA.get(),
                                                      vA.get());
    SB_LAUNCH_CHECK();
    const int w = radix_sort_pairs<u64>(kA.get(), kB.get(), vA.get(), vB.get(), n, 0, colBits + rowBits, s);
    k_st_unpack<<<grid_for(n), 256, 0, s>>>(w ? kB.get() : kA.get(), w ? vB.get() : vA.get(),
                                           L->arr[RPHM_SPARSE_VALUES].get(), L->arr[BSMR_REORDERED_ROWS].get(), n,
                                           colBits, st->row.get(), st->col.get(), st->idx.get());
    SB_LAUNCH_CHECK();
  }
  SB_CUDA(cudaStreamSynchronize(s));
  L->st = std::move(st);
  return L->st.get();
}

// ---- on-disk layout cache: reordering costs orders of magnitude more than one SDDMM pass,
// so a deployment keeps the layout keyed by (matrix, alpha, delta, block_size).  Versioned little-endian file:
//   "BSMRLAY1" | u32 version | bsmr_layout_info | u32 sparseChunk | u32 numDenseWork | u32 numSparseWork |
//   15 x (u64 length, u32 data[length]) | uint2 denseWork[] | uint2 sparseWork[]
namespace {
constexpr char kLayoutMagic[8] = {'B', 'S', 'M', 'R', 'L', 'A', 'Y', '1'};
struct File {
  FILE* f;
  File(const char* p, const char* m) : f(std::fopen(p, m)) {}
  ~File() { if (f) std::fclose(f); }
};
template <typename T>
void put(FILE* f, const T* p, size_t n) {
  if (n && std::fwrite(p, sizeof(T), n, f) != n) fail(SDDMM_E_ARG, "layout cache: short write");
}
template <typename T>
void get(FILE* f, T* p, size_t n) {
  if (n && std::fread(p, sizeof(T), n, f) != n) fail(SDDMM_E_ARG, "layout cache: short read / truncated file");
}
}  // namespace

void layout_save(const bsmr_layout* L, const char* path) {
  File fl(path, "wb");
  if (!fl.f) fail(SDDMM_E_ARG, "layout cache: cannot open %s for writing", path);
  const u32 version = 1;
  put(fl.f, kLayoutMagic, 8);
  put(fl.f, &version, 1);
  put(fl.f, &L->info, 1);
  put(fl.f, &L->sparseChunk, 1);
  put(fl.f, &L->numDenseWork, 1);
  put(fl.f, &L->numSparseWork, 1);
  std::vector<u32> host;
  for (int i = 0; i < BSMR_ARRAY_COUNT; ++i) {
    const u64 n = L->arr[i].size();
    put(fl.f, &n, 1);
    host.resize(n);
    if (n) SB_CUDA(cudaMemcpy(host.data(), L->arr[i].get(), n * 4, cudaMemcpyDeviceToHost));
    put(fl.f, host.data(), n);
  }
  std::vector<uint2> w(L->numDenseWork);
  if (L->numDenseWork) SB_CUDA(cudaMemcpy(w.data(), L->denseWork.get(), (size_t)L->numDenseWork * 8, cudaMemcpyDeviceToHost));
  put(fl.f, w.data(), w.size());
  w.resize(L->numSparseWork);
  if (L->numSparseWork) SB_CUDA(cudaMemcpy(w.data(), L->sparseWork.get(), (size_t)L->numSparseWork * 8, cudaMemcpyDeviceToHost));
  put(fl.f, w.data(), w.size());
  // optional full-tile plan
  const u32 hasTl = L->tl ? 1u : 0u;
  put(fl.f, &hasTl, 1);
  if (hasTl) {
    const TileLayout& T = *L->tl;
    const u32 hdr[4] = {T.numTiles, T.numEntries, T.tileRows, T.tileCols};
    put(fl.f, hdr, 4);
    std::vector<uint4> tiles(T.numTiles);
    if (T.numTiles) SB_CUDA(cudaMemcpy(tiles.data(), T.tiles.get(), (size_t)T.numTiles * 16, cudaMemcpyDeviceToHost));
    put(fl.f, tiles.data(), tiles.size());
    host.resize((size_t)T.numTiles * 640u);
    if (T.numTiles) SB_CUDA(cudaMemcpy(host.data(), T.rowMeta.get(), host.size() * 4, cudaMemcpyDeviceToHost));
    put(fl.f, host.data(), host.size());
    host.resize(T.numEntries);
    if (T.numEntries) SB_CUDA(cudaMemcpy(host.data(), T.idx.get(), host.size() * 4, cudaMemcpyDeviceToHost));
    put(fl.f, host.data(), host.size());
  }
}

bsmr_layout* layout_load(const char* path) {
  File fl(path, "rb");
  if (!fl.f) fail(SDDMM_E_ARG, "layout cache: cannot open %s", path);
  char magic[8];
  u32 version = 0;
  get(fl.f, magic, 8);
  get(fl.f, &version, 1);
  if (std::memcmp(magic, kLayoutMagic, 8) != 0 || version != 1) fail(SDDMM_E_ARG, "layout cache: %s is not a version-1 layout file", path);
  auto* L = new bsmr_layout();
  try {
    SB_CUDA(cudaGetDevice(&L->device));
    get(fl.f, &L->info, 1);
    get(fl.f, &L->sparseChunk, 1);
    get(fl.f, &L->numDenseWork, 1);
    get(fl.f, &L->numSparseWork, 1);
    // Every array is read to the host first and cross-checked before anything reaches the device: a truncated,
    // stale or foreign file must end in SDDMM_E_ARG, never in out-of-bounds reads inside the kernels.
    std::vector<std::vector<u32>> H(BSMR_ARRAY_COUNT);
    for (int i = 0; i < BSMR_ARRAY_COUNT; ++i) {
      u64 n = 0;
      get(fl.f, &n, 1);
      if (n > ((u64)1 << 34)) fail(SDDMM_E_ARG, "layout cache: implausible array length");
      H[i].resize(n);
      get(fl.f, H[i].data(), n);
    }
    std::vector<uint2> wd(L->numDenseWork <= ((u32)1 << 30) ? L->numDenseWork : 0), ws;
    if (wd.size() != L->numDenseWork) fail(SDDMM_E_ARG, "layout cache: implausible work-list length");
    get(fl.f, wd.data(), wd.size());
    if (L->numSparseWork > ((u32)1 << 30)) fail(SDDMM_E_ARG, "layout cache: implausible work-list length");
    ws.resize(L->numSparseWork);
    get(fl.f, ws.data(), ws.size());
    {
      const bsmr_layout_info& I = L->info;
      const u64 P = I.numRowPanels;
      auto bad = [](const char* what) { fail(SDDMM_E_ARG, "layout cache: inconsistent file (%s)", what); };
      auto len = [&](bsmr_array_id id) { return (u64)H[id].size(); };
      auto monotone = [&](bsmr_array_id id, u64 last, const char* what) {
        const auto& v = H[id];
        if (v.size() != P + 1 || v[0] != 0 || v[P] != last) bad(what);
        for (u64 i = 0; i < P; ++i) if (v[i] > v[i + 1]) bad(what);
      };
      auto below = [&](bsmr_array_id id, u64 limit, bool allowNull, const char* what) {
        for (u32 x : H[id]) if (x >= limit && !(allowNull && x == kNull)) bad(what);
      };
      if (P != ((u64)I.numRows + kPanel - 1) / kPanel || len(BSMR_REORDERED_ROWS) != I.numRows) bad("row count");
      below(BSMR_REORDERED_ROWS, I.M, false, "reorderedRows");
      monotone(BSMR_DENSE_COL_OFFSETS, len(BSMR_DENSE_COLS), "denseColOffsets");
      monotone(BSMR_SPARSE_COL_OFFSETS, len(BSMR_SPARSE_COLS), "sparseColOffsets");
      monotone(BSMR_SPARSE_VALUE_OFFSETS, I.numSparseValues, "sparseValueOffsets");
      monotone(RPHM_BLOCK_OFFSETS, I.numDenseBlocks, "blockOffsets");
      if (len(RPHM_SPARSE_VALUES) != I.numSparseValues || len(RPHM_SPARSE_RELATIVE_ROWS) != I.numSparseValues ||
          len(RPHM_SPARSE_COL_INDICES) != I.numSparseValues || len(RPHM_BLOCK_VALUES) != (u64)I.numDenseBlocks * 256u ||
          len(BSMR_DENSE_COLS) != (u64)I.numDenseBlocks * 16u)
        bad("array lengths");
      for (u64 p = 0; p <= P; ++p)
        if ((u64)H[RPHM_BLOCK_OFFSETS][p] * 16u != H[BSMR_DENSE_COL_OFFSETS][p]) bad("block / dense column offsets");
      below(BSMR_DENSE_COLS, (u64)I.N + 1, false, "denseCols");
      below(BSMR_SPARSE_COLS, (u64)I.N + 1, false, "sparseCols");
      below(RPHM_SPARSE_COL_INDICES, I.N, false, "sparseColIndices");
      below(RPHM_SPARSE_RELATIVE_ROWS, kPanel, false, "sparseRelativeRows");
      below(RPHM_SPARSE_VALUES, I.nnz, false, "sparseValues");
      below(RPHM_BLOCK_VALUES, I.nnz, true, "blockValues");
      if (len(RPHM_DENSE_ROW_PANEL_IDS) != I.numDenseThreadBlocks || len(RPHM_DENSE_COL_BLOCK_ITERS) != I.numDenseThreadBlocks ||
          len(RPHM_SPARSE_ROW_PANEL_IDS) != I.numSparseThreadBlocks || len(RPHM_SPARSE_COL_BLOCK_ITERS) != I.numSparseThreadBlocks)
        bad("reference work lists");
      below(RPHM_DENSE_ROW_PANEL_IDS, P, false, "denseRowPanelIds");
      below(RPHM_SPARSE_ROW_PANEL_IDS, P, false, "sparseRowPanelIds");
      if (L->sparseChunk < 32 || L->sparseChunk > 65536) bad("sparseChunk");
      for (const uint2& w : wd)
        if (w.x >= P || w.y >= H[RPHM_BLOCK_OFFSETS][w.x + 1] - H[RPHM_BLOCK_OFFSETS][w.x]) bad("dense work list");
      for (const uint2& w : ws)
        if (w.x >= P || w.y >= H[BSMR_SPARSE_VALUE_OFFSETS][w.x + 1] - H[BSMR_SPARSE_VALUE_OFFSETS][w.x]) bad("residual work list");
      if ((u64)I.numDenseValues + I.numSparseValues > I.nnz) bad("entry counts");
    }
    for (int i = 0; i < BSMR_ARRAY_COUNT; ++i) {
      const u64 n = H[i].size();
      L->arr[i].alloc(n ? n : 1, true);
      L->arr[i].n = n;
      if (n) SB_CUDA(cudaMemcpy(L->arr[i].get(), H[i].data(), n * 4, cudaMemcpyHostToDevice));
    }
    L->denseWork.alloc(L->numDenseWork ? L->numDenseWork : 1, true);
    if (L->numDenseWork) SB_CUDA(cudaMemcpy(L->denseWork.get(), wd.data(), wd.size() * 8, cudaMemcpyHostToDevice));
    L->sparseWork.alloc(L->numSparseWork ? L->numSparseWork : 1, true);
    if (L->numSparseWork) SB_CUDA(cudaMemcpy(L->sparseWork.get(), ws.data(), ws.size() * 8, cudaMemcpyHostToDevice));
    std::vector<u32> host;
    u32 hasTl = 0;
    get(fl.f, &hasTl, 1);
    if (hasTl) {
      auto T = std::make_unique<TileLayout>();
      u32 hdr[4];
      get(fl.f, hdr, 4);
      T->numTiles = hdr[0]; T->numEntries = hdr[1]; T->tileRows = hdr[2]; T->tileCols = hdr[3];
      if ((u64)T->numTiles > ((u64)1 << 30)) fail(SDDMM_E_ARG, "layout cache: implausible tile count");
      std::vector<uint4> tiles(T->numTiles);
      get(fl.f, tiles.data(), tiles.size());
      T->tiles.alloc(T->numTiles ? T->numTiles : 1, true);
      if (T->numTiles) SB_CUDA(cudaMemcpy(T->tiles.get(), tiles.data(), tiles.size() * 16, cudaMemcpyHostToDevice));
      host.resize((size_t)T->numTiles * 640u);
      get(fl.f, host.data(), host.size());
      T->rowMeta.alloc(host.size() ? host.size() : 1, true);
      if (!host.empty()) SB_CUDA(cudaMemcpy(T->rowMeta.get(), host.data(), host.size() * 4, cudaMemcpyHostToDevice));
      if ((u64)T->numEntries > L->info.nnz) fail(SDDMM_E_ARG, "layout cache: inconsistent file (tile entries)");
      for (const uint4& t : tiles)
        if (t.x >= T->tileRows || t.y >= T->tileCols || (u64)t.z + t.w > T->numEntries || t.w > 16384u)
          fail(SDDMM_E_ARG, "layout cache: inconsistent file (tile records)");
      host.resize(T->numEntries);
      get(fl.f, host.data(), host.size());
      for (u32 x : host)
        if (x >= L->info.nnz) fail(SDDMM_E_ARG, "layout cache: inconsistent file (tile CSR indices)");
      T->idx.alloc(host.size() ? host.size() : 1, true);
      if (!host.empty()) SB_CUDA(cudaMemcpy(T->idx.get(), host.data(), host.size() * 4, cudaMemcpyHostToDevice));
      build_quads(*T);
      L->tl = std::move(T);
    }
    return L;
  } catch (...) {
    delete L;
    throw;
  }
}

}  // namespace sb
