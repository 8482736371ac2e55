// reorder_rows.cu -- row-similarity reordering (a2-a6 of SURVEY.md 8a) as sm_100a kernels.
//
//   K1  encode      : per row, SPARSE histogram over column blocks (block id, count) + dispersion.
//                     Replaces kernel::calculateDispersion (src/rowReordering.cu:49-93), which
//                     writes a dense M x nbpr matrix.
//   K3  radix sorts : (dispersion, row) and (cluster, position), stable (rowReordering.cu:1060-1062,
//                     :986-990; thrust::host stable sorts in the reference).
//   K2  clustering  : one persistent cooperative kernel; same sequential semantics as the chain of
//                     1-CTA bsa_clustering kernels (rowReordering.cu:325-432) and the same fp32
//                     reduction tree (cudaUtil.cuh:13-45), evaluated on
//                     sparse histograms.
//
// Bit-exactness strategy for `sim > alpha`:
//   the reference value is  REDUCE(sum min(a_i,b_i)) / REDUCE(sum max(a_i,b_i))  in fp32 with a fixed
//   tree.  Over the index set the tree keeps, sum max = S_a + S_b - sum min in real arithmetic, so a
//   cheap fp32 estimate  m / (S_a + S_b - m)  from the intersection only is within a few ulp-sums of
//   the reference value (all terms are non-negative: relative error <= ~1e-4 worst case).  Pairs whose
//   estimate is farther than kTolRel from alpha are decided by the estimate; the (rare) others are
//   re-evaluated with the literal tree, using round-to-nearest intrinsics only.
#include <cooperative_groups.h>

#include <vector>

#include "primitives.cuh"
#include "reorder_rows.cuh"

namespace cg = cooperative_groups;

namespace sb {

// ------------------------------------------------------------------------------------------
// host-side exact restatements of the reference's launch geometry
// ------------------------------------------------------------------------------------------
u32 calc_block_size(u32 M, u32 N, u64 freeMem) {  // rowReordering.cu:1009-1025
  const float g = std::ceil((float)((size_t)M * (size_t)M * sizeof(u32)) / (float)(freeMem / 2));
  const float sm = std::ceil((float)((size_t)N * sizeof(u32)) / (float)(49152u / 2u));
  const u32 a = (u32)g, b = (u32)sm;
  const u32 bs = a > b ? a : b;
  return bs > 16 ? bs : 16;
}
u32 num_blocks_per_row(u32 N, u32 bs) { return (u32)(int)std::ceil((float)N / (float)bs); }  // :1035
u32 cluster_blockdim(u32 nbpr) {                                                              // :911-920
  if (nbpr < 32) return 32;
  int cand = (int)(32 * std::ceil((float)((int)nbpr / 4) / (float)32));
  cand = cand > 32 ? cand : 32;
  return (u32)(1024 < cand ? 1024 : cand);
}
u32 kept_warp_mask(u32 B) {  // cudaUtil.cuh:37-43: which per-warp partials reach shm[0]
  const u32 W = B / 32;
  u32 set[32];
  for (u32 w = 0; w < W; ++w) set[w] = 1u << w;
  for (u32 stride = B / 64; stride >= 1; stride >>= 1)
    for (u32 w = 0; w < stride; ++w) set[w] |= set[w + stride];
  return set[0];
}

// ------------------------------------------------------------------------------------------
// K1: encode.  One warp per row; columns of a row must be ascending (checked, else a sorted
// copy is made first).  Run-length encodes col / block_size with ballot/popc.
// ------------------------------------------------------------------------------------------
static __global__ void k_check_rows_sorted(const u32* __restrict__ rowOff, const u32* __restrict__ colIdx, u32 M,
                                           u32* __restrict__ unsortedFlag) {
  const u32 gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const u32 nw = (gridDim.x * blockDim.x) >> 5;
  for (u32 r = gw; r < M; r += nw) {
    const u32 b = rowOff[r], e = rowOff[r + 1];
    bool bad = false;
    for (u32 i = b + 1 + lane; i < e; i += 32) bad |= colIdx[i] <= colIdx[i - 1];
    if (__any_sync(0xffffffffu, bad)) {
      if (lane == 0) *unsortedFlag = 1;
    }
  }
}

static __global__ void k_make_row_col_keys(const u32* __restrict__ rowOff, const u32* __restrict__ colIdx, u32 M,
                                           u64* __restrict__ keys) {
  const u32 gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const u32 nw = (gridDim.x * blockDim.x) >> 5;
  for (u32 r = gw; r < M; r += nw) {
    const u32 b = rowOff[r], e = rowOff[r + 1];
    for (u32 i = b + lane; i < e; i += 32) keys[i] = ((u64)r << 32) | colIdx[i];
  }
}
static __global__ void k_keys_low32(const u64* __restrict__ keys, u32* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = (u32)keys[i];
}

__device__ __forceinline__ bool blk_kept(u32 blk, u32 B, u32 keptMask) { return (keptMask >> ((blk % B) >> 5)) & 1u; }

// Walks one row's sorted columns and calls emit(blockId, count, ordinal, keptOrdinal, isKept) once per
// distinct block, in ascending block order (ordinal = rank among all blocks of the row, keptOrdinal =
// rank among the blocks the reference's reduction tree keeps).  nbAll / nbKept are returned on all lanes.
template <typename Emit>
__device__ __forceinline__ void warp_rle_blocks(const u32* __restrict__ cols, u32 len, u32 bs, u32 B, u32 keptMask,
                                                u32& nbAll, u32& nbKept, Emit emit) {
  const u32 lane = threadIdx.x & 31;
  u32 carryBlk = kNull, carryCnt = 0;
  nbAll = 0;
  nbKept = 0;
  for (u32 base = 0; base < len; base += 32) {
    const u32 i = base + lane;
    const bool valid = i < len;
    const u32 blk = valid ? cols[i] / bs : kNull;
    u32 prev = __shfl_up_sync(0xffffffffu, blk, 1);
    if (lane == 0) prev = carryBlk;
    const bool head = valid && blk != prev;
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    const u32 nvalid = __popc(__ballot_sync(0xffffffffu, valid));
    if (heads == 0) {
      carryCnt += nvalid;
      continue;
    }
    const u32 firstHead = __ffs(heads) - 1;
    if (carryBlk != kNull) {  // the carried run closes at the first head of this chunk
      const bool ck = blk_kept(carryBlk, B, keptMask);
      if (lane == 0) emit(carryBlk, carryCnt + firstHead, nbAll, nbKept, ck);
      nbAll += 1;
      nbKept += ck;
    }
    const u32 lastHead = 31 - __clz(heads);
    const unsigned closedHeads = heads & ~(1u << lastHead);  // every head run but the last closes here
    const bool isClosed = head && lane != lastHead;
    const bool kh = isClosed && blk_kept(blk, B, keptMask);
    const unsigned keptHeads = __ballot_sync(0xffffffffu, kh);
    if (isClosed) {
      const unsigned rest = heads >> (lane + 1);  // non-zero: a later head exists
      const unsigned lt = (1u << lane) - 1u;
      emit(blk, (u32)__ffs(rest), nbAll + __popc(closedHeads & lt), nbKept + __popc(keptHeads & lt), kh);
    }
    nbAll += __popc(closedHeads);
    nbKept += __popc(keptHeads);
    carryBlk = __shfl_sync(0xffffffffu, blk, lastHead);
    carryCnt = nvalid - lastHead;
  }
  if (carryBlk != kNull) {
    const bool ck = blk_kept(carryBlk, B, keptMask);
    if (lane == 0) emit(carryBlk, carryCnt, nbAll, nbKept, ck);
    nbAll += 1;
    nbKept += ck;
  }
}

// pass 1: dispersion (rowReordering.cu:78-92, exact uint32 arithmetic) and #kept entries per row
static __global__ void __launch_bounds__(256) k_encode_count(const u32* __restrict__ rowOff,
                                                             const u32* __restrict__ cols, u32 M, u32 bs, u32 B,
                                                             u32 keptMask, u32* __restrict__ disp,
                                                             u32* __restrict__ keptCnt) {
  const u32 gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const u32 nw = (gridDim.x * blockDim.x) >> 5;
  for (u32 r = gw; r < M; r += nw) {
    const u32 b = rowOff[r], len = rowOff[r + 1] - b;
    u32 nb, kept;
    warp_rle_blocks(cols + b, len, bs, B, keptMask, nb, kept, [](u32, u32, u32, u32, bool) {});
    if (lane == 0) {
      // sum_b (bs - h_b) + nnz * nb  ==  nb*bs - nnz + nnz*nb   (mod 2^32), 0 for empty rows
      disp[r] = len ? nb * bs - len + len * nb : 0u;
      if (keptCnt) keptCnt[r] = kept;
    }
  }
}

// pass 2: write the kept (block, count) entries of row asc[pos] at encOff[pos], plus per-position
// meta {offset, length, sum of squares (uint32 wrap), sum of counts}.
static __global__ void __launch_bounds__(256) k_encode_fill(const u32* __restrict__ rowOff,
                                                            const u32* __restrict__ cols, u32 M, u32 bs, u32 B,
                                                            u32 keptMask, const u32* __restrict__ asc,
                                                            const u32* __restrict__ encOff, uint2* __restrict__ enc,
                                                            uint4* __restrict__ meta) {
  const u32 gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const u32 nw = (gridDim.x * blockDim.x) >> 5;
  for (u32 pos = gw; pos < M; pos += nw) {
    const u32 r = asc[pos];
    const u32 b = rowOff[r], len = rowOff[r + 1] - b;
    const u32 off = encOff[pos];
    uint2* out = enc + off;
    u32 ss = 0, s1 = 0, nb, kept;
    warp_rle_blocks(cols + b, len, bs, B, keptMask, nb, kept, [&](u32 blk, u32 cnt, u32, u32 ordKept, bool isKept) {
      if (isKept) {
        out[ordKept] = make_uint2(blk, cnt);
        ss += cnt * cnt;
        s1 += cnt;
      }
    });
    ss = __reduce_add_sync(0xffffffffu, ss);
    s1 = __reduce_add_sync(0xffffffffu, s1);
    if (lane == 0) meta[pos] = make_uint4(off, kept, ss, s1);
  }
}

// Bitmap signatures of the sparse encodings: bit (block & sigMask) of row `pos` is set for every kept block.
// With 32*W >= nbpr the map is the identity (an exact bitmap), otherwise a folding hash.  Used by the batched
// clustering kernel to bound the similarity from above without reading a candidate's entry list (see
// cb_sig_bound).  One warp per position; the signature array is zeroed beforehand.
static __global__ void __launch_bounds__(256) k_build_sig(const uint2* __restrict__ enc, const uint4* __restrict__ meta,
                                                          u32 M, u32 W, u32 sigMask, u32* __restrict__ sig) {
  const u32 gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const u32 nw = (gridDim.x * blockDim.x) >> 5;
  for (u32 pos = gw; pos < M; pos += nw) {
    const uint4 m = meta[pos];
    u32* out = sig + (size_t)pos * W;
    for (u32 j = lane; j < m.y; j += 32) {
      const u32 bit = enc[m.x + j].x & sigMask;
      atomicOr(out + (bit >> 5), 1u << (bit & 31u));
    }
  }
}

static __global__ void k_gather_u32(const u32* __restrict__ src, const u32* __restrict__ idx, u32* __restrict__ dst,
                                    size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[idx[i]];
}
static __global__ void k_count_zero(const u32* __restrict__ v, size_t n, u32* __restrict__ cnt) {
  u32 c = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    c += v[i] == 0;
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(cnt, c);
}
static __global__ void k_init_cid(u32* cid, size_t n, u32 zeroRows) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    cid[i] = i < zeroRows ? 0u : kNull;
}

// ------------------------------------------------------------------------------------------
// K2: clustering
// ------------------------------------------------------------------------------------------
constexpr int kClThreads = 256;
constexpr int kClWarps = kClThreads / 32;
constexpr float kTolRel = 1e-3f;
constexpr float kTolAbs = 1e-6f;

struct ClusterArgs {
  u32 M, start0, nbpr, B, keptMask;
  float alpha;
  const uint2* enc;
  const uint4* meta;  // by position: {off, len, ss, s1}
  u32* cid;           // by position
  u32* ctrl;          // [0..2] firstJoin slots, [3..5] firstReject slots, [6] exact-eval counter
  u32* numClustersOut;
};

__device__ __forceinline__ u32 ld_cg(const u32* p) { return __ldcg(p); }

// binary search of block id i in a sorted entry list; returns its count or 0
__device__ __forceinline__ u32 lookup_cnt(const uint2* __restrict__ e, u32 L, u32 i) {
  u32 lo = 0, hi = L;
  while (lo < hi) {
    const u32 mid = (lo + hi) >> 1;
    const u32 b = e[mid].x;
    if (b < i) lo = mid + 1; else hi = mid;
  }
  return (lo < L && e[lo].x == i) ? e[lo].y : 0u;
}

// Literal restatement of calculate_similarity_norm_weighted_jaccard (rowReordering.cu:235-293) +
// cuUtil::reduce_sum (cudaUtil.cuh:13-45) for ONE pair, executed by one warp that plays the B
// reference threads warp by warp.  Round-to-nearest intrinsics only: no contraction, no approx.
__device__ float exact_similarity_warp(const u32* __restrict__ rep, const uint2* __restrict__ ent, u32 L, u32 nbpr,
                                       u32 B, u32 keptMask, u32 ssRep, u32 ssCmp, u32 repStride = 1) {
  const u32 lane = threadIdx.x & 31;
  const float normRep = __fsqrt_rn(__uint2float_rn(ssRep));
  const float normCmp = __fsqrt_rn(__uint2float_rn(ssCmp));
  const u32 W = B >> 5;
  float smin = 0.f, smax = 0.f;  // lane w holds the reduced value of reference warp w
  for (u32 vw = 0; vw < W; ++vw) {
    if (!((keptMask >> vw) & 1u)) continue;  // never added into shm[0] by the reference's tree
    float pmin = 0.f, pmax = 0.f;
    for (u32 i = (vw << 5) + lane; i < nbpr; i += B) {
      const u32 r = rep[(size_t)i * repStride];
      const u32 c = lookup_cnt(ent, L, i);
      const float a = __fdiv_rn(__uint2float_rn(r), normRep);
      const float b = __fdiv_rn(__uint2float_rn(c), normCmp);
      pmin = __fadd_rn(pmin, fminf(a, b));
      pmax = __fadd_rn(pmax, fmaxf(a, b));
    }
#pragma unroll
    for (int w = 1; w < 32; w <<= 1) {
      pmin = __fadd_rn(pmin, __shfl_xor_sync(0xffffffffu, pmin, w));
      pmax = __fadd_rn(pmax, __shfl_xor_sync(0xffffffffu, pmax, w));
    }
    if (lane == vw) { smin = pmin; smax = pmax; }
  }
  for (u32 stride = B >> 6; stride >= 1; stride >>= 1) {
    const float omin = __shfl_down_sync(0xffffffffu, smin, stride);
    const float omax = __shfl_down_sync(0xffffffffu, smax, stride);
    if (lane < stride) { smin = __fadd_rn(smin, omin); smax = __fadd_rn(smax, omax); }
  }
  const float sim = __fdiv_rn(smin, smax);
  return __shfl_sync(0xffffffffu, sim, 0);
}

static __global__ void __launch_bounds__(kClThreads) k_cluster(ClusterArgs a) {
  extern __shared__ u32 rep[];  // dense accumulated histogram of the current cluster (kept blocks)
  __shared__ u32 sRed[kClWarps];
  __shared__ u32 sSsRep, sS1Rep;
  cg::grid_group grid = cg::this_grid();
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 totalWarps = gridDim.x * kClWarps;
  const u32 gw = blockIdx.x * kClWarps + warp;
  volatile u32* ctrl = a.ctrl;

  u32 cluster = 1, seed = a.start0, iter = 0;
  while (seed < a.M) {
    // ---- new cluster: rep = histogram of the seed row
    if (blockIdx.x == 0 && threadIdx.x == 0) a.cid[seed] = cluster;
    for (u32 i = threadIdx.x; i < a.nbpr; i += kClThreads) rep[i] = 0;
    __syncthreads();
    {
      const uint4 m = a.meta[seed];
      for (u32 j = threadIdx.x; j < m.y; j += kClThreads) {
        const uint2 e = a.enc[m.x + j];
        rep[e.x] = e.y;
      }
      if (threadIdx.x == 0) { sSsRep = m.z; sS1Rep = m.w; }
    }
    __syncthreads();
    u32 p = seed + 1, nextSeed = kNull;
    u32 chunk = totalWarps;
    while (p < a.M) {
      const u32 e = (a.M - p > chunk) ? p + chunk : a.M;
      const u32 slot = iter % 3;
      const u32 ssRep = sSsRep;
      const float nRepInv = ssRep ? 1.0f / sqrtf((float)ssRep) : 0.f;
      const float Sa = (float)sS1Rep * nRepInv;
      for (u32 pos = p + gw; pos < e; pos += totalWarps) {
        if (ld_cg(a.cid + pos) != kNull) continue;
        const uint4 m = a.meta[pos];
        const u32 ssCmp = m.z;
        bool join;
        if (ssRep == 0 || ssCmp == 0) {
          // rowReordering.cu:263-268: both zero -> 1.0f, one zero -> 0.0f
          const float sim = (ssRep == 0 && ssCmp == 0) ? 1.0f : 0.0f;
          join = sim > a.alpha;
        } else {
          const uint2* ent = a.enc + m.x;
          const float nCmpInv = 1.0f / sqrtf((float)ssCmp);
          float mn = 0.f;
          for (u32 j = lane; j < m.y; j += 32) {
            const uint2 en = ent[j];
            const u32 r = rep[en.x];
            mn += fminf((float)r * nRepInv, (float)en.y * nCmpInv);
          }
#pragma unroll
          for (int w = 16; w >= 1; w >>= 1) mn += __shfl_xor_sync(0xffffffffu, mn, w);
          const float den = Sa + (float)m.w * nCmpInv - mn;
          const float est = mn / den;
          const float tol = kTolRel * fabsf(a.alpha) + kTolAbs;
          if (den > 0.f && est > a.alpha + tol) join = true;
          else if (den > 0.f && est < a.alpha - tol) join = false;
          else {
            const float sim = exact_similarity_warp(rep, ent, m.y, a.nbpr, a.B, a.keptMask, ssRep, ssCmp);
            join = sim > a.alpha;
            if (lane == 0) atomicAdd(a.ctrl + 6, 1u);
          }
        }
        if (lane == 0) {
          if (join) atomicMin(a.ctrl + slot, pos);
          else if (nextSeed == kNull) atomicMin(a.ctrl + 3 + slot, pos);
        }
      }
      __threadfence();
      grid.sync();
      const u32 fj = ctrl[slot];
      const u32 fr = ctrl[3 + slot];
      if (blockIdx.x == 0 && threadIdx.x == 0) {
        const u32 rs = (iter + 2) % 3;
        ctrl[rs] = kNull;
        ctrl[3 + rs] = kNull;
      }
      // rejects before the first join are final for this cluster; the earliest one seeds the next
      if (nextSeed == kNull && fr != kNull && (fj == kNull || fr < fj)) nextSeed = fr;
      if (fj != kNull) {
        if (blockIdx.x == 0 && threadIdx.x == 0) a.cid[fj] = cluster;
        // rep += hist(fj); sum of squares updated exactly in uint32 (wraps like the reference's)
        const uint4 m = a.meta[fj];
        u32 dss = 0;
        for (u32 j = threadIdx.x; j < m.y; j += kClThreads) {
          const uint2 en = a.enc[m.x + j];
          const u32 r = rep[en.x];
          const u32 nr = r + en.y;
          dss += nr * nr - r * r;
          rep[en.x] = nr;
        }
        dss = __reduce_add_sync(0xffffffffu, dss);
        if (lane == 0) sRed[warp] = dss;
        __syncthreads();
        if (threadIdx.x == 0) {
          u32 t = 0;
          for (int w = 0; w < kClWarps; ++w) t += sRed[w];
          sSsRep += t;
          sS1Rep += m.w;
        }
        __syncthreads();
        p = fj + 1;
        chunk = totalWarps;
      } else {
        p = e;
        if (chunk < totalWarps * 16u) chunk *= 2;
      }
      ++iter;
    }
    if (nextSeed == kNull) break;
    seed = nextSeed;
    ++cluster;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.numClustersOut = cluster;
}

// ------------------------------------------------------------------------------------------
// K2b: batched clustering.  Same sequential semantics, but up to G consecutive clusters are scanned in
// ONE pass over the candidates (their representatives sit side by side in shared memory), so a
// candidate's entry list is read from HBM once per G clusters instead of once per cluster.
//
//   Equivalent formulation used here (leader clustering): a row, visited in dispersion order, joins the
//   FIRST open cluster (creation order) whose representative accepts it, where a representative holds the
//   rows that joined it earlier; a row no open cluster accepts seeds a new cluster.  Clusters c0..c0+g-1
//   of a batch are SPECULATED to be seeded by the next g unclustered positions s[0] < ... < s[g-1].
//   A window evaluates, for every unclustered position, the clusters whose seed precedes it, in order,
//   and reports the earliest position at which anything joins.  Everything before that position is
//   final (a speculated seed that was passed without being absorbed is thereby confirmed); the join is
//   applied and the scan resumes behind it.  If the joiner is itself a speculated seed s[j], clusters
//   j.. of the batch are dropped (nothing of theirs was final yet) and the batch continues with j.
// ------------------------------------------------------------------------------------------
constexpr int kCbThreads = 1024;
constexpr int kCbWarps = kCbThreads / 32;
constexpr int kCbMaxG = 8;

struct ClusterBatchArgs {
  u32 M, start0, nbpr, B, keptMask, G;
  u32 laneRows;  // != 0: rows with at most this many kept blocks are evaluated one per LANE (short-row matrices)
  float alpha;
  const uint2* enc;
  const uint4* meta;
  u32* cid;
  unsigned long long* slots;  // [3] earliest (pos << 8 | cluster-in-batch) that joins, per rotating window
  u32* accepts;               // [3] number of accepting candidates seen in the window (same rotation)
  u32* seeds;                 // [0] count, [1..8] positions of the next batch's seeds (written by block 0)
  u32* stats;                 // [0] exact evaluations, [1] clusters created, [2] windows, [3] batches,
                              // [4] lane-path redo rows, [5] candidates rejected by the signature bound
  const u32* sig;             // [M][W] bitmap signatures by position (nullptr / W == 0: filter off)
  u32 W, sigMask;
  float sigThr;               // reject when the upper bound of sim stays below this (alpha - tol, > 0)
};

// ---- signature upper bound -------------------------------------------------------------------------------
// With a_i = ca_i / |a| and b_i = cb_i / |b| over the kept blocks:   sum_i min(a_i, b_i) <= sum_{i shared} b_i
// = (sum_{i shared} cb_i) / |b|, and sum_{i shared} cb_i <= sh + (s1_b - popc(sig_b)), where sh = popc(sig_a & sig_b)
// counts the shared signature bits and the second term is the total count excess of b over one per set bit
// (valid for a folding hash too: a shared block always lands on a shared bit).  Symmetrically for a.  Since
// sim = smin / (S_a + S_b - smin) grows with smin, the bound on smin bounds sim.  The caller compares with
// alpha - tol, the same margin the estimate path uses, so a rejection here is a rejection there.
__device__ __forceinline__ bool cb_sig_may_join(float sh, float exRep, float nRepInv, float Sa, float exCmp,
                                                float nCmpInv, float Sb, float thr) {
  // ub / (Sa + Sb - ub) * 1.0002 < thr   <=>   ub * (1.0002 + thr) < thr * (Sa + Sb)      (den > 0: ub <= min(Sa, Sb))
  const float ub = fminf((sh + exRep) * nRepInv, (sh + exCmp) * nCmpInv);
  return !(ub * (1.0002f + thr) < thr * (Sa + Sb));
}

// ---- norm bound (no memory access beyond the meta record) ------------------------------------------------------
// sum_i min(a_i, b_i) <= min(S_a, S_b) and sum_i max(a_i, b_i) = S_a + S_b - sum_i min >= max(S_a, S_b), with
// S = (sum of counts) / |.|, so sim <= min(S_a, S_b) / max(S_a, S_b): a hub row (S ~ sqrt(#blocks)) can never join a
// cluster of short rows and vice versa.  Same margin and safety factor as the signature bound.
__device__ __forceinline__ bool cb_norm_may_join(float Sa, float Sb, float thr) {
  return !(fminf(Sa, Sb) * 1.0002f < thr * fmaxf(Sa, Sb));
}

// block 0 only: the next (up to G) unclustered positions at or after `from`, in order
__device__ void cb_find_seeds(const ClusterBatchArgs& a, u32 from, u32* sCount) {
  __shared__ u32 warpCnt[kCbWarps];
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) *sCount = 0;
  __syncthreads();
  for (u32 base = from; base < a.M; base += kCbThreads) {
    const u32 pos = base + threadIdx.x;
    const bool open = pos < a.M && __ldcg(a.cid + pos) == kNull;
    const unsigned bal = __ballot_sync(0xffffffffu, open);
    if (lane == 0) warpCnt[warp] = __popc(bal);
    __syncthreads();
    u32 before = *sCount;
    for (u32 w = 0; w < warp; ++w) before += warpCnt[w];
    const u32 rank = before + __popc(bal & ((1u << lane) - 1u));
    if (open && rank < a.G) a.seeds[1 + rank] = pos;
    __syncthreads();
    if (threadIdx.x == 0) {
      u32 t = *sCount;
      for (int w = 0; w < kCbWarps; ++w) t += warpCnt[w];
      *sCount = t;
    }
    __syncthreads();
    if (*sCount >= a.G) break;
  }
  if (threadIdx.x == 0) {
    a.seeds[0] = *sCount < a.G ? *sCount : a.G;
    __threadfence();
  }
  __syncthreads();
}


// applies "row at position f joins cluster jk": rep_jk += hist(f), sum of squares updated exactly in uint32
template <int TG>
__device__ __forceinline__ void cb_apply_join(const ClusterBatchArgs& a, u32* rep, u32* repSig, u32* sPop, u32* sSs,
                                              u32* sS1, u32* sRed, u32 f, u32 jk) {
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint4 m = a.meta[f];
  if (a.W) {  // signature of the representative = OR of its members'
    u32 added = 0;
    for (u32 w = threadIdx.x; w < a.W; w += kCbThreads) {
      const u32 x = a.sig[(size_t)f * a.W + w], old = repSig[w * TG + jk];
      added += __popc(x & ~old);
      repSig[w * TG + jk] = old | x;
    }
    added = __reduce_add_sync(0xffffffffu, added);
    if (lane == 0 && added) atomicAdd(sPop + jk, added);
  }
  u32 dss = 0;
  u32* rk = rep + jk;
  for (u32 j = threadIdx.x; j < m.y; j += kCbThreads) {
    const uint2 en = a.enc[m.x + j];
    const u32 r = rk[(size_t)en.x * TG];
    const u32 nr = r + en.y;
    dss += nr * nr - r * r;
    rk[(size_t)en.x * TG] = nr;
  }
  dss = __reduce_add_sync(0xffffffffu, dss);
  __syncthreads();
  if (lane == 0) sRed[warp] = dss;
  __syncthreads();
  if (threadIdx.x == 0) {
    u32 t = 0;
    for (int w = 0; w < kCbWarps; ++w) t += sRed[w];
    sSs[jk] += t;
    sS1[jk] += m.w;
  }
  __syncthreads();
}

// One candidate row against the first kmax clusters of the batch, the whole warp on the row (entries over
// lanes).  Returns the cluster it joins (warp-uniform) or kNull.  Clusters are tried in order; the estimate
// decides unless it is within `tol` of alpha, then the literal reduction tree does.
template <int TG>
__device__ __forceinline__ u32 cb_eval_row_warp(const ClusterBatchArgs& a, const u32* rep, const u32* repSig,
                                                const u32* sPop, const u32* sSs, const u32* sS1,
                                                const float (&nRepInv)[TG], const float (&Sa)[TG], const u32 pos,
                                                const uint4 m, const u32 kmax, const float tol, const u32 lane) {
  const uint2* ent = a.enc + m.x;
  const u32 ssCmp = m.z;
  const float nCmpInv = ssCmp ? 1.0f / sqrtf((float)ssCmp) : 0.f;
  if (a.sigThr > 0.f && ssCmp) {  // norm bound: decided from the meta record alone
    const float Sb = (float)m.w * nCmpInv;
    bool any = false;
#pragma unroll
    for (int k = 0; k < TG; ++k)
      if ((u32)k < kmax) any = any || sSs[k] == 0 || cb_norm_may_join(Sa[k], Sb, a.sigThr);
    if (!any) return kNull;
  }
  if (a.W && ssCmp && m.y * 2u > a.W) {  // reading the signature is cheaper than reading the entries
    u32 sh[TG], popB = 0;
#pragma unroll
    for (int k = 0; k < TG; ++k) sh[k] = 0;
    const u32* sg = a.sig + (size_t)pos * a.W;
    for (u32 w = lane; w < a.W; w += 32) {
      const u32 x = sg[w];
      popB += __popc(x);
      u32 rs[TG];
      if constexpr (TG >= 4) {
#pragma unroll
        for (int q = 0; q < TG / 4; ++q) {
          const uint4 v4 = *reinterpret_cast<const uint4*>(repSig + (size_t)w * TG + q * 4);
          rs[q * 4 + 0] = v4.x; rs[q * 4 + 1] = v4.y; rs[q * 4 + 2] = v4.z; rs[q * 4 + 3] = v4.w;
        }
      } else {
#pragma unroll
        for (int k = 0; k < TG; ++k) rs[k] = repSig[(size_t)w * TG + k];
      }
#pragma unroll
      for (int k = 0; k < TG; ++k) sh[k] += __popc(x & rs[k]);
    }
    popB = __reduce_add_sync(0xffffffffu, popB);
    const float exCmp = (float)(m.w - popB), Sb = (float)m.w * nCmpInv;
    bool any = false;
#pragma unroll
    for (int k = 0; k < TG; ++k) {
      if ((u32)k < kmax) {
        const u32 shk = __reduce_add_sync(0xffffffffu, sh[k]);
        any = any || sSs[k] == 0 ||
              cb_sig_may_join((float)shk, (float)(sS1[k] - sPop[k]), nRepInv[k], Sa[k], exCmp, nCmpInv, Sb, a.sigThr);
      }
    }
    if (!any) {
      if (lane == 0) atomicAdd(a.stats + 5, 1u);
      return kNull;
    }
  }
  float mn[TG];
#pragma unroll
  for (int k = 0; k < TG; ++k) mn[k] = 0.f;
  // 4 entry loads in flight per lane (the lists stream from HBM: latency, not bandwidth, is the limit)
  for (u32 j0 = lane; j0 < m.y; j0 += 128) {
    uint2 en4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const u32 j = j0 + u * 32;
      en4[u] = j < m.y ? ent[j] : make_uint2(0u, 0u);  // count 0 contributes min(.,0) = 0
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint2 en = en4[u];
      const float b = (float)en.y * nCmpInv;
      u32 rv[TG];
      if constexpr (TG >= 4) {
#pragma unroll
        for (int q = 0; q < TG / 4; ++q) {
          const uint4 v4 = *reinterpret_cast<const uint4*>(rep + (size_t)en.x * TG + q * 4);
          rv[q * 4 + 0] = v4.x; rv[q * 4 + 1] = v4.y; rv[q * 4 + 2] = v4.z; rv[q * 4 + 3] = v4.w;
        }
      } else if constexpr (TG == 2) {
        const uint2 v2 = *reinterpret_cast<const uint2*>(rep + (size_t)en.x * 2);
        rv[0] = v2.x; rv[1] = v2.y;
      } else {
        rv[0] = rep[en.x];
      }
#pragma unroll
      for (int k = 0; k < TG; ++k) mn[k] += fminf((float)rv[k] * nRepInv[k], b);  // clusters >= kmax are ignored below
    }
  }
#pragma unroll
  for (int k = 0; k < TG; ++k) {
    if ((u32)k < kmax) {
#pragma unroll
      for (int w = 16; w >= 1; w >>= 1) mn[k] += __shfl_xor_sync(0xffffffffu, mn[k], w);
    }
  }
  u32 joinK = kNull;
#pragma unroll
  for (int k = 0; k < TG; ++k) {
    if ((u32)k < kmax && joinK == kNull) {  // warp-uniform
      const u32 ssRep = sSs[k];
      bool join;
      if (ssRep == 0 || ssCmp == 0) {
        join = ((ssRep == 0 && ssCmp == 0) ? 1.0f : 0.0f) > a.alpha;  // rowReordering.cu:263-268
      } else {
        const float den = Sa[k] + (float)m.w * nCmpInv - mn[k];
        const float est = mn[k] / den;
        if (den > 0.f && est > a.alpha + tol) join = true;
        else if (den > 0.f && est < a.alpha - tol) join = false;
        else {
          const float sim = exact_similarity_warp(rep + k, ent, m.y, a.nbpr, a.B, a.keptMask, ssRep, ssCmp, TG);
          join = sim > a.alpha;
          if (lane == 0) atomicAdd(a.stats + 0, 1u);
        }
      }
      if (join) joinK = (u32)k;
    }
  }
  return joinK;
}

// The same decision for a SHORT row by one lane on its own (no shuffles).  If any cluster that has to be
// decided is ambiguous the lane gives up (*again = true) and the row is redone by cb_eval_row_warp.
template <int TG>
__device__ __forceinline__ u32 cb_eval_row_lane(const ClusterBatchArgs& a, const u32* rep, const u32* repSig,
                                                const u32* sPop, const u32* sSs, const u32* sS1,
                                                const float (&nRepInv)[TG], const float (&Sa)[TG], const u32 pos,
                                                const uint4 m, const u32 kmax, const float tol, bool* again) {
  const uint2* ent = a.enc + m.x;
  const u32 ssCmp = m.z;
  const float nCmpInv = ssCmp ? 1.0f / sqrtf((float)ssCmp) : 0.f;
  *again = false;
  if (a.sigThr > 0.f && ssCmp) {  // norm bound: decided from the meta record alone
    const float Sb = (float)m.w * nCmpInv;
    bool any = false;
#pragma unroll
    for (int k = 0; k < TG; ++k)
      if ((u32)k < kmax) any = any || sSs[k] == 0 || cb_norm_may_join(Sa[k], Sb, a.sigThr);
    if (!any) return kNull;
  }
  if (a.W && ssCmp && (m.y * 2u > a.W || m.y > a.laneRows)) {
    u32 sh[TG], popB = 0;
#pragma unroll
    for (int k = 0; k < TG; ++k) sh[k] = 0;
    const u32* sg = a.sig + (size_t)pos * a.W;
    for (u32 w = 0; w < a.W; ++w) {
      const u32 x = sg[w];
      popB += __popc(x);
      u32 rs[TG];
      if constexpr (TG >= 4) {
#pragma unroll
        for (int q = 0; q < TG / 4; ++q) {
          const uint4 v4 = *reinterpret_cast<const uint4*>(repSig + (size_t)w * TG + q * 4);
          rs[q * 4 + 0] = v4.x; rs[q * 4 + 1] = v4.y; rs[q * 4 + 2] = v4.z; rs[q * 4 + 3] = v4.w;
        }
      } else {
#pragma unroll
        for (int k = 0; k < TG; ++k) rs[k] = repSig[(size_t)w * TG + k];
      }
#pragma unroll
      for (int k = 0; k < TG; ++k) sh[k] += __popc(x & rs[k]);
    }
    const float exCmp = (float)(m.w - popB), Sb = (float)m.w * nCmpInv;
    bool any = false;
#pragma unroll
    for (int k = 0; k < TG; ++k)
      if ((u32)k < kmax)
        any = any || sSs[k] == 0 ||
              cb_sig_may_join((float)sh[k], (float)(sS1[k] - sPop[k]), nRepInv[k], Sa[k], exCmp, nCmpInv, Sb, a.sigThr);
    if (!any) return kNull;
  }
  if (m.y > a.laneRows) {  // a long row that may join: redone by a whole warp
    *again = true;
    return kNull;
  }
  float mn[TG];
#pragma unroll
  for (int k = 0; k < TG; ++k) mn[k] = 0.f;
  for (u32 j0 = 0; j0 < m.y; j0 += 4) {  // 4 independent entry loads in flight
    uint2 en4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) en4[u] = j0 + u < m.y ? ent[j0 + u] : make_uint2(0u, 0u);  // count 0 adds 0
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint2 en = en4[u];
      const float b = (float)en.y * nCmpInv;
      u32 rv[TG];
      if constexpr (TG >= 4) {
#pragma unroll
        for (int q = 0; q < TG / 4; ++q) {
          const uint4 v4 = *reinterpret_cast<const uint4*>(rep + (size_t)en.x * TG + q * 4);
          rv[q * 4 + 0] = v4.x; rv[q * 4 + 1] = v4.y; rv[q * 4 + 2] = v4.z; rv[q * 4 + 3] = v4.w;
        }
      } else if constexpr (TG == 2) {
        const uint2 v2 = *reinterpret_cast<const uint2*>(rep + (size_t)en.x * 2);
        rv[0] = v2.x; rv[1] = v2.y;
      } else {
        rv[0] = rep[en.x];
      }
#pragma unroll
      for (int k = 0; k < TG; ++k) mn[k] += fminf((float)rv[k] * nRepInv[k], b);
    }
  }
  u32 joinK = kNull;
#pragma unroll
  for (int k = 0; k < TG; ++k) {
    if ((u32)k < kmax && joinK == kNull && !*again) {
      const u32 ssRep = sSs[k];
      if (ssRep == 0 || ssCmp == 0) {
        if (((ssRep == 0 && ssCmp == 0) ? 1.0f : 0.0f) > a.alpha) joinK = (u32)k;
      } else {
        const float den = Sa[k] + (float)m.w * nCmpInv - mn[k];
        const float est = mn[k] / den;
        if (den > 0.f && est > a.alpha + tol) joinK = (u32)k;
        else if (!(den > 0.f && est < a.alpha - tol)) *again = true;
      }
    }
  }
  return joinK;
}

template <int TG>  // representatives interleaved: rep[block * TG + k], so one lookup serves all TG clusters
static __global__ void __launch_bounds__(kCbThreads, 1) k_cluster_batched(ClusterBatchArgs a) {
  extern __shared__ __align__(16) u32 rep[];  // nbpr x TG accumulated histograms, then W x TG signature words
  u32* repSig = rep + (((size_t)TG * a.nbpr + 3u) & ~(size_t)3u);
  __shared__ u32 sSeed[kCbMaxG], sSs[kCbMaxG], sS1[kCbMaxG], sPop[kCbMaxG];
  __shared__ u32 sRed[kCbWarps];
  __shared__ u32 sCount;
  __shared__ u32 sTodo[kCbThreads], sTodoCnt;
  cg::grid_group grid = cg::this_grid();
  const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 totalWarps = gridDim.x * kCbWarps;
  const u32 gw = blockIdx.x * kCbWarps + warp;
  if (threadIdx.x == 0) sTodoCnt = 0;
  volatile unsigned long long* slots = a.slots;
  volatile u32* seeds = a.seeds;
  const float tol = kTolRel * fabsf(a.alpha) + kTolAbs;

  u32 c0 = 1, iter = 0;
  bool calm = false;  // the previous batch saw no join: open the next one with the widest window
  if (blockIdx.x == 0) cb_find_seeds(a, a.start0, &sCount);
  __threadfence();
  grid.sync();
  while (true) {
    u32 g = seeds[0];
    if (g == 0) break;
    // ---- open the batch: representatives = histograms of the speculated seeds
    __syncthreads();
    if (threadIdx.x < g) {
      const u32 sd = seeds[1 + threadIdx.x];
      sSeed[threadIdx.x] = sd;
      const uint4 m = a.meta[sd];
      sSs[threadIdx.x] = m.z;
      sS1[threadIdx.x] = m.w;
    }
    if (threadIdx.x < kCbMaxG) sPop[threadIdx.x] = 0;
    for (u32 i = threadIdx.x; i < (u32)TG * a.nbpr; i += kCbThreads) rep[i] = 0;
    for (u32 i = threadIdx.x; i < (u32)TG * a.W; i += kCbThreads) repSig[i] = 0;
    __syncthreads();
    for (u32 i = threadIdx.x; i < g * a.W; i += kCbThreads) {
      const u32 k = i / a.W, w = i - k * a.W;
      const u32 x = a.sig[(size_t)sSeed[k] * a.W + w];
      repSig[w * TG + k] = x;
      if (x) atomicAdd(sPop + k, (u32)__popc(x));
    }
    for (u32 k = 0; k < g; ++k) {
      const uint4 m = a.meta[sSeed[k]];
      for (u32 j = threadIdx.x; j < m.y; j += kCbThreads) {
        const uint2 e = a.enc[m.x + j];
        rep[e.x * TG + k] = e.y;
      }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      a.cid[sSeed[0]] = c0;  // the first open position is a seed for certain
      atomicAdd(a.stats + 3, 1u);
    }
    __syncthreads();
    const u32 unit = a.laneRows ? 32u : 1u;  // positions one warp looks at per step
    u32 p = sSeed[0] + 1, chunk = (calm ? totalWarps * 16u : totalWarps) * unit;
    calm = true;
    while (p < a.M) {
      const u32 e = (a.M - p > chunk) ? p + chunk : a.M;
      const u32 slot = iter % 3;
      float nRepInv[TG], Sa[TG];
#pragma unroll
      for (int k = 0; k < TG; ++k) {
        const u32 ss = (u32)k < g ? sSs[k] : 0u;
        nRepInv[k] = ss ? 1.0f / sqrtf((float)ss) : 0.f;
        Sa[k] = (u32)k < g ? (float)sS1[k] * nRepInv[k] : 0.f;
      }
      // position -> warp: spread consecutive positions over the CTAs first (small windows use every SM)
      const u32 vw = warp * gridDim.x + blockIdx.x;
      if (a.laneRows) {
        // short-row matrices: a warp takes 32 consecutive positions, one per lane.  Rows that are long, or whose
        // estimate is ambiguous, go to a per-CTA list and are then redone one per WARP (a run of long rows
        // would otherwise be serialised inside one warp).
        const u32 span = totalWarps * 32u;
        const u32 nIter = (e - p + span - 1) / span;  // uniform over the grid
        for (u32 it = 0; it < nIter; ++it) {
          const u32 pos = p + it * span + vw * 32u + lane;
          u32 kmax = 0;
          uint4 m = make_uint4(0u, 0u, 0u, 0u);
          if (pos < e && __ldcg(a.cid + pos) == kNull) {
#pragma unroll
            for (int k = 0; k < TG; ++k) kmax += ((u32)k < g && sSeed[k] < pos) ? 1u : 0u;
            if (kmax) m = a.meta[pos];
          }
          u32 joinK = kNull;
          bool again = false;
          if (kmax) {
            joinK = cb_eval_row_lane<TG>(a, rep, repSig, sPop, sSs, sS1, nRepInv, Sa, pos, m, kmax, tol, &again);
          }
          if (again) {
            sTodo[atomicAdd(&sTodoCnt, 1u)] = pos;  // at most one per thread and step
            atomicAdd(a.stats + 4, 1u);
          }
          // lanes hold ascending positions: only the first accepting lane can be the window's first joiner
          const unsigned acc = __ballot_sync(0xffffffffu, joinK != kNull);
          if (acc && (int)lane == __ffs(acc) - 1) {
            atomicMin(a.slots + slot, ((unsigned long long)pos << 8) | (unsigned long long)joinK);
            atomicAdd(a.accepts + slot, (u32)__popc(acc));
          }
          __syncthreads();
          const u32 nTodo = sTodoCnt;
          for (u32 i = warp; i < nTodo; i += kCbWarps) {
            const u32 tp = sTodo[i];
            u32 ks = 0;
#pragma unroll
            for (int k = 0; k < TG; ++k) ks += ((u32)k < g && sSeed[k] < tp) ? 1u : 0u;
            const u32 jk = cb_eval_row_warp<TG>(a, rep, repSig, sPop, sSs, sS1, nRepInv, Sa, tp, a.meta[tp], ks, tol, lane);
            if (jk != kNull && lane == 0) {
              atomicMin(a.slots + slot, ((unsigned long long)tp << 8) | (unsigned long long)jk);
              atomicAdd(a.accepts + slot, 1u);
            }
          }
          __syncthreads();
          if (threadIdx.x == 0) sTodoCnt = 0;
        }
      } else {
        for (u32 pos = p + vw; pos < e; pos += totalWarps) {
          if (__ldcg(a.cid + pos) != kNull) continue;
          // clusters of the batch whose seed precedes this position (seeds are ascending)
          u32 kmax = 0;
#pragma unroll
          for (int k = 0; k < TG; ++k) kmax += ((u32)k < g && sSeed[k] < pos) ? 1u : 0u;
          if (kmax == 0) continue;
          const uint4 m = a.meta[pos];
          const u32 joinK = cb_eval_row_warp<TG>(a, rep, repSig, sPop, sSs, sS1, nRepInv, Sa, pos, m, kmax, tol, lane);
          if (joinK != kNull && lane == 0) {
            atomicMin(a.slots + slot, ((unsigned long long)pos << 8) | (unsigned long long)joinK);
            atomicAdd(a.accepts + slot, 1u);
          }
        }
      }
      __threadfence();
      grid.sync();
      const unsigned long long v = slots[slot];
      const u32 nAccept = ((volatile u32*)a.accepts)[slot];
      if (blockIdx.x == 0 && threadIdx.x == 0) {
        slots[(iter + 2) % 3] = ~0ull;
        a.accepts[(iter + 2) % 3] = 0u;
        atomicAdd(a.stats + 2, 1u);
      }
      const bool joined = v != ~0ull;
      const u32 f = joined ? (u32)(v >> 8) : 0u;
      const u32 jk = joined ? (u32)(v & 255u) : 0u;
      const u32 limit = joined ? f : e;  // positions in [p, limit) are final
      // speculated seeds that were passed without being absorbed are confirmed
      if (blockIdx.x == 0 && threadIdx.x >= 1 && threadIdx.x < g) {
        const u32 sd = sSeed[threadIdx.x];
        if (sd >= p && sd < limit) a.cid[sd] = c0 + threadIdx.x;
      }
      if (joined) {
        // was the joiner one of our speculated seeds?  then clusters from that one on never existed
        u32 newG = g;
        for (u32 j = 1; j < g; ++j)
          if (sSeed[j] == f) newG = j;
        if (blockIdx.x == 0 && threadIdx.x == 0) a.cid[f] = c0 + jk;
        cb_apply_join<TG>(a, rep, repSig, sPop, sSs, sS1, sRed, f, jk);
        g = newG;
        p = f + 1;
        // rows tend to join in streaks and only the first joiner of a window counts: look at the candidates right
        // behind it, then widen again while nothing joins.  A window costs its latency (one evaluation + one grid
        // barrier, ~10-20 us), not its width, so the first window after a join already uses a quarter of the warps.
        chunk = gridDim.x * (kCbWarps / 4) * unit;
        calm = false;
      } else {
        p = e;
        // every window costs a grid-wide barrier: with one candidate per lane a wasted position is cheap, so
        // widen faster
        if (chunk < totalWarps * 16u * unit) chunk *= a.laneRows ? 8u : 2u;
      }
      ++iter;
    }
    // ---- batch done: clusters c0 .. c0+g-1 are complete
    c0 += g;
    const u32 from = sSeed[g - 1] + 1;
    __syncthreads();
    if (blockIdx.x == 0) cb_find_seeds(a, from, &sCount);
    __threadfence();
    grid.sync();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) a.stats[1] = c0 - 1;
}

// ------------------------------------------------------------------------------------------
// final permutation
// ------------------------------------------------------------------------------------------
static __global__ void k_compose_perm(const u32* __restrict__ asc, const u32* __restrict__ indices, u32 skip, u32 n,
                                      u32* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = asc[indices[skip + i]];
}

// ------------------------------------------------------------------------------------------
// host drivers
// ------------------------------------------------------------------------------------------
struct SortedCols {
  const u32* cols;       // per-row ascending columns (either the input or an owned sorted copy)
  DevBuf<u32> owned;
};

static void make_sorted_cols(const u32* d_rowOff, const u32* d_colIdx, u32 M, u32 N, u32 nnz, cudaStream_t s,
                             SortedCols& out) {
  DevBuf<u32> flag(1);
  SB_CUDA(cudaMemsetAsync(flag.get(), 0, 4, s));
  k_check_rows_sorted<<<grid_for((size_t)M * 32), 256, 0, s>>>(d_rowOff, d_colIdx, M, flag.get());
  SB_LAUNCH_CHECK();
  u32 h = 0;
  SB_CUDA(cudaMemcpyAsync(&h, flag.get(), 4, cudaMemcpyDeviceToHost, s));
  SB_CUDA(cudaStreamSynchronize(s));
  if (!h) {
    out.cols = d_colIdx;
    return;
  }
  // file-order columns inside a row (src/Matrix.cpp:467 sorts by row only): sort (row, col) keys
  DevBuf<u64> ka(nnz), kb(nnz);
  k_make_row_col_keys<<<grid_for((size_t)M * 32), 256, 0, s>>>(d_rowOff, d_colIdx, M, ka.get());
  SB_LAUNCH_CHECK();
  const int which = radix_sort_pairs<u64>(ka.get(), kb.get(), nullptr, nullptr, nnz, 0, 32 + bits_for(M), s);
  (void)N;
  out.owned.alloc(nnz);
  k_keys_low32<<<grid_for(nnz), 256, 0, s>>>(which ? kb.get() : ka.get(), out.owned.get(), nnz);
  SB_LAUNCH_CHECK();
  SB_CUDA(cudaStreamSynchronize(s));
  out.cols = out.owned.get();
}

void dispersion_dev(const u32* d_rowOff, const u32* d_colIdx, u32 M, u32 N, u32 nnz, u32 bs, u32* d_disp,
                    u32* nbprOut, cudaStream_t s) {
  const u32 nbpr = num_blocks_per_row(N, bs);
  if (nbprOut) *nbprOut = nbpr;
  TempScope tempScope(s);
  SortedCols sc;
  make_sorted_cols(d_rowOff, d_colIdx, M, N, nnz, s, sc);
  const u32 B = cluster_blockdim(nbpr);
  k_encode_count<<<grid_for((size_t)M * 32), 256, 0, s>>>(d_rowOff, sc.cols, M, bs, B, kept_warp_mask(B), d_disp,
                                                         nullptr);
  SB_LAUNCH_CHECK();
  SB_CUDA(cudaStreamSynchronize(s));
}

void row_reorder_dev(const u32* d_rowOff, const u32* d_colIdx, u32 M, u32 N, u32 nnz, float alpha, u32 bs,
                     const bsmr_reorder_opts* opts, u32* d_reorderedRows, u32* numRows, int32_t* numClusters,
                     RowReorderStats* st, cudaStream_t s) {
  // options: explicit > environment > automatic
  bsmr_reorder_opts o{};
  if (opts) o = *opts;
  if (o.kernel == BSMR_CLUSTER_AUTO) {
    const char* e = getenv("SDDMM_B200_CLUSTER");
    o.kernel = (e && !strcmp(e, "legacy")) ? BSMR_CLUSTER_LEGACY : BSMR_CLUSTER_BATCHED;
  }
  if (o.batch == 0)
    if (const char* e = getenv("SDDMM_B200_CLUSTER_G")) { const int v = atoi(e); if (v >= 1 && v <= 8) o.batch = (u32)v; }
  int laneCfg = -1;  // -1 auto, 0 off, > 0 row-length threshold
  if (o.laneRows == BSMR_TRISTATE_OFF) laneCfg = 0;
  else if (o.laneRows == BSMR_TRISTATE_ON) laneCfg = 64;
  else if (const char* e = getenv("SDDMM_B200_CLUSTER_LANE")) laneCfg = atoi(e);
  if (M == 0) { *numRows = 0; if (numClusters) *numClusters = 0; return; }
  TempScope tempScope(s);
  const u32 nbpr = num_blocks_per_row(N, bs);
  const u32 B = cluster_blockdim(nbpr);
  const u32 keptMask = kept_warp_mask(B);
  if ((size_t)nbpr * 4 > 200 * 1024) fail(SDDMM_E_UNSUPPORTED, "nbpr=%u does not fit shared memory; use a larger block_size", nbpr);

  SortedCols sc;
  make_sorted_cols(d_rowOff, d_colIdx, M, N, nnz, s, sc);

  // K1 pass 1: dispersion + kept-entry counts
  DevBuf<u32> disp(M), keptCnt(M);
  k_encode_count<<<grid_for((size_t)M * 32), 256, 0, s>>>(d_rowOff, sc.cols, M, bs, B, keptMask, disp.get(),
                                                         keptCnt.get());
  SB_LAUNCH_CHECK();

  // K3: rows by (dispersion asc, row asc), stable
  DevBuf<u32> keyA(M), keyB(M), ascA(M), ascB(M);
  SB_CUDA(cudaMemcpyAsync(keyA.get(), disp.get(), (size_t)M * 4, cudaMemcpyDeviceToDevice, s));
  iota<u32>(ascA.get(), M, 0u, s);
  const int w1 = radix_sort_pairs<u32>(keyA.get(), keyB.get(), ascA.get(), ascB.get(), M, 0, 32, s);
  const u32* asc = w1 ? ascB.get() : ascA.get();

  DevBuf<u32> zc(1);
  SB_CUDA(cudaMemsetAsync(zc.get(), 0, 4, s));
  k_count_zero<<<grid_for(M), 256, 0, s>>>(disp.get(), M, zc.get());
  SB_LAUNCH_CHECK();
  u32 zeroRows = 0;
  SB_CUDA(cudaMemcpyAsync(&zeroRows, zc.get(), 4, cudaMemcpyDeviceToHost, s));
  SB_CUDA(cudaStreamSynchronize(s));

  // K1 pass 2: sparse encodings laid out in dispersion order
  DevBuf<u32> encOff((size_t)M + 1);
  k_gather_u32<<<grid_for(M), 256, 0, s>>>(keptCnt.get(), asc, encOff.get(), M);
  SB_LAUNCH_CHECK();
  SB_CUDA(cudaMemsetAsync(encOff.get() + M, 0, 4, s));
  exclusive_scan_u32(encOff.get(), encOff.get(), (size_t)M + 1, s);
  u32 totalEnt = 0;
  SB_CUDA(cudaMemcpyAsync(&totalEnt, encOff.get() + M, 4, cudaMemcpyDeviceToHost, s));
  SB_CUDA(cudaStreamSynchronize(s));
  DevBuf<uint2> enc(totalEnt ? totalEnt : 1);
  DevBuf<uint4> meta(M);
  k_encode_fill<<<grid_for((size_t)M * 32), 256, 0, s>>>(d_rowOff, sc.cols, M, bs, B, keptMask, asc, encOff.get(),
                                                        enc.get(), meta.get());
  SB_LAUNCH_CHECK();

  // K2: clustering
  DevBuf<u32> cid(M), ctrl(8), ncl(1);
  k_init_cid<<<grid_for(M), 256, 0, s>>>(cid.get(), M, zeroRows);
  SB_LAUNCH_CHECK();
  SB_CUDA(cudaMemsetAsync(ctrl.get(), 0xFF, 6 * 4, s));
  SB_CUDA(cudaMemsetAsync(ctrl.get() + 6, 0, 2 * 4, s));
  SB_CUDA(cudaMemsetAsync(ncl.get(), 0, 4, s));
  u32 exactEvals = 0;
  const bool legacy = o.kernel == BSMR_CLUSTER_LEGACY;
  DevBuf<unsigned long long> slots(4);
  DevBuf<u32> seedsBuf(16), statsBuf(8), acceptsBuf(4);
  if (zeroRows < M && !legacy) {
    ClusterBatchArgs a;
    a.M = M; a.start0 = zeroRows; a.nbpr = nbpr; a.B = B; a.keptMask = keptMask; a.alpha = alpha;
    a.enc = enc.get(); a.meta = meta.get(); a.cid = cid.get(); a.slots = slots.get(); a.seeds = seedsBuf.get();
    a.stats = statsBuf.get();
    a.accepts = acceptsBuf.get();
    // short-row matrices (graphs): evaluate one candidate per lane (SDDMM_B200_CLUSTER_LANE=0 turns it off)
    {
      const double avgEnt = (double)totalEnt / (double)(M - zeroRows);
      a.laneRows = laneCfg == 0 ? 0u : laneCfg > 0 ? (u32)laneCfg : (avgEnt <= 24.0 ? 64u : 0u);
    }
    SB_CUDA(cudaMemsetAsync(acceptsBuf.get(), 0, 16, s));
    // bitmap signatures (see cb_sig_may_join): ~8 bits per average kept entry, a power of two; an exact bitmap when
    // that covers nbpr.  Off when alpha - tol <= 0 (nothing can be rejected) or by option.
    DevBuf<u32> sigBuf;
    a.sig = nullptr; a.W = 0; a.sigMask = 0xFFFFFFFFu;
    a.sigThr = alpha - (kTolRel * fabsf(alpha) + kTolAbs);
    {
      int sigCfg = -1;
      if (o.signature == BSMR_TRISTATE_OFF) sigCfg = 0;
      else if (o.signature == BSMR_TRISTATE_ON) sigCfg = 1;
      else if (const char* e = getenv("SDDMM_B200_CLUSTER_SIG")) sigCfg = atoi(e) != 0;
      if (sigCfg != 0 && a.sigThr > 0.f && totalEnt) {
        const double avgEnt = (double)totalEnt / (double)(M - zeroRows);
        u32 bits = 64;
        // signature bits per average kept entry.  Lane regime (graphs): 16 -- R-MAT scale 22, whole row reorder at
        // 4 / 8 / 16 / 32 bits per entry: 37.6 / 29.5 / 26.6 / 28.9 s (fewer entry-list walks by single lanes vs
        // more signature bytes per candidate); the permutation is the same whatever the width.
        double perEnt = a.laneRows ? 16.0 : 8.0;
        if (const char* e = getenv("SDDMM_B200_CLUSTER_SIGBITS")) { const double v = atof(e); if (v >= 1.0 && v <= 64.0) perEnt = v; }
        while (bits < perEnt * avgEnt && bits < (1u << 16)) bits <<= 1;
        if (bits >= nbpr) { a.W = (nbpr + 31u) / 32u; a.sigMask = 0xFFFFFFFFu; }
        else { a.W = bits / 32u; a.sigMask = bits - 1u; }
        sigBuf.alloc((size_t)M * a.W);
        SB_CUDA(cudaMemsetAsync(sigBuf.get(), 0, (size_t)M * a.W * 4, s));
        k_build_sig<<<grid_for((size_t)M * 32), 256, 0, s>>>(enc.get(), meta.get(), M, a.W, a.sigMask, sigBuf.get());
        SB_LAUNCH_CHECK();
        a.sig = sigBuf.get();
      }
    }
    u32 G = (u32)((200u * 1024u) / (((size_t)nbpr + a.W + 1) * 4));
    if (o.batch && o.batch < G) G = o.batch;
    // lane regime: 4 clusters per batch unless asked otherwise (a lane's work and its chance of having to walk an
    // entry list both grow with the batch: R-MAT scale 20 / 22 at G = 8 vs 4: 5.8 vs 4.8 s, 26.6 vs 25.1 s); the warp
    // regime wants the widest batch (uniform 100k^2: 0.90 s at 8, 2.15 s at 4)
    if (a.laneRows && !o.batch && G > 4) G = 4;
    G = G >= 8 ? 8u : G >= 4 ? 4u : G >= 2 ? 2u : 1u;  // template instances
    a.G = G;
    SB_CUDA(cudaMemsetAsync(slots.get(), 0xFF, 4 * 8, s));
    SB_CUDA(cudaMemsetAsync(seedsBuf.get(), 0, 16 * 4, s));
    SB_CUDA(cudaMemsetAsync(statsBuf.get(), 0, 8 * 4, s));
    const size_t smem = ((((size_t)G * nbpr + 3u) & ~(size_t)3u) + (size_t)G * a.W) * 4;
    const void* kern = G == 8 ? (const void*)k_cluster_batched<8> : G == 4 ? (const void*)k_cluster_batched<4>
                     : G == 2 ? (const void*)k_cluster_batched<2> : (const void*)k_cluster_batched<1>;
    SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int perSm = 0;
    SB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kern, kCbThreads, smem));
    if (perSm < 1) fail(SDDMM_E_UNSUPPORTED, "batched clustering kernel does not fit on an SM (nbpr=%u)", nbpr);
    if (perSm > 2) perSm = 2;
    u32 grid = (u32)(perSm * device_sm_count());
    const u32 need = ceil_div(M - zeroRows, kCbWarps);
    if (grid > need) grid = need ? need : 1;
    if (const char* e = getenv("SDDMM_B200_CLUSTER_GRID")) { const int v = atoi(e); if (v >= 1 && (u32)v < grid) grid = (u32)v; }
    void* args[] = {&a};
    SB_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(kCbThreads), args, smem, s));
    SB_LAUNCH_CHECK();
    SB_CUDA(cudaMemcpyAsync(&exactEvals, statsBuf.get(), 4, cudaMemcpyDeviceToHost, s));
    if (getenv("SDDMM_B200_CLUSTER_DEBUG")) {
      u32 hs[8];
      SB_CUDA(cudaMemcpyAsync(hs, statsBuf.get(), sizeof hs, cudaMemcpyDeviceToHost, s));
      SB_CUDA(cudaStreamSynchronize(s));
      fprintf(stderr, "[cluster] rows %u entries %u G %u grid %u laneRows %u sigWords %u: exact %u clusters %u windows %u batches %u redo %u sig-rejected %u\n",
              M - zeroRows, totalEnt, G, grid, a.laneRows, a.W, hs[0], hs[1], hs[2], hs[3], hs[4], hs[5]);
    }
    SB_CUDA(cudaMemcpyAsync(ncl.get(), statsBuf.get() + 1, 4, cudaMemcpyDeviceToDevice, s));
  } else if (zeroRows < M) {
    ClusterArgs a;
    a.M = M; a.start0 = zeroRows; a.nbpr = nbpr; a.B = B; a.keptMask = keptMask; a.alpha = alpha;
    a.enc = enc.get(); a.meta = meta.get(); a.cid = cid.get(); a.ctrl = ctrl.get(); a.numClustersOut = ncl.get();
    const size_t smem = (size_t)nbpr * 4;
    SB_CUDA(cudaFuncSetAttribute(k_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int perSm = 0;
    SB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_cluster, kClThreads, smem));
    if (perSm < 1) fail(SDDMM_E_UNSUPPORTED, "clustering kernel does not fit on an SM (nbpr=%u)", nbpr);
    {
      static const int cap = [] { const char* e = getenv("SDDMM_B200_CLUSTER_CTAS_PER_SM"); return e ? atoi(e) : 6; }();
      if (perSm > cap) perSm = cap;
    }
    // no more CTAs than there are candidate rows to look at
    u32 grid = (u32)(perSm * device_sm_count());
    const u32 need = ceil_div(M - zeroRows, kClWarps);
    if (grid > need) grid = need ? need : 1;
    void* args[] = {&a};
    SB_CUDA(cudaLaunchCooperativeKernel((void*)k_cluster, dim3(grid), dim3(kClThreads), args, smem, s));
    SB_LAUNCH_CHECK();
    SB_CUDA(cudaMemcpyAsync(&exactEvals, ctrl.get() + 6, 4, cudaMemcpyDeviceToHost, s));
  }

  // :986-995  stable sort positions by cluster id; permutation = asc[indices]
  DevBuf<u32> cidB(M), idxA(M), idxB(M);
  iota<u32>(idxA.get(), M, 0u, s);
  const int w2 = radix_sort_pairs<u32>(cid.get(), cidB.get(), idxA.get(), idxB.get(), M, 0, 32, s);
  const u32* sortedCid = w2 ? cidB.get() : cid.get();
  const u32* indices = w2 ? idxB.get() : idxA.get();
  const u32 nR = M - zeroRows;  // :1081-1090: cluster 0 == the empty rows == the stripped prefix
  if (nR) {
    k_compose_perm<<<grid_for(nR), 256, 0, s>>>(asc, indices, zeroRows, nR, d_reorderedRows);
    SB_LAUNCH_CHECK();
  }
  *numRows = nR;
  // :996 quirk: cluster_cnt = sortedIds[indices[M-1]] + (zero_row_idx != 0)
  u32 lastIdx = 0, q = 0;
  SB_CUDA(cudaMemcpyAsync(&lastIdx, indices + (M - 1), 4, cudaMemcpyDeviceToHost, s));
  SB_CUDA(cudaStreamSynchronize(s));
  SB_CUDA(cudaMemcpyAsync(&q, sortedCid + lastIdx, 4, cudaMemcpyDeviceToHost, s));
  SB_CUDA(cudaStreamSynchronize(s));
  if (numClusters) *numClusters = (int32_t)(q + (zeroRows != 0 ? 1u : 0u));
  if (st) {
    st->nbpr = nbpr; st->blockDim = B; st->keptMask = keptMask; st->zeroRows = zeroRows;
    st->exactEvals = exactEvals; st->encEntries = totalEnt;
    u32 c = 0;
    SB_CUDA(cudaMemcpy(&c, ncl.get(), 4, cudaMemcpyDeviceToHost));
    st->clustersCreated = c;
  }
}

}  // namespace sb
