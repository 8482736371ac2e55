// sddmm_kernels.cu -- the two SDDMM kernels (a10, a11 of SURVEY.md 8a) and their launcher (a9).
//
//   k_sddmm_dense    : dense 16x16 blocks on the 5th-gen tensor cores (tcgen05.mma kind::tf32,
//                      accumulator in TMEM).  Replaces the WMMA m16n16k8 kernels
//                      src/sddmmKernel.cu:213-351 and :355-488.
//                      The problem is transposed: up to 8 dense blocks = 128 gathered
//                      B columns sit on the MMA M axis, the 16 panel rows on N (M=128, N=16, K=8).
//                      Operands are rounded to TF32 with cvt.rna exactly like the reference's
//                      wmma::__float_to_tf32 (sddmmKernel.cu:318-323) while they are staged into the
//                      128B-swizzled K-major shared-memory tiles the UMMA descriptors describe.
//   k_sddmm_residual : FP32 CUDA-core kernel over the residual COO entries, 128-bit loads of the
//                      gathered B rows, A panel tile in shared memory, 8 lanes per non-zero with a
//                      shuffle reduction.  Replaces src/sddmmKernel.cu:1994-2104 and :2109-2199.
#include <algorithm>

#include <cuda.h>
#include <cuda_fp16.h>  // CUtensorMap and its enums only; the encoder is fetched through cudaGetDriverEntryPoint

#include <cstring>
#include <initializer_list>
#include <map>
#include <mutex>
#include <utility>

#include "layout.cuh"
#include "primitives.cuh"
#include "sddmm_kernels.cuh"

namespace sb {

// =============================================================================================
// residual kernel
// =============================================================================================
// batched form (reference: sddmm_gpu_batch, src/sddmmKernel.cu:2764-2850): blockIdx.y = batch id, operands of
// batch b start at A + b*M*K, B + b*N*K, P + b*nnz (element strides below; all 0 for a single pass)
struct BatchStrides {
  size_t a = 0, b = 0, p = 0;
};

constexpr int kResThreads = 256;
constexpr int kResLanes = 8;  // lanes cooperating on one non-zero

__device__ __forceinline__ float4 ldg_nc_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

template <bool kL1>
__device__ __forceinline__ float4 ld_b(const float4* p) {
  if (kL1) return __ldg(p);
  return ldg_nc_f4(p);
}

template <bool kL1>
static __global__ void __launch_bounds__(kResThreads)
k_sddmm_residual(u32 M, u32 K4, const float4* __restrict__ A4, const float4* __restrict__ B4,
                 const u32* __restrict__ R, u32 nR, const u32* __restrict__ vOff, const u32* __restrict__ sVals,
                 const u32* __restrict__ sRows, const u32* __restrict__ sCols, const uint2* __restrict__ work,
                 u32 kSparseChunk, float* __restrict__ P, BatchStrides bs) {
  extern __shared__ float4 sA[];  // 16 rows x K4 float4
  A4 += (bs.a >> 2) * blockIdx.y;
  B4 += (bs.b >> 2) * blockIdx.y;
  P += bs.p * blockIdx.y;
  const uint2 w = work[blockIdx.x];
  const u32 p = w.x;
  const u32 segBeg = vOff[p] + w.y;
  const u32 segLim = vOff[p + 1];
  const u32 segEnd = (segLim - segBeg > kSparseChunk) ? segBeg + kSparseChunk : segLim;

  for (u32 i = threadIdx.x; i < 16u * K4; i += kResThreads) {
    const u32 r = i / K4, c = i - r * K4;
    const u32 ri = p * 16u + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ri < nR) {
      const u32 row = R[ri];
      if (row < M) v = __ldg(A4 + (size_t)row * K4 + c);
    }
    sA[i] = v;
  }
  __syncthreads();

  const u32 grp = threadIdx.x / kResLanes, gl = threadIdx.x % kResLanes;
  constexpr u32 kGroups = kResThreads / kResLanes;
  for (u32 e = segBeg + grp; e < segEnd; e += kGroups) {
    const u32 r = sRows[e];
    const u32 col = sCols[e];
    const float4* __restrict__ b = B4 + (size_t)col * K4;
    const float4* a = sA + r * K4;
    float acc0 = 0.f, acc1 = 0.f;
    u32 c = gl;
    for (; c + kResLanes < K4; c += 2 * kResLanes) {
      const float4 b0 = ld_b<kL1>(b + c);
      const float4 b1 = ld_b<kL1>(b + c + kResLanes);
      const float4 a0 = a[c];
      const float4 a1 = a[c + kResLanes];
      acc0 = fmaf(a0.x, b0.x, acc0); acc0 = fmaf(a0.y, b0.y, acc0);
      acc0 = fmaf(a0.z, b0.z, acc0); acc0 = fmaf(a0.w, b0.w, acc0);
      acc1 = fmaf(a1.x, b1.x, acc1); acc1 = fmaf(a1.y, b1.y, acc1);
      acc1 = fmaf(a1.z, b1.z, acc1); acc1 = fmaf(a1.w, b1.w, acc1);
    }
    if (c < K4) {
      const float4 b0 = ld_b<kL1>(b + c);
      const float4 a0 = a[c];
      acc0 = fmaf(a0.x, b0.x, acc0); acc0 = fmaf(a0.y, b0.y, acc0);
      acc0 = fmaf(a0.z, b0.z, acc0); acc0 = fmaf(a0.w, b0.w, acc0);
    }
    float acc = acc0 + acc1;
    // the four 8-lane groups of a warp run different trip counts: name only this group's lanes
    const unsigned gmask = 0xFFu << (threadIdx.x & 24u);
    acc += __shfl_xor_sync(gmask, acc, 4);
    acc += __shfl_xor_sync(gmask, acc, 2);
    acc += __shfl_xor_sync(gmask, acc, 1);
    if (gl == 0) P[sVals[e]] = acc;
  }
}


// ---- L2 eviction-policy loads / stores (sm_80+: createpolicy + .L2::cache_hint) -------------------------
__device__ __forceinline__ u64 l2_policy_evict_last() {
  u64 p;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ u64 l2_policy_evict_first() {
  u64 p;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ldg_f4_policy(const float4* p, u64 pol) {
  float4 v;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
      : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ u32 ldg_u32_policy(const u32* p, u64 pol) {
  u32 v;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void stg_f32_policy(float* p, float v, u64 pol) {
  asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(pol) : "memory");
}

// =============================================================================================
// residual kernel, super-panel form (K7b)
//   CTA = one segment of one super-panel (G row panels) whose A rows (16*G x K) sit in shared memory.
//   The super-panel's residual entries are sorted by (column, row); 8 lanes walk a CONTIGUOUS slice of the
//   segment, so the entries of one column meet the same 8 lanes back to back: the gathered B^T row is
//   fetched from L2 once (128-bit loads, 128 B per request) and then reused from registers.
//   Control flow is warp-uniform (every group runs the same number of 8-entry blocks, tails are
//   predicated), so all shuffles use the full mask; lane t of a group keeps the result of entry t of the
//   block and stores it itself.  L2->SM bytes per entry drop from 4K+16 to ~4K/(entries per column)+10.
// =============================================================================================
// kHints: gathered B^T rows are loaded with an L2 evict_last policy (and bypass L1), metadata and P stream with
// evict_first -- for matrices whose B does not fit the L2 (graphs), so that the hub columns stay resident.
// U > 0 ("gather mode", chosen when a (super-panel, column) run averages < 1.5 entries, i.e. graphs): there is no
// column run to reuse, so every entry loads its own B^T fragment and the loads of U consecutive entries are issued
// before the first FMA -- U x 4K bytes in flight per 8 lanes instead of one row (ncu, R-MAT scale 22: 24 of 30
// cycles per issue were long-scoreboard stalls with one row in flight).  U = 0: column-run reuse from registers.
// kHalfA (opt-in, sddmm_plan.operands = SDDMM_OPERANDS_FP16): the A tile sits in shared memory as fp16 (round to
// nearest; 11 significant bits, what the reference's TF32 dense path keeps), products and sums stay fp32.  Halves
// the shared-memory bytes per non-zero -- the limiter of reuse mode -- and doubles the rows per super-panel, hence
// the column-run reuse of B^T.  Values stay within the reference's checkData tolerance (tests), but they are no
// longer the fp32 results of the exact path, and |A| must stay below 65504.
// kHalfB (with kHalfA, K >= 64): the gathered B^T rows also come from an fp16 copy (k_round_b_half, once per pass), so
// the L2 -> SM gather -- the limiter on graphs -- halves as well: a lane holds 8 consecutive K values of both
// operands per 128-bit word and multiplies them with the mixed-precision fma.rn.f32.f16 (SASS FHFMA: fp16 x fp16
// product, exact in fp32, added to an fp32 accumulator) -- no conversion instructions at all.
__device__ __forceinline__ void fhfma2(float& acc, u32 a, u32 b) {
  asm("{\n\t.reg .b16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\t"
      "fma.rn.f32.f16 %0, al, bl, %0;\n\tfma.rn.f32.f16 %0, ah, bh, %0;\n\t}"
      : "+f"(acc)
      : "r"(a), "r"(b));
}
__device__ __forceinline__ void fhfma8(float& acc0, float& acc1, const uint4& a, const uint4& b) {
  fhfma2(acc0, a.x, b.x); fhfma2(acc1, a.y, b.y); fhfma2(acc0, a.z, b.z); fhfma2(acc1, a.w, b.w);
}
__device__ __forceinline__ uint4 ldg_u4_policy(const uint4* p, u64 pol) {
  uint4 v;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
      : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
      : "l"(p), "l"(pol));
  return v;
}

template <int NB, int kThreads, bool kHints, int U, bool kHalfA, bool kHalfB = false>
static __global__ void __launch_bounds__(kThreads, 1)
k_sddmm_residual_sp(u32 M, const float4* __restrict__ A4, const float4* __restrict__ B4, const u32* __restrict__ R,
                    u32 nR, u32 spRows, u32 segLen, const u32* __restrict__ spOff, const u32* __restrict__ spCol,
                    const unsigned short* __restrict__ spRow, const u32* __restrict__ spIdx,
                    const uint2* __restrict__ work, float* __restrict__ P, BatchStrides bs, u32 colMask) {
  extern __shared__ float4 sA[];  // spRows x K4 (fp32), or the same count of 4-half groups when kHalfA
  uint2* sH = reinterpret_cast<uint2*>(sA);
  A4 += (bs.a >> 2) * blockIdx.y;
  B4 += (bs.b >> 2) * blockIdx.y;
  P += bs.p * blockIdx.y;
  constexpr u32 K4 = 8 * NB;
  constexpr u32 kGroups = kThreads / 8;
  const uint2 w = work[blockIdx.x];
  const u32 sp = w.x;
  const u32 segBeg = spOff[sp] + w.y;
  const u32 segLim = spOff[sp + 1];
  const u32 segEnd = (segLim - segBeg > segLen) ? segBeg + segLen : segLim;

  for (u32 i = threadIdx.x; i < spRows * K4; i += kThreads) {
    const u32 r = i / K4, c = i - r * K4;
    const u32 ri = sp * spRows + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ri < nR) {
      const u32 row = R[ri];
      if (row < M) v = __ldg(A4 + (size_t)row * K4 + c);
    }
    if (kHalfA) {
      const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
      sH[i] = make_uint2(*reinterpret_cast<const u32*>(&lo), *reinterpret_cast<const u32*>(&hi));
    } else {
      sA[i] = v;
    }
  }
  __syncthreads();
  // 4 consecutive A values of the group's slice: one LDS.128 (fp32 tile) or one LDS.64 + 2 conversions (fp16 tile)
  auto lda = [&](u32 idx) -> float4 {
    if (!kHalfA) return sA[idx];
    const uint2 h = sH[idx];
    const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&h.x));
    const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
  };

  static_assert(!kHalfB || (kHalfA && NB >= 2), "fp16 B^T rows need the fp16 A tile and K >= 64");
  constexpr int NBH = kHalfB ? NB / 2 : 1;           // 128-bit words (8 halves) per lane and row
  constexpr u32 KH = 4 * NB;                          // 128-bit words per fp16 row
  const uint4* __restrict__ Bh = reinterpret_cast<const uint4*>(B4);  // kHalfB: B4 is the fp16 copy
  const uint4* sA8 = reinterpret_cast<const uint4*>(sA);
  const u32 grp = threadIdx.x >> 3, gl = threadIdx.x & 7u;
  u64 polB = 0, polS = 0;
  if (kHints) { polB = l2_policy_evict_last(); polS = l2_policy_evict_first(); }
  auto ld_meta4 = [&](const uint4* p) {
    if (!kHints) return __ldg(p);
    uint4 v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
        : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(polS));
    return v;
  };
  auto ld_meta1 = [&](const u32* p) { return kHints ? ldg_u32_policy(p, polS) : __ldg(p); };
  // Groups own 8-aligned (absolute index) slices of the segment, so metadata comes in aligned 8-entry blocks:
  // two 16-byte loads of columns and one of rows, the same addresses for the 8 lanes of a group.  Only the
  // first / last group see entries of a neighbouring segment in their blocks; those are masked out.
  const u32 aStart = segBeg & ~7u;
  const u32 per = (((segEnd - aStart + kGroups - 1) / kGroups) + 7u) & ~7u;  // multiple of 8
  const u32 nBlocks = per >> 3;                                              // uniform over the CTA
  const u32 aBeg = min(aStart + grp * per, (segEnd + 7u) & ~7u);
  const u32 gBeg = max(aBeg, segBeg);
  const u32 gEnd = min(segEnd, aBeg + per);
  float4 breg[kHalfB ? 1 : NB];
  uint4 bregH[NBH];
#pragma unroll
  for (int j = 0; j < (kHalfB ? 1 : NB); ++j) breg[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int j = 0; j < NBH; ++j) bregH[j] = make_uint4(0u, 0u, 0u, 0u);
  u32 prevCol = 0xFFFFFFFFu;
  const uint4* col4 = reinterpret_cast<const uint4*>(spCol);
  const uint4* row4 = reinterpret_cast<const uint4*>(spRow);
  uint4 c0 = make_uint4(~0u, ~0u, ~0u, ~0u), c1 = c0, r8 = make_uint4(0, 0, 0, 0);
  u32 mi = 0;
  if (aBeg < gEnd) {
    c0 = ld_meta4(col4 + (aBeg >> 2));
    c1 = ld_meta4(col4 + (aBeg >> 2) + 1);
    r8 = ld_meta4(row4 + (aBeg >> 3));
    mi = ld_meta1(spIdx + aBeg + gl);
  }
  for (u32 blk = 0; blk < nBlocks; ++blk) {
    const u32 base = aBeg + blk * 8u;
    uint4 nc0 = make_uint4(~0u, ~0u, ~0u, ~0u), nc1 = nc0, nr8 = make_uint4(0, 0, 0, 0);
    u32 ni = 0;
    if (base + 8u < gEnd) {  // next block's metadata
      nc0 = ld_meta4(col4 + ((base + 8u) >> 2));
      nc1 = ld_meta4(col4 + ((base + 8u) >> 2) + 1);
      nr8 = ld_meta4(row4 + ((base + 8u) >> 3));
      ni = ld_meta1(spIdx + base + 8u + gl);
    }
    const u32 cols[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
    const u32 rows[8] = {r8.x & 0xFFFFu, r8.x >> 16, r8.y & 0xFFFFu, r8.y >> 16,
                         r8.z & 0xFFFFu, r8.z >> 16, r8.w & 0xFFFFu, r8.w >> 16};
    float acc[8];
    if constexpr (U > 0 && kHalfB) {
      constexpr int UH = U * 2 > 8 ? 8 : U * 2;  // same bytes in flight per lane as the fp32 form
#pragma unroll
      for (int h = 0; h < 8 / UH; ++h) {
        uint4 bq[UH][NBH];
#pragma unroll
        for (int u = 0; u < UH; ++u) {
          const int t = h * UH + u;
          const u32 e = base + t;
          const bool live = e >= gBeg && e < gEnd;
          const uint4* __restrict__ b = Bh + (size_t)(cols[t] & colMask) * KH + gl;
          const u64 pol = (cols[t] & ~colMask) ? polS : polB;
#pragma unroll
          for (int j = 0; j < NBH; ++j)
            bq[u][j] = live ? (kHints ? ldg_u4_policy(b + j * 8, pol) : __ldg(b + j * 8)) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < UH; ++u) {
          const int t = h * UH + u;
          const u32 a = rows[t] * KH + gl;
          float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
          for (int j = 0; j < NBH; ++j) fhfma8(acc0, acc1, sA8[a + j * 8], bq[u][j]);
          acc[t] = acc0 + acc1;
        }
      }
    } else if constexpr (kHalfB) {
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const u32 e = base + t;
        const bool live = e >= gBeg && e < gEnd;
        const u32 col = cols[t];
        if (live && col != prevCol) {  // this group moves on to a new column
          const uint4* __restrict__ b = Bh + (size_t)(col & colMask) * KH + gl;
          const u64 pol = (col & ~colMask) ? polS : polB;
#pragma unroll
          for (int j = 0; j < NBH; ++j) bregH[j] = kHints ? ldg_u4_policy(b + j * 8, pol) : __ldg(b + j * 8);
          prevCol = col;
        }
        const u32 a = rows[t] * KH + gl;
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
        for (int j = 0; j < NBH; ++j) fhfma8(acc0, acc1, sA8[a + j * 8], bregH[j]);
        acc[t] = acc0 + acc1;
      }
    } else if constexpr (U > 0) {
#pragma unroll
      for (int h = 0; h < 8 / U; ++h) {
        float4 bq[U][NB];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int t = h * U + u;
          const u32 e = base + t;
          const bool live = e >= gBeg && e < gEnd;
          const float4* __restrict__ b = B4 + (size_t)(cols[t] & colMask) * K4 + gl;
          const u64 pol = (cols[t] & ~colMask) ? polS : polB;  // tail column: stream through the L2
          // (reusing the previous entry's fragment when the column repeats was tried: the register copy has to
          //  wait for that entry's load, and in-order issue then serialises the loads behind it -- 2x slower)
#pragma unroll
          for (int j = 0; j < NB; ++j)
            bq[u][j] = live ? (kHints ? ldg_f4_policy(b + j * 8, pol) : __ldg(b + j * 8)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int t = h * U + u;
          const u32 a = rows[t] * K4 + gl;
          float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            const float4 av = lda(a + j * 8), bv = bq[u][j];
            if (j & 1) {
              acc1 = fmaf(av.x, bv.x, acc1); acc1 = fmaf(av.y, bv.y, acc1);
              acc1 = fmaf(av.z, bv.z, acc1); acc1 = fmaf(av.w, bv.w, acc1);
            } else {
              acc0 = fmaf(av.x, bv.x, acc0); acc0 = fmaf(av.y, bv.y, acc0);
              acc0 = fmaf(av.z, bv.z, acc0); acc0 = fmaf(av.w, bv.w, acc0);
            }
          }
          acc[t] = acc0 + acc1;
        }
      }
    } else {
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const u32 e = base + t;
      const bool live = e >= gBeg && e < gEnd;
      const u32 col = cols[t];
      if (live && col != prevCol) {  // this group moves on to a new column
        const float4* __restrict__ b = B4 + (size_t)(col & colMask) * K4 + gl;
        const u64 pol = (col & ~colMask) ? polS : polB;  // tail column: stream through the L2
#pragma unroll
        for (int j = 0; j < NB; ++j) breg[j] = kHints ? ldg_f4_policy(b + j * 8, pol) : __ldg(b + j * 8);
        prevCol = col;
      }
      const u32 a = rows[t] * K4 + gl;
      float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const float4 av = lda(a + j * 8);
        if (j & 1) {
          acc1 = fmaf(av.x, breg[j].x, acc1); acc1 = fmaf(av.y, breg[j].y, acc1);
          acc1 = fmaf(av.z, breg[j].z, acc1); acc1 = fmaf(av.w, breg[j].w, acc1);
        } else {
          acc0 = fmaf(av.x, breg[j].x, acc0); acc0 = fmaf(av.y, breg[j].y, acc0);
          acc0 = fmaf(av.z, breg[j].z, acc0); acc0 = fmaf(av.w, breg[j].w, acc0);
        }
      }
      acc[t] = acc0 + acc1;
    }
    }
    // transposing butterfly: 7 shuffles reduce 8 entries over the 8 lanes; lane gl ends with entry gl
    float b4[4], b2[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = (gl & 4u) ? acc[i] : acc[i + 4];
      const float keep = (gl & 4u) ? acc[i + 4] : acc[i];
      b4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = (gl & 2u) ? b4[i] : b4[i + 2];
      const float keep = (gl & 2u) ? b4[i + 2] : b4[i];
      b2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2, 8);
    }
    const float send = (gl & 1u) ? b2[0] : b2[1];
    const float keep = (gl & 1u) ? b2[1] : b2[0];
    const float mine = keep + __shfl_xor_sync(0xffffffffu, send, 1, 8);
    const u32 me = base + gl;
    if (me >= gBeg && me < gEnd) {
      if (kHints) stg_f32_policy(P + mi, mine, polS);
      else P[mi] = mine;
    }
    c0 = nc0; c1 = nc1; r8 = nr8; mi = ni;
  }
}

// =============================================================================================
// residual kernel, row-stream form (K7c) -- for matrices whose residual has (almost) no column reuse inside a
// super-panel (graphs: B is gigabytes, every entry gathers its own B^T row from L2 / HBM, and the super-panel
// kernel keeps only ONE such row in flight per 8 lanes: ncu on R-MAT scale 22 shows long-scoreboard stalls at 24
// of 30 cycles per issue, DRAM at 50 %).  Here:
//   * entries are walked in reordered-ROW order (private StreamLayout), so a row's A fragment stays in registers
//     across its entries -- the L1/shared data pipe carries B only;
//   * LANES lanes (K/4 float4 split over 8, 16 or 32 lanes, NB float4 each) work on one entry; every lane issues the
//     gathered B^T loads of U = 8/NB consecutive entries before the first FMA, i.e. ~4 KB in flight per warp;
//   * B^T rows are loaded with an L2 evict_last policy and bypass L1; metadata and the P stores stream with
//     evict_first, so the hub columns of a power-law graph stay L2-resident instead of being washed out;
//   * a warp takes 32 consecutive entries (coalesced metadata), lane l ends up with the result of entry l through
//     a transposing butterfly (31 shuffles per 32 entries) and stores it.  Control flow is warp-uniform.
// Replaces src/sddmmKernel.cu:1994-2104 / :2109-2199 like the other residual kernels.
// =============================================================================================
template <int LANES, int NB>
static __global__ void __launch_bounds__(256)
k_sddmm_residual_stream(const float4* __restrict__ A4, const float4* __restrict__ B4, const u32* __restrict__ stRow,
                        const u32* __restrict__ stCol, const u32* __restrict__ stIdx, u32 n, float* __restrict__ P,
                        BatchStrides bs) {
  A4 += (bs.a >> 2) * blockIdx.y;
  B4 += (bs.b >> 2) * blockIdx.y;
  P += bs.p * blockIdx.y;
  constexpr u32 K4 = LANES * NB;            // float4 per row
  constexpr int U = NB == 1 ? 8 : NB == 2 ? 4 : 2;  // entries whose B^T fragments are in flight per lane
  constexpr int BODIES = LANES / 8;         // 8-entry bodies a group walks per 32-entry batch
  const u32 lane = threadIdx.x & 31u, gl = lane % LANES, g0 = lane - gl;
  const u64 polB = l2_policy_evict_last(), polS = l2_policy_evict_first();
  const u32 numBatches = (n + 31u) >> 5;
  const u32 warpsTotal = gridDim.x * (blockDim.x >> 5), gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const u32 per = (numBatches + warpsTotal - 1) / warpsTotal;
  const u32 bBeg = min(gw * per, numBatches), bEnd = min(bBeg + per, numBatches);

  u32 prevRow = kNull;
  float4 a[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) a[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  u32 myRow = 0, myCol = 0, myIdx = 0;
  if (bBeg < bEnd) {  // metadata arrays are padded by 32 entries
    const u32 e = bBeg * 32u + lane;
    myRow = ldg_u32_policy(stRow + e, polS);
    myCol = ldg_u32_policy(stCol + e, polS);
    myIdx = ldg_u32_policy(stIdx + e, polS);
  }
  for (u32 b = bBeg; b < bEnd; ++b) {
    u32 nRow = 0, nCol = 0, nIdx = 0;
    if (b + 1 < bEnd) {  // next batch's metadata
      const u32 e = (b + 1) * 32u + lane;
      nRow = ldg_u32_policy(stRow + e, polS);
      nCol = ldg_u32_policy(stCol + e, polS);
      nIdx = ldg_u32_policy(stIdx + e, polS);
    }
    const u32 left = n - b * 32u;  // entries of this batch that exist (>= 1)
    float v[BODIES];
#pragma unroll
    for (int body = 0; body < BODIES; ++body) {
      float acc[8];
#pragma unroll
      for (int h = 0; h < 8 / U; ++h) {
        float4 bq[U][NB];
        u32 rows[U];
        bool live[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const u32 src = g0 + (u32)(body * 8 + h * U + u);  // lane that owns this entry's metadata
          const u32 col = __shfl_sync(0xffffffffu, myCol, src);
          rows[u] = __shfl_sync(0xffffffffu, myRow, src);
          live[u] = src < left;
          const float4* __restrict__ bp = B4 + (size_t)col * K4 + gl;
#pragma unroll
          for (int j = 0; j < NB; ++j)
            bq[u][j] = live[u] ? ldg_f4_policy(bp + j * LANES, polB) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (live[u] && rows[u] != prevRow) {  // uniform inside the group
            const float4* __restrict__ ap = A4 + (size_t)rows[u] * K4 + gl;
#pragma unroll
            for (int j = 0; j < NB; ++j) a[j] = __ldg(ap + j * LANES);
            prevRow = rows[u];
          }
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            const float4 av = a[j], bv = bq[u][j];
            if (j & 1) {
              s1 = fmaf(av.x, bv.x, s1); s1 = fmaf(av.y, bv.y, s1); s1 = fmaf(av.z, bv.z, s1); s1 = fmaf(av.w, bv.w, s1);
            } else {
              s0 = fmaf(av.x, bv.x, s0); s0 = fmaf(av.y, bv.y, s0); s0 = fmaf(av.z, bv.z, s0); s0 = fmaf(av.w, bv.w, s0);
            }
          }
          acc[h * U + u] = s0 + s1;
        }
      }
      // transposing butterfly over lane bits 0..2: lane ends with entry (gl & 7) of this body, summed over the 8
      // lanes that share its upper bits
      float b4[4], b2[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float send = (gl & 4u) ? acc[i] : acc[i + 4];
        const float keep = (gl & 4u) ? acc[i + 4] : acc[i];
        b4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float send = (gl & 2u) ? b4[i] : b4[i + 2];
        const float keep = (gl & 2u) ? b4[i + 2] : b4[i];
        b2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
      }
      const float send = (gl & 1u) ? b2[0] : b2[1];
      const float keep = (gl & 1u) ? b2[1] : b2[0];
      v[body] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
    float mine = v[0];
    if constexpr (BODIES >= 2) {  // bit 3 picks the body inside a pair, the partner holds the other 8-lane partial
      float w[BODIES / 2];
#pragma unroll
      for (int i = 0; i < BODIES / 2; ++i) {
        const float send = (gl & 8u) ? v[2 * i] : v[2 * i + 1];
        const float keep = (gl & 8u) ? v[2 * i + 1] : v[2 * i];
        w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
      mine = w[0];
      if constexpr (BODIES == 4) {
        const float send = (gl & 16u) ? w[0] : w[1];
        const float keep = (gl & 16u) ? w[1] : w[0];
        mine = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
    }
    if (lane < left) stg_f32_policy(P + myIdx, mine, polS);
    myRow = nRow; myCol = nCol; myIdx = nIdx;
  }
}

// panels per super-panel for a given K (A tile <= ~192 KB), 0 if the super-panel kernel does not apply
static u32 superpanel_G(u32 K, bool halfA = false) {
  if (K % 32u || K > 512u) return 0;
  const u32 NB = K / 32u;
  if (NB != 1 && NB != 2 && NB != 4 && NB != 8 && NB != 16) return 0;
  u32 tileKB = 192u;
  if (const char* e = getenv("SDDMM_B200_SP_SMEM_KB")) { const int v = atoi(e); if (v >= 16 && v <= 216) tileKB = (u32)v; }
  u32 rows = (tileKB * 1024u) / (K * (halfA ? 2u : 4u));
  u32 G = rows / 16u;
  if (G > (halfA ? 128u : 64u)) G = halfA ? 128u : 64u;  // row ids inside a super-panel are 16-bit
  return G;
}

// =============================================================================================
// dense kernel: tcgen05 / TMEM
// =============================================================================================
constexpr int kDnThreads = 128;               // 4 warps == the 4 TMEM lane quarters
constexpr u32 kDnKChunk = 32;                 // floats per K step = one 128-byte swizzle row
constexpr u32 kDnRowsB = kDenseGroupBlocks * 16;  // 128 gathered B^T rows (MMA M)
constexpr u32 kDnStageBytes = kDnRowsB * 128 + 16 * 128;  // 16 KB + 2 KB, multiple of 1024
constexpr u32 kDnStages = 2;
constexpr u32 kDnTmemCols = 32;

__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ u32 mbar_try_wait(u64* bar, u32 parity) {
  u32 ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4 (unused for swizzled K-major),
//   [32,46) stride byte offset >> 4 (8 rows x 128 B = 1024), [46,48) version = 1, [61,64) layout 2.
__device__ __forceinline__ u64 umma_desc_sw128(u32 smemAddr) {
  u64 d = 0;
  d |= (u64)((smemAddr & 0x3FFFFu) >> 4);
  d |= (u64)(1024u >> 4) << 32;
  d |= (u64)1 << 46;
  d |= (u64)2 << 61;
  return d;
}
// kind::tf32 instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6), a/b=TF32 [7,10)/[10,13),
// K-major both, N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr u32 umma_idesc_tf32(u32 Mdim, u32 Ndim) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((Ndim >> 3) << 17) | ((Mdim >> 4) << 24);
}

__device__ __forceinline__ float tf32_rna(float x) {
  u32 r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// stores one 16-byte chunk of row `row` (128 B per row) at its 128B-swizzled position
__device__ __forceinline__ void st_swizzled(unsigned char* tile, u32 row, u32 chunk, float4 v) {
  const u32 off = (row >> 3) * 1024u + (row & 7u) * 128u + ((chunk ^ (row & 7u)) << 4);
  *reinterpret_cast<float4*>(tile + off) = v;
}

static __global__ void __launch_bounds__(kDnThreads)
k_sddmm_dense(u32 M, u32 N, u32 K, const float* __restrict__ A, const float* __restrict__ B,
              const u32* __restrict__ R, u32 nR, const u32* __restrict__ denseCols,
              const u32* __restrict__ blockOffsets, const u32* __restrict__ blockValues,
              const uint2* __restrict__ work, float* __restrict__ P, BatchStrides bs) {
  extern __shared__ __align__(1024) unsigned char smemRaw[];
  A += bs.a * blockIdx.y;
  B += bs.b * blockIdx.y;
  P += bs.p * blockIdx.y;
  // carve: stages (1024-aligned), then barriers / indices
  unsigned char* stages = smemRaw + ((1024u - (smem_u32(smemRaw) & 1023u)) & 1023u);
  __shared__ u64 mbar[kDnStages];
  __shared__ u32 tmemBase;
  __shared__ u32 sCols[kDnRowsB];
  __shared__ u32 sRowsA[16];

  const u32 tid = threadIdx.x, warp = tid >> 5;
  const uint2 w = work[blockIdx.x];
  const u32 p = w.x, firstBlk = w.y;
  const u32 blkBeg = blockOffsets[p], blkEnd = blockOffsets[p + 1];
  const u32 nBlk = min(kDenseGroupBlocks, blkEnd - blkBeg - firstBlk);
  const u32 colBase = (blkBeg + firstBlk) * 16u;  // denseColOffsets[p] == blockOffsets[p]*16

  // gathered column / row ids (sentinel N / missing rows -> zero rows, sddmmKernel.cu:279-306)
  sCols[tid] = (tid < nBlk * 16u) ? denseCols[colBase + tid] : N;
  if (tid < 16) {
    const u32 ri = p * 16u + tid;
    sRowsA[tid] = ri < nR ? R[ri] : M;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmemBase)),
                 "r"(kDnTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (u32 s = 0; s < kDnStages; ++s) mbar_init(&mbar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const u32 tmem = tmemBase;

  const u32 numChunks = (K + kDnKChunk - 1) / kDnKChunk;
  constexpr u32 idesc = umma_idesc_tf32(128, 16);

  // register-staged, software-pipelined operand gather: the loads of chunk kc+1 are in flight while chunk
  // kc is rounded (cvt.rna.tf32), stored to its swizzled stage, fenced and handed to the tensor core.
  float4 rb[8], ra;
  auto issue_loads = [&](u32 kc) {
    const u32 k0 = kc * kDnKChunk;
#pragma unroll
    for (u32 it = 0; it < 8; ++it) {
      const u32 idx = it * kDnThreads + tid;
      const u32 col = sCols[idx >> 3];
      const u32 k = k0 + (idx & 7u) * 4u;
      rb[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col < N && k < K) rb[it] = __ldg(reinterpret_cast<const float4*>(B + (size_t)col * K + k));
    }
    const u32 arow = sRowsA[tid >> 3];
    const u32 k = k0 + (tid & 7u) * 4u;
    ra = make_float4(0.f, 0.f, 0.f, 0.f);
    if (arow < M && k < K) ra = __ldg(reinterpret_cast<const float4*>(A + (size_t)arow * K + k));
  };
  issue_loads(0);
  for (u32 kc = 0; kc < numChunks; ++kc) {
    const u32 s = kc % kDnStages;
    if (kc >= kDnStages) mbar_wait(&mbar[s], ((kc / kDnStages) - 1) & 1u);  // MMA of chunk kc-2 released the stage
    unsigned char* tileB = stages + s * kDnStageBytes;   // 128 rows x 128 B (MMA "A" operand)
    unsigned char* tileA = tileB + kDnRowsB * 128;       // 16 rows x 128 B  (MMA "B" operand)
    // ---- gathered B^T rows: 8 lanes per row, 16 B each (coalesced 128 B per row), rounded like the reference
#pragma unroll
    for (u32 it = 0; it < 8; ++it) {
      const u32 idx = it * kDnThreads + tid;
      float4 v = rb[it];
      v.x = tf32_rna(v.x); v.y = tf32_rna(v.y); v.z = tf32_rna(v.z); v.w = tf32_rna(v.w);
      st_swizzled(tileB, idx >> 3, idx & 7u, v);
    }
    {
      float4 v = ra;
      v.x = tf32_rna(v.x); v.y = tf32_rna(v.y); v.z = tf32_rna(v.z); v.w = tf32_rna(v.w);
      st_swizzled(tileA, tid >> 3, tid & 7u, v);
    }
    if (kc + 1 < numChunks) issue_loads(kc + 1);
    // generic-proxy writes -> visible to the async proxy (tensor core reads smem through it)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const u64 dB = umma_desc_sw128(smem_u32(tileB));
      const u64 dA = umma_desc_sw128(smem_u32(tileA));
#pragma unroll
      for (u32 k = 0; k < kDnKChunk / 8; ++k) {
        const u32 acc = (kc | k) ? 1u : 0u;
        // advance 8 tf32 = 32 B inside the swizzle row: +2 in the (>>4) start-address field
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
            "l"(dB + 2ull * k), "l"(dA + 2ull * k), "r"(idesc), "r"(acc)
            : "memory");
      }
      // arrives on the stage barrier once the MMAs above have finished reading shared memory
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_u32(&mbar[s]))
                   : "memory");
    }
  }
  // last commit covers every earlier MMA (commits complete in order)
  {
    const u32 last = numChunks - 1;
    mbar_wait(&mbar[last % kDnStages], (last / kDnStages) & 1u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }

  // ---- epilogue: TMEM lane = gathered column slot j, register n = panel row
  u32 acc[16];
  const u32 taddr = tmem + ((warp * 32u) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]), "=r"(acc[4]), "=r"(acc[5]), "=r"(acc[6]), "=r"(acc[7]),
        "=r"(acc[8]), "=r"(acc[9]), "=r"(acc[10]), "=r"(acc[11]), "=r"(acc[12]), "=r"(acc[13]), "=r"(acc[14]),
        "=r"(acc[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  if (tid < nBlk * 16u) {
    const u32* bv = blockValues + ((size_t)(blkBeg + firstBlk) + (tid >> 4)) * 256u + (tid & 15u);
#pragma unroll
    for (u32 r = 0; r < 16; ++r) {
      const u32 idx = __ldg(bv + r * 16u);
      if (idx != kNull) P[idx] = __uint_as_float(acc[r]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kDnTmemCols) : "memory");
  }
}


// =============================================================================================
// full-tile kernel (K8): tcgen05 GEMM tiles with a sampled epilogue
//   One CTA = one non-empty 128-row x 128-column tile of the reordered S.  MMA M axis = the 128 rows
//   (A rows gathered through reorderedRows), N axis = the 128 columns (B^T rows, contiguous), K swept in
//   128-byte chunks: tcgen05.mma.cta_group::1.kind::tf32 M=128 N=128 K=8, accumulator = 128 TMEM columns.
//   Operands are staged exactly like in k_sddmm_dense (coalesced 128-bit loads, cvt.rna.tf32 like the
//   reference's wmma::__float_to_tf32, 128B-swizzled K-major tiles, 2-stage mbarrier ring, loads of the
//   next chunk in flight).  Epilogue: TMEM lane = row; each thread walks its row's 128-bit column mask and
//   stores the selected accumulators to P through the tile's CSR-index list.
//   Tile traffic is 2 x 128 x K x 4 B for up to 16384 outputs, so unlike the 16-row dense blocks the
//   operand gather is amortised over 8 panels; chosen by sddmm_launch when S is dense enough.
// =============================================================================================
constexpr int kTlThreads = 256;
constexpr u32 kTlStageBytes = 2u * 128u * 128u;  // A tile + B tile, 16 KB each
constexpr u32 kTlStages = 2;
constexpr u32 kTlTmemCols = 128;

static __global__ void __launch_bounds__(kTlThreads)
k_sddmm_tile(u32 M, u32 N, u32 K, const float* __restrict__ A, const float* __restrict__ B,
             const u32* __restrict__ R, u32 nR, const uint4* __restrict__ tiles, const u32* __restrict__ rowMeta,
             const u32* __restrict__ entIdx, float* __restrict__ P, BatchStrides bs) {
  extern __shared__ __align__(1024) unsigned char smemRaw[];
  A += bs.a * blockIdx.y;
  B += bs.b * blockIdx.y;
  P += bs.p * blockIdx.y;
  unsigned char* stages = smemRaw + ((1024u - (smem_u32(smemRaw) & 1023u)) & 1023u);
  __shared__ u64 mbar[kTlStages];
  __shared__ u32 tmemBase;
  __shared__ u32 sRows[128];

  const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  const uint4 tile = tiles[blockIdx.x];
  const u32 row0 = tile.x * 128u, col0 = tile.y * 128u;
  if (tid < 128) {
    const u32 ri = row0 + tid;
    sRows[tid] = ri < nR ? R[ri] : M;  // rows past the end read as zeros
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmemBase)),
                 "r"(kTlTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    for (u32 s = 0; s < kTlStages; ++s) mbar_init(&mbar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const u32 tmem = tmemBase;
  const u32 numChunks = (K + kDnKChunk - 1) / kDnKChunk;
  constexpr u32 idesc = umma_idesc_tf32(128, 128);

  // 2048 float4 per chunk (1024 of A, 1024 of B): 8 per thread; unit u -> (tile, row, 16-byte chunk)
  float4 rg[8];
  auto issue_loads = [&](u32 kc) {
    const u32 k0 = kc * kDnKChunk;
#pragma unroll
    for (u32 it = 0; it < 8; ++it) {
      const u32 u = it * kTlThreads + tid;  // 0..2047
      const u32 row = (u & 1023u) >> 3, ch = u & 7u;
      const u32 k = k0 + ch * 4u;
      rg[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (u < 1024u) {
        const u32 ar = sRows[row];
        if (ar < M && k < K) rg[it] = __ldg(reinterpret_cast<const float4*>(A + (size_t)ar * K + k));
      } else {
        const u32 c = col0 + row;
        if (c < N && k < K) rg[it] = __ldg(reinterpret_cast<const float4*>(B + (size_t)c * K + k));
      }
    }
  };
  issue_loads(0);
  for (u32 kc = 0; kc < numChunks; ++kc) {
    const u32 s = kc % kTlStages;
    if (kc >= kTlStages) mbar_wait(&mbar[s], ((kc / kTlStages) - 1) & 1u);
    unsigned char* tileA = stages + s * kTlStageBytes;
    unsigned char* tileB = tileA + 128u * 128u;
#pragma unroll
    for (u32 it = 0; it < 8; ++it) {
      const u32 u = it * kTlThreads + tid;
      float4 v = rg[it];
      v.x = tf32_rna(v.x); v.y = tf32_rna(v.y); v.z = tf32_rna(v.z); v.w = tf32_rna(v.w);
      st_swizzled(u < 1024u ? tileA : tileB, (u & 1023u) >> 3, u & 7u, v);
    }
    if (kc + 1 < numChunks) issue_loads(kc + 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const u64 dA = umma_desc_sw128(smem_u32(tileA));
      const u64 dB = umma_desc_sw128(smem_u32(tileB));
#pragma unroll
      for (u32 k = 0; k < kDnKChunk / 8; ++k) {
        const u32 acc = (kc | k) ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
            "l"(dA + 2ull * k), "l"(dB + 2ull * k), "r"(idesc), "r"(acc)
            : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                       smem_u32(&mbar[s]))
                   : "memory");
    }
  }
  {
    const u32 last = numChunks - 1;
    mbar_wait(&mbar[last % kTlStages], (last / kTlStages) & 1u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  // ---- epilogue.  Phase 1: warp w reads TMEM lanes 32*(w%4).. (rows) and the column half w/4 and compacts
  // the stored entries of its rows into shared memory (the operand stages are free once the last commit
  // has arrived), in the tile's entry order (row, col).  Phase 2: the whole CTA walks the tile's entry list
  // with coalesced, independent loads of the CSR indices and scatters the values.
  float* sOut = reinterpret_cast<float*>(stages);  // <= 16384 floats = the two stages
  {
    const u32 q4 = warp & 3u, half = warp >> 2;
    const u32 r = q4 * 32u + lane;
    const u32* meta = rowMeta + (size_t)blockIdx.x * 640u + r * 5u;
    const u32 m0 = meta[0], m1 = meta[1], m2 = meta[2], m3 = meta[3];
    u32 off = meta[4] - tile.z;  // relative to the tile's first entry
    if (half) off += __popc(m0) + __popc(m1);
#pragma unroll
    for (u32 qq = 0; qq < 2; ++qq) {
      const u32 q = half * 2u + qq;
      u32 acc[32];
      const u32 taddr = tmem + ((q4 * 32u) << 16) + q * 32u;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]), "=r"(acc[4]), "=r"(acc[5]), "=r"(acc[6]),
            "=r"(acc[7]), "=r"(acc[8]), "=r"(acc[9]), "=r"(acc[10]), "=r"(acc[11]), "=r"(acc[12]), "=r"(acc[13]),
            "=r"(acc[14]), "=r"(acc[15]), "=r"(acc[16]), "=r"(acc[17]), "=r"(acc[18]), "=r"(acc[19]), "=r"(acc[20]),
            "=r"(acc[21]), "=r"(acc[22]), "=r"(acc[23]), "=r"(acc[24]), "=r"(acc[25]), "=r"(acc[26]), "=r"(acc[27]),
            "=r"(acc[28]), "=r"(acc[29]), "=r"(acc[30]), "=r"(acc[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const u32 mw = q == 0 ? m0 : q == 1 ? m1 : q == 2 ? m2 : m3;
#pragma unroll
      for (u32 b = 0; b < 32; ++b) {
        if ((mw >> b) & 1u) {
          sOut[off] = __uint_as_float(acc[b]);
          ++off;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  {
    const u32 cnt = tile.w;
    const u32* __restrict__ idx = entIdx + tile.z;
    u32 e = tid;
    for (; e + 3u * kTlThreads < cnt; e += 4u * kTlThreads) {
      const u32 i0 = __ldg(idx + e), i1 = __ldg(idx + e + kTlThreads), i2 = __ldg(idx + e + 2u * kTlThreads),
                i3 = __ldg(idx + e + 3u * kTlThreads);
      P[i0] = sOut[e];
      P[i1] = sOut[e + kTlThreads];
      P[i2] = sOut[e + 2u * kTlThreads];
      P[i3] = sOut[e + 3u * kTlThreads];
    }
    for (; e < cnt; e += kTlThreads) P[__ldg(idx + e)] = sOut[e];
  }
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTlTmemCols) : "memory");
  }
}

// =============================================================================================
// tile kernel, TMA form (K9): the operands are first rounded to TF32 (cvt.rna, as everywhere) into
// row-gathered copies Ar [numRows x K], Br [N x K]; the tile kernel then has ONE thread issue
// cp.async.bulk.tensor loads of 128 x 32-float boxes (SWIZZLE_128B, the layout the UMMA descriptors
// expect; rows past the end and the K tail are zero-filled by the TMA unit), ONE thread issue the
// tcgen05.mma's, and full/empty mbarriers between them.  The 256 threads only meet again for the
// epilogue.  Compared with k_sddmm_tile this removes the LDG -> cvt -> STS work (55 % of the issue
// slots there) from the SM.
// =============================================================================================
static __global__ void __launch_bounds__(256) k_round_operands(u32 M, u32 N, u32 K4, const float4* __restrict__ A,
                                                               const float4* __restrict__ B,
                                                               const u32* __restrict__ R, u32 nR,
                                                               float4* __restrict__ Ar, float4* __restrict__ Br,
                                                               BatchStrides bs) {
  // programmatic dependent launch: the tile kernel behind this pass may be scheduled now; it waits
  // (griddepcontrol.wait) before its first read of the rounded copies
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  A += (bs.a >> 2) * blockIdx.y;
  B += (bs.b >> 2) * blockIdx.y;
  Ar += (size_t)nR * K4 * blockIdx.y;
  Br += (size_t)N * K4 * blockIdx.y;
  const size_t nA = (size_t)nR * K4, nB = (size_t)N * K4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nA + nB; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < nA) {
      const u32 r = (u32)(i / K4), c = (u32)(i - (size_t)r * K4);
      const u32 row = R[r];
      if (row < M) v = __ldg(A + (size_t)row * K4 + c);
    } else {
      v = __ldg(B + (i - nA));
    }
    v.x = tf32_rna(v.x); v.y = tf32_rna(v.y); v.z = tf32_rna(v.z); v.w = tf32_rna(v.w);
    if (i < nA) Ar[i] = v;
    else Br[i - nA] = v;
  }
}

// fp16 form of the copies (sddmm_plan.operands = SDDMM_OPERANDS_FP16): round to nearest, half the bytes
static __global__ void __launch_bounds__(256) k_round_operands_half(u32 M, u32 N, u32 K4, const float4* __restrict__ A,
                                                                    const float4* __restrict__ B,
                                                                    const u32* __restrict__ R, u32 nR,
                                                                    uint2* __restrict__ Ar, uint2* __restrict__ Br,
                                                                    BatchStrides bs) {
  // programmatic dependent launch: the tile kernel behind this pass may be scheduled now; it waits
  // (griddepcontrol.wait) before its first read of the rounded copies
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  A += (bs.a >> 2) * blockIdx.y;
  B += (bs.b >> 2) * blockIdx.y;
  Ar += (size_t)nR * K4 * blockIdx.y;
  Br += (size_t)N * K4 * blockIdx.y;
  const size_t nA = (size_t)nR * K4, nB = (size_t)N * K4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nA + nB; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < nA) {
      const u32 r = (u32)(i / K4), c = (u32)(i - (size_t)r * K4);
      const u32 row = R[r];
      if (row < M) v = __ldg(A + (size_t)row * K4 + c);
    } else {
      v = __ldg(B + (i - nA));
    }
    const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
    const uint2 o = make_uint2(*reinterpret_cast<const u32*>(&lo), *reinterpret_cast<const u32*>(&hi));
    if (i < nA) Ar[i] = o;
    else Br[i - nA] = o;
  }
}

// fp16 copy of the listed B^T rows only (the columns the residual references): the super-panel kernel's operand
static __global__ void __launch_bounds__(256) k_round_rows_half(const u32* __restrict__ list, u32 numList, u32 K4,
                                                                const float4* __restrict__ B, uint2* __restrict__ Bh,
                                                                size_t srcStride4, size_t dstStride4) {
  B += srcStride4 * blockIdx.y;
  Bh += dstStride4 * blockIdx.y;
  const size_t n = (size_t)numList * K4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const u32 li = (u32)(i / K4), c = (u32)(i - (size_t)li * K4);
    const size_t at = (size_t)__ldg(list + li) * K4 + c;
    const float4 v = __ldg(B + at);
    const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
    Bh[at] = make_uint2(*reinterpret_cast<const u32*>(&lo), *reinterpret_cast<const u32*>(&hi));
  }
}

// bounded wait: a protocol error must end in a launch failure, never in a hung GPU
__device__ __forceinline__ void mbar_wait_bounded(u64* bar, u32 parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}

__device__ __forceinline__ void tma_load_3d(u32 dstSmem, const void* map, u32 barSmem, u32 c0, u32 c1, u32 c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dstSmem), "l"(map), "r"(barSmem), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// kHalf: fp16 operand copies, 64 K-elements per 128-byte swizzle row, tcgen05.mma kind::f16 (K = 16 per instruction)
__host__ __device__ constexpr u32 umma_idesc_f16(u32 Mdim, u32 Ndim) {
  return (1u << 4) | ((Ndim >> 3) << 17) | ((Mdim >> 4) << 24);  // c = F32, a = b = F16 (format 0), K-major both
}

template <u32 kStagesT, bool kHalf>
static __global__ void __launch_bounds__(kTlThreads)
k_sddmm_tile_tma(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, u32 K,
                 const uint4* __restrict__ tiles, const u32* __restrict__ rowMeta, const u32* __restrict__ entIdx,
                 float* __restrict__ P, size_t pStride) {
  extern __shared__ __align__(1024) unsigned char smemRaw[];
  P += pStride * blockIdx.y;
  unsigned char* stages = smemRaw + ((1024u - (smem_u32(smemRaw) & 1023u)) & 1023u);
  __shared__ u64 fullBar[kStagesT], emptyBar[kStagesT], accBar;
  __shared__ u32 tmemBase;

  const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  const uint4 tile = tiles[blockIdx.x];
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmemBase)),
                 "r"(kTlTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (u32 s = 0; s < kStagesT; ++s) {
      mbar_init(&fullBar[s], 1);
      mbar_init(&emptyBar[s], 1);
    }
    mbar_init(&accBar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const u32 tmem = tmemBase;
  constexpr u32 kChunkElems = kHalf ? 64u : kDnKChunk;  // one 128-byte swizzle row of K
  const u32 numChunks = (K + kChunkElems - 1) / kChunkElems;
  constexpr u32 idesc = kHalf ? umma_idesc_f16(128, 128) : umma_idesc_tf32(128, 128);

  if (warp == 0) {
    // ---- producer: one thread feeds the stages through the TMA unit
    for (u32 kc = 0; kc < numChunks; ++kc) {
      const u32 s = kc % kStagesT;
      if (lane == 0) {
        if (kc >= kStagesT) mbar_wait_bounded(&emptyBar[s], ((kc / kStagesT) - 1) & 1u);
        const u32 bar = smem_u32(&fullBar[s]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kTlStageBytes) : "memory");
        const u32 dst = smem_u32(stages + s * kTlStageBytes);
        tma_load_3d(dst, &mapA, bar, kc * kChunkElems, tile.x * 128u, blockIdx.y);
        tma_load_3d(dst + 128u * 128u, &mapB, bar, kc * kChunkElems, tile.y * 128u, blockIdx.y);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---- MMA issuer
    for (u32 kc = 0; kc < numChunks; ++kc) {
      const u32 s = kc % kStagesT;
      if (lane == 0) {
        mbar_wait_bounded(&fullBar[s], (kc / kStagesT) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const u32 base = smem_u32(stages + s * kTlStageBytes);
        const u64 dA = umma_desc_sw128(base);
        const u64 dB = umma_desc_sw128(base + 128u * 128u);
#pragma unroll
        for (u32 k = 0; k < kDnKChunk / 8; ++k) {
          const u32 acc = (kc | k) ? 1u : 0u;
          if (kHalf)
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                "l"(dA + 2ull * k), "l"(dB + 2ull * k), "r"(idesc), "r"(acc)
                : "memory");
          else
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
              "l"(dA + 2ull * k), "l"(dB + 2ull * k), "r"(idesc), "r"(acc)
              : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         smem_u32(&emptyBar[s]))
                     : "memory");
        if (kc + 1 == numChunks)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                           smem_u32(&accBar))
                       : "memory");
      }
      __syncwarp();
    }
  }
  mbar_wait_bounded(&accBar, 0u);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  __syncthreads();  // every MMA has finished reading the stages: they become the epilogue's staging area

  float* sOut = reinterpret_cast<float*>(stages);
  {
    const u32 q4 = warp & 3u, half = warp >> 2;
    const u32 r = q4 * 32u + lane;
    const u32* meta = rowMeta + (size_t)blockIdx.x * 640u + r * 5u;
    const u32 m0 = meta[0], m1 = meta[1], m2 = meta[2], m3 = meta[3];
    u32 off = meta[4] - tile.z;
    if (half) off += __popc(m0) + __popc(m1);
#pragma unroll
    for (u32 qq = 0; qq < 2; ++qq) {
      const u32 q = half * 2u + qq;
      u32 acc[32];
      const u32 taddr = tmem + ((q4 * 32u) << 16) + q * 32u;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]), "=r"(acc[4]), "=r"(acc[5]), "=r"(acc[6]),
            "=r"(acc[7]), "=r"(acc[8]), "=r"(acc[9]), "=r"(acc[10]), "=r"(acc[11]), "=r"(acc[12]), "=r"(acc[13]),
            "=r"(acc[14]), "=r"(acc[15]), "=r"(acc[16]), "=r"(acc[17]), "=r"(acc[18]), "=r"(acc[19]), "=r"(acc[20]),
            "=r"(acc[21]), "=r"(acc[22]), "=r"(acc[23]), "=r"(acc[24]), "=r"(acc[25]), "=r"(acc[26]), "=r"(acc[27]),
            "=r"(acc[28]), "=r"(acc[29]), "=r"(acc[30]), "=r"(acc[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const u32 mw = q == 0 ? m0 : q == 1 ? m1 : q == 2 ? m2 : m3;
#pragma unroll
      for (u32 b = 0; b < 32; ++b) {
        if ((mw >> b) & 1u) {
          sOut[off] = __uint_as_float(acc[b]);
          ++off;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  {
    const u32 cnt = tile.w;
    const u32* __restrict__ idx = entIdx + tile.z;
    u32 e = tid;
    for (; e + 3u * kTlThreads < cnt; e += 4u * kTlThreads) {
      const u32 i0 = __ldg(idx + e), i1 = __ldg(idx + e + kTlThreads), i2 = __ldg(idx + e + 2u * kTlThreads),
                i3 = __ldg(idx + e + 3u * kTlThreads);
      P[i0] = sOut[e];
      P[i1] = sOut[e + kTlThreads];
      P[i2] = sOut[e + 2u * kTlThreads];
      P[i3] = sOut[e + 3u * kTlThreads];
    }
    for (; e < cnt; e += kTlThreads) P[__ldg(idx + e)] = sOut[e];
  }
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTlTmemCols) : "memory");
  }
}

// ---- cluster form: a 2x2 group of tiles runs as one 4-CTA cluster.  The A tile of a tile row is shared by the
// two CTAs of that row, the B tile of a tile column by the two CTAs of that column: every CTA fetches HALF of
// its A tile and HALF of its B tile (64-row boxes) and multicasts each half to itself and the partner, so the
// L2 -> SM operand traffic halves.  A stage may be refilled once the local MMA and both partners' MMAs are
// done with it: the local commit arrives on emptyBar, the partners' producers arrive remotely on peerBar.
__device__ __forceinline__ u32 cluster_ctarank() {
  u32 r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(u64* bar, u32 rank) {
  u32 remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster_bounded(u64* bar, u32 parity) {
  const long long t0 = clock64();
  u32 ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!ok && clock64() - t0 > 4000000000ll) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d_mc(u32 dstSmem, const void* map, u32 barSmem, u32 c0, u32 c1, u32 c2,
                                               unsigned short mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(dstSmem), "l"(map), "r"(barSmem), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
      : "memory");
}

template <u32 kStagesT>
static __global__ void __launch_bounds__(kTlThreads)
k_sddmm_tile_tma4(const __grid_constant__ CUtensorMap mapA64, const __grid_constant__ CUtensorMap mapB64, u32 K,
                  const uint2* __restrict__ quads, const u32* __restrict__ quadTiles, const uint4* __restrict__ tiles,
                  const u32* __restrict__ rowMeta, const u32* __restrict__ entIdx, float* __restrict__ P,
                  size_t pStride) {
  extern __shared__ __align__(1024) unsigned char smemRaw[];
  P += pStride * blockIdx.y;
  unsigned char* stages = smemRaw + ((1024u - (smem_u32(smemRaw) & 1023u)) & 1023u);
  __shared__ u64 fullBar[kStagesT], emptyBar[kStagesT], peerBar[kStagesT], accBar;
  __shared__ u32 tmemBase;

  const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  const u32 rank = cluster_ctarank(), dr = rank >> 1, dc = rank & 1u;
  const u32 quad = blockIdx.x >> 2;
  const uint2 qrc = quads[quad];
  const u32 tr = qrc.x * 2u + dr, tc = qrc.y * 2u + dc;
  const u32 tileIdx = quadTiles[quad * 4u + rank];
  const uint4 tile = tileIdx != kNull ? tiles[tileIdx] : make_uint4(tr, tc, 0u, 0u);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmemBase)),
                 "r"(kTlTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (u32 s = 0; s < kStagesT; ++s) {
      mbar_init(&fullBar[s], 1);
      mbar_init(&emptyBar[s], 1);
      mbar_init(&peerBar[s], 2);  // the row partner and the column partner
    }
    mbar_init(&accBar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();  // every member's barriers exist before any multicast or remote arrive can land
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const u32 tmem = tmemBase;
  const u32 numChunks = (K + kDnKChunk - 1) / kDnKChunk;
  constexpr u32 idesc = umma_idesc_tf32(128, 128);
  const u32 rowPartner = rank ^ 1u, colPartner = rank ^ 2u;
  const unsigned short maskRow = (unsigned short)((1u << rank) | (1u << rowPartner));
  const unsigned short maskCol = (unsigned short)((1u << rank) | (1u << colPartner));

  if (warp == 0) {
    for (u32 kc = 0; kc < numChunks; ++kc) {
      const u32 s = kc % kStagesT;
      if (lane == 0) {
        if (kc >= kStagesT) {
          const u32 par = ((kc / kStagesT) - 1) & 1u;
          mbar_wait_bounded(&emptyBar[s], par);      // my MMA is done with stage s ...
          mbar_arrive_remote(&peerBar[s], rowPartner);  // ... tell the CTAs that write into it ...
          mbar_arrive_remote(&peerBar[s], colPartner);
          mbar_wait_cluster_bounded(&peerBar[s], par);  // ... and wait until theirs are free for my halves
        }
        const u32 bar = smem_u32(&fullBar[s]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kTlStageBytes) : "memory");
        const u32 dst = smem_u32(stages + s * kTlStageBytes);
        tma_load_3d_mc(dst + dc * 8192u, &mapA64, bar, kc * kDnKChunk, tr * 128u + dc * 64u, blockIdx.y, maskRow);
        tma_load_3d_mc(dst + 16384u + dr * 8192u, &mapB64, bar, kc * kDnKChunk, tc * 128u + dr * 64u, blockIdx.y,
                       maskCol);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    for (u32 kc = 0; kc < numChunks; ++kc) {
      const u32 s = kc % kStagesT;
      if (lane == 0) {
        mbar_wait_bounded(&fullBar[s], (kc / kStagesT) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const u32 base = smem_u32(stages + s * kTlStageBytes);
        const u64 dA = umma_desc_sw128(base);
        const u64 dB = umma_desc_sw128(base + 128u * 128u);
#pragma unroll
        for (u32 k = 0; k < kDnKChunk / 8; ++k) {
          const u32 acc = (kc | k) ? 1u : 0u;
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
              "l"(dA + 2ull * k), "l"(dB + 2ull * k), "r"(idesc), "r"(acc)
              : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         smem_u32(&emptyBar[s]))
                     : "memory");
        if (kc + 1 == numChunks)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                           smem_u32(&accBar))
                       : "memory");
      }
      __syncwarp();
    }
  }
  mbar_wait_bounded(&accBar, 0u);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  __syncthreads();

  float* sOut = reinterpret_cast<float*>(stages);
  if (tileIdx != kNull) {
    const u32 q4 = warp & 3u, half = warp >> 2;
    const u32 r = q4 * 32u + lane;
    const u32* meta = rowMeta + (size_t)tileIdx * 640u + r * 5u;
    const u32 m0 = meta[0], m1 = meta[1], m2 = meta[2], m3 = meta[3];
    u32 off = meta[4] - tile.z;
    if (half) off += __popc(m0) + __popc(m1);
#pragma unroll
    for (u32 qq = 0; qq < 2; ++qq) {
      const u32 q = half * 2u + qq;
      u32 acc[32];
      const u32 taddr = tmem + ((q4 * 32u) << 16) + q * 32u;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]), "=r"(acc[4]), "=r"(acc[5]), "=r"(acc[6]),
            "=r"(acc[7]), "=r"(acc[8]), "=r"(acc[9]), "=r"(acc[10]), "=r"(acc[11]), "=r"(acc[12]), "=r"(acc[13]),
            "=r"(acc[14]), "=r"(acc[15]), "=r"(acc[16]), "=r"(acc[17]), "=r"(acc[18]), "=r"(acc[19]), "=r"(acc[20]),
            "=r"(acc[21]), "=r"(acc[22]), "=r"(acc[23]), "=r"(acc[24]), "=r"(acc[25]), "=r"(acc[26]), "=r"(acc[27]),
            "=r"(acc[28]), "=r"(acc[29]), "=r"(acc[30]), "=r"(acc[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const u32 mw = q == 0 ? m0 : q == 1 ? m1 : q == 2 ? m2 : m3;
#pragma unroll
      for (u32 b = 0; b < 32; ++b) {
        if ((mw >> b) & 1u) {
          sOut[off] = __uint_as_float(acc[b]);
          ++off;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tileIdx != kNull) {
    const u32 cnt = tile.w;
    const u32* __restrict__ idx = entIdx + tile.z;
    u32 e = tid;
    for (; e + 3u * kTlThreads < cnt; e += 4u * kTlThreads) {
      const u32 i0 = __ldg(idx + e), i1 = __ldg(idx + e + kTlThreads), i2 = __ldg(idx + e + 2u * kTlThreads),
                i3 = __ldg(idx + e + 3u * kTlThreads);
      P[i0] = sOut[e];
      P[i1] = sOut[e + kTlThreads];
      P[i2] = sOut[e + 2u * kTlThreads];
      P[i3] = sOut[e + 3u * kTlThreads];
    }
    for (; e < cnt; e += kTlThreads) P[__ldg(idx + e)] = sOut[e];
  }
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTlTmemCols) : "memory");
  }
  cluster_sync_all();  // nobody leaves while a partner could still signal one of its barriers
}

// =============================================================================================
// tile kernel, CTA-pair form (K10): persistent, warp-specialised, tcgen05.mma.cta_group::2.
//   Work unit = a 2x2 group of tiles (256 rows x 256 columns of the reordered S, `quads`), computed by the two CTAs
//   of a cluster as ONE M=256 N=256 MMA: CTA r stages its own 128 A rows and HALF of the B^T rows (128 of the 256
//   columns), the tensor cores of both SMs read both halves, and each CTA's TMEM receives 128 rows x 256 columns.
//   Per 128x128 output tile a CTA therefore pulls 128 operand rows through the L2 -> SM fabric instead of 256 -- the
//   limiter of the one-tile-per-CTA form (k_sddmm_tile_tma: 256 KB per tile at K=256).
//   Roles (576 threads): warp 0 = TMA producer (both CTAs; the loads signal the LEADER's full barrier through the
//   cta_group::2 form of cp.async.bulk.tensor), warp 1 = MMA issuer (leader only; tcgen05.commit multicasts the
//   stage-free and accumulator-ready arrivals to both CTAs), warps 2..17 = epilogue (TMEM -> per-row mask compaction
//   -> shared staging -> coalesced scatter through the tile's CSR-index list, as in the other tile kernels).
//   Two 256-column accumulators (all 512 TMEM columns): the epilogue of quad i runs under the main loop of quad
//   i+1, and the operand ring never drains between quads.  Every wait is bounded (a protocol error traps).
//   Same operands as K9: TF32-rounded (or fp16) row-gathered copies, 128-row SWIZZLE_128B boxes.
// =============================================================================================
constexpr int kTpThreads = 576;           // producer warp + MMA warp + 16 epilogue warps
constexpr u32 kTpEpiThreads = 512;
constexpr u32 kTpStages = 4;
constexpr u32 kTpStagingBytes = 128u * 128u * 4u;  // one 128x128 tile's stored entries at most
constexpr u32 kTpPrefetch = 12;  // CSR indices per epilogue thread and tile requested ahead (covers 6144 entries)
// AUTO's choice between K9 and K10 for K >= 64, from measurements (4096^2 masks, one CUDA graph per pass incl. the
// rounding pre-pass, K=256, K9 / K10): 70 % sparse 35.1 / 30.7 us, 90 % 33.0 / 29.0, 50 % 40.7 / 40.5; fp16 operands
// 27.0 / 24.9 us.  K = 64: 22.5 / 22.7 us (and 23.9 for the register-staged K8, which stays the choice below 64).
constexpr bool kPairDefault = true;

__device__ __forceinline__ u32 mapa_u32(u32 smemAddr, u32 rank) {
  u32 r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smemAddr), "r"(rank));
  return r;
}
// both CTAs of a pair load into their OWN shared memory and complete the transaction on the LEADER's barrier
__device__ __forceinline__ void tma_load_3d_pair(u32 dstSmem, const void* map, u32 leaderBar, u32 c0, u32 c1, u32 c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dstSmem), "l"(map), "r"(leaderBar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(u64* bar) {  // arrives on `bar` of BOTH CTAs when the MMAs retire
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)), "h"((unsigned short)3)
      : "memory");
}

// Remote arrive without a release fence: what it orders here are TMEM reads, already fenced by tcgen05.wait::ld +
// tcgen05.fence::before_thread_sync (a .release.cluster arrive costs a MEMBAR.ALL.GPU behind the scattered P stores).
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(u64* bar, u32 rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)), "r"(rank)
      : "memory");
}
// debug timeline of the pair kernel (SDDMM_B200_PAIR_DEBUG=1): clock64 stamps of CTA 0 / 1, 8 quads x 8 events
__device__ unsigned long long g_tpDbg[2][8][8];
#define SB_TP_STAMP(t, ev)                                                                    \
  do {                                                                                        \
    if (dbg && blockIdx.x < 2 && blockIdx.y == 0 && (t) < 8u) g_tpDbg[blockIdx.x][(t)][(ev)] = clock64(); \
  } while (0)

template <bool kHalf>
static __global__ void __launch_bounds__(kTpThreads, 1)
k_sddmm_tile_pair(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, u32 K,
                  u32 numQuads, const uint2* __restrict__ quads, const u32* __restrict__ quadTiles,
                  const uint4* __restrict__ tiles, const uint2* __restrict__ rowMetaT, const u32* __restrict__ entIdx,
                  float* __restrict__ P, size_t pStride, u32 dbg) {
  extern __shared__ __align__(1024) unsigned char smemRaw[];
  P += pStride * blockIdx.y;
  unsigned char* stages = smemRaw + ((1024u - (smem_u32(smemRaw) & 1023u)) & 1023u);
  float* sOut = reinterpret_cast<float*>(stages + kTpStages * kTlStageBytes);
  __shared__ u64 fullBar[kTpStages], emptyBar[kTpStages], accFull[2], accEmpty[2];
  __shared__ u32 tmemBase;

  const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  const u32 rank = cluster_ctarank();
  const u32 pairId = blockIdx.x >> 1, numPairs = gridDim.x >> 1;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmemBase)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (u32 s = 0; s < kTpStages; ++s) {
      mbar_init(&fullBar[s], 1);   // the leader's arrive.expect_tx (bytes of both CTAs' boxes)
      mbar_init(&emptyBar[s], 1);  // the MMA commit, multicast to both CTAs
    }
    for (u32 b = 0; b < 2; ++b) {
      mbar_init(&accFull[b], 1);    // the MMA commit of a quad's last chunk, multicast
      mbar_init(&accEmpty[b], 32);  // 16 epilogue warps of each CTA (used on the leader only)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers and TMEM exist before any remote arrive / multicast commit / MMA
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const u32 tmem = tmemBase;
  constexpr u32 kChunkElems = kHalf ? 64u : kDnKChunk;
  const u32 numChunks = (K + kChunkElems - 1) / kChunkElems;
  constexpr u32 idesc = kHalf ? umma_idesc_f16(256, 256) : umma_idesc_tf32(256, 256);

  if (warp == 0) {
    // ---- producer (both CTAs): A rows of this CTA's tile row, B^T rows of this CTA's half of the 256 columns
    // (launched with programmatic stream serialization: everything above overlapped the rounding pass; its copies
    //  are complete and visible after this wait -- nothing else in the kernel reads what that pass writes)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    u32 g = 0;
    for (u32 q = pairId; q < numQuads; q += numPairs) {
      const uint2 qrc = quads[q];
      const u32 rowA = (qrc.x * 2u + rank) * 128u, rowB = (qrc.y * 2u + rank) * 128u;
      for (u32 kc = 0; kc < numChunks; ++kc, ++g) {
        const u32 s = g % kTpStages;
        if (lane == 0) {
          if (g >= kTpStages) mbar_wait_bounded(&emptyBar[s], ((g / kTpStages) - 1) & 1u);
          if (rank == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&fullBar[s])),
                         "r"(2u * kTlStageBytes)
                         : "memory");
          const u32 bar = mapa_u32(smem_u32(&fullBar[s]), 0u);
          const u32 dst = smem_u32(stages + s * kTlStageBytes);
          tma_load_3d_pair(dst, &mapA, bar, kc * kChunkElems, rowA, blockIdx.y);
          tma_load_3d_pair(dst + 128u * 128u, &mapB, bar, kc * kChunkElems, rowB, blockIdx.y);
          if (kc + 1 == numChunks) SB_TP_STAMP(g / numChunks, 0);  // producer: all boxes of the quad requested
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer (leader CTA only): one M=256 N=256 accumulator per quad, alternating TMEM halves
    if (rank == 0) {
      u32 g = 0, t = 0;
      for (u32 q = pairId; q < numQuads; q += numPairs, ++t) {
        const u32 buf = t & 1u;
        if (lane == 0 && t >= 2u) mbar_wait_bounded(&accEmpty[buf], ((t >> 1) - 1u) & 1u);
        if (lane == 0) SB_TP_STAMP(t, 1);  // MMA: accumulator half free
        __syncwarp();
        for (u32 kc = 0; kc < numChunks; ++kc, ++g) {
          const u32 s = g % kTpStages;
          if (lane == 0) {
            mbar_wait_bounded(&fullBar[s], (g / kTpStages) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const u32 base = smem_u32(stages + s * kTlStageBytes);
            const u64 dA = umma_desc_sw128(base);
            const u64 dB = umma_desc_sw128(base + 128u * 128u);
#pragma unroll
            for (u32 k = 0; k < kDnKChunk / 8; ++k) {
              const u32 acc = (kc | k) ? 1u : 0u;
              if (kHalf)
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + buf * 256u),
                    "l"(dA + 2ull * k), "l"(dB + 2ull * k), "r"(idesc), "r"(acc)
                    : "memory");
              else
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + buf * 256u),
                    "l"(dA + 2ull * k), "l"(dB + 2ull * k), "r"(idesc), "r"(acc)
                    : "memory");
            }
            umma_commit_pair(&emptyBar[s]);
            if (kc + 1 == numChunks) {
              umma_commit_pair(&accFull[buf]);
              SB_TP_STAMP(t, 2);  // MMA: last chunk issued
            }
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ---- epilogue (both CTAs, 16 warps): this CTA's 128 rows x 256 columns = the quad's tiles (rank, 0), (rank, 1).
    // Warp (q4, cq) owns TMEM lanes 32*q4.. (rows; a warp may only touch the lane quarter of its id mod 4) and the
    // 32-column slice cq of each tile.  Nothing hides a global-load round trip inside this loop (the other 16 warps
    // of the SM are all here), so everything is requested a full step ahead: the next quad's tile headers at the
    // top of a quad, and a tile's row masks and CSR indices as soon as the previous tile's registers are free.
    const u32 et = tid - 64u, q4 = warp & 3u, cq = (warp - 2u) >> 2;
    const u32 r = q4 * 32u + lane;
    u32 cIdx[2] = {kNull, kNull}, cBeg[2] = {0u, 0u}, cCnt[2] = {0u, 0u}, pf[2][kTpPrefetch];
    uint2 mk[2];  // {mask word of my 32-column slice, offset of my first entry in the tile's staging order}
    auto load_hdr = [&](u32 q, u32 (&ti)[2], u32 (&beg)[2], u32 (&cnt)[2]) {
#pragma unroll
      for (u32 sub = 0; sub < 2; ++sub) {
        ti[sub] = q < numQuads ? __ldg(quadTiles + q * 4u + rank * 2u + sub) : kNull;
        beg[sub] = cnt[sub] = 0u;
        if (ti[sub] != kNull) {
          const uint4 tile = __ldg(tiles + ti[sub]);
          beg[sub] = tile.z;
          cnt[sub] = tile.w;
        }
      }
    };
    auto request = [&](u32 sub) {  // row masks + offset of my row, my first kTpPrefetch CSR indices
      if (cIdx[sub] != kNull) {
        mk[sub] = __ldg(rowMetaT + ((size_t)cIdx[sub] * 4u + cq) * 128u + r);
#pragma unroll
        for (u32 j = 0; j < kTpPrefetch; ++j) {
          const u32 e = et + j * kTpEpiThreads;
          pf[sub][j] = e < cCnt[sub] ? __ldg(entIdx + cBeg[sub] + e) : 0u;
        }
      }
    };
    load_hdr(pairId, cIdx, cBeg, cCnt);
    request(0);
    request(1);
    u32 t = 0;
    for (u32 q = pairId; q < numQuads; q += numPairs, ++t) {
      const u32 buf = t & 1u;
      u32 nIdx[2], nBeg[2], nCnt[2];
      load_hdr(q + numPairs, nIdx, nBeg, nCnt);
      const u32 arriveAt = cIdx[1] != kNull ? 1u : 0u;  // ONE accEmpty arrive per warp and quad, after its last TMEM read
      if (et == 0) SB_TP_STAMP(t, 3);  // epilogue: waiting for the accumulator
      mbar_wait_bounded(&accFull[buf], (t >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (et == 0) SB_TP_STAMP(t, 4);  // epilogue: accumulator ready
#pragma unroll
      for (u32 sub = 0; sub < 2; ++sub) {
        if (cIdx[sub] != kNull) {
          const u32 off = mk[sub].y, mw = mk[sub].x;
          u32 acc[32];
          const u32 taddr = tmem + ((q4 * 32u) << 16) + buf * 256u + sub * 128u + cq * 32u;
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
              : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]), "=r"(acc[4]), "=r"(acc[5]), "=r"(acc[6]),
                "=r"(acc[7]), "=r"(acc[8]), "=r"(acc[9]), "=r"(acc[10]), "=r"(acc[11]), "=r"(acc[12]), "=r"(acc[13]),
                "=r"(acc[14]), "=r"(acc[15]), "=r"(acc[16]), "=r"(acc[17]), "=r"(acc[18]), "=r"(acc[19]),
                "=r"(acc[20]), "=r"(acc[21]), "=r"(acc[22]), "=r"(acc[23]), "=r"(acc[24]), "=r"(acc[25]),
                "=r"(acc[26]), "=r"(acc[27]), "=r"(acc[28]), "=r"(acc[29]), "=r"(acc[30]), "=r"(acc[31])
              : "r"(taddr)
              : "memory");
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          // two independent offset chains (low / high 16 columns) instead of one of 32
          u32 offLo = off, offHi = off + __popc(mw & 0xFFFFu);
#pragma unroll
          for (u32 b = 0; b < 16; ++b) {
            if ((mw >> b) & 1u) {
              sOut[offLo] = __uint_as_float(acc[b]);
              ++offLo;
            }
            if ((mw >> (b + 16u)) & 1u) {
              sOut[offHi] = __uint_as_float(acc[b + 16u]);
              ++offHi;
            }
          }
        }
        if (sub == arriveAt) {  // the accumulator half is free for the quad after next
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_relaxed(&accEmpty[buf], 0u);
        }
        if (cIdx[sub] != kNull) {
          asm volatile("bar.sync 1, 512;" ::: "memory");  // the tile's values are staged
          if (et == 0) SB_TP_STAMP(t, 5 + sub);  // epilogue: tile `sub` compacted
          const u32 cnt = cCnt[sub];
          float v[kTpPrefetch];
#pragma unroll
          for (u32 j = 0; j < kTpPrefetch; ++j) {
            const u32 e = et + j * kTpEpiThreads;
            v[j] = e < cnt ? sOut[e] : 0.f;
          }
#pragma unroll
          for (u32 j = 0; j < kTpPrefetch; ++j) {
            const u32 e = et + j * kTpEpiThreads;
            if (e < cnt) P[pf[sub][j]] = v[j];
          }
          if (cnt > kTpPrefetch * kTpEpiThreads) {  // denser than the prefetch depth covers: the rest with loads here
            const u32* __restrict__ idx = entIdx + cBeg[sub];
            u32 e = et + kTpPrefetch * kTpEpiThreads;
            for (; e + 3u * kTpEpiThreads < cnt; e += 4u * kTpEpiThreads) {
              const u32 i0 = __ldg(idx + e), i1 = __ldg(idx + e + kTpEpiThreads), i2 = __ldg(idx + e + 2u * kTpEpiThreads),
                        i3 = __ldg(idx + e + 3u * kTpEpiThreads);
              P[i0] = sOut[e];
              P[i1] = sOut[e + kTpEpiThreads];
              P[i2] = sOut[e + 2u * kTpEpiThreads];
              P[i3] = sOut[e + 3u * kTpEpiThreads];
            }
            for (; e < cnt; e += kTpEpiThreads) P[__ldg(idx + e)] = sOut[e];
          }
          asm volatile("bar.sync 1, 512;" ::: "memory");  // staging area free again
        }
        // this tile's registers are free: request the same tile of the NEXT quad
        cIdx[sub] = nIdx[sub];
        cBeg[sub] = nBeg[sub];
        cCnt[sub] = nCnt[sub];
        request(sub);
      }
      if (et == 0) SB_TP_STAMP(t, 7);  // epilogue: quad done
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync_all();  // nobody leaves (or frees TMEM) while the partner can still signal or read
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// =============================================================================================
// dense-block kernel, TMA form (K6t): the kernel that consumes the BSMR dense blocks (replaces
// src/sddmmKernel.cu:213-351, whose gather is the scalar __ldg loop at :279-306) with both operands staged by
// the TMA unit.  Operands come from TF32-rounded copies (k_round_dense_rows: tcgen05 kind::tf32 truncates, the
// reference rounds) that hold ONLY the rows dense blocks touch:
//   * the 16 A rows of the panel: one ordinary box load (32 floats x 16 rows, SWIZZLE_128B);
//   * the 128 gathered B^T rows: 32 x cp.async.bulk.tensor.2d ... tile::gather4 -- four row indices (compact ids
//     of denseCols) plus the K coordinate per instruction, one instruction per lane of the producer warp, each
//     landing 4 x 128 B at its slot of the swizzled stage;
//   * full / empty mbarriers per stage, one MMA-issuing thread, accumulator 128 lanes x 16 TMEM columns, the same
//     blockValues epilogue as k_sddmm_dense.
// =============================================================================================
constexpr u32 kDtStages = 4;

__device__ __forceinline__ void tma_gather4_2d(u32 dstSmem, const void* map, u32 barSmem, u32 c0, u32 r0, u32 r1, u32 r2,
                                               u32 r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dstSmem), "l"(map), "r"(barSmem), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}

// rounds the rows the dense blocks touch into the compact copies: one warp per row
static __global__ void __launch_bounds__(256) k_round_dense_rows(u32 M, u32 K4, const float4* __restrict__ A,
                                                                 const float4* __restrict__ B, const u32* __restrict__ R,
                                                                 u32 nR, const u32* __restrict__ panelList, u32 numPanels,
                                                                 const u32* __restrict__ colList, u32 numCols,
                                                                 float4* __restrict__ Ar, float4* __restrict__ Br,
                                                                 BatchStrides bs) {
  A += (bs.a >> 2) * blockIdx.y;
  B += (bs.b >> 2) * blockIdx.y;
  Ar += (size_t)numPanels * 16u * K4 * blockIdx.y;
  Br += (size_t)numCols * K4 * blockIdx.y;
  const u32 lane = threadIdx.x & 31u;
  const u32 nA = numPanels * 16u;
  const size_t nw = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < (size_t)nA + numCols; w += nw) {
    const float4* src = nullptr;
    float4* dst;
    if (w < nA) {
      const u32 ri = panelList[w >> 4] * 16u + ((u32)w & 15u);
      if (ri < nR) {
        const u32 row = R[ri];
        if (row < M) src = A + (size_t)row * K4;
      }
      dst = Ar + w * K4;
    } else {
      src = B + (size_t)colList[w - nA] * K4;
      dst = Br + (w - nA) * K4;
    }
    for (u32 j = lane; j < K4; j += 32) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (src) v = __ldg(src + j);
      v.x = tf32_rna(v.x); v.y = tf32_rna(v.y); v.z = tf32_rna(v.z); v.w = tf32_rna(v.w);
      dst[j] = v;
    }
  }
}

constexpr int kDtThreads = 160;  // warps 0-3: TMA producers + epilogue (the 4 TMEM lane quarters), warp 4: MMA issuer

static __global__ void __launch_bounds__(kDtThreads)
k_sddmm_dense_tma(const __grid_constant__ CUtensorMap mapA16, const __grid_constant__ CUtensorMap mapBg, u32 K,
                  u32 numCompactCols, const u32* __restrict__ colCompact, const u32* __restrict__ blockOffsets,
                  const u32* __restrict__ blockValues, const uint2* __restrict__ work,
                  const u32* __restrict__ workRowA, float* __restrict__ P, size_t pStride) {
  extern __shared__ __align__(1024) unsigned char smemRaw[];
  P += pStride * blockIdx.y;
  unsigned char* stages = smemRaw + ((1024u - (smem_u32(smemRaw) & 1023u)) & 1023u);
  __shared__ u64 fullBar[kDtStages], emptyBar[kDtStages], accBar;
  __shared__ u32 tmemBase;

  const u32 tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  const uint2 w = work[blockIdx.x];
  const u32 p = w.x, firstBlk = w.y;
  const u32 blkBeg = blockOffsets[p], blkEnd = blockOffsets[p + 1];
  const u32 nBlk = min(kDenseGroupBlocks, blkEnd - blkBeg - firstBlk);
  const u32 colBase = (blkBeg + firstBlk) * 16u;  // denseColOffsets[p] == blockOffsets[p] * 16

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmemBase)),
                 "r"(kDnTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (u32 s = 0; s < kDtStages; ++s) {
      mbar_init(&fullBar[s], 4);  // one arrive.expect_tx per producer warp
      mbar_init(&emptyBar[s], 1);
    }
    mbar_init(&accBar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const u32 tmem = tmemBase;
  const u32 numChunks = (K + kDnKChunk - 1) / kDnKChunk;
  constexpr u32 idesc = umma_idesc_tf32(128, 16);

  if (warp < 4) {
    // ---- producers: warp q feeds gathered rows 32q .. 32q+31 of every stage, lanes 0..7 one tile::gather4 each
    // (4 row indices per instruction); warp 0 adds the panel's A box.  Slots past the group's last block and
    // sentinel columns read compact row 0: their accumulator lanes are never stored (blockValues holds NULL
    // there) and every TMEM lane depends on its own row only.
    u32 rows[4] = {0u, 0u, 0u, 0u};
    const u32 rowBase = blockIdx.y * numCompactCols;  // batch b's copy starts at row b * numCompactCols
    if (lane < 8u) {
#pragma unroll
      for (u32 i = 0; i < 4; ++i) {
        const u32 slot = warp * 32u + lane * 4u + i;
        rows[i] = rowBase + (slot < nBlk * 16u ? __ldg(colCompact + colBase + slot) : 0u);
      }
    }
    const u32 aRow = workRowA[blockIdx.x];
    const u32 myBytes = 32u * 128u + (warp == 0 ? 16u * 128u : 0u);
    for (u32 kc = 0; kc < numChunks; ++kc) {
      const u32 s = kc % kDtStages;
      const u32 bar = smem_u32(&fullBar[s]);
      const u32 dst = smem_u32(stages + s * kDnStageBytes);
      if (lane == 0) {
        if (kc >= kDtStages) mbar_wait_bounded(&emptyBar[s], ((kc / kDtStages) - 1) & 1u);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(myBytes) : "memory");
      }
      __syncwarp();
      if (lane < 8u)
        tma_gather4_2d(dst + (warp * 32u + lane * 4u) * 128u, &mapBg, bar, kc * kDnKChunk, rows[0], rows[1], rows[2], rows[3]);
      if (warp == 0 && lane == 0) tma_load_3d(dst + kDnRowsB * 128u, &mapA16, bar, kc * kDnKChunk, aRow, blockIdx.y);
    }
  } else {
    for (u32 kc = 0; kc < numChunks; ++kc) {
      const u32 s = kc % kDtStages;
      if (lane == 0) {
        mbar_wait_bounded(&fullBar[s], (kc / kDtStages) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const u32 base = smem_u32(stages + s * kDnStageBytes);
        const u64 dB = umma_desc_sw128(base);                   // 128 gathered B^T rows: MMA "A" operand
        const u64 dA = umma_desc_sw128(base + kDnRowsB * 128u);  // 16 panel rows: MMA "B" operand
#pragma unroll
        for (u32 k = 0; k < kDnKChunk / 8; ++k) {
          const u32 acc = (kc | k) ? 1u : 0u;
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
              "l"(dB + 2ull * k), "l"(dA + 2ull * k), "r"(idesc), "r"(acc)
              : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                         smem_u32(&emptyBar[s]))
                     : "memory");
        if (kc + 1 == numChunks)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                           smem_u32(&accBar))
                       : "memory");
      }
      __syncwarp();
    }
  }
  if (warp < 4) {
    mbar_wait_bounded(&accBar, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- epilogue: TMEM lane = gathered column slot j, register n = panel row
    u32 acc[16];
    const u32 taddr = tmem + ((warp * 32u) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]), "=r"(acc[4]), "=r"(acc[5]), "=r"(acc[6]), "=r"(acc[7]),
          "=r"(acc[8]), "=r"(acc[9]), "=r"(acc[10]), "=r"(acc[11]), "=r"(acc[12]), "=r"(acc[13]), "=r"(acc[14]),
          "=r"(acc[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (tid < nBlk * 16u) {
      const u32* bv = blockValues + ((size_t)(blkBeg + firstBlk) + (tid >> 4)) * 256u + (tid & 15u);
#pragma unroll
      for (u32 r = 0; r < 16; ++r) {
        const u32 idx = __ldg(bv + r * 16u);
        if (idx != kNull) P[idx] = __uint_as_float(acc[r]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kDnTmemCols) : "memory");
  }
}

// ---- K-independent index of the dense part (distinct columns, panels with dense blocks), built once per layout
static __global__ void k_mark_dense_cols(const u32* __restrict__ denseCols, size_t n, u32 N, u32* __restrict__ flag) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const u32 c = denseCols[i];
    if (c < N) flag[c] = 1u;
  }
}
static __global__ void k_mark_dense_panels(const u32* __restrict__ blockOffsets, u32 P, u32* __restrict__ flag) {
  for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x)
    flag[p] = blockOffsets[p + 1] > blockOffsets[p] ? 1u : 0u;
}
static __global__ void k_compact_ids(const u32* __restrict__ flag, const u32* __restrict__ ex, size_t n,
                                     u32* __restrict__ list) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    if (flag[i]) list[ex[i]] = (u32)i;
}
static __global__ void k_dense_col_compact(const u32* __restrict__ denseCols, size_t n, u32 N,
                                           const u32* __restrict__ ex, u32* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const u32 c = denseCols[i];
    out[i] = c < N ? ex[c] : 0u;
  }
}
static __global__ void k_dense_work_rows(const uint2* __restrict__ work, u32 n, const u32* __restrict__ exPanel,
                                         u32* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = exPanel[work[i].x] * 16u;
}

static u32 read_back_u32(const u32* d, cudaStream_t s) {
  u32 h = 0;
  SB_CUDA(cudaMemcpyAsync(&h, d, 4, cudaMemcpyDeviceToHost, s));
  SB_CUDA(cudaStreamSynchronize(s));
  return h;
}

static const bsmr_layout::DenseIndex* ensure_dense_index(const bsmr_layout* L, cudaStream_t s) {
  if (L->dix) return L->dix.get();
  auto d = std::make_unique<bsmr_layout::DenseIndex>();
  const bsmr_layout_info& I = L->info;
  const size_t nDc = L->arr[BSMR_DENSE_COLS].size();
  const u32 P = I.numRowPanels;
  if (nDc && P) {
    TempScope scope(s);
    DevBuf<u32> flagC((size_t)I.N + 1), exC((size_t)I.N + 1), flagP((size_t)P + 1), exP((size_t)P + 1);
    SB_CUDA(cudaMemsetAsync(flagC.get(), 0, ((size_t)I.N + 1) * 4, s));
    SB_CUDA(cudaMemsetAsync(flagP.get(), 0, ((size_t)P + 1) * 4, s));
    k_mark_dense_cols<<<grid_for(nDc), 256, 0, s>>>(L->arr[BSMR_DENSE_COLS].get(), nDc, I.N, flagC.get());
    SB_LAUNCH_CHECK();
    k_mark_dense_panels<<<grid_for(P), 256, 0, s>>>(L->arr[RPHM_BLOCK_OFFSETS].get(), P, flagP.get());
    SB_LAUNCH_CHECK();
    exclusive_scan_u32(flagC.get(), exC.get(), (size_t)I.N + 1, s);
    exclusive_scan_u32(flagP.get(), exP.get(), (size_t)P + 1, s);
    d->numCols = read_back_u32(exC.get() + I.N, s);
    d->numPanels = read_back_u32(exP.get() + P, s);
    d->colList.alloc(d->numCols ? d->numCols : 1, true);
    d->panelList.alloc(d->numPanels ? d->numPanels : 1, true);
    d->colCompact.alloc(nDc, true);
    d->workRowA.alloc(L->numDenseWork ? L->numDenseWork : 1, true);
    k_compact_ids<<<grid_for(I.N), 256, 0, s>>>(flagC.get(), exC.get(), I.N, d->colList.get());
    SB_LAUNCH_CHECK();
    k_compact_ids<<<grid_for(P), 256, 0, s>>>(flagP.get(), exP.get(), P, d->panelList.get());
    SB_LAUNCH_CHECK();
    k_dense_col_compact<<<grid_for(nDc), 256, 0, s>>>(L->arr[BSMR_DENSE_COLS].get(), nDc, I.N, exC.get(), d->colCompact.get());
    SB_LAUNCH_CHECK();
    if (L->numDenseWork) {
      k_dense_work_rows<<<grid_for(L->numDenseWork), 256, 0, s>>>(L->denseWork.get(), L->numDenseWork, exP.get(),
                                                                 d->workRowA.get());
      SB_LAUNCH_CHECK();
    }
    SB_CUDA(cudaStreamSynchronize(s));
  }
  L->dix = std::move(d);
  return L->dix.get();
}

// host side of the TMA kernels: workspaces + tensor maps (cached in the layout per K / batch count)
static void encode_map(void* out, const float* base, int rank, u32 K, u64 rows, u32 numBatch, u32 boxRows,
                       bool half = false) {
  using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeFn>(p);
  }();
  if (!fn) fail(SDDMM_E_UNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t esz = half ? 2u : 4u;
  const cuuint64_t dims[3] = {K, rows, numBatch};
  const cuuint64_t strides[2] = {(cuuint64_t)K * esz, (cuuint64_t)rows * K * esz};
  const cuuint32_t box[3] = {half ? 64u : kDnKChunk, boxRows, 1u};  // one 128-byte swizzle row of K
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  const CUresult rc = fn(reinterpret_cast<CUtensorMap*>(out),
                         half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                         const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) fail(SDDMM_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)rc);
}

static const bsmr_layout::TileTma* ensure_tile_tma(const bsmr_layout* L, u32 K, u32 numBatch, bool half = false) {
  const u64 key = ((u64)K << 32) | numBatch | (half ? 0x80000000ull : 0ull);
  {
    auto it = L->tma.find(key);
    if (it != L->tma.end()) return it->second.get();
  }
  auto t = std::make_unique<bsmr_layout::TileTma>();
  const bsmr_layout_info& I = L->info;
  const u32 nR = I.numRows ? I.numRows : 1u;
  t->K = K;
  t->numBatch = numBatch;
  // fp16 copies take half the floats' room (K is a multiple of 4, so K/2 floats per row hold K halves)
  t->rA.alloc((size_t)numBatch * nR * K / (half ? 2 : 1), true);  // outlives any scratch scope
  t->rB.alloc((size_t)numBatch * I.N * K / (half ? 2 : 1), true);
  encode_map(t->mapA, t->rA.get(), 3, K, nR, numBatch, 128u, half);
  encode_map(t->mapB, t->rB.get(), 3, K, I.N, numBatch, 128u, half);
  encode_map(t->mapA64, t->rA.get(), 3, K, nR, numBatch, 64u, half);
  encode_map(t->mapB64, t->rB.get(), 3, K, I.N, numBatch, 64u, half);
  SB_CUDA(cudaEventCreateWithFlags(&t->busy, cudaEventDisableTiming));
  return (L->tma[key] = std::move(t)).get();
}

static const bsmr_layout::HalfB* ensure_half_b(const bsmr_layout* L, u32 K, u32 numBatch) {
  const u64 key = ((u64)K << 32) | numBatch;
  {
    auto it = L->halfB.find(key);
    if (it != L->halfB.end()) return it->second.get();
  }
  auto t = std::make_unique<bsmr_layout::HalfB>();
  t->K = K;
  t->numBatch = numBatch;
  t->rB.alloc((size_t)numBatch * L->info.N * K / 2, true);  // K halves per row = K/2 floats' room
  SB_CUDA(cudaEventCreateWithFlags(&t->busy, cudaEventDisableTiming));
  return (L->halfB[key] = std::move(t)).get();
}
// whether SDDMM_OPERANDS_FP16 also reads B^T from an fp16 copy in the super-panel kernel (K >= 64)
static bool sp_half_b(u32 K) {
  const int cfg = [] { const char* e = getenv("SDDMM_B200_SP_HALF_B"); return e ? atoi(e) : 1; }();
  return cfg != 0 && K >= 64u;
}

static const bsmr_layout::DenseTma* ensure_dense_tma(const bsmr_layout* L, u32 K, u32 numBatch, cudaStream_t s) {
  const u64 key = ((u64)K << 32) | numBatch;
  {
    auto it = L->dtma.find(key);
    if (it != L->dtma.end()) return it->second.get();
  }
  const bsmr_layout::DenseIndex* d = ensure_dense_index(L, s);
  auto t = std::make_unique<bsmr_layout::DenseTma>();
  t->K = K;
  t->numBatch = numBatch;
  const u64 rowsA = (u64)(d->numPanels ? d->numPanels : 1u) * 16u, rowsB = d->numCols ? d->numCols : 1u;
  t->rA.alloc((size_t)numBatch * rowsA * K, true);
  t->rB.alloc((size_t)numBatch * rowsB * K, true);
  encode_map(t->mapA16, t->rA.get(), 3, K, rowsA, numBatch, 16u);
  encode_map(t->mapBg, t->rB.get(), 2, K, rowsB * numBatch, 1u, 1u);
  SB_CUDA(cudaEventCreateWithFlags(&t->busy, cudaEventDisableTiming));
  return (L->dtma[key] = std::move(t)).get();
}

// =============================================================================================
// plan resolution + launcher
// =============================================================================================
// L2 eviction policy of the super-panel kernel: wanted when B (N x K floats) cannot live in the 126 MB L2 next to A
// and the layout.  hub_budget = how many B^T rows may then be pinned (evict_last): 60 % of the L2.
static int l2_hint_cfg() {
  const char* e = getenv("SDDMM_B200_L2_HINTS");
  return e ? atoi(e) : -1;
}
static bool l2_hints_wanted(const bsmr_layout* L, u32 K) { return (size_t)L->info.N * K * 4 > ((size_t)48 << 20); }
static u32 hub_budget(const bsmr_layout* L, u32 K, bool halfB = false) {
  // read per call (getenv is cheap next to a launch), so that tests can switch the modes inside one process
  const int cfg = l2_hint_cfg();
  const int hubCfg = [] { const char* e = getenv("SDDMM_B200_L2_HUB_MB"); return e ? atoi(e) : -1; }();
  if (cfg == 0 || (cfg < 0 && !l2_hints_wanted(L, K))) return 0;
  if (hubCfg == 0) return 0;
  int dev = 0, l2 = 0;
  SB_CUDA(cudaGetDevice(&dev));
  SB_CUDA(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev));
  const size_t bytes = hubCfg > 0 ? (size_t)hubCfg << 20 : (size_t)l2 / 10 * 6;
  return (u32)std::min<size_t>(bytes / ((size_t)K * (halfB ? 2 : 4)), 0x7FFFFFFFu);
}
static u32 env_choice(const char* name, std::initializer_list<std::pair<const char*, u32>> table) {
  const char* e = getenv(name);
  if (!e) return 0u;
  for (const auto& kv : table)
    if (!strcmp(e, kv.first)) return kv.second;
  return 0u;
}

// The super-panel layout a pass uses and whether it runs in gather mode.  Gather mode (several B^T rows in flight per
// 8 lanes, no column-run reuse) when a (super-panel, column) run averages fewer than 2.5 entries -- graphs -- and the
// rows are short: K <= 64, or any K with fp16 B^T rows.  Gather mode has nothing to gain from tall super-panels (their
// only purpose is run reuse) and runs on half-height ones: 96 KB A tiles, shorter tile-load phases, more CTAs in
// flight (R-MAT scale 22, 192 -> 96 KB: K=64 2.06 -> 1.90 ms, fp16 K=64 1.67 -> 1.49 ms, fp16 K=256 4.92 -> 4.47 ms;
// reuse mode at K=256 loses: 7.40 -> 7.57 ms and keeps the full height).
struct SpChoice {
  const SuperPanelLayout* sp;
  bool gather;
};
static SpChoice choose_superpanels(const bsmr_layout* L, u32 K, bool halfA, bool halfB, cudaStream_t s) {
  const int gatherCfg = [] { const char* e = getenv("SDDMM_B200_SP_GATHER"); return e ? atoi(e) : -1; }();
  const u32 G = superpanel_G(K, halfA), hub = hub_budget(L, K, halfB);
  const SuperPanelLayout* sp = ensure_superpanels(L, G, hub, s);
  const bool gather = gatherCfg >= 0 ? gatherCfg != 0
                                     : ((halfB || K <= 64u) && (double)sp->numEntries < 2.5 * (double)sp->numRuns);
  if (gather && G >= 2u && !getenv("SDDMM_B200_SP_SMEM_KB")) sp = ensure_superpanels(L, G / 2u, hub, s);
  return {sp, gather};
}

// whether AUTO picks the CTA-pair tile kernel (K10) over the one-tile-per-CTA TMA kernel (K9)
static bool pair_default() {
  static const bool v = [] { const char* e = getenv("SDDMM_B200_TILE_PAIR"); return e ? atoi(e) != 0 : kPairDefault; }();
  return v;
}

void plan_default(sddmm_plan* out) {
  std::memset(out, 0, sizeof *out);
  out->plan = env_choice("SDDMM_B200_PLAN", {{"bsmr", SDDMM_PLAN_BSMR}, {"full", SDDMM_PLAN_TILE}, {"tile", SDDMM_PLAN_TILE}});
  out->dense = env_choice("SDDMM_B200_DENSE", {{"reg", SDDMM_DENSE_REG}, {"tma", SDDMM_DENSE_TMA}});
  out->residual = env_choice("SDDMM_B200_RESIDUAL", {{"0", SDDMM_RESIDUAL_PANEL}, {"panel", SDDMM_RESIDUAL_PANEL},
                                                     {"1", SDDMM_RESIDUAL_SUPERPANEL}, {"sp", SDDMM_RESIDUAL_SUPERPANEL},
                                                     {"2", SDDMM_RESIDUAL_STREAM}, {"stream", SDDMM_RESIDUAL_STREAM}});
  out->tile = env_choice("SDDMM_B200_TILE", {{"reg", SDDMM_TILE_REG}, {"tma1", SDDMM_TILE_TMA}, {"tma", SDDMM_TILE_TMA},
                                             {"tma4", SDDMM_TILE_TMA_CLUSTER}, {"pair", SDDMM_TILE_TMA_PAIR}});
  out->operands = env_choice("SDDMM_B200_OPERANDS", {{"fp16", SDDMM_OPERANDS_FP16}});
  if (const char* e = getenv("SDDMM_B200_TILE_STAGES")) { const int v = atoi(e); if (v >= 2 && v <= 4) out->tileStages = (u32)v; }
}

// Resolves every AUTO.  Cost model (cycles per SM, calibrated on B200): a tile moves 2*128*K*4 bytes from L2
// (~54 B/clk/SM) and its stores cost ~1 clk each; the BSMR kernels cost about 0.03*K + 1.5 clk per non-zero.
void plan_resolve(const bsmr_layout* L, u32 K, u32 numBatch, const sddmm_plan* in, sddmm_plan* out) {
  if (numBatch > 65535u) fail(SDDMM_E_ARG, "numBatch=%u exceeds the grid's y extent", numBatch);
  if (K == 0 || (K & 3u)) fail(SDDMM_E_ARG, "K=%u must be a positive multiple of 4", K);
  sddmm_plan p;
  if (in) p = *in; else plan_default(&p);
  if (p.plan > SDDMM_PLAN_TILE || p.dense > SDDMM_DENSE_TMA || p.residual > SDDMM_RESIDUAL_STREAM ||
      p.tile > SDDMM_TILE_TMA_PAIR || (p.tileStages && (p.tileStages < 2 || p.tileStages > 4)) ||
      p.operands > SDDMM_OPERANDS_FP16)
    fail(SDDMM_E_ARG, "sddmm_plan holds an unknown selector");
  const bsmr_layout_info& I = L->info;
  const bool haveTiles = L->tl && L->tl->numTiles;
  if (p.plan == SDDMM_PLAN_TILE && !haveTiles) {
    if (L->tl || (u64)I.numDenseValues + I.numSparseValues == 0) p.plan = SDDMM_PLAN_BSMR;  // nothing stored: nothing to launch
    else fail(SDDMM_E_UNSUPPORTED, "SDDMM_PLAN_TILE: this layout was built without the full-tile layout (BSMR_BUILD_TILES_*)");
  }
  if (p.plan == SDDMM_PLAN_AUTO) {
    p.plan = SDDMM_PLAN_BSMR;
    if (haveTiles) {
      const double tileCost = (double)L->tl->numTiles * (1024.0 * K / 54.0 + 600.0) + (double)L->tl->numEntries;
      const double covered = (double)I.numDenseValues + (double)I.numSparseValues;  // this shard's entries
      if (tileCost < covered * (0.03 * K + 1.5)) p.plan = SDDMM_PLAN_TILE;
    }
  }
  if (p.plan == SDDMM_PLAN_TILE) {
    // REG: register-staged tiles (K < 64: the rounding pre-pass of the TMA forms costs more than it saves there);
    // TMA: one CTA per tile fed by TMA; TMA_PAIR: persistent CTA pairs (AUTO for K >= 64); TMA_CLUSTER: 2x2 clusters with multicast (opt-in: measured slower)
    if (p.operands == SDDMM_OPERANDS_FP16) {
      if (p.tile == SDDMM_TILE_REG || p.tile == SDDMM_TILE_TMA_CLUSTER)
        fail(SDDMM_E_UNSUPPORTED, "SDDMM_OPERANDS_FP16 is implemented by the TMA tile kernels only (tile = SDDMM_TILE_TMA / _TMA_PAIR)");
      if (p.tile == SDDMM_TILE_AUTO) p.tile = pair_default() ? SDDMM_TILE_TMA_PAIR : SDDMM_TILE_TMA;
    }
    if (p.tile == SDDMM_TILE_AUTO) p.tile = K >= 64 ? (pair_default() ? SDDMM_TILE_TMA_PAIR : SDDMM_TILE_TMA) : SDDMM_TILE_REG;
    if ((p.tile == SDDMM_TILE_TMA_CLUSTER || p.tile == SDDMM_TILE_TMA_PAIR) && !L->tl->numQuads) p.tile = SDDMM_TILE_TMA;
    if (p.tile == SDDMM_TILE_TMA_PAIR) p.tileStages = kTpStages;  // fixed ring depth
    if (!p.tileStages) p.tileStages = 2;
    p.dense = SDDMM_DENSE_AUTO;
    p.residual = SDDMM_RESIDUAL_AUTO;
  } else {
    p.tile = SDDMM_TILE_AUTO;
    p.tileStages = 0;
    if (!L->numDenseWork) p.dense = SDDMM_DENSE_AUTO;
    else if (p.dense == SDDMM_DENSE_AUTO) p.dense = SDDMM_DENSE_REG;
    if (!L->numSparseWork) p.residual = SDDMM_RESIDUAL_AUTO;
    else {
      const u32 G = superpanel_G(K, p.operands == SDDMM_OPERANDS_FP16);
      if ((p.residual == SDDMM_RESIDUAL_SUPERPANEL || p.residual == SDDMM_RESIDUAL_STREAM) && !G)
        fail(SDDMM_E_UNSUPPORTED, "SDDMM_RESIDUAL_%s needs K in {32, 64, 128, 256, 512} (K=%u)",
             p.residual == SDDMM_RESIDUAL_STREAM ? "STREAM" : "SUPERPANEL", K);
      if (p.residual == SDDMM_RESIDUAL_AUTO) {
        if (!G) p.residual = SDDMM_RESIDUAL_PANEL;
        else {
          // expected entries per (super-panel, column) run: below ~1.5 a fetched B^T row is not reused from
          // registers and the super-panel kernel only pays for its shared-memory staging -> row-stream kernel
          // the row-stream kernel stays opt-in: measured on R-MAT scale 22 it is slower than the super-panel kernel at
          // every K (2.35 / 4.0 / 6.4 / 9.1 / 15.4 ms vs 2.07 / 2.46 / 3.58 / 7.5 / 15.5 ms at K = 32 ... 512; its
          // per-row A fragment loads are exposed and it issues twice the instructions), and on uniform 100k^2
          // (3.86 vs 3.24 ms at K=128)
          p.residual = SDDMM_RESIDUAL_SUPERPANEL;
        }
      }
      if (p.residual == SDDMM_RESIDUAL_PANEL && (size_t)16 * K * sizeof(float) > 200 * 1024)
        fail(SDDMM_E_UNSUPPORTED, "K=%u too large for the residual kernel's A tile", K);
    }
  }
  *out = p;
}

// builds everything a run with this (resolved) plan needs; after it a run only enqueues
void plan_prepare(const bsmr_layout* L, u32 K, u32 numBatch, const sddmm_plan& p, cudaStream_t s) {
  if (p.plan == SDDMM_PLAN_TILE) {
    if (p.tile != SDDMM_TILE_REG) ensure_tile_tma(L, K, numBatch, p.operands == SDDMM_OPERANDS_FP16);
    return;
  }
  if (p.dense == SDDMM_DENSE_TMA) ensure_dense_tma(L, K, numBatch, s);
  if (p.residual == SDDMM_RESIDUAL_SUPERPANEL) {
    const bool halfA = p.operands == SDDMM_OPERANDS_FP16, halfB = halfA && sp_half_b(K);
    choose_superpanels(L, K, halfA, halfB, s);
    if (halfB) ensure_half_b(L, K, numBatch);
  }
  if (p.residual == SDDMM_RESIDUAL_STREAM) ensure_stream(L, s);
}

// The rounded-operand workspaces are shared by every pass of one (K, numBatch): passes on different streams are
// ordered through the workspace's event.  Inside a stream capture the event is left alone (a captured event
// cannot be waited on by later non-captured work); a graph's replays are ordered by their launch stream.
static bool is_capturing(cudaStream_t s) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  return cudaStreamIsCapturing(s, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone;
}
static void workspace_acquire(cudaEvent_t busy, cudaStream_t s) {
  if (!is_capturing(s)) SB_CUDA(cudaStreamWaitEvent(s, busy, 0));
}
static void workspace_release(cudaEvent_t busy, cudaStream_t s) {
  if (!is_capturing(s)) SB_CUDA(cudaEventRecord(busy, s));
}

template <typename Kern>
static void set_smem(Kern kern, size_t smem) {
  // per device and cheap: set on every launch path instead of caching a per-process flag
  SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
}

void sddmm_launch(const bsmr_layout* L, u32 K, const float* dA, const float* dB, float* dP, cudaStream_t denseStream,
                  cudaStream_t sparseStream, int which, u32 numBatch, const sddmm_plan* plan) {
  if (numBatch == 0) return;
  sddmm_plan p;
  plan_resolve(L, K, numBatch, plan, &p);
  const bsmr_layout_info& I = L->info;
  auto arr = [&](bsmr_array_id id) { return L->arr[id].get(); };
  BatchStrides bst;
  if (numBatch > 1) {
    bst.a = (size_t)I.M * K;
    bst.b = (size_t)I.N * K;
    bst.p = I.nnz;
  }
  const u32 K4 = K / 4;
  if (p.plan == SDDMM_PLAN_TILE) {
    if (!(which & kLaunchDense)) return;  // every stored entry lives in exactly one tile: the "dense" side does it all
    if (p.tile == SDDMM_TILE_REG) {
      const size_t smem = (size_t)kTlStages * kTlStageBytes + 1024;
      set_smem(k_sddmm_tile, smem);
      k_sddmm_tile<<<dim3(L->tl->numTiles, numBatch), kTlThreads, smem, denseStream>>>(
          I.M, I.N, K, dA, dB, arr(BSMR_REORDERED_ROWS), I.numRows, L->tl->tiles.get(), L->tl->rowMeta.get(),
          L->tl->idx.get(), dP, bst);
      SB_LAUNCH_CHECK();
      return;
    }
    const bool half = p.operands == SDDMM_OPERANDS_FP16;
    const bsmr_layout::TileTma* t = ensure_tile_tma(L, K, numBatch, half);
    const size_t smem = (size_t)p.tileStages * kTlStageBytes + 1024;
    workspace_acquire(t->busy, denseStream);  // an earlier pass on another stream may still read rA / rB
    const size_t work = ((size_t)I.numRows + I.N) * K4;
    const dim3 rgrid((unsigned)std::min<size_t>((work + 255) / 256, (size_t)device_sm_count() * 16), numBatch);
    if (half)
      k_round_operands_half<<<rgrid, 256, 0, denseStream>>>(
          I.M, I.N, K4, reinterpret_cast<const float4*>(dA), reinterpret_cast<const float4*>(dB),
          arr(BSMR_REORDERED_ROWS), I.numRows, reinterpret_cast<uint2*>(t->rA.get()),
          reinterpret_cast<uint2*>(t->rB.get()), bst);
    else
      k_round_operands<<<rgrid, 256, 0, denseStream>>>(
          I.M, I.N, K4, reinterpret_cast<const float4*>(dA), reinterpret_cast<const float4*>(dB),
          arr(BSMR_REORDERED_ROWS), I.numRows, reinterpret_cast<float4*>(t->rA.get()),
          reinterpret_cast<float4*>(t->rB.get()), bst);
    SB_LAUNCH_CHECK();
    if (p.tile == SDDMM_TILE_TMA_PAIR) {
      auto kp = half ? k_sddmm_tile_pair<true> : k_sddmm_tile_pair<false>;
      const size_t psmem = (size_t)kTpStages * kTlStageBytes + kTpStagingBytes + 1024;
      set_smem(kp, psmem);
      cudaLaunchConfig_t cfg{};
      cfg.blockDim = dim3(kTpThreads);
      cfg.dynamicSmemBytes = psmem;
      cfg.stream = denseStream;
      cudaLaunchAttribute at[2];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // prologue under the rounding pass's tail
      at[1].val.programmaticStreamSerializationAllowed = 1;
      static const bool pdl = [] { const char* e = getenv("SDDMM_B200_PAIR_PDL"); return !e || atoi(e) != 0; }();
      cfg.attrs = at;
      cfg.numAttrs = pdl ? 2 : 1;
      // persistent: as many CTA pairs as can be resident at once (a static round-robin schedule must not queue pairs)
      cfg.gridDim = dim3((unsigned)device_sm_count() & ~1u, 1);
      static std::mutex occMu;
      static std::map<std::pair<int, bool>, int> occ;  // (device, fp16 form) -> resident CTA pairs: the query costs
      int maxPairs = 0;                                  // tens of microseconds on the host, a pass about as much
      {
        int dev = 0;
        SB_CUDA(cudaGetDevice(&dev));
        std::lock_guard<std::mutex> lk(occMu);
        auto it = occ.find({dev, half});
        if (it == occ.end()) {
          SB_CUDA(cudaOccupancyMaxActiveClusters(&maxPairs, kp, &cfg));
          occ[{dev, half}] = maxPairs;
        } else {
          maxPairs = it->second;
        }
      }
      if (maxPairs < 1) fail(SDDMM_E_CUDA, "k_sddmm_tile_pair: no CTA pair fits on this device");
      u32 perBatch = std::max<u32>(1u, (u32)maxPairs / numBatch);
      if (const char* e = getenv("SDDMM_B200_PAIR_GRID")) { const int v = atoi(e); if (v >= 1) perBatch = std::min<u32>(perBatch, (u32)v); }
      cfg.gridDim = dim3(2u * std::min<u32>(L->tl->numQuads, perBatch), numBatch);
      const u32 dbgOn = [] { const char* e = getenv("SDDMM_B200_PAIR_DEBUG"); return e && atoi(e) ? 1u : 0u; }();
      SB_CUDA(cudaLaunchKernelEx(&cfg, kp, *reinterpret_cast<const CUtensorMap*>(t->mapA),
                                 *reinterpret_cast<const CUtensorMap*>(t->mapB), K, (u32)L->tl->numQuads,
                                 (const uint2*)L->tl->quads.get(), (const u32*)L->tl->quadTiles.get(),
                                 (const uint4*)L->tl->tiles.get(), (const uint2*)L->tl->rowMetaT.get(),
                                 (const u32*)L->tl->idx.get(), dP, bst.p, dbgOn));
      SB_LAUNCH_CHECK();
      if (dbgOn && !is_capturing(denseStream)) {
        unsigned long long h[2][8][8];
        SB_CUDA(cudaStreamSynchronize(denseStream));
        SB_CUDA(cudaMemcpyFromSymbol(h, g_tpDbg, sizeof h));
        for (int c = 0; c < 2; ++c)
          for (int t = 0; t < 8; ++t) {
            if (!h[c][t][7]) continue;
            fprintf(stderr, "[pair dbg] cta %d quad %d:", c, t);
            for (int e = 0; e < 8; ++e) fprintf(stderr, " %lld", (long long)(h[c][t][e] ? h[c][t][e] - h[c][0][3] : 0));
            fprintf(stderr, "\n");
          }
      }
    } else if (p.tile == SDDMM_TILE_TMA_CLUSTER) {
      auto kq = p.tileStages == 2 ? k_sddmm_tile_tma4<2> : p.tileStages == 3 ? k_sddmm_tile_tma4<3> : k_sddmm_tile_tma4<4>;
      set_smem(kq, smem);
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(L->tl->numQuads * 4u, numBatch);
      cfg.blockDim = dim3(kTlThreads);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = denseStream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      SB_CUDA(cudaLaunchKernelEx(&cfg, kq, *reinterpret_cast<const CUtensorMap*>(t->mapA64),
                                 *reinterpret_cast<const CUtensorMap*>(t->mapB64), K,
                                 (const uint2*)L->tl->quads.get(), (const u32*)L->tl->quadTiles.get(),
                                 (const uint4*)L->tl->tiles.get(), (const u32*)L->tl->rowMeta.get(),
                                 (const u32*)L->tl->idx.get(), dP, bst.p));
      SB_LAUNCH_CHECK();
    } else {
      auto kt = half ? (p.tileStages == 2 ? k_sddmm_tile_tma<2, true> : p.tileStages == 3 ? k_sddmm_tile_tma<3, true>
                                                                                      : k_sddmm_tile_tma<4, true>)
                     : (p.tileStages == 2 ? k_sddmm_tile_tma<2, false> : p.tileStages == 3 ? k_sddmm_tile_tma<3, false>
                                                                                       : k_sddmm_tile_tma<4, false>);
      set_smem(kt, smem);
      kt<<<dim3(L->tl->numTiles, numBatch), kTlThreads, smem, denseStream>>>(
          *reinterpret_cast<const CUtensorMap*>(t->mapA), *reinterpret_cast<const CUtensorMap*>(t->mapB), K,
          L->tl->tiles.get(), L->tl->rowMeta.get(), L->tl->idx.get(), dP, bst.p);
      SB_LAUNCH_CHECK();
    }
    workspace_release(t->busy, denseStream);
    return;
  }
  if (L->numDenseWork && (which & kLaunchDense)) {
    if (p.dense == SDDMM_DENSE_TMA) {
      const bsmr_layout::DenseTma* t = ensure_dense_tma(L, K, numBatch, denseStream);
      const bsmr_layout::DenseIndex* d = L->dix.get();
      workspace_acquire(t->busy, denseStream);
      const size_t rows = (size_t)d->numPanels * 16u + d->numCols;
      k_round_dense_rows<<<dim3((unsigned)std::min<size_t>((rows + 7) / 8, 148 * 16), numBatch), 256, 0, denseStream>>>(
          I.M, K4, reinterpret_cast<const float4*>(dA), reinterpret_cast<const float4*>(dB), arr(BSMR_REORDERED_ROWS),
          I.numRows, d->panelList.get(), d->numPanels, d->colList.get(), d->numCols,
          reinterpret_cast<float4*>(t->rA.get()), reinterpret_cast<float4*>(t->rB.get()), bst);
      SB_LAUNCH_CHECK();
      const size_t smem = (size_t)kDtStages * kDnStageBytes + 1024;
      set_smem(k_sddmm_dense_tma, smem);
      k_sddmm_dense_tma<<<dim3(L->numDenseWork, numBatch), kDtThreads, smem, denseStream>>>(
          *reinterpret_cast<const CUtensorMap*>(t->mapA16), *reinterpret_cast<const CUtensorMap*>(t->mapBg), K,
          d->numCols ? d->numCols : 1u, d->colCompact.get(), arr(RPHM_BLOCK_OFFSETS), arr(RPHM_BLOCK_VALUES),
          L->denseWork.get(), d->workRowA.get(), dP, bst.p);
      SB_LAUNCH_CHECK();
      workspace_release(t->busy, denseStream);
    } else {
      const size_t smem = (size_t)kDnStages * kDnStageBytes + 1024;
      set_smem(k_sddmm_dense, smem);
      k_sddmm_dense<<<dim3(L->numDenseWork, numBatch), kDnThreads, smem, denseStream>>>(
          I.M, I.N, K, dA, dB, arr(BSMR_REORDERED_ROWS), I.numRows, arr(BSMR_DENSE_COLS), arr(RPHM_BLOCK_OFFSETS),
          arr(RPHM_BLOCK_VALUES), L->denseWork.get(), dP, bst);
      SB_LAUNCH_CHECK();
    }
  }
  if (L->numSparseWork && (which & kLaunchSparse)) {
    if (p.residual == SDDMM_RESIDUAL_STREAM) {
      const StreamLayout* st = ensure_stream(L, sparseStream);
      if (st->numEntries) {
        const u32 numBatches = (st->numEntries + 31u) / 32u;
        auto launch = [&](auto kern) {
          // one resident wave: every warp walks ONE contiguous range of 32-entry batches (its rows' A fragments and
          // the metadata stream stay local), so the grid is what fits on the device at once
          int perSm = 0;
          SB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kern, 256, 0));
          const u32 grid = std::min<u32>((numBatches + 7u) / 8u, (u32)device_sm_count() * (u32)std::max(perSm, 1));
          kern<<<dim3(grid, numBatch), 256, 0, sparseStream>>>(reinterpret_cast<const float4*>(dA),
                                                              reinterpret_cast<const float4*>(dB), st->row.get(),
                                                              st->col.get(), st->idx.get(), st->numEntries, dP, bst);
        };
        switch (K / 32u) {
          case 1: launch(k_sddmm_residual_stream<8, 1>); break;
          case 2: launch(k_sddmm_residual_stream<16, 1>); break;
          case 4: launch(k_sddmm_residual_stream<32, 1>); break;
          case 8: launch(k_sddmm_residual_stream<32, 2>); break;
          default: launch(k_sddmm_residual_stream<32, 4>); break;
        }
        SB_LAUNCH_CHECK();
      }
    } else if (p.residual == SDDMM_RESIDUAL_SUPERPANEL) {
      const bool halfA = p.operands == SDDMM_OPERANDS_FP16, halfB = halfA && sp_half_b(K);
      const SpChoice spc = choose_superpanels(L, K, halfA, halfB, sparseStream);
      const SuperPanelLayout* sp = spc.sp;
      if (sp->numWork) {
        const size_t smem = (size_t)sp->rows * K * (halfA ? 2 : 4);
        const float* bSrc = dB;
        BatchStrides kst = bst;
        const bsmr_layout::HalfB* hb = nullptr;
        if (halfB) {  // fp16 copy of B^T, rewritten every pass (B may change between calls)
          hb = ensure_half_b(L, K, numBatch);
          workspace_acquire(hb->busy, sparseStream);
          const size_t work = (size_t)sp->numUsedCols * K4;
          const dim3 rgrid((unsigned)std::min<size_t>((work + 255) / 256, (size_t)device_sm_count() * 16), numBatch);
          k_round_rows_half<<<rgrid, 256, 0, sparseStream>>>(
              sp->usedCols.get(), sp->numUsedCols, K4, reinterpret_cast<const float4*>(dB),
              reinterpret_cast<uint2*>(hb->rB.get()), numBatch > 1 ? (size_t)I.N * K4 : 0, (size_t)I.N * K4);
          SB_LAUNCH_CHECK();
          bSrc = hb->rB.get();
          if (numBatch > 1) kst.b = (size_t)I.N * K / 2;  // batch stride of the copy, in floats
        }
        auto launch = [&](auto kern, int threads) {
          set_smem(kern, smem);
          kern<<<dim3(sp->numWork, numBatch), threads, smem, sparseStream>>>(
              I.M, reinterpret_cast<const float4*>(dA), reinterpret_cast<const float4*>(bSrc), arr(BSMR_REORDERED_ROWS),
              I.numRows, sp->rows, sp->segLen, sp->off.get(), sp->col.get(), sp->row.get(), sp->idx.get(),
              sp->work.get(), dP, kst, sp->colMask);
        };
        // eviction hints when B (N x K floats) cannot live in the 126 MB L2 next to A and the layout; gather mode
        // (several B^T rows in flight, no column-run reuse) when a run averages fewer than 1.5 entries
        const int hintCfg = l2_hint_cfg();
        const bool hints = hintCfg >= 0 ? hintCfg != 0 : l2_hints_wanted(L, K);
        // measured on R-MAT scale 22: gather mode 2.07 -> 1.17 ms at K=32 and 2.46 -> 2.03 ms at K=64, but 3.58 -> 4.00
        // ms at K=128 and worse above (there one row per 8 lanes already keeps 64 KB per SM in flight and the L2
        // fabric, ~8.5 TB/s of gathered rows, is the limit), hence K <= 64.  With fp16 B^T rows a row is half as long
        // and gather mode wins at every K (K = 64 / 256 / 512: 2.65 -> 1.78, 6.89 -> 5.41, 10.9 -> 10.6 ms).
        // Threshold: 2.5 entries per run -- the REORDERED graph averages a little over 1.5 (its hub runs are long, most
        // runs are still singletons) and gathers faster too (K=32: 2.01 -> 1.38 ms, K=64: 2.43 -> 2.13 ms, fp16 K=256:
        // 6.33 -> 4.94 ms); the uniform 1 % matrix (3.8 entries per run, 7.7 with the fp16 tile) keeps reuse mode.
        // A third form, two independent streams per group with a B^T register set each (two rows in flight AND run
        // reuse), measured 7.44 vs 7.27 ms at K=256: at that row length bandwidth, not the round trip, is the limit.
        const bool gather = spc.gather;  // choose_superpanels
#define SB_SP_CASE2(NBv, THRv, Uv, HALFv)                                                         \
  do {                                                                                            \
    if (gather) { if (hints) launch(k_sddmm_residual_sp<NBv, THRv, true, Uv, HALFv>, THRv);       \
                  else launch(k_sddmm_residual_sp<NBv, THRv, false, Uv, HALFv>, THRv); }          \
    else { if (hints) launch(k_sddmm_residual_sp<NBv, THRv, true, 0, HALFv>, THRv);               \
           else launch(k_sddmm_residual_sp<NBv, THRv, false, 0, HALFv>, THRv); }                  \
  } while (0)
#define SB_SP_CASE3(NBv, THRv, Uv)                                                                         \
  do {                                                                                                     \
    if (gather) { if (hints) launch(k_sddmm_residual_sp<NBv, THRv, true, Uv, true, true>, THRv);           \
                  else launch(k_sddmm_residual_sp<NBv, THRv, false, Uv, true, true>, THRv); }              \
    else { if (hints) launch(k_sddmm_residual_sp<NBv, THRv, true, 0, true, true>, THRv);                   \
           else launch(k_sddmm_residual_sp<NBv, THRv, false, 0, true, true>, THRv); }                      \
  } while (0)
#define SB_SP_CASE(NBv, THRv, Uv)                    \
  do {                                               \
    if (halfA) SB_SP_CASE2(NBv, THRv, Uv, true);     \
    else SB_SP_CASE2(NBv, THRv, Uv, false);          \
  } while (0)
        if (halfB) {
          switch (K / 32u) {
            case 2: SB_SP_CASE3(2, 1024, 4); break;
            case 4: SB_SP_CASE3(4, 1024, 2); break;
            case 8: SB_SP_CASE3(8, 512, 2); break;
            default: SB_SP_CASE3(16, 512, 1); break;
          }
        } else
        switch (K / 32u) {
          case 1: SB_SP_CASE(1, 1024, 8); break;
          case 2: SB_SP_CASE(2, 1024, 4); break;
          case 4: SB_SP_CASE(4, 1024, 2); break;
          case 8: SB_SP_CASE(8, 512, 2); break;
          default: SB_SP_CASE(16, 512, 1); break;
        }
#undef SB_SP_CASE
#undef SB_SP_CASE2
#undef SB_SP_CASE3
        SB_LAUNCH_CHECK();
        if (hb) workspace_release(hb->busy, sparseStream);
      }
    } else {
      const size_t smem = (size_t)16 * K * sizeof(float);
      static const bool useL1 = [] { const char* e = getenv("SDDMM_B200_L1"); return !e || atoi(e) != 0; }();
      auto kern = useL1 ? k_sddmm_residual<true> : k_sddmm_residual<false>;
      if (smem > 48 * 1024) set_smem(kern, smem);
      kern<<<dim3(L->numSparseWork, numBatch), kResThreads, smem, sparseStream>>>(
          I.M, K4, reinterpret_cast<const float4*>(dA), reinterpret_cast<const float4*>(dB), arr(BSMR_REORDERED_ROWS),
          I.numRows, arr(BSMR_SPARSE_VALUE_OFFSETS), arr(RPHM_SPARSE_VALUES), arr(RPHM_SPARSE_RELATIVE_ROWS),
          arr(RPHM_SPARSE_COL_INDICES), L->sparseWork.get(), L->sparseChunk, dP, bst);
      SB_LAUNCH_CHECK();
    }
  }
}

}  // namespace sb
