/*
 * oracle/bsmr_oracle.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * See bsmr_oracle.h for the contract and the parity status (PINNED).
 *
 * Plain C restatement of the reference algorithm.  It is deliberately literal
 * (dense histograms, O(#clusters * M * nbpr) clustering) so that it is easy to
 * audit against the reference; it is only meant for sizes that finish in
 * seconds.  Compile with -ffp-contract=off: every fp32 operation must round
 * exactly like the reference's (IEEE RN add/div/sqrt, no FMA).
 */
#include "bsmr_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint32_t u32;

/* ------------------------------------------------------------------------- */
/* rowReordering.cu:1009-1025                                                */
u32 oracle_block_size(u32 M, u32 N, uint64_t free_mem) {
  /* (size_t)M*M*sizeof(UIN) / static_cast<float>(freeMem/2): size_t -> float */
  const float gmem_num = (float)((size_t)M * (size_t)M * sizeof(u32));
  const float gmem_den = (float)(free_mem / 2);
  const u32 min_gmem = (u32)ceilf(gmem_num / gmem_den);
  const float smem_num = (float)((size_t)N * sizeof(u32));
  const float smem_den = (float)(49152u / 2u); /* TensorCoreConfig.cuh:14 */
  const u32 min_smem = (u32)ceilf(smem_num / smem_den);
  const u32 bs = min_gmem > min_smem ? min_gmem : min_smem;
  return bs > 16 ? bs : 16;
}

/* rowReordering.cu:1035 */
u32 oracle_num_blocks_per_row(u32 N, u32 block_size) {
  return (u32)(int)ceilf((float)N / (float)block_size);
}

/* rowReordering.cu:911-920 */
u32 oracle_cluster_blockdim(u32 nbpr) {
  if (nbpr < 32) return 32;
  const int num_scan_iterate = 4;
  int cand = (int)(32 * ceil((float)((int)nbpr / num_scan_iterate) / (float)32));
  cand = cand > 32 ? cand : 32;
  return (u32)(1024 < cand ? 1024 : cand);
}

/* cudaUtil.cuh:37-43: halving tree over per-warp partials, stride = B/64. */
void oracle_kept_warps(u32 blockdim, uint8_t* kept) {
  const u32 W = blockdim / 32;
  /* track, per warp slot, the SET of warps summed into it, as a bitmask */
  uint32_t set[32];
  for (u32 w = 0; w < W; ++w) set[w] = 1u << w;
  for (u32 stride = blockdim / 64; stride >= 1; stride >>= 1)
    for (u32 w = 0; w < stride; ++w) set[w] |= set[w + stride];
  for (u32 w = 0; w < W; ++w) kept[w] = (uint8_t)((set[0] >> w) & 1u);
}

/* ------------------------------------------------------------------------- */
/* cudaUtil.cuh:13-45 reduce_sum as seen by thread 0 (the only value used).   */
static float reduce_f32(float* v /* B, clobbered */, u32 B) {
  const u32 W = B / 32;
  float s[32];
  for (u32 w = 0; w < W; ++w) {
    float* l = v + 32 * w;
    /* xor butterfly 1,2,4,8,16 == balanced pairwise tree for lane 0 */
    for (u32 step = 1; step < 32; step <<= 1)
      for (u32 i = 0; i < 32; i += 2 * step) l[i] = l[i] + l[i + step];
    s[w] = l[0];
  }
  for (u32 stride = B / 64; stride >= 1; stride >>= 1)
    for (u32 w = 0; w < stride; ++w) s[w] = s[w] + s[w + stride];
  return s[0];
}
static u32 reduce_u32(const u32* v, u32 B) {
  const u32 W = B / 32;
  u32 s[32];
  for (u32 w = 0; w < W; ++w) {
    u32 acc = 0;
    for (u32 i = 0; i < 32; ++i) acc += v[32 * w + i];
    s[w] = acc;
  }
  for (u32 stride = B / 64; stride >= 1; stride >>= 1)
    for (u32 w = 0; w < stride; ++w) s[w] += s[w + stride];
  return s[0];
}

/* rowReordering.cu:235-293 */
float oracle_similarity(const u32* rep, const u32* cmp, u32 nbpr, u32 B) {
  u32 ur[1024], uc[1024];
  float fmin_[1024], fmax_[1024];
  memset(ur, 0, sizeof(u32) * B);
  memset(uc, 0, sizeof(u32) * B);
  for (u32 i = 0; i < nbpr; ++i) { /* thread t = i % B accumulates i ascending */
    const u32 t = i % B;
    ur[t] += rep[i] * rep[i]; /* int*int wraps on the device: same bits */
    uc[t] += cmp[i] * cmp[i];
  }
  const u32 ss_rep = reduce_u32(ur, B);
  const u32 ss_cmp = reduce_u32(uc, B);
  if (ss_rep == 0 && ss_cmp == 0) return 1.0f;
  if (ss_rep == 0 || ss_cmp == 0) return 0.0f;
  const float norm_rep = sqrtf((float)ss_rep);
  const float norm_cmp = sqrtf((float)ss_cmp);
  for (u32 t = 0; t < B; ++t) fmin_[t] = 0.0f, fmax_[t] = 0.0f;
  for (u32 i = 0; i < nbpr; ++i) {
    const u32 t = i % B;
    const float a = ((float)rep[i]) / norm_rep;
    const float b = ((float)cmp[i]) / norm_cmp;
    fmin_[t] = fmin_[t] + fminf(a, b);
    fmax_[t] = fmax_[t] + fmaxf(a, b);
  }
  const float min_sum = reduce_f32(fmin_, B);
  const float max_sum = reduce_f32(fmax_, B);
  return min_sum / max_sum;
}

/* ------------------------------------------------------------------------- */
/* rowReordering.cu:49-93                                                     */
void oracle_encode_dispersion(const u32* rowOff, const u32* colIdx, u32 M,
                              u32 N, u32 block_size, u32* enc, u32* disp) {
  const u32 nbpr = oracle_num_blocks_per_row(N, block_size);
  u32* h = (u32*)calloc(nbpr ? nbpr : 1, sizeof(u32));
  for (u32 r = 0; r < M; ++r) {
    const u32 nz = rowOff[r + 1] - rowOff[r];
    if (enc) memset(enc + (size_t)r * nbpr, 0, sizeof(u32) * nbpr);
    if (nz == 0) { /* early return at :66-68; buffers were memset to 0 */
      disp[r] = 0;
      continue;
    }
    memset(h, 0, sizeof(u32) * nbpr);
    for (u32 i = rowOff[r]; i < rowOff[r + 1]; ++i) h[colIdx[i] / block_size]++;
    u32 tmp = 0, nb = 0;
    for (u32 b = 0; b < nbpr; ++b) {
      if (h[b]) {
        nb++;
        tmp += block_size - h[b];
      }
    }
    /* reduce_sum over 128 threads (4 warps: nothing dropped), uint32 wrap */
    disp[r] = tmp + nz * nb;
    if (enc) memcpy(enc + (size_t)r * nbpr, h, sizeof(u32) * nbpr);
  }
  free(h);
}

/* stable LSD radix sort of (key,val) pairs == thrust::host stable sort_by_key
 * (parallelAlgorithm.cu:43-48; thrust/system/detail/sequential/sort.inl) */
static void stable_sort_by_key(u32* key, u32* val, size_t n) {
  u32* k2 = (u32*)malloc(sizeof(u32) * (n ? n : 1));
  u32* v2 = (u32*)malloc(sizeof(u32) * (n ? n : 1));
  for (int pass = 0; pass < 4; ++pass) {
    size_t cnt[257];
    memset(cnt, 0, sizeof(cnt));
    const int sh = pass * 8;
    for (size_t i = 0; i < n; ++i) cnt[((key[i] >> sh) & 255u) + 1]++;
    for (int d = 0; d < 256; ++d) cnt[d + 1] += cnt[d];
    for (size_t i = 0; i < n; ++i) {
      const size_t p = cnt[(key[i] >> sh) & 255u]++;
      k2[p] = key[i];
      v2[p] = val[i];
    }
    memcpy(key, k2, sizeof(u32) * n);
    memcpy(val, v2, sizeof(u32) * n);
  }
  free(k2);
  free(v2);
}

/* rowReordering.cu:1027-1095, :893-1007, :325-432 */
int oracle_row_reorder(const u32* rowOff, const u32* colIdx, u32 M, u32 N,
                       float alpha, u32 block_size, u32* reorderedRows,
                       u32* numRows, int32_t* clusterCnt, u32* clusterOfRow,
                       u32* ascending_out) {
  const u32 nbpr = oracle_num_blocks_per_row(N, block_size);
  const u32 B = oracle_cluster_blockdim(nbpr);
  u32* enc = (u32*)malloc(sizeof(u32) * (size_t)M * nbpr + 4);
  u32* disp = (u32*)malloc(sizeof(u32) * M);
  if (!enc || !disp) return -1;
  oracle_encode_dispersion(rowOff, colIdx, M, N, block_size, enc, disp);

  /* :1060-1062 stable sort rows by dispersion ascending */
  u32* keys = (u32*)malloc(sizeof(u32) * M);
  u32* asc = (u32*)malloc(sizeof(u32) * M);
  for (u32 i = 0; i < M; ++i) keys[i] = disp[i], asc[i] = i;
  stable_sort_by_key(keys, asc, M);

  /* :939-949 leading zero-dispersion rows form cluster 0 */
  u32* cid = (u32*)malloc(sizeof(u32) * M);
  for (u32 i = 0; i < M; ++i) cid[i] = ORACLE_NULL_VALUE;
  u32 zero_row_idx = 0;
  while (zero_row_idx < M && disp[asc[zero_row_idx]] == 0) cid[zero_row_idx++] = 0;

  /* :325-432: equivalent sequential semantics of the mutex-chained kernels */
  u32* rep = (u32*)malloc(sizeof(u32) * (nbpr ? nbpr : 1));
  u32 cluster = 1, start = zero_row_idx;
  while (start < M) {
    cid[start] = cluster;
    memcpy(rep, enc + (size_t)asc[start] * nbpr, sizeof(u32) * nbpr);
    u32 next_start = ORACLE_NULL_VALUE;
    for (u32 idx = start + 1; idx < M; ++idx) {
      if (cid[idx] != ORACLE_NULL_VALUE) continue;
      const u32* cmp = enc + (size_t)asc[idx] * nbpr;
      const float sim = oracle_similarity(rep, cmp, nbpr, B);
      if (sim > alpha) {
        cid[idx] = cluster;
        for (u32 i = 0; i < nbpr; ++i) rep[i] += cmp[i];
      } else if (next_start == ORACLE_NULL_VALUE) {
        next_start = idx;
      }
    }
    if (next_start == ORACLE_NULL_VALUE) break;
    start = next_start;
    cluster++;
  }

  /* :986-995 stable sort positions by cluster id, compose with ascending */
  u32* skey = (u32*)malloc(sizeof(u32) * M);
  u32* ind = (u32*)malloc(sizeof(u32) * M);
  for (u32 i = 0; i < M; ++i) skey[i] = cid[i], ind[i] = i;
  stable_sort_by_key(skey, ind, M);
  u32* perm = (u32*)malloc(sizeof(u32) * M);
  for (u32 i = 0; i < M; ++i) perm[i] = asc[ind[i]];
  /* :996 (quirk: indexes the SORTED id array by a position of the unsorted one) */
  if (clusterCnt) *clusterCnt = M ? (int32_t)(skey[ind[M - 1]] + (u32)(zero_row_idx != 0)) : 0;
  if (clusterOfRow)
    for (u32 i = 0; i < M; ++i) clusterOfRow[asc[i]] = cid[i];
  if (ascending_out) memcpy(ascending_out, asc, sizeof(u32) * M);

  /* :1081-1090 strip leading empty rows */
  u32 s = 0;
  while (s < M && rowOff[perm[s] + 1] - rowOff[perm[s]] == 0) ++s;
  *numRows = M - s;
  memcpy(reorderedRows, perm + s, sizeof(u32) * (M - s));

  free(perm); free(ind); free(skey); free(rep); free(cid);
  free(asc); free(keys); free(disp); free(enc);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* The same clustering with CANDIDATE PRUNING (validation of the round-2 device
 * design; also what lets the oracle reach matrices whose dense M x nbpr encoding
 * does not fit).  For alpha >= 0 a row that shares no non-zero block with the
 * representative has min_sum == 0, hence sim == 0, and can never join; the only
 * other way to join is the "both norms are zero" rule (:263-268).  So for every
 * cluster only the rows on the posting lists of the representative's non-zero
 * blocks (lists of the blocks a joiner adds are merged in from its position on)
 * and the rows with a zero norm are evaluated -- in ascending position, with the
 * literal reduction tree restricted to the blocks that are non-zero on either
 * side (adding +0.0f to a partial sum does not change it, and every thread still
 * sees its blocks in ascending order).  Must return exactly what
 * oracle_row_reorder returns; tests/test_oracle_cpu.py checks that.            */
typedef struct { u32 blk, cnt; } ent_t;

static float similarity_sparse(const u32* rep, const u32* repNz, u32 nRepNz, const ent_t* cmp, u32 nCmp,
                               u32 B, u32* ur, u32* uc, float* fmn, float* fmx, u32* mark /* nbpr, zero */) {
  /* integer norms: per-thread partials over the non-zero blocks (order irrelevant, wrapping adds) */
  memset(ur, 0, sizeof(u32) * B);
  memset(uc, 0, sizeof(u32) * B);
  for (u32 k = 0; k < nRepNz; ++k) { const u32 i = repNz[k]; ur[i % B] += rep[i] * rep[i]; }
  for (u32 k = 0; k < nCmp; ++k) uc[cmp[k].blk % B] += cmp[k].cnt * cmp[k].cnt;
  const u32 ss_rep = reduce_u32(ur, B), ss_cmp = reduce_u32(uc, B);
  if (ss_rep == 0 && ss_cmp == 0) return 1.0f;
  if (ss_rep == 0 || ss_cmp == 0) return 0.0f;
  const float norm_rep = sqrtf((float)ss_rep), norm_cmp = sqrtf((float)ss_cmp);
  for (u32 t = 0; t < B; ++t) fmn[t] = 0.0f, fmx[t] = 0.0f;
  /* union of the two sorted block lists, ascending */
  for (u32 k = 0; k < nCmp; ++k) mark[cmp[k].blk] = cmp[k].cnt;
  u32 a = 0, b = 0;
  while (a < nRepNz || b < nCmp) {
    u32 i;
    if (b >= nCmp || (a < nRepNz && repNz[a] <= cmp[b].blk)) { i = repNz[a]; if (b < nCmp && cmp[b].blk == i) ++b; ++a; }
    else i = cmp[b++].blk;
    const u32 t = i % B;
    const float x = ((float)rep[i]) / norm_rep, y = ((float)mark[i]) / norm_cmp;
    fmn[t] = fmn[t] + fminf(x, y);
    fmx[t] = fmx[t] + fmaxf(x, y);
  }
  for (u32 k = 0; k < nCmp; ++k) mark[cmp[k].blk] = 0;
  const float min_sum = reduce_f32(fmn, B), max_sum = reduce_f32(fmx, B);
  return min_sum / max_sum;
}

static int cmp_u32_asc(const void* x, const void* y) {
  const u32 a = *(const u32*)x, b = *(const u32*)y;
  return a < b ? -1 : a > b;
}

int oracle_row_reorder_pruned(const u32* rowOff, const u32* colIdx, u32 M, u32 N, float alpha, u32 block_size,
                              u32* reorderedRows, u32* numRows, int32_t* clusterCnt, u32* clusterOfRow,
                              u32* ascending_out, uint64_t* evaluations, int flags) {
  if (!(alpha >= 0.0f)) return -2; /* the pruning argument needs alpha >= 0 */
  const int prefixFilter = flags & 1;
  const u32 nbpr = oracle_num_blocks_per_row(N, block_size);
  const u32 B = oracle_cluster_blockdim(nbpr);
  u32* disp = (u32*)malloc(sizeof(u32) * (M ? M : 1));
  oracle_encode_dispersion(rowOff, colIdx, M, N, block_size, NULL, disp);
  u32* keys = (u32*)malloc(sizeof(u32) * (M ? M : 1));
  u32* asc = (u32*)malloc(sizeof(u32) * (M ? M : 1));
  for (u32 i = 0; i < M; ++i) keys[i] = disp[i], asc[i] = i;
  stable_sort_by_key(keys, asc, M);
  u32* cid = (u32*)malloc(sizeof(u32) * (M ? M : 1));
  for (u32 i = 0; i < M; ++i) cid[i] = ORACLE_NULL_VALUE;
  u32 zero_row_idx = 0;
  while (zero_row_idx < M && disp[asc[zero_row_idx]] == 0) cid[zero_row_idx++] = 0;

  /* sparse encodings by POSITION (blocks ascending), posting lists by block (positions ascending) */
  size_t* eOff = (size_t*)malloc(sizeof(size_t) * ((size_t)M + 1));
  u32* h = (u32*)calloc(nbpr ? nbpr : 1, sizeof(u32));
  u32* tb = (u32*)malloc(sizeof(u32) * (nbpr ? nbpr : 1));
  size_t total = 0;
  for (u32 p = 0; p < M; ++p) { /* count distinct blocks per row */
    const u32 r = asc[p];
    u32 nt = 0;
    for (u32 i = rowOff[r]; i < rowOff[r + 1]; ++i) { const u32 b = colIdx[i] / block_size; if (h[b]++ == 0) tb[nt++] = b; }
    for (u32 k = 0; k < nt; ++k) h[tb[k]] = 0;
    eOff[p] = total;
    total += nt;
  }
  eOff[M] = total;
  ent_t* ent = (ent_t*)malloc(sizeof(ent_t) * (total ? total : 1));
  size_t* pOff = (size_t*)calloc((size_t)nbpr + 1, sizeof(size_t));
  for (u32 p = 0; p < M; ++p) {
    const u32 r = asc[p];
    u32 nt = 0;
    for (u32 i = rowOff[r]; i < rowOff[r + 1]; ++i) { const u32 b = colIdx[i] / block_size; if (h[b]++ == 0) tb[nt++] = b; }
    qsort(tb, nt, sizeof(u32), cmp_u32_asc);
    for (u32 k = 0; k < nt; ++k) { ent[eOff[p] + k].blk = tb[k]; ent[eOff[p] + k].cnt = h[tb[k]]; pOff[tb[k] + 1]++; h[tb[k]] = 0; }
  }
  for (u32 b = 0; b < nbpr; ++b) pOff[b + 1] += pOff[b];
  u32* post = (u32*)malloc(sizeof(u32) * (total ? total : 1));
  size_t* fill = (size_t*)malloc(sizeof(size_t) * ((size_t)nbpr + 1));
  memcpy(fill, pOff, sizeof(size_t) * ((size_t)nbpr + 1));
  for (u32 p = 0; p < M; ++p)
    for (size_t k = eOff[p]; k < eOff[p + 1]; ++k) post[fill[ent[k].blk]++] = p; /* ascending p */

  u32 ur[1024], uc[1024];
  float fmn[1024], fmx[1024];
  u32* mark = (u32*)calloc(nbpr ? nbpr : 1, sizeof(u32));
  /* rows whose integer norm reduces to zero (all their blocks in dropped warps, or a wrap) */
  u32* zeroNorm = (u32*)malloc(sizeof(u32) * (M ? M : 1));
  u32 nZeroNorm = 0;
  for (u32 p = zero_row_idx; p < M; ++p) {
    memset(uc, 0, sizeof(u32) * B);
    for (size_t k = eOff[p]; k < eOff[p + 1]; ++k) uc[ent[k].blk % B] += ent[k].cnt * ent[k].cnt;
    if (reduce_u32(uc, B) == 0) zeroNorm[nZeroNorm++] = p;
  }

  /* PREFIX FILTER (flags & 1).  sim = min_sum / max_sum <= (rep mass on the shared blocks) / (rep mass), both over
   * the blocks the reduction keeps, so a row that shares only blocks of a set F with mass(F) <= alpha * mass can
   * never join.  F = the most frequent (longest posting list) blocks of the representative while that holds, with a
   * 1e-3 relative margin for the fp32 tree sums; blocks in dropped warps carry no mass and are always in F.
   * Candidates then come from the posting lists of the OTHER blocks only.                                       */
  uint8_t keptWarp[32];
  oracle_kept_warps(B, keptWarp);
  uint8_t* hasIt = (uint8_t*)calloc(nbpr ? nbpr : 1, 1); /* block already feeds the candidate stream */
  u32* byFreq = (u32*)malloc(sizeof(u32) * (nbpr ? nbpr : 1));
  u32* rep = (u32*)calloc(nbpr ? nbpr : 1, sizeof(u32));
  u32* repNz = (u32*)malloc(sizeof(u32) * (nbpr ? nbpr : 1));
  /* candidate iterators: one per non-zero block of the representative (+ the zero-norm list); a plain array
   * scanned for its minimum is enough for a CPU checker */
  size_t* itCur = (size_t*)malloc(sizeof(size_t) * ((size_t)nbpr + 2));
  size_t* itEnd = (size_t*)malloc(sizeof(size_t) * ((size_t)nbpr + 2));
  const u32** itArr = (const u32**)malloc(sizeof(u32*) * ((size_t)nbpr + 2));
  uint64_t evals = 0;
  u32 cluster = 1, start = zero_row_idx;
  while (start < M) {
    cid[start] = cluster;
    u32 nRepNz = 0, nIt = 0;
    for (size_t k = eOff[start]; k < eOff[start + 1]; ++k) {
      const u32 b = ent[k].blk;
      rep[b] = ent[k].cnt;
      repNz[nRepNz++] = b;
    }
    itArr[nIt] = zeroNorm; itCur[nIt] = 0; itEnd[nIt] = nZeroNorm; ++nIt;
    u32 last = start; /* candidates must lie strictly behind this position */
#define ADD_SOURCES()                                                                                         \
    do {                                                                                                        \
      if (!prefixFilter) {                                                                                      \
        for (u32 k_ = 0; k_ < nRepNz; ++k_) {                                                                   \
          const u32 b_ = repNz[k_];                                                                             \
          if (!hasIt[b_]) { hasIt[b_] = 1; itArr[nIt] = post; itCur[nIt] = pOff[b_]; itEnd[nIt] = pOff[b_ + 1]; ++nIt; } \
        }                                                                                                       \
      } else {                                                                                                  \
        uint64_t mass_ = 0;                                                                                     \
        u32 nk_ = 0;                                                                                            \
        for (u32 k_ = 0; k_ < nRepNz; ++k_) {                                                                   \
          const u32 b_ = repNz[k_];                                                                             \
          if (keptWarp[(b_ % B) / 32]) { mass_ += rep[b_]; byFreq[nk_++] = b_; }                                \
        }                                                                                                       \
        /* most frequent first (insertion sort: the lists are short) */                                         \
        for (u32 i_ = 1; i_ < nk_; ++i_) {                                                                      \
          const u32 v_ = byFreq[i_];                                                                            \
          const size_t lv_ = pOff[v_ + 1] - pOff[v_];                                                           \
          u32 j_ = i_;                                                                                          \
          while (j_ > 0 && (pOff[byFreq[j_ - 1] + 1] - pOff[byFreq[j_ - 1]]) < lv_) { byFreq[j_] = byFreq[j_ - 1]; --j_; } \
          byFreq[j_] = v_;                                                                                      \
        }                                                                                                       \
        const double budget_ = ((double)alpha * (1.0 - 1e-3) - 1e-6) * (double)mass_;                           \
        uint64_t acc_ = 0;                                                                                      \
        u32 skip_ = 0;                                                                                          \
        while (skip_ < nk_ && (double)(acc_ + rep[byFreq[skip_]]) <= budget_) { acc_ += rep[byFreq[skip_]]; ++skip_; } \
        for (u32 k_ = skip_; k_ < nk_; ++k_) {                                                                  \
          const u32 b_ = byFreq[k_];                                                                            \
          if (!hasIt[b_]) { hasIt[b_] = 1; itArr[nIt] = post; itCur[nIt] = pOff[b_]; itEnd[nIt] = pOff[b_ + 1]; ++nIt; } \
        }                                                                                                       \
      }                                                                                                         \
    } while (0)
    ADD_SOURCES();
    for (;;) {
      u32 best = ORACLE_NULL_VALUE;
      for (u32 t = 0; t < nIt; ++t) {
        while (itCur[t] < itEnd[t] && itArr[t][itCur[t]] <= last) ++itCur[t];
        if (itCur[t] < itEnd[t] && itArr[t][itCur[t]] < best) best = itArr[t][itCur[t]];
      }
      if (best == ORACLE_NULL_VALUE) break;
      last = best;
      if (cid[best] != ORACLE_NULL_VALUE) continue;
      ++evals;
      const float sim = similarity_sparse(rep, repNz, nRepNz, ent + eOff[best], (u32)(eOff[best + 1] - eOff[best]), B,
                                          ur, uc, fmn, fmx, mark);
      if (sim > alpha) {
        cid[best] = cluster;
        for (size_t k = eOff[best]; k < eOff[best + 1]; ++k) {
          const u32 b = ent[k].blk;
          if (rep[b] == 0) { /* a new block of the representative */
            u32 at = nRepNz++;
            while (at > 0 && repNz[at - 1] > b) { repNz[at] = repNz[at - 1]; --at; }
            repNz[at] = b;
          }
          rep[b] += ent[k].cnt;
        }
        ADD_SOURCES(); /* new blocks (and blocks that left F) feed the stream from `last` on */
      }
    }
    for (u32 k = 0; k < nRepNz; ++k) rep[repNz[k]] = 0, hasIt[repNz[k]] = 0;
#undef ADD_SOURCES
    u32 next_start = start + 1;
    while (next_start < M && cid[next_start] != ORACLE_NULL_VALUE) ++next_start;
    if (next_start >= M) break;
    start = next_start;
    cluster++;
  }
  if (evaluations) *evaluations = evals;

  u32* skey = (u32*)malloc(sizeof(u32) * (M ? M : 1));
  u32* ind = (u32*)malloc(sizeof(u32) * (M ? M : 1));
  for (u32 i = 0; i < M; ++i) skey[i] = cid[i], ind[i] = i;
  stable_sort_by_key(skey, ind, M);
  u32* perm = (u32*)malloc(sizeof(u32) * (M ? M : 1));
  for (u32 i = 0; i < M; ++i) perm[i] = asc[ind[i]];
  if (clusterCnt) *clusterCnt = M ? (int32_t)(skey[ind[M - 1]] + (u32)(zero_row_idx != 0)) : 0;
  if (clusterOfRow)
    for (u32 i = 0; i < M; ++i) clusterOfRow[asc[i]] = cid[i];
  if (ascending_out) memcpy(ascending_out, asc, sizeof(u32) * M);
  u32 s0 = 0;
  while (s0 < M && rowOff[perm[s0] + 1] - rowOff[perm[s0]] == 0) ++s0;
  *numRows = M - s0;
  memcpy(reorderedRows, perm + s0, sizeof(u32) * (M - s0));
  free(perm); free(ind); free(skey); free(itArr); free(itEnd); free(itCur); free(repNz); free(rep); free(zeroNorm);
  free(byFreq); free(hasIt);
  free(mark); free(fill); free(post); free(pOff); free(ent); free(tb); free(h); free(eOff); free(cid); free(asc);
  free(keys); free(disp);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* BSMR.cpp:48 (float ceil; exact for numRows <= 2^24, which bounds every case
 * the reference can run; integer form used beyond that) */
u32 oracle_num_panels(u32 numRows) { return (numRows + ORACLE_PANEL - 1) / ORACLE_PANEL; }

/* colReordering.cu:274-404 + :244-271.  Single pass that can be called twice:
 * offsets are always written; cols only when non-NULL. */
int oracle_col_reorder(const u32* rowOff, const u32* colIdx, u32 M, u32 N,
                       const u32* R, u32 nR, float delta, u32* dOff, u32* sOff,
                       u32* vOff, u32* denseCols, u32* sparseCols) {
  (void)M;
  const u32 P = oracle_num_panels(nR);
  /* :246  static_cast<UIN>(std::ceil(delta * BLOCK_SIZE)) in float */
  const u32 T = (u32)ceilf(delta * (float)(ORACLE_PANEL * ORACLE_BLOCK_COLS));
  u32* cnt = (u32*)calloc(N ? N : 1, sizeof(u32));
  u32 cap = 1024;
  u32* cols = (u32*)malloc(sizeof(u32) * cap);
  u32* keys = (u32*)malloc(sizeof(u32) * cap);
  u32* cols2 = (u32*)malloc(sizeof(u32) * cap);
  u32* touched = (u32*)malloc(sizeof(u32) * cap);
  dOff[0] = sOff[0] = vOff[0] = 0;
  for (u32 p = 0; p < P; ++p) {
    const u32 r0 = p * ORACLE_PANEL;
    const u32 r1 = (r0 + ORACLE_PANEL < nR) ? r0 + ORACLE_PANEL : nR;
    u32 nt = 0;
    for (u32 ri = r0; ri < r1; ++ri) {
      const u32 row = R[ri];
      for (u32 i = rowOff[row]; i < rowOff[row + 1]; ++i) {
        const u32 c = colIdx[i];
        if (cnt[c]++ == 0) {
          if (nt + 16 >= cap) {
            cap *= 2;
            cols = (u32*)realloc(cols, sizeof(u32) * cap);
            keys = (u32*)realloc(keys, sizeof(u32) * cap);
            cols2 = (u32*)realloc(cols2, sizeof(u32) * cap);
            touched = (u32*)realloc(touched, sizeof(u32) * cap);
          }
          touched[nt++] = c;
        }
      }
    }
    /* :314-331 non-empty columns in ascending column order */
    {
      /* sort touched ascending (radix via stable_sort_by_key on itself) */
      memcpy(cols, touched, sizeof(u32) * nt);
      memcpy(keys, touched, sizeof(u32) * nt);
      stable_sort_by_key(keys, cols, nt);
    }
    /* :333-336 stable sort by count descending (counts in 1..16): counting sort */
    u32 bucket[18];
    memset(bucket, 0, sizeof(bucket));
    for (u32 i = 0; i < nt; ++i) bucket[16 - cnt[cols[i]] + 1]++;
    for (int b = 0; b < 17; ++b) bucket[b + 1] += bucket[b];
    for (u32 i = 0; i < nt; ++i) cols2[bucket[16 - cnt[cols[i]]]++] = cols[i];
    /* :338-343 pad to a multiple of 16 with sentinel N / count 0 */
    const u32 padded = (nt + 15u) / 16u * 16u;
    /* :250-261 dense groups */
    u32 nd = 0, nnz_sparse = 0;
    for (u32 g = 0; g < padded; g += 16) {
      u32 sum = 0;
      for (u32 i = g; i < g + 16 && i < nt; ++i) sum += cnt[cols2[i]];
      if (sum >= T) nd += 16;
    }
    for (u32 i = nd; i < nt; ++i) nnz_sparse += cnt[cols2[i]];
    if (denseCols)
      for (u32 i = 0; i < nd; ++i) denseCols[dOff[p] + i] = i < nt ? cols2[i] : N;
    if (sparseCols)
      for (u32 i = nd; i < padded; ++i) sparseCols[sOff[p] + (i - nd)] = i < nt ? cols2[i] : N;
    dOff[p + 1] = dOff[p] + nd;
    sOff[p + 1] = sOff[p] + (padded - nd);
    vOff[p + 1] = vOff[p] + nnz_sparse;
    for (u32 i = 0; i < nt; ++i) cnt[touched[i]] = 0;
  }
  free(touched); free(cols2); free(keys); free(cols); free(cnt);
  return (int)P;
}

/* ------------------------------------------------------------------------- */
/* BSMR.cpp:83-265 */
int oracle_rphm_build(const u32* rowOff, const u32* colIdx, u32 M, u32 N,
                      const u32* R, u32 nR, const u32* dOff, const u32* denseCols,
                      const u32* sOff, const u32* sparseCols, const u32* vOff,
                      u32* blockOffsets, u32* blockValues, u32* sparseValues,
                      u32* sparseRelativeRows, u32* sparseColIndices) {
  (void)M;
  const u32 P = oracle_num_panels(nR);
  blockOffsets[0] = 0;
  for (u32 p = 0; p < P; ++p) /* :125-136 */
    blockOffsets[p + 1] = blockOffsets[p] + (dOff[p + 1] - dOff[p] + 15u) / 16u;
  const size_t nbv = (size_t)blockOffsets[P] * 256u;
  for (size_t i = 0; i < nbv; ++i) blockValues[i] = ORACLE_NULL_VALUE; /* :142 */
  /* col -> (csr index) per row, and per panel col -> list of (r, idx) */
  u32* where = (u32*)malloc(sizeof(u32) * ((size_t)N + 1));
  for (u32 c = 0; c <= N; ++c) where[c] = ORACLE_NULL_VALUE;
  /* :143-174 dense part */
  for (u32 ri = 0; ri < nR; ++ri) {
    const u32 row = R[ri], p = ri / 16, lr = ri % 16;
    for (u32 i = rowOff[row]; i < rowOff[row + 1]; ++i) where[colIdx[i]] = i;
    const size_t base = (size_t)blockOffsets[p] * 256u;
    for (u32 k = dOff[p], count = 0; k < dOff[p + 1]; ++k, ++count) {
      const u32 col = denseCols[k];
      if (col < N && where[col] != ORACLE_NULL_VALUE)
        blockValues[base + (size_t)(count / 16) * 256u + lr * 16u + (count % 16)] = where[col];
    }
    for (u32 i = rowOff[row]; i < rowOff[row + 1]; ++i) where[colIdx[i]] = ORACLE_NULL_VALUE;
  }
  /* :177-219 residual: for col in sparseCols(panel) order, rows in panel order */
  u32* slot = (u32*)malloc(sizeof(u32) * ((size_t)N + 1)); /* col -> sparse rank+1 */
  memset(slot, 0, sizeof(u32) * ((size_t)N + 1));
  for (u32 p = 0; p < P; ++p) {
    const u32 r0 = p * 16, r1 = (r0 + 16 < nR) ? r0 + 16 : nR;
    const u32 ns = sOff[p + 1] - sOff[p];
    /* count per sparse col, then prefix, then fill in (col order, row order) */
    u32* cstart = (u32*)calloc((size_t)ns + 1, sizeof(u32));
    for (u32 k = 0; k < ns; ++k) {
      const u32 c = sparseCols[sOff[p] + k];
      if (c < N) slot[c] = k + 1;
    }
    for (u32 ri = r0; ri < r1; ++ri) {
      const u32 row = R[ri];
      for (u32 i = rowOff[row]; i < rowOff[row + 1]; ++i)
        if (slot[colIdx[i]]) cstart[slot[colIdx[i]]]++;
    }
    for (u32 k = 0; k < ns; ++k) cstart[k + 1] += cstart[k];
    for (u32 ri = r0; ri < r1; ++ri) {
      const u32 row = R[ri];
      for (u32 i = rowOff[row]; i < rowOff[row + 1]; ++i) {
        const u32 s = slot[colIdx[i]];
        if (!s) continue;
        const u32 pos = vOff[p] + cstart[s - 1]++;
        sparseRelativeRows[pos] = ri % 16;
        sparseValues[pos] = i;
        sparseColIndices[pos] = colIdx[i];
      }
    }
    for (u32 k = 0; k < ns; ++k) {
      const u32 c = sparseCols[sOff[p] + k];
      if (c < N) slot[c] = 0;
    }
    free(cstart);
  }
  free(slot);
  free(where);
  return 0;
}

/* BSMR.cpp:99-119, :221-246 */
void oracle_work_lists(u32 P, const u32* dOff, const u32* vOff, u32* numDenseTB,
                       u32* maxDenseBlocks, u32* dIds, u32* dIters,
                       u32* numSparseTB, u32* maxSparseTB, u32* sIds, u32* sIters) {
  u32 nd = 0, md = 0, ns = 0, ms = 0;
  for (u32 p = 0; p < P; ++p) {
    const u32 nb = (dOff[p + 1] - dOff[p] + 15u) / 16u;
    if (nb > md) md = nb;
    const u32 tb = (nb + 3u) / 4u;
    for (u32 i = 0; i < tb; ++i) {
      if (dIds) dIds[nd + i] = p;
      if (dIters) dIters[nd + i] = dOff[p] / 16u + i * 4u;
    }
    nd += tb;
    const u32 nsd = vOff[p + 1] - vOff[p];
    const u32 stb = (nsd + 127u) / 128u;
    if (stb > ms) ms = stb;
    for (u32 i = 0; i < stb; ++i) {
      if (sIds) sIds[ns + i] = p;
      if (sIters) sIters[ns + i] = i * 128u;
    }
    ns += stb;
  }
  if (numDenseTB) *numDenseTB = nd;
  if (maxDenseBlocks) *maxDenseBlocks = md;
  if (numSparseTB) *numSparseTB = ns;
  if (maxSparseTB) *maxSparseTB = ms;
}

/* ------------------------------------------------------------------------- */
/* BSMR.cpp:953-994 calculateNumDenseBlocksAndAverageDensityInOriginalMatrix:
 * 16x16 blocks of the UNreordered matrix, panel-major then column-block
 * ascending (the order of the reference's two loops, which fixes the order of
 * its float accumulation); edge blocks use their clipped size.                */
static int cmp_u32(const void* a, const void* b) {
  const u32 x = *(const u32*)a, y = *(const u32*)b;
  return x < y ? -1 : x > y;
}
void oracle_original_block_stats(const u32* rowOff, const u32* colIdx, u32 M, u32 N, float delta,
                                 u32* numDenseBlocks, float* averageDensity) {
  const u32 nRP = (M + 15u) / 16u, nCB = (N + 15u) / 16u;
  u32* cnt = (u32*)calloc(nCB ? nCB : 1, sizeof(u32));
  u32* touched = (u32*)malloc(sizeof(u32) * (nCB ? nCB : 1));
  u32 nDense = 0;
  float total = 0.0f;
  for (u32 rp = 0; rp < nRP; ++rp) {
    const u32 r0 = rp * 16u, r1 = (r0 + 16u < M) ? r0 + 16u : M;
    u32 nt = 0;
    for (u32 r = r0; r < r1; ++r)
      for (u32 i = rowOff[r]; i < rowOff[r + 1]; ++i) {
        const u32 cb = colIdx[i] / 16u;
        if (cnt[cb]++ == 0) touched[nt++] = cb;
      }
    qsort(touched, nt, sizeof(u32), cmp_u32);
    for (u32 t = 0; t < nt; ++t) {
      const u32 cb = touched[t];
      const u32 c0 = cb * 16u, c1 = (c0 + 16u < N) ? c0 + 16u : N;
      const float blockSize = (float)((r1 - r0) * (c1 - c0));
      const float density = (float)cnt[cb] / blockSize;
      if (density >= delta) {
        total += density;
        ++nDense;
      }
      cnt[cb] = 0;
    }
  }
  free(cnt);
  free(touched);
  *numDenseBlocks = nDense;
  *averageDensity = nDense > 0 ? total / (float)nDense : 0.0f;
}

/* BSMR.cpp:826-925 evaluationReordering.  out6 = numDenseBlock, numDenseThreadBlocks,
 * numSparseThreadBlocks, numSparseData, numDenseData, (spare); *averageDensity as :917.
 * Only the DENSE column blocks get a column set in the reference (:860-871), so the
 * sparse blocks never count; numSparseData counts entries whose column is one of the
 * panel's sparse columns (:873-880, :899-901).                                  */
void oracle_evaluation_reordering(const u32* rowOff, const u32* colIdx, u32 M, u32 N, u32 nnz,
                                  const u32* R, u32 nR, const u32* dOff, const u32* dCols,
                                  const u32* sOff, const u32* sCols, const u32* vOff, float delta,
                                  u32* out6, float* averageDensity) {
  (void)M;
  const u32 P = (nR + 15u) / 16u;
  int* blockOfCol = (int*)malloc(sizeof(int) * ((size_t)N + 1));   /* dense block id of a column, -1 */
  unsigned char* isSparse = (unsigned char*)calloc((size_t)N + 1, 1);
  for (u32 c = 0; c <= N; ++c) blockOfCol[c] = -1;
  u32 numDenseBlocks = 0, numDenseTB = 0, numSparseTB = 0, numSparseData = 0;
  float totalDensity = 0.0f;
  for (u32 p = 0; p < P; ++p) {
    const u32 nDenseBlk = (u32)ceilf((float)(dOff[p + 1] - dOff[p]) / 16.0f);
    numDenseTB += (u32)ceilf((float)nDenseBlk / 4.0f);
    numSparseTB += (u32)ceilf((float)(vOff[p + 1] - vOff[p]) / 128.0f);
    u32* nnzIn = (u32*)calloc(nDenseBlk ? nDenseBlk : 1, sizeof(u32));
    for (u32 i = dOff[p]; i < dOff[p + 1]; ++i) blockOfCol[dCols[i]] = (int)((i - dOff[p]) / 16u);
    for (u32 i = sOff[p]; i < sOff[p + 1]; ++i) isSparse[sCols[i]] = 1;
    const u32 i0 = p * 16u, i1 = (i0 + 16u < nR) ? i0 + 16u : nR;
    for (u32 ir = i0; ir < i1; ++ir) {
      const u32 row = R[ir];
      for (u32 idx = rowOff[row]; idx < rowOff[row + 1]; ++idx) {
        const u32 col = colIdx[idx];
        if (blockOfCol[col] >= 0) ++nnzIn[blockOfCol[col]];
        if (isSparse[col]) ++numSparseData;
      }
    }
    for (u32 b = 0; b < nDenseBlk; ++b)
      if (nnzIn[b] > 0) {
        const float density = (float)nnzIn[b] / 256.0f;
        totalDensity += density;
        if (density >= delta) ++numDenseBlocks;
      }
    for (u32 i = dOff[p]; i < dOff[p + 1]; ++i) blockOfCol[dCols[i]] = -1;
    for (u32 i = sOff[p]; i < sOff[p + 1]; ++i) isSparse[sCols[i]] = 0;
    free(nnzIn);
  }
  free(blockOfCol);
  free(isSparse);
  out6[0] = numDenseBlocks;
  out6[1] = numDenseTB;
  out6[2] = numSparseTB;
  out6[3] = numSparseData;
  out6[4] = nnz - numSparseData;
  out6[5] = 0;
  /* :917  `totalDensity / numDenseBlocks > 0 ? totalDensity / numDenseBlocks : 0.0f` (0/0 -> NaN -> 0) */
  {
    const float q = totalDensity / (float)(int)numDenseBlocks;
    *averageDensity = q > 0 ? q : 0.0f;
  }
}

/* ------------------------------------------------------------------------- */
/* host.cpp:44-76: sequential-k fp32 dot per non-zero, OpenMP over rows.       */
void oracle_sddmm_cpu(const float* A, const float* B, const u32* rowOff,
                      const u32* colIdx, u32 M, u32 K, float* P, int threads) {
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#else
  (void)threads;
#endif
#pragma omp parallel for schedule(static)
  for (long row = 0; row < (long)M; ++row) {
    const float* a = A + (size_t)row * K;
    for (u32 i = rowOff[row]; i < rowOff[row + 1]; ++i) {
      const float* b = B + (size_t)colIdx[i] * K;
      float val = 0.0f;
      for (u32 k = 0; k < K; ++k) val += a[k] * b[k];
      P[i] = val;
    }
  }
}

/* checkData.hpp:14-30 */
size_t oracle_check_data(const float* a, const float* b, size_t n) {
  size_t errors = 0;
  for (size_t i = 0; i < n; ++i) {
    const float d = fabsf(a[i] - b[i]);
    if (d < 1e-5f) continue;
    float mx = fabsf(a[i]) > fabsf(b[i]) ? fabsf(a[i]) : fabsf(b[i]);
    if (mx < 1e-3f) mx = 1e-3f;
    if (!((d / mx) < 1e-3f)) ++errors;
  }
  return errors;
}

/* ------------------------------------------------------------------------- */
/* Matrix.cpp:398-480 + :374-396 + :236-250                                   */
static const char* next_word(const char* s, char* out, size_t cap) {
  while (*s == ' ' || *s == '\t' || *s == '\r') ++s;
  size_t n = 0;
  while (*s && *s != ' ' && *s != '\t' && *s != '\r' && *s != '\n') {
    if (n + 1 < cap) out[n++] = *s;
    ++s;
  }
  out[n] = 0;
  return s;
}

typedef struct { u32 r, c; } rc_t;
static int rc_cmp(const void* a, const void* b) {
  const rc_t* x = (const rc_t*)a; const rc_t* y = (const rc_t*)b;
  if (x->r != y->r) return x->r < y->r ? -1 : 1;
  if (x->c != y->c) return x->c < y->c ? -1 : 1;
  return 0;
}

int oracle_load_mtx(const char* path, u32* M, u32* N, u32* nnz, u32** rowOffOut,
                    u32** colIdxOut, float** valuesOut) {
  FILE* f = fopen(path, "r");
  if (!f) return -1;
  char* line = NULL; size_t lcap = 0; char w[64];
  ssize_t got;
  while ((got = getline(&line, &lcap, f)) >= 0 && line[0] == '%') {}
  if (got < 0) { fclose(f); free(line); return -2; }
  const char* s = next_word(line, w, sizeof w); *M = (u32)atoi(w);
  s = next_word(s, w, sizeof w); *N = (u32)atoi(w);
  s = next_word(s, w, sizeof w); *nnz = (u32)strtod(w[0] ? w : "0", NULL);
  const u32 n = *nnz;
  u32* ri = (u32*)malloc(sizeof(u32) * (n ? n : 1));
  u32* ci = (u32*)malloc(sizeof(u32) * (n ? n : 1));
  float* va = (float*)malloc(sizeof(float) * (n ? n : 1));
  u32 idx = 0;
  int rc = 0;
  while ((got = getline(&line, &lcap, f)) >= 0) {
    if (got == 0 || line[0] == '\n' || line[0] == 0) continue; /* empty line skipped */
    s = next_word(line, w, sizeof w); if (!w[0]) continue;
    const u32 r = (u32)atoi(w);
    s = next_word(s, w, sizeof w); const u32 c = (u32)atoi(w);
    s = next_word(s, w, sizeof w);
    const float v = w[0] ? (float)strtod(w, NULL) : 0.0f;
    if (idx >= n) { rc = -3; break; } /* too many elements */
    ri[idx] = r - 1; ci[idx] = c - 1; va[idx] = v; ++idx;
  }
  fclose(f); free(line);
  if (!rc && idx < n) rc = -4; /* not enough */
  if (!rc) {
    rc_t* set = (rc_t*)malloc(sizeof(rc_t) * (n ? n : 1));
    for (u32 i = 0; i < n && !rc; ++i) {
      if (ri[i] >= *M || ci[i] >= *N) rc = -5;
      set[i].r = ri[i]; set[i].c = ci[i];
    }
    if (!rc) {
      qsort(set, n, sizeof(rc_t), rc_cmp);
      for (u32 i = 1; i < n; ++i)
        if (set[i].r == set[i - 1].r && set[i].c == set[i - 1].c) { rc = -6; break; }
    }
    free(set);
  }
  if (!rc && n <= 1) rc = -7;
  if (rc) { free(ri); free(ci); free(va); return rc; }
  /* :467 stable sort by ROW ONLY: columns keep file order within a row */
  u32* ro = (u32*)calloc((size_t)*M + 1, sizeof(u32));
  for (u32 i = 0; i < n; ++i) ro[ri[i] + 1]++;
  for (u32 r = 0; r < *M; ++r) ro[r + 1] += ro[r];
  u32* pos = (u32*)malloc(sizeof(u32) * ((size_t)*M + 1));
  memcpy(pos, ro, sizeof(u32) * ((size_t)*M + 1));
  u32* co = (u32*)malloc(sizeof(u32) * n);
  float* vo = (float*)malloc(sizeof(float) * n);
  for (u32 i = 0; i < n; ++i) {
    const u32 p = pos[ri[i]]++;
    co[p] = ci[i]; vo[p] = va[i];
  }
  free(pos); free(ri); free(ci); free(va);
  *rowOffOut = ro; *colIdxOut = co; *valuesOut = vo;
  return 0;
}

void oracle_free(void* p) { free(p); }
