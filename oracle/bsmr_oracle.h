/*
 * oracle/bsmr_oracle.h  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the reference hot path
 *   reorder -> BSMR split -> RPHM layout -> SDDMM
 * of CX9898/sddmm-gpu.  Every function cites the reference file:line it follows.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (libsddmm_b200.so) never
 * links, loads or calls anything in here.
 *
 * Parity status: PINNED.  Column reordering, sddmm_cpu, checkData and the
 * Matrix-Market loader are pinned against the reference's own host code built
 * from /root/reference (oracle/_ref/libref_cpu.so, fixtures in tests/golden/);
 * the row reordering and RPHM arrays are pinned against the reference GPU
 * binary run on a B200 (oracle/_ref/ref_dump, fixtures in tests/golden/ref_gpu/).
 */
#ifndef BSMR_ORACLE_H
#define BSMR_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_NULL_VALUE 0xFFFFFFFFu /* include/TensorCoreConfig.cuh:11-12 */
#define ORACLE_PANEL 16u              /* include/BSMR.hpp:8-10 */
#define ORACLE_BLOCK_COLS 16u

/* rowReordering.cu:1009-1025 (free_mem = cudaMemGetInfo's free bytes). */
uint32_t oracle_block_size(uint32_t M, uint32_t N, uint64_t free_mem);

/* rowReordering.cu:1035  nbpr = ceil((float)N / (float)block_size). */
uint32_t oracle_num_blocks_per_row(uint32_t N, uint32_t block_size);

/* rowReordering.cu:911-920  clustering blockDim. */
uint32_t oracle_cluster_blockdim(uint32_t nbpr);

/* cudaUtil.cuh:27-45 emulation: which warps survive the halving tree.
 * kept[w] = 1 if warp w contributes to the value thread 0 sees. */
void oracle_kept_warps(uint32_t blockdim, uint8_t* kept /* blockdim/32 */);

/* rowReordering.cu:49-93: dense encodings (M x nbpr, may be NULL) + dispersion. */
void oracle_encode_dispersion(const uint32_t* rowOff, const uint32_t* colIdx,
                              uint32_t M, uint32_t N, uint32_t block_size,
                              uint32_t* enc /* M*nbpr or NULL */,
                              uint32_t* disp /* M */);

/* rowReordering.cu:235-293 + cudaUtil.cuh:13-45: similarity of two dense
 * histograms with the reference's exact reduction tree. */
float oracle_similarity(const uint32_t* rep, const uint32_t* cmp, uint32_t nbpr,
                        uint32_t blockdim);

/* rowReordering.cu:1027-1095 + :893-1007 + :325-432.
 * Outputs: reorderedRows (cap M, non-empty rows only), *numRows,
 * *clusterCnt (reference's quirky value, :996), clusterOfRow (optional, M,
 * indexed by ORIGINAL row id; 0 = empty row), ascending (optional, M,
 * dispersion order).  Returns 0 on success. */
int oracle_row_reorder(const uint32_t* rowOff, const uint32_t* colIdx,
                       uint32_t M, uint32_t N, float alpha, uint32_t block_size,
                       uint32_t* reorderedRows, uint32_t* numRows,
                       int32_t* clusterCnt, uint32_t* clusterOfRow,
                       uint32_t* ascending);

/* colReordering.cu:244-404.  Two-call protocol: pass NULL arrays to get sizes.
 * denseColOffsets / sparseColOffsets / sparseValueOffsets have numPanels+1
 * entries.  Returns number of row panels. */
uint32_t oracle_num_panels(uint32_t numRows);
int oracle_col_reorder(const uint32_t* rowOff, const uint32_t* colIdx,
                       uint32_t M, uint32_t N,
                       const uint32_t* reorderedRows, uint32_t numRows,
                       float delta,
                       uint32_t* denseColOffsets, uint32_t* sparseColOffsets,
                       uint32_t* sparseValueOffsets,
                       uint32_t* denseCols /* cap: denseColOffsets[P] */,
                       uint32_t* sparseCols /* cap: sparseColOffsets[P] */);

/* BSMR.cpp:83-265. blockOffsets has P+1 entries; blockValues has
 * blockOffsets[P]*256 entries; sparse arrays have sparseValueOffsets[P]. */
int oracle_rphm_build(const uint32_t* rowOff, const uint32_t* colIdx,
                      uint32_t M, uint32_t N,
                      const uint32_t* reorderedRows, uint32_t numRows,
                      const uint32_t* denseColOffsets, const uint32_t* denseCols,
                      const uint32_t* sparseColOffsets, const uint32_t* sparseCols,
                      const uint32_t* sparseValueOffsets,
                      uint32_t* blockOffsets, uint32_t* blockValues,
                      uint32_t* sparseValues, uint32_t* sparseRelativeRows,
                      uint32_t* sparseColIndices);

/* BSMR.cpp:99-119 and :221-246 work lists.  Pass NULL outputs to get counts. */
/* oracle_row_reorder with candidate pruning through an inverted index over the non-zero blocks (alpha >= 0 only,
 * returns -2 otherwise).  Same outputs; *evaluations receives the number of similarity evaluations made. */
int oracle_row_reorder_pruned(const uint32_t* rowOff, const uint32_t* colIdx, uint32_t M, uint32_t N, float alpha,
                              uint32_t block_size, uint32_t* reorderedRows, uint32_t* numRows, int32_t* clusterCnt,
                              uint32_t* clusterOfRow, uint32_t* ascending_out, uint64_t* evaluations,
                              int flags /* bit 0: prefix filter on the representative's block masses */);
/* BSMR.cpp:953-994 and :826-925 (the statistics the reference logs after every run) */
void oracle_original_block_stats(const uint32_t* rowOff, const uint32_t* colIdx, uint32_t M, uint32_t N,
                                 float delta, uint32_t* numDenseBlocks, float* averageDensity);
void oracle_evaluation_reordering(const uint32_t* rowOff, const uint32_t* colIdx, uint32_t M, uint32_t N,
                                  uint32_t nnz, const uint32_t* R, uint32_t nR, const uint32_t* dOff,
                                  const uint32_t* dCols, const uint32_t* sOff, const uint32_t* sCols,
                                  const uint32_t* vOff, float delta, uint32_t* out6, float* averageDensity);
void oracle_work_lists(uint32_t P, const uint32_t* denseColOffsets,
                       const uint32_t* sparseValueOffsets,
                       uint32_t* numDenseTB, uint32_t* maxDenseBlocks,
                       uint32_t* denseRowPanelIds, uint32_t* denseColBlockIters,
                       uint32_t* numSparseTB, uint32_t* maxSparseTB,
                       uint32_t* sparseRowPanelIds, uint32_t* sparseColBlockIters);

/* host.cpp:44-76. A row-major MxK, B column-major KxN (B[k + col*K]).
 * threads <= 0 -> omp default. */
void oracle_sddmm_cpu(const float* A, const float* B, const uint32_t* rowOff,
                      const uint32_t* colIdx, uint32_t M, uint32_t K, float* P,
                      int threads);

/* checkData.hpp:14-30. returns number of mismatching elements. */
size_t oracle_check_data(const float* a, const float* b, size_t n);

/* Matrix.cpp:398-480. Returns 0 ok, <0 error code mirroring the reference's
 * failure branches.  Outputs are malloc'ed; free with oracle_free. */
int oracle_load_mtx(const char* path, uint32_t* M, uint32_t* N, uint32_t* nnz,
                    uint32_t** rowOff, uint32_t** colIdx, float** values);
void oracle_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
