/*
 * include/sddmm_b200.h -- the drop-in boundary of the B200-native SDDMM engine.
 *
 * A plain C ABI (no C++/torch types) over libsddmm_b200.so.  Each entry point names the
 * interface of CX9898/sddmm-gpu (paths relative to the reference root) that it replaces; the
 * reference-side binding a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - uint32_t indices everywhere (the reference's UIN, include/TensorCoreConfig.cuh:10);
 *     SDDMM_NULL_VALUE = 0xFFFFFFFF (TensorCoreConfig.cuh:11-12); row panel = 16 rows, column
 *     block = 16 columns (include/BSMR.hpp:8-10).
 *   - A is row-major M x K; B is COLUMN-major K x N (N contiguous rows of K floats);
 *     P has one float per stored (row, col) of S, in the CSR order of colIdx
 *     (layout contract of include/Matrix.hpp + src/sddmmKernel.cu:2518-2537).
 *   - P[i] = sum_k A[row,k] * B[k,col].  The values of S are NOT read (src/host.cpp:62-73).
 *   - Pointers named d_* are device pointers on the current CUDA device, h_* are host pointers.
 *   - Every function returning int returns 0 on success and a non-zero SDDMM_E_* code on error;
 *     sddmm_last_error() returns a thread-local human-readable message.  (The reference returns
 *     void and prints; the C++ shim in sddmm-gpu_b200/csrc/host keeps those signatures.)
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - There is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     SDDMM_E_CUDA.
 *   - Threading: the library keeps its streams, scratch arena, launch counter and error string per
 *     host thread (streams additionally per device), so different threads may work on DIFFERENT
 *     layouts concurrently.  One bsmr_layout must not be used by two threads at once: it caches
 *     K-dependent private layouts and staging buffers inside the object (the reference is not
 *     re-entrant at all).  Passes on ONE layout may be enqueued on different streams;
 *     the tile-TMA plan's rounded-operand workspace is shared per (K, numBatch), so such passes are
 *     serialised on the device by an event inside the layout (they never corrupt each other).
 *   - A layout belongs to the CUDA device that was current when it was built; every entry point
 *     that takes a layout fails with SDDMM_E_ARG when another device is current.
 *   - Run calls are enqueue-only once sddmm_prepare() has been called for the (K, numBatch, plan)
 *     in use; without it the first run for a new K builds the K-dependent private layouts
 *     (allocations, sorts, one stream synchronisation) and is NOT capturable in a CUDA graph.
 */
#ifndef SDDMM_B200_H
#define SDDMM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDDMM_NULL_VALUE 0xFFFFFFFFu
#define SDDMM_ROW_PANEL 16u
#define SDDMM_BLOCK_COLS 16u

enum {
  SDDMM_OK = 0,
  SDDMM_E_ARG = 1,      /* bad argument (null pointer, K % 4 != 0, capacity too small ...) */
  SDDMM_E_CUDA = 2,     /* CUDA runtime error / no device */
  SDDMM_E_NOMEM = 3,    /* device allocation failed */
  SDDMM_E_UNSUPPORTED = 4
};

/* ---- library ------------------------------------------------------------------------------ */
int sddmm_b200_abi_version(void);            /* bumps when this header changes incompatibly */
const char* sddmm_last_error(void);
/* number of kernels this library has launched on this thread since the last reset (bench.py's
 * gpu_launches) */
uint64_t sddmm_launch_count(void);
void sddmm_launch_count_reset(void);

/* ---- a2: histogram block width --------------------------------------------------------------
 * replaces calculateBlockSize(const CSR&)          src/rowReordering.cu:1009-1025
 * free_mem_bytes == 0 -> query cudaMemGetInfo like the reference does (result then depends on
 * the device's free memory); pass an explicit value for reproducibility. */
uint32_t bsmr_calc_block_size(uint32_t M, uint32_t N, uint64_t free_mem_bytes);

/* ---- a3-a6: row-similarity reordering ---------------------------------------------------------
 * replaces bsa_rowReordering_gpu(matrix, alpha, block_size, num_clusters, time)
 *                                                  src/rowReordering.cu:1027-1095
 *          (calculateDispersion :49-93/:478-501, host sorts :1060-1062/:990,
 *           get_permutation_gpu :893-1007, bsa_clustering :325-432, zero-row strip :1081-1090)
 * and, through it, BSMR::rowReordering             src/BSMR.cpp:27-50.
 * block_size == 0 -> bsmr_calc_block_size(M, N, 0).
 * d_reorderedRows: capacity M; *numRows receives the number of non-empty rows written.
 * numClusters (optional) receives the reference's value (incl. its :996 quirk).
 * ms (optional) receives device time in milliseconds. */
int bsmr_row_reorder_dev(const uint32_t* d_rowOff, const uint32_t* d_colIdx, uint32_t M, uint32_t N,
                         uint32_t nnz, float alpha, uint32_t block_size, uint32_t* d_reorderedRows,
                         uint32_t* numRows, int32_t* numClusters, float* ms, void* stream);
/* host-buffer form (what BSMR::rowReordering consumes/produces: host CSR in, host vector out) */
int bsmr_row_reorder(const uint32_t* h_rowOff, const uint32_t* h_colIdx, uint32_t M, uint32_t N,
                     uint32_t nnz, float alpha, uint32_t block_size, uint32_t* h_reorderedRows,
                     uint32_t* numRows, int32_t* numClusters, float* ms);
/* Same result (the permutation is the reference's, bit for bit, whatever the options); the options only choose
 * HOW the clustering kernel gets there, so that every variant can be parity-tested on its own.
 * A zero-initialised struct = defaults (= the environment variables of DESIGN.md section 9, else automatic). */
enum { BSMR_CLUSTER_AUTO = 0, BSMR_CLUSTER_LEGACY = 1, BSMR_CLUSTER_BATCHED = 2 };
enum { BSMR_TRISTATE_AUTO = 0, BSMR_TRISTATE_OFF = 1, BSMR_TRISTATE_ON = 2 };
typedef struct {
  uint32_t kernel;     /* BSMR_CLUSTER_*: one cluster per sweep (k_cluster) / up to `batch` clusters per sweep */
  uint32_t batch;      /* 0 = as many as fit shared memory, else 1, 2, 4 or 8 (template instances)            */
  uint32_t laneRows;   /* BSMR_TRISTATE_*: one candidate row per LANE for short-row matrices                    */
  uint32_t signature;  /* BSMR_TRISTATE_*: bitmap-signature upper bound that rejects candidates early          */
  uint32_t reserved[4];
} bsmr_reorder_opts;
int bsmr_row_reorder_dev_ex(const uint32_t* d_rowOff, const uint32_t* d_colIdx, uint32_t M, uint32_t N,
                            uint32_t nnz, float alpha, uint32_t block_size, const bsmr_reorder_opts* opts,
                            uint32_t* d_reorderedRows, uint32_t* numRows, int32_t* numClusters, float* ms,
                            void* stream);
int bsmr_row_reorder_ex(const uint32_t* h_rowOff, const uint32_t* h_colIdx, uint32_t M, uint32_t N,
                        uint32_t nnz, float alpha, uint32_t block_size, const bsmr_reorder_opts* opts,
                        uint32_t* h_reorderedRows, uint32_t* numRows, int32_t* numClusters, float* ms);

/* optional introspection used by the parity tests (K1 encode / dispersion, a3):
 * d_dispersion[M]; returns nbpr in *numBlocksPerRow. */
int bsmr_dispersion_dev(const uint32_t* d_rowOff, const uint32_t* d_colIdx, uint32_t M, uint32_t N,
                        uint32_t nnz, uint32_t block_size, uint32_t* d_dispersion,
                        uint32_t* numBlocksPerRow, void* stream);

/* ---- a7 + a8: column reordering, dense/sparse split and the RPHM device layout ---------------
 * replaces colReordering_cpu(...)                   src/colReordering.cu:274-404
 *          BSMR::colReordering                      src/BSMR.cpp:52-81
 *          RPHM::RPHM(matrix, bsmr)                 src/BSMR.cpp:83-265
 * The layout object owns device copies of every array named in include/BSMR.hpp:39-49, 85-104.
 * panelBegin/panelEnd select a contiguous range of row panels of the reordered matrix (multi-GPU
 * row-panel shards); pass 0 / UINT32_MAX for all panels.  Offsets inside a shard
 * are relative to the shard; sparseValues / blockValues still hold GLOBAL CSR indices. */
typedef struct bsmr_layout bsmr_layout;

int bsmr_layout_build_dev(const uint32_t* d_rowOff, const uint32_t* d_colIdx, uint32_t M, uint32_t N,
                          uint32_t nnz, const uint32_t* d_reorderedRows, uint32_t numRows, float delta,
                          uint32_t panelBegin, uint32_t panelEnd, bsmr_layout** out, float* msColReorder,
                          float* msRphm, void* stream);
int bsmr_layout_build(const uint32_t* h_rowOff, const uint32_t* h_colIdx, uint32_t M, uint32_t N,
                      uint32_t nnz, const uint32_t* h_reorderedRows, uint32_t numRows, float delta,
                      bsmr_layout** out, float* msColReorder, float* msRphm);
/* flags: whether the private 128x128 full-tile layout of the tile plan is built next to the BSMR/RPHM arrays
 * (AUTO: when at least ~1 % of the tile slots are stored; the BSMR/RPHM arrays are identical either way). */
enum { BSMR_BUILD_TILES_AUTO = 0, BSMR_BUILD_TILES_ALWAYS = 1, BSMR_BUILD_TILES_NEVER = 2 };
int bsmr_layout_build_dev_ex(const uint32_t* d_rowOff, const uint32_t* d_colIdx, uint32_t M, uint32_t N,
                             uint32_t nnz, const uint32_t* d_reorderedRows, uint32_t numRows, float delta,
                             uint32_t panelBegin, uint32_t panelEnd, uint32_t flags, bsmr_layout** out,
                             float* msColReorder, float* msRphm, void* stream);
int bsmr_layout_build_ex(const uint32_t* h_rowOff, const uint32_t* h_colIdx, uint32_t M, uint32_t N,
                         uint32_t nnz, const uint32_t* h_reorderedRows, uint32_t numRows, float delta,
                         uint32_t flags, bsmr_layout** out, float* msColReorder, float* msRphm);
void bsmr_layout_destroy(bsmr_layout*);
/* On-disk cache of a layout (no reference counterpart; SURVEY.md 8f rank 4: reordering costs 10^4 x one
 * SDDMM pass, so deployments persist it keyed by (matrix, alpha, delta, block_size)).  Versioned file. */
int bsmr_layout_save(const bsmr_layout*, const char* path);
int bsmr_layout_load(const char* path, bsmr_layout** out);

typedef enum {
  BSMR_REORDERED_ROWS = 0,      /* [numRows]      BSMR::reorderedRows()      */
  BSMR_DENSE_COLS = 1,          /* [denseColOffsets[P]]                       */
  BSMR_DENSE_COL_OFFSETS = 2,   /* [P+1]                                      */
  BSMR_SPARSE_COLS = 3,         /* [sparseColOffsets[P]] (sentinel N included)*/
  BSMR_SPARSE_COL_OFFSETS = 4,  /* [P+1]                                      */
  BSMR_SPARSE_VALUE_OFFSETS = 5,/* [P+1]                                      */
  RPHM_BLOCK_OFFSETS = 6,       /* [P+1]                                      */
  RPHM_BLOCK_VALUES = 7,        /* [numBlocks*256] CSR index or NULL_VALUE    */
  RPHM_SPARSE_VALUES = 8,       /* [numSparse] CSR index                      */
  RPHM_SPARSE_RELATIVE_ROWS = 9,/* [numSparse] 0..15                          */
  RPHM_SPARSE_COL_INDICES = 10, /* [numSparse]                                */
  RPHM_DENSE_ROW_PANEL_IDS = 11,
  RPHM_DENSE_COL_BLOCK_ITERS = 12,
  RPHM_SPARSE_ROW_PANEL_IDS = 13,
  RPHM_SPARSE_COL_BLOCK_ITERS = 14,
  BSMR_ARRAY_COUNT = 15
} bsmr_array_id;

typedef struct {
  uint32_t M, N, nnz;
  uint32_t numRows;        /* non-empty rows (|reorderedRows|) covered by this layout */
  uint32_t numRowPanels;   /* BSMR::numRowPanels()                                    */
  uint32_t panelBegin;     /* first global panel of this shard                        */
  uint32_t numDenseBlocks; /* RPHM::getNumDenseBlocks()                               */
  uint32_t numSparseValues;
  uint32_t numDenseValues; /* stored entries that fall in dense blocks                */
  uint32_t maxNumDenseColBlocksInRowPanel;
  uint32_t maxNumSparseColBlocksInRowPanel;
  uint32_t numDenseThreadBlocks;
  uint32_t numSparseThreadBlocks;
} bsmr_layout_info;

/* ---- f2: the reordering statistics the reference logs after every run ------------------------
 * replaces evaluationReordering(matrix, bsmr, logger)                src/BSMR.cpp:826-925
 *   numDenseBlock   -> [bsmr_numDenseBlock]   dense blocks whose density (stored entries / 256) >= delta
 *   averageDensity  -> [bsmr_averageDensity]  density summed over every non-empty dense block, divided by
 *                                             numDenseBlock (the reference's expression, :917)
 *   num*ThreadBlocks-> [bsmr_numDenseThreadBlocks] / [bsmr_numSparseThreadBlocks]   (:843-849)
 *   numSparseData / numDenseData -> [bsmr_numSparseData] / [bsmr_numDenseData] = nnz - numSparseData (:923-924)
 * and calculateNumDenseBlocksAndAverageDensityInOriginalMatrix(delta, matrix)   src/BSMR.cpp:953-994
 *   -> [original_numDenseBlock], [original_averageDensity]: 16x16 blocks of the UNreordered matrix
 *      (edge blocks use their clipped size) with density >= delta, and the mean density of those.
 * Both run on the device (a key sort / a reduction) instead of the reference's host loops. */
typedef struct {
  uint32_t numDenseBlock;
  float averageDensity;
  uint32_t numDenseThreadBlocks, numSparseThreadBlocks;
  uint32_t numDenseData, numSparseData;
} bsmr_eval;
int bsmr_layout_eval(const bsmr_layout*, float delta, bsmr_eval* out);
int bsmr_original_block_stats_dev(const uint32_t* d_rowOff, const uint32_t* d_colIdx, uint32_t M, uint32_t N,
                                  uint32_t nnz, float delta, uint32_t* numDenseBlocks, float* averageDensity,
                                  void* stream);
int bsmr_original_block_stats(const uint32_t* h_rowOff, const uint32_t* h_colIdx, uint32_t M, uint32_t N,
                              uint32_t nnz, float delta, uint32_t* numDenseBlocks, float* averageDensity);

int bsmr_layout_get_info(const bsmr_layout*, bsmr_layout_info* out);
size_t bsmr_layout_array_len(const bsmr_layout*, bsmr_array_id which);
const uint32_t* bsmr_layout_array_dev(const bsmr_layout*, bsmr_array_id which);
int bsmr_layout_array_to_host(const bsmr_layout*, bsmr_array_id which, uint32_t* h_dst, size_t capacity);

/* ---- a9-a11: the SDDMM itself ---------------------------------------------------------------
 * replaces sddmm_gpu(M, N, K, dA, dB, rphm, dP, logger)   src/sddmmKernel.cu:2539-2663
 *          sddmm_gpu_k32(...)                             src/sddmmKernel.cu:2665-2762
 *          and the kernels they launch (:213-351, :355-488, :1994-2104, :2109-2199).
 * One call = one pass: dense-block tcgen05 kernel and residual CUDA-core kernel, concurrently.
 * K must be a positive multiple of 4 (16-byte rows); the reference asks for K % 32 == 0.
 * Entries of d_P not covered by this layout (other shards) are left untouched. */
int sddmm_run_dev(const bsmr_layout*, uint32_t K, const float* d_A, const float* d_B, float* d_P,
                  void* stream);
/* ---- kernel plan, chosen PER CALL ---------------------------------------------------------------
 * The reference picks its kernels at compile time (commented-out launch sites, src/sddmmKernel.cu:2583-2646).
 * Here every kernel that can serve a pass is selectable per call, so each one is parity-tested by name:
 *   plan     BSMR  = dense 16x16 blocks (tcgen05) || residual (CUDA cores): the reference's split  (:2575, :2617)
 *            TILE  = every stored entry through whole 128x128 tcgen05 tiles with a sampled epilogue
 *   dense    REG   = k_sddmm_dense      (operands staged through registers, cvt.rna.tf32 on the way)
 *            TMA   = k_sddmm_dense_tma  (TMA tile::gather4 of the denseCols rows from TF32-rounded copies)
 *   residual PANEL = k_sddmm_residual   (one 16-row panel per CTA; any K % 4 == 0)
 *            SUPERPANEL = k_sddmm_residual_sp (K in {32, 64, 128, 256, 512}; A rows of 16*G panels in shared
 *                         memory, B^T rows reused from registers along column runs: masks, uniform matrices)
 *            STREAM = k_sddmm_residual_stream (same K; entries in row order, A fragment in registers, up to 8
 *                         gathered B^T rows in flight per lane, L2 eviction hints: graphs, where B >> L2)
 *   tile     REG / TMA / TMA_CLUSTER = k_sddmm_tile / k_sddmm_tile_tma / k_sddmm_tile_tma4
 *            TMA_PAIR = k_sddmm_tile_pair: persistent CTA pairs, one 256x256 tcgen05.mma.cta_group::2 accumulator
 *                       per 2x2 group of tiles, double-buffered in TMEM (epilogue under the next main loop)
 * AUTO everywhere = the library's cost model (the defaults of sddmm_run_dev).  A choice that cannot serve the
 * call (SUPERPANEL with K = 36, TILE on a layout built with BSMR_BUILD_TILES_NEVER ...) fails with
 * SDDMM_E_UNSUPPORTED instead of silently running something else. */
enum { SDDMM_PLAN_AUTO = 0, SDDMM_PLAN_BSMR = 1, SDDMM_PLAN_TILE = 2 };
enum { SDDMM_DENSE_AUTO = 0, SDDMM_DENSE_REG = 1, SDDMM_DENSE_TMA = 2 };
enum { SDDMM_RESIDUAL_AUTO = 0, SDDMM_RESIDUAL_PANEL = 1, SDDMM_RESIDUAL_SUPERPANEL = 2, SDDMM_RESIDUAL_STREAM = 3 };
enum { SDDMM_TILE_AUTO = 0, SDDMM_TILE_REG = 1, SDDMM_TILE_TMA = 2, SDDMM_TILE_TMA_CLUSTER = 3, SDDMM_TILE_TMA_PAIR = 4 };
/* operands  EXACT = the reference's arithmetic: TF32 (round-to-nearest) operands into the tensor cores for dense
 *                   blocks / tiles, fp32 operands for the residual; fp32 accumulation everywhere (the default)
 *           FP16  = fp16 (round-to-nearest) copies of the operands where that halves the limiting traffic: the A
 *                   tile of the super-panel residual kernel in shared memory, both operands of the TMA tile kernel
 *                   (tcgen05 kind::f16).  fp32 accumulation.  Same 11 significant bits as TF32, so results stay
 *                   within the reference's checkData tolerance, but |operand| must stay below 65504 and residual
 *                   values are no longer the exact path's.  Opt-in only; AUTO never picks it. */
enum { SDDMM_OPERANDS_EXACT = 0, SDDMM_OPERANDS_FP16 = 1 };
typedef struct {
  uint32_t plan, dense, residual, tile;
  uint32_t tileStages;   /* 0 = default, else 2..4 operand stages of the TMA tile kernels                      */
  uint32_t operands;     /* SDDMM_OPERANDS_*                                                                   */
  uint32_t reserved[2];
} sddmm_plan;
/* fills *out with the defaults: AUTO unless an SDDMM_B200_* environment variable says otherwise */
void sddmm_plan_default(sddmm_plan* out);
/* resolves every AUTO of `in` (NULL = defaults) for this layout / K / numBatch: *out names the kernels a run
 * with `in` WILL launch (dense / residual are AUTO in *out when that part has no work) */
int sddmm_plan_resolve(const bsmr_layout*, uint32_t K, uint32_t numBatch, const sddmm_plan* in, sddmm_plan* out);
/* builds (and keeps, keyed by K / numBatch) every K-dependent private layout and workspace the plan needs, so
 * that later runs with the same arguments only enqueue work (CUDA-graph capturable). */
int sddmm_prepare(const bsmr_layout*, uint32_t K, uint32_t numBatch, const sddmm_plan* plan);
int sddmm_run_dev_ex(const bsmr_layout*, uint32_t K, uint32_t numBatch, const float* d_A, const float* d_B,
                     float* d_P, const sddmm_plan* plan, void* stream);
/* replaces sddmm_gpu_batch(numBatch, M, N, K, nnz, dA, dB, rphm, dP, time)   src/sddmmKernel.cu:2764-2850
 * One layout, numBatch independent (A, B, P) triples stored back to back: batch b uses
 * d_A + b*M*K, d_B + b*N*K, d_P + b*nnz (the reference's blockIdx.z indexing, :1280-1283).
 * One launch per kernel covers every batch. */
int sddmm_run_batch_dev(const bsmr_layout*, uint32_t K, uint32_t numBatch, const float* d_A, const float* d_B,
                        float* d_P, void* stream);
/* `iters` timed passes after `warmup` untimed ones; ms_* are per-pass means (CUDA events on the
 * launching streams).  Any of the three outputs may be NULL. */
int sddmm_run_timed_dev(const bsmr_layout*, uint32_t K, const float* d_A, const float* d_B, float* d_P,
                        int warmup, int iters, float* msDense, float* msSparse, float* msTotal);
/* replaces sddmm_gpu(const Matrix&, const Matrix&, const RPHM&, CSR&, Logger&)
 *                                                         src/sddmmKernel.cu:2518-2537
 * host A, B in; host P out; H2D and D2H inside the call. msTotal (optional) = whole call. */
int sddmm_run_host(const bsmr_layout*, uint32_t K, const float* h_A, const float* h_B, float* h_P,
                   float* msTotal);

/* Pipelined form of sddmm_run_host for callers that stream many (A, B) batches through one layout:
 * the call only ENQUEUES  H2D(A, B) -> SDDMM pass -> D2H(P)  for `slot` (0 or 1) on three internal streams
 * and returns; slot s may be reused after sddmm_host_sync() or once two later calls have been enqueued and
 * synced.  Host buffers must be page-locked for the copies to overlap.  With two slots the H2D of batch
 * i+1, the kernels of batch i and the D2H of batch i-1 run concurrently (PCIe is full duplex).
 * (The reference's host overload, src/sddmmKernel.cu:2518-2537, is synchronous; this is its streaming twin.) */
int sddmm_run_host_async(const bsmr_layout*, uint32_t K, const float* h_A, const float* h_B, float* h_P, int slot);
int sddmm_host_sync(const bsmr_layout*);
/* Host <-> device bytes of the most recent sddmm_run_host / _async call on this layout.  Both host entry points move
 * only what the pass reads when they can: if h_A and h_B are page-locked and the layout references at most 85 % of
 * the rows of A plus the columns of S (graphs leave many empty), a gather kernel reads just the referenced A rows
 * (reorderedRows) and B^T rows through the mapped host pointers instead of copying the whole arrays
 * (SDDMM_B200_H2D = auto | full | gather).  (The reference copies whole arrays, src/sddmmKernel.cu:2518-2537.) */
int sddmm_host_traffic(const bsmr_layout*, uint64_t* h2dBytes, uint64_t* d2hBytes);

/* ---- whole path on host buffers ----------------------------------------------------------------
 * replaces sddmm(options, A, B, P, logger)                src/sddmm.cu:10-39
 * = row reorder -> layout -> one SDDMM pass.  `layoutOut` (optional) receives the layout so the
 * caller can read the BSMR/RPHM arrays or re-run; otherwise it is destroyed before returning. */
typedef struct {
  float rowReorderMs, colReorderMs, rphmMs, sddmmMs;
  int32_t numClusters;
  uint32_t blockSize;
} sddmm_stats;
int sddmm_host(const uint32_t* h_rowOff, const uint32_t* h_colIdx, uint32_t M, uint32_t N, uint32_t nnz,
               uint32_t K, const float* h_A, const float* h_B, float alpha, float delta,
               uint32_t block_size, float* h_P, sddmm_stats* stats, bsmr_layout** layoutOut);

/* ---- f1: CSR assembly on the device for the file loaders ------------------------------------------
 * replaces the sort / duplicate-check / offset stages of
 *   sparseMatrix::CSR<float>::initializeFromMtxFile          src/Matrix.cpp:398-480
 *     (std::set duplicate check :447-461, thrust::host stable sort by row :467, rowOffsets :470-479)
 * for triplets a loader has already parsed and range-checked: n entries in FILE order, 0-based.
 * Output: rows ascending, FILE order kept inside a row (the reference sorts by row only).  *hasDuplicate = 1 when
 * some (row, col) appears twice (the reference rejects such a file); the outputs are then untouched.
 * h_vals / h_values may be NULL (pattern only). */
int sddmm_coo_to_csr(const uint32_t* h_rows, const uint32_t* h_cols, const float* h_vals, uint32_t nnz, uint32_t M,
                     uint32_t N, uint32_t* h_rowOff, uint32_t* h_colIdx, float* h_values, int* hasDuplicate);

/* ---- e: multi-GPU row-panel sharding ----------------------------------------------------------
 * (no reference counterpart: the reference is single-GPU.)  Cuts the P row panels of the
 * reordered matrix into `numShards` contiguous ranges with equal non-zero counts; cuts[s] ..
 * cuts[s+1] is shard s (numShards+1 entries).  Host arrays in, host array out. */
int bsmr_shard_plan(const uint32_t* h_rowOff, const uint32_t* h_reorderedRows, uint32_t numRows,
                    uint32_t numShards, uint32_t* h_cuts);
/* the same cuts from device-resident arrays (a 33 M-row order never has to visit the host); h_cuts on the host */
int bsmr_shard_plan_dev(const uint32_t* d_rowOff, const uint32_t* d_reorderedRows, uint32_t numRows,
                        uint32_t numShards, uint32_t* h_cuts, void* stream);

/* One PROCESS per GPU.  NCCL (bound at run time with dlopen; SDDMM_B200_NCCL_LIB overrides the library name) is
 * used for the two one-time replications only -- rank 0's row order and B -- and for the optional final merge of P.
 * The steady state (sddmm_mgpu_run) has no collective: every rank computes the P entries of its own panel range.
 *   rank 0:      sddmm_mgpu_unique_id(id)          -> ship the 128 bytes to every rank (file, MPI, torch.distributed)
 *   every rank:  cudaSetDevice(local); sddmm_mgpu_init(rank, world, id, &g);
 *                rank 0 runs bsmr_row_reorder_dev; then every rank calls
 *                sddmm_mgpu_shard(g, ...)          -> broadcasts (numRows, reorderedRows) from rank 0, cuts the panels
 *                                                     into `world` nnz-balanced ranges (the bsmr_shard_plan rule) and
 *                                                     builds THIS rank's layout (panelBegin/panelEnd = its range)
 *                sddmm_mgpu_bcast(g, d_B, bytes, 0, stream)   once per B
 *                sddmm_mgpu_run(g, layout, K, d_A, d_B, d_P, stream)   per pass, no communication
 *                sddmm_mgpu_gather(g, d_P, nnz, stream)       optional: every rank ends with the full P
 * d_reorderedRows must have capacity M on every rank; *numRows is an input on rank 0 and an output elsewhere.
 * h_cuts (optional, world + 1 entries) receives the panel cuts.  world == 1 needs no id and no NCCL. */
#define SDDMM_MGPU_ID_BYTES 128
typedef struct sddmm_mgpu sddmm_mgpu;
int sddmm_mgpu_unique_id(void* id128);
int sddmm_mgpu_init(int rank, int world, const void* id128, sddmm_mgpu** out);
void sddmm_mgpu_destroy(sddmm_mgpu*);
int sddmm_mgpu_shard(sddmm_mgpu*, const uint32_t* d_rowOff, const uint32_t* d_colIdx, uint32_t M, uint32_t N,
                     uint32_t nnz, uint32_t* d_reorderedRows, uint32_t* numRows, float delta, uint32_t flags,
                     bsmr_layout** out, uint32_t* h_cuts, float* msColReorder, float* msRphm, void* stream);
/* Cost calibration of the cuts (optional, collective): equal non-zero counts are not equal times when the shards'
 * B^T rows hit the L2 at different rates (power-law graphs: the shard holding the hub rows is cheaper per non-zero).
 * Every rank passes the device time of its last pass(es); a non-zero of shard r is charged ms_r / nnz_r, the cuts
 * move so that every rank's estimated time is equal, and this rank's layout is rebuilt when its range changed
 * (*layout is destroyed and replaced).  One or two rounds settle it; not part of the steady state. */
int sddmm_mgpu_rebalance(sddmm_mgpu*, const uint32_t* d_rowOff, const uint32_t* d_colIdx, uint32_t M, uint32_t N,
                         uint32_t nnz, const uint32_t* d_reorderedRows, uint32_t numRows, float delta, uint32_t flags,
                         float myMs, bsmr_layout** layout, uint32_t* h_cuts, void* stream);
/* the rule sddmm_mgpu_rebalance applies, on host arrays (no device, no communicator): prefix[numPanels + 1] of the
 * per-panel non-zero counts, the old cuts, the measured time of every shard -> new cuts (numShards + 1 entries) */
int bsmr_rebalance_cuts(const uint64_t* h_panelNnzPrefix, uint32_t numPanels, const uint32_t* h_oldCuts,
                        const float* h_ms, uint32_t numShards, uint32_t* h_newCuts);
int sddmm_mgpu_bcast(sddmm_mgpu*, void* d_buf, size_t bytes, int root, void* stream);
int sddmm_mgpu_run(sddmm_mgpu*, const bsmr_layout*, uint32_t K, const float* d_A, const float* d_B, float* d_P,
                   void* stream);
int sddmm_mgpu_gather(sddmm_mgpu*, float* d_P, size_t count, void* stream);
/* Host-buffer pass of a multi-GPU job (the multi-GPU form of sddmm_run_host, collective, synchronous): every rank
 * holds the same host A (M x K) and B (N x K) but copies only its 1/world slice of each over its own PCIe link;
 * the slices are all-gathered over NVLink (so the host is read ONCE per job step, not once per rank), every rank
 * computes its panel range, the disjoint pieces of P are summed onto `root`, and root copies the whole P to h_P
 * (h_P may be NULL elsewhere).  msTotal (optional) = this rank's device time for the whole call.
 * With page-locked h_A / h_B and a matrix that leaves at least 15 % of its rows + columns unreferenced, a rank reads
 * only 1/world of the REFERENCED rows (non-empty rows of S, referenced columns; lists kept by sddmm_mgpu_shard)
 * through the mapped pointers into a packed buffer; the packed buffers are all-gathered and unpacked on the device. */
int sddmm_mgpu_run_host(sddmm_mgpu*, const bsmr_layout*, uint32_t K, const float* h_A, const float* h_B, float* h_P,
                        int root, float* msTotal);
/* host -> device bytes THIS rank moved in its most recent sddmm_mgpu_run_host call */
int sddmm_mgpu_host_traffic(const sddmm_mgpu*, uint64_t* h2dBytes);

#ifdef __cplusplus
}
#endif
#endif /* SDDMM_B200_H */
