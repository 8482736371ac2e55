// host_api.cpp -- C++ mirror of the reference's host API for the SDDMM path, on top of the C ABI.
//   BSMR / RPHM           <- src/BSMR.cpp:16-265
//   sddmm / sddmm_testMode / checkSddmm <- src/sddmm.cu:10-118
//   sddmm_gpu (both overloads)          <- src/sddmmKernel.cu:2518-2663
// Error convention of the reference: void functions, failures are printed to stderr and execution
// continues (include/cudaErrorCheck.cuh:12-16); the library's error string is what gets printed.
#include <cuda_runtime_api.h>
#include <omp.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <fstream>

#include "sddmm.hpp"

namespace {
bool ok(int rc, const char* what) {
  if (rc != SDDMM_OK) std::fprintf(stderr, "%s failed (%d): %s\n", what, rc, sddmm_last_error());
  return rc == SDDMM_OK;
}
std::vector<UIN> fetch(const bsmr_layout* l, bsmr_array_id id) {
  std::vector<UIN> v(bsmr_layout_array_len(l, id));
  if (!v.empty()) ok(bsmr_layout_array_to_host(l, id, v.data(), v.size()), "bsmr_layout_array_to_host");
  return v;
}
}  // namespace

UIN calculateBlockSize(const sparseMatrix::CSR<float>& m, uint64_t freeMem) {
  return bsmr_calc_block_size(m.row(), m.col(), freeMem);
}

BSMR::BSMR(float alpha, float delta, const sparseMatrix::CSR<float>& matrix, int numIterations, UIN blockSize) {
  rowReordering(alpha, matrix, numIterations, blockSize);
  colReordering(delta, matrix, reorderedRows_, numIterations);
}

void BSMR::rowReordering(float alpha, const sparseMatrix::CSR<float>& m, int numIterations, UIN blockSize) {
  blockSize_ = blockSize ? blockSize : calculateBlockSize(m);
  std::vector<UIN> out(m.row());
  UIN n = 0;
  int32_t ncl = 1;
  float total = 0.f;
  for (int it = 0; it < numIterations; ++it) {
    float ms = 0.f;
    if (!ok(bsmr_row_reorder(m.rowOffsets().data(), m.colIndices().data(), m.row(), m.col(), m.nnz(), alpha, blockSize_,
                             out.data(), &n, &ncl, &ms), "bsmr_row_reorder"))
      return;
    total += ms;
  }
  out.resize(n);
  reorderedRows_ = std::move(out);
  numClusters_ = ncl;
  rowReorderingTime_ = total / numIterations;
  numRowPanels_ = static_cast<int>((reorderedRows_.size() + ROW_PANEL_SIZE - 1) / ROW_PANEL_SIZE);
}

void BSMR::colReordering(float delta, const sparseMatrix::CSR<float>& m, const std::vector<UIN>& reorderedRows,
                         int numIterations) {
  if (!reorderedRows.empty()) {
    reorderedRows_ = reorderedRows;
    numRowPanels_ = static_cast<int>((reorderedRows_.size() + ROW_PANEL_SIZE - 1) / ROW_PANEL_SIZE);
  }
  float totalC = 0.f, totalR = 0.f;
  for (int it = 0; it < numIterations; ++it) {
    bsmr_layout* l = nullptr;
    float msC = 0.f, msR = 0.f;
    if (!ok(bsmr_layout_build(m.rowOffsets().data(), m.colIndices().data(), m.row(), m.col(), m.nnz(),
                              reorderedRows_.data(), static_cast<UIN>(reorderedRows_.size()), delta, &l, &msC, &msR),
            "bsmr_layout_build"))
      return;
    layout_ = LayoutPtr(l, [](bsmr_layout* p) { bsmr_layout_destroy(p); });
    totalC += msC;
    totalR += msR;
  }
  colReorderingTime_ = totalC / numIterations;
  rphmTime_ = totalR / numIterations;
  denseCols_ = fetch(layout_.get(), BSMR_DENSE_COLS);
  denseColOffsets_ = fetch(layout_.get(), BSMR_DENSE_COL_OFFSETS);
  sparseCols_ = fetch(layout_.get(), BSMR_SPARSE_COLS);
  sparseColOffsets_ = fetch(layout_.get(), BSMR_SPARSE_COL_OFFSETS);
  sparseValueOffsets_ = fetch(layout_.get(), BSMR_SPARSE_VALUE_OFFSETS);
}

RPHM::RPHM(const sparseMatrix::CSR<float>&, const BSMR& bsmr) : layout_(bsmr.layout()) {
  if (layout_) ok(bsmr_layout_get_info(layout_.get(), &info_), "bsmr_layout_get_info");
}
std::vector<UIN> RPHM::hostCopy(bsmr_array_id id) const { return fetch(layout_.get(), id); }

// ---- sddmm_gpu: host-matrix overload (H2D A, B; one pass; D2H P) -----------------------------------
void sddmm_gpu(const Matrix<float>& A, const Matrix<float>& B, const RPHM& rphm, sparseMatrix::CSR<float>& P,
               Logger& logger) {
  if (!rphm.layout()) { std::fprintf(stderr, "sddmm_gpu: empty RPHM\n"); return; }
  std::vector<float> out(P.nnz());
  float ms = 0.f;
  if (!ok(sddmm_run_host(rphm.layout().get(), A.col(), A.data(), B.data(), out.data(), &ms), "sddmm_run_host")) return;
  P.setValues() = out;
  // kernel-only timing like the reference's sddmmTime_ (device-resident operands, numITER_ passes)
  const size_t nA = A.size(), nB = B.size();
  float *dA = nullptr, *dB = nullptr, *dP = nullptr;
  if (cudaMalloc((void**)&dA, nA * 4) == cudaSuccess && cudaMalloc((void**)&dB, nB * 4) == cudaSuccess &&
      cudaMalloc((void**)&dP, (size_t)(P.nnz() ? P.nnz() : 1) * 4) == cudaSuccess) {
    cudaMemcpy(dA, A.data(), nA * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), nB * 4, cudaMemcpyHostToDevice);
    sddmm_gpu(P.row(), P.col(), A.col(), dA, dB, rphm, dP, logger);
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dP);
}

// ---- raw device-pointer overload -------------------------------------------------------------------
void sddmm_gpu(UIN, UIN, UIN K, const float* dA, const float* dB, const RPHM& rphm, float* dP, Logger& logger) {
  float d = 0.f, s = 0.f, t = 0.f;
  if (!ok(sddmm_run_timed_dev(rphm.layout().get(), K, dA, dB, dP, 3, logger.numITER_ > 0 ? logger.numITER_ : 10, &d, &s,
                              &t), "sddmm_run_timed_dev"))
    return;
  logger.sddmmTime_ = t;
  sddmm_plan resolved{};
  if (sddmm_plan_resolve(rphm.layout().get(), K, 1, nullptr, &resolved) == SDDMM_OK)
    logger.residualCta_ = resolved.residual == SDDMM_RESIDUAL_SUPERPANEL ? (K >= 256u ? 512u : 1024u) : 256u;
  logger.denseTime_ = d;
  logger.sparseTime_ = s;
}

void evaluationReordering(const sparseMatrix::CSR<float>& matrix, const RPHM& rphm, Logger& logger) {
  bsmr_eval ev{};
  if (!ok(bsmr_layout_eval(rphm.layout().get(), logger.delta_, &ev), "bsmr_layout_eval")) return;
  uint32_t nd = 0;
  float ad = 0.f;
  ok(bsmr_original_block_stats(matrix.rowOffsets().data(), matrix.colIndices().data(), matrix.row(), matrix.col(),
                               matrix.nnz(), logger.delta_, &nd, &ad), "bsmr_original_block_stats");
  logger.numDenseBlock_ = static_cast<int>(ev.numDenseBlock);
  logger.averageDensity_ = ev.averageDensity;
  logger.numDenseThreadBlocks_ = static_cast<int>(ev.numDenseThreadBlocks);
  logger.numSparseThreadBlocks_ = static_cast<int>(ev.numSparseThreadBlocks);
  logger.numSparseData_ = ev.numSparseData;
  logger.numDenseData_ = ev.numDenseData;
  logger.originalNumDenseBlock_ = static_cast<int>(nd);
  logger.originalAverageDensity_ = ad;
}

void sddmm(const Options& options, const Matrix<float>& A, const Matrix<float>& B, sparseMatrix::CSR<float>& P,
           Logger& logger) {
  UIN bs = options.blockSize();
  if (!bs) bs = calculateBlockSize(P, options.freeMemForBlockSize());
  BSMR bsmr(options.similarityThresholdAlpha(), options.blockDensityThresholdDelta(), P, 1, bs);
  logger.rowReorderingTime_ = bsmr.rowReorderingTime();
  logger.colReorderingTime_ = bsmr.colReorderingTime();
  logger.reorderingTime_ = bsmr.reorderingTime();
  logger.rphmTime_ = bsmr.rphmTime();
  logger.numRowPanels_ = bsmr.numRowPanels();
  logger.numClusters_ = bsmr.numClusters();
  logger.blockSize_ = bsmr.blockSize();
  RPHM rphm(P, bsmr);
  sddmm_gpu(A, B, rphm, P, logger);
  evaluationReordering(P, rphm, logger);  // src/sddmm.cu:32
}

// ---- multi-GPU: one process per GPU, driven through the C ABI's sddmm_mgpu_* entry points ----------------
namespace {
struct DevMem {
  void* p = nullptr;
  explicit DevMem(size_t bytes) { if (cudaMalloc(&p, bytes ? bytes : 4) != cudaSuccess) p = nullptr; }
  ~DevMem() { cudaFree(p); }
  template <typename T> T* as() const { return static_cast<T*>(p); }
};
// rank 0 publishes the NCCL id through a file (written under a temporary name, then renamed); the others poll
bool exchange_id_through_file(const Options& o, unsigned char* id) {
  const std::string path = o.idFile();
  if (path.empty()) { std::fprintf(stderr, "multi-GPU run needs -u <id file>\n"); return false; }
  if (o.rank() == 0) {
    if (!ok(sddmm_mgpu_unique_id(id), "sddmm_mgpu_unique_id")) return false;
    const std::string tmp = path + ".tmp";
    std::ofstream f(tmp, std::ios::binary);
    f.write(reinterpret_cast<const char*>(id), SDDMM_MGPU_ID_BYTES);
    f.close();
    return std::rename(tmp.c_str(), path.c_str()) == 0;
  }
  for (int tries = 0; tries < 6000; ++tries) {  // up to ~60 s
    std::ifstream f(path, std::ios::binary);
    if (f.good() && f.read(reinterpret_cast<char*>(id), SDDMM_MGPU_ID_BYTES)) return true;
    struct timespec ts = {0, 10 * 1000 * 1000};
    nanosleep(&ts, nullptr);
  }
  std::fprintf(stderr, "rank %d: no NCCL id appeared in %s\n", o.rank(), path.c_str());
  return false;
}
}  // namespace

bool sddmm_multiGpu(const Options& options, const Matrix<float>& A, const Matrix<float>& B, sparseMatrix::CSR<float>& P,
                    Logger& logger) {
  const int rank = options.rank(), world = options.world();
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { std::fprintf(stderr, "no CUDA device\n"); return false; }
  const char* lr = std::getenv("LOCAL_RANK");
  cudaSetDevice((lr ? std::atoi(lr) : rank) % ndev);
  unsigned char id[SDDMM_MGPU_ID_BYTES] = {0};
  if (world > 1 && !exchange_id_through_file(options, id)) return false;
  sddmm_mgpu* g = nullptr;
  if (!ok(sddmm_mgpu_init(rank, world, id, &g), "sddmm_mgpu_init")) return false;
  const UIN M = P.row(), N = P.col(), nnz = P.nnz(), K = A.col();
  DevMem ro((size_t)(M + 1) * 4), ci((size_t)nnz * 4), rr((size_t)M * 4), dA(A.size() * 4), dB(B.size() * 4), dP((size_t)nnz * 4);
  bool good = ro.p && ci.p && rr.p && dA.p && dB.p && dP.p;
  bsmr_layout* lay = nullptr;
  if (good) {
    cudaMemcpy(ro.p, P.rowOffsets().data(), (size_t)(M + 1) * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(ci.p, P.colIndices().data(), (size_t)nnz * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dA.p, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    if (rank == 0) cudaMemcpy(dB.p, B.data(), B.size() * 4, cudaMemcpyHostToDevice);  // the others receive it
    cudaMemset(dP.p, 0, (size_t)nnz * 4);
    UIN numRows = 0;
    int32_t ncl = 0;
    float msRow = 0.f, msCol = 0.f, msRphm = 0.f;
    UIN bs = options.blockSize() ? options.blockSize() : calculateBlockSize(P, options.freeMemForBlockSize());
    if (rank == 0)
      good = ok(bsmr_row_reorder_dev(ro.as<UIN>(), ci.as<UIN>(), M, N, nnz, options.similarityThresholdAlpha(), bs,
                                     rr.as<UIN>(), &numRows, &ncl, &msRow, nullptr), "bsmr_row_reorder_dev");
    std::vector<UIN> cuts(world + 1, 0);
    good = good && ok(sddmm_mgpu_shard(g, ro.as<UIN>(), ci.as<UIN>(), M, N, nnz, rr.as<UIN>(), &numRows,
                                       options.blockDensityThresholdDelta(), BSMR_BUILD_TILES_AUTO, &lay, cuts.data(),
                                       &msCol, &msRphm, nullptr), "sddmm_mgpu_shard");
    good = good && ok(sddmm_mgpu_bcast(g, dB.p, B.size() * 4, 0, nullptr), "sddmm_mgpu_bcast");
    if (good) {
      bsmr_layout_info info{};
      bsmr_layout_get_info(lay, &info);
      float d = 0.f, s = 0.f, t = 0.f;
      good = ok(sddmm_run_timed_dev(lay, K, dA.as<float>(), dB.as<float>(), dP.as<float>(), 3,
                                    logger.numITER_ > 0 ? logger.numITER_ : 10, &d, &s, &t), "sddmm_run_timed_dev");
      good = good && ok(sddmm_mgpu_run(g, lay, K, dA.as<float>(), dB.as<float>(), dP.as<float>(), nullptr), "sddmm_mgpu_run");
      good = good && ok(sddmm_mgpu_gather(g, dP.as<float>(), nnz, nullptr), "sddmm_mgpu_gather");
      cudaDeviceSynchronize();
      std::vector<float> out(nnz);
      cudaMemcpy(out.data(), dP.p, (size_t)nnz * 4, cudaMemcpyDeviceToHost);
      P.setValues() = out;
      logger.sddmmTime_ = t; logger.denseTime_ = d; logger.sparseTime_ = s;
      logger.rowReorderingTime_ = msRow; logger.colReorderingTime_ = msCol; logger.reorderingTime_ = msRow + msCol;
      logger.rphmTime_ = msRphm; logger.numClusters_ = ncl; logger.blockSize_ = bs;
      logger.numRowPanels_ = static_cast<int>(info.numRowPanels);
      logger.rank_ = rank; logger.world_ = world;
      logger.shardNnz_ = info.numDenseValues + info.numSparseValues;
      logger.shardPanelBegin_ = cuts[rank]; logger.shardPanelEnd_ = cuts[rank + 1];
      logger.numDenseData_ = info.numDenseValues; logger.numSparseData_ = info.numSparseValues;
      logger.numDenseBlock_ = static_cast<int>(info.numDenseBlocks);
      logger.numDenseThreadBlocks_ = static_cast<int>(info.numDenseThreadBlocks);
      logger.numSparseThreadBlocks_ = static_cast<int>(info.numSparseThreadBlocks);
    }
  } else {
    std::fprintf(stderr, "rank %d: device allocation failed\n", rank);
  }
  if (lay) bsmr_layout_destroy(lay);
  sddmm_mgpu_destroy(g);
  return good;
}

// ---- checker (verification only) ---------------------------------------------------------------------
bool checkSddmm(const Matrix<float>& A, const Matrix<float>& B, const sparseMatrix::CSR<float>& S,
                const sparseMatrix::CSR<float>& P) {
  const UIN K = A.col();
  size_t errors = 0;
#pragma omp parallel for reduction(+ : errors)
  for (long row = 0; row < static_cast<long>(S.row()); ++row) {
    for (UIN i = S.rowOffsets()[row]; i < S.rowOffsets()[row + 1]; ++i) {
      const float* a = A.data() + static_cast<size_t>(row) * K;
      const float* b = B.data() + static_cast<size_t>(S.colIndices()[i]) * K;
      float v = 0.f;
      for (UIN k = 0; k < K; ++k) v += a[k] * b[k];
      const float d = std::fabs(v - P.values()[i]);
      if (d < 1e-5f) continue;
      const float mx = std::max(std::max(std::fabs(v), std::fabs(P.values()[i])), 1e-3f);
      if (!(d / mx < 1e-3f)) ++errors;
    }
  }
  if (errors) {
    std::printf("[checkData : NO PASS Error rate : %2.2f%%]\n", 100.0f * errors / static_cast<float>(P.nnz()));
    return false;
  }
  std::printf("| Pass! Result validates successfully.\n");
  return true;
}

// ---- alpha x delta x K sweep, one log file per (K, alpha, delta)  (src/sddmm.cu:62-118) ---------------
void sddmm_testMode(const Options& options, sparseMatrix::CSR<float>& P) {
  const float alphas[] = {0.1f, 0.3f, 0.5f, 0.7f, 0.9f};
  const float deltas[] = {0.0f, 0.1f, 0.3f, 0.5f, 0.7f, 0.9f, 1.1f};
  const UIN Ks[] = {32, 64, 128, 256};
  auto trimmed = [](float v) {
    char b[32];
    std::snprintf(b, sizeof b, "%g", v);
    return std::string(b);
  };
  BSMR bsmr;
  for (float alpha : alphas) {
    bsmr.rowReordering(alpha, P, 1, options.blockSize());
    for (float delta : deltas) {
      bsmr.colReordering(delta, P);
      RPHM rphm(P, bsmr);
      for (UIN k : Ks) {
        Matrix<float> A(P.row(), k, MatrixStorageOrder::row_major);
        A.makeData(1);
        Matrix<float> B(k, P.col(), MatrixStorageOrder::col_major);
        B.makeData(2);
        Logger logger;
        logger.getInformation(options);
        logger.getInformation(P);
        logger.getInformation(A, B);
        logger.alpha_ = alpha;
        logger.delta_ = delta;
        logger.rowReorderingTime_ = bsmr.rowReorderingTime();
        logger.colReorderingTime_ = bsmr.colReorderingTime();
        logger.reorderingTime_ = bsmr.reorderingTime();
        logger.numRowPanels_ = bsmr.numRowPanels();
        logger.numClusters_ = bsmr.numClusters();
        sddmm_gpu(A, B, rphm, P, logger);
        evaluationReordering(P, rphm, logger);  // src/sddmm.cu:102
        const std::string f = options.outputLogDirectory() + "BSMR_k_" + std::to_string(k) + "_a_" + trimmed(alpha) +
                              "_d_" + trimmed(delta) + ".log";
        std::ofstream fout(f, std::ios::app);
        if (fout.fail()) {
          std::fprintf(stderr, "Error, failed to open log file: %s\n", f.c_str());
          return;
        }
        fout << "\n---New data---\n";
        logger.printLogInformation(fout);
      }
    }
  }
}
