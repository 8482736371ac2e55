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
