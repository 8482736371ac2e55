// oracle/ref_cpu_shim.cpp -- TEST INFRASTRUCTURE.
// extern "C" doors onto the UNMODIFIED reference host code, compiled from
// /root/reference where it lies (see oracle/Makefile, target _ref/libref_cpu.so).
// Used only to pin oracle/bsmr_oracle.c and, optionally, as the CPU baseline
// (cpu_baseline.kind == "reference").  No reference source is copied here.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "BSMR.hpp"       // reference: colReordering_cpu  (src/colReordering.cu:274)
#include "Matrix.hpp"     // reference: Matrix<T>, sparseMatrix::CSR<T>
#include "checkData.hpp"  // reference: checkOneData       (include/checkData.hpp:14-30)
#include "host.hpp"       // reference: sddmm_cpu<T>       (src/host.cpp:44)

struct RefCtx { Matrix<float> A, B; sparseMatrix::CSR<float> S, P; };

#include <omp.h>

extern "C" {

// colReordering_cpu on caller-provided CSR + row order.  Two-call protocol:
// call once with cols == nullptr to learn sizes from the offsets.
int ref_col_reorder(const uint32_t* rowOff, const uint32_t* colIdx, uint32_t M, uint32_t N,
                    uint32_t nnz, const uint32_t* reorderedRows, uint32_t numRows, float delta,
                    uint32_t* denseColOffsets, uint32_t* sparseColOffsets,
                    uint32_t* sparseValueOffsets, uint32_t* denseCols, uint32_t* sparseCols) {
  std::vector<float> vals(nnz, 1.0f);
  sparseMatrix::CSR<float> S(M, N, nnz, rowOff, colIdx, vals.data());
  std::vector<UIN> R(reorderedRows, reorderedRows + numRows);
  const UIN P = static_cast<UIN>(std::ceil(static_cast<float>(R.size()) / ROW_PANEL_SIZE));
  std::vector<UIN> dc, dco, sc, sco, svo;
  float t = 0.f;
  colReordering_cpu(S, P, R, delta, dc, dco, sc, sco, svo, t);
  std::memcpy(denseColOffsets, dco.data(), sizeof(UIN) * (P + 1));
  std::memcpy(sparseColOffsets, sco.data(), sizeof(UIN) * (P + 1));
  std::memcpy(sparseValueOffsets, svo.data(), sizeof(UIN) * (P + 1));
  if (denseCols) std::memcpy(denseCols, dc.data(), sizeof(UIN) * dc.size());
  if (sparseCols) std::memcpy(sparseCols, sc.data(), sizeof(UIN) * sc.size());
  return static_cast<int>(P);
}

// sddmm_cpu<float> (OpenMP inside the reference).
void ref_sddmm_cpu(const float* A, const float* B, const uint32_t* rowOff, const uint32_t* colIdx,
                   uint32_t M, uint32_t N, uint32_t K, uint32_t nnz, float* P) {
  Matrix<float> mA(M, K, MatrixStorageOrder::row_major, A);
  Matrix<float> mB(K, N, MatrixStorageOrder::col_major, B);
  std::vector<float> vals(nnz, 0.0f);
  sparseMatrix::CSR<float> S(M, N, nnz, rowOff, colIdx, vals.data());
  sparseMatrix::CSR<float> Pm(S);
  sddmm_cpu(mA, mB, S, Pm);
  std::memcpy(P, Pm.values().data(), sizeof(float) * nnz);
}

// Same loop but with Matrix objects built once by the caller (for timing).
void* ref_sddmm_prepare(const float* A, const float* B, const uint32_t* rowOff,
                        const uint32_t* colIdx, uint32_t M, uint32_t N, uint32_t K, uint32_t nnz) {
  std::vector<float> vals(nnz, 0.0f);
  auto* c = new RefCtx{Matrix<float>(M, K, MatrixStorageOrder::row_major, A),
                    Matrix<float>(K, N, MatrixStorageOrder::col_major, B),
                    sparseMatrix::CSR<float>(M, N, nnz, rowOff, colIdx, vals.data()),
                    sparseMatrix::CSR<float>()};
  c->P = c->S;
  return c;
}
void ref_sddmm_run(void* ctx, float* P /* may be null */) {
  auto* c = static_cast<RefCtx*>(ctx);
  sddmm_cpu(c->A, c->B, c->S, c->P);
  if (P) std::memcpy(P, c->P.values().data(), sizeof(float) * c->P.values().size());
}
// OpenMP thread count the reference's `#pragma omp parallel for` (src/host.cpp:52) will really use; n > 0 sets it
// first (torchrun exports OMP_NUM_THREADS=1 to its ranks, which would silently serialise the CPU baseline)
int ref_omp_threads(int n) {
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
}

void ref_sddmm_release(void* ctx) {
  delete static_cast<RefCtx*>(ctx);
}

// evaluationReordering (src/BSMR.cpp:826-925) on a BSMR built from a caller-provided row order through the
// reference's own BSMR::colReordering (CPU).  outi = numDenseBlock, numDenseThreadBlocks, numSparseThreadBlocks,
// numSparseData, numDenseData, originalNumDenseBlock; outf = averageDensity, originalAverageDensity.
void ref_evaluation_reordering(const uint32_t* rowOff, const uint32_t* colIdx, uint32_t M, uint32_t N,
                               uint32_t nnz, const uint32_t* reorderedRows, uint32_t numRows, float delta,
                               int* outi, float* outf) {
  std::vector<float> vals(nnz, 1.0f);
  sparseMatrix::CSR<float> S(M, N, nnz, rowOff, colIdx, vals.data());
  std::vector<UIN> R(reorderedRows, reorderedRows + numRows);
  BSMR bsmr;
  bsmr.colReordering(delta, S, R, 1);
  Logger lg;
  lg.delta_ = delta;
  evaluationReordering(S, bsmr, lg);
  outi[0] = lg.numDenseBlock_;
  outi[1] = lg.numDenseThreadBlocks_;
  outi[2] = lg.numSparseThreadBlocks_;
  outi[3] = lg.numSparseData_;
  outi[4] = lg.numDenseData_;
  outi[5] = lg.originalNumDenseBlock_;
  outf[0] = lg.averageDensity_;
  outf[1] = lg.originalAverageDensity_;
}

// checkOneData<float> applied element-wise; returns #mismatches.
size_t ref_check_data(const float* a, const float* b, size_t n) {
  size_t e = 0;
  for (size_t i = 0; i < n; ++i)
    if (!checkOneData<float>(a[i], b[i])) ++e;
  return e;
}

// Matrix-Market loader. Returns 0 ok / 1 failure. Arrays are copied out via a
// second call (ref_mtx_copy) so the caller can size buffers.
static sparseMatrix::CSR<float>* g_loaded = nullptr;
int ref_mtx_load(const char* path, uint32_t* M, uint32_t* N, uint32_t* nnz) {
  delete g_loaded;
  g_loaded = new sparseMatrix::CSR<float>();
  if (!g_loaded->initializeFromMatrixFile(path)) {
    delete g_loaded;
    g_loaded = nullptr;
    return 1;
  }
  *M = g_loaded->row();
  *N = g_loaded->col();
  *nnz = g_loaded->nnz();
  return 0;
}
void ref_mtx_copy(uint32_t* rowOff, uint32_t* colIdx, float* values) {
  std::memcpy(rowOff, g_loaded->rowOffsets().data(), sizeof(UIN) * (g_loaded->row() + 1));
  std::memcpy(colIdx, g_loaded->colIndices().data(), sizeof(UIN) * g_loaded->nnz());
  std::memcpy(values, g_loaded->values().data(), sizeof(float) * g_loaded->nnz());
}

}  // extern "C"

// src/BSMR.cpp also references the two GPU entry points of src/rowReordering.cu; they are never reached from the
// doors above (no GPU in the CPU suite) and only exist so that libref_cpu.so has no undefined symbols.
#include <cstdlib>
UIN calculateBlockSize(const sparseMatrix::CSR<float>&) { std::abort(); }
std::vector<UIN> bsa_rowReordering_gpu(const sparseMatrix::CSR<float>&, float, UIN, int&, float&) { std::abort(); }
