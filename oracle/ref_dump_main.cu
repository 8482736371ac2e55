// oracle/ref_dump_main.cu -- TEST INFRASTRUCTURE.
// A dump-only driver for the UNMODIFIED reference (objects compiled from
// /root/reference/src by oracle/Makefile into oracle/_ref/).  It runs the
// reference's own  BSMR -> RPHM -> sddmm_gpu  pipeline (src/sddmm.cu:10-39) on a
// case file written by tests/cases.py and writes every array the parity tests
// compare, as raw little-endian files, plus the reference's timings.
//
//   ref_dump <case.bin> <alpha> <delta> <outdir> [numIter=10]
//
// case.bin: u32 magic 'SDM1', M, N, nnz, K; rowOff[M+1]; colIdx[nnz]; A[M*K] f32
//           (row-major); B[N*K] f32 (column-major KxN, i.e. B^T rows).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "BSMR.hpp"
#include "Logger.hpp"
#include "Matrix.hpp"
#include "checkData.hpp"
#include "host.hpp"
#include "sddmm.hpp"
#include "sddmmKernel.cuh"

template <typename T>
static void dump(const std::string& dir, const char* name, const std::vector<T>& v) {
  const std::string p = dir + "/" + name;
  FILE* f = fopen(p.c_str(), "wb");
  if (!f) { fprintf(stderr, "cannot write %s\n", p.c_str()); exit(3); }
  if (!v.empty()) fwrite(v.data(), sizeof(T), v.size(), f);
  fclose(f);
}

int main(int argc, char** argv) {
  if (argc < 5) {
    fprintf(stderr, "usage: ref_dump case.bin alpha delta outdir [numIter]\n");
    return 2;
  }
  const float alpha = std::stof(argv[2]);
  const float delta = std::stof(argv[3]);
  const std::string out = argv[4];
  const int numIter = argc > 5 ? atoi(argv[5]) : 10;

  FILE* f = fopen(argv[1], "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
  uint32_t hdr[5];
  if (fread(hdr, 4, 5, f) != 5 || hdr[0] != 0x314D4453u) { fprintf(stderr, "bad case file\n"); return 2; }
  const uint32_t M = hdr[1], N = hdr[2], nnz = hdr[3], K = hdr[4];
  std::vector<UIN> rowOff(M + 1), colIdx(nnz);
  std::vector<float> A((size_t)M * K), B((size_t)N * K), vals(nnz, 1.0f);
  if (fread(rowOff.data(), 4, M + 1, f) != M + 1 || fread(colIdx.data(), 4, nnz, f) != nnz ||
      fread(A.data(), 4, A.size(), f) != A.size() || fread(B.data(), 4, B.size(), f) != B.size()) {
    fprintf(stderr, "short case file\n");
    return 2;
  }
  fclose(f);

  sparseMatrix::CSR<float> S(M, N, nnz, rowOff, colIdx, vals);
  Matrix<float> mA(M, K, MatrixStorageOrder::row_major, A);
  Matrix<float> mB(K, N, MatrixStorageOrder::col_major, B);

  size_t freeMem = 0, totalMem = 0;
  cudaMemGetInfo(&freeMem, &totalMem);
  const UIN blockSize = calculateBlockSize(S);

  // src/sddmm.cu:10-39, spelled out so the intermediates can be dumped
  BSMR bsmr(alpha, delta, S, 1);
  RPHM rphm(S, bsmr);
  // layout arrays first: a faulting SDDMM kernel poisons the context (on sm_100 the reference's
  // K>32 residual kernel dies with cudaErrorIllegalInstruction: its __shfl_xor_sync mask
  // `(1 << tId) | (1 << (tId ^ 1))` (src/sddmmKernel.cu:2096) is 0 for threadIdx.x >= 32)
  dump(out, "reorderedRows.u32", bsmr.reorderedRows());
  dump(out, "denseCols.u32", bsmr.denseCols());
  dump(out, "denseColOffsets.u32", bsmr.denseColOffsets());
  dump(out, "sparseCols.u32", bsmr.sparseCols());
  dump(out, "sparseColOffsets.u32", bsmr.sparseColOffsets());
  dump(out, "sparseValueOffsets.u32", bsmr.sparseValueOffsets());
  dump(out, "blockOffsets.u32", d2h(rphm.blockOffsets()));
  dump(out, "blockValues.u32", d2h(rphm.blockValues()));
  dump(out, "sparseValues.u32", d2h(rphm.sparseValues()));
  dump(out, "sparseRelativeRows.u32", d2h(rphm.sparseRelativeRows()));
  dump(out, "sparseColIndices.u32", d2h(rphm.sparseColIndices()));
  dump(out, "denseRowPanelIds.u32", d2h(rphm.denseRowPanelIds()));
  dump(out, "denseColBlockIters.u32", d2h(rphm.denseColBlockIters()));
  dump(out, "sparseRowPanelIds.u32", d2h(rphm.sparseRowPanelIds()));
  dump(out, "sparseColBlockIters.u32", d2h(rphm.sparseColBlockIters()));
  const UIN numDenseBlocks = rphm.getNumDenseBlocks();

  Logger logger;
  logger.numITER_ = numIter;
  sparseMatrix::CSR<float> P(S);
  sddmm_gpu(mA, mB, rphm, P, logger);
  cudaDeviceSynchronize();
  const cudaError_t err = cudaGetLastError();
  dump(out, "P.f32", P.values());

  // the reference's own oracle + tolerance (src/sddmm.cu:41-59)
  sparseMatrix::CSR<float> Pcpu(S);
  sddmm_cpu(mA, mB, S, Pcpu);
  size_t numErr = 0;
  for (size_t i = 0; i < nnz; ++i)
    if (!checkOneData<float>(Pcpu.values()[i], P.values()[i])) ++numErr;
  dump(out, "P_cpu.f32", Pcpu.values());

  const std::string metaPath = out + "/meta.txt";
  FILE* m = fopen(metaPath.c_str(), "w");
  fprintf(m, "M %u\nN %u\nnnz %u\nK %u\nalpha %.9g\ndelta %.9g\n", M, N, nnz, K, alpha, delta);
  fprintf(m, "block_size %u\nfree_mem %zu\nnum_row_panels %d\nnum_clusters %d\n", blockSize, freeMem,
          bsmr.numRowPanels(), bsmr.numClusters());
  fprintf(m, "row_reorder_ms %.6f\ncol_reorder_ms %.6f\nsddmm_ms %.6f\n", bsmr.rowReorderingTime(),
          bsmr.colReorderingTime(), logger.sddmmTime_);
  fprintf(m, "gflops %.6f\n", 2.0 * nnz * K / (logger.sddmmTime_ * 1e6));
  fprintf(m, "num_dense_blocks %u\nnum_sparse_values %u\n", numDenseBlocks,
          bsmr.sparseValueOffsets().empty() ? 0u : bsmr.sparseValueOffsets().back());
  fprintf(m, "max_dense_blocks %u\nnum_dense_tb %u\nnum_sparse_tb %u\nmax_sparse_tb %u\n",
          rphm.maxNumDenseColBlocksInRowPanel(), rphm.numDenseThreadBlocks(),
          rphm.numSparseThreadBlocks(), rphm.maxNumSparseColBlocksInRowPanel());
  fprintf(m, "check_errors %zu\ncuda_error %d\n", numErr, (int)err);
  fclose(m);
  printf("ref_dump ok: M=%u N=%u nnz=%u K=%u bs=%u panels=%d rowms=%.3f colms=%.3f sddmm_ms=%.4f gflops=%.1f errs=%zu cuda=%d\n",
         M, N, nnz, K, blockSize, bsmr.numRowPanels(), bsmr.rowReorderingTime(),
         bsmr.colReorderingTime(), logger.sddmmTime_, 2.0 * nnz * K / (logger.sddmmTime_ * 1e6),
         numErr, (int)err);
  return 0;
}
