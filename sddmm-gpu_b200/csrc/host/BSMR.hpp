// BSMR.hpp -- BSMR / RPHM with the public surface of the reference's include/BSMR.hpp (:21-63, :79-159).
// Host vectors are filled from the device layout that libsddmm_b200 builds; all compute is in the library.
#pragma once
#include <memory>
#include <vector>

#include "Matrix.hpp"
#include "sddmm_b200.h"

constexpr UIN ROW_PANEL_SIZE = SDDMM_ROW_PANEL;
constexpr UIN BLOCK_COL_SIZE = SDDMM_BLOCK_COLS;
constexpr UIN BLOCK_SIZE = ROW_PANEL_SIZE * BLOCK_COL_SIZE;

using LayoutPtr = std::shared_ptr<bsmr_layout>;

UIN calculateBlockSize(const sparseMatrix::CSR<float>& matrix, uint64_t freeMem = 0);  // rowReordering.cu:1009

class BSMR {
 public:
  BSMR() = default;
  BSMR(float similarityThreshold, float blockDensityThreshold, const sparseMatrix::CSR<float>& matrix,
       int numIterations = 1, UIN blockSize = 0);
  void rowReordering(float similarityThreshold, const sparseMatrix::CSR<float>& matrix, int numIterations = 1,
                     UIN blockSize = 0);
  void colReordering(float blockDensityThreshold, const sparseMatrix::CSR<float>& matrix,
                     const std::vector<UIN>& reorderedRows = std::vector<UIN>(), int numIterations = 1);

  int numRowPanels() const { return numRowPanels_; }
  const std::vector<UIN>& reorderedRows() const { return reorderedRows_; }
  const std::vector<UIN>& denseCols() const { return denseCols_; }
  const std::vector<UIN>& denseColOffsets() const { return denseColOffsets_; }
  const std::vector<UIN>& sparseCols() const { return sparseCols_; }
  const std::vector<UIN>& sparseColOffsets() const { return sparseColOffsets_; }
  const std::vector<UIN>& sparseValueOffsets() const { return sparseValueOffsets_; }
  int numClusters() const { return numClusters_; }
  float rowReorderingTime() const { return rowReorderingTime_; }
  float colReorderingTime() const { return colReorderingTime_; }
  float reorderingTime() const { return rowReorderingTime_ + colReorderingTime_; }
  float rphmTime() const { return rphmTime_; }
  UIN blockSize() const { return blockSize_; }
  const LayoutPtr& layout() const { return layout_; }

 private:
  int numRowPanels_ = 0;
  std::vector<UIN> reorderedRows_, denseCols_, denseColOffsets_, sparseCols_, sparseColOffsets_, sparseValueOffsets_;
  int numClusters_ = 1;
  float rowReorderingTime_ = 0.f, colReorderingTime_ = 0.f, rphmTime_ = 0.f;
  UIN blockSize_ = 0;
  LayoutPtr layout_;
};

// The reference keeps the RPHM arrays in dev::vector members; here they stay inside the layout object
// and are exposed as device pointers (+ host copies on request).
class RPHM {
 public:
  RPHM() = default;
  RPHM(const sparseMatrix::CSR<float>& matrix, const BSMR& bsmr);
  UIN numRowPanels() const { return info_.numRowPanels; }
  UIN maxNumDenseColBlocksInRowPanel() const { return info_.maxNumDenseColBlocksInRowPanel; }
  UIN maxNumSparseColBlocksInRowPanel() const { return info_.maxNumSparseColBlocksInRowPanel; }
  UIN numDenseThreadBlocks() const { return info_.numDenseThreadBlocks; }
  UIN numSparseThreadBlocks() const { return info_.numSparseThreadBlocks; }
  UIN getNumDenseBlocks() const { return info_.numDenseBlocks; }
  UIN numDenseValues() const { return info_.numDenseValues; }
  UIN numSparseValues() const { return info_.numSparseValues; }
  const UIN* devicePtr(bsmr_array_id id) const { return bsmr_layout_array_dev(layout_.get(), id); }
  std::vector<UIN> hostCopy(bsmr_array_id id) const;
  std::vector<UIN> blockValues() const { return hostCopy(RPHM_BLOCK_VALUES); }
  std::vector<UIN> blockOffsets() const { return hostCopy(RPHM_BLOCK_OFFSETS); }
  std::vector<UIN> sparseValues() const { return hostCopy(RPHM_SPARSE_VALUES); }
  std::vector<UIN> sparseRelativeRows() const { return hostCopy(RPHM_SPARSE_RELATIVE_ROWS); }
  std::vector<UIN> sparseColIndices() const { return hostCopy(RPHM_SPARSE_COL_INDICES); }
  const LayoutPtr& layout() const { return layout_; }

 private:
  LayoutPtr layout_;
  bsmr_layout_info info_{};
};

struct Logger;
// evaluationReordering(matrix, bsmr, logger)  <- src/BSMR.cpp:826-925 (+ :953-994): fills the logger's
// numDenseBlock_/averageDensity_/thread-block and data counts and the original-matrix statistics.
// The layout of `rphm` is what the BSMR object describes; the counting runs on the device.
void evaluationReordering(const sparseMatrix::CSR<float>& matrix, const RPHM& rphm, Logger& logger);
