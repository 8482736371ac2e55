// Matrix.hpp -- host data model with the surface of the reference's include/Matrix.hpp
// (Matrix<T> :39-164, sparseMatrix::CSR<T> :195-296) as far as the SDDMM path uses it.
// Own implementation; the Matrix Market loader follows the behaviour of src/Matrix.cpp:398-480
// (skip '%' lines, "M N nnz", 1-based triplets, value optional -> 0, duplicates / out-of-range /
// nnz <= 1 rejected, stable sort by ROW ONLY so columns keep file order inside a row).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

using UIN = uint32_t;
constexpr UIN NULL_VALUE = 0xFFFFFFFFu;

enum MatrixStorageOrder { row_major, col_major };

template <typename T>
class Matrix {
 public:
  Matrix() = delete;
  Matrix(UIN row, UIN col, MatrixStorageOrder order)
      : row_(row), col_(col), storageOrder_(order), leadingDimension_(order == row_major ? col : row),
        values_(static_cast<size_t>(row) * col) {}
  Matrix(UIN row, UIN col, MatrixStorageOrder order, const std::vector<T>& values)
      : row_(row), col_(col), storageOrder_(order), leadingDimension_(order == row_major ? col : row), values_(values) {}
  Matrix(UIN row, UIN col, MatrixStorageOrder order, const T* values)
      : row_(row), col_(col), storageOrder_(order), leadingDimension_(order == row_major ? col : row),
        values_(values, values + static_cast<size_t>(row) * col) {}

  // uniform [0,2) like src/Matrix.cpp:117-138, but seeded and race-free (the reference shares one
  // std::mt19937 across OpenMP threads, so its data differ run to run)
  void makeData(uint64_t seed = 5489u);

  UIN row() const { return row_; }
  UIN col() const { return col_; }
  UIN leadingDimension() const { return leadingDimension_; }
  size_t size() const { return values_.size(); }
  MatrixStorageOrder storageOrder() const { return storageOrder_; }
  const std::vector<T>& values() const { return values_; }
  const T* data() const { return values_.data(); }
  T& operator[](size_t i) { return values_[i]; }
  const T& operator[](size_t i) const { return values_[i]; }

 private:
  UIN row_, col_;
  MatrixStorageOrder storageOrder_;
  UIN leadingDimension_;
  std::vector<T> values_;
};

namespace sparseMatrix {

class DataBase {
 public:
  UIN row() const { return row_; }
  UIN col() const { return col_; }
  UIN nnz() const { return nnz_; }
  float getSparsity() const { return 1.0f - static_cast<float>(nnz_) / (static_cast<float>(row_) * col_); }

 protected:
  UIN row_ = 0, col_ = 0, nnz_ = 0;
};

template <typename T>
class CSR : public DataBase {
 public:
  CSR() = default;
  CSR(UIN row, UIN col, UIN nnz, const std::vector<UIN>& rowOffsets, const std::vector<UIN>& colIndices,
      const std::vector<T>& values)
      : rowOffsets_(rowOffsets), colIndices_(colIndices), values_(values) { row_ = row; col_ = col; nnz_ = nnz; }
  CSR(UIN row, UIN col, UIN nnz, const UIN* rowOffsets, const UIN* colIndices, const T* values)
      : rowOffsets_(rowOffsets, rowOffsets + row + 1), colIndices_(colIndices, colIndices + nnz),
        values_(values, values + nnz) { row_ = row; col_ = col; nnz_ = nnz; }
  CSR(UIN row, UIN col, UIN nnz, const std::vector<UIN>& rowOffsets, const std::vector<UIN>& colIndices)
      : rowOffsets_(rowOffsets), colIndices_(colIndices), values_(nnz, 0) { row_ = row; col_ = col; nnz_ = nnz; }

  bool initializeFromMatrixFile(const std::string& file);  // suffix dispatch, src/Matrix.cpp:280-294
  bool initializeFromMtxFile(const std::string& file);     // src/Matrix.cpp:398-480
  bool initializeFromSmtxFile(const std::string& file);    // src/Matrix.cpp:296-371 (DLMC native format)
  bool initializeFromGraphDataset(const std::string& file);  // SNAP-style .txt edge list, src/Matrix.cpp:483-580
  bool outputToMarketMatrixFile(const std::string& fileName) const;

  const std::vector<UIN>& rowOffsets() const { return rowOffsets_; }
  const std::vector<UIN>& colIndices() const { return colIndices_; }
  const std::vector<T>& values() const { return values_; }
  std::vector<T>& setValues() { return values_; }

 private:
  std::vector<UIN> rowOffsets_, colIndices_;
  std::vector<T> values_;
};

}  // namespace sparseMatrix
