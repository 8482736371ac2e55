// Matrix.cpp -- loaders / generators of the host data model (see Matrix.hpp).
// The .mtx loader reads the whole file once and parses numbers in place (the reference goes line by
// line through std::stod and a std::set, O(nnz log nnz) with large constants);
// duplicate detection is a sort of (row, col) keys.  Accept/reject behaviour and the resulting CSR
// (row-only stable order) are those of src/Matrix.cpp:398-480.
#include "Matrix.hpp"

#include <cuda_runtime_api.h>
#include <omp.h>
#include <parallel/algorithm>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <random>
#include <unordered_map>

#include "sddmm_b200.h"

template <typename T>
void Matrix<T>::makeData(uint64_t seed) {
  std::mt19937 gen(static_cast<uint32_t>(seed));
  std::uniform_real_distribution<float> dist(0.0f, 2.0f);
  for (size_t i = 0; i < values_.size(); ++i) values_[i] = static_cast<T>(dist(gen));
}
template class Matrix<float>;

namespace {

bool read_file(const std::string& path, std::vector<char>& buf) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  std::fseek(f, 0, SEEK_END);
  const long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  buf.resize(static_cast<size_t>(n) + 1);
  const size_t got = std::fread(buf.data(), 1, static_cast<size_t>(n), f);
  std::fclose(f);
  buf[got] = 0;
  buf.resize(got + 1);
  return true;
}

inline const char* skip_blank(const char* p) {
  while (*p == ' ' || *p == '\t' || *p == '\r') ++p;
  return p;
}
inline const char* next_line(const char* p) {
  while (*p && *p != '\n') ++p;
  return *p ? p + 1 : p;
}
inline bool at_eol(const char* p) { return *p == '\n' || *p == 0; }

// device CSR assembly (libsddmm_b200's sddmm_coo_to_csr) for files of >= 4 M entries when a CUDA device exists
bool use_device_loader(size_t nnz) {
  const char* e = std::getenv("SDDMM_B200_LOADER");
  if (e && !std::strcmp(e, "host")) return false;
  int n = 0;
  const bool haveDev = cudaGetDeviceCount(&n) == cudaSuccess && n > 0;
  if (e && !std::strcmp(e, "device")) return true;  // forced: fails loudly without a device
  return haveDev && nnz >= (static_cast<size_t>(1) << 22);
}

std::string suffix_of(const std::string& f) {
  const size_t dot = f.find_last_of('.');
  return dot == std::string::npos ? "" : f.substr(dot);
}

}  // namespace

namespace sparseMatrix {

template <typename T>
bool CSR<T>::initializeFromMatrixFile(const std::string& file) {
  const std::string s = suffix_of(file);
  if (s == ".mtx" || s == ".mmio") return initializeFromMtxFile(file);
  if (s == ".smtx") return initializeFromSmtxFile(file);
  if (s == ".txt") return initializeFromGraphDataset(file);
  std::cerr << "Error, file format is not supported : " << file << std::endl;
  return false;
}

template <typename T>
bool CSR<T>::initializeFromMtxFile(const std::string& file) {
  std::vector<char> buf;
  if (!read_file(file, buf)) {
    std::cerr << "Error, file cannot be opened : " << file << std::endl;
    return false;
  }
  std::cout << "sparseMatrix::CSR initialize from file : " << file << std::endl;
  const char* p = buf.data();
  while (*p == '%') p = next_line(p);  // skip comments
  char* e = nullptr;
  p = skip_blank(p);
  row_ = static_cast<UIN>(std::strtol(p, &e, 10)); p = skip_blank(e);
  col_ = static_cast<UIN>(std::strtol(p, &e, 10)); p = skip_blank(e);
  nnz_ = at_eol(p) ? 0u : static_cast<UIN>(std::strtod(p, &e));
  p = next_line(p);
  // ---- parallel parse: the body is cut into one slab per thread at line boundaries; every slab is parsed
  // into its own vectors and the slabs are concatenated in file order (the order inside a row is part of
  // the contract, src/Matrix.cpp:467).
  const char* body = p;
  const char* fileEnd = buf.data() + buf.size() - 1;  // points at the terminating 0
  const int nThreads = std::max(1, omp_get_max_threads());
  std::vector<const char*> cut(nThreads + 1, fileEnd);
  cut[0] = body;
  for (int t = 1; t < nThreads; ++t) {
    const char* q = body + (fileEnd - body) / nThreads * t;
    while (q < fileEnd && *q != '\n') ++q;
    cut[t] = q < fileEnd ? q + 1 : fileEnd;
  }
  for (int t = 1; t <= nThreads; ++t)
    if (cut[t] < cut[t - 1]) cut[t] = cut[t - 1];
  std::vector<std::vector<UIN>> tri(nThreads), tci(nThreads);
  std::vector<std::vector<T>> tva(nThreads);
#pragma omp parallel for schedule(static, 1)
  for (int t = 0; t < nThreads; ++t) {
    const char* q = cut[t];
    const char* qe = cut[t + 1];
    char* e2 = nullptr;
    while (q < qe && *q) {
      const char* w = skip_blank(q);
      if (at_eol(w)) { q = next_line(q); continue; }  // empty line
      const long r = std::strtol(w, &e2, 10); w = skip_blank(e2);
      const long c = std::strtol(w, &e2, 10); w = skip_blank(e2);
      T v = static_cast<T>(0);
      if (!at_eol(w)) v = static_cast<T>(std::strtod(w, &e2));
      tri[t].push_back(static_cast<UIN>(r - 1));
      tci[t].push_back(static_cast<UIN>(c - 1));
      tva[t].push_back(v);
      q = next_line(q);
    }
  }
  size_t total = 0;
  for (int t = 0; t < nThreads; ++t) total += tri[t].size();
  if (total > nnz_) {
    std::cerr << "Error, file " << file << " too many elements, exceeding the number nnz!" << std::endl;
    return false;
  }
  if (total < nnz_) {
    std::cerr << "Error, file " << file << " elements is not enough!" << std::endl;
    return false;
  }
  std::vector<UIN> ri(nnz_), ci(nnz_);
  std::vector<T> va(nnz_);
  {
    size_t o = 0;
    for (int t = 0; t < nThreads; ++t) {
      std::copy(tri[t].begin(), tri[t].end(), ri.begin() + o);
      std::copy(tci[t].begin(), tci[t].end(), ci.begin() + o);
      std::copy(tva[t].begin(), tva[t].end(), va.begin() + o);
      o += tri[t].size();
    }
  }
  bool tooBig = false;
#pragma omp parallel for reduction(|| : tooBig)
  for (long i = 0; i < static_cast<long>(nnz_); ++i) tooBig = tooBig || ri[i] >= row_ || ci[i] >= col_;
  if (tooBig) {
    std::cerr << "Error, file " << file << " row or col is too big!" << std::endl;
    return false;
  }
  // CSR assembly: on the device (sddmm_coo_to_csr: duplicate check = key sort, stable radix sort by row, histogram +
  // scan) for big files when a GPU is there, else the host stages below.  SDDMM_B200_LOADER=device|host forces one.
  if (use_device_loader(nnz_)) {
    int dup = 0;
    rowOffsets_.assign(static_cast<size_t>(row_) + 1, 0);
    colIndices_.resize(nnz_);
    std::vector<float> vin(va.begin(), va.end()), vout(nnz_);
    const int rc = sddmm_coo_to_csr(ri.data(), ci.data(), vin.data(), nnz_, row_, col_, rowOffsets_.data(),
                                    colIndices_.data(), vout.data(), &dup);
    if (rc != SDDMM_OK) {
      std::cerr << "Error, device CSR build failed: " << sddmm_last_error() << std::endl;
      return false;
    }
    if (dup) {
      std::cerr << "Error, matrix has duplicate data!" << std::endl;
      return false;
    }
    if (nnz_ <= 1) {
      std::cerr << "Warning, file " << file << " nnz is 1, this is not a valid matrix!" << std::endl;
      return false;
    }
    values_.assign(vout.begin(), vout.end());
    return true;
  }
  std::vector<uint64_t> keys(nnz_);
#pragma omp parallel for
  for (long i = 0; i < static_cast<long>(nnz_); ++i) keys[i] = (static_cast<uint64_t>(ri[i]) << 32) | ci[i];
  __gnu_parallel::sort(keys.begin(), keys.end());
  if (std::adjacent_find(keys.begin(), keys.end()) != keys.end()) {
    std::cerr << "Error, matrix has duplicate data!" << std::endl;
    return false;
  }
  if (nnz_ <= 1) {
    std::cerr << "Warning, file " << file << " nnz is 1, this is not a valid matrix!" << std::endl;
    return false;
  }
  // stable by row only: columns keep FILE order inside a row (src/Matrix.cpp:467)
  rowOffsets_.assign(static_cast<size_t>(row_) + 1, 0);
  for (UIN i = 0; i < nnz_; ++i) rowOffsets_[ri[i] + 1]++;
  for (UIN r = 0; r < row_; ++r) rowOffsets_[r + 1] += rowOffsets_[r];
  std::vector<UIN> pos(rowOffsets_.begin(), rowOffsets_.end() - 1);
  colIndices_.resize(nnz_);
  values_.resize(nnz_);
  for (UIN i = 0; i < nnz_; ++i) {
    const UIN d = pos[ri[i]]++;
    colIndices_[d] = ci[i];
    values_[d] = va[i];
  }
  return true;
}

template <typename T>
bool CSR<T>::initializeFromSmtxFile(const std::string& file) {
  std::vector<char> buf;
  if (!read_file(file, buf)) {
    std::cerr << "Error, file cannot be opened : " << file << std::endl;
    return false;
  }
  const char* p = buf.data();
  while (*p == '%') p = next_line(p);
  auto next_int = [&](const char*& q) -> long {
    while (*q == ' ' || *q == '\t' || *q == '\r' || *q == ',') ++q;
    char* e = nullptr;
    const long v = std::strtol(q, &e, 10);
    q = e;
    return v;
  };
  row_ = static_cast<UIN>(next_int(p));
  col_ = static_cast<UIN>(next_int(p));
  nnz_ = static_cast<UIN>(next_int(p));
  if (nnz_ == 0) {
    std::cerr << "Error, file " << file << " nnz is 0!" << std::endl;
    return false;
  }
  p = next_line(p);
  rowOffsets_.resize(static_cast<size_t>(row_) + 1);
  for (size_t i = 0; i < rowOffsets_.size(); ++i) rowOffsets_[i] = static_cast<UIN>(next_int(p));
  p = next_line(p);
  colIndices_.resize(nnz_);
  values_.assign(nnz_, static_cast<T>(1));
  for (UIN i = 0; i < nnz_; ++i) colIndices_[i] = static_cast<UIN>(next_int(p));
  for (UIN r = 0; r < row_; ++r) {  // duplicate check per row
    std::vector<UIN> c(colIndices_.begin() + rowOffsets_[r], colIndices_.begin() + rowOffsets_[r + 1]);
    std::sort(c.begin(), c.end());
    if (std::adjacent_find(c.begin(), c.end()) != c.end()) {
      std::cerr << "Error, matrix has duplicate data!" << std::endl;
      return false;
    }
  }
  return true;
}

// SNAP edge list (src/Matrix.cpp:483-580): '#' header lines carry "Nodes: n" and "Edges: e"; every other line is
// "from to [value]"; node ids are renumbered in order of first appearance (from before to); duplicates and
// counts that disagree with the header are errors; rows end up stably sorted, columns in file order.
template <typename T>
bool CSR<T>::initializeFromGraphDataset(const std::string& file) {
  std::vector<char> buf;
  if (!read_file(file, buf)) {
    std::cerr << "Error, file cannot be opened : " << file << std::endl;
    return false;
  }
  std::cout << "sparseMatrix::CSR initialize from file : " << file << std::endl;
  const char* p = buf.data();
  row_ = col_ = nnz_ = 0;
  while (*p == '#') {
    const char* eol = p;
    while (*eol && *eol != '\n') ++eol;
    const std::string line(p, eol);
    size_t at;
    if ((at = line.find("Nodes: ")) != std::string::npos) row_ = col_ = static_cast<UIN>(std::strtol(line.c_str() + at + 7, nullptr, 10));
    if ((at = line.find("Edges: ")) != std::string::npos) nnz_ = static_cast<UIN>(std::strtol(line.c_str() + at + 7, nullptr, 10));
    p = next_line(p);
  }
  if (!row_ || !col_ || !nnz_) {
    std::cerr << "Error, file " << file << " row or col or nnz not initialized!" << std::endl;
    return false;
  }
  std::vector<UIN> ri, ci;
  std::vector<T> va;
  ri.reserve(nnz_); ci.reserve(nnz_); va.reserve(nnz_);
  std::unordered_map<UIN, UIN> id;
  id.reserve(static_cast<size_t>(row_) * 2);
  auto renumber = [&](UIN node) {
    auto it = id.find(node);
    if (it != id.end()) return it->second;
    const UIN n = static_cast<UIN>(id.size());
    id.emplace(node, n);
    return n;
  };
  char* e = nullptr;
  while (*p) {
    if (at_eol(p)) { p = next_line(p); continue; }  // empty line
    const char* w = p;
    const UIN a = static_cast<UIN>(std::strtol(w, &e, 10)); w = skip_blank(e);
    const UIN b = static_cast<UIN>(std::strtol(w, &e, 10)); w = skip_blank(e);
    const T v = at_eol(w) ? static_cast<T>(0) : static_cast<T>(std::strtod(w, &e));
    const UIN ra = renumber(a), rb = renumber(b);
    if (ri.size() >= nnz_) {
      std::cerr << "Error, file " << file << " too many elements, exceeding the number nnz!" << std::endl;
      return false;
    }
    ri.push_back(ra); ci.push_back(rb); va.push_back(v);
    p = next_line(p);
  }
  if (ri.size() < nnz_) {
    std::cerr << "Error, file " << file << " elements is not enough!" << std::endl;
    return false;
  }
  std::vector<uint64_t> keys(nnz_);
  for (UIN i = 0; i < nnz_; ++i) {
    if (ri[i] >= row_ || ci[i] >= col_) {
      std::cerr << "Error, file " << file << " row or col is too big!" << std::endl;
      return false;
    }
    keys[i] = (static_cast<uint64_t>(ri[i]) << 32) | ci[i];
  }
  __gnu_parallel::sort(keys.begin(), keys.end());
  if (std::adjacent_find(keys.begin(), keys.end()) != keys.end()) {
    std::cerr << "Error, matrix has duplicate data!" << std::endl;
    return false;
  }
  rowOffsets_.assign(static_cast<size_t>(row_) + 1, 0);
  for (UIN i = 0; i < nnz_; ++i) rowOffsets_[ri[i] + 1]++;
  for (UIN r = 0; r < row_; ++r) rowOffsets_[r + 1] += rowOffsets_[r];
  std::vector<UIN> pos(rowOffsets_.begin(), rowOffsets_.end() - 1);
  colIndices_.resize(nnz_);
  values_.resize(nnz_);
  for (UIN i = 0; i < nnz_; ++i) {
    const UIN d = pos[ri[i]]++;
    colIndices_[d] = ci[i];
    values_[d] = va[i];
  }
  return true;
}

template <typename T>
bool CSR<T>::outputToMarketMatrixFile(const std::string& fileName) const {
  std::ofstream out(fileName);
  if (!out) return false;
  out << "%%MatrixMarket matrix coordinate real general\n" << row_ << " " << col_ << " " << nnz_ << "\n";
  for (UIN r = 0; r < row_; ++r)
    for (UIN i = rowOffsets_[r]; i < rowOffsets_[r + 1]; ++i)
      out << (r + 1) << " " << (colIndices_[i] + 1) << " " << values_[i] << "\n";
  return true;
}

template class CSR<float>;

}  // namespace sparseMatrix
