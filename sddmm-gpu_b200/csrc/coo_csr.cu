// coo_csr.cu -- CSR assembly on the device for the file loaders.
// Replaces, for the triplets a loader has parsed, the three host stages of the reference's
//   sparseMatrix::CSR<float>::initializeFromMtxFile   src/Matrix.cpp:398-480
//     * duplicate check through a std::set                            :447-461  -> one 64-bit key sort + adjacent compare
//     * thrust::host stable sort BY ROW ONLY (file order inside a row) :467      -> stable LSD radix sort, payload = position
//     * rowOffsets from the sorted rows                               :470-479  -> histogram + exclusive scan
// with the reference's accept / reject rules left to the caller (range check, nnz <= 1, header count).
#include "primitives.cuh"

namespace sb {
namespace {
__global__ void k_coo_keys(const u32* __restrict__ rows, const u32* __restrict__ cols, size_t n, int colBits,
                           u64* __restrict__ keys) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    keys[i] = ((u64)rows[i] << colBits) | cols[i];
}
__global__ void k_adjacent_equal(const u64* __restrict__ keys, size_t n, u32* __restrict__ flag) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i + 1 < n; i += (size_t)gridDim.x * blockDim.x)
    if (keys[i] == keys[i + 1]) *flag = 1u;
}
__global__ void k_row_hist(const u32* __restrict__ rows, size_t n, u32* __restrict__ cnt) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    atomicAdd(cnt + rows[i], 1u);
}
__global__ void k_gather_entries(const u32* __restrict__ perm, const u32* __restrict__ cols, const float* __restrict__ vals,
                                 size_t n, u32* __restrict__ colIdx, float* __restrict__ values) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const u32 s = perm[i];
    colIdx[i] = cols[s];
    if (values) values[i] = vals ? vals[s] : 0.f;
  }
}
}  // namespace

// d_rows / d_cols / d_vals: n triplets in file order (0-based, already range-checked).  Outputs on the device.
// Returns true when some (row, col) appears twice (outputs are then unspecified).
bool coo_to_csr_dev(const u32* d_rows, const u32* d_cols, const float* d_vals, u32 n, u32 M, u32 N, u32* d_rowOff,
                    u32* d_colIdx, float* d_values, cudaStream_t s) {
  TempScope scope(s);
  SB_CUDA(cudaMemsetAsync(d_rowOff, 0, ((size_t)M + 1) * 4, s));
  if (n == 0) {
    SB_CUDA(cudaStreamSynchronize(s));
    return false;
  }
  // ---- duplicates
  u32 dup = 0;
  {
    DevBuf<u64> ka(n), kb(n);
    DevBuf<u32> flag(1);
    SB_CUDA(cudaMemsetAsync(flag.get(), 0, 4, s));
    const int colBits = bits_for(N ? N - 1 : 0), rowBits = bits_for(M ? M - 1 : 0);
    k_coo_keys<<<grid_for(n), 256, 0, s>>>(d_rows, d_cols, n, colBits, ka.get());
    SB_LAUNCH_CHECK();
    const int w = radix_sort_pairs<u64>(ka.get(), kb.get(), nullptr, nullptr, n, 0, colBits + rowBits, s);
    k_adjacent_equal<<<grid_for(n), 256, 0, s>>>(w ? kb.get() : ka.get(), n, flag.get());
    SB_LAUNCH_CHECK();
    SB_CUDA(cudaMemcpyAsync(&dup, flag.get(), 4, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaStreamSynchronize(s));
  }
  if (dup) return true;
  // ---- stable by row only
  DevBuf<u32> ra(n), rb(n), pa(n), pb(n);
  SB_CUDA(cudaMemcpyAsync(ra.get(), d_rows, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
  iota<u32>(pa.get(), n, 0u, s);
  const int w = radix_sort_pairs<u32>(ra.get(), rb.get(), pa.get(), pb.get(), n, 0, bits_for(M ? M - 1 : 0), s);
  k_gather_entries<<<grid_for(n), 256, 0, s>>>(w ? pb.get() : pa.get(), d_cols, d_vals, n, d_colIdx, d_values);
  SB_LAUNCH_CHECK();
  // ---- row offsets
  k_row_hist<<<grid_for(n), 256, 0, s>>>(d_rows, n, d_rowOff);
  SB_LAUNCH_CHECK();
  exclusive_scan_u32(d_rowOff, d_rowOff, (size_t)M + 1, s);
  SB_CUDA(cudaStreamSynchronize(s));
  return false;
}

}  // namespace sb

using namespace sb;

extern "C" int sddmm_coo_to_csr(const uint32_t* h_rows, const uint32_t* h_cols, const float* h_vals, uint32_t nnz,
                                uint32_t M, uint32_t N, uint32_t* h_rowOff, uint32_t* h_colIdx, float* h_values,
                                int* hasDuplicate) {
  try {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
      fail(SDDMM_E_CUDA, "no CUDA device available: libsddmm_b200 has no CPU fallback");
    if (!h_rowOff || (nnz && (!h_rows || !h_cols || !h_colIdx)) || !hasDuplicate) fail(SDDMM_E_ARG, "invalid argument: null pointer");
    *hasDuplicate = 0;
    DevBuf<u32> dr(nnz ? nnz : 1), dc(nnz ? nnz : 1), ro((size_t)M + 1), ci(nnz ? nnz : 1);
    DevBuf<float> dv, vo;
    if (h_vals && h_values) { dv.alloc(nnz ? nnz : 1); vo.alloc(nnz ? nnz : 1); }
    SB_CUDA(cudaMemcpy(dr.get(), h_rows, (size_t)nnz * 4, cudaMemcpyHostToDevice));
    SB_CUDA(cudaMemcpy(dc.get(), h_cols, (size_t)nnz * 4, cudaMemcpyHostToDevice));
    if (dv.get()) SB_CUDA(cudaMemcpy(dv.get(), h_vals, (size_t)nnz * 4, cudaMemcpyHostToDevice));
    if (coo_to_csr_dev(dr.get(), dc.get(), dv.get(), nnz, M, N, ro.get(), ci.get(), vo.get(), nullptr)) {
      *hasDuplicate = 1;
      return SDDMM_OK;
    }
    SB_CUDA(cudaMemcpy(h_rowOff, ro.get(), ((size_t)M + 1) * 4, cudaMemcpyDeviceToHost));
    SB_CUDA(cudaMemcpy(h_colIdx, ci.get(), (size_t)nnz * 4, cudaMemcpyDeviceToHost));
    if (vo.get()) SB_CUDA(cudaMemcpy(h_values, vo.get(), (size_t)nnz * 4, cudaMemcpyDeviceToHost));
    else if (h_values) std::memset(h_values, 0, (size_t)nnz * 4);
  } catch (const sb::Error& e) {
    sb::set_last_error(e.what());
    return e.code;
  } catch (const std::exception& e) {
    sb::set_last_error(e.what());
    return SDDMM_E_CUDA;
  }
  return SDDMM_OK;
}
