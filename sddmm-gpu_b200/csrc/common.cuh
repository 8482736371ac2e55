// common.cuh -- error handling, RAII device buffers, launch accounting.
#pragma once

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges show up in nsys / ncu --nvtx, cost nothing otherwise

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

#include "../../include/sddmm_b200.h"

namespace sb {

using u32 = uint32_t;
using u64 = uint64_t;

constexpr u32 kNull = SDDMM_NULL_VALUE;
constexpr u32 kPanel = SDDMM_ROW_PANEL;
constexpr u32 kBlockCols = SDDMM_BLOCK_COLS;

// ---- errors -------------------------------------------------------------------------------
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string& m);

[[noreturn]] inline void fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw Error(code, buf);
}

#define SB_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      ::sb::fail(e__ == cudaErrorMemoryAllocation ? SDDMM_E_NOMEM : SDDMM_E_CUDA,              \
                 "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

// ---- launch accounting (sddmm_launch_count) -----------------------------------------------
extern thread_local u64 g_launches;
#define SB_LAUNCH_CHECK()            \
  do {                               \
    ++::sb::g_launches;              \
    SB_CUDA(cudaGetLastError());     \
  } while (0)

// ---- RAII device buffer ---------------------------------------------------------------------
// Inside a TempScope, scratch allocations come from a per-thread arena: a few large cudaMalloc chunks
// (one cudaMalloc of 4 GB costs under a millisecond on this driver, while growing the stream-ordered
// pool by the same amount costs ~0.6 s and forty 100 MB cudaMallocs ~90 ms) handed out stack-fashion.
// Everything inside a scope runs on the scope's stream, so a block released by the host and handed out
// again is only touched by work enqueued later on that stream.  The arena is returned to the driver
// when the outermost scope ends (a chunk of <= 64 MB stays cached for small problems).  Buffers that
// outlive the scope -- the layout's own arrays -- are allocated with alloc(count, /*persistent=*/true).
struct TempState {
  bool active = false;
  cudaStream_t stream = nullptr;
};
extern thread_local TempState g_temp;
struct TempScope {
  TempState saved;
  explicit TempScope(cudaStream_t s, size_t hintBytes = 0);
  ~TempScope();
  TempScope(const TempScope&) = delete;
  TempScope& operator=(const TempScope&) = delete;
};
void* temp_alloc(size_t bytes);
void temp_free(void* p);

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  bool pooled = false;
  DevBuf() = default;
  explicit DevBuf(size_t count) { alloc(count); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n), pooled(o.pooled) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; pooled = o.pooled; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(size_t count, bool persistent = false) {
    release();
    n = count;
    if (!count) return;
    if (g_temp.active && !persistent) {
      p = static_cast<T*>(temp_alloc(count * sizeof(T)));
      pooled = true;
    } else {
      SB_CUDA(cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T)));
      pooled = false;
    }
  }
  void release() {
    if (p) {
      if (pooled) temp_free(p);
      else cudaFree(p);
    }
    p = nullptr;
    n = 0;
  }
  T* get() const { return p; }
  size_t size() const { return n; }
};

struct Timer {
  cudaEvent_t a = nullptr, b = nullptr;
  cudaStream_t s;
  explicit Timer(cudaStream_t st) : s(st) {
    SB_CUDA(cudaEventCreate(&a));
    SB_CUDA(cudaEventCreate(&b));
  }
  ~Timer() {
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
  }
  void start() { SB_CUDA(cudaEventRecord(a, s)); }
  float stop() {
    SB_CUDA(cudaEventRecord(b, s));
    SB_CUDA(cudaEventSynchronize(b));
    float ms = 0.f;
    SB_CUDA(cudaEventElapsedTime(&ms, a, b));
    return ms;
  }
};

// NVTX range over a stage of the path (reorder / layout / SDDMM pass), closed when the scope ends
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

inline u32 ceil_div(u32 a, u32 b) { return (a + b - 1) / b; }
inline u64 ceil_div64(u64 a, u64 b) { return (a + b - 1) / b; }

int device_sm_count();

}  // namespace sb
