// common.cuh -- error handling, RAII device buffers, launch accounting.
#pragma once

#include <cuda_runtime.h>

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
// Inside a TempScope allocations are stream-ordered (cudaMallocAsync / cudaFreeAsync on the scope's
// stream, pool kept warm): the many short-lived scratch buffers of the sort / scan pipelines then cost
// neither a device synchronisation nor a real cudaMalloc.  Outside a scope: plain cudaMalloc / cudaFree.
struct TempState {
  bool active = false;
  cudaStream_t stream = nullptr;
};
extern thread_local TempState g_temp;
struct TempScope {
  TempState saved;
  explicit TempScope(cudaStream_t s);
  ~TempScope() { g_temp = saved; }
};

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
  void alloc(size_t count) {
    release();
    n = count;
    if (!count) return;
    if (g_temp.active) {
      SB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&p), count * sizeof(T), g_temp.stream));
      pooled = true;
    } else {
      SB_CUDA(cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T)));
      pooled = false;
    }
  }
  void release() {
    if (p) {
      if (pooled && g_temp.active) cudaFreeAsync(p, g_temp.stream);
      else cudaFree(p);  // also valid for stream-ordered allocations (synchronises)
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

inline u32 ceil_div(u32 a, u32 b) { return (a + b - 1) / b; }
inline u64 ceil_div64(u64 a, u64 b) { return (a + b - 1) / b; }

int device_sm_count();

}  // namespace sb
