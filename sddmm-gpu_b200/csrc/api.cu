// api.cu -- the C ABI of include/sddmm_b200.h.  Thin: argument checks, H2D/D2H for the host-buffer
// forms, exception -> error code translation.  No CPU compute path exists behind any entry point.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include "eval_stats.cuh"
#include "layout.cuh"
#include "reorder_rows.cuh"
#include "sddmm_kernels.cuh"

namespace sb {
thread_local u64 g_launches = 0;
thread_local TempState g_temp;
namespace {
// Scratch arena behind TempScope (see common.cuh).
struct Arena {
  struct Block { size_t off, size; bool live; };
  struct Chunk { char* base; size_t cap, top; std::vector<Block> blocks; };
  std::vector<Chunk> chunks;
  int depth = 0;
  size_t total = 0;
  static constexpr size_t kKeep = (size_t)64 << 20;
  static constexpr size_t kMaxStep = (size_t)4 << 30;

  void add_chunk(size_t cap) {
    Chunk c{nullptr, cap, 0, {}};
    static const bool dbg = getenv("SDDMM_B200_ARENA_DEBUG") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    SB_CUDA(cudaMalloc(reinterpret_cast<void**>(&c.base), cap));
    if (dbg)
      fprintf(stderr, "[arena] cudaMalloc %.1f MB: %.3f ms (total %.1f MB)\n", cap / 1048576.0,
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(),
              (total + cap) / 1048576.0);
    chunks.push_back(std::move(c));
    total += cap;
  }
  void* alloc(size_t bytes) {
    bytes = (bytes + 255) & ~(size_t)255;
    for (auto& c : chunks)
      if (c.cap - c.top >= bytes) {
        c.blocks.push_back({c.top, bytes, true});
        c.top += bytes;
        return c.base + (c.top - bytes);
      }
    add_chunk(std::max(bytes, std::max(kKeep, std::min(total, kMaxStep))));  // geometric, then 4 GB steps
    Chunk& c = chunks.back();
    c.blocks.push_back({0, bytes, true});
    c.top = bytes;
    return c.base;
  }
  void free(void* ptr) {
    char* q = static_cast<char*>(ptr);
    for (auto& c : chunks) {
      if (q < c.base || q >= c.base + c.cap) continue;
      const size_t off = (size_t)(q - c.base);
      for (size_t i = c.blocks.size(); i-- > 0;)
        if (c.blocks[i].off == off) { c.blocks[i].live = false; break; }
      while (!c.blocks.empty() && !c.blocks.back().live) {
        c.top = c.blocks.back().off;
        c.blocks.pop_back();
      }
      return;
    }
  }
  void trim() {  // outermost scope ended: give the big chunks back, keep at most one small one
    static const bool dbg = getenv("SDDMM_B200_ARENA_DEBUG") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    struct Report {
      bool on; std::chrono::steady_clock::time_point t0; size_t n;
      ~Report() {
        if (on) fprintf(stderr, "[arena] trim of %zu chunks: %.3f ms\n", n,
                        std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
      }
    } report{dbg, t0, chunks.size()};
    bool kept = false;
    std::vector<Chunk> keep;
    for (auto& c : chunks) {
      if (!kept && c.cap <= kKeep && c.blocks.empty()) {
        kept = true;
        keep.push_back(std::move(c));
      } else {
        cudaFree(c.base);
      }
    }
    chunks = std::move(keep);
    total = 0;
    for (auto& c : chunks) total += c.cap;
  }
};
thread_local Arena g_arena;
}  // namespace

void* temp_alloc(size_t bytes) { return g_arena.alloc(bytes); }
void temp_free(void* p) { g_arena.free(p); }

TempScope::TempScope(cudaStream_t s, size_t hintBytes) : saved(g_temp) {
  if (g_arena.depth++ == 0 && hintBytes > g_arena.total) {
    try {
      g_arena.add_chunk(hintBytes - g_arena.total > Arena::kKeep ? hintBytes - g_arena.total : Arena::kKeep);
    } catch (...) {
      --g_arena.depth;
      throw;
    }
  }
  g_temp.active = true;
  g_temp.stream = s;
}
TempScope::~TempScope() {
  if (--g_arena.depth == 0) {
    cudaStreamSynchronize(g_temp.stream);  // work that used arena memory is done before it is returned
    g_arena.trim();
  }
  g_temp = saved;
}
static thread_local std::string g_err;
void set_last_error(const std::string& m) { g_err = m; }
int device_sm_count() {
  int dev = 0, n = 0;
  SB_CUDA(cudaGetDevice(&dev));
  SB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  return n;
}
}  // namespace sb

namespace sb {
void shard_cuts_from_prefix(const std::vector<u64>& pre, u32 numShards, u32* cuts);  // mgpu.cu
}
using namespace sb;

#define API_BEGIN try {
#define API_END                                   \
  }                                               \
  catch (const sb::Error& e) {                    \
    sb::set_last_error(e.what());                 \
    return e.code;                                \
  }                                               \
  catch (const std::exception& e) {               \
    sb::set_last_error(e.what());                 \
    return SDDMM_E_CUDA;                          \
  }                                               \
  return SDDMM_OK;

static void require(bool ok, const char* what) {
  if (!ok) fail(SDDMM_E_ARG, "invalid argument: %s", what);
}
static void require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    fail(SDDMM_E_CUDA, "no CUDA device available (%s): libsddmm_b200 has no CPU fallback",
         e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
}
// a layout's arrays, cached private layouts and workspaces live on the device it was built on
static void require_layout_device(const bsmr_layout* L) {
  int dev = -1;
  SB_CUDA(cudaGetDevice(&dev));
  if (dev != L->device)
    fail(SDDMM_E_ARG, "layout belongs to CUDA device %d but device %d is current", L->device, dev);
}

extern "C" {

int sddmm_b200_abi_version(void) { return 2; }
const char* sddmm_last_error(void) { return sb::g_err.c_str(); }
uint64_t sddmm_launch_count(void) { return sb::g_launches; }
void sddmm_launch_count_reset(void) { sb::g_launches = 0; }

uint32_t bsmr_calc_block_size(uint32_t M, uint32_t N, uint64_t free_mem_bytes) {
  if (free_mem_bytes == 0) {
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess || fr == 0) {
      sb::set_last_error("bsmr_calc_block_size: cudaMemGetInfo failed (no device?)");
      return 0;
    }
    free_mem_bytes = fr;
  }
  return calc_block_size(M, N, free_mem_bytes);
}

int bsmr_row_reorder_dev_ex(const uint32_t* d_rowOff, const uint32_t* d_colIdx, uint32_t M, uint32_t N, uint32_t nnz,
                            float alpha, uint32_t block_size, const bsmr_reorder_opts* opts, uint32_t* d_reorderedRows,
                            uint32_t* numRows, int32_t* numClusters, float* ms, void* stream) {
  API_BEGIN
  require_device();
  require(d_rowOff && d_colIdx && d_reorderedRows && numRows, "null pointer");
  if (block_size == 0) block_size = bsmr_calc_block_size(M, N, 0);
  require(block_size > 0, "block_size");
  if (opts)
    require(opts->kernel <= BSMR_CLUSTER_BATCHED && opts->laneRows <= BSMR_TRISTATE_ON &&
                opts->signature <= BSMR_TRISTATE_ON &&
                (opts->batch == 0 || opts->batch == 1 || opts->batch == 2 || opts->batch == 4 || opts->batch == 8),
            "bsmr_reorder_opts holds an unknown selector");
  cudaStream_t s = (cudaStream_t)stream;
  NvtxRange nvtx("sddmm_b200: row reorder (encode + sort + cluster)");
  Timer t(s);
  t.start();
  row_reorder_dev(d_rowOff, d_colIdx, M, N, nnz, alpha, block_size, opts, d_reorderedRows, numRows, numClusters, nullptr,
                  s);
  const float el = t.stop();
  if (ms) *ms = el;
  API_END
}
int bsmr_row_reorder_dev(const uint32_t* d_rowOff, const uint32_t* d_colIdx, uint32_t M, uint32_t N, uint32_t nnz,
                         float alpha, uint32_t block_size, uint32_t* d_reorderedRows, uint32_t* numRows,
                         int32_t* numClusters, float* ms, void* stream) {
  return bsmr_row_reorder_dev_ex(d_rowOff, d_colIdx, M, N, nnz, alpha, block_size, nullptr, d_reorderedRows, numRows,
                                 numClusters, ms, stream);
}

int bsmr_row_reorder_ex(const uint32_t* h_rowOff, const uint32_t* h_colIdx, uint32_t M, uint32_t N, uint32_t nnz,
                        float alpha, uint32_t block_size, const bsmr_reorder_opts* opts, uint32_t* h_reorderedRows,
                        uint32_t* numRows, int32_t* numClusters, float* ms) {
  API_BEGIN
  require_device();
  require(h_rowOff && h_colIdx && h_reorderedRows && numRows, "null pointer");
  DevBuf<u32> ro((size_t)M + 1), ci(nnz ? nnz : 1), out(M ? M : 1);
  SB_CUDA(cudaMemcpy(ro.get(), h_rowOff, ((size_t)M + 1) * 4, cudaMemcpyHostToDevice));
  SB_CUDA(cudaMemcpy(ci.get(), h_colIdx, (size_t)nnz * 4, cudaMemcpyHostToDevice));
  const int rc = bsmr_row_reorder_dev_ex(ro.get(), ci.get(), M, N, nnz, alpha, block_size, opts, out.get(), numRows,
                                         numClusters, ms, nullptr);
  if (rc) return rc;
  SB_CUDA(cudaMemcpy(h_reorderedRows, out.get(), (size_t)(*numRows) * 4, cudaMemcpyDeviceToHost));
  API_END
}
int bsmr_row_reorder(const uint32_t* h_rowOff, const uint32_t* h_colIdx, uint32_t M, uint32_t N, uint32_t nnz,
                     float alpha, uint32_t block_size, uint32_t* h_reorderedRows, uint32_t* numRows,
                     int32_t* numClusters, float* ms) {
  return bsmr_row_reorder_ex(h_rowOff, h_colIdx, M, N, nnz, alpha, block_size, nullptr, h_reorderedRows, numRows,
                             numClusters, ms);
}

int bsmr_dispersion_dev(const uint32_t* d_rowOff, const uint32_t* d_colIdx, uint32_t M, uint32_t N, uint32_t nnz,
                        uint32_t block_size, uint32_t* d_dispersion, uint32_t* numBlocksPerRow, void* stream) {
  API_BEGIN
  require_device();
  require(d_rowOff && d_colIdx && d_dispersion, "null pointer");
  if (block_size == 0) block_size = bsmr_calc_block_size(M, N, 0);
  dispersion_dev(d_rowOff, d_colIdx, M, N, nnz, block_size, d_dispersion, numBlocksPerRow, (cudaStream_t)stream);
  API_END
}

int bsmr_layout_build_dev_ex(const uint32_t* d_rowOff, const uint32_t* d_colIdx, uint32_t M, uint32_t N, uint32_t nnz,
                             const uint32_t* d_reorderedRows, uint32_t numRows, float delta, uint32_t panelBegin,
                             uint32_t panelEnd, uint32_t flags, bsmr_layout** out, float* msColReorder, float* msRphm,
                             void* stream) {
  API_BEGIN
  require_device();
  require(d_rowOff && d_colIdx && out && (d_reorderedRows || numRows == 0), "null pointer");
  require(flags <= BSMR_BUILD_TILES_NEVER, "flags");
  NvtxRange nvtx("sddmm_b200: column reorder + RPHM layout");
  *out = layout_build_dev(d_rowOff, d_colIdx, M, N, nnz, d_reorderedRows, numRows, delta, panelBegin, panelEnd, flags,
                          msColReorder, msRphm, (cudaStream_t)stream);
  API_END
}
int bsmr_layout_build_dev(const uint32_t* d_rowOff, const uint32_t* d_colIdx, uint32_t M, uint32_t N, uint32_t nnz,
                          const uint32_t* d_reorderedRows, uint32_t numRows, float delta, uint32_t panelBegin,
                          uint32_t panelEnd, bsmr_layout** out, float* msColReorder, float* msRphm, void* stream) {
  return bsmr_layout_build_dev_ex(d_rowOff, d_colIdx, M, N, nnz, d_reorderedRows, numRows, delta, panelBegin, panelEnd,
                                  BSMR_BUILD_TILES_AUTO, out, msColReorder, msRphm, stream);
}

int bsmr_layout_build_ex(const uint32_t* h_rowOff, const uint32_t* h_colIdx, uint32_t M, uint32_t N, uint32_t nnz,
                         const uint32_t* h_reorderedRows, uint32_t numRows, float delta, uint32_t flags,
                         bsmr_layout** out, float* msColReorder, float* msRphm) {
  API_BEGIN
  require_device();
  require(h_rowOff && h_colIdx && out && (h_reorderedRows || numRows == 0), "null pointer");
  require(flags <= BSMR_BUILD_TILES_NEVER, "flags");
  DevBuf<u32> ro((size_t)M + 1), ci(nnz ? nnz : 1), rr(numRows ? numRows : 1);
  SB_CUDA(cudaMemcpy(ro.get(), h_rowOff, ((size_t)M + 1) * 4, cudaMemcpyHostToDevice));
  SB_CUDA(cudaMemcpy(ci.get(), h_colIdx, (size_t)nnz * 4, cudaMemcpyHostToDevice));
  SB_CUDA(cudaMemcpy(rr.get(), h_reorderedRows, (size_t)numRows * 4, cudaMemcpyHostToDevice));
  *out = layout_build_dev(ro.get(), ci.get(), M, N, nnz, rr.get(), numRows, delta, 0, 0xFFFFFFFFu, flags, msColReorder,
                          msRphm, nullptr);
  API_END
}
int bsmr_layout_build(const uint32_t* h_rowOff, const uint32_t* h_colIdx, uint32_t M, uint32_t N, uint32_t nnz,
                      const uint32_t* h_reorderedRows, uint32_t numRows, float delta, bsmr_layout** out,
                      float* msColReorder, float* msRphm) {
  return bsmr_layout_build_ex(h_rowOff, h_colIdx, M, N, nnz, h_reorderedRows, numRows, delta, BSMR_BUILD_TILES_AUTO, out,
                              msColReorder, msRphm);
}

void bsmr_layout_destroy(bsmr_layout* L) { delete L; }

int bsmr_layout_save(const bsmr_layout* L, const char* path) {
  API_BEGIN
  require_device();
  require(L && path, "null pointer");
  layout_save(L, path);
  API_END
}
int bsmr_layout_load(const char* path, bsmr_layout** out) {
  API_BEGIN
  require_device();
  require(path && out, "null pointer");
  *out = layout_load(path);
  API_END
}

int bsmr_layout_eval(const bsmr_layout* L, float delta, bsmr_eval* out) {
  API_BEGIN
  require_device();
  require(L && out, "null pointer");
  layout_eval(L, delta, out, nullptr);
  API_END
}
int bsmr_original_block_stats_dev(const uint32_t* d_rowOff, const uint32_t* d_colIdx, uint32_t M, uint32_t N,
                                  uint32_t nnz, float delta, uint32_t* numDenseBlocks, float* averageDensity,
                                  void* stream) {
  API_BEGIN
  require_device();
  require(d_rowOff && (d_colIdx || nnz == 0) && numDenseBlocks && averageDensity, "null pointer");
  original_block_stats_dev(d_rowOff, d_colIdx, M, N, nnz, delta, numDenseBlocks, averageDensity,
                           static_cast<cudaStream_t>(stream));
  API_END
}
int bsmr_original_block_stats(const uint32_t* h_rowOff, const uint32_t* h_colIdx, uint32_t M, uint32_t N,
                              uint32_t nnz, float delta, uint32_t* numDenseBlocks, float* averageDensity) {
  API_BEGIN
  require_device();
  require(h_rowOff && (h_colIdx || nnz == 0) && numDenseBlocks && averageDensity, "null pointer");
  DevBuf<u32> ro((size_t)M + 1), ci(nnz ? nnz : 1);
  SB_CUDA(cudaMemcpy(ro.get(), h_rowOff, ((size_t)M + 1) * 4, cudaMemcpyHostToDevice));
  SB_CUDA(cudaMemcpy(ci.get(), h_colIdx, (size_t)nnz * 4, cudaMemcpyHostToDevice));
  original_block_stats_dev(ro.get(), ci.get(), M, N, nnz, delta, numDenseBlocks, averageDensity, nullptr);
  API_END
}

int bsmr_layout_get_info(const bsmr_layout* L, bsmr_layout_info* out) {
  API_BEGIN
  require(L && out, "null pointer");
  *out = L->info;
  API_END
}
size_t bsmr_layout_array_len(const bsmr_layout* L, bsmr_array_id which) {
  if (!L || which < 0 || which >= BSMR_ARRAY_COUNT) return 0;
  return L->arr[which].size();
}
const uint32_t* bsmr_layout_array_dev(const bsmr_layout* L, bsmr_array_id which) {
  if (!L || which < 0 || which >= BSMR_ARRAY_COUNT) return nullptr;
  return L->arr[which].get();
}
int bsmr_layout_array_to_host(const bsmr_layout* L, bsmr_array_id which, uint32_t* h_dst, size_t capacity) {
  API_BEGIN
  require(L && which >= 0 && which < BSMR_ARRAY_COUNT, "layout / array id");
  const size_t n = L->arr[which].size();
  require(capacity >= n && (h_dst || n == 0), "capacity too small");
  if (n) SB_CUDA(cudaMemcpy(h_dst, L->arr[which].get(), n * 4, cudaMemcpyDeviceToHost));
  API_END
}

// ---- SDDMM ------------------------------------------------------------------------------------
namespace {
struct Streams {
  cudaStream_t dense = nullptr, sparse = nullptr;
  cudaEvent_t fork = nullptr, joinD = nullptr, joinS = nullptr;
  Streams() {
    SB_CUDA(cudaStreamCreateWithFlags(&dense, cudaStreamNonBlocking));
    SB_CUDA(cudaStreamCreateWithFlags(&sparse, cudaStreamNonBlocking));
    SB_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
    SB_CUDA(cudaEventCreateWithFlags(&joinD, cudaEventDisableTiming));
    SB_CUDA(cudaEventCreateWithFlags(&joinS, cudaEventDisableTiming));
  }
};
Streams& streams() {
  // per host thread AND per device (streams and events belong to the device that was current at creation);
  // leaked at exit on purpose
  static thread_local std::map<int, Streams*>* byDev = nullptr;
  if (!byDev) byDev = new std::map<int, Streams*>();
  int dev = 0;
  SB_CUDA(cudaGetDevice(&dev));
  Streams*& s = (*byDev)[dev];
  if (!s) s = new Streams();
  return *s;
}
// dense || residual, forked from and joined back into `s`
void run_once(const bsmr_layout* L, u32 K, const float* dA, const float* dB, float* dP, cudaStream_t s,
              u32 numBatch = 1, const sddmm_plan* plan = nullptr) {
  require_layout_device(L);
  NvtxRange nvtx("sddmm_b200: SDDMM pass");
  sddmm_plan p;
  plan_resolve(L, K, numBatch ? numBatch : 1, plan, &p);
  if (p.plan == SDDMM_PLAN_BSMR && L->numDenseWork && L->numSparseWork) {
    Streams& st = streams();
    plan_prepare(L, K, numBatch, p, s);  // no-op when everything is cached; never inside the fork
    SB_CUDA(cudaEventRecord(st.fork, s));
    SB_CUDA(cudaStreamWaitEvent(st.dense, st.fork, 0));
    SB_CUDA(cudaStreamWaitEvent(st.sparse, st.fork, 0));
    sddmm_launch(L, K, dA, dB, dP, st.dense, st.sparse, kLaunchBoth, numBatch, &p);
    SB_CUDA(cudaEventRecord(st.joinD, st.dense));
    SB_CUDA(cudaEventRecord(st.joinS, st.sparse));
    SB_CUDA(cudaStreamWaitEvent(s, st.joinD, 0));
    SB_CUDA(cudaStreamWaitEvent(s, st.joinS, 0));
  } else {
    sddmm_launch(L, K, dA, dB, dP, s, s, kLaunchBoth, numBatch, &p);
  }
}
}  // namespace

void sddmm_plan_default(sddmm_plan* out) {
  if (out) plan_default(out);
}
int sddmm_plan_resolve(const bsmr_layout* L, uint32_t K, uint32_t numBatch, const sddmm_plan* in, sddmm_plan* out) {
  API_BEGIN
  require(L && out, "null pointer");
  plan_resolve(L, K, numBatch ? numBatch : 1, in, out);
  API_END
}
int sddmm_prepare(const bsmr_layout* L, uint32_t K, uint32_t numBatch, const sddmm_plan* plan) {
  API_BEGIN
  require_device();
  require(L && numBatch > 0, "arguments");
  require_layout_device(L);
  sddmm_plan p;
  plan_resolve(L, K, numBatch, plan, &p);
  cudaStream_t s = streams().dense;
  NvtxRange nvtx("sddmm_b200: prepare (K-dependent private layouts)");
  plan_prepare(L, K, numBatch, p, s);
  SB_CUDA(cudaStreamSynchronize(s));
  API_END
}
int sddmm_run_dev_ex(const bsmr_layout* L, uint32_t K, uint32_t numBatch, const float* d_A, const float* d_B, float* d_P,
                     const sddmm_plan* plan, void* stream) {
  API_BEGIN
  require_device();
  require(L && d_A && d_B && d_P && numBatch > 0, "arguments");
  run_once(L, K, d_A, d_B, d_P, (cudaStream_t)stream, numBatch, plan);
  API_END
}

int sddmm_run_dev(const bsmr_layout* L, uint32_t K, const float* d_A, const float* d_B, float* d_P, void* stream) {
  API_BEGIN
  require_device();
  require(L && d_A && d_B && d_P, "null pointer");
  run_once(L, K, d_A, d_B, d_P, (cudaStream_t)stream);
  API_END
}

int sddmm_run_batch_dev(const bsmr_layout* L, uint32_t K, uint32_t numBatch, const float* d_A, const float* d_B,
                        float* d_P, void* stream) {
  API_BEGIN
  require_device();
  require(L && d_A && d_B && d_P, "null pointer");
  run_once(L, K, d_A, d_B, d_P, (cudaStream_t)stream, numBatch);
  API_END
}

int sddmm_run_timed_dev(const bsmr_layout* L, uint32_t K, const float* d_A, const float* d_B, float* d_P, int warmup,
                        int iters, float* msDense, float* msSparse, float* msTotal) {
  API_BEGIN
  require_device();
  require(L && d_A && d_B && d_P && iters > 0 && warmup >= 0, "arguments");
  require_layout_device(L);
  Streams& st = streams();
  cudaStream_t s = st.dense;  // the launching stream of the combined pass
  for (int i = 0; i < warmup; ++i) run_once(L, K, d_A, d_B, d_P, s);
  SB_CUDA(cudaStreamSynchronize(s));
  Timer t(s);
  t.start();
  for (int i = 0; i < iters; ++i) run_once(L, K, d_A, d_B, d_P, s);
  const float tot = t.stop() / iters;
  if (msTotal) *msTotal = tot;
  // each kernel alone (its own stream, nothing else running): the per-kernel roofline figures
  if (msDense) {
    *msDense = 0.f;
    if (L->numDenseWork) {
      Timer td(st.dense);
      td.start();
      for (int i = 0; i < iters; ++i) sddmm_launch(L, K, d_A, d_B, d_P, st.dense, st.dense, kLaunchDense);
      *msDense = td.stop() / iters;
    }
  }
  if (msSparse) {
    *msSparse = 0.f;
    if (L->numSparseWork) {
      Timer ts(st.sparse);
      ts.start();
      for (int i = 0; i < iters; ++i) sddmm_launch(L, K, d_A, d_B, d_P, st.sparse, st.sparse, kLaunchSparse);
      *msSparse = ts.stop() / iters;
    }
  }
  API_END
}

}  // extern "C"

namespace sb_hostcopy {
using namespace sb;
// Host -> device copy of the operands of one pass.  Whole arrays by cudaMemcpyAsync, or -- when both host buffers are
// page-locked (device-accessible through UVA) and the layout references clearly fewer rows than the arrays hold (an
// R-MAT graph leaves half of its rows and columns empty) -- a gather kernel that reads ONLY the referenced A rows
// (reorderedRows) and B^T rows (ensure_host_refs) through the mapped pointers: PCIe carries what the pass reads.
// Unreferenced rows of the staging buffers keep their (zeroed) contents; no kernel result depends on them.
static __global__ void __launch_bounds__(256) k_gather_rows_h2d(const u32* __restrict__ listA, u32 nA, const u32* __restrict__ listB,
                                                         u32 nB, u32 limA, u32 limB, u32 K4, const float4* hA,
                                                         const float4* hB, float4* __restrict__ dA,
                                                         float4* __restrict__ dB) {
  const u32 lane = threadIdx.x & 31u;
  const u32 warps = (gridDim.x * blockDim.x) >> 5, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (u32 i = gw; i < nA + nB; i += warps) {
    const bool isA = i < nA;
    const u32 row = isA ? listA[i] : listB[i - nA];
    if (row >= (isA ? limA : limB)) continue;
    const float4* src = (isA ? hA : hB) + (size_t)row * K4;
    float4* dst = (isA ? dA : dB) + (size_t)row * K4;
    for (u32 c0 = 0; c0 < K4; c0 += 128u) {  // up to 4 x 16 bytes in flight per lane
      float4 v[4];
#pragma unroll
      for (u32 u = 0; u < 4; ++u) {
        const u32 c = c0 + u * 32u + lane;
        if (c < K4) v[u] = src[c];
      }
#pragma unroll
      for (u32 u = 0; u < 4; ++u) {
        const u32 c = c0 + u * 32u + lane;
        if (c < K4) dst[c] = v[u];
      }
    }
  }
}

static int h2d_mode() {  // SDDMM_B200_H2D = auto | full | gather
  const char* e = getenv("SDDMM_B200_H2D");
  if (!e) return 0;
  return !strcmp(e, "full") ? 1 : !strcmp(e, "gather") ? 2 : 0;
}
static bool pinned_host(const void* p, const void** dev) {
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  if (at.type != cudaMemoryTypeHost || !at.devicePointer) return false;
  *dev = at.devicePointer;
  return true;
}
static size_t h2d_operands(const bsmr_layout* L, u32 K, const float* hA, const float* hB, float* dA, float* dB, cudaStream_t s) {
  const bsmr_layout_info& I = L->info;
  const size_t nA = (size_t)I.M * K, nB = (size_t)I.N * K;
  const int mode = h2d_mode();
  const void *mA = nullptr, *mB = nullptr;
  if (mode != 1 && !(K & 3u) && pinned_host(hA, &mA) && pinned_host(hB, &mB)) {
    const bsmr_layout::HostRefs* hr = ensure_host_refs(L, s);
    const size_t refRows = (size_t)I.numRows + hr->numCols;
    if (mode == 2 || refRows * 100 <= ((size_t)I.M + I.N) * 85) {
      // few CTAs: the kernel waits on PCIe, it must not crowd the SDDMM pass of the other slot off the SMs
      static const int ctas = [] { const char* e = getenv("SDDMM_B200_H2D_CTAS"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 128; }();
      k_gather_rows_h2d<<<ctas, 256, 0, s>>>(L->arr[BSMR_REORDERED_ROWS].get(), I.numRows, hr->cols.get(), hr->numCols,
                                            I.M, I.N, K / 4, static_cast<const float4*>(mA),
                                            static_cast<const float4*>(mB), reinterpret_cast<float4*>(dA),
                                            reinterpret_cast<float4*>(dB));
      SB_LAUNCH_CHECK();
      return refRows * K * 4;
    }
  }
  SB_CUDA(cudaMemcpyAsync(dA, hA, nA * 4, cudaMemcpyHostToDevice, s));
  SB_CUDA(cudaMemcpyAsync(dB, hB, nB * 4, cudaMemcpyHostToDevice, s));
  return (nA + nB) * 4;
}
}  // namespace sb_hostcopy
using sb_hostcopy::h2d_operands;

extern "C" {

int sddmm_host_traffic(const bsmr_layout* L, uint64_t* h2dBytes, uint64_t* d2hBytes) {
  API_BEGIN
  require(L, "null layout");
  if (h2dBytes) *h2dBytes = L->lastH2DBytes;
  if (d2hBytes) *d2hBytes = L->lastD2HBytes;
  API_END
}

int sddmm_run_host(const bsmr_layout* L, uint32_t K, const float* h_A, const float* h_B, float* h_P, float* msTotal) {
  API_BEGIN
  require_device();
  require(L && h_A && h_B && h_P, "null pointer");
  require_layout_device(L);
  const bsmr_layout_info& I = L->info;
  cudaStream_t s = streams().dense;
  Timer t(s);
  t.start();
  const size_t nA = (size_t)I.M * K, nB = (size_t)I.N * K, nP = I.nnz ? I.nnz : 1;
  if (L->wsA.size() < nA) { L->wsA.alloc(nA, true); SB_CUDA(cudaMemsetAsync(L->wsA.get(), 0, nA * 4, s)); }
  if (L->wsB.size() < nB) { L->wsB.alloc(nB, true); SB_CUDA(cudaMemsetAsync(L->wsB.get(), 0, nB * 4, s)); }
  if (L->wsP.size() < nP) {
    L->wsP.alloc(nP, true);
    SB_CUDA(cudaMemsetAsync(L->wsP.get(), 0, nP * 4, s));
  }
  float *dA = L->wsA.get(), *dB = L->wsB.get(), *dP = L->wsP.get();
  L->lastH2DBytes = h2d_operands(L, K, h_A, h_B, dA, dB, s);
  L->lastD2HBytes = (unsigned long long)I.nnz * 4;
  // (the reference zero-fills P, sddmmKernel.cu:2525; every entry this layout covers is overwritten by the pass --
  //  the NaN-canary test proves it -- so a full layout needs no memset; a row-panel shard leaves foreign entries
  //  as they were, hence the one-time zeroing when the staging buffer is created)
  run_once(L, K, dA, dB, dP, s);
  SB_CUDA(cudaMemcpyAsync(h_P, dP, (size_t)I.nnz * 4, cudaMemcpyDeviceToHost, s));
  const float el = t.stop();
  if (msTotal) *msTotal = el;
  API_END
}

int sddmm_run_host_async(const bsmr_layout* L, uint32_t K, const float* h_A, const float* h_B, float* h_P, int slot) {
  API_BEGIN
  require_device();
  require(L && h_A && h_B && h_P && (slot == 0 || slot == 1), "arguments");
  require_layout_device(L);
  const bsmr_layout_info& I = L->info;
  if (!L->pipe) {
    auto pp = std::make_unique<bsmr_layout::HostPipe>();
    SB_CUDA(cudaStreamCreateWithFlags(&pp->h2d, cudaStreamNonBlocking));
    SB_CUDA(cudaStreamCreateWithFlags(&pp->comp, cudaStreamNonBlocking));
    SB_CUDA(cudaStreamCreateWithFlags(&pp->d2h, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      SB_CUDA(cudaEventCreateWithFlags(&pp->evH2D[i], cudaEventDisableTiming));
      SB_CUDA(cudaEventCreateWithFlags(&pp->evComp[i], cudaEventDisableTiming));
      SB_CUDA(cudaEventCreateWithFlags(&pp->evD2H[i], cudaEventDisableTiming));
    }
    L->pipe = std::move(pp);
  }
  bsmr_layout::HostPipe& P = *L->pipe;
  const size_t nA = (size_t)I.M * K, nB = (size_t)I.N * K, nP = I.nnz ? I.nnz : 1;
  if (P.A[slot].size() < nA || P.B[slot].size() < nB || P.P[slot].size() < nP) {
    SB_CUDA(cudaDeviceSynchronize());  // growing a slot: nothing may still be using it
    if (P.A[slot].size() < nA) { P.A[slot].alloc(nA, true); SB_CUDA(cudaMemset(P.A[slot].get(), 0, nA * 4)); }
    if (P.B[slot].size() < nB) { P.B[slot].alloc(nB, true); SB_CUDA(cudaMemset(P.B[slot].get(), 0, nB * 4)); }
    if (P.P[slot].size() < nP) {
      P.P[slot].alloc(nP, true);
      SB_CUDA(cudaMemset(P.P[slot].get(), 0, nP * 4));  // once; see sddmm_run_host
    }
  }
  // H2D of this batch may start once the previous pass on this slot has consumed A/B
  SB_CUDA(cudaStreamWaitEvent(P.h2d, P.evComp[slot], 0));
  L->lastH2DBytes = h2d_operands(L, K, h_A, h_B, P.A[slot].get(), P.B[slot].get(), P.h2d);
  L->lastD2HBytes = (unsigned long long)I.nnz * 4;
  SB_CUDA(cudaEventRecord(P.evH2D[slot], P.h2d));
  // the pass needs the operands and a P buffer whose previous contents have left for the host
  SB_CUDA(cudaStreamWaitEvent(P.comp, P.evH2D[slot], 0));
  SB_CUDA(cudaStreamWaitEvent(P.comp, P.evD2H[slot], 0));
  run_once(L, K, P.A[slot].get(), P.B[slot].get(), P.P[slot].get(), P.comp);
  SB_CUDA(cudaEventRecord(P.evComp[slot], P.comp));
  SB_CUDA(cudaStreamWaitEvent(P.d2h, P.evComp[slot], 0));
  SB_CUDA(cudaMemcpyAsync(h_P, P.P[slot].get(), (size_t)I.nnz * 4, cudaMemcpyDeviceToHost, P.d2h));
  SB_CUDA(cudaEventRecord(P.evD2H[slot], P.d2h));
  API_END
}

int sddmm_host_sync(const bsmr_layout* L) {
  API_BEGIN
  require(L, "null layout");
  if (L->pipe) {
    SB_CUDA(cudaStreamSynchronize(L->pipe->h2d));
    SB_CUDA(cudaStreamSynchronize(L->pipe->comp));
    SB_CUDA(cudaStreamSynchronize(L->pipe->d2h));
  }
  API_END
}

int sddmm_host(const uint32_t* h_rowOff, const uint32_t* h_colIdx, uint32_t M, uint32_t N, uint32_t nnz, uint32_t K,
               const float* h_A, const float* h_B, float alpha, float delta, uint32_t block_size, float* h_P,
               sddmm_stats* stats, bsmr_layout** layoutOut) {
  API_BEGIN
  require_device();
  require(h_rowOff && h_colIdx && h_A && h_B && h_P, "null pointer");
  if (block_size == 0) block_size = bsmr_calc_block_size(M, N, 0);
  DevBuf<u32> ro((size_t)M + 1), ci(nnz ? nnz : 1), rr(M ? M : 1);
  SB_CUDA(cudaMemcpy(ro.get(), h_rowOff, ((size_t)M + 1) * 4, cudaMemcpyHostToDevice));
  SB_CUDA(cudaMemcpy(ci.get(), h_colIdx, (size_t)nnz * 4, cudaMemcpyHostToDevice));
  u32 numRows = 0;
  int32_t ncl = 0;
  float msRow = 0.f, msCol = 0.f, msRphm = 0.f, msRun = 0.f;
  int rc = bsmr_row_reorder_dev(ro.get(), ci.get(), M, N, nnz, alpha, block_size, rr.get(), &numRows, &ncl, &msRow,
                                nullptr);
  if (rc) return rc;
  bsmr_layout* L = nullptr;
  rc = bsmr_layout_build_dev(ro.get(), ci.get(), M, N, nnz, rr.get(), numRows, delta, 0, 0xFFFFFFFFu, &L, &msCol,
                             &msRphm, nullptr);
  if (rc) return rc;
  rc = sddmm_run_host(L, K, h_A, h_B, h_P, &msRun);
  if (stats) {
    stats->rowReorderMs = msRow; stats->colReorderMs = msCol; stats->rphmMs = msRphm; stats->sddmmMs = msRun;
    stats->numClusters = ncl; stats->blockSize = block_size;
  }
  if (rc || !layoutOut) bsmr_layout_destroy(L);
  else *layoutOut = L;
  if (rc) return rc;
  API_END
}

int bsmr_shard_plan(const uint32_t* h_rowOff, const uint32_t* h_reorderedRows, uint32_t numRows, uint32_t numShards,
                    uint32_t* h_cuts) {
  API_BEGIN
  require(h_rowOff && (h_reorderedRows || numRows == 0) && h_cuts && numShards > 0, "arguments");
  const u32 P = (numRows + kPanel - 1) / kPanel;
  std::vector<u64> pre((size_t)P + 1, 0);
  for (u32 p = 0; p < P; ++p) {
    u64 c = 0;
    for (u32 i = p * kPanel; i < (p + 1) * kPanel && i < numRows; ++i) {
      const u32 r = h_reorderedRows[i];
      c += h_rowOff[r + 1] - h_rowOff[r];
    }
    pre[p + 1] = pre[p] + c;
  }
  shard_cuts_from_prefix(pre, numShards, h_cuts);
  API_END
}

}  // extern "C"
