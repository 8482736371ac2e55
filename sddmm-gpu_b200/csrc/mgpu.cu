// mgpu.cu -- multi-GPU layer (the reference is single-GPU, so there is no counterpart to cite):
// one process per GPU, row panels of the reordered matrix cut into nnz-balanced contiguous ranges, B and the row
// order replicated ONCE over NCCL (NVLink 5 / NVSwitch), no collective in the steady state.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy torch ships is found when the caller is a torchrun
// rank; SDDMM_B200_NCCL_LIB names another one), so that the library loads and the single-GPU path works on a box
// without NCCL.  The 128-byte communicator id travels out of band: the caller moves it from rank 0 to the other
// ranks with whatever it has (a file, MPI, torch.distributed) -- see INTEGRATION.md.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>

#include <vector>

#include "layout.cuh"
#include "primitives.cuh"
#include "sddmm_kernels.cuh"

namespace sb {
namespace {

struct NcclId { char internal[SDDMM_MGPU_ID_BYTES]; };
using ncclComm_t = void*;
constexpr int kNcclUint8 = 1, kNcclFloat32 = 7, kNcclSum = 0;

struct Nccl {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, NcclId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

const Nccl& nccl() {
  static Nccl n = [] {
    Nccl r;
    const char* names[] = {getenv("SDDMM_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      if (!nm || !*nm) continue;
      r.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (r.handle) break;
    }
    if (!r.handle) return r;
    auto sym = [&](const char* s) { return dlsym(r.handle, s); };
    r.GetUniqueId = reinterpret_cast<decltype(r.GetUniqueId)>(sym("ncclGetUniqueId"));
    r.CommInitRank = reinterpret_cast<decltype(r.CommInitRank)>(sym("ncclCommInitRank"));
    r.CommDestroy = reinterpret_cast<decltype(r.CommDestroy)>(sym("ncclCommDestroy"));
    r.Broadcast = reinterpret_cast<decltype(r.Broadcast)>(sym("ncclBroadcast"));
    r.AllReduce = reinterpret_cast<decltype(r.AllReduce)>(sym("ncclAllReduce"));
    r.AllGather = reinterpret_cast<decltype(r.AllGather)>(sym("ncclAllGather"));
    r.Reduce = reinterpret_cast<decltype(r.Reduce)>(sym("ncclReduce"));
    r.GetErrorString = reinterpret_cast<decltype(r.GetErrorString)>(sym("ncclGetErrorString"));
    if (!r.GetUniqueId || !r.CommInitRank || !r.CommDestroy || !r.Broadcast || !r.AllReduce || !r.AllGather || !r.Reduce)
      r.handle = nullptr;
    return r;
  }();
  if (!n.handle)
    fail(SDDMM_E_UNSUPPORTED, "NCCL is not available (dlopen libnccl.so.2 failed; set SDDMM_B200_NCCL_LIB): %s",
         dlerror() ? dlerror() : "missing symbols");
  return n;
}

void nccl_check(int rc, const char* what) {
  if (rc != 0) {
    const Nccl& n = nccl();
    fail(SDDMM_E_CUDA, "%s failed: %s", what, n.GetErrorString ? n.GetErrorString(rc) : "NCCL error");
  }
}

// per-panel stored-entry counts of the reordered matrix
__global__ void k_panel_nnz(const u32* __restrict__ rowOff, const u32* __restrict__ R, u32 numRows, u32 P,
                            u32* __restrict__ out) {
  for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
    u32 c = 0;
    for (u32 i = (u32)p * kPanel; i < ((u32)p + 1) * kPanel && i < numRows; ++i) {
      const u32 r = R[i];
      c += rowOff[r + 1] - rowOff[r];
    }
    out[p] = c;
  }
}

}  // namespace

// nnz-balanced contiguous panel ranges from a prefix sum over per-panel counts (the rule of bsmr_shard_plan)
void shard_cuts_from_prefix(const std::vector<u64>& pre, u32 numShards, u32* cuts) {
  const u32 P = (u32)pre.size() - 1;
  const u64 total = pre[P];
  cuts[0] = 0;
  u32 p = 0;
  for (u32 s = 1; s < numShards; ++s) {
    const u64 target = (total * s + numShards / 2) / numShards;
    while (p < P && pre[p] < target) ++p;
    if (p > 0 && target - pre[p - 1] < pre[p] - target) --p;  // the closer of p-1 / p
    if (p < cuts[s - 1]) p = cuts[s - 1];
    cuts[s] = p;
  }
  cuts[numShards] = P;
}

void shard_plan_dev(const u32* d_rowOff, const u32* d_R, u32 numRows, u32 numShards, u32* cuts, cudaStream_t s,
                    std::vector<u64>* preOut = nullptr) {
  const u32 P = (numRows + kPanel - 1) / kPanel;
  std::vector<u64> pre((size_t)P + 1, 0);
  if (P) {
    DevBuf<u32> cnt(P);
    k_panel_nnz<<<grid_for(P), 256, 0, s>>>(d_rowOff, d_R, numRows, P, cnt.get());
    SB_LAUNCH_CHECK();
    std::vector<u32> h(P);
    SB_CUDA(cudaMemcpyAsync(h.data(), cnt.get(), (size_t)P * 4, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaStreamSynchronize(s));
    for (u32 p = 0; p < P; ++p) pre[p + 1] = pre[p] + h[p];
  }
  shard_cuts_from_prefix(pre, numShards, cuts);
  if (preOut) *preOut = std::move(pre);
}

// Cost-calibrated cuts: shard r took ms[r] for its pre[cuts[r+1]] - pre[cuts[r]] non-zeros, so a non-zero of that
// range is charged ms[r] / nnz_r; the new cuts give every rank the same estimated time.
void rebalance_cuts(const std::vector<u64>& pre, const std::vector<u32>& oldCuts, const std::vector<float>& ms,
                    std::vector<u32>& newCuts) {
  const u32 world = (u32)ms.size(), P = (u32)pre.size() - 1;
  std::vector<double> dens(world, 0.0);
  double total = 0.0;
  for (u32 r = 0; r < world; ++r) {
    const double n = (double)(pre[oldCuts[r + 1]] - pre[oldCuts[r]]);
    dens[r] = n > 0 ? (double)ms[r] / n : 0.0;
    total += n > 0 ? (double)ms[r] : 0.0;
  }
  newCuts.assign((size_t)world + 1, 0);
  newCuts[world] = P;
  double acc = 0.0;
  u32 s = 1, r = 0;
  for (u32 p = 0; p < P && s < world; ++p) {
    while (r + 1 < world && p >= oldCuts[r + 1]) ++r;
    const double c = dens[r] * (double)(pre[p + 1] - pre[p]);
    const double target = total * s / world;
    if (acc + c >= target) {
      // cut before or after panel p, whichever lands closer to the target
      newCuts[s] = (target - acc < acc + c - target) ? p : p + 1;
      if (newCuts[s] < newCuts[s - 1]) newCuts[s] = newCuts[s - 1];
      ++s;
      if (s < world && acc + c >= total * s / world) { --p; continue; }  // a heavy panel can close several shards
    }
    acc += c;
  }
  for (; s < world; ++s) newCuts[s] = P;
}

}  // namespace sb

using namespace sb;

struct sddmm_mgpu {
  int rank = 0, world = 1, device = 0;
  void* comm = nullptr;
  std::vector<u32> cuts;
  std::vector<u64> pre;  // prefix sum of per-panel non-zero counts of the reordered matrix (kept for rebalancing)
  // sddmm_mgpu_run_host: device staging (grown on demand), two streams and their events
  DevBuf<float> wsA, wsB, wsP, wsPfull;
  // the rows of A (non-empty rows of S) and of B^T (referenced columns) the WHOLE job reads: with page-locked host
  // buffers every rank gathers 1/world of each list over its PCIe link into a packed buffer, the packed buffers are
  // all-gathered over NVLink and unpacked on the device -- PCIe and NVLink carry only rows some rank reads
  DevBuf<u32> refRows, refCols;
  u32 nRefRows = 0, nRefCols = 0;
  DevBuf<float> packA, packB;
  unsigned long long lastH2DBytes = 0;
  cudaStream_t sCopy = nullptr, sWork = nullptr;
  cudaEvent_t evA = nullptr, evB = nullptr, evDone = nullptr, evT0 = nullptr, evT1 = nullptr;
  ~sddmm_mgpu() {
    for (cudaEvent_t e : {evA, evB, evDone, evT0, evT1}) if (e) cudaEventDestroy(e);
    if (sCopy) cudaStreamDestroy(sCopy);
    if (sWork) cudaStreamDestroy(sWork);
  }
};

// rows list[beg..end) of a page-locked host array (read through its mapped pointer) into the packed buffer at the
// same list positions; and the inverse on the device after the all-gather
static __global__ void __launch_bounds__(256) k_pack_rows_h2d(const u32* __restrict__ list, u32 beg, u32 end, u32 K4,
                                                              const float4* hsrc, float4* __restrict__ packed) {
  const u32 lane = threadIdx.x & 31u;
  const u32 warps = (gridDim.x * blockDim.x) >> 5, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (u32 i = beg + gw; i < end; i += warps) {
    const float4* src = hsrc + (size_t)list[i] * K4;
    float4* dst = packed + (size_t)i * K4;
    for (u32 c0 = 0; c0 < K4; c0 += 128u) {
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
static __global__ void __launch_bounds__(256) k_unpack_rows(const u32* __restrict__ list, u32 n, u32 K4,
                                                            const float4* __restrict__ packed, float4* __restrict__ dst) {
  const u32 lane = threadIdx.x & 31u;
  const u32 warps = (gridDim.x * blockDim.x) >> 5, gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  for (u32 i = gw; i < n; i += warps) {
    const float4* src = packed + (size_t)i * K4;
    float4* d = dst + (size_t)list[i] * K4;
    for (u32 c = lane; c < K4; c += 32u) d[c] = src[c];
  }
}
static bool host_mapped(const void* p, const void** dev) {
  cudaPointerAttributes at{};
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  if (at.type != cudaMemoryTypeHost || !at.devicePointer) return false;
  *dev = at.devicePointer;
  return true;
}

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

static void need(bool ok, const char* what) {
  if (!ok) fail(SDDMM_E_ARG, "invalid argument: %s", what);
}

extern "C" {

int sddmm_mgpu_unique_id(void* id128) {
  API_BEGIN
  need(id128 != nullptr, "null id buffer");
  NcclId id;
  nccl_check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
  std::memcpy(id128, id.internal, SDDMM_MGPU_ID_BYTES);
  API_END
}

int sddmm_mgpu_init(int rank, int world, const void* id128, sddmm_mgpu** out) {
  API_BEGIN
  need(out && world >= 1 && rank >= 0 && rank < world && (world == 1 || id128), "rank / world / id");
  auto g = std::make_unique<sddmm_mgpu>();
  g->rank = rank;
  g->world = world;
  SB_CUDA(cudaGetDevice(&g->device));
  if (world > 1) {
    NcclId id;
    std::memcpy(id.internal, id128, SDDMM_MGPU_ID_BYTES);
    nccl_check(nccl().CommInitRank(&g->comm, world, id, rank), "ncclCommInitRank");
  }
  *out = g.release();
  API_END
}

void sddmm_mgpu_destroy(sddmm_mgpu* g) {
  if (!g) return;
  if (g->comm) nccl().CommDestroy(g->comm);
  delete g;
}

int sddmm_mgpu_bcast(sddmm_mgpu* g, void* d_buf, size_t bytes, int root, void* stream) {
  API_BEGIN
  need(g && (d_buf || bytes == 0) && root >= 0 && root < g->world, "arguments");
  if (g->world > 1 && bytes)
    nccl_check(nccl().Broadcast(d_buf, d_buf, bytes, kNcclUint8, root, g->comm, (cudaStream_t)stream), "ncclBroadcast");
  API_END
}

int sddmm_mgpu_shard(sddmm_mgpu* g, const uint32_t* d_rowOff, const uint32_t* d_colIdx, uint32_t M, uint32_t N,
                     uint32_t nnz, uint32_t* d_reorderedRows, uint32_t* numRows, float delta, uint32_t flags,
                     bsmr_layout** out, uint32_t* h_cuts, float* msColReorder, float* msRphm, void* stream) {
  API_BEGIN
  need(g && d_rowOff && d_colIdx && d_reorderedRows && numRows && out, "null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  // rank 0's row order reaches every rank: the clustering is one sequential chain, it is computed once
  if (g->world > 1) {
    DevBuf<u32> n(1);
    SB_CUDA(cudaMemcpyAsync(n.get(), numRows, 4, cudaMemcpyHostToDevice, s));
    nccl_check(nccl().Broadcast(n.get(), n.get(), 4, kNcclUint8, 0, g->comm, s), "ncclBroadcast(numRows)");
    SB_CUDA(cudaMemcpyAsync(numRows, n.get(), 4, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaStreamSynchronize(s));
    need(*numRows <= M, "numRows from rank 0 exceeds M");
    if (*numRows)
      nccl_check(nccl().Broadcast(d_reorderedRows, d_reorderedRows, (size_t)*numRows * 4, kNcclUint8, 0, g->comm, s),
                 "ncclBroadcast(reorderedRows)");
  }
  g->nRefRows = *numRows;
  g->refRows.alloc(*numRows ? *numRows : 1u, true);
  if (*numRows)
    SB_CUDA(cudaMemcpyAsync(g->refRows.get(), d_reorderedRows, (size_t)*numRows * 4, cudaMemcpyDeviceToDevice, s));
  g->nRefCols = distinct_values_dev(d_colIdx, nnz, nullptr, 0, N, g->refCols, s);
  g->cuts.assign((size_t)g->world + 1, 0);
  shard_plan_dev(d_rowOff, d_reorderedRows, *numRows, (u32)g->world, g->cuts.data(), s, &g->pre);
  if (h_cuts) std::memcpy(h_cuts, g->cuts.data(), g->cuts.size() * 4);
  *out = layout_build_dev(d_rowOff, d_colIdx, M, N, nnz, d_reorderedRows, *numRows, delta, g->cuts[g->rank],
                          g->cuts[g->rank + 1], flags, msColReorder, msRphm, s);
  API_END
}

int sddmm_mgpu_rebalance(sddmm_mgpu* g, const uint32_t* d_rowOff, const uint32_t* d_colIdx, uint32_t M, uint32_t N,
                         uint32_t nnz, const uint32_t* d_reorderedRows, uint32_t numRows, float delta, uint32_t flags,
                         float myMs, bsmr_layout** layout, uint32_t* h_cuts, void* stream) {
  API_BEGIN
  need(g && d_rowOff && d_colIdx && d_reorderedRows && layout && *layout && myMs >= 0.f, "arguments");
  need(g->cuts.size() == (size_t)g->world + 1 && !g->pre.empty(), "sddmm_mgpu_shard must come first");
  cudaStream_t s = (cudaStream_t)stream;
  std::vector<float> ms((size_t)g->world, 0.f);
  ms[g->rank] = myMs;
  if (g->world > 1) {  // every rank learns every rank's time: a sum over one-hot vectors
    DevBuf<float> d((size_t)g->world);
    SB_CUDA(cudaMemcpyAsync(d.get(), ms.data(), ms.size() * 4, cudaMemcpyHostToDevice, s));
    nccl_check(nccl().AllReduce(d.get(), d.get(), ms.size(), kNcclFloat32, kNcclSum, g->comm, s), "ncclAllReduce(times)");
    SB_CUDA(cudaMemcpyAsync(ms.data(), d.get(), ms.size() * 4, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaStreamSynchronize(s));
  }
  std::vector<u32> nc;
  rebalance_cuts(g->pre, g->cuts, ms, nc);
  if (h_cuts) std::memcpy(h_cuts, nc.data(), nc.size() * 4);
  if (nc != g->cuts) {
    bsmr_layout* fresh = layout_build_dev(d_rowOff, d_colIdx, M, N, nnz, d_reorderedRows, numRows, delta, nc[g->rank],
                                          nc[g->rank + 1], flags, nullptr, nullptr, s);
    bsmr_layout_destroy(*layout);
    *layout = fresh;
    g->cuts = nc;
  }
  API_END
}

int sddmm_mgpu_run(sddmm_mgpu* g, const bsmr_layout* L, uint32_t K, const float* d_A, const float* d_B, float* d_P,
                   void* stream) {
  // the steady state has no collective: a shard's pass is an ordinary pass over its own panel range
  if (!g) { sb::set_last_error("sddmm_mgpu_run: null handle"); return SDDMM_E_ARG; }
  return sddmm_run_dev(L, K, d_A, d_B, d_P, stream);
}

int sddmm_mgpu_run_host(sddmm_mgpu* g, const bsmr_layout* L, uint32_t K, const float* h_A, const float* h_B, float* h_P,
                        int root, float* msTotal) {
  API_BEGIN
  need(g && L && h_A && h_B && root >= 0 && root < g->world && (h_P || g->rank != root), "arguments");
  bsmr_layout_info I{};
  bsmr_layout_get_info(L, &I);
  const size_t W = (size_t)g->world;
  // every rank brings 1/world of A and of B over its own PCIe link; the slices meet over NVLink (all-gather)
  auto slice = [&](size_t n) { return (((n + W - 1) / W) + 3) & ~(size_t)3; };
  const size_t nA = (size_t)I.M * K, nB = (size_t)I.N * K, nP = I.nnz ? I.nnz : 1;
  const size_t cA = slice(nA), cB = slice(nB);
  if (!g->sCopy) {
    SB_CUDA(cudaStreamCreateWithFlags(&g->sCopy, cudaStreamNonBlocking));
    SB_CUDA(cudaStreamCreateWithFlags(&g->sWork, cudaStreamNonBlocking));
    for (cudaEvent_t* e : {&g->evA, &g->evB, &g->evDone}) SB_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    SB_CUDA(cudaEventCreate(&g->evT0));
    SB_CUDA(cudaEventCreate(&g->evT1));
  }
  if (g->wsA.size() < cA * W) { g->wsA.alloc(cA * W, true); SB_CUDA(cudaMemset(g->wsA.get(), 0, cA * W * 4)); }
  if (g->wsB.size() < cB * W) { g->wsB.alloc(cB * W, true); SB_CUDA(cudaMemset(g->wsB.get(), 0, cB * W * 4)); }
  if (g->wsP.size() < nP) {  // zeroed once: a shard's pass never writes foreign entries, so the sum below IS the merge
    g->wsP.alloc(nP, true);
    SB_CUDA(cudaMemset(g->wsP.get(), 0, nP * 4));
  }
  if (g->rank == root && g->wsPfull.size() < nP) g->wsPfull.alloc(nP, true);
  float *dA = g->wsA.get(), *dB = g->wsB.get(), *dP = g->wsP.get();
  const size_t r = (size_t)g->rank;
  auto h2d_slice = [&](float* d, const float* h, size_t n, size_t c) {
    const size_t beg = std::min(n, r * c), end = std::min(n, (r + 1) * c);
    if (end > beg) SB_CUDA(cudaMemcpyAsync(d + beg, h + beg, (end - beg) * 4, cudaMemcpyHostToDevice, g->sCopy));
  };
  SB_CUDA(cudaEventRecord(g->evT0, g->sCopy));
  // referenced rows only, when both host buffers are page-locked and the lists are clearly shorter than the arrays
  const void *mA = nullptr, *mB = nullptr;
  const char* h2dEnv = getenv("SDDMM_B200_H2D");
  const bool forceFull = h2dEnv && !strcmp(h2dEnv, "full"), forceGather = h2dEnv && !strcmp(h2dEnv, "gather");
  const bool refsKnown = g->world > 1 && (g->nRefRows || g->nRefCols) && !(K & 3u) && !forceFull;
  const bool sparseRefs = forceGather || ((size_t)g->nRefRows + g->nRefCols) * 100 <= ((size_t)I.M + I.N) * 85;
  if (refsKnown && sparseRefs && host_mapped(h_A, &mA) && host_mapped(h_B, &mB)) {
    const u32 K4 = K / 4;
    const size_t perA = (g->nRefRows + W - 1) / W, perB = (g->nRefCols + W - 1) / W;
    if (g->packA.size() < perA * W * K) g->packA.alloc(perA * W * K, true);
    if (g->packB.size() < perB * W * K) g->packB.alloc(perB * W * K, true);
    auto pack = [&](const u32* list, u32 n, size_t per, const void* hsrc, float* packed) {
      const u32 beg = (u32)std::min<size_t>(n, r * per), end = (u32)std::min<size_t>(n, (r + 1) * per);
      if (end > beg)
        k_pack_rows_h2d<<<128, 256, 0, g->sCopy>>>(list, beg, end, K4, static_cast<const float4*>(hsrc),
                                                   reinterpret_cast<float4*>(packed));
      SB_LAUNCH_CHECK();
    };
    auto unpack = [&](const u32* list, u32 n, const float* packed, float* dst) {
      if (n) k_unpack_rows<<<grid_for((size_t)n * 32), 256, 0, g->sWork>>>(list, n, K4,
                                                                          reinterpret_cast<const float4*>(packed),
                                                                          reinterpret_cast<float4*>(dst));
      SB_LAUNCH_CHECK();
    };
    pack(g->refRows.get(), g->nRefRows, perA, mA, g->packA.get());
    SB_CUDA(cudaEventRecord(g->evA, g->sCopy));
    pack(g->refCols.get(), g->nRefCols, perB, mB, g->packB.get());
    SB_CUDA(cudaEventRecord(g->evB, g->sCopy));
    SB_CUDA(cudaStreamWaitEvent(g->sWork, g->evA, 0));
    nccl_check(nccl().AllGather(g->packA.get() + r * perA * K, g->packA.get(), perA * K, kNcclFloat32, g->comm, g->sWork),
               "ncclAllGather(A rows)");
    unpack(g->refRows.get(), g->nRefRows, g->packA.get(), dA);
    SB_CUDA(cudaStreamWaitEvent(g->sWork, g->evB, 0));
    nccl_check(nccl().AllGather(g->packB.get() + r * perB * K, g->packB.get(), perB * K, kNcclFloat32, g->comm, g->sWork),
               "ncclAllGather(B rows)");
    unpack(g->refCols.get(), g->nRefCols, g->packB.get(), dB);
    g->lastH2DBytes = (unsigned long long)(std::min<size_t>(g->nRefRows, (r + 1) * perA) - std::min<size_t>(g->nRefRows, r * perA) +
                                           std::min<size_t>(g->nRefCols, (r + 1) * perB) - std::min<size_t>(g->nRefCols, r * perB)) * K * 4;
  } else {
  h2d_slice(dA, h_A, nA, cA);
  SB_CUDA(cudaEventRecord(g->evA, g->sCopy));
  h2d_slice(dB, h_B, nB, cB);
  SB_CUDA(cudaEventRecord(g->evB, g->sCopy));
  SB_CUDA(cudaStreamWaitEvent(g->sWork, g->evA, 0));
  if (g->world > 1) nccl_check(nccl().AllGather(dA + r * cA, dA, cA, kNcclFloat32, g->comm, g->sWork), "ncclAllGather(A)");
  SB_CUDA(cudaStreamWaitEvent(g->sWork, g->evB, 0));
  if (g->world > 1) nccl_check(nccl().AllGather(dB + r * cB, dB, cB, kNcclFloat32, g->comm, g->sWork), "ncclAllGather(B)");
    g->lastH2DBytes = (unsigned long long)((std::min(nA, (r + 1) * cA) - std::min(nA, r * cA)) +
                                           (std::min(nB, (r + 1) * cB) - std::min(nB, r * cB))) * 4;
  }
  {
    const int rc = sddmm_run_dev(L, K, dA, dB, dP, g->sWork);
    if (rc) return rc;
  }
  float* out = dP;
  if (g->world > 1) {
    out = g->rank == root ? g->wsPfull.get() : nullptr;
    nccl_check(nccl().Reduce(dP, out, (size_t)I.nnz, kNcclFloat32, kNcclSum, root, g->comm, g->sWork), "ncclReduce(P)");
  }
  if (g->rank == root && I.nnz) SB_CUDA(cudaMemcpyAsync(h_P, out, (size_t)I.nnz * 4, cudaMemcpyDeviceToHost, g->sWork));
  SB_CUDA(cudaEventRecord(g->evDone, g->sWork));
  SB_CUDA(cudaStreamWaitEvent(g->sCopy, g->evDone, 0));
  SB_CUDA(cudaEventRecord(g->evT1, g->sCopy));
  SB_CUDA(cudaEventSynchronize(g->evT1));
  if (msTotal) SB_CUDA(cudaEventElapsedTime(msTotal, g->evT0, g->evT1));
  API_END
}

int sddmm_mgpu_host_traffic(const sddmm_mgpu* g, uint64_t* h2dBytes) {
  API_BEGIN
  need(g && h2dBytes, "arguments");
  *h2dBytes = g->lastH2DBytes;
  API_END
}

int sddmm_mgpu_gather(sddmm_mgpu* g, float* d_P, size_t count, void* stream) {
  API_BEGIN
  need(g && (d_P || count == 0), "arguments");
  // shards write disjoint CSR positions and leave the rest untouched: with P zero-initialised a sum IS the merge
  if (g->world > 1 && count)
    nccl_check(nccl().AllReduce(d_P, d_P, count, kNcclFloat32, kNcclSum, g->comm, (cudaStream_t)stream), "ncclAllReduce");
  API_END
}

int bsmr_rebalance_cuts(const uint64_t* h_panelNnzPrefix, uint32_t numPanels, const uint32_t* h_oldCuts,
                        const float* h_ms, uint32_t numShards, uint32_t* h_newCuts) {
  API_BEGIN
  need(h_panelNnzPrefix && h_oldCuts && h_ms && h_newCuts && numShards > 0, "arguments");
  need(h_oldCuts[0] == 0 && h_oldCuts[numShards] == numPanels, "old cuts must span [0, numPanels]");
  for (uint32_t r = 0; r < numShards; ++r) need(h_oldCuts[r] <= h_oldCuts[r + 1] && h_ms[r] >= 0.f, "old cuts / times");
  std::vector<u64> pre(h_panelNnzPrefix, h_panelNnzPrefix + (size_t)numPanels + 1);
  std::vector<u32> oc(h_oldCuts, h_oldCuts + (size_t)numShards + 1), nc;
  std::vector<float> ms(h_ms, h_ms + numShards);
  rebalance_cuts(pre, oc, ms, nc);
  std::memcpy(h_newCuts, nc.data(), nc.size() * 4);
  API_END
}

int bsmr_shard_plan_dev(const uint32_t* d_rowOff, const uint32_t* d_reorderedRows, uint32_t numRows, uint32_t numShards,
                        uint32_t* h_cuts, void* stream) {
  API_BEGIN
  need(d_rowOff && (d_reorderedRows || numRows == 0) && h_cuts && numShards > 0, "arguments");
  shard_plan_dev(d_rowOff, d_reorderedRows, numRows, numShards, h_cuts, (cudaStream_t)stream);
  API_END
}

}  // extern "C"
