// layout.cuh -- the device-resident BSMR + RPHM layout object behind `bsmr_layout` (C ABI).
#pragma once
#include <map>
#include <memory>

#include "common.cuh"

namespace sb {
// Private residual layout of the "super-panel" kernel (K7b): G consecutive row panels form a super-panel
// whose A rows (16*G x K floats) fit in shared memory; its residual entries are sorted by (column, row)
// so that one B^T row fetched from L2 is reused from registers by every entry of that column.
struct SuperPanelLayout {
  u32 G = 0, rows = 0, numSp = 0, segLen = 0, numWork = 0, numEntries = 0, numRuns = 0;
  // L2 residency classes (graphs, B >> L2): the `hubCols` most referenced columns -- as many B^T rows as may stay
  // in the L2 -- keep bit 31 of `col` clear and are loaded evict_last; every other ("tail") column has bit 31 set
  // and streams through evict_first, so one-time rows cannot wash the hub rows out.  colMask strips the flag
  // (0xFFFFFFFF = unflagged layout).
  u32 hubCols = 0, hubMinDegree = 0, colMask = 0xFFFFFFFFu;
  u64 hubEntries = 0;
  DevBuf<u32> off;              // [numSp+1] entry ranges per super-panel
  DevBuf<u32> col, idx;         // per entry: column, CSR index
  DevBuf<unsigned short> row;   // per entry: row inside the super-panel
  DevBuf<uint2> work;           // (super-panel, first entry of the segment inside it), segment-major
  // distinct columns the residual references, ascending: the only B^T rows the fp16 copy of SDDMM_OPERANDS_FP16 has
  // to rewrite per pass (half of the columns of an R-MAT graph are empty)
  u32 numUsedCols = 0;
  DevBuf<u32> usedCols;
};
}  // namespace sb

namespace sb {
// Private residual layout of the row-stream kernel (K7c): the residual entries in reordered-ROW order (then
// column), so that the entries of one row are consecutive and its A fragment stays in registers.
struct StreamLayout {
  u32 numEntries = 0;
  DevBuf<u32> row;  // per entry: ORIGINAL row id (index into A)
  DevBuf<u32> col;  // per entry: column (index into B^T)
  DevBuf<u32> idx;  // per entry: CSR index (index into P)
};
}  // namespace sb

namespace sb {
// Private layout of the "full tile" plan (K8): the reordered matrix is cut into 128-row x 128-column tiles
// (rows in BSMR order, columns in natural order); every non-empty tile is one tcgen05 GEMM tile whose
// epilogue picks the stored entries with per-row bitmasks.  Used when S is dense enough that computing
// whole tiles on the tensor cores beats gathering per non-zero (DLMC-style masks), see sddmm_launch.
struct TileLayout {
  u32 numTiles = 0, numEntries = 0, tileRows = 0, tileCols = 0;
  DevBuf<uint4> tiles;     // per non-empty tile: {tile row, tile col, first entry, entry count}
  DevBuf<u32> rowMeta;     // per tile 128 x 5 words: 4 mask words (which of the 128 columns are stored) + entry offset
  DevBuf<u32> idx;         // per entry (sorted by tile, row, col): CSR index
  // 2x2 groups of tiles for the cluster form of the tile kernel (operands multicast inside a 4-CTA cluster):
  // quads[q] = {quad row, quad col}; quadTiles[4q + 2*dr + dc] = index into tiles[] or 0xFFFFFFFF (no entries)
  u32 numQuads = 0;
  DevBuf<uint2> quads;
  DevBuf<u32> quadTiles;
  // rowMeta transposed for the CTA-pair kernel, whose epilogue warp (row quarter, 32-column slice cq) wants ONE
  // coalesced 8-byte load per thread: rowMetaT[(tile * 4 + cq) * 128 + row] = {mask word cq of the row, offset of the
  // row's first stored entry inside slice cq relative to the tile's first entry}
  DevBuf<uint2> rowMetaT;
};
void build_quads(TileLayout& T);  // from T.tiles (host pass over the tile list)
}  // namespace sb

struct bsmr_layout {
  bsmr_layout_info info{};
  sb::DevBuf<sb::u32> arr[BSMR_ARRAY_COUNT];  // indexed by bsmr_array_id
  // work lists consumed by OUR kernels (the reference-shaped ones above are kept for API parity)
  sb::DevBuf<uint2> denseWork;   // (local panel, first dense block of the group inside the panel)
  sb::DevBuf<uint2> sparseWork;  // (local panel, first residual entry of the chunk inside the panel)
  sb::u32 numDenseWork = 0, numSparseWork = 0;
  sb::u32 sparseChunk = 1024;  // residual entries per sparse work item
  int device = 0;
  // device staging buffers of the host-buffer entry point (sddmm_run_host), grown on demand and kept
  // so that repeated calls do not pay cudaMalloc/cudaFree
  mutable sb::DevBuf<float> wsA, wsB, wsP;
  // TMA forms (tile plan and BSMR dense blocks): TF32-rounded, row-gathered copies of the operands (rewritten every
  // pass) and the tensor maps describing them.  One workspace per (K, numBatch), kept for the layout's lifetime
  // (sddmm_prepare builds it ahead of the first run).  `busy` serialises passes that share the workspace from
  // different streams: a pass waits for it before rounding and records it after its last reader.
  struct TileTma {
    sb::u32 K = 0, numBatch = 0;
    sb::DevBuf<float> rA, rB;                  // [numBatch][numRows][K], [numBatch][N][K]
    alignas(64) unsigned char mapA[128], mapB[128];      // CUtensorMap (box 128 rows), kept opaque here
    alignas(64) unsigned char mapA64[128], mapB64[128];  // box 64 rows: the halves a cluster member multicasts
    cudaEvent_t busy = nullptr;
    ~TileTma() { if (busy) cudaEventDestroy(busy); }
  };
  mutable std::map<sb::u64, std::unique_ptr<TileTma>> tma;              // key = K << 32 | numBatch
  // TMA form of the BSMR dense-block kernel.  K-independent index (built once): the distinct columns that
  // appear in denseCols and the panels that own dense blocks, each numbered compactly, so that the rounded
  // copies hold ONLY the rows dense blocks touch (a graph's dense part uses a sliver of B).
  struct DenseIndex {
    sb::u32 numCols = 0, numPanels = 0;
    sb::DevBuf<sb::u32> colList;      // [numCols]   distinct dense columns, ascending
    sb::DevBuf<sb::u32> colCompact;   // [|denseCols|] compact id of denseCols[i] (sentinel N -> 0)
    sb::DevBuf<sb::u32> panelList;    // [numPanels] local panels with >= 1 dense block, ascending
    sb::DevBuf<sb::u32> workRowA;     // [numDenseWork] first compact A row of the work item's panel
  };
  mutable std::unique_ptr<DenseIndex> dix;
  struct DenseTma {
    sb::u32 K = 0, numBatch = 0;
    sb::DevBuf<float> rA, rB;             // [numBatch][16*numPanels][K], [numBatch][numCols][K], TF32-rounded
    alignas(64) unsigned char mapA16[128];  // 3D, box 32 floats x 16 rows: one row panel
    alignas(64) unsigned char mapBg[128];   // 2D, box 32 floats x 1 row: tile::gather4 of four rows per instruction
    cudaEvent_t busy = nullptr;
    ~DenseTma() { if (busy) cudaEventDestroy(busy); }
  };
  mutable std::map<sb::u64, std::unique_ptr<DenseTma>> dtma;            // key = K << 32 | numBatch
  // fp16 copy of B^T for the super-panel residual kernel under SDDMM_OPERANDS_FP16 (rewritten every pass)
  struct HalfB {
    sb::u32 K = 0, numBatch = 0;
    sb::DevBuf<float> rB;  // [numBatch][N][K] halves
    cudaEvent_t busy = nullptr;
    ~HalfB() { if (busy) cudaEventDestroy(busy); }
  };
  mutable std::map<sb::u64, std::unique_ptr<HalfB>> halfB;              // key = K << 32 | numBatch
  mutable std::map<sb::u64, std::unique_ptr<sb::SuperPanelLayout>> sp;  // key = G | hub budget << 32
  mutable std::unique_ptr<sb::StreamLayout> st;                         // K-independent, built on first use
  std::unique_ptr<sb::TileLayout> tl;                // built with the layout when S is dense enough to consider it
  // two-slot pipeline of sddmm_run_host_async
  struct HostPipe {
    cudaStream_t h2d = nullptr, comp = nullptr, d2h = nullptr;
    cudaEvent_t evH2D[2] = {nullptr, nullptr}, evComp[2] = {nullptr, nullptr}, evD2H[2] = {nullptr, nullptr};
    sb::DevBuf<float> A[2], B[2], P[2];
    ~HostPipe() {
      for (int i = 0; i < 2; ++i) {
        if (evH2D[i]) cudaEventDestroy(evH2D[i]);
        if (evComp[i]) cudaEventDestroy(evComp[i]);
        if (evD2H[i]) cudaEventDestroy(evD2H[i]);
      }
      if (h2d) cudaStreamDestroy(h2d);
      if (comp) cudaStreamDestroy(comp);
      if (d2h) cudaStreamDestroy(d2h);
    }
  };
  mutable std::unique_ptr<HostPipe> pipe;
  // host-buffer entry points: the distinct columns this layout references (dense blocks + residual), ascending, so
  // that a pass moves only the B^T rows it reads over PCIe (the A rows it reads are `reorderedRows`); built on first use
  struct HostRefs {
    sb::u32 numCols = 0;
    sb::DevBuf<sb::u32> cols;
  };
  mutable std::unique_ptr<HostRefs> hostRefs;
  mutable unsigned long long lastH2DBytes = 0, lastD2HBytes = 0;  // of the most recent host-buffer pass
};

namespace sb {

constexpr u32 kDenseGroupBlocks = 8;   // 8 x 16 = 128 gathered B columns per tcgen05 tile (MMA M)
constexpr u32 kSparseChunkDefault = 2048;  // residual entries per CTA work item (env SDDMM_B200_CHUNK overrides)

bsmr_layout* layout_build_dev(const u32* d_rowOff, const u32* d_colIdx, u32 M, u32 N, u32 nnz,
                              const u32* d_reorderedRows, u32 numRows, float delta, u32 panelBegin, u32 panelEnd,
                              u32 tileFlags, float* msCol, float* msRphm, cudaStream_t s);

void layout_save(const bsmr_layout* L, const char* path);
bsmr_layout* layout_load(const char* path);

// builds (once, then cached) the super-panel layout for G panels per super-panel; returns it
// hubBudget: number of B^T rows that may be kept L2-resident (0 = no residency classes)
const SuperPanelLayout* ensure_superpanels(const bsmr_layout* L, u32 G, u32 hubBudget, cudaStream_t s);
// the distinct values < N of up to two u32 device lists, ascending, into `out`; returns their number
u32 distinct_values_dev(const u32* a, size_t na, const u32* b, size_t nb, u32 N, DevBuf<u32>& out, cudaStream_t s);
// builds (once) the list of columns the layout references (host-buffer passes copy only those B^T rows)
const bsmr_layout::HostRefs* ensure_host_refs(const bsmr_layout* L, cudaStream_t s);
// builds (once) the row-ordered residual layout of the row-stream kernel
const StreamLayout* ensure_stream(const bsmr_layout* L, cudaStream_t s);

}  // namespace sb
