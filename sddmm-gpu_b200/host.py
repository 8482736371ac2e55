"""Python host-side mirror of the reference's operator interface for the hot path.

Names and argument meaning follow the reference (include/BSMR.hpp:21-63, :79-159; include/sddmm.hpp:8-21;
include/sddmmKernel.cuh:19-51); every method forwards to the C ABI of include/sddmm_b200.h.  The C++
mirror of the same interface lives in csrc/host/ (that is what a C++ caller links); this module is what
tests/ and bench.py drive.  numpy arrays = host buffers, torch CUDA tensors = device buffers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import ARRAY_IDS, LayoutInfo, Plan, ReorderOpts, Stats, check

NULL_VALUE = 0xFFFFFFFF
ROW_PANEL_SIZE = 16
BLOCK_COL_SIZE = 16


def _np_u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _np_f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a):
    """host numpy array or torch tensor -> raw pointer"""
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()


def make_plan(plan="auto", dense="auto", residual="auto", tile="auto", tile_stages=0, operands="exact"):
    """sddmm_plan by name: plan auto|bsmr|tile, dense auto|reg|tma, residual auto|panel|superpanel|stream,
    tile auto|reg|tma|tma_cluster.  None where the library's defaults (environment, cost model) should decide."""
    return Plan(_lib.PLAN[plan], _lib.DENSE[dense], _lib.RESIDUAL[residual], _lib.TILE[tile], int(tile_stages),
                _lib.OPERANDS[operands])


def make_reorder_opts(kernel="auto", batch=0, lane_rows="auto", signature="auto"):
    return ReorderOpts(_lib.CLUSTER[kernel], int(batch), _lib.TRISTATE[lane_rows], _lib.TRISTATE[signature])


def _plan_ptr(plan):
    return C.byref(plan) if plan is not None else None


def plan_resolve(layout, K, numBatch=1, plan=None):
    """The kernels a pass with `plan` will launch on this layout, as a dict of names."""
    lay = layout.layout() if hasattr(layout, "layout") else layout
    out = Plan()
    check(_lib.lib().sddmm_plan_resolve(lay.handle, int(K), int(numBatch), _plan_ptr(plan), C.byref(out)))
    inv = lambda d, v: next(k for k, x in d.items() if x == v)
    return dict(plan=inv(_lib.PLAN, out.plan), dense=inv(_lib.DENSE, out.dense), residual=inv(_lib.RESIDUAL, out.residual),
                tile=inv(_lib.TILE, out.tile), tile_stages=int(out.tileStages), operands=inv(_lib.OPERANDS, out.operands))


def sddmm_prepare(layout, K, numBatch=1, plan=None):
    """Front-loads every K-dependent private layout / workspace so that later passes only enqueue kernels."""
    lay = layout.layout() if hasattr(layout, "layout") else layout
    check(_lib.lib().sddmm_prepare(lay.handle, int(K), int(numBatch), _plan_ptr(plan)))


def calculateBlockSize(S, free_mem_bytes=0):
    """calculateBlockSize(matrix)  src/rowReordering.cu:1009-1025 (free memory injectable)."""
    return int(_lib.lib().bsmr_calc_block_size(S.M, S.N, int(free_mem_bytes)))


def launch_count(reset=False):
    L = _lib.lib()
    n = int(L.sddmm_launch_count())
    if reset:
        L.sddmm_launch_count_reset()
    return n


# ------------------------------------------------------------------------------------------------
class Layout:
    """Owns a `bsmr_layout*` (device-resident BSMR + RPHM arrays)."""

    def __init__(self, handle):
        self._h = C.c_void_p(handle)
        info = LayoutInfo()
        check(_lib.lib().bsmr_layout_get_info(self._h, C.byref(info)))
        self.info = info

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None and getattr(_lib, "_lib", None) is not None:  # not during interpreter teardown
            _lib._lib.bsmr_layout_destroy(h)

    @property
    def handle(self):
        return self._h

    def array(self, name) -> np.ndarray:
        L = _lib.lib()
        aid = ARRAY_IDS[name]
        n = int(L.bsmr_layout_array_len(self._h, aid))
        out = np.zeros(max(1, n), dtype=np.uint32)
        check(L.bsmr_layout_array_to_host(self._h, aid, out.ctypes.data, n))
        return out[:n]

    def arrays(self):
        return {k: self.array(k) for k in ARRAY_IDS}

    def save(self, path):
        """persist the layout (reorder cache)"""
        check(_lib.lib().bsmr_layout_save(self._h, str(path).encode()))

    @staticmethod
    def load(path):
        h = C.c_void_p()
        check(_lib.lib().bsmr_layout_load(str(path).encode(), C.byref(h)))
        return Layout(h.value)

    def array_dev_ptr(self, name):
        return int(_lib.lib().bsmr_layout_array_dev(self._h, ARRAY_IDS[name]) or 0)


# ------------------------------------------------------------------------------------------------
class BSMR:
    """BSMR(similarityThreshold, blockDensityThreshold, matrix)   include/BSMR.hpp:21-63."""

    def __init__(self, alpha=None, delta=None, S=None, block_size=0):
        self._S = None
        self._layout = None
        self.numClusters_ = 1
        self.rowReorderingTime_ = 0.0
        self.colReorderingTime_ = 0.0
        self.rphmTime_ = 0.0
        self.reorderedRows_ = np.zeros(0, np.uint32)
        self.blockSize_ = block_size
        if S is not None:
            self.rowReordering(alpha, S, block_size=block_size)
            self.colReordering(delta, S)

    # BSMR::rowReordering  src/BSMR.cpp:27-50
    def rowReordering(self, alpha, S, numIterations=1, block_size=0, opts=None):
        L = _lib.lib()
        bs = block_size or self.blockSize_ or calculateBlockSize(S)
        out = np.zeros(max(1, S.M), dtype=np.uint32)
        n, ncl, ms = C.c_uint32(0), C.c_int32(0), C.c_float(0)
        tot = 0.0
        for _ in range(numIterations):
            check(L.bsmr_row_reorder_ex(S.row_off.ctypes.data, S.col_idx.ctypes.data, S.M, S.N, S.nnz, float(alpha), bs,
                                        C.byref(opts) if opts is not None else None, out.ctypes.data, C.byref(n),
                                        C.byref(ncl), C.byref(ms)))
            tot += ms.value
        self.reorderedRows_ = out[: n.value].copy()
        self.numClusters_ = ncl.value
        self.rowReorderingTime_ = tot / numIterations
        self.blockSize_ = bs
        self._S = S
        return self

    # BSMR::colReordering  src/BSMR.cpp:52-81  (+ the RPHM arrays, built by the same device pass)
    def colReordering(self, delta, S, reorderedRows=None, numIterations=1, tiles="auto"):
        L = _lib.lib()
        if reorderedRows is not None and len(reorderedRows):
            self.reorderedRows_ = _np_u32(reorderedRows)
        R = self.reorderedRows_
        h = C.c_void_p()
        msC, msR = C.c_float(0), C.c_float(0)
        check(L.bsmr_layout_build_ex(S.row_off.ctypes.data, S.col_idx.ctypes.data, S.M, S.N, S.nnz,
                                     R.ctypes.data if R.size else None, R.size, float(delta), _lib.BUILD_TILES[tiles],
                                     C.byref(h), C.byref(msC), C.byref(msR)))
        self._layout = Layout(h.value)
        self.colReorderingTime_ = msC.value
        self.rphmTime_ = msR.value
        self._S = S
        return self

    def numRowPanels(self): return int(self._layout.info.numRowPanels) if self._layout else (len(self.reorderedRows_) + 15) // 16
    def reorderedRows(self): return self.reorderedRows_
    def denseCols(self): return self._layout.array("denseCols")
    def denseColOffsets(self): return self._layout.array("denseColOffsets")
    def sparseCols(self): return self._layout.array("sparseCols")
    def sparseColOffsets(self): return self._layout.array("sparseColOffsets")
    def sparseValueOffsets(self): return self._layout.array("sparseValueOffsets")
    def numClusters(self): return self.numClusters_
    def rowReorderingTime(self): return self.rowReorderingTime_
    def colReorderingTime(self): return self.colReorderingTime_
    def reorderingTime(self): return self.rowReorderingTime_ + self.colReorderingTime_
    def layout(self): return self._layout


class RPHM:
    """RPHM(matrix, bsmr)   include/BSMR.hpp:79-159.  The device arrays already exist inside the
    layout object BSMR::colReordering built; this class only exposes them under the reference's names."""

    def __init__(self, S, bsmr: BSMR):
        self._l = bsmr.layout()
        self._S = S

    def layout(self): return self._l
    def numRowPanels(self): return int(self._l.info.numRowPanels)
    def maxNumDenseColBlocksInRowPanel(self): return int(self._l.info.maxNumDenseColBlocksInRowPanel)
    def maxNumSparseColBlocksInRowPanel(self): return int(self._l.info.maxNumSparseColBlocksInRowPanel)
    def numDenseThreadBlocks(self): return int(self._l.info.numDenseThreadBlocks)
    def numSparseThreadBlocks(self): return int(self._l.info.numSparseThreadBlocks)
    def getNumDenseBlocks(self): return int(self._l.info.numDenseBlocks)
    def reorderedRows(self): return self._l.array("reorderedRows")
    def denseCols(self): return self._l.array("denseCols")
    def blockValues(self): return self._l.array("blockValues")
    def blockOffsets(self): return self._l.array("blockOffsets")
    def sparseValueOffsets(self): return self._l.array("sparseValueOffsets")
    def sparseValues(self): return self._l.array("sparseValues")
    def sparseRelativeRows(self): return self._l.array("sparseRelativeRows")
    def sparseColIndices(self): return self._l.array("sparseColIndices")
    def denseRowPanelIds(self): return self._l.array("denseRowPanelIds")
    def denseColBlockIters(self): return self._l.array("denseColBlockIters")
    def sparseRowPanelIds(self): return self._l.array("sparseRowPanelIds")
    def sparseColBlockIters(self): return self._l.array("sparseColBlockIters")


# ------------------------------------------------------------------------------------------------
def sddmm_gpu(A, B, rphm_or_layout, P=None, plan=None):
    """sddmm_gpu(matrixA, matrixB, rphm, matrixP, logger)  src/sddmmKernel.cu:2518-2537 (host buffers:
    H2D of A and B, one pass, D2H of P)  -- or the raw device-pointer overload :2539-2663 when A, B, P
    are torch CUDA tensors.  Returns (P, ms)."""
    L = _lib.lib()
    lay = rphm_or_layout.layout() if hasattr(rphm_or_layout, "layout") else rphm_or_layout
    if isinstance(A, np.ndarray):
        A, B = _np_f32(A), _np_f32(B)
        K = A.shape[1]
        if P is None:
            P = np.zeros(max(1, lay.info.nnz), dtype=np.float32)
        ms = C.c_float(0)
        check(L.sddmm_run_host(lay.handle, K, A.ctypes.data, B.ctypes.data, P.ctypes.data, C.byref(ms)))
        return P[: lay.info.nnz], ms.value
    import torch

    K = A.shape[1]
    if P is None:
        P = torch.zeros(max(1, lay.info.nnz), dtype=torch.float32, device=A.device)
    stream = torch.cuda.current_stream().cuda_stream
    check(L.sddmm_run_dev_ex(lay.handle, K, 1, A.data_ptr(), B.data_ptr(), P.data_ptr(), _plan_ptr(plan),
                             C.c_void_p(stream)))
    return P, None


def sddmm_gpu_batch(A, B, layout, P=None, plan=None):
    """sddmm_gpu_batch(numBatch, M, N, K, nnz, dA, dB, rphm, dP, time)  src/sddmmKernel.cu:2764-2850:
    A [numBatch, M, K], B [numBatch, N, K], P [numBatch, nnz] torch CUDA tensors; one layout for all."""
    import torch

    lay = layout.layout() if hasattr(layout, "layout") else layout
    nb, _, K = A.shape
    if P is None:
        P = torch.zeros((nb, max(1, lay.info.nnz)), dtype=torch.float32, device=A.device)
    stream = torch.cuda.current_stream().cuda_stream
    check(_lib.lib().sddmm_run_dev_ex(lay.handle, K, nb, A.data_ptr(), B.data_ptr(), P.data_ptr(), _plan_ptr(plan),
                                      C.c_void_p(stream)))
    return P


def sddmm_gpu_async(A, B, layout, P, slot):
    """Streaming twin of the host-buffer sddmm_gpu: enqueue H2D(A,B) -> pass -> D2H(P) on `slot` (0/1) and
    return immediately; `sddmm_gpu_sync(layout)` waits.  A, B, P: page-locked numpy arrays."""
    lay = layout.layout() if hasattr(layout, "layout") else layout
    check(_lib.lib().sddmm_run_host_async(lay.handle, A.shape[1], A.ctypes.data, B.ctypes.data, P.ctypes.data, int(slot)))


def host_traffic(lay):
    """(h2d_bytes, d2h_bytes) of the most recent host-buffer pass on this layout (sddmm_host_traffic)."""
    a, b = C.c_uint64(0), C.c_uint64(0)
    check(_lib.lib().sddmm_host_traffic(lay.handle, C.byref(a), C.byref(b)))
    return int(a.value), int(b.value)


def sddmm_gpu_sync(layout):
    lay = layout.layout() if hasattr(layout, "layout") else layout
    check(_lib.lib().sddmm_host_sync(lay.handle))


def sddmm_gpu_timed(A, B, layout, P, warmup=3, iters=10):
    """Mean device milliseconds per pass (dense, residual, both concurrently) -- the reference's
    `sddmmTime_` loop (src/sddmmKernel.cu:2561-2659) with a warm-up."""
    L = _lib.lib()
    lay = layout.layout() if hasattr(layout, "layout") else layout
    d, s, t = C.c_float(0), C.c_float(0), C.c_float(0)
    check(L.sddmm_run_timed_dev(lay.handle, A.shape[1], A.data_ptr(), B.data_ptr(), P.data_ptr(), warmup, iters,
                                C.byref(d), C.byref(s), C.byref(t)))
    return dict(dense_ms=d.value, sparse_ms=s.value, total_ms=t.value)


def sddmm(S, A, B, alpha=0.3, delta=0.3, block_size=0, keep_layout=True):
    """sddmm(options, A, B, P, logger)  src/sddmm.cu:10-39: reorder -> layout -> one SDDMM, host buffers.
    Returns a dict with P, every BSMR/RPHM array, timings and the number of kernels launched."""
    L = _lib.lib()
    A, B = _np_f32(A), _np_f32(B)
    P = np.zeros(max(1, S.nnz), dtype=np.float32)
    st = Stats()
    h = C.c_void_p()
    L.sddmm_launch_count_reset()
    check(L.sddmm_host(S.row_off.ctypes.data, S.col_idx.ctypes.data, S.M, S.N, S.nnz, A.shape[1], A.ctypes.data,
                       B.ctypes.data, float(alpha), float(delta), int(block_size), P.ctypes.data, C.byref(st),
                       C.byref(h)))
    launches = int(L.sddmm_launch_count())
    lay = Layout(h.value)
    out = dict(P=P[: S.nnz], block_size=int(st.blockSize), numClusters=int(st.numClusters),
               rowReorderMs=st.rowReorderMs, colReorderMs=st.colReorderMs, rphmMs=st.rphmMs, sddmmMs=st.sddmmMs,
               gpu_launches=launches, numDenseBlocks=int(lay.info.numDenseBlocks),
               numSparseValues=int(lay.info.numSparseValues), numDenseValues=int(lay.info.numDenseValues),
               numRowPanels=int(lay.info.numRowPanels))
    out.update(lay.arrays())
    if keep_layout:
        out["layout"] = lay
    return out


# ------------------------------------------------------------------------------------------------
# device-resident forms (torch CUDA tensors)
def row_reorder_dev(row_off_t, col_idx_t, M, N, alpha, block_size=0, opts=None):
    import torch

    L = _lib.lib()
    out = torch.empty(max(1, M), dtype=torch.int32, device=row_off_t.device)
    n, ncl, ms = C.c_uint32(0), C.c_int32(0), C.c_float(0)
    stream = torch.cuda.current_stream().cuda_stream
    check(L.bsmr_row_reorder_dev_ex(row_off_t.data_ptr(), col_idx_t.data_ptr(), M, N, col_idx_t.numel(), float(alpha),
                                    int(block_size), C.byref(opts) if opts is not None else None, out.data_ptr(),
                                    C.byref(n), C.byref(ncl), C.byref(ms), C.c_void_p(stream)))
    return out[: n.value], ncl.value, ms.value


def layout_build_dev(row_off_t, col_idx_t, M, N, reordered_rows_t, delta, panel_begin=0, panel_end=0xFFFFFFFF,
                     tiles="auto"):
    import torch

    L = _lib.lib()
    h = C.c_void_p()
    msC, msR = C.c_float(0), C.c_float(0)
    stream = torch.cuda.current_stream().cuda_stream
    check(L.bsmr_layout_build_dev_ex(row_off_t.data_ptr(), col_idx_t.data_ptr(), M, N, col_idx_t.numel(),
                                     reordered_rows_t.data_ptr(), reordered_rows_t.numel(), float(delta), panel_begin,
                                     panel_end, _lib.BUILD_TILES[tiles], C.byref(h), C.byref(msC), C.byref(msR),
                                     C.c_void_p(stream)))
    return Layout(h.value), msC.value, msR.value


def dispersion_dev(row_off_t, col_idx_t, M, N, block_size):
    import torch

    L = _lib.lib()
    out = torch.empty(max(1, M), dtype=torch.int32, device=row_off_t.device)
    nb = C.c_uint32(0)
    check(L.bsmr_dispersion_dev(row_off_t.data_ptr(), col_idx_t.data_ptr(), M, N, col_idx_t.numel(), int(block_size),
                                out.data_ptr(), C.byref(nb), None))
    return out[:M], nb.value


def evaluationReordering(S, layout, delta):
    """evaluationReordering(matrix, bsmr, logger)  src/BSMR.cpp:826-925 plus the original-matrix statistics
    (:953-994): the dict of values the reference logs, computed on the device."""
    L = _lib.lib()
    lay = layout.layout() if hasattr(layout, "layout") else layout
    ev = _lib.Eval()
    check(L.bsmr_layout_eval(lay.handle, float(delta), C.byref(ev)))
    nd, ad = C.c_uint32(0), C.c_float(0)
    check(L.bsmr_original_block_stats(S.row_off.ctypes.data, S.col_idx.ctypes.data, S.M, S.N, S.nnz, float(delta),
                                      C.byref(nd), C.byref(ad)))
    return dict(numDenseBlock=int(ev.numDenseBlock), averageDensity=float(ev.averageDensity),
                numDenseThreadBlocks=int(ev.numDenseThreadBlocks), numSparseThreadBlocks=int(ev.numSparseThreadBlocks),
                numDenseData=int(ev.numDenseData), numSparseData=int(ev.numSparseData),
                originalNumDenseBlock=int(nd.value), originalAverageDensity=float(ad.value))


def shard_plan(S, reordered_rows, num_shards):
    L = _lib.lib()
    R = _np_u32(reordered_rows)
    cuts = np.zeros(num_shards + 1, dtype=np.uint32)
    check(L.bsmr_shard_plan(S.row_off.ctypes.data, R.ctypes.data, R.size, num_shards, cuts.ctypes.data))
    return cuts
