"""ctypes door onto oracle/liboracle.so and oracle/_ref/libref_cpu.so.

TEST INFRASTRUCTURE ONLY (see oracle/bsmr_oracle.h).  May be imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by the
product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NULL_VALUE = 0xFFFFFFFF

_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """Compile the C restatement and, when /root/reference is present, the reference pieces."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref and os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-s", "-j8", "-C", HERE, "ref_cpu", "ref_gpu"], check=True)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = C.CDLL(path)
        L.oracle_block_size.restype = C.c_uint32
        L.oracle_block_size.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
        L.oracle_num_blocks_per_row.restype = C.c_uint32
        L.oracle_num_blocks_per_row.argtypes = [C.c_uint32, C.c_uint32]
        L.oracle_cluster_blockdim.restype = C.c_uint32
        L.oracle_cluster_blockdim.argtypes = [C.c_uint32]
        L.oracle_kept_warps.argtypes = [C.c_uint32, C.c_void_p]
        L.oracle_encode_dispersion.argtypes = [_u32p, _u32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, _u32p]
        L.oracle_similarity.restype = C.c_float
        L.oracle_similarity.argtypes = [_u32p, _u32p, C.c_uint32, C.c_uint32]
        L.oracle_row_reorder.restype = C.c_int
        L.oracle_row_reorder.argtypes = [_u32p, _u32p, C.c_uint32, C.c_uint32, C.c_float, C.c_uint32, _u32p,
                                         C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.c_void_p, C.c_void_p]
        L.oracle_num_panels.restype = C.c_uint32
        L.oracle_num_panels.argtypes = [C.c_uint32]
        L.oracle_col_reorder.restype = C.c_int
        L.oracle_col_reorder.argtypes = [_u32p, _u32p, C.c_uint32, C.c_uint32, _u32p, C.c_uint32, C.c_float,
                                         _u32p, _u32p, _u32p, C.c_void_p, C.c_void_p]
        L.oracle_rphm_build.restype = C.c_int
        L.oracle_rphm_build.argtypes = [_u32p, _u32p, C.c_uint32, C.c_uint32, _u32p, C.c_uint32, _u32p, _u32p,
                                        _u32p, _u32p, _u32p, _u32p, _u32p, _u32p, _u32p, _u32p]
        L.oracle_work_lists.argtypes = [C.c_uint32, _u32p, _u32p] + [C.c_void_p] * 8
        L.oracle_original_block_stats.argtypes = [_u32p, _u32p, C.c_uint32, C.c_uint32, C.c_float, C.c_void_p, C.c_void_p]
        L.oracle_original_block_stats.restype = None
        L.oracle_evaluation_reordering.argtypes = [_u32p, _u32p, C.c_uint32, C.c_uint32, C.c_uint32, _u32p, C.c_uint32,
                                                   _u32p, _u32p, _u32p, _u32p, _u32p, C.c_float, _u32p, C.c_void_p]
        L.oracle_evaluation_reordering.restype = None
        L.oracle_sddmm_cpu.argtypes = [_f32p, _f32p, _u32p, _u32p, C.c_uint32, C.c_uint32, _f32p, C.c_int]
        L.oracle_check_data.restype = C.c_size_t
        L.oracle_check_data.argtypes = [_f32p, _f32p, C.c_size_t]
        L.oracle_load_mtx.restype = C.c_int
        L.oracle_load_mtx.argtypes = [C.c_char_p] + [C.POINTER(C.c_uint32)] * 3 + [C.POINTER(C.c_void_p)] * 3
        L.oracle_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def ref_available() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libref_cpu.so"))


def ref():
    """The reference's own host code (colReordering_cpu, sddmm_cpu, loader, checkData)."""
    global _ref
    if _ref is None:
        L = C.CDLL(os.path.join(HERE, "_ref", "libref_cpu.so"))
        L.ref_col_reorder.restype = C.c_int
        L.ref_col_reorder.argtypes = [_u32p, _u32p, C.c_uint32, C.c_uint32, C.c_uint32, _u32p, C.c_uint32,
                                      C.c_float, _u32p, _u32p, _u32p, C.c_void_p, C.c_void_p]
        L.ref_evaluation_reordering.argtypes = [_u32p, _u32p, C.c_uint32, C.c_uint32, C.c_uint32, _u32p, C.c_uint32,
                                                C.c_float, C.c_void_p, C.c_void_p]
        L.ref_evaluation_reordering.restype = None
        L.ref_sddmm_cpu.argtypes = [_f32p, _f32p, _u32p, _u32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _f32p]
        L.ref_sddmm_prepare.restype = C.c_void_p
        L.ref_sddmm_prepare.argtypes = [_f32p, _f32p, _u32p, _u32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        L.ref_sddmm_run.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_sddmm_release.argtypes = [C.c_void_p]
        L.ref_omp_threads.restype = C.c_int
        L.ref_omp_threads.argtypes = [C.c_int]
        L.ref_check_data.restype = C.c_size_t
        L.ref_check_data.argtypes = [_f32p, _f32p, C.c_size_t]
        L.ref_mtx_load.restype = C.c_int
        L.ref_mtx_load.argtypes = [C.c_char_p] + [C.POINTER(C.c_uint32)] * 3
        L.ref_mtx_copy.argtypes = [_u32p, _u32p, _f32p]
        _ref = L
    return _ref


# --------------------------------------------------------------------------- oracle wrappers
def block_size(M, N, free_mem):
    return int(lib().oracle_block_size(M, N, int(free_mem)))


def nbpr(N, bs):
    return int(lib().oracle_num_blocks_per_row(N, bs))


def cluster_blockdim(n):
    return int(lib().oracle_cluster_blockdim(n))


def kept_warps(blockdim):
    k = np.zeros(blockdim // 32, dtype=np.uint8)
    lib().oracle_kept_warps(blockdim, k.ctypes.data)
    return k


def encode_dispersion(S, bs, dense=True):
    n = nbpr(S.N, bs)
    enc = np.zeros((S.M, n), dtype=np.uint32) if dense else None
    disp = np.zeros(S.M, dtype=np.uint32)
    lib().oracle_encode_dispersion(S.row_off, S.col_idx, S.M, S.N, bs, enc.ctypes.data if dense else None, disp)
    return enc, disp


def similarity(rep, cmp, blockdim):
    rep = np.ascontiguousarray(rep, dtype=np.uint32)
    cmp = np.ascontiguousarray(cmp, dtype=np.uint32)
    return float(lib().oracle_similarity(rep, cmp, rep.shape[0], blockdim))


def row_reorder(S, alpha, bs):
    """-> dict(reorderedRows, numClusters, clusterOfRow, ascending)"""
    out = np.zeros(S.M, dtype=np.uint32)
    n = C.c_uint32(0)
    cc = C.c_int32(0)
    cof = np.zeros(S.M, dtype=np.uint32)
    asc = np.zeros(S.M, dtype=np.uint32)
    rc = lib().oracle_row_reorder(S.row_off, S.col_idx, S.M, S.N, float(alpha), bs, out, C.byref(n), C.byref(cc),
                                  cof.ctypes.data, asc.ctypes.data)
    assert rc == 0
    return dict(reorderedRows=out[: n.value].copy(), numClusters=cc.value, clusterOfRow=cof, ascending=asc)


def row_reorder_pruned(S, alpha, bs, prefix_filter=False):
    """oracle_row_reorder_pruned: the same clustering through an inverted index (alpha >= 0); adds `evaluations`."""
    L = lib()
    L.oracle_row_reorder_pruned.restype = C.c_int
    L.oracle_row_reorder_pruned.argtypes = [_u32p, _u32p, C.c_uint32, C.c_uint32, C.c_float, C.c_uint32, _u32p,
                                            C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.c_void_p, C.c_void_p,
                                            C.POINTER(C.c_uint64), C.c_int]
    out = np.zeros(max(1, S.M), dtype=np.uint32)
    n, cc, ev = C.c_uint32(0), C.c_int32(0), C.c_uint64(0)
    cof = np.zeros(max(1, S.M), dtype=np.uint32)
    asc = np.zeros(max(1, S.M), dtype=np.uint32)
    rc = L.oracle_row_reorder_pruned(S.row_off, S.col_idx, S.M, S.N, float(alpha), bs, out, C.byref(n), C.byref(cc),
                                     cof.ctypes.data, asc.ctypes.data, C.byref(ev), 1 if prefix_filter else 0)
    assert rc == 0, rc
    return dict(reorderedRows=out[: n.value].copy(), numClusters=cc.value, clusterOfRow=cof[: S.M],
                ascending=asc[: S.M], evaluations=int(ev.value))


def col_reorder(S, reordered_rows, delta):
    R = np.ascontiguousarray(reordered_rows, dtype=np.uint32)
    P = int(lib().oracle_num_panels(R.shape[0]))
    dO = np.zeros(P + 1, np.uint32)
    sO = np.zeros(P + 1, np.uint32)
    vO = np.zeros(P + 1, np.uint32)
    lib().oracle_col_reorder(S.row_off, S.col_idx, S.M, S.N, R, R.shape[0], float(delta), dO, sO, vO, None, None)
    dC = np.zeros(max(1, int(dO[-1])), np.uint32)
    sC = np.zeros(max(1, int(sO[-1])), np.uint32)
    lib().oracle_col_reorder(S.row_off, S.col_idx, S.M, S.N, R, R.shape[0], float(delta), dO, sO, vO,
                             dC.ctypes.data, sC.ctypes.data)
    return dict(numRowPanels=P, denseColOffsets=dO, sparseColOffsets=sO, sparseValueOffsets=vO,
                denseCols=dC[: int(dO[-1])], sparseCols=sC[: int(sO[-1])])


def rphm_build(S, reordered_rows, cr):
    R = np.ascontiguousarray(reordered_rows, dtype=np.uint32)
    P = cr["numRowPanels"]
    dO, sO, vO = cr["denseColOffsets"], cr["sparseColOffsets"], cr["sparseValueOffsets"]
    nblk = int(((np.diff(dO.astype(np.int64)) + 15) // 16).sum())
    bO = np.zeros(P + 1, np.uint32)
    bV = np.zeros(max(1, nblk * 256), np.uint32)
    ns = int(vO[-1])
    sV = np.zeros(max(1, ns), np.uint32)
    sR = np.zeros(max(1, ns), np.uint32)
    sC = np.zeros(max(1, ns), np.uint32)
    dC = np.ascontiguousarray(cr["denseCols"]) if cr["denseCols"].size else np.zeros(1, np.uint32)
    spC = np.ascontiguousarray(cr["sparseCols"]) if cr["sparseCols"].size else np.zeros(1, np.uint32)
    lib().oracle_rphm_build(S.row_off, S.col_idx, S.M, S.N, R, R.shape[0], dO, dC, sO, spC, vO, bO, bV, sV, sR, sC)
    out = dict(blockOffsets=bO, blockValues=bV[: nblk * 256], sparseValues=sV[:ns], sparseRelativeRows=sR[:ns],
               sparseColIndices=sC[:ns])
    # work lists (BSMR.cpp:99-119, :221-246)
    cnt = (C.c_uint32 * 4)()
    ptrs = [C.c_void_p(C.addressof(cnt) + 4 * i) for i in range(4)]
    lib().oracle_work_lists(P, dO, vO, ptrs[0], ptrs[1], None, None, ptrs[2], ptrs[3], None, None)
    nd, md, nsb, ms = (int(x) for x in cnt)
    dI = np.zeros(max(1, nd), np.uint32); dT = np.zeros(max(1, nd), np.uint32)
    sI = np.zeros(max(1, nsb), np.uint32); sT = np.zeros(max(1, nsb), np.uint32)
    lib().oracle_work_lists(P, dO, vO, None, None, dI.ctypes.data, dT.ctypes.data, None, None, sI.ctypes.data,
                            sT.ctypes.data)
    out.update(numDenseThreadBlocks=nd, maxNumDenseColBlocksInRowPanel=md, numSparseThreadBlocks=nsb,
               maxNumSparseColBlocksInRowPanel=ms, denseRowPanelIds=dI[:nd], denseColBlockIters=dT[:nd],
               sparseRowPanelIds=sI[:nsb], sparseColBlockIters=sT[:nsb])
    return out


def sddmm_cpu(S, A, B, threads=0):
    A = np.ascontiguousarray(A, np.float32)
    B = np.ascontiguousarray(B, np.float32)
    P = np.zeros(max(1, S.nnz), np.float32)
    lib().oracle_sddmm_cpu(A, B, S.row_off, S.col_idx, S.M, A.shape[1], P, int(threads))
    return P[: S.nnz]


def check_data(a, b):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return int(lib().oracle_check_data(a, b, a.shape[0]))


def load_mtx(path):
    M, N, nnz = C.c_uint32(), C.c_uint32(), C.c_uint32()
    ro, ci, va = C.c_void_p(), C.c_void_p(), C.c_void_p()
    rc = lib().oracle_load_mtx(path.encode(), C.byref(M), C.byref(N), C.byref(nnz), C.byref(ro), C.byref(ci), C.byref(va))
    if rc != 0:
        return rc, None
    row_off = np.ctypeslib.as_array(C.cast(ro, C.POINTER(C.c_uint32)), (M.value + 1,)).copy()
    col_idx = np.ctypeslib.as_array(C.cast(ci, C.POINTER(C.c_uint32)), (nnz.value,)).copy()
    vals = np.ctypeslib.as_array(C.cast(va, C.POINTER(C.c_float)), (nnz.value,)).copy()
    for p in (ro, ci, va):
        lib().oracle_free(p)
    return 0, (M.value, N.value, row_off, col_idx, vals)


# --------------------------------------------------------------------------- reference wrappers
class _quiet_stderr:
    """The reference's CudaTimeCalculator prints a CUDA error per event call on a GPU-less host
    (include/CudaTimeCalculator.cuh:7-11); that is noise, not a failure of the host algorithm."""

    def __enter__(self):
        import sys
        sys.stderr.flush()
        self._saved = os.dup(2)
        self._null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self._null, 2)

    def __exit__(self, *a):
        os.dup2(self._saved, 2)
        os.close(self._null)
        os.close(self._saved)


def ref_col_reorder(S, reordered_rows, delta):
    R = np.ascontiguousarray(reordered_rows, dtype=np.uint32)
    P = (R.shape[0] + 15) // 16
    dO = np.zeros(P + 1, np.uint32); sO = np.zeros(P + 1, np.uint32); vO = np.zeros(P + 1, np.uint32)
    with _quiet_stderr():
        ref().ref_col_reorder(S.row_off, S.col_idx, S.M, S.N, S.nnz, R, R.shape[0], float(delta), dO, sO, vO, None, None)
        dC = np.zeros(max(1, int(dO[-1])), np.uint32); sC = np.zeros(max(1, int(sO[-1])), np.uint32)
        ref().ref_col_reorder(S.row_off, S.col_idx, S.M, S.N, S.nnz, R, R.shape[0], float(delta), dO, sO, vO,
                              dC.ctypes.data, sC.ctypes.data)
    return dict(numRowPanels=P, denseColOffsets=dO, sparseColOffsets=sO, sparseValueOffsets=vO,
                denseCols=dC[: int(dO[-1])], sparseCols=sC[: int(sO[-1])])


_EVAL_KEYS = ("numDenseBlock", "numDenseThreadBlocks", "numSparseThreadBlocks", "numSparseData", "numDenseData")


def evaluation_reordering(S, reordered_rows, cr, delta):
    """evaluationReordering + original-matrix statistics (BSMR.cpp:826-994) -> dict of the logged values."""
    R = np.ascontiguousarray(reordered_rows, dtype=np.uint32)
    out6 = np.zeros(6, np.uint32)
    avg = C.c_float(0)
    dC = np.ascontiguousarray(cr["denseCols"]) if cr["denseCols"].size else np.zeros(1, np.uint32)
    sC = np.ascontiguousarray(cr["sparseCols"]) if cr["sparseCols"].size else np.zeros(1, np.uint32)
    lib().oracle_evaluation_reordering(S.row_off, S.col_idx, S.M, S.N, S.nnz, R, R.shape[0], cr["denseColOffsets"], dC,
                                       cr["sparseColOffsets"], sC, cr["sparseValueOffsets"], float(delta), out6,
                                       C.byref(avg))
    nd, ad = C.c_uint32(0), C.c_float(0)
    lib().oracle_original_block_stats(S.row_off, S.col_idx, S.M, S.N, float(delta), C.byref(nd), C.byref(ad))
    out = dict(zip(_EVAL_KEYS, (int(x) for x in out6[:5])))
    out.update(averageDensity=avg.value, originalNumDenseBlock=nd.value, originalAverageDensity=ad.value)
    return out


def ref_evaluation_reordering(S, reordered_rows, delta):
    """The reference's own evaluationReordering on a BSMR built by its CPU colReordering."""
    R = np.ascontiguousarray(reordered_rows, dtype=np.uint32)
    oi = (C.c_int * 6)()
    of = (C.c_float * 2)()
    with _quiet_stderr():
        ref().ref_evaluation_reordering(S.row_off, S.col_idx, S.M, S.N, S.nnz, R, R.shape[0], float(delta), oi, of)
    out = dict(zip(_EVAL_KEYS, (int(x) for x in oi[:5])))
    out.update(averageDensity=of[0], originalNumDenseBlock=int(oi[5]), originalAverageDensity=of[1])
    return out


def ref_sddmm_cpu(S, A, B):
    A = np.ascontiguousarray(A, np.float32); B = np.ascontiguousarray(B, np.float32)
    P = np.zeros(max(1, S.nnz), np.float32)
    ref().ref_sddmm_cpu(A, B, S.row_off, S.col_idx, S.M, S.N, A.shape[1], S.nnz, P)
    return P[: S.nnz]


def ref_check_data(a, b):
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    return int(ref().ref_check_data(a, b, a.shape[0]))


def ref_load_mtx(path):
    M, N, nnz = C.c_uint32(), C.c_uint32(), C.c_uint32()
    rc = ref().ref_mtx_load(path.encode(), C.byref(M), C.byref(N), C.byref(nnz))
    if rc != 0:
        return rc, None
    ro = np.zeros(M.value + 1, np.uint32); ci = np.zeros(nnz.value, np.uint32); va = np.zeros(nnz.value, np.float32)
    ref().ref_mtx_copy(ro, ci, va)
    return 0, (M.value, N.value, ro, ci, va)
