"""ctypes binding of libsddmm_b200.so (the C ABI of include/sddmm_b200.h).

The library is the ONLY compute path: if it is missing this module raises, and every compute entry
point of the library fails with SDDMM_E_CUDA when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsddmm_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "sddmm_b200.h")

ARRAY_IDS = dict(
    reorderedRows=0, denseCols=1, denseColOffsets=2, sparseCols=3, sparseColOffsets=4, sparseValueOffsets=5,
    blockOffsets=6, blockValues=7, sparseValues=8, sparseRelativeRows=9, sparseColIndices=10,
    denseRowPanelIds=11, denseColBlockIters=12, sparseRowPanelIds=13, sparseColBlockIters=14)


class LayoutInfo(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "M", "N", "nnz", "numRows", "numRowPanels", "panelBegin", "numDenseBlocks", "numSparseValues",
        "numDenseValues", "maxNumDenseColBlocksInRowPanel", "maxNumSparseColBlocksInRowPanel",
        "numDenseThreadBlocks", "numSparseThreadBlocks")]


class Eval(C.Structure):
    _fields_ = [("numDenseBlock", C.c_uint32), ("averageDensity", C.c_float), ("numDenseThreadBlocks", C.c_uint32),
                ("numSparseThreadBlocks", C.c_uint32), ("numDenseData", C.c_uint32), ("numSparseData", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("rowReorderMs", C.c_float), ("colReorderMs", C.c_float), ("rphmMs", C.c_float),
                ("sddmmMs", C.c_float), ("numClusters", C.c_int32), ("blockSize", C.c_uint32)]


class Plan(C.Structure):
    """sddmm_plan (include/sddmm_b200.h): which kernels serve a pass; 0 = AUTO everywhere."""
    _fields_ = [("plan", C.c_uint32), ("dense", C.c_uint32), ("residual", C.c_uint32), ("tile", C.c_uint32),
                ("tileStages", C.c_uint32), ("operands", C.c_uint32), ("reserved", C.c_uint32 * 2)]


class ReorderOpts(C.Structure):
    """bsmr_reorder_opts: how the clustering kernel gets to the (always identical) permutation."""
    _fields_ = [("kernel", C.c_uint32), ("batch", C.c_uint32), ("laneRows", C.c_uint32), ("signature", C.c_uint32),
                ("reserved", C.c_uint32 * 4)]


PLAN = dict(auto=0, bsmr=1, tile=2)
DENSE = dict(auto=0, reg=1, tma=2)
RESIDUAL = dict(auto=0, panel=1, superpanel=2, stream=3)
TILE = dict(auto=0, reg=1, tma=2, tma_cluster=3, tma_pair=4)
OPERANDS = dict(exact=0, fp16=1)
CLUSTER = dict(auto=0, legacy=1, batched=2)
TRISTATE = dict(auto=0, off=1, on=2)
BUILD_TILES = dict(auto=0, always=1, never=2)


class SddmmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsddmm_b200 error {code}: {msg}")
        self.code = code


def declared_symbols():
    """Every function name include/sddmm_b200.h declares (used by the CPU-side ABI test)."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:sddmm|bsmr)_[a-z0-9_]+)\s*\(", src)))


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python __graft_entry__.py` (build()); there is no fallback path")
    L = C.CDLL(LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    if missing:
        raise ImportError(f"libsddmm_b200.so does not export {missing}")
    vp, u32, u64, f32, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_float, C.c_int32
    pf32, pu32, pi32 = C.POINTER(f32), C.POINTER(u32), C.POINTER(i32)
    L.sddmm_b200_abi_version.restype = C.c_int
    L.sddmm_last_error.restype = C.c_char_p
    L.sddmm_launch_count.restype = u64
    L.sddmm_launch_count_reset.restype = None
    L.bsmr_calc_block_size.restype = u32
    L.bsmr_calc_block_size.argtypes = [u32, u32, u64]
    L.bsmr_row_reorder_dev.argtypes = [vp, vp, u32, u32, u32, f32, u32, vp, pu32, pi32, pf32, vp]
    L.bsmr_row_reorder.argtypes = [vp, vp, u32, u32, u32, f32, u32, vp, pu32, pi32, pf32]
    L.bsmr_dispersion_dev.argtypes = [vp, vp, u32, u32, u32, u32, vp, pu32, vp]
    L.bsmr_layout_build_dev.argtypes = [vp, vp, u32, u32, u32, vp, u32, f32, u32, u32, C.POINTER(vp), pf32, pf32, vp]
    L.bsmr_layout_build.argtypes = [vp, vp, u32, u32, u32, vp, u32, f32, C.POINTER(vp), pf32, pf32]
    L.bsmr_row_reorder_dev_ex.argtypes = [vp, vp, u32, u32, u32, f32, u32, C.POINTER(ReorderOpts), vp, pu32, pi32, pf32, vp]
    L.bsmr_row_reorder_ex.argtypes = [vp, vp, u32, u32, u32, f32, u32, C.POINTER(ReorderOpts), vp, pu32, pi32, pf32]
    L.bsmr_layout_build_dev_ex.argtypes = [vp, vp, u32, u32, u32, vp, u32, f32, u32, u32, u32, C.POINTER(vp), pf32, pf32, vp]
    L.bsmr_layout_build_ex.argtypes = [vp, vp, u32, u32, u32, vp, u32, f32, u32, C.POINTER(vp), pf32, pf32]
    L.sddmm_plan_default.argtypes = [C.POINTER(Plan)]
    L.sddmm_plan_default.restype = None
    L.sddmm_plan_resolve.argtypes = [vp, u32, u32, C.POINTER(Plan), C.POINTER(Plan)]
    L.sddmm_prepare.argtypes = [vp, u32, u32, C.POINTER(Plan)]
    L.sddmm_run_dev_ex.argtypes = [vp, u32, u32, vp, vp, vp, C.POINTER(Plan), vp]
    L.bsmr_layout_destroy.argtypes = [vp]
    L.bsmr_layout_destroy.restype = None
    L.bsmr_layout_get_info.argtypes = [vp, C.POINTER(LayoutInfo)]
    L.bsmr_layout_array_len.argtypes = [vp, C.c_int]
    L.bsmr_layout_array_len.restype = C.c_size_t
    L.bsmr_layout_array_dev.argtypes = [vp, C.c_int]
    L.bsmr_layout_array_dev.restype = vp
    L.bsmr_layout_array_to_host.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.bsmr_layout_save.argtypes = [vp, C.c_char_p]
    L.bsmr_layout_load.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.sddmm_run_dev.argtypes = [vp, u32, vp, vp, vp, vp]
    L.sddmm_run_batch_dev.argtypes = [vp, u32, u32, vp, vp, vp, vp]
    L.sddmm_run_timed_dev.argtypes = [vp, u32, vp, vp, vp, C.c_int, C.c_int, pf32, pf32, pf32]
    L.sddmm_run_host.argtypes = [vp, u32, vp, vp, vp, pf32]
    L.sddmm_run_host_async.argtypes = [vp, u32, vp, vp, vp, C.c_int]
    L.sddmm_host_sync.argtypes = [vp]
    L.sddmm_host_traffic.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    L.sddmm_host.argtypes = [vp, vp, u32, u32, u32, u32, vp, vp, f32, f32, u32, vp, C.POINTER(Stats), C.POINTER(vp)]
    L.bsmr_shard_plan.argtypes = [vp, vp, u32, u32, vp]
    L.sddmm_coo_to_csr.argtypes = [vp, vp, vp, u32, u32, u32, vp, vp, vp, C.POINTER(C.c_int)]
    L.bsmr_shard_plan_dev.argtypes = [vp, vp, u32, u32, vp, vp]
    L.sddmm_mgpu_unique_id.argtypes = [vp]
    L.sddmm_mgpu_init.argtypes = [C.c_int, C.c_int, vp, C.POINTER(vp)]
    L.sddmm_mgpu_destroy.argtypes = [vp]
    L.sddmm_mgpu_destroy.restype = None
    L.sddmm_mgpu_shard.argtypes = [vp, vp, vp, u32, u32, u32, vp, pu32, f32, u32, C.POINTER(vp), vp, pf32, pf32, vp]
    L.sddmm_mgpu_rebalance.argtypes = [vp, vp, vp, u32, u32, u32, vp, u32, f32, u32, f32, C.POINTER(vp), vp, vp]
    L.bsmr_rebalance_cuts.argtypes = [vp, u32, vp, vp, u32, vp]
    L.sddmm_mgpu_bcast.argtypes = [vp, vp, C.c_size_t, C.c_int, vp]
    L.sddmm_mgpu_run.argtypes = [vp, vp, u32, vp, vp, vp, vp]
    L.sddmm_mgpu_gather.argtypes = [vp, vp, C.c_size_t, vp]
    L.sddmm_mgpu_run_host.argtypes = [vp, vp, u32, vp, vp, vp, C.c_int, pf32]
    L.sddmm_mgpu_host_traffic.argtypes = [vp, C.POINTER(u64)]
    L.bsmr_layout_eval.argtypes = [vp, f32, C.POINTER(Eval)]
    L.bsmr_original_block_stats_dev.argtypes = [vp, vp, u32, u32, u32, f32, vp, vp, vp]
    L.bsmr_original_block_stats.argtypes = [vp, vp, u32, u32, u32, f32, vp, vp]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise SddmmError(rc, lib().sddmm_last_error().decode(errors="replace"))
