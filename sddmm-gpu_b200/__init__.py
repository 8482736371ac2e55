"""sddmm-gpu_b200: a B200-native (sm_100a) SDDMM engine behind the host API of CX9898/sddmm-gpu.

Import name: `sddmm_gpu_b200` (the directory name has a hyphen; `__graft_entry__.load_package()`
and tests/conftest.py register it).  The compute path is `libsddmm_b200.so` (hand-written CUDA
behind the C-ABI of include/sddmm_b200.h); nothing in here falls back to the CPU.
"""
from . import generators  # noqa: F401
from ._lib import SddmmError, declared_symbols, lib  # noqa: F401
from .host import (BSMR, RPHM, Layout, calculateBlockSize, dispersion_dev, launch_count,  # noqa: F401
                   layout_build_dev, row_reorder_dev, sddmm, sddmm_gpu, sddmm_gpu_async, sddmm_gpu_batch, sddmm_gpu_sync,
                   sddmm_gpu_timed, shard_plan, evaluationReordering, make_plan, make_reorder_opts, plan_resolve,
                   sddmm_prepare, host_traffic)

__all__ = ["generators", "lib", "declared_symbols", "SddmmError", "BSMR", "RPHM", "Layout", "calculateBlockSize",
           "sddmm", "sddmm_gpu", "sddmm_gpu_async", "sddmm_gpu_batch", "sddmm_gpu_sync", "sddmm_gpu_timed", "row_reorder_dev", "layout_build_dev", "dispersion_dev",
           "shard_plan", "launch_count", "evaluationReordering", "make_plan", "make_reorder_opts", "plan_resolve",
           "sddmm_prepare", "host_traffic"]
