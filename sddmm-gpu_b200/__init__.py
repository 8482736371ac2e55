"""sddmm-gpu_b200: a B200-native (sm_100a) SDDMM engine behind the host API of CX9898/sddmm-gpu.

Import name: `sddmm_gpu_b200` (the directory name has a hyphen; `__graft_entry__.load_package()`
and tests/conftest.py register it).  The compute path is `libsddmm_b200.so` (hand-written CUDA
behind the C-ABI of include/sddmm_b200.h); nothing in here falls back to the CPU.
"""
from . import generators  # noqa: F401

__all__ = ["generators"]
