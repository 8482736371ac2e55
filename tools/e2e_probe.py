"""Development aid: end-to-end step time of sddmm_run_host_async (two slots, pinned host buffers) on an R-MAT layout
with identity row order -- for tuning the referenced-rows gather (SDDMM_B200_H2D, SDDMM_B200_H2D_CTAS).

    SDDMM_B200_H2D_CTAS=256 python tools/e2e_probe.py [--scale 22] [--K 256] [--steps 6]
"""
import argparse
import json
import os
import sys
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--K", type=int, default=256)
    ap.add_argument("--steps", type=int, default=6)
    a = ap.parse_args()
    import torch
    pkg = load_package()
    ro, ci, M = pkg.generators.rmat_device(a.scale, 16, 4)
    row_off = ro.cpu().numpy().view(np.uint32)
    nnz = int(ci.numel())
    lens = np.diff(row_off.astype(np.int64))
    R = torch.from_numpy(np.nonzero(lens)[0].astype(np.int32)).cuda()
    lay, _, _ = pkg.layout_build_dev(ro, ci, M, M, R, 0.3)
    K = a.K
    hA = (torch.rand((M, K)) * 2).pin_memory()
    hB = (torch.rand((M, K)) * 2).pin_memory()
    hP = [torch.zeros(nnz).pin_memory() for _ in range(2)]
    nA, nB, nPs = hA.numpy(), hB.numpy(), [t.numpy() for t in hP]
    for i in range(2):
        pkg.sddmm_gpu_async(nA, nB, lay, nPs[i], i)
    pkg.sddmm_gpu_sync(lay)
    t0 = time.perf_counter()
    for i in range(a.steps):
        pkg.sddmm_gpu_async(nA, nB, lay, nPs[i & 1], i & 1)
    pkg.sddmm_gpu_sync(lay)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / a.steps
    h2d, d2h = pkg.host_traffic(lay)
    print(json.dumps(dict(K=K, ms_per_step=round(ms, 2), h2d_bytes=h2d, d2h_bytes=d2h, h2d_gbs=round(h2d / ms / 1e6, 1),
                          ctas=os.environ.get("SDDMM_B200_H2D_CTAS", "default"), mode=os.environ.get("SDDMM_B200_H2D", "auto"))))


if __name__ == "__main__":
    main()
