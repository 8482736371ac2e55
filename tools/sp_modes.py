"""Development aid: the super-panel residual kernel's modes (gather / reuse / fp16 / L2 policy) on ONE R-MAT
layout, identity or reordered row order; one JSON line per (K, mode).  Modes are environment switches the library
reads per call.

    python tools/sp_modes.py [--scale 22] [--reorder] [--Ks 32,64,256]
"""
import argparse
import json
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

MODES = {
    "default": {},
    "gather": {"SDDMM_B200_SP_GATHER": "1"},
    "reuse": {"SDDMM_B200_SP_GATHER": "0"},
    "nohints": {"SDDMM_B200_L2_HINTS": "0"},
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--reorder", action="store_true")
    ap.add_argument("--Ks", default="32,64,128,256")
    ap.add_argument("--modes", default="default,gather,reuse")
    ap.add_argument("--operands", default="exact,fp16")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--persist-mb", type=int, default=-1)
    a = ap.parse_args()
    import torch
    pkg = load_package()
    if a.persist_mb >= 0:
        import ctypes
        torch.zeros(1, device="cuda")
        rt = ctypes.CDLL("libcudart.so")
        print("persisting L2 limit rc", rt.cudaDeviceSetLimit(6, ctypes.c_size_t(a.persist_mb << 20)), flush=True)
    ro, ci, M = pkg.generators.rmat_device(a.scale, 16, 4)
    S = types.SimpleNamespace(M=M, N=M, nnz=int(ci.numel()), row_off=ro.cpu().numpy().view(np.uint32))
    if a.reorder:
        bs = pkg.calculateBlockSize(S, 180 * 10 ** 9)
        R, ncl, row_ms = pkg.row_reorder_dev(ro, ci, M, M, 0.3, bs)
    else:
        lens = np.diff(S.row_off.astype(np.int64))
        R = torch.from_numpy(np.nonzero(lens)[0].astype(np.int32)).cuda()
    lay, _, _ = pkg.layout_build_dev(ro, ci, M, M, R, 0.3)
    rows = np.random.default_rng(0).choice(M, 24, replace=False)
    for K in [int(x) for x in a.Ks.split(",")]:
        g = torch.Generator(device="cuda")
        g.manual_seed(1001)
        dA = torch.rand((M, K), device="cuda", generator=g) * 2
        dB = torch.rand((M, K), device="cuda", generator=g) * 2
        dP = torch.zeros(S.nnz, dtype=torch.float32, device="cuda")
        for op in a.operands.split(","):
            plan = pkg.make_plan(plan="bsmr", residual="superpanel", operands=op)
            for mode in a.modes.split(","):
                for k in ("SDDMM_B200_SP_GATHER", "SDDMM_B200_L2_HINTS"):
                    os.environ.pop(k, None)
                os.environ.update(MODES[mode])
                pkg.sddmm_prepare(lay, K, 1, plan)
                for _ in range(3):
                    pkg.sddmm_gpu(dA, dB, lay, dP, plan=plan)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.iters):
                    pkg.sddmm_gpu(dA, dB, lay, dP, plan=plan)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / a.iters
                worst = 0.0
                for r in rows:
                    b, e = int(S.row_off[r]), int(S.row_off[r + 1])
                    if e > b:
                        ref = (dA[int(r)].double()[None, :] * dB[ci[b:e].to(torch.int64)].double()).sum(1)
                        worst = max(worst, float(((dP[b:e].double() - ref).abs() / ref.abs().clamp_min(1e-3)).max()))
                print(json.dumps(dict(K=K, operands=op, mode=mode, ms=round(ms, 4), reordered=bool(a.reorder),
                                      gflops=round(2.0 * S.nnz * K / (ms * 1e-3) / 1e9), max_rel_err=worst)), flush=True)
        del dA, dB, dP
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
