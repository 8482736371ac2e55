"""Kernel micro-benchmark (development aid): SDDMM kernels only, on a layout built from a given row order
(identity by default, so the 100k-row clustering is skipped).  Not the judged bench (that is bench.py).

    python tools/kbench.py [--workload uniform100k|bern4096|rmat] [--K 128] [--delta 0.3] [--iters 10]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="uniform100k")
    ap.add_argument("--K", type=int, default=128)
    ap.add_argument("--delta", type=float, default=0.3)
    ap.add_argument("--alpha", type=float, default=0.3)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--reorder", action="store_true", help="run the real row reordering instead of identity")
    ap.add_argument("--scale", type=int, default=18)
    ap.add_argument("--sparsity", type=float, default=0.7)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--rebuild", action="store_true", help="build the layout twice and print both timings")
    ap.add_argument("--Ks", default="", help="comma list: sweep K on one layout, one JSON line per K")
    a = ap.parse_args()
    import torch
    pkg = load_package()
    gen = pkg.generators
    if a.workload == "uniform100k":
        S = gen.uniform_random(100_000, 100_000, 0.01, 2)
    elif a.workload == "uniform20k":
        S = gen.uniform_random(20_000, 20_000, 0.01, 2)
    elif a.workload == "bern4096":
        S = gen.bernoulli_mask(4096, 4096, a.sparsity, 30)
    elif a.workload == "dlmc4096":
        S = gen.dlmc_magnitude_mask(4096, 4096, a.sparsity, 33)
    elif a.workload == "rmat":
        S = gen.rmat(a.scale, 16, 4)
    else:
        raise SystemExit("unknown workload")
    ro = torch.from_numpy(S.row_off.view(np.int32)).cuda()
    ci = torch.from_numpy(S.col_idx.view(np.int32)).cuda()
    if a.reorder:
        R, ncl, row_ms = pkg.row_reorder_dev(ro, ci, S.M, S.N, a.alpha, 0)
    else:
        lens = np.diff(S.row_off.astype(np.int64))
        R = torch.from_numpy(np.nonzero(lens)[0].astype(np.int32)).cuda()
        ncl, row_ms = -1, 0.0
    lay, col_ms, rphm_ms = pkg.layout_build_dev(ro, ci, S.M, S.N, R, a.delta)
    if a.rebuild:  # second build in the same process: scratch pool already grown
        del lay
        lay, col2, rphm2 = pkg.layout_build_dev(ro, ci, S.M, S.N, R, a.delta)
        print(json.dumps(dict(first_build_ms=[col_ms, rphm_ms], second_build_ms=[col2, rphm2])), flush=True)
    for K in ([int(x) for x in a.Ks.split(",")] if a.Ks else [a.K]):
        run_k(a, torch, pkg, gen, S, lay, K, ncl, row_ms, col_ms, rphm_ms)


def run_k(a, torch, pkg, gen, S, lay, K, ncl, row_ms, col_ms, rphm_ms):
    g = torch.Generator(device="cuda")
    g.manual_seed(1001)
    dA = torch.rand((S.M, K), device="cuda", generator=g) * 2  # U[0,2) like Matrix::makeData
    dB = torch.rand((S.N, K), device="cuda", generator=g) * 2
    dP = torch.zeros(max(1, S.nnz), dtype=torch.float32, device="cuda")
    t = pkg.sddmm_gpu_timed(dA, dB, lay, dP, warmup=3, iters=a.iters)
    i = lay.info
    out = dict(workload=a.workload, M=S.M, N=S.N, nnz=S.nnz, K=K, delta=a.delta, dense_blocks=int(i.numDenseBlocks),
               dense_nnz=int(i.numDenseValues), residual_nnz=int(i.numSparseValues), row_ms=row_ms, clusters=ncl,
               col_ms=col_ms, rphm_ms=rphm_ms, **t)
    out["gflops_total"] = 2.0 * S.nnz * K / (t["total_ms"] * 1e-3) / 1e9
    if t["sparse_ms"] > 0:
        out["residual_gflops"] = 2.0 * i.numSparseValues * K / (t["sparse_ms"] * 1e-3) / 1e9
    if t["dense_ms"] > 0:
        out["dense_useful_gflops"] = 2.0 * i.numDenseValues * K / (t["dense_ms"] * 1e-3) / 1e9
        out["dense_padded_tflops"] = 2.0 * 256 * i.numDenseBlocks * K / (t["dense_ms"] * 1e-3) / 1e12
    if a.check:  # sampled rows against fp64 on the device
        torch.cuda.synchronize()
        rows = np.random.default_rng(0).choice(S.M, 32, replace=False)
        ci = torch.from_numpy(S.col_idx.view(np.int32)).cuda()
        worst = 0.0
        for r in rows:
            b, e = int(S.row_off[r]), int(S.row_off[r + 1])
            if e > b:
                ref = (dA[int(r)].double()[None, :] * dB[ci[b:e].to(torch.int64)].double()).sum(1)
                worst = max(worst, float(((dP[b:e].double() - ref).abs() / ref.abs().clamp_min(1e-3)).max()))
        out["max_rel_err_sample"] = worst
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
