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
    ap.add_argument("--plan", default="auto"); ap.add_argument("--dense", default="auto")
    ap.add_argument("--residual", default="auto"); ap.add_argument("--tile", default="auto")
    ap.add_argument("--stages", type=int, default=0); ap.add_argument("--tiles", default="auto")
    ap.add_argument("--operands", default="exact")
    ap.add_argument("--host-gen", action="store_true", help="R-MAT from the host generator (tests' matrices)")
    ap.add_argument("--graph", action="store_true", help="forced plans: time replays of ONE CUDA graph per pass")
    ap.add_argument("--persist-mb", type=int, default=-1, help="experiment: cudaLimitPersistingL2CacheSize in MB")
    a = ap.parse_args()
    import torch
    pkg = load_package()
    gen = pkg.generators
    if a.persist_mb >= 0:
        import ctypes
        torch.cuda.init(); torch.zeros(1, device="cuda")
        rt = ctypes.CDLL("libcudart.so")
        rc = rt.cudaDeviceSetLimit(6, ctypes.c_size_t(a.persist_mb << 20))  # cudaLimitPersistingL2CacheSize
        v = ctypes.c_size_t(0); rt.cudaDeviceGetLimit(ctypes.byref(v), 6)
        print(f"persisting L2 limit: rc={rc} now {v.value >> 20} MB", flush=True)
    if a.workload == "uniform100k":
        S = gen.uniform_random(100_000, 100_000, 0.01, 2)
    elif a.workload == "uniform20k":
        S = gen.uniform_random(20_000, 20_000, 0.01, 2)
    elif a.workload == "bern4096":
        S = gen.bernoulli_mask(4096, 4096, a.sparsity, 30)
    elif a.workload == "dlmc4096":
        S = gen.dlmc_magnitude_mask(4096, 4096, a.sparsity, 33)
    elif a.workload == "rmat" and a.host_gen:
        S = gen.rmat(a.scale, 16, 4)
    elif a.workload == "rmat":
        import types
        ro, ci, M = gen.rmat_device(a.scale, 16, 4)
        S = types.SimpleNamespace(M=M, N=M, nnz=int(ci.numel()), row_off=ro.cpu().numpy().view(np.uint32),
                                  col_idx=None)
    elif a.workload == "blockscat":
        S = gen.block_structured_scattered(16384, 16384, 64, 512, 0.6, 41, noise=0.0005)
    else:
        raise SystemExit("unknown workload")
    if a.workload != "rmat" or a.host_gen:
        ro = torch.from_numpy(S.row_off.view(np.int32)).cuda()
        ci = torch.from_numpy(S.col_idx.view(np.int32)).cuda()
    a.ci = ci
    a.plan_obj = pkg.make_plan(a.plan, a.dense, a.residual, a.tile, a.stages, a.operands)
    if a.reorder:
        bs = pkg.calculateBlockSize(S, 180 * 10 ** 9)
        R, ncl, row_ms = pkg.row_reorder_dev(ro, ci, S.M, S.N, a.alpha, bs)
    else:
        lens = np.diff(S.row_off.astype(np.int64))
        R = torch.from_numpy(np.nonzero(lens)[0].astype(np.int32)).cuda()
        ncl, row_ms = -1, 0.0
    lay, col_ms, rphm_ms = pkg.layout_build_dev(ro, ci, S.M, S.N, R, a.delta, tiles=a.tiles)
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
    names = pkg.plan_resolve(lay, K, 1, a.plan_obj)
    if (a.plan, a.dense, a.residual, a.tile, a.operands) == ("auto",) * 4 + ("exact",):
        t = pkg.sddmm_gpu_timed(dA, dB, lay, dP, warmup=3, iters=a.iters)
    else:  # forced kernels: time whole passes with events on the current stream
        pkg.sddmm_prepare(lay, K, 1, a.plan_obj)
        for _ in range(3):
            pkg.sddmm_gpu(dA, dB, lay, dP, plan=a.plan_obj)
        torch.cuda.synchronize()
        step = lambda: pkg.sddmm_gpu(dA, dB, lay, dP, plan=a.plan_obj)  # noqa: E731
        if a.graph:  # launch-bound regime: host launch cost out of the picture
            side = torch.cuda.Stream()
            with torch.cuda.stream(side):
                step()
                torch.cuda.synchronize()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=side):
                    step()
            step = gr.replay
            for _ in range(3):
                step()
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            step()
        e1.record()
        torch.cuda.synchronize()
        t = dict(dense_ms=0.0, sparse_ms=0.0, total_ms=e0.elapsed_time(e1) / a.iters)
    i = lay.info
    out = dict(workload=a.workload, M=S.M, N=S.N, nnz=S.nnz, K=K, delta=a.delta, dense_blocks=int(i.numDenseBlocks),
               dense_nnz=int(i.numDenseValues), residual_nnz=int(i.numSparseValues), row_ms=row_ms, clusters=ncl,
               col_ms=col_ms, rphm_ms=rphm_ms, kernels=names, **t)
    out["gflops_total"] = 2.0 * S.nnz * K / (t["total_ms"] * 1e-3) / 1e9
    if t["sparse_ms"] > 0:
        out["residual_gflops"] = 2.0 * i.numSparseValues * K / (t["sparse_ms"] * 1e-3) / 1e9
    if t["dense_ms"] > 0:
        out["dense_useful_gflops"] = 2.0 * i.numDenseValues * K / (t["dense_ms"] * 1e-3) / 1e9
        out["dense_padded_tflops"] = 2.0 * 256 * i.numDenseBlocks * K / (t["dense_ms"] * 1e-3) / 1e12
    if a.check:  # sampled rows against fp64 on the device
        torch.cuda.synchronize()
        rows = np.random.default_rng(0).choice(S.M, 32, replace=False)
        ci = a.ci
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
