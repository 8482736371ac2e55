"""Strong scaling of ONE matrix over the ranks of a torchrun job (SURVEY.md 8e, config-5 style):
row panels of the (re)ordered S are cut into nnz-balanced ranges (bsmr_shard_plan), every rank builds the
layout of its own range, B is replicated once with an NCCL broadcast, and each step every rank computes its
share of P with no collective.  Time per step = max over ranks (CUDA events), throughput = 2*nnz*K/t.

    python -m torch.distributed.run --nproc-per-node N tools/strong_scaling.py --scale 22 --K 128 [--reorder]

Without --reorder the row order is the identity over the non-empty rows (the clustering of a 4M-row R-MAT
takes minutes); the sharding / layout / SDDMM path is the same either way.
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


def rmat_device(torch, scale, edge_factor, seed, abcd=(0.57, 0.19, 0.19, 0.05)):
    """R-MAT edges drawn, de-duplicated and turned into CSR on the current GPU.  Returns (row_off int32[M+1],
    col_idx int32[nnz] sorted inside each row, M).  Input generation only -- torch is plumbing here."""
    n = edge_factor << scale
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    a, b, c, _ = abcd
    key = torch.zeros(n, dtype=torch.int64, device="cuda")  # row << scale | col
    for lvl in range(scale):
        r = torch.rand(n, device="cuda", generator=g)
        rowbit = (r >= a + b)
        colbit = ((r >= a) & (r < a + b)) | (r >= a + b + c)
        key |= (rowbit.to(torch.int64) << (scale + scale - 1 - lvl)) | (colbit.to(torch.int64) << (scale - 1 - lvl))
        del r, rowbit, colbit
    key = torch.unique(key)  # sorted, duplicates removed
    M = 1 << scale
    rows = key >> scale
    ci = (key & (M - 1)).to(torch.int32)
    del key
    counts = torch.bincount(rows, minlength=M)
    del rows
    ro = torch.zeros(M + 1, dtype=torch.int64, device="cuda")
    ro[1:] = torch.cumsum(counts, 0)
    del counts
    assert int(ro[-1]) < 2 ** 32
    ro32 = (ro & 0xFFFFFFFF).to(torch.int64)
    ro32 = torch.where(ro32 >= 2 ** 31, ro32 - 2 ** 32, ro32).to(torch.int32)  # uint32 bit pattern
    torch.cuda.empty_cache()
    return ro32, ci, M


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--K", type=int, default=128)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--reorder", action="store_true")
    ap.add_argument("--device-gen", action="store_true", help="generate the R-MAT matrix on the GPU (config 5)")
    ap.add_argument("--seed", type=int, default=5)
    ap.add_argument("--check", action="store_true", help="verify sampled rows of this rank's shard in fp64")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    pkg = load_package()
    gen = pkg.generators
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K = a.K
    gen_ms = 0.0
    if a.device_gen:
        # config 5: the matrix never exists on the host.  Philox with a fixed seed gives every
        # rank the same edges; a checksum all-reduce below asserts it.
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ro, ci, M = rmat_device(torch, a.scale, 16, a.seed)
        e1.record(); torch.cuda.synchronize()
        gen_ms = e0.elapsed_time(e1)
        S = types.SimpleNamespace(M=M, N=M, nnz=int(ci.numel()), row_off=ro.cpu().numpy().view(np.uint32))
        if world > 1:
            chk = torch.tensor([S.nnz, int(ci[:: 1009].to(torch.int64).sum())], dtype=torch.int64, device="cuda")
            lo, hi = chk.clone(), chk.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            assert bool((lo == hi).all()), "ranks generated different matrices"
    else:
        S = gen.rmat(a.scale, 16, 4)  # same seed on every rank: same matrix
        ro = torch.from_numpy(S.row_off.view(np.int32)).cuda()
        ci = torch.from_numpy(S.col_idx.view(np.int32)).cuda()
    if a.reorder:
        R, _, row_ms = pkg.row_reorder_dev(ro, ci, S.M, S.N, 0.3, 0)
        Rh = R.cpu().numpy().view(np.uint32)
    else:
        Rh = np.nonzero(np.diff(S.row_off.astype(np.int64)))[0].astype(np.uint32)
        R = torch.from_numpy(Rh.view(np.int32)).cuda()
        row_ms = 0.0
    cuts = pkg.shard_plan(S, Rh, world)
    lay, col_ms, rphm_ms = pkg.layout_build_dev(ro, ci, S.M, S.N, R, 0.3, int(cuts[rank]), int(cuts[rank + 1]))
    mine = int(lay.info.numDenseValues + lay.info.numSparseValues)
    # A: every rank only needs its rows, kept full-size here for simplicity; B replicated from rank 0
    dA = torch.rand((S.M, K), device="cuda") * 2
    dB = torch.rand((S.N, K), device="cuda") * 2
    bc_ms = 0.0
    if world > 1:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dist.broadcast(dB, src=0); e1.record(); torch.cuda.synchronize()
        bc_ms = e0.elapsed_time(e1)
    dP = torch.zeros(S.nnz, device="cuda")
    for _ in range(3):
        pkg.sddmm_gpu(dA, dB, lay, dP)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        pkg.sddmm_gpu(dA, dB, lay, dP)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    t = torch.tensor([ms, float(mine)], dtype=torch.float64, device="cuda")
    mx = t.clone()
    if world > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        tot = t.clone(); dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    else:
        tot = t
    worst = -1.0
    if a.check:  # fp64 on the device, rows of this rank's own panel range
        lo, hi = int(cuts[rank]) * 16, min(int(cuts[rank + 1]) * 16, Rh.size)
        pick = Rh[np.linspace(lo, hi - 1, 16).astype(np.int64)] if hi > lo else []
        roh = S.row_off
        worst = 0.0
        for r in pick:
            b, e = int(roh[r]), int(roh[r + 1])
            cols = ci[b:e].to(torch.int64)
            ref = (dA[int(r)].double()[None, :] * dB[cols].double()).sum(1)
            err = ((dP[b:e].double() - ref).abs() / ref.abs().clamp_min(1e-3)).max()
            worst = max(worst, float(err))
        w = torch.tensor([worst], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(w, op=dist.ReduceOp.MAX)
        worst = float(w[0])
    if rank == 0:
        print(json.dumps(dict(workload=f"R-MAT scale {a.scale}", device_generated=a.device_gen, gen_ms=gen_ms,
                              max_rel_err_sample=worst, hbm_gb_allocated=torch.cuda.max_memory_allocated() / 1e9, M=S.M, nnz=S.nnz, K=K, n_gpus=world, ms_per_step=float(mx[0]),
                              gflops=2.0 * S.nnz * K / (float(mx[0]) * 1e-3) / 1e9, covered_nnz=int(tot[1]),
                              max_shard_nnz=int(mx[1]), b_broadcast_ms=bc_ms, row_reorder_ms=row_ms, reordered=a.reorder,
                              col_reorder_ms=col_ms, rphm_ms=rphm_ms)), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
