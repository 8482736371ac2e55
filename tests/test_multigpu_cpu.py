"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: the out-of-band exchange of the library's NCCL
communicator id (multigpu.exchange_unique_id), nnz-balanced panel shards (bsmr_shard_plan: the rule
sddmm_mgpu_shard applies on the device), the disjoint-merge property sddmm_mgpu_gather relies on, and the
packed referenced-rows exchange of sddmm_mgpu_run_host.  Each rank's
share of P is produced by the oracle here (test infrastructure); the product kernels and the NCCL calls themselves
run in the -m gpu tests and under bench.py --gpus N."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cases import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from __graft_entry__ import load_package
        pkg = load_package()
        from sddmm_gpu_b200 import multigpu as mg
        from oracle import oracle as O
        gen = pkg.generators
        S = gen.rmat(11, 8, 6)
        K = 32
        A, B = gen.dense_operands(S.M, S.N, K)
        # the 128-byte communicator id is made on rank 0 (ncclGetUniqueId needs no GPU) and reaches every rank
        ident = mg.exchange_unique_id()
        assert len(ident) == mg.ID_BYTES and any(ident)
        gathered = [None] * world
        dist.all_gather_object(gathered, ident)
        assert all(g == gathered[0] for g in gathered)
        # there is no CPU fallback behind the communicator either
        with pytest.raises(pkg.SddmmError):
            mg.MultiGpu(rank, world, ident)
        # rank 0 owns the row order (here from the oracle) and B; both are replicated once
        Rt = torch.zeros(S.M, dtype=torch.int32)
        n = torch.zeros(1, dtype=torch.int64)
        if rank == 0:
            R0 = O.row_reorder(S, 0.3, 16)["reorderedRows"]
            Rt[: len(R0)] = torch.from_numpy(R0.view(np.int32))
            n[0] = len(R0)
        dist.broadcast(n, src=0)
        dist.broadcast(Rt, src=0)
        R = Rt[: int(n)].numpy().view(np.uint32)
        Bt = torch.from_numpy(B.copy()) if rank == 0 else torch.zeros((S.N, K))
        dist.broadcast(Bt, src=0)
        assert np.array_equal(Bt.numpy(), B)
        cuts = pkg.shard_plan(S, R, world)
        p0, p1 = int(cuts[rank]), int(cuts[rank + 1])
        assert cuts[0] == 0 and cuts[-1] == (len(R) + 15) // 16
        rows = R[p0 * 16: min(p1 * 16, len(R))]
        # this rank's disjoint share of P
        Pfull = O.sddmm_cpu(S, A, Bt.numpy())
        mine = np.zeros(S.nnz, np.float32)
        for r in rows:
            b, e = int(S.row_off[r]), int(S.row_off[r + 1])
            mine[b:e] = Pfull[b:e]
        nnz_mine = int(sum(int(S.row_off[r + 1]) - int(S.row_off[r]) for r in rows))
        Pt = torch.from_numpy(mine)
        dist.all_reduce(Pt)  # disjoint pieces, zeros elsewhere: the sum IS the merge (sddmm_mgpu_gather)
        ok = np.array_equal(Pt.numpy(), Pfull)
        # host pass of the whole job (sddmm_mgpu_run_host): every rank packs ITS share of the referenced rows of A
        # (the non-empty rows of S, in R's order), the packed buffers are all-gathered, row i of the list is unpacked
        # to row R[i]: every referenced row arrives exactly once, nothing else is touched
        per = -(-len(R) // world)
        beg, end = mg.packed_share(len(R), rank, world)
        packed = torch.zeros((per, K))
        packed[: end - beg] = torch.from_numpy(A[R[beg:end].astype(np.int64)])
        parts = [torch.zeros((per, K)) for _ in range(world)]
        dist.all_gather(parts, packed)
        allp = torch.cat(parts)[: len(R)].numpy()
        dA = np.zeros_like(A)
        dA[R.astype(np.int64)] = allp
        ref_rows = np.zeros(S.M, bool)
        ref_rows[R.astype(np.int64)] = True
        ok = ok and np.array_equal(dA[ref_rows], A[ref_rows]) and not dA[~ref_rows].any()
        shares = [mg.packed_share(len(R), r, world) for r in range(world)]
        ok = ok and shares[0][0] == 0 and shares[-1][1] == len(R) and all(shares[i][1] == shares[i + 1][0] for i in range(world - 1))
        q.put((rank, ok, nnz_mine, S.nnz, [int(c) for c in cuts]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_shard_broadcast_merge():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert all(r[1] for r in res)
    assert res[0][4] == res[1][4]                      # same plan on every rank
    total = res[0][3]
    assert res[0][2] + res[1][2] == total              # shards cover every non-zero exactly once
    assert abs(res[0][2] - res[1][2]) < 0.2 * total    # balanced by nnz


def test_rebalance_cuts_equalises_estimated_time():
    """bsmr_rebalance_cuts (the rule behind sddmm_mgpu_rebalance), host only: with a cost per non-zero that differs
    between the old shards, the new cuts equalise the estimated time; equal times leave balanced cuts alone."""
    sys.path.insert(0, ROOT)
    from __graft_entry__ import load_package
    pkg = load_package()
    L = pkg.lib()
    rng = np.random.default_rng(5)
    P, world = 4000, 4
    cnt = rng.integers(1, 400, P).astype(np.uint64)
    pre = np.zeros(P + 1, np.uint64)
    np.cumsum(cnt, out=pre[1:])
    old = np.array([0, 1000, 2000, 3000, P], np.uint32)
    true_cost = np.where(np.arange(P) < 1500, 3.0, 1.0) * cnt  # the first 1500 panels cost 3x per non-zero
    for _ in range(4):  # a few rounds settle it (the cost density inside an old shard is not uniform)
        ms = np.array([true_cost[old[r]:old[r + 1]].sum() for r in range(world)], np.float32)
        new = np.zeros(world + 1, np.uint32)
        assert L.bsmr_rebalance_cuts(pre.ctypes.data, P, old.ctypes.data, ms.ctypes.data, world, new.ctypes.data) == 0
        assert new[0] == 0 and new[-1] == P and np.all(np.diff(new.astype(np.int64)) >= 0)
        old = new
    t = np.array([true_cost[old[r]:old[r + 1]].sum() for r in range(world)])
    assert t.max() / t.mean() < 1.03, t
    # already balanced: unchanged up to one panel
    ms = np.array([float(pre[old[r + 1]] - pre[old[r]]) for r in range(world)], np.float32)
    nnz_cuts = np.zeros(world + 1, np.uint32)
    eq = np.array([0, 0, 0, 0, P], np.uint32)
    for s in range(1, world):
        eq[s] = int(np.searchsorted(pre, int(pre[-1]) * s // world))
    ms = np.array([float(pre[eq[r + 1]] - pre[eq[r]]) for r in range(world)], np.float32)
    assert L.bsmr_rebalance_cuts(pre.ctypes.data, P, eq.ctypes.data, ms.ctypes.data, world, nnz_cuts.ctypes.data) == 0
    assert np.all(np.abs(nnz_cuts.astype(np.int64) - eq.astype(np.int64)) <= 1)
