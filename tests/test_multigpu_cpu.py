"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: row-order broadcast, nnz-balanced panel
shards, B replication, and the disjoint-merge of P.  Each rank's share of P is produced by the oracle
here (test infrastructure) -- the product kernels are covered by the -m gpu tests."""
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
        # rank 0 owns the row order (here from the oracle) and B; both are replicated once
        R = O.row_reorder(S, 0.3, 16)["reorderedRows"] if rank == 0 else None
        R = mg.broadcast_row_order(R, S.M, src=0)
        Bt = torch.from_numpy(B.copy()) if rank == 0 else torch.zeros((S.N, K))
        mg.replicate_B(Bt, src=0)
        assert np.array_equal(Bt.numpy(), B)
        p0, p1, cuts = mg.my_panel_range(S, R)
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
        mg.merge_P(Pt)
        ok = np.array_equal(Pt.numpy(), Pfull)
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
