"""Multi-GPU layer: one process per GPU, torch.distributed for the plumbing.

After row reordering the 16-row panels are independent units (every kernel indexes by row panel and
writes disjoint CSR positions), so the path shards with NO steady-state collective:
  * rank 0's row order is broadcast once (the clustering is a sequential chain: it is not sharded),
  * bsmr_shard_plan cuts the panels into contiguous ranges with equal non-zero counts,
  * every rank builds the layout of its own range (bsmr_layout_build_dev with panelBegin/panelEnd),
  * B is replicated once with a broadcast over NCCL/NVLink; A and P stay where they are,
  * an optional final all-reduce(SUM) merges the disjoint pieces of P when one full P is wanted.
The host logic below is backend-agnostic (gloo on CPU in the tests, nccl on the GPUs).
"""
from __future__ import annotations

import numpy as np

from . import host


def _dist():
    import torch.distributed as dist
    return dist


def broadcast_row_order(reordered_rows, M, src=0, device="cpu"):
    """Rank `src` provides the permutation (uint32 numpy array); every rank returns it."""
    import torch
    dist = _dist()
    n = torch.tensor([0 if reordered_rows is None else len(reordered_rows)], dtype=torch.int64, device=device)
    dist.broadcast(n, src=src)
    buf = torch.zeros(int(n.item()), dtype=torch.int32, device=device)
    if dist.get_rank() == src:
        buf.copy_(torch.from_numpy(np.ascontiguousarray(reordered_rows, dtype=np.uint32).view(np.int32)))
    dist.broadcast(buf, src=src)
    assert int(n.item()) <= M
    return buf.cpu().numpy().view(np.uint32)


def replicate_B(B_tensor, src=0):
    """One broadcast of B (K x N column-major, i.e. N x K rows) before the steady state."""
    _dist().broadcast(B_tensor, src=src)
    return B_tensor


def my_panel_range(S, reordered_rows, rank=None, world=None):
    dist = _dist()
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    cuts = host.shard_plan(S, reordered_rows, world)
    return int(cuts[rank]), int(cuts[rank + 1]), cuts


def merge_P(P_tensor):
    """Disjoint pieces (zeros elsewhere) -> the full P on every rank."""
    _dist().all_reduce(P_tensor)
    return P_tensor


class ShardedSDDMM:
    """Strong-scaling form: one S, row panels split over the ranks of the default process group."""

    def __init__(self, S, alpha=0.3, delta=0.3, block_size=0, device=None):
        import torch
        dist = _dist()
        self.S, self.rank, self.world = S, dist.get_rank(), dist.get_world_size()
        dev = device or torch.device("cuda", torch.cuda.current_device())
        self.ro = torch.from_numpy(S.row_off.view(np.int32)).to(dev)
        self.ci = torch.from_numpy(S.col_idx.view(np.int32)).to(dev)
        R = None
        self.row_ms = 0.0
        if self.rank == 0:
            Rt, self.num_clusters, self.row_ms = host.row_reorder_dev(self.ro, self.ci, S.M, S.N, alpha, block_size)
            R = Rt.cpu().numpy().view(np.uint32)
        self.R = broadcast_row_order(R, S.M, src=0, device=dev)
        self.p0, self.p1, self.cuts = my_panel_range(S, self.R, self.rank, self.world)
        Rt = torch.from_numpy(self.R.view(np.int32)).to(dev)
        self.layout, self.col_ms, self.rphm_ms = host.layout_build_dev(self.ro, self.ci, S.M, S.N, Rt, delta, self.p0,
                                                                       self.p1)

    def run(self, dA, dB, dP):
        return host.sddmm_gpu(dA, dB, self.layout, dP)
