"""Multi-GPU layer over the C ABI's sddmm_mgpu_* entry points: ONE process per GPU.

After row reordering the 16-row panels are independent units (every kernel indexes by row panel and writes
disjoint CSR positions, reference src/sddmmKernel.cu:239, :2028), so the path shards with NO steady-state
collective:
  * rank 0 computes the row order once (the clustering is a sequential chain: it is not sharded);
  * sddmm_mgpu_shard broadcasts it (NCCL), cuts the panels into `world` contiguous ranges with equal non-zero
    counts (the bsmr_shard_plan rule) and builds THIS rank's layout (panelBegin/panelEnd = its range);
  * sddmm_mgpu_bcast replicates B once over NVLink; A and P stay where they are;
  * sddmm_mgpu_run is an ordinary pass over the rank's own panels; sddmm_mgpu_gather (optional) merges P.
The library's NCCL communicator needs its 128-byte id moved from rank 0 to every rank out of band;
`exchange_unique_id` does that with torch.distributed (any backend: nccl under torchrun, gloo in the CPU tests).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib, host
from ._lib import check

ID_BYTES = 128


def _point_at_bundled_nccl():
    """torch ships libnccl.so.2 inside site-packages; name it for the library's dlopen unless the caller did."""
    if os.environ.get("SDDMM_B200_NCCL_LIB"):
        return
    try:
        import nvidia.nccl as n  # noqa: PLC0415
        p = os.path.join(os.path.dirname(n.__file__ or list(n.__path__)[0]), "lib", "libnccl.so.2")
        if os.path.exists(p):
            os.environ["SDDMM_B200_NCCL_LIB"] = p
    except Exception:
        pass


def unique_id() -> bytes:
    """sddmm_mgpu_unique_id: rank 0 creates the communicator id."""
    _point_at_bundled_nccl()
    buf = C.create_string_buffer(ID_BYTES)
    check(_lib.lib().sddmm_mgpu_unique_id(buf))
    return buf.raw


def exchange_unique_id(make_id=unique_id, src=0, device="cpu") -> bytes:
    """Rank `src` makes the id, every rank of the default torch.distributed group returns it (one 128-byte
    broadcast -- the only thing torch.distributed carries for the product path)."""
    import torch
    import torch.distributed as dist

    t = torch.zeros(ID_BYTES, dtype=torch.uint8, device=device)
    if dist.get_rank() == src:
        raw = make_id()
        assert len(raw) == ID_BYTES
        t.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def packed_share(n: int, rank: int, world: int):
    """The slice [beg, end) of an n-entry referenced-row list that `rank` brings over its own PCIe link in
    sddmm_mgpu_run_host (csrc/mgpu.cu): ceil(n / world) consecutive list positions per rank, the tail clipped.  The
    packed buffers of all ranks are all-gathered and row i of the list is unpacked to row list[i] of the operand."""
    per = -(-n // world) if world > 0 else n
    return min(n, rank * per), min(n, (rank + 1) * per)


class MultiGpu:
    """Owns a `sddmm_mgpu*`: this process's rank of the library's NCCL communicator on the current CUDA device."""

    def __init__(self, rank: int, world: int, id_bytes: bytes | None = None):
        _point_at_bundled_nccl()
        self.rank, self.world = int(rank), int(world)
        h = C.c_void_p()
        idbuf = C.create_string_buffer(id_bytes, ID_BYTES) if id_bytes is not None else None
        check(_lib.lib().sddmm_mgpu_init(self.rank, self.world, idbuf, C.byref(h)))
        self._h = h

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h and getattr(_lib, "_lib", None) is not None:
            _lib._lib.sddmm_mgpu_destroy(h)

    __del__ = close

    def shard(self, row_off_t, col_idx_t, M, N, reordered_rows_t, num_rows, delta, tiles="auto"):
        """sddmm_mgpu_shard: `reordered_rows_t` has capacity M on every rank and is valid on rank 0 (num_rows
        entries).  Returns (layout of this rank's panel range, cuts, numRows, col_ms, rphm_ms)."""
        import torch

        n = C.c_uint32(int(num_rows))
        cuts = np.zeros(self.world + 1, dtype=np.uint32)
        h = C.c_void_p()
        msC, msR = C.c_float(0), C.c_float(0)
        stream = torch.cuda.current_stream().cuda_stream
        check(_lib.lib().sddmm_mgpu_shard(self._h, row_off_t.data_ptr(), col_idx_t.data_ptr(), M, N, col_idx_t.numel(),
                                          reordered_rows_t.data_ptr(), C.byref(n), float(delta),
                                          _lib.BUILD_TILES[tiles], C.byref(h), cuts.ctypes.data, C.byref(msC),
                                          C.byref(msR), C.c_void_p(stream)))
        return host.Layout(h.value), cuts, int(n.value), msC.value, msR.value

    def rebalance(self, layout, row_off_t, col_idx_t, M, N, reordered_rows_t, num_rows, delta, my_ms, tiles="auto"):
        """sddmm_mgpu_rebalance (collective): move the panel cuts so that every rank's estimated time is equal, given
        the device time `my_ms` of this rank's last pass.  Returns (layout, cuts); the old layout object is consumed."""
        import torch

        cuts = np.zeros(self.world + 1, dtype=np.uint32)
        h = C.c_void_p(layout.handle.value)
        layout._h = None  # ownership moves into the call (it destroys / replaces the layout)
        stream = torch.cuda.current_stream().cuda_stream
        check(_lib.lib().sddmm_mgpu_rebalance(self._h, row_off_t.data_ptr(), col_idx_t.data_ptr(), M, N,
                                              col_idx_t.numel(), reordered_rows_t.data_ptr(), int(num_rows), float(delta),
                                              _lib.BUILD_TILES[tiles], float(my_ms), C.byref(h), cuts.ctypes.data,
                                              C.c_void_p(stream)))
        return host.Layout(h.value), cuts

    def bcast(self, tensor, root=0):
        """sddmm_mgpu_bcast: one replication of a device tensor (B) from `root`, on the current stream."""
        import torch

        stream = torch.cuda.current_stream().cuda_stream
        check(_lib.lib().sddmm_mgpu_bcast(self._h, tensor.data_ptr(), tensor.numel() * tensor.element_size(), int(root),
                                          C.c_void_p(stream)))
        return tensor

    def run(self, layout, A, B, P):
        """sddmm_mgpu_run: this rank's share of one pass; no communication."""
        import torch

        stream = torch.cuda.current_stream().cuda_stream
        check(_lib.lib().sddmm_mgpu_run(self._h, layout.handle, A.shape[1], A.data_ptr(), B.data_ptr(), P.data_ptr(),
                                        C.c_void_p(stream)))
        return P

    def run_host(self, layout, A, B, P=None, root=0):
        """sddmm_mgpu_run_host: host numpy A, B in (every rank the same arrays; each copies only its slice), host P
        out on `root`.  Returns this rank's device milliseconds for the call."""
        ms = C.c_float(0)
        check(_lib.lib().sddmm_mgpu_run_host(self._h, layout.handle, A.shape[1], A.ctypes.data, B.ctypes.data,
                                             P.ctypes.data if P is not None else None, int(root), C.byref(ms)))
        return ms.value

    def host_traffic(self):
        """host -> device bytes this rank moved in its most recent run_host call (sddmm_mgpu_host_traffic)."""
        b = C.c_uint64(0)
        check(_lib.lib().sddmm_mgpu_host_traffic(self._h, C.byref(b)))
        return int(b.value)

    def gather(self, P):
        """sddmm_mgpu_gather: disjoint pieces (zeros elsewhere) -> the full P on every rank."""
        import torch

        stream = torch.cuda.current_stream().cuda_stream
        check(_lib.lib().sddmm_mgpu_gather(self._h, P.data_ptr(), P.numel(), C.c_void_p(stream)))
        return P


class ShardedSDDMM:
    """One S, row panels split over the ranks of a torchrun job (strong scaling; BASELINE config 5).
    `row_off_t` / `col_idx_t`: the SAME CSR on every rank (device tensors).  `reorder=False` keeps the identity
    order over the non-empty rows (label it: the permutation is then not the reference's)."""

    def __init__(self, mg: MultiGpu, row_off_t, col_idx_t, M, N, alpha=0.3, delta=0.3, block_size=0, reorder=True,
                 opts=None, tiles="auto"):
        import torch

        self.mg, self.M, self.N = mg, M, N
        R = torch.zeros(max(1, M), dtype=torch.int32, device=row_off_t.device)
        n, self.num_clusters, self.row_ms = 0, 0, 0.0
        if mg.rank == 0:
            if reorder:
                Rr, self.num_clusters, self.row_ms = host.row_reorder_dev(row_off_t, col_idx_t, M, N, alpha, block_size, opts)
            else:
                lens = row_off_t[1:] - row_off_t[:-1]  # uint32 bit patterns in int32: != 0 is all that matters
                Rr = torch.nonzero(lens != 0).flatten().to(torch.int32)
            n = Rr.numel()
            R[:n] = Rr
        self.layout, self.cuts, self.num_rows, self.col_ms, self.rphm_ms = mg.shard(row_off_t, col_idx_t, M, N, R, n,
                                                                                   delta, tiles)
        self._ro, self._ci, self._Rfull, self._delta, self._tiles = row_off_t, col_idx_t, R, delta, tiles
        self.R = R[: self.num_rows]
        self.reordered = bool(reorder)
        info = self.layout.info
        self.my_nnz = int(info.numDenseValues) + int(info.numSparseValues)

    def replicate_B(self, B):
        return self.mg.bcast(B, 0)

    def calibrate(self, dA, dB, dP, rounds=2, passes=3):
        """Cost-calibrated cuts: time `passes` passes on every rank, let sddmm_mgpu_rebalance move the cuts, repeat.
        Setup work like the reordering itself (not part of a timed step).  Returns the per-round cuts."""
        import torch

        history = []
        for _ in range(rounds):
            host.sddmm_prepare(self.layout, dA.shape[1])
            self.run(dA, dB, dP)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(passes):
                self.run(dA, dB, dP)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / passes
            self.layout, self.cuts = self.mg.rebalance(self.layout, self._ro, self._ci, self.M, self.N, self._Rfull,
                                                       self.num_rows, self._delta, ms, self._tiles)
            info = self.layout.info
            self.my_nnz = int(info.numDenseValues) + int(info.numSparseValues)
            history.append([int(c) for c in self.cuts])
        return history

    def run(self, dA, dB, dP):
        return self.mg.run(self.layout, dA, dB, dP)
