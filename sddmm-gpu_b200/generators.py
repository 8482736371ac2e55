"""Seeded synthetic sampling matrices S (CSR, uint32) and dense operands A, B.

These are the input classes BASELINE.json names:
  * uniform-random (config 2), Bernoulli masks / DLMC-style pruned masks (config 3),
  * R-MAT power-law graphs (configs 4, 5), Zipf "documents x words" (nips surrogate, config 1).

A is row-major M x K, B is COLUMN-major K x N (stored as N rows of K floats, i.e. B^T
row-major), both ~ U[0,2) fp32 like the reference's Matrix::makeData (src/Matrix.cpp:117-138),
but seeded and reproducible (the reference's generator is racy).

Within a row the column indices are ascending unless `shuffle_cols=True` (the reference's
loader keeps FILE order inside a row, src/Matrix.cpp:467, so unsorted rows are legal input).
"""
from __future__ import annotations

import numpy as np


class CSRPattern:
    """Sparsity pattern of S in the reference's CSR layout (include/Matrix.hpp:195-296)."""

    __slots__ = ("M", "N", "row_off", "col_idx", "name")

    def __init__(self, M, N, row_off, col_idx, name=""):
        self.M, self.N = int(M), int(N)
        self.row_off = np.ascontiguousarray(row_off, dtype=np.uint32)
        self.col_idx = np.ascontiguousarray(col_idx, dtype=np.uint32)
        self.name = name
        assert self.row_off.shape == (self.M + 1,)
        assert int(self.row_off[-1]) == self.col_idx.shape[0]

    @property
    def nnz(self):
        return int(self.col_idx.shape[0])

    def rows(self):
        return np.repeat(np.arange(self.M, dtype=np.uint32), np.diff(self.row_off.astype(np.int64)))


def _from_flat_positions(M, N, pos, name):
    """pos: sorted unique int64 positions in the row-major M*N grid."""
    r = pos // N
    c = (pos - r * N).astype(np.uint32)
    counts = np.bincount(r, minlength=M).astype(np.int64)
    row_off = np.zeros(M + 1, dtype=np.int64)
    np.cumsum(counts, out=row_off[1:])
    return CSRPattern(M, N, row_off.astype(np.uint32), c, name)


def _from_coo(M, N, r, c, name, shuffle_cols=False, seed=0):
    key = np.unique(r.astype(np.int64) * N + c.astype(np.int64))
    S = _from_flat_positions(M, N, key, name)
    if shuffle_cols:
        S = shuffle_within_rows(S, seed)
    return S


def shuffle_within_rows(S: CSRPattern, seed=0) -> CSRPattern:
    rng = np.random.default_rng(seed)
    rows = S.rows().astype(np.int64)
    tie = rng.random(S.nnz)
    order = np.lexsort((tie, rows))
    return CSRPattern(S.M, S.N, S.row_off, S.col_idx[order], S.name + "+shuf")


def uniform_random(M, N, density, seed, name=None) -> CSRPattern:
    """Each (i,j) kept independently with prob `density` (exact Bernoulli via geometric gaps)."""
    rng = np.random.default_rng(seed)
    total = M * N
    exp = int(total * density)
    chunks, last, got = [], -1, 0
    while True:
        n = max(1024, int((exp - got) * 1.05) + 1024)
        gaps = rng.geometric(density, size=n).astype(np.int64)
        pos = last + np.cumsum(gaps)
        keep = pos < total
        chunks.append(pos[keep])
        if not keep.all():
            break
        last = int(pos[-1])
        got += n
    pos = np.concatenate(chunks)
    return _from_flat_positions(M, N, pos, name or f"uniform_{M}x{N}_d{density}_s{seed}")


def bernoulli_mask(M, N, sparsity, seed, name=None) -> CSRPattern:
    """Random-pruning mask: keep with prob 1-sparsity (DLMC 'random pruning' look-alike)."""
    return uniform_random(M, N, 1.0 - sparsity, seed, name or f"bern_{M}x{N}_s{sparsity}_seed{seed}")


def dlmc_magnitude_mask(M, N, sparsity, seed, name=None) -> CSRPattern:
    """Magnitude-pruning look-alike: per-row keep the top (1-sparsity) of |N(0,1)*rowscale*colscale|
    with log-normal row/col scales, which gives the column-density skew real DLMC masks have."""
    rng = np.random.default_rng(seed)
    rs = np.exp(0.5 * rng.standard_normal(M)).astype(np.float32)
    cs = np.exp(0.75 * rng.standard_normal(N)).astype(np.float32)
    keep = max(1, int(round(N * (1.0 - sparsity))))
    cols = np.empty((M, keep), dtype=np.uint32)
    for i in range(M):
        w = np.abs(rng.standard_normal(N).astype(np.float32)) * cs * rs[i]
        idx = np.argpartition(-w, keep - 1)[:keep]
        cols[i] = np.sort(idx)
    row_off = (np.arange(M + 1, dtype=np.int64) * keep).astype(np.uint32)
    return CSRPattern(M, N, row_off, cols.reshape(-1), name or f"dlmc_{M}x{N}_s{sparsity}_seed{seed}")


def rmat(scale, edge_factor, seed, abcd=(0.57, 0.19, 0.19, 0.05), name=None) -> CSRPattern:
    """R-MAT graph, 2^scale vertices, edge_factor*2^scale edges before dedup, no vertex permutation."""
    rng = np.random.default_rng(seed)
    n = 1 << scale
    ne = edge_factor * n
    a, b, c, _ = abcd
    r = np.zeros(ne, np.int64)
    col = np.zeros(ne, np.int64)
    for _lvl in range(scale):
        u = rng.random(ne)
        rb = u >= a + b
        cb = ((u >= a) & (u < a + b)) | (u >= a + b + c)
        r = (r << 1) | rb
        col = (col << 1) | cb
    return _from_coo(n, n, r, col, name or f"rmat_s{scale}_ef{edge_factor}_seed{seed}")


def rmat_device(scale, edge_factor, seed, abcd=(0.57, 0.19, 0.19, 0.05)):
    """The same R-MAT recipe drawn, de-duplicated and turned into CSR ON THE CURRENT GPU (BASELINE configs 4 / 5:
    2^25 vertices never exist on the host).  Returns (row_off int32[M+1] (uint32 bit patterns), col_idx int32[nnz]
    ascending inside each row, M).  Philox with a fixed seed gives every rank of a job the same edges.
    Input generation only: torch is plumbing here, the product path never sees it."""
    import torch

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
    ro32 = torch.where(ro >= 2 ** 31, ro - 2 ** 32, ro).to(torch.int32)  # uint32 bit pattern
    del ro
    torch.cuda.empty_cache()
    return ro32, ci, M


def block_structured_scattered(M, N, n_groups, cols_per_group, row_fill, seed, noise=0.0, name=None) -> CSRPattern:
    """Block-structured like `block_structured`, but rows of a group are interleaved over the whole matrix and
    the group's columns are scattered over [0, N): the structure only shows after BSMR's row + column reordering
    (the matrices the reference's dense/sparse split is made for)."""
    rng = np.random.default_rng(seed)
    grp = rng.integers(0, n_groups, M)
    cols = [np.sort(rng.choice(N, cols_per_group, replace=False)) for _ in range(n_groups)]
    rr, cc = [], []
    for g in range(n_groups):
        rows = np.nonzero(grp == g)[0]
        keep = rng.random((rows.size, cols_per_group)) < row_fill
        ri, ci = np.nonzero(keep)
        rr.append(rows[ri])
        cc.append(cols[g][ci])
    if noise > 0:
        nn = int(M * N * noise)
        rr.append(rng.integers(0, M, nn))
        cc.append(rng.integers(0, N, nn))
    return _from_coo(M, N, np.concatenate(rr), np.concatenate(cc),
                     name or f"blockscat_{M}x{N}_g{n_groups}_c{cols_per_group}_seed{seed}")


def zipf_docs(M, N, nnz, seed, name=None) -> CSRPattern:
    """'Documents x words' with Zipf(1.0) word frequencies: surrogate for the missing
    dataset/nips.mtx (1500 x 12419, nnz 746316)."""
    rng = np.random.default_rng(seed)
    p = 1.0 / np.arange(1, N + 1)
    p /= p.sum()
    per = nnz // M
    cols, counts = [], []
    for i in range(M):
        k = per + (1 if i < nnz - per * M else 0)
        cols.append(np.sort(rng.choice(N, size=k, replace=False, p=p)))
        counts.append(k)
    row_off = np.zeros(M + 1, dtype=np.int64)
    np.cumsum(counts, out=row_off[1:])
    return CSRPattern(M, N, row_off.astype(np.uint32), np.concatenate(cols), name or f"zipf_{M}x{N}_seed{seed}")


def block_structured(M, N, n_groups, cols_per_group, row_fill, seed, noise=0.0, name=None) -> CSRPattern:
    """Rows belong to one of `n_groups` groups (interleaved), each group owns a random column set and
    each row keeps a fraction `row_fill` of it: row clustering recovers the groups and column
    reordering yields dense 16x16 blocks.  Exercises the dense (tensor-core) path."""
    rng = np.random.default_rng(seed)
    group_cols = [np.sort(rng.choice(N, size=cols_per_group, replace=False)) for _ in range(n_groups)]
    r_list, c_list = [], []
    for i in range(M):
        g = int(rng.integers(n_groups))
        gc = group_cols[g]
        sel = gc[rng.random(gc.shape[0]) < row_fill]
        if noise > 0:
            extra = rng.choice(N, size=max(1, int(noise * N)), replace=False)
            sel = np.union1d(sel, extra)
        r_list.append(np.full(sel.shape[0], i, np.int64))
        c_list.append(sel.astype(np.int64))
    return _from_coo(M, N, np.concatenate(r_list), np.concatenate(c_list),
                     name or f"blocks_{M}x{N}_g{n_groups}_seed{seed}")


def with_empty_rows(S: CSRPattern, every=7) -> CSRPattern:
    """Drop all non-zeros of every `every`-th row (the reference strips empty rows, rowReordering.cu:1081-1090)."""
    rows = S.rows()
    keep = (rows % every) != 0
    counts = np.bincount(rows[keep], minlength=S.M).astype(np.int64)
    row_off = np.zeros(S.M + 1, dtype=np.int64)
    np.cumsum(counts, out=row_off[1:])
    return CSRPattern(S.M, S.N, row_off.astype(np.uint32), S.col_idx[keep], S.name + "+empty")


def dense_operands(M, N, K, seed_a=1001, seed_b=1002):
    """A (M x K row-major) and B (stored N x K: column-major K x N), U[0,2) fp32."""
    A = (np.random.default_rng(seed_a).random((M, K), dtype=np.float32) * np.float32(2.0)).astype(np.float32)
    B = (np.random.default_rng(seed_b).random((N, K), dtype=np.float32) * np.float32(2.0)).astype(np.float32)
    return A, B


def write_mtx(path, S: CSRPattern, order="col", values=True):
    """Matrix Market coordinate file.  order='col' writes entries sorted by (col,row) like SuiteSparse
    files (so the loader's row-only stable sort yields ascending columns); order='rowrev' writes rows
    ascending but columns DESCENDING inside a row (exercises the file-order quirk)."""
    r = S.rows().astype(np.int64)
    c = S.col_idx.astype(np.int64)
    if order == "col":
        o = np.lexsort((r, c))
    elif order == "rowrev":
        o = np.lexsort((-c, r))
    else:
        o = np.arange(S.nnz)
    r, c = r[o], c[o]
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n%% generated by sddmm-gpu_b200.generators\n")
        f.write("%d %d %d\n" % (S.M, S.N, S.nnz))
        if values:
            np.savetxt(f, np.stack([r + 1, c + 1, np.ones_like(r)], 1), fmt="%d")
        else:
            np.savetxt(f, np.stack([r + 1, c + 1], 1), fmt="%d")


def write_smtx(path, S: CSRPattern):
    """DLMC `.smtx` file (the reference's second loader, src/Matrix.cpp:296-371): `M, N, nnz`, then one line
    of row offsets, then one line of column indices (0-based); values are implicitly 1."""
    with open(path, "w") as f:
        f.write("%d, %d, %d\n" % (S.M, S.N, S.nnz))
        f.write(" ".join(str(int(x)) for x in S.row_off) + "\n")
        f.write(" ".join(str(int(x)) for x in S.col_idx) + "\n")


def write_snap_txt(path, S: CSRPattern, seed=0):
    """SNAP-style edge list (the reference's `.txt` loader, src/Matrix.cpp:483-580): '#' header with
    `Nodes: n Edges: e`, then `from<TAB>to` lines.  Node ids are scrambled (the loader renumbers them in order of
    first appearance) and the edge order is shuffled; S must be square."""
    assert S.M == S.N
    rng = np.random.default_rng(seed)
    label = rng.permutation(10 * S.M)[: S.M]  # non-contiguous ids
    r, c = S.rows().astype(np.int64), S.col_idx.astype(np.int64)
    o = rng.permutation(S.nnz)
    with open(path, "w") as f:
        f.write("# Directed graph: synthetic\n# generated by sddmm-gpu_b200.generators\n")
        f.write("# Nodes: %d Edges: %d\n# FromNodeId\tToNodeId\n" % (S.M, S.nnz))
        np.savetxt(f, np.stack([label[r[o]], label[c[o]]], 1), fmt="%d", delimiter="\t")


def write_case_bin(path, S: CSRPattern, A, B):
    """Case file consumed by oracle/_ref/ref_dump (oracle/ref_dump_main.cu)."""
    K = A.shape[1]
    with open(path, "wb") as f:
        np.array([0x314D4453, S.M, S.N, S.nnz, K], dtype=np.uint32).tofile(f)
        S.row_off.tofile(f)
        S.col_idx.tofile(f)
        np.ascontiguousarray(A, dtype=np.float32).tofile(f)
        np.ascontiguousarray(B, dtype=np.float32).tofile(f)
