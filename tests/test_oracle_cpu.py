"""CPU tests (-m "not gpu"): the oracle against the reference's golden vectors, host logic, C-ABI exports.

Pins:
  * tests/golden/ref_gpu/*.npz  -- arrays produced by the UNMODIFIED reference GPU pipeline on a B200
    (tests/make_ref_goldens.py): row permutation for every case, full BSMR/RPHM arrays for small ones.
  * oracle/_ref/libref_cpu.so   -- the reference's own colReordering_cpu / sddmm_cpu / loader / checkData,
    compared live when the library is present (it is built from /root/reference by build()).
"""
import json
import os

import numpy as np
import pytest

from cases import GOLDEN, gen, operands, pkg, small_cases
from oracle import oracle as O

REF_GPU = os.path.join(GOLDEN, "ref_gpu")


def _golden_specs():
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    import make_ref_goldens as m
    return {c[0]: c for c in m.cases(big=False)}


SPECS = _golden_specs()


@pytest.mark.parametrize("name", sorted(SPECS))
def test_oracle_matches_reference_gpu_golden(name):
    """oracle/bsmr_oracle.c reproduces the reference's reorderedRows (and, where stored, every BSMR/RPHM
    array) bit for bit."""
    path = os.path.join(REF_GPU, name + ".npz")
    if not os.path.exists(path):
        pytest.skip("golden not generated yet")
    _, S, K, alpha, delta, _ = SPECS[name]
    g = np.load(path)
    bs = int(g["block_size"])
    rr = O.row_reorder(S, alpha, bs)
    assert np.array_equal(rr["reorderedRows"], g["reorderedRows"])
    assert rr["numClusters"] == int(g["num_clusters"])
    if "denseCols" in g.files:
        cr = O.col_reorder(S, g["reorderedRows"], delta)
        for k in ("denseCols", "denseColOffsets", "sparseCols", "sparseColOffsets", "sparseValueOffsets"):
            assert np.array_equal(cr[k], g[k]), k
        rp = O.rphm_build(S, g["reorderedRows"], cr)
        for k in ("blockOffsets", "blockValues", "sparseValues", "sparseRelativeRows", "sparseColIndices",
                  "denseRowPanelIds", "denseColBlockIters", "sparseRowPanelIds", "sparseColBlockIters"):
            assert np.array_equal(rp[k], g[k]), k


def test_golden_summary_says_oracle_equals_reference():
    """The on-box comparison recorded by make_ref_goldens.py: every array of every case equal."""
    p = os.path.join(REF_GPU, "summary.json")
    if not os.path.exists(p):
        pytest.skip("no summary")
    recs = json.load(open(p))
    checked = 0
    for r in recs:
        eq = r.get("oracle_equals_reference")
        if not eq:
            continue
        bad = [k for k, v in eq.items() if v is False]
        assert not bad, (r["name"], bad)
        checked += 1
    assert checked >= 20


needs_ref = pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref/libref_cpu.so not built")


@needs_ref
@pytest.mark.parametrize("delta", [0.0, 0.1, 0.3, 0.5, 0.7, 0.9, 1.1])
def test_col_reorder_equals_reference_host_code(delta):
    for S in (gen.uniform_random(512, 512, 0.1, 7), gen.rmat(11, 8, 3), gen.zipf_docs(100, 3000, 9000, 2),
              gen.with_empty_rows(gen.bernoulli_mask(200, 333, 0.7, 4), 5)):
        bs = 16
        R = O.row_reorder(S, 0.3, bs)["reorderedRows"]
        a, b = O.col_reorder(S, R, delta), O.ref_col_reorder(S, R, delta)
        for k in ("denseCols", "denseColOffsets", "sparseCols", "sparseColOffsets", "sparseValueOffsets"):
            assert np.array_equal(a[k], b[k]), (S.name, k)


@needs_ref
@pytest.mark.parametrize("K", [32, 64, 100])
def test_sddmm_cpu_bitexact_with_reference(K):
    S = gen.rmat(10, 8, 1)
    A, B = operands(S, K)
    assert np.array_equal(O.sddmm_cpu(S, A, B, threads=2), O.ref_sddmm_cpu(S, A, B))


@needs_ref
def test_check_data_equals_reference():
    rng = np.random.default_rng(0)
    a = rng.random(20000, dtype=np.float32) * 50
    b = a * (1 + rng.normal(0, 7e-4, a.shape).astype(np.float32))
    b[::97] += 2e-5
    a[::101] = 0
    assert O.check_data(a, b) == O.ref_check_data(a, b)
    assert O.check_data(a, a) == 0


@needs_ref
@pytest.mark.parametrize("order", ["col", "rowrev"])
def test_mtx_loader_equals_reference(tmp_path, order):
    S = gen.with_empty_rows(gen.uniform_random(64, 80, 0.1, 3), 5)
    p = str(tmp_path / "a.mtx")
    gen.write_mtx(p, S, order=order)
    rc, mine = O.load_mtx(p)
    rc2, ref = O.ref_load_mtx(p)
    assert rc == 0 and rc2 == 0
    assert mine[0] == ref[0] and mine[1] == ref[1]
    for x, y in zip(mine[2:], ref[2:]):
        assert np.array_equal(x, y)
    if order == "rowrev":  # file order inside a row is kept (src/Matrix.cpp:467)
        r0 = mine[3][mine[2][1]:mine[2][2]]
        assert np.all(np.diff(r0.astype(np.int64)) < 0) or r0.size < 2


@needs_ref
def test_mtx_loader_rejects_like_reference(tmp_path):
    p = str(tmp_path / "dup.mtx")
    open(p, "w").write("%%MatrixMarket\n3 3 3\n1 1 1\n2 2 1\n1 1 1\n")
    assert O.load_mtx(p)[0] != 0 and O.ref_load_mtx(p)[0] != 0
    p = str(tmp_path / "one.mtx")
    open(p, "w").write("%%MatrixMarket\n3 3 1\n1 1 1\n")
    assert O.load_mtx(p)[0] != 0 and O.ref_load_mtx(p)[0] != 0
    p = str(tmp_path / "noval.mtx")
    open(p, "w").write("%%MatrixMarket\n3 3 2\n1 2\n3 1\n")
    rc, m = O.load_mtx(p)
    rc2, r = O.ref_load_mtx(p)
    assert rc == 0 and rc2 == 0 and np.array_equal(m[3], r[3]) and np.array_equal(m[4], r[4])


def test_block_size_rule():
    assert O.block_size(1500, 12419, 180e9) == 16
    assert O.block_size(100000, 100000, 180e9) == 17       # SMEM term: ceil(400000/24576)
    assert O.block_size(4194304, 4194304, 180e9) >= 683
    assert O.nbpr(12419, 16) == 777 and O.cluster_blockdim(777) == 224
    assert O.kept_warps(224).tolist() == [1, 1, 0, 1, 1, 0, 0]
    assert O.kept_warps(96).tolist() == [1, 1, 0]
    assert O.kept_warps(256).tolist() == [1] * 8


def test_layout_invariants_from_reference_checkers():
    """The invariants of check_rowReordering / check_colReordering / check_rphm (src/BSMR.cpp:444-824)."""
    for name, (S, alpha, delta, K) in small_cases().items():
        bs = O.block_size(S.M, S.N, 180e9)
        R = O.row_reorder(S, alpha, bs)["reorderedRows"]
        lens = np.diff(S.row_off.astype(np.int64))
        assert np.array_equal(np.sort(R), np.nonzero(lens)[0]), name          # every non-empty row once
        cr = O.col_reorder(S, R, delta)
        rp = O.rphm_build(S, R, cr)
        bv = rp["blockValues"]
        used = np.concatenate([bv[bv != O.NULL_VALUE], rp["sparseValues"]])
        assert np.array_equal(np.sort(used), np.arange(S.nnz)), name          # each nnz exactly once
        assert np.all(cr["denseColOffsets"] % 16 == 0)


def test_c_abi_library_loads_and_exports_every_declared_symbol():
    syms = pkg.declared_symbols()
    assert len(syms) >= 18 and "sddmm_run_dev" in syms and "bsmr_row_reorder_dev" in syms
    L = pkg.lib()  # raises ImportError if the .so or any symbol is missing
    assert L.sddmm_b200_abi_version() == 2


def test_host_traffic_entry_points_check_their_arguments():
    """sddmm_host_traffic / sddmm_mgpu_host_traffic (the bytes a host-buffer pass moved): present in the ABI and
    strict about null handles -- an argument error, never a crash; no device is touched."""
    import ctypes as C
    L = pkg.lib()
    a, b = C.c_uint64(7), C.c_uint64(7)
    assert L.sddmm_host_traffic(None, C.byref(a), C.byref(b)) == 1      # SDDMM_E_ARG
    assert L.sddmm_mgpu_host_traffic(None, C.byref(a)) == 1
    assert b"null" in L.sddmm_last_error().lower() or b"argument" in L.sddmm_last_error().lower()


def test_product_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    S = gen.uniform_random(64, 64, 0.1, 1)
    A, B = operands(S, 32)
    with pytest.raises(pkg.SddmmError):
        pkg.sddmm(S, A, B)


def test_shard_plan_balances_nnz():
    S = gen.rmat(12, 8, 4)
    R = np.nonzero(np.diff(S.row_off.astype(np.int64)))[0].astype(np.uint32)
    cuts = pkg.shard_plan(S, R, 4)
    P = (len(R) + 15) // 16
    assert cuts[0] == 0 and cuts[-1] == P and np.all(np.diff(cuts.astype(np.int64)) >= 0)
    lens = np.diff(S.row_off.astype(np.int64))[R]
    per = [lens[cuts[i] * 16: cuts[i + 1] * 16].sum() for i in range(4)]
    assert max(per) - min(per) <= 0.25 * S.nnz / 4 + lens.max() * 16


def _fnv(a):
    h = 1469598103934665603
    for b in np.ascontiguousarray(a).tobytes():
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.mark.parametrize("order", ["col", "rowrev"])
def test_cpp_host_loader_matches_oracle(tmp_path, order):
    """csrc/host/Matrix.cpp (parallel in-place parser) builds the same CSR as the oracle's restatement of
    src/Matrix.cpp:398-480 (itself pinned to the reference loader above); `BSMR-sddmm -x 1` needs no GPU."""
    import subprocess
    from cases import ROOT
    exe = os.path.join(ROOT, "sddmm-gpu_b200", "BSMR-sddmm")
    if not os.access(exe, os.X_OK):
        pytest.skip("CLI not built")
    S = gen.with_empty_rows(gen.rmat(9, 8, 3), 5)
    p = str(tmp_path / "m.mtx")
    gen.write_mtx(p, S, order=order)
    rc, (M, N, ro, ci, va) = O.load_mtx(p)
    assert rc == 0
    out = subprocess.run([exe, "-f", p, "-x", "1"], capture_output=True, text=True, timeout=60).stdout
    line = [l for l in out.splitlines() if l.startswith("[loader")][0]
    tok = line.strip("[]").split()
    got = dict(zip(tok[2::2], tok[3::2]))
    assert int(got["M"]) == M and int(got["N"]) == N and int(got["nnz"]) == len(ci)
    assert int(got["rowOff"], 16) == _fnv(ro) and int(got["colIdx"], 16) == _fnv(ci) and int(got["values"], 16) == _fnv(va)
    # rejects what the reference rejects
    bad = str(tmp_path / "dup.mtx")
    open(bad, "w").write("%%MatrixMarket\n3 3 3\n1 1 1\n2 2 1\n1 1 1\n")
    r = subprocess.run([exe, "-f", bad, "-x", "1"], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "duplicate" in r.stderr


@needs_ref
@pytest.mark.parametrize("delta", [0.1, 0.3, 0.6])
def test_evaluation_reordering_equals_reference_host_code(delta):
    """oracle_evaluation_reordering / oracle_original_block_stats against the reference's own evaluationReordering
    (src/BSMR.cpp:826-994), run on the CPU through its BSMR::colReordering."""
    for S in (gen.block_structured(512, 768, 6, 96, 0.8, seed=11, noise=0.004), gen.rmat(10, 8, 3),
              gen.with_empty_rows(gen.bernoulli_mask(200, 333, 0.6, 4), 5), gen.uniform_random(37, 50, 0.3, 9),
              gen.dlmc_magnitude_mask(256, 256, 0.7, 30)):
        R = O.row_reorder(S, 0.3, 16)["reorderedRows"]
        cr = O.col_reorder(S, R, delta)
        a, b = O.evaluation_reordering(S, R, cr, delta), O.ref_evaluation_reordering(S, R, delta)
        for k in b:
            if isinstance(b[k], float):
                assert a[k] == pytest.approx(b[k], rel=1e-6, abs=1e-7), (S.name, k, a[k], b[k])
            else:
                assert a[k] == b[k], (S.name, k, a[k], b[k])


@needs_ref
def test_cpp_host_smtx_loader_matches_reference(tmp_path):
    """csrc/host/Matrix.cpp's `.smtx` loader (DLMC format) against the reference's own (src/Matrix.cpp:296-371)
    through `BSMR-sddmm -x 1`; duplicates inside a row are rejected by both."""
    import subprocess
    from cases import ROOT
    exe = os.path.join(ROOT, "sddmm-gpu_b200", "BSMR-sddmm")
    if not os.access(exe, os.X_OK):
        pytest.skip("CLI not built")
    S = gen.shuffle_within_rows(gen.with_empty_rows(gen.dlmc_magnitude_mask(96, 200, 0.8, 5), 3), 2)
    p = str(tmp_path / "m.smtx")
    gen.write_smtx(p, S)
    rc, (M, N, ro, ci, va) = O.ref_load_mtx(p)
    assert rc == 0 and M == S.M and N == S.N and np.array_equal(ci, S.col_idx) and np.all(va == 1.0)
    out = subprocess.run([exe, "-f", p, "-x", "1"], capture_output=True, text=True, timeout=60).stdout
    tok = [l for l in out.splitlines() if l.startswith("[loader")][0].strip("[]").split()
    got = dict(zip(tok[2::2], tok[3::2]))
    assert int(got["M"]) == M and int(got["N"]) == N and int(got["nnz"]) == len(ci)
    assert int(got["rowOff"], 16) == _fnv(ro) and int(got["colIdx"], 16) == _fnv(ci) and int(got["values"], 16) == _fnv(va)
    bad = str(tmp_path / "dup.smtx")
    open(bad, "w").write("2, 4, 3\n0 2 3\n1 1 2\n")
    assert O.ref_load_mtx(bad)[0] != 0
    r = subprocess.run([exe, "-f", bad, "-x", "1"], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "duplicate" in r.stderr


_MTX_QUIRKS = {
    "crlf": "%%MatrixMarket matrix coordinate real general\r\n3 4 4\r\n1 1 1.5\r\n2 3 -2\r\n3 4 1e-3\r\n1 2 7\r\n",
    "blank_lines": "%%MatrixMarket x\n% c\n3 4 4\n\n1 1 1.5\n\n2 3 -2\n3 4 1e-3\n1 2 7\n\n",
    "tabs": "%%MatrixMarket x\n3\t4\t4\n1\t1\t1.5\n2\t3\t-2\n3\t4\t1e-3\n1\t2\t7\n",
    "pattern": "%%MatrixMarket matrix coordinate pattern general\n3 4 4\n1 1\n2 3\n3 4\n1 2\n",
    "no_trailing_newline": "%%MatrixMarket x\n3 4 4\n1 1 1.5\n2 3 -2\n3 4 1e-3\n1 2 7",
    "trailing_spaces": "%%MatrixMarket x\n3 4 4 \n1 1 1.5 \n2 3 -2  \n3 4 1e-3\n1 2 7 \n",
    "int_values": "%%MatrixMarket matrix coordinate integer general\n3 4 4\n1 1 2\n2 3 -2\n3 4 5\n1 2 7\n",
    "too_few": "%%MatrixMarket x\n3 4 5\n1 1 1.5\n2 3 -2\n3 4 1e-3\n1 2 7\n",
    "too_many": "%%MatrixMarket x\n3 4 3\n1 1 1.5\n2 3 -2\n3 4 1e-3\n1 2 7\n",
    "row_oob": "%%MatrixMarket x\n3 4 4\n1 1 1.5\n4 3 -2\n3 4 1e-3\n1 2 7\n",
    "col_oob": "%%MatrixMarket x\n3 4 4\n1 1 1.5\n2 5 -2\n3 4 1e-3\n1 2 7\n",
    "zero_index": "%%MatrixMarket x\n3 4 4\n0 1 1.5\n2 3 -2\n3 4 1e-3\n1 2 7\n",
    "dup": "%%MatrixMarket x\n3 4 4\n1 1 1.5\n2 3 -2\n1 1 1e-3\n1 2 7\n",
    "nnz1": "%%MatrixMarket x\n3 4 1\n1 1 1.5\n",
}


@needs_ref
@pytest.mark.parametrize("name", sorted(_MTX_QUIRKS))
def test_cpp_host_loader_quirks_match_reference(tmp_path, name):
    """accept / reject decisions and the loaded arrays of csrc/host/Matrix.cpp against the reference's loader
    (src/Matrix.cpp:398-480) on line-ending, separator, value-format and error cases.  (Leading blanks and
    comments between entries make the reference throw from std::stoi; those are outside its contract.)"""
    import subprocess
    from cases import ROOT
    exe = os.path.join(ROOT, "sddmm-gpu_b200", "BSMR-sddmm")
    if not os.access(exe, os.X_OK):
        pytest.skip("CLI not built")
    p = str(tmp_path / (name + ".mtx"))
    with open(p, "w", newline="") as f:
        f.write(_MTX_QUIRKS[name])
    rc, res = O.ref_load_mtx(p)
    r = subprocess.run([exe, "-f", p, "-x", "1"], capture_output=True, text=True, timeout=60)
    line = [l for l in r.stdout.splitlines() if l.startswith("[loader")]
    if rc != 0:
        assert r.returncode != 0 and not line
        return
    M, N, ro, ci, va = res
    assert r.returncode == 0 and line
    tok = line[0].strip("[]").split()
    got = dict(zip(tok[2::2], tok[3::2]))
    assert int(got["M"]) == M and int(got["N"]) == N and int(got["nnz"]) == len(ci)
    assert int(got["rowOff"], 16) == _fnv(ro) and int(got["colIdx"], 16) == _fnv(ci) and int(got["values"], 16) == _fnv(va)
    # the oracle's restatement agrees as well
    rc2, res2 = O.load_mtx(p)
    assert rc2 == 0 and np.array_equal(res2[2], ro) and np.array_equal(res2[3], ci) and np.array_equal(res2[4], va)


@needs_ref
def test_col_reorder_and_evaluation_random_matrices_vs_reference():
    """randomised sweep (fixed seeds): tiny ragged shapes, any row order (not just the clustering's), every delta
    regime -- oracle column reordering + statistics against the reference host code."""
    rng = np.random.default_rng(20240)
    for trial in range(40):
        M, N = int(rng.integers(1, 90)), int(rng.integers(1, 140))
        dens = float(rng.choice([0.02, 0.1, 0.35, 0.8]))
        S = gen.uniform_random(M, N, dens, int(rng.integers(1, 1 << 30)))
        if S.nnz < 2:
            continue
        nz = np.nonzero(np.diff(S.row_off.astype(np.int64)))[0].astype(np.uint32)
        R = rng.permutation(nz).astype(np.uint32)
        delta = float(rng.choice([0.0, 0.05, 0.3, 0.5, 0.99, 1.1]))
        a, b = O.col_reorder(S, R, delta), O.ref_col_reorder(S, R, delta)
        for k in ("denseCols", "denseColOffsets", "sparseCols", "sparseColOffsets", "sparseValueOffsets"):
            assert np.array_equal(a[k], b[k]), (trial, M, N, dens, delta, k)
        ea, eb = O.evaluation_reordering(S, R, a, delta), O.ref_evaluation_reordering(S, R, delta)
        for k in eb:
            if isinstance(eb[k], float):
                assert ea[k] == pytest.approx(eb[k], rel=1e-6, abs=1e-7), (trial, k)
            else:
                assert ea[k] == eb[k], (trial, M, N, dens, delta, k, ea[k], eb[k])


@pytest.mark.parametrize("name", sorted(n for n, c in small_cases().items() if c[1] >= 0))
def test_pruned_clustering_equals_literal_oracle(name):
    """oracle_row_reorder_pruned (candidates from an inverted index over non-zero blocks + the zero-norm rows,
    sparse evaluation of the literal reduction tree) must reproduce the literal oracle exactly: the pruning
    argument behind the round-2 device design, checked on every parity case incl. the dropped-warp ones."""
    S, alpha, _delta, _K = small_cases()[name]
    bs = O.block_size(S.M, S.N, 180e9)
    a = O.row_reorder(S, alpha, bs)
    for prefix in (False, True):  # True: also skip the representative's most frequent blocks while their mass <= alpha
        b = O.row_reorder_pruned(S, alpha, bs, prefix)
        assert np.array_equal(a["reorderedRows"], b["reorderedRows"]), prefix
        assert a["numClusters"] == b["numClusters"]
        assert np.array_equal(a["clusterOfRow"], b["clusterOfRow"])


def test_pruned_clustering_random_and_edge_cases():
    rng = np.random.default_rng(77)
    for trial in range(25):
        M, N = int(rng.integers(1, 120)), int(rng.integers(1, 400))
        S = gen.with_empty_rows(gen.uniform_random(M, N, float(rng.choice([0.01, 0.05, 0.3])), int(rng.integers(1, 1 << 30))), 2)
        alpha = float(rng.choice([0.0, 0.1, 0.3, 0.6, 0.95, 1.0]))
        bs = int(rng.choice([16, 23, 64]))
        a, b = O.row_reorder(S, alpha, bs), O.row_reorder_pruned(S, alpha, bs, bool(trial & 1))
        assert np.array_equal(a["reorderedRows"], b["reorderedRows"]), (trial, M, N, alpha, bs)
        assert a["numClusters"] == b["numClusters"]


@pytest.mark.skipif(not os.environ.get("SDDMM_SLOW_TESTS"), reason="8 minutes of CPU; set SDDMM_SLOW_TESTS=1")
def test_pruned_oracle_reproduces_reference_gpu_at_65k_rows():
    """R-MAT scale 16 (65 536 rows, 24 047 clusters): the pruned oracle against the permutation the UNMODIFIED
    reference GPU pipeline produced on a B200 (tests/golden/ref_gpu/rmat16_k32.npz).  The literal oracle cannot
    reach this size (dense M x nbpr encoding, all-pairs evaluation).  Last run: equal, 137 384 802 evaluations."""
    import make_ref_goldens as m
    S = [c for c in m.cases(big=True) if c[0] == "rmat16_k32"][0][1]
    z = np.load(os.path.join(REF_GPU, "rmat16_k32.npz"))
    r = O.row_reorder_pruned(S, 0.3, int(z["block_size"]))
    assert np.array_equal(r["reorderedRows"], z["reorderedRows"]) and r["numClusters"] == int(z["num_clusters"])


@needs_ref
def test_cpp_host_graph_txt_loader_matches_reference(tmp_path):
    """SNAP-style `.txt` edge lists (src/Matrix.cpp:483-580: '# Nodes: n Edges: e' header, node ids renumbered
    in order of first appearance, shuffled edges): csrc/host/Matrix.cpp against the reference's own loader."""
    import subprocess
    from cases import ROOT
    exe = os.path.join(ROOT, "sddmm-gpu_b200", "BSMR-sddmm")
    if not os.access(exe, os.X_OK):
        pytest.skip("CLI not built")
    S = gen.rmat(9, 6, 12)
    p = str(tmp_path / "g.txt")
    gen.write_snap_txt(p, S, seed=3)
    rc, (M, N, ro, ci, va) = O.ref_load_mtx(p)
    assert rc == 0 and M == S.M and len(ci) == S.nnz
    out = subprocess.run([exe, "-f", p, "-x", "1"], capture_output=True, text=True, timeout=60).stdout
    tok = [l for l in out.splitlines() if l.startswith("[loader")][0].strip("[]").split()
    got = dict(zip(tok[2::2], tok[3::2]))
    assert int(got["M"]) == M and int(got["N"]) == N and int(got["nnz"]) == len(ci)
    assert int(got["rowOff"], 16) == _fnv(ro) and int(got["colIdx"], 16) == _fnv(ci) and int(got["values"], 16) == _fnv(va)
    for name, text in (("dup", "# Nodes: 3 Edges: 3\n5\t7\n7\t9\n5\t7\n"), ("few", "# Nodes: 3 Edges: 4\n5\t7\n7\t9\n9\t5\n"),
                       ("many", "# Nodes: 3 Edges: 2\n5\t7\n7\t9\n9\t5\n"), ("nodes", "# Nodes: 2 Edges: 3\n5\t7\n7\t9\n9\t5\n"),
                       ("nohdr", "5\t7\n7\t9\n")):
        q = str(tmp_path / (name + ".txt"))
        open(q, "w").write(text)
        assert O.ref_load_mtx(q)[0] != 0, name
        r = subprocess.run([exe, "-f", q, "-x", "1"], capture_output=True, text=True, timeout=60)
        assert r.returncode != 0, name


def test_cli_log_is_readable_by_the_reference_analyser(tmp_path):
    """`scripts/analyze_results.cpp` of the reference (compiled unmodified into oracle/_ref/analyze_results) must
    parse the [key : value] record our CLI writes: file, M, N, NNZ, sparsity, K and bsmr_gflops end up in its CSV.
    (`-x 2` prints the record with a nominal 1 ms pass and needs no GPU.)"""
    import subprocess
    from cases import ROOT
    exe = os.path.join(ROOT, "sddmm-gpu_b200", "BSMR-sddmm")
    ana = os.path.join(ROOT, "oracle", "_ref", "analyze_results")
    if not (os.access(exe, os.X_OK) and os.access(ana, os.X_OK)):
        pytest.skip("CLI or oracle/_ref/analyze_results not built")
    S = gen.rmat(9, 8, 3)
    mtx = str(tmp_path / "m.mtx")
    gen.write_mtx(mtx, S, order="col")
    log = str(tmp_path / "results.log")
    with open(log, "w") as f:
        for K in (64,):
            out = subprocess.run([exe, "-f", mtx, "-k", str(K), "-x", "2"], capture_output=True, text=True, timeout=60)
            assert out.returncode == 0
            f.write(out.stdout[out.stdout.index("---New data---"):])
    r = subprocess.run([ana, log], capture_output=True, text=True, timeout=120, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-300:]
    csv = [l for l in open(str(tmp_path / "results_64.csv")).read().splitlines() if l]
    assert csv[0].startswith("file,M,N,NNZ,Sparsity,K,BSMR")
    row = csv[1].split(",")
    assert row[0] == mtx and int(row[1]) == S.M and int(row[2]) == S.N and int(row[3]) == S.nnz and int(row[5]) == 64
    # gflops at 1 ms; two decimals, as in the reference's log (std::fixed/setprecision(2) stay set after "sparsity")
    assert float(row[6]) == pytest.approx(2.0 * S.nnz * 64 / 1e6, abs=0.006)


@pytest.mark.parametrize("name", ["bern4096s70_k64", "dlmc4096s90_k64"])
def test_pruned_oracle_matches_reference_gpu_golden_4096(name):
    """the 4096 x 4096 config-3 masks: permutation of the unmodified reference GPU pipeline (golden) == pruned oracle."""
    import make_ref_goldens as m
    S, alpha = [(c[1], c[3]) for c in m.cases(big=True) if c[0] == name][0]
    z = np.load(os.path.join(REF_GPU, name + ".npz"))
    r = O.row_reorder_pruned(S, alpha, int(z["block_size"]))
    assert np.array_equal(r["reorderedRows"], z["reorderedRows"]) and r["numClusters"] == int(z["num_clusters"])
