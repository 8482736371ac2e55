"""GPU parity tests (-m gpu): the CUDA path, through the C ABI, against the oracle on the same inputs.

Bit-exact for everything integer (dispersion, permutation, BSMR/RPHM arrays); the reference's own
checkData tolerance (include/checkData.hpp:14-30: |a-b| < 1e-5 or |a-b|/max(|a|,|b|,1e-3) < 1e-3) for P.
"""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from cases import GOLDEN, ROOT, gen, operands, pkg, small_cases
from oracle import oracle as O

pytestmark = pytest.mark.gpu

CASES = small_cases()
LAYOUT_KEYS = ("denseCols", "denseColOffsets", "sparseCols", "sparseColOffsets", "sparseValueOffsets")
RPHM_KEYS = ("blockOffsets", "blockValues", "sparseValues", "sparseRelativeRows", "sparseColIndices",
             "denseRowPanelIds", "denseColBlockIters", "sparseRowPanelIds", "sparseColBlockIters")


def _bs(S):
    return O.block_size(S.M, S.N, 180e9)


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    assert torch.cuda.is_available()
    return torch


def _dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32) if a.dtype == np.uint32 else np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("name", sorted(CASES))
def test_dispersion_bitexact(name, torch_mod):
    S, alpha, delta, K = CASES[name]
    bs = _bs(S)
    d, nb = pkg.dispersion_dev(_dev(torch_mod, S.row_off), _dev(torch_mod, S.col_idx), S.M, S.N, bs)
    _, disp = O.encode_dispersion(S, bs, dense=False)
    assert nb == O.nbpr(S.N, bs)
    assert np.array_equal(d.cpu().numpy().view(np.uint32), disp)


@pytest.mark.parametrize("name", sorted(CASES))
def test_row_reorder_bitexact(name):
    S, alpha, delta, K = CASES[name]
    bs = _bs(S)
    b = pkg.BSMR().rowReordering(alpha, S, block_size=bs)
    rr = O.row_reorder(S, alpha, bs)
    assert np.array_equal(b.reorderedRows(), rr["reorderedRows"])
    assert b.numClusters() == rr["numClusters"]


@pytest.mark.parametrize("alpha", [0.0, 0.2, 0.45, 0.6, 0.75, 1.0, -0.5])
def test_row_reorder_alpha_sweep(alpha):
    S = gen.block_structured(300, 2000, 7, 120, 0.7, seed=5, noise=0.003)
    bs = _bs(S)
    b = pkg.BSMR().rowReordering(alpha, S, block_size=bs)
    rr = O.row_reorder(S, alpha, bs)
    assert np.array_equal(b.reorderedRows(), rr["reorderedRows"])


@pytest.mark.parametrize("bs", [16, 23, 64, 200])
def test_row_reorder_explicit_block_size(bs):
    S = gen.rmat(11, 8, 9)
    b = pkg.BSMR().rowReordering(0.25, S, block_size=bs)
    rr = O.row_reorder(S, 0.25, bs)
    assert np.array_equal(b.reorderedRows(), rr["reorderedRows"])


@pytest.mark.parametrize("name", sorted(CASES))
def test_layout_bitexact(name):
    """column reordering + RPHM arrays from the ORACLE's row order (isolates this stage)."""
    S, alpha, delta, K = CASES[name]
    R = O.row_reorder(S, alpha, _bs(S))["reorderedRows"]
    b = pkg.BSMR().colReordering(delta, S, R)
    cr = O.col_reorder(S, R, delta)
    rp = O.rphm_build(S, R, cr)
    got = b.layout().arrays()
    for k in LAYOUT_KEYS:
        assert np.array_equal(got[k], cr[k]), k
    for k in RPHM_KEYS:
        assert np.array_equal(got[k], rp[k]), k
    info = b.layout().info
    assert info.numDenseBlocks == rp["blockOffsets"][-1]
    assert info.numDenseThreadBlocks == rp["numDenseThreadBlocks"]
    assert info.numSparseThreadBlocks == rp["numSparseThreadBlocks"]
    assert info.maxNumDenseColBlocksInRowPanel == rp["maxNumDenseColBlocksInRowPanel"]
    assert info.maxNumSparseColBlocksInRowPanel == rp["maxNumSparseColBlocksInRowPanel"]


@pytest.mark.parametrize("name", sorted(CASES))
def test_evaluation_reordering_matches_oracle(name):
    """the statistics the reference logs after every run (evaluationReordering, BSMR.cpp:826-994), computed on the
    device; integers exact, the two average densities to float rounding."""
    S, alpha, delta, K = CASES[name]
    R = O.row_reorder(S, alpha, _bs(S))["reorderedRows"]
    b = pkg.BSMR().colReordering(delta, S, R)
    got = pkg.evaluationReordering(S, b, delta)
    want = O.evaluation_reordering(S, R, O.col_reorder(S, R, delta), delta)
    for k, v in want.items():
        if isinstance(v, float):
            assert got[k] == pytest.approx(v, rel=2e-6, abs=1e-7), (k, got[k], v)
        else:
            assert got[k] == v, (k, got[k], v)


def test_original_block_stats_edges_and_threshold():
    """clipped edge blocks (M, N not multiples of 16), every threshold side, empty rows."""
    S = gen.with_empty_rows(gen.dlmc_magnitude_mask(203, 117, 0.5, 3), 4)
    R = np.nonzero(np.diff(S.row_off.astype(np.int64)))[0].astype(np.uint32)
    for delta in (0.0, 0.25, 0.5, 0.75, 1.0, 1.5):
        lay = pkg.BSMR().colReordering(delta, S, R)
        got = pkg.evaluationReordering(S, lay, delta)
        want = O.evaluation_reordering(S, R, O.col_reorder(S, R, delta), delta)
        assert got["originalNumDenseBlock"] == want["originalNumDenseBlock"], delta
        assert got["originalAverageDensity"] == pytest.approx(want["originalAverageDensity"], rel=2e-6, abs=1e-7)
        assert got["numDenseBlock"] == want["numDenseBlock"] and got["numSparseData"] == want["numSparseData"]
        assert got["averageDensity"] == pytest.approx(want["averageDensity"], rel=2e-6, abs=1e-7)


@pytest.mark.parametrize("delta", [0.0, 0.05, 0.5, 0.9, 1.1])
def test_layout_delta_sweep(delta):
    S = gen.dlmc_magnitude_mask(256, 512, 0.8, 12)
    R = O.row_reorder(S, 0.3, 16)["reorderedRows"]
    got = pkg.BSMR().colReordering(delta, S, R).layout().arrays()
    cr = O.col_reorder(S, R, delta)
    rp = O.rphm_build(S, R, cr)
    for k in LAYOUT_KEYS:
        assert np.array_equal(got[k], cr[k]), k
    for k in RPHM_KEYS:
        assert np.array_equal(got[k], rp[k]), k


@pytest.mark.parametrize("name", sorted(CASES))
def test_sddmm_values_within_reference_tolerance(name):
    """whole path on host buffers == sddmm(options, A, B, P, logger), checked with checkData."""
    S, alpha, delta, K = CASES[name]
    A, B = operands(S, K)
    res = pkg.sddmm(S, A, B, alpha=alpha, delta=delta, block_size=_bs(S))
    Pref = O.sddmm_cpu(S, A, B)
    assert O.check_data(Pref, res["P"]) == 0
    assert res["gpu_launches"] > 0
    rr = O.row_reorder(S, alpha, _bs(S))
    assert np.array_equal(res["reorderedRows"], rr["reorderedRows"])


@pytest.mark.parametrize("K", [4, 32, 36, 64, 96, 128, 256, 512])
@pytest.mark.parametrize("delta", [0.0, 0.3, 1.1])
def test_sddmm_k_sweep(K, delta):
    """delta=0: everything through the tcgen05 dense kernel; 1.1: everything residual; 0.3: both."""
    S = gen.block_structured(200, 300, 4, 64, 0.8, seed=2, noise=0.01)
    A, B = operands(S, K)
    res = pkg.sddmm(S, A, B, alpha=0.3, delta=delta, block_size=16)
    if delta == 0.0:
        assert res["numSparseValues"] == 0
    if delta > 1.0:
        assert res["numDenseBlocks"] == 0
    assert O.check_data(O.sddmm_cpu(S, A, B), res["P"]) == 0


def test_residual_kernel_is_fp32_accurate():
    """residual path is fp32 FMA: much tighter than the tf32 tolerance."""
    S = gen.rmat(11, 8, 2)
    A, B = operands(S, 128)
    res = pkg.sddmm(S, A, B, alpha=0.3, delta=1.1, block_size=16)
    Pref = O.sddmm_cpu(S, A, B)
    rel = np.abs(res["P"] - Pref) / np.maximum(np.abs(Pref), 1e-3)
    assert rel.max() < 2e-5


def test_linearity_and_idempotence_device_api(torch_mod):
    """size-independent properties: P(2A, B) == 2 P(A, B) exactly (power-of-two scale), re-running is idempotent."""
    torch = torch_mod
    S = gen.bernoulli_mask(1024, 2048, 0.97, 8)
    A, B = operands(S, 64)
    ro, ci = _dev(torch, S.row_off), _dev(torch, S.col_idx)
    R, ncl, _ = pkg.row_reorder_dev(ro, ci, S.M, S.N, 0.3, 16)
    lay, _, _ = pkg.layout_build_dev(ro, ci, S.M, S.N, R, 0.3)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    P1, _ = pkg.sddmm_gpu(dA, dB, lay)
    P1 = P1.clone()
    P2, _ = pkg.sddmm_gpu(dA * 2, dB, lay)
    torch.cuda.synchronize()
    assert torch.equal(P2[: S.nnz], 2 * P1[: S.nnz])
    P3, _ = pkg.sddmm_gpu(dA, dB, lay)
    torch.cuda.synchronize()
    assert torch.equal(P3[: S.nnz], P1[: S.nnz])


def test_full_size_config2_properties(torch_mod):
    """BASELINE config 2 at FULL size (uniform 100k x 100k, ~1e8 stored entries, K=128), too big for the oracle:
    size-independent properties instead.  (i) every P entry is written exactly once (NaN canary), (ii) exact
    linearity under a power-of-two scale, (iii) idempotence, (iv) 64 sampled rows against fp64 on the device,
    (v) the layout's entry counts add up to nnz and the statistics agree with them.  Row order = identity (the
    3 s clustering of this matrix is covered by bench.py; every cluster is a singleton)."""
    torch = torch_mod
    S = gen.uniform_random(100_000, 100_000, 0.01, 2)
    K = 128
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    dA = torch.rand((S.M, K), device="cuda", generator=g) * 2
    dB = torch.rand((S.N, K), device="cuda", generator=g) * 2
    ro, ci = _dev(torch, S.row_off), _dev(torch, S.col_idx)
    R = torch.arange(S.M, dtype=torch.int32, device="cuda")
    lay, _, _ = pkg.layout_build_dev(ro, ci, S.M, S.N, R, 0.3)
    info = lay.info
    assert int(info.numDenseValues) + int(info.numSparseValues) == S.nnz
    P1 = torch.full((S.nnz,), float("nan"), device="cuda")
    pkg.sddmm_gpu(dA, dB, lay, P1)
    torch.cuda.synchronize()
    assert not torch.isnan(P1).any()
    P2 = torch.zeros(S.nnz, device="cuda")
    pkg.sddmm_gpu(dA * 2, dB, lay, P2)
    P3 = torch.zeros(S.nnz, device="cuda")
    pkg.sddmm_gpu(dA, dB, lay, P3)
    torch.cuda.synchronize()
    assert torch.equal(P2, 2 * P1) and torch.equal(P3, P1)
    rows = np.random.default_rng(1).choice(S.M, 64, replace=False)
    for r in rows:
        b, e = int(S.row_off[r]), int(S.row_off[r + 1])
        ref = (dA[int(r)].double()[None, :] * dB[ci[b:e].to(torch.int64)].double()).sum(1)
        err = (P1[b:e].double() - ref).abs()
        assert bool(((err < 1e-5) | (err / ref.abs().clamp_min(1e-3) < 1e-3)).all())  # checkData.hpp:14-30
    ev = pkg.evaluationReordering(S, lay, 0.3)
    assert ev["numSparseData"] == int(info.numSparseValues) and ev["numDenseData"] == S.nnz - ev["numSparseData"]
    assert ev["originalNumDenseBlock"] == 0  # 1 % density: no 16x16 block reaches 77 entries


def test_sharded_layouts_cover_every_nonzero_once(torch_mod):
    """row-panel shards (multi-GPU layer): the union of the shards' outputs == the single layout's."""
    torch = torch_mod
    S = gen.rmat(12, 8, 4)
    A, B = operands(S, 32)
    ro, ci = _dev(torch, S.row_off), _dev(torch, S.col_idx)
    R, _, _ = pkg.row_reorder_dev(ro, ci, S.M, S.N, 0.3, 16)
    cuts = pkg.shard_plan(S, R.cpu().numpy().view(np.uint32), 3)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    full, _, _ = pkg.layout_build_dev(ro, ci, S.M, S.N, R, 0.3)
    Pfull, _ = pkg.sddmm_gpu(dA, dB, full)
    P = torch.full((S.nnz,), float("nan"), device="cuda")
    for s in range(3):
        lay, _, _ = pkg.layout_build_dev(ro, ci, S.M, S.N, R, 0.3, int(cuts[s]), int(cuts[s + 1]))
        pkg.sddmm_gpu(dA, dB, lay, P)
    torch.cuda.synchronize()
    assert not torch.isnan(P).any()
    assert O.check_data(Pfull[: S.nnz].cpu().numpy(), P.cpu().numpy()) == 0


def test_medium_uniform_roundtrip(torch_mod):
    """20000 x 20000 at 1% (config-2 class, scaled): permutation validity + values on a row sample."""
    torch = torch_mod
    S = gen.uniform_random(20000, 20000, 0.01, 2)
    K = 128
    A, B = operands(S, K)
    ro, ci = _dev(torch, S.row_off), _dev(torch, S.col_idx)
    R, ncl, ms = pkg.row_reorder_dev(ro, ci, S.M, S.N, 0.3, 0)
    Rh = R.cpu().numpy().view(np.uint32)
    lens = np.diff(S.row_off.astype(np.int64))
    assert np.array_equal(np.sort(Rh), np.nonzero(lens)[0])
    lay, _, _ = pkg.layout_build_dev(ro, ci, S.M, S.N, R, 0.3)
    assert lay.info.numDenseValues + lay.info.numSparseValues == S.nnz
    P, _ = pkg.sddmm_gpu(torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda(), lay)
    torch.cuda.synchronize()
    Ph = P[: S.nnz].cpu().numpy()
    rows = np.random.default_rng(0).choice(S.M, 64, replace=False)
    for r in rows:
        b, e = int(S.row_off[r]), int(S.row_off[r + 1])
        ref = (A[r][None, :] * B[S.col_idx[b:e]]).sum(1, dtype=np.float64)
        assert np.all(np.abs(Ph[b:e] - ref) / np.maximum(np.abs(ref), 1e-3) < 1e-3)


REF_DUMP = os.path.join(ROOT, "oracle", "_ref", "ref_dump")


@pytest.mark.skipif(not os.path.exists(REF_DUMP), reason="reference GPU binary not shipped")
@pytest.mark.parametrize("name", ["blocks512_a07_d03", "rmat12_a03_d03", "w3_600x4800_a05_d03"])
def test_against_live_reference_binary(name):
    """Same input through the UNMODIFIED reference pipeline (oracle/_ref/ref_dump) and through ours."""
    S, alpha, delta, _K = CASES[name]
    K = 32  # the reference's K>32 kernels fault on sm_100
    A, B = operands(S, K)
    tmp = tempfile.mkdtemp()
    case = os.path.join(tmp, "case.bin")
    gen.write_case_bin(case, S, A, B)
    p = subprocess.run([REF_DUMP, case, repr(alpha), repr(delta), tmp], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-500:]
    ref = {k: np.fromfile(os.path.join(tmp, k + ".u32"), dtype=np.uint32) for k in ("reorderedRows",) + LAYOUT_KEYS + RPHM_KEYS}
    Pref = np.fromfile(os.path.join(tmp, "P.f32"), dtype=np.float32)
    bs = [int(l.split()[1]) for l in open(os.path.join(tmp, "meta.txt")) if l.startswith("block_size")][0]
    res = pkg.sddmm(S, A, B, alpha=alpha, delta=delta, block_size=bs)
    for k in ref:
        assert np.array_equal(res[k], ref[k]), k
    assert O.check_data(Pref, res["P"]) == 0


def test_cli_binary_matches_reference_contract(tmp_path):
    """BSMR-sddmm -f file -k K -a alpha -d delta (src/main.cu:6-42, include/Options.hpp): runs the C++
    host mirror end to end, validates with the host checker and prints the reference's [key : value] log."""
    exe = os.path.join(ROOT, "sddmm-gpu_b200", "BSMR-sddmm")
    assert os.access(exe, os.X_OK), "CLI not built"
    S = gen.block_structured(512, 768, 6, 96, 0.8, seed=11, noise=0.004)
    mtx = str(tmp_path / "blocks.mtx")
    gen.write_mtx(mtx, S, order="col")
    p = subprocess.run([exe, "-f", mtx, "-k", "64", "-a", "0.3", "-d", "0.3", "-c", "1"], capture_output=True,
                       text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-400:]
    out = p.stdout
    assert "Pass! Result validates successfully." in out
    kv = dict(l.strip("[]").split(" : ", 1) for l in out.replace("], [", "]\n[").splitlines() if " : " in l and l.startswith("["))
    assert int(kv["NNZ"]) == S.nnz and int(kv["K"]) == 64
    rr = O.row_reorder(S, 0.3, 16)
    cr = O.col_reorder(S, rr["reorderedRows"], 0.3)
    assert int(kv["NumRowPanel"]) == cr["numRowPanels"]
    assert int(kv["bsmr_numDenseBlock"]) == int(cr["denseColOffsets"][-1]) // 16
    assert int(kv["bsmr_numSparseData"]) == int(cr["sparseValueOffsets"][-1])
    assert float(kv["bsmr_gflops"]) > 0
    # every key scripts/analyze_results.cpp parses is there, with evaluationReordering's values (src/sddmm.cu:32)
    ev = O.evaluation_reordering(S, rr["reorderedRows"], cr, 0.3)
    assert int(kv["original_numDenseBlock"]) == ev["originalNumDenseBlock"]
    assert float(kv["original_averageDensity"]) == pytest.approx(ev["originalAverageDensity"], abs=0.006)  # %.2f
    assert float(kv["bsmr_averageDensity"]) == pytest.approx(ev["averageDensity"], abs=0.006)
    assert int(kv["bsmr_numDenseData"]) == ev["numDenseData"]
    assert int(kv["bsmr_numDenseThreadBlocks"]) == ev["numDenseThreadBlocks"]
    assert int(kv["bsmr_numSparseThreadBlocks"]) == ev["numSparseThreadBlocks"]
    for key in ("blockDim_dense", "blockDim_sparse", "bsmr_threadBlockRatio", "bsmr_rowReordering", "bsmr_colReordering",
                "bsmr_reordering", "bsmr_numClusters", "bsmr_alpha", "bsmr_delta", "bsmr_sddmm"):
        assert key in kv, key
    assert "[bsmr_dataRatio: " in out
    # positional form: prog file K
    p = subprocess.run([exe, mtx, "32"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "[K : 32]" in p.stdout


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()


def test_pipelined_host_api_equals_synchronous(torch_mod):
    """sddmm_run_host_async (two slots) returns exactly what sddmm_run_host returns, for every batch."""
    torch = torch_mod
    S = gen.bernoulli_mask(2048, 1024, 0.95, 3)
    K = 64
    lay = pkg.BSMR(0.3, 0.3, S, block_size=16).layout()
    outs, refs = [], []
    bufs = [torch.zeros(S.nnz, dtype=torch.float32).pin_memory() for _ in range(2)]
    ins = []
    for i in range(5):
        A, B = gen.dense_operands(S.M, S.N, K, seed_a=100 + i, seed_b=200 + i)
        ins.append((torch.from_numpy(A).pin_memory(), torch.from_numpy(B).pin_memory()))
        refs.append(pkg.sddmm_gpu(A, B, lay)[0].copy())
    for i, (a, b) in enumerate(ins):
        if i >= 2:
            pkg.sddmm_gpu_sync(lay)  # slot reuse: previous result must have been consumed
            outs.append(bufs[i & 1].numpy().copy()) if False else None
        pkg.sddmm_gpu_async(a.numpy(), b.numpy(), lay, bufs[i & 1].numpy(), i & 1)
        pkg.sddmm_gpu_sync(lay)
        assert np.array_equal(bufs[i & 1].numpy(), refs[i]), i
    # back-to-back without intermediate syncs: last two results
    for i in (3, 4):
        pkg.sddmm_gpu_async(ins[i][0].numpy(), ins[i][1].numpy(), lay, bufs[i & 1].numpy(), i & 1)
    pkg.sddmm_gpu_sync(lay)
    assert np.array_equal(bufs[1].numpy(), refs[3]) and np.array_equal(bufs[0].numpy(), refs[4])


@pytest.mark.parametrize("plan_kind", ["bsmr", "tile"])
def test_host_passes_copy_only_referenced_rows(plan_kind, torch_mod, monkeypatch):
    """Host entry points with page-locked A / B: the referenced-rows gather (A rows of reorderedRows, B^T rows of the
    referenced columns, read through the mapped pointers) gives bit-identical results to whole-array copies and
    moves fewer bytes; pageable buffers fall back to whole-array copies.  Operands change between passes."""
    torch = torch_mod
    S = gen.with_empty_rows(gen.rmat(11, 8, 4), every=3)   # many empty rows AND columns
    K = 64
    b = pkg.BSMR().rowReordering(0.3, S, block_size=16)
    b.colReordering(0.3, S, tiles="always" if plan_kind == "tile" else "never")
    lay = b.layout()
    used_cols = int(np.unique(S.col_idx).size)
    used_rows = int((np.diff(S.row_off.astype(np.int64)) > 0).sum())
    assert used_cols + used_rows < 0.85 * (S.M + S.N)
    if plan_kind == "tile":
        monkeypatch.setenv("SDDMM_B200_PLAN", "full")
    out = torch.zeros(S.nnz, dtype=torch.float32).pin_memory()
    for it in range(3):
        A, B = gen.dense_operands(S.M, S.N, K, seed_a=300 + it, seed_b=400 + it)
        ref = O.sddmm_cpu(S, A, B)
        pa, pb = torch.from_numpy(A).pin_memory(), torch.from_numpy(B).pin_memory()
        results = {}
        for mode in ("full", "auto"):
            monkeypatch.setenv("SDDMM_B200_H2D", mode)
            P, _ = pkg.sddmm_gpu(pa.numpy(), pb.numpy(), lay)
            results[mode] = (P.copy(), pkg.host_traffic(lay))
            assert O.check_data(ref, P) == 0, (mode, it)
            pkg.sddmm_gpu_async(pa.numpy(), pb.numpy(), lay, out.numpy(), it & 1)
            pkg.sddmm_gpu_sync(lay)
            assert np.array_equal(out.numpy(), P), (mode, it, "async")
        assert np.array_equal(results["full"][0], results["auto"][0])
        assert results["full"][1] == (4 * K * (S.M + S.N), 4 * S.nnz)
        assert results["auto"][1] == (4 * K * (used_rows + used_cols), 4 * S.nnz)
        monkeypatch.setenv("SDDMM_B200_H2D", "auto")
        P2, _ = pkg.sddmm_gpu(A, B, lay)  # pageable numpy buffers: whole-array copies
        assert pkg.host_traffic(lay)[0] == 4 * K * (S.M + S.N)
        assert np.array_equal(P2, results["full"][0])


def test_batched_sddmm_matches_per_batch_calls(torch_mod):
    """sddmm_gpu_batch (src/sddmmKernel.cu:2764-2850): numBatch (A, B, P) triples back to back, one layout."""
    torch = torch_mod
    S = gen.block_structured(300, 400, 4, 64, 0.8, seed=2, noise=0.01)
    K, nb = 64, 3
    lay = pkg.BSMR(0.3, 0.3, S, block_size=16).layout()
    A = torch.rand((nb, S.M, K), device="cuda") * 2
    B = torch.rand((nb, S.N, K), device="cuda") * 2
    P = pkg.sddmm_gpu_batch(A, B, lay)
    torch.cuda.synchronize()
    for b in range(nb):
        Pb, _ = pkg.sddmm_gpu(A[b].contiguous(), B[b].contiguous(), lay)
        torch.cuda.synchronize()
        assert torch.equal(P[b, : S.nnz], Pb[: S.nnz])
        assert O.check_data(O.sddmm_cpu(S, A[b].cpu().numpy(), B[b].cpu().numpy()), P[b, : S.nnz].cpu().numpy()) == 0


def test_layout_cache_roundtrip(tmp_path, torch_mod):
    """bsmr_layout_save / bsmr_layout_load: identical arrays and identical SDDMM results."""
    S = gen.rmat(11, 8, 5)
    A, B = operands(S, 32)
    lay = pkg.BSMR(0.3, 0.3, S, block_size=16).layout()
    P1, _ = pkg.sddmm_gpu(A, B, lay)
    path = tmp_path / "layout.bsmr"
    lay.save(path)
    lay2 = pkg.Layout.load(path)
    a1, a2 = lay.arrays(), lay2.arrays()
    for k in a1:
        assert np.array_equal(a1[k], a2[k]), k
    assert bytes(lay.info) == bytes(lay2.info)
    P2, _ = pkg.sddmm_gpu(A, B, lay2)
    assert np.array_equal(P1, P2)
    open(tmp_path / "bad.bsmr", "wb").write(b"nonsense")
    with pytest.raises(pkg.SddmmError):
        pkg.Layout.load(tmp_path / "bad.bsmr")


def test_strong_scaling_shards_two_ranks_nccl(torch_mod):
    """ShardedSDDMM on 2 ranks (both on cuda:0 here; one GPU each under torchrun): NCCL is exercised by
    bench.py --gpus N; here the sharded plan is checked in-process by emulating the two ranks."""
    torch = torch_mod
    S = gen.rmat(12, 8, 4)
    A, B = operands(S, 64)
    ro, ci = _dev(torch, S.row_off), _dev(torch, S.col_idx)
    R, _, _ = pkg.row_reorder_dev(ro, ci, S.M, S.N, 0.3, 16)
    Rh = R.cpu().numpy().view(np.uint32)
    cuts = pkg.shard_plan(S, Rh, 2)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    P = torch.zeros(S.nnz, device="cuda")
    covered = 0
    for r in range(2):
        lay, _, _ = pkg.layout_build_dev(ro, ci, S.M, S.N, R, 0.3, int(cuts[r]), int(cuts[r + 1]))
        covered += lay.info.numDenseValues + lay.info.numSparseValues
        pkg.sddmm_gpu(dA, dB, lay, P)
    torch.cuda.synchronize()
    assert covered == S.nnz
    assert O.check_data(O.sddmm_cpu(S, A, B), P.cpu().numpy()) == 0


def test_error_paths_are_loud():
    """bad arguments -> error codes with a message, never a silent fallback"""
    S = gen.uniform_random(64, 64, 0.2, 1)
    lay = pkg.BSMR(0.3, 0.3, S, block_size=16).layout()
    A, B = operands(S, 6)  # K % 4 != 0
    with pytest.raises(pkg.SddmmError) as e:
        pkg.sddmm_gpu(A, B, lay)
    assert e.value.code == 1 and "multiple of 4" in str(e.value)


@pytest.mark.parametrize("shape", [(1, 40), (17, 5), (33, 2000), (300, 15)])
def test_tiny_and_ragged_shapes(shape):
    """single row, fewer than 16 columns, partial last panel: whole path vs oracle"""
    M, N = shape
    rng = np.random.default_rng(M * 1000 + N)
    dens = 0.4
    mask = rng.random((M, N)) < dens
    mask[0, 0] = True
    mask[-1, -1] = True
    r, c = np.nonzero(mask)
    from sddmm_gpu_b200.generators import _from_coo
    S = _from_coo(M, N, r, c, f"tiny{M}x{N}")
    A, B = operands(S, 32)
    bs = _bs(S)
    res = pkg.sddmm(S, A, B, alpha=0.3, delta=0.3, block_size=bs)
    rr = O.row_reorder(S, 0.3, bs)
    assert np.array_equal(res["reorderedRows"], rr["reorderedRows"])
    cr = O.col_reorder(S, rr["reorderedRows"], 0.3)
    rp = O.rphm_build(S, rr["reorderedRows"], cr)
    for k in LAYOUT_KEYS:
        assert np.array_equal(res[k], cr[k]), k
    for k in RPHM_KEYS:
        assert np.array_equal(res[k], rp[k]), k
    assert O.check_data(O.sddmm_cpu(S, A, B), res["P"]) == 0


def _golden_cases():
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    import make_ref_goldens as m
    have = {f[:-4] for f in os.listdir(os.path.join(GOLDEN, "ref_gpu")) if f.endswith(".npz")}
    # one entry per distinct (matrix, alpha): the *_k256 / *_k128 twins share the permutation of their *_k64 / *_k32 case
    return {c[0]: c for c in m.cases(big=True) if c[0] in have and not c[0].endswith(("_k256", "_k128"))}


_GOLDEN_CASES = _golden_cases()


@pytest.mark.parametrize("name", sorted(_GOLDEN_CASES))
def test_row_reorder_matches_reference_gpu_goldens(name):
    """our row reordering against the permutation the UNMODIFIED reference GPU pipeline produced on a B200
    (tests/golden/ref_gpu, tests/make_ref_goldens.py) -- directly, without the oracle in between; includes the
    4096 x 4096 masks, the nips surrogate and R-MAT scale 16 (65 536 rows, one-candidate-per-lane sweep)."""
    _, S, _K, alpha, _delta, _ = _GOLDEN_CASES[name]
    g = np.load(os.path.join(GOLDEN, "ref_gpu", name + ".npz"))
    b = pkg.BSMR().rowReordering(alpha, S, block_size=int(g["block_size"]))
    assert np.array_equal(b.reorderedRows(), g["reorderedRows"])
    assert b.numClusters() == int(g["num_clusters"])


def test_cli_test_mode_sweep_feeds_the_reference_analyser(tmp_path):
    """`BSMR-sddmm -f file -t 1 -l dir` (sddmm_testMode, src/sddmm.cu:62-118; the invocation of
    scripts/test_script.sh:92): the alpha x delta x K sweep writes one log per (K, alpha, delta) = 140 files, each
    holding one [key : value] record; the reference's own analyser (scripts/analyze_results.cpp, compiled unmodified
    into oracle/_ref/analyze_results) reads ALL of them and emits its per-K CSVs with the best BSMR GFLOP/s."""
    exe = os.path.join(ROOT, "sddmm-gpu_b200", "BSMR-sddmm")
    ana = os.path.join(ROOT, "oracle", "_ref", "analyze_results")
    assert os.access(exe, os.X_OK), "CLI not built"
    if not os.access(ana, os.X_OK):
        pytest.skip("oracle/_ref/analyze_results not shipped")
    S = gen.block_structured(256, 384, 4, 64, 0.8, seed=5, noise=0.01)
    mtx = str(tmp_path / "sweep.mtx")
    gen.write_mtx(mtx, S, order="col")
    logdir = tmp_path / "logs"
    logdir.mkdir()
    p = subprocess.run([exe, "-f", mtx, "-t", "1", "-l", str(logdir) + "/", "-b", "16"], capture_output=True, text=True,
                       timeout=900)
    assert p.returncode == 0, p.stderr[-400:]
    logs = sorted(os.listdir(logdir))
    assert len(logs) == 5 * 7 * 4, len(logs)  # alphas x deltas x Ks
    assert "BSMR_k_128_a_0.3_d_0.3.log" in logs and "BSMR_k_32_a_0.1_d_0.log" in logs and "BSMR_k_256_a_0.9_d_1.1.log" in logs
    # every record carries the sweep's own (alpha, delta, K), a positive kernel time, and delta's effect on the split
    seen = {}
    for name in logs:
        txt = open(logdir / name).read()
        assert txt.count("---New data---") == 1
        kv = dict(l.strip("[]").split(" : ", 1) for l in txt.replace("], [", "]\n[").splitlines() if " : " in l and l.startswith("["))
        k, a, d = int(kv["K"]), float(kv["bsmr_alpha"]), float(kv["bsmr_delta"])
        assert name == f"BSMR_k_{k}_a_{a:g}_d_{d:g}.log"
        assert float(kv["bsmr_sddmm"]) > 0 and float(kv["bsmr_gflops"]) > 0 and int(kv["NNZ"]) == S.nnz
        seen[(k, a, d)] = (int(kv["bsmr_numDenseData"]), int(kv["bsmr_numSparseData"]))
    for (k, a, d), (nd, ns) in seen.items():
        assert nd + ns == S.nnz
        if d > 1.0:
            assert nd == 0
        if d == 0.0:
            assert ns == 0
    # the layout of one sweep point against the oracle (same alpha/delta, block size 16)
    rr = O.row_reorder(S, 0.5, 16)
    cr = O.col_reorder(S, rr["reorderedRows"], 0.3)
    assert seen[(64, 0.5, 0.3)][1] == int(cr["sparseValueOffsets"][-1])
    # the reference's analyser reads the 140 files: one invocation per K (it writes results_<K>.csv for the K of the
    # records it was given, scripts/analyze_results.cpp:785-789), 35 (alpha, delta) logs each
    for k in (32, 64, 128, 256):
        mine = [str(logdir / n) for n in logs if n.startswith(f"BSMR_k_{k}_")]
        assert len(mine) == 35
        r = subprocess.run([ana] + mine, capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
        assert r.returncode == 0, r.stderr[-400:]
        csv = [l for l in open(logdir / f"results_{k}.csv").read().splitlines() if l]
        assert csv[0].startswith("file,M,N,NNZ,Sparsity,K,BSMR")
        row = csv[1].split(",")
        assert row[0] == mtx and int(row[3]) == S.nnz and int(row[5]) == k and float(row[6]) > 0


def test_layout_cache_rejects_tampered_files(tmp_path):
    """bsmr_layout_load cross-checks the file before anything reaches the device: truncated, padded, or edited
    files (an offset array that no longer matches, a CSR index past nnz, a work item past its panel) end in
    SDDMM_E_ARG, never in out-of-bounds reads inside the kernels."""
    S = gen.block_structured(300, 400, 4, 64, 0.8, seed=2, noise=0.01)
    lay = pkg.BSMR(0.3, 0.3, S, block_size=16).layout()
    good = tmp_path / "good.bsmr"
    lay.save(good)
    raw = bytearray(open(good, "rb").read())
    pkg.Layout.load(good)  # sanity: the untouched file loads

    def expect_reject(data, what):
        p = tmp_path / (what + ".bsmr")
        open(p, "wb").write(bytes(data))
        with pytest.raises(pkg.SddmmError) as e:
            pkg.Layout.load(p)
        assert e.value.code == 1, what

    expect_reject(raw[: len(raw) // 2], "truncated")
    hdr = 8 + 4 + 13 * 4 + 3 * 4  # magic, version, bsmr_layout_info, sparseChunk, numDenseWork, numSparseWork
    n0 = int.from_bytes(raw[hdr: hdr + 8], "little")  # length of reorderedRows
    assert n0 == lay.info.numRows
    bad = bytearray(raw)
    bad[hdr + 8: hdr + 12] = (S.M + 5).to_bytes(4, "little")  # a row id past M
    expect_reject(bad, "row_out_of_range")
    bad = bytearray(raw)
    bad[8 + 4 + 2 * 4: 8 + 4 + 3 * 4] = (7).to_bytes(4, "little")  # info.nnz = 7: every CSR index is now past nnz
    expect_reject(bad, "nnz_too_small")
    bad = bytearray(raw)
    bad[8 + 4 + 3 * 4: 8 + 4 + 4 * 4] = (lay.info.numRows + 16).to_bytes(4, "little")  # info.numRows disagrees with the array
    expect_reject(bad, "row_count")
    # an offset array edited so that it no longer ends at the array it indexes
    off = hdr + 8 + 4 * n0          # -> denseCols (length + data)
    n1 = int.from_bytes(raw[off: off + 8], "little")
    off2 = off + 8 + 4 * n1         # -> denseColOffsets
    n2 = int.from_bytes(raw[off2: off2 + 8], "little")
    bad = bytearray(raw)
    last = off2 + 8 + 4 * (n2 - 1)
    bad[last: last + 4] = (n1 + 16).to_bytes(4, "little")
    expect_reject(bad, "offsets")


# ---- BASELINE configs at FULL size (the oracle cannot run these: size-independent properties instead) --------------
def _full_size_properties(torch, ro, ci, ro_host, M, N, nnz, K, R, label, plan=None):
    """(i) every P entry written exactly once (NaN canary), (ii) exact linearity under a power-of-two scale,
    (iii) idempotence, (iv) sampled rows against fp64 on the device with the reference's checkData rule,
    (v) the layout's entry counts add up to nnz."""
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    dA = torch.rand((M, K), device="cuda", generator=g) * 2
    dB = torch.rand((N, K), device="cuda", generator=g) * 2
    lay, _, _ = pkg.layout_build_dev(ro, ci, M, N, R, 0.3)
    info = lay.info
    assert int(info.numDenseValues) + int(info.numSparseValues) == nnz, label
    P1 = torch.full((nnz,), float("nan"), device="cuda")
    pkg.sddmm_gpu(dA, dB, lay, P1, plan=plan)
    torch.cuda.synchronize()
    assert not torch.isnan(P1).any(), label
    P2 = torch.zeros(nnz, device="cuda")
    pkg.sddmm_gpu(dA * 2, dB, lay, P2, plan=plan)
    torch.cuda.synchronize()
    assert torch.equal(P2, 2 * P1), label
    P2.zero_()
    pkg.sddmm_gpu(dA, dB, lay, P2, plan=plan)
    torch.cuda.synchronize()
    assert torch.equal(P2, P1), label
    rows = np.random.default_rng(1).choice(M, 48, replace=False)
    for r in rows:
        b, e = int(ro_host[r]), int(ro_host[r + 1])
        if e <= b:
            continue
        ref = (dA[int(r)].double()[None, :] * dB[ci[b:e].to(torch.int64)].double()).sum(1)
        err = (P1[b:e].double() - ref).abs()
        assert bool(((err < 1e-5) | (err / ref.abs().clamp_min(1e-3) < 1e-3)).all()), (label, int(r))  # checkData.hpp:14-30
    return info


def test_full_size_config4_rmat22_properties(torch_mod):
    """BASELINE config 4 at full size: R-MAT scale 22 (4.19 M rows, ~65 M stored entries), K=128, generated on the
    device; identity row order over the non-empty rows (the 37 s clustering is exercised by bench.py)."""
    torch = torch_mod
    ro, ci, M = gen.rmat_device(22, 16, 4)
    ro_host = ro.cpu().numpy().view(np.uint32)
    R = torch.nonzero((ro[1:] - ro[:-1]) != 0).flatten().to(torch.int32)
    info = _full_size_properties(torch, ro, ci, ro_host, M, M, int(ci.numel()), 128, R, "config 4")
    assert info.numRowPanels == (R.numel() + 15) // 16


@pytest.mark.parametrize("sparsity,kind", [(0.70, "bern"), (0.80, "dlmc"), (0.90, "dlmc"), (0.95, "bern"), (0.98, "dlmc")])
@pytest.mark.parametrize("K", [64, 256])
def test_full_size_config3_masks_values_vs_oracle(sparsity, kind, K):
    """BASELINE config 3 at full size: 4096 x 4096 masks at 70-98 % sparsity, K = 64 / 256, whole pipeline on host
    buffers (reorder -> layout -> SDDMM with whatever plan the cost model picks) against the oracle's sddmm_cpu +
    checkData; the permutation against the oracle's pruned clustering."""
    S = gen.bernoulli_mask(4096, 4096, sparsity, 30) if kind == "bern" else gen.dlmc_magnitude_mask(4096, 4096, sparsity, 33)
    A, B = operands(S, K)
    res = pkg.sddmm(S, A, B, alpha=0.3, delta=0.3, block_size=16)
    assert O.check_data(O.sddmm_cpu(S, A, B), res["P"]) == 0
    assert res["numDenseValues"] + res["numSparseValues"] == S.nnz
    lens = np.diff(S.row_off.astype(np.int64))
    assert np.array_equal(np.sort(res["reorderedRows"]), np.nonzero(lens)[0])


def test_full_size_config5_rmat25_properties(torch_mod):
    """BASELINE config 5 at full size: R-MAT scale 25 (33.5 M rows, ~529 M stored entries) generated on the device;
    K=64 keeps the operands at 8.6 GB each (the K=256 run is bench.py's); identity row order.  Exercises 64-bit
    addressing (M*K > 2^31 elements) and the 2^29-entry layouts."""
    torch = torch_mod
    free, _ = torch.cuda.mem_get_info()
    if free < 90e9:
        pytest.skip("needs ~90 GB of free device memory")
    ro, ci, M = gen.rmat_device(25, 16, 5)
    ro_host = ro.cpu().numpy().view(np.uint32)
    R = torch.nonzero((ro[1:] - ro[:-1]) != 0).flatten().to(torch.int32)
    _full_size_properties(torch, ro, ci, ro_host, M, M, int(ci.numel()), 64, R, "config 5")
    del ro, ci, R
    torch.cuda.empty_cache()
