"""GPU parity matrix (-m gpu): EVERY kernel that can serve an SDDMM pass, selected by name through the per-call
plan (sddmm_run_dev_ex / sddmm_plan), at every K class and batch count, against the oracle's sddmm_cpu with the
reference's checkData tolerance (include/checkData.hpp:14-30).  One test id per (kernel, K, batch, matrix).

  dense_reg    k_sddmm_dense        BSMR dense blocks, register-staged operands   (src/sddmmKernel.cu:213-351, :355-488)
  dense_tma    k_sddmm_dense_tma    BSMR dense blocks, TMA tile::gather4 operands (same reference kernels)
  res_panel    k_sddmm_residual     residual, one panel per CTA, any K % 4 == 0   (src/sddmmKernel.cu:1994-2104, :2109-2199)
  res_sp       k_sddmm_residual_sp  residual, super-panels, K in {32,...,512}     (same reference kernels)
  res_stream   k_sddmm_residual_stream  residual, row order, deep gather pipeline, K in {32,...,512}
  tile_reg     k_sddmm_tile         128x128 tcgen05 tiles, register-staged
  tile_tma     k_sddmm_tile_tma     128x128 tcgen05 tiles, TMA-fed
  tile_tma4    k_sddmm_tile_tma4    2x2 clusters, multicast TMA
  tile_pair    k_sddmm_tile_pair    persistent CTA pairs, 256x256 tcgen05.mma.cta_group::2, double-buffered TMEM
  res_sp_fp16 / tile_tma_fp16       the opt-in fp16-operand forms of K7b and K9 (fp32 accumulation, same tolerance)
and the clustering kernels (k_cluster, k_cluster_batched<1|2|4|8>, lane sweep on/off, signature filter on/off)
against the permutations of the unmodified reference GPU pipeline (tests/golden/ref_gpu).
"""
import os

import numpy as np
import pytest

from cases import GOLDEN, gen, operands, pkg
from oracle import oracle as O

pytestmark = pytest.mark.gpu

KS = [32, 64, 96, 128, 256, 512]
SP_KS = {32, 64, 128, 256, 512}

# kernel name -> (delta that routes EVERY entry to it, plan keywords, what plan_resolve must report)
KERNELS = {
    "dense_reg": (0.0, dict(plan="bsmr", dense="reg"), dict(plan="bsmr", dense="reg")),
    "dense_tma": (0.0, dict(plan="bsmr", dense="tma"), dict(plan="bsmr", dense="tma")),
    "res_panel": (1.1, dict(plan="bsmr", residual="panel"), dict(plan="bsmr", residual="panel")),
    "res_sp": (1.1, dict(plan="bsmr", residual="superpanel"), dict(plan="bsmr", residual="superpanel")),
    "res_stream": (1.1, dict(plan="bsmr", residual="stream"), dict(plan="bsmr", residual="stream")),
    # opt-in fp16 operand copies (sddmm_plan.operands): fp16 A tile in the super-panel kernel, kind::f16 tile kernel
    "res_sp_fp16": (1.1, dict(plan="bsmr", residual="superpanel", operands="fp16"),
                    dict(plan="bsmr", residual="superpanel", operands="fp16")),
    "tile_tma_fp16": (0.3, dict(plan="tile", tile="tma", operands="fp16"), dict(plan="tile", tile="tma", operands="fp16")),
    "tile_reg": (0.3, dict(plan="tile", tile="reg"), dict(plan="tile", tile="reg")),
    "tile_tma": (0.3, dict(plan="tile", tile="tma"), dict(plan="tile", tile="tma")),
    "tile_tma3": (0.3, dict(plan="tile", tile="tma", tile_stages=3), dict(plan="tile", tile="tma", tile_stages=3)),
    "tile_tma4": (0.3, dict(plan="tile", tile="tma_cluster"), dict(plan="tile", tile="tma_cluster")),
    "tile_pair": (0.3, dict(plan="tile", tile="tma_pair"), dict(plan="tile", tile="tma_pair")),
    "tile_pair_fp16": (0.3, dict(plan="tile", tile="tma_pair", operands="fp16"),
                       dict(plan="tile", tile="tma_pair", operands="fp16")),
}


def _matrices():
    return {
        "blocks": gen.block_structured(200, 300, 4, 64, 0.8, seed=2, noise=0.01),
        "rmat11": gen.rmat(11, 8, 4),
    }


MATS = _matrices()
_LAYOUTS = {}


def _layout(mname, delta):
    """one layout per (matrix, delta), with the full-tile layout always built so every plan can be forced"""
    key = (mname, delta)
    if key not in _LAYOUTS:
        S = MATS[mname]
        b = pkg.BSMR().rowReordering(0.3, S, block_size=16)
        b.colReordering(delta, S, tiles="always")
        _LAYOUTS[key] = b.layout()
    return _LAYOUTS[key]


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    assert torch.cuda.is_available()
    return torch


def _run(torch, S, lay, K, nb, plan):
    rng = np.random.default_rng(K * 7 + nb)
    A = (rng.random((nb, S.M, K), dtype=np.float32) * np.float32(2)).astype(np.float32)
    B = (rng.random((nb, S.N, K), dtype=np.float32) * np.float32(2)).astype(np.float32)
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    P = torch.full((nb, max(1, S.nnz)), float("nan"), device="cuda")  # NaN canary: every entry must be written
    if nb == 1:
        pkg.sddmm_gpu(dA[0], dB[0], lay, P[0], plan=plan)
    else:
        pkg.sddmm_gpu_batch(dA, dB, lay, P, plan=plan)
    torch.cuda.synchronize()
    Ph = P.cpu().numpy()
    for b in range(nb):
        assert not np.isnan(Ph[b, : S.nnz]).any(), f"batch {b}: entries never written"
        assert O.check_data(O.sddmm_cpu(S, A[b], B[b]), Ph[b, : S.nnz]) == 0, f"batch {b}"


@pytest.mark.parametrize("mname", sorted(MATS))
@pytest.mark.parametrize("nb", [1, 3])
@pytest.mark.parametrize("K", KS)
@pytest.mark.parametrize("kernel", sorted(KERNELS))
def test_kernel_parity(kernel, K, nb, mname, torch_mod):
    delta, kw, expect = KERNELS[kernel]
    S = MATS[mname]
    lay = _layout(mname, delta)
    plan = pkg.make_plan(**kw)
    if kernel in ("res_sp", "res_stream", "res_sp_fp16") and K not in SP_KS:
        with pytest.raises(pkg.SddmmError) as e:  # an impossible choice fails loudly, it never silently runs another kernel
            pkg.plan_resolve(lay, K, nb, plan)
        assert e.value.code == 4
        return
    got = pkg.plan_resolve(lay, K, nb, plan)
    for k, v in expect.items():
        assert got[k] == v, (got, expect)
    if delta == 0.0:
        assert lay.info.numSparseValues == 0 and lay.info.numDenseBlocks > 0
    if delta > 1.0:
        assert lay.info.numDenseBlocks == 0 and lay.info.numSparseValues == S.nnz
    _run(torch_mod, S, lay, K, nb, plan)


@pytest.mark.parametrize("pairs", [1, 3, 7])
@pytest.mark.parametrize("K", [32, 128, 200])
@pytest.mark.parametrize("operands_", ["exact", "fp16"])
def test_tile_pair_persistent_loop(pairs, K, operands_, torch_mod, monkeypatch):
    """K10 with the grid capped at a few CTA pairs: every pair walks MANY 256x256 quads, so the operand ring wraps,
    both TMEM accumulators are reused (full / empty phases flip) and quads with missing tiles are crossed."""
    monkeypatch.setenv("SDDMM_B200_PAIR_GRID", str(pairs))
    S = MATS["rmat11"]
    lay = _layout("rmat11", 0.3)
    assert lay.info.numRows > 256
    _run(torch_mod, S, lay, K, 2, pkg.make_plan(plan="tile", tile="tma_pair", operands=operands_))


@pytest.mark.parametrize("hints", ["0", "1"])
@pytest.mark.parametrize("gather", ["0", "1"])
@pytest.mark.parametrize("operands_", ["exact", "fp16"])
@pytest.mark.parametrize("K", [32, 64, 128, 256, 512])
def test_superpanel_kernel_modes(K, operands_, gather, hints, torch_mod, monkeypatch):
    """K7b under every (gather mode, L2 eviction policy with hub / tail column classes) combination, with fp32
    operands and with the fp16 copies (fp16 A tile; from K = 64 also fp16 B^T rows multiplied by FHFMA)."""
    monkeypatch.setenv("SDDMM_B200_SP_GATHER", gather)
    monkeypatch.setenv("SDDMM_B200_L2_HINTS", hints)
    monkeypatch.setenv("SDDMM_B200_L2_HUB_MB", "1")  # 1 MB of hub rows: rmat11 then has hub AND tail columns
    S = MATS["rmat11"]
    lay = _layout("rmat11", 1.1)
    _run(torch_mod, S, lay, K, 2, pkg.make_plan(plan="bsmr", residual="superpanel", operands=operands_))


@pytest.mark.parametrize("K", [36, 100, 520])
def test_residual_panel_kernel_odd_K(K, torch_mod):
    """K7 is the only residual kernel for K outside {32,...,512}: AUTO must pick it and it must be right."""
    S = MATS["rmat11"]
    lay = _layout("rmat11", 1.1)
    assert pkg.plan_resolve(lay, K, 1, pkg.make_plan(plan="bsmr"))["residual"] == "panel"
    _run(torch_mod, S, lay, K, 1, pkg.make_plan(plan="bsmr"))


@pytest.mark.parametrize("dense", ["reg", "tma"])
@pytest.mark.parametrize("residual", ["panel", "superpanel", "stream"])
@pytest.mark.parametrize("K", [64, 256])
def test_mixed_dense_and_residual(dense, residual, K, torch_mod):
    """delta = 0.3: dense blocks and residual entries in the same pass, on two streams."""
    S = MATS["blocks"]
    lay = _layout("blocks", 0.3)
    assert lay.info.numDenseBlocks > 0 and lay.info.numSparseValues > 0
    _run(torch_mod, S, lay, K, 2, pkg.make_plan(plan="bsmr", dense=dense, residual=residual))


def test_prepare_then_graph_capture(torch_mod):
    """after sddmm_prepare a pass only enqueues kernels: it can be captured in a CUDA graph and replayed."""
    torch = torch_mod
    S = MATS["blocks"]
    K = 128
    for kw in (dict(plan="tile", tile="tma"), dict(plan="tile", tile="tma_pair"),
               dict(plan="bsmr", dense="tma", residual="superpanel")):
        lay = _layout("blocks", 0.3)
        plan = pkg.make_plan(**kw)
        pkg.sddmm_prepare(lay, K, 1, plan)
        A, B = operands(S, K)
        dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
        P = torch.zeros(S.nnz, device="cuda")
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            pkg.sddmm_gpu(dA, dB, lay, P, plan=plan)  # warm-up on the capture stream
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                pkg.sddmm_gpu(dA, dB, lay, P, plan=plan)
        P.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert O.check_data(O.sddmm_cpu(S, A, B), P.cpu().numpy()) == 0


@pytest.mark.parametrize("tile", ["tma", "tma_pair"])
def test_two_streams_share_one_layout(tile, torch_mod):
    """tile-TMA plans, one layout, two streams: the shared rounded workspace is serialised by the layout's event."""
    torch = torch_mod
    S = MATS["blocks"]
    K = 128
    lay = _layout("blocks", 0.3)
    plan = pkg.make_plan(plan="tile", tile=tile)
    pkg.sddmm_prepare(lay, K, 1, plan)
    ops = [gen.dense_operands(S.M, S.N, K, seed_a=50 + i, seed_b=60 + i) for i in range(4)]
    dev = [(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()) for a, b in ops]
    outs = [torch.zeros(S.nnz, device="cuda") for _ in ops]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    for rep in range(5):
        for i, (a, b) in enumerate(dev):
            with torch.cuda.stream(streams[i & 1]):
                pkg.sddmm_gpu(a, b, lay, outs[i], plan=plan)
    torch.cuda.synchronize()
    for (a, b), out in zip(ops, outs):
        assert O.check_data(O.sddmm_cpu(S, a, b), out.cpu().numpy()) == 0


# ---- clustering kernels against the reference GPU pipeline's permutations -------------------------------------
def _golden(name):
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    import make_ref_goldens as m
    case = {c[0]: c for c in m.cases(big=True)}[name]
    return case, np.load(os.path.join(GOLDEN, "ref_gpu", name + ".npz"))


CLUSTER_VARIANTS = {
    "legacy": dict(kernel="legacy"),
    "batched1": dict(kernel="batched", batch=1),
    "batched2": dict(kernel="batched", batch=2),
    "batched4": dict(kernel="batched", batch=4),
    "batched8": dict(kernel="batched", batch=8),
    "lane_on": dict(kernel="batched", lane_rows="on"),
    "lane_off": dict(kernel="batched", lane_rows="off"),
    "sig_on": dict(kernel="batched", signature="on"),
    "sig_off": dict(kernel="batched", signature="off"),
    "sig_on_lane_on": dict(kernel="batched", signature="on", lane_rows="on"),
}


@pytest.mark.parametrize("name", ["blocks512_a03_d03", "blocks512_a07_d03", "uni512_a09_d01", "rmat12_a03_d03",
                                  "rmat12_a01_d01", "w3_600x4800_a05_d03", "w5_zipf400x9600_a04_d03",
                                  "nips_surrogate_a03_d03", "dlmc1024s70_a03_d03", "blocks512empty_a03_d03"])
@pytest.mark.parametrize("variant", sorted(CLUSTER_VARIANTS))
def test_cluster_kernel_variants_match_reference_gpu(variant, name):
    (_, S, _K, alpha, _delta, _), g = _golden(name)
    opts = pkg.make_reorder_opts(**CLUSTER_VARIANTS[variant])
    b = pkg.BSMR().rowReordering(alpha, S, block_size=int(g["block_size"]), opts=opts)
    assert np.array_equal(b.reorderedRows(), g["reorderedRows"])
    assert b.numClusters() == int(g["num_clusters"])
