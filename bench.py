#!/usr/bin/env python
"""bench.py -- effective SDDMM GFLOP/s (2*nnz*K/t), the reference's headline metric
(include/Logger.hpp:178-180), on BASELINE.json's configs, on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--scale S] [--K K]

Headline workload (every N, including 1): ONE R-MAT power-law matrix (BASELINE configs 4/5 class; scale 22 by
default, generated on the device with the same seed on every rank), the north star's pipeline end to end:
    row reorder ONCE on rank 0 (bit-exact to the reference's permutation)  ->  sddmm_mgpu_shard (row order
    broadcast over NCCL, nnz-balanced contiguous row-panel ranges, one layout per rank)  ->  B replicated ONCE
    (sddmm_mgpu_bcast, communicator warmed up before the timer)  ->  steady state with NO collective.
* A "step" is one SDDMM pass (dense-block tcgen05 kernel || residual CUDA-core kernel) of every rank over its own
  panels with the layout resident: what the reference's sddmmTime_ loop times (src/sddmmKernel.cu:2561-2659).
  Reordering / layout-build times are reported beside it in `config` (as the reference's log does).
* `value`  : 2 * nnz_total * K / max-over-ranks step time (CUDA events on the launching stream, barrier +
             synchronize on both sides); inputs resident in HBM.  scaling = strong (one matrix, fixed work).
* `e2e`    : the same metric through the host-buffer entry point sddmm_run_host_async (pinned host A, B in, host
             P out; H2D + D2H of every step inside the timed region), every rank on its own shard.
* `roofline`: rank 0's dominant kernel, ALGORITHMIC bytes (DESIGN.md "bytes per unit") / its mean launch time.
* `cpu_baseline` (N=1): the reference's OpenMP sddmm_cpu (oracle/_ref/libref_cpu.so when present, else the oracle
             port) on a bounded row sample of the same S / A / B, on this box's host cores.
* `config.k_sweep`: the metric at K = 32 ... 512 on the same layouts; `config.also` (N=1): the other BASELINE
             configs as nested records (config 2 uniform 100k^2, config 3 masks 4096^2, a block-structured
             matrix where the BSMR dense/sparse split is what wins, config 5 R-MAT scale 25).
--impl reference times the reference's CPU implementation of the path as the reference arm (rank 0 only).
--workload uniform100k | dlmc4096_s70 | ... runs one of the single-GPU records as the headline instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "effective SDDMM GFLOP/s (2*nnz*K/t)"
UNIT = "GFLOP/s"
FREE_MEM_FOR_BLOCK_SIZE = 180 * 10 ** 9  # fixed free-memory figure: reproducible block size

WORKLOADS = {
    # BASELINE configs[3] / [4]: one R-MAT matrix, row panels sharded over the ranks
    "rmat": dict(kind="rmat", scale=22, ef=16, K=256, seed=4,
                 desc="synthetic R-MAT power-law graph, edge factor 16, (a,b,c,d)=(.57,.19,.19,.05), one matrix "
                      "row-panel sharded over the ranks"),
    # BASELINE configs[1]
    "uniform100k": dict(kind="uniform", M=100_000, N=100_000, density=0.01, K=128, seed=2,
                        desc="synthetic uniform-random 100k x 100k, 1% density, K=128"),
    # configs[2]: DLMC-style pruned masks
    "dlmc4096_s70": dict(kind="bernoulli", M=4096, N=4096, sparsity=0.70, K=256, seed=30,
                         desc="synthetic pruned mask 4096 x 4096 at 70% sparsity, K=256"),
    "dlmc4096_s90": dict(kind="dlmc", M=4096, N=4096, sparsity=0.90, K=64, seed=33,
                         desc="synthetic DLMC-style magnitude-pruned mask 4096 x 4096 at 90% sparsity, K=64"),
    "dlmc4096_s98": dict(kind="dlmc", M=4096, N=4096, sparsity=0.98, K=64, seed=34,
                         desc="synthetic DLMC-style magnitude-pruned mask 4096 x 4096 at 98% sparsity, K=64"),
    # block structure hidden by interleaved rows / scattered columns: what BSMR's reordering + split is for
    "blockscat16k": dict(kind="blockscat", M=16384, N=16384, groups=64, cols=512, fill=0.6, noise=0.0005, K=128,
                         seed=41, desc="synthetic block-structured 16384^2 (64 row groups x 512 scattered columns, "
                                       "60% fill, 0.05% noise), K=128"),
    "small": dict(kind="uniform", M=4096, N=4096, density=0.02, K=64, seed=2, desc="smoke-size uniform"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=float(d["hbm_gbs"]), tf=float(d.get("bf16_tflops", 0)), which="measured")
    return dict(hbm_gbs=6650.0, tf=1590.0, which="fallback")


def make_pattern(gen, w, args):
    """host-generated patterns (everything except R-MAT, which is generated on the device)"""
    seed = w["seed"]
    if w["kind"] == "uniform":
        M = args.rows or w["M"]
        return gen.uniform_random(M, w["N"] if not args.rows else min(w["N"], max(M, 1024)), w["density"], seed)
    if w["kind"] == "bernoulli":
        return gen.bernoulli_mask(w["M"], w["N"], w["sparsity"], seed)
    if w["kind"] == "dlmc":
        return gen.dlmc_magnitude_mask(w["M"], w["N"], w["sparsity"], seed)
    if w["kind"] == "blockscat":
        return gen.block_structured_scattered(w["M"], w["N"], w["groups"], w["cols"], w["fill"], seed, noise=w["noise"])
    if w["kind"] == "rmat":
        return gen.rmat(args.scale or w["scale"], w["ef"], seed)
    raise ValueError(w["kind"])


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                if out.strip():
                    self.rows.append([x.strip() for x in out.strip().splitlines()[0].split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for i, n in enumerate(names):
                if len(r) > 3 + i and r[3 + i].lower().startswith("active"):
                    reasons.add(n)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(self.rows))


def algorithmic_bytes(rows_touched, cols_touched, M, K, nnz_kernel, n_blocks, kernel):
    """SURVEY.md 8(d) compulsory-traffic model: every operand element touched once.
    residual: A + B rows touched + (col, relRow, csrIdx: 12 B) + P (4 B) per entry + row offsets;
    dense   : A + B rows touched + per block 1024 B blockValues + 64 B denseCols + 4 B per stored entry of P."""
    base = 4.0 * K * (rows_touched + cols_touched) + 4.0 * (M + 1)
    if kernel == "residual":
        return base + 16.0 * nnz_kernel
    return base + 1088.0 * n_blocks + 4.0 * nnz_kernel


def traffic_entry(key):
    """dram bytes per launch from a committed `ncu --set full` capture (never measured inside a bench run: a run
    under the profiler is not a bench run) -- returned with its source, or (None, None)."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        d = json.load(open(tp)).get(key)
        if isinstance(d, dict):
            return d.get("bytes"), d.get("source")
        if d is not None:
            return d, "profiles/traffic.json"
    return None, None


# ------------------------------------------------------------------------------------------------
def omp_threads_for_cpu_arm():
    """torchrun exports OMP_NUM_THREADS=1 to its ranks; the CPU arm must use (and report) the box's cores."""
    from oracle import oracle as O
    want = os.cpu_count() or 1
    if O.ref_available():
        return int(O.ref().ref_omp_threads(want))
    return want


def cpu_reference_gflops(row_off, col_idx, M, N, A, B, K, sample_rows, repeats=3):
    """The reference's OpenMP sddmm_cpu on the first `sample_rows` rows (same S/A/B).  -> dict"""
    from oracle import oracle as O
    rows = int(min(sample_rows, M))
    nnz = int(row_off[rows])
    ro = np.ascontiguousarray(row_off[: rows + 1])
    ci = np.ascontiguousarray(col_idx[:nnz])
    As = np.ascontiguousarray(A[:rows])
    cores = omp_threads_for_cpu_arm()
    best = float("inf")
    if O.ref_available():
        kind = "reference"
        ctx = O.ref().ref_sddmm_prepare(As, B, ro, ci, rows, N, K, nnz)
        for _ in range(repeats):
            t0 = time.perf_counter()
            O.ref().ref_sddmm_run(ctx, None)
            best = min(best, time.perf_counter() - t0)
        O.ref().ref_sddmm_release(ctx)
    else:
        kind = "port"
        P = np.zeros(max(1, nnz), np.float32)
        for _ in range(repeats):
            t0 = time.perf_counter()
            O.lib().oracle_sddmm_cpu(As, B, ro, ci, rows, K, P, cores)
            best = min(best, time.perf_counter() - t0)
    return dict(value=2.0 * nnz * K / best / 1e9, unit=UNIT, cores=cores, kind=kind, seconds=best,
                sample=f"first {rows} rows ({nnz} nnz) of the same S with the same A/B, best of {repeats}, "
                       f"OpenMP threads = {cores} (omp_get_max_threads)")


def base_config(desc, M, N, nnz, K, alpha, delta):
    """the keys BOTH arms print, in this order"""
    return dict(workload=desc, M=int(M), N=int(N), nnz=int(nnz), K=int(K), alpha=alpha, delta=delta)


def host_pattern_for(args, w):
    """(row_off u32, col_idx u32, M, N, A, B) on the host, for the CPU arm: the SAME matrix and operands as the GPU
    arm (R-MAT and its operands come from the device generators when a GPU is there, as in run_ours)."""
    from __graft_entry__ import load_package
    gen = load_package().generators
    K = args.K or w["K"]
    if w["kind"] == "rmat":
        try:
            import torch
            have_gpu = torch.cuda.is_available()
        except Exception:
            have_gpu = False
        if have_gpu:
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
            ro, ci, M = gen.rmat_device(args.scale or w["scale"], w["ef"], w["seed"])
            g = torch.Generator(device="cuda")
            g.manual_seed(1001)
            A = (torch.rand((M, K), device="cuda", generator=g) * 2).cpu().numpy()
            B = (torch.rand((M, K), device="cuda", generator=g) * 2).cpu().numpy()
            return ro.cpu().numpy().view(np.uint32), ci.cpu().numpy().view(np.uint32), M, M, A, B
    S = make_pattern(gen, w, args)
    A, B = gen.dense_operands(S.M, S.N, K)
    return S.row_off, S.col_idx, S.M, S.N, A, B


def run_reference(args, w):
    """Reference arm: the reference's own CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from __graft_entry__ import load_package
    gen = load_package().generators
    K = args.K or w["K"]
    ro, ci, M, N, A, B = host_pattern_for(args, w)
    nnz = int(ro[-1])
    # bounded sample: a few seconds of CPU work per step (gather-bound graphs run at ~1.5 GFLOP/s on 16 cores,
    # L2-friendly uniform matrices at ~8)
    sample_rows = min(M, max(256, int(M * min(1.0, (4e9 if w["kind"] == "rmat" else 2e10) / max(1.0, 2.0 * nnz * K)))))
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_reference_gflops(ro, ci, M, N, A, B, K, sample_rows, repeats=1)
        if i >= args.warmup:
            vals.append(r)
    v = float(np.mean([r["value"] for r in vals]))
    ms = float(np.mean([r["seconds"] for r in vals])) * 1e3
    scale = args.scale or w.get("scale")
    desc = w["desc"] + (f", scale {scale}" if w["kind"] == "rmat" else "")
    line = dict(metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms, higher_is_better=True, scaling="strong" if w["kind"] == "rmat" else "weak",
                vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
                config=base_config(desc, M, N, nnz, K, args.alpha, args.delta),
                cpu_baseline=dict(value=v, unit=UNIT, cores=vals[0]["cores"], kind=vals[0]["kind"],
                                  sample=vals[0]["sample"]),
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class Ctx:
    """per-process state shared by the records of one bench run"""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        from __graft_entry__ import load_package
        self.torch, self.dist = torch, dist
        self.pkg = load_package()
        self.gen = self.pkg.generators
        self.pkg.lib()  # fail loudly if the CUDA library is missing
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.sm_count = torch.cuda.get_device_properties(self.local).multi_processor_count
        self.mg = None
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            from sddmm_gpu_b200 import multigpu as mgmod
            ident = mgmod.exchange_unique_id(device="cuda")
            self.mg = mgmod.MultiGpu(self.rank, self.world, ident)
            # warm the library's communicator up (connection set-up is not part of any timed broadcast)
            t = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
            self.mg.bcast(t, 0)
            torch.cuda.synchronize()
        else:
            from sddmm_gpu_b200 import multigpu as mgmod
            self.mg = mgmod.MultiGpu(0, 1, None)
        self.args = args

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, vals):
        t = self.torch.tensor(vals, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def sum_over_ranks(self, vals):
        t = self.torch.tensor(vals, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(x) for x in t]

    def timed_steps(self, fn, warmup, steps):
        """W untimed steps, then exactly K steps between barrier + synchronize, CUDA events on the launching
        stream; returns this rank's ms per step"""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        stream = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        self.barrier()
        return e0.elapsed_time(e1) / steps


def device_operands(ctx, M, N, K, seed=1001):
    torch = ctx.torch
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)  # same seed on every rank: same A and (before the broadcast overwrites it) same B
    dA = torch.rand((M, K), device="cuda", generator=g) * 2  # U[0,2) like Matrix::makeData (src/Matrix.cpp:132)
    dB = torch.rand((N, K), device="cuda", generator=g) * 2
    return dA, dB


def fp64_row_check(torch, ro_host, ci_t, dA, dB, P, rows):
    """sampled rows against an fp64 recomputation on the device, with the reference's checkData rule"""
    worst = 0.0
    for r in rows:
        b, e = int(ro_host[r]), int(ro_host[r + 1])
        if e <= b:
            continue
        ref = (dA[int(r)].double()[None, :] * dB[ci_t[b:e].to(torch.int64)].double()).sum(1)
        err = (P[b:e].double() - ref).abs()
        bad = ~((err < 1e-5) | (err / ref.abs().clamp_min(1e-3) < 1e-3))  # include/checkData.hpp:14-30
        assert not bool(bad.any()), f"bench result check failed on row {r}"
        worst = max(worst, float((err / ref.abs().clamp_min(1e-3)).max()))
    return worst


def kernel_roofline(ctx, lay, K, kt, rows_touched, cols_touched, M, nnz, workload_key, plan_names):
    pk = peaks()
    info = lay.info
    dense_dominant = kt["dense_ms"] > kt["sparse_ms"]
    if dense_dominant:
        # tensor-core kernel (128x128 tcgen05 tiles or 16x16 BSMR blocks): useful flops vs the tf32 peak, taken as
        # half of the measured dense bf16 peak (tf32 runs at half the bf16 rate)
        kms = kt["dense_ms"]
        kname = ("k_round_operands + k_sddmm_tile_pair (cta_group::2)" if plan_names["plan"] == "tile" and plan_names["tile"] == "tma_pair"
                 else "k_round_operands + k_sddmm_tile_tma" if plan_names["plan"] == "tile" and plan_names["tile"] != "reg"
                 else "k_sddmm_tile" if plan_names["plan"] == "tile"
                 else "k_round_dense_rows + k_sddmm_dense_tma" if plan_names["dense"] == "tma" else "k_sddmm_dense")
        kname += " (tcgen05 kind::tf32)"
        nnz_k = nnz if plan_names["plan"] == "tile" else int(info.numDenseValues)
        byts = algorithmic_bytes(rows_touched, cols_touched, M, K, nnz_k, int(info.numDenseBlocks), "dense")
        achieved = 2.0 * nnz_k * K / (kms * 1e-3) / 1e12 if kms > 0 else 0.0
        peak = pk["tf"] / 2.0
        tr, src = traffic_entry(f"{workload_key}:dense")
        return dict(bound="tensor", achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak, traffic=tr,
                    traffic_source=src, kernel=kname, kernel_ms=kms, algorithmic_bytes=byts,
                    algorithmic_flops=2.0 * nnz_k * K, bsmr_block_padded_flops=2.0 * 256.0 * info.numDenseBlocks * K,
                    peak_source=pk["which"] + " bf16 / 2",
                    hbm_view=dict(achieved_gbs=byts / (kms * 1e-3) / 1e9 if kms > 0 else 0.0, peak_gbs=pk["hbm_gbs"]),
                    dense_kernel_ms=kt["dense_ms"], residual_kernel_ms=kt["sparse_ms"])
    kms = kt["sparse_ms"]
    kname = ("k_sddmm_residual_sp (fp32 CUDA cores, super-panel)" if plan_names["residual"] == "superpanel"
             else "k_sddmm_residual (fp32 CUDA cores, panel)")
    nres = int(info.numSparseValues)
    byts = algorithmic_bytes(rows_touched, cols_touched, M, K, nres, 0, "residual")
    achieved = byts / (kms * 1e-3) / 1e9 if kms > 0 else 0.0
    gather_bytes = float(nres) * (4.0 * K + 16.0) + 4.0 * K * rows_touched + 4.0 * (M + 1)  # SURVEY 8(d) no-reuse model
    sm_hz = 1.965e9
    tr, src = traffic_entry(f"{workload_key}:residual")
    return dict(bound="hbm", achieved=achieved, peak=pk["hbm_gbs"], unit="GB/s", frac=achieved / pk["hbm_gbs"],
                traffic=tr, traffic_source=src, kernel=kname, kernel_ms=kms, algorithmic_bytes=byts,
                peak_source=pk["which"],
                gather_model=dict(bytes=gather_bytes, achieved_gbs=gather_bytes / (kms * 1e-3) / 1e9 if kms > 0 else 0.0),
                smem_model=dict(bytes=4.0 * K * nres, floor_ms=4.0 * K * nres / (ctx.sm_count * 128 * sm_hz) * 1e3,
                                note=f"A operand: 4K bytes per non-zero through LDS.128, {ctx.sm_count} SMs x 128 B/clk"),
                dense_kernel_ms=kt["dense_ms"], residual_kernel_ms=kt["sparse_ms"])


# ------------------------------------------------------------------------------------------------
def record_rmat(ctx, w, headline=True, scale=None, K=None, reorder=True, steps=None, warmup=None, k_sweep=True,
                e2e=True, cpu=True):
    """ONE R-MAT matrix over all ranks (strong scaling).  Returns the JSON record (rank 0) or None."""
    torch, pkg, args = ctx.torch, ctx.pkg, ctx.args
    from sddmm_gpu_b200 import multigpu as mgmod
    scale = scale or args.scale or w["scale"]
    K = K or args.K or w["K"]
    steps = steps or args.steps
    warmup = warmup if warmup is not None else args.warmup
    alpha, delta = args.alpha, args.delta
    t0 = time.time()
    ro, ci, M = ctx.gen.rmat_device(scale, w["ef"], w["seed"])
    N, nnz = M, int(ci.numel())
    torch.cuda.synchronize()
    gen_s = time.time() - t0
    # every rank generated the same matrix?
    chk = ctx.max_over_ranks([float(nnz), -float(nnz), float(int(ci[::1009].to(torch.int64).sum()) % (1 << 40))])
    assert chk[0] == -chk[1] == float(nnz), "ranks generated different matrices"

    bs = pkg.calculateBlockSize(type("S", (), dict(M=M, N=N))(), FREE_MEM_FOR_BLOCK_SIZE)
    t0 = time.time()
    sh = mgmod.ShardedSDDMM(ctx.mg, ro, ci, M, N, alpha=alpha, delta=delta, block_size=bs, reorder=reorder)
    torch.cuda.synchronize()
    setup_s = time.time() - t0
    lay, info = sh.layout, sh.layout.info
    my_nnz = sh.my_nnz
    tot = ctx.sum_over_ranks([float(my_nnz)])
    assert int(tot[0]) == nnz, f"shards cover {int(tot[0])} of {nnz} stored entries"
    mx_nnz = ctx.max_over_ranks([float(my_nnz)])[0]

    dA, dB = device_operands(ctx, M, N, K)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sh.replicate_B(dB)  # once; the communicator was warmed up in Ctx
    e1.record()
    torch.cuda.synchronize()
    bcast_ms = e0.elapsed_time(e1)
    dP = torch.zeros(max(1, nnz), dtype=torch.float32, device="cuda")
    # cost-calibrated cuts (setup, like the reordering): equal nnz is not equal time on a power-law graph
    cut_history = []
    if ctx.world > 1 and args.rebalance > 0:
        cut_history = sh.calibrate(dA, dB, dP, rounds=args.rebalance)
        lay, info = sh.layout, sh.layout.info
        my_nnz = sh.my_nnz
        tot = ctx.sum_over_ranks([float(my_nnz)])
        assert int(tot[0]) == nnz, f"rebalanced shards cover {int(tot[0])} of {nnz} stored entries"
        mx_nnz = ctx.max_over_ranks([float(my_nnz)])[0]
        dP.zero_()
    pkg.sddmm_prepare(lay, K)

    sampler = ClockSampler(ctx.local)
    sampler.start()
    pkg.launch_count(reset=True)
    my_ms = ctx.timed_steps(lambda: sh.run(dA, dB, dP), warmup, steps)
    launches_per_step = pkg.launch_count() // max(1, warmup + steps)
    kt = pkg.sddmm_gpu_timed(dA, dB, lay, dP, warmup=2, iters=max(3, min(steps, 10)))
    clocks = sampler.stop()
    ms_max = ctx.max_over_ranks([my_ms])[0]
    value = 2.0 * nnz * K / (ms_max * 1e-3) / 1e9

    # result check: sampled rows of THIS rank's panel range, fp64 on the device
    ro_host = ro.cpu().numpy().view(np.uint32)
    Rh = sh.R.cpu().numpy().view(np.uint32)
    lo, hi = int(sh.cuts[ctx.rank]) * 16, min(int(sh.cuts[ctx.rank + 1]) * 16, Rh.size)
    pick = Rh[np.linspace(lo, hi - 1, 16).astype(np.int64)] if hi > lo else []
    worst = fp64_row_check(torch, ro_host, ci, dA, dB, dP, pick)
    worst = ctx.max_over_ranks([worst])[0]

    # ---- K sweep on the same layouts (BASELINE: "at K=32..512")
    sweep = []
    if k_sweep:
        for Ks in (32, 64, 128, 256, 512):
            if Ks == K:
                sweep.append(dict(K=Ks, ms_per_step=ms_max, value=value))
                continue
            if 2.0 * 4.0 * Ks * (M + N) > 60e9:
                continue
            a2, b2 = device_operands(ctx, M, N, Ks)
            pkg.sddmm_prepare(lay, Ks)
            ms = ctx.timed_steps(lambda: sh.run(a2, b2, dP), 3, 5)
            ms = ctx.max_over_ranks([ms])[0]
            ent = dict(K=Ks, ms_per_step=ms, value=2.0 * nnz * Ks / (ms * 1e-3) / 1e9)
            if headline:  # the opt-in fp16 operand copies at this K (not the default arithmetic)
                p16 = pkg.make_plan(plan="bsmr", residual="superpanel", operands="fp16")
                pkg.sddmm_prepare(lay, Ks, 1, p16)
                m16 = ctx.timed_steps(lambda: pkg.sddmm_gpu(a2, b2, lay, dP, plan=p16), 3, 5)
                ent["fp16_operands_ms_per_step"] = ctx.max_over_ranks([m16])[0]
            sweep.append(ent)
            del a2, b2
        torch.cuda.empty_cache()

    # ---- opt-in fp16 operand copies (sddmm_plan.operands = FP16) on the same layouts: fp16 A tile in shared memory and
    # an fp16 copy of the referenced B^T rows rewritten inside every timed pass, FHFMA products, fp32 accumulation.
    # NOT the default arithmetic (the headline stays exact); checked with the same checkData rule.
    fp16_rec = None
    if headline:
        plan16 = pkg.make_plan(plan="bsmr", residual="superpanel", operands="fp16")
        pkg.sddmm_prepare(lay, K, 1, plan16)
        ms16 = ctx.timed_steps(lambda: pkg.sddmm_gpu(dA, dB, lay, dP, plan=plan16), 3, max(3, min(steps, 10)))
        worst16 = fp64_row_check(torch, ro_host, ci, dA, dB, dP, pick)
        ms16, worst16 = ctx.max_over_ranks([ms16, worst16])
        fp16_rec = dict(K=K, ms_per_step=ms16, value=2.0 * nnz * K / (ms16 * 1e-3) / 1e9, max_rel_err_sample=worst16,
                        dtype="fp16 operand copies (opt-in sddmm_plan.operands), fp32 accumulate",
                        note="conversion of the referenced B^T rows to fp16 is inside every timed pass")
        for ent in sweep:
            if ent["K"] == K:
                ent["fp16_operands_ms_per_step"] = ms16
        sh.run(dA, dB, dP)  # leave the exact result in dP
        torch.cuda.synchronize()

    # ---- end to end through the host-buffer entry point, every rank on its own shard (pinned host memory;
    # each step copies its own A and B in and its whole P out inside the timed region)
    e2e_rec = None
    if e2e:
        e2e_steps = max(2, min(steps, 8))
        hA = dA.cpu().pin_memory()
        hB = dB.cpu().pin_memory()
        hP = [torch.zeros(max(1, nnz), dtype=torch.float32).pin_memory() for _ in range(2)]
        nA, nB, nPs = hA.numpy(), hB.numpy(), [t.numpy() for t in hP]
        W = ctx.world
        if W == 1:
            # one GPU: the streaming host-buffer entry point, two slots (H2D of step i+1, kernels of step i and D2H
            # of step i-1 overlap); every step moves all of A and B in and all of P out
            for i in range(2):
                pkg.sddmm_gpu_async(nA, nB, lay, nPs[i], i)
            pkg.sddmm_gpu_sync(lay)
            ctx.barrier()
            t0 = time.perf_counter()
            for i in range(e2e_steps):
                pkg.sddmm_gpu_async(nA, nB, lay, nPs[i & 1], i & 1)
            pkg.sddmm_gpu_sync(lay)
            torch.cuda.synchronize()
            e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
            assert np.array_equal(nPs[0], nPs[1]), "pipelined slots disagree"
            api = "sddmm_run_host_async, 2 slots (H2D / kernels / D2H of consecutive steps overlap)"
            moved = pkg.host_traffic(lay)  # what the entry point really moved (referenced rows only when it can)
            hPt = torch.from_numpy(nPs[(e2e_steps - 1) & 1]).cuda()
            fp64_row_check(torch, ro_host, ci, dA, dB, hPt, pick)
            del hPt
        else:
            # N GPUs: sddmm_mgpu_run_host -- every rank copies 1/N of A and of B over its own PCIe link, the slices are
            # all-gathered over NVLink, P is summed onto rank 0, which copies it to the host: the job reads the host
            # ONCE per step
            for i in range(2):
                ctx.mg.run_host(lay, nA, nB, nPs[0] if ctx.rank == 0 else None, 0)
            ctx.barrier()
            t0 = time.perf_counter()
            for i in range(e2e_steps):
                ctx.mg.run_host(lay, nA, nB, nPs[i & 1] if ctx.rank == 0 else None, 0)
            torch.cuda.synchronize()
            e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
            api = ("sddmm_mgpu_run_host: 1/N of the referenced rows of A and B per rank over PCIe (gathered through the "
                   "mapped pinned buffers), all-gather over NVLink, P reduced to rank 0 and copied out")
            moved = (int(ctx.sum_over_ranks([float(ctx.mg.host_traffic())])[0]), int(4 * nnz))
            if ctx.rank == 0:  # the merged result on the host against fp64, rows from every shard
                pick_all = Rh[np.linspace(0, Rh.size - 1, 24).astype(np.int64)]
                hPt = torch.from_numpy(nPs[(e2e_steps - 1) & 1]).cuda()
                fp64_row_check(torch, ro_host, ci, dA, dB, hPt, pick_all)
                del hPt
        ctx.barrier()
        # host-copy ceiling: the same bytes per step moved by bare cudaMemcpyAsync (both directions concurrently),
        # all ranks at once -- what the box's PCIe / host memory system gives N ranks, kernels and NVLink excluded
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        fa, fb = dA.view(-1), dB.view(-1)
        ha, hb = hA.view(-1), hB.view(-1)
        ca, cb = -(-fa.numel() // W), -(-fb.numel() // W)
        sa = slice(min(fa.numel(), ctx.rank * ca), min(fa.numel(), (ctx.rank + 1) * ca))
        sb_ = slice(min(fb.numel(), ctx.rank * cb), min(fb.numel(), (ctx.rank + 1) * cb))
        ctx.barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            with torch.cuda.stream(s_in):
                fa[sa].copy_(ha[sa], non_blocking=True)
                fb[sb_].copy_(hb[sb_], non_blocking=True)
            if ctx.rank == 0:
                with torch.cuda.stream(s_out):
                    hP[i & 1].copy_(dP, non_blocking=True)
        torch.cuda.synchronize()
        ceil_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        ctx.barrier()
        e2e_ms_max, ceil_ms_max = ctx.max_over_ranks([e2e_ms, ceil_ms])
        full_in = int(4 * K * (M + N))
        h2d_b, d2h_b = moved
        e2e_rec = dict(value=2.0 * nnz * K / (e2e_ms_max * 1e-3) / 1e9, unit=UNIT, ms_per_step=e2e_ms_max,
                       steps=e2e_steps, h2d_bytes_per_step=int(h2d_b), d2h_bytes_per_step=int(d2h_b),
                       bytes_note=((f"whole job: the referenced rows of A and B ({h2d_b / full_in:.0%} of the arrays) "
                                    "enter once per step (1/N per rank), P leaves once (rank 0)") if W > 1 else
                                   ("the A rows and B^T rows the pass reads (non-empty rows / referenced columns of S: "
                                    f"{h2d_b / full_in:.0%} of the arrays) gathered from the pinned host buffers, all of "
                                    "P out, every step" if h2d_b < full_in else
                                    "all of A and B in, all of P out, every step")),
                       api=api,
                       host_copy_ceiling=dict(ms_per_step=ceil_ms_max, value=2.0 * nnz * K / (ceil_ms_max * 1e-3) / 1e9,
                                              note="ALL of A and B in and P out per step by bare pinned cudaMemcpyAsync "
                                                   "on two streams, all ranks concurrently, no kernels, no NVLink"))
        del hA, hB, hP, ha, hb

    cpu_rec = None
    if cpu and ctx.world == 1 and ctx.rank == 0 and not args.no_cpu:
        ci_host = ci.cpu().numpy().view(np.uint32)
        A_h, B_h = dA.cpu().numpy(), dB.cpu().numpy()
        rows = int(min(M, max(256, M * min(1.0, args.cpu_gflop * 1e9 / max(1.0, 2.0 * nnz * K)))))
        c = cpu_reference_gflops(ro_host, ci_host, M, N, A_h, B_h, K, rows)
        cpu_rec = dict(value=c["value"], unit=UNIT, cores=c["cores"], kind=c["kind"], sample=c["sample"])
        del A_h, B_h

    plan_names = pkg.plan_resolve(lay, K)
    rec = None
    if ctx.rank == 0:
        rows_touched = int(info.numRows)
        cols_touched = int(torch.unique(ci).numel())
        roof = kernel_roofline(ctx, lay, K, kt, rows_touched, cols_touched, M, my_nnz, f"rmat{scale}_k{K}", plan_names)
        roof["note"] = ("rank 0's shard; bytes = 4K(rows of the shard + distinct columns of S) + 16 B per residual "
                        "entry + row offsets (SURVEY.md 8d compulsory model)")
        desc = w["desc"] + f", scale {scale}"
        cfg = base_config(desc, M, N, nnz, K, alpha, delta)
        cfg.update(block_size=int(bs), reordered=bool(reorder), parallelism=f"one matrix, nnz-balanced row-panel ranges x{ctx.world}, B replicated once (NCCL), no steady-state collective",
                   largest_shard_nnz_share=mx_nnz / max(1, nnz), rebalance_rounds=len(cut_history),
                   panel_cuts=[int(c) for c in sh.cuts], l2="working set (A+B+layout+P) >> 126 MB L2, no flush between steps",
                   num_row_panels_rank0=int(info.numRowPanels), num_clusters=int(sh.num_clusters),
                   dense_blocks_rank0=int(info.numDenseBlocks), dense_nnz_rank0=int(info.numDenseValues),
                   residual_nnz_rank0=int(info.numSparseValues), row_reorder_ms=sh.row_ms, col_reorder_ms=sh.col_ms,
                   rphm_build_ms=sh.rphm_ms, shard_setup_s=round(setup_s, 2), b_broadcast_ms=bcast_ms,
                   b_broadcast_gbs=(4.0 * K * N / (bcast_ms * 1e-3) / 1e9) if ctx.world > 1 and bcast_ms > 0 else None,
                   datagen_s=round(gen_s, 2), kernels=plan_names, max_rel_err_sample=worst, k_sweep=sweep)
        if fp16_rec:
            cfg["fp16_operands"] = fp16_rec
        rec = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=ctx.world, steps=steps, warmup=warmup,
                   ms_per_step=ms_max, higher_is_better=True, scaling="strong", vs_baseline=None,
                   dtype="tf32 (dense blocks, fp32 accumulate) / f32 (residual)", data="synthetic", config=cfg,
                   clocks=clocks, gpu_launches=int(launches_per_step * steps), gpu_launches_per_step=int(launches_per_step),
                   gpu_launches_note="rank 0's kernels inside the timed region", roofline=roof)
        if e2e_rec:
            rec["e2e"] = e2e_rec
        if cpu_rec:
            rec["cpu_baseline"] = cpu_rec
    del dA, dB, dP, sh, lay
    torch.cuda.empty_cache()
    return rec


def record_single(ctx, name, w, K=None, steps=None, warmup=None, e2e=True, cpu=True, headline=False, plan_kw=None):
    """single-GPU record (configs 2 / 3 and the block-structured case): whole pipeline on this GPU.
    plan_kw: keywords of host.make_plan for a non-default plan (the opt-in fp16-operand records)."""
    torch, pkg, args, gen = ctx.torch, ctx.pkg, ctx.args, ctx.gen
    K = K or (args.K if headline and args.K else w["K"])
    steps = steps or args.steps
    warmup = warmup if warmup is not None else args.warmup
    alpha, delta = args.alpha, args.delta
    t0 = time.time()
    S = make_pattern(gen, w, args)
    gen_s = time.time() - t0
    A, B = gen.dense_operands(S.M, S.N, K)
    ro = torch.from_numpy(S.row_off.view(np.int32)).cuda()
    ci = torch.from_numpy(S.col_idx.view(np.int32)).cuda()
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    dP = torch.zeros(max(1, S.nnz), dtype=torch.float32, device="cuda")
    bs = pkg.calculateBlockSize(S, FREE_MEM_FOR_BLOCK_SIZE)
    R, ncl, row_ms = pkg.row_reorder_dev(ro, ci, S.M, S.N, alpha, bs)
    lay, col_ms, rphm_ms = pkg.layout_build_dev(ro, ci, S.M, S.N, R, delta)
    info = lay.info
    plan = pkg.make_plan(**plan_kw) if plan_kw else None
    pkg.sddmm_prepare(lay, K, 1, plan)
    plan_names = pkg.plan_resolve(lay, K, 1, plan)

    sampler = ClockSampler(ctx.local)
    sampler.start()
    pkg.launch_count(reset=True)
    # launch-bound regime (4096^2 masks: tens of microseconds per pass): one CUDA graph per step
    use_graph = S.nnz * K < 2e9
    if use_graph:
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            pkg.sddmm_gpu(dA, dB, lay, dP, plan=plan)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                pkg.sddmm_gpu(dA, dB, lay, dP, plan=plan)
        launches_per_step = pkg.launch_count() // 2
        step = graph.replay
    else:
        pkg.sddmm_gpu(dA, dB, lay, dP, plan=plan)
        launches_per_step = pkg.launch_count()
        step = lambda: pkg.sddmm_gpu(dA, dB, lay, dP, plan=plan)  # noqa: E731
    ms = ctx.timed_steps(step, warmup, steps)
    if plan is None:
        kt = pkg.sddmm_gpu_timed(dA, dB, lay, dP, warmup=2, iters=max(3, min(steps, 10)))
    else:  # forced plan: one kernel class serves the pass; its time is the step time
        kt = dict(dense_ms=ms if plan_names["plan"] == "tile" else 0.0, sparse_ms=0.0 if plan_names["plan"] == "tile" else ms,
                  total_ms=ms)
    clocks = sampler.stop()
    value = 2.0 * S.nnz * K / (ms * 1e-3) / 1e9
    rows = np.random.default_rng(0).choice(S.M, min(8, S.M), replace=False)
    worst = fp64_row_check(torch, S.row_off, ci, dA, dB, dP, rows)

    e2e_rec = None
    if e2e:
        hA, hB = torch.from_numpy(A).pin_memory(), torch.from_numpy(B).pin_memory()
        hP = [torch.zeros(max(1, S.nnz), dtype=torch.float32).pin_memory() for _ in range(2)]
        nA, nB, nPs = hA.numpy(), hB.numpy(), [t.numpy() for t in hP]
        for _ in range(2):
            pkg.sddmm_gpu(nA, nB, lay, nPs[0])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            pkg.sddmm_gpu(nA, nB, lay, nPs[0])
        torch.cuda.synchronize()
        sync_ms = (time.perf_counter() - t0) * 1e3 / steps
        for i in range(2):
            pkg.sddmm_gpu_async(nA, nB, lay, nPs[i], i)
        pkg.sddmm_gpu_sync(lay)
        t0 = time.perf_counter()
        for i in range(steps):
            pkg.sddmm_gpu_async(nA, nB, lay, nPs[i & 1], i & 1)
        pkg.sddmm_gpu_sync(lay)
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / steps
        assert np.array_equal(nPs[0], nPs[1]), "pipelined slots disagree"
        h2d_b, d2h_b = pkg.host_traffic(lay)
        e2e_rec = dict(value=2.0 * S.nnz * K / (e2e_ms * 1e-3) / 1e9, unit=UNIT, ms_per_step=e2e_ms,
                       h2d_bytes_per_step=int(h2d_b), d2h_bytes_per_step=int(d2h_b),
                       api="sddmm_run_host_async, 2 slots (H2D / kernels / D2H of consecutive steps overlap)",
                       sync_api_ms_per_step=sync_ms, sync_api_value=2.0 * S.nnz * K / (sync_ms * 1e-3) / 1e9)
    cpu_rec = None
    if cpu and not args.no_cpu:
        rows_c = int(min(S.M, max(256, S.M * min(1.0, args.cpu_gflop * 1e9 / max(1.0, 2.0 * S.nnz * K)))))
        c = cpu_reference_gflops(S.row_off, S.col_idx, S.M, S.N, A, B, K, rows_c)
        cpu_rec = dict(value=c["value"], unit=UNIT, cores=c["cores"], kind=c["kind"], sample=c["sample"])
    cols_touched = int(np.unique(S.col_idx).size)
    roof = kernel_roofline(ctx, lay, K, kt, int(info.numRows), cols_touched, S.M, S.nnz, f"{name}_k{K}", plan_names)
    cfg = base_config(w["desc"], S.M, S.N, S.nnz, K, alpha, delta)
    all_tf32 = plan_names["plan"] == "tile"
    fp16 = plan_names.get("operands") == "fp16"
    cfg.update(block_size=int(bs), reordered=True, parallelism="single GPU", cuda_graph_per_step=bool(use_graph),
               l2=("working set >> 126 MB L2, no flush" if 4.0 * K * (S.M + S.N) + 20.0 * S.nnz > 200e6
                   else "working set fits the 126 MB L2: L2-resident by design (the reference's loop is the same)"),
               num_row_panels=int(info.numRowPanels), num_clusters=int(ncl), dense_blocks=int(info.numDenseBlocks),
               dense_nnz=int(info.numDenseValues), residual_nnz=int(info.numSparseValues), row_reorder_ms=row_ms,
               col_reorder_ms=col_ms, rphm_build_ms=rphm_ms, datagen_s=round(gen_s, 2), kernels=plan_names,
               max_rel_err_sample=worst)
    rec = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=1, steps=steps, warmup=warmup, ms_per_step=ms,
               higher_is_better=True, scaling="weak", vs_baseline=None,
               dtype=("fp16 operand copies (opt-in sddmm_plan.operands), fp32 accumulate" if fp16
                      else "tf32 for ALL entries (whole 128x128 tcgen05 tiles), fp32 accumulate" if all_tf32
                      else "tf32 (dense blocks, fp32 accumulate) / f32 (residual)"),
               data="synthetic", config=cfg, clocks=clocks, gpu_launches=int(launches_per_step * steps),
               gpu_launches_per_step=int(launches_per_step), roofline=roof)
    if e2e_rec:
        rec["e2e"] = e2e_rec
    if cpu_rec:
        rec["cpu_baseline"] = cpu_rec
    del dA, dB, dP, lay
    torch.cuda.empty_cache()
    return rec


def slim(rec):
    """nested `also` records: drop the bulky parts"""
    out = {k: rec[k] for k in ("value", "unit", "ms_per_step", "dtype", "n_gpus") if k in rec}
    out["config"] = rec["config"]
    out["roofline"] = {k: rec["roofline"][k] for k in ("bound", "achieved", "peak", "unit", "frac", "kernel", "kernel_ms",
                                                        "dense_kernel_ms", "residual_kernel_ms") if k in rec["roofline"]}
    if "e2e" in rec:
        out["e2e"] = {k: rec["e2e"][k] for k in ("value", "ms_per_step") if k in rec["e2e"]}
    return out


def run_ours(args, w):
    ctx = Ctx(args)
    if w["kind"] != "rmat":
        if ctx.world > 1:
            raise SystemExit("bench.py: --gpus N > 1 shards ONE R-MAT matrix (--workload rmat); the other workloads are "
                             "single-GPU records")
        rec = record_single(ctx, args.workload, w, headline=True)
        print(json.dumps(rec), flush=True)
        return
    rec = record_rmat(ctx, w)
    also = {}
    if not args.no_also:
        # config 5: R-MAT scale 25, K=256, the largest config, at every N (identity row order: the sequential
        # clustering of 33.5 M rows does not finish in a bench run; labelled in the record)
        if args.cfg5:
            try:
                r5 = record_rmat(ctx, w, headline=False, scale=25, K=256, reorder=False, steps=5, warmup=3, k_sweep=False,
                                 e2e=False, cpu=False)
                if r5:
                    also["config5_rmat25_k256"] = slim(r5)
            except Exception as e:  # never lose the headline to an optional record
                also["config5_rmat25_k256"] = dict(error=str(e)[:300])
        if ctx.world == 1:
            fp16_sp = dict(plan="bsmr", residual="superpanel", operands="fp16")
            fp16_tile = dict(plan="tile", operands="fp16")
            for name, K, plan_kw in (("uniform100k", 128, None), ("dlmc4096_s70", 256, None), ("dlmc4096_s70", 64, None),
                                     ("dlmc4096_s90", 256, None), ("dlmc4096_s98", 64, None), ("blockscat16k", 128, None),
                                     # opt-in fp16 operand copies: NOT the default arithmetic, labelled in `dtype`
                                     ("uniform100k", 128, fp16_sp), ("dlmc4096_s70", 256, fp16_tile)):
                key = f"{name}_k{K}" + ("_fp16_operands" if plan_kw else "")
                try:
                    r = record_single(ctx, name, WORKLOADS[name], K=K, steps=10, warmup=3,
                                      e2e=(name == "uniform100k" and plan_kw is None), cpu=False, plan_kw=plan_kw)
                    also[key] = slim(r) if "e2e" not in r else {**slim(r), "e2e": r["e2e"]}
                except Exception as e:
                    also[key] = dict(error=str(e)[:300])
    if ctx.rank == 0:
        if also:
            rec["config"]["also"] = also
        print(json.dumps(rec), flush=True)
    if ctx.world > 1:
        ctx.mg.close()
        ctx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="rmat", choices=sorted(WORKLOADS))
    ap.add_argument("--K", type=int, default=0)
    ap.add_argument("--alpha", type=float, default=0.3)
    ap.add_argument("--delta", type=float, default=0.3)
    ap.add_argument("--rows", type=int, default=0, help="debug: override the row count of a uniform workload")
    ap.add_argument("--scale", type=int, default=0, help="R-MAT scale override")
    ap.add_argument("--cpu-gflop", type=float, default=8.0, help="GFLOP of work in the bounded cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="headline record only")
    ap.add_argument("--rebalance", type=int, default=2, help="cost-calibration rounds of the panel cuts (N > 1)")
    ap.add_argument("--cfg5", type=int, default=1, help="also run config 5 (R-MAT scale 25, K=256) at this N")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
