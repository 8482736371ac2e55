#!/usr/bin/env python
"""bench.py -- effective SDDMM GFLOP/s (2*nnz*K/t), the reference's headline metric
(include/Logger.hpp:178-180), on BASELINE.json's configs, on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

* A "step" is one SDDMM pass (dense-block tcgen05 kernel || residual CUDA-core kernel) over one (A, B)
  batch with the BSMR/RPHM layout resident, exactly what the reference's sddmmTime_ loop times
  (src/sddmmKernel.cu:2561-2659).  Reordering / layout-build times are reported beside it in `config`.
* N > 1 (torchrun): every rank owns an independent row-panel shard (its own 100k-row slab of a taller
  matrix) and all of B, which rank 0 broadcasts ONCE over NCCL before the timed region; there is no
  collective in the steady state.  scaling = weak.
* `value`  : inputs resident in HBM, CUDA events on the launching stream, max over ranks.
* `e2e`    : the same metric through the host-buffer entry point sddmm_run_host (pinned host A, B in,
             host P out; H2D + D2H inside the timed region).
* `roofline`: dominant kernel, ALGORITHMIC bytes (DESIGN.md "bytes per unit") / its mean launch time.
* `cpu_baseline`: the reference's OpenMP sddmm_cpu (oracle/_ref/libref_cpu.so when present, else the
             oracle port) on a bounded row sample of the same workload, on this box's host cores.
--impl reference times that CPU implementation as the reference arm.
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

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on at N=1
    "uniform100k": dict(kind="uniform", M=100_000, N=100_000, density=0.01, K=128, seed=2,
                        desc="synthetic uniform-random 100k x 100k, 1% density, K=128"),
    # configs[2]: DLMC-style pruned masks
    "dlmc4096_s70": dict(kind="bernoulli", M=4096, N=4096, sparsity=0.70, K=256, seed=30,
                         desc="synthetic pruned mask 4096 x 4096 at 70% sparsity, K=256"),
    "dlmc4096_s90": dict(kind="dlmc", M=4096, N=4096, sparsity=0.90, K=64, seed=33,
                         desc="synthetic DLMC-style magnitude-pruned mask 4096 x 4096 at 90% sparsity, K=64"),
    # configs[3] (scaled by --scale): R-MAT power-law
    "rmat": dict(kind="rmat", scale=18, ef=16, K=128, seed=4, desc="synthetic R-MAT power-law graph"),
    "small": dict(kind="uniform", M=4096, N=4096, density=0.02, K=64, seed=2, desc="smoke-size uniform"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=float(d["hbm_gbs"]), tf=float(d.get("bf16_tflops", 0)), which="measured")
    return dict(hbm_gbs=6650.0, tf=1590.0, which="fallback")


def make_pattern(gen, w, rank, args):
    seed = w["seed"] + 1000 * rank
    if w["kind"] == "uniform":
        M = args.rows or w["M"]
        return gen.uniform_random(M, w["N"] if not args.rows else min(w["N"], max(M, 1024)), w["density"], seed)
    if w["kind"] == "bernoulli":
        return gen.bernoulli_mask(w["M"], w["N"], w["sparsity"], seed)
    if w["kind"] == "dlmc":
        return gen.dlmc_magnitude_mask(w["M"], w["N"], w["sparsity"], seed)
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


def algorithmic_bytes(M, N, K, nnz_kernel, n_blocks, kernel):
    """SURVEY.md 8(d) compulsory-traffic model: every operand element touched once.
    residual: A + B + (col, relRow, csrIdx: 12 B) + P (4 B) per entry + row offsets;
    dense   : A + B + per block 1024 B blockValues + 64 B denseCols + 4 B per stored entry of P."""
    base = 4.0 * K * (M + N) + 4.0 * (M + 1)
    if kernel == "residual":
        return base + 16.0 * nnz_kernel
    return base + 1088.0 * n_blocks + 4.0 * nnz_kernel


# ------------------------------------------------------------------------------------------------
def cpu_reference_gflops(S, A, B, K, sample_rows, repeats=3):
    """The reference's OpenMP sddmm_cpu on the first `sample_rows` rows (same S/A/B).  -> dict"""
    from oracle import oracle as O
    rows = min(sample_rows, S.M)
    nnz = int(S.row_off[rows])
    ro = np.ascontiguousarray(S.row_off[: rows + 1])
    ci = np.ascontiguousarray(S.col_idx[:nnz])
    As = np.ascontiguousarray(A[:rows])
    cores = os.cpu_count() or 1
    best = float("inf")
    if O.ref_available():
        kind = "reference"
        ctx = O.ref().ref_sddmm_prepare(As, B, ro, ci, rows, S.N, K, nnz)
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
            O.lib().oracle_sddmm_cpu(As, B, ro, ci, rows, K, P, 0)
            best = min(best, time.perf_counter() - t0)
    return dict(value=2.0 * nnz * K / best / 1e9, unit=UNIT, cores=cores, kind=kind, seconds=best,
                sample=f"first {rows} rows ({nnz} nnz) of the same S with the same A/B, best of {repeats}, "
                       f"OMP threads = {cores}")


def run_reference(args, w):
    """Reference arm: the reference's own CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from __graft_entry__ import load_package
    gen = load_package().generators
    S = make_pattern(gen, w, 0, args)
    K = args.K or w["K"]
    A, B = gen.dense_operands(S.M, S.N, K)
    # bounded sample: ~2 s of CPU work per step
    sample_rows = min(S.M, max(256, int(S.M * min(1.0, 4e9 / max(1.0, 2.0 * S.nnz * K)))))
    vals = []
    for i in range(args.warmup + args.steps):
        r = cpu_reference_gflops(S, A, B, K, sample_rows, repeats=1)
        if i >= args.warmup:
            vals.append(r)
    v = float(np.mean([r["value"] for r in vals]))
    ms = float(np.mean([r["seconds"] for r in vals])) * 1e3
    line = dict(metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference",
                config=dict(workload=w["desc"], M=S.M, N=S.N, nnz=S.nnz, K=K),
                cpu_baseline=dict(value=v, unit=UNIT, cores=vals[0]["cores"], kind=vals[0]["kind"],
                                  sample=vals[0]["sample"]),
                e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def run_ours(args, w):
    import torch
    import torch.distributed as dist

    from __graft_entry__ import load_package
    pkg = load_package()
    gen = pkg.generators
    pkg.lib()  # fail loudly if the CUDA library is missing

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K = args.K or w["K"]
    alpha, delta = args.alpha, args.delta

    # ---- synthetic inputs: this rank's row slab of S and A; B replicated from rank 0 (NCCL, once)
    t0 = time.time()
    S = make_pattern(gen, w, rank, args)
    A = (np.random.default_rng(1001 + rank).random((S.M, K), dtype=np.float32) * np.float32(2)).astype(np.float32)
    if rank == 0:
        B = (np.random.default_rng(1002).random((S.N, K), dtype=np.float32) * np.float32(2)).astype(np.float32)
        dB = torch.from_numpy(B).cuda()
    else:
        dB = torch.empty((S.N, K), dtype=torch.float32, device="cuda")
    bcast_ms = 0.0
    if world > 1:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.broadcast(dB, src=0)
        e1.record()
        torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
        B = dB.cpu().numpy()
    gen_s = time.time() - t0

    ro = torch.from_numpy(S.row_off.view(np.int32)).cuda()
    ci = torch.from_numpy(S.col_idx.view(np.int32)).cuda()
    dA = torch.from_numpy(A).cuda()
    dP = torch.zeros(max(1, S.nnz), dtype=torch.float32, device="cuda")

    # ---- reorder + layout (reported, not part of the step: same accounting as the reference's log)
    bs = pkg.calculateBlockSize(S, 180 * 10 ** 9)  # fixed free-memory figure: reproducible block size (H3)
    R, ncl, row_ms = pkg.row_reorder_dev(ro, ci, S.M, S.N, alpha, bs)
    lay, col_ms, rphm_ms = pkg.layout_build_dev(ro, ci, S.M, S.N, R, delta)
    info = lay.info
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing
    for _ in range(args.warmup):
        pkg.sddmm_gpu(dA, dB, lay, dP)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    pkg.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        pkg.sddmm_gpu(dA, dB, lay, dP)
    e1.record(stream)
    barrier()
    launches = pkg.launch_count()
    ms_step = e0.elapsed_time(e1) / args.steps
    # per-kernel durations (each alone on its stream, CUDA events on that stream) for the roofline
    kt = pkg.sddmm_gpu_timed(dA, dB, lay, dP, warmup=2, iters=max(3, args.steps))
    clocks = sampler.stop()

    # ---- end to end through the host-buffer entry points (pinned host memory).
    # (1) synchronous call per step (the reference-shaped sddmm_gpu(Matrix...) overload);
    # (2) its streaming twin: two slots, so the H2D of step i+1, the kernels of step i and the D2H of step
    #     i-1 overlap.  Every step still copies its own A and B in and its whole P out inside the timed region.
    hA = torch.from_numpy(A).pin_memory()
    hB = torch.from_numpy(B).pin_memory()
    hP = [torch.zeros(max(1, S.nnz), dtype=torch.float32).pin_memory() for _ in range(2)]
    nA, nB, nPs = hA.numpy(), hB.numpy(), [t.numpy() for t in hP]
    nP = nPs[0]
    for _ in range(min(2, args.warmup)):
        pkg.sddmm_gpu(nA, nB, lay, nP)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pkg.sddmm_gpu(nA, nB, lay, nP)
    torch.cuda.synchronize()
    e2e_sync_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    for i in range(2):
        pkg.sddmm_gpu_async(nA, nB, lay, nPs[i], i)
    pkg.sddmm_gpu_sync(lay)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        pkg.sddmm_gpu_async(nA, nB, lay, nPs[i & 1], i & 1)
    pkg.sddmm_gpu_sync(lay)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    barrier()
    assert np.array_equal(nPs[0], nPs[1]), "pipelined slots disagree"
    # spot-check the e2e result against a float64 recomputation of a few rows (not timed)
    rows = np.random.default_rng(0).choice(S.M, 8, replace=False)
    for r in rows:
        b, e = int(S.row_off[r]), int(S.row_off[r + 1])
        if e > b:
            ref = (A[r][None, :].astype(np.float64) * B[S.col_idx[b:e]]).sum(1)
            err = np.abs(nP[b:e] - ref) / np.maximum(np.abs(ref), 1e-3)
            assert err.max() < 1e-3, f"bench result check failed on row {r}: {err.max()}"

    # ---- reduce over ranks: time = max, work = sum
    tt = torch.tensor([ms_step, e2e_ms, e2e_sync_ms], dtype=torch.float64, device="cuda")
    nn = torch.tensor([float(S.nnz)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(nn, op=dist.ReduceOp.SUM)
    ms_step_max, e2e_ms_max, e2e_sync_ms_max = float(tt[0]), float(tt[1]), float(tt[2])
    total_nnz = float(nn[0])
    value = 2.0 * total_nnz * K / (ms_step_max * 1e-3) / 1e9
    e2e_value = 2.0 * total_nnz * K / (e2e_ms_max * 1e-3) / 1e9

    if rank == 0:
        pk = peaks()
        dense_dominant = kt["dense_ms"] > kt["sparse_ms"]
        gather_bytes = float(S.nnz) * (4.0 * K + 16.0) + 4.0 * K * S.M + 4.0 * (S.M + 1)  # SURVEY 8(d) no-reuse model
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(f"{args.workload}:{'dense' if dense_dominant else 'residual'}")
        if dense_dominant:
            # tensor-core kernel (128x128 tcgen05 tiles or 16x16 BSMR blocks): padded flops vs the tf32 peak,
            # taken as half of the measured dense bf16 peak (tf32 runs at half the bf16 rate)
            kname, kms = "k_round_operands + k_sddmm_tile_tma / k_sddmm_tile / k_sddmm_dense (tcgen05 kind::tf32)", kt["dense_ms"]
            tiles = (-(-S.M // 128)) * (-(-S.N // 128))
            padded = 2.0 * 16384.0 * tiles * K if info.numDenseBlocks == 0 or True else 0.0
            padded_blocks = 2.0 * 256.0 * info.numDenseBlocks * K
            flops = max(padded_blocks, 2.0 * S.nnz * K)
            byts = algorithmic_bytes(S.M, S.N, K, S.nnz, 0, "dense")
            achieved = 2.0 * S.nnz * K / (kms * 1e-3) / 1e12 if kms > 0 else 0.0
            peak = pk["tf"] / 2.0
            roof = dict(bound="tensor", achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak,
                        traffic=traffic, kernel=kname, kernel_ms=kms, algorithmic_bytes=byts,
                        algorithmic_flops=2.0 * S.nnz * K, dense_tile_flops_if_all_tiles=padded,
                        bsmr_block_padded_flops=padded_blocks, peak_source=pk["which"] + " bf16 / 2",
                        hbm_view=dict(achieved_gbs=byts / (kms * 1e-3) / 1e9 if kms > 0 else 0.0, peak_gbs=pk["hbm_gbs"]),
                        dense_kernel_ms=kt["dense_ms"], residual_kernel_ms=kt["sparse_ms"])
        else:
            kname, kms = "k_sddmm_residual_sp (fp32 CUDA cores, super-panel)", kt["sparse_ms"]
            byts = algorithmic_bytes(S.M, S.N, K, info.numSparseValues, 0, "residual")
            achieved = byts / (kms * 1e-3) / 1e9 if kms > 0 else 0.0
            roof = dict(bound="hbm", achieved=achieved, peak=pk["hbm_gbs"], unit="GB/s", frac=achieved / pk["hbm_gbs"],
                        traffic=traffic, kernel=kname, kernel_ms=kms, algorithmic_bytes=byts, peak_source=pk["which"],
                        gather_model=dict(bytes=gather_bytes, achieved_gbs=gather_bytes / (kms * 1e-3) / 1e9 if kms > 0 else 0.0),
                        smem_model=dict(bytes=4.0 * K * info.numSparseValues,
                                        floor_ms=4.0 * K * info.numSparseValues / (148 * 128 * 1.965e9) * 1e3,
                                        note="A operand: 4K bytes per non-zero through LDS.128, 148 SMs x 128 B/clk"),
                        dense_kernel_ms=kt["dense_ms"], residual_kernel_ms=kt["sparse_ms"])
        cpu = None
        if world == 1 and not args.no_cpu:
            c = cpu_reference_gflops(S, A, B, K, args.cpu_rows)
            cpu = dict(value=c["value"], unit=UNIT, cores=c["cores"], kind=c["kind"], sample=c["sample"])
        line = dict(
            metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
            ms_per_step=ms_step_max, higher_is_better=True, scaling="weak", vs_baseline=None,
            dtype="tf32 (dense blocks, fp32 accumulate) / f32 (residual)", data="synthetic",
            config=dict(workload=w["desc"], per_gpu=dict(M=S.M, N=S.N, nnz=S.nnz), K=K, alpha=alpha, delta=delta,
                        block_size=bs, total_nnz=int(total_nnz), parallelism=f"row-panel shards x{world}, B replicated",
                        l2="working set (A+B+layout+P) >> 126 MB L2, no flush between steps",
                        num_row_panels=int(info.numRowPanels), num_clusters=int(ncl),
                        dense_blocks=int(info.numDenseBlocks), dense_nnz=int(info.numDenseValues),
                        residual_nnz=int(info.numSparseValues), row_reorder_ms=row_ms, col_reorder_ms=col_ms,
                        rphm_build_ms=rphm_ms, b_broadcast_ms=bcast_ms, datagen_s=round(gen_s, 2)),
            clocks=clocks,
            e2e=dict(value=e2e_value, unit=UNIT, ms_per_step=e2e_ms_max,
                     h2d_bytes_per_step=int(4 * K * (S.M + S.N)), d2h_bytes_per_step=int(4 * S.nnz),
                     api="sddmm_run_host_async, 2 slots (H2D / kernels / D2H of consecutive steps overlap)",
                     sync_api_ms_per_step=e2e_sync_ms_max,
                     sync_api_value=2.0 * total_nnz * K / (e2e_sync_ms_max * 1e-3) / 1e9),
            gpu_launches=int(launches), roofline=roof)
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="uniform100k", choices=sorted(WORKLOADS))
    ap.add_argument("--K", type=int, default=0)
    ap.add_argument("--alpha", type=float, default=0.3)
    ap.add_argument("--delta", type=float, default=0.3)
    ap.add_argument("--rows", type=int, default=0, help="debug: override the row count of a uniform workload")
    ap.add_argument("--scale", type=int, default=0, help="R-MAT scale override")
    ap.add_argument("--cpu-rows", type=int, default=8192, help="rows of S in the bounded cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
