"""Generate golden vectors by RUNNING THE UNMODIFIED REFERENCE on a GPU box.

    gpurun -- python tests/make_ref_goldens.py [--big]

For every seeded case below this script writes a case file, runs oracle/_ref/ref_dump (the
reference's own BSMR -> RPHM -> sddmm_gpu pipeline, built by oracle/Makefile from
/root/reference with only the arch flag changed) and stores, under gpurun_out/ref_gpu/:
  <case>.npz   reorderedRows + the BSMR/RPHM arrays (small cases: full arrays; all cases: sha256)
  summary.json per-case reference timings, and whether oracle/bsmr_oracle.c reproduces every
               array bit-for-bit on the same input (checked right here, on the box).
The committed copies live in tests/golden/ref_gpu/ (copied from gpurun_out by hand after the run).
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

gen = load_package().generators
from oracle import oracle as O  # noqa: E402

U32 = ["reorderedRows", "denseCols", "denseColOffsets", "sparseCols", "sparseColOffsets", "sparseValueOffsets",
       "blockOffsets", "blockValues", "sparseValues", "sparseRelativeRows", "sparseColIndices",
       "denseRowPanelIds", "denseColBlockIters", "sparseRowPanelIds", "sparseColBlockIters"]


def cases(big=False):
    """(name, pattern, K, alpha, delta, store_full)"""
    c = []
    S = gen.uniform_random(512, 512, 0.10, 7)
    c.append(("uni512_a03_d03", S, 32, 0.3, 0.3, True))
    c.append(("uni512_a03_d00", S, 32, 0.3, 0.0, True))
    c.append(("uni512_a09_d01", S, 64, 0.9, 0.1, True))
    S = gen.block_structured(512, 768, 6, 96, 0.8, seed=11, noise=0.004)
    c.append(("blocks512_a03_d03", S, 64, 0.3, 0.3, True))
    c.append(("blocks512_a05_d05", S, 64, 0.5, 0.5, True))
    c.append(("blocks512_a01_d11", S, 32, 0.1, 1.1, True))
    c.append(("blocks512_a07_d03", S, 64, 0.7, 0.3, True))
    c.append(("blocks512_a08_d03", S, 64, 0.8, 0.3, False))
    c.append(("blocks512shuf_a03_d03", gen.shuffle_within_rows(S, 3), 64, 0.3, 0.3, True))
    c.append(("blocks512empty_a03_d03", gen.with_empty_rows(S, 7), 64, 0.3, 0.3, True))
    S = gen.rmat(12, 8, 4)
    c.append(("rmat12_a03_d03", S, 32, 0.3, 0.3, False))
    c.append(("rmat12_a01_d01", S, 32, 0.1, 0.1, False))
    c.append(("rmat12_a07_d03", S, 128, 0.7, 0.3, False))
    S = gen.dlmc_magnitude_mask(1024, 1024, 0.7, 30)
    c.append(("dlmc1024s70_a03_d03", S, 64, 0.3, 0.3, False))
    c.append(("dlmc1024s70_a05_d07", S, 64, 0.5, 0.7, False))
    c.append(("dlmc1024s70_a08_d03", S, 64, 0.8, 0.3, False))
    S = gen.bernoulli_mask(1024, 1024, 0.9, 31)
    c.append(("bern1024s90_a03_d03", S, 256, 0.3, 0.3, False))
    # non-power-of-two warp counts in the clustering reduction: W=3, 5, 7
    S = gen.block_structured(600, 4800, 5, 400, 0.5, seed=21, noise=0.002)
    c.append(("w3_600x4800_a03_d03", S, 32, 0.3, 0.3, False))
    c.append(("w3_600x4800_a05_d03", S, 32, 0.5, 0.3, False))
    c.append(("w3_600x4800_a06_d03", S, 32, 0.6, 0.3, False))
    S = gen.zipf_docs(400, 9600, 60000, 22)
    c.append(("w5_zipf400x9600_a03_d03", S, 32, 0.3, 0.3, False))
    c.append(("w5_zipf400x9600_a04_d03", S, 32, 0.4, 0.3, False))
    c.append(("w5_zipf400x9600_a05_d03", S, 32, 0.5, 0.3, False))
    S = gen.zipf_docs(1500, 12419, 746316, 1)
    c.append(("nips_surrogate_a03_d03", S, 32, 0.3, 0.3, False))
    c.append(("nips_surrogate_a05_d01", S, 32, 0.5, 0.1, False))
    c.append(("nips_surrogate_a06_d03", S, 32, 0.6, 0.3, False))
    if big:
        S = gen.bernoulli_mask(4096, 4096, 0.7, 30)
        c.append(("bern4096s70_k64", S, 64, 0.3, 0.3, False))
        c.append(("bern4096s70_k256", S, 256, 0.3, 0.3, False))
        S = gen.dlmc_magnitude_mask(4096, 4096, 0.9, 33)
        c.append(("dlmc4096s90_k64", S, 64, 0.3, 0.3, False))
        c.append(("dlmc4096s90_k256", S, 256, 0.3, 0.3, False))
        S = gen.rmat(16, 16, 4)
        c.append(("rmat16_k32", S, 32, 0.3, 0.3, False))
        c.append(("rmat16_k128", S, 128, 0.3, 0.3, False))
    return c


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def run_case(name, S, K, alpha, delta, store_full, outdir, check_oracle=True):
    check_oracle = check_oracle and S.M <= 8192  # the literal oracle is O(#clusters * M * nbpr)
    A, B = gen.dense_operands(S.M, S.N, K)
    tmp = tempfile.mkdtemp(prefix="refdump_")
    case = os.path.join(tmp, "case.bin")
    gen.write_case_bin(case, S, A, B)
    t0 = time.time()
    p = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ref_dump"), case, repr(alpha), repr(delta), tmp],
                       capture_output=True, text=True, timeout=1500)
    wall = time.time() - t0
    rec = dict(name=name, M=S.M, N=S.N, nnz=S.nnz, K=K, alpha=alpha, delta=delta, wall_s=round(wall, 3),
               rc=p.returncode, stdout_tail=p.stdout.strip().splitlines()[-1:] if p.stdout else [],
               stderr_tail=p.stderr.strip().splitlines()[-3:] if p.stderr else [])
    if p.returncode != 0:
        return rec
    meta = {}
    for line in open(os.path.join(tmp, "meta.txt")):
        k, v = line.split()
        try:
            meta[k] = int(v)
        except ValueError:
            meta[k] = float(v)
    rec["meta"] = meta
    arrs = {k: np.fromfile(os.path.join(tmp, k + ".u32"), dtype=np.uint32) for k in U32}
    P = np.fromfile(os.path.join(tmp, "P.f32"), dtype=np.float32)
    Pcpu = np.fromfile(os.path.join(tmp, "P_cpu.f32"), dtype=np.float32)
    rec["sha"] = {k: sha(v) for k, v in arrs.items()}
    rec["sizes"] = {k: int(v.size) for k, v in arrs.items()}
    save = dict(reorderedRows=arrs["reorderedRows"], block_size=np.uint32(meta["block_size"]),
                num_clusters=np.int32(meta["num_clusters"]))
    if store_full:
        save.update(arrs)
        save["P"] = P
    np.savez_compressed(os.path.join(outdir, name + ".npz"), **save)
    # ---- oracle vs reference, on the box
    if check_oracle:
        bs = int(meta["block_size"])
        t0 = time.time()
        rr = O.row_reorder(S, alpha, bs)
        cr = O.col_reorder(S, arrs["reorderedRows"], delta)
        rp = O.rphm_build(S, arrs["reorderedRows"], cr)
        rec["oracle_s"] = round(time.time() - t0, 3)
        eq = {"reorderedRows": bool(np.array_equal(rr["reorderedRows"], arrs["reorderedRows"])),
              "numClusters": bool(rr["numClusters"] == meta["num_clusters"])}
        for k in ("denseCols", "denseColOffsets", "sparseCols", "sparseColOffsets", "sparseValueOffsets"):
            eq[k] = bool(np.array_equal(cr[k], arrs[k]))
        for k in ("blockOffsets", "blockValues", "sparseValues", "sparseRelativeRows", "sparseColIndices",
                  "denseRowPanelIds", "denseColBlockIters", "sparseRowPanelIds", "sparseColBlockIters"):
            eq[k] = bool(np.array_equal(rp[k], arrs[k]))
        Po = O.sddmm_cpu(S, A, B)
        eq["P_cpu_bitexact"] = bool(np.array_equal(Po, Pcpu))
        eq["P_gpu_vs_oracle_errors"] = O.check_data(Po, P)
        if not eq["reorderedRows"]:
            d = np.nonzero(rr["reorderedRows"] != arrs["reorderedRows"])[0] if rr["reorderedRows"].size == arrs[
                "reorderedRows"].size else None
            eq["first_diff"] = None if d is None or d.size == 0 else int(d[0])
        rec["oracle_equals_reference"] = eq
    for fn in os.listdir(tmp):
        os.remove(os.path.join(tmp, fn))
    os.rmdir(tmp)
    return rec


def main():
    big = "--big" in sys.argv
    outdir = os.path.join(ROOT, "gpurun_out", "ref_gpu")
    os.makedirs(outdir, exist_ok=True)
    recs = []
    for c in cases(big):
        rec = run_case(*c, outdir=outdir)
        recs.append(rec)
        print(json.dumps({k: rec[k] for k in rec if k not in ("sha", "sizes")}), flush=True)
        with open(os.path.join(outdir, "summary.json"), "w") as f:
            json.dump(recs, f, indent=1)


if __name__ == "__main__":
    main()
