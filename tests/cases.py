"""Shared seeded parity cases (small enough for the literal oracle to finish in seconds)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

pkg = load_package()
gen = pkg.generators

GOLDEN = os.path.join(ROOT, "tests", "golden")


def small_cases():
    """name -> (S, alpha, delta, K)"""
    c = {}
    S = gen.uniform_random(512, 512, 0.10, 7)
    c["uni512_a03_d03"] = (S, 0.3, 0.3, 32)
    c["uni512_a09_d01"] = (S, 0.9, 0.1, 64)
    S = gen.block_structured(512, 768, 6, 96, 0.8, seed=11, noise=0.004)
    c["blocks512_a03_d03"] = (S, 0.3, 0.3, 64)
    c["blocks512_a07_d03"] = (S, 0.7, 0.3, 64)
    c["blocks512_a08_d00"] = (S, 0.8, 0.0, 32)
    c["blocks512_a01_d11"] = (S, 0.1, 1.1, 32)
    c["blocks512shuf_a03_d03"] = (gen.shuffle_within_rows(S, 3), 0.3, 0.3, 64)
    c["blocks512empty_a05_d05"] = (gen.with_empty_rows(S, 7), 0.5, 0.5, 128)
    S = gen.rmat(12, 8, 4)
    c["rmat12_a03_d03"] = (S, 0.3, 0.3, 32)
    c["rmat12_a01_d01"] = (S, 0.1, 0.1, 64)
    S = gen.dlmc_magnitude_mask(1024, 1024, 0.7, 30)
    c["dlmc1024s70_a03_d03"] = (S, 0.3, 0.3, 64)
    c["dlmc1024s70_a08_d03"] = (S, 0.8, 0.3, 256)
    S = gen.block_structured(600, 4800, 5, 400, 0.5, seed=21, noise=0.002)
    c["w3_600x4800_a05_d03"] = (S, 0.5, 0.3, 32)   # clustering blockDim 96 = 3 warps (warps dropped)
    S = gen.zipf_docs(400, 9600, 60000, 22)
    c["w5_zipf400x9600_a04_d03"] = (S, 0.4, 0.3, 32)  # 5 warps
    S = gen.zipf_docs(300, 12419, 40000, 5)
    c["w7_zipf300x12419_a03_d01"] = (S, 0.3, 0.1, 32)  # 7 warps (the nips shape's nbpr=777)
    # ragged: 37 rows (last panel partial), tiny N
    S = gen.uniform_random(37, 50, 0.3, 9)
    c["ragged37x50_a03_d03"] = (S, 0.3, 0.3, 32)
    return c


def operands(S, K):
    return gen.dense_operands(S.M, S.N, K)
