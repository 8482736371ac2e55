python - <<'PY'
import sys, json, os, time
sys.path.insert(0, '.')
import numpy as np, torch
from __graft_entry__ import load_package
pkg = load_package(); gen = pkg.generators
cases = [("nips", gen.zipf_docs(1500, 12419, 746316, 1)), ("bern4096s70", gen.bernoulli_mask(4096, 4096, 0.7, 30)),
         ("dlmc4096s90", gen.dlmc_magnitude_mask(4096, 4096, 0.9, 33)), ("rmat16", gen.rmat(16, 16, 4)),
         ("uni20k", gen.uniform_random(20000, 20000, 0.01, 2)), ("rmat18", gen.rmat(18, 16, 4))]
for name, S in cases:
    ro = torch.from_numpy(S.row_off.view(np.int32)).cuda(); ci = torch.from_numpy(S.col_idx.view(np.int32)).cuda()
    best = 1e9
    for rep in range(3):
        R, ncl, ms = pkg.row_reorder_dev(ro, ci, S.M, S.N, 0.3, 16 if S.N <= 98304 else 0)
        best = min(best, ms)
    lay, cms, rms = pkg.layout_build_dev(ro, ci, S.M, S.N, R, 0.3)
    lay, cms, rms = pkg.layout_build_dev(ro, ci, S.M, S.N, R, 0.3)
    print(name, "rows", S.M, "clusters", ncl, "row_reorder best ms %.2f" % best, "col %.2f rphm %.2f" % (cms, rms), flush=True)
PY
