"""Per-kernel SASS opcode histogram of libsddmm_b200.so (evidence that the hot kernels are Blackwell-native:
UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA, UTCBAR = tcgen05.commit, .2CTA = the cta_group::2 forms, FHFMA = fma.rn.f32.f16; B200_PROFILING.md table).

    python tools/sass_histogram.py > profiles/sass_r02.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sddmm-gpu_b200", "libsddmm_b200.so")
KEYS = ["UTCHMMA", "2CTA", "UTCBAR", "LDTM", "UTMALDG", "UTMALDG.2D.GATHER4", "MULTICAST", "SYNCS", "HMMA", "FFMA", "FHFMA",
        "LDG", "LDS", "STS", "STG", "SHFL", "POPC", "REDUX", "ATOM"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    demangle = lambda s: subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip()
    cur, hist = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = demangle(m.group(1))
            cur = re.sub(r"\(.*", "", cur).replace("void sb::", "").replace("(anonymous namespace)::", "")
            hist[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1)
            hist[cur]["total"] += 1
            for k in KEYS:
                if op.startswith(k) or (k == "MULTICAST" and "MULTICAST" in op) or (k == "2CTA" and ".2CTA" in op) or (k == "UTMALDG.2D.GATHER4" and "GATHER4" in op):
                    hist[cur][k] += 1
    print(f"# SASS opcode counts per kernel, {os.path.relpath(LIB, ROOT)} (cuobjdump -sass, sm_100a)")
    print("# " + " ".join(f"{k:>8}" for k in ["total"] + KEYS) + "  kernel")
    for name, h in hist.items():
        print("  " + " ".join(f"{h.get(k, 0):>8}" for k in ["total"] + KEYS) + "  " + name)
    tot = collections.Counter()
    for h in hist.values():
        tot.update(h)
    print("# library totals: " + ", ".join(f"{k}={tot.get(k, 0)}" for k in KEYS))


if __name__ == "__main__":
    sys.exit(main())
