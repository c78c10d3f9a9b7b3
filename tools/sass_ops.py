#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that show what a kernel is built from (cuobjdump -sass of the shipped
library): bulk copies (UBLKCP = cp.async.bulk, the TMA engine), st.async into peer CTAs (STAS), mbarrier operations
(SYNCS), cluster barriers (UCGABAR), register re-allocation (USETMAXREG), cp.async (LDGSTS), packed FP32 arithmetic
(FFMA2 / FADD2 / FMUL2) next to the scalar forms, tensor-core and tensor-map instructions (none by design).
usage: python tools/sass_ops.py [liquid-dsp_b200/libliquid_cuda.so] > profiles/rNN_sass_ops.txt"""
import collections
import os
import re
import subprocess
import sys

OPS = ["UBLKCP", "UTMALDG", "UTMASTG", "STAS", "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT", "USETMAXREG", "LDGSTS", "FFMA2", "FADD2",
       "FMUL2", "FFMA", "FADD", "FMUL", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "HMMA", "UTCHMMA"]


def main():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "liquid-dsp_b200", "libliquid_cuda.so")
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    rows = []
    for fn in re.split(r"\n\s+Function : ", txt)[1:]:
        name = fn.split("\n")[0].strip()
        dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
        dem = re.sub(r"\(anonymous namespace\)::", "", dem)
        cut = dem.rfind(">(")
        dem = dem[:cut + 1] if cut >= 0 else re.sub(r"\(.*$", "", dem)
        dem = dem.replace("(int)", "").replace("(bool)", "")
        c = collections.Counter()
        n = 0
        for m in re.finditer(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", fn):
            n += 1
            c[m.group(1)] += 1
        rows.append((dem, n, c))
    print("# %s: %d kernels (sm_100a)" % (os.path.basename(lib), len(rows)))
    print("# columns: instructions, then " + " ".join(OPS))
    tot = collections.Counter()
    for dem, n, c in sorted(rows):
        print("%-100s %6d  %s" % (dem[:100], n, " ".join("%s:%d" % (o, c[o]) for o in OPS if c[o])))
        tot.update({o: c[o] for o in OPS})
    print("# totals: " + " ".join("%s:%d" % (o, tot[o]) for o in OPS))


if __name__ == "__main__":
    main()
