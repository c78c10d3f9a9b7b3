#!/usr/bin/env python
"""Compare the prototypes of include/liquid_cuda.h with the reference's include/liquid.h (after preprocessing, so the
reference's API-generating macros are expanded): every function this library exports under a reference name must
take the same parameters.  Enums of the reference are passed as int here (same ABI), `float complex` is
liquid_float_complex in both.  Needs /root/reference; run in the build container:  python tools/check_signatures.py"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("LIQUID_REF", "/root/reference")


def prototypes(header, incs):
    src = subprocess.check_output(["gcc", "-E", "-P", "-x", "c", "-std=c99"] + sum((["-I", i] for i in incs), []) + [header]).decode()
    src = re.sub(r"\s+", " ", src)
    out = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b([a-z][A-Za-z0-9_]*)\s*\(([^()]*)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if ret.startswith("typedef") or not ret:
            continue
        out[name] = (ret, args)
    return out


ENUMS = ("liquid_firfilt_type", "liquid_resamp_type", "liquid_firdespm_btype", "liquid_firdespm_wtype")


def norm(t):
    t = re.sub(r"\b_Complex float\b|\bfloat _Complex\b|\bfloat complex\b", "liquid_float_complex", t)
    for e in ENUMS:
        t = re.sub(r"\b%s\b" % e, "int", t)
    t = re.sub(r"\bconst\b", "", t)
    t = re.sub(r"\b_[A-Za-z0-9_]+\b", "", t)          # parameter names
    t = re.sub(r"\bstruct ([a-z0-9_]+)_s \*", r"\1", t)   # opaque handles
    t = t.replace("unsigned int", "unsigned")
    return re.sub(r"\s+", "", t)


def main():
    ours = prototypes(os.path.join(ROOT, "include", "liquid_cuda.h"), [])
    ref = prototypes(os.path.join(REF, "include", "liquid.h"), [os.path.join(ROOT, "oracle", "refcfg")])
    shared = sorted(set(ours) & set(ref))
    bad = []
    for name in shared:
        a, b = ours[name], ref[name]
        if norm(a[0]) != norm(b[0]) or norm(a[1]) != norm(b[1]):
            bad.append((name, a, b))
    print("%d functions exported under a reference name, %d with differing prototypes" % (len(shared), len(bad)))
    for name, a, b in bad:
        print("  %s\n    here: %s (%s)\n    ref : %s (%s)" % (name, a[0], a[1], b[0], b[1]))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
