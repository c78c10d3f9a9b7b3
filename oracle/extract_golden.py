#!/usr/bin/env python
"""Extract the reference's own golden vectors / known-answer tests for the channelizer path
into tests/golden/ref_kat.json.  TEST INFRASTRUCTURE.

Reads (never copies source from) /root/reference:
  src/filter/tests/data/firfilt_crcf_data_h{4x8,7x16,13x32,23x64}.c   h / x / y literal arrays
  src/fft/tests/data/fft_data_<N>.c                                   x / y literal arrays
  src/dotprod/tests/dotprod_crcf_autotest.c                           rand01 / rand02 KATs
  src/filter/tests/firpfb_autotest.c                                  48-tap / 12-sample KAT
  src/filter/tests/firdes_autotest.c                                  liquid_firdes_rrcos(2,3,0.3) taps
  src/math/tests/math_bessel_autotest.c                               liquid_besseli0f points
Only numbers are taken (the numeric test vectors), together with the tolerance each reference
test uses.  Run in the dev container (where /root/reference exists); the JSON is committed.
"""
import glob
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ref_kat.json")

NUM = r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?"


def strip_comments(s):
    s = re.sub(r"/\*.*?\*/", "", s, flags=re.S)
    return re.sub(r"//[^\n]*", "", s)


def arrays(path):
    """all `float [complex] name[...] = { ... };` initialisers of a C file -> {name: list}"""
    src = strip_comments(open(path).read())
    out = {}
    for m in re.finditer(r"float\s+(complex\s+)?(\w+)\s*\[[^\]]*\]\s*=\s*\{(.*?)\}\s*;", src, flags=re.S):
        is_c, name, body = bool(m.group(1)), m.group(2), m.group(3)
        if is_c:
            vals = []
            for item in body.split(","):
                item = item.strip()
                if not item:
                    continue
                mm = re.match(r"^(%s)\s*\+\s*(%s)\s*\*\s*_Complex_I$" % (NUM, NUM), item)
                if not mm:
                    raise ValueError("cannot parse complex literal %r in %s" % (item, path))
                vals.append([float(mm.group(1)), float(mm.group(2))])
        else:
            vals = [float(v.rstrip("f")) for v in re.findall(NUM + "f?", body)]
        out[name] = vals
    return out


def scalars_complex(path, name):
    src = strip_comments(open(path).read())
    res = []
    for m in re.finditer(r"float\s+complex\s+%s\s*=\s*(%s)\s*([-+])\s*(%s)\s*\*\s*_Complex_I" % (name, NUM, NUM), src):
        im = float(m.group(3)) * (-1 if m.group(2) == "-" else 1)
        res.append([float(m.group(1)), im])
    return res


kat = {"source": "liquid-dsp reference tests (numeric vectors only); see oracle/extract_golden.py"}

# firfilt_crcf golden vectors; tolerance 1e-3 (firfilt_runtest.c:75)
kat["firfilt_crcf"] = {"tol": 1e-3, "cases": {}}
for tag in ("h4x8", "h7x16", "h13x32", "h23x64"):
    a = arrays(os.path.join(REF, "src/filter/tests/data/firfilt_crcf_data_%s.c" % tag))
    kat["firfilt_crcf"]["cases"][tag] = {k.split("_")[-1]: v for k, v in a.items()}

# fft golden vectors; tolerance 2e-4 abs (fft_runtest.c:35)
kat["fft"] = {"tol": 2e-4, "cases": {}}
for path in sorted(glob.glob(os.path.join(REF, "src/fft/tests/data/fft_data_*.c"))):
    n = int(re.search(r"fft_data_(\d+)\.c", path).group(1))
    a = arrays(path)
    kat["fft"]["cases"][str(n)] = {"x": a["fft_test_x%d" % n], "y": a["fft_test_y%d" % n]}

# dotprod_crcf KATs; tolerance 1e-3 (dotprod_crcf_autotest.c:50)
path = os.path.join(REF, "src/dotprod/tests/dotprod_crcf_autotest.c")
src = strip_comments(open(path).read())
tests = scalars_complex(path, "test")
blocks = re.split(r"void\s+autotest_dotprod_crcf_rand0[12]\s*\(\)", src)[1:3]
kat["dotprod_crcf"] = {"tol": 1e-3, "cases": []}
for blk, t in zip(blocks, tests):
    tmp = "/tmp/_dp_blk.c"
    open(tmp, "w").write(blk)
    a = arrays(tmp)
    kat["dotprod_crcf"]["cases"].append({"h": a["h"], "x": a["x"], "y": t})

# firpfb KAT (rrrf in the reference; the crcf object reduces to it for real input); tol 1e-4
a = arrays(os.path.join(REF, "src/filter/tests/firpfb_autotest.c"))
kat["firpfb"] = {"tol": 1e-4, "M": 4, "h": a["h"], "x": a["noise"], "y": a["test"]}

# liquid_firdes_rrcos(k=2, m=3, beta=0.3, dt=0) taps; tol 1e-5 (firdes_autotest.c:59-90)
src = strip_comments(open(os.path.join(REF, "src/filter/tests/firdes_autotest.c")).read())
blk = src[src.index("autotest_liquid_firdes_rrcos"):]
open("/tmp/_rrc_blk.c", "w").write(blk[:blk.index("autotest_liquid_firdes_rkaiser")])
kat["firdes_rrcos"] = {"tol": 1e-5, "k": 2, "m": 3, "beta": 0.3, "h": arrays("/tmp/_rrc_blk.c")["h0"]}

# liquid_besseli0f points; tol 1e-3 (math_bessel_autotest.c:96-106)
src = strip_comments(open(os.path.join(REF, "src/math/tests/math_bessel_autotest.c")).read())
pts = re.findall(r"liquid_besseli0f\((%s)f?\)\s*,\s*(%s)f?" % (NUM, NUM), src)
kat["besseli0f"] = {"tol": 1e-3, "points": [[float(a), float(b)] for a, b in pts]}

with open(OUT, "w") as f:
    json.dump(kat, f, indent=0, separators=(",", ":"))
print("wrote", os.path.normpath(OUT), {k: (len(v.get("cases", v)) if isinstance(v, dict) else 1) for k, v in kat.items()})
