"""CPU-only checks of the product library: it builds for sm_100a, loads, exports every symbol
include/liquid_cuda.h declares, and its host-side prototype design equals the reference's.
No compute calls that need a GPU."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="session")
def L():
    import __graft_entry__ as g
    g.build()
    import liquid_dsp_b200
    return liquid_dsp_b200


def _header_functions():
    src = open(os.path.join(ROOT, "include", "liquid_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)            # preprocessor lines
    src = re.sub(r"typedef\s+[^;]*\(\s*\*\s*\w+\s*\)\s*\([^;]*\)\s*;", "", src)   # function-pointer typedefs (callbacks)
    names = set()
    for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", src):
        names.add(m.group(1))
    return names


def test_library_exports_every_declared_symbol(L):
    declared = _header_functions()
    assert len(declared) > 80
    out = subprocess.check_output(["nm", "-D", "--defined-only", L.LIB_PATH]).decode()
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = sorted(declared - exported)
    assert not missing, missing
    # and the Python mirror binds exactly that set
    assert set(L.DECLARED_SYMBOLS) == declared


def test_library_carries_sm_100a_code(L):
    out = subprocess.check_output(["/usr/local/cuda/bin/cuobjdump", "-lelf", L.LIB_PATH]).decode()
    assert "sm_100a" in out


def test_product_does_not_reference_the_oracle():
    """the product path must never route through oracle/ (or any CPU fallback)"""
    pkg = os.path.join(ROOT, "liquid-dsp_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".cu", ".cuh", ".py", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in text and "oracle." not in text and "from oracle" not in text, f


def test_host_design_kaiser_matches_reference(L, refout):
    for name in [k for k in refout if k.startswith("kaiser_") and not k.endswith("_args")]:
        n, fc, As = refout[name + "_args"]
        h = L.liquid_firdes_kaiser(int(n), np.float32(fc), np.float32(As))
        assert np.abs(h - refout[name]).max() <= 5e-7, name


def test_host_design_kaiser_equals_oracle_bit_for_bit(L, orc):
    for n, fc, As in [(129, 1 / 32, 60.0), (14337, 1 / 1024, 60.0), (97, 0.1, 40.0), (33, 0.2, 15.0), (73729, 1 / 4096, 60.0)]:
        a = L.liquid_firdes_kaiser(n, np.float32(fc), As)
        b = orc.firdes_kaiser(n, np.float32(fc), As)
        assert np.array_equal(a, b), (n, np.abs(a - b).max())


def test_host_design_rrcos(L, orc, kat, refout):
    k = kat["firdes_rrcos"]
    h = L.liquid_firdes_rrcos(k["k"], k["m"], k["beta"])
    assert np.abs(h - np.array(k["h"])).max() < k["tol"]
    for name in [n for n in refout if n.startswith("rrcos_") and not n.endswith("_args")]:
        kk, m, beta = refout[name + "_args"]
        a = L.liquid_firdes_rrcos(int(kk), int(m), np.float32(beta))
        assert np.abs(a - refout[name]).max() <= 1e-6
        assert np.array_equal(a, orc.firdes_rrcos(int(kk), int(m), np.float32(beta)))


def test_kaiser_beta(L, orc):
    for As in [10.0, 21.0, 35.0, 50.0, 60.0, 80.0, -60.0]:
        assert L.kaiser_beta_As(As) == orc.lib().orc_kaiser_beta_As(As)


def test_fft_shift_host(L):
    for n in (8, 9):
        x = np.arange(n).astype(np.complex64)
        y = x.copy()
        L.lib.fft_shift(y, n)
        n2 = n // 2
        want = x.copy()
        want[:n2], want[n2:2 * n2] = x[n2:2 * n2].copy(), x[:n2].copy()
        assert np.array_equal(y, want)


def test_host_design_rnyquist_kaiser(L, orc, refrn):
    """liquid_firdes_arkaiser / _rkaiser / _prototype(KAISER) of the library (host code, no GPU needed)
    equal the reference fixture and the oracle bit for bit"""
    for name in sorted(k for k in refrn if k.startswith("args_")):
        tag = name[5:]
        k, m, beta, dt = refrn[name]
        k, m, beta, dt = int(k), int(m), np.float32(beta), np.float32(dt)
        assert np.array_equal(L.liquid_firdes_arkaiser(k, m, beta, dt), refrn["taps_arkaiser_" + tag]), tag
        assert np.array_equal(L.liquid_firdes_rkaiser(k, m, beta, dt), refrn["taps_rkaiser_" + tag]), tag
        assert np.array_equal(L.liquid_firdes_prototype(L.LIQUID_FIRFILT_KAISER, k, m, beta, dt), refrn["taps_nkaiser_" + tag]), tag
        assert np.array_equal(L.liquid_firdes_prototype(L.LIQUID_FIRFILT_RKAISER, k, m, beta, dt),
                              orc.firdes_rkaiser(k, m, beta, dt))
        rms, peak = L.liquid_filter_isi(refrn["taps_rkaiser_" + tag], k, m)
        assert np.allclose([rms, peak], refrn["isi_rkaiser_" + tag], rtol=1e-6)
    for df, N, want in refrn["req_As"]:
        assert L.lib.estimate_req_filter_As(df, int(N)) == np.float32(want)


def test_estimate_req_filter_len(L, refrn2):
    """Kaiser's length estimate (host code) at the fixture's points from the compiled reference"""
    for (df, As), want in zip(refrn2["len_points"], refrn2["len_values"]):
        assert L.lib.estimate_req_filter_len(float(df), float(As)) == want


BAD_CALLS = [
    # (python expression on the ctypes mirror `L`, function named in the message)   reference file:line of the check
    ("L.lib.firpfbch2_crcf_create_kaiser(0, 7, 3, 60.0)", "firpfbch2_crcf_create_kaiser"),        # firpfbch2.c:140-146 (odd M)
    ("L.lib.firpfbch2_crcf_create_kaiser(0, 8, 0, 60.0)", "firpfbch2_crcf_create_kaiser"),        # :147-150 (m = 0)
    ("L.lib.firpfbch_crcf_create_kaiser(0, 0, 3, 60.0)", "firpfbch_crcf_create_kaiser"),          # firpfbch.c:156-159
    ("L.lib.firdecim_crcf_create_kaiser(1, 3, 60.0)", "decim_crcf_create_kaiser"),                # firdecim.c:93-96
    ("L.lib.firinterp_crcf_create_kaiser(1, 3, 60.0)", "firinterp_crcf_create_kaiser"),           # firinterp.c:91-94
    ("L.lib.resamp2_crcf_create(1, 0.0, 60.0)", "resamp2_crcf_create"),                           # resamp2.c:70-73
    ("L.lib.resamp2_crcf_create(4, 0.75, 60.0)", "resamp2_crcf_create"),                          # resamp2.c:79-82
    ("L.lib.msresamp2_crcf_create(1, 17, 0.25, 0.0, 60.0)", "msresamp2_crcf_create"),             # msresamp2.c:89-92
    ("L.lib.msresamp2_crcf_create(1, 3, 0.5, 0.0, 60.0)", "msresamp2_crcf_create"),               # msresamp2.c:95-98
    ("L.lib.fftfilt_crcf_create(np.ones(9, np.float32), 9, 4)", "fftfilt_crcf_create"),           # fftfilt.c:77-83
    ("L.lib.asgramcf_create(1)", "asgramcf_create"),                                              # asgram.c:51-54
    ("L.lib.estimate_req_filter_len(0.7, 60.0)", "estimate_req_filter_len"),                      # firdes.c:55-58
    ("L.lib.spgramcf_create_kaiser(1, 1, 3.0)", "spgramcf_create"),                               # spgram.c:63-66
    ("L.lib.ofdmframegen_create(63, 8, 2, None)", "ofdmframegen_create"),                         # ofdmframegen.c:94-96
    ("L.lib.ofdmframegen_create(64, 8, 9, None)", "ofdmframegen_create"),                         # ofdmframegen.c:100-102
    ("L.lib.ofdmframegen_create(8, 2, 1, np.zeros(8, np.uint8).ctypes.data)", "ofdmframegen_create"),   # :121-123
]


@pytest.mark.parametrize("expr,fn", BAD_CALLS)
def test_invalid_arguments_follow_the_reference_error_convention(expr, fn):
    """the reference prints `error: <fn>(), <message>` to stderr and exits with status 1 (e.g. firpfbch2.c:72-81);
    the argument checks come before any device work, so this runs without a GPU"""
    code = ("import sys, importlib, numpy as np; sys.path.insert(0, %r); "
            "L = importlib.import_module('liquid-dsp_b200'); %s" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), expr))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 1, (expr, r.returncode, r.stderr[-300:])
    assert ("error: %s" % fn) in r.stderr, (expr, r.stderr[-300:])


def test_host_design_parks_mcclellan_and_hM3_equal_the_reference_bit_for_bit(L, refrn2):
    """firdespm_run (band-pass type: low-pass with weights, even length, exponential / linear weights, several bands),
    liquid_firdes_prototype(PM) and liquid_firdes_hM3 (up to 101 equiripple designs each, picked by inter-symbol
    interference) against taps of the compiled reference: the designer runs in double precision with the reference's
    operation order, so the float32 taps are identical"""
    for k, m, beta in refrn2["pm_cases"]:
        k, m = int(k), int(m)
        tag = "k%d_m%d_b%03d" % (k, m, round(beta * 100))
        assert np.array_equal(L.liquid_firdes_hM3(k, m, np.float32(beta)), refrn2["taps_hM3_" + tag]), tag
        assert np.array_equal(L.liquid_firdes_prototype(L.LIQUID_FIRFILT_hM3, k, m, np.float32(beta)), refrn2["taps_hM3_" + tag]), tag
        assert np.array_equal(L.liquid_firdes_prototype(L.LIQUID_FIRFILT_PM, k, m, np.float32(beta)), refrn2["taps_pm_" + tag]), tag
    for name in sorted(k for k in refrn2 if k.startswith("pmrun_") and k.endswith("_h")):
        tag = name[:-2]
        want = refrn2[name]
        got = L.firdespm_run(len(want), refrn2[tag + "_bands"], refrn2[tag + "_des"], refrn2[tag + "_weights"], refrn2[tag + "_wtype"])
        assert np.array_equal(got, want), tag
    # the object form gives the same taps
    tag = "pmrun_bp77"
    wt = np.ascontiguousarray(refrn2[tag + "_wtype"], np.int32)
    import ctypes as C
    q = L.lib.firdespm_create(77, 3, refrn2[tag + "_bands"], refrn2[tag + "_des"], refrn2[tag + "_weights"],
                              wt.ctypes.data_as(C.POINTER(C.c_int)), L.LIQUID_FIRDESPM_BANDPASS)
    h = np.empty(77, np.float32)
    L.lib.firdespm_execute(q, h)
    L.lib.firdespm_destroy(q)
    assert np.array_equal(h, refrn2[tag + "_h"])
    # what an equiripple low-pass is: pass band near 1, stop band small
    h = L.firdespm_run(51, [0.0, 0.2, 0.25, 0.5], [1.0, 0.0], [1.0, 10.0])
    H = np.abs(np.fft.fft(h, 1024))[:513]
    f = np.arange(513) / 1024.0
    assert np.abs(H[f <= 0.2] - 1).max() < 0.1 and H[f >= 0.25].max() < 0.01


def test_host_design_remaining_prototype_families(L, refrn2):
    """liquid_firdes_prototype for RCOS, FEXP / FSECH / FARCSECH and their root versions, GMSKTX / GMSKRX against the
    compiled reference's taps: the closed forms bit for bit, the frequency-sampled ones to the rounding of the
    reference's float FFT (the receive GMSK design divides two such responses: 2e-5)"""
    tol = {3: 0.0, 11: 0.0, 4: 2e-6, 5: 2e-6, 6: 2e-6, 13: 2e-6, 14: 2e-6, 15: 2e-6, 12: 2e-5}
    for ftype, t in tol.items():
        for k, m, beta, dt in refrn2["proto_cases"]:
            k, m = int(k), int(m)
            want = refrn2["taps_proto%d_k%d_m%d" % (ftype, k, m)]
            got = L.liquid_firdes_prototype(ftype, k, m, np.float32(beta), np.float32(dt))
            if t == 0.0:
                assert np.array_equal(got, want), (ftype, k, m)
            else:
                assert np.abs(got - want).max() <= t * np.abs(want).max(), (ftype, k, m)


def test_c_program_builds_and_links_against_the_drop_in(L, tmp_path):
    """the boundary is a C ABI: include/liquid_cuda.h compiles as C99 and as C++11, and a plain C program that calls
    the reference's own function names links against libliquid_cuda.so and runs (host-side design calls: no GPU)"""
    inc = os.path.join(ROOT, "include")
    libdir = os.path.dirname(L.LIB_PATH)
    src = tmp_path / "app.c"
    src.write_text(
        '#include <stdio.h>\n#include <liquid_cuda.h>\n'
        'int main(void) {\n'
        '    float h[57];\n'
        '    liquid_firdes_kaiser(57, 0.2f, 60.0f, 0.0f, h);\n'
        '    float g[2 * 4 * 3 + 1];\n'
        '    liquid_firdes_prototype(LIQUID_FIRFILT_hM3, 4, 3, 0.3f, 0.0f, g);\n'
        '    printf("%.8e %u %.8e\\n", h[28], estimate_req_filter_len(0.1f, 60.0f), g[12]);\n'
        '    return 0;\n}\n')
    exe = tmp_path / "app"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", inc, str(src), "-o", str(exe), "-L", libdir,
                           "-lliquid_cuda", "-lm", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-300:]
    centre, n_len, g12 = out.stdout.split()
    assert abs(float(centre) - float(L.liquid_firdes_kaiser(57, 0.2, 60.0)[28])) < 1e-9 and int(n_len) == 36
    assert abs(float(g12) - float(L.liquid_firdes_hM3(4, 3, 0.3)[12])) < 1e-7
    cpp = tmp_path / "hdr.cpp"
    cpp.write_text('#include <liquid_cuda.h>\nint main() { return 0; }\n')
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Werror", "-I", inc, "-fsyntax-only", str(cpp)])


@pytest.mark.parametrize("name,symbol", [("firdes_kaiser_example", "liquid_firdes_kaiser"), ("firdespm_example", "firdespm_run")])
def test_reference_design_example_unmodified_on_the_drop_in(L, tmp_path, name, symbol):
    """the reference's own filter-design example programs (compiled unmodified against the compiled reference by
    oracle/Makefile) run as they are and with libliquid_cuda.so preloaded: the dynamic linker binds the design call
    to this library (LD_DEBUG=bindings) and the programs write the same numbers -- host code, no GPU involved"""
    exe = os.path.join(ROOT, "oracle", "_ref", "examples", name)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/examples not built (needs /root/reference at build time)")
    outs = []
    for tag, preload in (("ref", None), ("cuda", L.LIB_PATH)):
        d = tmp_path / tag
        d.mkdir()
        env = dict(os.environ)
        env["LD_DEBUG"] = "bindings"
        if preload:
            env["LD_PRELOAD"] = preload
        r = subprocess.run([exe], cwd=str(d), env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, (tag, r.stderr[-300:])
        bound_here = any("libliquid_cuda.so" in ln.split(" to ")[-1] and ("`%s'" % symbol) in ln
                         for ln in r.stderr.splitlines() if "binding file" in ln and " to " in ln)
        assert bound_here == bool(preload), (tag, symbol)
        text = (d / (name + ".m")).read_text()
        outs.append(np.array([float(v) for v in re.findall(r"[-+]?\d+\.\d+(?:[eE][-+]?\d+)?", text)]))
    ref, got = outs
    assert ref.shape == got.shape and ref.size > 10 and np.abs(ref - got).max() <= 1e-6 * max(1.0, np.abs(ref).max())


def test_exported_prototypes_equal_the_reference_header():
    """every function exported under a reference name takes exactly the reference's parameters: both headers are
    preprocessed (the reference generates its API with macros) and compared type by type -- needs /root/reference,
    so this runs in the build container only (tools/check_signatures.py)"""
    if not os.path.isdir("/root/reference/include"):
        pytest.skip("/root/reference is not present on this machine")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "check_signatures.py")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout[-2000:]
    n = int(r.stdout.split()[0])
    assert n >= 150, r.stdout


_F = "C.c_float(%r)"
ERROR_TEXT_CALLS = [
    ("resamp2_crcf_create", "(4, C.c_float(0.75), C.c_float(60.0))"), ("estimate_req_filter_len", "(C.c_float(0.7), C.c_float(60.0))"),
    ("estimate_req_filter_len", "(C.c_float(0.1), C.c_float(-1.0))"), ("liquid_firdes_hM3", "(1, 3, C.c_float(0.3), C.c_float(0.0), None)"),
    ("liquid_firdes_rcos", "(0, 3, C.c_float(0.3), C.c_float(0.0), None)"), ("liquid_firdes_fexp", "(2, 0, C.c_float(0.3), C.c_float(0.0), None)"),
    ("liquid_firdes_gmsktx", "(2, 3, C.c_float(1.3), C.c_float(0.0), None)"), ("liquid_firdes_gmskrx", "(0, 3, C.c_float(0.3), C.c_float(0.0), None)"),
    ("liquid_firdes_kaiser", "(0, C.c_float(0.2), C.c_float(60.0), C.c_float(0.0), None)"),
    ("liquid_firdes_kaiser", "(11, C.c_float(0.7), C.c_float(60.0), C.c_float(0.0), None)"),
    ("liquid_firdes_kaiser", "(11, C.c_float(0.2), C.c_float(60.0), C.c_float(0.9), None)"),
    ("liquid_firdes_arkaiser", "(1, 3, C.c_float(0.3), C.c_float(0.0), None)"), ("liquid_firdes_rkaiser", "(2, 0, C.c_float(0.3), C.c_float(0.0), None)"),
    ("liquid_firdes_rkaiser", "(2, 3, C.c_float(1.3), C.c_float(0.0), None)"), ("liquid_firdes_arkaiser", "(2, 3, C.c_float(0.3), C.c_float(2.0), None)"),
    ("liquid_firdes_rrcos", "(2, 3, C.c_float(1.3), C.c_float(0.0), None)"), ("liquid_firdes_prototype", "(77, 2, 3, C.c_float(0.3), C.c_float(0.0), None)"),
    ("firfilt_crcf_create", "(None, 0)"), ("firfilt_crcf_create_rnyquist", "(9, 1, 3, C.c_float(0.3), C.c_float(0.0))"),
    ("firfilt_crcf_create_rect", "(0,)"), ("firfilt_crcf_create_rect", "(2000,)"), ("firfilt_crcf_create_kaiser", "(0, C.c_float(0.2), C.c_float(60.0), C.c_float(0.0))"),
    ("firpfbch_crcf_create", "(5, 4, 3, None)"), ("firpfbch_crcf_create", "(0, 0, 3, None)"), ("firpfbch_crcf_create", "(0, 4, 0, None)"),
    ("firpfbch_crcf_create_kaiser", "(0, 4, 0, C.c_float(60.0))"), ("firpfbch_crcf_create_kaiser", "(0, 0, 3, C.c_float(60.0))"),
    ("firpfbch_crcf_create_rnyquist", "(0, 4, 3, C.c_float(0.3), 99)"),
    ("firpfbch2_crcf_create", "(0, 7, 3, None)"), ("firpfbch2_crcf_create", "(0, 8, 0, None)"),
    ("firpfbch2_crcf_create_kaiser", "(0, 7, 3, C.c_float(60.0))"), ("firpfbch2_crcf_create_kaiser", "(0, 8, 0, C.c_float(60.0))"),
    ("firpfb_crcf_create", "(0, None, 10)"), ("firpfb_crcf_create", "(4, None, 0)"), ("firpfb_crcf_create_kaiser", "(0, 3, C.c_float(0.3), C.c_float(60.0))"),
    ("firpfb_crcf_create_rnyquist", "(9, 0, 2, 3, C.c_float(0.3))"), ("firpfb_crcf_create_rnyquist", "(9, 4, 1, 3, C.c_float(0.3))"),
    ("firdecim_crcf_create", "(0, None, 10)"), ("firdecim_crcf_create", "(4, None, 0)"), ("firdecim_crcf_create_kaiser", "(4, 0, C.c_float(60.0))"),
    ("firdecim_crcf_create_kaiser", "(1, 3, C.c_float(60.0))"), ("firdecim_crcf_create_prototype", "(9, 1, 3, C.c_float(0.3), C.c_float(0.0))"),
    ("firinterp_crcf_create", "(1, None, 10)"), ("firinterp_crcf_create", "(4, None, 2)"), ("firinterp_crcf_create_kaiser", "(4, 0, C.c_float(60.0))"),
    ("firinterp_crcf_create_kaiser", "(1, 3, C.c_float(60.0))"), ("firinterp_crcf_create_prototype", "(9, 4, 0, C.c_float(0.3), C.c_float(0.0))"),
    ("fftfilt_crcf_create", "(None, 0, 8)"), ("spgramcf_create", "(64, None, 65)"), ("spgramcf_create", "(64, None, 0)"),
    ("spgramcf_create_kaiser", "(64, 65, C.c_float(3.0))"), ("spgramcf_create_kaiser", "(1, 1, C.c_float(3.0))"), ("spgramcf_create_default", "(1,)"),
    ("ofdmframegen_create", "(1, 0, 0, None)"), ("ofdmframegen_create", "(63, 8, 2, None)"), ("ofdmframegen_create", "(64, 65, 2, None)"),
    ("ofdmframegen_create", "(64, 8, 9, None)"), ("msresamp2_crcf_create", "(1, 3, C.c_float(0.0), C.c_float(0.0), C.c_float(60.0))"),
    ("msresamp2_crcf_create", "(1, 17, C.c_float(0.25), C.c_float(0.0), C.c_float(60.0))"), ("resamp2_crcf_create", "(1, C.c_float(0.0), C.c_float(60.0))"),
    ("asgramcf_create", "(0,)"),
]


def test_error_messages_and_exit_status_equal_the_reference(L):
    """the same invalid call made against this library and against the compiled reference (each in its own process):
    identical first message line on stderr and identical exit status, for some sixty bad calls across every object family"""
    ref_lib = os.path.join(ROOT, "oracle", "_ref", "libliquid_ref_sse.so")
    if not os.path.exists(ref_lib):
        pytest.skip("oracle/_ref not built")
    script = ("import ctypes as C, sys\nL = C.CDLL(sys.argv[1])\n"
              "fn, args = sys.argv[2], eval(sys.argv[3])\ngetattr(L, fn)(*args)\n")
    diffs = []
    for fn, args in ERROR_TEXT_CALLS:
        seen = []
        for lib in (L.LIB_PATH, ref_lib):
            r = subprocess.run([sys.executable, "-c", script, lib, fn, args], capture_output=True, text=True, timeout=60)
            lines = [ln for ln in r.stderr.splitlines() if ln.startswith(("error", "warning"))]
            seen.append((r.returncode, lines[0] if lines else "<none>"))
        if seen[0] != seen[1] or seen[0][0] != 1:
            diffs.append((fn, args, seen))
    assert not diffs, diffs
