#!/usr/bin/env python
"""Golden vectors for SURVEY.md 8(f) item 3, second part: resamp2_crcf and msresamp2_crcf, generated HERE by the
real reference compiled into oracle/_ref (make -C oracle ref; the baseline x86-64 build, no FMA contraction) and
committed as tests/golden/ref_outputs_next2.npz.  /root/reference does not exist on the GPU box; this script is
the record of how the fixture was made.

  hb_*     resamp2_crcf(m, f0, As): the five modes run unit by unit through the reference's per-call API on one
           seeded input (filter: y0 | y1), and a mixed-mode run on one object (decim, filter, interp, analyzer,
           synthesizer, filter in turn) that pins the shared window / toggle state
  ms_*     msresamp2_crcf(type, stages, fc, 0, As): outputs, get_delay and the stage semi-lengths implied by
           estimate_req_filter_len
  len_*    estimate_req_filter_len points
  asg_*    asgramcf(nfft) after a write of a tone plus noise: the character string and the peak (default and custom scale)
  ofdm_*   ofdmframegen(M, cp_len, taper_len, allocation): a whole frame S0a | S0b | S1 | data symbols | tail
  taps_hM3_* / taps_pm_* / pmrun_* / pfbch_hM3_*   liquid_firdes_hM3, liquid_firdes_prototype(PM), firdespm_run and
           firpfbch_crcf_create_rnyquist(hM3) outputs
  ff_*     fftfilt_crcf(h, n) over several blocks (random taps / a Kaiser low-pass, with and without set_scale)
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
os.environ["LIQUID_ORACLE_REF_VARIANT"] = "_sse"
sys.path.insert(0, os.path.dirname(HERE))
from oracle import oracle as O  # noqa: E402

out = {}
for m, f0, As in [(2, 0.0, 40.0), (4, 0.0, 60.0), (7, 0.1, 70.0), (12, -0.25, 80.0), (5, 0.5, 60.0)]:
    tag = "hb_m%d_f%+04d" % (m, round(f0 * 100))
    x = O.synth_input(600, seed=0x5EEDB000 + m)
    out[tag + "_args"] = np.array([m, f0, As], np.float64)
    out[tag + "_x"] = x
    y0, y1 = O.ref_resamp2("filter", m, f0, As, x)
    out[tag + "_filter"] = np.concatenate([y0, y1])
    for mode in ("analyzer", "synthesizer", "decim", "interp"):
        out[tag + "_" + mode] = O.ref_resamp2(mode, m, f0, As, x)
    out[tag + "_mixed"] = O.ref_resamp2_mixed(m, f0, As, x)

for typ, S, fc, As in [(0, 1, 0.25, 60.0), (0, 3, 0.4, 60.0), (0, 5, 0.1, 80.0), (1, 1, 0.25, 60.0),
                       (1, 3, 0.4, 60.0), (1, 4, 0.45, 70.0), (1, 6, 0.2, 60.0), (0, 0, 0.2, 60.0)]:
    tag = "ms_t%d_S%d_fc%03d" % (typ, S, round(fc * 100))
    n = 40 if typ == 0 else 30 * (1 << S)
    x = O.synth_input(n, seed=0x5EEDC000 + 16 * typ + S)
    y, delay = O.ref_msresamp2_block(typ, S, fc, 0.0, As, x)
    out[tag + "_args"] = np.array([typ, S, fc, As, delay], np.float64)
    out[tag + "_x"] = x
    out[tag + "_y"] = y

# fftfilt_crcf: seeded random taps and a Kaiser low-pass; power-of-two and other block sizes; a set_scale case
rng = np.random.default_rng(20261020)
for h_len, n, scale in [(5, 8, None), (21, 32, None), (64, 64, 0.5), (33, 100, None), (129, 128, 2.0), (200, 512, None),
                        (17, 30, None)]:
    tag = "ff_h%d_n%d" % (h_len, n)
    h = rng.standard_normal(h_len).astype(np.float32) if h_len != 129 else O.firdes_kaiser(129, 0.2, 60.0)
    x = O.synth_input(n * (9 if n < 200 else 4), seed=0x5EEDD000 + n)
    out[tag + "_args"] = np.array([h_len, n, -1.0 if scale is None else scale], np.float64)
    out[tag + "_h"] = h
    out[tag + "_x"] = x
    out[tag + "_y"] = O.ref_fftfilt_blocks(h, n, x, scale)

# asgramcf: a tone plus noise, default and custom scale
for nfft, offset, scale in [(64, None, None), (72, -20.0, 5.0), (256, -30.0, 4.0)]:
    tag = "asg_n%d" % nfft
    t = np.arange(3 * nfft + 5)
    x = (O.synth_input(len(t), seed=0x5EEDE000 + nfft) * np.float32(0.05)
         + np.exp(2j * np.pi * 0.21 * t).astype(np.complex64)).astype(np.complex64)
    ascii_, pv, pf = O.ref_asgramcf(nfft, x, offset, scale)
    out[tag + "_args"] = np.array([nfft, np.nan if offset is None else offset, np.nan if scale is None else scale], np.float64)
    out[tag + "_x"] = x
    out[tag + "_ascii"] = np.frombuffer(ascii_.encode("ascii"), np.uint8).copy()
    out[tag + "_peak"] = np.array([pv, pf], np.float32)

# ofdmframegen: whole frames (S0a | S0b | S1 | data symbols | tail), default and custom allocations, power-of-two
# and other subcarrier counts, with and without taper
def custom_alloc(M):
    p = np.full(M, 2, np.uint8)
    p[0] = p[M // 2] = 0
    p[1] = p[M - 1] = p[3] = 1
    return p


for M, cp, tl, custom, nsym in [(64, 16, 4, False, 9), (64, 16, 0, False, 3), (120, 20, 8, False, 5), (256, 32, 32, False, 40),
                                (32, 8, 2, True, 70), (1024, 128, 16, False, 3), (10, 2, 1, True, 4)]:
    tag = "ofdm_M%d_cp%d_t%d" % (M, cp, tl)
    p = custom_alloc(M) if custom else None
    x = O.synth_input(M * nsym, seed=0x5EEDF000 + M)
    out[tag + "_args"] = np.array([M, cp, tl, int(custom)], np.int64)
    out[tag + "_x"] = x
    out[tag + "_frame"] = O.ofdm_frame(O.RefOfdmframegen(M, cp, tl, p), x)

# Parks-McClellan designs and the harris-Moerder prototype (8f item 1, the last prototype of create_rnyquist): taps from
# the reference (both builds agree bit for bit: the designer runs in double precision), and firpfbch_crcf outputs
# with the hM3 prototype.  (k=3, m=7, beta=0.9 and its like trip an assert inside the reference's own extremal search
# -- firdespm.c:544 -- and are left out.)
Rr = O.ref()
c_f_p2 = np.ctypeslib.ndpointer(np.float32, flags="C")
Rr.liquid_firdes_hM3.argtypes = [C.c_uint, C.c_uint, C.c_float, C.c_float, c_f_p2]
Rr.liquid_firdes_prototype.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_float, C.c_float, c_f_p2]
Rr.firdespm_run.argtypes = [C.c_uint, C.c_uint, c_f_p2, c_f_p2, c_f_p2, C.POINTER(C.c_int), C.c_int, c_f_p2]
pm_cases = []
for k, m, beta in [(2, 3, 0.3), (4, 5, 0.5), (8, 3, 0.4), (2, 12, 0.2), (16, 4, 0.35), (3, 7, 0.5), (5, 9, 0.25), (64, 2, 0.5)]:
    n = 2 * k * m + 1
    h = np.empty(n, np.float32)
    Rr.liquid_firdes_hM3(k, m, beta, 0.0, h)
    out["taps_hM3_k%d_m%d_b%03d" % (k, m, round(beta * 100))] = h.copy()
    Rr.liquid_firdes_prototype(2, k, m, beta, 0.0, h)                 # LIQUID_FIRFILT_PM
    out["taps_pm_k%d_m%d_b%03d" % (k, m, round(beta * 100))] = h.copy()
    pm_cases.append((k, m, beta))
out["pm_cases"] = np.array(pm_cases, np.float64)
# firdespm_run on its own: low-pass with weights, even length, exponential / linear weights, four bands
for tag, h_len, bands, des, wts, wtype in [
        ("lp51", 51, [0.0, 0.2, 0.25, 0.5], [1.0, 0.0], [1.0, 10.0], [0, 0]),
        ("lp40", 40, [0.0, 0.1, 0.15, 0.5], [1.0, 0.0], [1.0, 1.0], [0, 1]),
        ("bp77", 77, [0.0, 0.1, 0.15, 0.3, 0.35, 0.5], [0.0, 1.0, 0.0], [4.0, 1.0, 4.0], [0, 0, 2]),
        ("mb64", 64, [0.0, 0.08, 0.12, 0.2, 0.25, 0.33, 0.38, 0.5], [1.0, 0.0, 0.5, 0.0], [1.0, 2.0, 1.0, 3.0], [0, 0, 0, 0])]:
    h = np.empty(h_len, np.float32)
    wt = np.array(wtype, np.int32)
    Rr.firdespm_run(h_len, len(des), np.array(bands, np.float32), np.array(des, np.float32), np.array(wts, np.float32),
                    wt.ctypes.data_as(C.POINTER(C.c_int)), 0, h)
    out["pmrun_%s_bands" % tag] = np.array(bands, np.float32)
    out["pmrun_%s_des" % tag] = np.array(des, np.float32)
    out["pmrun_%s_weights" % tag] = np.array(wts, np.float32)
    out["pmrun_%s_wtype" % tag] = wt
    out["pmrun_%s_h" % tag] = h
# the other families of liquid_firdes_prototype (rcos, fexp / fsech / farcsech and their roots, gmsk tx / rx)
proto_cases = [(2, 3, 0.3, 0.0), (4, 5, 0.5, 0.0), (8, 3, 0.4, 0.25), (2, 12, 0.2, 0.0), (16, 4, 0.35, 0.0), (3, 7, 0.9, -0.5)]
out["proto_cases"] = np.array(proto_cases, np.float64)
for ftype in (3, 4, 5, 6, 11, 12, 13, 14, 15):
    for k, m, beta, dt in proto_cases:
        h = np.empty(2 * k * m + 1, np.float32)
        Rr.liquid_firdes_prototype(ftype, k, m, beta, dt, h)
        out["taps_proto%d_k%d_m%d" % (ftype, k, m)] = h.copy()
for M, m, beta in [(8, 3, 0.4), (64, 2, 0.5)]:
    x = O.synth_input(M * 60, seed=0x5EEDAA00 + M)
    out["pfbch_hM3_M%d_m%d_x" % (M, m)] = x
    out["pfbch_hM3_M%d_m%d_args" % (M, m)] = np.array([M, m, beta], np.float64)
    out["pfbch_hM3_a_M%d_m%d_y" % (M, m)] = O.ref_firpfbch_frames(O.ANALYZER, M, x, rnyquist=(m, beta, 10))
    out["pfbch_hM3_s_M%d_m%d_y" % (M, m)] = O.ref_firpfbch_frames(O.SYNTHESIZER, M, x, rnyquist=(m, beta, 10))

R = O.ref()
R.estimate_req_filter_len.restype = C.c_uint
R.estimate_req_filter_len.argtypes = [C.c_float, C.c_float]
pts = [(0.05, 60.0), (0.1, 60.0), (0.2, 80.0), (0.1375, 60.0), (0.01, 100.0), (0.3, 20.0), (0.22, 60.0)]
out["len_points"] = np.array(pts, np.float32)
out["len_values"] = np.array([R.estimate_req_filter_len(df, As) for df, As in pts], np.uint32)

path = os.path.join(HERE, "..", "tests", "golden", "ref_outputs_next2.npz")
np.savez_compressed(path, **out)
print("wrote", os.path.normpath(path), "%d arrays, %.0f KB" % (len(out), os.path.getsize(path) / 1024))
