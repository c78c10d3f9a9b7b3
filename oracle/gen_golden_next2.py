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

R = O.ref()
R.estimate_req_filter_len.restype = C.c_uint
R.estimate_req_filter_len.argtypes = [C.c_float, C.c_float]
pts = [(0.05, 60.0), (0.1, 60.0), (0.2, 80.0), (0.1375, 60.0), (0.01, 100.0), (0.3, 20.0), (0.22, 60.0)]
out["len_points"] = np.array(pts, np.float32)
out["len_values"] = np.array([R.estimate_req_filter_len(df, As) for df, As in pts], np.uint32)

path = os.path.join(HERE, "..", "tests", "golden", "ref_outputs_next2.npz")
np.savez_compressed(path, **out)
print("wrote", os.path.normpath(path), "%d arrays, %.0f KB" % (len(out), os.path.getsize(path) / 1024))
