#!/usr/bin/env python
"""Golden vectors for the first "next" row of SURVEY.md 8(f): the root-Nyquist Kaiser prototypes and
firpfbch_cccf (complex taps), generated HERE by the
real reference compiled into oracle/_ref (make -C oracle ref) and committed as
tests/golden/ref_outputs_next.npz.  /root/reference does not exist on the GPU box; this script is
the record of how the fixture was made.

  taps_*     liquid_firdes_arkaiser / liquid_firdes_rkaiser / liquid_firdes_prototype(KAISER) from the
             baseline x86-64 build of the reference (libliquid_ref_sse.so: no FMA contraction, so the C
             restatement and the host design code can be held to it bit for bit)
  pfbch_*    firpfbch_crcf_create_rnyquist(ARKAISER / RKAISER) analysis and synthesis outputs, from the same
             baseline build (with FMA contraction the rkaiser search settles on a rho that differs in the
             5th digit, which moves every tap by ~1e-5: fine for a filter, useless as a pin)
  scalars    estimate_req_filter_As, rkaiser_approximate_rho, liquid_filter_isi points
  cccf_*     firpfbch_cccf_create with seeded complex taps: analysis and synthesis outputs
  mr_*       firdecim_crcf / firinterp_crcf (8f item 3): Kaiser objects and seeded random taps
  spgram_*   spgramcf (8f item 2): execute / execute_psd after a write, accumulate_psd in two calls
             (alpha 0.3: the first call runs with alpha = 1 by the reference's quirk) + write_accumulation,
             and estimate_psd, on a tone-plus-noise input
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
os.environ["LIQUID_ORACLE_REF_VARIANT"] = "_sse"
sys.path.insert(0, os.path.dirname(HERE))
from oracle import oracle as O  # noqa: E402

c_f_p = np.ctypeslib.ndpointer(np.float32, flags="C")
S = C.CDLL(os.path.join(HERE, "_ref", "libliquid_ref_sse.so"))
for fn in ("liquid_firdes_arkaiser", "liquid_firdes_rkaiser"):
    getattr(S, fn).restype = None
    getattr(S, fn).argtypes = [C.c_uint, C.c_uint, C.c_float, C.c_float, c_f_p]
S.liquid_firdes_prototype.restype = None
S.liquid_firdes_prototype.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_float, C.c_float, c_f_p]
S.estimate_req_filter_As.restype = C.c_float
S.estimate_req_filter_As.argtypes = [C.c_float, C.c_uint]
S.rkaiser_approximate_rho.restype = C.c_float
S.rkaiser_approximate_rho.argtypes = [C.c_uint, C.c_float]
S.liquid_filter_isi.restype = None
S.liquid_filter_isi.argtypes = [c_f_p, C.c_uint, C.c_uint, C.POINTER(C.c_float), C.POINTER(C.c_float)]

KAISER, ARKAISER, RKAISER = 1, 7, 8
out = {}
cases = [(2, 3, 0.3, 0.0), (4, 5, 0.5, 0.0), (8, 3, 0.4, 0.0), (16, 7, 0.25, 0.0), (64, 4, 0.9, 0.0),
         (3, 12, 0.1, 0.0), (5, 25, 0.35, 0.0), (4, 3, 0.3, 0.25), (256, 3, 0.5, 0.0)]
for k, m, beta, dt in cases:
    tag = "k%d_m%d_b%03d_dt%03d" % (k, m, round(beta * 100), round(dt * 100))
    for name, fn in (("arkaiser", S.liquid_firdes_arkaiser), ("rkaiser", S.liquid_firdes_rkaiser)):
        h = np.empty(2 * k * m + 1, np.float32)
        fn(k, m, beta, dt, h)
        out["taps_%s_%s" % (name, tag)] = h
    h = np.empty(2 * k * m + 1, np.float32)
    S.liquid_firdes_prototype(KAISER, k, m, beta, dt, h)
    out["taps_nkaiser_%s" % tag] = h
    rms, peak = C.c_float(), C.c_float()
    S.liquid_filter_isi(out["taps_rkaiser_%s" % tag], k, m, C.byref(rms), C.byref(peak))
    out["isi_rkaiser_%s" % tag] = np.array([rms.value, peak.value], np.float32)
    out["args_%s" % tag] = np.array([k, m, beta, dt], np.float64)
out["req_As"] = np.array([[df, N, S.estimate_req_filter_As(df, N)] for df, N in
                          [(0.05, 33), (0.01, 129), (0.2, 13), (0.003, 2049), (0.1, 7)]], np.float64)
out["approx_rho"] = np.array([[m, b, S.rkaiser_approximate_rho(m, b)] for m, b in
                              [(1, 0.2), (3, 0.5), (7, 0.25), (12, 0.9), (22, 0.05), (30, 0.35)]], np.float64)

# channelizers built on the prototypes (baseline build, see above)
for ftype, name in ((ARKAISER, "arkaiser"), (RKAISER, "rkaiser")):
    for M, m, beta in [(8, 3, 0.4), (32, 4, 0.3)]:
        x = O.synth_input(40 * M, seed=0x5EED7000 + M + ftype)
        out["pfbch_%s_M%d_m%d_x" % (name, M, m)] = x
        out["pfbch_%s_a_M%d_m%d_y" % (name, M, m)] = O.ref_firpfbch_frames(O.ANALYZER, M, x, rnyquist=(m, beta, ftype))
        out["pfbch_%s_s_M%d_m%d_y" % (name, M, m)] = O.ref_firpfbch_frames(O.SYNTHESIZER, M, x, rnyquist=(m, beta, ftype))
        out["pfbch_%s_M%d_m%d_args" % (name, M, m)] = np.array([M, m, beta, ftype], np.float64)

# firpfbch_cccf with genuinely complex taps
rng = np.random.default_rng(20261018)
for M, p in [(8, 6), (16, 8), (12, 5), (64, 4)]:
    h = ((rng.standard_normal(M * p) + 1j * rng.standard_normal(M * p)) / p).astype(np.complex64)
    x = O.synth_input(40 * M, seed=0x5EED8000 + M)
    out["cccf_M%d_p%d_h" % (M, p)] = h
    out["cccf_M%d_p%d_x" % (M, p)] = x
    out["cccf_M%d_p%d_a_y" % (M, p)] = O.ref_firpfbch_cccf_frames(O.ANALYZER, M, p, h, x)
    out["cccf_M%d_p%d_s_y" % (M, p)] = O.ref_firpfbch_cccf_frames(O.SYNTHESIZER, M, p, h, x)

# firdecim_crcf / firinterp_crcf (8f item 3)
rng3 = np.random.default_rng(20261019)
for M, m in [(2, 4), (4, 5), (8, 3), (5, 2), (16, 6)]:
    x = O.synth_input(M * 300, seed=0x5EEDA000 + M)
    h = rng3.standard_normal(2 * M * m + 3).astype(np.float32)
    tag = "mr_M%d_m%d" % (M, m)
    out[tag + "_x"] = x
    out[tag + "_h"] = h
    out[tag + "_decim_kaiser"] = O.ref_firdecim_block(M, x, kaiser=(m, 60.0))
    out[tag + "_decim_h"] = O.ref_firdecim_block(M, x, h=h)
    out[tag + "_interp_kaiser"] = O.ref_firinterp_block(M, x[:300], kaiser=(m, 60.0))
    out[tag + "_interp_h"] = O.ref_firinterp_block(M, x[:300], h=h)

# spgramcf
for nfft, wl, beta in [(64, 32, 10.0), (256, 200, 6.0), (1024, 512, 10.0), (100, 60, 8.0), (16, 16, 3.0)]:
    n = 6 * nfft + 37
    t = np.arange(n)
    x = (O.synth_input(n, seed=0x5EED9000 + nfft) * np.float32(0.1)
         + np.exp(2j * np.pi * 0.17 * t).astype(np.complex64)).astype(np.complex64)
    tag = "spgram_n%d_w%d" % (nfft, wl)
    out[tag + "_args"] = np.array([nfft, wl, beta], np.float64)
    out[tag + "_x"] = x
    q = O.RefSpgramcf(nfft, window_len=wl, beta=beta)
    q.write(x[: wl + 5])
    out[tag + "_X"] = q.execute()
    out[tag + "_psd"] = q.execute_psd()
    q.reset()
    q.accumulate_psd(x[: 2 * nfft + 3], 0.3)
    out[tag + "_acc1"] = q.write_accumulation()
    q.accumulate_psd(x[2 * nfft + 3:], 0.3)
    out[tag + "_acc2"] = q.write_accumulation()
    out[tag + "_est"] = q.estimate_psd(x)
    del q

path = os.path.join(HERE, "..", "tests", "golden", "ref_outputs_next.npz")
np.savez_compressed(path, **out)
print("wrote", os.path.normpath(path), "%d arrays, %.0f KB" % (len(out), os.path.getsize(path) / 1024))
