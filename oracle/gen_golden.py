#!/usr/bin/env python
"""Generate tests/golden/ref_outputs.npz from the REAL reference (oracle/_ref), on seeded inputs.
TEST INFRASTRUCTURE.  Run in the dev container after `make -C oracle ref`; the .npz is committed
so the parity tests can check against reference outputs where /root/reference does not exist.

Nothing pins liquid_firdes_kaiser taps, firpfbch2 analysis-only outputs or any M > 64 in the
reference's own tests, so these fixtures are the pin for those.
"""
import os
import sys

import numpy as np

sys.path[0] = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
from oracle import oracle as O  # noqa: E402

R = O.ref()
out = {}


def lcg_input(n):
    """the firpfbch2 autotest's input (firpfbch2_crcf_autotest.c:43-51)"""
    s, xs = 1, np.empty(n, np.complex64)
    for i in range(n):
        s = (s * 524287) % 1031
        xs[i] = np.float32(s) / np.float32(1031) - np.float32(0.5)
    return xs


# prototype designs
for n, fc, As in [(129, 1 / 32, 60.0), (161, 1 / 16, 60.0), (64, 0.25, 60.0), (14337, 1 / 1024, 60.0),
                  (2561, 0.5 / 256, 60.0), (97, 0.1, 40.0), (33, 0.2, 15.0)]:
    h = np.empty(n, np.float32)
    R.liquid_firdes_kaiser(n, fc, As, 0.0, h)
    out["kaiser_n%d_As%d" % (n, int(As))] = h
    out["kaiser_n%d_As%d_args" % (n, int(As))] = np.array([n, fc, As], np.float64)
for k, m, beta in [(2, 3, 0.3), (16, 4, 0.5), (8, 5, 0.25)]:
    h = np.empty(2 * k * m + 1, np.float32)
    R.liquid_firdes_rrcos(k, m, beta, 0.0, h)
    out["rrcos_k%d_m%d" % (k, m)] = h
    out["rrcos_k%d_m%d_args" % (k, m)] = np.array([k, m, beta], np.float64)

# firpfbch2 analysis / synthesis (white synthetic input and the autotest LCG input)
for M, m in [(8, 5), (16, 5), (64, 3), (6, 2), (256, 4)]:
    nfr = 48 if M <= 64 else 24
    x = O.synth_input(nfr * M // 2, seed=0x5EED0000 + M)
    out["pfbch2_a_M%d_m%d_x" % (M, m)] = x
    out["pfbch2_a_M%d_m%d_y" % (M, m)] = O.ref_firpfbch2_frames(O.ANALYZER, M, m, 60.0, x)
    X = O.synth_input(nfr * M, seed=0x5EED1000 + M)
    out["pfbch2_s_M%d_m%d_x" % (M, m)] = X
    out["pfbch2_s_M%d_m%d_y" % (M, m)] = O.ref_firpfbch2_frames(O.SYNTHESIZER, M, m, 60.0, X)
x = lcg_input(16 * 8 * 5)
out["pfbch2_a_M16_m5_lcg_x"] = x
out["pfbch2_a_M16_m5_lcg_y"] = O.ref_firpfbch2_frames(O.ANALYZER, 16, 5, 60.0, x)

# firpfbch analysis / synthesis, Kaiser and RRC prototypes
for M, m in [(4, 3), (16, 4), (10, 2), (64, 5)]:
    x = O.synth_input(40 * M, seed=0x5EED2000 + M)
    out["pfbch_a_M%d_m%d_x" % (M, m)] = x
    out["pfbch_a_M%d_m%d_y" % (M, m)] = O.ref_firpfbch_frames(O.ANALYZER, M, x, kaiser=(m, 60.0))
    out["pfbch_s_M%d_m%d_y" % (M, m)] = O.ref_firpfbch_frames(O.SYNTHESIZER, M, x, kaiser=(m, 60.0))
LIQUID_FIRFILT_RRC = 9
x = O.synth_input(40 * 8, seed=0x5EED3000)
out["pfbch_rrc_M8_m3_x"] = x
out["pfbch_rrc_a_M8_m3_y"] = O.ref_firpfbch_frames(O.ANALYZER, 8, x, rnyquist=(3, 0.4, LIQUID_FIRFILT_RRC))
out["pfbch_rrc_s_M8_m3_y"] = O.ref_firpfbch_frames(O.SYNTHESIZER, 8, x, rnyquist=(3, 0.4, LIQUID_FIRFILT_RRC))

# firfilt block filter (64-tap Kaiser low-pass of BASELINE config 2, fc = 0.25)
x = O.synth_input(2048, seed=0x5EED4000)
out["firfilt_k64_x"] = x
out["firfilt_k64_y"] = O.ref_firfilt_block(x, kaiser=(64, 0.25, 60.0, 0.0))
out["firfilt_k64_scale_y"] = O.ref_firfilt_block(x, kaiser=(64, 0.25, 60.0, 0.0), scale=0.375)

# FFT sizes the channelizers use that the reference's vectors do not cover
for n in (128, 256, 1024, 4096):
    x = O.synth_input(n, seed=0x5EED5000 + n)
    out["fft_%d_x" % n] = x
    out["fft_%d_fwd" % n] = O.ref_fft(x, O.FFT_FORWARD)
    out["fft_%d_bwd" % n] = O.ref_fft(x, O.FFT_BACKWARD)

path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ref_outputs.npz")
np.savez_compressed(path, **out)
print("wrote", os.path.normpath(path), "%d arrays, %.0f KB" % (len(out), os.path.getsize(path) / 1024))
