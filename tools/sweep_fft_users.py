#!/usr/bin/env python
"""Rates of the transform users over their parameters: ofdmframegen (M x cyclic prefix x taper), spgramcf (nfft x window),
fftfilt_crcf (block length).  python tools/sweep_fft_users.py [log2_samples]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import liquid_dsp_b200 as L
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
L.set_device(0)
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream()
peak = 6531.6
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = float(json.load(open(p))["hbm_gbs"])
n = 1 << log2n
x = torch.empty(2 * n, dtype=torch.float32, device=dev)
y = torch.empty(2 * 2 * n, dtype=torch.float32, device=dev)
L.synth_input_dev(x.data_ptr(), n, stream=st.cuda_stream)

def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(3):
        fn()
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 3

print("ofdmframegen writesymbols: fraction of the 8 + 8 (M + cp) / M byte roofline; columns (cp, taper) = (0,0) (M/8,0) (M/8,4) (M/4,8)")
for M in (32, 64, 100, 128, 256, 512, 1024, 2048, 4096):
    row = []
    for cp, tl in ((0, 0), (M // 8, 0), (M // 8, 4), (M // 4, 8)):
        q = L.ofdmframegen(M, cp, min(tl, cp))
        t = timed(lambda: q.writesymbols_dev(x.data_ptr(), n // M, y.data_ptr(), st.cuda_stream))
        row.append("%.2f" % ((8 + 8 * (M + cp) / M) * n / (t * 1e-3) / 1e9 / peak))
    print("  M=%-5d %s" % (M, "  ".join(row)), flush=True)
print("spgramcf accumulate_psd: G input samples/s; columns window = nfft/4, nfft/2, nfft (a transform every window/2 samples)")
for nfft in (64, 128, 256, 300, 512, 1024, 2048, 4096, 8192):
    row = []
    for wl in (nfft // 4, nfft // 2, nfft):
        q = L.spgramcf.create_kaiser(nfft, wl, 8.0)
        t = timed(lambda: q.accumulate_psd_dev(x.data_ptr(), 0.01, n, st.cuda_stream))
        row.append("%6.1f" % (n / (t * 1e-3) / 1e9))
    print("  nfft=%-5d %s" % (nfft, "  ".join(row)), flush=True)
print("fftfilt_crcf: G samples/s; filter of n + 1 taps in blocks of n")
row = []
for nb in (16, 32, 64, 100, 128, 256, 512, 1024, 2048, 4096):
    h = np.asarray(L.liquid_firdes_kaiser(nb + 1, 0.2, 60.0, 0.0), np.float32)
    q = L.fftfilt_crcf(h, nb)
    t = timed(lambda: q.execute_block_dev(x.data_ptr(), n // nb, y.data_ptr(), st.cuda_stream))
    row.append("n=%d: %.1f" % (nb, n / (t * 1e-3) / 1e9))
print("  " + "   ".join(row), flush=True)
