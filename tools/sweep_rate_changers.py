#!/usr/bin/env python
"""Roofline fraction of the rate changers over their parameters (resamp2 m, firdecim / firinterp M x m, msresamp2 stages):
finds the parameter values that fall off the fast kernels.  python tools/sweep_rate_changers.py [log2_samples]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
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

def frac(bytes_total, ms):
    return bytes_total / (ms * 1e-3) / 1e9 / peak

print("resamp2_crcf (columns m = 2 .. 12): decimator / interpolator / analyzer")
for mode, bpi in (("decim", 12), ("interp", 24), ("analyzer", 16)):
    row = []
    for m in range(2, 13):
        q = L.resamp2_crcf(m, 0.0, 60.0)
        n_units = n // 2 if mode != "interp" else n // 2
        t = timed(lambda: q.execute_block_dev(mode, x.data_ptr(), n_units, y.data_ptr(), st.cuda_stream))
        n_in = n if mode != "interp" else n // 2
        row.append("%.2f" % frac(bpi * n_in, t))
    print("  %-9s %s" % (mode, "  ".join(row)), flush=True)
for name, cls in (("firdecim_crcf", L.firdecim_crcf), ("firinterp_crcf", L.firinterp_crcf)):
    print(name, "(rows M, columns m = 2 .. 8)")
    for M in (2, 3, 4, 8, 16, 32):
        row = []
        for m in range(2, 9):
            q = cls.create_kaiser(M, m, 60.0)
            if name == "firdecim_crcf":
                t = timed(lambda: q.execute_block_dev(x.data_ptr(), n // M, y.data_ptr(), st.cuda_stream))
                row.append("%.2f" % frac((8 + 8 / M) * n, t))
            else:
                ni = n // M
                t = timed(lambda: q.execute_block_dev(x.data_ptr(), ni, y.data_ptr(), st.cuda_stream))
                row.append("%.2f" % frac((8 + 8 * M) * ni, t))
        print("  M=%-3d %s" % (M, "  ".join(row)), flush=True)
print("msresamp2_crcf (columns stages = 1 .. 6): decimator / interpolator")
for typ, nm in ((L.LIQUID_RESAMP_DECIM, "decim"), (L.LIQUID_RESAMP_INTERP, "interp")):
    row = []
    for S in range(1, 7):
        q = L.msresamp2_crcf(typ, S, 0.4, 0.0, 60.0)
        nu = n >> S
        t = timed(lambda: q.execute_block_dev(x.data_ptr(), nu, y.data_ptr(), st.cuda_stream))
        b = (8 + 8 / (1 << S)) * n if nm == "decim" else (8 + 8 * (1 << S)) * nu
        row.append("%.2f" % frac(b, t))
    print("  %-7s %s" % (nm, "  ".join(row)), flush=True)
