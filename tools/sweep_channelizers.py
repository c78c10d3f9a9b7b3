#!/usr/bin/env python
"""Roofline fraction of both channelizers, analysis and synthesis, over M and m: finds the shapes that fall off the
dedicated kernels.  python tools/sweep_channelizers.py [log2_samples]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import json
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
x = torch.zeros(4 * n + (1 << 22), dtype=torch.float32, device=dev)
y = torch.empty(4 * n, dtype=torch.float32, device=dev)
L.synth_input_dev(x.data_ptr(), 2 * n, stream=st.cuda_stream)

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

Ms = tuple(int(v) for v in os.environ.get("SWEEP_M", "4,8,16,32,64,128,256,512,1024,2048,4096").split(","))
ms_ = (1, 2, 3, 4, 5, 6, 7, 8, 9)
for kind in ("firpfbch2 analysis", "firpfbch2 synthesis", "firpfbch analysis", "firpfbch synthesis"):
    print(kind, "(fraction of the HBM roofline; rows M, columns m = %s)" % (ms_,), flush=True)
    for M in Ms:
        row = []
        for m in ms_:
            if kind == "firpfbch2 analysis":
                q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
                t = timed(lambda: q.execute_block_dev(x.data_ptr(), n // (M // 2), y.data_ptr(), st.cuda_stream))
                frac = 24 * n / (t * 1e-3) / 1e9 / peak
            elif kind == "firpfbch2 synthesis":
                q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_SYNTHESIZER, M, m, 60.0)
                t = timed(lambda: q.execute_block_dev(x.data_ptr(), (n // 2) // (M // 2), y.data_ptr(), st.cuda_stream))
                frac = 24 * (n // 2) / (t * 1e-3) / 1e9 / peak
            else:
                typ = L.LIQUID_ANALYZER if kind.endswith("analysis") else L.LIQUID_SYNTHESIZER
                q = L.firpfbch_crcf.create_kaiser(typ, M, m, 60.0)
                t = timed(lambda: q.execute_block_dev(x.data_ptr(), n // M, y.data_ptr(), st.cuda_stream))
                frac = 16 * n / (t * 1e-3) / 1e9 / peak
            q.destroy()
            row.append("%.2f" % frac)
        print("  M=%-5d %s" % (M, "  ".join(row)), flush=True)
