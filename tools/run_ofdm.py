#!/usr/bin/env python
"""ofdmframegen data symbols on device-resident values, for ncu captures: python tools/run_ofdm.py M cp taper log2n"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import liquid_dsp_b200 as L
M, cp, tl, log2n = (int(v) for v in sys.argv[1:5])
L.set_device(0)
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream()
n_s = 1 << log2n
x = torch.empty(2 * n_s, dtype=torch.float32, device=dev)
L.synth_input_dev(x.data_ptr(), n_s, stream=st.cuda_stream)
y = torch.empty(2 * (n_s // M) * (M + cp), dtype=torch.float32, device=dev)
q = L.ofdmframegen(M, cp, tl)
for _ in range(2):
    q.writesymbols_dev(x.data_ptr(), n_s // M, y.data_ptr(), st.cuda_stream)
torch.cuda.synchronize()
print("ok")
