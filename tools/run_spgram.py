#!/usr/bin/env python
"""spgramcf_accumulate_psd on device-resident samples, for ncu captures: python tools/run_spgram.py nfft log2n"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import liquid_dsp_b200 as L
nfft, log2n = int(sys.argv[1]), int(sys.argv[2])
L.set_device(0)
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream()
n = 1 << log2n
x = torch.empty(2 * n, dtype=torch.float32, device=dev)
L.synth_input_dev(x.data_ptr(), n, stream=st.cuda_stream)
q = L.spgramcf.create_default(nfft)
for _ in range(2):
    q.accumulate_psd_dev(x.data_ptr(), 0.01, n, st.cuda_stream)
torch.cuda.synchronize()
print("ok")
