#!/usr/bin/env python
"""One batched transform on device-resident data, for ncu captures: python tools/run_fft.py n log2_points"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import liquid_dsp_b200 as L
nn, log2n = int(sys.argv[1]), int(sys.argv[2])
L.set_device(0)
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream()
n = 1 << log2n
x = torch.empty(2 * n, dtype=torch.float32, device=dev)
y = torch.empty(2 * n, dtype=torch.float32, device=dev)
L.synth_input_dev(x.data_ptr(), n, stream=st.cuda_stream)
hx, hy = np.zeros(nn, np.complex64), np.zeros(nn, np.complex64)
p = L.fftplan(nn, hx, hy, L.LIQUID_FFT_FORWARD)
for _ in range(2):
    p.execute_batch_dev(x.data_ptr(), y.data_ptr(), n // nn, st.cuda_stream)
torch.cuda.synchronize()
print("ok")
