#!/usr/bin/env python
"""One launch of BASELINE config 2 at a reduced size (64-tap firfilt, n_streams x 2^20 samples) for ncu captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import liquid_dsp_b200 as L
ns, n = int(sys.argv[1]) if len(sys.argv) > 1 else 128, 1 << 20
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 64
L.set_device(0)
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream()
q = L.firfilt_crcf.create_kaiser(nt, 0.25, 60.0, 0.0)
x = torch.empty(2 * ns * n, dtype=torch.float32, device=dev)
y = torch.empty(2 * ns * n, dtype=torch.float32, device=dev)
L.synth_input_dev(x.data_ptr(), ns * n, stream=st.cuda_stream)
for _ in range(2):
    q.execute_streams_dev(x.data_ptr(), ns, n, y.data_ptr(), st.cuda_stream)
torch.cuda.synchronize()
print("ok")
