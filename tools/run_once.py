#!/usr/bin/env python
"""One launch of a channelizer shape on device-resident data, for `ncu` captures:
   python tools/run_once.py pfbch2|pfbch|pfbch2s|pfbchs M m log2n [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import liquid_dsp_b200 as L

kind, M, m, log2n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
L.set_device(0)
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream()
n = 1 << log2n
if kind == "pfbch2":
    q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
    H = q.get_state_len()
    x = torch.zeros(2 * (H + n), dtype=torch.float32, device=dev)
    y = torch.empty(4 * n, dtype=torch.float32, device=dev)
    L.synth_input_dev(x.data_ptr() + 8 * H, n, stream=st.cuda_stream)
    fn = lambda: q.execute_segment_dev(x.data_ptr(), x.data_ptr() + 8 * H, 0, n // (M // 2), y.data_ptr(), st.cuda_stream)
elif kind == "pfbch2s":
    q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_SYNTHESIZER, M, m, 60.0)
    x = torch.empty(4 * n, dtype=torch.float32, device=dev)
    y = torch.empty(2 * n, dtype=torch.float32, device=dev)
    L.synth_input_dev(x.data_ptr(), 2 * n, stream=st.cuda_stream)
    fn = lambda: q.execute_block_dev(x.data_ptr(), n // (M // 2), y.data_ptr(), st.cuda_stream)
else:
    q = L.firpfbch_crcf.create_kaiser(L.LIQUID_ANALYZER if kind == "pfbch" else L.LIQUID_SYNTHESIZER, M, m, 60.0)
    x = torch.empty(2 * n, dtype=torch.float32, device=dev)
    y = torch.empty(2 * n, dtype=torch.float32, device=dev)
    L.synth_input_dev(x.data_ptr(), n, stream=st.cuda_stream)
    fn = lambda: q.execute_block_dev(x.data_ptr(), n // M, y.data_ptr(), st.cuda_stream)
for _ in range(reps):
    fn()
torch.cuda.synchronize()
print("ok", kind, M, m, log2n)
