#!/usr/bin/env python
"""Multi-GPU parity check of the time-sharded path (run under torchrun on N GPUs, NCCL):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29533 tests/multi_gpu_check.py [--log2-samples 24]

Every rank channelises its time segment of ONE synthetic stream (M=1024, m=7) after receiving the
2mM-sample halo from its left neighbour over NCCL, then checks the frames around its segment start
(the ones that depend on the halo) and a block in the middle against the CPU oracle, which is fed
from 2mM samples before the block (state = last 2mM - M/2 samples).  Test
infrastructure: lives in tests/, may use the oracle."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2-samples", type=int, default=24)
    ap.add_argument("--channels", type=int, default=1024)
    ap.add_argument("--semi", type=int, default=7)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import liquid_dsp_b200 as L
    import liquid_dsp_b200.multi as multi
    from oracle import oracle as O

    M, m = args.channels, args.semi
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    L.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 1 << args.log2_samples
    n_frames = n // (M // 2)
    q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
    H = q.get_state_len()
    hf = H // (M // 2)
    xbuf = torch.zeros(2 * (H + n), dtype=torch.float32, device=dev)
    ybuf = torch.empty(2 * n_frames * M, dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream()
    d_halo = xbuf.data_ptr()
    d_x = d_halo + 8 * H
    L.synth_input_dev(d_x, n, seed=0x5EED0000, first=rank * n, stream=st.cuda_stream)
    torch.cuda.synchronize()
    head, tail = xbuf[: 2 * H], xbuf[2 * n: 2 * (H + n)]
    reqs = multi.exchange_halo(dist, head, tail, rank, world)
    first = rank * n_frames
    q.execute_segment_dev(d_x, d_x + 8 * hf * (M // 2), first + hf, n_frames - hf, ybuf.data_ptr() + 8 * hf * M, st.cuda_stream)
    for r in reqs:
        r.wait()
    q.execute_segment_dev(d_halo, d_x, first, hf, ybuf.data_ptr(), st.cuda_stream)
    torch.cuda.synchronize()
    y = ybuf.cpu().numpy().view(np.complex64).reshape(n_frames, M)

    def oracle_block(f0, nf):
        """frames [f0, f0+nf) of this rank's segment from the global stream"""
        g0 = rank * n + f0 * (M // 2)                       # global index of the block's first sample
        start = max(0, g0 - H)
        x = O.synth_input(g0 - start + nf * (M // 2), seed=0x5EED0000, first=start)
        out = O.Firpfbch2(O.ANALYZER, M, m, As=60.0).execute_frames(x)
        return out[(g0 - start) // (M // 2):], x

    worst = 0.0
    for f0, nf in ((0, 64), (n_frames // 2, 32), (n_frames - 32, 32)):
        want, x = oracle_block(f0, nf)
        got = y[f0:f0 + nf]
        num = (np.abs(got.astype(np.complex128) - want) ** 2).sum(axis=0)
        den = (np.abs(want.astype(np.complex128)) ** 2).sum(axis=0)
        rel = float(np.sqrt(num / den).max())
        mx = float(np.abs(got - want).max() / np.sqrt(np.mean(np.abs(x) ** 2)))
        assert rel < 1e-5 and mx < 1e-4, (rank, f0, rel, mx)
        worst = max(worst, rel)
    t = torch.tensor([worst], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("multi_gpu_check OK: world=%d, 2^%d samples per GPU, worst per-channel rel-RMS %.2e" % (world, args.log2_samples, t.item()))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
