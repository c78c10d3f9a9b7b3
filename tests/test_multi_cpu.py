"""world_size-2 gloo tests of the multi-rank host logic (CPU only): time-segment sharding with the
filter-history halo exchanged between neighbours reproduces the single-stream result; stream
sharding partitions without overlap.  The per-rank compute is done by the CPU oracle here -- the
point of these tests is the sharding / halo logic of liquid-dsp_b200/multi.py, which bench.py and a
GPU deployment use unchanged with the NCCL backend."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, M, m, n_frames_total, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    import importlib.util
    spec = importlib.util.spec_from_file_location("lc_multi", os.path.join(ROOT, "liquid-dsp_b200", "multi.py"))
    multi = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(multi)

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    H = 2 * m * M                                   # firpfbch2_crcf_get_state_len
    lo, hi = multi.segment_of_rank(n_frames_total, rank, world)
    assert lo % 2 == 0
    n = (hi - lo) * (M // 2)
    x_seg = O.synth_input(n, seed=99, first=lo * (M // 2))      # this rank's samples of the one stream
    buf = torch.zeros(2 * (H + n), dtype=torch.float32)         # head-room | segment
    buf[2 * H:] = torch.from_numpy(x_seg.view(np.float32).copy())
    head, tail = buf[: 2 * H], buf[2 * n: 2 * (H + n)]
    for r in multi.exchange_halo(dist, head, tail, rank, world):
        r.wait()
    # per-rank compute with the halo as prior history: run the oracle from the zero state over
    # (halo | segment) and drop the halo's frames -- identical to starting from the carried state
    full = buf.numpy().view(np.complex64)
    y = O.Firpfbch2(O.ANALYZER, M, m, As=60.0).execute_frames(full)[H // (M // 2):]
    np.save(os.path.join(out_dir, "y_%d.npy" % rank), y)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("M,m,n_frames", [(64, 3, 203), (16, 5, 90)])
def test_time_sharded_halo_exchange_world2(tmp_path, orc, M, m, n_frames):
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, M, m, n_frames, str(tmp_path)), nprocs=world, join=True)
    y = np.concatenate([np.load(tmp_path / ("y_%d.npy" % r)) for r in range(world)])
    x = orc.synth_input(n_frames * (M // 2), seed=99)
    want = orc.Firpfbch2(orc.ANALYZER, M, m, As=60.0).execute_frames(x)
    assert y.shape == want.shape
    assert np.array_equal(y, want)          # same float32 operations in the same order: bit exact


def test_segment_and_stream_partitions():
    import importlib.util
    spec = importlib.util.spec_from_file_location("lc_multi", os.path.join(ROOT, "liquid-dsp_b200", "multi.py"))
    multi = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(multi)
    for total in (0, 1, 7, 8, 100, 2 ** 21, 2 ** 21 + 3):
        for world in (1, 2, 3, 4, 8):
            segs = [multi.segment_of_rank(total, r, world) for r in range(world)]
            assert segs[0][0] == 0 and segs[-1][1] == total
            for (a, b), (c, d) in zip(segs, segs[1:]):
                assert b == c and a <= b
            assert all(lo % 2 == 0 for lo, hi in segs if hi > lo)
            st = [multi.streams_of_rank(total, r, world) for r in range(world)]
            assert st[0][0] == 0 and st[-1][1] == total
            assert all(b == c for (_, b), (c, _) in zip(st, st[1:]))
            assert max(b - a for a, b in st) - min(b - a for a, b in st) <= 1
