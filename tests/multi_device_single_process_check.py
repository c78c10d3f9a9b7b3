#!/usr/bin/env python
"""One process, several GPUs (run by hand on a multi-GPU box: python tests/multi_device_single_process_check.py):
the same channelizer calls on every visible device in turn -- liquid_cuda_set_device(d), create, execute,
destroy -- must give bit-identical results.  Exercises the per-device kernel configuration (opt-in shared
memory sizes, cluster occupancy) that a process which only ever uses one device never notices."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import liquid_dsp_b200 as L
    from oracle import oracle as O
    n_dev = L.device_count()
    assert n_dev >= 1
    results = []
    for d in range(n_dev):
        L.set_device(d)
        out = {}
        for M, m, nfr in ((1024, 7, 200), (4096, 9, 100), (256, 4, 64)):
            x = O.synth_input(nfr * M // 2, seed=0x700 + M)
            q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
            out["pfbch2_%d" % M] = q.execute_block(x)
            q.destroy()
        x = O.synth_input(16 * 40 * 256, seed=0x701)
        q = L.firpfbch_crcf.create_kaiser(L.LIQUID_ANALYZER, 256, 5, 60.0)
        out["pfbch_256"] = q.execute_block(x)
        q.destroy()
        q = L.firfilt_crcf.create_kaiser(64, 0.25, 60.0, 0.0)
        out["firfilt"] = q.execute_block(x[:100000])
        q.destroy()
        q = L.spgramcf.create_default(1024)
        out["spgram"] = q.estimate_psd(x[:200000])
        q.destroy()
        results.append(out)
    for d in range(1, n_dev):
        for k in results[0]:
            assert np.array_equal(results[0][k], results[d][k]), (d, k)
    want = O.Firpfbch2(O.ANALYZER, 1024, 7, As=60.0).execute_frames(O.synth_input(200 * 512, seed=0x700 + 1024))
    assert np.abs(results[0]["pfbch2_1024"] - want).max() < 1e-5
    print("multi_device_single_process_check OK: %d device(s), %d outputs each identical" % (n_dev, len(results[0])))


if __name__ == "__main__":
    main()
