#!/usr/bin/env python
"""Secondary benchmarks: the other BASELINE.json configs (parity-test cases in bench.py's contract,
measured here for the record).  One JSON line per config, device-resident synthetic data, CUDA
events, inputs larger than L2.  Not the driver's headline bench (that is bench.py).

  python tools/bench_configs.py [config ...] [--steps K]

  c1 c2 c4 c5   BASELINE configs 1, 2, 4, 5 (c3 is bench.py)          c3m c5m  c3 / c5 against the filter length m
  g             shape-generic one-kernel analyzers (M = 256, 64, 2048)  s        synthesizers (firpfbch2 M=1024, firpfbch M=1024 / 256)
  rb            the reference's own benchmark shapes (M = 4 ... 1024, m = 2), firpfbch2 analysis / synthesis, firpfbch analysis
  sm            the smallest shapes, M = 4 ... 32
  t ts tp       firpfbch2 tile kernels (analysis / synthesis), M = 64 ... 512; firpfbch analysis M = 64 ... 1024
  fft           fft_execute_batch_dev, n = 16 / 256 / 1024 / 4096       d        dotprod_crcf batch entry point
  p             firpfb_crcf block form
  f1            firpfbch_cccf (8f item 1)                                f2       spgramcf_accumulate_psd (8f item 2)
  f3            firdecim_crcf / firinterp_crcf                           f4       resamp2_crcf modes, msresamp2_crcf cascades
  f5            fftfilt_crcf                                             f6       ofdmframegen data symbols (8f item 4)
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="*", default=["c1", "c2", "c4", "c5"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    args = ap.parse_args()
    run(args.configs, args.steps, args.warmup, lambda line: print(json.dumps(line), flush=True))


def run(configs, steps, warmup, sink, device_index=0):
    """measure `configs` (names as in the module docstring) on the current process's GPU; every result line (a dict)
    goes to `sink` -- bench.py collects the BASELINE configs this way for its `other_configs` entry"""
    import types
    args = types.SimpleNamespace(configs=list(configs), steps=steps, warmup=warmup)
    import torch
    import liquid_dsp_b200 as L
    L.set_device(device_index)
    dev = torch.device("cuda", device_index)
    st = torch.cuda.current_stream()
    peak = 6531.6
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = L.launch_count()
        e0.record(st)
        for _ in range(args.steps):
            fn()
        e1.record(st)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps, (L.launch_count() - l0) / args.steps

    def report(name, n_in, bytes_per_sample, ms, launches, extra=None):
        gbs = bytes_per_sample * n_in / (ms * 1e-3) / 1e9
        line = {"config": name, "metric": "input throughput", "value": n_in / (ms * 1e-3) / 1e6, "unit": "MS/s",
                "ms_per_step": ms, "launches_per_step": launches,
                "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                             "algorithmic_bytes_per_sample": bytes_per_sample}}
        if extra:
            line.update(extra)
        sink(line)

    if "c1" in args.configs:
        # firpfbch_crcf analysis M=16, m=4 (p=8), one long stream (BASELINE configs[0] shape, GPU-sized)
        M, m, n = 16, 4, 1 << 28
        q = L.firpfbch_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
        x = torch.empty(2 * n, dtype=torch.float32, device=dev)
        y = torch.empty(2 * n, dtype=torch.float32, device=dev)
        L.synth_input_dev(x.data_ptr(), n, stream=st.cuda_stream)
        ms, ln = timed(lambda: q.execute_block_dev(x.data_ptr(), n // M, y.data_ptr(), st.cuda_stream))
        report("c1: firpfbch_crcf analysis M=16 m=4, 2^28 samples", n, 16, ms, ln)
        del x, y
    if "c2" in args.configs:
        # firfilt_crcf_execute_block, 64-tap Kaiser low-pass, 1024 independent streams x 2^20 samples
        ns, n = 1024, 1 << 20
        q = L.firfilt_crcf.create_kaiser(64, 0.25, 60.0, 0.0)
        x = torch.empty(2 * ns * n, dtype=torch.float32, device=dev)
        y = torch.empty(2 * ns * n, dtype=torch.float32, device=dev)
        L.synth_input_dev(x.data_ptr(), ns * n, stream=st.cuda_stream)
        ms, ln = timed(lambda: q.execute_streams_dev(x.data_ptr(), ns, n, y.data_ptr(), st.cuda_stream))
        fp_peak = 148 * 4 * 0.5 * 32 * 1.92e9 / 64 / 1e6      # MS/s at 100 % FMA-pipe use (64 packed FMA per sample)
        report("c2: firfilt_crcf 64 taps, 1024 streams x 2^20 samples", ns * n, 16, ms, ln,
               {"fp32_pipe_bound_MSps": fp_peak})
        del x, y
    if "ff" in args.configs:
        # firfilt_crcf block form across filter lengths (256 streams x 2^20 samples): which lengths sit on which kernel
        for nt in (8, 16, 24, 32, 48, 64, 65, 96, 128, 160, 192, 256, 257):
            ns, n = 256, 1 << 20
            q = L.firfilt_crcf.create_kaiser(nt, 0.25, 60.0, 0.0)
            x = torch.empty(2 * ns * n, dtype=torch.float32, device=dev)
            y = torch.empty(2 * ns * n, dtype=torch.float32, device=dev)
            L.synth_input_dev(x.data_ptr(), ns * n, stream=st.cuda_stream)
            ms, ln = timed(lambda: q.execute_streams_dev(x.data_ptr(), ns, n, y.data_ptr(), st.cuda_stream))
            fp_peak = 148 * 4 * 0.5 * 32 * 1.92e9 / nt / 1e6      # MS/s at 100 % FMA-pipe use (nt packed FMA per sample)
            report("ff: firfilt_crcf %d taps, 256 streams x 2^20 samples" % nt, ns * n, 16, ms, ln, {"fp32_pipe_bound_MSps": fp_peak})
            del x, y
    if "c4" in args.configs:
        # firpfbch_crcf synthesis + analysis round trip, M=256, m=5, 32 streams x 2^22 samples on this GPU
        M, m, ns, n = 256, 5, 32, 1 << 22
        qs = [L.firpfbch_crcf.create_kaiser(L.LIQUID_SYNTHESIZER, M, m, 60.0) for _ in range(ns)]
        qa = [L.firpfbch_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0) for _ in range(ns)]
        x = torch.empty(2 * ns * n, dtype=torch.float32, device=dev)
        t = torch.empty(2 * ns * n, dtype=torch.float32, device=dev)
        y = torch.empty(2 * ns * n, dtype=torch.float32, device=dev)
        L.synth_input_dev(x.data_ptr(), ns * n, stream=st.cuda_stream)

        def rt():
            # all streams of each half in one launch (stateless multi-stream form)
            qs[0].execute_streams_dev(x.data_ptr(), ns, n // M, t.data_ptr(), st.cuda_stream)
            qa[0].execute_streams_dev(t.data_ptr(), ns, n // M, y.data_ptr(), st.cuda_stream)
        ms, ln = timed(rt)
        report("c4: firpfbch_crcf synthesis+analysis round trip M=256 m=5, 32 streams x 2^22 samples", ns * n, 32, ms, ln)
        ms, ln = timed(lambda: qa[0].execute_streams_dev(t.data_ptr(), ns, n // M, y.data_ptr(), st.cuda_stream))
        report("c4a: analysis half only (firpfbch_crcf analysis M=256 m=5, 32 streams x 2^22 samples)", ns * n, 16, ms, ln)
        del x, t, y
    if "c5" in args.configs:
        # firpfbch2_crcf analysis M=4096, m=9 on a 2^29-sample segment (one GPU's share of the sharded stream)
        M, m, n = 4096, 9, 1 << 29
        q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
        H = q.get_state_len()
        x = torch.zeros(2 * (H + n), dtype=torch.float32, device=dev)
        y = torch.empty(4 * n, dtype=torch.float32, device=dev)
        L.synth_input_dev(x.data_ptr() + 8 * H, n, stream=st.cuda_stream)
        ms, ln = timed(lambda: q.execute_segment_dev(x.data_ptr(), x.data_ptr() + 8 * H, 0, n // (M // 2), y.data_ptr(), st.cuda_stream))
        report("c5: firpfbch2_crcf analysis M=4096 m=9, 2^29-sample segment", n, 24, ms, ln)
    if "c3m" in args.configs:
        # the headline kernel at other filter lengths (m = 2 is the reference's own benchmark shape,
        # firpfbch2_crcf_benchmark.c:75): how the rate moves with the MAC work per sample
        for m in (2, 4, 5, 7):
            M, n = 1024, 1 << 29
            q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
            H = q.get_state_len()
            x = torch.zeros(2 * (H + n), dtype=torch.float32, device=dev)
            y = torch.empty(4 * n, dtype=torch.float32, device=dev)
            L.synth_input_dev(x.data_ptr() + 8 * H, n, stream=st.cuda_stream)
            ms, ln = timed(lambda: q.execute_segment_dev(x.data_ptr(), x.data_ptr() + 8 * H, 0, n // (M // 2), y.data_ptr(), st.cuda_stream))
            report("c3m: firpfbch2_crcf analysis M=1024 m=%d, 2^29 samples" % m, n, 24, ms, ln)
            del x, y
    if "m36" in args.configs:
        # m = 3 and m = 6 on the dedicated kernels (instantiated in round 2; these shapes ran on the shape-generic path before)
        for M in (16, 256, 1024):
            for m in (3, 6):
                n = 1 << 27
                q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
                H = q.get_state_len()
                x = torch.zeros(2 * (H + n), dtype=torch.float32, device=dev)
                y = torch.empty(4 * n, dtype=torch.float32, device=dev)
                L.synth_input_dev(x.data_ptr() + 8 * H, n, stream=st.cuda_stream)
                ms, ln = timed(lambda: q.execute_segment_dev(x.data_ptr(), x.data_ptr() + 8 * H, 0, n // (M // 2), y.data_ptr(), st.cuda_stream))
                report("m36: firpfbch2_crcf analysis M=%d m=%d, 2^27 samples" % (M, m), n, 24, ms, ln)
                q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_SYNTHESIZER, M, m, 60.0)
                ms, ln = timed(lambda: q.execute_block_dev(y.data_ptr(), (n // 2) // (M // 2), x.data_ptr(), st.cuda_stream))
                report("m36: firpfbch2_crcf synthesis M=%d m=%d, 2^26 output samples" % (M, m), n // 2, 24, ms, ln)
                for typ, name in ((L.LIQUID_ANALYZER, "analysis"), (L.LIQUID_SYNTHESIZER, "synthesis")):
                    q = L.firpfbch_crcf.create_kaiser(typ, M, m, 60.0)
                    ms, ln = timed(lambda: q.execute_block_dev(x.data_ptr(), n // M, y.data_ptr(), st.cuda_stream))
                    report("m36: firpfbch_crcf %s M=%d m=%d, 2^27 samples" % (name, M, m), n, 16, ms, ln)
                del x, y
    if "c5m" in args.configs:
        # the cluster kernel (M=4096) at the filter lengths it is instantiated for
        for m in (2, 3, 4, 5, 6, 7, 8, 9):
            M, n = 4096, 1 << 28
            q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
            H = q.get_state_len()
            x = torch.zeros(2 * (H + n), dtype=torch.float32, device=dev)
            y = torch.empty(4 * n, dtype=torch.float32, device=dev)
            L.synth_input_dev(x.data_ptr() + 8 * H, n, stream=st.cuda_stream)
            ms, ln = timed(lambda: q.execute_segment_dev(x.data_ptr(), x.data_ptr() + 8 * H, 0, n // (M // 2), y.data_ptr(), st.cuda_stream))
            report("c5m: firpfbch2_crcf analysis M=4096 m=%d, 2^28 samples" % m, n, 24, ms, ln)
            del x, y
    if "s" in args.configs:
        # synthesizers (a4, a6): firpfbch2 M=1024 m=7 (2^27 output samples) and firpfbch M=1024 m=2 / M=256 m=5
        M, m, n = 1024, 7, 1 << 27                      # n = output samples; input = 2 n channel samples
        q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_SYNTHESIZER, M, m, 60.0)
        X = torch.empty(4 * n, dtype=torch.float32, device=dev)
        y = torch.empty(2 * n, dtype=torch.float32, device=dev)
        L.synth_input_dev(X.data_ptr(), 2 * n, stream=st.cuda_stream)
        ms, ln = timed(lambda: q.execute_block_dev(X.data_ptr(), n // (M // 2), y.data_ptr(), st.cuda_stream))
        report("s: firpfbch2_crcf synthesis M=1024 m=7, 2^27 output samples", n, 24, ms, ln,
               {"note": "throughput in OUTPUT samples/s; algorithmic bytes 16 in + 8 out per output sample"})
        del X, y
        for M, m in ((1024, 2), (256, 5), (64, 2), (128, 2), (512, 2), (512, 5)):
            n = 1 << 27
            q = L.firpfbch_crcf.create_kaiser(L.LIQUID_SYNTHESIZER, M, m, 60.0)
            X = torch.empty(2 * n, dtype=torch.float32, device=dev)
            y = torch.empty(2 * n, dtype=torch.float32, device=dev)
            L.synth_input_dev(X.data_ptr(), n, stream=st.cuda_stream)
            ms, ln = timed(lambda: q.execute_block_dev(X.data_ptr(), n // M, y.data_ptr(), st.cuda_stream))
            report("s: firpfbch_crcf synthesis M=%d m=%d, 2^27 samples (one stream, stateful)" % (M, m), n, 16, ms, ln)
            del X, y
    if "g" in args.configs:
        # shapes without a dedicated fused kernel: the one-kernel shape-generic analyzers (power-of-two M)
        for M, m, n in ((256, 4, 1 << 27), (64, 5, 1 << 27), (2048, 3, 1 << 27)):
            q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
            H = q.get_state_len()
            x = torch.zeros(2 * (H + n), dtype=torch.float32, device=dev)
            y = torch.empty(4 * n, dtype=torch.float32, device=dev)
            L.synth_input_dev(x.data_ptr() + 8 * H, n, stream=st.cuda_stream)
            ms, ln = timed(lambda: q.execute_segment_dev(x.data_ptr(), x.data_ptr() + 8 * H, 0, n // (M // 2), y.data_ptr(), st.cuda_stream))
            report("g: firpfbch2_crcf analysis M=%d m=%d (shape-generic), 2^27 samples" % (M, m), n, 24, ms, ln)
            del x, y
    if "t" in args.configs:
        # the reference's own benchmark shapes below M = 1024 (firpfbch2_crcf_benchmark.c:70-82: m = 2) and m = 5 / 7:
        # the fused tile kernels of pfbch2_tile.cu
        for M, m in ((64, 2), (64, 5), (128, 5), (256, 2), (256, 5), (512, 2), (512, 5), (512, 7)):
            n = 1 << 28
            q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
            H = q.get_state_len()
            x = torch.zeros(2 * (H + n), dtype=torch.float32, device=dev)
            y = torch.empty(4 * n, dtype=torch.float32, device=dev)
            L.synth_input_dev(x.data_ptr() + 8 * H, n, stream=st.cuda_stream)
            ms, ln = timed(lambda: q.execute_segment_dev(x.data_ptr(), x.data_ptr() + 8 * H, 0, n // (M // 2), y.data_ptr(), st.cuda_stream))
            report("t: firpfbch2_crcf analysis M=%d m=%d (tile kernel), 2^28 samples" % (M, m), n, 24, ms, ln)
            del x, y
    if "fftc" in args.configs:
        # composite transform sizes (mixed-radix Stockham; LIQUID_CUDA_FFT_NO_MIXED=1 gives the table-driven DFT they ran on before)
        import numpy as np
        for nfft in (100, 300, 1000, 3000, 6000):
            npts = 1 << 26
            batch = npts // nfft
            xin, yout = np.zeros(nfft, np.complex64), np.zeros(nfft, np.complex64)
            p = L.fftplan(nfft, xin, yout, L.LIQUID_FFT_FORWARD)
            x = torch.empty(2 * batch * nfft, dtype=torch.float32, device=dev)
            y = torch.empty(2 * batch * nfft, dtype=torch.float32, device=dev)
            L.synth_input_dev(x.data_ptr(), batch * nfft, stream=st.cuda_stream)
            ms, ln = timed(lambda: p.execute_batch_dev(x.data_ptr(), y.data_ptr(), batch, st.cuda_stream))
            report("fftc: fft_execute_batch_dev n=%d, 2^26 points" % nfft, batch * nfft, 16, ms, ln)
            del x, y
    if "rb" in args.configs:
        # the reference's own benchmark list, shape by shape (src/multichannel/bench/firpfbch2_crcf_benchmark.c:70-90 analysis and
        # synthesis, firpfbch_crcf_benchmark.c:80-85 analysis; M = 4 ... 1024, m = 2), 2^27 samples each, one stateful stream
        for kind in ("firpfbch2 analysis", "firpfbch2 synthesis", "firpfbch analysis"):
            for M in (4, 16, 64, 256, 512, 1024):
                m, n = 2, 1 << 27
                if kind == "firpfbch2 analysis":
                    q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
                    n_in, n_out, nfr, bps = n, 2 * n, n // (M // 2), 24
                elif kind == "firpfbch2 synthesis":
                    q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_SYNTHESIZER, M, m, 60.0)
                    n_in, n_out, nfr, bps = 2 * n, n, n // (M // 2), 24
                else:
                    q = L.firpfbch_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
                    n_in, n_out, nfr, bps = n, n, n // M, 16
                x = torch.empty(2 * n_in, dtype=torch.float32, device=dev)
                y = torch.empty(2 * n_out, dtype=torch.float32, device=dev)
                L.synth_input_dev(x.data_ptr(), n_in, stream=st.cuda_stream)
                ms, ln = timed(lambda: q.execute_block_dev(x.data_ptr(), nfr, y.data_ptr(), st.cuda_stream))
                report("rb: %s M=%d m=2" % (kind, M), n, bps, ms, ln)
                q.destroy()
                del x, y
    if "sm" in args.configs:
        # the smallest shapes of the reference's benchmarks (firpfbch2_crcf_benchmark.c:70-90, firpfbch_crcf_benchmark.c:80-81), m = 2
        for kind, M in (("a2", 4), ("a2", 8), ("a2", 16), ("a2", 32), ("s2", 4), ("s2", 16), ("s2", 32), ("a1", 4), ("a1", 8), ("a1", 32)):
            m, n = 2, 1 << 27
            if kind == "a2":
                q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
                x = torch.empty(2 * n, dtype=torch.float32, device=dev)
                y = torch.empty(4 * n, dtype=torch.float32, device=dev)
                L.synth_input_dev(x.data_ptr(), n, stream=st.cuda_stream)
                ms, ln = timed(lambda: q.execute_block_dev(x.data_ptr(), n // (M // 2), y.data_ptr(), st.cuda_stream))
                report("sm: firpfbch2_crcf analysis M=%d m=2, 2^27 samples" % M, n, 24, ms, ln)
            elif kind == "s2":
                q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_SYNTHESIZER, M, m, 60.0)
                x = torch.empty(4 * n, dtype=torch.float32, device=dev)
                y = torch.empty(2 * n, dtype=torch.float32, device=dev)
                L.synth_input_dev(x.data_ptr(), 2 * n, stream=st.cuda_stream)
                ms, ln = timed(lambda: q.execute_block_dev(x.data_ptr(), n // (M // 2), y.data_ptr(), st.cuda_stream))
                report("sm: firpfbch2_crcf synthesis M=%d m=2, 2^27 output samples" % M, n, 24, ms, ln)
            else:
                q = L.firpfbch_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
                x = torch.empty(2 * n, dtype=torch.float32, device=dev)
                y = torch.empty(2 * n, dtype=torch.float32, device=dev)
                L.synth_input_dev(x.data_ptr(), n, stream=st.cuda_stream)
                ms, ln = timed(lambda: q.execute_block_dev(x.data_ptr(), n // M, y.data_ptr(), st.cuda_stream))
                report("sm: firpfbch_crcf analysis M=%d m=2, 2^27 samples" % M, n, 16, ms, ln)
            del x, y
    if "tp" in args.configs:
        # firpfbch_crcf analysis at the reference's own benchmark shapes (firpfbch_crcf_benchmark.c:80-85: m = 2) and m = 5
        for M, m in ((64, 2), (64, 5), (128, 2), (256, 2), (512, 2), (512, 5), (1024, 2)):
            n = 1 << 28
            q = L.firpfbch_crcf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
            x = torch.empty(2 * n, dtype=torch.float32, device=dev)
            y = torch.empty(2 * n, dtype=torch.float32, device=dev)
            L.synth_input_dev(x.data_ptr(), n, stream=st.cuda_stream)
            ms, ln = timed(lambda: q.execute_block_dev(x.data_ptr(), n // M, y.data_ptr(), st.cuda_stream))
            report("tp: firpfbch_crcf analysis M=%d m=%d, 2^28 samples (one stream, stateful)" % (M, m), n, 16, ms, ln)
            del x, y
    if "ts" in args.configs:
        # the reference's synthesizer benchmark shapes (firpfbch2_crcf_benchmark.c:84-90: m = 2) and m = 4 / 5:
        # the fused tile synthesizer of pfbch2_tile.cu; throughput in OUTPUT samples/s, 16 B in + 8 B out per output
        for M, m in ((64, 2), (64, 5), (128, 2), (256, 2), (256, 4), (256, 5), (512, 2), (512, 5), (1024, 2), (1024, 4)):
            n = 1 << 27
            q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_SYNTHESIZER, M, m, 60.0)
            X = torch.empty(4 * n, dtype=torch.float32, device=dev)
            y = torch.empty(2 * n, dtype=torch.float32, device=dev)
            L.synth_input_dev(X.data_ptr(), 2 * n, stream=st.cuda_stream)
            ms, ln = timed(lambda: q.execute_block_dev(X.data_ptr(), n // (M // 2), y.data_ptr(), st.cuda_stream))
            report("ts: firpfbch2_crcf synthesis M=%d m=%d (tile kernel), 2^27 output samples" % (M, m), n, 24, ms, ln)
            del X, y
    if "f1" in args.configs:
        # firpfbch_cccf analysis, M=256, p=10, one stream of 2^27 samples --
        # (a) a real prototype (create_kaiser): rides the crcf fused kernel; (b) complex taps: generic kernels
        import numpy as np
        M, m, n = 256, 5, 1 << 27
        x = torch.empty(2 * n, dtype=torch.float32, device=dev)
        y = torch.empty(2 * n, dtype=torch.float32, device=dev)
        L.synth_input_dev(x.data_ptr(), n, stream=st.cuda_stream)
        q = L.firpfbch_cccf.create_kaiser(L.LIQUID_ANALYZER, M, m, 60.0)
        ms, ln = timed(lambda: q.execute_block_dev(x.data_ptr(), n // M, y.data_ptr(), st.cuda_stream))
        report("f1a: firpfbch_cccf analysis M=256 m=5, real Kaiser prototype (crcf fused kernel), 2^27 samples", n, 16, ms, ln)
        rng = np.random.default_rng(1)
        h = ((rng.standard_normal(M * 2 * m) + 1j * rng.standard_normal(M * 2 * m)) / (2 * m)).astype(np.complex64)
        q = L.firpfbch_cccf.create(L.LIQUID_ANALYZER, M, 2 * m, h)
        ms, ln = timed(lambda: q.execute_block_dev(x.data_ptr(), n // M, y.data_ptr(), st.cuda_stream))
        report("f1b: firpfbch_cccf analysis M=256 p=10, complex taps (tile kernel, two FFMA2 per tap), 2^27 samples", n, 16, ms, ln)
        del x, y
    if "f3" in args.configs:
        # firdecim_crcf (M=8, m=6: 96 taps, 2^28 inputs) and firinterp_crcf (M=8, m=6, 2^25 inputs)
        M, m = 8, 6
        n = 1 << 28
        x = torch.empty(2 * n, dtype=torch.float32, device=dev)
        y = torch.empty(2 * n, dtype=torch.float32, device=dev)
        L.synth_input_dev(x.data_ptr(), n, stream=st.cuda_stream)
        q = L.firdecim_crcf.create_kaiser(M, m, 60.0)
        ms, ln = timed(lambda: q.execute_block_dev(x.data_ptr(), n // M, y.data_ptr(), st.cuda_stream))
        report("f3a: firdecim_crcf M=8 m=6 (96 taps), 2^28 input samples", n, 8 + 8 / M, ms, ln)
        ni = n // M
        qi = L.firinterp_crcf.create_kaiser(M, m, 60.0)
        ms, ln = timed(lambda: qi.execute_block_dev(x.data_ptr(), ni, y.data_ptr(), st.cuda_stream))
        report("f3b: firinterp_crcf M=8 m=6, 2^25 input samples (2^28 outputs)", ni, 8 + 8 * M, ms, ln,
               {"note": "throughput in INPUT samples/s; algorithmic bytes 8 in + 8 M out per input sample"})
        del x, y
    if "f4" in args.configs:
        # SURVEY 8(f) item 3, second part: resamp2_crcf decimator / interpolator (m=7: 14 taps on the filter branch)
        # and msresamp2_crcf decimating by 2^4, 2^28 input samples (interpolator: 2^27 in)
        n = 1 << 28
        x = torch.empty(2 * n, dtype=torch.float32, device=dev)
        y = torch.empty(2 * n, dtype=torch.float32, device=dev)
        L.synth_input_dev(x.data_ptr(), n, stream=st.cuda_stream)
        q = L.resamp2_crcf(7, 0.0, 60.0)
        ms, ln = timed(lambda: q.execute_block_dev("decim", x.data_ptr(), n // 2, y.data_ptr(), st.cuda_stream))
        report("f4a: resamp2_crcf decimator m=7, 2^28 input samples", n, 8 + 4, ms, ln)
        ms, ln = timed(lambda: q.execute_block_dev("interp", x.data_ptr(), n // 2, y.data_ptr(), st.cuda_stream))
        report("f4b: resamp2_crcf interpolator m=7, 2^27 input samples (2^28 outputs)", n // 2, 8 + 16, ms, ln,
               {"note": "throughput in INPUT samples/s; algorithmic bytes 8 in + 16 out per input sample"})
        ms, ln = timed(lambda: q.execute_block_dev("analyzer", x.data_ptr(), n // 2, y.data_ptr(), st.cuda_stream))
        report("f4c: resamp2_crcf analyzer m=7, 2^28 input samples", n, 16, ms, ln)
        qm = L.msresamp2_crcf(L.LIQUID_RESAMP_DECIM, 4, 0.4, 0.0, 60.0)
        ms, ln = timed(lambda: qm.execute_block_dev(x.data_ptr(), n // 16, y.data_ptr(), st.cuda_stream))
        report("f4d: msresamp2_crcf decimator 4 stages (m=%s), 2^28 input samples" % qm.stage_semilengths(), n,
               8 + 8 / 16, ms, ln, {"note": "algorithmic bytes: input read once, output written once; the "
                                            "intermediate streams of the stage-by-stage form are extra traffic"})
        qi = L.msresamp2_crcf(L.LIQUID_RESAMP_INTERP, 4, 0.4, 0.0, 60.0)
        ms, ln = timed(lambda: qi.execute_block_dev(x.data_ptr(), n // 16, y.data_ptr(), st.cuda_stream))
        report("f4e: msresamp2_crcf interpolator 4 stages, 2^24 input samples (2^28 outputs)", n // 16, 8 + 8 * 16, ms, ln,
               {"note": "throughput in INPUT samples/s; algorithmic bytes 8 in + 128 out per input sample"})
        del x, y
    if "f5" in args.configs:
        # fftfilt_crcf, 2^28 samples: 1025 taps in blocks of 1024, 257 taps in blocks of 256
        import numpy as np
        n_s = 1 << 28
        x = torch.empty(2 * n_s, dtype=torch.float32, device=dev)
        y = torch.empty(2 * n_s, dtype=torch.float32, device=dev)
        L.synth_input_dev(x.data_ptr(), n_s, stream=st.cuda_stream)
        for h_len, n in ((129, 128), (257, 256), (513, 512), (1025, 1024), (2049, 2048)):
            h = L.liquid_firdes_kaiser(h_len, 0.2, 60.0, 0.0)
            q = L.fftfilt_crcf(h, n)
            ms, ln = timed(lambda: q.execute_block_dev(x.data_ptr(), n_s // n, y.data_ptr(), st.cuda_stream))
            report("f5: fftfilt_crcf %d taps, blocks of %d (2n = %d point transforms), 2^28 samples" % (h_len, n, 2 * n),
                   n_s, 16, ms, ln)
        del x, y
    if "f6" in args.configs:
        # ofdmframegen data symbols, default allocation, 2^27 subcarrier values per step
        n_s = 1 << 27
        x = torch.empty(2 * n_s, dtype=torch.float32, device=dev)
        L.synth_input_dev(x.data_ptr(), n_s, stream=st.cuda_stream)
        for M, cp, tl in ((64, 16, 4), (256, 32, 8), (512, 64, 8), (1024, 128, 16), (2048, 256, 32), (4096, 512, 64)):
            y = torch.empty(2 * (n_s // M) * (M + cp), dtype=torch.float32, device=dev)
            q = L.ofdmframegen(M, cp, tl)
            ms, ln = timed(lambda: q.writesymbols_dev(x.data_ptr(), n_s // M, y.data_ptr(), st.cuda_stream))
            report("f6: ofdmframegen M=%d cp=%d taper=%d, 2^27 subcarrier values (%d symbols)" % (M, cp, tl, n_s // M),
                   n_s, 8 + 8 * (M + cp) / M, ms, ln,
                   {"note": "throughput in subcarrier values/s; algorithmic bytes 8 in + 8 (M + cp)/M out each"})
            del y
        del x
    if "p" in args.configs:
        # a13: firpfb_crcf (M = 32 filters, m = 6: 12 taps each), a new sample and an arbitrary filter index per output
        # (the symbol-timing use), 2^28 samples
        n_s = 1 << 28
        x = torch.empty(2 * n_s, dtype=torch.float32, device=dev)
        y = torch.empty(2 * n_s, dtype=torch.float32, device=dev)
        L.synth_input_dev(x.data_ptr(), n_s, stream=st.cuda_stream)
        idx = (torch.arange(n_s, device=dev, dtype=torch.int64) * 7 % 32).to(torch.int32)
        q = L.firpfb_crcf.create_kaiser(32, 6, 0.5, 60.0)
        ms, ln = timed(lambda: q.execute_block_dev(x.data_ptr(), idx.data_ptr(), n_s, y.data_ptr(), st.cuda_stream))
        report("p: firpfb_crcf M=32 m=6 (12 taps per output, filter index per sample), 2^28 samples", n_s, 20, ms, ln,
               {"note": "algorithmic bytes 8 in + 4 index + 8 out per sample"})
        del x, y, idx
    if "d" in args.configs:
        # a8: dotprod_crcf as an object: 2^22 / 2^18 independent dot products of 64 / 1024 taps over disjoint vectors
        import numpy as np
        n_s = 1 << 28
        x = torch.empty(2 * n_s, dtype=torch.float32, device=dev)
        L.synth_input_dev(x.data_ptr(), n_s, stream=st.cuda_stream)
        for nt in (8, 16, 32, 64, 128, 256, 1024):
            y = torch.empty(2 * (n_s // nt), dtype=torch.float32, device=dev)
            q = L.dotprod_crcf(L.liquid_firdes_kaiser(nt, 0.2, 60.0, 0.0))
            ms, ln = timed(lambda: q.execute_batch_dev(x.data_ptr(), n_s // nt, nt, y.data_ptr(), st.cuda_stream))
            report("d: dotprod_crcf_execute_batch_dev %d taps, %d vectors (2^28 samples)" % (nt, n_s // nt), n_s, 8 + 8 / nt, ms, ln)
            del y
        del x
    if "fft" in args.configs:
        # a11: the batched complex FFT on its own (fft_execute_batch_dev), 2^28 points per step, out of place
        import numpy as np
        for nn in (4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192):
            n = 1 << 28
            x = torch.empty(2 * n, dtype=torch.float32, device=dev)
            y = torch.empty(2 * n, dtype=torch.float32, device=dev)
            L.synth_input_dev(x.data_ptr(), n, stream=st.cuda_stream)
            hx, hy = np.zeros(nn, np.complex64), np.zeros(nn, np.complex64)
            p = L.fftplan(nn, hx, hy, L.LIQUID_FFT_FORWARD)
            ms, ln = timed(lambda: p.execute_batch_dev(x.data_ptr(), y.data_ptr(), n // nn, st.cuda_stream))
            report("fft: fft_execute_batch_dev n=%d, 2^28 points" % nn, n, 16, ms, ln)
            del x, y
    if "f2" in args.configs:
        # spgramcf_accumulate_psd on a device-resident block: default object (Kaiser window
        # nfft/2, a transform every nfft/4 samples ... window_len/2), 2^28 samples
        for nfft in (256, 1024, 4096):
            n = 1 << 28
            x = torch.empty(2 * n, dtype=torch.float32, device=dev)
            L.synth_input_dev(x.data_ptr(), n, stream=st.cuda_stream)
            q = L.spgramcf.create_default(nfft)
            ms, ln = timed(lambda: q.accumulate_psd_dev(x.data_ptr(), 0.01, n, st.cuda_stream))
            report("f2: spgramcf_accumulate_psd nfft=%d window=%d (one transform per %d samples), 2^28 samples"
                   % (nfft, nfft // 2, nfft // 4), n, 8, ms, ln,
                   {"transforms_per_step": n // (nfft // 4), "note": "algorithmic bytes: the input, read once"})
            del x


if __name__ == "__main__":
    main()
