#!/usr/bin/env python
"""bench.py -- headline benchmark of the channelizer hot path (BASELINE.json metric:
"firpfbch2 M=1024 input MS/s @1/2/4/8 B200; % of HBM roofline").

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W

A step is one pass of firpfbch2_crcf analysis (M=1024, m=7, As=60 dB Kaiser prototype) over one
2^30-sample complex stream segment per GPU (BASELINE.json configs[2]); inputs are synthetic and
already resident in HBM for `value`.  With N>1 the N segments are consecutive time segments of ONE
long stream: every step each rank sends the last 2mM input samples of its segment to its right
neighbour with NCCL send/recv (the filter-history halo), the interior frames run while the halo is in
flight -- all inside the C entry point firpfbch2_crcf_execute_sharded_dev (csrc/multi.cu); torch.distributed
only carries the rendezvous (the NCCL unique id), the barrier and the max-over-ranks of the timing.  Weak
scaling: per-GPU work is fixed.  After the timed region every rank > 0 checks its halo and its halo-dependent
frames bit for bit ("halo_parity"); a mismatch fails the run.

--workload c2 / c4 run BASELINE.json configs[1] / configs[3] instead, sharded by STREAM across the ranks (no
collective); at N=1 the default run also measures configs 1, 2, 4, 5 briefly after the timed region and
reports them under "other_configs" (same code as tools/bench_configs.py).

`e2e` is the same metric through the host-pointer C ABI call firpfbch2_crcf_execute_block with
pinned host buffers: H2D of the input and D2H of the channelised output inside the timed region.

Only the cpu_baseline leg and --impl reference execute anything under oracle/ (the compiled
reference in oracle/_ref when present, else the oracle port) -- as the baseline, never as the
thing measured.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

M, SEMI, AS = 1024, 7, 60.0
BYTES_PER_SAMPLE = 24            # algorithmic: 8 B read + 16 B written per input sample
# dram__bytes_read.sum + dram__bytes_write.sum of the fused kernel from one `ncu --set full` capture
# (profiles/r02_ncu_fused_v9_summary.txt: 1.090297 GB + 2.088802 GB for 2^27 samples) -> bytes per sample
NCU_DRAM_BYTES_PER_SAMPLE = (1.090297e9 + 2.088802e9) / (1 << 27)
NCU_SOURCE = "profiles/r02_ncu_fused_v9_summary.txt"
NCU_LOG2 = 27
KERNEL = "pfbch2_fused_1024_kernel<14,0> (fused polyphase MAC + 1024-point IFFT)"
METRIC = "firpfbch2_crcf analysis M=1024 m=7 input throughput"
UNIT = "MS/s"


def select_workload(name):
    """--workload c3 (default, the headline: BASELINE.json configs[2]) or c5 (configs[4]: M=4096, m=9,
    the wideband channelizer that is time-sharded across GPUs)"""
    global M, SEMI, NCU_DRAM_BYTES_PER_SAMPLE, NCU_SOURCE, NCU_LOG2, KERNEL, METRIC
    if name == "c5":
        M, SEMI = 4096, 9
        # profiles/r01_ncu_cluster_4096_v4_summary.txt: 4.304220 GB + 8.536757 GB for 2^29 samples
        NCU_DRAM_BYTES_PER_SAMPLE = (4.304220e9 + 8.536757e9) / (1 << 29)
        NCU_SOURCE, NCU_LOG2 = "profiles/r01_ncu_cluster_4096_v4_summary.txt", 29
        KERNEL = "pfbch2_cluster_4096_kernel<18,0> (cluster of 8 CTAs per frame: polyphase MAC + 64x64 IFFT over DSMEM)"
        METRIC = "firpfbch2_crcf analysis M=4096 m=9 input throughput"


def workload_name(n_samples, n_gpus):
    return ("firpfbch2_crcf analysis M=%d m=%d As=%g Kaiser, 2^%d complex samples per GPU"
            % (M, SEMI, AS, n_samples.bit_length() - 1)
            + ("" if n_gpus == 1 else ", %d consecutive time segments of one stream, NCCL halo exchange" % n_gpus))


# --------------------------------------------------------------------------------------------
# clocks sampled DURING the timed region (NVML in a background thread; nvidia-smi CLI is too slow
# for a region that lasts tens of milliseconds)
# --------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index, period=0.002):
        self.samples, self.reasons, self.stop_flag = [], set(), False
        self.period = period
        self.armed = False              # the thread is started before the timed region, samples count from arm()
        self.max_mhz = None
        self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:          # noqa: BLE001
            self.nv = None
            self.err = str(e)

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonGpuIdle", 0x1): "gpu_idle",
            getattr(nv, "nvmlClocksEventReasonApplicationsClocksSetting", 0x2): "applications_clocks_setting",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSyncBoost", 0x10): "sync_boost",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self.stop_flag:
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                if not self.armed:
                    time.sleep(self.period)
                    continue
                self.samples.append(mhz)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:   # noqa: BLE001
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:       # noqa: BLE001
                pass
            time.sleep(self.period)

    def start(self):
        if self.nv:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()

    def arm(self):
        self.armed = True

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "no NVML samples"}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# --------------------------------------------------------------------------------------------
# CPU baseline (reference implementation on the host cores) -- the one place oracle/ is executed
# --------------------------------------------------------------------------------------------
def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_cpu_reference(n_per_core, cores):
    """(MS/s aggregate, kind, sample description): `cores` independent instances, one per core,
    each on its own synthetic stream of n_per_core samples, started together"""
    from oracle import oracle as O
    if O.ref_available():
        secs = O.ref_bench(0, M, SEMI, AS, n_per_core, cores)
        kind = "reference"
    else:
        O.build(ref=False)
        secs = O.bench_firpfbch2_analysis(M, SEMI, AS, n_per_core, cores)
        kind = "port"
    value = cores * n_per_core / secs / 1e6
    sample = ("%d instances (one per host core), independent synthetic streams of 2^%d samples each, "
              "wall time of the slowest thread %.3f s" % (cores, n_per_core.bit_length() - 1, secs))
    return value, kind, sample


def impl_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = cpu_cores()
    n_per_core = 1 << args.ref_log2_samples
    for _ in range(args.warmup):
        run_cpu_reference(n_per_core // 4, cores)
    t_total, kind, sample = 0.0, None, None
    vals = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        v, kind, sample = run_cpu_reference(n_per_core, cores)
        t_total += time.perf_counter() - t0
        vals.append(v)
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(1 << 30, 1),
                   "note": "reference CPU implementation (liquid-dsp firpfbch2_crcf_execute, internal FFT, SSE dotprod) "
                           "on the host cores; each step is a bounded sample of the workload: " + sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------
# stream-sharded workloads (BASELINE.json configs[1] and configs[3]): no collective on the data path
# --------------------------------------------------------------------------------------------
def main_streams(args):
    import torch
    import torch.distributed as dist
    import liquid_dsp_b200 as L
    import liquid_dsp_b200.multi as multi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    L.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    st = torch.cuda.current_stream()
    if args.workload == "c2":
        total_streams, n, bytes_per_sample = 1024, 1 << 20, 16
        metric = "firfilt_crcf_execute_block 64-tap Kaiser low-pass input throughput"
        lo, hi = multi.streams_of_rank(total_streams, rank, world)
        ns = hi - lo
        q = L.firfilt_crcf.create_kaiser(64, 0.25, 60.0, 0.0)
        x = torch.empty(2 * ns * n, dtype=torch.float32, device=dev)
        y = torch.empty(2 * ns * n, dtype=torch.float32, device=dev)
        for s_i in range(ns):       # stream s of the job has its own seed
            L.synth_input_dev(x.data_ptr() + 8 * s_i * n, n, seed=0x5EED0000 + lo + s_i, stream=st.cuda_stream)

        def step():
            q.execute_streams_dev(x.data_ptr(), ns, n, y.data_ptr(), st.cuda_stream)
        workload = ("firfilt_crcf 64 taps (Kaiser, fc=0.25, As=60), %d independent streams x 2^20 samples, sharded by stream over "
                    "%d GPU(s): %d streams on rank 0" % (total_streams, world, ns))
        kernel = "firfilt_regblock_kernel<64,16>"
    else:
        total_streams, n, bytes_per_sample = 256, 1 << 22, 32
        Mc, mc = 256, 5
        metric = "firpfbch_crcf synthesis+analysis round trip M=256 m=5 input throughput"
        lo, hi = multi.streams_of_rank(total_streams, rank, world)
        ns = hi - lo
        qs = L.firpfbch_crcf.create_kaiser(L.LIQUID_SYNTHESIZER, Mc, mc, 60.0)
        qa = L.firpfbch_crcf.create_kaiser(L.LIQUID_ANALYZER, Mc, mc, 60.0)
        x = torch.empty(2 * ns * n, dtype=torch.float32, device=dev)
        t = torch.empty(2 * ns * n, dtype=torch.float32, device=dev)
        y = torch.empty(2 * ns * n, dtype=torch.float32, device=dev)
        for s_i in range(ns):
            L.synth_input_dev(x.data_ptr() + 8 * s_i * n, n, seed=0x5EED0000 + lo + s_i, stream=st.cuda_stream)

        def step():
            qs.execute_streams_dev(x.data_ptr(), ns, n // Mc, t.data_ptr(), st.cuda_stream)
            qa.execute_streams_dev(t.data_ptr(), ns, n // Mc, y.data_ptr(), st.cuda_stream)
        workload = ("firpfbch_crcf synthesis then analysis, M=256 m=5 As=60, %d independent streams x 2^22 samples, sharded by "
                    "stream over %d GPU(s): %d streams on rank 0" % (total_streams, world, ns))
        kernel = "pfbch_synth_fused_256_kernel<10> + pfbch_fused_256_kernel<10> (two API calls per step)"

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
    for _ in range(max(args.warmup, 0)):
        step()
    sync_all()
    sampler = ClockSampler(local_rank, period=0.002 if world == 1 else 0.010) if rank == 0 else None
    launches0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler:
        sampler.start()
    time.sleep(0.05)
    sync_all()
    if sampler:
        sampler.arm()
    e0.record(st)
    for _ in range(args.steps):
        step()
    e1.record(st)
    sync_all()
    clocks = sampler.stop() if sampler else None
    launches = L.launch_count() - launches0
    local_ms = e0.elapsed_time(e1)
    total_ms = local_ms
    if world > 1:
        tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    value = total_streams * n * args.steps / (total_ms * 1e-3) / 1e6
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else 6650.0
    achieved = bytes_per_sample * ns * n * args.steps / (local_ms * 1e-3) / 1e9
    if rank == 0:
        line = {"metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "l2": "inputs larger than L2 per GPU per step, no flush needed",
                           "timing": "CUDA events on the launching stream, max over ranks",
                           "sharding": "by stream, contiguous blocks per rank (liquid_dsp_b200.multi.streams_of_rank), no collective"},
                "clocks": clocks, "e2e": None, "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": None, "kernel": kernel, "algorithmic_bytes_per_sample": bytes_per_sample,
                             "note": "rank 0's share of the job / rank 0's event time"},
                "cpu_baseline": None}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2-samples", type=int, default=30, help="input samples per GPU per step (log2)")
    ap.add_argument("--e2e-log2-samples", type=int, default=28, help="input samples per e2e step (log2)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-log2-samples", type=int, default=24, help="cpu_baseline samples per core (log2)")
    ap.add_argument("--ref-log2-samples", type=int, default=22, help="--impl reference samples per core per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--debug-steps", action="store_true", help="N>1: print every rank's per-step times to stderr")
    ap.add_argument("--reserve-sms", type=int, default=2,
                    help="N>1 only: SMs the persistent kernel leaves to the NCCL kernels and the halo launch")
    ap.add_argument("--workload", default="c3", choices=["c3", "c5", "c2", "c4"],
                    help="c3 = headline (M=1024, m=7); c5 = BASELINE.json configs[4] (M=4096, m=9), both time-sharded; "
                         "c2 / c4 = configs[1] / configs[3], sharded by stream")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the brief runs of configs 1, 2, 4, 5 (N=1 only)")
    ap.add_argument("--other-steps", type=int, default=3)
    ap.add_argument("--python-exchange", action="store_true",
                    help="N>1: the round-1 path (torch.distributed send/recv driven from Python) instead of the C driver")
    args = ap.parse_args()
    if args.workload in ("c2", "c4"):
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "the reference arm times the headline workload only"}))
            return 0
        return main_streams(args)
    select_workload(args.workload)
    if args.impl == "reference":
        return impl_reference(args)

    import torch
    import torch.distributed as dist
    import liquid_dsp_b200 as L
    import liquid_dsp_b200.multi as multi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this benchmark has no CPU path (use --impl reference)")
    torch.cuda.set_device(local_rank)
    L.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # The halo is 112-576 KB: one NCCL channel (= one CTA per rank) moves it.  With the default channel
        # count the send/recv kernel is a grid of CTAs that must all be co-resident with their peers'; next
        # to a persistent kernel that owns (almost) every SM they are not, and ranks stall each other for a
        # whole kernel duration.  One channel + --reserve-sms 2 lets the exchange run beside the interior frames.
        os.environ.setdefault("NCCL_MAX_P2P_NCHANNELS", "1")
        os.environ.setdefault("NCCL_MIN_P2P_NCHANNELS", "1")
        dist.init_process_group("nccl", device_id=dev)

    n = 1 << args.log2_samples                     # input samples per GPU per step
    n_frames = n // (M // 2)
    q = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, SEMI, AS)
    H = q.get_state_len()                          # 2mM samples of history = the halo
    halo_frames = H // (M // 2)                    # 4m frames depend on the halo

    # device-resident input with H samples of headroom in front (halo lands there), output [n_frames][M]
    xbuf = torch.empty(2 * (H + n), dtype=torch.float32, device=dev)
    ybuf = torch.empty(2 * n_frames * M, dtype=torch.float32, device=dev)
    xbuf[: 2 * H].zero_()
    stream = torch.cuda.current_stream()
    d_halo = xbuf.data_ptr()
    d_x = d_halo + 8 * H
    d_y = ybuf.data_ptr()
    # rank r holds samples [r*n, (r+1)*n) of the one long synthetic stream
    L.synth_input_dev(d_x, n, seed=0x5EED0000, first=rank * n, stream=stream.cuda_stream)
    torch.cuda.synchronize()

    first_frame = rank * n_frames                  # even: every segment starts with flag = 0
    tail = xbuf[2 * n: 2 * (H + n)]                # last H samples of the local segment
    head = xbuf[: 2 * H]

    # N > 1: the halo exchange and the 4m halo-dependent frames live on a side stream, and the persistent
    # channelizer kernel leaves two SMs free (--reserve-sms), so that the NCCL send/recv kernels and the
    # small halo launch run BESIDE the interior frames instead of queueing behind a kernel that owns every SM
    side = torch.cuda.Stream(device=dev) if world > 1 else None
    comm = None
    if world > 1 and args.python_exchange:
        L.set_reserved_sms(args.reserve_sms)
    elif world > 1:
        # the C multi-GPU driver: rank 0 makes the NCCL unique id, torch.distributed carries it to the others
        ids = [L.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = L.Comm(rank, world, ids[0])
        if args.reserve_sms != 2:
            L.set_reserved_sms(args.reserve_sms)

    def step():
        if comm is not None:
            q.execute_sharded_dev(comm, d_halo, first_frame, n_frames, d_y, stream.cuda_stream)
        elif world > 1:
            with torch.cuda.stream(side):
                reqs = multi.exchange_halo(dist, head, tail, rank, world)      # NCCL send/recv of 2mM samples
            # interior frames do not see the halo: run them while it is in flight
            q.execute_segment_dev(d_x + 8 * (halo_frames * (M // 2)) - 8 * H, d_x + 8 * halo_frames * (M // 2),
                                  first_frame + halo_frames, n_frames - halo_frames,
                                  d_y + 8 * halo_frames * M, stream.cuda_stream)
            with torch.cuda.stream(side):
                for r in reqs:
                    r.wait()
                q.execute_segment_dev(d_halo, d_x, first_frame, halo_frames, d_y, side.cuda_stream)
            stream.wait_stream(side)                      # the step ends when both parts have
        else:
            q.execute_segment_dev(d_halo, d_x, first_frame, n_frames, d_y, stream.cuda_stream)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 0)):
        step()
    sync_all()

    # N > 1: rank 0 alone samples, and less often -- eight processes polling NVML every 2 ms contend for the
    # driver's locks and stall each other's kernel launches for milliseconds at a time
    sampler = ClockSampler(local_rank, period=0.002 if world == 1 else 0.010) if rank == 0 else None
    launches0 = L.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    # the sampler thread's start-up and first NVML call happen before the barrier, not inside the timed region
    if sampler:
        sampler.start()
    time.sleep(0.05)
    sync_all()                                           # every rank: barrier + device synchronize
    if sampler:
        sampler.arm()
    ev[0].record(stream)
    for i in range(args.steps):
        step()
        ev[i + 1].record(stream)
    sync_all()
    clocks = sampler.stop() if sampler else None
    launches = L.launch_count() - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    if world > 1:
        if args.debug_steps:
            print("rank %d step ms: %s" % (rank, " ".join("%.2f" % v for v in step_ms)), file=sys.stderr, flush=True)
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * n * args.steps / (total_ms * 1e-3) / 1e6          # MS/s, whole job

    # multi-GPU parity, checked on the buffers the timed steps left behind: (1) the head-room holds the left
    # neighbour's last 2mM samples bit for bit (the input is counter-based, so the rank regenerates them);
    # (2) the 4m halo-dependent frames equal, bit for bit, the same launch over the regenerated halo; (3) the
    # first 8m frames agree with ONE single-GPU segment that starts 4m frames inside the left neighbour's range
    # (different kernels take part -- the fused kernel there, the shape-generic one for the short halo launch --
    # so this last comparison is to 2e-6 of the output scale, not bitwise)
    halo_parity = None
    if world > 1:
        ok = True
        if rank > 0:
            z = torch.empty(2 * 4 * H, dtype=torch.float32, device=dev)          # stream[rank*n - 2H, rank*n + 2H)
            L.synth_input_dev(z.data_ptr(), 4 * H, seed=0x5EED0000, first=rank * n - 2 * H, stream=stream.cuda_stream)
            torch.cuda.synchronize()
            ok &= bool(torch.equal(z[2 * H: 4 * H].view(torch.int32), xbuf[: 2 * H].view(torch.int32)))
            yh = torch.empty(2 * halo_frames * M, dtype=torch.float32, device=dev)
            q.execute_segment_dev(z.data_ptr() + 8 * H, d_x, first_frame, halo_frames, yh.data_ptr(), stream.cuda_stream)
            torch.cuda.synchronize()
            ok &= bool(torch.equal(yh.view(torch.int32), ybuf[: 2 * halo_frames * M].view(torch.int32)))
            ym = torch.empty(2 * 3 * halo_frames * M, dtype=torch.float32, device=dev)
            q.execute_segment_dev(z.data_ptr(), z.data_ptr() + 8 * H, first_frame - halo_frames, 3 * halo_frames,
                                  ym.data_ptr(), stream.cuda_stream)
            torch.cuda.synchronize()
            ref = ym[2 * halo_frames * M:]
            got = ybuf[: 2 * 2 * halo_frames * M]
            ok &= bool((ref - got).abs().max().item() <= 2e-6 * max(ref.abs().max().item(), 1e-30))
            del z, yh, ym
        t = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        halo_parity = bool(t.item() > 0.5)

    # roofline of the dominant kernel: algorithmic bytes per launch / average launch duration.
    # One step on one GPU is one pass of the fused channelizer kernel(s) over n samples; the
    # duration is the CUDA-event time of the step on the launching stream (this rank).
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    avg_step_ms = sum(step_ms) / len(step_ms)
    achieved = BYTES_PER_SAMPLE * n / (avg_step_ms * 1e-3) / 1e9
    kernels_per_step = launches / max(args.steps, 1)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": NCU_DRAM_BYTES_PER_SAMPLE * n / 1e9, "traffic_unit": "GB per launch",
                "traffic_source": "ncu --set full capture at 2^%d samples per launch (%s), scaled per sample"
                                  % (NCU_LOG2, NCU_SOURCE),
                "peak_source": peak_src,
                "kernel": KERNEL + ", %.0f launch(es) per step" % kernels_per_step,
                "algorithmic_bytes_per_launch": BYTES_PER_SAMPLE * n,
                "avg_launch_ms": avg_step_ms, "frac_of_8TBps_nominal": achieved / 8000.0}

    # end to end through the host-pointer C ABI (pinned host buffers, copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        import ctypes as C
        import numpy as np
        ne = 1 << args.e2e_log2_samples
        nfe = ne // (M // 2)
        hx = L.pinned_empty(ne)
        hy = L.pinned_empty(nfe * M)
        # fill the pinned input with this rank's part of the synthetic stream (generated on the device, copied back)
        L.synth_input_dev(d_x, min(ne, n), seed=0x5EED0000, first=rank * ne, stream=stream.cuda_stream)
        torch.cuda.synchronize()
        L.lib.liquid_cuda_memcpy_d2h(hx.ctypes.data_as(C.c_void_p), d_x, min(ne, n) * 8)
        qe = L.firpfbch2_crcf.create_kaiser(L.LIQUID_ANALYZER, M, SEMI, AS)

        def e2e_step(x, y):
            # N > 1: the ranks' host arrays are consecutive segments of one stream; the C driver moves the halo
            if comm is not None:
                qe.execute_sharded(comm, x, rank * nfe, out=y)
            else:
                qe.reset()
                qe.execute_block(x, out=y)

        def e2e_time(x, y, steps):
            e2e_step(x, y)                                   # warm-up (allocates staging buffers)
            sync_all()
            t0 = time.perf_counter()
            for _ in range(steps):
                e2e_step(x, y)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            return world * ne * steps / dt / 1e6

        v_pinned = e2e_time(hx, hy, args.e2e_steps)
        # what the link gives on its own: one pinned D2H copy of the output / H2D copy of the input, nothing else running
        def copy_rate(nbytes, d2h):
            torch.cuda.synchronize()
            best = 0.0
            for _ in range(3):
                t0 = time.perf_counter()
                if d2h:
                    L.lib.liquid_cuda_memcpy_d2h(hy.ctypes.data_as(C.c_void_p), d_y, nbytes)
                else:
                    L.lib.liquid_cuda_memcpy_h2d(d_x, hx.ctypes.data_as(C.c_void_p), nbytes)
                best = max(best, nbytes / (time.perf_counter() - t0) / 1e9)
            return best
        probe_bytes = min(ne * 8, 1 << 30)
        sync_all()
        d2h_alone, h2d_alone = copy_rate(probe_bytes, True), copy_rate(probe_bytes, False)
        d2h_rate = (v_pinned / world) * 1e6 * 16 / 1e9       # GB/s of channelised output per GPU
        e2e = {"value": v_pinned, "unit": UNIT,
               "h2d_bytes_per_step": ne * 8, "d2h_bytes_per_step": nfe * M * 8,
               "samples_per_step_per_gpu": ne, "steps": args.e2e_steps,
               "api": ("firpfbch2_crcf_execute_sharded (host pointers, pinned; NCCL halo, then the chunked pipeline)"
                       if comm is not None else
                       "firpfbch2_crcf_execute_block (host pointers, pinned, chunked H2D/compute/D2H pipeline)"),
               "link": {"d2h_alone_GBps": d2h_alone, "h2d_alone_GBps": h2d_alone, "d2h_achieved_GBps_per_gpu": d2h_rate,
                        "note": "all ranks probe at the same time" if world > 1 else "single copy, idle machine"},
               "link_frac": d2h_rate / d2h_alone if d2h_alone > 0 else None}
        if world == 1:
            # a caller's malloc'd arrays: plain pageable copies (default), page-locked for the duration of each call, or
            # page-locked once by the caller (liquid_cuda_host_register) and reused
            px, py = np.empty(ne, np.complex64), np.empty(nfe * M, np.complex64)
            px[:] = hx
            py[:] = 0
            pageable = {}
            for mode in ("register", "direct"):
                os.environ["LIQUID_CUDA_PAGEABLE"] = mode
                pageable[mode] = e2e_time(px, py, max(1, args.e2e_steps - 1))
            os.environ.pop("LIQUID_CUDA_PAGEABLE", None)
            L.lib.liquid_cuda_host_register(px.ctypes.data_as(C.c_void_p), px.nbytes)
            L.lib.liquid_cuda_host_register(py.ctypes.data_as(C.c_void_p), py.nbytes)
            v_reg = e2e_time(px, py, max(1, args.e2e_steps - 1))
            L.lib.liquid_cuda_host_unregister(px.ctypes.data_as(C.c_void_p))
            L.lib.liquid_cuda_host_unregister(py.ctypes.data_as(C.c_void_p))
            e2e["pageable_host_arrays"] = {"unit": UNIT, "plain_pageable_copies (default)": pageable["direct"],
                                           "page_locked_per_call (LIQUID_CUDA_PAGEABLE=register)": pageable["register"],
                                           "registered_once_by_the_caller (liquid_cuda_host_register)": v_reg}
            del px, py
        qe.destroy()
        L.pinned_free(hx)
        L.pinned_free(hy)

    # the other BASELINE configs, briefly (N = 1): same code as tools/bench_configs.py
    other = None
    if world == 1 and not args.no_other_configs and args.workload == "c3":
        del xbuf, ybuf
        torch.cuda.empty_cache()
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_configs
        other = {}

        def keep(line):
            other[line["config"].split(":")[0]] = {"config": line["config"], "value": line["value"], "unit": line["unit"],
                                                   "ms_per_step": line["ms_per_step"], "roofline_frac": line["roofline"]["frac"],
                                                   "bytes_per_sample": line["roofline"]["algorithmic_bytes_per_sample"]}
        for cfg in ("c1", "c2", "c4", "c5"):
            try:
                bench_configs.run([cfg], args.other_steps, 2, keep, device_index=local_rank)
            except Exception as e:          # noqa: BLE001
                other[cfg] = {"error": str(e)}
            torch.cuda.empty_cache()
        # the reference's own benchmark list (firpfbch2_crcf_benchmark.c / firpfbch_crcf_benchmark.c: M = 4 ... 1024, m = 2)
        shapes = {}
        try:
            bench_configs.run(["rb"], args.other_steps, 2,
                              lambda line: shapes.__setitem__(line["config"][4:], {"GSps": round(line["value"] / 1e3, 1),
                                                                                   "roofline_frac": round(line["roofline"]["frac"], 3)}),
                              device_index=local_rank)
        except Exception as e:              # noqa: BLE001
            shapes["error"] = str(e)
        other["reference_benchmark_shapes"] = {"note": "2^27 samples per shape, one stateful stream, same timing code as the configs above; "
                                                       "24 B per sample (firpfbch2) / 16 B per sample (firpfbch) against the measured HBM peak",
                                               "shapes": shapes}
        torch.cuda.empty_cache()
        # the kernels the channelizers reduce to or feed (SURVEY a8 / a9 / a11 / a13 and the 8(f) rows), shape by shape:
        # firfilt over filter lengths, dotprod batch, firpfb, the batched FFT over sizes, ofdmframegen, spgramcf, fftfilt
        callees = {}
        for cfg in ("ff", "d", "p", "fft", "f6", "f2", "f5"):
            try:
                bench_configs.run([cfg], args.other_steps, 2,
                                  lambda line: callees.__setitem__(line["config"], {"GSps": round(line["value"] / 1e3, 1),
                                                                                    "roofline_frac": round(line["roofline"]["frac"], 3),
                                                                                    "bytes_per_sample": line["roofline"]["algorithmic_bytes_per_sample"]}),
                                  device_index=local_rank)
            except Exception as e:          # noqa: BLE001
                callees[cfg] = {"error": str(e)}
            torch.cuda.empty_cache()
        other["callees_and_next_rows"] = callees

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = cpu_cores()
        v, kind, sample = run_cpu_reference(1 << args.cpu_log2_samples, cores)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(n, world),
                       "l2": "inputs larger than L2: %d MiB read + %d MiB written per GPU per step, no flush needed"
                             % (n * 8 >> 20, n * 16 >> 20),
                       "timing": "CUDA events on the launching stream, max over ranks"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu_baseline,
        }
        if halo_parity is not None:
            line["halo_parity"] = halo_parity
            line["config"]["exchange"] = ("torch.distributed batch_isend_irecv (Python)" if comm is None else
                                          "firpfbch2_crcf_execute_sharded_dev: ncclSend/ncclRecv of 2mM samples inside the C driver")
        if other is not None:
            line["other_configs"] = other
        print(json.dumps(line))
    if comm is not None:
        comm.destroy()
    if world > 1:
        dist.destroy_process_group()
    if halo_parity is False:
        print("bench.py: halo parity check FAILED", file=sys.stderr)
        return 3
    return 0


if __name__ == "__main__":
    sys.exit(main())
