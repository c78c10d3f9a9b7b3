#!/usr/bin/env python
"""What the host link gives: pinned D2H, H2D and both at once (16 B out per 8 B in, the channelizer's ratio).

  python tools/pcie_probe.py                 one GPU (device 0)
  python tools/pcie_probe.py --all           every GPU of the box AT THE SAME TIME, one process per GPU started together:
                                             the host-side ceiling the multi-GPU end-to-end numbers run into
"""
import argparse
import os
import subprocess
import sys
import time


def probe(dev, start_at):
    import torch
    torch.cuda.set_device(dev)
    n = 1 << 30
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    h2 = torch.empty(n // 2, dtype=torch.uint8).pin_memory()
    d2 = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    h.copy_(d, non_blocking=True)
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    while time.time() < start_at:          # all processes begin their timed copies together
        time.sleep(0.001)
    out = []
    for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
        t0 = time.perf_counter()
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        out.append("%s %.1f GB/s" % (name, 5 * n / (time.perf_counter() - t0) / 1e9))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        with torch.cuda.stream(s1):
            h.copy_(d, non_blocking=True)
        with torch.cuda.stream(s2):
            d2.copy_(h2, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out.append("duplex d2h %.1f + h2d %.1f GB/s" % (5 * n / dt / 1e9, 5 * n / 2 / dt / 1e9))
    print("gpu %d: %s" % (dev, "; ".join(out)), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--all", action="store_true")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--start-at", type=float, default=0.0)
    args = ap.parse_args()
    if not args.all:
        probe(args.device, args.start_at)
        return
    import torch
    n = torch.cuda.device_count()
    start = time.time() + 25.0            # leaves time for the imports and the pinned allocations
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--device", str(i), "--start-at", str(start)])
             for i in range(n)]
    for p in procs:
        p.wait()
    print("%d GPUs probed concurrently" % n)


if __name__ == "__main__":
    main()
