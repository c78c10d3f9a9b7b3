#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source sass` dump: stall samples per reason, the hottest
instructions, and shared-memory instructions with excessive wavefronts.
usage: ncu -i X.ncu-rep --page source --csv --print-source sass > src.csv; python tools/ncu_src_summary.py src.csv [top]"""
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(open(path)))
    # find the header row
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    col = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]

    def f(r, name):
        try:
            return float(r[col[name]])
        except Exception:  # noqa: BLE001
            return 0.0
    total = sum(f(r, "# Samples") for r in body)
    print("instructions: %d   samples: %d" % (len(body), total))
    reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = {h: sum(f(r, h) for r in body) for h in reasons}
    print("stall samples by reason:")
    for h, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        if v > 0:
            print("  %-26s %8d  %5.1f%%" % (h, v, 100 * v / max(total, 1)))
    print("hottest instructions:")
    for idx, r in sorted(enumerate(body), key=lambda ir: -f(ir[1], "# Samples"))[:top]:
        main_reason = max(reasons, key=lambda h: f(r, h))
        print("  #%-5d %6d  %-22s %s" % (idx, f(r, "# Samples"), main_reason, r[col["Source"]].strip()[:90]))
    print("shared-memory instructions with excessive wavefronts:")
    ex = [(f(r, "L1 Wavefronts Shared Excessive"), f(r, "L1 Wavefronts Shared"), idx, r) for idx, r in enumerate(body)]
    for e, w, idx, r in sorted(ex, key=lambda t: -t[0])[:top]:
        if e > 0:
            print("  #%-5d excess %10d of %10d  %s" % (idx, e, w, r[col["Source"]].strip()[:90]))


if __name__ == "__main__":
    main()
