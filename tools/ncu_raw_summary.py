#!/usr/bin/env python
"""Key metrics of the first kernel in an `ncu --page raw --csv` dump, in the format of profiles/*_summary.txt.
usage: ncu -i X.ncu-rep --page raw --csv > raw.csv; python tools/ncu_raw_summary.py raw.csv"""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg",
        "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread", "launch__block_size",
        "launch__grid_size", "launch__cluster_size", "launch__cluster_max_active",
        "launch__shared_mem_per_block_dynamic", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    for k in KEYS:
        if k in col:
            print("%-80s %s %s" % (k, vals[col[k]], units[col[k]]))
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
            print("%-80s %s" % (h, vals[col[h]]))


if __name__ == "__main__":
    main()
