#!/usr/bin/env python
"""Summaries of ncu output for profiles/: per-kernel share of a launch list (CSV of
`ncu --metrics gpu__time_duration.sum`) and selected metrics of a `--set full` report
(`ncu -i rep --page raw --csv`).

    python profiles_tool.py launches gpurun_out/launches.csv
    python profiles_tool.py raw gpurun_out/prof_raw.csv
"""
import collections
import csv
import re
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg", "sm__cycles_elapsed.max"]


def short(name):
    name = re.sub(r"^void (hmg::)?", "", name)
    return re.sub(r"\(.*$", "", name)


def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    t, n = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        k = short(r[ik])
        t[k] += float(r[iv].replace(",", ""))
        n[k] += 1
    tot = sum(t.values())
    unit = rows[1][hdr.index("Metric Unit")]
    print(f"{'kernel':60s} {'launches':>8s} {'total ' + unit:>14s} {'avg':>10s} {'share':>7s}")
    for k, v in t.most_common():
        print(f"{k:60s} {n[k]:8d} {v:14.1f} {v / n[k]:10.1f} {100 * v / tot:6.1f}%")
    print(f"{'all':60s} {sum(n.values()):8d} {tot:14.1f}")


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("-----", short(r[hdr.index("Kernel Name")]))
        for i, h in enumerate(hdr):
            stall = "issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(r[i] or 0) > 0.2
            if h in KEYS or stall:
                print(f"  {h:80s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2])
