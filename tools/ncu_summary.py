#!/usr/bin/env python
"""Summarise ncu output for profiles/: (1) a launch-list CSV (--metrics gpu__time_duration.sum[,dram bytes]) -> per-kernel
totals and shares; (2) a --set full .ncu-rep -> the handful of metrics the roofline discussion in DESIGN.md uses."""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = rows[0]
    ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    per = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        per.setdefault(r[ii], {"name": r[ki]})[r[mi]] = v
    agg = collections.OrderedDict()
    for d in per.values():
        name = d["name"].split("(")[0].replace("psacb200::", "")[:70]
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    tot = sum(a[1] for a in agg.values())
    print("%-72s %5s %12s %7s %14s" % ("kernel", "n", "time", "share", "dram bytes"))
    for k, a in agg.items():
        print("%-72s %5d %12.0f %6.1f%% %14.0f" % (k, a[0], a[1], 100 * a[1] / tot, a[2]))
    print("total %.0f (ncu units as in the csv: time ns unless stated; bytes as reported)" % tot)


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_warps", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio" ]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== %s" % r[h.index("Kernel Name")][:100])
        for w in WANT:
            if w in h:
                i = h.index(w)
                print("  %-72s %s %s" % (w, r[i], units[i]))
        for i, name in enumerate(h):
            if "warp_issue_stalled" in name and name.endswith("_per_warp_active.pct"):
                try:
                    if float(r[i]) >= 3.0:
                        print("  %-72s %s %s" % (name, r[i], units[i]))
                except ValueError:
                    pass


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print("#", p)
        (rep if p.endswith(".ncu-rep") else launches)(p)
