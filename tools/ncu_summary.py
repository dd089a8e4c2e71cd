#!/usr/bin/env python
"""Summarise ncu output for profiles/: a launch list (csv from
`ncu --metrics gpu__time_duration.sum --csv`) and/or a full capture
(.ncu-rep from `ncu --set full`).  Usage:
    tools/ncu_summary.py launches <launches.csv>
    tools/ncu_summary.py full <prof.ncu-rep>
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "launch__registers_per_thread", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__grid_size", "launch__block_size",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[ki][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# launch list: {path}  (unit {rows[1][ui]}; cold-cache, serialised: compare SHARES)")
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k:72s} n={c:4d} total={t:14.1f} avg={t / c:12.1f} share={100 * t / tot:5.1f}%")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"# full capture: {path}")
    for d in data:
        print("---", d[hdr.index("Kernel Name")][:90])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:66s} {d[i]:>18s} {units[i]}")
        stalls = [(float(d[i]), h) for i, h in enumerate(hdr)
                  if h.startswith("smsp__average_warp") and h.endswith("_per_issue_active.ratio")
                  and d[i] not in ("", "n/a")]
        for v, h in sorted(stalls, reverse=True)[:6]:
            print(f"  stall {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:40s} {v:8.3f} warps/issue")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
