#!/usr/bin/env python
"""Summarise ncu output for profiles/:  launch list CSV -> per-kernel table;  .ncu-rep -> selected raw metrics.

    python tools/ncu_summary.py launches gpurun_out/launches.csv
    python tools/ncu_summary.py full gpurun_out/prof.ncu-rep
"""
import collections
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        a = agg.setdefault(r[ki][:90], [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print("# per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none` (unit: %s; cold-cache, serialised)" % rows[hdr + 1][ui])
    print("%-92s %6s %14s %12s %7s" % ("kernel", "n", "total", "avg", "share"))
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-92s %6d %14.0f %12.0f %7.3f" % (k, n, t, t / n, t / tot))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]
    names = [r[h.index("Kernel Name")][:60] for r in rows[2:]]
    print("# selected metrics of `ncu --set full --clock-control none` capture %s" % path)
    print("# launches:", names)
    for w in WANT:
        if w in h:
            i = h.index(w)
            print("%-80s %-14s %s" % (w, rows[1][i], [r[i] for r in rows[2:]]))


def traffic(path):
    """JSON of DRAM bytes (read + write) and duration per launch, per kernel, for bench.py's roofline.traffic."""
    import json
    import re
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]

    def col(name, r):
        i = h.index(name)
        v = float(r[i].replace(",", ""))
        u = units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "second": 1,
                    "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9}.get(u, 1)

    agg = {}
    for r in rows[2:]:
        name = re.sub(r"^void |\(.*$", "", r[h.index("Kernel Name")])
        name = re.sub(r"<.*", "", name).replace("sb::", "")
        a = agg.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "seconds": 0.0})
        a["launches"] += 1
        a["dram_bytes"] += col("dram__bytes_read.sum", r) + col("dram__bytes_write.sum", r)
        a["seconds"] += col("gpu__time_duration.sum", r)
    for a in agg.values():
        a["dram_bytes_per_launch"] = a.pop("dram_bytes") / a["launches"]
        a["us_per_launch_under_ncu"] = 1e6 * a.pop("seconds") / a["launches"]
    print(json.dumps({"source": path, "kernels": agg}, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2])
