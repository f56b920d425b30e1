#!/usr/bin/env python
"""Top source lines of a kernel by warp-stall samples and by executed instructions:
ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv;  python tools/ncu_hot_lines.py src.csv [N]"""
import csv
import sys


def num(x):
    try:
        return int(x)
    except (TypeError, ValueError):
        return 0


rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out, cur, hdr = [], "", None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 6 and r[0] == "Line No":
        hdr = r
    elif len(r) > 6 and r[0].isdigit():
        d = dict(zip(hdr, r))
        out.append(dict(samples=num(d["# Samples"]), file=cur, line=int(r[0]), src=r[1][:110], lsb=num(d.get("stall_long_sb")),
                        bar=num(d.get("stall_barrier")), ssb=num(d.get("stall_short_sb")), wait=num(d.get("stall_wait")),
                        inst=num(d.get("Instructions Executed")), wf=num(d.get("L1 Wavefronts Shared")),
                        wfx=num(d.get("L1 Wavefronts Shared Excessive"))))
tot = sum(o["samples"] for o in out) or 1
ti = sum(o["inst"] for o in out) or 1
print("# total samples %d, total warp instructions %d" % (tot, ti))
print("# samples share | inst share | file:line long_sb barrier short_sb wait smem_wavefronts(excess) | source")
for o in sorted(out, key=lambda o: -o["samples"])[:top]:
    print("%6d %5.1f%% | %9d %5.1f%% | %s:%d lsb=%d bar=%d ssb=%d wait=%d wf=%d(%d) | %s" % (
        o["samples"], 100 * o["samples"] / tot, o["inst"], 100 * o["inst"] / ti, o["file"], o["line"], o["lsb"], o["bar"], o["ssb"],
        o["wait"], o["wf"], o["wfx"], o["src"]))
