#!/usr/bin/env python
"""Top source lines of a kernel by warp-stall samples:  ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv;
python tools/ncu_hot_lines.py src.csv [N]"""
import csv
import sys

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
        out.append((int(d["# Samples"] or 0), cur, int(r[0]), r[1][:100], d.get("stall_long_sb"), d.get("stall_barrier"),
                    d.get("stall_short_sb"), d.get("stall_wait"), d.get("Instructions Executed")))
tot = sum(o[0] for o in out) or 1
print("# total samples %d" % tot)
print("# samples share file:line long_sb barrier short_sb wait inst | source")
for o in sorted(out, reverse=True)[:top]:
    print("%6d %5.1f%% %s:%d lsb=%s bar=%s ssb=%s wait=%s inst=%s | %s" % (o[0], 100 * o[0] / tot, o[1], o[2], o[4], o[5], o[6], o[7], o[8], o[3]))
