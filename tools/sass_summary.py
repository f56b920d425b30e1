#!/usr/bin/env python
"""SASS evidence of the Blackwell data-movement engine in the shipped library: per kernel, the number of
UTMALDG / UTMASTG (cp.async.bulk.tensor), UBLKCP (cp.async.bulk), SYNCS (mbarrier), LDGSTS and plain LDG / STG instructions.

    python tools/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "scarlet_b200", "lib", "libscarlet_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts, cur = collections.OrderedDict(), None
pat = re.compile(r"\b(UTMALDG|UTMASTG|UBLKCP|SYNCS|LDGSTS|LDG|STG|LDS|STS|UTMACMDFLUSH)\b")
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    if cur:
        m = pat.search(line)
        if m:
            counts[cur][m.group(1)] += 1
cols = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDG", "STG", "LDS", "STS"]
print("# cuobjdump -sass scarlet_b200/lib/libscarlet_b200.so (sm_100a), instruction counts per kernel")
print("# UTMALDG/UTMASTG = cp.async.bulk.tensor (TMA tile load/store), UBLKCP = cp.async.bulk, SYNCS = mbarrier operations")
print("%-78s " % "kernel" + " ".join("%8s" % c for c in cols))
for k, c in counts.items():
    if not any(c[x] for x in ("UTMALDG", "UTMASTG", "UBLKCP")) and "--all" not in sys.argv:
        continue
    print("%-78s " % k[:78] + " ".join("%8d" % c[x] for x in cols))
tot = collections.Counter()
for c in counts.values():
    tot.update(c)
print("# library totals: " + ", ".join("%s %d" % (x, tot[x]) for x in cols))
