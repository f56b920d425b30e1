#!/usr/bin/env python
"""bench.measure_dynamic alone, a few repetitions (host-driven path: how much does it vary?)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = sys.argv[:1]
import bench  # noqa: E402

args = bench.parse()
env = bench.Env()
for rep in range(3):
    t = time.perf_counter()
    d = bench.measure_dynamic(env, args)
    print("rep", rep, d["e2e"]["value"], d.get("replans"), "wall %.1f s" % (time.perf_counter() - t), flush=True)
t = time.perf_counter()
d = bench.measure_dynamic(env, args, S=256, iters=200, start_box=41)
print("200:", d["e2e"]["value"], d.get("replans"), "wall %.1f s" % (time.perf_counter() - t), flush=True)
