#!/usr/bin/env python
"""cProfile of BlendBatch.fit with dynamic boxes (where does the host-driven path spend its time?)"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scarlet_b200 import BlendBatch, synthetic  # noqa: E402

S, iters, start_box = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
cfg = dict(synthetic.CONFIGS["cfg2"], B=start_box, resizing=True, config_id=22)
base = [synthetic.make_scene(cfg, i) for i in range(16)]
for rep in range(3):
    blends = [synthetic.make_blend(base[i % 16], precision=32) for i in range(S)]
    batch = BlendBatch(blends, precision=32)
    t0 = time.perf_counter()
    if rep == 2:
        pr = cProfile.Profile()
        pr.enable()
    res = batch.fit(max_iter=iters, e_rel=1e-9, upload_observations=True)
    if rep == 2:
        pr.disable()
    dt = time.perf_counter() - t0
    print("rep", rep, "%.3f s" % dt, sum(r[0] for r in res) / dt, batch.replans, flush=True)
    batch.close()
pstats.Stats(pr).sort_stats("cumtime").print_stats(28)
