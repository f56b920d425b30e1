#!/usr/bin/env python
"""Does running two half-batches as two plans on two CUDA streams overlap the latency-bound update kernel of one half
with the spectral kernels of the other?  (diagnostic; GPU box)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scarlet_b200 import BlendBatch, _native, synthetic  # noqa: E402

S, iters = int(sys.argv[1]) if len(sys.argv) > 1 else 96, 50
base = [synthetic.make_scene("cfg3", i) for i in range(16)]
opts = _native.fit_opts(max_iter=iters, e_rel=1e-3, fixed_iterations=True, check_every=10 ** 6)
for nsplit in (1, 2, 3, 4):
    plans = []
    for k in range(nsplit):
        n = S // nsplit
        blends = [synthetic.make_blend(base[(k * n + i) % 16]) for i in range(n)]
        plans.append(BlendBatch(blends).plan)
    for rep in range(3):
        for p in plans:
            p.forget_state()
            p.upload_parameters(state=True)
        for p in plans:
            p.sync()
        t0 = time.perf_counter()
        for p in plans:
            p.fit_enqueue(opts, iters)
        for p in plans:
            p.sync()
        dt = time.perf_counter() - t0
    print("plans=%d scenes=%d: %.2f ms per iteration, %.0f scene-iterations/s" % (nsplit, n * nsplit, 1e3 * dt / iters, n * nsplit * iters / dt), flush=True)
    for p in plans:
        p.close()
