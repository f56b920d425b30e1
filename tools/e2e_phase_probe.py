#!/usr/bin/env python
"""Wall-clock of the stages of one end-to-end step (restart, copy-in, loop, copy-out, finish), run one after the other:
    python tools/e2e_phase_probe.py cfg4 [scenes] [iters]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from scarlet_b200 import BlendBatch  # noqa: E402
from scarlet_b200.blend import _fit_options  # noqa: E402

config = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
S = int(sys.argv[2]) if len(sys.argv) > 2 else bench.DEFAULT_SCENES[config]
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 50
base = [bench._make_scene(config, i) for i in range(min(S, 16))]
blends = [bench._make_blend(config, base[i % len(base)]) for i in range(S)]
batch = BlendBatch(blends)
plan = batch.plans[0]
init = plan.pack_current()[0]
opts = _fit_options(iters, 1e-3, 1, 0, dict(fixed_iterations=True), 10 ** 6)
for rep in range(3):
    t = [time.perf_counter()]
    plan.forget_state(values=init)
    for b in blends:
        b.loss.clear()
    t.append(time.perf_counter())
    h2d = batch._copy_in(0, True)
    t.append(time.perf_counter())
    out = plan.fit(opts)
    t.append(time.perf_counter())
    d2h = batch._copy_out(0, out)
    t.append(time.perf_counter())
    batch._finish([out + (h2d, d2h)])
    t.append(time.perf_counter())
    names = ["restart", "copy_in", "loop", "copy_out", "finish"]
    print(config, S, "rep", rep, " ".join("%s=%.1fms" % (n, 1e3 * (b - a)) for n, a, b in zip(names, t, t[1:])), "h2d=%.0fMB d2h=%.0fMB" % (h2d / 1e6, d2h / 1e6))
batch.close()
