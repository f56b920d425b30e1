#!/usr/bin/env python
"""Stage intervals of a BatchPipeline run (does the device loop of batch k slow down while k+1 copies in and k-1 copies out?):
    python tools/pipeline_probe.py [config] [scenes] [n_batches_in_sequence]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from scarlet_b200 import BatchPipeline, BlendBatch  # noqa: E402

config = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
S = int(sys.argv[2]) if len(sys.argv) > 2 else bench.DEFAULT_SCENES[config]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 8
base = [bench._make_scene(config, i) for i in range(min(S, 16))]
batches = [BlendBatch([bench._make_blend(config, base[i % len(base)]) for i in range(S)]) for _ in range(2)]
init = {id(b): [p.pack_current()[0] for p in b.plans] for b in batches}


def restart(k, b):
    for p, vals in zip(b.plans, init[id(b)]):
        p.forget_state(values=vals)
    for bl in b.blends:
        bl.loss.clear()


pipe = BatchPipeline(depth=2)
for rep in range(2):
    pipe.run([batches[k % 2] for k in range(n)], max_iter=50, e_rel=1e-3, fixed_iterations=True, check_every=10 ** 6,
             upload_observations=True, prepare=restart)
t0 = min(t["copy_in"][0] for t in pipe.timings)
for t in sorted(pipe.timings, key=lambda t: t["k"]):
    print("k=%d  copy_in %.1f-%.1f (%.1f ms)  loop %.1f-%.1f (%.1f ms)  copy_out %.1f-%.1f (%.1f ms)" % (
        t["k"], *(1e3 * (v - t0) for v in t["copy_in"]), 1e3 * (t["copy_in"][1] - t["copy_in"][0]),
        *(1e3 * (v - t0) for v in t["loop"]), 1e3 * (t["loop"][1] - t["loop"][0]),
        *(1e3 * (v - t0) for v in t["copy_out"]), 1e3 * (t["copy_out"][1] - t["copy_out"][0])))
