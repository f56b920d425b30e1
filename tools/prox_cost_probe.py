#!/usr/bin/env python
"""How much of the update stage is the proximal loop?  Stage profile with prox_max_iter = 1, 3, 10 (GPU box):
    python tools/prox_cost_probe.py [config] [scenes]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from scarlet_b200 import BlendBatch, _native  # noqa: E402

config = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
S = int(sys.argv[2]) if len(sys.argv) > 2 else bench.DEFAULT_SCENES[config]
base = [bench._make_scene(config, i) for i in range(min(S, 16))]
batch = BlendBatch([bench._make_blend(config, base[i % len(base)]) for i in range(S)])
plan = batch.plan
init = plan.pack_current()[0]
for pm in (1, 3, 10):
    opts = _native.fit_opts(max_iter=50, e_rel=1e-3, min_iter=1, prox_max_iter=pm, check_every=10 ** 6, fixed_iterations=True)
    for rep in range(2):
        sed, morph, cen = init
        _native.check(_native.lib().sb_plan_upload_params(plan._handle, 0, _native.ptr(sed), _native.ptr(morph), _native.ptr(cen)))
        _native.check(_native.lib().sb_plan_zero_state(plan._handle))
        stages = plan.profile(opts, 10)
    print(config, S, "prox_max_iter", pm, "source_update %.4f ms" % stages["source_update"], flush=True)
