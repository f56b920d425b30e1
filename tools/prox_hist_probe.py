#!/usr/bin/env python
"""How many of the <= 10 proximal sub-iterations (blend.py:145) does the grouped update kernel actually run?

    python tools/prox_hist_probe.py cfg3 96 100   ->  histogram per window of 10 iterations (GPU box; diagnostic)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scarlet_b200 import BlendBatch, _native, synthetic  # noqa: E402


def main():
    config, S, iters = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    uniq = min(S, 16)
    base = [synthetic.make_scene(config, i) for i in range(uniq)]
    batch = BlendBatch([synthetic.make_blend(base[i % uniq]) for i in range(S)])
    plan = batch.plan
    plan.upload_parameters(state=True)
    opts = _native.fit_opts(max_iter=iters, e_rel=1e-3, min_iter=1, prox_max_iter=10, check_every=10 ** 6, fixed_iterations=True)
    plan.prox_histogram(True)
    rows = []
    for stop in range(10, iters + 1, 10):
        opts.run_until = stop
        plan.fit(opts)
        opts.resume = 1
        h = plan.prox_histogram(True)
        tot = max(int(h.sum()), 1)
        mean = float((h * np.arange(16)).sum() / tot)
        rows.append(dict(iterations=[stop - 10, stop], hist=h[:12].tolist(), mean=mean))
        print("%s it %3d-%3d  mean %.2f  hist(1..10) %s" % (config, stop - 10, stop, mean, h[1:11].tolist()), flush=True)
    plan.prox_histogram(False)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(dict(config=config, scenes=S, rows=rows), open(os.path.join(ROOT, "gpurun_out", "prox_hist_%s.json" % config), "w"), indent=1)


if __name__ == "__main__":
    main()
