#!/usr/bin/env python
"""Print the float32-vs-oracle parity margins of a config after N iterations (diagnostic; GPU box)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import scenes  # noqa: E402
from scarlet_b200 import synthetic  # noqa: E402


def rel_peak(a, b):
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / max(np.abs(b).max(), 1e-300))


def main():
    config, n_iter = sys.argv[1], int(sys.argv[2])
    scene_ids = [int(x) for x in sys.argv[3:]] or [0]
    for sid in scene_ids:
        scene = synthetic.make_scene(config, sid)
        o = scenes.build_oracle(scene, frame_dtype=np.float32)
        o.fit(max_iter=n_iter, e_rel=1e-3, min_iter=10 ** 9)
        for prec in (32, 64):
            blend = synthetic.make_blend(scene, precision=prec)
            blend.fit(max_iter=n_iter, e_rel=1e-3, min_iter=10 ** 9, check_every=10 ** 6)
            sed = max(rel_peak(s.parameters[0], os_.spectrum.x) for s, os_ in zip(blend.sources, o.sources))
            morph = max(rel_peak(s.parameters[1], os_.image.x) for s, os_ in zip(blend.sources, o.sources) if os_.kind == "extended")
            loss = float(np.abs(np.array(blend.loss) / np.array(o.loss) - 1).max())
            print("%s scene %d iters %d precision %d grid %s: sed %.2e morph %.2e loss %.2e" % (
                config, sid, n_iter, prec, os.environ.get("SB_REFERENCE_GRID", "device"), sed, morph, loss), flush=True)


if __name__ == "__main__":
    main()
