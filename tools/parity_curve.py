#!/usr/bin/env python
"""Error-versus-iteration curves of the CUDA fitting path against the CPU oracle (GPU box; diagnostic + evidence).

    python tools/parity_curve.py cfg3 100 [scene ids...]  ->  gpurun_out/parity_curve_cfg3.json

Four trajectories of the same seeded scene are compared at checkpoints 1, 10, 20, 50, 100, ... iterations:

    O64   the oracle in the reference's arithmetic (float64 FFTs and morphologies, float32 model cube)
    O32   the oracle with ``float32_arithmetic()``: float32 FFTs, float32 morphologies and optimiser state -- an
          independent float32 implementation of the same algorithm
    G64   the CUDA float64 twin
    G32   the CUDA float32 product path

``G64 vs O64`` isolates algorithmic differences (should sit at rounding level); ``O32 vs O64`` measures how far float32
rounding ALONE moves the trajectory of this scene; ``G32 vs O64`` is the product's deviation from the reference
arithmetic.  A float32 tolerance in tests/ is justified where G32-vs-O64 is no larger than O32-vs-O64.
All errors are max |a-b| / max |b| over the array (``rel_peak``): model cube, K x C spectrum matrix, worst single
morphology image; loss relative.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import scarlet_oracle as so  # noqa: E402
from oracle import scenes  # noqa: E402
from scarlet_b200 import synthetic  # noqa: E402


def rel_peak(a, b):
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / max(np.abs(b).max(), 1e-300))


def snapshot_oracle(o):
    return dict(model=np.array(o.get_model(), dtype=np.float64),
                sed=np.array([np.asarray(s.spectrum.x, dtype=np.float64) for s in o.sources]),
                morph=[np.array(s.image.x, dtype=np.float64) for s in o.sources if s.kind == "extended"],
                loss=float(o.loss[-1]))


def oracle_run(scene, checkpoints, f32):
    snaps = {}

    def build():
        return scenes.build_oracle(scene, frame_dtype=np.float32)

    def cb(it):
        if it + 1 in checkpoints:
            snaps[it + 1] = snapshot_oracle(o)

    if f32:
        with so.float32_arithmetic():
            o = build()
            o.fit(max_iter=max(checkpoints), e_rel=1e-3, min_iter=10 ** 9, callback=cb)
    else:
        o = build()
        o.fit(max_iter=max(checkpoints), e_rel=1e-3, min_iter=10 ** 9, callback=cb)
    return snaps


def gpu_run(scene, checkpoints, precision):
    snaps = {}
    for n in checkpoints:
        blend = synthetic.make_blend(scene, precision=precision)
        blend.fit(max_iter=n, e_rel=1e-3, min_iter=10 ** 9, check_every=10 ** 6)
        snaps[n] = dict(model=np.array(blend.get_model(), dtype=np.float64),
                        sed=np.array([np.asarray(s.parameters[0], dtype=np.float64) for s in blend.sources]),
                        morph=[np.array(s.parameters[1], dtype=np.float64) for s in blend.sources if s.parameters[1].ndim == 2],
                        loss=float(blend.loss[-1]))
        blend._plan.close()
    return snaps


def compare(a, b):
    return dict(model=rel_peak(a["model"], b["model"]), sed=rel_peak(a["sed"], b["sed"]),
                morph=max(rel_peak(x, y) for x, y in zip(a["morph"], b["morph"])) if b["morph"] else 0.0,
                loss=abs(a["loss"] / b["loss"] - 1))


def main():
    config, n_max = sys.argv[1], int(sys.argv[2])
    ids = [int(x) for x in sys.argv[3:]] or [0]
    checkpoints = [n for n in (1, 10, 20, 30, 50, 100, 150, 200) if n <= n_max]
    out = dict(config=config, checkpoints=checkpoints, scenes={})
    for sid in ids:
        scene = synthetic.make_scene(config, sid)
        t0 = time.time()
        o64 = oracle_run(scene, checkpoints, False)
        o32 = oracle_run(scene, checkpoints, True)
        g64 = gpu_run(scene, checkpoints, 64)
        g32 = gpu_run(scene, checkpoints, 32)
        rows = {}
        for n in checkpoints:
            # the oracle snapshot after iteration index n-1 holds n updates; a GPU fit of n iterations holds n updates too
            rows[n] = dict(G64_vs_O64=compare(g64[n], o64[n]), O32_vs_O64=compare(o32[n], o64[n]),
                           G32_vs_O64=compare(g32[n], o64[n]), G32_vs_O32=compare(g32[n], o32[n]))
            print("%s scene %d it %3d | G64-O64 model %.1e sed %.1e morph %.1e | O32-O64 model %.1e sed %.1e morph %.1e | "
                  "G32-O64 model %.1e sed %.1e morph %.1e loss %.1e" % (
                      config, sid, n, rows[n]["G64_vs_O64"]["model"], rows[n]["G64_vs_O64"]["sed"], rows[n]["G64_vs_O64"]["morph"],
                      rows[n]["O32_vs_O64"]["model"], rows[n]["O32_vs_O64"]["sed"], rows[n]["O32_vs_O64"]["morph"],
                      rows[n]["G32_vs_O64"]["model"], rows[n]["G32_vs_O64"]["sed"], rows[n]["G32_vs_O64"]["morph"],
                      rows[n]["G32_vs_O64"]["loss"]), flush=True)
        out["scenes"][str(sid)] = dict(rows=rows, seconds=time.time() - t0)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "parity_curve_%s.json" % config), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
