#!/usr/bin/env python
"""Small workloads for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): every kernel family of the
fitting path runs at least once -- grouped + generic update kernels, point sources, the fused spectral kernels in both
precisions, the shifting (Toeplitz) path, the resampling kernels, the single-operator entry points.

    compute-sanitizer --tool racecheck python tools/sanitizer_workload.py [names...]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from scarlet_b200 import BlendBatch, synthetic  # noqa: E402


def tiny(precision):
    b = synthetic.make_blend(synthetic.make_scene("tiny", 0), precision=precision)
    b.fit(max_iter=3, e_rel=1e-3, min_iter=10 ** 9)
    b.get_model()


def cfg3_iteration():
    scenes = [synthetic.make_scene("cfg3", i) for i in range(2)]
    batch = BlendBatch([synthetic.make_blend(s) for s in scenes])
    batch.fit(max_iter=1, e_rel=1e-3, fixed_iterations=True)
    batch.close()


def cfg5_iteration():
    scenes = [synthetic.make_scene("cfg5", i) for i in range(2)]
    batch = BlendBatch([synthetic.make_blend(s) for s in scenes])
    batch.fit(max_iter=1, e_rel=1e-3, fixed_iterations=True)
    batch.close()


def shifting():
    cfg = dict(synthetic.CONFIGS["tiny"], shifting=True)
    b = synthetic.make_blend(synthetic.make_scene(cfg, 1), precision=32)
    b.fit(max_iter=2, e_rel=1e-3, min_iter=10 ** 9)


def multires():
    import multires_scene
    _, blend, _, _ = multires_scene.product_scene(32)
    blend.fit(max_iter=2, e_rel=1e-3, min_iter=10 ** 9)


def multires_rotated():
    """rotated ResolutionRenderer: k_rot_partial (cp.async staging), k_rot_residual, k_rot_adjoint, both precisions"""
    import multires_scene
    for precision in (32, 64):
        _, blend, _, _ = multires_scene.product_scene(precision, rotated=True)
        blend.fit(max_iter=2, e_rel=1e-3, min_iter=10 ** 9)


def psf_shift():
    """ConvolutionRenderer(psf_shift=...): k_psf_kernel / k_psf_corr / k_psf_update + the K^ refresh"""
    import scarlet_b200 as sb
    sc = synthetic.make_scene("tiny", 2)
    frame = sb.Frame(sc["images"].shape, psf=sb.GaussianPSF(sigma=(0.8,) * 3), channels=sc["channels"])
    obs = sb.Observation(sc["images"].copy(), psf=sb.ImagePSF(sc["obs_psf"].copy()), weights=sc["weights"].copy(), channels=sc["channels"])
    obs.match(frame, renderer=sb.renderer.ConvolutionRenderer(obs, frame, psf_shift=np.array([0.1, -0.2])))
    srcs = []
    for s_ in sc["sources"]:
        if s_["kind"] != "extended":
            continue
        B = s_["morph"].shape[0]
        srcs.append(sb.ExtendedSource(frame, s_["center"], obs, spectrum=s_["sed"].copy(), morphology=s_["morph"].copy(),
                                      bbox=sb.Box((B, B), origin=s_["origin"]), resizing=False))
    sb.Blend(srcs, obs).fit(max_iter=2, e_rel=1e-3, min_iter=10 ** 9)


def dynamic_batch():
    """BlendBatch with resizing sources: per-scene run state, k_inspect, sb_plan_set_sources"""
    cfg = dict(synthetic.CONFIGS["tiny"], resizing=True)
    blends = [synthetic.make_blend(synthetic.make_scene(cfg, i), precision=32) for i in range(3)]
    batch = BlendBatch(blends)
    batch.fit(max_iter=22, e_rel=1e-3)
    batch.close()


def big_box():
    """129 x 129 box: beyond the 16-bit byte offsets of the grouped kernel -> generic kernel"""
    cfg = dict(synthetic.CONFIGS["cfg2"], B=129, N=160, n_ext=2)
    b = synthetic.make_blend(synthetic.make_scene(cfg, 0), precision=32)
    b.fit(max_iter=1, e_rel=1e-3, min_iter=10 ** 9)


WORK = dict(tiny32=lambda: tiny(32), tiny64=lambda: tiny(64), cfg3=cfg3_iteration, cfg5=cfg5_iteration, shifting=shifting,
            multires=multires, multires_rotated=multires_rotated, psf_shift=psf_shift, dynamic_batch=dynamic_batch, big_box=big_box)

if __name__ == "__main__":
    for name in (sys.argv[1:] or list(WORK)):
        WORK[name]()
        print("sanitizer workload '%s' done" % name, flush=True)
