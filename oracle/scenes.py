"""Build the CPU oracle's objects from a synthetic scene dict (``scarlet_b200.synthetic.make_scene`` output --
plain arrays only).  TEST INFRASTRUCTURE ONLY."""
import numpy as np

from . import scarlet_oracle as so


def build_oracle(scene, frame_dtype=np.float32, sed_dtype=np.float32):
    C, N = scene["C"], scene["N"]
    cfg = scene["config"]
    model_psf = so.GaussianPSFOracle((scene["model_sigma"],) * C)
    obs = so.ObservationOracle(scene["images"], scene["weights"], so.ImagePSFOracle(scene["obs_psf"]), frame_dtype=frame_dtype)
    obs.match((C, N, N), model_psf)
    min_step = obs.channel_noise_rms()
    sources = []
    for s in scene["sources"]:
        if s["kind"] == "extended":
            sources.append(so.ExtendedSourceOracle(s["sed"], s["morph"], s["origin"], min_step=min_step, monotonic="angle",
                                                   symmetric=cfg["symmetric"], sed_dtype=sed_dtype))
        else:
            sources.append(so.PointSourceOracle(s["sed"], s["center"], model_psf, min_step=min_step, sed_dtype=sed_dtype))
    return so.SceneOracle((C, N, N), model_psf, sources, [obs], frame_dtype=frame_dtype)
