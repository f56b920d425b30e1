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
                                                   symmetric=cfg["symmetric"], sed_dtype=sed_dtype,
                                                   resizing=bool(cfg.get("resizing", False)),
                                                   shift=(np.array(s["center"]) - np.round(s["center"])) if cfg.get("shifting") else None))
        else:
            sources.append(so.PointSourceOracle(s["sed"], s["center"], model_psf, min_step=min_step, sed_dtype=sed_dtype))
    return so.SceneOracle((C, N, N), model_psf, sources, [obs], frame_dtype=frame_dtype)


def build_multires_oracle(scene, setup, frame_dtype=np.float32, sed_dtype=np.float32):
    """cfg4 (two observations on different pixel grids) from a ``synthetic.make_multires_scene`` dict.  ``setup`` carries
    the host set-up products of the low-resolution renderer (``frame_shape``, ``model_psf`` image, ``lr_kernel`` = padded
    difference kernel, ``lr_shifts``, ``lr_h``, ``hr_origin``), which are pinned to the reference's own set-up by
    tests/test_host_api.py::test_multiresolution_setup_vs_reference_fixture."""
    frame_shape = tuple(int(v) for v in setup["frame_shape"])
    model_psf = so.ImagePSFOracle(setup["model_psf"])
    if setup.get("lr_other_shifts") is not None:  # rotated grids (tests/test_host_api.py::test_rotated_multiresolution_setup_...)
        lr = so.RotatedResolutionObservationOracle(scene["lr_images"], scene["lr_weights"], setup["lr_kernel"], setup["lr_shifts"],
                                                   setup["lr_other_shifts"], setup["lr_h"], small_axis=setup["lr_small_axis"],
                                                   frame_dtype=frame_dtype, channel_offset=0)
    else:
        lr = so.ResolutionObservationOracle(scene["lr_images"], scene["lr_weights"], setup["lr_kernel"], setup["lr_shifts"], setup["lr_h"],
                                            frame_dtype=frame_dtype, channel_offset=0)
    lr.match(frame_shape, None)
    hr = so.ObservationOracle(scene["hr_images"], scene["hr_weights"], so.ImagePSFOracle(scene["hr_psfs"]), frame_dtype=frame_dtype,
                              channel_offset=5, origin=tuple(int(v) for v in setup["hr_origin"]))
    hr.match(frame_shape, model_psf)
    min_step = np.concatenate([lr.channel_noise_rms(), hr.channel_noise_rms()])
    sources = [so.ExtendedSourceOracle(s["sed"], s["morph"], s["origin"], min_step=min_step, monotonic="angle", symmetric=True,
                                       sed_dtype=sed_dtype) for s in scene["sources"]]
    return so.SceneOracle(frame_shape, model_psf, sources, [lr, hr], frame_dtype=frame_dtype)


def multires_setup(blend):
    """The set-up products ``build_multires_oracle`` needs, read off a product Blend (host objects only)."""
    obs_lr, obs_hr = blend.observations
    r = obs_lr.renderer
    return dict(frame_shape=blend.frame.shape, model_psf=np.asarray(blend.frame.psf.get_model(), dtype=np.float64),
                lr_kernel=np.asarray(r.diff_kernel.image, dtype=np.float64), lr_shifts=np.asarray(r.shifts), lr_h=float(r.h),
                hr_origin=obs_hr.renderer.origin, lr_other_shifts=np.asarray(r.other_shifts) if r.isrot else None,
                lr_small_axis=bool(getattr(r, "small_axis", True)))
