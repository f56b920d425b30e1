"""The two-observation multi-resolution test scene (tests/golden/multires.npz inputs) as product objects and as oracle
objects, with a few ExtendedSource components.  Shared by the CPU and GPU tests."""
import numpy as np

from conftest import golden


def _sources(frame_shape, n_src=3, B=15, seed=3):
    rng = np.random.default_rng(seed)
    C, Ny, Nx = frame_shape
    out = []
    for k in range(n_src):
        cy, cx = rng.uniform(22, Ny - 22), rng.uniform(22, Nx - 22)
        py, px = int(np.round(cy)), int(np.round(cx))
        y, x = np.mgrid[:B, :B] - B // 2
        rs, q, th = rng.uniform(1.5, 3.0), rng.uniform(0.6, 1.0), rng.uniform(0, np.pi)
        u = np.cos(th) * x + np.sin(th) * y
        v = (-np.sin(th) * x + np.cos(th) * y) / q
        morph = np.exp(-np.sqrt(u * u + v * v) / rs)
        out.append(dict(center=(cy, cx), origin=(py - B // 2, px - B // 2), morph=morph / morph.max(),
                        sed=(rng.uniform(5, 50) * rng.dirichlet(np.ones(C)) * C).astype(np.float32)))
    return out


def product_scene(precision=32, rotated=False):
    """-> (golden, blend, obs_lr, obs_hr) built from scarlet_b200 objects; ``rotated``: the low-resolution grid is turned by 25
    degrees against the model frame (tests/golden/multires_rot.npz)."""
    import scarlet_b200 as sb
    from scarlet_b200.wcs import AffineWCS
    g = golden("multires_rot.npz" if rotated else "multires.npz")
    obs_hr = sb.Observation(g["hr_images"].copy(), psf=sb.ImagePSF(g["hr_psfs"].copy()), weights=g["hr_weights"].copy(),
                            wcs=AffineWCS(g["hr_cd"], crpix=g["hr_crpix"]), channels=["h0", "h1", "h2"])
    obs_lr = sb.Observation(g["lr_images"].copy(), psf=sb.ImagePSF(g["lr_psfs"].copy()), weights=g["lr_weights"].copy(),
                            wcs=AffineWCS(g["lr_cd"], crpix=g["lr_crpix"]), channels=["l0", "l1", "l2", "l3", "l4"])
    frame = sb.Frame.from_observations([obs_lr, obs_hr], coverage="union")
    if precision == 64:
        frame = sb.Frame(frame.shape, channels=frame.channels, psf=frame.psf, wcs=frame.wcs, dtype=np.float64)
        obs_lr.match(frame)
        obs_hr.match(frame)
    observations = [obs_lr, obs_hr]
    srcs = []
    for s in _sources(frame.shape):
        B = s["morph"].shape[0]
        sky = frame.get_sky_coord(np.array(s["center"]))
        srcs.append(sb.ExtendedSource(frame, sky, observations, spectrum=s["sed"].copy(), morphology=s["morph"].copy(),
                                      bbox=sb.Box((B, B), origin=s["origin"]), monotonic="angle", symmetric=True, resizing=False))
    return g, sb.Blend(srcs, observations, precision=precision), obs_lr, obs_hr


def oracle_scene(frame_dtype=np.float32, rotated=False, setup_of=None):
    """The same scene from oracle objects; set-up products of the low-resolution renderer come from the fixture (float64
    frame), or -- ``setup_of`` = the product's low-resolution renderer -- from the product's own set-up: with a float32 frame
    the reference computes the difference kernel in float32 arithmetic, 1e-4 away from its float64 one (both pinned to the
    reference's by tests/test_host_api.py)."""
    from oracle import scarlet_oracle as so
    g = dict(golden("multires_rot.npz" if rotated else "multires.npz"))
    tag = "" if rotated else "64"  # the rotated fixture is made with a float64 frame throughout
    if setup_of is not None:
        g["lr_diff_kernel" + tag] = np.asarray(setup_of.diff_kernel.image, dtype=np.float64)
        g["lr_shifts" + tag], g["lr_h" + tag] = np.asarray(setup_of.shifts, dtype=np.float64), float(setup_of.h)
        g["model_psf" + tag] = np.asarray(setup_of.model_frame.psf.get_model(), dtype=np.float64)
        if rotated:
            g["lr_other_shifts"] = np.asarray(setup_of.other_shifts, dtype=np.float64)
    frame_shape = tuple(int(v) for v in g["frame_shape" + tag])
    model_psf = so.ImagePSFOracle(g["model_psf" + tag])
    if rotated:
        lr = so.RotatedResolutionObservationOracle(g["lr_images"], g["lr_weights"], g["lr_diff_kernel"], g["lr_shifts"],
                                                   g["lr_other_shifts"], float(g["lr_h"]), small_axis=bool(g["lr_small_axis"]),
                                                   frame_dtype=frame_dtype, channel_offset=0)
    else:
        lr = so.ResolutionObservationOracle(g["lr_images"], g["lr_weights"], g["lr_diff_kernel" + tag], g["lr_shifts" + tag],
                                            float(g["lr_h" + tag]), frame_dtype=frame_dtype, channel_offset=0)
    lr.match(frame_shape, None)
    hr = so.ObservationOracle(g["hr_images"], g["hr_weights"], so.ImagePSFOracle(g["hr_psfs"]), frame_dtype=frame_dtype,
                              channel_offset=5, origin=tuple(int(v) for v in g["hr_model_slice_start" + tag]))
    hr.match(frame_shape, model_psf)
    min_step = np.concatenate([lr.channel_noise_rms(), hr.channel_noise_rms()])
    srcs = [so.ExtendedSourceOracle(s["sed"], s["morph"], s["origin"], min_step=min_step, monotonic="angle", symmetric=True)
            for s in _sources(frame_shape)]
    return g, so.SceneOracle(frame_shape, model_psf, srcs, [lr, hr], frame_dtype=frame_dtype)
