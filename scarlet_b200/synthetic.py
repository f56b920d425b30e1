"""Seeded synthetic multi-band scenes for the benchmark and the parity tests (pure NumPy/SciPy; no CUDA, no oracle).

Follows the generator specified in SURVEY.md 8(d): model PSF GaussianPSF(0.8); observed PSFs either Gaussian
(cfg2) or Moffat images (cfg3/cfg5); exponential elliptical galaxies with log-uniform amplitudes and Dirichlet
colours; point sources at sub-pixel positions; unit-variance Gaussian noise (weights = 1); fixed B x B boxes,
``resizing=False``, ``shifting=False``.  One documented deviation: the initial morphologies are the truth
profiles with perturbed shape parameters (already monotonic / normalised) instead of "truth + noise pushed
through the constraint chain", so that generating a scene needs neither the GPU nor the oracle.
"""
import numpy as np
from scipy import signal

CONFIGS = {
    # name: frame, sources, observed PSF, constraints, default iteration count
    "cfg2": dict(C=5, N=128, n_ext=10, n_pt=0, psf="gaussian", P=21, B=41, symmetric=False, iters=200, config_id=2),
    "cfg3": dict(C=5, N=256, n_ext=20, n_pt=5, psf="moffat", P=41, B=41, symmetric=True, iters=100, config_id=3),
    "cfg5": dict(C=5, N=128, n_ext=12, n_pt=0, psf="moffat", P=41, B=41, symmetric=False, iters=100, config_id=5),
    "tiny": dict(C=3, N=40, n_ext=3, n_pt=1, psf="gaussian", P=15, B=15, symmetric=True, iters=30, config_id=9),
}
MODEL_SIGMA = 0.8


def _moffat(P, fwhm, beta=2.5):
    alpha = fwhm / (2 * np.sqrt(2 ** (1 / beta) - 1))
    y, x = np.mgrid[:P, :P] - P // 2
    img = (1 + (x * x + y * y) / alpha ** 2) ** (-beta)
    return img / img.sum()


def _profile(B, dy, dx, rs, q, theta):
    y, x = np.mgrid[:B, :B] - B // 2
    y = y - dy
    x = x - dx
    ct, st = np.cos(theta), np.sin(theta)
    u = ct * x + st * y
    v = (-st * x + ct * y) / q
    img = np.exp(-np.sqrt(u * u + v * v) / rs)
    return img / img.max()


def make_scene(config="cfg2", scene_id=0, per_scene_psf=True):
    """-> dict of plain arrays describing one scene (data, PSFs, truth and initial parameters)."""
    cfg = dict(CONFIGS[config]) if isinstance(config, str) else dict(config)
    rng = np.random.default_rng(1000 * cfg["config_id"] + scene_id)
    C, N, B, P = cfg["C"], cfg["N"], cfg["B"], cfg["P"]
    from . import fft as sfft
    from .psf import GaussianPSF
    model_psf = GaussianPSF(sigma=(MODEL_SIGMA,) * C)
    jitter = rng.uniform(0.9, 1.1) if per_scene_psf else 1.0
    if cfg["psf"] == "gaussian":
        sig = np.linspace(1.2, 2.0, C) * jitter
        obs_psf = GaussianPSF(sig, boxsize=P).get_model()
    else:
        fw = np.linspace(2.8, 4.2, C) * jitter
        obs_psf = np.stack([_moffat(P, f) for f in fw])
    diff = sfft.match_psf(obs_psf.astype(np.float32), model_psf.get_model().astype(np.float32), padding=10).image

    margin = min(12, N // 4)
    sources, truth = [], np.zeros((C, N, N))
    for k in range(cfg["n_ext"] + cfg["n_pt"]):
        cy, cx = rng.uniform(margin, N - margin, size=2)
        amp = np.exp(rng.uniform(np.log(20), np.log(2000)))
        sed = amp * rng.dirichlet(np.ones(C)) * C
        if k < cfg["n_ext"]:
            py, px = int(np.round(cy)), int(np.round(cx))
            rs, q, th = rng.uniform(1.5, 4.0), rng.uniform(0.5, 1.0), rng.uniform(0, np.pi)
            rs = min(rs, B / 10.0)
            tm = _profile(B, cy - py, cx - px, rs, q, th)
            init = _profile(B, 0.0, 0.0, rs * rng.uniform(0.7, 1.3), min(1.0, q * rng.uniform(0.8, 1.2)), th + rng.normal(0, 0.2))
            origin = (py - B // 2, px - B // 2)
            src = dict(kind="extended", center=(cy, cx), origin=origin, sed_true=sed, morph_true=tm,
                       sed=(sed * rng.uniform(0.7, 1.3, C)).astype(np.float32), morph=init)
        else:
            b = model_psf.bbox.shape[1]
            py, px = int(np.round(cy)), int(np.round(cx))
            origin = (py - b // 2, px - b // 2)
            box_center = np.array([origin[0] + b / 2, origin[1] + b / 2])  # the reference's convention (morphology.py:505)
            tm = model_psf.get_model(offset=np.array([cy, cx]) - box_center)[0]
            src = dict(kind="point", center_true=(cy, cx), origin=origin, sed_true=sed, morph_true=tm,
                       sed=(sed * rng.uniform(0.7, 1.3, C)).astype(np.float32),
                       center=(cy + rng.uniform(-0.3, 0.3), cx + rng.uniform(-0.3, 0.3)))
        oy, ox = src["origin"]
        bb = src["morph_true"].shape[0]
        y0, y1, x0, x1 = max(0, oy), min(N, oy + bb), max(0, ox), min(N, ox + bb)
        truth[:, y0:y1, x0:x1] += sed[:, None, None] * src["morph_true"][None, y0 - oy:y1 - oy, x0 - ox:x1 - ox]
        sources.append(src)
    clean = np.stack([signal.fftconvolve(truth[c], diff[c], mode="same") for c in range(C)])
    images = (clean + rng.standard_normal(clean.shape)).astype(np.float32)
    return dict(config=cfg, scene_id=scene_id, C=C, N=N, images=images, weights=np.ones_like(images),
                obs_psf=obs_psf, model_sigma=MODEL_SIGMA, sources=sources, channels=[str(c) for c in range(C)])


def make_blend(scene, precision=32, device=None):
    """Build the scarlet_b200 objects (Frame, Observation, sources, Blend) for a scene dict."""
    import scarlet_b200 as sb
    C = scene["C"]
    cfg = scene["config"]
    model_psf = sb.GaussianPSF(sigma=(scene["model_sigma"],) * C)
    # the float64 twin also keeps the frame (PSF images, difference kernel) in float64, like the float64 oracle
    frame = sb.Frame(scene["images"].shape, psf=model_psf, channels=scene["channels"],
                     dtype=np.float32 if precision == 32 else np.float64)
    obs = sb.Observation(scene["images"].copy(), psf=sb.ImagePSF(scene["obs_psf"].copy()), weights=scene["weights"].copy(),
                         channels=scene["channels"])
    obs.match(frame)
    sources = []
    for s in scene["sources"]:
        if s["kind"] == "extended":
            B = s["morph"].shape[0]
            sources.append(sb.ExtendedSource(frame, s["center"], obs, spectrum=s["sed"].copy(), morphology=s["morph"].copy(),
                                             bbox=sb.Box((B, B), origin=s["origin"]), monotonic="angle",
                                             symmetric=cfg["symmetric"], resizing=bool(cfg.get("resizing", False)),
                                             shifting=bool(cfg.get("shifting", False))))
        else:
            sources.append(sb.PointSource(frame, s["center"], obs, spectrum=s["sed"].copy()))
    return sb.Blend(sources, obs, precision=precision, device=device)


def algorithmic_bytes(scene_or_cfg, fft_shape, elem=4):
    """algorithmic bytes per iteration per scene."""
    cfg = scene_or_cfg["config"] if "config" in scene_or_cfg else scene_or_cfg
    C, N = cfg["C"], cfg["N"]
    Fy, Fx = fft_shape
    Fc = Fy * (Fx // 2 + 1)
    src = cfg["n_ext"] * cfg["B"] ** 2 * (14 + C) + cfg["n_pt"] * 81 * (14 + C)
    return elem * (6 * C * Fy * Fx + 3 * C * N * N + src) + 2 * elem * (6 * C * Fc)


# ---------------------------------------------------------------------------------------------------
# BASELINE config 4: two observations on different pixel grids (5 bands at 0.2"/px + 3 bands at 0.03"/px)
# ---------------------------------------------------------------------------------------------------
# (narrow 9x9 high-resolution PSFs: the reference's PSF matching is a plain Fourier division, which is only stable when
#  the model PSF is narrow; they also make the reference's grid 228 + 9 + 3 = 240, a length of the fused kernels)
CFG4 = dict(hr_n=200, lr_n=30, hr_scale=0.03, lr_scale=0.2, n_ext=8, B=41, hr_P=9, lr_P=15, config_id=4)


def _gauss_img(P, sig):
    y, x = np.mgrid[:P, :P] - P // 2
    out = np.stack([np.exp(-(x * x + y * y) / (2 * s * s)) for s in sig])
    return out / out.sum(axis=(1, 2))[:, None, None]


_MR_CACHE = {}


def _multires_observations(scene, dtype=np.float32):
    """Frame + matched observations of a cfg4 scene.  Every cfg4 scene shares PSFs and WCS, so the (expensive, host-side)
    renderer set-up is done once per dtype and re-attached through ``Observation.match(frame, renderer=...)``."""
    import scarlet_b200 as sb
    from .wcs import AffineWCS
    cfg = scene["config"]
    hr_c, lr_c = (cfg["hr_n"] - 1) / 2.0, (cfg["lr_n"] - 1) / 2.0
    key = (tuple(sorted(cfg.items())), np.dtype(dtype).name)
    cached = _MR_CACHE.get(key)
    wcs_hr = cached["wcs_hr"] if cached else AffineWCS(np.diag([cfg["hr_scale"]] * 2), crpix=(hr_c, hr_c))
    ang = np.deg2rad(cfg.get("lr_angle", 0.0))  # low-resolution grid turned against the high-resolution one (rotated ResolutionRenderer)
    rot = np.array([[np.cos(ang), -np.sin(ang)], [np.sin(ang), np.cos(ang)]]) if ang else np.eye(2)
    wcs_lr = cached["wcs_lr"] if cached else AffineWCS(cfg["lr_scale"] * rot, crpix=(lr_c, lr_c))
    obs_hr = sb.Observation(scene["hr_images"].copy(), psf=sb.ImagePSF(scene["hr_psfs"].copy()), weights=scene["hr_weights"].copy(),
                            wcs=wcs_hr, channels=["h0", "h1", "h2"])
    obs_lr = sb.Observation(scene["lr_images"].copy(), psf=sb.ImagePSF(scene["lr_psfs"].copy()), weights=scene["lr_weights"].copy(),
                            wcs=wcs_lr, channels=["l0", "l1", "l2", "l3", "l4"])
    if cached:
        frame = cached["frame"]
        obs_lr.match(frame, renderer=cached["r_lr"])
        obs_hr.match(frame, renderer=cached["r_hr"])
        return frame, obs_lr, obs_hr
    frame = sb.Frame.from_observations([obs_lr, obs_hr], coverage="union")
    if dtype is np.float64:
        frame = sb.Frame(frame.shape, channels=frame.channels, psf=frame.psf, wcs=frame.wcs, dtype=np.float64)
        obs_lr.match(frame)
        obs_hr.match(frame)
    _MR_CACHE[key] = dict(frame=frame, r_lr=obs_lr.renderer, r_hr=obs_hr.renderer, wcs_hr=wcs_hr, wcs_lr=wcs_lr)
    return frame, obs_lr, obs_hr


def make_multires_scene(scene_id=0, config=None):
    """Plain arrays of one cfg4 scene: truth galaxies in the common model frame, observed once through the
    low-resolution resampling renderer and once through the high-resolution convolution, unit-variance noise."""
    cfg = dict(CFG4 if config is None else config)
    rng = np.random.default_rng(1000 * cfg["config_id"] + scene_id)
    hr_n, lr_n, B = cfg["hr_n"], cfg["lr_n"], cfg["B"]
    scene = dict(config=cfg, scene_id=scene_id,
                 hr_images=np.zeros((3, hr_n, hr_n), np.float32), lr_images=np.zeros((5, lr_n, lr_n), np.float32),
                 hr_weights=np.ones((3, hr_n, hr_n), np.float32), lr_weights=np.ones((5, lr_n, lr_n), np.float32),
                 hr_psfs=_gauss_img(cfg["hr_P"], np.linspace(0.8, 1.0, 3)), lr_psfs=_gauss_img(cfg["lr_P"], np.linspace(1.0, 1.4, 5)))
    frame, obs_lr, obs_hr = _multires_observations(scene, np.float64)
    C, Ny, Nx = frame.shape
    margin = B // 2 + 4
    sources, truth = [], np.zeros((C, Ny, Nx))
    lo_y, hi_y, lo_x, hi_x = margin, Ny - margin, margin, Nx - margin
    if cfg.get("lr_angle"):  # the union frame of rotated grids has corners no observation covers: keep the galaxies on the data
        oy, ox = obs_hr.renderer.origin
        lo_y, hi_y, lo_x, hi_x = oy + margin, oy + hr_n - margin, ox + margin, ox + hr_n - margin
    for k in range(cfg["n_ext"]):
        cy, cx = rng.uniform(lo_y, hi_y), rng.uniform(lo_x, hi_x)
        py, px = int(np.round(cy)), int(np.round(cx))
        rs, q, th = rng.uniform(2.0, 4.0), rng.uniform(0.5, 1.0), rng.uniform(0, np.pi)
        sed = np.exp(rng.uniform(np.log(20), np.log(500))) * rng.dirichlet(np.ones(C)) * C
        tm = _profile(B, cy - py, cx - px, rs, q, th)
        init = _profile(B, 0.0, 0.0, rs * rng.uniform(0.7, 1.3), min(1.0, q * rng.uniform(0.8, 1.2)), th + rng.normal(0, 0.2))
        origin = (py - B // 2, px - B // 2)
        truth[:, origin[0]:origin[0] + B, origin[1]:origin[1] + B] += sed[:, None, None] * tm[None]
        sources.append(dict(kind="extended", center=(cy, cx), origin=origin, sed_true=sed, morph_true=tm,
                            sed=(sed * rng.uniform(0.7, 1.3, C)).astype(np.float32), morph=init))
    lr_clean = obs_lr.renderer.get_model()(truth)
    r2 = obs_hr.renderer
    hr_sub = truth[r2.channel_offset:r2.channel_offset + 3]
    hr_conv = np.stack([signal.fftconvolve(hr_sub[c], np.asarray(r2.diff_kernel.image[c], dtype=np.float64), mode="same") for c in range(3)])
    oy, ox = r2.origin
    scene["hr_images"] = (hr_conv[:, oy:oy + hr_n, ox:ox + hr_n] + rng.standard_normal((3, hr_n, hr_n))).astype(np.float32)
    scene["lr_images"] = (lr_clean + rng.standard_normal(lr_clean.shape)).astype(np.float32)
    scene["sources"] = sources
    scene["frame_shape"] = (int(C), int(Ny), int(Nx))
    return scene


def algorithmic_bytes_multires(cfg, elem=4, frame=(8, 228, 228), fft_shape=(240, 240)):
    """SURVEY.md 8(d) accounting extended to cfg4 (two observations of one model frame): every logical stage reads its
    inputs and writes its outputs once.

    high-resolution observation (ConvolutionRenderer, 3 bands, 200 x 200): the single-observation formula --
        6 real-grid passes + 3 image passes + 6 complex passes;
    low-resolution observation (ResolutionRenderer, 5 bands, 30 x 30; renderer.py:262-547 as the Parseval form of
    csrc/spectral.cuh): forward  = padded model write + read (2 real), M^ write + read + K^ read (3 complex),
                                   row-resampled spectrum T1 write + read (2 x C H Fxc complex), rendered + data + weights (3 images);
                        adjoint  = U write + read (2 x C H Fxc complex), K^ Q^ write + read + K^ read (3 complex),
                                   gradient grid write + read (2 real);
    per source: B^2 (14 + C_model) as in the single-observation formula."""
    c = dict(CFG4)
    c.update(cfg if "hr_n" in cfg else {})
    Cm, Ny, Nx = frame
    Fy, Fx = fft_shape
    Fxc = Fx // 2 + 1
    Fc = Fy * Fxc
    hr_C, lr_C, H, W = 3, 5, c["hr_n"], c["lr_n"]
    hr = elem * (6 * hr_C * Fy * Fx + 3 * hr_C * H * H) + 2 * elem * (6 * hr_C * Fc)
    lr = elem * (4 * lr_C * Fy * Fx + 3 * lr_C * W * W) + 2 * elem * (6 * lr_C * Fc + 4 * lr_C * W * Fxc)
    src = elem * c["n_ext"] * c["B"] ** 2 * (14 + Cm)
    return hr + lr + src


def make_multires_blend(scene, precision=32, device=None):
    """scarlet_b200 objects of a cfg4 scene: -> Blend over [low-resolution, high-resolution] observations."""
    import scarlet_b200 as sb
    frame, obs_lr, obs_hr = _multires_observations(scene, np.float32 if precision == 32 else np.float64)
    observations = [obs_lr, obs_hr]
    srcs = []
    for s in scene["sources"]:
        B = s["morph"].shape[0]
        srcs.append(sb.ExtendedSource(frame, frame.get_sky_coord(np.array(s["center"])), observations, spectrum=s["sed"].copy(),
                                      morphology=s["morph"].copy(), bbox=sb.Box((B, B), origin=s["origin"]), monotonic="angle",
                                      symmetric=True, resizing=False))
    return sb.Blend(srcs, observations, precision=precision, device=device)
