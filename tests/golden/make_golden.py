"""Generate the golden fixtures in ``tests/golden`` from the REFERENCE'S OWN forward code.

Runs only in the build container (needs the read-only reference tree at ``/root/reference``; it is
executed through the import shim ``oracle/ref_shim.py`` because autograd/proxmin/astropy are absent).
The fixtures are small ``.npz`` files that travel with the repo, so the GPU box never needs the
reference tree.  Re-run with:  ``python tests/golden/make_golden.py``

What is written (all values produced by reference code, float64 unless the reference says otherwise):

* ``obs_render_loss.npz``  -- the scenario of the reference's ``tests/test_observation.py:12-47``.
* ``hsc_cosmos_35.npz``    -- BASELINE config 1: data/hsc_cosmos_35.npz, 3 ``ExtendedSource`` initialised by the
                             reference; inputs, initial parameters, boxes, diff kernel, model, render, logL,
                             spectrum steps, the morphology constraint chain applied to perturbed images,
                             and central finite differences of the reference forward for a few parameters.
* ``point_extended.npz``   -- data/psf_unmatched_sim.npz: 2 ``PointSource`` + 2 ``ExtendedSource``; same contents,
                             plus point-source morphology models at sub-pixel centres.
* ``prox_chain.npz``       -- constraint-chain outputs (monotonic angle/flat/nearest x symmetric on/off) on seeded
                             random 41x41 / 21x21 / 20x31 images.
* ``monotonic_weights.npz``-- ``getRadialMonotonicWeights`` for several shapes / kinds / centres.
* ``multires_rot.npz``     -- the rotated branch of ``ResolutionRenderer``: set-up products (shifts along both axes, kernel) and render.
* ``psf_shift.npz``        -- ``ConvolutionRenderer(psf_shift=...)``: shifted difference kernel, render, logL and finite
                             differences of logL wrt the shift, for two shifts.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

REF_DATA = os.path.join(ref_shim.REFERENCE_ROOT, "data")


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print("wrote %s (%.1f kB)" % (name, os.path.getsize(path) / 1e3))


def obs_render_loss(sc):
    """tests/test_observation.py:12-47 through reference code."""
    shape0, s0 = (3, 13, 13), 0.9
    model_psf = sc.psf.GaussianPSF(s0, boxsize=shape0[1])
    model_psf_image = model_psf.get_model()
    shape = (3, 43, 43)
    channels = list(range(shape[0]))
    frame = sc.frame.Frame(shape, psf=model_psf, channels=channels)
    origin = (0, shape[1] // 2 - shape0[1] // 2, shape[2] // 2 - shape0[2] // 2)
    bbox = sc.bbox.Box(shape0, origin=origin)
    model = np.zeros(shape)
    box = np.stack([model_psf_image[0] for _ in range(shape[0])], axis=0)
    bbox.insert_into(model, box)
    sigmas = np.array([2.1, 1.1, 3.5])
    psf = sc.psf.GaussianPSF(sigmas, boxsize=shape[1])
    images = np.ones(shape)
    obs = sc.observation.Observation(images, psf=psf, channels=channels)
    obs.match(frame)
    rendered = obs.render(model)
    logL = obs.get_log_likelihood(model)
    save("obs_render_loss.npz", model_sigma=s0, model_boxsize=shape0[1], obs_sigmas=sigmas, obs_boxsize=shape[1],
         model=model, model_psf_image=np.asarray(model_psf_image), obs_psf_image=np.asarray(psf.get_model()),
         diff_kernel=obs.renderer.diff_kernel.image, rendered=rendered, logL=logL, images=images)


def _finite_diff(blend, obs, params, which, eps=1e-4):
    """Central differences of the reference forward, evaluated with a float64 model frame (the float32
    default frame makes the loss too noisy to difference).  The loss is quadratic in any single sed or
    morphology element, so the central difference is exact up to rounding; for centres it is O(eps^2)."""
    frame = blend.frame
    keep = frame.dtype
    frame.dtype = np.float64
    out = []
    try:
        for (pi, idx) in which:
            p = params[pi]
            old = p[idx].copy() if hasattr(p[idx], "copy") else p[idx]
            vals, xs = [], []
            for sgn in (+1, -1):
                p[idx] = old + sgn * eps
                xs.append(float(p[idx]))
                model = blend.get_model()
                vals.append(-obs.get_log_likelihood(model))
            p[idx] = old
            out.append((vals[0] - vals[1]) / (xs[0] - xs[1]))
    finally:
        frame.dtype = keep
    return np.array(out)


def _scene_dump(sc, prefix, frame, obs, sources, extra):
    blend = sc.blend.Blend(sources, obs)
    model = blend.get_model()
    rendered = obs.render(model)
    logL = obs.get_log_likelihood(model)
    out = dict(extra)
    out.update(model=model, rendered=rendered, logL=logL, log_norm=obs.log_norm,
               diff_kernel=obs.renderer.diff_kernel.image,
               noise_rms_band=np.array(np.mean(obs.noise_rms, axis=(1, 2))), n_sources=len(sources))
    for k, src in enumerate(sources):
        ps = src.parameters
        out["src%d_kind" % k] = type(src).__name__
        out["src%d_origin" % k] = np.array(src.bbox.origin)
        out["src%d_shape" % k] = np.array(src.bbox.shape)
        for p in ps:
            out["src%d_%s" % (k, p.name)] = np.asarray(p)
            st = p.step(p, it=0) if callable(p.step) else p.step
            out["src%d_%s_step" % (k, p.name)] = np.asarray(st, dtype=np.float64)
        out["src%d_model" % k] = np.asarray(src.get_model())
    return blend, out


def hsc_cosmos_35(sc):
    d = np.load(os.path.join(REF_DATA, "hsc_cosmos_35.npz"))
    images, variance, psfs = d["images"], d["variance"], d["psfs"]
    weights = 1 / variance
    channels = [str(f) for f in d["filters"]]
    centers = [(float(s["y"]), float(s["x"])) for s in d["catalog"]][:3]
    model_psf = sc.psf.GaussianPSF(sigma=(0.8,) * len(channels))
    frame = sc.frame.Frame(images.shape, psf=model_psf, channels=channels)
    obs = sc.observation.Observation(images, psf=sc.psf.ImagePSF(psfs.copy()), weights=weights, channels=channels)
    obs.match(frame)
    sources = [sc.source.ExtendedSource(frame, c, obs, resizing=False) for c in centers]
    blend, out = _scene_dump(sc, "hsc", frame, obs, sources,
                             dict(images=images, weights=weights.astype(np.float32), psfs=psfs,
                                  model_sigma=0.8, centers=np.array(centers)))
    params = blend.parameters
    # promote the float32 spectra to float64 for the finite differences only
    which = [(0, 2), (1, (20, 20)), (1, (5, 30)), (3, 0), (4, (30, 30)), (6, 4), (7, (20, 18)), (7, (0, 0))]
    fd = []
    for pi, idx in which:
        p = params[pi]
        eps = 1e-2 * max(abs(float(p[idx])), 1e-2) if p.dtype == np.float32 else 1e-4
        fd.append(_finite_diff(blend, obs, params, [(pi, idx)], eps=eps)[0])
    out["fd_which"] = np.array([(pi,) + (tuple(np.atleast_1d(idx)) + (-1,))[:2] for pi, idx in which])
    out["fd_grad"] = np.array(fd)
    # the constraint chain on perturbed morphologies (one prox application each)
    rng = np.random.default_rng(35)
    for k, src in enumerate(sources):
        img = np.asarray(params[3 * k + 1]).copy()
        pert = np.maximum(img + 0.05 * rng.standard_normal(img.shape), 0)
        out["src%d_perturbed" % k] = pert.copy()
        out["src%d_chain" % k] = np.asarray(params[3 * k + 1].constraint(pert.copy(), 0))
    save("hsc_cosmos_35.npz", **out)


def point_extended(sc):
    d = np.load(os.path.join(REF_DATA, "psf_unmatched_sim.npz"), allow_pickle=True)
    images, psfs = d["images"], d["psfs"]
    channels = [str(f) for f in d["filters"]]
    cat = d["catalog"]
    _, first = np.unique(cat["index"], return_index=True)
    cat = cat[np.sort(first)]
    rng = np.random.default_rng(7)
    weights = (1.0 / (0.05 + 0.02 * rng.random(images.shape))).astype(np.float32)
    weights[:, :3, :5] = 0  # a masked corner: exercises log_norm / noise_rms masking
    model_psf = sc.psf.GaussianPSF(sigma=(0.8,) * len(channels))
    frame = sc.frame.Frame(images.shape, psf=model_psf, channels=channels)
    obs = sc.observation.Observation(images, psf=sc.psf.ImagePSF(psfs.astype(np.float64)), weights=weights,
                                     channels=channels)
    obs.match(frame)
    sources, kinds, centers = [], [], []
    for row in cat:
        c = (float(row["y"]) + 0.23, float(row["x"]) - 0.31)
        if row["is_star"] and kinds.count("P") < 2:
            sources.append(sc.source.PointSource(frame, c, obs))
            kinds.append("P")
            centers.append(c)
        elif (not row["is_star"]) and kinds.count("E") < 2:
            sources.append(sc.source.ExtendedSource(frame, c, obs, resizing=False))
            kinds.append("E")
            centers.append(c)
    blend, out = _scene_dump(sc, "pe", frame, obs, sources,
                             dict(images=images, weights=weights, psfs=psfs.astype(np.float64), model_sigma=0.8,
                                  centers=np.array(centers), kinds=np.array(kinds)))
    params = blend.parameters
    names = [p.name for p in params]
    which = []
    for pi, n in enumerate(names):
        if n == "center":
            which += [(pi, 0), (pi, 1)]
        elif n == "spectrum":
            which += [(pi, 1)]
    fd = []
    for pi, idx in which:
        p = params[pi]
        eps = 1e-2 * max(abs(float(p[idx])), 1e-2) if p.dtype == np.float32 else 1e-4
        fd.append(_finite_diff(blend, obs, params, [(pi, idx)], eps=eps)[0])
    out["fd_which"] = np.array(which)
    out["fd_grad"] = np.array(fd)
    out["param_names"] = np.array(names)
    # point-source morphology at a few sub-pixel centres (reference PointSourceMorphology.get_model)
    k = kinds.index("P")
    morph = sources[k].children[1]
    cen = np.asarray(morph.center).copy()
    offs = np.array([[0.0, 0.0], [0.3, -0.2], [-0.49, 0.49], [0.11, 0.07]])
    out["ps_index"] = k
    out["ps_offsets"] = offs
    out["ps_models"] = np.stack([np.asarray(morph.get_model(sc.parameter.Parameter(cen + o, name="center")))
                                 for o in offs])
    save("point_extended.npz", **out)


def prox_chain(sc):
    rng = np.random.default_rng(2024)
    out = {}
    n = 0
    for shape in [(41, 41), (21, 21), (20, 31)]:
        yy, xx = np.meshgrid(np.arange(shape[0]) - shape[0] // 2, np.arange(shape[1]) - shape[1] // 2, indexing="ij")
        base = np.exp(-np.hypot(yy, xx) / 4.0)
        for kind in ("angle", "flat", "nearest"):
            for sym in (False, True):
                for min_grad in (0.0, 0.1):
                    X = base + 0.1 * rng.standard_normal(shape)
                    cons = [sc.constraint.MonotonicityConstraint(neighbor_weight=kind, min_gradient=min_grad)]
                    if sym:
                        cons.append(sc.constraint.SymmetryConstraint())
                    cons += [sc.constraint.PositivityConstraint(), sc.constraint.CenterOnConstraint(),
                             sc.constraint.NormalizationConstraint("max")]
                    chain = sc.constraint.ConstraintChain(*cons)
                    out["in%d" % n] = X.copy()
                    out["out%d" % n] = np.asarray(chain(X.copy(), 0))
                    out["cfg%d" % n] = np.array([kind, str(int(sym)), str(min_grad)])
                    n += 1
    out["n"] = n
    save("prox_chain.npz", **out)


def monotonic_weights(sc):
    out = {}
    n = 0
    for shape in [(5, 5), (7, 9), (10, 8), (21, 21), (6, 6)]:
        for kind in ("flat", "angle", "nearest"):
            for center in (None, (shape[0] // 2, shape[1] // 2), (1, 2)):
                w = sc.operator.getRadialMonotonicWeights(shape, neighbor_weight=kind, center=center)
                out["w%d" % n] = w
                out["cfg%d" % n] = np.array([shape[0], shape[1], -1 if center is None else center[0],
                                             -1 if center is None else center[1]])
                out["kind%d" % n] = np.array(kind)
                n += 1
    out["n"] = n
    save("monotonic_weights.npz", **out)


def multires_inputs(seed=7):
    """Plain arrays of the two-observation multi-resolution scenario (also used by the tests to rebuild the scene)."""
    def gauss(P, sig):
        y, x = np.mgrid[:P, :P] - P // 2
        out = np.stack([np.exp(-(x * x + y * y) / (2 * s * s)) for s in sig])
        return out / out.sum(axis=(1, 2))[:, None, None]

    rng = np.random.default_rng(seed)
    return dict(hr_images=rng.standard_normal((3, 40, 40)).astype(np.float32), lr_images=rng.standard_normal((5, 8, 8)).astype(np.float32),
                hr_psfs=gauss(15, [1.8, 2.0, 2.2]), lr_psfs=gauss(11, [1.0, 1.1, 1.2, 1.3, 1.4]),
                hr_cd=np.diag([0.03, 0.03]), lr_cd=np.diag([0.2, 0.2]), hr_crpix=np.array([19.5, 19.5]), lr_crpix=np.array([3.5, 3.5]),
                lr_weights=rng.uniform(0.5, 2.0, (5, 8, 8)).astype(np.float32), hr_weights=rng.uniform(0.5, 2.0, (3, 40, 40)).astype(np.float32))


def multires(sc):
    """Two observations on different pixel grids (5 bands at 0.2"/px, 3 bands at 0.03"/px): Frame.from_observations,
    ResolutionRenderer set-up and render through the reference's own code (affine WCS stand-in: astropy is absent)."""
    from scarlet_b200.wcs import AffineWCS
    RefWCS = type("RefWCS", (AffineWCS, sys.modules["astropy.wcs"].WCS), {})
    inp = multires_inputs()
    out = dict(inp)
    for dtype, tag in ((np.float32, ""), (np.float64, "64")):
        obs_hr = sc.observation.Observation(inp["hr_images"].copy(), psf=sc.psf.ImagePSF(inp["hr_psfs"].copy()), weights=inp["hr_weights"].copy(),
                                            wcs=RefWCS(inp["hr_cd"], crpix=inp["hr_crpix"]), channels=["h0", "h1", "h2"])
        obs_lr = sc.observation.Observation(inp["lr_images"].copy(), psf=sc.psf.ImagePSF(inp["lr_psfs"].copy()), weights=inp["lr_weights"].copy(),
                                            wcs=RefWCS(inp["lr_cd"], crpix=inp["lr_crpix"]), channels=["l0", "l1", "l2", "l3", "l4"])
        frame = sc.frame.Frame.from_observations([obs_lr, obs_hr], coverage="union")
        if dtype is np.float64:  # same frame in float64: the reference then keeps its resampling operator in float64
            frame = sc.frame.Frame(frame.shape, channels=frame.channels, psf=frame.psf, wcs=frame.wcs, dtype=np.float64)
            obs_lr.match(frame)
            obs_hr.match(frame)
        r, r2 = obs_lr.renderer, obs_hr.renderer
        assert type(r).__name__ == "ResolutionRenderer" and type(r2).__name__ == "ConvolutionRenderer"
        rng = np.random.default_rng(11)
        yy, xx = np.mgrid[:frame.shape[1], :frame.shape[2]]
        model = np.stack([rng.uniform(1, 5) * np.exp(-((yy - rng.uniform(25, 50)) ** 2 + (xx - rng.uniform(25, 50)) ** 2) / (2 * rng.uniform(2, 5) ** 2))
                          for _ in range(frame.shape[0])]) + 0.01 * rng.random(frame.shape)
        model = model.astype(dtype)
        out.update({"frame_shape" + tag: np.array(frame.shape), "model_psf" + tag: frame.psf.get_model(), "model" + tag: model,
                    "lr_h" + tag: np.array(r.h), "lr_fft_shape" + tag: np.array(r._fft_shape), "lr_shifts" + tag: np.array(r.shifts),
                    "lr_diff_kernel" + tag: np.asarray(r.diff_kernel.image), "lr_small_axis" + tag: np.array(r.small_axis),
                    "lr_rendered" + tag: obs_lr.render(model), "lr_logL" + tag: np.array(obs_lr.get_log_likelihood(model)),
                    "hr_diff_kernel" + tag: np.asarray(r2.diff_kernel.image), "hr_rendered" + tag: obs_hr.render(model),
                    "hr_logL" + tag: np.array(obs_hr.get_log_likelihood(model)),
                    "hr_model_slice_start" + tag: np.array([r2.slices[1][1].start, r2.slices[1][2].start]),
                    "model_crpix" + tag: np.array(frame.wcs.wcs.crpix)})
    # coverage="intersection" with the high-resolution observation named as the reference (obs_id=1), float64 frame: compact products
    obs_hr = sc.observation.Observation(inp["hr_images"].copy(), psf=sc.psf.ImagePSF(inp["hr_psfs"].copy()), weights=inp["hr_weights"].copy(),
                                        wcs=RefWCS(inp["hr_cd"], crpix=inp["hr_crpix"]), channels=["h0", "h1", "h2"])
    obs_lr = sc.observation.Observation(inp["lr_images"].copy(), psf=sc.psf.ImagePSF(inp["lr_psfs"].copy()), weights=inp["lr_weights"].copy(),
                                        wcs=RefWCS(inp["lr_cd"], crpix=inp["lr_crpix"]), channels=["l0", "l1", "l2", "l3", "l4"])
    frame = sc.frame.Frame.from_observations([obs_lr, obs_hr], obs_id=1, coverage="intersection")
    frame = sc.frame.Frame(frame.shape, channels=frame.channels, psf=frame.psf, wcs=frame.wcs, dtype=np.float64)
    obs_lr.match(frame)
    obs_hr.match(frame)
    r, r2 = obs_lr.renderer, obs_hr.renderer
    rng = np.random.default_rng(21)
    yy, xx = np.mgrid[:frame.shape[1], :frame.shape[2]]
    model = np.stack([rng.uniform(1, 5) * np.exp(-((yy - rng.uniform(10, 30)) ** 2 + (xx - rng.uniform(10, 30)) ** 2) / (2 * rng.uniform(2, 5) ** 2))
                      for _ in range(frame.shape[0])]) + 0.01 * rng.random(frame.shape)
    out.update(isect_frame_shape=np.array(frame.shape), isect_model_crpix=np.array(frame.wcs.wcs.crpix), isect_model=model,
               isect_lr_renderer=np.array(type(r).__name__), isect_hr_renderer=np.array(type(r2).__name__),
               isect_lr_h=np.array(r.h), isect_lr_fft_shape=np.array(r._fft_shape), isect_lr_shifts=np.array(r.shifts),
               isect_lr_rendered=obs_lr.render(model), isect_lr_logL=np.array(obs_lr.get_log_likelihood(model)),
               isect_hr_rendered=obs_hr.render(model), isect_hr_logL=np.array(obs_hr.get_log_likelihood(model)))
    save("multires.npz", **out)


def multires_rot(sc):
    """The ROTATED branch of ResolutionRenderer (renderer.py:318-363, 498-524): the low-resolution grid of ``multires`` turned by
    25 degrees.  One environment patch: renderer.py:443 builds ``np.array(mk_shifter(...))`` from two arrays of different
    length, which NumPy >= 1.24 refuses (older NumPy made an object array); the reference is handed an object array of its own
    two shifters.  Nothing else is touched."""
    from scarlet_b200.wcs import AffineWCS
    RefWCS = type("RefWCS", (AffineWCS, sys.modules["astropy.wcs"].WCS), {})
    orig = sc.interpolation.mk_shifter

    def mk_shifter(shape, real=False):
        sy, sx = orig(shape, real=real)
        out = np.empty(2, dtype=object)
        out[0], out[1] = sy, sx
        return out

    sc.interpolation.mk_shifter = mk_shifter
    try:
        inp = multires_inputs()
        th = np.deg2rad(25.0)
        lr_cd = 0.2 * np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        out = dict(inp)
        out["lr_cd"] = lr_cd
        obs_hr = sc.observation.Observation(inp["hr_images"].copy(), psf=sc.psf.ImagePSF(inp["hr_psfs"].copy()), weights=inp["hr_weights"].copy(),
                                            wcs=RefWCS(inp["hr_cd"], crpix=inp["hr_crpix"]), channels=["h0", "h1", "h2"])
        obs_lr = sc.observation.Observation(inp["lr_images"].copy(), psf=sc.psf.ImagePSF(inp["lr_psfs"].copy()), weights=inp["lr_weights"].copy(),
                                            wcs=RefWCS(lr_cd, crpix=inp["lr_crpix"]), channels=["l0", "l1", "l2", "l3", "l4"])
        frame = sc.frame.Frame.from_observations([obs_lr, obs_hr], coverage="union")
        frame = sc.frame.Frame(frame.shape, channels=frame.channels, psf=frame.psf, wcs=frame.wcs, dtype=np.float64)
        obs_lr.match(frame)
        obs_hr.match(frame)
        r, r2 = obs_lr.renderer, obs_hr.renderer
        assert type(r).__name__ == "ResolutionRenderer" and r.isrot and type(r2).__name__ == "ConvolutionRenderer"
        rng = np.random.default_rng(12)
        yy, xx = np.mgrid[:frame.shape[1], :frame.shape[2]]
        model = np.stack([rng.uniform(1, 5) * np.exp(-((yy - rng.uniform(30, 60)) ** 2 + (xx - rng.uniform(30, 60)) ** 2) / (2 * rng.uniform(2, 5) ** 2))
                          for _ in range(frame.shape[0])]) + 0.01 * rng.random(frame.shape)
        out.update(frame_shape=np.array(frame.shape), model_psf=frame.psf.get_model(), model=model, lr_h=np.array(r.h),
                   lr_angle=np.array([float(r.angle[0]), float(r.angle[1])]), lr_fft_shape=np.array(r._fft_shape),
                   lr_shifts=np.array(r.shifts), lr_other_shifts=np.array(r.other_shifts), lr_small_axis=np.array(r.small_axis),
                   lr_diff_kernel=np.asarray(r.diff_kernel.image), lr_rendered=obs_lr.render(model),
                   lr_logL=np.array(obs_lr.get_log_likelihood(model)), hr_rendered=obs_hr.render(model),
                   hr_model_slice_start=np.array([r2.slices[1][1].start, r2.slices[1][2].start]), model_crpix=np.array(frame.wcs.wcs.crpix))
        save("multires_rot.npz", **out)
    finally:
        sc.interpolation.mk_shifter = orig


def source_recipes(sc):
    """The other branches of the reference's ``ExtendedSource`` factory on data/hsc_cosmos_35.npz: a two-component source
    (K=2) and a compact one, initialised by the reference."""
    d = np.load(os.path.join(REF_DATA, "hsc_cosmos_35.npz"))
    images, variance, psfs = d["images"], d["variance"], d["psfs"]
    channels = [str(f) for f in d["filters"]]
    centers = [(float(s["y"]), float(s["x"])) for s in d["catalog"]][:3]
    frame = sc.frame.Frame(images.shape, psf=sc.psf.GaussianPSF(sigma=(0.8,) * len(channels)), channels=channels)
    obs = sc.observation.Observation(images, psf=sc.psf.ImagePSF(psfs.copy()), weights=1 / variance, channels=channels)
    obs.match(frame)
    out = dict(centers=np.array(centers))
    multi = sc.source.ExtendedSource(frame, centers[1], obs, K=2)
    for k, comp in enumerate(multi.children):
        out["multi%d_spectrum" % k] = np.asarray(comp.parameters[0])
        out["multi%d_image" % k] = np.asarray(comp.parameters[1])
        out["multi%d_origin" % k] = np.array(comp.bbox.origin)
        out["multi%d_spectrum_step" % k] = np.asarray(comp.parameters[0].step(comp.parameters[0], it=0))
    compact = sc.source.ExtendedSource(frame, centers[2], obs, compact=True)
    out.update(compact_spectrum=np.asarray(compact.parameters[0]), compact_image=np.asarray(compact.parameters[1]),
               compact_origin=np.array(compact.bbox.origin))
    save("source_recipes.npz", **out)


def init_helpers(sc):
    import importlib
    """The reference's user-facing initialisation helpers on data/hsc_cosmos_35.npz: ``get_psf_spectrum`` (with SNR) at the
    first five catalogue positions, the component count ``init_source`` settles on, and ``set_spectra_to_match`` applied to
    three single-component sources (inputs: the rendered unit-spectrum models, outputs: the solved spectra)."""
    d = np.load(os.path.join(REF_DATA, "hsc_cosmos_35.npz"))
    images, variance, psfs = d["images"], d["variance"], d["psfs"]
    channels = [str(f) for f in d["filters"]]
    centers = [(float(s["y"]), float(s["x"])) for s in d["catalog"]][:5]
    frame = sc.frame.Frame(images.shape, psf=sc.psf.GaussianPSF(sigma=(0.8,) * len(channels)), channels=channels)
    weights = 1 / variance
    weights[:, 5:9, 30:34] = 0  # a masked patch, so that the mask handling is exercised
    obs = sc.observation.Observation(images, psf=sc.psf.ImagePSF(psfs.copy()), weights=weights, channels=channels)
    obs.match(frame)
    out = dict(centers=np.array(centers), weights=weights.astype(np.float32))
    spec, snr = [], []
    for c in centers:
        a, b = sc.initialization.get_psf_spectrum(c, obs, compute_snr=True)
        spec.append(a), snr.append(b)
    out["psf_spectrum"], out["psf_snr"] = np.array(spec), np.array(snr)
    # detection image of the first source with the masked patch in the weights (initialization.py:213-284)
    pix_spectrum = sc.initialization.get_pixel_spectrum(centers[0], obs)
    out["detect"], out["detect_std"] = (a.astype(np.float32) for a in sc.initialization.build_initialization_image(obs, spectra=pix_spectrum))
    out["detect_flat"], out["detect_flat_std"] = (a.astype(np.float32) for a in sc.initialization.build_initialization_image(obs))
    del obs._detect
    out["edge_spectrum"] = sc.initialization.get_psf_spectrum((1.0, 2.0), obs)  # PSF box sticks out of the image
    # component counts of init_source for a few min_snr values (the cap is floor(psf_snr / min_snr))
    ks = []
    for min_snr in (50, 200, 1000):
        row = []
        for c in centers[:3]:
            src = sc.initialization.init_source(frame, c, obs, max_components=2, min_snr=min_snr)
            row.append(len(src.children) if isinstance(src, sc.component.CombinedComponent) else
                       (0 if type(src).__name__ == "CompactExtendedSource" else 1))
        ks.append(row)
    out["init_source_K"] = np.array(ks)
    sources = [sc.source.ExtendedSource(frame, c, obs, resizing=False) for c in centers[:3]]
    for k, src in enumerate(sources):
        src.parameters[0][:] = 1
        out["src%d_image" % k] = np.asarray(src.parameters[1])
        out["src%d_origin" % k] = np.array(src.bbox.origin)
    out["unit_rendered"] = np.stack([obs.render(src.get_model(frame=frame)) for src in sources]).astype(np.float32)
    sc.initialization.set_spectra_to_match(sources, obs)
    out["matched_spectra"] = np.stack([np.asarray(src.parameters[0]) for src in sources])
    # spectra of the PSF-shaped recipes keep the dtype of the data (in-place divisions, initialization.py:58-67, source.py:119,302)
    clean = sc.observation.Observation(images, psf=sc.psf.ImagePSF(psfs.copy()), weights=1 / variance, channels=channels)
    clean.match(frame)
    point = sc.source.PointSource(frame, centers[0], clean)
    out["point_spectrum"], out["point_center"] = np.asarray(point.parameters[0]), np.asarray(point.parameters[1])
    out["point_spectrum_step"] = np.asarray(point.parameters[0].step(point.parameters[0], it=0))
    compact = sc.source.ExtendedSource(frame, centers[2], clean, compact=True)
    out["compact_spectrum"], out["compact_image"] = np.asarray(compact.parameters[0]), np.asarray(compact.parameters[1])
    out["compact_origin"] = np.array(compact.bbox.origin)
    # functional PSF models (psf.py:80-201)
    moffat = sc.psf.MoffatPSF(alpha=[4.7, 3.0, 2.2], beta=[1.5, 2.5, 3.0], boxsize=21)
    out["moffat"], out["moffat_offset"] = moffat.get_model(), moffat.get_model(offset=(0.3, -0.2))
    out["moffat_same"] = sc.psf.MoffatPSF(alpha=[2.0, 2.0], beta=[2.0, 2.0]).get_model()
    out["imagepsf_offset"] = sc.psf.ImagePSF(psfs.copy()).get_model(offset=(0.3, -0.45))
    out["gauss_offset"] = sc.psf.GaussianPSF(sigma=[0.8, 1.3], boxsize=11).get_model(offset=(0.25, -0.4))
    # off-centre symmetry operators of the source initialisation (operator.py:207-271)
    ref_operator = importlib.import_module(sc.__name__ + ".operator")
    rng = np.random.default_rng(207)
    sym_cases = [((9, 11), None, None), ((9, 11), (2, 7), None), ((9, 11), (6, 3), 0.0), ((8, 10), (5, 2), None), ((8, 10), (1, 8), -1.0),
                 ((7, 7), (3, 3), None)]
    out["sym_shapes"] = np.array([c[0] for c in sym_cases])
    out["sym_centers"] = np.array([(-1, -1) if c[1] is None else c[1] for c in sym_cases])
    out["sym_fills"] = np.array([np.nan if c[2] is None else c[2] for c in sym_cases])
    for i, (shape, center, fill) in enumerate(sym_cases):
        X = rng.random(shape)
        out["sym%d_in" % i] = X.copy()
        out["sym%d_out" % i] = np.array(ref_operator.prox_uncentered_symmetry(X.copy(), 0, center=center, algorithm="sdss", fill=fill))
    # box trimming (initialization.py:173-210): blob near the centre index, explicit box size, centre outside the support, NaN
    yy, xx = np.mgrid[:40, :46]
    blob = np.exp(-((yy - 17) ** 2 / 30.0 + (xx - 25) ** 2 / 18.0))
    blob[3, 4] = np.nan
    trim_cases = [((17, 25), 0.01, None), ((17, 25), 0.2, None), ((16, 27), 1e-4, None), ((17, 25), 0.01, 11), ((2, 40), 0.5, None)]
    out["trim_in"] = blob
    out["trim_args"] = np.array([(c[0][0], c[0][1], c[1], -1 if c[2] is None else c[2]) for c in trim_cases])
    for i, (ci, thr, bs) in enumerate(trim_cases):
        m, bb = sc.initialization.trim_morphology(ci, blob.copy(), bg_thresh=thr, boxsize=bs)
        out["trim%d_out" % i], out["trim%d_origin" % i] = m, np.array(bb.origin)
    # dynamic box of an image morphology (morphology.py:52-68,132-207): shrink, grow, stay
    rng = np.random.default_rng(132)
    yy, xx = np.mgrid[:31, :31]
    def blob(sig):
        return np.exp(-((yy - 15) ** 2 + (xx - 15) ** 2) / (2.0 * sig ** 2))
    box_cases = []
    small = blob(1.5)
    small[small < 1e-3] = 0                       # empty outer rings -> shrink
    box_cases.append((small, 1e-3 * rng.standard_normal((31, 31)), np.full((31, 31), 1e-2)))
    wide = blob(9.0)                               # flux at the edges, optimiser pulling outwards -> grow
    m_out = -np.ones((31, 31)) * 5.0
    box_cases.append((wide, m_out, np.full((31, 31), 1e-4)))
    mid = blob(4.0)                                # nothing to do
    v_mid = np.full((31, 31), 1e-2)
    v_mid[0, :] = 0                                # zero second moments are masked out of the edge statistics
    box_cases.append((mid, 1e-3 * rng.standard_normal((31, 31)), v_mid))
    for i, (img, m, v) in enumerate(box_cases):
        par = sc.parameter.Parameter(img.copy(), name="image", step=1e-2, m=m.copy(), v=v.copy(), vhat=v.copy() * 2)
        fr = sc.frame.Frame((1, 61, 61), channels=["r"])
        morph = sc.morphology.ImageMorphology(fr, par, bbox=sc.bbox.Box((31, 31), origin=(10, 12)), resizing=True)
        try:
            morph.update()
            changed = 0
        except sc.model.UpdateException:
            changed = 1
        new = morph.parameters[0]
        out["box%d_image" % i], out["box%d_m" % i], out["box%d_v" % i] = img, m, v
        out["box%d_changed" % i] = np.array(changed)
        out["box%d_new_image" % i], out["box%d_new_m" % i] = np.asarray(new), np.asarray(new.m)
        out["box%d_new_v" % i], out["box%d_new_vhat" % i] = np.asarray(new.v), np.asarray(new.vhat)
        out["box%d_new_origin" % i], out["box%d_new_step" % i] = np.array(morph.bbox.origin), np.array(float(new.step))
    # image moments (measure.py:108-150) of a small cube and of a single image
    ref_measure = importlib.import_module(sc.__name__ + ".measure")
    rng = np.random.default_rng(108)
    cube, wgt = rng.random((3, 9, 7)), rng.random((3, 9, 7))
    out["mom_cube"], out["mom_weight"] = cube, wgt
    for name, args in (("default", {}), ("centroid", dict(centroid=np.array([2.5, 4.25]))), ("weighted", dict(N=3, weight=wgt))):
        M = ref_measure.moments(cube, **args)
        out["mom_%s_keys" % name] = np.array(sorted(M))
        out["mom_%s_vals" % name] = np.array([M[k] for k in sorted(M)])
    # matched-filter SNR of a model cube with zero-weight pixels under the source (measure.py:60-104)
    snr_model = sources[0].get_model(frame=frame)
    out["snr_rendered"] = np.asarray(obs.render(snr_model), dtype=np.float32)
    out["snr_value"] = np.array(ref_measure.snr(snr_model, obs))
    shifted = np.roll(snr_model, (-26, 18), axis=(1, 2))  # the same source moved onto the masked patch
    out["snr_rendered_masked"] = np.asarray(obs.render(shifted), dtype=np.float32)
    out["snr_value_masked"] = np.array(ref_measure.snr(shifted, obs))
    M2 = ref_measure.moments(cube[0], N=1)
    out["mom_image_vals"] = np.array([M2[k] for k in sorted(M2)])
    save("init_helpers.npz", **out)


def psf_shift(sc):
    """``ConvolutionRenderer(psf_shift=...)`` (renderer.py:172-177, 220-227, 252-254; fft.shift fft.py:399-428) through the
    reference's own code: the shifted difference kernel, the rendered model, logL and central differences of logL wrt the shift
    (float64 frame, as in ``_finite_diff``)."""
    rng = np.random.default_rng(21)
    C, Ny, Nx, P = 3, 34, 30, 15
    channels = list(range(C))
    model_psf = sc.psf.GaussianPSF(sigma=(0.8,) * C)
    frame = sc.frame.Frame((C, Ny, Nx), psf=model_psf, channels=channels, dtype=np.float64)
    y, x = np.mgrid[:P, :P] - P // 2
    psfs = np.stack([(1 + (x * x + 1.3 * y * y + 0.4 * x * y) / (1.1 + 0.25 * c) ** 2) ** -2.5 for c in range(C)])
    psfs /= psfs.sum(axis=(1, 2))[:, None, None]
    model = np.zeros((C, Ny, Nx))
    seds, morphs = [], []
    for k in range(4):
        cy, cx = rng.uniform(6, Ny - 6), rng.uniform(6, Nx - 6)
        yy, xx = np.mgrid[:Ny, :Nx]
        seds.append(rng.uniform(5, 50, C))
        morphs.append(np.exp(-np.hypot(yy - cy, xx - cx) / rng.uniform(1.0, 2.5)))
        model += seds[-1][:, None, None] * morphs[-1][None]
    images = rng.standard_normal((C, Ny, Nx)) + model * 0.9
    weights = rng.uniform(0.5, 2.0, (C, Ny, Nx))
    out = dict(psfs=psfs, model=model, images=images, weights=weights, model_sigma=0.8, seds=np.array(seds), morphs=np.array(morphs))
    for tag, shift in (("a", (0.3, -0.45)), ("b", (-1.2, 0.8))):
        obs = sc.observation.Observation(images.copy(), psf=sc.psf.ImagePSF(psfs.copy()), weights=weights.copy(), channels=channels)
        renderer = sc.renderer.ConvolutionRenderer(obs, frame, psf_shift=np.array(shift))
        obs.match(frame, renderer=renderer)
        sh = renderer.parameters[0]
        kernel = sc.fft.shift(renderer.diff_kernel.image, np.asarray(sh), fft_shape=None, axes=(-2, -1), return_Fourier=True).image
        rendered = obs.render(model, *obs.parameters)
        logL = obs.get_log_likelihood(model, *obs.parameters)
        fd = []
        for j in range(2):
            vals = []
            for sgn in (+1, -1):
                sh[j] = shift[j] + sgn * 1e-4  # the renderer reads its own Parameter (model.py:71-110)
                vals.append(-obs.get_log_likelihood(model, *obs.parameters))
            sh[j] = shift[j]
            fd.append((vals[0] - vals[1]) / 2e-4)
        out.update({"shift_" + tag: np.array(shift), "kernel_" + tag: np.asarray(kernel), "rendered_" + tag: np.asarray(rendered),
                    "logL_" + tag: logL, "dloss_dshift_" + tag: np.array(fd), "step_" + tag: float(sh.step)})
        out["diff_kernel"] = np.asarray(renderer.diff_kernel.image)
    save("psf_shift.npz", **out)


if __name__ == "__main__":
    sc = ref_shim.install()
    which = sys.argv[1:] or ["obs_render_loss", "hsc_cosmos_35", "point_extended", "prox_chain", "monotonic_weights", "multires",
                             "source_recipes", "init_helpers", "psf_shift", "multires_rot"]
    for name in which:
        globals()[name](sc)
