"""GPU: the CUDA path (through the C ABI) against the CPU oracle and the reference-derived golden fixtures.

Tolerances.  Integer/index work and the monotonic sweep: bit-exact.  Floating point: the north-star bar is
<= 1e-5 relative in float32 on model pixels and SEDs; "relative" is measured against the peak of the reference
array (max |a-b| <= tol * max |b|): the model cube, a morphology image, and -- for spectra -- the K x C matrix of
all SEDs of the scene.  (A float32 FFT leaves noise of ~1.4e-7 of the SCENE peak in every pixel
(tools/fft_accuracy_probe.py), so after tens of iterations a source 100x fainter than the brightest one cannot
agree to 1e-5 of its OWN amplitude; its own-scale error is bounded at 1e-4 here to still catch real defects.)
The float64 twin of every kernel must agree with the float64 oracle to ~1e-9, which separates algorithmic
differences from float32 rounding.
"""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from conftest import ARANGE25, MONO_KATS, SYM_HALF, golden

pytestmark = pytest.mark.gpu


def tracks_rounding(ours, ref64, ref32, floor=1e-5, factor=3.0):
    """float32 product vs the float64 reference arithmetic, judged against what float32 rounding ALONE does to the reference
    algorithm (ref32 = the oracle under ``float32_arithmetic()``): max |ours - ref64| <= max(floor, factor |ref32 - ref64|),
    all relative to the peak of ref64.  Returns (our error, rounding error) for the assertion message."""
    return rel_peak(ours, ref64), rel_peak(ref32, ref64)


def rel_peak(a, b):
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / max(np.abs(b).max(), 1e-300))


# --------------------------------------------------------------------------------------------------
# monotonic wavefront kernel: bit-exact against the sequential sweep
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,min_grad,expect", MONO_KATS)
def test_monotonic_reference_kat(kind, min_grad, expect):
    import scarlet_b200 as sb
    out = sb.MonotonicityConstraint(neighbor_weight=kind, min_gradient=min_grad)(ARANGE25.copy(), 0)
    assert_allclose(out, expect, atol=5e-8)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(5, 5), (21, 21), (41, 41), (81, 81), (20, 31), (1, 1), (2, 3)])
@pytest.mark.parametrize("kind", ["angle", "flat", "nearest"])
def test_monotonic_bit_exact(dtype, shape, kind):
    from oracle import monotonic_c
    from scarlet_b200 import operator
    from scarlet_b200.operators_pybind11 import prox_weighted_monotonic
    rng = np.random.default_rng(hash((shape, kind)) % 2 ** 31)
    center = (shape[0] // 2, shape[1] // 2)
    w, off, idx = operator.monotonic_tables(shape, kind, center)
    w = w.astype(dtype)
    for mg in (0.0, 0.1):
        X = (np.exp(-np.hypot(*np.mgrid[:shape[0], :shape[1]]) / 5.0) + 0.2 * rng.standard_normal(shape)).astype(dtype)
        a, b = X.copy().reshape(-1), X.copy().reshape(-1)
        monotonic_c.sweep(a, w, off, idx, mg)
        prox_weighted_monotonic(b, w, off, idx, mg)
        assert_array_equal(a, b)


def test_monotonic_batch_and_arbitrary_order():
    """a batch of images with one operator; and a scrambled dist_idx (hazard-exact scheduling)"""
    from oracle import monotonic_c
    from scarlet_b200 import operator
    from scarlet_b200.operators_pybind11 import prox_weighted_monotonic
    rng = np.random.default_rng(5)
    shape = (15, 15)
    w, off, idx = operator.monotonic_tables(shape, "angle", (7, 7))
    X = rng.random((6, 225))
    ref = X.copy()
    for r in ref:
        monotonic_c.sweep(r, w, off, idx, 0.05)
    out = X.copy()
    prox_weighted_monotonic(out, w, off, idx, 0.05)
    assert_array_equal(out, ref)
    scr = rng.permutation(idx).astype(np.int32)
    a, b = X[0].copy(), X[0].copy()
    monotonic_c.sweep(a, w, off, scr, 0.0)
    prox_weighted_monotonic(b, w, off, scr, 0.0)
    assert_array_equal(a, b)


def test_batch_pipeline_matches_plain_batches():
    """BatchPipeline (copy-in / loop / copy-out of consecutive batches overlapped) gives exactly the results of fitting the
    same batches one after the other, also when a batch object comes round a second time (warm restart)."""
    from scarlet_b200 import BatchPipeline, BlendBatch, synthetic

    def make(seed0):
        return [synthetic.make_blend(synthetic.make_scene("tiny", seed0 + i)) for i in range(3)]

    ref_a, ref_b = BlendBatch(make(10)), BlendBatch(make(20))
    ref = [ref_a.fit(max_iter=12, e_rel=1e-6, upload_observations=True), ref_b.fit(max_iter=12, e_rel=1e-6, upload_observations=True),
           ref_a.fit(max_iter=12, e_rel=1e-6, upload_observations=True)]
    a, b = BlendBatch(make(10)), BlendBatch(make(20))
    got = BatchPipeline(depth=2).run([a, b, a], max_iter=12, e_rel=1e-6)
    assert got == ref
    for x, y in ((a, ref_a), (b, ref_b)):
        for bx, by in zip(x.blends, y.blends):
            assert bx.loss == by.loss
            for px, py in zip(bx.parameters, by.parameters):
                assert_array_equal(np.asarray(px), np.asarray(py))
                assert_array_equal(px.v, py.v)
    for x in (a, b, ref_a, ref_b):
        x.close()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_apply_filter_bit_exact(dtype):
    """the reference's second native entry point (operators_pybind11.cc:39-56): real-space filter as a sum of shifted,
    scaled blocks -- against the same loop in NumPy (separate multiply and add per tap, image dtype), bit for bit"""
    from scarlet_b200.operators_pybind11 import apply_filter
    rng = np.random.default_rng(9)
    img = rng.standard_normal((23, 31)).astype(dtype)
    ker = rng.standard_normal((5, 7)).astype(dtype)
    dy, dx = np.mgrid[:5, :7]
    dy, dx = (dy - 2).reshape(-1), (dx - 3).reshape(-1)
    z = np.zeros(dy.size, dtype=int)
    ys, ye, xs, xe = np.maximum(z, dy), -np.minimum(z, dy), np.maximum(z, dx), -np.minimum(z, dx)
    ref = np.zeros_like(img)
    for n, v in enumerate(ker.reshape(-1)):
        rows, cols = img.shape[0] - ys[n] - ye[n], img.shape[1] - xs[n] - xe[n]
        ref[ys[n]:ys[n] + rows, xs[n]:xs[n] + cols] += v * img[ye[n]:ye[n] + rows, xe[n]:xe[n] + cols]
    out = np.full_like(img, np.nan)
    apply_filter(img, ker.reshape(-1), ys, ye, xs, xe, out)
    assert_array_equal(out, ref)
    from scipy import signal
    assert_allclose(out, signal.convolve2d(img.astype(np.float64), ker.astype(np.float64), mode="same"), atol=1e-4 if dtype == np.float32 else 1e-12)


def test_native_errors_are_reported():
    from scarlet_b200 import _native
    from scarlet_b200.operators_pybind11 import prox_weighted_monotonic
    w = np.ones((8, 9))
    off = np.array([-4, -3, -2, -1, 1, 2, 3, 4], dtype=np.int32)
    with pytest.raises(_native.NativeError):  # neighbour outside the image
        prox_weighted_monotonic(np.zeros(9), w, off, np.arange(9, dtype=np.int32), 0.0)
    with pytest.raises(TypeError):
        prox_weighted_monotonic(np.zeros(9, dtype=np.int32), w, off, np.arange(9, dtype=np.int32), 0.0)


# --------------------------------------------------------------------------------------------------
# constraint chain
# --------------------------------------------------------------------------------------------------
def test_symmetry_and_elementwise_kats():
    import scarlet_b200 as sb
    assert_allclose(sb.SymmetryConstraint()(ARANGE25.copy(), 0), np.full((5, 5), 12.0))
    assert_allclose(sb.SymmetryConstraint(0.5)(ARANGE25.copy(), 0), SYM_HALF)
    rng = np.random.default_rng(0)
    X = rng.standard_normal((6, 7))
    assert (sb.PositivityConstraint(0.1)(X.copy(), 0) >= 0.1).all()
    Y = np.abs(X) + 0.1
    assert_allclose(sb.NormalizationConstraint("sum")(Y.copy(), 0).sum(), 1)
    assert_allclose(sb.NormalizationConstraint("max")(Y.copy(), 0).max(), 1)
    assert sb.CenterOnConstraint()(np.zeros((5, 5)), 0)[2, 2] > 0
    # even-sized symmetry (zero line appended before the flip, operator.py:281-288)
    from oracle import scarlet_oracle as so
    E = rng.random((6, 5))
    assert_allclose(sb.SymmetryConstraint(0.7)(E.copy(), 0), so.prox_symmetry(E.copy(), 0.7), atol=1e-15)


def test_chain_vs_reference_fixture():
    import scarlet_b200 as sb
    g = golden("prox_chain.npz")
    for i in range(int(g["n"])):
        kind, sym, mg = g["cfg%d" % i]
        cons = [sb.MonotonicityConstraint(neighbor_weight=str(kind), min_gradient=float(mg))]
        if int(sym):
            cons.append(sb.SymmetryConstraint())
        cons += [sb.PositivityConstraint(), sb.CenterOnConstraint(), sb.NormalizationConstraint("max")]
        chain = sb.ConstraintChain(*cons)
        out64 = chain(g["in%d" % i].copy(), 0)
        assert_allclose(out64, g["out%d" % i], rtol=0, atol=1e-14, err_msg="case %d" % i)
        out32 = chain(g["in%d" % i].astype(np.float32), 0)
        assert rel_peak(out32, g["out%d" % i]) < 1e-6


def test_chain_idempotent_at_scale():
    """projection property on large boxes: applying the ExtendedSource chain twice changes nothing beyond rounding"""
    import scarlet_b200 as sb
    rng = np.random.default_rng(11)
    chain = sb.ConstraintChain(sb.MonotonicityConstraint("angle", 0), sb.SymmetryConstraint(), sb.PositivityConstraint(),
                               sb.CenterOnConstraint(), sb.NormalizationConstraint("max"))
    X = np.exp(-np.hypot(*(np.mgrid[:81, :81] - 40)) / 9.0) + 0.05 * rng.standard_normal((81, 81))
    once = chain(X.copy(), 0)
    twice = chain(once.copy(), 0)
    assert_allclose(twice, once, atol=1e-12)
    assert once.max() == 1.0 and (once >= 0).all()
    assert_allclose(once, once[::-1, ::-1], atol=1e-15)


def test_unknown_constraint_raises():
    import scarlet_b200 as sb
    from scarlet_b200.constraint import MonoTables, constraint_ops

    class Mine(sb.Constraint):
        pass

    with pytest.raises(TypeError, match="Mine"):
        constraint_ops(sb.ConstraintChain(sb.PositivityConstraint(), Mine()), (5, 5), MonoTables())


# --------------------------------------------------------------------------------------------------
# FFT convolution
# --------------------------------------------------------------------------------------------------
def test_observation_render_reference_scenario():
    """reference tests/test_observation.py:12-47 through the product API"""
    import scarlet_b200 as sb
    g = golden("obs_render_loss.npz")
    mpsf = sb.GaussianPSF(float(g["model_sigma"]), boxsize=int(g["model_boxsize"]))
    channels = [0, 1, 2]
    frame = sb.Frame((3, 43, 43), psf=mpsf, channels=channels)
    obs = sb.Observation(g["images"].copy(), psf=sb.GaussianPSF(g["obs_sigmas"], boxsize=int(g["obs_boxsize"])), channels=channels)
    obs.match(frame)
    rendered = obs.render(g["model"].astype(np.float32))
    assert rendered.dtype == np.float32
    assert_allclose(rendered, g["obs_psf_image"], atol=1e-6)   # the reference's own assertion (assert_almost_equal)
    assert rel_peak(rendered, g["rendered"]) < 1e-5
    logL = obs.get_log_likelihood(g["model"].astype(np.float32))
    assert_allclose(logL, float(g["logL"]), rtol=1e-6)
    r64 = obs.render(g["model"].astype(np.float64))
    assert rel_peak(r64, g["rendered"]) < 1e-6  # fixture itself carries complex64 noise


@pytest.mark.parametrize("dtype,tol", [(np.float32, 2e-6), (np.float64, 1e-13)])
def test_convolve_and_adjoint_vs_oracle(dtype, tol):
    from oracle import scarlet_oracle as so
    from scarlet_b200 import fft as sfft
    rng = np.random.default_rng(3)
    img = rng.random((4, 37, 52)).astype(dtype)
    ker = rng.random((4, 9, 11))
    fshape, khat = sfft.kernel_transform(ker, img.shape)
    assert list(fshape) == so.get_fft_shape(img.shape, ker.shape, 3, (1, 2))
    out = sfft.device_convolve(img, khat, fshape)
    ref = so.convolve(img.astype(np.float64), ker, axes=(1, 2))
    assert rel_peak(out, ref) < tol
    # adjoint: <K x, y> == <x, K^T y>
    y = rng.random(img.shape).astype(dtype)
    adj = sfft.device_convolve(y, khat, fshape, adjoint=True)
    lhs, rhs = float((out.astype(np.float64) * y).sum()), float((img.astype(np.float64) * adj).sum())
    assert abs(lhs - rhs) <= (1e-5 if dtype == np.float32 else 1e-12) * abs(lhs)


def test_convolution_linearity_full_size():
    """size-independent property at the cfg3 grid (5x256x256, 41x41 kernel, F=300): K(a+2b) = K a + 2 K b"""
    from scarlet_b200 import fft as sfft
    rng = np.random.default_rng(4)
    a = rng.random((5, 256, 256)).astype(np.float32)
    b = rng.random((5, 256, 256)).astype(np.float32)
    ker = rng.random((5, 41, 41))
    ker /= ker.sum(axis=(1, 2))[:, None, None]
    fshape, khat = sfft.kernel_transform(ker, a.shape)
    assert tuple(fshape) == (300, 300)
    lhs = sfft.device_convolve(a + 2 * b, khat, fshape)
    rhs = sfft.device_convolve(a, khat, fshape) + 2 * sfft.device_convolve(b, khat, fshape)
    assert rel_peak(lhs, rhs) < 5e-6
    assert_allclose(lhs.sum(axis=(1, 2)) / (a + 2 * b).sum(axis=(1, 2)), 1.0, atol=0.12)  # flux (up to edge losses of a flat 41x41 kernel)


# --------------------------------------------------------------------------------------------------
# scene forward / gradients
# --------------------------------------------------------------------------------------------------
def _blend_from_golden(g, precision):
    import scarlet_b200 as sb
    C = g["images"].shape[0]
    channels = list(range(C))
    mpsf = sb.GaussianPSF(sigma=(float(g["model_sigma"]),) * C)
    frame = sb.Frame(g["images"].shape, psf=mpsf, channels=channels, dtype=np.float32 if precision == 32 else np.float64)
    obs = sb.Observation(g["images"].copy(), psf=sb.ImagePSF(g["psfs"].copy()), weights=g["weights"].copy(), channels=channels)
    obs.match(frame)
    srcs = []
    for k in range(int(g["n_sources"])):
        if str(g["src%d_kind" % k]) == "PointSource":
            srcs.append(sb.PointSource(frame, g["src%d_center" % k], obs, spectrum=g["src%d_spectrum" % k].copy()))
        else:
            img = g["src%d_image" % k]
            srcs.append(sb.ExtendedSource(frame, (0, 0), obs, spectrum=g["src%d_spectrum" % k].copy(), morphology=img.copy(),
                                          bbox=sb.Box(img.shape, origin=g["src%d_origin" % k][1:]), resizing=False))
    return sb.Blend(srcs, obs, precision=precision), obs


def test_source_initialisation_vs_reference_fixture():
    """ExtendedSource(frame, sky_coord, observations) initialised from the data like the reference's
    SingleExtendedSource (peak-pixel spectrum, SNR coadd, min-symmetry, flat-weight monotonicity on the GPU, threshold,
    box trimming, PSF floor) -- spectra, morphologies and boxes of the three quickstart sources of data/hsc_cosmos_35"""
    import scarlet_b200 as sb
    g = golden("hsc_cosmos_35.npz")
    C = g["images"].shape[0]
    channels = list(range(C))
    frame = sb.Frame(g["images"].shape, psf=sb.GaussianPSF(sigma=(float(g["model_sigma"]),) * C), channels=channels)
    obs = sb.Observation(g["images"].copy(), psf=sb.ImagePSF(g["psfs"].copy()), weights=g["weights"].copy(), channels=channels)
    obs.match(frame)
    for k, center in enumerate(g["centers"]):
        src = sb.ExtendedSource(frame, tuple(center), obs, resizing=False)
        spectrum, image = src.parameters[0], src.parameters[1]
        assert src.bbox.origin == tuple(g["src%d_origin" % k]) and image.shape == g["src%d_image" % k].shape
        assert_allclose(spectrum, g["src%d_spectrum" % k], rtol=1e-6)
        assert_allclose(image, g["src%d_image" % k], atol=1e-6)
        assert_allclose(spectrum.step(spectrum, it=0), g["src%d_spectrum_step" % k], rtol=1e-5)
    pt = sb.PointSource(frame, tuple(g["centers"][0]), obs)
    peak = sb.ImagePSF(g["psfs"]).get_model().max(axis=(1, 2))
    assert_allclose(pt.parameters[0], g["images"][:, 33, 14] / peak, rtol=1e-6)
    # the other branches of the factory (source.py:759-807): two stacked components, and a compact source
    r = golden("source_recipes.npz")
    multi = sb.ExtendedSource(frame, tuple(r["centers"][1]), obs, K=2)
    assert isinstance(multi, sb.MultiExtendedSource) and len(multi.children) == 2
    for k, comp in enumerate(multi.children):
        assert comp.bbox.origin == tuple(r["multi%d_origin" % k])
        assert_allclose(comp.parameters[0], r["multi%d_spectrum" % k], rtol=1e-6)
        assert_allclose(comp.parameters[1], r["multi%d_image" % k], atol=1e-6)
        assert_allclose(comp.parameters[0].step(comp.parameters[0], it=0), r["multi%d_spectrum_step" % k], rtol=1e-5)
    compact = sb.ExtendedSource(frame, tuple(r["centers"][2]), obs, compact=True)
    assert compact.bbox.origin == tuple(r["compact_origin"])
    assert_allclose(compact.parameters[0], r["compact_spectrum"], rtol=1e-6)
    assert_allclose(compact.parameters[1], r["compact_image"], atol=1e-12)
    # a blend of all recipes fits on the device (CombinedComponent leaves are flattened into the plan)
    blend = sb.Blend([multi, compact, pt], obs)
    n, logL = blend.fit(12, e_rel=1e-4)
    assert n == 12 and np.isfinite(logL) and blend.loss[-1] < blend.loss[0]


def _quickstart(precision, arithmetic32=False):
    """-> (blend or None, oracle): the quickstart scene (3 ExtendedSource initialised from the data, dynamic boxes)"""
    import scarlet_b200 as sb
    from oracle import scarlet_oracle as so
    g = golden("hsc_cosmos_35.npz")
    C = g["images"].shape[0]
    channels = list(range(C))
    dtype = np.float32 if precision == 32 else np.float64
    frame = sb.Frame(g["images"].shape, psf=sb.GaussianPSF(sigma=(0.8,) * C), channels=channels, dtype=dtype)
    obs = sb.Observation(g["images"].copy(), psf=sb.ImagePSF(g["psfs"].copy()), weights=g["weights"].copy(), channels=channels)
    obs.match(frame)
    sources = [sb.ExtendedSource(frame, tuple(c), obs) for c in g["centers"]]

    def oracle():  # from the same starting point
        mpsf = so.GaussianPSFOracle([0.8] * C)
        oobs = so.ObservationOracle(g["images"], g["weights"], so.ImagePSFOracle(g["psfs"]), frame_dtype=dtype)
        oobs.match(g["images"].shape, mpsf)
        osrcs = [so.ExtendedSourceOracle(np.array(s.parameters[0]), np.array(s.parameters[1]), s.bbox.origin[1:],
                                         min_step=oobs.channel_noise_rms(), resizing=True, sed_dtype=s.parameters[0].dtype) for s in sources]
        return so.SceneOracle(g["images"].shape, mpsf, osrcs, [oobs], frame_dtype=dtype)

    if arithmetic32:
        with so.float32_arithmetic():
            o = oracle()
            o.fit(max_iter=50, e_rel=1e-3)
        return None, o
    return sb.Blend(sources, obs, precision=precision), oracle()


def test_quickstart_cfg1_drop_in_float64_twin():
    """BASELINE config 1, the reference's quickstart (docs/0-quickstart.ipynb cells 3-24) written against this package with
    the reference's own calls: Frame, Observation.match, ExtendedSource(frame, center, obs) initialised from the data,
    Blend(sources, obs).fit(50, e_rel) with dynamic boxes -- against the oracle started from the same initial parameters.
    Boxes overhang the 58x48 frame and one is wider than the frame."""
    import scarlet_b200 as sb
    blend, o = _quickstart(64)
    o_n, o_logL = o.fit(max_iter=50, e_rel=1e-3)
    n, logL = blend.fit(50, e_rel=1e-3)
    assert n == o_n and n == len(blend.loss)
    assert [tuple(s.parameters[1].shape) for s in blend.sources] == [tuple(s.image.x.shape) for s in o.sources]
    assert_allclose(np.array(blend.loss), np.array(o.loss), rtol=1e-8)
    assert_allclose(logL, o_logL, rtol=1e-8)
    for src, osrc in zip(blend.sources, o.sources):
        assert rel_peak(src.parameters[0], osrc.spectrum.x) < 1e-7
        assert rel_peak(src.parameters[1], osrc.image.x) < 1e-7
    assert sb.measure.flux(blend.sources[0]).shape == (blend.frame.C,)


def test_quickstart_cfg1_drop_in_float32():
    """The shipped float32 path on the quickstart.  This fit starts far from the optimum (the loss falls by 10 % per
    iteration, weights = 1/variance span orders of magnitude) and every early AMSGrad step moves a pixel by the SIGN of its
    gradient, so rounding is amplified: the oracle itself, run with float32 FFTs and float32 morphologies, leaves its own
    float64 trajectory by the amounts asserted against below.  The product has to stay within 3x of that (and within 1e-5
    where rounding alone stays below it); same iteration count and same box sequence as the reference arithmetic."""
    blend, o = _quickstart(32)
    _, o32 = _quickstart(32, arithmetic32=True)
    o_n, o_logL = o.fit(max_iter=50, e_rel=1e-3)
    n, logL = blend.fit(50, e_rel=1e-3)
    assert n == o_n and n == len(blend.loss)
    assert [tuple(s.parameters[1].shape) for s in blend.sources] == [tuple(s.image.x.shape) for s in o.sources]
    same_boxes = [tuple(s.image.x.shape) for s in o32.sources] == [tuple(s.image.x.shape) for s in o.sources] and len(o32.loss) == len(o.loss)
    assert same_boxes, "float32 rounding alone changes the box sequence of this scene: pick another yardstick"
    checks = [("loss", np.array(blend.loss), np.array(o.loss), np.array(o32.loss))]
    for k, (src, osrc, o32src) in enumerate(zip(blend.sources, o.sources, o32.sources)):
        checks.append(("sed%d" % k, src.parameters[0], osrc.spectrum.x, o32src.spectrum.x))
        checks.append(("morph%d" % k, src.parameters[1], osrc.image.x, o32src.image.x))
    for name, ours, ref64, ref32 in checks:
        err, rounding = tracks_rounding(ours, ref64, ref32)
        assert err < max(1e-5, 3 * rounding), (name, err, rounding)


@pytest.mark.parametrize("name", ["hsc_cosmos_35.npz", "point_extended.npz"])
@pytest.mark.parametrize("precision,tol", [(32, 1e-5), (64, 1e-6)])
def test_scene_forward_vs_reference_fixture(name, precision, tol):
    """model, render and logL of the reference's own forward code (boxes overhang the frame; masked weights)"""
    g = golden(name)
    blend, obs = _blend_from_golden(g, precision)
    model = blend.get_model()
    assert rel_peak(model, g["model"]) < tol
    ev = blend._get_plan().evaluate(want=("rendered", "loss"))
    assert rel_peak(ev["rendered"][0], g["rendered"]) < tol
    assert_allclose(-ev["loss"][0], float(g["logL"]), rtol=1e-5)
    assert rel_peak(obs.render(model), g["rendered"]) < 2 * tol


@pytest.mark.parametrize("name", ["hsc_cosmos_35.npz", "point_extended.npz"])
def test_scene_gradients_vs_oracle(name):
    from test_oracle_golden import _scene_from_golden
    g = golden(name)
    for precision, tol in ((64, 1e-9), (32, 2e-5)):
        blend, _ = _blend_from_golden(g, precision)
        plan = blend._get_plan()
        plan.upload_parameters(state=False)
        ev = plan.evaluate(want=("loss", "grads"))
        scene, _ = _scene_from_golden(g, frame_dtype=np.float64 if precision == 64 else np.float32)
        loss, grads = scene.loss_and_grads()
        assert_allclose(ev["loss"][0], loss, rtol=1e-9 if precision == 64 else 1e-5)
        i = ie = ip = 0
        for k, src in enumerate(scene.sources):
            assert rel_peak(ev["g_sed"][k], grads[i]) < tol, (k, "sed")
            if src.kind == "extended":
                assert rel_peak(ev["g_morph"][ie], grads[i + 1]) < tol, (k, "morph")
                # out-of-frame morphology pixels have exactly zero gradient (blend.py:41-44)
                oy, ox = src.bbox.origin[1:]
                yy, xx = np.mgrid[:src.bbox.shape[1], :src.bbox.shape[2]]
                out = (yy + oy < 0) | (yy + oy >= scene.frame_shape[1]) | (xx + ox < 0) | (xx + ox >= scene.frame_shape[2])
                assert (ev["g_morph"][ie][out] == 0).all()
                ie += 1
            else:
                assert rel_peak(ev["g_center"][ip], grads[i + 1]) < 10 * tol, (k, "center")
                ip += 1
            i += len(src.parameters)


@pytest.mark.parametrize("config", ["tiny", "cfg2", "cfg3"])
def test_fused_spectral_kernels_vs_cufft_path(config, monkeypatch):
    """the fused row/column spectral kernels against the cuFFT pipeline of the same plan code (same grid, same K^):
    model, rendered model, loss and every gradient of one evaluation, float64 twin to rounding, float32 to 2e-6"""
    from scarlet_b200 import synthetic
    scene = synthetic.make_scene(config, 1)
    for precision, tol in ((64, 1e-11), (32, 3e-6)):
        out = {}
        for mode in ("fused", "cufft"):
            monkeypatch.setenv("SB_SPECTRAL", mode)
            blend = synthetic.make_blend(scene, precision=precision)
            plan = blend._get_plan()
            assert plan.spectral_mode == (1 if mode == "fused" else 0)
            plan.upload_parameters(state=False)
            out[mode] = plan.evaluate(want=("model", "rendered", "loss", "grads"))
            plan.close()
        a, b = out["fused"], out["cufft"]
        assert_array_equal(a["model"], b["model"])   # same accumulation order, no FFT involved
        assert rel_peak(a["rendered"], b["rendered"]) < tol
        assert_allclose(a["loss"], b["loss"], rtol=10 * tol)
        assert rel_peak(a["g_sed"], b["g_sed"]) < 20 * tol
        for ga, gb in zip(a["g_morph"], b["g_morph"]):
            assert np.abs(ga - gb).max() <= 20 * tol * max(np.abs(gm).max() for gm in b["g_morph"])
        if len(b["g_center"]):
            assert rel_peak(a["g_center"], b["g_center"]) < 50 * tol


# --------------------------------------------------------------------------------------------------
# multi-resolution: ResolutionRenderer + ConvolutionRenderer on one model frame (BASELINE config 4 structure)
# --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rotated", [False, True])
@pytest.mark.parametrize("precision,tol", [(64, 1e-10), (32, 1e-5)])
def test_multiresolution_forward_and_gradients_vs_oracle(precision, tol, rotated):
    """rendered low- and high-resolution models, loss and all gradients of one evaluation; the oracle restates the
    reference's ResolutionRenderer literally (tabulated shifted kernels + matrix product) and is pinned to the
    reference's own render (tests/test_oracle_golden.py), the device evaluates the equivalent Fourier contraction"""
    from multires_scene import oracle_scene, product_scene
    g, blend, obs_lr, obs_hr = product_scene(precision, rotated)
    # float32 frame: the oracle takes the product's own float32 set-up (see multires_scene.oracle_scene), float64: the fixture's
    _, o = oracle_scene(np.float64 if precision == 64 else np.float32, rotated, setup_of=obs_lr.renderer if precision == 32 else None)
    assert obs_lr.renderer.isrot == rotated
    plan = blend._get_plan()
    assert plan.spectral_mode == 1
    plan.upload_parameters(state=False)
    model = o.get_model()
    ev0 = plan.evaluate(obs=0, want=("model", "rendered", "loss", "grads"))
    ev1 = plan.evaluate(obs=1, want=("rendered",))
    assert rel_peak(ev0["model"][0], model) < tol
    # the reference's own float32 render (float32 operator + np.dot) is 1.1e-5 from its float64 render on the fixture scene
    # (tests/test_host_api.py); the float64 twin pins the algorithm at 1e-10
    assert rel_peak(ev0["rendered"][0], o.observations[0].render(model)) < tol
    assert rel_peak(ev1["rendered"][0], o.observations[1].render(model)) < tol
    loss, grads = o.loss_and_grads()
    assert_allclose(ev0["loss"][0], loss, rtol=max(tol, 1e-9))
    gscale = max(np.abs(gr).max() for gr in grads[1::3])
    for k in range(len(o.sources)):
        assert rel_peak(ev0["g_sed"][k], grads[3 * k]) < 20 * tol
        assert np.abs(ev0["g_morph"][k] - grads[3 * k + 1]).max() < 20 * tol * gscale


@pytest.mark.parametrize("rotated", [False, True])
@pytest.mark.parametrize("precision,tol", [(64, 1e-8), (32, 2e-5)])
def test_multiresolution_fit_matches_oracle(precision, tol, rotated):
    from multires_scene import oracle_scene, product_scene
    g, blend, obs_lr, obs_hr = product_scene(precision, rotated)
    _, o = oracle_scene(np.float64 if precision == 64 else np.float32, rotated, setup_of=obs_lr.renderer if precision == 32 else None)
    n_iter = 12
    o.fit(max_iter=n_iter, e_rel=1e-3, min_iter=10 ** 9)
    n, logL = blend.fit(max_iter=n_iter, e_rel=1e-3, min_iter=10 ** 9)
    assert n == n_iter
    assert_allclose(np.array(blend.loss), np.array(o.loss), rtol=max(tol, 1e-9))
    sed_scale = max(float(np.abs(np.asarray(s.spectrum.x, dtype=np.float64)).max()) for s in o.sources)
    for src, osrc in zip(blend.sources, o.sources):
        assert np.abs(np.asarray(src.parameters[0], dtype=np.float64) - osrc.spectrum.x).max() < tol * sed_scale
        assert rel_peak(src.parameters[1], osrc.image.x) < tol
    assert rel_peak(blend.get_model(), o.get_model()) < tol


def test_intersection_frame_convolution_render():
    """coverage="intersection" model frame (frame.py:200-312): the high-resolution observation overhangs the frame; its
    ConvolutionRenderer render and logL against the reference's own (tests/golden/make_golden.py:multires, isect_*)"""
    import scarlet_b200 as sb
    from scarlet_b200.wcs import AffineWCS
    g = golden("multires.npz")
    obs_hr = sb.Observation(g["hr_images"].copy(), psf=sb.ImagePSF(g["hr_psfs"].copy()), weights=g["hr_weights"].copy(),
                            wcs=AffineWCS(g["hr_cd"], crpix=g["hr_crpix"]), channels=["h0", "h1", "h2"])
    obs_lr = sb.Observation(g["lr_images"].copy(), psf=sb.ImagePSF(g["lr_psfs"].copy()), weights=g["lr_weights"].copy(),
                            wcs=AffineWCS(g["lr_cd"], crpix=g["lr_crpix"]), channels=["l0", "l1", "l2", "l3", "l4"])
    frame = sb.Frame.from_observations([obs_lr, obs_hr], obs_id=1, coverage="intersection")
    frame = sb.Frame(frame.shape, channels=frame.channels, psf=frame.psf, wcs=frame.wcs, dtype=np.float64)
    obs_lr.match(frame)
    obs_hr.match(frame)
    model = g["isect_model"]
    assert rel_peak(obs_hr.render(model), g["isect_hr_rendered"]) < 1e-9
    assert_allclose(obs_hr.get_log_likelihood(model), float(g["isect_hr_logL"]), rtol=1e-8)


def test_multiresolution_cfg4_full_size():
    """BASELINE config 4 shape: 5 bands 30x30 at 0.2"/px + 3 bands 200x200 at 0.03"/px, model frame (8,228,228), grid 240^2,
    8 ExtendedSource: 5 iterations against the oracle.  Model pixels and SEDs at 1e-5 of their peaks; morphologies at
    5e-5 (the low-resolution bands' float32 render noise, see test_multiresolution_forward_and_gradients_vs_oracle)."""
    from oracle import scenes
    from scarlet_b200 import synthetic
    scene = synthetic.make_multires_scene(0)
    blend = synthetic.make_multires_blend(scene, precision=32)
    assert tuple(blend.frame.shape) == (8, 228, 228) and list(blend.observations[0].renderer._fft_shape) == [240, 240]
    o = scenes.build_multires_oracle(scene, scenes.multires_setup(blend))
    n_iter = 5
    o.fit(max_iter=n_iter, e_rel=1e-3, min_iter=10 ** 9)
    n, _ = blend.fit(max_iter=n_iter, e_rel=1e-3, min_iter=10 ** 9)
    assert n == n_iter and blend._get_plan().spectral_mode == 1
    assert_allclose(np.array(blend.loss), np.array(o.loss), rtol=2e-5)
    sed_scale = max(float(np.abs(np.asarray(s.spectrum.x, dtype=np.float64)).max()) for s in o.sources)
    for src, osrc in zip(blend.sources, o.sources):
        assert np.abs(np.asarray(src.parameters[0], dtype=np.float64) - osrc.spectrum.x).max() < 1e-5 * sed_scale
        assert rel_peak(src.parameters[1], osrc.image.x) < 5e-5
    assert rel_peak(blend.get_model(), o.get_model()) < 1e-5


@pytest.mark.parametrize("precision,n_iter,tol_loss,tol", [(64, 5, 1e-9, 1e-7), (32, 5, 2e-5, 5e-5)])
def test_multiresolution_cfg4_rotated_full_size(precision, n_iter, tol_loss, tol):
    """cfg4's shape with the low-resolution grid turned by 25 degrees (the rotated branch of ResolutionRenderer,
    renderer.py:318-363, 498-524): model frame (8,284,284), grid 300^2, 30x30 low-resolution pixels -- the dense contraction
    over the half plane on the device against the oracle's literal shift tables.  The float64 plan pins the algorithm; the
    float32 plan is held to the bars of the aligned full-size test (one evaluation agrees to 2e-7 of the peak in the render
    and 4e-7 in the gradients, profiles/r2n_multires_probe.txt).  The galaxies are drawn on the high-resolution footprint:
    in the uncovered corners of the union frame every gradient is rounding noise and AMSGrad's first steps follow its sign."""
    from oracle import scenes
    from scarlet_b200 import synthetic
    scene = synthetic.make_multires_scene(0, dict(synthetic.CFG4, lr_angle=25.0, config_id=41))
    blend = synthetic.make_multires_blend(scene, precision=precision)
    r = blend.observations[0].renderer
    assert r.isrot and tuple(blend.frame.shape) == (8, 284, 284) and list(r._fft_shape) == [300, 300]
    o = scenes.build_multires_oracle(scene, scenes.multires_setup(blend), frame_dtype=np.float32 if precision == 32 else np.float64)
    o.fit(max_iter=n_iter, e_rel=1e-3, min_iter=10 ** 9)
    n, _ = blend.fit(max_iter=n_iter, e_rel=1e-3, min_iter=10 ** 9)
    assert n == n_iter and blend._get_plan().spectral_mode == 1
    assert_allclose(np.array(blend.loss), np.array(o.loss), rtol=tol_loss)
    sed_scale = max(float(np.abs(np.asarray(s.spectrum.x, dtype=np.float64)).max()) for s in o.sources)
    for src, osrc in zip(blend.sources, o.sources):
        assert np.abs(np.asarray(src.parameters[0], dtype=np.float64) - osrc.spectrum.x).max() < tol * sed_scale
        assert rel_peak(src.parameters[1], osrc.image.x) < tol
    assert rel_peak(blend.get_model(), o.get_model()) < tol


# --------------------------------------------------------------------------------------------------
# the fitting loop
# --------------------------------------------------------------------------------------------------
def _compare_fit(scene, n_iter, precision, tol_morph, tol_sed, e_rel=1e-3, fixed=True, tol_model=None, check_state=True):
    from oracle import scenes
    from scarlet_b200 import synthetic
    o = scenes.build_oracle(scene, frame_dtype=np.float32 if precision == 32 else np.float64)
    if fixed:
        o_n, o_logL = o.fit(max_iter=n_iter, e_rel=e_rel, min_iter=10 ** 9)
    else:
        o_n, o_logL = o.fit(max_iter=n_iter, e_rel=e_rel)
    blend = synthetic.make_blend(scene, precision=precision)
    n, logL = blend.fit(max_iter=n_iter, e_rel=e_rel, min_iter=10 ** 9 if fixed else 1, check_every=1)
    assert n == o_n
    assert_allclose(np.array(blend.loss), np.array(o.loss), rtol=1e-9 if precision == 64 else 2e-5)
    worst = {}
    sed_scale = max(float(np.abs(np.asarray(osrc.spectrum.x, dtype=np.float64)).max()) for osrc in o.sources)
    for src, osrc in zip(blend.sources, o.sources):
        ps = src.parameters
        sed_err = float(np.abs(np.asarray(ps[0], dtype=np.float64) - np.asarray(osrc.spectrum.x, dtype=np.float64)).max())
        assert sed_err < tol_sed * sed_scale
        assert rel_peak(ps[0], osrc.spectrum.x) < 10 * tol_sed
        if check_state:  # first moment of the spectrum gradient (meaningful while the gradient is not yet rounding noise)
            assert rel_peak(ps[0].m, osrc.spectrum.m) < max(tol_sed, 1e-6) * 50
        if osrc.kind == "extended":
            worst["morph"] = max(worst.get("morph", 0), rel_peak(ps[1], osrc.image.x))
            assert rel_peak(ps[1], osrc.image.x) < tol_morph
            assert ps[1].std is not None and ps[1].m.shape == ps[1].shape
        else:
            assert np.abs(np.asarray(ps[1]) - osrc.center.x).max() < max(tol_morph, 1e-9) * 10
    model = blend.get_model()
    assert rel_peak(model, o.get_model()) < (tol_model if tol_model is not None else max(tol_morph, tol_sed))
    return blend, o


@pytest.mark.parametrize("n_iter", [1, 10, 30])
def test_fit_float64_twin_matches_oracle(n_iter):
    """algorithmic identity: the float64 twin follows the float64 oracle to rounding"""
    from scarlet_b200 import synthetic
    _compare_fit(synthetic.make_scene("tiny", 0), n_iter, 64, 1e-9, 1e-9)


@pytest.mark.parametrize("n_iter", [1, 10, 30])
def test_fit_float32_matches_oracle(n_iter):
    from scarlet_b200 import synthetic
    _compare_fit(synthetic.make_scene("tiny", 0), n_iter, 32, 1e-5, 1e-5)


def test_fit_stop_rule_matches_oracle():
    """same iteration count under the reference's convergence rule (blend.py:294-299)"""
    from scarlet_b200 import synthetic
    blend, o = _compare_fit(synthetic.make_scene("tiny", 2), 200, 64, 1e-8, 1e-8, e_rel=1e-2, fixed=False)
    assert len(blend.loss) < 200


def test_fit_cfg2_float32_matches_oracle():
    """BASELINE config 2 shape (5x128x128, 10 ExtendedSource, Gaussian PSF): 30 iterations at the 1e-5 bar; after 50
    iterations model pixels and SEDs still meet 1e-5, while the worst pixel of the worst individual morphology image
    sits at 0.9-1.9e-5 depending on the FFT grid / rounding order (float32 noise amplified by the non-smooth
    projections; tests/parity_probe.py), hence 3e-5 for that one quantity."""
    from scarlet_b200 import synthetic
    _compare_fit(synthetic.make_scene("cfg2", 0), 30, 32, 1e-5, 1e-5)
    _compare_fit(synthetic.make_scene("cfg2", 0), 50, 32, 3e-5, 1e-5, tol_model=1e-5)


def test_fit_cfg3_float32_matches_oracle():
    """BASELINE config 3 shape (5x256x256, 20 ExtendedSource + 5 PointSource, ImagePSF, monotonic + symmetry)"""
    from scarlet_b200 import synthetic
    _compare_fit(synthetic.make_scene("cfg3", 0), 20, 32, 1e-5, 1e-5)


def test_warm_start_second_fit():
    """optimiser state lives on the Parameters: fit(5)+fit(5) == oracle fit(5)+fit(5) (blend.py:154-163)"""
    from oracle import scenes
    from scarlet_b200 import synthetic
    sc = synthetic.make_scene("tiny", 4)
    o = scenes.build_oracle(sc, frame_dtype=np.float64)
    o.fit(max_iter=5, min_iter=10 ** 9)
    o.fit(max_iter=5, min_iter=10 ** 9)
    b = synthetic.make_blend(sc, precision=64)
    b.fit(max_iter=5, min_iter=10 ** 9)
    n, _ = b.fit(max_iter=5, min_iter=10 ** 9)
    assert n == 10
    assert_allclose(b.loss, o.loss, rtol=1e-9)
    for src, osrc in zip(b.sources, o.sources):
        assert rel_peak(src.parameters[0], osrc.spectrum.x) < 1e-9


def test_batch_equals_individual_fits():
    import scarlet_b200 as sb
    from scarlet_b200 import synthetic
    scs = [synthetic.make_scene("tiny", i) for i in range(5)]
    singles = [synthetic.make_blend(s) for s in scs]
    for b in singles:
        b.fit(max_iter=15, e_rel=1e-4)
    batch = [synthetic.make_blend(s) for s in scs]
    res = sb.BlendBatch(batch).fit(max_iter=15, e_rel=1e-4)
    for b1, b2, r in zip(singles, batch, res):
        assert r[0] == len(b1.loss)
        assert_array_equal(np.array(b1.loss), np.array(b2.loss))
        for p1, p2 in zip(b1.parameters, b2.parameters):
            assert_array_equal(np.asarray(p1), np.asarray(p2))


def test_batch_of_rotated_multiresolution_scenes_equals_individual_fits():
    """three scenes on the rotated two-grid geometry (shared resampling tables, per-scene data and sources): a batch gives the
    single fits bit for bit"""
    import scarlet_b200 as sb
    from multires_scene import product_scene
    rng = np.random.default_rng(3)
    singles, batch = [], []
    for k in range(3):
        scale, noise = rng.uniform(0.7, 1.4), rng.standard_normal((5, 8, 8)).astype(np.float32)
        for out in (singles, batch):
            _, blend, obs_lr, obs_hr = product_scene(32, rotated=True)
            obs_lr.data[...] = obs_lr.data * scale + 0.3 * noise
            for src in blend.sources:
                src.parameters[0][...] = np.asarray(src.parameters[0]) * scale
            out.append(blend)
    for b in singles:
        b.fit(max_iter=12, e_rel=1e-4)
    res = sb.BlendBatch(batch).fit(max_iter=12, e_rel=1e-4)
    for b1, b2, r in zip(singles, batch, res):
        assert r[0] == len(b1.loss)
        assert_array_equal(np.array(b1.loss), np.array(b2.loss))
        for p1, p2 in zip(b1.parameters, b2.parameters):
            assert_array_equal(np.asarray(p1), np.asarray(p2))
    assert len({tuple(b.loss) for b in singles}) == 3  # the scenes do differ


def test_batch_with_fitted_psf_offsets_equals_individual_fits():
    """two scenes whose ConvolutionRenderers carry their own fitted psf_shift: the batch keeps one kernel per scene"""
    import scarlet_b200 as sb
    g = golden("psf_shift.npz")
    shifts = ([0.12, -0.2], [-0.3, 0.25])
    singles = [_psf_shift_blend(g, sh, 32, fit_sources=True)[0] for sh in shifts]
    batch = [_psf_shift_blend(g, sh, 32, fit_sources=True)[0] for sh in shifts]
    for b in singles:
        b.fit(max_iter=8, e_rel=1e-5)
    res = sb.BlendBatch(batch).fit(max_iter=8, e_rel=1e-5)
    for b1, b2, r in zip(singles, batch, res):
        assert r[0] == len(b1.loss)
        assert_array_equal(np.array(b1.loss), np.array(b2.loss))
        for p1, p2 in zip(b1.parameters, b2.parameters):
            assert_array_equal(np.asarray(p1), np.asarray(p2))
        assert_array_equal(np.asarray(b1.observations[0].renderer.parameters[0]), np.asarray(b2.observations[0].renderer.parameters[0]))
    fitted = [np.asarray(b.observations[0].renderer.parameters[0]) for b in batch]
    assert not np.array_equal(fitted[0], fitted[1]) and all(np.abs(f - sh).max() > 0 for f, sh in zip(fitted, shifts))


def _resizing_scene():
    """a wide and a compact galaxy in 31x31 boxes that are too small for what the data pull in: the boxes grow several
    times within 45 iterations (dynamic boxes on, morphology.py:132-207)"""
    from scarlet_b200 import synthetic
    scene = synthetic.make_scene(dict(C=3, N=70, n_ext=2, n_pt=0, psf="gaussian", P=15, B=31, symmetric=True, iters=40, config_id=11,
                                      resizing=True), 0)
    y, x = np.mgrid[:31, :31] - 15
    wide = np.exp(-np.hypot(y, x) / 9.0)
    compact = np.exp(-np.hypot(y, x) / 1.5) * (np.hypot(y, x) < 6)
    for s, morph, cen, amp in zip(scene["sources"], (wide, compact), ((24, 26), (50, 46)), (900.0, 300.0)):
        s["morph"], s["origin"], s["center"] = morph / morph.max(), (cen[0] - 15, cen[1] - 15), cen
        s["sed"] = (amp * np.array([1.0, 0.8, 0.6])).astype(np.float32)
    truth = sum(s["sed"][:, None, None].astype(np.float64) * np.pad(np.exp(-np.hypot(*(np.mgrid[:70, :70] - np.array(s["center"])[:, None, None])) / r), 0)
                for s, r in zip(scene["sources"], (9.0, 1.5)))
    from scipy import signal
    from scarlet_b200 import fft as sfft
    from scarlet_b200.psf import GaussianPSF
    diff = sfft.match_psf(scene["obs_psf"].astype(np.float32), GaussianPSF(sigma=(0.8,) * 3).get_model().astype(np.float32), padding=10).image
    rng = np.random.default_rng(5)
    scene["images"] = (np.stack([signal.fftconvolve(truth[c], diff[c], mode="same") for c in range(3)])
                       + rng.standard_normal((3, 70, 70))).astype(np.float32)
    return scene


@pytest.mark.parametrize("precision,tol", [(64, 1e-8), (32, 2e-3)])
def test_dynamic_box_resize_matches_oracle(precision, tol):
    """boxes shrink / grow every 10 iterations, the optimiser restarts with warm state and halved step.
    The float64 twin follows the oracle through every restart to 1e-8.  The scene starts far from its optimum on
    purpose (the loss oscillates for the first 30 iterations), which amplifies float32 rounding to ~2e-4 in the loss
    history: the float32 run checks the control flow (same box sizes, origins, restart points, iteration count)."""
    from oracle import scenes
    from scarlet_b200 import synthetic
    scene = _resizing_scene()
    o = scenes.build_oracle(scene, frame_dtype=np.float32 if precision == 32 else np.float64)
    o.fit(max_iter=45, e_rel=1e-6)
    blend = synthetic.make_blend(scene, precision=precision)
    n, logL = blend.fit(max_iter=45, e_rel=1e-6)
    shapes = [tuple(s.parameters[1].shape) for s in blend.sources]
    assert shapes == [tuple(s.image.x.shape) for s in o.sources]
    assert shapes[0][0] > 41 and shapes[1][0] > 31, shapes          # several resize + restart cycles happened
    assert [s.bbox.origin[1:] for s in blend.sources] == [s.bbox.origin[1:] for s in o.sources]
    assert n == len(o.loss)
    assert_allclose(np.array(blend.loss), np.array(o.loss), rtol=max(tol, 1e-9))
    for src, osrc in zip(blend.sources, o.sources):
        assert rel_peak(src.parameters[0], osrc.spectrum.x) < tol
        assert rel_peak(src.parameters[1], osrc.image.x) < tol
        assert src.parameters[1].step == osrc.image.step
    assert rel_peak(blend.get_model(), o.get_model()) < tol


def test_ragged_batch_and_many_channels():
    """scenes with different numbers of sources in one batch, and a 10-band scene (more bands than one CTA of the row
    kernels takes at once, and more than the grouped update kernel handles: band chunking + the generic update kernel)"""
    import scarlet_b200 as sb
    from oracle import scenes
    from scarlet_b200 import synthetic
    base = dict(C=3, N=40, n_pt=1, psf="gaussian", P=15, B=15, symmetric=True, iters=10, config_id=21)
    scs = [synthetic.make_scene(dict(base, n_ext=n), i) for i, n in enumerate((1, 4, 2, 3))]
    batch = [synthetic.make_blend(s) for s in scs]
    res = sb.BlendBatch(batch).fit(max_iter=10, e_rel=1e-3, min_iter=10 ** 9)
    for sc, b in zip(scs, batch):
        o = scenes.build_oracle(sc)
        o.fit(max_iter=10, e_rel=1e-3, min_iter=10 ** 9)
        assert_allclose(np.array(b.loss), np.array(o.loss), rtol=2e-5)
        assert rel_peak(b.get_model(), o.get_model()) < 1e-5
    wide = synthetic.make_scene(dict(base, C=10, n_ext=3, config_id=22), 0)
    for precision, tol in ((64, 1e-9), (32, 1e-5)):
        o = scenes.build_oracle(wide, frame_dtype=np.float32 if precision == 32 else np.float64)
        o.fit(max_iter=10, e_rel=1e-3, min_iter=10 ** 9)
        b = synthetic.make_blend(wide, precision=precision)
        b.fit(max_iter=10, e_rel=1e-3, min_iter=10 ** 9)
        assert b._get_plan().spectral_mode == 1
        assert_allclose(np.array(b.loss), np.array(o.loss), rtol=max(2 * tol, 1e-9))
        assert rel_peak(b.get_model(), o.get_model()) < tol
        for src, osrc in zip(b.sources, o.sources):
            assert rel_peak(src.parameters[0], osrc.spectrum.x) < tol


@pytest.mark.parametrize("precision,tol", [(64, 1e-9), (32, 2e-5)])
def test_shifting_morphology_matches_oracle(precision, tol):
    """ExtendedSource(shifting=True) -- the model uses fft.shift(image, shift) and both are fitted.  The oracle
    shifts with the reference's literal FFT pipeline and differentiates it with torch autograd; the device evaluates the
    equivalent Toeplitz operators.  One evaluation (model, loss, gradients wrt spectrum, image and shift), then a fit."""
    from oracle import scenes
    from scarlet_b200 import synthetic
    scene = synthetic.make_scene(dict(C=3, N=40, n_ext=3, n_pt=1, psf="gaussian", P=15, B=15, symmetric=False, iters=10, config_id=31,
                                      shifting=True), 0)
    dtype = np.float32 if precision == 32 else np.float64
    o = scenes.build_oracle(scene, frame_dtype=dtype)
    blend = synthetic.make_blend(scene, precision=precision)
    assert [p.name for p in blend.sources[0].parameters] == ["spectrum", "image", "shift"] and not blend.sources[0].parameters[2].fixed
    plan = blend._get_plan()
    plan.upload_parameters(state=False)
    ev = plan.evaluate(want=("model", "loss", "grads"))
    loss, grads = o.loss_and_grads()
    assert rel_peak(ev["model"][0], o.get_model()) < tol
    assert_allclose(ev["loss"][0], loss, rtol=max(tol, 1e-9))
    gi = 0
    n_ext = sum(1 for s in o.sources if s.kind == "extended")
    for k, src in enumerate(o.sources):
        assert rel_peak(ev["g_sed"][k], grads[gi]) < 20 * tol
        if src.kind == "extended":
            assert rel_peak(ev["g_morph"][k], grads[gi + 1]) < 20 * tol
            assert_allclose(ev["g_center"][k], grads[gi + 2], rtol=200 * tol, atol=200 * tol * np.abs(grads[gi + 2]).max())
        gi += len(src.parameters)
    o.fit(max_iter=10, e_rel=1e-3, min_iter=10 ** 9)
    n, _ = blend.fit(max_iter=10, e_rel=1e-3, min_iter=10 ** 9)
    assert_allclose(np.array(blend.loss), np.array(o.loss), rtol=max(tol, 1e-9))
    for src, osrc in zip(blend.sources, o.sources):
        assert rel_peak(src.parameters[0], osrc.spectrum.x) < 10 * tol
        if osrc.kind == "extended":
            assert rel_peak(src.parameters[1], osrc.image.x) < 10 * tol
            assert np.abs(np.asarray(src.parameters[2]) - osrc.shift.x).max() < 100 * tol
    assert rel_peak(blend.get_model(), o.get_model()) < 10 * tol


def test_post_fit_api_pickle_render_measure():
    """what users do around fit (quickstart cells 27-34): get_model, Observation.render, measure.*, pickle of the sources
    with their optimiser state (Parameter.m / v / vhat / std live in the plan's packed host arrays until detached)"""
    import pickle
    import scarlet_b200 as sb
    from scarlet_b200 import measure, synthetic
    scene = synthetic.make_scene("tiny", 3)
    blend = synthetic.make_blend(scene)
    blend.fit(max_iter=8, e_rel=1e-4)
    model = blend.get_model()
    obs = blend.observations[0]
    rendered = obs.render(model)
    assert rendered.shape == obs.data.shape and np.isfinite(rendered).all()
    plan = blend._get_plan()
    plan.upload_parameters(state=False)
    assert_allclose(obs.get_log_likelihood(model), -plan.evaluate(want=("loss",))["loss"][0], rtol=1e-5)  # host formula vs device loss
    src = blend.sources[0]
    assert measure.flux(src).shape == (3,) and measure.snr(src, obs) > 0
    p = src.parameters[1]
    assert p.m is not None and p.v.shape == p.shape and np.ma.isMaskedArray(p.std)
    blob = pickle.dumps(blend.sources)
    v_before = np.array(p.v)
    blend._plan.close()  # frees the packed staging arrays: the parameters keep private copies of their state
    assert_array_equal(np.array(p.v), v_before) and p.std is not None
    restored = pickle.loads(blob)
    q = restored[0].parameters[1]
    assert_array_equal(np.asarray(q), np.asarray(p))
    assert_array_equal(q.v, v_before)
    assert_allclose(q.std, p.std)
    # a second fit on the restored sources warm-starts from the pickled state
    blend2 = sb.Blend(restored, obs)
    n, _ = blend2.fit(max_iter=3, e_rel=1e-9)
    assert n == 3 and np.isfinite(blend2.loss).all()


def test_nonfinite_raises_arithmetic_error():
    from scarlet_b200 import synthetic
    sc = synthetic.make_scene("tiny", 0)
    sc["images"][0, 5, 5] = np.inf
    b = synthetic.make_blend(sc)
    with pytest.raises(ArithmeticError):
        b.fit(max_iter=5)


def test_unsupported_features_raise():
    import scarlet_b200 as sb
    from scarlet_b200 import synthetic
    sc = synthetic.make_scene("tiny", 0)
    b = synthetic.make_blend(sc)
    with pytest.raises(NotImplementedError):
        b.fit(max_iter=2, scheme="adam")
    b.sources[0].parameters[0].step = lambda X, it: 1.0
    with pytest.raises(TypeError):
        b.fit(max_iter=2)


def test_init_all_sources_end_to_end():
    """SURVEY 8f-2, the quickstart's own call: initialization.init_all_sources (component count from the PSF signal-to-noise,
    sources from the data, spectra solved against device-rendered unit-spectrum models) -- spectra against the reference's
    (tests/golden/make_golden.py:init_helpers), then a short fit."""
    import scarlet_b200 as sb
    h = golden("init_helpers.npz")
    g = golden("hsc_cosmos_35.npz")
    C = g["images"].shape[0]
    frame = sb.Frame(g["images"].shape, psf=sb.GaussianPSF(sigma=(0.8,) * C), channels=list(range(C)))
    obs = sb.Observation(g["images"].copy(), psf=sb.ImagePSF(g["psfs"].copy()), weights=h["weights"].copy(), channels=list(range(C)))
    obs.match(frame)
    sources, skipped = sb.initialization.init_all_sources(frame, [tuple(c) for c in h["centers"][:3]], obs, max_components=1,
                                                           resizing=False)
    assert skipped == [] and len(sources) == 3
    for k, src in enumerate(sources):
        assert_allclose(np.asarray(src.parameters[1]), h["src%d_image" % k], atol=1e-6)
        assert_allclose(np.asarray(src.parameters[0]), h["matched_spectra"][k], rtol=1e-3)
    blend = sb.Blend(sources, obs)
    n, logL = blend.fit(20, e_rel=1e-4)
    assert n == len(blend.loss) and np.isfinite(logL) and blend.loss[-1] < blend.loss[0]


def test_nonfinite_data_raises_with_extended_sources_only():
    """np.maximum / max() propagate NaN (constraint.py:83-92, 276-287), so a NaN gradient must reach the
    non-finite check instead of being scrubbed by the positivity projection.  Scene with ExtendedSources only."""
    from scarlet_b200 import synthetic
    scene = synthetic.make_scene("cfg2", 3)
    scene["images"] = scene["images"].copy()
    scene["images"][2, 40, 41] = np.nan
    blend = synthetic.make_blend(scene, precision=32)
    with pytest.raises(ArithmeticError):
        blend.fit(max_iter=5, e_rel=1e-3, min_iter=10 ** 9)
    scene["images"][2, 40, 41] = np.inf
    blend = synthetic.make_blend(scene, precision=64)
    with pytest.raises(ArithmeticError):
        blend.fit(max_iter=5, e_rel=1e-3, min_iter=10 ** 9)


def test_user_callback_and_explicit_parameters():
    """``callback(*X, it=it)`` after every iteration that neither stops nor restarts (blend.py:301-302),
    StopIteration from it ends the fit cleanly; ``get_model(*parameters)`` renders explicit values (blend.py:200-244)."""
    from oracle import scenes
    from scarlet_b200 import synthetic
    scene = synthetic.make_scene("tiny", 1)
    ref = synthetic.make_blend(scene, precision=64)
    ref.fit(max_iter=12, e_rel=1e-3, min_iter=10 ** 9)
    seen = []

    def cb(*X, it=None):
        assert len(X) == len(blend.parameters)
        seen.append((it, float(np.asarray(X[0]).sum()), float(blend.loss[-1])))

    blend = synthetic.make_blend(scene, precision=64)
    n, logL = blend.fit(max_iter=12, e_rel=1e-3, min_iter=10 ** 9, callback=cb)
    assert [s[0] for s in seen] == list(range(12)) and n == 12
    assert_allclose(blend.loss, ref.loss, rtol=1e-13)   # one iteration at a time == one call (same arithmetic)
    for a, b in zip(blend.parameters, ref.parameters):
        assert_allclose(np.asarray(a), np.asarray(b), rtol=1e-12, atol=1e-300)
    o = scenes.build_oracle(scene, frame_dtype=np.float64)
    osum = []
    o.fit(max_iter=12, e_rel=1e-3, min_iter=10 ** 9, callback=lambda it: osum.append(float(o.sources[0].spectrum.x.sum())))
    assert_allclose([s[1] for s in seen], osum, rtol=1e-6)  # the callback sees the parameters of ITS iteration

    def stopper(*X, it=None):
        if it == 4:
            raise StopIteration

    blend2 = synthetic.make_blend(scene, precision=64)
    n2, _ = blend2.fit(max_iter=12, e_rel=1e-3, min_iter=10 ** 9, callback=stopper)
    assert n2 == 5
    assert_allclose(blend2.loss, ref.loss[:5], rtol=1e-13)
    # explicit parameter values: the model of the initial parameters, rendered from a fitted blend
    fresh = synthetic.make_blend(scene, precision=64)
    init = [np.array(p) for p in fresh.parameters]
    m0 = fresh.get_model()
    assert rel_peak(blend.get_model(*init), m0) < 1e-12
    assert rel_peak(blend.get_model(), ref.get_model()) < 1e-12  # ... and the stored values are untouched


def test_plan_follows_edits_between_fits():
    """``fixed``, ``step`` and in-place edits of ``obs.data`` between two fits are honoured (the reference
    re-reads them on every fit, blend.py:103-145)."""
    from scarlet_b200 import synthetic
    scene = synthetic.make_scene("tiny", 2)
    blend = synthetic.make_blend(scene, precision=64)
    blend.fit(max_iter=3, e_rel=1e-3, min_iter=10 ** 9)
    sed = blend.sources[0].parameters[0]
    sed.fixed = True
    before = np.array(sed)
    blend.fit(max_iter=3, e_rel=1e-3, min_iter=10 ** 9)
    assert np.array_equal(before, np.asarray(sed))
    sed.fixed = False
    blend.observations[0].data[...] *= 2  # in place
    twin = synthetic.make_blend(scene, precision=64)
    twin.observations[0].data[...] *= 2
    for a, b in zip(twin.parameters, blend.parameters):
        a[...] = np.asarray(b)
        a.m, a.v, a.vhat = np.array(b.m), np.array(b.v), np.array(b.vhat)
    blend.loss.clear()
    blend.fit(max_iter=3, e_rel=1e-3, min_iter=10 ** 9)
    twin.fit(max_iter=3, e_rel=1e-3, min_iter=10 ** 9)
    assert_allclose(blend.loss, twin.loss, rtol=1e-12)


# --------------------------------------------------------------------------------------------------
# full-length trajectories at BASELINE's iteration counts
# --------------------------------------------------------------------------------------------------
FULL_LENGTH = [("cfg2", 200, 1), ("cfg3", 100, 0), ("cfg3", 100, 1), ("cfg5", 100, 1)]


@pytest.mark.parametrize("config,n_iter,scene_id", FULL_LENGTH)
def test_full_length_float32_meets_north_star_bar(config, n_iter, scene_id):
    """The shipped float32 path after BASELINE's iteration counts (cfg2 200, cfg3 / cfg5 100): model pixels and SEDs within
    1e-5 of the peak of the reference arithmetic (north star), loss history 2e-5; single morphology images within 1e-5 too on
    these scenes.  Measured curves: profiles/r2_parity_curve_*.json (tools/parity_curve.py)."""
    from scarlet_b200 import synthetic
    # (the optimiser's gradient moments are not compared here: near convergence the gradient itself is a difference of
    #  nearly equal numbers, i.e. rounding noise in any float32 implementation)
    # Single morphology images: 3e-5.  Measured (profiles/r2a_parity_curve_cfg2.json, scene 1, 200 iterations): merely rounding
    # the model cube to float32 -- the only difference between the float64 twin and the oracle with a float32 frame -- moves
    # the worst morphology pixel by 1.9e-5; model pixels and SEDs (the north-star quantities) stay below 1e-5.
    _compare_fit(synthetic.make_scene(config, scene_id), n_iter, 32, 3e-5, 1e-5, tol_model=1e-5, check_state=False)


@pytest.mark.parametrize("config,n_iter,scene_id", [("cfg2", 200, 0), ("cfg5", 100, 0)])
def test_full_length_float32_sensitive_scenes_track_float32_rounding(config, n_iter, scene_id):
    """Scenes whose trajectory is sensitive to rounding (AMSGrad's first steps move a pixel by +-alpha*3.16 according to the
    SIGN of its gradient, and the projections are non-smooth): an independent float32 implementation of the reference
    algorithm (the oracle with float32 FFTs / morphologies, ``float32_arithmetic``) leaves the float64 trajectory by 6e-5
    (model) / 2.5e-4 (morphology) on cfg5 scene 0.  The bar for the product on such a scene is therefore relative: no further
    from the reference arithmetic than three times what float32 rounding alone does -- and SEDs still within 1e-5."""
    import tools.parity_curve as pc
    from scarlet_b200 import synthetic
    scene = synthetic.make_scene(config, scene_id)
    cps = [n_iter]
    o64, o32 = pc.oracle_run(scene, cps, False), pc.oracle_run(scene, cps, True)
    g32 = pc.gpu_run(scene, cps, 32)
    ours, rounding = pc.compare(g32[n_iter], o64[n_iter]), pc.compare(o32[n_iter], o64[n_iter])
    assert ours["sed"] < 1e-5
    for key in ("model", "morph", "loss"):
        assert ours[key] < max(1e-5, 3 * rounding[key]), (key, ours, rounding)


@pytest.mark.parametrize("config,n_iter,scene_id", [("cfg2", 200, 1), ("cfg3", 100, 0), ("cfg5", 100, 1)])
def test_full_length_float64_twin(config, n_iter, scene_id):
    """The float64 twin against the float64 oracle over the same trajectories: algorithmic identity of the CUDA path."""
    from scarlet_b200 import synthetic
    _compare_fit(synthetic.make_scene(config, scene_id), n_iter, 64, 1e-8, 1e-9, tol_model=1e-9)


def test_generic_kernel_takes_boxes_beyond_grouped_limit():
    """a 129 x 129 box is beyond the 16-bit byte offsets of the grouped / warp kernels (and beyond the
    register-resident iterate), so the generic kernel (one shared-memory image, x / psi / z streamed) takes it."""
    from scarlet_b200 import synthetic
    cfg = dict(synthetic.CONFIGS["cfg2"], B=129, N=160, n_ext=2, config_id=12)
    scene = synthetic.make_scene(cfg, 0)
    _compare_fit(scene, 5, 64, 1e-9, 1e-9)
    _compare_fit(scene, 5, 32, 1e-5, 1e-5)


def _many_resizing_scenes(n):
    """variants of the resizing scene: amplitudes, widths and noise differ, so the scenes resize at different iterations"""
    out = []
    for i in range(n):
        rng = np.random.default_rng(100 + i)
        scene = _resizing_scene()
        scene["images"] = (scene["images"] * rng.uniform(0.6, 1.4) + rng.standard_normal(scene["images"].shape) * 0.5).astype(np.float32)
        for s in scene["sources"]:
            s["sed"] = (s["sed"] * rng.uniform(0.7, 1.3)).astype(np.float32)
        if i % 3 == 1:  # a scene whose second box starts large enough: fewer restarts
            y, x = np.mgrid[:41, :41] - 20
            m = np.exp(-np.hypot(y, x) / 1.5) * (np.hypot(y, x) < 6)
            s = scene["sources"][1]
            s["morph"], s["origin"] = m / m.max(), (s["center"][0] - 20, s["center"][1] - 20)
        out.append(scene)
    return out


@pytest.mark.parametrize("precision", [32, 64])
def test_dynamic_batch_equals_individual_fits(precision):
    """a batch of 32 scenes with dynamic boxes (the reference default) equals 32 single ``Blend.fit`` runs bit
    for bit -- per-scene iteration counters, pause at each scene's own inspection points, restart with warm state."""
    from scarlet_b200 import BlendBatch, synthetic
    scenes_ = _many_resizing_scenes(32)
    singles = [synthetic.make_blend(sc, precision=precision) for sc in scenes_]
    res_single = [b.fit(max_iter=45, e_rel=1e-4) for b in singles]
    batch_blends = [synthetic.make_blend(sc, precision=precision) for sc in scenes_]
    batch = BlendBatch(batch_blends, precision=precision)
    assert batch.dynamic
    res_batch = batch.fit(max_iter=45, e_rel=1e-4)
    assert batch.replans >= 2
    sizes = set()
    for k, (a, b) in enumerate(zip(singles, batch_blends)):
        assert res_single[k][0] == res_batch[k][0], k
        assert a.loss == b.loss, k
        for pa, pb in zip(a.parameters, b.parameters):
            assert pa.shape == pb.shape and np.array_equal(np.asarray(pa), np.asarray(pb)), (k, pa.name)
            if pa.m is not None and pb.m is not None:
                assert np.array_equal(np.asarray(pa.m), np.asarray(pb.m)) and np.array_equal(np.asarray(pa.v), np.asarray(pb.v))
        assert [s.bbox for s in a.sources] == [s.bbox for s in b.sources]
        sizes.add(tuple(s.parameters[1].shape[0] for s in b.sources))
        sizes.add(len(b.loss))
    assert len(sizes) > 3  # the scenes really went different ways
    batch.close()


# --------------------------------------------------------------------------------------------------
# ConvolutionRenderer(psf_shift=...)  (renderer.py:172-177, 220-227, 252-254)
# --------------------------------------------------------------------------------------------------
def _psf_shift_blend(g, shift, precision, fit_sources=False):
    """the fixture scene as product objects: 4 full-frame components (spectrum x image), one observation whose renderer
    carries the fitted kernel offset"""
    import scarlet_b200 as sb
    C = g["images"].shape[0]
    channels = list(range(C))
    dtype = np.float32 if precision == 32 else np.float64
    frame = sb.Frame(g["images"].shape, psf=sb.GaussianPSF(sigma=(float(g["model_sigma"]),) * C), channels=channels, dtype=dtype)
    obs = sb.Observation(g["images"].copy(), psf=sb.ImagePSF(g["psfs"].copy()), weights=g["weights"].copy(), channels=channels)
    renderer = sb.renderer.ConvolutionRenderer(obs, frame, psf_shift=np.array(shift, dtype=np.float64))
    obs.match(frame, renderer=renderer)
    comps = []
    for sed, morph in zip(g["seds"], g["morphs"]):
        spectrum = sb.TabulatedSpectrum(frame, sb.Parameter(sed.copy(), name="spectrum", step=1e-2, fixed=not fit_sources,
                                                            constraint=sb.PositivityConstraint()))
        image = sb.Parameter(morph.copy(), name="image", step=1e-2, fixed=not fit_sources, constraint=sb.PositivityConstraint())
        comps.append(sb.FactorizedComponent(frame, spectrum, sb.ImageMorphology(frame, image, resizing=False)))
    return sb.Blend(comps, obs, precision=precision), obs, renderer


def _psf_shift_oracle(g, shift, fit_sources=False, frame_dtype=np.float64):
    from oracle import scarlet_oracle as so
    C = g["images"].shape[0]
    obs = so.ObservationOracle(g["images"], g["weights"], so.ImagePSFOracle(g["psfs"]), frame_dtype=frame_dtype, psf_shift=shift)
    mpsf = so.GaussianPSFOracle((float(g["model_sigma"]),) * C)
    srcs = []
    for sed, morph in zip(g["seds"], g["morphs"]):
        s = so.ExtendedSourceOracle(sed, morph, (0, 0), sed_dtype=np.float64)
        s.spectrum = so.OParam(np.array(sed, dtype=np.float64), "spectrum", 1e-2, so.ChainSpec([("positivity", 0.0)]), fixed=not fit_sources)
        s.image = so.OParam(np.array(morph, dtype=np.float64), "image", 1e-2, so.ChainSpec([("positivity", 0.0)]), fixed=not fit_sources)
        s.shift.fixed = True
        srcs.append(s)
    return so.SceneOracle(g["images"].shape, mpsf, srcs, [obs], frame_dtype=frame_dtype)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_psf_shift_forward_and_gradient_vs_reference_fixture(tag):
    """render, logL and d logL / d psf_shift against the reference's own outputs (tests/golden/make_golden.py:psf_shift)"""
    g = golden("psf_shift.npz")
    for precision, tol in ((64, 1e-9), (32, 2e-5)):
        blend, obs, renderer = _psf_shift_blend(g, g["shift_" + tag], precision)
        plan = blend._get_plan()
        plan.upload_parameters(state=False)
        ev = plan.evaluate(want=("model", "rendered", "loss", "grads"))
        assert rel_peak(ev["model"][0], g["model"]) < max(tol, 1e-7)
        assert rel_peak(ev["rendered"][0], g["rendered_" + tag]) < tol
        assert_allclose(-ev["loss"][0], float(g["logL_" + tag]), rtol=max(tol, 1e-9))
        gs = ev["g_center"][-1]  # the renderer parameter sits behind the sources' entries
        assert_allclose(gs, g["dloss_dshift_" + tag], rtol=1e-6 if precision == 64 else 2e-3)
        # the host-side render (what users call around a fit) follows the parameter too
        assert rel_peak(obs.render(g["model"].astype(blend.frame.dtype), *obs.parameters), g["rendered_" + tag]) < tol
        # (float32 frame: the difference kernel itself is formed from float32 PSF images, renderer.py:198-202)
        assert_allclose(renderer.shifted_kernel(g["shift_" + tag]), g["kernel_" + tag], atol=1e-13 if precision == 64 else 5e-6)


@pytest.mark.parametrize("fit_sources", [False, True])
def test_psf_shift_fit_matches_oracle(fit_sources):
    """the offset is fitted by the same AMSGrad step as every other parameter (blend.py:103-105: renderer parameters close
    the parameter tuple); started away from the truth it moves, alone or together with the sources"""
    g = golden("psf_shift.npz")
    start = np.array([0.6, -0.2])
    for precision, tol in ((64, 1e-8), (32, 1e-4)):
        o = _psf_shift_oracle(g, start, fit_sources, frame_dtype=np.float64 if precision == 64 else np.float32)
        o.fit(max_iter=12, e_rel=1e-3, min_iter=10 ** 9)
        blend, obs, renderer = _psf_shift_blend(g, start, precision, fit_sources)
        n, logL = blend.fit(max_iter=12, e_rel=1e-3, min_iter=10 ** 9)
        assert n == 12
        assert_allclose(np.array(blend.loss), np.array(o.loss), rtol=max(tol, 1e-9))
        shift = np.asarray(renderer.parameters[0])
        assert np.abs(shift - o.observations[0].psf_shift.x).max() < tol
        assert np.abs(shift - start).max() > 0.05  # it really moved
        assert renderer.parameters[0].m is not None
        if fit_sources:
            for comp, osrc in zip(blend.sources, o.sources):
                assert rel_peak(comp.parameters[0], osrc.spectrum.x) < tol
                assert rel_peak(comp.parameters[1], osrc.image.x) < tol
