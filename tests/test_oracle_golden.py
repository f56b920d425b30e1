"""CPU: pin the oracle (oracle/scarlet_oracle.py) against the reference's own known-answer tests and against
fixtures produced by the reference's forward code (tests/golden/make_golden.py)."""
import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from conftest import ARANGE25, MONO_KATS, SYM_HALF, golden
from oracle import scarlet_oracle as so


@pytest.mark.parametrize("kind,min_grad,expect", MONO_KATS)
def test_monotonic_reference_kat(kind, min_grad, expect):
    """reference tests/test_constraint.py:93-138"""
    for fn in (so.prox_monotonic, so.prox_monotonic_python):
        out = fn(ARANGE25.copy(), kind, min_grad)
        assert_allclose(out, expect, atol=5e-8)


def test_symmetry_reference_kat():
    """reference tests/test_constraint.py:140-163"""
    assert_allclose(so.prox_symmetry(ARANGE25.copy(), 1.0), np.full((5, 5), 12.0))
    assert_allclose(so.prox_symmetry(ARANGE25.copy(), 0.5), SYM_HALF)


def test_positivity_normalization_center_on():
    """reference tests/test_constraint.py:8-33,165-172"""
    rng = np.random.default_rng(0)
    X = rng.standard_normal((6, 7))
    assert (so.prox_positivity(X, 0.1) >= 0.1).all()
    Y = np.abs(X) + 0.1
    assert_allclose(so.prox_normalization(Y.copy(), "sum").sum(), 1)
    assert_allclose(so.prox_normalization(Y.copy(), "max").max(), 1)
    assert so.prox_center_on(np.zeros((5, 5)))[2, 2] > 0


def test_pad_center_reference_kat():
    """reference tests/test_fft.py:12-77"""
    a_pad = so.pad_to(np.ones((1, 1)), (5, 4))
    assert a_pad[2, 2] == 1 and a_pad.sum() == 1
    assert np.fft.ifftshift(a_pad)[0, 0] == 1
    a0 = np.arange(10).reshape(5, 2)
    a_pad = so.pad_to(a0, (9, 11))
    assert_array_equal(a_pad[2:7, 5:7], a0)
    assert a_pad.sum() == a0.sum()
    sh = np.fft.ifftshift(a_pad)
    assert_array_equal(sh[:3, :2], [[4, 5], [6, 7], [8, 9]])
    assert_array_equal(sh[7:, :2], [[0, 1], [2, 3]])
    assert_array_equal(so.centered(a_pad, (5, 2)), a0)


def test_match_psf_roundtrip():
    """reference tests/test_fft.py:88-124, narrow -> wide direction (the direction the renderer uses).  The
    reference's wide -> narrow assertions only hold because its Fourier object caches the k-space ratio on the
    same grid; the fitting path passes the kernel IMAGE (renderer.py:229-233), so that cache is not on the path."""
    p1 = so.GaussianPSFOracle([1.0], boxsize=41).get_model()
    p2 = so.GaussianPSFOracle([2.0], boxsize=41).get_model()
    p3 = so.GaussianPSFOracle([1.0, 2.0, 3.0], boxsize=41).get_model()
    assert_allclose(so.convolve(p1, so.match_psf(p2, p1)), p2, atol=1e-7)
    k = so.match_psf(p3, p1)
    assert k.shape == (3, 41, 41)
    assert_allclose(so.convolve(k, p1), p3, atol=1e-7)


def test_weights_vs_reference():
    g = golden("monotonic_weights.npz")
    for i in range(int(g["n"])):
        H, W, cy, cx = g["cfg%d" % i]
        center = None if cy < 0 else (int(cy), int(cx))
        w = so.monotonic_weights((int(H), int(W)), str(g["kind%d" % i]), center)
        assert_allclose(w, g["w%d" % i], atol=1e-15, err_msg="case %d" % i)


def test_chain_vs_reference():
    g = golden("prox_chain.npz")
    for i in range(int(g["n"])):
        kind, sym, mg = g["cfg%d" % i]
        chain = so.extended_source_chain(str(kind), bool(int(sym)), float(mg))
        out = chain(g["in%d" % i].copy())
        assert_allclose(out, g["out%d" % i], rtol=0, atol=1e-14, err_msg="case %d" % i)


def test_observation_render_and_loglike():
    """reference tests/test_observation.py:12-47, values from the reference's own code"""
    g = golden("obs_render_loss.npz")
    mpsf = so.GaussianPSFOracle([float(g["model_sigma"])] * 3, boxsize=int(g["model_boxsize"]))
    opsf = so.GaussianPSFOracle(g["obs_sigmas"], boxsize=int(g["obs_boxsize"]))
    assert_allclose(mpsf.get_model(), g["model_psf_image"], atol=1e-15)
    assert_allclose(opsf.get_model(), g["obs_psf_image"], atol=1e-15)
    obs = so.ObservationOracle(g["images"], None, opsf, frame_dtype=np.float32)
    obs.match((3, 43, 43), mpsf)
    # the fixture was produced by the reference under NumPy 2.x, whose FFT of the float32-cast PSF images runs in
    # complex64 (reference-era NumPy upcast to complex128, which is what the oracle does): agreement ~1e-7
    assert_allclose(obs.diff_kernel, g["diff_kernel"], atol=2e-7)
    rendered = obs.render(g["model"].astype(np.float32))
    assert_allclose(rendered, g["rendered"], atol=2e-7)
    assert_allclose(rendered, g["obs_psf_image"], atol=2e-7)  # the reference test's own assertion
    assert_allclose(-obs.neg_log_likelihood(g["model"].astype(np.float32)), float(g["logL"]), rtol=1e-8)


def _scene_from_golden(g, frame_dtype=np.float32):
    C = g["images"].shape[0]
    mpsf = so.GaussianPSFOracle([float(g["model_sigma"])] * C)
    obs = so.ObservationOracle(g["images"], g["weights"], so.ImagePSFOracle(g["psfs"]), frame_dtype=frame_dtype)
    obs.match(g["images"].shape, mpsf)
    srcs = []
    for k in range(int(g["n_sources"])):
        if str(g["src%d_kind" % k]) == "PointSource":
            srcs.append(so.PointSourceOracle(g["src%d_spectrum" % k], g["src%d_center" % k], mpsf, min_step=g["noise_rms_band"]))
        else:
            srcs.append(so.ExtendedSourceOracle(g["src%d_spectrum" % k], g["src%d_image" % k], g["src%d_origin" % k][1:],
                                                min_step=g["noise_rms_band"]))
    return so.SceneOracle(g["images"].shape, mpsf, srcs, [obs], frame_dtype=frame_dtype), obs


@pytest.mark.parametrize("name", ["hsc_cosmos_35.npz", "point_extended.npz"])
def test_scene_forward_vs_reference(name):
    g = golden(name)
    scene, obs = _scene_from_golden(g)
    assert_allclose(obs.diff_kernel, g["diff_kernel"], atol=1e-8)
    assert_allclose(obs.log_norm, float(g["log_norm"]), rtol=1e-9)  # reference sums log(rms) in float32
    for k, src in enumerate(scene.sources):
        assert_allclose(src.get_model(), g["src%d_model" % k], rtol=1e-6, atol=1e-9)
        assert src.bbox.origin == tuple(g["src%d_origin" % k])
        assert_allclose(src.spectrum.step_size(0), g["src%d_spectrum_step" % k], rtol=1e-6)
    model = scene.get_model()
    assert_allclose(model, g["model"], rtol=1e-6, atol=1e-6)
    assert_allclose(obs.render(model), g["rendered"], rtol=1e-5, atol=1e-5 * np.abs(g["rendered"]).max())
    assert_allclose(-scene.loss_only(), float(g["logL"]), rtol=1e-6)


@pytest.mark.parametrize("name", ["hsc_cosmos_35.npz", "point_extended.npz"])
def test_gradients_vs_reference_finite_differences(name):
    """hand adjoints == central differences of the REFERENCE forward (float64 frame)"""
    g = golden(name)
    scene, _ = _scene_from_golden(g, frame_dtype=np.float64)
    _, grads = scene.loss_and_grads()
    for row, fd in zip(g["fd_which"], g["fd_grad"]):
        row = [int(r) for r in np.atleast_1d(row)]
        gp = np.asarray(grads[row[0]])
        val = gp[tuple(r for r in row[1:] if r >= 0)] if gp.ndim == len([r for r in row[1:] if r >= 0]) else gp[row[1]]
        assert_allclose(val, fd, rtol=2e-4, atol=1e-6 * max(1.0, abs(fd)))  # FD noise: complex64 FFT in the fixture run


def test_point_source_morphology_vs_reference():
    g = golden("point_extended.npz")
    scene, _ = _scene_from_golden(g)
    src = scene.sources[int(g["ps_index"])]
    cen = src.center.x.copy()
    for off, expect in zip(g["ps_offsets"], g["ps_models"]):
        assert_allclose(src.morph(cen + off), expect, atol=1e-15)


def test_gradients_vs_torch_autograd():
    """independent check: torch CPU float64 autograd over the same forward"""
    import torch
    from scarlet_b200 import synthetic
    from oracle import scenes
    sc = synthetic.make_scene("tiny", 3)
    o = scenes.build_oracle(sc, frame_dtype=np.float64, sed_dtype=np.float64)
    loss, grads = o.loss_and_grads()
    obs = o.observations[0]
    C, N = sc["C"], sc["N"]
    fshape = so.get_fft_shape((C, N, N), obs.diff_kernel.shape, 3, (1, 2))
    khat = torch.from_numpy(obs._kernel_fft(fshape))
    params = []
    model = torch.zeros((C, N, N), dtype=torch.float64)
    for src in o.sources:
        sed = torch.tensor(np.asarray(src.spectrum.x, dtype=np.float64), requires_grad=True)
        if src.kind != "extended":
            params.append((sed, None))
            m = torch.from_numpy(src.morph()[0])
        else:
            m = torch.tensor(src.image.x, requires_grad=True)
            params.append((sed, m))
        fs, ms = so.overlapped_slices(o.frame_box, src.bbox)
        full = torch.zeros((C, N, N), dtype=torch.float64)
        full[fs] = (sed[:, None, None] * m[None])[ms]
        model = model + full
    pad = torch.zeros((C, fshape[0], fshape[1]), dtype=torch.float64)
    sy, sx = (fshape[0] - N + 1) // 2, (fshape[1] - N + 1) // 2
    pad[:, sy:sy + N, sx:sx + N] = model
    conv = torch.fft.fftshift(torch.fft.irfftn(torch.fft.rfftn(torch.fft.ifftshift(pad, dim=(1, 2)), dim=(1, 2)) * khat,
                                               s=fshape, dim=(1, 2)), dim=(1, 2))[:, sy:sy + N, sx:sx + N]
    tl = obs.log_norm + 0.5 * (torch.from_numpy(obs.weights.astype(np.float64)) * (conv - torch.from_numpy(obs.data.astype(np.float64))) ** 2).sum()
    tl.backward()
    assert_allclose(float(tl), loss, rtol=1e-12)
    i = 0
    for src, (sed, m) in zip(o.sources, params):
        assert_allclose(grads[i], sed.grad.numpy(), rtol=1e-9, atol=1e-9)
        if m is not None:
            assert_allclose(grads[i + 1], m.grad.numpy(), rtol=1e-9, atol=1e-8)
        i += len(src.parameters)


def test_fit_runs_and_descends():
    from scarlet_b200 import synthetic
    from oracle import scenes
    o = scenes.build_oracle(synthetic.make_scene("tiny", 1))
    n, logL = o.fit(max_iter=25, e_rel=1e-9)
    assert n == 25 and o.loss[-1] < 0.5 * o.loss[0]
    for p in o.parameters:
        assert np.isfinite(p.x).all() and p.std is not None


def test_resolution_oracle_vs_reference_fixture():
    """the literal restatement of ResolutionRenderer (renderer.py:262-547) against the reference's own render and
    log-likelihood; its adjoint against the dot-product identity"""
    from oracle import scarlet_oracle as so
    g = golden("multires.npz")
    o = so.ResolutionObservationOracle(g["lr_images"], g["lr_weights"], g["lr_diff_kernel64"], g["lr_shifts64"], float(g["lr_h64"]),
                                       frame_dtype=np.float64)
    o.match(tuple(g["frame_shape64"]), None)
    lr = o.render(g["model64"])
    assert_allclose(lr, g["lr_rendered64"], atol=1e-12 * np.abs(g["lr_rendered64"]).max())
    assert_allclose(-o.neg_log_likelihood(g["model64"]), float(g["lr_logL64"]), rtol=1e-12)
    rng = np.random.default_rng(0)
    R, M = rng.standard_normal(lr.shape), rng.standard_normal(g["model64"].shape)
    lhs, rhs = (o.render(M) * R).sum(), (M * o.render_adjoint(R)).sum()
    assert abs(lhs - rhs) < 1e-12 * abs(lhs)
    hr = so.ObservationOracle(g["hr_images"], g["hr_weights"], so.ImagePSFOracle(g["hr_psfs"]), frame_dtype=np.float64, channel_offset=5,
                              origin=tuple(int(v) for v in g["hr_model_slice_start64"]))
    hr.match(tuple(g["frame_shape64"]), so.ImagePSFOracle(g["model_psf64"]))
    assert_allclose(hr.render(g["model64"]), g["hr_rendered64"], atol=1e-12 * np.abs(g["hr_rendered64"]).max())
    assert_allclose(-hr.neg_log_likelihood(g["model64"]), float(g["hr_logL64"]), rtol=1e-12)


def test_psf_shift_oracle_vs_reference_fixture():
    """``ConvolutionRenderer(psf_shift=...)`` (renderer.py:172-177, 220-227): the oracle's shifted kernel, render and logL equal
    the reference's own outputs, its hand gradient wrt the shift equals central differences of the reference's logL."""
    from oracle import scarlet_oracle as so
    g = golden("psf_shift.npz")
    C = g["images"].shape[0]
    for tag in "ab":
        obs = so.ObservationOracle(g["images"], g["weights"], so.ImagePSFOracle(g["psfs"]), frame_dtype=np.float64, psf_shift=g["shift_" + tag])
        obs.match(g["model"].shape, so.GaussianPSFOracle((float(g["model_sigma"]),) * C))
        assert_allclose(obs.diff_kernel, g["diff_kernel"], atol=1e-14)
        assert_allclose(obs.shifted_kernel(), g["kernel_" + tag], atol=1e-14)
        assert_allclose(obs.render(g["model"]), g["rendered_" + tag], atol=1e-11)
        assert_allclose(-obs.neg_log_likelihood(g["model"]), float(g["logL_" + tag]), rtol=1e-13)
        assert_allclose(obs.param_grads(g["model"])[0], g["dloss_dshift_" + tag], rtol=1e-6)
