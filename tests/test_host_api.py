"""CPU: host-side mirror of the reference API, table builders, the C-ABI export list, and the no-fallback rule."""
import ctypes
import os
import re

import numpy as np
import pytest
from numpy.testing import assert_allclose, assert_array_equal

from conftest import ROOT, golden


def test_box_geometry():
    """reference tests/test_bbox.py:6-54"""
    from scarlet_b200 import Box, overlapped_slices
    x = np.zeros((5, 6))
    x[1:3, 2:5] = 1
    b = Box.from_data(x)
    assert b.shape == (2, 3) and b.origin == (1, 2)
    assert b.contains((1, 2)) and not b.contains((3, 2))
    img = np.arange(30).reshape(5, 6)
    assert_array_equal(b.extract_from(img), img[1:3, 2:5])
    # boxes hanging over the edge
    over = Box((3, 3), origin=(-1, 4))
    sub = over.extract_from(img)
    assert_array_equal(sub[1:, :2], img[0:2, 4:6])
    assert (sub[0] == 0).all() and (sub[:, 2] == 0).all()
    canvas = np.zeros((5, 6), dtype=int)
    over.insert_into(canvas, np.ones((3, 3), dtype=int))
    assert canvas.sum() == 4
    s1, s2 = overlapped_slices(Box((5, 6)), over)
    assert s1 == (slice(0, 2), slice(4, 6)) and s2 == (slice(1, 3), slice(0, 2))
    assert (Box((2,)) @ Box((3, 4), origin=(1, 1))).shape == (2, 3, 4)
    assert (Box((3, 3)) | Box((2, 2), origin=(4, 4))).shape == (6, 6)
    assert (Box((3, 3)) & Box((2, 2), origin=(4, 4))).shape == (0, 0)
    assert Box((3, 3), origin=(1, 1)) + (1, 2) == Box((3, 3), origin=(2, 3))


def test_pad_center_conventions():
    """reference tests/test_fft.py:12-77"""
    from scarlet_b200 import fft
    a_pad = fft._pad(np.ones((1, 1)), (5, 4))
    truth = np.zeros((5, 4))
    truth[2, 2] = 1
    assert_array_equal(a_pad, truth)
    a0 = np.arange(10).reshape(5, 2)
    a_pad = fft._pad(a0, (9, 11))
    truth = np.zeros((9, 11), dtype=int)
    truth[2:7, 5:7] = a0
    assert_array_equal(a_pad, truth)
    assert_array_equal(fft._centered(a_pad, (5, 2)), a0)
    from oracle import scarlet_oracle as so
    for s1, s2 in (((5, 58, 48), (5, 43, 43)), ((5, 128, 128), (5, 21, 21)), ((5, 256, 256), (5, 41, 41)), ((3, 17, 20), (3, 8, 6))):
        assert fft._get_fft_shape(s1, s2, 3, (1, 2)) == so.get_fft_shape(s1, s2, 3, (1, 2))
    assert fft._get_fft_shape((5, 256, 256), (5, 41, 41), 3, (1, 2)) == [300, 300]


def test_weight_tables_vs_reference_fixture():
    from scarlet_b200 import operator
    g = golden("monotonic_weights.npz")
    for i in range(int(g["n"])):
        H, W, cy, cx = g["cfg%d" % i]
        center = None if cy < 0 else (int(cy), int(cx))
        w = operator.getRadialMonotonicWeights((int(H), int(W)), str(g["kind%d" % i]), center)
        assert_allclose(w, g["w%d" % i], atol=1e-15, err_msg="case %d" % i)
    w, off, idx = operator.monotonic_tables((5, 7), "angle", (2, 3))
    assert w.shape == (8, 35) and idx.size == 34 and 2 * 7 + 3 not in idx
    assert_array_equal(off, [-8, -7, -6, -1, 1, 6, 7, 8])


def test_diff_kernel_vs_reference_fixture():
    import scarlet_b200 as sb
    g = golden("hsc_cosmos_35.npz")
    C = g["images"].shape[0]
    frame = sb.Frame(g["images"].shape, psf=sb.GaussianPSF(sigma=(0.8,) * C), channels=list(range(C)))
    obs = sb.Observation(g["images"].copy(), psf=sb.ImagePSF(g["psfs"].copy()), weights=g["weights"].copy(), channels=list(range(C)))
    obs.match(frame)
    assert type(obs.renderer).__name__ == "ConvolutionRenderer"
    assert_allclose(obs.renderer.diff_kernel.image, g["diff_kernel"], atol=2e-7)
    assert_allclose(obs.log_norm, float(g["log_norm"]), rtol=1e-9)
    assert_allclose(np.array(np.mean(obs.noise_rms, axis=(1, 2))), g["noise_rms_band"], rtol=1e-6)
    fshape, khat = obs.renderer.kernel_transform()
    assert tuple(fshape) == (108, 96) and khat.shape == (5, 108, 49)
    pg = golden("obs_render_loss.npz")
    assert_allclose(sb.GaussianPSF(pg["obs_sigmas"], boxsize=43).get_model(), pg["obs_psf_image"], atol=1e-15)


def test_parameter_roundtrip_and_model_tree():
    import pickle
    import scarlet_b200 as sb
    from scarlet_b200 import synthetic
    p = sb.Parameter(np.arange(3.0), name="x", step=0.1, m=np.ones(3))
    q = pickle.loads(pickle.dumps(p))
    assert q.name == "x" and q.step == 0.1 and (q.m == 1).all() and (q[1:]).name == "x"
    sc = synthetic.make_scene("tiny", 0)
    b = synthetic.make_blend.__wrapped__(sc) if hasattr(synthetic.make_blend, "__wrapped__") else None
    frame = sb.Frame(sc["images"].shape, psf=sb.GaussianPSF(sigma=(0.8,) * 3), channels=sc["channels"])
    obs = sb.Observation(sc["images"], psf=sb.ImagePSF(sc["obs_psf"].copy()), weights=sc["weights"], channels=sc["channels"])
    obs.match(frame)
    s = sc["sources"][0]
    src = sb.ExtendedSource(frame, s["center"], obs, spectrum=s["sed"], morphology=s["morph"], resizing=False)
    names = [p.name for p in src.parameters]
    assert names == ["spectrum", "image", "shift"]  # order of model.py:51-54 / morphology.py:112-113
    assert src.bbox.shape == (3, 15, 15)
    assert src.get_model().shape == (3, 15, 15)
    assert_allclose(src.parameters[0].step(src.parameters[0], it=0), max(1.0, 0.01 * s["sed"].mean()))
    pt = sb.PointSource(frame, (20.3, 11.8), obs, spectrum=s["sed"])
    assert [p.name for p in pt.parameters] == ["spectrum", "center"] and pt.bbox.shape == (3, 9, 9)
    from oracle import scarlet_oracle as so
    po = so.PointSourceOracle(s["sed"], (20.3, 11.8), so.GaussianPSFOracle([0.8] * 3))
    assert_allclose(pt.get_model(), po.get_model(), rtol=1e-6)
    assert pt.bbox.origin == po.bbox.origin
    bad = np.array([1.0, np.nan, 2.0])
    with pytest.raises(ArithmeticError):
        sb.TabulatedSpectrum(frame, bad)


def test_abi_exports_every_declared_symbol():
    """the shared library must export exactly what include/scarlet_b200.h declares (no compute calls here)"""
    from scarlet_b200 import _build, _native
    header = open(os.path.join(ROOT, "include", "scarlet_b200.h")).read()
    declared = set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(_native.SYMBOLS), declared ^ set(_native.SYMBOLS)
    path = _build.build()
    handle = ctypes.CDLL(path)
    for name in declared:
        assert hasattr(handle, name), name
    lib = _native.lib()
    assert lib.sb_version().startswith(b"scarlet_b200")
    assert lib.sb_stage_name(0) == b"render"


def test_no_cpu_fallback(has_gpu):
    """without a CUDA device every compute entry point fails loudly"""
    if has_gpu:
        pytest.skip("a GPU is present")
    import scarlet_b200 as sb
    from scarlet_b200 import _native, synthetic
    with pytest.raises(_native.NativeError, match="no CUDA device"):
        sb.PositivityConstraint()(np.zeros((3, 3)), 0)
    b = synthetic.make_blend(synthetic.make_scene("tiny", 0))
    with pytest.raises(_native.NativeError, match="no CUDA device"):
        b.fit(max_iter=2)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "scarlet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "/root/reference" not in src, f


def test_algorithmic_bytes_formula():
    """SURVEY.md 8(d) figures"""
    from scarlet_b200 import synthetic
    assert abs(synthetic.algorithmic_bytes(synthetic.CONFIGS["cfg2"], (160, 160)) / 1e6 - 8.44) < 0.05
    assert abs(synthetic.algorithmic_bytes(synthetic.CONFIGS["cfg5"], (180, 180)) / 1e6 - 10.3) < 0.1


def test_host_gather_scatter_roundtrip():
    """the C pack/unpack helpers behind Parameter transfer (no GPU involved)"""
    from scarlet_b200 import _native as nat
    rng = np.random.default_rng(0)
    arrs = [rng.random(5).astype(np.float32), rng.random((7, 3)), rng.random(1), rng.random((4, 4)).astype(np.float32)]
    ptrs = np.array([a.ctypes.data for a in arrs], dtype=np.uint64)
    counts = np.array([a.size for a in arrs], dtype=np.int64)
    f32 = np.array([a.dtype == np.float32 for a in arrs], dtype=np.int32)
    packed = np.zeros(int(counts.sum()))
    nat.check(nat.lib().sb_host_gather_f64(nat.ptr(packed), nat.ptr(ptrs), nat.ptr(counts), nat.ptr(f32), len(arrs)))
    assert_array_equal(packed, np.concatenate([a.reshape(-1).astype(np.float64) for a in arrs]))
    keep = [a.copy() for a in arrs]
    packed2 = packed * 2
    nat.check(nat.lib().sb_host_scatter_f64(nat.ptr(packed2), nat.ptr(ptrs), nat.ptr(counts), nat.ptr(f32), len(arrs)))
    for a, k in zip(arrs, keep):
        assert_allclose(a, 2 * k, rtol=1e-7)


def test_parameter_lazy_state_links():
    """m/v/vhat/std of a Parameter linked to a plan's packed host arrays (what download_parameters leaves behind)"""
    import pickle
    import scarlet_b200 as sb
    from scarlet_b200._plan import _HostStore
    from scarlet_b200.parameter import _StateLink
    store = _HostStore()
    for key in _HostStore.KEYS:
        store.arrays[key] = dict(morph=np.arange(12.0) + 100 * _HostStore.KEYS.index(key))
    p = sb.Parameter(np.zeros((2, 3)), name="image", step=0.01)
    p.__dict__["_link"] = _StateLink(store, "morph", 6, 12, (2, 3))
    assert p.m is None and p.std is None          # store not valid yet: nothing to show
    store.valid = True
    assert_array_equal(p.v, (np.arange(6.0) + 206).reshape(2, 3))
    assert p.v.base is not None                    # a view, not a copy
    assert_allclose(p.std, 1 / np.sqrt(p.v))
    store.arrays["v"]["morph"][6] = 0
    assert p.std.mask[0, 0] and not p.std.mask[1, 2]   # masked where v == 0 (blend.py:189-192)
    p.m = np.ones((2, 3))                          # explicit assignment wins
    assert (p.m == 1).all()
    q = pickle.loads(pickle.dumps(p))
    assert_array_equal(q.v, p.v)
    assert q.name == "image"
    assert (q.m == 1).all() and q.std is not None
    assert (p[0]).name == "image" and p[0].v.shape == (2, 3)   # attributes travel by reference to views


def _multires_scene(dtype=np.float32):
    import scarlet_b200 as sb
    from scarlet_b200.wcs import AffineWCS
    g = golden("multires.npz")
    obs_hr = sb.Observation(g["hr_images"].copy(), psf=sb.ImagePSF(g["hr_psfs"].copy()), weights=g["hr_weights"].copy(),
                            wcs=AffineWCS(g["hr_cd"], crpix=g["hr_crpix"]), channels=["h0", "h1", "h2"])
    obs_lr = sb.Observation(g["lr_images"].copy(), psf=sb.ImagePSF(g["lr_psfs"].copy()), weights=g["lr_weights"].copy(),
                            wcs=AffineWCS(g["lr_cd"], crpix=g["lr_crpix"]), channels=["l0", "l1", "l2", "l3", "l4"])
    frame = sb.Frame.from_observations([obs_lr, obs_hr], coverage="union")
    if dtype is np.float64:
        frame = sb.Frame(frame.shape, channels=frame.channels, psf=frame.psf, wcs=frame.wcs, dtype=np.float64)
        obs_lr.match(frame)
        obs_hr.match(frame)
    return g, frame, obs_lr, obs_hr


def test_multiresolution_setup_vs_reference_fixture():
    """Frame.from_observations + ResolutionRenderer set-up against the reference's own (tests/golden/make_golden.py:multires),
    and the Fourier identity the device evaluates against the reference's render"""
    g, frame, obs_lr, obs_hr = _multires_scene(np.float64)
    assert tuple(frame.shape) == tuple(g["frame_shape64"])
    assert_allclose(frame.psf.get_model(), g["model_psf64"], atol=1e-15)
    assert_allclose(frame.wcs.wcs.crpix, g["model_crpix64"])
    r, r2 = obs_lr.renderer, obs_hr.renderer
    assert type(r).__name__ == "ResolutionRenderer" and type(r2).__name__ == "ConvolutionRenderer"
    assert_allclose(r.h, float(g["lr_h64"]))
    assert list(r._fft_shape) == list(g["lr_fft_shape64"]) and bool(r.small_axis) == bool(g["lr_small_axis64"])
    assert_allclose(r.shifts, g["lr_shifts64"], atol=1e-12)
    assert_allclose(r.diff_kernel.image, g["lr_diff_kernel64"], atol=1e-12)
    assert_allclose(r2.diff_kernel.image, g["hr_diff_kernel64"], atol=1e-12)
    assert r2.origin == tuple(g["hr_model_slice_start64"]) and r.channel_offset == 0 and r2.channel_offset == 5
    lr = obs_lr.render(g["model64"])
    assert_allclose(lr, g["lr_rendered64"], atol=1e-11 * np.abs(g["lr_rendered64"]).max())
    # the reference's default float32 frame (float32 model, float32 resampling operator) only agrees to float32 rounding
    assert np.abs(lr - g["lr_rendered"]).max() < 5e-5 * np.abs(g["lr_rendered"]).max()
    assert_allclose(obs_lr.get_log_likelihood(g["model64"]), float(g["lr_logL64"]), rtol=1e-10)


def test_multiresolution_intersection_coverage_vs_reference_fixture():
    """Frame.from_observations(obs_id=1, coverage="intersection") (frame.py:200-312) and both renderers on that frame against
    the reference's own products (tests/golden/make_golden.py:multires, keys isect_*)."""
    import scarlet_b200 as sb
    from scarlet_b200.wcs import AffineWCS
    g = golden("multires.npz")
    obs_hr = sb.Observation(g["hr_images"].copy(), psf=sb.ImagePSF(g["hr_psfs"].copy()), weights=g["hr_weights"].copy(),
                            wcs=AffineWCS(g["hr_cd"], crpix=g["hr_crpix"]), channels=["h0", "h1", "h2"])
    obs_lr = sb.Observation(g["lr_images"].copy(), psf=sb.ImagePSF(g["lr_psfs"].copy()), weights=g["lr_weights"].copy(),
                            wcs=AffineWCS(g["lr_cd"], crpix=g["lr_crpix"]), channels=["l0", "l1", "l2", "l3", "l4"])
    frame = sb.Frame.from_observations([obs_lr, obs_hr], obs_id=1, coverage="intersection")
    assert tuple(frame.shape) == tuple(g["isect_frame_shape"])
    assert_allclose(frame.wcs.wcs.crpix, g["isect_model_crpix"], atol=1e-12)
    frame = sb.Frame(frame.shape, channels=frame.channels, psf=frame.psf, wcs=frame.wcs, dtype=np.float64)
    obs_lr.match(frame)
    obs_hr.match(frame)
    r = obs_lr.renderer
    assert type(r).__name__ == str(g["isect_lr_renderer"]) and type(obs_hr.renderer).__name__ == str(g["isect_hr_renderer"])
    assert_allclose(r.h, float(g["isect_lr_h"]))
    assert list(r._fft_shape) == list(g["isect_lr_fft_shape"])
    assert_allclose(r.shifts, g["isect_lr_shifts"], atol=1e-11)
    model = g["isect_model"]
    assert_allclose(obs_lr.render(model), g["isect_lr_rendered"], atol=1e-11 * np.abs(g["isect_lr_rendered"]).max())
    assert_allclose(obs_lr.get_log_likelihood(model), float(g["isect_lr_logL"]), rtol=1e-10)
    # (the high-resolution render runs on the device: tests/test_gpu_parity.py::test_intersection_frame_convolution_render)


def _multires_rot_scene():
    import scarlet_b200 as sb
    from scarlet_b200.wcs import AffineWCS
    g = golden("multires_rot.npz")
    obs_hr = sb.Observation(g["hr_images"].copy(), psf=sb.ImagePSF(g["hr_psfs"].copy()), weights=g["hr_weights"].copy(),
                            wcs=AffineWCS(g["hr_cd"], crpix=g["hr_crpix"]), channels=["h0", "h1", "h2"])
    obs_lr = sb.Observation(g["lr_images"].copy(), psf=sb.ImagePSF(g["lr_psfs"].copy()), weights=g["lr_weights"].copy(),
                            wcs=AffineWCS(g["lr_cd"], crpix=g["lr_crpix"]), channels=["l0", "l1", "l2", "l3", "l4"])
    frame = sb.Frame.from_observations([obs_lr, obs_hr], coverage="union")
    frame = sb.Frame(frame.shape, channels=frame.channels, psf=frame.psf, wcs=frame.wcs, dtype=np.float64)
    obs_lr.match(frame)
    obs_hr.match(frame)
    return g, frame, obs_lr, obs_hr


def test_rotated_multiresolution_setup_and_render_vs_reference_fixture():
    """The rotated branch of ResolutionRenderer (renderer.py:318-363, 498-524): set-up products and the half-plane multiplier
    form of the render against the reference's own (tests/golden/make_golden.py:multires_rot); the oracle's literal
    restatement and its adjoint against the same fixture."""
    from oracle import scarlet_oracle as so
    g, frame, obs_lr, obs_hr = _multires_rot_scene()
    assert tuple(frame.shape) == tuple(g["frame_shape"])
    r = obs_lr.renderer
    assert type(r).__name__ == "ResolutionRenderer" and r.isrot
    assert_allclose([float(r.angle[0]), float(r.angle[1])], g["lr_angle"], atol=1e-14)
    assert_allclose(r.h, float(g["lr_h"]))
    assert list(r._fft_shape) == list(g["lr_fft_shape"]) and bool(r.small_axis) == bool(g["lr_small_axis"])
    assert_allclose(r.shifts, g["lr_shifts"], atol=1e-11)
    assert_allclose(r.other_shifts, g["lr_other_shifts"], atol=1e-11)
    assert_allclose(r.diff_kernel.image, g["lr_diff_kernel"], atol=1e-12)
    peak = np.abs(g["lr_rendered"]).max()
    assert_allclose(obs_lr.render(g["model"]), g["lr_rendered"], atol=1e-11 * peak)
    assert_allclose(obs_lr.get_log_likelihood(g["model"]), float(g["lr_logL"]), rtol=1e-10)
    o = so.RotatedResolutionObservationOracle(g["lr_images"], g["lr_weights"], g["lr_diff_kernel"], g["lr_shifts"], g["lr_other_shifts"],
                                              float(g["lr_h"]), small_axis=bool(g["lr_small_axis"]), frame_dtype=np.float64)
    o.match(tuple(int(v) for v in g["frame_shape"]), None)
    assert_allclose(o.render(g["model"]), g["lr_rendered"], atol=1e-11 * peak)
    rng = np.random.default_rng(0)
    G, M = rng.standard_normal(g["lr_rendered"].shape), rng.standard_normal(g["model"].shape)
    assert_allclose((o.render(M) * G).sum(), (o.render_adjoint(G) * M).sum(), rtol=1e-12)


def test_plan_classifies_renderers_without_a_device():
    """host half of the plan (no GPU): which device observation kind a matched renderer maps to, and what is refused"""
    import scarlet_b200 as sb
    from multires_scene import product_scene
    from scarlet_b200._plan import DevicePlan
    for rotated, kind, tables in ((False, 2, ("Ey", "Ex")), (True, 3, ("A", "B"))):
        _, blend, obs_lr, obs_hr = product_scene(32, rotated)
        m_lr, m_hr = DevicePlan._obs_meta(blend, 0), DevicePlan._obs_meta(blend, 1)
        assert (m_lr["kind"], m_hr["kind"]) == (kind, 0)
        assert all(t in m_lr["operator"] for t in tables) and m_lr["operator"]["rotated"] == rotated
        Fy, Fx = m_lr["fshape"]
        assert m_lr["operator"]["khat"].shape == (5, Fy, Fx // 2 + 1)
        if rotated:  # half-plane multiplier tables, one per low-resolution row / column
            assert m_lr["operator"]["A"].shape == (8, Fy, Fx // 2 + 1) and m_lr["operator"]["B"].shape == (8, Fy, Fx // 2 + 1)
    # a renderer the device path does not know
    class Other:  # (stands for any user-written Renderer subclass)
        pass
    obs_hr.renderer = Other()
    with pytest.raises(TypeError):
        DevicePlan._obs_meta(blend, 1)


def test_edge_pull_equals_the_masked_array_expression():
    """ImageMorphology.update's grow rule (morphology.py:165-177) is evaluated on the four edges only: bit-identical to the
    reference's masked-array expression over the whole image, zeros of v masked, fully masked edges -> nan"""
    import warnings
    import numpy.ma as ma
    from scarlet_b200.morphology import _edge_pull
    rng = np.random.default_rng(0)
    edges = ((slice(None), 0), (slice(None), -1), (0, slice(None)), (-1, slice(None)))
    for t in range(60):
        B = int(rng.choice([15, 21, 31, 41, 51]))
        m, v = rng.standard_normal((B, B)), rng.uniform(0, 1, (B, B)) ** 4
        v[rng.uniform(size=(B, B)) < rng.choice([0, 0.1, 0.9])] = 0
        if t % 7 == 0:
            v[:, 0] = 0
        data = (rng.uniform(size=(B, B)) * (rng.uniform(size=(B, B)) > 0.3)).astype(np.float32)
        step = 0.01 / 2 ** int(rng.integers(0, 3))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            gp = (-m / np.sqrt(np.sqrt(ma.masked_equal(v, 0))) * step) * (data > 0)
            ref = np.array((gp[:, 0].mean(), gp[:, -1].mean(), gp[0, :].mean(), gp[-1, :].mean()))
        new = np.array([_edge_pull(m[sl], v[sl], data[sl], step) for sl in edges])
        assert np.array_equal(ref, new, equal_nan=True)


def test_measure_helpers():
    """scarlet/measure.py:6-59 on a component and on a plain cube (host reductions; no device involved)"""
    import scarlet_b200 as sb
    from scarlet_b200 import measure, synthetic
    sc = synthetic.make_scene("tiny", 0)
    frame = sb.Frame(sc["images"].shape, psf=sb.GaussianPSF(sigma=(0.8,) * 3), channels=sc["channels"])
    s = sc["sources"][0]
    B = s["morph"].shape[0]
    src = sb.ExtendedSource(frame, s["center"], None, spectrum=s["sed"], morphology=s["morph"], bbox=sb.Box((B, B), origin=s["origin"]),
                            resizing=False)
    model = s["sed"][:, None, None].astype(np.float64) * s["morph"][None]
    assert_allclose(measure.flux(src), model.sum(axis=(1, 2)), rtol=1e-6)
    c, y, x = measure.max_pixel(src)
    assert (y, x) == (s["origin"][0] + B // 2, s["origin"][1] + B // 2)
    cen = measure.centroid(src)
    yy, xx = np.mgrid[:B, :B]
    assert_allclose(cen[1:], [(yy * s["morph"]).sum() / s["morph"].sum() + s["origin"][0], (xx * s["morph"]).sum() / s["morph"].sum() + s["origin"][1]], rtol=1e-6)
    assert_allclose(measure.flux(model), model.sum(axis=(1, 2)))


def test_batch_pipeline_orders_stages_and_overlaps_copies():
    """BatchPipeline (host logic, no GPU): copy-in and loop stages are entered in sequence order, one loop at a time; the
    next batch's copy-in runs while the current loop is busy; a batch object used twice is finished before it is reused."""
    import threading
    import time
    from scarlet_b200.blend import BatchPipeline
    log, lock = [], threading.Lock()

    def note(*ev):
        with lock:
            log.append(ev + (time.perf_counter(),))

    class FakePlan:
        def __init__(self, name):
            self.name = name

        def fit(self, opts):
            note("loop+", self.name)
            time.sleep(0.08)
            note("loop-", self.name)
            return (self.name, "loss", "status")

    class FakeBatch:
        def __init__(self, name):
            self.name, self.plans, self.uses = name, [FakePlan(name)], 0

        def _each(self, fn):
            return [fn(0)]

        def _copy_in(self, i, upload_observations):
            note("in+", self.name)
            time.sleep(0.03)
            note("in-", self.name)
            return 7

        def _copy_out(self, i, out):
            note("out+", self.name)
            time.sleep(0.03)
            note("out-", self.name)
            return 5

        def _finish(self, outs):
            self.uses += 1
            return (self.name, self.uses, outs[0][3], outs[0][4])

    a, b = FakeBatch("a"), FakeBatch("b")
    prepared = []
    pipe = BatchPipeline(depth=2)
    res = pipe.run([a, b, a, b, a], max_iter=3, prepare=lambda k, batch: prepared.append((k, batch.name)))
    assert res == [("a", 1, 7, 5), ("b", 1, 7, 5), ("a", 2, 7, 5), ("b", 2, 7, 5), ("a", 3, 7, 5)]
    # the stage intervals the pipeline records (bench.py derives the steady-state step period from them)
    tm = sorted(pipe.timings, key=lambda t: t["k"])
    assert [t["k"] for t in tm] == [0, 1, 2, 3, 4]
    for t in tm:
        assert t["copy_in"][0] <= t["copy_in"][1] <= t["loop"][0] <= t["loop"][1] <= t["copy_out"][1]
    assert all(t2["loop"][0] >= t1["loop"][1] for t1, t2 in zip(tm, tm[1:]))  # one loop at a time
    assert tm[1]["copy_in"][0] < tm[0]["loop"][1]                             # copies overlap the neighbour's loop
    assert prepared == [(0, "a"), (1, "b"), (2, "a"), (3, "b"), (4, "a")]
    order = [e[1] for e in log if e[0] == "loop+"]
    assert order == ["a", "b", "a", "b", "a"]
    # loops never overlap
    depth = 0
    for ev in log:
        if ev[0] == "loop+":
            depth += 1
            assert depth == 1
        elif ev[0] == "loop-":
            depth -= 1
    # the second batch copies in while the first one loops
    t = {(e[0], i): e[2] for i, e in enumerate(log)}
    first_loop_end = [e[2] for e in log if e[0] == "loop-"][0]
    second_in_start = [e[2] for e in log if e[0] == "in+"][1]
    assert second_in_start < first_loop_end
    # a reused batch object: its next copy-in starts only after its previous copy-out ended
    ins = [e[2] for e in log if e[0] == "in+" and e[1] == "a"]
    outs = [e[2] for e in log if e[0] == "out-" and e[1] == "a"]
    assert ins[1] > outs[0] and ins[2] > outs[1]


def test_batch_pipeline_propagates_errors():
    from scarlet_b200.blend import BatchPipeline

    class Bad:
        plans = [None]

        def _copy_in(self, i, u):
            raise RuntimeError("copy failed")

    import pytest
    with pytest.raises(RuntimeError, match="copy failed"):
        BatchPipeline(depth=2).run([Bad(), Bad(), Bad()], max_iter=1)


def _hsc_observation(weights=None):
    import scarlet_b200 as sb
    g = golden("hsc_cosmos_35.npz")
    C = g["images"].shape[0]
    frame = sb.Frame(g["images"].shape, psf=sb.GaussianPSF(sigma=(0.8,) * C), channels=list(range(C)))
    obs = sb.Observation(g["images"].copy(), psf=sb.ImagePSF(g["psfs"].copy()), weights=(g["weights"] if weights is None else weights).copy(),
                         channels=list(range(C)))
    obs.match(frame)
    return frame, obs


def test_get_psf_spectrum_vs_reference_fixture():
    """initialization.get_psf_spectrum (PSF-projected spectrum + matched-filter SNR, masked pixels left out, PSF box allowed
    to stick out of the image) against the reference's own outputs on data/hsc_cosmos_35 (make_golden.py:init_helpers)"""
    from scarlet_b200 import initialization as init
    h = golden("init_helpers.npz")
    frame, obs = _hsc_observation(weights=h["weights"])
    for k, center in enumerate(h["centers"]):
        spectrum, snr = init.get_psf_spectrum(tuple(center), obs, compute_snr=True)
        assert_allclose(spectrum, h["psf_spectrum"][k], rtol=2e-6)
        assert_allclose(snr, h["psf_snr"][k], rtol=2e-6)
    assert_allclose(init.get_psf_spectrum((1.0, 2.0), obs), h["edge_spectrum"], rtol=2e-6)
    per_obs = init.get_psf_spectrum(tuple(h["centers"][0]), (obs, obs), concat=False)
    assert len(per_obs) == 2 and per_obs[0].shape == (5,)


def test_init_source_component_count_and_fallback(monkeypatch):
    """initialization.init_source / init_all_sources control flow (initialization.py:287-490): component count capped by
    floor(psf_snr / min_snr), one component fewer after every ArithmeticError, compact source as the last resort, skipped
    list -- the counts equal the reference's on data/hsc_cosmos_35 (source construction itself is stubbed: it needs the GPU)"""
    import scarlet_b200.source as source_module
    from scarlet_b200 import initialization as init
    h = golden("init_helpers.npz")
    frame, obs = _hsc_observation(weights=h["weights"])
    calls = []

    class Fake:
        def __init__(self, K, compact, fail):
            self.K, self.compact, self.fail = K, compact, fail

        def check_parameters(self):
            if self.fail:
                raise ArithmeticError("not finite")

    bad = set()

    def fake_factory(model_frame, sky_coord, observations, K=1, compact=False, **kw):
        calls.append((K, compact))
        return Fake(K, compact, (0 if compact else K) in bad)

    monkeypatch.setattr(source_module, "ExtendedSource", fake_factory)
    for row, min_snr in zip(h["init_source_K"], (50, 200, 1000)):
        got = [init.init_source(frame, tuple(c), obs, max_components=2, min_snr=min_snr).K for c in h["centers"][:3]]
        assert got == list(row)
    # fallback chain 2 -> 1 -> compact; without fallback the error propagates
    bad.update({2, 1})
    calls.clear()
    src = init.init_source(frame, tuple(h["centers"][0]), obs, max_components=2)
    assert src.compact and calls == [(2, False), (1, False), (1, True)]
    bad.add(0)
    assert init.init_source(frame, tuple(h["centers"][0]), obs, max_components=2) is None
    with pytest.raises(ArithmeticError):
        init.init_source(frame, tuple(h["centers"][0]), obs, max_components=2, fallback=False)
    # init_all_sources: exceptions are collected when silent
    def exploding(model_frame, sky_coord, observations, **kw):
        if sky_coord == tuple(h["centers"][1]):
            raise ValueError("boom")
        return Fake(1, False, False)

    monkeypatch.setattr(source_module, "ExtendedSource", exploding)
    sources, skipped = init.init_all_sources(frame, [tuple(c) for c in h["centers"][:3]], obs, silent=True, set_spectra=False)
    assert len(sources) == 2 and skipped == [1]
    with pytest.raises(ValueError):
        init.init_all_sources(frame, [tuple(c) for c in h["centers"][:3]], obs, silent=False, set_spectra=False)


def test_set_spectra_to_match_vs_reference_fixture(monkeypatch):
    """initialization.set_spectra_to_match: weighted linear least squares for the amplitudes of all components per channel,
    given their unit-spectrum rendered models -- against the reference's result for three sources on data/hsc_cosmos_35 with
    a masked patch.  Observation.render (device) is replaced by the reference's rendered models from the fixture; a duplicated
    component shares its twin's spectrum."""
    import scarlet_b200 as sb
    from scarlet_b200 import initialization as init
    from scarlet_b200.bbox import Box
    h = golden("init_helpers.npz")
    frame, obs = _hsc_observation(weights=h["weights"])
    sources = []
    for k in range(3):
        img = h["src%d_image" % k]
        oy, ox = (int(v) for v in h["src%d_origin" % k][-2:])
        sources.append(sb.ExtendedSource(frame, tuple(h["centers"][k]), obs, spectrum=np.full(5, 3.0, dtype=np.float32), morphology=img,
                                         bbox=Box(img.shape, origin=(oy, ox)), resizing=False))
    for src in sources:  # the closing constraint projection runs on the device; the fixture's spectra are positive anyway
        src.parameters[0].constraint = None
    rendered = iter(h["unit_rendered"])
    seen = []

    def fake_render(model, *parameters):
        seen.append(np.asarray(model))
        return next(rendered).astype(np.float64)

    monkeypatch.setattr(obs, "render", fake_render, raising=False)
    init.set_spectra_to_match(sources, obs)
    assert len(seen) == 3
    for k, src in enumerate(sources):
        # the models handed to the renderer carry a flat unit spectrum
        assert_allclose(seen[k].sum(axis=(1, 2)), np.full(5, h["src%d_image" % k].sum()), rtol=1e-5)
        assert_allclose(np.asarray(src.parameters[0]), h["matched_spectra"][k], rtol=2e-4)
    # a duplicate of source 0 is left out of the solve and gets the same spectrum
    twin = sb.ExtendedSource(frame, tuple(h["centers"][0]), obs, spectrum=np.ones(5, dtype=np.float32), morphology=h["src0_image"],
                             bbox=sources[0].children[1].bbox.copy() if hasattr(sources[0].children[1].bbox, "copy") else None, resizing=False)
    twin.parameters[0].constraint = None
    rendered = iter(h["unit_rendered"])
    seen.clear()
    init.set_spectra_to_match(sources + [twin], obs)
    assert len(seen) == 3
    assert_allclose(np.asarray(twin.parameters[0]), np.asarray(sources[0].parameters[0]))


def test_function_psf_models_vs_reference_fixture():
    """MoffatPSF / GaussianPSF images (centre-sampled Moffat, pixel-integrated Gaussian, sub-pixel offsets, the single-plane
    form when all bands share a width) against the reference's own (psf.py:80-201)"""
    import scarlet_b200 as sb
    h = golden("init_helpers.npz")
    moffat = sb.MoffatPSF(alpha=[4.7, 3.0, 2.2], beta=[1.5, 2.5, 3.0], boxsize=21)
    assert moffat.bbox.shape == (3, 21, 21) and moffat.bbox.origin == (0, -10, -10)
    assert_allclose(moffat.get_model(), h["moffat"], rtol=1e-12)
    assert_allclose(moffat.get_model(offset=(0.3, -0.2)), h["moffat_offset"], rtol=1e-12)
    same = sb.MoffatPSF(alpha=[2.0, 2.0], beta=[2.0, 2.0]).get_model()
    assert same.shape == h["moffat_same"].shape == (1, 11, 11)
    assert_allclose(same, h["moffat_same"], rtol=1e-12)
    assert_allclose(sb.GaussianPSF(sigma=[0.8, 1.3], boxsize=11).get_model(offset=(0.25, -0.4)), h["gauss_offset"], rtol=1e-10)
    with pytest.raises(AssertionError):
        sb.MoffatPSF(integrate=True)
    g = golden("hsc_cosmos_35.npz")
    assert_allclose(sb.ImagePSF(g["psfs"].copy()).get_model(offset=(0.3, -0.45)), h["imagepsf_offset"], atol=1e-12)


def test_measure_moments_vs_reference_fixture():
    """measure.moments (same keys, same axis conventions, default centroid, explicit centroid, weight function, 2-D input)
    against the reference's numbers"""
    from scarlet_b200 import measure
    h = golden("init_helpers.npz")
    cube, wgt = h["mom_cube"], h["mom_weight"]
    for name, args in (("default", {}), ("centroid", dict(centroid=np.array([2.5, 4.25]))), ("weighted", dict(N=3, weight=wgt))):
        M = measure.moments(cube, **args)
        assert [tuple(k) for k in h["mom_%s_keys" % name]] == sorted(M)
        assert_allclose(np.array([M[k] for k in sorted(M)]), h["mom_%s_vals" % name], rtol=1e-12)
    M2 = measure.moments(cube[0], N=1)
    assert_allclose(np.array([M2[k] for k in sorted(M2)]), h["mom_image_vals"], rtol=1e-12)

    class Fake:
        def get_model(self):
            return cube

    assert_allclose(measure.moments(Fake(), N=0)[0, 0], cube.sum(axis=(1, 2)))


def test_detection_image_and_snr_with_masked_pixels_vs_reference_fixture():
    """build_initialization_image and measure.snr on data with a zero-weight patch: the reference lets such pixels in with
    the value underneath the mask of noise_rms (unit variance); both functions follow it (fixture from the reference)."""
    from scarlet_b200 import initialization as init, measure
    h = golden("init_helpers.npz")
    frame, obs = _hsc_observation(weights=h["weights"])
    assert (h["weights"][:, 5:9, 30:34] == 0).all()
    spectrum = init.get_pixel_spectrum(tuple(h["centers"][0]), obs)
    detect, std = init.build_initialization_image(obs, spectra=spectrum)
    assert_allclose(detect, h["detect"], rtol=1e-6, atol=1e-6 * np.abs(h["detect"]).max())
    assert_allclose(std, h["detect_std"], rtol=1e-6)
    detect, std = init.build_initialization_image(obs)
    assert_allclose(detect, h["detect_flat"], rtol=1e-6, atol=1e-6 * np.abs(h["detect_flat"]).max())
    assert_allclose(std, h["detect_flat_std"], rtol=1e-6)
    for tag in ("", "_masked"):  # Observation.render (device) replaced by the reference's rendered model
        obs.render = lambda model, rendered=h["snr_rendered" + tag]: rendered
        assert_allclose(measure.snr(np.zeros(frame.shape), obs), float(h["snr_value" + tag]), rtol=1e-5)


def test_uncentered_symmetry_vs_reference_fixture():
    """operator.prox_uncentered_symmetry(algorithm="sdss") -- minimum of 180-degree partners about the peak or an explicit
    off-centre pixel, odd and even shapes, with and without fill -- against the reference's outputs"""
    from scarlet_b200 import operator
    h = golden("init_helpers.npz")
    for i, (shape, center, fill) in enumerate(zip(h["sym_shapes"], h["sym_centers"], h["sym_fills"])):
        X = h["sym%d_in" % i].copy()
        assert X.shape == tuple(shape)
        out = operator.prox_uncentered_symmetry(X, 0, center=None if center[0] < 0 else tuple(int(v) for v in center), algorithm="sdss",
                                                fill=None if np.isnan(fill) else float(fill))
        assert_allclose(out, h["sym%d_out" % i], rtol=0, atol=0)
    with pytest.raises(NotImplementedError):
        operator.prox_uncentered_symmetry(np.ones((5, 5)), 0, algorithm="kspace")


def test_trim_morphology_vs_reference_fixture():
    """initialization.trim_morphology / get_minimal_boxsize: threshold, smallest standard box around the centre index that
    holds the support, explicit box size, centre outside the support, NaN pixels -- against the reference's outputs"""
    from scarlet_b200 import initialization as init
    h = golden("init_helpers.npz")
    for i, (cy, cx, thr, bs) in enumerate(h["trim_args"]):
        morph, box = init.trim_morphology((int(cy), int(cx)), h["trim_in"].copy(), bg_thresh=float(thr), boxsize=None if bs < 0 else int(bs))
        assert box.origin == tuple(h["trim%d_origin" % i]) and morph.shape == h["trim%d_out" % i].shape
        assert_array_equal(morph, h["trim%d_out" % i])
    assert [init.get_minimal_boxsize(s) for s in (0, 21, 22, 31, 32, 100)] == [21, 21, 31, 31, 41, 101]


def test_image_morphology_update_vs_reference_fixture():
    """ImageMorphology.update (dynamic box): shrink when the outer rings are empty, grow by linear-ramp padding
    when the next AMSGrad step pulls flux to an edge (zero second moments masked out), step halved, optimiser state cut /
    zero-padded, UpdateException raised -- new image, m, v, vhat, box and step against the reference's own update"""
    import warnings
    import scarlet_b200 as sb
    h = golden("init_helpers.npz")
    frame = sb.Frame((1, 61, 61), channels=["r"])
    for i in range(3):
        img, m, v = h["box%d_image" % i], h["box%d_m" % i], h["box%d_v" % i]
        par = sb.Parameter(img.copy(), name="image", step=1e-2, m=m.copy(), v=v.copy(), vhat=v.copy() * 2)
        morph = sb.ImageMorphology(frame, par, bbox=sb.Box((31, 31), origin=(10, 12)), resizing=True)
        changed = 0
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            try:
                morph.update()
            except sb.UpdateException:
                changed = 1
        assert changed == int(h["box%d_changed" % i])
        new = morph.parameters[0]
        assert tuple(morph.bbox.origin) == tuple(h["box%d_new_origin" % i]) and morph.bbox.shape == h["box%d_new_image" % i].shape
        assert float(new.step) == float(h["box%d_new_step" % i])
        assert_allclose(np.asarray(new), h["box%d_new_image" % i], rtol=1e-12)
        for key in ("m", "v", "vhat"):
            assert_allclose(np.asarray(getattr(new, key)), h["box%d_new_%s" % (i, key)], rtol=1e-12)
    # a fixed or non-resizing image never changes
    par = sb.Parameter(h["box0_image"].copy(), name="image", step=1e-2)
    still = sb.ImageMorphology(frame, par, bbox=sb.Box((31, 31), origin=(10, 12)), resizing=False)
    still.update()
    assert still.bbox.shape == (31, 31)


def test_psf_shaped_recipes_keep_the_data_dtype():
    """PointSource and the compact ExtendedSource (both built on the host): spectra bit-equal to the reference's INCLUDING the
    dtype -- the reference divides the float32 peak-pixel values in place, and a float32 spectrum is rounded to float32 after
    every step of the fit"""
    import scarlet_b200 as sb
    h = golden("init_helpers.npz")
    g = golden("hsc_cosmos_35.npz")
    frame, obs = _hsc_observation()
    point = sb.PointSource(frame, tuple(h["centers"][0]), obs)
    assert point.parameters[0].dtype == h["point_spectrum"].dtype == np.float32
    assert_array_equal(np.asarray(point.parameters[0]), h["point_spectrum"])
    assert_array_equal(np.asarray(point.parameters[1]), h["point_center"])
    assert_array_equal(point.parameters[0].step(point.parameters[0], it=0), h["point_spectrum_step"])
    compact = sb.ExtendedSource(frame, tuple(h["centers"][2]), obs, compact=True)
    assert compact.parameters[0].dtype == h["compact_spectrum"].dtype == np.float32
    assert_array_equal(np.asarray(compact.parameters[0]), h["compact_spectrum"])
    assert_allclose(np.asarray(compact.parameters[1]), h["compact_image"], rtol=1e-12)
    assert tuple(compact.bbox.origin) == tuple(h["compact_origin"])
    assert g["images"].dtype == np.float32
