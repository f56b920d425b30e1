"""Import shim that lets the reference's own *forward* code run in the build container.

TEST INFRASTRUCTURE ONLY.  This module is used by ``tests/golden/make_golden.py`` (in the build
container, where ``/root/reference`` is mounted read-only) to execute the reference's forward path
(PSF models, Frame, Observation.match/render/get_log_likelihood, constraints, Blend.get_model) from
its own files, so that the CPU oracle in ``oracle/scarlet_oracle.py`` can be pinned against the real
thing and golden fixtures can be written to ``tests/golden``.  Nothing here is product code and nothing
here ships reference sources: modules are imported from where they lie.

What is stubbed (the packages are absent in this image):
  * ``autograd``      -> NumPy re-exports, ``primitive`` = identity, ``defvjp`` = no-op
  * ``proxmin``       -> names only (the optimiser cannot be executed; it is restated in the oracle)
  * ``astropy``       -> empty ``astropy.wcs.WCS`` class (isinstance checks only)
  * ``scarlet.operators_pybind11`` -> the oracle's C restatement of ``prox_weighted_monotonic``
  * ``scarlet/__init__.py`` is bypassed (it imports matplotlib-dependent modules)
"""
import importlib
import os
import sys
import types

import numpy as _np
import scipy.special as _sp

REFERENCE_ROOT = os.environ.get("SCARLET_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "scarlet"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    """Register the stand-in modules and return the synthetic ``scarlet`` package."""
    if "scarlet" in sys.modules and getattr(sys.modules["scarlet"], "_is_ref_shim", False):
        return sys.modules["scarlet"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)

    # ---- autograd ------------------------------------------------------------------------
    anp = types.ModuleType("autograd.numpy")
    for k in dir(_np):
        if not k.startswith("__"):
            setattr(anp, k, getattr(_np, k))
    # numpy>=2 refuses generators in np.stack (reference psf.py uses one)
    _stack = _np.stack
    anp.stack = lambda arrays, *a, **kw: _stack(list(arrays), *a, **kw)
    anp.fft = _np.fft
    anp.random = _np.random
    anp.linalg = _np.linalg
    sys.modules["autograd.numpy"] = anp

    class _Box:  # ArrayBox stand-in
        @staticmethod
        def register(cls):
            return None

    class _VSpace:
        mappings = {_np.ndarray: None}

        @staticmethod
        def register(cls, vspace_maker=None):
            return None

    boxes = _mod("autograd.numpy.numpy_boxes", ArrayBox=_Box)
    anp.numpy_boxes = boxes
    core = _mod("autograd.core", VSpace=_VSpace)
    ext = _mod("autograd.extend", primitive=lambda f: f, defvjp=lambda *a, **k: None)
    asp_special = _mod("autograd.scipy.special", erfc=_sp.erfc, erf=_sp.erf, gamma=_sp.gamma)
    asp = _mod("autograd.scipy", special=asp_special)

    def _no_grad(*a, **k):
        raise RuntimeError("autograd is not available; gradients are restated in the oracle")

    _mod("autograd", numpy=anp, core=core, extend=ext, scipy=asp, grad=_no_grad)

    # ---- proxmin (names only) -----------------------------------------------------------------
    def _unavailable(*a, **k):
        raise RuntimeError("proxmin is not available in this image")

    pops = _mod("proxmin.operators", prox_unity_plus=_unavailable, prox_hard=_unavailable,
                prox_soft=_unavailable, prox_hard_plus=_unavailable, prox_plus=_unavailable)
    putils = _mod("proxmin.utils", l2sq=lambda x: (x ** 2).sum())
    palg = _mod("proxmin.algorithms")
    for n in ("adam", "nadam", "amsgrad", "padam", "adamx", "radam"):
        setattr(palg, "_%s_phi_psi" % n, _unavailable)
    _mod("proxmin", operators=pops, utils=putils, algorithms=palg, adaprox=_unavailable)

    # ---- astropy (isinstance only) --------------------------------------------------------------
    class WCS:  # noqa: N801
        pass

    wcs = _mod("astropy.wcs", WCS=WCS)
    _mod("astropy", wcs=wcs)

    # ---- scarlet package without its __init__ ----------------------------------------------------
    pkg = types.ModuleType("scarlet")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "scarlet")]
    pkg._is_ref_shim = True
    sys.modules["scarlet"] = pkg

    # native module: the oracle's C restatement of operators_pybind11.cc:14-36
    from oracle import monotonic_c

    def prox_weighted_monotonic(flat_img, weights, offsets, dist_idx, min_gradient):
        monotonic_c.sweep(flat_img, weights, offsets, dist_idx, min_gradient)

    def _native_unavailable(*a, **k):
        raise RuntimeError("operators_pybind11 symbol not restated (out of scope)")

    _mod("scarlet.operators_pybind11", prox_weighted_monotonic=prox_weighted_monotonic,
         apply_filter=_native_unavailable, get_valid_monotonic_pixels=_native_unavailable,
         linear_interpolate_invalid_pixels=_native_unavailable)

    for name in ("bbox", "cache", "fft", "interpolation", "constraint", "operator", "parameter", "model",
                 "psf", "frame", "renderer", "observation", "spectrum"):
        setattr(pkg, name, importlib.import_module("scarlet." + name))
    # morphology needs wavelet + initialization; import what loads, tolerate the rest
    for name in ("wavelet", "initialization", "morphology", "component", "source", "blend"):
        try:
            setattr(pkg, name, importlib.import_module("scarlet." + name))
        except Exception as exc:  # pragma: no cover - informational
            setattr(pkg, "_import_error_" + name, exc)
    return pkg
