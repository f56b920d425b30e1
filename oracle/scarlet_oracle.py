"""CPU oracle: NumPy float64 restatement of scarlet's proximal-gradient fitting path.

TEST INFRASTRUCTURE ONLY -- only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module.  The product
(``scarlet_b200``) never routes through it.

What it restates (all citations are into the reference tree, ``scarlet/``):

* box geometry ....................... ``bbox.py:279-301``
* pad / centre / FFT-shape rules ..... ``fft.py:9-36, 82-167``
* k-space convolution / PSF matching . ``fft.py:200-273, 316-396``; ``renderer.py:164-259``
* pixel-integrated Gaussian PSF ...... ``psf.py:9-17, 80-142``
* scene model (sed x morph, insertion)  ``component.py:144-171``; ``blend.py:17-46, 200-244``
* Gaussian log-likelihood ............ ``observation.py:116-186``
* gradients (hand adjoints) .......... what ``autograd.grad`` yields at ``blend.py:118``; explicit
                                       forms in ``lite/models.py:206-216, 364-367, 537-545``
* proximal operators ................. ``constraint.py:58-114, 183-234, 262-287``;
                                       ``operator.py:10-96, 274-293, 530-667``;
                                       ``operators_pybind11.cc:14-36`` (C restatement: ``monotonic.c``)
* optimiser loop ..................... ``blend.py:85-198, 276-302``; ``parameter.py:126-129``;
                                       inner update mirrored in-tree at ``lite/parameters.py:274-305``

PARITY STATUS.  The forward path, the constraints and the geometry are PINNED: against the
reference's own known-answer tests (``tests/test_constraint.py:93-163``, ``tests/test_fft.py:12-124``,
``tests/test_observation.py:12-47``, ``tests/test_component.py``, ``tests/test_bbox.py``) and against
outputs of the reference's own forward code executed in the build container through
``oracle/ref_shim.py`` (fixtures in ``tests/golden``, generator ``tests/golden/make_golden.py``).
The gradients are pinned indirectly: hand adjoints agree with finite differences of the reference
forward and with torch-CPU autograd over the same forward.  The OPTIMISER ARITHMETIC IS
"PARITY UNPINNED": ``proxmin`` (``setup.py:147``, ``proxmin>=0.6.11``, not vendored, not installable
here) supplies ``adaprox``/AMSGrad; ``adaprox_step`` below restates its published algorithm
(Reddi, Kale & Kumar 2018 AMSGrad without bias correction + proximal sub-iterations in the
``psi`` metric), cross-checked only against the in-tree mirror ``lite/parameters.py:274-305``.
No test in the reference pins a ``Blend.fit`` trajectory.

Precision follows the reference: model cube in the frame dtype (float32, ``frame.py:29``,
``blend.py:241``), spectra in the data dtype (float32), morphologies/centres float64, FFTs in
complex128 (reference-era NumPy upcast float32 input), reductions in float64.
"""
from __future__ import annotations

import math

import numpy as np
from scipy import fftpack, special

from . import monotonic_c

# Arithmetic of the restatement.  The defaults are the reference's (see "Precision" above).  ``float32_arithmetic()``
# switches the FFT convolutions to float32/complex64 and keeps the morphology images and their optimiser state in
# float32 -- an independent float32 implementation of the same algorithm, used ONLY to measure how far float32 rounding
# alone moves a trajectory (tools/parity_curve.py: the justification of the float32 tolerances in tests/).
ARITH = {"fft": np.float64, "morph": np.float64}


class float32_arithmetic:
    def __enter__(self):
        self.saved = dict(ARITH)
        ARITH.update(fft=np.float32, morph=np.float32)
        return self

    def __exit__(self, *exc):
        ARITH.update(self.saved)
        return False

# ----------------------------------------------------------------------------------------------
# geometry  (bbox.py)
# ----------------------------------------------------------------------------------------------


class OBox:
    """N-d integer box: ``shape`` and ``origin`` (minimum corner).  bbox.py:4-38."""

    def __init__(self, shape, origin=None):
        self.shape = tuple(int(s) for s in shape)
        self.origin = tuple(int(o) for o in (origin if origin is not None else (0,) * len(self.shape)))

    @property
    def stop(self):
        return tuple(o + s for o, s in zip(self.origin, self.shape))

    def __and__(self, other):
        lo = [max(a, b) for a, b in zip(self.origin, other.origin)]
        hi = [min(a, b) for a, b in zip(self.stop, other.stop)]
        return OBox([max(0, h - l) for l, h in zip(lo, hi)], lo)

    def __repr__(self):
        return "OBox(shape=%s, origin=%s)" % (self.shape, self.origin)


def overlapped_slices(b1, b2):
    """Slices into arrays bounded by ``b1`` and ``b2`` covering their intersection.  bbox.py:279-301."""
    ov = b1 & b2
    s1 = tuple(slice(o - a, o - a + s) for o, a, s in zip(ov.origin, b1.origin, ov.shape))
    s2 = tuple(slice(o - a, o - a + s) for o, a, s in zip(ov.origin, b2.origin, ov.shape))
    return s1, s2


# ----------------------------------------------------------------------------------------------
# padding / centring / FFT shapes  (fft.py)
# ----------------------------------------------------------------------------------------------


def pad_to(arr, newshape, axes=None):
    """Zero-pad ``arr`` to ``newshape``; the extra pixel of an odd difference goes in FRONT, which
    puts an odd array's centre on the centre-right pixel of an even target.  fft.py:82-113."""
    arr = np.asarray(arr)
    if axes is None:
        axes = range(arr.ndim)
    axes = [a % arr.ndim for a in (axes if hasattr(axes, "__len__") or hasattr(axes, "__iter__") else [axes])]
    out_shape = list(arr.shape)
    start = [0] * arr.ndim
    for n, ax in enumerate(axes):
        extra = int(newshape[n]) - arr.shape[ax]
        if extra < 0:
            raise ValueError("cannot pad to a smaller shape")
        start[ax] = (extra + 1) // 2
        out_shape[ax] = int(newshape[n])
    out = np.zeros(out_shape, dtype=arr.dtype)
    out[tuple(slice(s, s + n) for s, n in zip(start, arr.shape))] = arr
    return out


def centered(arr, newshape):
    """Inverse of :func:`pad_to` over all axes.  fft.py:9-36."""
    newshape = np.asarray(newshape)
    cur = np.array(arr.shape)
    if np.any(newshape > cur):
        raise ValueError("arr must be at least as large as newshape")
    start = (cur - newshape + 1) // 2
    return arr[tuple(slice(int(s), int(s + n)) for s, n in zip(start, newshape))]


def get_fft_shape(shape1, shape2, padding=3, axes=None):
    """Fast (5-smooth) FFT lengths for the linear convolution of two arrays.  fft.py:116-167.

    Sum of the sizes + ``padding`` per transformed axis -> ``next_fast_len``; the last axis is forced
    even, the second-to-last only when the second operand is even there."""
    shape1 = np.asarray(shape1)
    shape2 = np.asarray(shape2)
    if len(shape1) != len(shape2):
        raise ValueError("both operands need the same number of dimensions")
    if axes is None:
        total = shape1 + shape2
    else:
        total = np.array([shape1[a] + shape2[a] for a in axes])
    fast = [int(fftpack.next_fast_len(int(s + padding))) for s in total]
    while fast[-1] % 2:
        fast[-1] = int(fftpack.next_fast_len(fast[-1] + 1))
    if shape2[-2] % 2 == 0:
        while fast[-2] % 2:
            fast[-2] = int(fftpack.next_fast_len(fast[-2] + 1))
    return fast


def forward_fft(image, fft_shape, axes):
    """pad -> ifftshift -> rfftn over ``axes``.  fft.py:255-273.  float64/complex128 throughout."""
    image = np.asarray(image, dtype=ARITH["fft"])
    padded = pad_to(image, fft_shape, axes)
    return np.fft.rfftn(np.fft.ifftshift(padded, axes), axes=axes)


def inverse_fft(image_fft, fft_shape, image_shape, axes):
    """irfftn -> fftshift -> centre crop to ``image_shape``.  fft.py:200-243."""
    img = np.fft.irfftn(image_fft, fft_shape, axes=axes)
    img = np.fft.fftshift(img, axes=axes)
    return centered(img, image_shape)


def kspace_combine(image1, image2, padding, divide, out_shape, axes):
    """Multiply (or divide) two real images in k-space on the common fast grid.  fft.py:316-331."""
    image1 = np.asarray(image1)
    image2 = np.asarray(image2)
    if image1.ndim != image2.ndim:
        raise ValueError("both images need the same number of axes")
    fshape = get_fft_shape(image1.shape, image2.shape, padding, axes)
    a = forward_fft(image1, fshape, axes)
    b = forward_fft(image2, fshape, axes)
    return inverse_fft(a / b if divide else a * b, fshape, out_shape, axes)


def match_psf(psf1, psf2, padding=3, axes=(-2, -1)):
    """Difference kernel that maps ``psf2`` onto ``psf1`` (``psf1 = kernel * psf2``).  fft.py:334-365.
    The output takes the shape of the operand with more entries along axis 0 (ties -> ``psf1``)."""
    psf1 = np.asarray(psf1)
    psf2 = np.asarray(psf2)
    shape = psf2.shape if psf1.shape[0] < psf2.shape[0] else psf1.shape
    return kspace_combine(psf1, psf2, padding, True, shape, axes)


def convolve(image, kernel, padding=3, axes=(-2, -1)):
    """FFT convolution, output cropped to ``image.shape``.  fft.py:368-396."""
    return kspace_combine(image, kernel, padding, False, np.asarray(image).shape, axes)


# ----------------------------------------------------------------------------------------------
# PSF models  (psf.py)
# ----------------------------------------------------------------------------------------------


def normalize_psf(image):
    """Unit sum per band.  psf.py:9-17."""
    return image / image.sum(axis=(1, 2))[:, None, None]


def gaussian_pixel_integral(x, sigma):
    """Integral of exp(-t^2/2 sigma^2) over the pixel [x-1/2, x+1/2].  psf.py:129-142."""
    s2 = math.sqrt(2.0) * sigma
    return math.sqrt(math.pi / 2) * sigma * (
        1 - special.erfc((0.5 - x) / s2) + 1 - special.erfc((2 * x + 1) / (2 * s2)))


def gaussian_pixel_integral_deriv(x, sigma):
    """d/dx of :func:`gaussian_pixel_integral`."""
    return np.exp(-((x + 0.5) ** 2) / (2 * sigma ** 2)) - np.exp(-((x - 0.5) ** 2) / (2 * sigma ** 2))


class GaussianPSFOracle:
    """Pixel-integrated circular Gaussian per band on an odd box centred on the origin.
    psf.py:39-142.  ``get_model`` returns (1,b,b) when all sigmas are equal, else (C,b,b)."""

    def __init__(self, sigma, boxsize=None):
        self.sigma = np.atleast_1d(np.asarray(sigma, dtype=np.float64))
        if boxsize is None:
            boxsize = int(np.ceil(10 * np.max(self.sigma)))
        if boxsize % 2 == 0:
            boxsize += 1
        self.boxsize = int(boxsize)
        self.coords = np.arange(self.boxsize) - self.boxsize // 2
        self.is_same = bool(np.all(self.sigma == self.sigma[0]))
        self.bbox = OBox((len(self.sigma), self.boxsize, self.boxsize), (0, -(self.boxsize // 2), -(self.boxsize // 2)))

    def get_model(self, offset=None):
        oy, ox = (0.0, 0.0) if offset is None else (float(offset[0]), float(offset[1]))
        sig = self.sigma[:1] if self.is_same else self.sigma
        planes = [gaussian_pixel_integral(self.coords - oy, s)[:, None]
                  * gaussian_pixel_integral(self.coords - ox, s)[None, :] for s in sig]
        return normalize_psf(np.stack(planes, axis=0))


class ImagePSFOracle:
    """PSF given as an image (2-D or per band), normalised on construction.  psf.py:205-234."""

    def __init__(self, image):
        image = np.array(image, dtype=np.float64)
        if image.ndim == 2:
            image = image[None]
        self.image = normalize_psf(image)
        self.bbox = OBox(self.image.shape, (0, -(image.shape[1] // 2), -(image.shape[2] // 2)))

    def get_model(self, offset=None):
        if offset is not None:
            raise NotImplementedError("Fourier-shifted ImagePSF is outside the oracle's scope (SURVEY f-3)")
        return self.image.copy()


# ----------------------------------------------------------------------------------------------
# proximal operators  (constraint.py, operator.py, operators_pybind11.cc)
# ----------------------------------------------------------------------------------------------

NEIGHBOURS = ((-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1))


def radial_order(shape, center):
    """Flat pixel indices sorted by distance from ``center``.  operator.py:10-48."""
    cy, cx = int(center[0]), int(center[1])
    yy, xx = np.meshgrid(np.arange(shape[0]) - cy, np.arange(shape[1]) - cx, indexing="ij")
    return np.argsort(np.sqrt(xx ** 2 + yy ** 2).ravel(), kind="stable")


def monotonic_weights(shape, neighbor_weight="flat", center=None):
    """(8, H*W) neighbour weights of the radial monotonicity operator.  operator.py:591-667.

    A neighbour counts iff it is inside the image and STRICTLY closer to the centre; ``angle``
    weighs it by the cosine between the direction to the centre and the direction to the neighbour
    (normalised to unit sum), ``flat`` weighs all such neighbours equally, ``nearest`` keeps only the
    best aligned one."""
    if neighbor_weight not in ("flat", "angle", "nearest"):
        raise ValueError(neighbor_weight)
    H, W = int(shape[0]), int(shape[1])
    if center is None:
        center = ((H - 1) // 2, (W - 1) // 2)
    cy, cx = int(center[0]), int(center[1])
    yy, xx = np.meshgrid(np.arange(H) - cy, np.arange(W) - cx, indexing="ij")
    d2 = xx ** 2 + yy ** 2
    toward = np.arctan2((-yy).astype(np.float64), (-xx).astype(np.float64))  # direction pixel -> centre (integer negation: no -0.0)
    # the reference evaluates the centre pixel's own angle as arctan2(0,-1); it has no valid neighbour
    cosw = np.zeros((8, H, W))
    for n, (dy, dx) in enumerate(NEIGHBOURS):
        ny, nx = np.arange(H)[:, None] + dy, np.arange(W)[None, :] + dx
        inside = (ny >= 0) & (ny < H) & (nx >= 0) & (nx < W)
        nd2 = (xx + dx) ** 2 + (yy + dy) ** 2
        valid = inside & (nd2 < d2)
        cosw[n] = np.where(valid, np.cos(toward - math.atan2(dy, dx)), 0.0)
    cosw = cosw.reshape(8, H * W)
    if neighbor_weight == "nearest":
        best = np.argmax(cosw, axis=0)
        out = np.zeros_like(cosw)
        out[best, np.arange(H * W)] = 1
        out[:, cy * W + cx] = 0
        return out
    if neighbor_weight == "flat":
        cosw[cosw != 0] = 1
    norm = cosw.sum(axis=0)
    norm[norm == 0] = 1
    return cosw / norm[None, :]


_MONO_CACHE = {}


def prox_monotonic(X, neighbor_weight="flat", min_gradient=0.1, center=None):
    """Radial monotonicity projection of a 2-D image, in place.  constraint.py:203-234 with
    ``use_mask=False``, ``fit_center_radius=0``; sweep = monotonic.c (operators_pybind11.cc:14-36)."""
    shape = X.shape
    if center is None:
        center = (shape[0] // 2, shape[1] // 2)
    key = (shape, tuple(center), neighbor_weight)
    if key not in _MONO_CACHE:
        w = monotonic_weights(shape, neighbor_weight, center)
        order = radial_order(shape, center)[1:].astype(np.int32)
        offs = np.array([shape[1] * dy + dx for dy, dx in NEIGHBOURS], dtype=np.int32)
        _MONO_CACHE[key] = (w, order, offs)
    w, order, offs = _MONO_CACHE[key]
    if not (X.flags.c_contiguous and X.dtype in (np.float32, np.float64)):
        raise TypeError("prox_monotonic needs a C-contiguous float array")
    monotonic_c.sweep(X.reshape(-1), w, offs, order, min_gradient)
    return X


def prox_monotonic_python(X, neighbor_weight="flat", min_gradient=0.1, center=None):
    """Pure-Python version of the sweep (small cases only; validates the C file)."""
    shape = X.shape
    if center is None:
        center = (shape[0] // 2, shape[1] // 2)
    w = monotonic_weights(shape, neighbor_weight, center)
    order = radial_order(shape, center)[1:]
    offs = [shape[1] * dy + dx for dy, dx in NEIGHBOURS]
    flat = X.reshape(-1)
    for p in order:
        ref = 0.0
        for i in range(8):
            if w[i, p] > 0:
                ref += flat[p + offs[i]] * w[i, p]
        flat[p] = min(flat[p], ref * (1 - min_gradient))
    return X


def prox_symmetry(X, strength=1.0):
    """Blend with the 180-degree rotation; even axes are extended by one zero row/column first.
    operator.py:274-293."""
    H, W = X.shape
    ext = np.zeros((H + (H % 2 == 0), W + (W % 2 == 0)), dtype=X.dtype)
    ext[:H, :W] = X
    rot = ext[::-1, ::-1]
    out = 0.5 * strength * (ext + rot) + (1 - strength) * ext
    return out[:H, :W]


def prox_positivity(X, zero=0.0):
    """constraint.py:83-92."""
    return np.maximum(X, zero)


def prox_center_on(X, tiny=1e-6):
    """Floor the centre pixel (shape//2), in place.  constraint.py:276-287."""
    c = (X.shape[0] // 2, X.shape[1] // 2)
    X[c] = max(X[c], tiny)
    return X


def prox_normalization(X, kind="sum"):
    """Divide by the sum or the maximum, in place.  constraint.py:95-114."""
    X /= X.sum() if kind == "sum" else X.max()
    return X


class ChainSpec:
    """An ordered list of constraint op descriptors, e.g.
    ``[("monotonic", "angle", 0.0), ("symmetry", 1.0), ("positivity", 0.0), ("center_on", 1e-6),
    ("normalization", "max")]`` applied ``repeat`` times.  constraint.py:58-80."""

    def __init__(self, ops, repeat=1):
        self.ops = list(ops)
        self.repeat = int(repeat)

    def __call__(self, X, step=None):
        for _ in range(self.repeat):
            for op in self.ops:
                kind = op[0]
                if kind == "monotonic":
                    X = prox_monotonic(np.ascontiguousarray(X), op[1], op[2])
                elif kind == "symmetry":
                    X = prox_symmetry(X, op[1])
                elif kind == "positivity":
                    X = prox_positivity(X, op[1])
                elif kind == "center_on":
                    X = prox_center_on(X, op[1])
                elif kind == "normalization":
                    X = prox_normalization(X, op[1])
                else:
                    raise TypeError("unknown constraint op %r" % (kind,))
        return X


def extended_source_chain(monotonic="angle", symmetric=False, min_grad=0.0):
    """Constraint chain of ``ExtendedSourceMorphology``.  morphology.py:644-669."""
    ops = []
    if monotonic:
        ops.append(("monotonic", "angle" if monotonic is True else monotonic, float(min_grad)))
    if symmetric:
        ops.append(("symmetry", 1.0))
    ops += [("positivity", 0.0), ("center_on", 1e-6), ("normalization", "max")]
    return ChainSpec(ops)


# ----------------------------------------------------------------------------------------------
# parameters, sources
# ----------------------------------------------------------------------------------------------


class OParam:
    """Optimisation parameter with AMSGrad state (parameter.py:9-84)."""

    def __init__(self, value, name, step, prox=None, fixed=False):
        self.x = np.array(value)
        self.name = name
        self.step = step  # float, ndarray, or callable(x, it)
        self.prox = prox
        self.fixed = fixed
        self.m = self.v = self.vhat = None
        self.std = None

    def step_size(self, it):
        return self.step(self.x, it) if callable(self.step) else self.step


def relative_step(x, it, factor=0.1, minimum=0.0):
    """parameter.py:126-129."""
    return np.maximum(minimum, factor * x.mean())


class ExtendedSourceOracle:
    """``FactorizedComponent(TabulatedSpectrum, ExtendedSourceMorphology)`` with ``shifting=False``,
    ``resizing=False``.  source.py:367-450, morphology.py:607-688, spectrum.py:32-71.

    Parameters in reference order: spectrum, image, shift (the free, unused ``shift`` of
    morphology.py:112-113 is carried for index parity; its gradient is identically zero)."""

    kind = "extended"

    def __init__(self, sed, morph, origin, min_step=0.0, monotonic="angle", symmetric=False, min_grad=0.0,
                 sed_dtype=np.float32, resizing=False, shift=None):
        self.resizing = resizing
        self.shifting = shift is not None
        sed = np.asarray(sed, dtype=sed_dtype)
        morph = np.array(morph, dtype=ARITH["morph"])
        self.min_step = np.asarray(min_step, dtype=np.float64)
        self.spectrum = OParam(sed, "spectrum", lambda x, it: relative_step(x, it, 1e-2, self.min_step),
                               ChainSpec([("positivity", 1e-20)]))
        self.image = OParam(morph, "image", 1e-2, extended_source_chain(monotonic, symmetric, min_grad))
        # morphology.py:112-113 (free, unused, step 1e-2) or, with shifting=True, morphology.py:672-675 (step 1e-1)
        self.shift = OParam(np.zeros(2), "shift", 1e-2, None) if shift is None else OParam(np.array(shift, dtype=np.float64), "shift", 1e-1, None)
        C = sed.shape[0]
        self.bbox = OBox((C,) + morph.shape, (0, int(origin[0]), int(origin[1])))

    @property
    def parameters(self):
        return (self.spectrum, self.image, self.shift)

    def get_model(self, values=None):
        sed, morph, shift = (self.spectrum.x, self.image.x, self.shift.x) if values is None else values[:3]
        if self.shifting:
            morph = fourier_shift(morph, shift)
        return sed[:, None, None] * morph[None, :, :]

    def param_grads(self, gbox, values=None):
        sed, morph, shift = (self.spectrum.x, self.image.x, self.shift.x) if values is None else values[:3]
        g_shifted = np.einsum("c,cyx->yx", np.asarray(sed, dtype=np.float64), gbox)
        if not self.shifting:
            return (np.einsum("cyx,yx->c", gbox, morph), g_shifted, np.zeros(2))
        shifted, g_morph, g_shift = fourier_shift_vjp(morph, shift, g_shifted)
        return (np.einsum("cyx,yx->c", gbox, shifted), g_morph, g_shift)

    def _new_image(self, data, m, v, vhat, origin_shift):
        old = self.image
        self.image = OParam(np.array(data, dtype=ARITH["morph"]), "image", old.step / 2, old.prox, old.fixed)
        self.image.m, self.image.v, self.image.vhat = np.array(m), np.array(v), np.array(vhat)
        C = self.bbox.shape[0]
        oy, ox = self.bbox.origin[1] + origin_shift, self.bbox.origin[2] + origin_shift
        self.bbox = OBox((C,) + self.image.x.shape, (0, oy, ox))

    def update(self):
        """Dynamic box of ``ImageMorphology.update`` / ``shrink_box`` (morphology.py:52-68, 132-207).  Returns True
        when the box changed (the reference raises ``UpdateException``)."""
        image = self.image
        if not self.resizing or image.fixed:
            return False
        x = image.x
        size = max(x.shape)
        dist = 0
        while (np.all(x[dist, :] <= 0) and np.all(x[-dist - 1, :] <= 0) and np.all(x[:, dist] <= 0) and np.all(x[:, -dist - 1] <= 0)):
            dist += 1
        newsize = minimal_boxsize(size - 2 * dist)
        if newsize < size:
            d = (size - newsize) // 2
            sl = (slice(d, d + newsize), slice(d, d + newsize))
            self._new_image(x[sl], image.m[sl], image.v[sl], image.vhat[sl], d)
            return True
        if image.m is not None:
            gu = -image.m / np.sqrt(np.sqrt(np.ma.masked_equal(image.v, 0))) * image.step
            pull = gu * (x > 0)
            edge = np.array((pull[:, 0].mean(), pull[:, -1].mean(), pull[0, :].mean(), pull[-1, :].mean()))
            if np.any(edge > 0.1):
                newsize = minimal_boxsize(size + 1)
                pad = (newsize - size) // 2
                self._new_image(np.pad(x, pad, mode="linear_ramp"), np.pad(image.m, pad, mode="constant"),
                                np.pad(image.v, pad, mode="constant"), np.pad(image.vhat, pad, mode="constant"), -pad)
                return True
        return False


def fourier_shift(image, shift, padding=10):
    """``fft.shift(image, shift, return_Fourier=False)`` (fft.py:399-428), literally: centre-pad to the fast shape of
    (image, image, padding 10), ifftshift, rfftn, multiply by exp(-2 pi i (fftfreq_y s0 + rfftfreq_x s1)), irfftn,
    fftshift, centre-crop."""
    image = np.asarray(image, dtype=np.float64)
    fshape = get_fft_shape(image.shape, image.shape, padding, (0, 1))
    spec = forward_fft(image, fshape, (0, 1))
    ramp = np.exp(-2j * np.pi * np.fft.fftfreq(fshape[0]) * shift[0])[:, None] * np.exp(-2j * np.pi * np.fft.rfftfreq(fshape[1]) * shift[1])[None, :]
    return np.real(inverse_fft(spec * ramp, fshape, image.shape, (0, 1)))


def fourier_shift_vjp(image, shift, g_out):
    """(shifted image, d<g_out, shifted>/d image, d<g_out, shifted>/d shift) by torch double-precision autograd through
    the same pipeline as ``fourier_shift`` (torch.fft follows numpy's rfftn / irfftn semantics, including what the
    real inverse transform does with the non-Hermitian Nyquist row that a fractional shift along y creates)."""
    import torch
    img = torch.tensor(np.asarray(image, dtype=np.float64), requires_grad=True)
    sh = torch.tensor(np.asarray(shift, dtype=np.float64), requires_grad=True)
    By, Bx = img.shape
    Fy, Fx = get_fft_shape((By, Bx), (By, Bx), 10, (0, 1))
    oy, ox = (Fy - By + 1) // 2, (Fx - Bx + 1) // 2
    pad = torch.zeros((Fy, Fx), dtype=torch.float64)
    pad = torch.nn.functional.pad(img, (ox, Fx - Bx - ox, oy, Fy - By - oy))
    spec = torch.fft.rfftn(torch.fft.ifftshift(pad))
    fy = torch.tensor(np.fft.fftfreq(Fy))
    fx = torch.tensor(np.fft.rfftfreq(Fx))
    ramp = torch.exp(-2j * np.pi * fy * sh[0])[:, None] * torch.exp(-2j * np.pi * fx * sh[1])[None, :]
    out = torch.fft.fftshift(torch.fft.irfftn(spec * ramp, s=(Fy, Fx)))[oy:oy + By, ox:ox + Bx]
    (out * torch.tensor(np.asarray(g_out, dtype=np.float64))).sum().backward()
    return out.detach().numpy(), img.grad.numpy(), sh.grad.numpy()


def minimal_boxsize(size, min_size=21, increment=10):
    """initialization.py:173-177."""
    boxsize = min_size
    while boxsize < size:
        boxsize += increment
    return boxsize


class PointSourceOracle:
    """``FactorizedComponent(TabulatedSpectrum, PointSourceMorphology)`` for a Gaussian model PSF.
    source.py:92-128, morphology.py:476-513.  Parameters: spectrum, center."""

    kind = "point"

    def __init__(self, sed, center, model_psf, min_step=0.0, sed_dtype=np.float32):
        assert isinstance(model_psf, GaussianPSFOracle)
        self.psf = model_psf
        sed = np.asarray(sed, dtype=sed_dtype)
        self.min_step = np.asarray(min_step, dtype=np.float64)
        self.spectrum = OParam(sed, "spectrum", lambda x, it: relative_step(x, it, 1e-2, self.min_step),
                               ChainSpec([("positivity", 1e-20)]))
        center = np.array(center, dtype=np.float64)
        self.center = OParam(center, "center", 3e-2, None)
        pix = np.round(center).astype(int)
        b = model_psf.boxsize
        self.bbox = OBox((sed.shape[0], b, b), (0, int(pix[0]) - b // 2, int(pix[1]) - b // 2))
        # morphology.py:505: mean of the (start, stop) bounds, i.e. origin + size/2
        self.box_center = np.array([self.bbox.origin[1] + b / 2, self.bbox.origin[2] + b / 2])

    @property
    def parameters(self):
        return (self.spectrum, self.center)

    def morph(self, center=None):
        center = self.center.x if center is None else center
        return self.psf.get_model(offset=center - self.box_center)

    def get_model(self, values=None):
        sed, center = (self.spectrum.x, self.center.x) if values is None else values[:2]
        return sed[:, None, None] * self.morph(center)

    def param_grads(self, gbox, values=None):
        sed, center = (self.spectrum.x, self.center.x) if values is None else values[:2]
        off = center - self.box_center
        sed64 = np.asarray(sed, dtype=np.float64)
        sig = self.psf.sigma[:1] if self.psf.is_same else self.psf.sigma
        Y = self.psf.coords
        g_sed = np.zeros(len(sed64))
        g_center = np.zeros(2)
        for k, s in enumerate(sig):
            chans = range(len(sed64)) if self.psf.is_same else [k]
            fy = gaussian_pixel_integral(Y - off[0], s)
            fx = gaussian_pixel_integral(Y - off[1], s)
            dfy = -gaussian_pixel_integral_deriv(Y - off[0], s)  # d fy / d center_y
            dfx = -gaussian_pixel_integral_deriv(Y - off[1], s)
            Sy, Sx = fy.sum(), fx.sum()
            ny, nx = fy / Sy, fx / Sx
            dny = dfy / Sy - fy * dfy.sum() / Sy ** 2
            dnx = dfx / Sx - fx * dfx.sum() / Sx ** 2
            M = ny[:, None] * nx[None, :]
            Gm = np.zeros_like(M)
            for c in chans:
                g_sed[c] = (gbox[c] * M).sum()
                Gm += sed64[c] * gbox[c]
            g_center[0] += (Gm * (dny[:, None] * nx[None, :])).sum()
            g_center[1] += (Gm * (ny[:, None] * dnx[None, :])).sum()
        return (g_sed, g_center)


# ----------------------------------------------------------------------------------------------
# observation / renderer
# ----------------------------------------------------------------------------------------------


class ObservationOracle:
    """Data cube + weights + PSF, matched to a model frame with the FFT ``ConvolutionRenderer``
    (observation.py:9-186, renderer.py:164-259).  ``channel_offset`` is the start of this
    observation's channels inside the model frame (channel map = slice, renderer.py:26-51);
    ``origin`` the (y,x) position of the data's pixel (0,0) in the model frame (pure translation)."""

    def __init__(self, data, weights, psf, frame_dtype=np.float32, channel_offset=0, origin=(0, 0), psf_shift=None):
        # renderer parameter of ConvolutionRenderer(psf_shift=...) (renderer.py:172-177): a free (dy, dx), step 1e-2, no constraint
        self.psf_shift = None if psf_shift is None else OParam(np.array(psf_shift, dtype=np.float64), "psf_shift", 1e-2, None)
        self.data = np.asarray(data, dtype=frame_dtype)
        self.weights = (np.ones(self.data.shape, dtype=frame_dtype) if weights is None
                        else np.asarray(weights, dtype=frame_dtype))
        assert self.data.shape == self.weights.shape
        self.psf = psf
        self.channel_offset = int(channel_offset)
        self.origin = (int(origin[0]), int(origin[1]))
        self.frame_dtype = frame_dtype
        self.diff_kernel = None
        self._khat = {}

    # -- setup ------------------------------------------------------------------------------
    def match(self, frame_shape, model_psf, padding=10):
        """renderer.py:165-202: diff kernel = IFFT(FFT(psf_obs)/FFT(psf_model)) with padding 10,
        both PSF images cast to the frame dtype first; overlap slices between the data box and the
        model frame."""
        self.frame_shape = tuple(frame_shape)
        C, H, W = self.data.shape
        if model_psf is None or self.psf is None:
            self.diff_kernel = None
        else:
            p_obs = self.psf.get_model().astype(self.frame_dtype)
            p_mod = model_psf.get_model().astype(self.frame_dtype)
            self.diff_kernel = match_psf(p_obs, p_mod, padding=padding)
        data_box = OBox((C, H, W), (0,) + self.origin)  # renderer.py:186-195 (channel axis taken from the frame)
        frame_box = OBox((C,) + self.frame_shape[1:], (0, 0, 0))
        self.data_slices, self.model_slices = overlapped_slices(data_box, frame_box)
        return self

    @property
    def parameters(self):
        return () if self.psf_shift is None else (self.psf_shift,)

    def shifted_kernel(self, shift=None):
        """renderer.py:220-227: the difference kernel moved by ``psf_shift`` with ``fft.shift`` (per band, axes (-2, -1))."""
        sh = self.psf_shift.x if shift is None else shift
        return np.stack([fourier_shift(k, sh) for k in self.diff_kernel])

    def _kernel_fft(self, fshape):
        if self.psf_shift is not None:  # the kernel changes with the parameter: transformed on every call (as fft.convolve does)
            return forward_fft(self.shifted_kernel(), fshape, (1, 2))
        key = tuple(fshape)
        key = key + (ARITH["fft"],)
        if key not in self._khat:
            saved, ARITH["fft"] = ARITH["fft"], np.float64  # K^ is always formed in double, then stored in the working precision
            try:
                k = forward_fft(self.diff_kernel, fshape, (1, 2))
            finally:
                ARITH["fft"] = saved
            self._khat[key] = k.astype(np.complex64) if saved == np.float32 else k
        return self._khat[key]

    # -- forward ------------------------------------------------------------------------------
    def render(self, model):
        """map channels -> convolve -> match shape.  renderer.py:247-259, 130-146."""
        C = self.data.shape[0]
        sub = model[self.channel_offset:self.channel_offset + C]
        if self.diff_kernel is not None:
            fshape = get_fft_shape(sub.shape, self.diff_kernel.shape, 3, (1, 2))
            conv = inverse_fft(forward_fft(sub, fshape, (1, 2)) * self._kernel_fft(fshape), fshape, sub.shape, (1, 2))
        else:
            conv = sub
        ds, ms = self.data_slices, self.model_slices
        if any(ds[d].stop - ds[d].start != self.data.shape[d] for d in (1, 2)):
            out = np.zeros(self.data.shape, dtype=self.frame_dtype)
            out[ds] = conv[ms]
            return out
        return conv[ms]

    def render_adjoint(self, grad_render):
        """Vector-Jacobian product of :meth:`render`: same FFT pipeline with conj(K^)."""
        C = self.data.shape[0]
        sub_shape = (C,) + self.frame_shape[1:]
        emb = np.zeros(sub_shape, dtype=ARITH["fft"])
        emb[self.model_slices] = grad_render[self.data_slices]
        if self.diff_kernel is not None:
            fshape = get_fft_shape(sub_shape, self.diff_kernel.shape, 3, (1, 2))
            emb = inverse_fft(forward_fft(emb, fshape, (1, 2)) * np.conj(self._kernel_fft(fshape)), fshape,
                              sub_shape, (1, 2))
        g = np.zeros(self.frame_shape, dtype=np.float64)
        g[self.channel_offset:self.channel_offset + C] = emb
        return g

    # -- likelihood ----------------------------------------------------------------------------
    @property
    def noise_rms(self):
        """1/sqrt(w) with w==0 masked.  observation.py:116-124."""
        w = self.weights
        return np.ma.masked_array(1 / np.sqrt(np.where(w == 0, 1, w)), mask=(w == 0))

    @property
    def log_norm(self):
        """observation.py:172-186."""
        w = self.weights.astype(np.float64)
        good = w != 0
        return good.sum() / 2 * np.log(2 * np.pi) + np.log(1 / np.sqrt(w[good])).sum()

    def channel_noise_rms(self):
        """Per-band mean noise rms used as the spectrum's minimum step.  source.py:413-415."""
        return np.array(np.mean(self.noise_rms, axis=(1, 2)))

    def neg_log_likelihood(self, model, want_grad=False):
        rendered = self.render(model)
        diff = rendered.astype(np.float64) - self.data
        w = self.weights
        nll = self.log_norm + np.sum(w * diff ** 2) / 2
        if not want_grad:
            return nll
        return nll, self.render_adjoint(w * diff)

    def param_grads(self, model):
        """Gradient of the negative log-likelihood wrt the renderer parameters (``psf_shift``): the render is linear in the
        shifted kernel K_s -- rendered[y, x] = sum_uv K_s[u, v] M[y - u + P//2, x - v + P//2] ("same" convolution) -- so
        dL/dK_s is the correlation of the residual with the model, pulled back to the shift through ``fourier_shift``."""
        if self.psf_shift is None:
            return ()
        from scipy import signal
        C = self.data.shape[0]
        sub = np.asarray(model[self.channel_offset:self.channel_offset + C], dtype=np.float64)
        diff = self.render(model).astype(np.float64) - self.data
        r = np.zeros(sub.shape)
        r[self.model_slices] = (self.weights * diff)[self.data_slices]
        P = self.diff_kernel.shape[-1]
        g = np.zeros(2)
        for c in range(C):
            mp = np.pad(sub[c], P // 2)
            gk = np.ascontiguousarray(signal.correlate(mp, r[c], mode="valid", method="fft")[::-1, ::-1])
            g += fourier_shift_vjp(self.diff_kernel[c], self.psf_shift.x, gk)[2]
        return (g,)


class ResolutionObservationOracle(ObservationOracle):
    """Observation on a coarser, aligned pixel grid rendered with the reference's ``ResolutionRenderer`` algorithm
    (renderer.py:262-547), restated literally: the padded difference kernel Fourier-shifted to every low-resolution
    row is tabulated once (``_resconv_op``, 352-363, ``sinc_shift`` 414-476); a render Fourier-shifts the padded model
    to every low-resolution column and contracts with that table (478-547).  The set-up products (padded difference
    kernel, ``shifts``, ``h``, ``_fft_shape``) are inputs: they are pinned against the reference's own set-up in
    tests/test_host_api.py::test_multiresolution_setup_vs_reference_fixture.  The adjoint is the transpose of the
    same two linear maps (a real Fourier shift by ``s`` transposes to the shift by ``-s``)."""

    def __init__(self, data, weights, diff_kernel_padded, shifts, h, frame_dtype=np.float32, channel_offset=0):
        super().__init__(data, weights, None, frame_dtype=frame_dtype, channel_offset=channel_offset)
        self.kernel = np.asarray(diff_kernel_padded, dtype=np.float64)  # (C, Fy, Fx), centred in the grid
        self.fshape = self.kernel.shape[1:]
        self.shifts = np.asarray(shifts, dtype=np.float64)
        self.h = float(h)
        # rows of the operator: kernel shifted along y to every low-resolution row, times h^2
        self.op = self.h ** 2 * np.stack([self._shift(self.kernel, s, axis=1) for s in self.shifts[0]], axis=1)  # (C, ny, Fy, Fx)

    @staticmethod
    def _shift(arr, s, axis):
        """Fourier shift by ``s`` pixels along ``axis`` with a real transform (mk_shifter(real=True) + irfft)."""
        n = arr.shape[axis]
        spec = np.fft.rfft(arr, axis=axis)
        shape = [1] * arr.ndim
        shape[axis] = spec.shape[axis]
        ramp = np.exp(-2j * np.pi * np.fft.rfftfreq(n) * s).reshape(shape)
        return np.fft.irfft(spec * ramp, n=n, axis=axis)

    def match(self, frame_shape, model_psf, padding=10):
        self.frame_shape = tuple(frame_shape)
        Fy, Fx = self.fshape
        Ny, Nx = self.frame_shape[1:]
        self.pad0 = ((Fy - Ny + 1) // 2, (Fx - Nx + 1) // 2)  # fft._pad: centre-right rule
        return self

    def render(self, model):
        C = self.data.shape[0]
        sub = np.asarray(model[self.channel_offset:self.channel_offset + C], dtype=np.float64)
        Fy, Fx = self.fshape
        Ny, Nx = self.frame_shape[1:]
        padded = np.zeros((C, Fy, Fx))
        padded[:, self.pad0[0]:self.pad0[0] + Ny, self.pad0[1]:self.pad0[1] + Nx] = sub
        # model shifted along x by -xs_j for every low-resolution column j: (C, nx, Fy, Fx)
        conv = np.stack([self._shift(padded, -s, axis=2) for s in self.shifts[1]], axis=1)
        return np.einsum("ciyx,cjyx->cij", self.op, conv).astype(self.frame_dtype)

    def render_adjoint(self, grad_render):
        C = self.data.shape[0]
        Ny, Nx = self.frame_shape[1:]
        g = np.asarray(grad_render, dtype=np.float64)
        t = np.einsum("cij,ciyx->cjyx", g, self.op)  # d/d conv
        padded = sum(self._shift(t[:, j], s, axis=2) for j, s in enumerate(self.shifts[1]))
        out = np.zeros(self.frame_shape, dtype=np.float64)
        out[self.channel_offset:self.channel_offset + C] = padded[:, self.pad0[0]:self.pad0[0] + Ny, self.pad0[1]:self.pad0[1] + Nx]
        return out


class RotatedResolutionObservationOracle(ResolutionObservationOracle):
    """The rotated branch of ``ResolutionRenderer`` (renderer.py:318-363, 498-524): the observation's pixel grid is turned by
    an angle against the model frame, so every low-resolution row needs the kernel shifted along BOTH axes (``shifts`` =
    (Y cos, -Y sin) per row) and every column the model shifted along both axes (``other_shifts`` = (X sin, X cos)); the
    render is again the contraction of the two tables.  Restated literally: ``sinc_shift`` with axes (1, 2) is an ``rfftn``
    over (y, x), the phase ramp ``exp(shifter_y s0) exp(shifter_x s1)`` with ``mk_shifter(real=False)`` (full frequencies
    along y, half along x), and ``irfftn`` back (renderer.py:414-476); the centre shifts of ``Fourier.fft`` / ``from_fft``
    cancel in the contraction.  ``small_axis`` (data_frame.Nx <= data_frame.Ny) decides which table carries which index."""

    def __init__(self, data, weights, diff_kernel_padded, shifts, other_shifts, h, small_axis=True, frame_dtype=np.float32,
                 channel_offset=0):
        ObservationOracle.__init__(self, data, weights, None, frame_dtype=frame_dtype, channel_offset=channel_offset)
        self.kernel = np.asarray(diff_kernel_padded, dtype=np.float64)
        self.fshape = self.kernel.shape[1:]
        self.shifts = np.asarray(shifts, dtype=np.float64)
        self.other_shifts = np.asarray(other_shifts, dtype=np.float64)
        self.h, self.small_axis = float(h), bool(small_axis)
        self.op = self.h ** 2 * np.stack([self._shift2(self.kernel, s0, s1) for s0, s1 in self.shifts.T], axis=1)  # (C, nq, Fy, Fx)

    @staticmethod
    def _shift2(arr, s0, s1):
        Fy, Fx = arr.shape[-2:]
        spec = np.fft.rfftn(arr, axes=(-2, -1))
        ramp = np.exp(-2j * np.pi * np.fft.fftfreq(Fy) * s0)[:, None] * np.exp(-2j * np.pi * np.fft.rfftfreq(Fx) * s1)[None, :]
        return np.fft.irfftn(spec * ramp, s=(Fy, Fx), axes=(-2, -1))

    def render(self, model):
        C = self.data.shape[0]
        sub = np.asarray(model[self.channel_offset:self.channel_offset + C], dtype=np.float64)
        Fy, Fx = self.fshape
        Ny, Nx = self.frame_shape[1:]
        padded = np.zeros((C, Fy, Fx))
        padded[:, self.pad0[0]:self.pad0[0] + Ny, self.pad0[1]:self.pad0[1] + Nx] = sub
        conv = np.stack([self._shift2(padded, -s0, -s1) for s0, s1 in self.other_shifts.T], axis=1)  # (C, nr, Fy, Fx)
        out = np.einsum("cqyx,cryx->cqr", self.op, conv)
        return (out if self.small_axis else out.transpose(0, 2, 1)).astype(self.frame_dtype)

    def render_adjoint(self, grad_render):
        """transpose of ``render``: a shift by (s0, s1) with these transform semantics transposes to the shift by (-s0, -s1)"""
        C = self.data.shape[0]
        Ny, Nx = self.frame_shape[1:]
        g = np.asarray(grad_render, dtype=np.float64)
        g = g if self.small_axis else g.transpose(0, 2, 1)
        t = np.einsum("cqr,cqyx->cryx", g, self.op)
        padded = sum(self._shift2(t[:, r], s0, s1) for r, (s0, s1) in enumerate(self.other_shifts.T))
        out = np.zeros(self.frame_shape, dtype=np.float64)
        out[self.channel_offset:self.channel_offset + C] = padded[:, self.pad0[0]:self.pad0[0] + Ny, self.pad0[1]:self.pad0[1] + Nx]
        return out


# ----------------------------------------------------------------------------------------------
# the blend
# ----------------------------------------------------------------------------------------------


class ArithmeticErrorNonFinite(ArithmeticError):
    pass


class SceneOracle:
    """``Blend(sources, observations)`` restated.  blend.py:49-308."""

    def __init__(self, frame_shape, model_psf, sources, observations, frame_dtype=np.float32):
        self.frame_shape = tuple(int(s) for s in frame_shape)
        self.model_psf = model_psf
        self.sources = list(sources)
        self.observations = list(observations)
        self.frame_dtype = frame_dtype
        self.frame_box = OBox(self.frame_shape)
        self.loss = []
        for obs in self.observations:
            if getattr(obs, "frame_shape", None) is None:
                obs.match(self.frame_shape, model_psf)

    @property
    def parameters(self):
        """blend.py:103-105: the sources' parameters, then the observations' (renderer) parameters."""
        return tuple(p for s in self.sources for p in s.parameters) + tuple(p for o in self.observations for p in getattr(o, "parameters", ()))

    def get_model(self, values=None):
        """Sum of the boxed source models inside the frame.  blend.py:200-244, 17-27."""
        full = np.zeros(self.frame_shape, dtype=self.frame_dtype)
        i = 0
        for src in self.sources:
            n = len(src.parameters)
            vals = None if values is None else values[i:i + n]
            i += n
            fs, ms = overlapped_slices(self.frame_box, src.bbox)
            full[fs] += src.get_model(vals)[ms]
        return full

    def loss_and_grads(self, values=None):
        """Negative log-likelihood summed over observations and its gradient wrt every parameter
        (what ``autograd.grad(_loss_func)`` returns, blend.py:118, 259-274)."""
        model = self.get_model(values)
        total = 0.0
        g_model = np.zeros(self.frame_shape, dtype=np.float64)
        for obs in self.observations:
            nll, g = obs.neg_log_likelihood(model, want_grad=True)
            total += nll
            g_model += g
        grads = []
        i = 0
        for src in self.sources:
            n = len(src.parameters)
            vals = None if values is None else values[i:i + n]
            i += n
            fs, ms = overlapped_slices(self.frame_box, src.bbox)
            gbox = np.zeros(src.bbox.shape, dtype=np.float64)
            gbox[ms] = g_model[fs]
            grads.extend(src.param_grads(gbox, vals))
        for obs in self.observations:
            if getattr(obs, "parameters", ()):
                grads.extend(obs.param_grads(model))
        return total, grads

    def loss_only(self, values=None):
        model = self.get_model(values)
        return sum(obs.neg_log_likelihood(model) for obs in self.observations)

    def fit(self, max_iter=200, e_rel=1e-3, min_iter=1, prox_max_iter=10, b1=0.9, b2=0.999, eps=1e-8,
            callback=None):
        """``Blend.fit`` with ``scheme="amsgrad"`` (blend.py:85-198): one ``adaprox`` call per pass of the outer loop;
        after its iterations 10, 20, ... the sources may adapt their boxes (``_callback``, blend.py:276-302), which
        aborts the call and restarts it with warm state and ``it = len(self.loss)``.  Returns (n_iter, logL)."""
        it_outer = 0
        while it_outer < max_iter:
            params = self.parameters
            for p in params:
                sdt = np.float32 if (p.name == "image" and p.x.dtype == np.float32) else np.float64  # float32_arithmetic() only
                if p.m is None:
                    p.m = np.zeros(p.x.shape, dtype=sdt)
                if p.v is None:
                    p.v = np.zeros(p.x.shape, dtype=sdt)
                if p.vhat is None:
                    p.vhat = np.zeros(p.x.shape, dtype=sdt)
            restarted = False
            for it in range(max_iter - it_outer):  # proxmin's own counter starts at 0 in every call
                loss, grads = self.loss_and_grads()
                self.loss.append(loss)
                steps = [p.step_size(it) for p in params]
                for p, g, a in zip(params, grads, steps):
                    if p.fixed:
                        continue
                    adaprox_step(p, g, a, it, e_rel=e_rel, prox_max_iter=prox_max_iter, b1=b1, b2=b2, eps=eps)
                # ---- Blend._callback, blend.py:276-302 ----
                for p in params:
                    if not np.isfinite(p.x).all():
                        raise ArithmeticErrorNonFinite("parameter '%s' is not finite" % p.name)
                if it > 0 and it % 10 == 0:
                    changed = [src.update() for src in self.sources if hasattr(src, "update")]
                    if any(changed):
                        it_outer = len(self.loss)
                        restarted = True
                        break
                if it > min_iter and abs(self.loss[-1] - self.loss[-2]) < e_rel * abs(self.loss[-1]):
                    break
                if callback is not None:
                    callback(it)
            if not restarted:
                break
        for p in self.parameters:
            if p.v is not None:
                p.std = 1 / np.sqrt(np.ma.masked_equal(p.v, 0))
        return len(self.loss), -self.loss[-1]


def amsgrad_phi_psi(it, g, m, v, vhat, b1, b2, eps, overwrite_vhat_at_it0=True):
    """AMSGrad direction ``phi`` and metric ``psi`` -- proxmin's ``_amsgrad_phi_psi`` as recalled
    (PARITY UNPINNED, see module docstring): first/second moments without bias correction, running
    maximum of the second moment, ``psi = sqrt(max(vhat, eps))``.  Moments are updated in place."""
    m[...] = (1 - b1) * g + b1 * m
    v[...] = (1 - b2) * g ** 2 + b2 * v
    if it == 0 and overwrite_vhat_at_it0:
        vhat[...] = v
    else:
        vhat[...] = np.maximum(vhat, v)
    floor = np.maximum(vhat, eps) if eps > 0 else vhat
    return m, np.sqrt(floor)


def adaprox_step(p, g, alpha, it, e_rel=1e-3, prox_max_iter=10, b1=0.9, b2=0.999, eps=1e-8):
    """One ``proxmin.adaprox`` update of one parameter (call site blend.py:165-180; structure as in
    lite/parameters.py:274-305 minus the lite-only first-step damping): gradient step in the AMSGrad
    metric, then up to ``prox_max_iter`` proximal sub-iterations
    ``z <- prox(z - psi/max(psi) * (z - x), alpha/max(psi))`` until ``|dz|^2 <= e_rel^2 |z|^2``.
    The parameter array is updated in place (keeps its dtype)."""
    g = np.asarray(g, dtype=np.float64)
    phi, psi = amsgrad_phi_psi(it, g, p.m, p.v, p.vhat, b1, b2, eps)
    p.x -= alpha * phi / psi  # float64 arithmetic, rounded once into the parameter's dtype
    if p.prox is not None:
        z = p.x.copy()
        gamma = alpha / np.max(psi)
        for _ in range(prox_max_iter):
            z_new = p.prox(z - gamma / alpha * psi * (z - p.x), gamma)
            done = ((z_new - z) ** 2).sum() <= e_rel ** 2 * (z ** 2).sum()
            z = z_new
            if done:
                break
        p.x[...] = z
