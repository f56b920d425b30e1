"""Host-side construction of the proximal operators' tables and the single-operator CUDA entry points.

Mirrors the slice of scarlet/operator.py that the fitting path uses: ``sort_by_radius`` 10-48,
``prox_weighted_monotonic`` 62-96, ``getRadialMonotonicWeights`` 591-667, ``prox_soft_symmetry`` 274-293.
Table construction is setup work on the host (as in the reference, cached per shape); applying an operator
always runs on the GPU through the C ABI -- there is no host implementation of the sweep in this package.
"""
import math

import numpy as np

from . import _native as nat

NEIGHBOUR_COORDS = ((-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1))


def _centre(shape, center):
    if center is None:
        return (shape[0] - 1) >> 1, (shape[1] - 1) >> 1
    return int(center[0]), int(center[1])


def sort_by_radius(shape, center=None):
    """Flat pixel indices ordered by distance from ``center`` (ties in arbitrary but fixed order)."""
    cy, cx = _centre(shape, center)
    y = np.arange(shape[0]) - cy
    x = np.arange(shape[1]) - cx
    r = np.sqrt((x[None, :] ** 2 + y[:, None] ** 2).astype(np.float64))
    return np.argsort(r.reshape(-1), kind="stable")


def getRadialMonotonicWeights(shape, neighbor_weight="flat", center=None):
    """(8, H*W) weights of the 8 neighbours of every pixel for the radial monotonicity operator.

    Neighbour n of pixel p (order ``NEIGHBOUR_COORDS``) gets a non-zero weight iff it lies inside the image and is
    strictly closer to the centre than p.  ``angle``: cosine of the angle between the direction p->centre and the
    direction p->neighbour, normalised to unit sum per pixel; ``flat``: equal weights; ``nearest``: only the best
    aligned neighbour, weight 1.
    """
    if neighbor_weight not in ("flat", "angle", "nearest"):
        raise AssertionError("neighbor_weight must be 'flat', 'angle' or 'nearest'")
    H, W = int(shape[0]), int(shape[1])
    cy, cx = _centre(shape, center)
    Y = (np.arange(H) - cy)[:, None] * np.ones((1, W), dtype=np.int64)
    X = np.ones((H, 1), dtype=np.int64) * (np.arange(W) - cx)[None, :]
    dist2 = X * X + Y * Y
    to_centre = np.arctan2((-Y).astype(np.float64), (-X).astype(np.float64))
    rows = np.arange(H)[:, None]
    cols = np.arange(W)[None, :]
    cosw = np.zeros((8, H, W), dtype=np.float64)
    for n, (dy, dx) in enumerate(NEIGHBOUR_COORDS):
        inside = (rows + dy >= 0) & (rows + dy < H) & (cols + dx >= 0) & (cols + dx < W)
        closer = (X + dx) ** 2 + (Y + dy) ** 2 < dist2
        cosw[n] = np.where(inside & closer, np.cos(to_centre - math.atan2(dy, dx)), 0.0)
    cosw = cosw.reshape(8, H * W)
    if neighbor_weight == "nearest":
        out = np.zeros_like(cosw)
        out[np.argmax(cosw, axis=0), np.arange(H * W)] = 1
        out[:, cy * W + cx] = 0
        return out
    if neighbor_weight == "flat":
        cosw[cosw != 0] = 1
    total = cosw.sum(axis=0)
    total[total == 0] = 1
    return cosw / total[None, :]


def monotonic_tables(shape, neighbor_weight="flat", center=None):
    """(weights[8,N] float64, offsets[8] int32, dist_idx[N-1] int32): the arguments of the native sweep."""
    width = int(shape[1])
    didx = sort_by_radius(shape, center)
    offsets = np.array([width * dy + dx for dy, dx in NEIGHBOUR_COORDS], dtype=np.int32)
    weights = getRadialMonotonicWeights(shape, neighbor_weight=neighbor_weight, center=center)
    return np.ascontiguousarray(weights), offsets, np.ascontiguousarray(didx[1:], dtype=np.int32)


def native_prox_weighted_monotonic(flat_img, weights, offsets, dist_idx, min_gradient, device=None):
    """Drop-in for ``scarlet.operators_pybind11.prox_weighted_monotonic`` (operators_pybind11.cc:14-36, bound
    for float32 and float64 at :243-246): in-place sweep over ``flat_img``, executed by the CUDA wavefront
    kernel.  ``flat_img`` may also be 2-D ``(n_img, n_pix)`` to process a batch with one operator."""
    if not isinstance(flat_img, np.ndarray) or not flat_img.flags.c_contiguous:
        raise TypeError("flat_img must be a C-contiguous ndarray (it is mutated in place)")
    if flat_img.dtype == np.float32:
        fn, dt = nat.lib().sb_monotonic_f32, np.float32
    elif flat_img.dtype == np.float64:
        fn, dt = nat.lib().sb_monotonic_f64, np.float64
    else:
        raise TypeError("flat_img must be float32 or float64")
    n_img = 1 if flat_img.ndim == 1 else flat_img.shape[0]
    n_pix = flat_img.shape[-1]
    w = nat.as_array(weights, dt)
    off = nat.as_array(offsets, np.int32)
    idx = nat.as_array(dist_idx, np.int32)
    if w.shape != (off.size, n_pix):
        raise TypeError("weights must have shape (len(offsets), n_pix)")
    dev = nat.default_device() if device is None else device
    nat.check(fn(nat.ptr(flat_img), nat.ptr(w), nat.ptr(off), int(off.size), nat.ptr(idx), int(idx.size), int(n_pix),
                 float(min_gradient), int(n_img), int(dev)))
    return None


def prox_weighted_monotonic(shape, neighbor_weight="flat", min_gradient=0.1, center=None):
    """Build the monotonicity operator for images of ``shape``; returns ``prox(X, step) -> X`` (in place)."""
    weights, offsets, didx = monotonic_tables(shape, neighbor_weight, center)

    def prox(X, step):
        native_prox_weighted_monotonic(X.reshape(-1), weights, offsets, didx, min_gradient)
        return X

    return prox


def prox_soft_symmetry(X, step, strength=1):
    """``strength/2 (X + rot180 X) + (1-strength) X`` -- on the GPU through the constraint-chain kernel."""
    from .constraint import SymmetryConstraint
    return SymmetryConstraint(strength)(X, step)


# --------------------------------------------------------------------------------------------------
# operators used by the source initialisation (scarlet/operator.py:207-271)
# --------------------------------------------------------------------------------------------------
def prox_sdss_symmetry(X, step):
    """Symmetrise by the MINIMUM of every pixel and its 180-degree partner (in place)."""
    X[:] = np.minimum(X, X[::-1, ::-1])
    return X


def uncentered_operator(X, func, center=None, fill=None, **kwargs):
    """Apply ``func`` on the largest sub-array of ``X`` that is centred on ``center`` (default: the peak); the rest
    keeps its values, or is set to ``fill``."""
    py, px = np.unravel_index(np.argmax(X), X.shape) if center is None else center
    cy, cx = np.array(X.shape) // 2
    if py == cy and px == cx:
        return func(X, **kwargs)
    dy, dx = int(2 * (py - cy)) + (X.shape[0] % 2 == 0), int(2 * (px - cx)) + (X.shape[1] % 2 == 0)
    ysl = slice(None, dy) if dy < 0 else slice(dy, None)
    xsl = slice(None, dx) if dx < 0 else slice(dx, None)
    if fill is not None:
        out = np.full(X.shape, fill, dtype=X.dtype)
        out[ysl, xsl] = func(X[ysl, xsl], **kwargs)
        X[:] = out
    else:
        X[ysl, xsl] = func(X[ysl, xsl], **kwargs)
    return X


def prox_uncentered_symmetry(X, step, center=None, algorithm="sdss", fill=None, strength=0.5):
    """Symmetry about an off-centre pixel.  Only the variants used on the fitting path and its initialisation exist here:
    ``"sdss"`` (minimum of partners) and ``"soft"``; the k-space variant needs ``fft.shift``."""
    if algorithm == "sdss":
        return uncentered_operator(X, prox_sdss_symmetry, center, step=step, fill=fill)
    if algorithm == "soft":
        return uncentered_operator(X, prox_soft_symmetry, center, step=step, strength=strength, fill=fill)
    raise NotImplementedError("symmetry algorithm %r is outside the device path" % (algorithm,))


def windowed_monotonic(image, center_index, neighbor_weight="flat", min_gradient=0, half_width=127):
    """Radial monotonicity about ``center_index`` on a whole detection image (what ``SingleExtendedSource.init_morph`` does
    with ``prox_weighted_monotonic(im.shape, center=...)``, source.py:488-498), executed by the CUDA wavefront kernel.

    The sweep only ever reads pixels that are closer to the centre, so a square window centred on the source gives the
    same values as the whole image inside the window; the window is capped at 255 x 255 pixels (the kernel's 16-bit pixel
    indices), pixels outside it are cleared -- sources wider than that cannot be initialised here."""
    H, W = image.shape
    cy, cx = int(center_index[0]), int(center_index[1])
    y0, y1, x0, x1 = max(0, cy - half_width), min(H, cy + half_width + 1), max(0, cx - half_width), min(W, cx + half_width + 1)
    win = np.ascontiguousarray(image[y0:y1, x0:x1], dtype=np.float64)
    prox = prox_weighted_monotonic(win.shape, neighbor_weight=neighbor_weight, min_gradient=min_gradient, center=(cy - y0, cx - x0))
    win = prox(win, 0).reshape(win.shape)
    out = np.zeros_like(image, dtype=np.float64)
    out[y0:y1, x0:x1] = win
    return out
