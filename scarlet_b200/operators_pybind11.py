"""Same-name, same-signature stand-in for the reference's native module ``scarlet.operators_pybind11``
(operators_pybind11.cc:234-260) for the one symbol on the fitting path."""
from .operator import native_prox_weighted_monotonic as prox_weighted_monotonic  # noqa: F401


def _not_on_path(name):
    def f(*a, **k):
        raise NotImplementedError("%s is outside the B200 fitting path (SURVEY.md section 2.2)" % name)
    return f


def apply_filter(image, values, y_start, y_end, x_start, x_end, result):
    """Same signature as the reference binding (operators_pybind11.cc:39-56, :248-249): ``result`` (a contiguous array of the
    image's shape and dtype) is overwritten with the sum of shifted, scaled copies of ``image``; executed on the GPU."""
    import numpy as np
    from . import _native as nat
    image = np.asarray(image)
    if image.ndim != 2 or image.dtype not in (np.float32, np.float64):
        raise TypeError("image must be a 2-D float32 or float64 array")
    if not (isinstance(result, np.ndarray) and result.flags.c_contiguous and result.shape == image.shape and result.dtype == image.dtype):
        raise TypeError("result must be a C-contiguous array of the image's shape and dtype (it is overwritten)")
    img = np.ascontiguousarray(image)
    vals = np.ascontiguousarray(values, dtype=image.dtype).reshape(-1)
    idx = [np.ascontiguousarray(a, dtype=np.int32).reshape(-1) for a in (y_start, y_end, x_start, x_end)]
    if any(a.size != vals.size for a in idx):
        raise ValueError("values and the four bound arrays must have the same length")
    fn = nat.lib().sb_apply_filter_f32 if image.dtype == np.float32 else nat.lib().sb_apply_filter_f64
    nat.check(fn(nat.ptr(img), img.shape[0], img.shape[1], nat.ptr(vals), nat.ptr(idx[0]), nat.ptr(idx[1]), nat.ptr(idx[2]), nat.ptr(idx[3]),
                 int(vals.size), nat.ptr(result), nat.default_device()))
    return None
get_valid_monotonic_pixels = _not_on_path("get_valid_monotonic_pixels")
linear_interpolate_invalid_pixels = _not_on_path("linear_interpolate_invalid_pixels")
