"""Same-name, same-signature stand-in for the reference's native module ``scarlet.operators_pybind11``
(operators_pybind11.cc:234-260) for the one symbol on the fitting path."""
from .operator import native_prox_weighted_monotonic as prox_weighted_monotonic  # noqa: F401


def _not_on_path(name):
    def f(*a, **k):
        raise NotImplementedError("%s is outside the B200 fitting path (SURVEY.md section 2.2)" % name)
    return f


apply_filter = _not_on_path("apply_filter")
get_valid_monotonic_pixels = _not_on_path("get_valid_monotonic_pixels")
linear_interpolate_invalid_pixels = _not_on_path("linear_interpolate_invalid_pixels")
