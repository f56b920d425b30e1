"""Post-fit measurements on component models: the step right after the fitting path.

Same call signatures as scarlet/measure.py (``max_pixel`` 6-21, ``flux`` 24-38, ``centroid`` 41-59, ``snr`` 62-104, ``moments`` 108-150): each
takes a component (anything with ``get_model()`` and ``bbox``) or a plain (C, Ny, Nx) cube.  Plain reductions over
models that are already on the host after ``Blend.fit`` wrote the parameters back; ``snr`` renders through the
observations (device convolution for ``ConvolutionRenderer``)."""
import numpy as np


def _cube(component, frame=None):
    """-> (model cube, origin of its box)"""
    if hasattr(component, "get_model"):
        model = component.get_model(frame=frame) if frame is not None else component.get_model()
        return np.asarray(model), np.asarray((0, 0, 0) if frame is not None else component.bbox.origin)
    return np.asarray(component), np.zeros(np.ndim(component), dtype=int)


def max_pixel(component):
    """(channel, y, x) of the brightest model pixel, in frame coordinates."""
    model, origin = _cube(component)
    return tuple(np.array(np.unravel_index(int(np.argmax(model)), model.shape)) + origin)


def flux(component):
    """Model flux per channel."""
    return _cube(component)[0].sum(axis=(1, 2))


def centroid(component):
    """Flux-weighted mean position (channel, y, x) in frame coordinates."""
    model, origin = _cube(component)
    total = model.sum()
    grids = np.indices(model.shape)
    return np.array([(g * model).sum() / total for g in grids]) + origin


def snr(component, observations):
    """Matched-filter signal-to-noise with the rendered model itself as weight function (Erben et al. 2001, eq. 16,
    summed over all pixels and bands of all observations)."""
    if not hasattr(observations, "__iter__"):
        observations = (observations,)
    model, _ = _cube(component, frame=observations[0].model_frame)
    signal = weight2 = 0.0
    for obs in observations:
        rendered = np.asarray(obs.render(model), dtype=np.float64)
        w = rendered / rendered.sum(axis=(-2, -1))[:, None, None]
        # like the reference, the variance of a zero-weight pixel is what lies underneath the mask of noise_rms (1: numpy.ma
        # keeps the numerator of 1/sqrt(w) there) -- its np.concatenate of the masked arrays drops the masks
        var = np.asarray(np.ma.getdata(obs.noise_rms), dtype=np.float64) ** 2
        signal += float((rendered * w).sum())
        weight2 += float((var * w ** 2).sum())
    return signal / np.sqrt(weight2)


def moments(component, N=2, centroid=None, weight=None):
    """Weighted image moments up to order ``N`` as a dict ``{(p, q): value per channel}`` (measure.py:108-150).

    Conventions are the reference's, kept as they are so that numbers agree: the first key index ``p`` is the power of the
    coordinate along the LAST image axis measured from ``centroid[0]``, the second, ``q``, the power of the coordinate along
    the second-to-last axis measured from ``centroid[1]``; the default centroid is ``shape // 2`` of the whole array (for a
    cube its first two entries are ``C // 2`` and ``Ny // 2``)."""
    model = np.asarray(component.get_model() if hasattr(component, "get_model") else component)
    if weight is None:
        weight = 1
    else:
        assert model.shape == np.shape(weight)
    if centroid is None:
        centroid = np.array(model.shape) // 2
    along_rows, along_cols = np.indices(model.shape[-2:], dtype=np.float64)
    u = along_cols - centroid[0]
    v = along_rows - centroid[1]
    out = {}
    for n in range(N + 1):
        for p in range(n + 1):
            out[p, n - p] = (u ** p * v ** (n - p) * model * weight).sum(axis=(-2, -1))
    return out
