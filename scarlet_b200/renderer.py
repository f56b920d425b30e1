"""Model frame -> observation frame mappings.  Mirrors scarlet/renderer.py: ``Renderer`` 13-83,
``NullRenderer`` 86-94, ``match_shape`` 130-161, ``ConvolutionRenderer`` 164-259.

Setup (channel map, data/model overlap, difference kernel, K^) happens here on the host; the convolution itself
runs on the device: inside the fitting loop as part of the plan, and for a stand-alone ``renderer(model)`` call
through ``sb_fft_convolve_*``.
"""
import numpy as np

from . import fft
from .bbox import Box, overlapped_slices
from .model import Model


class Renderer(Model):
    def __init__(self, data_frame, model_frame, *parameters):
        self.data_frame = data_frame
        self.model_frame = model_frame
        self.channel_map = self.get_channel_map(data_frame, model_frame)
        super().__init__(*parameters)

    def __call__(self, model, *parameters):
        return self.get_model(*parameters)(model)

    def get_channel_map(self, data_frame, model_frame):
        if list(data_frame.channels) == list(model_frame.channels):
            return None
        idx = [list(model_frame.channels).index(c) for c in data_frame.channels]
        lo, hi = min(idx), max(idx)
        if hi + 1 - lo == len(idx) and idx == list(range(lo, hi + 1)):
            return slice(lo, hi + 1)
        raise NotImplementedError("non-contiguous channel maps are not supported (the reference returns the index "
                                  "list, which its own map_channels cannot apply either)")

    def map_channels(self, model):
        if self.channel_map is None:
            return model
        return model[self.channel_map]

    @property
    def channel_offset(self):
        return 0 if self.channel_map is None else self.channel_map.start


def _spatial_slices(data_frame, model_frame):
    """Overlap of the data pixels with the model frame (translation only)."""
    pix = np.atleast_2d(data_frame.convert_pixel_to(model_frame))
    ll = np.round(pix.min(axis=0)).astype(int)
    ur = np.round(pix.max(axis=0)).astype(int) + 1
    data_box = model_frame.bbox[0] @ Box.from_bounds((ll[0], ur[0]), (ll[1], ur[1]))
    return overlapped_slices(data_box, model_frame.bbox), (int(ll[0]), int(ll[1]))


def match_shape(model, data_frame, slices):
    data_slices, model_slices = slices
    if any(data_slices[d].stop - data_slices[d].start != data_frame.shape[d] for d in (-2, -1)):
        out = np.zeros(data_frame.shape, dtype=data_frame.dtype)
        out[data_slices] = model[model_slices]
        return out
    return model[model_slices]


class NullRenderer(Renderer):
    """Observation and model share the PSF: channel selection + shape matching only."""

    def __init__(self, data_frame, model_frame):
        super().__init__(data_frame, model_frame)
        self.slices, self.origin = _spatial_slices(data_frame, model_frame)
        self.diff_kernel = None

    def get_model(self, *parameters):
        return lambda model: match_shape(self.map_channels(model), self.data_frame, self.slices)


class ConvolutionRenderer(Renderer):
    def __init__(self, data_frame, model_frame, *parameters, convolution_type="fft", padding=10, psf_shift=None):
        if psf_shift is not None:
            raise NotImplementedError("psf_shift is outside the device path (SURVEY 2.1: out of scope)")
        if convolution_type != "fft":
            raise NotImplementedError("real-space convolution is a 'next' row (SURVEY f-4)")
        super().__init__(data_frame, model_frame, *parameters)
        self._convolution_type = convolution_type
        self.slices, self.origin = _spatial_slices(data_frame, model_frame)
        psf_obs = fft.Fourier(data_frame.psf.get_model().astype(model_frame.dtype))
        psf_model = fft.Fourier(model_frame.psf.get_model().astype(model_frame.dtype))
        self.diff_kernel = fft.match_psf(psf_obs, psf_model, padding=padding)
        self._khat = None

    def kernel_transform(self):
        """(fft_shape, K^) of the difference kernel on the grid of the model-frame sub-cube."""
        if self._khat is None:
            sub_shape = (self.data_frame.C,) + tuple(self.model_frame.shape[1:])
            self._khat = fft.kernel_transform(self.diff_kernel.image, sub_shape, padding=3)
        return self._khat

    def device_kernel(self):
        """(fft_shape, (y0, x0), float64 kernel image) for the device fitting loop, which transforms the kernel itself."""
        sub_shape = (self.data_frame.C,) + tuple(self.model_frame.shape[1:])
        ker = np.ascontiguousarray(self.diff_kernel.image, dtype=np.float64)
        fshape, origin = fft.device_grid(sub_shape, ker.shape, padding=3)
        return fshape, origin, ker

    def convolve(self, model, convolution_type=None, psf_shift=None):
        fshape, khat = self.kernel_transform()
        return fft.device_convolve(np.asarray(model), khat, fshape)

    def get_model(self, *parameters):
        def transform(model):
            return match_shape(self.convolve(self.map_channels(model)), self.data_frame, self.slices)
        return transform


class ResolutionRenderer(Renderer):
    def __init__(self, *a, **k):
        raise NotImplementedError("ResolutionRenderer (multi-resolution k-space resampling) is scheduled after the "
                                  "same-resolution path meets its bar (SURVEY 8a-17)")
