"""Observation: a data cube with per-pixel weights and a PSF, tied to a model frame by a renderer.

Behaviour follows scarlet/observation.py (``match`` 59-114, ``noise_rms`` 116-124, ``render`` 131-145,
``get_log_likelihood`` 147-170, ``log_norm`` 172-186); the implementation is this package's own.  During a fit the device
evaluates the likelihood (csrc/spectral.cuh: k_spec_residual); the methods here are the host-side forms users call around
a fit, and the constants (``log_norm``, ``noise_rms``) the plan builder uploads.
"""
import numpy as np
import numpy.ma as ma

from .frame import Frame
from .renderer import ConvolutionRenderer, NullRenderer, Renderer, ResolutionRenderer


def _pick_renderer(obs, model_frame):
    """Which built-in renderer maps ``model_frame`` onto ``obs`` (observation.py:87-112):

    same PSF object            -> nothing to do (NullRenderer)
    same WCS object            -> PSF matching only (ConvolutionRenderer, FFT)
    different WCS              -> compare pixel scale and orientation: equal within machine precision means PSF matching
                                  only, anything else needs resampling (ResolutionRenderer)
    """
    if obs.psf is model_frame.psf:
        return NullRenderer(obs, model_frame)
    if obs.psf is None or model_frame.psf is None:
        raise AssertionError("PSF matching needs a PSF on the observation and on the model frame")
    if obs.wcs is model_frame.wcs:
        return ConvolutionRenderer(obs, model_frame, convolution_type="fft")
    if obs.wcs is None or model_frame.wcs is None:
        raise AssertionError("resampling needs a WCS on the observation and on the model frame")
    from . import interpolation
    tiny = np.finfo(float).eps
    angle, scale_ratio = interpolation.get_angles(obs.wcs, model_frame.wcs)
    aligned = abs(scale_ratio - 1) < tiny and np.abs(angle[1]) ** 2 < tiny
    if aligned:
        return ConvolutionRenderer(obs, model_frame, convolution_type="fft")
    return ResolutionRenderer(obs, model_frame)


class Observation(Frame):
    def __init__(self, data, channels, psf=None, weights=None, wcs=None, padding=10):
        super().__init__(data.shape, wcs=wcs, psf=psf, channels=channels, dtype=data.dtype)
        if weights is None:
            weights = np.ones(data.shape, dtype=data.dtype)
        if weights.shape != data.shape:
            raise AssertionError("Weights needs to have same shape as data")
        self.data, self.weights = data, weights
        self._padding = padding

    def match(self, model_frame, renderer=None):
        """Adopt the model frame's dtype for the cubes and attach the renderer (given, or chosen from PSF / WCS)."""
        self.model_frame = model_frame
        if model_frame.dtype != self.dtype:
            self.dtype = model_frame.dtype
            self.data = self.data.astype(self.dtype)
            if type(self.weights) is np.ndarray:
                self.weights = self.weights.astype(self.dtype)
        if renderer is not None and not isinstance(renderer, Renderer):
            raise AssertionError("renderer must be a Renderer")
        self.renderer = _pick_renderer(self, model_frame) if renderer is None else renderer
        return self

    # ---- noise model: independent Gaussian pixels, variance 1/weight; weight 0 marks a masked pixel ----------------
    @property
    def noise_rms(self):
        cached = self.__dict__.get("_noise_rms")
        if cached is None:
            cached = 1 / np.sqrt(ma.masked_equal(self.weights, 0))
            ma.set_fill_value(cached, np.inf)
            self._noise_rms = cached
        return cached

    @property
    def log_norm(self):
        """Normalisation of the Gaussian likelihood over the unmasked pixels: n/2 log(2 pi) + sum log(rms)."""
        cached = self.__dict__.get("_log_norm")
        if cached is None:
            rms = self.noise_rms
            n_good = np.prod(self.data.shape) - np.sum(ma.getmaskarray(rms))
            with np.errstate(divide="ignore"):
                cached = n_good / 2 * np.log(2 * np.pi) + np.log(rms).sum()
            self._log_norm = cached
        return cached

    @property
    def parameters(self):
        return self.renderer.parameters

    def render(self, model, *parameters):
        return self.renderer(model, *parameters)

    def get_log_likelihood(self, model, *parameters, noise_factor=0):
        if noise_factor > 0:
            raise NotImplementedError("noise injection (host RNG) is outside the device path")
        residual = self.render(model, *parameters) - self.data
        return -self.log_norm - np.sum(self.weights * residual ** 2) / 2
