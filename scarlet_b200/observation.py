"""Observation: data cube + weights + PSF, matched to a model frame.  Mirrors scarlet/observation.py
(``match`` 59-114, ``noise_rms`` 116-124, ``render`` 131-145, ``get_log_likelihood`` 147-170, ``log_norm`` 172-186)."""
import numpy as np
import numpy.ma as ma

from .frame import Frame
from .renderer import ConvolutionRenderer, NullRenderer, Renderer, ResolutionRenderer


class Observation(Frame):
    def __init__(self, data, channels, psf=None, weights=None, wcs=None, padding=10):
        super().__init__(data.shape, wcs=wcs, psf=psf, channels=channels, dtype=data.dtype)
        self.data = data
        self.weights = weights if weights is not None else np.ones(data.shape, dtype=data.dtype)
        assert self.weights.shape == self.data.shape, "Weights needs to have same shape as data"
        self._padding = padding

    def match(self, model_frame, renderer=None):
        self.model_frame = model_frame
        if self.dtype != model_frame.dtype:
            self.dtype = model_frame.dtype
            self.data = self.data.astype(model_frame.dtype)
            if type(self.weights) is np.ndarray:
                self.weights = self.weights.astype(model_frame.dtype)
        if renderer is None:
            if self.psf is model_frame.psf:
                self.renderer = NullRenderer(self, model_frame)
            else:
                assert self.psf is not None and model_frame.psf is not None
                if self.wcs is model_frame.wcs:
                    self.renderer = ConvolutionRenderer(self, model_frame, convolution_type="fft")
                else:
                    from . import interpolation
                    assert self.wcs is not None and model_frame.wcs is not None
                    angle, h = interpolation.get_angles(self.wcs, model_frame.wcs)
                    same_res = abs(h - 1) < np.finfo(float).eps
                    same_rot = (np.abs(angle[1]) ** 2) < np.finfo(float).eps
                    if same_res and same_rot:
                        self.renderer = ConvolutionRenderer(self, model_frame, convolution_type="fft")
                    else:
                        self.renderer = ResolutionRenderer(self, model_frame)
        else:
            assert isinstance(renderer, Renderer)
            self.renderer = renderer
        return self

    @property
    def noise_rms(self):
        if not hasattr(self, "_noise_rms"):
            self._noise_rms = 1 / np.sqrt(ma.masked_equal(self.weights, 0))
            ma.set_fill_value(self._noise_rms, np.inf)
        return self._noise_rms

    @property
    def parameters(self):
        return self.renderer.parameters

    def render(self, model, *parameters):
        return self.renderer(model, *parameters)

    def get_log_likelihood(self, model, *parameters, noise_factor=0):
        if noise_factor > 0:
            raise NotImplementedError("noise injection (host RNG) is outside the device path")
        model_ = self.render(model, *parameters)
        return -self.log_norm - np.sum(self.weights * (model_ - self.data) ** 2) / 2

    @property
    def log_norm(self):
        if not hasattr(self, "_log_norm"):
            rms = self.noise_rms
            D = np.prod(self.data.shape) - np.sum(ma.getmaskarray(rms))
            self._log_norm = D / 2 * np.log(2 * np.pi)
            with np.errstate(divide="ignore"):
                self._log_norm += np.log(rms).sum()
        return self._log_norm
