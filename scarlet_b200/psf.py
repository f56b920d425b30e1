"""PSF models (host setup).  Mirrors scarlet/psf.py: ``normalize`` 9-17, ``FunctionPSF`` 39-77,
``GaussianPSF`` 80-142, ``MoffatPSF`` 143-201, ``ImagePSF`` 205-234.  PSF images are evaluated on the host once per scene to build
the difference kernel; the per-iteration evaluation of a point source's PSF image at a moving sub-pixel centre
happens on the device (csrc/kernels.cuh: point_planes / update_point)."""
import numpy as np
from scipy import special

from .bbox import Box
from .model import Model
from .parameter import Parameter, prepare_param


def normalize(image):
    """Unit sum in every band (in place for plain arrays)."""
    sums = image.sum(axis=(1, 2))
    if isinstance(image, Parameter):
        image._data[...] = image._data / sums[:, None, None]
    else:
        image /= sums[:, None, None]
    return image


class PSF(Model):
    def get_model(self, *parameters, offset=None):
        raise NotImplementedError


class FunctionPSF(PSF):
    def __init__(self, *parameters, integrate=True, boxsize=None):
        super().__init__(*parameters)
        self.integrate = integrate
        if boxsize is None:
            boxsize = 15
        if boxsize % 2 == 0:
            boxsize += 1
        p0 = self.get_parameter(0, *parameters)
        self.bbox = Box((len(p0), boxsize, boxsize), origin=(0, -(boxsize // 2), -(boxsize // 2)))
        self._Y = np.arange(boxsize) - boxsize // 2
        self._X = np.arange(boxsize) - boxsize // 2
        self.is_same = bool(np.all(p0 == p0[0]))


class GaussianPSF(FunctionPSF):
    """Circular Gaussian, integrated over pixels, per band."""

    def __init__(self, sigma, integrate=True, boxsize=None):
        sigma = prepare_param(sigma, "sigma", fixed=True)
        if boxsize is None:
            boxsize = int(np.ceil(10 * np.max(sigma)))
        super().__init__(sigma, integrate=integrate, boxsize=boxsize)

    @property
    def sigma(self):
        return self.get_parameter(0)

    def get_model(self, *parameters, offset=None):
        sigma = self.get_parameter(0, *parameters)
        oy, ox = (0, 0) if offset is None else (offset[0], offset[1])
        sig = [sigma[0]] if self.is_same else list(sigma)
        planes = [self._f(self._Y - oy, s)[:, None] * self._f(self._X - ox, s)[None, :] for s in sig]
        return normalize(np.stack(planes, axis=0))

    def _f(self, X, sigma):
        if not self.integrate:
            return np.exp(-(X ** 2) / (2 * sigma ** 2))
        s2 = np.sqrt(2) * sigma
        return np.sqrt(np.pi / 2) * sigma * (1 - special.erfc((0.5 - X) / s2) + 1 - special.erfc((2 * X + 1) / (2 * s2)))


class MoffatPSF(FunctionPSF):
    """Circular Moffat profile ``(1 + r^2 / alpha^2)^-beta`` per band, sampled at the pixel centres (the reference has no
    pixel-integrated form either, psf.py:143-201) and normalised to unit sum over its box.  Usable as the PSF of an
    observation; the device code for point sources assumes a Gaussian MODEL-frame PSF."""

    def __init__(self, alpha=4.7, beta=1.5, integrate=False, boxsize=None):
        alpha = prepare_param(alpha, "alpha", fixed=True)
        beta = prepare_param(beta, "beta", fixed=True)
        assert len(alpha) == len(beta)
        assert integrate is False, "In-pixel integration not implemented (yet)!"
        if boxsize is None:
            boxsize = int(np.ceil(5 * np.max(alpha)))
        super().__init__(alpha, beta, integrate=integrate, boxsize=boxsize)

    def get_model(self, *parameters, offset=None):
        alpha, beta = self.get_parameter(0, *parameters), self.get_parameter(1, *parameters)
        oy, ox = (0, 0) if offset is None else (offset[0], offset[1])
        pairs = [(alpha[0], beta[0])] if self.is_same else list(zip(alpha, beta))  # "same" looks at alpha only, like the reference
        r2 = (self._Y - oy)[:, None] ** 2 + (self._X - ox)[None, :] ** 2
        return normalize(np.stack([(1 + r2 / a ** 2) ** -b for a, b in pairs], axis=0))


class ImagePSF(PSF):
    """PSF from a centred image (2-D or per band); normalised on construction."""

    def __init__(self, image):
        image = np.array(image, dtype=np.float64)
        if image.ndim == 2:
            image = image[None]
        image = prepare_param(normalize(image), "image", fixed=True)
        super().__init__(image)
        self.bbox = Box(image.shape, origin=(0, -(image.shape[1] // 2), -(image.shape[2] // 2)))

    def get_model(self, *parameters, offset=None):
        image = np.array(self.get_parameter(0, *parameters)._data if not parameters else parameters[0])
        if offset is not None:  # band by band through the host Fourier shift (psf.py:228-234, fft.py:399-428)
            from . import fft
            image = np.stack([fft.shift(plane, offset, return_Fourier=False) for plane in image], axis=0)
        return image
