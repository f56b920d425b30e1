"""Frame metadata.  Mirrors scarlet/frame.py:12-153 (host only)."""
import logging

import numpy as np

from .bbox import Box
from .psf import PSF, ImagePSF

logger = logging.getLogger("scarlet_b200.frame")


class Frame:
    def __init__(self, shape, channels, wcs=None, psf=None, dtype=np.float32):
        self._bbox = Box(shape)
        assert len(channels) == self.C
        self.channels = channels
        self.wcs = wcs  # duck-typed; only ``None`` (pure pixel frames) reaches the device path in this round
        if psf is None:
            logger.warning("No PSF specified. Possible, but dangerous!")
            self._psf = None
        else:
            self._psf = psf if isinstance(psf, PSF) else ImagePSF(psf)
        self.dtype = dtype

    @property
    def bbox(self):
        return self._bbox

    @property
    def shape(self):
        return self._bbox.shape

    @property
    def C(self):
        return self._bbox.shape[0]

    @property
    def Ny(self):
        return self._bbox.shape[1]

    @property
    def Nx(self):
        return self._bbox.shape[2]

    @property
    def psf(self):
        return self._psf

    def get_pixel(self, sky_coord):
        sky = np.array(sky_coord, dtype=np.float64).reshape(-1, 2)
        if self.wcs is not None:
            pixel = np.flip(np.array(self.wcs.celestial.world_to_pixel_values(sky)).reshape(-1, 2), axis=-1)
        else:
            pixel = sky
        return pixel[0] if pixel.size == 2 else pixel

    def get_sky_coord(self, pixel):
        pix = np.array(pixel, dtype=np.float64).reshape(-1, 2)
        if self.wcs is not None:
            sky = np.array(self.wcs.celestial.pixel_to_world_values(np.flip(pix, axis=-1)))
        else:
            sky = pix
        return sky[0] if sky.size == 2 else sky

    def convert_pixel_to(self, target, pixel=None):
        if pixel is None:
            y, x = np.indices(self.shape[-2:], dtype=np.float64)
            pixel = np.stack((y.flatten(), x.flatten()), axis=1)
        out = target.get_pixel(self.get_sky_coord(pixel))
        return out
