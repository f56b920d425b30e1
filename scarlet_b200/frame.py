"""Frame metadata.  Mirrors scarlet/frame.py:12-153 and ``from_observations`` 155-287 (host only)."""
import logging

import numpy as np

from . import interpolation
from .bbox import Box
from .psf import PSF, ImagePSF

logger = logging.getLogger("scarlet_b200.frame")


class Frame:
    def __init__(self, shape, channels, wcs=None, psf=None, dtype=np.float32):
        self._bbox = Box(shape)
        assert len(channels) == self.C
        self.channels = channels
        self.wcs = wcs  # duck-typed (astropy.wcs.WCS or scarlet_b200.wcs.AffineWCS); the reference asserts astropy
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

    @staticmethod
    def from_observations(observations, model_psf=None, model_wcs=None, obs_id=None, coverage="union"):
        """Common model frame of several observations: highest resolution, smallest PSF, union / intersection of the
        footprints padded by the widest PSF; matches every observation to it (frame.py:155-287)."""
        assert coverage in ["union", "intersection"]
        if not hasattr(observations, "__iter__"):
            observations = (observations,)
        pix_tab, channels = [], []
        fat_psf_size = small_psf_size = None
        model_psf_temp, psf_h = None, None
        for c, obs in enumerate(observations):
            channels = channels + list(obs.channels)
            h_temp = interpolation.get_pixel_size(np.asarray(interpolation.get_affine(obs.wcs)))
            pix_tab.append(h_temp)
            for psf in obs.psf.get_model():
                psf_size = interpolation.get_psf_size(psf) * h_temp
                if fat_psf_size is None or psf_size > fat_psf_size:
                    fat_psf_size = psf_size
                if obs_id is None or c == obs_id:
                    if model_psf is None and (small_psf_size is None or psf_size < small_psf_size):
                        small_psf_size = psf_size
                        model_psf_temp = ImagePSF(psf[np.newaxis, :, :])
                        psf_h = h_temp
        obs_ref = observations[int(np.where(np.array(pix_tab) == np.min(pix_tab))[0][0])] if obs_id is None else observations[obs_id]
        if model_wcs is None:
            model_wcs = obs_ref.wcs
        h = interpolation.get_pixel_size(np.asarray(interpolation.get_affine(model_wcs)))
        if model_psf is None:
            if psf_h > h:
                angle, h = interpolation.get_angles(model_wcs, obs.wcs)
                model_psf = ImagePSF(interpolation.sinc_interp_inplace(model_psf_temp.get_model(), psf_h, h, angle))
            else:
                model_psf = model_psf_temp
        model_frame = Frame((len(channels), 0, 0), channels=channels, psf=model_psf, wcs=model_wcs)
        model_box = None
        for c, obs in enumerate(observations):
            if model_frame.wcs is obs.wcs:
                this_box = obs_ref.bbox[-2:]
            else:
                coord = obs.convert_pixel_to(model_frame)
                y_min, x_min = int(np.floor(np.min(coord[:, 0]))), int(np.floor(np.min(coord[:, 1])))
                y_max, x_max = int(np.ceil(np.max(coord[:, 0]))), int(np.ceil(np.max(coord[:, 1])))
                this_box = Box.from_bounds((y_min, y_max + 1), (x_min, x_max + 1))
            if c == 0:
                model_box = this_box
            elif coverage == "union":
                model_box |= this_box
            else:
                model_box &= this_box
        pad = int(np.round(fat_psf_size / h / 2))
        model_box = Box(tuple(s + 2 * pad for s in model_box.shape), origin=tuple(o - pad for o in model_box.origin))
        model_wcs = model_wcs.deepcopy()
        model_wcs.wcs.crpix -= model_box.origin  # sic: (y, x) origin subtracted from the (x, y) reference pixel (frame.py:274)
        model_wcs.array_shape = model_box.shape
        model_frame = Frame((len(channels),) + tuple(model_box.shape), channels=channels, psf=model_psf, wcs=model_wcs)
        for obs in observations:
            obs.match(model_frame)
        return model_frame
