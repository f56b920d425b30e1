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
        """Common model frame of several observations (frame.py:155-287): the pixel grid of the finest observation (or of
        ``obs_id``), the narrowest PSF among the eligible bands as model PSF, the union / intersection of the footprints grown
        by half the widest PSF; every observation is matched to the result."""
        assert coverage in ["union", "intersection"]
        observations = tuple(observations) if hasattr(observations, "__iter__") else (observations,)
        scales = [interpolation.get_pixel_size(np.asarray(interpolation.get_affine(o.wcs))) for o in observations]
        channels = [ch for o in observations for ch in o.channels]

        # PSF widths of all bands in sky units: the widest pads the frame, the narrowest eligible one becomes the model PSF
        widths = [(interpolation.get_psf_size(img) * scales[k], k, img) for k, o in enumerate(observations) for img in o.psf.get_model()]
        widest = max(w for w, _, _ in widths)
        reference = observations[int(np.argmin(scales))] if obs_id is None else observations[obs_id]  # first finest grid
        if model_wcs is None:
            model_wcs = reference.wcs
        h = interpolation.get_pixel_size(np.asarray(interpolation.get_affine(model_wcs)))
        if model_psf is None:
            eligible = [t for t in widths if obs_id is None or t[1] == obs_id]
            _, k_narrow, img = eligible[int(np.argmin([t[0] for t in eligible]))]  # first of the narrowest
            model_psf = ImagePSF(img[np.newaxis, :, :])
            if scales[k_narrow] > h:  # tabulated on a coarser grid than the model's: resample it
                # (sic: the reference measures the rotation against the LAST observation's grid, frame.py:229)
                angle, h = interpolation.get_angles(model_wcs, observations[-1].wcs)
                model_psf = ImagePSF(interpolation.sinc_interp_inplace(model_psf.get_model(), scales[k_narrow], h, angle))

        # footprints in the pixel grid of the model
        grid = Frame((len(channels), 0, 0), channels=channels, psf=model_psf, wcs=model_wcs)
        area = None
        for o in observations:
            if grid.wcs is o.wcs:
                foot = reference.bbox[-2:]
            else:
                yx = o.convert_pixel_to(grid)
                lo, hi = np.floor(yx.min(axis=0)).astype(int), np.ceil(yx.max(axis=0)).astype(int)
                foot = Box.from_bounds((int(lo[0]), int(hi[0]) + 1), (int(lo[1]), int(hi[1]) + 1))
            if area is None:
                area = foot
            elif coverage == "union":
                area |= foot
            else:
                area &= foot
        margin = int(np.round(widest / h / 2))
        area = Box(tuple(n + 2 * margin for n in area.shape), origin=tuple(o - margin for o in area.origin))
        model_wcs = model_wcs.deepcopy()
        model_wcs.wcs.crpix -= area.origin  # sic: (y, x) origin subtracted from the (x, y) reference pixel (frame.py:274)
        model_wcs.array_shape = area.shape
        model_frame = Frame((len(channels),) + tuple(area.shape), channels=channels, psf=model_psf, wcs=model_wcs)
        for o in observations:
            o.match(model_frame)
        return model_frame
