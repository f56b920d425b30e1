"""Minimal affine world-coordinate system for multi-resolution scenes when astropy is not installed.

The reference takes ``astropy.wcs.WCS`` objects and only ever touches, on the fitting path: ``wcs.wcs.pc`` or ``wcs.cd``
(interpolation.py:378-384), ``wcs.celestial.world_to_pixel_values / pixel_to_world_values`` (frame.py:94-97, 116-120),
``wcs.deepcopy()``, ``wcs.wcs.crpix`` and ``wcs.array_shape`` (frame.py:273-275).  ``AffineWCS`` offers exactly that
surface for a tangent-plane-free, purely linear mapping  world = crval + cd @ (pixel - crpix)  (x/y order, 0-based
pixels), which is all the synthetic multi-resolution configuration needs.  A real ``astropy.wcs.WCS`` works too: every
access in ``Frame`` / ``ResolutionRenderer`` is duck-typed.
"""
import copy

import numpy as np


class _Inner:
    def __init__(self, pc, crpix):
        self.pc = np.array(pc, dtype=np.float64)
        self.crpix = np.array(crpix, dtype=np.float64)


class AffineWCS:
    def __init__(self, cd, crpix=(0.0, 0.0), crval=(0.0, 0.0), array_shape=None):
        self.wcs = _Inner(cd, crpix)
        self.crval = np.array(crval, dtype=np.float64)
        self.array_shape = array_shape

    @property
    def celestial(self):
        return self

    def pixel_to_world_values(self, pix):
        pix = np.asarray(pix, dtype=np.float64).reshape(-1, 2)  # (x, y)
        return (pix - self.wcs.crpix) @ self.wcs.pc.T + self.crval

    def world_to_pixel_values(self, sky):
        sky = np.asarray(sky, dtype=np.float64).reshape(-1, 2)
        return (sky - self.crval) @ np.linalg.inv(self.wcs.pc).T + self.wcs.crpix

    def deepcopy(self):
        return copy.deepcopy(self)
