"""Morphologies of factorized components.  Mirrors scarlet/morphology.py: ``Morphology`` 31-68,
``ImageMorphology`` 71-207, ``PointSourceMorphology`` 476-513, ``ExtendedSourceMorphology`` 607-688."""
import numpy as np

from .bbox import Box
from .constraint import (CenterOnConstraint, ConstraintChain, MonotonicityConstraint, NormalizationConstraint,
                         PositivityConstraint, SymmetryConstraint)
from .frame import Frame
from .model import Model
from .parameter import Parameter, prepare_param, relative_step
from .psf import PSF


class Morphology(Model):
    def __init__(self, frame, *parameters, bbox=None):
        assert isinstance(frame, Frame)
        self.frame = frame
        if bbox is None:
            bbox = frame.bbox
        assert isinstance(bbox, Box)
        self.bbox = bbox
        super().__init__(*parameters)


class ImageMorphology(Morphology):
    """Free-form image.  ``shifting`` (Fourier sub-pixel shift) and ``resizing`` (dynamic box) are accepted for
    API parity; the device path currently fits ``shifting=False`` and treats the box as fixed (see Blend.fit)."""

    def __init__(self, frame, image, bbox=None, shifting=False, shift=None, resizing=True):
        if isinstance(image, Parameter):
            assert image.name == "image"
        else:
            image = Parameter(image, name="image", step=relative_step, constraint=PositivityConstraint())
        if bbox is None:
            assert frame.bbox[1:].shape == image.shape
            bbox = Box(image.shape)
        else:
            assert bbox.shape == image.shape
        self.resizing = resizing
        self.shifting = shifting
        if shift is None:
            # the reference creates this free-but-unused parameter too (morphology.py:112-113)
            shift = Parameter(np.zeros(2), name="shift", step=1e-2, fixed=self.shifting)
        else:
            assert shift.shape == (2,)
            if isinstance(shift, Parameter):
                assert shift.name == "shift"
            else:
                shift = Parameter(shift, name="shift", step=1e-2)
        super().__init__(frame, image, shift, bbox=bbox)

    def get_model(self, *parameters):
        if self.shifting:
            raise NotImplementedError("Fourier-shifted morphologies are a 'next' row (SURVEY f-3)")
        return self.get_parameter(0, *parameters)


class PointSourceMorphology(Morphology):
    """The model-frame PSF evaluated at a free sub-pixel ``center``."""

    def __init__(self, frame, center):
        assert frame.psf is not None and isinstance(frame.psf, PSF)
        self.psf = frame.psf
        pixel_center = tuple(np.round(center).astype("int"))
        bbox = self.psf.bbox + (0, *pixel_center)
        self.center = prepare_param(center, name="center")
        super().__init__(frame, self.center, bbox=bbox)

    def get_model(self, *parameters):
        center = self.get_parameter(0, *parameters)
        box_center = np.mean(self.bbox.bounds[1:], axis=1)
        return self.psf.get_model(offset=np.asarray(center) - box_center)

    @property
    def integral(self):
        return self.psf.get_model().sum()


class ExtendedSourceMorphology(ImageMorphology):
    """Free-form galaxy image: monotonic (optionally symmetric), positive, centre pinned, peak-normalised."""

    def __init__(self, frame, center, image, bbox=None, monotonic="angle", symmetric=False, min_grad=0,
                 shifting=False, resizing=True):
        constraints = []
        if monotonic is True:
            monotonic = "angle"
        elif monotonic is False:
            monotonic = None
        if monotonic is not None:
            constraints.append(MonotonicityConstraint(neighbor_weight=monotonic, min_gradient=min_grad))
        if symmetric:
            constraints.append(SymmetryConstraint())
        constraints += [PositivityConstraint(), CenterOnConstraint(), NormalizationConstraint("max")]
        image = Parameter(image, name="image", step=1e-2, constraint=ConstraintChain(*constraints))
        self.pixel_center = np.round(center).astype("int")
        shift = Parameter(center - self.pixel_center, name="shift", step=1e-1) if shifting else None
        self.shift = shift
        super().__init__(frame, image, bbox=bbox, shifting=shifting, shift=shift, resizing=resizing)

    @property
    def center(self):
        if self.shift is not None:
            return self.pixel_center + self.shift
        return self.pixel_center
