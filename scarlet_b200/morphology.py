"""Morphologies of factorized components.  Mirrors scarlet/morphology.py: ``Morphology`` 31-68,
``ImageMorphology`` 71-207, ``PointSourceMorphology`` 476-513, ``ExtendedSourceMorphology`` 607-688."""
import numpy as np

from .bbox import Box
from .constraint import (CenterOnConstraint, ConstraintChain, MonotonicityConstraint, NormalizationConstraint,
                         PositivityConstraint, SymmetryConstraint)
from .frame import Frame
from .bbox import overlapped_slices
from .model import Model, UpdateException
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

    def shrink_box(self, image, thresh=0):
        """Peel the onion: drop outer rings that are entirely <= thresh, down to the next allowed box size
        (morphology.py:52-68)."""
        size = max(image.shape)
        dist = 0
        while (np.all(image[dist, :] <= thresh) and np.all(image[-dist - 1, :] <= thresh)
               and np.all(image[:, dist] <= thresh) and np.all(image[:, -dist - 1] <= thresh)):
            dist += 1
        newsize = get_minimal_boxsize(size - 2 * dist)
        if newsize < size:
            dist = (size - newsize) // 2
            self.bbox = Box((newsize, newsize), origin=tuple(o + dist for o in self.bbox.origin))


def get_minimal_boxsize(size, min_size=21, increment=10):
    """Smallest allowed box size >= size (initialization.py:173-177)."""
    boxsize = min_size
    while boxsize < size:
        boxsize += increment
    return boxsize


class ImageMorphology(Morphology):
    """Free-form image.  ``resizing=True``: every 10 iterations ``Blend.fit`` calls ``update`` (host), which shrinks the
    box when its outer rings are empty or grows it when the optimiser keeps pulling flux towards an edge; the device
    loop is then re-planned with the new shapes and warm optimiser state (morphology.py:132-207, blend.py:196-198).
    ``shifting=True``: the model is the image translated by the free sub-pixel ``shift`` parameter (``fft.shift``); image
    and shift are both fitted (SURVEY f-3)."""

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
        image = self.get_parameter(0, *parameters)
        if self.shifting:  # morphology.py:124-130
            from . import fft
            return fft.shift(np.asarray(image), np.asarray(self.get_parameter(1, *parameters)), return_Fourier=False)
        return image

    def _replace_image(self, old, data, m, v, vhat):
        """New image Parameter (private, contiguous copies: the old arrays may live in a plan's staging memory), step halved."""
        def own(a):
            return None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        image = Parameter(np.ascontiguousarray(data), name=old.name, prior=old.prior, constraint=old.constraint, step=old.step / 2,
                          fixed=old.fixed, m=own(m), v=own(v), vhat=own(vhat))
        self._parameters = (image,) + self._parameters[1:]

    def update(self):
        """Dynamic box: shrink / grow and raise ``UpdateException`` when the box changed (morphology.py:132-207)."""
        import numpy.ma as ma
        image = self._parameters[0]
        if not self.resizing or image.fixed:
            return
        bbox = self.bbox.copy()
        self.shrink_box(image._data)
        if bbox != self.bbox:
            sl, _ = overlapped_slices(bbox, self.bbox)
            m, v, vhat = image.m, image.v, image.vhat
            self._replace_image(image, image._data[sl], None if m is None else m[sl], None if v is None else v[sl],
                                None if vhat is None else vhat[sl])
            raise UpdateException
        elif image.m is not None:
            # next gradient update, in units of the peak (= 1 at the centre): does the optimiser pull flux to an edge?
            gu = -image.m / np.sqrt(np.sqrt(ma.masked_equal(image.v, 0))) * image.step
            gu_pull = gu * (image._data > 0)
            edge_pull = np.array((gu_pull[:, 0].mean(), gu_pull[:, -1].mean(), gu_pull[0, :].mean(), gu_pull[-1, :].mean()))
            if np.any(edge_pull > 0.1):
                size = max(bbox.shape)
                newsize = get_minimal_boxsize(size + 1)
                pad = (newsize - size) // 2
                def grown(a):
                    return None if a is None else np.pad(a, pad, mode="constant")
                self._replace_image(image, np.pad(image._data, pad, mode="linear_ramp"), grown(image.m), grown(image.v), grown(image.vhat))
                self.bbox = Box((newsize, newsize), origin=tuple(o - pad for o in self.bbox.origin))
                raise UpdateException


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
