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
    """A spatial model confined to ``bbox`` of the model frame."""

    def __init__(self, frame, *parameters, bbox=None):
        if not isinstance(frame, Frame):
            raise AssertionError("frame must be a Frame")
        box = frame.bbox if bbox is None else bbox
        if not isinstance(box, Box):
            raise AssertionError("bbox must be a Box")
        self.frame, self.bbox = frame, box
        super().__init__(*parameters)

    def shrink_box(self, image, thresh=0):
        """Count the complete outer rings of ``image`` that hold nothing above ``thresh`` and cut the (square) box down to
        the smallest allowed size that still contains the rest (morphology.py:52-68).  Only ``self.bbox`` changes."""
        size = max(image.shape)

        def ring_is_empty(d):
            edges = (image[d, :], image[-d - 1, :], image[:, d], image[:, -d - 1])
            return all(np.all(e <= thresh) for e in edges)

        empty_rings = 0
        while ring_is_empty(empty_rings):
            empty_rings += 1
        target = get_minimal_boxsize(size - 2 * empty_rings)
        if target < size:
            trim = (size - target) // 2
            self.bbox = Box((target, target), origin=tuple(o + trim for o in self.bbox.origin))


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
        if not isinstance(image, Parameter):  # bare array: positive image with a step relative to its mean
            image = Parameter(image, name="image", step=relative_step, constraint=PositivityConstraint())
        elif image.name != "image":
            raise AssertionError("the image parameter must be named 'image'")
        if bbox is None:  # the image then has to cover the whole frame
            bbox = Box(image.shape)
            if frame.bbox[1:].shape != image.shape:
                raise AssertionError("an image without a box must have the shape of the frame")
        elif bbox.shape != image.shape:
            raise AssertionError("box and image shapes differ")
        self.resizing, self.shifting = resizing, shifting
        # The second parameter always exists: the reference also creates a free 'shift' that nothing reads when
        # shifting=False (morphology.py:112-113); it keeps the parameter tuple aligned with the reference's.
        if shift is None:
            shift = Parameter(np.zeros(2), name="shift", step=1e-2, fixed=self.shifting)
        elif isinstance(shift, Parameter):
            if shift.name != "shift" or shift.shape != (2,):
                raise AssertionError("the shift parameter must be named 'shift' and hold (dy, dx)")
        else:
            shift = Parameter(np.asarray(shift), name="shift", step=1e-2)
            if shift.shape != (2,):
                raise AssertionError("shift must hold (dy, dx)")
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
            # (the reference forms the masked quotient over the whole image, morphology.py:165-170; only the four edges enter,
            # evaluated here with the same operations in the same order: pixels with v == 0 are left out of the mean)
            edge_pull = np.array([_edge_pull(image.m[sl], image.v[sl], image._data[sl], image.step)
                                  for sl in ((slice(None), 0), (slice(None), -1), (0, slice(None)), (-1, slice(None)))])
            if np.any(edge_pull > 0.1):
                size = max(bbox.shape)
                newsize = get_minimal_boxsize(size + 1)
                pad = (newsize - size) // 2
                def grown(a):  # zero padding (numpy.pad(mode="constant") without its generic machinery)
                    if a is None:
                        return None
                    out = np.zeros(tuple(n + 2 * pad for n in a.shape), dtype=a.dtype)
                    out[(slice(pad, -pad),) * a.ndim] = a
                    return out
                self._replace_image(image, np.pad(image._data, pad, mode="linear_ramp"), grown(image.m), grown(image.v), grown(image.vhat))
                self.bbox = Box((newsize, newsize), origin=tuple(o - pad for o in self.bbox.origin))
                raise UpdateException


def _edge_pull(m, v, data, step):
    """mean over one edge of the next gradient update ``-m / sqrt(sqrt(v)) * step`` where the image is positive, pixels with
    ``v == 0`` masked out -- what ``numpy.ma`` gives for the reference's expression (nan when every pixel is masked)"""
    ok = v != 0
    n = int(np.count_nonzero(ok))
    if n == 0:
        return np.nan
    gu = np.zeros(m.shape, dtype=np.result_type(m, v, np.float64))
    gu[ok] = (-m[ok] / np.sqrt(np.sqrt(v[ok])) * step) * (data[ok] > 0)
    return gu.sum() / n


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
        # projection order of morphology.py:644-669: monotonic -> [symmetric] -> positive -> centre floor -> peak = 1
        weighting = "angle" if monotonic is True else (None if monotonic is False else monotonic)
        chain = [MonotonicityConstraint(neighbor_weight=weighting, min_gradient=min_grad)] if weighting is not None else []
        if symmetric:
            chain.append(SymmetryConstraint())
        chain.extend((PositivityConstraint(), CenterOnConstraint(), NormalizationConstraint("max")))
        image = Parameter(image, name="image", step=1e-2, constraint=ConstraintChain(*chain))
        self.pixel_center = np.round(center).astype("int")
        shift = Parameter(center - self.pixel_center, name="shift", step=1e-1) if shifting else None
        self.shift = shift
        super().__init__(frame, image, bbox=bbox, shifting=shifting, shift=shift, resizing=resizing)

    @property
    def center(self):
        if self.shift is not None:
            return self.pixel_center + self.shift
        return self.pixel_center
