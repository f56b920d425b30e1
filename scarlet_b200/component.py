"""Model-tree nodes that live in a bounding box of the model frame.

Behaviour follows scarlet/component.py (``Component`` 12-116, ``FactorizedComponent`` 119-193, ``CombinedComponent``
229-290 with operation "add"); the implementation is this package's own.  These host objects only describe a scene: the
fitting loop reads their parameters, boxes and constraints once per plan (``_plan.py``) and renders on the device;
``get_model`` here is the NumPy path users call around a fit (measurements, plots, initialisation).
"""
import numpy as np

from .bbox import Box, overlapped_slices
from .frame import Frame
from .model import Model, UpdateException
from .morphology import Morphology
from .spectrum import Spectrum


def _paste(target_box, source_box, values, dtype=None):
    """``values`` (filling ``source_box``) embedded in a zero array that fills ``target_box``; only the overlap survives."""
    canvas = np.zeros(target_box.shape, dtype=values.dtype if dtype is None else dtype)
    into, outof = overlapped_slices(target_box, source_box)
    canvas[into] = values[outof]
    return canvas


def _union(boxes):
    hull = boxes[0].copy()
    for b in boxes[1:]:
        hull = hull | b
    return hull


class Component(Model):
    """A model confined to ``bbox`` inside ``frame``.  Changing either re-derives the cached overlap slices."""

    def __init__(self, frame, *parameters, children=None, bbox=None):
        if not isinstance(frame, Frame):
            raise AssertionError("frame must be a Frame")
        box = frame.bbox if bbox is None else bbox
        if not isinstance(box, Box):
            raise AssertionError("bbox must be a Box")
        self._frame, self._bbox = frame, box
        self._relink()
        super().__init__(*parameters, children=children)

    def _relink(self):
        # where the box sits in the frame / which part of the box is inside the frame
        self._model_frame_slices, self._model_slices = overlapped_slices(self._frame.bbox, self._bbox)

    bbox = property(lambda self: self._bbox)
    frame = property(lambda self: self._frame)

    @bbox.setter
    def bbox(self, box):
        self._bbox = self._frame.bbox if box is None else box
        self._relink()

    @frame.setter
    def frame(self, frame):
        self._frame = frame
        self._relink()

    def model_to_box(self, bbox=None, model=None):
        """The component's model re-expressed in another box (default: the whole model frame), zero outside its own."""
        values = self.get_model() if model is None else model
        if bbox is None or bbox == self._frame.bbox:
            out = np.zeros(self._frame.bbox.shape, dtype=values.dtype)
            out[self._model_frame_slices] = values[self._model_slices]
            return out
        return _paste(bbox, self._bbox, values)


class FactorizedComponent(Component):
    """One source: spectrum (C,) times morphology (By, Bx) -- or a per-band morphology (C, By, Bx) -- in a box."""

    def __init__(self, frame, spectrum, morphology):
        if not isinstance(spectrum, Spectrum) or not isinstance(morphology, Morphology):
            raise AssertionError("FactorizedComponent needs a Spectrum and a Morphology")
        super().__init__(frame, children=[spectrum, morphology], bbox=self._box_of(spectrum, morphology))

    @staticmethod
    def _box_of(spectrum, morphology):
        return spectrum.bbox @ morphology.bbox[-2:]

    spectrum = property(lambda self: self.children[0])
    morphology = property(lambda self: self.children[1])

    def get_model(self, *parameters, frame=None):
        amplitudes, image = (np.asarray(m) for m in self.get_models_of_children(*parameters))
        if image.ndim not in (2, 3):
            raise AttributeError("morphology must be 2D or 3D")
        cube = amplitudes.reshape(-1, 1, 1) * (image if image.ndim == 3 else image[np.newaxis])
        return cube if frame is None else self.model_to_box(frame.bbox, cube)

    def update(self):
        """Children may adapt themselves (dynamic morphology box, component.py:173-181).  If one did, the component box is
        re-derived before the interruption travels on to the optimiser."""
        interrupted = None
        for child in self.children:
            try:
                child.update()
            except UpdateException as exc:
                self.bbox = self._box_of(*self.children)
                interrupted = exc
                break
        if interrupted is not None:
            raise interrupted


class CombinedComponent(Component):
    """Sum of components that share one frame (a multi-component source, or the whole blend)."""

    def __init__(self, components, operation="add"):
        components = list(components)
        if not components:
            raise AssertionError("CombinedComponent needs at least one component")
        frame = components[0].frame
        if not all(isinstance(c, Component) and c.frame is frame for c in components):
            raise AssertionError("all members must be Components of the same frame")
        if operation != "add":
            raise NotImplementedError("only the additive combination is on the device path")
        super().__init__(frame, children=components, bbox=components[0].bbox)
        self.operation = operation

    def update(self):
        interrupted = None
        for child in self.children:
            try:
                child.update()
            except UpdateException as exc:
                self.bbox = _union([c.bbox for c in self.children])
                interrupted = exc
                break
        if interrupted is not None:
            raise interrupted

    def get_model(self, *parameters, frame=None):
        parts = self.get_models_of_children(*parameters, frame=None)
        hull = _union([self.bbox] + [c.bbox for c in self.children[1:]])
        total = np.zeros(hull.shape)
        for child, part in zip(self.children, parts):
            into, outof = overlapped_slices(hull, child.bbox)
            total[into] += part[outof]
        return total if frame is None else _paste(frame.bbox, hull, total)
