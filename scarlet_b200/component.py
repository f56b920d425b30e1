"""Model-tree nodes.  Mirrors scarlet/component.py: ``Component`` 12-116, ``FactorizedComponent`` 119-193,
``CombinedComponent`` 229-290 (operation "add")."""
import numpy as np

from .bbox import Box, overlapped_slices
from .frame import Frame
from .model import Model, UpdateException
from .morphology import Morphology
from .spectrum import Spectrum


class Component(Model):
    def __init__(self, frame, *parameters, children=None, bbox=None):
        assert isinstance(frame, Frame)
        if bbox is None:
            bbox = frame.bbox
        assert isinstance(bbox, Box)
        self._bbox = bbox
        self.frame = frame
        super().__init__(*parameters, children=children)

    @property
    def bbox(self):
        return self._bbox

    @bbox.setter
    def bbox(self, b):
        self._bbox = self._frame.bbox if b is None else b
        self._model_frame_slices, self._model_slices = overlapped_slices(self._frame.bbox, self._bbox)

    @property
    def frame(self):
        return self._frame

    @frame.setter
    def frame(self, f):
        self._frame = f
        self._model_frame_slices, self._model_slices = overlapped_slices(self._frame.bbox, self._bbox)

    def model_to_box(self, bbox=None, model=None):
        if model is None:
            model = self.get_model()
        if bbox is None or bbox == self.frame.bbox:
            bbox = self.frame.bbox
            frame_slices, model_slices = self._model_frame_slices, self._model_slices
        else:
            frame_slices, model_slices = overlapped_slices(bbox, self.bbox)
        result = np.zeros(bbox.shape, dtype=model.dtype)
        result[frame_slices] = model[model_slices]
        return result


class FactorizedComponent(Component):
    """One source: spectrum (C,) x morphology (By,Bx) [or (C,By,Bx)] inside a bounding box."""

    def __init__(self, frame, spectrum, morphology):
        assert isinstance(spectrum, Spectrum)
        assert isinstance(morphology, Morphology)
        super().__init__(frame, children=[spectrum, morphology], bbox=spectrum.bbox @ morphology.bbox[-2:])

    def get_model(self, *parameters, frame=None):
        spectrum, morphology = self.get_models_of_children(*parameters)
        spectrum, morphology = np.asarray(spectrum), np.asarray(morphology)
        if morphology.ndim == 2:
            model = spectrum[:, None, None] * morphology[None, :, :]
        elif morphology.ndim == 3:
            model = spectrum[:, None, None] * morphology
        else:
            raise AttributeError("morphology must be 2D or 3D")
        if frame is not None:
            model = self.model_to_box(frame.bbox, model)
        return model

    def update(self):
        """Let the children adapt (dynamic morphology box); re-derive the component box (component.py:173-181)."""
        for child in self.children:
            try:
                child.update()
            except UpdateException as e:
                spectrum, morphology = self.children
                self.bbox = spectrum.bbox @ morphology.bbox[-2:]
                raise e

    @property
    def spectrum(self):
        return self.children[0]

    @property
    def morphology(self):
        return self.children[1]


class CombinedComponent(Component):
    def __init__(self, components, operation="add"):
        assert len(components)
        frame = components[0].frame
        for c in components:
            assert isinstance(c, Component)
            assert c.frame is frame
        super().__init__(frame, children=components, bbox=components[0].bbox)
        if operation != "add":
            raise NotImplementedError("only the additive combination is on the device path")
        self.operation = operation

    def update(self):
        for child in self.children:
            try:
                child.update()
            except UpdateException as e:
                box = self.children[0].bbox.copy()
                for c in self.children[1:]:
                    box = box | c.bbox
                self.bbox = box
                raise e

    def get_model(self, *parameters, frame=None):
        models = self.get_models_of_children(*parameters, frame=None)
        bbox = self.bbox
        for c in self.children[1:]:
            bbox = bbox | c.bbox
        model = np.zeros(bbox.shape)
        for c, m in zip(self.children, models):
            sl, msl = overlapped_slices(bbox, c.bbox)
            model[sl] += m[msl]
        if frame is not None:
            out = np.zeros(frame.bbox.shape, dtype=model.dtype)
            fsl, msl = overlapped_slices(frame.bbox, bbox)
            out[fsl] = model[msl]
            model = out
        return model
