"""Spectral factor of a factorized component: a 1-D array over the channels of the model frame (scarlet/spectrum.py;
``TabulatedSpectrum`` 32-71 is the free-form table every source recipe on the fitting path uses)."""
import functools

from .bbox import Box
from .constraint import PositivityConstraint
from .frame import Frame
from .model import Model
from .parameter import Parameter, relative_step

SPECTRUM_FLOOR = 1e-20  # amplitudes are kept slightly positive
SPECTRUM_STEP = 1e-2    # step = 1 % of the mean amplitude ...


def _table_parameter(values, min_step):
    """... but never below ``min_step`` (the sources pass the per-band noise rms)."""
    return Parameter(values, name="spectrum", constraint=PositivityConstraint(zero=SPECTRUM_FLOOR),
                     step=functools.partial(relative_step, factor=SPECTRUM_STEP, minimum=min_step))


class Spectrum(Model):
    """Base class: parameters plus the channel box the spectrum occupies in ``frame``."""

    def __init__(self, frame, *parameters, bbox=None):
        assert isinstance(frame, Frame) and isinstance(bbox, Box)
        self.frame, self.bbox = frame, bbox
        super().__init__(*parameters)


class TabulatedSpectrum(Spectrum):
    """One free amplitude per channel.  ``spectrum`` may be a ready-made ``Parameter`` named "spectrum" (kept as is, with
    its own step and constraint) or plain values; without ``bbox`` it must span all channels of the frame."""

    def __init__(self, frame, spectrum, bbox=None, min_step=0):
        table = spectrum if isinstance(spectrum, Parameter) else _table_parameter(spectrum, min_step)
        assert table.name == "spectrum"
        if bbox is None:
            assert table.shape == frame.bbox[0].shape
        channel_box = Box(table.shape) if bbox is None else bbox
        assert channel_box.shape == table.shape
        super().__init__(frame, table, bbox=channel_box)

    def get_model(self, *parameters):
        return self.get_parameter(0, *parameters)
