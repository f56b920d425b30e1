"""scarlet_b200 -- B200-native implementation of scarlet's proximal-gradient fitting path (``Blend.fit``).

The public names mirror the reference package for everything on that path; the per-iteration work runs in
hand-written CUDA (sm_100a) + cuFFT behind the C ABI in ``include/scarlet_b200.h``.  No CPU fallback.
"""
from . import fft, initialization, measure, operator  # noqa: F401
from .bbox import Box, overlapped_slices  # noqa: F401
from .blend import BatchPipeline, Blend, BlendBatch  # noqa: F401
from .cache import Cache  # noqa: F401
from .component import CombinedComponent, Component, FactorizedComponent  # noqa: F401
from .constraint import (CenterOnConstraint, Constraint, ConstraintChain, MonotonicityConstraint,  # noqa: F401
                         NormalizationConstraint, PositivityConstraint, SymmetryConstraint)
from .frame import Frame  # noqa: F401
from .model import Model, UpdateException  # noqa: F401
from .morphology import ExtendedSourceMorphology, ImageMorphology, Morphology, PointSourceMorphology  # noqa: F401
from .observation import Observation  # noqa: F401
from .parameter import Parameter, relative_step  # noqa: F401
from .prior import Prior  # noqa: F401
from .psf import PSF, FunctionPSF, GaussianPSF, ImagePSF, MoffatPSF  # noqa: F401
from .renderer import ConvolutionRenderer, NullRenderer, Renderer  # noqa: F401
from .source import (CompactExtendedSource, ExtendedSource, MultiExtendedSource, PointSource,  # noqa: F401
                     SingleExtendedSource)
from .spectrum import Spectrum, TabulatedSpectrum  # noqa: F401

__version__ = "0.1.0"
