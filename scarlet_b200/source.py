"""Source recipes.  Mirrors the constructors of scarlet/source.py ``PointSource`` 92-128 and
``SingleExtendedSource`` 367-450 for the part that is ON the fitting path: parameter layout, step sizes
(spectrum steps floored by the per-band noise rms, source.py:412-416) and constraint chains.

Initialisation of spectra / morphologies from the data (scarlet/initialization.py) happens when
``spectrum`` / ``morphology`` are not passed in: peak-pixel spectrum, symmetric + monotonic cut-out of the
spectrum-weighted detection image (single-component ``ExtendedSource``; the compact and multi-component variants of the
reference are not covered).
"""
import numpy as np

from .bbox import Box
from .component import FactorizedComponent
from .morphology import ExtendedSourceMorphology, PointSourceMorphology
from .parameter import Parameter
from .spectrum import TabulatedSpectrum


def _noise_rms(observations):
    if observations is None:
        return 0
    if not hasattr(observations, "__iter__"):
        observations = (observations,)
    return np.concatenate([np.array(np.mean(obs.noise_rms, axis=(1, 2))) for obs in observations]).reshape(-1)


class PointSource(FactorizedComponent):
    """Model-PSF shaped source at a free sub-pixel centre."""

    def __init__(self, model_frame, sky_coord, observations, spectrum=None):
        if spectrum is None:  # peak pixel, corrected for the PSF peak (source.py:118-120)
            from . import initialization as init
            spectrum = init.get_pixel_spectrum(sky_coord, observations, correct_psf=True)
        center = Parameter(np.array(model_frame.get_pixel(sky_coord), dtype=np.float64), name="center", step=3e-2)
        morphology = PointSourceMorphology(model_frame, center)
        spec = TabulatedSpectrum(model_frame, np.asarray(spectrum), min_step=_noise_rms(observations))
        super().__init__(model_frame, spec, morphology)
        self.center = morphology.center


class ExtendedSource(FactorizedComponent):
    """Free-form monotonic (optionally symmetric) galaxy model in a square box around ``sky_coord``."""

    def __init__(self, model_frame, sky_coord, observations, spectrum=None, morphology=None, bbox=None,
                 monotonic="angle", symmetric=False, min_grad=0, shifting=False, resizing=True, thresh=1.0, boxsize=None):
        if spectrum is None or morphology is None:  # SingleExtendedSource.__init__, source.py:407-434
            from . import initialization as init
            obs_list = observations if hasattr(observations, "__iter__") else (observations,)
            spectra = init.get_pixel_spectrum(sky_coord, obs_list, concat=False)
            if spectrum is None:
                spectrum = np.concatenate(spectra).reshape(-1)
            if morphology is None:
                detect, std = init.build_initialization_image(obs_list, spectra=spectra)
                morphology, bbox = init.extended_morphology(model_frame, sky_coord, detect, std, thresh=thresh, symmetric=True,
                                                            monotonic="flat", min_grad=0, boxsize=boxsize)
        center = np.asarray(model_frame.get_pixel(sky_coord), dtype=np.float64)
        morphology = np.asarray(morphology)
        if bbox is None:
            pix = np.round(center).astype(int)
            bbox = Box(morphology.shape, origin=(int(pix[0]) - morphology.shape[0] // 2, int(pix[1]) - morphology.shape[1] // 2))
        morph = ExtendedSourceMorphology(model_frame, center, morphology, bbox=bbox, monotonic=monotonic,
                                         symmetric=symmetric, min_grad=min_grad, shifting=shifting, resizing=resizing)
        spec = TabulatedSpectrum(model_frame, np.asarray(spectrum), min_step=_noise_rms(observations))
        super().__init__(model_frame, spec, morph)
        self.center = morph.center
