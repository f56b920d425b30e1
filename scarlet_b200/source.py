"""Source recipes.  Mirrors the constructors of scarlet/source.py ``PointSource`` 92-128 and
``SingleExtendedSource`` 367-450 for the part that is ON the fitting path: parameter layout, step sizes
(spectrum steps floored by the per-band noise rms, source.py:412-416) and constraint chains.

Initialisation of spectra / morphologies from the data (scarlet/initialization.py) happens when
``spectrum`` / ``morphology`` are not passed in: peak-pixel spectrum, symmetric + monotonic cut-out of the
spectrum-weighted detection image.  ``ExtendedSource`` is the reference's factory (source.py:759-807): compact
(point-source shaped), single- or multi-component (``K`` layers split at flux percentiles) extended sources.
"""
import numpy as np

import logging

from .bbox import Box
from .component import CombinedComponent, FactorizedComponent
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


logger = logging.getLogger("scarlet_b200.source")


class SingleExtendedSource(FactorizedComponent):
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


class CompactExtendedSource(FactorizedComponent):
    """Extended-source model started from the shape of the model PSF, spectrum from the PSF-corrected peak pixel
    (source.py:249-363)."""

    def __init__(self, model_frame, sky_coord, observations, shifting=False, resizing=True, boxsize=None):
        from . import initialization as init
        assert model_frame.psf is not None
        obs_list = observations if hasattr(observations, "__iter__") else (observations,)
        morph, bbox = init.compact_morphology(model_frame, sky_coord, boxsize=boxsize)
        center = np.asarray(model_frame.get_pixel(sky_coord), dtype=np.float64)
        morphology = ExtendedSourceMorphology(model_frame, center, morph, bbox=bbox, monotonic="angle", symmetric=False, min_grad=0,
                                              shifting=shifting, resizing=resizing)
        spectrum = init.get_pixel_spectrum(sky_coord, obs_list, correct_psf=True)
        spectrum /= morph.sum()  # in place: keeps the dtype of the data, like the reference
        super().__init__(model_frame, TabulatedSpectrum(model_frame, spectrum, min_step=_noise_rms(obs_list)), morphology)
        self.center = morphology.center


class MultiExtendedSource(CombinedComponent):
    """K stacked components: the single-component morphology is cut at flux percentiles of its peak; every layer above a cut
    starts from the part of the profile exceeding it, every layer below is flattened there; all layers start with the
    peak-pixel spectrum (source.py:615-746)."""

    def __init__(self, model_frame, sky_coord, observations, K=2, flux_percentiles=None, thresh=1.0, shifting=False, resizing=True,
                 boxsize=None):
        if flux_percentiles is None:
            flux_percentiles = (25,)
        assert K == len(flux_percentiles) + 1
        obs_list = observations if hasattr(observations, "__iter__") else (observations,)
        base = SingleExtendedSource(model_frame, sky_coord, obs_list, thresh=thresh, boxsize=boxsize)
        spectrum = base.children[0].parameters[0]._data
        morphs, boxes = self.init_morphs(base.children[1], flux_percentiles, sky_coord)
        center = np.asarray(model_frame.get_pixel(sky_coord), dtype=np.float64)
        noise_rms = _noise_rms(obs_list)
        components = []
        for k in range(K):
            spec = TabulatedSpectrum(model_frame, spectrum.copy(), min_step=noise_rms / 10)
            morph = ExtendedSourceMorphology(model_frame, center, morphs[k], bbox=boxes[k], monotonic="angle", symmetric=False, min_grad=0,
                                             shifting=shifting, resizing=resizing)
            self.center = morph.center
            components.append(FactorizedComponent(model_frame, spec, morph))
        super().__init__(components)

    @staticmethod
    def init_morphs(morphology, flux_percentiles, sky_coord=None):
        morph = np.asarray(morphology.get_model())
        K = len(flux_percentiles) + 1
        layers = np.zeros((K,) + morph.shape, dtype=morph.dtype)
        layers[0] = morph
        peak, below = morph.max(), 0
        for k, perc in enumerate(np.sort(flux_percentiles), start=1):
            cut = perc * peak / 100
            above = morph > cut
            layers[k - 1][above] = cut - below
            layers[k][above] = morph[above] - cut
            below = cut
        for k in range(K):
            if np.all(layers[k] <= 0):
                logger.warning("Zero or negative morphology for component %d at %s", k, sky_coord)
            layers[k] /= layers[k].max()
        return layers, tuple(morphology.bbox.copy() for _ in range(K))


def ExtendedSource(model_frame, sky_coord, observations, K=1, flux_percentiles=None, thresh=1.0, compact=False, shifting=False,
                   resizing=True, boxsize=None, **single_kwargs):
    """The reference's factory (source.py:759-807): compact, single- or multi-component extended source.  Extra keywords
    (``spectrum=, morphology=, bbox=, monotonic=, symmetric=, min_grad=``) go to the single-component recipe and let a
    caller start from explicit parameters instead of the data."""
    if compact:
        return CompactExtendedSource(model_frame, sky_coord, observations, shifting=shifting, resizing=resizing, boxsize=boxsize)
    if K == 1:
        return SingleExtendedSource(model_frame, sky_coord, observations, thresh=thresh, shifting=shifting, resizing=resizing,
                                    boxsize=boxsize, **single_kwargs)
    return MultiExtendedSource(model_frame, sky_coord, observations, K=K, flux_percentiles=flux_percentiles, thresh=thresh,
                               shifting=shifting, resizing=resizing, boxsize=boxsize)
