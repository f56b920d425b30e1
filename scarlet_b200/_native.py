"""ctypes binding of the C ABI in ``include/scarlet_b200.h``.

There is no CPU fallback: if the library is missing it is built with nvcc; if that fails, or no CUDA
device is present when a compute entry point is called, an error is raised.
"""
import ctypes as C
import os

import numpy as np

from . import _build

SB_MAX_CHAIN_OPS = 8
SB_MAX_CHANNELS = 16
SB_MAX_OBS = 4
SB_N_STAGES = 10
SB_ERR_NONFINITE = -4
SCENE_RUN, SCENE_CONVERGED, SCENE_PAUSED, SCENE_EXHAUSTED, SCENE_FAILED, SCENE_CONV_PENDING = 0, 1, 2, 4, 8, 16

OP_MONOTONIC, OP_SYMMETRY, OP_POSITIVITY, OP_CENTER_ON, OP_NORMALIZE = 1, 2, 3, 4, 5


class sb_op(C.Structure):
    _fields_ = [("code", C.c_int32), ("iarg", C.c_int32), ("farg", C.c_double)]


class sb_chain_desc(C.Structure):
    _fields_ = [("n_ops", C.c_int32), ("repeat", C.c_int32), ("ops", sb_op * SB_MAX_CHAIN_OPS)]


class sb_mono_desc(C.Structure):
    _fields_ = [("n_pix", C.c_int32), ("n_off", C.c_int32), ("n_idx", C.c_int32), ("_pad", C.c_int32),
                ("weights", C.c_void_p), ("offsets", C.c_void_p), ("dist_idx", C.c_void_p)]


class sb_obs_desc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("chan_off", C.c_int32),
                ("oy", C.c_int32), ("ox", C.c_int32), ("Fy", C.c_int32), ("Fx", C.c_int32), ("khat_shared", C.c_int32),
                ("psf_shift", C.c_int32), ("shift_Fy", C.c_int32), ("shift_Fx", C.c_int32), ("shift_fixed", C.c_int32),
                ("shift_step", C.c_double)]


class sb_source_desc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("By", C.c_int32), ("Bx", C.c_int32), ("oy", C.c_int32), ("ox", C.c_int32),
                ("chain", C.c_int32), ("sed_chain", C.c_int32), ("sed_is_f32", C.c_int32), ("morph_fixed", C.c_int32),
                ("sed_fixed", C.c_int32), ("shifting", C.c_int32), ("shift_Fy", C.c_int32), ("shift_Fx", C.c_int32),
                ("resizing", C.c_int32), ("shift_step", C.c_double), ("morph_step", C.c_double),
                ("sed_step_factor", C.c_double), ("sed_step_min", C.c_double * SB_MAX_CHANNELS)]


class sb_batch_desc(C.Structure):
    _fields_ = [("precision", C.c_int32), ("n_scenes", C.c_int32), ("C", C.c_int32), ("Ny", C.c_int32), ("Nx", C.c_int32),
                ("n_obs", C.c_int32), ("obs", sb_obs_desc * SB_MAX_OBS), ("n_sources", C.c_int32), ("n_chains", C.c_int32),
                ("n_mono", C.c_int32), ("psf_boxsize", C.c_int32), ("psf_sigma", C.c_double * SB_MAX_CHANNELS),
                ("scene_src_start", C.c_void_p), ("sources", C.c_void_p), ("chains", C.c_void_p), ("mono", C.c_void_p)]


class sb_fit_opts(C.Structure):
    _fields_ = [("max_iter", C.c_int32), ("min_iter", C.c_int32), ("prox_max_iter", C.c_int32), ("check_every", C.c_int32),
                ("fixed_iterations", C.c_int32), ("overwrite_vhat_at_it0", C.c_int32), ("resume", C.c_int32),
                ("run_until", C.c_int32), ("e_rel", C.c_double),
                ("b1", C.c_double), ("b2", C.c_double), ("eps", C.c_double), ("pause_every", C.c_int32), ("_pad1", C.c_int32)]


# every symbol declared in include/scarlet_b200.h: (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "sb_last_error": (C.c_char_p, []),
    "sb_device_count": (C.c_int, []),
    "sb_version": (C.c_char_p, []),
    "sb_source_hash": (C.c_char_p, []),
    "sb_fft_supported_length": (C.c_int, [C.c_int]),
    "sb_plan_spectral_mode": (C.c_int, [_P]),
    "sb_plan_prox_histogram": (C.c_int, [_P, C.c_int, _P]),
    "sb_plan_scene_control": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "sb_plan_scene_status": (C.c_int, [_P, _P, _P, _P]),
    "sb_plan_run": (C.c_int, [_P, C.POINTER(sb_fit_opts), C.c_int, _P]),
    "sb_plan_download_loss": (C.c_int, [_P, _P, C.c_int]),
    "sb_plan_inspect": (C.c_int, [_P, _P]),
    "sb_plan_set_sources": (C.c_int, [_P, C.POINTER(sb_batch_desc)]),
    "sb_plan_upload_loss": (C.c_int, [_P, _P, C.c_int]),
    "sb_plan_create": (C.c_int, [C.POINTER(sb_batch_desc), C.c_int, C.POINTER(_P)]),
    "sb_plan_destroy": (None, [_P]),
    "sb_plan_device_bytes": (C.c_int64, [_P]),
    "sb_plan_upload_observation": (C.c_int, [_P, C.c_int, _P, _P, _P, _P]),
    "sb_plan_upload_observation_f64": (C.c_int, [_P, C.c_int, _P, _P, _P, _P]),
    "sb_plan_upload_kernels": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int]),
    "sb_plan_zero_state": (C.c_int, [_P]),
    "sb_plan_upload_resampling": (C.c_int, [_P, C.c_int, _P, _P, C.c_double]),
    "sb_plan_upload_resampling_rot": (C.c_int, [_P, C.c_int, _P, _P, C.c_double]),
    "sb_host_gather_f64": (C.c_int, [_P, _P, _P, _P, C.c_int64]),
    "sb_host_scatter_f64": (C.c_int, [_P, _P, _P, _P, C.c_int64]),
    "sb_host_alloc": (_P, [C.c_int64]),
    "sb_host_free": (None, [_P]),
    "sb_plan_upload_params": (C.c_int, [_P, C.c_int, _P, _P, _P]),
    "sb_plan_download_params": (C.c_int, [_P, C.c_int, _P, _P, _P]),
    "sb_plan_evaluate": (C.c_int, [_P, C.c_int, _P, _P, _P, _P, _P, _P]),
    "sb_plan_fit": (C.c_int, [_P, C.POINTER(sb_fit_opts), _P, _P, _P]),
    "sb_plan_fit_enqueue": (C.c_int, [_P, C.POINTER(sb_fit_opts), C.c_int]),
    "sb_plan_sync": (C.c_int, [_P]),
    "sb_plan_timer_start": (C.c_int, [_P]),
    "sb_plan_timer_stop": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "sb_plan_profile_iterations": (C.c_int, [_P, C.POINTER(sb_fit_opts), C.c_int, _P]),
    "sb_stage_name": (C.c_char_p, [C.c_int]),
    "sb_plan_kernel_launches": (C.c_int64, [_P]),
    "sb_plan_stream": (_P, [_P]),
    "sb_plan_device_params": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_int64), C.POINTER(_P), C.POINTER(C.c_int64),
                                        C.POINTER(C.c_int)]),
    "sb_monotonic_f32": (C.c_int, [_P, _P, _P, C.c_int, _P, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int]),
    "sb_monotonic_f64": (C.c_int, [_P, _P, _P, C.c_int, _P, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int]),
    "sb_prox_chain_f32": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.POINTER(sb_chain_desc), _P, C.c_int, C.c_int]),
    "sb_prox_chain_f64": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.POINTER(sb_chain_desc), _P, C.c_int, C.c_int]),
    "sb_apply_filter_f32": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, _P, _P, C.c_int, _P, C.c_int]),
    "sb_apply_filter_f64": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, _P, _P, C.c_int, _P, C.c_int]),
    "sb_fft_convolve_f32": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.c_int, C.c_int, _P, C.c_int]),
    "sb_fft_convolve_f64": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, C.c_int, C.c_int, C.c_int, _P, C.c_int]),
}

_lib = None


class NativeError(RuntimeError):
    """An entry point of libscarlet_b200 returned an error code."""


def lib():
    """Load (building if necessary) the CUDA library.  Raises if it cannot be had -- there is no fallback."""
    global _lib
    if _lib is None:
        path = _build.LIB
        if not _build.up_to_date():  # missing, or built from other sources than the ones next to it: never load a stale ABI
            path = _build.build()
        handle = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)  # AttributeError if the library does not export what the header declares
            fn.restype = res
            fn.argtypes = args
        built = handle.sb_source_hash().decode()
        if built != _build.source_hash():
            raise NativeError("libscarlet_b200.so was built from other sources (%s) than csrc/ (%s): rebuild with "
                              "python -m scarlet_b200._build" % (built, _build.source_hash()))
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().sb_last_error().decode("utf-8", "replace")
        raise NativeError("scarlet_b200 native call failed (%d): %s" % (rc, msg))


def ptr(a):
    """Pointer to a C-contiguous ndarray (None -> NULL)."""
    if a is None:
        return None
    assert a.flags.c_contiguous
    return a.ctypes.data


def default_device():
    return int(os.environ.get("LOCAL_RANK", "0")) if lib().sb_device_count() > 1 else 0


def fit_opts(max_iter=200, e_rel=1e-3, min_iter=1, prox_max_iter=10, check_every=10, fixed_iterations=False,
             b1=0.9, b2=0.999, eps=1e-8, overwrite_vhat_at_it0=True, resume=False, run_until=0, pause_every=0):
    return sb_fit_opts(int(max_iter), int(min_iter), int(prox_max_iter), int(check_every), int(bool(fixed_iterations)),
                       int(bool(overwrite_vhat_at_it0)), int(bool(resume)), int(run_until), float(e_rel), float(b1), float(b2),
                       float(eps), int(pause_every), 0)


def as_array(x, dtype):
    return np.ascontiguousarray(x, dtype=dtype)
