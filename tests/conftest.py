import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """Tests marked ``gpu`` are skipped (not failed) on a machine without a CUDA device."""
    gpu_items = [it for it in items if "gpu" in it.keywords]
    if not gpu_items:
        return
    try:
        from scarlet_b200 import _native
        have = _native.lib().sb_device_count() > 0
    except Exception:
        have = False
    if not have:
        skip = pytest.mark.skip(reason="no CUDA device: scarlet_b200 has no CPU fallback")
        for it in gpu_items:
            it.add_marker(skip)


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def has_gpu():
    from scarlet_b200 import _native
    return _native.lib().sb_device_count() > 0


# known-answer vectors of the reference's own tests (tests/test_constraint.py:93-163), restated
ARANGE25 = np.arange(25, dtype=float).reshape(5, 5)
MONO_NEAREST_0 = np.array([[0, 1, 2, 3, 4], [5, 6, 7, 8, 9], [10, 11, 12, 12, 12], [11, 12, 12, 12, 12], [12, 12, 12, 12, 12]], float)
MONO_ANGLE_0 = np.array([
    [0.000000000, 1.000000000, 2.000000000, 3.000000000, 4.000000000],
    [5.000000000, 6.000000000, 7.000000000, 8.000000000, 9.000000000],
    [9.742640687, 11.000000000, 12.000000000, 12.000000000, 10.828427125],
    [11.030627697, 11.707106781, 12.000000000, 12.000000000, 11.771236166],
    [11.556349186, 11.868867239, 11.914213562, 11.983249156, 11.928090416]])
MONO_ANGLE_025 = np.array([
    [0.000000000, 1.000000000, 2.000000000, 3.000000000, 4.000000000],
    [5.000000000, 6.000000000, 7.000000000, 7.242640687, 5.806841831],
    [5.801461031, 9.000000000, 12.000000000, 9.000000000, 6.074431804],
    [5.895545844, 7.681980515, 9.000000000, 7.681980515, 5.935521488],
    [4.988519641, 5.949655012, 6.170941546, 5.949655012, 4.997301087]])
SYM_HALF = np.array([[6.0, 6.5, 7.0, 7.5, 8.0], [8.5, 9.0, 9.5, 10.0, 10.5], [11.0, 11.5, 12.0, 12.5, 13.0],
                     [13.5, 14.0, 14.5, 15.0, 15.5], [16.0, 16.5, 17.0, 17.5, 18.0]])
MONO_KATS = [("nearest", 0.0, MONO_NEAREST_0), ("angle", 0.0, MONO_ANGLE_0), ("angle", 0.25, MONO_ANGLE_025)]
