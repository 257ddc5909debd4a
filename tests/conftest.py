import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def K():
    from tools import synth
    return synth.K_DEFAULT.copy()


@pytest.fixture(scope="session")
def frames():
    """First 12 synthetic frames along the fr1/plant path: (depth [n,480,640], R [n,3,3], t [n,3])."""
    from tools import synth
    return synth.render_sequence(12)


@pytest.fixture(scope="session")
def gpu_lib():
    import tracking_sdf_b200 as T
    L = T.load_library()          # raises if not built: no fallback
    if L.tsdf_device_count() < 1:
        pytest.fail("GPU test selected but no CUDA device is visible")
    return L


def rot_angle(Ra, Rb):
    dR = np.asarray(Ra) @ np.asarray(Rb).T
    return float(np.arccos(np.clip((np.trace(dR) - 1.0) / 2.0, -1.0, 1.0)))
