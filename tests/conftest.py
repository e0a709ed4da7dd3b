import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_helpers():
    return np.load(os.path.join(GOLDEN, "helpers.npz"))


@pytest.fixture(scope="session")
def golden_icp_small():
    return np.load(os.path.join(GOLDEN, "icp_small.npz"))


@pytest.fixture(scope="session")
def golden_icp_full():
    return np.load(os.path.join(GOLDEN, "icp_full.npz"))


def pose_errors(T, T_ref):
    """(relative translation error, rotation geodesic angle [rad]) of a 4x4 pose vs a reference."""
    T = np.asarray(T, dtype=np.float64).reshape(4, 4)
    T_ref = np.asarray(T_ref, dtype=np.float64).reshape(4, 4)
    dt = np.linalg.norm(T[:3, 3] - T_ref[:3, 3]) / max(np.linalg.norm(T_ref[:3, 3]), 1e-12)
    dR = T_ref[:3, :3].T @ T[:3, :3]
    # atan2 form of the geodesic angle: arccos((tr-1)/2) turns 1e-7 entry noise of float32
    # (not exactly orthonormal) matrices into 3e-4 rad; the skew part does not.
    skew = 0.5 * np.array([dR[2, 1] - dR[1, 2], dR[0, 2] - dR[2, 0], dR[1, 0] - dR[0, 1]])
    ang = np.arctan2(np.linalg.norm(skew), (np.trace(dR) - 1.0) / 2.0)
    return dt, ang
