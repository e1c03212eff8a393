import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        from kiss_icp_b200 import _native
        return _native.lib().kb_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def O():
    """the oracle (CPU restatement of the reference) — checker only"""
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def K():
    import kiss_icp_b200
    return kiss_icp_b200


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    return np.load(path)


def canon_map(vox, cnt, pts):
    """voxel dump -> ascending (x,y,z) voxel order, per-voxel point order preserved"""
    order = np.lexsort((vox[:, 2], vox[:, 1], vox[:, 0]))
    starts = np.concatenate([[0], np.cumsum(cnt)])[:-1]
    if len(order) == 0:
        return vox, cnt, pts
    return vox[order], cnt[order], np.concatenate([pts[starts[i]:starts[i] + cnt[i]] for i in order])


def pose_error(A, B):
    """(translation error [m], rotation error [rad]) between two 4x4 poses"""
    dt = float(np.linalg.norm(A[:3, 3] - B[:3, 3]))
    R = A[:3, :3].T @ B[:3, :3]
    dr = float(np.arccos(np.clip((np.trace(R) - 1.0) / 2.0, -1.0, 1.0)))
    return dt, dr
