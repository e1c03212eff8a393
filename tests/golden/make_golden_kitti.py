"""Generates tests/golden/golden_kitti_v1.npz — known answers for `_correct_kitti_scan`
(python/kiss_icp/pybind/kiss_icp_pybind.cpp:127-138). Like golden_v1 these come from this project's oracle
(the reference cannot be built offline): PARITY UNPINNED.   Run:  python tests/golden/make_golden_kitti.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

rng = np.random.default_rng(20260923)
pts = (rng.normal(size=(509, 3)) * [30.0, 30.0, 2.0]).astype(np.float32).astype(np.float64)  # fp32-representable like KITTI .bin
pts = np.concatenate([pts, [[0.0, 0.0, 0.0], [0.0, 0.0, 7.5], [12.0, 0.0, 0.0]]])  # origin, on the axis, in the plane
path = os.path.join(ROOT, "tests", "golden", "golden_kitti_v1.npz")
np.savez_compressed(path, pts=pts, corrected=O.correct_kitti_scan(pts))
print("wrote", path, os.path.getsize(path), "bytes")
