"""Preprocessor — python/kiss_icp/preprocess.py:30-51 surface over kb_preprocessor_*."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N
from .config import KISSConfig


def get_preprocessor(config: KISSConfig):
    return Preprocessor(
        max_range=config.data.max_range,
        min_range=config.data.min_range,
        deskew=config.data.deskew,
        max_num_threads=config.registration.max_num_threads,
    )


class Preprocessor:
    def __init__(self, max_range, min_range, deskew, max_num_threads):
        self._h = N.vp()
        N.check(N.lib().kb_preprocessor_create(float(max_range), float(min_range), int(bool(deskew)),
                                               int(max_num_threads), C.byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and N._lib is not None:
            N._lib.kb_preprocessor_destroy(h)
            self._h = None

    def preprocess(self, frame: np.ndarray, timestamps: np.ndarray, relative_motion: np.ndarray):
        pts = N.points_arg(frame)
        ts = np.ascontiguousarray(np.asarray(timestamps).ravel(), dtype=np.float64)
        T = N.mat4_arg(relative_motion)
        out = np.empty_like(pts)
        n = N.sz(0)
        N.check(N.lib().kb_preprocessor_preprocess(self._h, N.ptr(pts), len(pts), N.ptr(ts), len(ts), N.ptr(T),
                                                   N.ptr(out), len(pts), C.byref(n)))
        return out[: n.value]


def correct_kitti_scan(frame: np.ndarray) -> np.ndarray:
    """`kiss_icp_pybind._correct_kitti_scan` (kiss_icp_pybind.cpp:127-138; used by datasets/kitti.py:44-48,68):
    the KITTI-only intrinsic correction, every point rotated by 0.205 deg about normalized(pt x e_z)."""
    pts = N.points_arg(frame)
    out = np.empty_like(pts)
    N.check(N.lib().kb_correct_kitti_scan(N.ptr(pts), len(pts), N.ptr(out)))
    return out
