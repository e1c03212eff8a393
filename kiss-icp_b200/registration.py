"""Registration — python/kiss_icp/registration.py:36-65 surface over kb_registration_*."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N
from .config import KISSConfig
from .mapping import VoxelHashMap


def get_registration(config: KISSConfig):
    return Registration(
        max_num_iterations=config.registration.max_num_iterations,
        convergence_criterion=config.registration.convergence_criterion,
        max_num_threads=config.registration.max_num_threads,
    )


class Registration:
    def __init__(self, max_num_iterations: int, convergence_criterion: float, max_num_threads: int = 0):
        self._h = N.vp()
        N.check(N.lib().kb_registration_create(int(max_num_iterations), float(convergence_criterion),
                                               int(max_num_threads), C.byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and N._lib is not None:
            N._lib.kb_registration_destroy(h)
            self._h = None

    def align_points_to_map(self, points: np.ndarray, voxel_map: VoxelHashMap, initial_guess: np.ndarray,
                            max_correspondance_distance: float, kernel: float) -> np.ndarray:
        pts = N.points_arg(points)
        g = N.mat4_arg(initial_guess)
        out = np.empty((4, 4))
        N.check(N.lib().kb_registration_align_points_to_map(self._h, N.ptr(pts), len(pts), voxel_map._h, N.ptr(g),
                                                            float(max_correspondance_distance), float(kernel), N.ptr(out)))
        return out

    @property
    def last_iterations(self) -> int:
        it = N.i32(0)
        N.check(N.lib().kb_registration_last_iterations(self._h, C.byref(it)))
        return it.value

    def build_system(self, points_in_map_frame, voxel_map: VoxelHashMap, max_correspondance_distance: float, kernel: float):
        """one DataAssociation + BuildLinearSystem pass (Registration.cpp:60-121): (JTJ, JTr, n)."""
        pts = N.points_arg(points_in_map_frame)
        JTJ = np.empty((6, 6))
        JTr = np.empty(6)
        nc = N.i32(0)
        N.check(N.lib().kb_registration_build_system(self._h, N.ptr(pts), len(pts), voxel_map._h,
                                                     float(max_correspondance_distance), float(kernel), N.ptr(JTJ),
                                                     N.ptr(JTr), C.byref(nc)))
        return JTJ, JTr, nc.value
