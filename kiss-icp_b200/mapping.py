"""VoxelHashMap — same Python surface as python/kiss_icp/mapping.py:37-68, backed by the HBM
voxel table (C-ABI kb_map_*)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N
from .config import KISSConfig


def get_voxel_hash_map(config: KISSConfig):
    return VoxelHashMap(
        voxel_size=config.mapping.voxel_size,
        max_distance=config.data.max_range,
        max_points_per_voxel=config.mapping.max_points_per_voxel,
    )


class VoxelHashMap:
    def __init__(self, voxel_size: float, max_distance: float, max_points_per_voxel: int, _borrowed=None, _owner=None):
        self._owner = _owner
        self._borrowed = _borrowed is not None
        if self._borrowed:
            self._h = N.vp(_borrowed)
        else:
            self._h = N.vp()
            N.check(N.lib().kb_map_create(float(voxel_size), float(max_distance), int(max_points_per_voxel), C.byref(self._h)))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and not self._borrowed and N._lib is not None:
            N._lib.kb_map_destroy(h)
            self._h = None

    # -- reference surface ---------------------------------------------------------------------
    def clear(self):
        N.check(N.lib().kb_map_clear(self._h))

    def empty(self):
        e = N.i32(0)
        N.check(N.lib().kb_map_empty(self._h, C.byref(e)))
        return bool(e.value)

    def update(self, points: np.ndarray, pose: np.ndarray = np.eye(4)):
        """points + (4,4) pose  -> VoxelHashMap::Update(points, pose)   (VoxelHashMap.cpp:89-95)
        points + (3,) origin    -> VoxelHashMap::Update(points, origin) (VoxelHashMap.cpp:83-87)"""
        pts = N.points_arg(points)
        x = np.ascontiguousarray(pose, dtype=np.float64)
        if x.shape == (4, 4):
            N.check(N.lib().kb_map_update_pose(self._h, N.ptr(pts), len(pts), N.ptr(x)))
        elif x.shape == (3,):
            N.check(N.lib().kb_map_update_origin(self._h, N.ptr(pts), len(pts), N.ptr(x)))
        else:
            raise TypeError("_update(): incompatible function arguments (pose (4,4) or origin (3,))")

    def add_points(self, points):
        pts = N.points_arg(points)
        N.check(N.lib().kb_map_add_points(self._h, N.ptr(pts), len(pts)))

    def remove_far_away_points(self, origin):
        o = np.ascontiguousarray(origin, dtype=np.float64).reshape(3)
        N.check(N.lib().kb_map_remove_far(self._h, N.ptr(o)))

    def point_cloud(self) -> np.ndarray:
        n = N.sz(0)
        N.check(N.lib().kb_map_pointcloud(self._h, None, 0, C.byref(n)))
        out = np.empty((n.value, 3))
        if n.value:
            N.check(N.lib().kb_map_pointcloud(self._h, N.ptr(out), n.value, C.byref(n)))
        return out

    # -- extras (no reference counterpart in Python; C++ has GetClosestNeighbor) ----------------
    def compact(self):
        """rebuild the table at the smallest capacity (load <= 0.5, no tombstones)"""
        N.check(N.lib().kb_map_compact(self._h))

    def num_points(self) -> int:
        n = N.sz(0)
        N.check(N.lib().kb_map_num_points(self._h, C.byref(n)))
        return n.value

    def num_voxels(self) -> int:
        n = N.sz(0)
        N.check(N.lib().kb_map_num_voxels(self._h, C.byref(n)))
        return n.value

    def dump(self):
        """(voxels (V,3) int32 ascending, counts (V,), points (P,3) per-voxel insertion order)."""
        nv, npnt = self.num_voxels(), self.num_points()
        vox = np.empty((nv, 3), dtype=np.int32)
        cnt = np.empty(nv, dtype=np.int32)
        pts = np.empty((npnt, 3))
        a, b = N.sz(0), N.sz(0)
        N.check(N.lib().kb_map_dump(self._h, N.ptr(vox), N.ptr(cnt), N.ptr(pts), nv, npnt, C.byref(a), C.byref(b)))
        return vox, cnt, pts

    def closest_neighbors(self, queries):
        """batched VoxelHashMap::GetClosestNeighbor (VoxelHashMap.cpp:46-70)."""
        q = N.points_arg(queries)
        outp = np.empty_like(q)
        outd = np.empty(len(q))
        N.check(N.lib().kb_map_closest_neighbors(self._h, N.ptr(q), len(q), N.ptr(outp), N.ptr(outd)))
        return outp, outd
