"""ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product path.

ctypes front-end to ``oracle/libkiss_oracle.so`` (the dependency-free CPU restatement of the
reference's registration hot path, see oracle_core.hpp for the file:line citations).

Only ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py`` may
import this module. PARITY UNPINNED: the reference ships no golden vectors for this path and
cannot be built offline; this restatement is the project's definition of "reference output".
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libkiss_oracle.so")


def build(force: bool = False) -> str:
    """Compile the oracle with the recipe committed beside it (oracle/Makefile)."""
    srcs = [os.path.join(_HERE, f) for f in ("oracle_c_api.cpp", "oracle_core.hpp", "oracle_math.hpp", "oracle_robin.hpp")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_map_create.restype = C.c_void_p
        L.oracle_map_create.argtypes = [C.c_double, C.c_double, C.c_uint]
        L.oracle_pipeline_create.restype = C.c_void_p
        L.oracle_pipeline_create.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double,
                                             C.c_int, C.c_double, C.c_int, C.c_int]
        L.oracle_pipeline_map.restype = C.c_void_p
        L.oracle_pipeline_sigma.restype = C.c_double
        for name in ("oracle_map_num_voxels", "oracle_map_num_points", "oracle_map_pointcloud", "oracle_map_dump",
                     "oracle_voxel_downsample", "oracle_preprocess", "oracle_robin_trace"):
            getattr(L, name).restype = C.c_long
        L.oracle_voxel_hash.restype = C.c_uint
        _lib = L
    return _lib


def _a(x, dtype=np.float64):
    return np.ascontiguousarray(x, dtype=dtype)


def _p(arr):
    return arr.ctypes.data_as(C.c_void_p)


def _pts(x):
    x = _a(x)
    if x.ndim != 2 or x.shape[1] != 3:
        raise ValueError("points must be (N, 3)")
    return x


# ------------------------------------------------------------------ math
def se3_exp(a):
    M = np.empty((4, 4))
    lib().oracle_se3_exp(_p(_a(a)), _p(M))
    return M


def se3_log(M):
    a = np.empty(6)
    if lib().oracle_se3_log(_p(_a(M)), _p(a)):
        raise ValueError("not an SE(3) matrix")
    return a


def se3_mul(A, B):
    M = np.empty((4, 4))
    if lib().oracle_se3_mul(_p(_a(A)), _p(_a(B)), _p(M)):
        raise ValueError("not an SE(3) matrix")
    return M


def se3_inverse(A):
    M = np.empty((4, 4))
    if lib().oracle_se3_inverse(_p(_a(A)), _p(M)):
        raise ValueError("not an SE(3) matrix")
    return M


def se3_act(A, pts):
    pts = _pts(pts)
    out = np.empty_like(pts)
    if lib().oracle_se3_act(_p(_a(A)), _p(pts), C.c_long(len(pts)), _p(out)):
        raise ValueError("not an SE(3) matrix")
    return out


def ldlt6_solve(A, b):
    x = np.empty(6)
    lib().oracle_ldlt6_solve(_p(_a(A)), _p(_a(b)), _p(x))
    return x


# ------------------------------------------------------------------ voxel utils
def point_to_voxel(pts, voxel_size):
    pts = _pts(pts)
    out = np.empty((len(pts), 3), dtype=np.int32)
    lib().oracle_point_to_voxel(_p(pts), C.c_long(len(pts)), C.c_double(voxel_size), _p(out))
    return out


def voxel_hash(x, y, z):
    return int(lib().oracle_voxel_hash(C.c_int(int(x)), C.c_int(int(y)), C.c_int(int(z))))


def voxel_down_sample(pts, voxel_size):
    pts = _pts(pts)
    out = np.empty_like(pts)
    n = lib().oracle_voxel_downsample(_p(pts), C.c_long(len(pts)), C.c_double(voxel_size), _p(out))
    return out[:n].copy()


# ------------------------------------------------------------------ VoxelHashMap
class VoxelHashMap:
    def __init__(self, voxel_size, max_distance, max_points_per_voxel, _handle=None, _owner=None):
        self.voxel_size = float(voxel_size)
        self.max_distance = float(max_distance)
        self.max_points_per_voxel = int(max_points_per_voxel)
        self._owner = _owner  # keeps a pipeline alive when this is a borrowed view
        self._h = _handle or lib().oracle_map_create(C.c_double(voxel_size), C.c_double(max_distance),
                                                     C.c_uint(max_points_per_voxel))
        self._borrowed = _handle is not None

    def __del__(self):
        if getattr(self, "_h", None) and not self._borrowed and _lib is not None:
            _lib.oracle_map_destroy(C.c_void_p(self._h))
            self._h = None

    def clear(self):
        lib().oracle_map_clear(C.c_void_p(self._h))

    def empty(self):
        return bool(lib().oracle_map_empty(C.c_void_p(self._h)))

    def num_voxels(self):
        return int(lib().oracle_map_num_voxels(C.c_void_p(self._h)))

    def num_points(self):
        return int(lib().oracle_map_num_points(C.c_void_p(self._h)))

    def add_points(self, pts):
        pts = _pts(pts)
        lib().oracle_map_add_points(C.c_void_p(self._h), _p(pts), C.c_long(len(pts)))

    def remove_far_away_points(self, origin):
        lib().oracle_map_remove_far(C.c_void_p(self._h), _p(_a(origin)))

    def update(self, pts, pose_or_origin):
        pts = _pts(pts)
        x = _a(pose_or_origin)
        if x.shape == (4, 4):
            if lib().oracle_map_update_pose(C.c_void_p(self._h), _p(pts), C.c_long(len(pts)), _p(x)):
                raise ValueError("not an SE(3) matrix")
        else:
            lib().oracle_map_update_origin(C.c_void_p(self._h), _p(pts), C.c_long(len(pts)), _p(x))

    def point_cloud(self):
        out = np.empty((self.num_points(), 3))
        n = lib().oracle_map_pointcloud(C.c_void_p(self._h), _p(out))
        return out[:n]

    def dump(self):
        """(voxels (V,3) int32, counts (V,), points (P,3)) in reference iteration order."""
        nv, npnt = self.num_voxels(), self.num_points()
        vox = np.empty((nv, 3), dtype=np.int32)
        cnt = np.empty(nv, dtype=np.int32)
        pts = np.empty((npnt, 3))
        lib().oracle_map_dump(C.c_void_p(self._h), _p(vox), _p(cnt), _p(pts))
        return vox, cnt, pts

    def closest_neighbors(self, queries, nthreads=0):
        q = _pts(queries)
        outp = np.empty_like(q)
        outd = np.empty(len(q))
        lib().oracle_map_closest_neighbors(C.c_void_p(self._h), _p(q), C.c_long(len(q)), _p(outp), _p(outd),
                                           C.c_int(nthreads))
        return outp, outd


def build_system(vmap, pts_in_map_frame, max_dist, kernel, nthreads=1):
    pts = _pts(pts_in_map_frame)
    JTJ = np.empty((6, 6))
    JTr = np.empty(6)
    nc = C.c_int(0)
    lib().oracle_build_system(C.c_void_p(vmap._h), _p(pts), C.c_long(len(pts)), C.c_double(max_dist),
                              C.c_double(kernel), C.c_int(nthreads), _p(JTJ), _p(JTr), C.byref(nc))
    return JTJ, JTr, nc.value


def align_points_to_map(vmap, pts, guess, max_dist, kernel, max_iter=500, conv=1e-4, nthreads=0):
    pts = _pts(pts)
    out = np.empty((4, 4))
    it = C.c_int(0)
    if lib().oracle_align_points_to_map(C.c_void_p(vmap._h), _p(pts), C.c_long(len(pts)), _p(_a(guess)),
                                        C.c_double(max_dist), C.c_double(kernel), C.c_int(max_iter),
                                        C.c_double(conv), C.c_int(nthreads), _p(out), C.byref(it)):
        raise ValueError("not an SE(3) matrix")
    return out, it.value


def preprocess(pts, timestamps, motion, max_range, min_range, deskew):
    pts = _pts(pts)
    ts = _a(timestamps).reshape(-1)
    out = np.empty_like(pts)
    n = lib().oracle_preprocess(_p(pts), C.c_long(len(pts)), _p(ts), C.c_long(len(ts)), _p(_a(motion)),
                                C.c_double(max_range), C.c_double(min_range), C.c_int(int(deskew)), _p(out))
    if n == -1:
        raise ValueError("not an SE(3) matrix")
    if n == -2:
        raise IndexError("timestamps shorter than frame (std::out_of_range in the reference)")
    return out[:n].copy()


def correct_kitti_scan(pts):
    pts = _pts(pts)
    out = np.empty_like(pts)
    lib().oracle_correct_kitti_scan(_p(pts), C.c_long(len(pts)), _p(out))
    return out


def threshold_update(model_sse, num_samples, deviation, min_motion_th, max_range):
    sse = C.c_double(model_sse)
    ns = C.c_int(num_samples)
    if lib().oracle_threshold_update(C.byref(sse), C.byref(ns), _p(_a(deviation)), C.c_double(min_motion_th),
                                     C.c_double(max_range)):
        raise ValueError("not an SE(3) matrix")
    return sse.value, ns.value


class KissICP:
    """oracle twin of kiss_icp::pipeline::KissICP (pipeline/KissICP.hpp:56-96)."""

    def __init__(self, voxel_size=1.0, max_range=100.0, min_range=0.0, max_points_per_voxel=20, min_motion_th=0.1,
                 initial_threshold=2.0, max_num_iterations=500, convergence_criterion=1e-4, max_num_threads=0,
                 deskew=True):
        self._h = lib().oracle_pipeline_create(C.c_double(voxel_size), C.c_double(max_range), C.c_double(min_range),
                                               C.c_int(max_points_per_voxel), C.c_double(min_motion_th),
                                               C.c_double(initial_threshold), C.c_int(max_num_iterations),
                                               C.c_double(convergence_criterion), C.c_int(max_num_threads),
                                               C.c_int(int(deskew)))
        self._cfg = (voxel_size, max_range, max_points_per_voxel)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.oracle_pipeline_destroy(C.c_void_p(self._h))
            self._h = None

    def register_frame(self, frame, timestamps=(), want_clouds=True):
        pts = _pts(frame)
        ts = _a(timestamps).reshape(-1)
        npre, nsrc = C.c_long(0), C.c_long(0)
        rc = lib().oracle_pipeline_register_frame(C.c_void_p(self._h), _p(pts), C.c_long(len(pts)), _p(ts),
                                                  C.c_long(len(ts)), C.byref(npre), C.byref(nsrc))
        if rc:
            raise IndexError("timestamps shorter than frame (std::out_of_range in the reference)")
        if not want_clouds:
            return None, None
        pre = np.empty((npre.value, 3))
        src = np.empty((nsrc.value, 3))
        lib().oracle_pipeline_last_clouds(C.c_void_p(self._h), _p(pre), _p(src))
        return pre, src

    @property
    def pose(self):
        M = np.empty((4, 4))
        lib().oracle_pipeline_pose(C.c_void_p(self._h), _p(M))
        return M

    @property
    def delta(self):
        M = np.empty((4, 4))
        lib().oracle_pipeline_delta(C.c_void_p(self._h), _p(M))
        return M

    @property
    def sigma(self):
        return float(lib().oracle_pipeline_sigma(C.c_void_p(self._h)))

    @property
    def last_iterations(self):
        return int(lib().oracle_pipeline_last_iterations(C.c_void_p(self._h)))

    @property
    def local_map(self):
        vs, mr, mp = self._cfg
        return VoxelHashMap(vs, mr, mp, _handle=lib().oracle_pipeline_map(C.c_void_p(self._h)), _owner=self)


def num_threads():
    return int(lib().oracle_num_threads())
