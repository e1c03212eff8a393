"""KissICP — same surface as python/kiss_icp/kiss_icp.py:33-80.

Two execution modes:
  * fused (default): ``register_frame`` is ONE call into kb_pipeline_register_frame, i.e. one
    persistent CUDA kernel doing kiss_icp::pipeline::KissICP::RegisterFrame
    (cpp/kiss_icp/pipeline/KissICP.cpp:35-68) with all state resident in HBM.
  * modular (``fused=False`` or a fixed threshold, which only the Python reference has): the
    reference's own module-by-module sequence (kiss_icp.py:43-75) through the per-module C-ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N
from .config import KISSConfig, load_config
from .mapping import VoxelHashMap, get_voxel_hash_map
from .preprocess import get_preprocessor
from .registration import get_registration
from .threshold import get_threshold_estimator
from .voxelization import voxel_down_sample


class _PipelineThreshold:
    """read-only view of the device-resident adaptive threshold of a fused pipeline"""

    def __init__(self, owner):
        self._owner = owner

    def get_threshold(self):
        return self._owner._sigma_next()


class KissICP:
    def __init__(self, config: KISSConfig | None = None, fused: bool = True):
        self.config = config if config is not None else load_config()
        if self.config.mapping.voxel_size is None:
            self.config.mapping.voxel_size = float(self.config.data.max_range / 100.0)
        self.fused = bool(fused) and self.config.adaptive_threshold.fixed_threshold is None
        self._h = None
        if self.fused:
            c = N.Config()
            c.voxel_size = self.config.mapping.voxel_size
            c.max_range = self.config.data.max_range
            c.min_range = self.config.data.min_range
            c.max_points_per_voxel = self.config.mapping.max_points_per_voxel
            c.min_motion_th = self.config.adaptive_threshold.min_motion_th
            c.initial_threshold = self.config.adaptive_threshold.initial_threshold
            c.max_num_iterations = self.config.registration.max_num_iterations
            c.convergence_criterion = self.config.registration.convergence_criterion
            c.max_num_threads = self.config.registration.max_num_threads
            c.deskew = int(self.config.data.deskew)
            self._h = N.vp()
            N.check(N.lib().kb_pipeline_create(C.byref(c), C.byref(self._h)))
            self.local_map = VoxelHashMap(c.voxel_size, c.max_range, c.max_points_per_voxel,
                                          _borrowed=N.lib().kb_pipeline_voxel_map(self._h), _owner=self)
            self.adaptive_threshold = _PipelineThreshold(self)
        else:
            self._last_pose = np.eye(4)
            self._last_delta = np.eye(4)
            self.adaptive_threshold = get_threshold_estimator(self.config)
            self.preprocessor = get_preprocessor(self.config)
            self.registration = get_registration(self.config)
            self.local_map = get_voxel_hash_map(self.config)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h and N._lib is not None:
            N._lib.kb_pipeline_destroy(h)
            self._h = None

    # -- pose()/delta() (KissICP.hpp:81-85) -----------------------------------------------------
    @property
    def last_pose(self):
        if not self.fused:
            return self._last_pose
        M = np.empty((4, 4))
        N.check(N.lib().kb_pipeline_pose(self._h, N.ptr(M)))
        return M

    @last_pose.setter
    def last_pose(self, T):
        if not self.fused:
            self._last_pose = np.array(T, dtype=np.float64)
        else:
            N.check(N.lib().kb_pipeline_set_pose(self._h, N.ptr(N.mat4_arg(T))))

    @property
    def last_delta(self):
        if not self.fused:
            return self._last_delta
        M = np.empty((4, 4))
        N.check(N.lib().kb_pipeline_delta(self._h, N.ptr(M)))
        return M

    @last_delta.setter
    def last_delta(self, T):
        if not self.fused:
            self._last_delta = np.array(T, dtype=np.float64)
        else:
            N.check(N.lib().kb_pipeline_set_delta(self._h, N.ptr(N.mat4_arg(T))))

    def _sigma_next(self):
        s = N.dbl(0)
        N.check(N.lib().kb_pipeline_threshold(self._h, C.byref(s)))
        return s.value

    @property
    def last_sigma(self):
        s = N.dbl(0)
        N.check(N.lib().kb_pipeline_last_sigma(self._h, C.byref(s)))
        return s.value

    @property
    def last_iterations(self):
        if not self.fused:
            return self.registration.last_iterations
        it = N.i32(0)
        N.check(N.lib().kb_pipeline_last_iterations(self._h, C.byref(it)))
        return it.value

    @property
    def last_profile_us(self):
        """device-side phase durations of the last fused RegisterFrame [us]:
        preprocess, downsample 0.5v, downsample 1.5v, ICP, map update, epilogue"""
        us = np.zeros(6)
        N.check(N.lib().kb_pipeline_last_profile(self._h, N.ptr(us), 6))
        return us

    def set_profiling(self, enabled: bool):
        """in-kernel phase timestamps (last_profile_us, history phase_us): off by default, ~10 us/scan when on"""
        N.check(N.lib().kb_pipeline_set_profiling(self._h, int(bool(enabled))))

    def grow_retries(self) -> int:
        """frames the kernel vetoed because the voxel table was sized too optimistically (grown + replayed)"""
        v = C.c_ulonglong(0)
        N.check(N.lib().kb_pipeline_grow_retries(self._h, C.byref(v)))
        return int(v.value)

    def start_history(self, capacity: int):
        """log per-frame statistics of the next ``capacity`` fused RegisterFrame calls (host side)"""
        N.check(N.lib().kb_pipeline_set_history(self._h, int(capacity)))

    def history(self):
        n = N.sz(0)
        N.check(N.lib().kb_pipeline_get_history(self._h, None, 0, C.byref(n)))
        arr = (N.FrameStats * n.value)()
        if n.value:
            N.check(N.lib().kb_pipeline_get_history(self._h, arr, n.value, C.byref(n)))
        return list(arr)

    # -- RegisterFrame ---------------------------------------------------------------------------
    def register_frame(self, frame, timestamps, return_clouds: bool = True):
        """-> (preprocessed frame, source) like the reference; ``return_clouds=False`` skips the
        two device-to-host cloud copies (the pose is available as ``last_pose`` either way)."""
        if not self.fused:
            return self._register_frame_modular(frame, timestamps)
        ts = np.ascontiguousarray(np.asarray(timestamps).ravel(), dtype=np.float64)
        f = np.asarray(frame)
        if f.dtype == np.float32 and f.ndim == 2 and f.shape[1] == 3:
            # float32 clouds (KITTI .bin, PointCloud2) go to the device as they are: half the H2D bytes,
            # widened exactly on the device (the reference does .astype(np.float64) on the host)
            pts = np.ascontiguousarray(f)
            N.check(N.lib().kb_pipeline_register_frame_f32(self._h, N.ptr(pts), len(pts), N.ptr(ts), len(ts)))
        else:
            pts = N.points_arg(frame)
            N.check(N.lib().kb_pipeline_register_frame(self._h, N.ptr(pts), len(pts), N.ptr(ts), len(ts)))
        if not return_clouds:
            return None, None
        a, b = N.sz(0), N.sz(0)
        N.check(N.lib().kb_pipeline_last_cloud_sizes(self._h, C.byref(a), C.byref(b)))
        pre = np.empty((a.value, 3))
        src = np.empty((b.value, 3))
        N.check(N.lib().kb_pipeline_last_clouds(self._h, N.ptr(pre), a.value, N.ptr(src), b.value))
        return pre, src

    def register_frames(self, frames, timestamps=None):
        """Register a whole sequence (the dataset loop of python/kiss_icp/pipeline.py:106-112) -> poses (K,4,4).

        Same results as K ``register_frame`` calls; the frames are queued on the device (copy of frame k+1
        overlaps the registration of frame k, results are read behind the queue). ``frames`` is a sequence of
        (N_k,3) arrays, all float64 or all float32; ``timestamps`` a matching sequence (or None)."""
        K = len(frames)
        if timestamps is None:
            timestamps = [np.empty(0)] * K
        if len(timestamps) != K:
            raise ValueError("frames and timestamps differ in length")
        if not self.fused:
            poses = np.empty((K, 4, 4))
            for k in range(K):
                self._register_frame_modular(frames[k], timestamps[k])
                poses[k] = self._last_pose
            return poses
        arrs = [np.asarray(f) for f in frames]
        f32 = K > 0 and all(a.dtype == np.float32 and a.ndim == 2 and a.shape[1] == 3 for a in arrs)
        pts = [np.ascontiguousarray(a) for a in arrs] if f32 else [N.points_arg(a) for a in arrs]
        ts = [np.ascontiguousarray(np.asarray(t).ravel(), dtype=np.float64) for t in timestamps]
        return self._register_frames_raw([N.ptr(a) for a in pts], [len(a) for a in pts], [N.ptr(t) for t in ts],
                                         [len(t) for t in ts], 1 if f32 else 0)

    def _register_frames_raw(self, xyz_ptrs, sizes, ts_ptrs, ts_sizes, layout):
        """raw-pointer form (host or device addresses): layout 0 host f64, 1 host f32, 2 device f64"""
        K = len(xyz_ptrs)
        as_vp = lambda v: v if isinstance(v, C.c_void_p) else C.c_void_p(int(v) if v else None)
        X = (C.c_void_p * K)(*[as_vp(v) for v in xyz_ptrs])
        T = (C.c_void_p * K)(*[as_vp(v) for v in ts_ptrs])
        n = (N.sz * K)(*sizes)
        nt = (N.sz * K)(*ts_sizes)
        poses = np.empty((K, 4, 4))
        N.check(N.lib().kb_pipeline_register_frames(self._h, X, n, T, nt, K, int(layout), N.ptr(poses)))
        return poses

    def _register_frame_modular(self, frame, timestamps):
        # python/kiss_icp/kiss_icp.py:43-75, line for line, on the per-module device API
        frame = self.preprocessor.preprocess(frame, np.asarray(timestamps), self._last_delta)
        source, frame_downsample = self.voxelize(frame)
        sigma = self.adaptive_threshold.get_threshold()
        initial_guess = self._last_pose @ self._last_delta
        new_pose = self.registration.align_points_to_map(
            points=source,
            voxel_map=self.local_map,
            initial_guess=initial_guess,
            max_correspondance_distance=3 * sigma,
            kernel=sigma,
        )
        model_deviation = np.linalg.inv(initial_guess) @ new_pose
        self.adaptive_threshold.update_model_deviation(model_deviation)
        self.local_map.update(frame_downsample, new_pose)
        self._last_delta = np.linalg.inv(self._last_pose) @ new_pose
        self._last_pose = new_pose
        return frame, source

    def voxelize(self, iframe):
        if self.fused:
            pts = N.points_arg(iframe)
            src = np.empty_like(pts)
            ds = np.empty_like(pts)
            a, b = N.sz(0), N.sz(0)
            N.check(N.lib().kb_pipeline_voxelize(self._h, N.ptr(pts), len(pts), N.ptr(src), len(pts), C.byref(a),
                                                 N.ptr(ds), len(pts), C.byref(b)))
            return src[: a.value], ds[: b.value]
        frame_downsample = voxel_down_sample(iframe, self.config.mapping.voxel_size * 0.5)
        source = voxel_down_sample(frame_downsample, self.config.mapping.voxel_size * 1.5)
        return source, frame_downsample
