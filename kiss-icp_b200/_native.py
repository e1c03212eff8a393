"""ctypes binding of the C-ABI (include/kiss_icp_b200.h) + the in-tree build recipe.

The shared library is built IN-TREE (kiss-icp_b200/libkiss_icp_b200.so) with nvcc for sm_100a so
that it travels with the repo snapshot to the GPU box. There is no CPU fallback: if the
library is missing it is an ImportError-grade failure, and without a CUDA device every compute
entry point returns KB_ERR_NO_DEVICE, surfaced here as ``NoDeviceError``.
"""
from __future__ import annotations

import atexit
import ctypes as C
import os
import shutil
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.environ.get("KB_LIB") or os.path.join(_HERE, "libkiss_icp_b200.so")  # KB_LIB: A/B builds side by side
CSRC = os.path.join(_HERE, "csrc")
HEADER = os.path.join(ROOT, "include", "kiss_icp_b200.h")
DEBUG_HEADER = os.path.join(ROOT, "include", "kiss_icp_b200_debug.h")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",  # FP64 products/sums round like the x86-64 reference build (no FMA contraction)
    "-shared", "-Xcompiler", "-fPIC",
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [HEADER, DEBUG_HEADER]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/api.cu -> libkiss_icp_b200.so (sm_100a). No-op when up to date."""
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in sources())
    if force or stale:
        cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB_PATH, os.path.join(CSRC, "api.cu")]
        if verbose:
            print(" ".join(cmd))
        env = dict(os.environ)
        env.pop("CXX", None)  # the image's CXX wrapper is not usable as nvcc host compiler
        env.pop("CC", None)
        subprocess.check_call(cmd, env=env)
    return LIB_PATH


class KissB200Error(RuntimeError):
    pass


class NoDeviceError(KissB200Error):
    pass


KB_OK, KB_ERR_INVALID_ARG, KB_ERR_NOT_SE3, KB_ERR_OUT_OF_RANGE, KB_ERR_CUDA, KB_ERR_CAPACITY, KB_ERR_NO_DEVICE = range(7)

_lib = None
vp, dbl, sz, i32, u32 = C.c_void_p, C.c_double, C.c_size_t, C.c_int, C.c_uint


class Config(C.Structure):
    """kb_config == KISSConfig (cpp/kiss_icp/pipeline/KissICP.hpp:36-54)."""
    _fields_ = [("voxel_size", dbl), ("max_range", dbl), ("min_range", dbl), ("max_points_per_voxel", i32),
                ("min_motion_th", dbl), ("initial_threshold", dbl), ("max_num_iterations", i32),
                ("convergence_criterion", dbl), ("max_num_threads", i32), ("deskew", i32)]


class FrameStats(C.Structure):
    """kb_frame_stats"""
    _fields_ = [("pose", dbl * 16), ("phase_us", dbl * 6), ("icp_queries", dbl), ("icp_candidates", dbl),
                ("iterations", i32), ("n_points_in", i32), ("n_preprocessed", i32), ("n_downsampled", i32),
                ("n_source", i32), ("map_points", i32), ("map_voxels", i32), ("team", i32)]


# every exported symbol of include/kiss_icp_b200.h: name -> (restype, argtypes)
SIGNATURES = {
    "kb_last_error": (C.c_char_p, []),
    "kb_version": (C.c_char_p, []),
    "kb_device_count": (i32, []),
    "kb_set_device": (i32, [i32]),
    "kb_set_stream": (i32, [vp]),
    "kb_set_grid_blocks": (i32, [i32]),
    "kb_map_create": (i32, [dbl, dbl, u32, C.POINTER(vp)]),
    "kb_map_destroy": (i32, [vp]),
    "kb_map_clear": (i32, [vp]),
    "kb_map_empty": (i32, [vp, C.POINTER(i32)]),
    "kb_map_update_origin": (i32, [vp, vp, sz, vp]),
    "kb_map_update_pose": (i32, [vp, vp, sz, vp]),
    "kb_map_add_points": (i32, [vp, vp, sz]),
    "kb_map_remove_far": (i32, [vp, vp]),
    "kb_map_pointcloud": (i32, [vp, vp, sz, C.POINTER(sz)]),
    "kb_map_num_points": (i32, [vp, C.POINTER(sz)]),
    "kb_map_num_voxels": (i32, [vp, C.POINTER(sz)]),
    "kb_map_dump": (i32, [vp, vp, vp, vp, sz, sz, C.POINTER(sz), C.POINTER(sz)]),
    "kb_map_closest_neighbors": (i32, [vp, vp, sz, vp, vp]),
    "kb_map_closest_neighbors_dev": (i32, [vp, vp, sz, vp, vp]),
    "kb_map_compact": (i32, [vp]),
    "kb_map_params": (i32, [vp, C.POINTER(dbl), C.POINTER(dbl), C.POINTER(u32)]),
    "kb_map_query_bytes_dev": (i32, [vp, vp, sz, C.POINTER(dbl)]),
    "kb_map_sync": (i32, [vp]),
    "kb_registration_create": (i32, [i32, dbl, i32, C.POINTER(vp)]),
    "kb_registration_destroy": (i32, [vp]),
    "kb_registration_align_points_to_map": (i32, [vp, vp, sz, vp, vp, dbl, dbl, vp]),
    "kb_registration_last_iterations": (i32, [vp, C.POINTER(i32)]),
    "kb_registration_build_system": (i32, [vp, vp, sz, vp, dbl, dbl, vp, vp, C.POINTER(i32)]),
    "kb_preprocessor_create": (i32, [dbl, dbl, i32, i32, C.POINTER(vp)]),
    "kb_preprocessor_destroy": (i32, [vp]),
    "kb_preprocessor_preprocess": (i32, [vp, vp, sz, vp, sz, vp, vp, sz, C.POINTER(sz)]),
    "kb_threshold_create": (i32, [dbl, dbl, dbl, C.POINTER(vp)]),
    "kb_threshold_destroy": (i32, [vp]),
    "kb_threshold_compute": (i32, [vp, C.POINTER(dbl)]),
    "kb_threshold_update_model_deviation": (i32, [vp, vp]),
    "kb_voxel_down_sample": (i32, [vp, sz, dbl, vp, sz, C.POINTER(sz)]),
    "kb_correct_kitti_scan": (i32, [vp, sz, vp]),
    "kb_correct_kitti_scan_dev": (i32, [vp, sz, vp]),
    "kb_config_default": (None, [C.POINTER(Config)]),
    "kb_pipeline_create": (i32, [C.POINTER(Config), C.POINTER(vp)]),
    "kb_pipeline_destroy": (i32, [vp]),
    "kb_pipeline_register_frame": (i32, [vp, vp, sz, vp, sz]),
    "kb_pipeline_register_frame_f32": (i32, [vp, vp, sz, vp, sz]),
    "kb_pipeline_register_frame_dev": (i32, [vp, vp, sz, vp, sz]),
    "kb_pipeline_register_frames": (i32, [vp, vp, vp, vp, vp, sz, i32, vp]),
    "kb_pipeline_grow_retries": (i32, [vp, C.POINTER(C.c_ulonglong)]),
    "kb_pipeline_last_cloud_sizes": (i32, [vp, C.POINTER(sz), C.POINTER(sz)]),
    "kb_pipeline_last_clouds": (i32, [vp, vp, sz, vp, sz]),
    "kb_pipeline_voxelize": (i32, [vp, vp, sz, vp, sz, C.POINTER(sz), vp, sz, C.POINTER(sz)]),
    "kb_pipeline_pose": (i32, [vp, vp]),
    "kb_pipeline_delta": (i32, [vp, vp]),
    "kb_pipeline_set_pose": (i32, [vp, vp]),
    "kb_pipeline_set_delta": (i32, [vp, vp]),
    "kb_pipeline_voxel_map": (vp, [vp]),
    "kb_pipeline_last_sigma": (i32, [vp, C.POINTER(dbl)]),
    "kb_pipeline_debug_stamps": (i32, [vp, vp, i32]),
    "kb_pipeline_history_stamps": (i32, [vp, vp, sz, C.POINTER(sz)]),
    "kb_debug_ldlt6": (i32, [vp, vp, vp, vp]),
    "kb_debug_icp_solve": (i32, [vp, vp, vp, vp, vp, vp]),
    "kb_debug_icp_schur": (i32, [vp, vp, vp, C.POINTER(i32)]),
    "kb_debug_barrier_ns": (i32, [i32, C.POINTER(dbl)]),
    "kb_pipeline_last_ds_profile": (i32, [vp, vp]),
    "kb_pipeline_last_map_profile": (i32, [vp, vp]),
    "kb_pipeline_last_cache_stats": (i32, [vp, vp]),
    "kb_pipeline_last_icp_work": (i32, [vp, C.POINTER(dbl), C.POINTER(dbl)]),
    "kb_pipeline_threshold": (i32, [vp, C.POINTER(dbl)]),
    "kb_pipeline_last_iterations": (i32, [vp, C.POINTER(i32)]),
    "kb_pipeline_last_profile": (i32, [vp, vp, i32]),
    "kb_pipeline_set_profiling": (i32, [vp, i32]),
    "kb_any_device_stuck": (i32, []),
    "kb_pipeline_set_history": (i32, [vp, sz]),
    "kb_pipeline_get_history": (i32, [vp, vp, sz, C.POINTER(sz)]),
    "kb_pipeline_launch_count": (i32, [vp, C.POINTER(C.c_ulonglong)]),
}


def lib():
    """Load the CUDA library. Fails loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). kiss_icp_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
        atexit.register(_exit_without_teardown_if_stuck)
    return _lib


def _exit_without_teardown_if_stuck():
    """A launch that never ended (reported as an error by the call that waited for it) would block the CUDA context's
    teardown at interpreter exit forever: leave without it."""
    if _lib is not None and _lib.kb_any_device_stuck():
        import sys
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(1)


def check(status: int):
    if status == KB_OK:
        return
    msg = lib().kb_last_error().decode()
    if status == KB_ERR_NO_DEVICE:
        raise NoDeviceError(msg)
    if status == KB_ERR_OUT_OF_RANGE:
        raise IndexError(msg)  # pybind11 maps std::out_of_range to IndexError
    if status == KB_ERR_NOT_SE3:
        raise ValueError(msg)
    raise KissB200Error(f"kb_status {status}: {msg}")


def points_arg(points) -> np.ndarray:
    """ndarray (N,3) float64 C-contiguous, as py_array_to_vectors_double does
    (python/kiss_icp/pybind/stl_vector_eigen.h:67-80: c_style | forcecast, else cast_error)."""
    a = np.asarray(points)
    if a.ndim != 2 or a.shape[1] != 3:
        raise RuntimeError("Unable to cast Python instance to C++ type (expected an (N, 3) array)")
    return np.ascontiguousarray(a, dtype=np.float64)


def mat4_arg(T) -> np.ndarray:
    a = np.ascontiguousarray(T, dtype=np.float64)
    if a.shape != (4, 4):
        raise RuntimeError("Unable to cast Python instance to C++ type (expected a (4, 4) array)")
    return a


def ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)
