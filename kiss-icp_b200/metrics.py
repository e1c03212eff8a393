"""Trajectory metrics — python/kiss_icp/metrics.py:30-39 surface (`sequence_error`, `absolute_trajectory_error`).

Host code, like in the reference (cpp/kiss_icp/metrics/Metrics.cpp:33-189 runs on the CPU once per sequence; it is
not part of the per-scan hot path and nothing here touches the device). Restated from the reference's rules:
KITTI dev-kit segment errors over 100..800 m every 10 frames (Metrics.cpp:35-36,101-145), the average converted
with the reference's constants (note `/ 3.14 * 180`, Metrics.cpp:161), and ATE after an SVD alignment without scale
(Eigen::umeyama(source, target, false), Metrics.cpp:166-189). Both return float32 pairs like std::tuple<float, float>.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

LENGTHS = (100.0, 200.0, 300.0, 400.0, 500.0, 600.0, 700.0, 800.0)  # Metrics.cpp:35
STEP_SIZE = 10  # Metrics.cpp:106 ("every second")


def _poses(a) -> np.ndarray:
    p = np.asarray(a, dtype=np.float64)
    if p.ndim != 3 or p.shape[1:] != (4, 4):
        raise ValueError("poses must have shape (N, 4, 4)")
    return p


def _trajectory_distances(poses: np.ndarray) -> np.ndarray:
    # Metrics.cpp:47-63
    d = np.linalg.norm(np.diff(poses[:, :3, 3], axis=0), axis=1)
    return np.concatenate([[0.0], np.cumsum(d)])


def _sequence_errors(gt: np.ndarray, res: np.ndarray):
    dist = _trajectory_distances(gt)
    out = []
    for first in range(0, len(gt), STEP_SIZE):
        for length in LENGTHS:
            beyond = np.nonzero(dist[first:] > dist[first] + length)[0]  # LastFrameFromSegmentLength, :65-74
            if len(beyond) == 0:
                continue
            last = first + int(beyond[0])
            delta_gt = np.linalg.inv(gt[first]) @ gt[last]
            delta_res = np.linalg.inv(res[first]) @ res[last]
            err = np.linalg.inv(delta_res) @ delta_gt
            d = 0.5 * (err[0, 0] + err[1, 1] + err[2, 2] - 1.0)
            r_err = float(np.arccos(max(min(d, 1.0), -1.0)))  # RotationError, :76-82
            t_err = float(np.linalg.norm(err[:3, 3]))  # TranslationError, :84-89
            out.append((r_err / length, t_err / length))
    return out


def sequence_error(gt_poses: np.ndarray, results_poses: np.ndarray) -> Tuple[float, float]:
    """-> (average translational error [%], average rotational error [deg/m]) — SeqError, Metrics.cpp:151-165.
    A trajectory shorter than 100 m has no segment: the reference divides 0 by 0, so do we (nan, nan)."""
    gt, res = _poses(gt_poses), _poses(results_poses)
    if len(gt) != len(res):
        raise ValueError("different number of poses in ground truth and estimate")
    err = _sequence_errors(gt, res)
    with np.errstate(invalid="ignore", divide="ignore"):
        n = np.float64(len(err))
        t = np.float64(sum(e[1] for e in err)) / n
        r = np.float64(sum(e[0] for e in err)) / n
    return float(np.float32(100.0 * t)), float(np.float32(r / 3.14 * 180.0))


def _umeyama_no_scale(src: np.ndarray, dst: np.ndarray) -> np.ndarray:
    """Eigen::umeyama(src, dst, with_scaling=false) for 3 x N point sets -> 4x4"""
    n = src.shape[1]
    mu_s, mu_d = src.mean(axis=1, keepdims=True), dst.mean(axis=1, keepdims=True)
    sigma = (dst - mu_d) @ (src - mu_s).T / n
    U, _, Vt = np.linalg.svd(sigma)
    S = np.ones(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        S[2] = -1.0
    T = np.eye(4)
    T[:3, :3] = U @ np.diag(S) @ Vt
    T[:3, 3] = (mu_d - T[:3, :3] @ mu_s).ravel()
    return T


def _angle_axis_angle(R: np.ndarray) -> float:
    """Eigen::AngleAxisd(Matrix3d).angle(): matrix -> quaternion (Eigen's four branches) -> 2 atan2(|vec|, |w|)"""
    t = R[0, 0] + R[1, 1] + R[2, 2]
    if t > 0.0:
        t = np.sqrt(t + 1.0)
        w = 0.5 * t
        t = 0.5 / t
        vec = np.array([(R[2, 1] - R[1, 2]) * t, (R[0, 2] - R[2, 0]) * t, (R[1, 0] - R[0, 1]) * t])
    else:
        i = 0
        if R[1, 1] > R[0, 0]:
            i = 1
        if R[2, 2] > R[i, i]:
            i = 2
        j, k = (i + 1) % 3, (i + 2) % 3
        t = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0)
        vec = np.zeros(3)
        vec[i] = 0.5 * t
        t = 0.5 / t
        w = (R[k, j] - R[j, k]) * t
        vec[j] = (R[j, i] + R[i, j]) * t
        vec[k] = (R[k, i] + R[i, k]) * t
    return float(2.0 * np.arctan2(np.linalg.norm(vec), abs(w)))


def absolute_trajectory_error(gt_poses: np.ndarray, results_poses: np.ndarray) -> Tuple[float, float]:
    """-> (rotational RMSE [rad], translational RMSE [m]) after aligning the estimate to the ground truth —
    AbsoluteTrajectoryError, Metrics.cpp:166-189"""
    gt, res = _poses(gt_poses), _poses(results_poses)
    if len(gt) != len(res):
        raise ValueError("different number of poses in ground truth and estimate")
    T_align = _umeyama_no_scale(res[:, :3, 3].T, gt[:, :3, 3].T)
    rot, trans = 0.0, 0.0
    for T_res, T_gt in zip(res, gt):
        delta = np.linalg.inv(T_align @ T_res) @ T_gt
        theta = _angle_axis_angle(delta[:3, :3])
        rot += theta * theta
        trans += float(delta[:3, 3] @ delta[:3, 3])
    return float(np.float32(np.sqrt(rot / len(gt)))), float(np.float32(np.sqrt(trans / len(gt))))
