"""Host-side mirror of the reference interface: config rules, argument checking, synthetic
data determinism. (CPU only.)"""
import numpy as np
import pytest


def test_config_rules_match_parser():
    import kiss_icp_b200 as K
    c = K.load_config()
    assert c.mapping.voxel_size == 1.0  # max_range / 100 (config/parser.py:78-79)
    c = K.load_config(max_range=50.0)
    assert c.mapping.voxel_size == 0.5
    c = K.load_config(max_range=5.0, data__min_range=10.0)
    assert c.data.min_range == 0.0  # parser.py:73-75
    c = K.load_config(voxel_size=0.3, deskew=False)
    assert c.mapping.voxel_size == 0.3 and c.data.deskew is False
    assert (c.registration.max_num_iterations, c.registration.convergence_criterion) == (500, 1e-4)
    assert (c.adaptive_threshold.initial_threshold, c.adaptive_threshold.min_motion_th) == (2.0, 0.1)
    assert c.mapping.max_points_per_voxel == 20


def test_public_surface_matches_reference_names():
    import kiss_icp_b200 as K
    for name in ("KissICP", "VoxelHashMap", "Registration", "Preprocessor", "AdaptiveThreshold", "FixedThreshold",
                 "voxel_down_sample", "get_voxel_hash_map", "get_registration", "get_preprocessor", "get_threshold_estimator"):
        assert hasattr(K, name)
    for meth in ("clear", "empty", "update", "add_points", "remove_far_away_points", "point_cloud"):
        assert hasattr(K.VoxelHashMap, meth)
    assert hasattr(K.Registration, "align_points_to_map") and hasattr(K.Preprocessor, "preprocess")
    assert hasattr(K.KissICP, "register_frame") and hasattr(K.KissICP, "voxelize")


def test_shape_errors_like_pybind_cast_error():
    from kiss_icp_b200 import _native as N
    with pytest.raises(RuntimeError):
        N.points_arg(np.zeros((5, 2)))
    with pytest.raises(RuntimeError):
        N.points_arg(np.zeros(9))
    with pytest.raises(RuntimeError):
        N.mat4_arg(np.eye(3))
    a = N.points_arg(np.zeros((4, 3), dtype=np.float32)[::2])  # forcecast + c_style
    assert a.dtype == np.float64 and a.flags.c_contiguous and a.shape == (2, 3)


def test_fixed_threshold_python_only_class():
    import kiss_icp_b200 as K
    c = K.load_config(adaptive_threshold__fixed_threshold=0.7)
    th = K.get_threshold_estimator(c)
    assert isinstance(th, K.FixedThreshold) and th.get_threshold() == 0.7
    th.update_model_deviation(np.eye(4))
    assert th.get_threshold() == 0.7


def test_synthetic_streams_are_deterministic_and_shaped():
    from kiss_icp_b200 import synthetic
    a = synthetic.small_shape(seed=4, beams=8, cols=64)
    b = synthetic.small_shape(seed=4, beams=8, cols=64)
    pa, ta = a.scan(3)
    pb, tb = b.scan(3)
    assert np.array_equal(pa, pb) and ta.size == 0 and tb.size == 0
    assert pa.shape[1] == 3 and 0 < len(pa) <= 8 * 64
    assert np.array_equal(pa, pa.astype(np.float32).astype(np.float64))  # fp32-representable like KITTI .bin
    c = synthetic.small_shape(seed=4, beams=8, cols=64, stamps="column")
    pc, tc = c.scan(3)
    assert len(tc) == len(pc) and tc.min() >= 0 and tc.max() < 1
    assert np.allclose(a.pose(0)[:3, :3] @ a.pose(0)[:3, :3].T, np.eye(3), atol=1e-12)
    assert np.linalg.norm(a.pose(1)[:3, 3] - a.pose(0)[:3, 3]) < 0.2  # starts from rest
    assert 0.8 < np.linalg.norm(a.pose(101)[:3, 3] - a.pose(100)[:3, 3]) < 1.2  # ~10 m/s at 10 Hz


def test_metrics_sequence_error_and_ate():
    """kiss-icp_b200/metrics.py (Metrics.cpp:33-189 restated): analytic drifts on a curved 3-D trajectory"""
    from scipy.spatial.transform import Rotation as R
    from kiss_icp_b200 import metrics as M
    n = 900
    gt = np.tile(np.eye(4), (n, 1, 1))
    s = np.arange(n) * 1.0  # ~1 m per frame along a gentle helix
    gt[:, 0, 3], gt[:, 1, 3], gt[:, 2, 3] = 2000 * np.sin(s / 2000), 2000 * (1 - np.cos(s / 2000)), 0.01 * s
    for k in range(n):
        gt[k, :3, :3] = R.from_euler("z", s[k] / 2000).as_matrix()
    assert max(M.sequence_error(gt, gt)) < 1e-6   # (acos near 1 amplifies rounding to ~1e-8 rad, like in the reference)
    assert M.absolute_trajectory_error(gt, gt)[1] < 1e-6
    # a rigidly displaced estimate: relative segment errors vanish, ATE vanishes after the alignment
    T = np.eye(4)
    T[:3, :3] = R.from_euler("xyz", [0.1, -0.2, 0.3]).as_matrix()
    T[:3, 3] = [5.0, 6.0, 7.0]
    moved = np.array([T @ g for g in gt])
    te, re = M.sequence_error(gt, moved)
    assert te < 1e-6 and re < 1e-6
    ar, at = M.absolute_trajectory_error(gt, moved)
    assert ar < 1e-6 and at < 1e-6
    # odometry that over-estimates every step by 1 %: the KITTI translational error is ~1 %
    rel = [np.linalg.inv(gt[k]) @ gt[k + 1] for k in range(n - 1)]
    est = [np.eye(4)]
    for d in rel:
        d = d.copy()
        d[:3, 3] *= 1.01
        est.append(est[-1] @ d)
    te, re = M.sequence_error(np.array([np.linalg.inv(gt[0]) @ g for g in gt]), np.array(est))
    assert 0.95 < te < 1.1 and re < 1e-4
    # constant yaw-rate bias of 1e-4 rad per frame (~1e-4 rad/m): rotational error in the reference's units
    est = [np.eye(4)]
    bias = np.eye(4)
    bias[:3, :3] = R.from_euler("z", 1e-4).as_matrix()
    for d in rel:
        est.append(est[-1] @ d @ bias)
    te, re = M.sequence_error(np.array([np.linalg.inv(gt[0]) @ g for g in gt]), np.array(est))
    assert abs(re - 1e-4 / 3.14 * 180.0) < 0.1 * 1e-4 / 3.14 * 180.0
    # too short for a 100 m segment: 0 / 0 like the reference
    assert all(np.isnan(v) for v in M.sequence_error(gt[:50], gt[:50]))
    with pytest.raises(ValueError):
        M.sequence_error(gt[:10], gt[:9])
    # angle extraction, all quaternion branches
    for rv in ([0.0, 0.0, 0.0], [0.3, -0.2, 0.1], [3.0, 0.0, 0.0], [0.0, 3.1, 0.0], [0.0, 0.1, 3.1], [2.2, -2.2, 0.0]):
        assert abs(M._angle_axis_angle(R.from_rotvec(rv).as_matrix()) - np.linalg.norm(rv)) < 1e-9
