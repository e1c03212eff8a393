"""The oracle's Sophus/Eigen restatement (oracle/oracle_math.hpp) against independent
implementations available here: scipy Rotation, numpy.linalg. (CPU only.)"""
import numpy as np
import pytest
from scipy.spatial.transform import Rotation

rng = np.random.default_rng(1)


def rand_se3(O, scale_t=5.0, scale_r=1.0):
    return O.se3_exp(np.concatenate([rng.normal(size=3) * scale_t, rng.normal(size=3) * scale_r]))


def test_so3_exp_matches_scipy(O):
    for _ in range(50):
        w = rng.normal(size=3) * rng.choice([1e-12, 1e-6, 0.1, 1.0, 3.0])
        M = O.se3_exp(np.concatenate([np.zeros(3), w]))
        assert np.allclose(M[:3, :3], Rotation.from_rotvec(w).as_matrix(), atol=1e-14)
        assert np.allclose(M[:3, 3], 0) and np.allclose(M[3], [0, 0, 0, 1])


def test_se3_exp_translation_is_left_jacobian(O):
    for _ in range(20):
        a = rng.normal(size=6)
        w = a[3:]
        th = np.linalg.norm(w)
        W = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        V = np.eye(3) + (1 - np.cos(th)) / th**2 * W + (th - np.sin(th)) / th**3 * W @ W
        assert np.allclose(O.se3_exp(a)[:3, 3], V @ a[:3], atol=1e-13)


def test_se3_log_inverts_exp(O):
    for scale in (1e-11, 1e-5, 0.3, 2.5):
        for _ in range(20):
            a = np.concatenate([rng.normal(size=3) * 3, rng.normal(size=3) * scale])
            if np.linalg.norm(a[3:]) > 3.0:
                a[3:] *= 3.0 / np.linalg.norm(a[3:])
            assert np.allclose(O.se3_log(O.se3_exp(a)), a, atol=1e-11, rtol=1e-11)


def test_group_ops_match_matrix_algebra(O):
    for _ in range(20):
        A, B = rand_se3(O), rand_se3(O)
        assert np.allclose(O.se3_mul(A, B), A @ B, atol=1e-13)
        assert np.allclose(O.se3_inverse(A), np.linalg.inv(A), atol=1e-13)
        p = rng.normal(size=(7, 3)) * 10
        assert np.allclose(O.se3_act(A, p), p @ A[:3, :3].T + A[:3, 3], atol=1e-12)


def test_not_se3_is_rejected(O):
    bad = np.eye(4)
    bad[0, 0] = 1.1
    with pytest.raises(ValueError):
        O.se3_log(bad)
    refl = np.diag([1.0, 1.0, -1.0, 1.0])
    with pytest.raises(ValueError):
        O.se3_inverse(refl)


def test_ldlt6_matches_numpy_on_spd(O):
    for _ in range(30):
        J = rng.normal(size=(40, 6))
        A = J.T @ J
        b = rng.normal(size=6)
        assert np.allclose(O.ldlt6_solve(A, b), np.linalg.solve(A, b), rtol=1e-9, atol=1e-12)


def test_ldlt6_zero_matrix_gives_zero(O):
    # no correspondences -> JTJ = 0 -> Eigen's LDLT solve returns 0 (Registration.cpp:156)
    assert np.array_equal(O.ldlt6_solve(np.zeros((6, 6)), np.ones(6)), np.zeros(6))


def test_ldlt6_rank_deficient_is_finite(O):
    J = rng.normal(size=(10, 3))
    A = np.zeros((6, 6))
    A[:3, :3] = J.T @ J
    b = np.concatenate([rng.normal(size=3), np.zeros(3)])
    x = O.ldlt6_solve(A, b)
    assert np.all(np.isfinite(x))
    assert np.allclose(A @ x, b, atol=1e-9)


def test_ldlt6_uses_only_lower_triangle(O):
    J = rng.normal(size=(20, 6))
    A = J.T @ J
    b = rng.normal(size=6)
    A2 = A.copy()
    A2[np.triu_indices(6, 1)] = 123.0
    assert np.array_equal(O.ldlt6_solve(A, b), O.ldlt6_solve(A2, b))


def test_threshold_update_matches_formula(O):
    # Threshold.cpp:38-49: err = |t| + 2 R sin(theta/2); accumulate only when err > min_motion
    for _ in range(20):
        a = np.concatenate([rng.normal(size=3) * 0.3, rng.normal(size=3) * 0.02])
        T = O.se3_exp(a)
        theta = np.linalg.norm(Rotation.from_matrix(T[:3, :3]).as_rotvec())
        err = np.linalg.norm(T[:3, 3]) + 2 * 100.0 * np.sin(theta / 2)
        sse, n = O.threshold_update(4.0, 1, T, 0.1, 100.0)
        if err > 0.1:
            assert n == 2 and np.isclose(sse, 4.0 + err * err, rtol=1e-12)
        else:
            assert n == 1 and sse == 4.0
    sse, n = O.threshold_update(4.0, 1, np.eye(4), 0.1, 100.0)
    assert (sse, n) == (4.0, 1)
