// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product path.
//
// CPU restatement of the third-party arithmetic the reference's registration
// hot path relies on. None of these libraries is vendored in /root/reference
// and none is installed in this image, so their published algorithms are
// restated here from the pinned upstream versions:
//
//   Sophus 1.24.6  (cpp/kiss_icp/3rdparty/sophus/sophus.cmake:28)   SO3 / SE3
//   Eigen  3.4.0   (cpp/kiss_icp/3rdparty/eigen/eigen.cmake:31)     Quaternion, AngleAxis, LDLT
//
// PARITY UNPINNED: the reference ships no golden vectors for this path
// (python/tests/test_kiss_icp.py:1-4 is an import check) and cannot be built
// offline, so this restatement is cross-checked only against independent
// implementations available here (scipy Rotation, numpy.linalg, brute force).
//
// Call sites in the reference that define which operations are needed:
//   SE3::exp            Registration.cpp:157, Preprocessing.cpp:78
//   SE3::log            Preprocessing.cpp:68
//   SE3 * point         Registration.cpp:57, Preprocessing.cpp:79, VoxelHashMap.cpp:92
//   SE3 * SE3, inverse  Registration.cpp:161,166; KissICP.cpp:47,57,62
//   SO3::hat            Registration.cpp:86
//   SE3(Matrix4d), matrix(), rotationMatrix()  kiss_icp_pybind.cpp:68,84,99,103,117; Threshold.cpp:40
//   Matrix6d::ldlt().solve   Registration.cpp:156
//   AngleAxisd(Matrix3d).angle()  Threshold.cpp:40
#pragma once

#include <array>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>

namespace oracle {

struct Vec3 {
    double x, y, z;
};
inline Vec3 operator+(const Vec3 &a, const Vec3 &b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(const Vec3 &a, const Vec3 &b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator*(double s, const Vec3 &a) { return {s * a.x, s * a.y, s * a.z}; }
inline Vec3 cross(const Vec3 &a, const Vec3 &b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// Eigen fixed-size-3 squaredNorm / norm: (x*x + y*y) + z*z, norm = sqrt(squaredNorm)
inline double squaredNorm(const Vec3 &a) { return (a.x * a.x + a.y * a.y) + a.z * a.z; }
inline double norm(const Vec3 &a) { return std::sqrt(squaredNorm(a)); }

struct Mat3 {
    double m[3][3];
};
inline Mat3 mat3_identity() { return {{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}}; }
inline Mat3 matmul(const Mat3 &a, const Mat3 &b) {
    Mat3 c;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            c.m[i][j] = (a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j]) + a.m[i][2] * b.m[2][j];
    return c;
}
inline Vec3 matvec(const Mat3 &a, const Vec3 &v) {
    return {(a.m[0][0] * v.x + a.m[0][1] * v.y) + a.m[0][2] * v.z,
            (a.m[1][0] * v.x + a.m[1][1] * v.y) + a.m[1][2] * v.z,
            (a.m[2][0] * v.x + a.m[2][1] * v.y) + a.m[2][2] * v.z};
}
// Sophus SO3::hat
inline Mat3 hat(const Vec3 &w) { return {{{0, -w.z, w.y}, {w.z, 0, -w.x}, {-w.y, w.x, 0}}}; }

constexpr double kSophusEps = 1e-10;  // Sophus::Constants<double>::epsilon()

// Unit quaternion, Eigen coefficient order (x, y, z, w).
struct Quat {
    double x, y, z, w;
};

// Sophus SO3Base::normalize(): coeffs /= norm.
inline Quat quat_normalized(const Quat &q) {
    const double len = std::sqrt(((q.x * q.x + q.y * q.y) + q.z * q.z) + q.w * q.w);
    return {q.x / len, q.y / len, q.z / len, q.w / len};
}

// Sophus SO3 operator*: explicit Hamilton product, then the quaternion ctor normalises.
inline Quat quat_mul(const Quat &a, const Quat &b) {
    Quat r;
    r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
    r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
    return quat_normalized(r);
}

// Eigen QuaternionBase::_transformVector (Sophus SO3 * point).
inline Vec3 quat_rotate(const Quat &q, const Vec3 &v) {
    const Vec3 qv{q.x, q.y, q.z};
    Vec3 uv = cross(qv, v);
    uv = uv + uv;
    return (v + q.w * uv) + cross(qv, uv);
}

// Eigen QuaternionBase::toRotationMatrix.
inline Mat3 quat_to_matrix(const Quat &q) {
    const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    Mat3 r;
    r.m[0][0] = 1.0 - (tyy + tzz);
    r.m[0][1] = txy - twz;
    r.m[0][2] = txz + twy;
    r.m[1][0] = txy + twz;
    r.m[1][1] = 1.0 - (txx + tzz);
    r.m[1][2] = tyz - twx;
    r.m[2][0] = txz - twy;
    r.m[2][1] = tyz + twx;
    r.m[2][2] = 1.0 - (txx + tyy);
    return r;
}

// Eigen quaternion-from-rotation-matrix (Shoemake 1987), as in Quaternion.h
// quaternionbase_assign_impl<Other,3,3>::run.
inline Quat quat_from_matrix(const Mat3 &mat) {
    Quat q;
    double c[4];  // x y z w
    double t = mat.m[0][0] + mat.m[1][1] + mat.m[2][2];
    if (t > 0.0) {
        t = std::sqrt(t + 1.0);
        c[3] = 0.5 * t;
        t = 0.5 / t;
        c[0] = (mat.m[2][1] - mat.m[1][2]) * t;
        c[1] = (mat.m[0][2] - mat.m[2][0]) * t;
        c[2] = (mat.m[1][0] - mat.m[0][1]) * t;
    } else {
        int i = 0;
        if (mat.m[1][1] > mat.m[0][0]) i = 1;
        if (mat.m[2][2] > mat.m[i][i]) i = 2;
        const int j = (i + 1) % 3;
        const int k = (j + 1) % 3;
        t = std::sqrt(mat.m[i][i] - mat.m[j][j] - mat.m[k][k] + 1.0);
        c[i] = 0.5 * t;
        t = 0.5 / t;
        c[3] = (mat.m[k][j] - mat.m[j][k]) * t;
        c[j] = (mat.m[j][i] + mat.m[i][j]) * t;
        c[k] = (mat.m[k][i] + mat.m[i][k]) * t;
    }
    q.x = c[0];
    q.y = c[1];
    q.z = c[2];
    q.w = c[3];
    return q;
}

// Sophus SO3::expAndTheta.
inline Quat so3_exp(const Vec3 &omega, double *theta_out) {
    const double theta_sq = squaredNorm(omega);
    double imag_factor, real_factor, theta;
    if (theta_sq < kSophusEps * kSophusEps) {
        theta = 0.0;
        const double theta_po4 = theta_sq * theta_sq;
        imag_factor = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
        real_factor = 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * theta_po4;
    } else {
        theta = std::sqrt(theta_sq);
        const double half_theta = 0.5 * theta;
        const double sin_half_theta = std::sin(half_theta);
        imag_factor = sin_half_theta / theta;
        real_factor = std::cos(half_theta);
    }
    if (theta_out) *theta_out = theta;
    return {imag_factor * omega.x, imag_factor * omega.y, imag_factor * omega.z, real_factor};
}

// Sophus SO3::logAndTheta.
inline Vec3 so3_log(const Quat &q, double *theta_out) {
    const double squared_n = (q.x * q.x + q.y * q.y) + q.z * q.z;
    const double w = q.w;
    double two_atan_nbyw_by_n, theta;
    if (squared_n < kSophusEps * kSophusEps) {
        const double squared_w = w * w;
        two_atan_nbyw_by_n = 2.0 / w - (2.0 / 3.0) * (squared_n) / (w * squared_w);
        theta = 2.0 * squared_n / w;
    } else {
        const double n = std::sqrt(squared_n);
        const double atan_nbyw = (w < 0.0) ? std::atan2(-n, -w) : std::atan2(n, w);
        two_atan_nbyw_by_n = 2.0 * atan_nbyw / n;
        theta = two_atan_nbyw_by_n * n;
    }
    if (theta_out) *theta_out = theta;
    return {two_atan_nbyw_by_n * q.x, two_atan_nbyw_by_n * q.y, two_atan_nbyw_by_n * q.z};
}

struct SE3 {
    Quat q{0, 0, 0, 1};
    Vec3 t{0, 0, 0};
};

inline Vec3 se3_act(const SE3 &T, const Vec3 &p) { return quat_rotate(T.q, p) + T.t; }

inline SE3 se3_mul(const SE3 &a, const SE3 &b) {
    SE3 r;
    r.q = quat_mul(a.q, b.q);
    r.t = a.t + quat_rotate(a.q, b.t);
    return r;
}

// Sophus SE3::inverse: invR = SO3(conjugate) (ctor normalises); t' = invR * (t * -1).
inline SE3 se3_inverse(const SE3 &a) {
    SE3 r;
    r.q = quat_normalized(Quat{-a.q.x, -a.q.y, -a.q.z, a.q.w});
    r.t = quat_rotate(r.q, Vec3{a.t.x * -1.0, a.t.y * -1.0, a.t.z * -1.0});
    return r;
}

// Sophus SO3::leftJacobian(omega, theta) as used by SE3::exp.
inline Mat3 so3_left_jacobian(const Vec3 &omega, double theta) {
    const double theta_sq = theta * theta;
    const Mat3 Omega = hat(omega);
    const Mat3 Omega_sq = matmul(Omega, Omega);
    Mat3 V = mat3_identity();
    if (theta_sq < kSophusEps * kSophusEps) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) V.m[i][j] += 0.5 * Omega.m[i][j];
    } else {
        const double a = (1.0 - std::cos(theta)) / theta_sq;
        const double b = (theta - std::sin(theta)) / (theta_sq * theta);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) V.m[i][j] = (V.m[i][j] + a * Omega.m[i][j]) + b * Omega_sq.m[i][j];
    }
    return V;
}

// Sophus SE3::exp; tangent = (upsilon, omega).
inline SE3 se3_exp(const double a[6]) {
    const Vec3 upsilon{a[0], a[1], a[2]};
    const Vec3 omega{a[3], a[4], a[5]};
    double theta;
    SE3 r;
    r.q = so3_exp(omega, &theta);
    const Mat3 V = so3_left_jacobian(omega, theta);
    r.t = matvec(V, upsilon);
    return r;
}

// Sophus SE3::log.
inline void se3_log(const SE3 &T, double out[6]) {
    double theta;
    const Vec3 omega = so3_log(T.q, &theta);
    const Mat3 Omega = hat(omega);
    const Mat3 Omega_sq = matmul(Omega, Omega);
    Mat3 V_inv = mat3_identity();
    if (std::abs(theta) < kSophusEps) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                V_inv.m[i][j] = (V_inv.m[i][j] - 0.5 * Omega.m[i][j]) + (1. / 12.) * Omega_sq.m[i][j];
    } else {
        const double half_theta = 0.5 * theta;
        const double c =
            (1.0 - theta * std::cos(half_theta) / (2.0 * std::sin(half_theta))) / (theta * theta);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                V_inv.m[i][j] = (V_inv.m[i][j] - 0.5 * Omega.m[i][j]) + c * Omega_sq.m[i][j];
    }
    const Vec3 u = matvec(V_inv, T.t);
    out[0] = u.x;
    out[1] = u.y;
    out[2] = u.z;
    out[3] = omega.x;
    out[4] = omega.y;
    out[5] = omega.z;
}

// Row-major 4x4 <-> SE3. from_matrix returns false where Sophus would SOPHUS_ENSURE-fail
// (non-orthogonal R, det <= 0, bad last row).
inline bool se3_from_matrix(const double M[16], SE3 *out) {
    Mat3 R;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R.m[i][j] = M[4 * i + j];
    // isOrthogonal: (R R^T - I).norm() < eps ; det > 0
    double fro = 0.0;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += R.m[i][k] * R.m[j][k];
            s -= (i == j) ? 1.0 : 0.0;
            fro += s * s;
        }
    const double det = R.m[0][0] * (R.m[1][1] * R.m[2][2] - R.m[1][2] * R.m[2][1]) -
                       R.m[0][1] * (R.m[1][0] * R.m[2][2] - R.m[1][2] * R.m[2][0]) +
                       R.m[0][2] * (R.m[1][0] * R.m[2][1] - R.m[1][1] * R.m[2][0]);
    if (!(std::sqrt(fro) < kSophusEps) || !(det > 0.0)) return false;
    if (!(std::abs(M[12]) < kSophusEps && std::abs(M[13]) < kSophusEps &&
          std::abs(M[14]) < kSophusEps && std::abs(M[15] - 1.0) < kSophusEps))
        return false;
    out->q = quat_from_matrix(R);
    out->t = {M[3], M[7], M[11]};
    return true;
}

inline void se3_to_matrix(const SE3 &T, double M[16]) {
    const Mat3 R = quat_to_matrix(T.q);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[4 * i + j] = R.m[i][j];
    M[3] = T.t.x;
    M[7] = T.t.y;
    M[11] = T.t.z;
    M[12] = M[13] = M[14] = 0.0;
    M[15] = 1.0;
}

// Eigen::AngleAxisd(Matrix3d).angle(): quaternion from matrix, then
// angle = 2*atan2(|vec|, |w|) (AngleAxis::operator=(QuaternionBase)).
inline double angle_axis_angle(const Mat3 &R) {
    const Quat q = quat_from_matrix(R);
    double n = std::sqrt((q.x * q.x + q.y * q.y) + q.z * q.z);
    if (n < DBL_EPSILON) {
        // stableNorm
        const double m = std::fmax(std::fabs(q.x), std::fmax(std::fabs(q.y), std::fabs(q.z)));
        if (m > 0.0) {
            const double a = q.x / m, b = q.y / m, c = q.z / m;
            n = m * std::sqrt((a * a + b * b) + c * c);
        } else {
            n = 0.0;
        }
    }
    if (n != 0.0) return 2.0 * std::atan2(n, std::fabs(q.w));
    return 0.0;
}

// Eigen::LDLT<Matrix6d>::compute + solve (Lower, unblocked, diagonal pivoting; pivots with
// |d| <= DBL_MIN give 0 in the solve). A is symmetric 6x6 row-major, only the lower
// triangle is read. Inner products are summed left-to-right (Eigen's packet order is not
// reproducible at the bit level; differences are O(1e-16) relative).
inline void ldlt6_solve(const double A_in[36], const double b[6], double x[6]) {
    constexpr int N = 6;
    double mat[N][N];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) mat[i][j] = A_in[N * i + j];
    int transpositions[N];
    double temp[N];

    for (int k = 0; k < N; ++k) {
        // largest |diagonal| in the trailing corner; first maximum wins
        int big = k;
        double best = std::fabs(mat[k][k]);
        for (int i = k + 1; i < N; ++i) {
            const double v = std::fabs(mat[i][i]);
            if (v > best) {
                best = v;
                big = i;
            }
        }
        transpositions[k] = big;
        if (k != big) {
            const int s = N - big - 1;
            for (int j = 0; j < k; ++j) std::swap(mat[k][j], mat[big][j]);
            for (int i = 0; i < s; ++i) std::swap(mat[N - s + i][k], mat[N - s + i][big]);
            std::swap(mat[k][k], mat[big][big]);
            for (int i = k + 1; i < big; ++i) {
                const double tmp = mat[i][k];
                mat[i][k] = mat[big][i];
                mat[big][i] = tmp;
            }
        }
        const int rs = N - k - 1;
        if (k > 0) {
            for (int j = 0; j < k; ++j) temp[j] = mat[j][j] * mat[k][j];
            double acc = 0.0;
            for (int j = 0; j < k; ++j) acc += mat[k][j] * temp[j];
            mat[k][k] -= acc;
            for (int i = 0; i < rs; ++i) {
                double a2 = 0.0;
                for (int j = 0; j < k; ++j) a2 += mat[k + 1 + i][j] * temp[j];
                mat[k + 1 + i][k] -= a2;
            }
        }
        const double realAkk = mat[k][k];
        const bool pivot_is_valid = std::fabs(realAkk) > 0.0;
        if (k == 0 && !pivot_is_valid) {
            for (int j = 0; j < N; ++j) transpositions[j] = j;
            break;
        }
        if (rs > 0 && pivot_is_valid)
            for (int i = 0; i < rs; ++i) mat[k + 1 + i][k] /= realAkk;
    }

    double d[N];
    for (int i = 0; i < N; ++i) d[i] = b[i];
    // dst = P b
    for (int k = 0; k < N; ++k)
        if (transpositions[k] != k) std::swap(d[k], d[transpositions[k]]);
    // L^-1 (unit lower)
    for (int i = 0; i < N; ++i) {
        double acc = d[i];
        for (int j = 0; j < i; ++j) acc -= mat[i][j] * d[j];
        d[i] = acc;
    }
    // pseudo-inverse of D
    for (int i = 0; i < N; ++i) {
        if (std::fabs(mat[i][i]) > DBL_MIN)
            d[i] /= mat[i][i];
        else
            d[i] = 0.0;
    }
    // L^-T
    for (int i = N - 1; i >= 0; --i) {
        double acc = d[i];
        for (int j = i + 1; j < N; ++j) acc -= mat[j][i] * d[j];
        d[i] = acc;
    }
    // P^T
    for (int k = N - 1; k >= 0; --k)
        if (transpositions[k] != k) std::swap(d[k], d[transpositions[k]]);
    for (int i = 0; i < N; ++i) x[i] = d[i];
}

}  // namespace oracle
