// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product path.
//
// Dependency-free CPU restatement of the reference's per-scan registration hot path.
// Every function cites the reference file:line it follows (paths relative to
// /root/reference/cpp/kiss_icp/). OpenMP stands in for oneTBB (scheduling only).
//
// PARITY UNPINNED: the reference has no golden vectors for this path and cannot be
// compiled offline (Eigen/Sophus/oneTBB/tsl::robin_map are fetched, not vendored).
#pragma once

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <tuple>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "oracle_math.hpp"
#include "oracle_robin.hpp"

namespace oracle {

using Points = std::vector<Vec3>;

// core/VoxelUtils.hpp:33-37 — floor of a DIVISION, static_cast<int>.
inline Voxel PointToVoxel(const Vec3 &p, double voxel_size) {
    return Voxel{static_cast<int>(std::floor(p.x / voxel_size)),
                 static_cast<int>(std::floor(p.y / voxel_size)),
                 static_cast<int>(std::floor(p.z / voxel_size))};
}

// core/VoxelUtils.cpp:7-21 — first point per voxel (input order), emitted in robin_map
// iteration order.
inline Points VoxelDownsample(const Points &frame, double voxel_size) {
    RobinMap<Vec3> grid;
    grid.reserve(frame.size());
    for (const auto &point : frame) {
        const Voxel voxel = PointToVoxel(point, voxel_size);
        if (!grid.contains(voxel)) {
            Vec3 p = point;
            grid.insert(voxel, std::move(p));
        }
    }
    Points out;
    out.reserve(grid.size());
    for (const auto &b : grid.buckets())
        if (!b.empty()) out.push_back(b.value);
    return out;
}

// core/VoxelHashMap.cpp:35-41
static const int kVoxelShifts[27][3] = {
    {0, 0, 0},   {1, 0, 0},   {-1, 0, 0},  {0, 1, 0},   {0, -1, 0},  {0, 0, 1},   {0, 0, -1},
    {1, 1, 0},   {1, -1, 0},  {-1, 1, 0},  {-1, -1, 0}, {1, 0, 1},   {1, 0, -1},  {-1, 0, 1},
    {-1, 0, -1}, {0, 1, 1},   {0, 1, -1},  {0, -1, 1},  {0, -1, -1}, {1, 1, 1},   {1, 1, -1},
    {1, -1, 1},  {1, -1, -1}, {-1, 1, 1},  {-1, 1, -1}, {-1, -1, 1}, {-1, -1, -1}};

// core/VoxelHashMap.hpp:38-57
struct VoxelHashMap {
    VoxelHashMap(double voxel_size, double max_distance, unsigned int max_points_per_voxel)
        : voxel_size_(voxel_size), max_distance_(max_distance), max_points_per_voxel_(max_points_per_voxel) {}

    void Clear() { map_.clear(); }
    bool Empty() const { return map_.is_empty(); }

    // core/VoxelHashMap.cpp:46-70
    std::tuple<Vec3, double> GetClosestNeighbor(const Vec3 &query) const {
        const Voxel voxel = PointToVoxel(query, voxel_size_);
        Vec3 closest_neighbor{0, 0, 0};
        double closest_distance = std::numeric_limits<double>::max();
        for (const auto &s : kVoxelShifts) {
            const Voxel qv{voxel.x + s[0], voxel.y + s[1], voxel.z + s[2]};
            const Points *points = map_.find(qv);
            if (points != nullptr) {
                // std::min_element with comparator norm(lhs-q) < norm(rhs-q): first minimum wins
                const Vec3 *neighbor = &(*points)[0];
                double nd = norm(*neighbor - query);
                for (size_t i = 1; i < points->size(); ++i) {
                    const double d = norm((*points)[i] - query);
                    if (d < nd) {
                        nd = d;
                        neighbor = &(*points)[i];
                    }
                }
                const double distance = norm(*neighbor - query);
                if (distance < closest_distance) {
                    closest_neighbor = *neighbor;
                    closest_distance = distance;
                }
            }
        }
        return std::make_tuple(closest_neighbor, closest_distance);
    }

    // core/VoxelHashMap.cpp:72-81
    Points Pointcloud() const {
        Points points;
        points.reserve(map_.size() * static_cast<size_t>(max_points_per_voxel_));
        for (const auto &b : map_.buckets())
            if (!b.empty()) points.insert(points.end(), b.value.begin(), b.value.end());
        points.shrink_to_fit();
        return points;
    }

    // core/VoxelHashMap.cpp:83-87
    void Update(const Points &points, const Vec3 &origin) {
        AddPoints(points);
        RemovePointsFarFromLocation(origin);
    }

    // core/VoxelHashMap.cpp:89-95
    void Update(const Points &points, const SE3 &pose) {
        Points transformed(points.size());
        std::transform(points.cbegin(), points.cend(), transformed.begin(),
                       [&](const Vec3 &p) { return se3_act(pose, p); });
        Update(transformed, pose.t);
    }

    // core/VoxelHashMap.cpp:97-119 — sequential, order dependent.
    void AddPoints(const Points &points) {
        const double map_resolution = std::sqrt(voxel_size_ * voxel_size_ / max_points_per_voxel_);
        for (const auto &point : points) {
            const Voxel voxel = PointToVoxel(point, voxel_size_);
            Points *voxel_points = map_.find(voxel);
            if (voxel_points != nullptr) {
                if (voxel_points->size() == max_points_per_voxel_ ||
                    std::any_of(voxel_points->cbegin(), voxel_points->cend(), [&](const Vec3 &vp) {
                        return norm(vp - point) < map_resolution;
                    })) {
                    continue;
                }
                voxel_points->emplace_back(point);
            } else {
                Points fresh;
                fresh.reserve(max_points_per_voxel_);
                fresh.emplace_back(point);
                map_.insert(voxel, std::move(fresh));
            }
        }
    }

    // core/VoxelHashMap.cpp:121-132 — tests only the first-inserted point of each voxel.
    void RemovePointsFarFromLocation(const Vec3 &origin) {
        const double max_distance2 = max_distance_ * max_distance_;
        size_t it = map_.next_occupied(0);
        while (it < map_.bucket_count()) {
            const Vec3 &pt = map_.buckets()[it].value.front();
            if (squaredNorm(pt - origin) >= max_distance2) {
                it = map_.erase_at(it);
            } else {
                it = map_.next_occupied(it + 1);
            }
        }
    }

    size_t NumVoxels() const { return map_.size(); }

    double voxel_size_;
    double max_distance_;
    unsigned int max_points_per_voxel_;
    RobinMap<Points> map_;
};

// core/Registration.hpp:33-45, core/Registration.cpp:43-167
struct Registration {
    Registration(int max_num_iteration, double convergence_criterion, int max_num_threads)
        : max_num_iterations_(max_num_iteration),
          convergence_criterion_(convergence_criterion),
          max_num_threads_(max_num_threads) {}

    struct LinearSystem {
        double JTJ[36];
        double JTr[6];
    };

    // core/Registration.cpp:60-78 (DataAssociation) + :80-121 (BuildLinearSystem), fused per
    // point. Summation order: static chunks per thread, combined in thread order
    // (the reference's TBB order is nondeterministic; semantics identical).
    LinearSystem BuildSystem(const Points &source, const VoxelHashMap &voxel_map,
                             double max_correspondance_distance, double kernel_scale,
                             int *n_corr_out) const {
        int nthreads = 1;
#ifdef _OPENMP
        nthreads = max_num_threads_ > 0 ? max_num_threads_ : omp_get_max_threads();
#endif
        std::vector<LinearSystem> partial(nthreads);
        std::vector<int> ncorr(nthreads, 0);
        for (auto &p : partial) {
            std::fill(std::begin(p.JTJ), std::end(p.JTJ), 0.0);
            std::fill(std::begin(p.JTr), std::end(p.JTr), 0.0);
        }
        const long n = static_cast<long>(source.size());
#pragma omp parallel num_threads(nthreads)
        {
            int tid = 0, nt = 1;
#ifdef _OPENMP
            tid = omp_get_thread_num();
            nt = omp_get_num_threads();
#endif
            const long lo = n * tid / nt, hi = n * (tid + 1) / nt;
            LinearSystem &acc = partial[tid];
            for (long i = lo; i < hi; ++i) {
                const Vec3 &s = source[i];
                const auto [target, distance] = voxel_map.GetClosestNeighbor(s);
                if (!(distance < max_correspondance_distance)) continue;
                ++ncorr[tid];
                // residual = source - target ; J_r = [I | -hat(source)]   (:84-88)
                const Vec3 r = s - target;
                const double residual2 = squaredNorm(r);
                // GM weight k^2 / (k + r^2)^2   (:96-98)
                const double w = (kernel_scale * kernel_scale) /
                                 ((kernel_scale + residual2) * (kernel_scale + residual2));
                double J[3][6] = {{1, 0, 0, 0, s.z, -s.y}, {0, 1, 0, -s.z, 0, s.x}, {0, 0, 1, s.y, -s.x, 0}};
                // J^T * w * J  and  J^T * w * r   (:112-113)
                for (int a = 0; a < 6; ++a) {
                    for (int b = 0; b < 6; ++b) {
                        double v = 0.0;
                        for (int k = 0; k < 3; ++k) v += (J[k][a] * w) * J[k][b];
                        acc.JTJ[6 * a + b] += v;
                    }
                    double v = 0.0;
                    const double rr[3] = {r.x, r.y, r.z};
                    for (int k = 0; k < 3; ++k) v += (J[k][a] * w) * rr[k];
                    acc.JTr[a] += v;
                }
            }
        }
        LinearSystem total = partial[0];
        int nc = ncorr[0];
        for (int t = 1; t < nthreads; ++t) {
            for (int i = 0; i < 36; ++i) total.JTJ[i] += partial[t].JTJ[i];
            for (int i = 0; i < 6; ++i) total.JTr[i] += partial[t].JTr[i];
            nc += ncorr[t];
        }
        if (n_corr_out) *n_corr_out = nc;
        return total;
    }

    // core/Registration.cpp:138-167
    SE3 AlignPointsToMap(const Points &frame, const VoxelHashMap &voxel_map, const SE3 &initial_guess,
                         double max_distance, double kernel_scale, int *iterations_out = nullptr) const {
        if (iterations_out) *iterations_out = 0;
        if (voxel_map.Empty()) return initial_guess;
        Points source = frame;
        for (auto &p : source) p = se3_act(initial_guess, p);  // TransformPoints :55-58
        SE3 T_icp;
        int j = 0;
        for (; j < max_num_iterations_; ++j) {
            const LinearSystem ls = BuildSystem(source, voxel_map, max_distance, kernel_scale, nullptr);
            double rhs[6], dx[6];
            for (int i = 0; i < 6; ++i) rhs[i] = -ls.JTr[i];
            ldlt6_solve(ls.JTJ, rhs, dx);                       // :156
            const SE3 estimation = se3_exp(dx);                 // :157
            for (auto &p : source) p = se3_act(estimation, p);  // :159
            T_icp = se3_mul(estimation, T_icp);                 // :161
            double n2 = 0.0;
            for (int i = 0; i < 6; ++i) n2 += dx[i] * dx[i];
            if (std::sqrt(n2) < convergence_criterion_) {       // :163
                ++j;
                break;
            }
        }
        if (iterations_out) *iterations_out = j;
        return se3_mul(T_icp, initial_guess);  // :166
    }

    int max_num_iterations_;
    double convergence_criterion_;
    int max_num_threads_;
};

// core/Preprocessing.hpp:32-45, core/Preprocessing.cpp:40-95
struct Preprocessor {
    Preprocessor(double max_range, double min_range, bool deskew, int max_num_threads)
        : max_range_(max_range), min_range_(min_range), deskew_(deskew), max_num_threads_(max_num_threads) {}

    // returns false where the reference would throw std::out_of_range (timestamps.at(idx), :76-77)
    bool Preprocess(const Points &frame, const std::vector<double> &timestamps, const SE3 &relative_motion,
                    Points *out) const {
        Points deskewed;
        const Points *src = &frame;
        if (deskew_ && !timestamps.empty()) {
            if (timestamps.size() < frame.size()) return false;
            const auto mm = std::minmax_element(timestamps.cbegin(), timestamps.cend());
            const double min_time = *mm.first, max_time = *mm.second;
            double omega[6];
            se3_log(relative_motion, omega);
            deskewed.resize(frame.size());
            const long n = static_cast<long>(frame.size());
#pragma omp parallel for schedule(static)
            for (long idx = 0; idx < n; ++idx) {
                const double stamp = (timestamps[idx] - min_time) / (max_time - min_time);
                double a[6];
                for (int i = 0; i < 6; ++i) a[i] = (stamp - 1.0) * omega[i];
                deskewed[idx] = se3_act(se3_exp(a), frame[idx]);
            }
            src = &deskewed;
        }
        out->clear();
        out->reserve(src->size());
        for (const auto &point : *src) {
            const double point_range = norm(point);
            if (point_range < max_range_ && point_range > min_range_) out->emplace_back(point);
        }
        return true;
    }

    double max_range_, min_range_;
    bool deskew_;
    int max_num_threads_;
};

// _correct_kitti_scan, python/kiss_icp/pybind/kiss_icp_pybind.cpp:127-138 (the lambda the KITTI loader applies before
// RegisterFrame, datasets/kitti.py:44-48,68). Eigen 3.4 semantics restated: Vector3d::normalized() divides by
// sqrt(squaredNorm) when squaredNorm > 0 and returns the vector unchanged otherwise; AngleAxisd * v =
// toRotationMatrix() * v with Eigen's Rodrigues form (Geometry/AngleAxis.h, toRotationMatrix).
inline Mat3 AngleAxisToRotationMatrix(double angle, const Vec3 &axis) {
    Mat3 res;
    const Vec3 sin_axis = std::sin(angle) * axis;
    const double c = std::cos(angle);
    const Vec3 cos1_axis = (1.0 - c) * axis;
    double tmp;
    tmp = cos1_axis.x * axis.y;
    res.m[0][1] = tmp - sin_axis.z;
    res.m[1][0] = tmp + sin_axis.z;
    tmp = cos1_axis.x * axis.z;
    res.m[0][2] = tmp + sin_axis.y;
    res.m[2][0] = tmp - sin_axis.y;
    tmp = cos1_axis.y * axis.z;
    res.m[1][2] = tmp - sin_axis.x;
    res.m[2][1] = tmp + sin_axis.x;
    res.m[0][0] = cos1_axis.x * axis.x + c;
    res.m[1][1] = cos1_axis.y * axis.y + c;
    res.m[2][2] = cos1_axis.z * axis.z + c;
    return res;
}
inline Points CorrectKITTIScan(const Points &frame) {
    const double VERTICAL_ANGLE_OFFSET = (0.205 * M_PI) / 180.0;
    Points out(frame.size());
    for (size_t i = 0; i < frame.size(); ++i) {
        const Vec3 &pt = frame[i];
        Vec3 rotationVector = cross(pt, Vec3{0., 0., 1.});
        const double z = squaredNorm(rotationVector);
        if (z > 0.0) {
            const double nrm = std::sqrt(z);
            rotationVector = Vec3{rotationVector.x / nrm, rotationVector.y / nrm, rotationVector.z / nrm};
        }
        out[i] = matvec(AngleAxisToRotationMatrix(VERTICAL_ANGLE_OFFSET, rotationVector), pt);
    }
    return out;
}

// core/Threshold.hpp:29-47, core/Threshold.cpp:30-49
struct AdaptiveThreshold {
    AdaptiveThreshold(double initial_threshold, double min_motion_threshold, double max_range)
        : min_motion_threshold_(min_motion_threshold),
          max_range_(max_range),
          model_sse_(initial_threshold * initial_threshold),
          num_samples_(1) {}

    void UpdateModelDeviation(const SE3 &current_deviation) {
        const double theta = angle_axis_angle(quat_to_matrix(current_deviation.q));
        const double delta_rot = 2.0 * max_range_ * std::sin(theta / 2.0);
        const double delta_trans = norm(current_deviation.t);
        const double model_error = delta_trans + delta_rot;
        if (model_error > min_motion_threshold_) {
            model_sse_ += model_error * model_error;
            num_samples_++;
        }
    }
    double ComputeThreshold() const { return std::sqrt(model_sse_ / num_samples_); }

    double min_motion_threshold_, max_range_, model_sse_;
    int num_samples_;
};

// pipeline/KissICP.hpp:36-54
struct KISSConfig {
    double voxel_size = 1.0;
    double max_range = 100.0;
    double min_range = 0.0;
    int max_points_per_voxel = 20;
    double min_motion_th = 0.1;
    double initial_threshold = 2.0;
    int max_num_iterations = 500;
    double convergence_criterion = 0.0001;
    int max_num_threads = 0;
    bool deskew = true;
};

// pipeline/KissICP.hpp:56-96, pipeline/KissICP.cpp:35-75
class KissICP {
public:
    explicit KissICP(const KISSConfig &config)
        : config_(config),
          preprocessor_(config.max_range, config.min_range, config.deskew, config.max_num_threads),
          registration_(config.max_num_iterations, config.convergence_criterion, config.max_num_threads),
          local_map_(config.voxel_size, config.max_range, static_cast<unsigned>(config.max_points_per_voxel)),
          adaptive_threshold_(config.initial_threshold, config.min_motion_th, config.max_range) {}

    // KissICP.cpp:70-75
    void Voxelize(const Points &frame, Points *source, Points *frame_downsample) const {
        *frame_downsample = VoxelDownsample(frame, config_.voxel_size * 0.5);
        *source = VoxelDownsample(*frame_downsample, config_.voxel_size * 1.5);
    }

    // KissICP.cpp:35-68
    bool RegisterFrame(const Points &frame, const std::vector<double> &timestamps, Points *preprocessed_out,
                       Points *source_out) {
        Points preprocessed;
        if (!preprocessor_.Preprocess(frame, timestamps, last_delta_, &preprocessed)) return false;
        Points source, frame_downsample;
        Voxelize(preprocessed, &source, &frame_downsample);
        const double sigma = adaptive_threshold_.ComputeThreshold();
        const SE3 initial_guess = se3_mul(last_pose_, last_delta_);
        const SE3 new_pose = registration_.AlignPointsToMap(source, local_map_, initial_guess, 3.0 * sigma,
                                                            sigma, &last_iterations_);
        const SE3 model_deviation = se3_mul(se3_inverse(initial_guess), new_pose);
        adaptive_threshold_.UpdateModelDeviation(model_deviation);
        local_map_.Update(frame_downsample, new_pose);
        last_delta_ = se3_mul(se3_inverse(last_pose_), new_pose);
        last_pose_ = new_pose;
        last_sigma_ = sigma;
        if (preprocessed_out) *preprocessed_out = std::move(preprocessed);
        if (source_out) *source_out = std::move(source);
        return true;
    }

    SE3 last_pose_;
    SE3 last_delta_;
    int last_iterations_ = 0;
    double last_sigma_ = 0.0;
    KISSConfig config_;
    Preprocessor preprocessor_;
    Registration registration_;
    VoxelHashMap local_map_;
    AdaptiveThreshold adaptive_threshold_;
};

}  // namespace oracle
