// kiss_icp/pipeline/KissICP.hpp for the B200 backend: the class ros/src/OdometryServer.cpp:80,162,165,222 uses
// (cpp/kiss_icp/pipeline/KissICP.hpp:36-96), forwarding to the C-ABI of kiss_icp_b200.h. Compile it where Eigen and
// Sophus exist (they do on a ROS machine, not in this repo's build image: tests/test_cpp_adaptor.py compiles it
// against two minimal stand-in headers) and link libkiss_icp_b200 instead of kiss_icp_pipeline / kiss_icp_core.
//
// Differences a caller can see: pose() / delta() are returned by const reference only (set them with SetPose /
// SetDelta: the state lives on the device), VoxelMap() is replaced by LocalMap(), and where Sophus would abort on a
// non-SE(3) matrix a std::invalid_argument is thrown.
#pragma once

#include <Eigen/Core>
#include <sophus/se3.hpp>
#include <stdexcept>
#include <tuple>
#include <vector>

#include "kiss_icp_b200.h"

namespace kiss_icp::pipeline {

// Same field names, types and defaults as the reference's KISSConfig (KissICP.hpp:36-54) so that callers that fill
// it field by field (ros/src/OdometryServer.cpp:60-80) compile unchanged; kb_config is its C mirror.
struct KISSConfig {
    double voxel_size = 1.0;            // local map / downsampling grid [m]
    double max_range = 100.0;           // crop + map radius [m]
    double min_range = 0.0;
    int max_points_per_voxel = 20;
    double min_motion_th = 0.1;         // adaptive threshold: ignore model deviations below this [m]
    double initial_threshold = 2.0;     // sigma before any motion was observed [m]
    int max_num_iterations = 500;       // ICP
    double convergence_criterion = 0.0001;
    int max_num_threads = 0;            // accepted, ignored on the GPU
    bool deskew = true;                 // needs per-point timestamps
};

class KissICP {
public:
    using Vector3dVector = std::vector<Eigen::Vector3d>;
    using Vector3dVectorTuple = std::tuple<Vector3dVector, Vector3dVector>;
    static_assert(sizeof(Eigen::Vector3d) == 3 * sizeof(double), "std::vector<Eigen::Vector3d>::data() must be double[n][3]");

    explicit KissICP(const KISSConfig &c) {
        kb_config k;
        kb_config_default(&k);
        k.voxel_size = c.voxel_size;
        k.max_range = c.max_range;
        k.min_range = c.min_range;
        k.max_points_per_voxel = c.max_points_per_voxel;
        k.min_motion_th = c.min_motion_th;
        k.initial_threshold = c.initial_threshold;
        k.max_num_iterations = c.max_num_iterations;
        k.convergence_criterion = c.convergence_criterion;
        k.max_num_threads = c.max_num_threads;
        k.deskew = c.deskew ? 1 : 0;
        Check(kb_pipeline_create(&k, &p_));
        Refresh();
    }
    ~KissICP() { kb_pipeline_destroy(p_); }
    KissICP(const KissICP &) = delete;
    KissICP &operator=(const KissICP &) = delete;

    // RegisterFrame(frame, timestamps) -> (preprocessed frame, source), KissICP.cpp:35-68
    Vector3dVectorTuple RegisterFrame(const Vector3dVector &frame, const std::vector<double> &timestamps) {
        Check(kb_pipeline_register_frame(p_, Data(frame), frame.size(), timestamps.data(), timestamps.size()));
        size_t np = 0, ns = 0;
        Check(kb_pipeline_last_cloud_sizes(p_, &np, &ns));
        Vector3dVector pre(np), src(ns);  // only materialised because the signature returns them
        Check(kb_pipeline_last_clouds(p_, MutableData(pre), np, MutableData(src), ns));
        Refresh();
        return {std::move(pre), std::move(src)};
    }
    // Voxelize(frame) -> (source, frame_downsample), KissICP.cpp:70-75
    Vector3dVectorTuple Voxelize(const Vector3dVector &frame) const {
        Vector3dVector src(frame.size()), ds(frame.size());
        size_t ns = 0, nd = 0;
        Check(kb_pipeline_voxelize(p_, Data(frame), frame.size(), MutableData(src), src.size(), &ns, MutableData(ds), ds.size(), &nd));
        src.resize(ns);
        ds.resize(nd);
        return {std::move(src), std::move(ds)};
    }
    Vector3dVector LocalMap() const {
        kb_map *m = kb_pipeline_voxel_map(p_);
        size_t n = 0;
        Check(kb_map_pointcloud(m, nullptr, 0, &n));
        Vector3dVector out(n);
        if (n) Check(kb_map_pointcloud(m, MutableData(out), n, &n));
        return out;
    }
    const Sophus::SE3d &pose() const { return last_pose_; }
    const Sophus::SE3d &delta() const { return last_delta_; }
    void SetPose(const Sophus::SE3d &T) {
        double M[16];
        ToRowMajor(T, M);
        Check(kb_pipeline_set_pose(p_, M));
        Refresh();
    }
    void SetDelta(const Sophus::SE3d &T) {
        double M[16];
        ToRowMajor(T, M);
        Check(kb_pipeline_set_delta(p_, M));
        Refresh();
    }

private:
    static const double *Data(const Vector3dVector &v) { return v.empty() ? nullptr : v.front().data(); }
    static double *MutableData(Vector3dVector &v) { return v.empty() ? nullptr : v.front().data(); }
    static void Check(int st) {
        if (st == KB_OK) return;
        if (st == KB_ERR_OUT_OF_RANGE) throw std::out_of_range(kb_last_error());  // timestamps.at(idx), Preprocessing.cpp:76-77
        if (st == KB_ERR_NOT_SE3) throw std::invalid_argument(kb_last_error());
        throw std::runtime_error(kb_last_error());
    }
    static Sophus::SE3d FromRowMajor(const double M[16]) {  // the C-ABI is row-major, Eigen's default is column-major
        Eigen::Matrix4d E;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) E(i, j) = M[4 * i + j];
        return Sophus::SE3d(E);
    }
    static void ToRowMajor(const Sophus::SE3d &T, double M[16]) {
        const Eigen::Matrix4d E = T.matrix();
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) M[4 * i + j] = E(i, j);
    }
    void Refresh() {
        double M[16];
        Check(kb_pipeline_pose(p_, M));
        last_pose_ = FromRowMajor(M);
        Check(kb_pipeline_delta(p_, M));
        last_delta_ = FromRowMajor(M);
    }

    kb_pipeline *p_ = nullptr;
    Sophus::SE3d last_pose_, last_delta_;
};

}  // namespace kiss_icp::pipeline
