// kiss_icp/pipeline/KissICP.hpp for the B200 backend: the class ros/src/OdometryServer.cpp:80,162,165,222 uses
// (cpp/kiss_icp/pipeline/KissICP.hpp:36-96), forwarding to the C-ABI of kiss_icp_b200.h. Compile it where Eigen and
// Sophus exist (they do on a ROS machine, not in this repo's build image: tests/test_cpp_adaptor.py compiles it
// against two minimal stand-in headers) and link libkiss_icp_b200 instead of kiss_icp_pipeline / kiss_icp_core.
//
// Differences a caller can see: pose() / delta() also exist in their mutable form (OdometryServer.cpp:162,165 writes
// through them): the state lives on the device, so a modified value is written back at the next RegisterFrame;
// VoxelMap() returns the kiss_icp::VoxelHashMap adaptor of core/VoxelHashMap.hpp borrowing the pipeline's map; where
// Sophus would abort on a non-SE(3) matrix a std::invalid_argument is thrown.
#pragma once

#include <Eigen/Core>
#include <sophus/se3.hpp>
#include <stdexcept>
#include <tuple>
#include <vector>

#include "kiss_icp/core/Registration.hpp"
#include "kiss_icp/core/VoxelHashMap.hpp"
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

    explicit KissICP(const KISSConfig &c) : p_(Create(c)), local_map_(VoxelHashMap::Borrowed{}, kb_pipeline_voxel_map(p_)) { Refresh(); }
    ~KissICP() {
        local_map_.map_handle_ = nullptr;  // borrowed from the pipeline, which owns and frees it
        kb_pipeline_destroy(p_);
    }
    KissICP(const KissICP &) = delete;
    KissICP &operator=(const KissICP &) = delete;

    // RegisterFrame(frame, timestamps) -> (preprocessed frame, source), KissICP.cpp:35-68
    Vector3dVectorTuple RegisterFrame(const Vector3dVector &frame, const std::vector<double> &timestamps) {
        WriteBack();
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
    std::vector<Eigen::Vector3d> LocalMap() const { return local_map_.Pointcloud(); }

    const VoxelHashMap &VoxelMap() const { return local_map_; }
    VoxelHashMap &VoxelMap() { return local_map_; }

    const Sophus::SE3d &pose() const { return last_pose_; }
    Sophus::SE3d &pose() { return last_pose_; }  // written back to the device at the next RegisterFrame

    const Sophus::SE3d &delta() const { return last_delta_; }
    Sophus::SE3d &delta() { return last_delta_; }

private:
    static const double *Data(const Vector3dVector &v) { return b200_detail::Data(v); }
    static double *MutableData(Vector3dVector &v) { return b200_detail::MutableData(v); }
    static void Check(int st) { b200_detail::Check(st); }
    static kb_pipeline *Create(const KISSConfig &c) {
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
        kb_pipeline *p = nullptr;
        Check(kb_pipeline_create(&k, &p));
        return p;
    }
    static bool Same(const Sophus::SE3d &a, const Sophus::SE3d &b) {
        const Eigen::Matrix4d A = a.matrix(), B = b.matrix();
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j)
                if (A(i, j) != B(i, j)) return false;
        return true;
    }
    void Refresh() {
        double M[16];
        Check(kb_pipeline_pose(p_, M));
        last_pose_ = synced_pose_ = b200_detail::FromRowMajor(M);
        Check(kb_pipeline_delta(p_, M));
        last_delta_ = synced_delta_ = b200_detail::FromRowMajor(M);
    }
    void WriteBack() {  // the caller assigned through the mutable pose() / delta() since the last call
        double M[16];
        if (!Same(last_pose_, synced_pose_)) {
            b200_detail::ToRowMajor(last_pose_, M);
            Check(kb_pipeline_set_pose(p_, M));
            synced_pose_ = last_pose_;
        }
        if (!Same(last_delta_, synced_delta_)) {
            b200_detail::ToRowMajor(last_delta_, M);
            Check(kb_pipeline_set_delta(p_, M));
            synced_delta_ = last_delta_;
        }
    }

    kb_pipeline *p_ = nullptr;
    VoxelHashMap local_map_;  // borrows the pipeline's map
    Sophus::SE3d last_pose_, last_delta_;
    Sophus::SE3d synced_pose_, synced_delta_;  // what the device holds
};

}  // namespace kiss_icp::pipeline
