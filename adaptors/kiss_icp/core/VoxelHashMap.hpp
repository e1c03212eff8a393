// kiss_icp/core/VoxelHashMap.hpp for the B200 backend: kiss_icp::VoxelHashMap with the reference's signatures
// (cpp/kiss_icp/core/VoxelHashMap.hpp:38-57) forwarding to the kb_map_* entry points of kiss_icp_b200.h. The voxel
// table lives in HBM; this object is a handle. Differences a caller can see: there is no `map_` member (the
// tsl::robin_map is the thing being replaced), the object is movable but not copyable, and Pointcloud() lists the
// voxels in ascending (x, y, z) order instead of robin_map iteration order (per-voxel point order is the reference's).
#pragma once

#include <Eigen/Core>
#include <sophus/se3.hpp>
#include <stdexcept>
#include <tuple>
#include <utility>
#include <vector>

#include "kiss_icp_b200.h"

namespace kiss_icp {

namespace b200_detail {
inline void Check(int st) {
    if (st == KB_OK) return;
    if (st == KB_ERR_OUT_OF_RANGE) throw std::out_of_range(kb_last_error());  // timestamps.at(idx), Preprocessing.cpp:76-77
    if (st == KB_ERR_NOT_SE3) throw std::invalid_argument(kb_last_error());    // where Sophus would abort
    throw std::runtime_error(kb_last_error());
}
static_assert(sizeof(Eigen::Vector3d) == 3 * sizeof(double), "std::vector<Eigen::Vector3d>::data() must be double[n][3]");
inline const double *Data(const std::vector<Eigen::Vector3d> &v) { return v.empty() ? nullptr : v.front().data(); }
inline double *MutableData(std::vector<Eigen::Vector3d> &v) { return v.empty() ? nullptr : v.front().data(); }
inline Sophus::SE3d FromRowMajor(const double M[16]) {  // the C-ABI is row-major, Eigen's default is column-major
    Eigen::Matrix4d E;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) E(i, j) = M[4 * i + j];
    return Sophus::SE3d(E);
}
inline void ToRowMajor(const Sophus::SE3d &T, double M[16]) {
    const Eigen::Matrix4d E = T.matrix();
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) M[4 * i + j] = E(i, j);
}
}  // namespace b200_detail

struct VoxelHashMap {
    explicit VoxelHashMap(double voxel_size, double max_distance, unsigned int max_points_per_voxel)
        : voxel_size_(voxel_size), max_distance_(max_distance), max_points_per_voxel_(max_points_per_voxel) {
        b200_detail::Check(kb_map_create(voxel_size, max_distance, max_points_per_voxel, &map_handle_));
    }
    // the local map of a pipeline (kb_pipeline_voxel_map): owned by the pipeline, kb_map_destroy is a no-op on it
    struct Borrowed {};
    VoxelHashMap(Borrowed, kb_map *handle) : map_handle_(handle) {
        b200_detail::Check(kb_map_params(handle, &voxel_size_, &max_distance_, &max_points_per_voxel_));
    }
    ~VoxelHashMap() { kb_map_destroy(map_handle_); }
    VoxelHashMap(VoxelHashMap &&o) noexcept
        : voxel_size_(o.voxel_size_),
          max_distance_(o.max_distance_),
          max_points_per_voxel_(o.max_points_per_voxel_),
          map_handle_(std::exchange(o.map_handle_, nullptr)) {}
    VoxelHashMap(const VoxelHashMap &) = delete;
    VoxelHashMap &operator=(const VoxelHashMap &) = delete;

    inline void Clear() { b200_detail::Check(kb_map_clear(map_handle_)); }
    inline bool Empty() const {
        int e = 1;
        b200_detail::Check(kb_map_empty(map_handle_, &e));
        return e != 0;
    }
    void Update(const std::vector<Eigen::Vector3d> &points, const Eigen::Vector3d &origin) {  // VoxelHashMap.cpp:83-87
        b200_detail::Check(kb_map_update_origin(map_handle_, b200_detail::Data(points), points.size(), origin.data()));
    }
    void Update(const std::vector<Eigen::Vector3d> &points, const Sophus::SE3d &pose) {  // VoxelHashMap.cpp:89-95
        double M[16];
        b200_detail::ToRowMajor(pose, M);
        b200_detail::Check(kb_map_update_pose(map_handle_, b200_detail::Data(points), points.size(), M));
    }
    void AddPoints(const std::vector<Eigen::Vector3d> &points) {  // VoxelHashMap.cpp:97-119
        b200_detail::Check(kb_map_add_points(map_handle_, b200_detail::Data(points), points.size()));
    }
    void RemovePointsFarFromLocation(const Eigen::Vector3d &origin) {  // VoxelHashMap.cpp:121-132
        b200_detail::Check(kb_map_remove_far(map_handle_, origin.data()));
    }
    std::vector<Eigen::Vector3d> Pointcloud() const {  // VoxelHashMap.cpp:72-81
        size_t n = 0;
        b200_detail::Check(kb_map_pointcloud(map_handle_, nullptr, 0, &n));
        std::vector<Eigen::Vector3d> out(n);
        if (n) b200_detail::Check(kb_map_pointcloud(map_handle_, b200_detail::MutableData(out), n, &n));
        return out;
    }
    // VoxelHashMap.cpp:46-70; a miss returns (Zero, DBL_MAX) like the reference. One query per call is a launch per
    // call: batch with kb_map_closest_neighbors where the caller has many.
    std::tuple<Eigen::Vector3d, double> GetClosestNeighbor(const Eigen::Vector3d &query) const {
        Eigen::Vector3d p;
        double d = 0.0;
        b200_detail::Check(kb_map_closest_neighbors(map_handle_, query.data(), 1, p.data(), &d));
        return {p, d};
    }

    double voxel_size_ = 0.0;
    double max_distance_ = 0.0;
    unsigned int max_points_per_voxel_ = 0;
    kb_map *map_handle_ = nullptr;  // in place of `tsl::robin_map<Voxel, std::vector<Eigen::Vector3d>> map_`
};

}  // namespace kiss_icp
