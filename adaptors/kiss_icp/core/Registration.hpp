// kiss_icp/core/Registration.hpp for the B200 backend: kiss_icp::Registration with the reference's signature
// (cpp/kiss_icp/core/Registration.hpp:33-45) forwarding to kb_registration_* of kiss_icp_b200.h.
// max_num_threads is accepted and ignored (there is no TBB pool to size; Registration.cpp:126-136).
#pragma once

#include <Eigen/Core>
#include <sophus/se3.hpp>
#include <utility>
#include <vector>

#include "VoxelHashMap.hpp"

namespace kiss_icp {

struct Registration {
    explicit Registration(int max_num_iteration, double convergence_criterion, int max_num_threads)
        : max_num_iterations_(max_num_iteration),
          convergence_criterion_(convergence_criterion),
          max_num_threads_(max_num_threads) {
        b200_detail::Check(kb_registration_create(max_num_iteration, convergence_criterion, max_num_threads, &handle_));
    }
    ~Registration() { kb_registration_destroy(handle_); }
    Registration(Registration &&o) noexcept
        : max_num_iterations_(o.max_num_iterations_),
          convergence_criterion_(o.convergence_criterion_),
          max_num_threads_(o.max_num_threads_),
          handle_(std::exchange(o.handle_, nullptr)) {}
    Registration(const Registration &) = delete;
    Registration &operator=(const Registration &) = delete;

    // Registration.cpp:138-167: the whole loop runs on the device, one launch
    Sophus::SE3d AlignPointsToMap(const std::vector<Eigen::Vector3d> &frame,
                                  const VoxelHashMap &voxel_map,
                                  const Sophus::SE3d &initial_guess,
                                  const double max_correspondence_distance,
                                  const double kernel_scale) {
        double G[16], out[16];
        b200_detail::ToRowMajor(initial_guess, G);
        b200_detail::Check(kb_registration_align_points_to_map(handle_, b200_detail::Data(frame), frame.size(), voxel_map.map_handle_,
                                                               G, max_correspondence_distance, kernel_scale, out));
        return b200_detail::FromRowMajor(out);
    }

    int max_num_iterations_;
    double convergence_criterion_;
    int max_num_threads_;
    kb_registration *handle_ = nullptr;
};

}  // namespace kiss_icp
