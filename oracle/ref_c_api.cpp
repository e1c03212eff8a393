// ORACLE / TEST INFRASTRUCTURE ONLY. C-ABI over the REAL reference (kiss_icp::pipeline::KissICP compiled from the sources
// where they lie under /root/reference, `make -C oracle ref`), so that the restatement in oracle_core.hpp can be checked
// against the reference itself wherever its dependencies (Eigen3, Sophus, oneTBB, tsl::robin_map) are installed. In this
// image they are not: the target reports that and builds nothing; tests/test_oracle_vs_ref.py is skipped then.
#include <cstddef>
#include <cstring>
#include <vector>

#include "kiss_icp/pipeline/KissICP.hpp"

namespace {
using kiss_icp::pipeline::KISSConfig;
using kiss_icp::pipeline::KissICP;

void to_row_major(const Sophus::SE3d &T, double out[16]) {
    const Eigen::Matrix4d M = T.matrix();  // column-major
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) out[4 * r + c] = M(r, c);
}
}  // namespace

extern "C" {
// KISSConfig field by field (pipeline/KissICP.hpp:36-54)
void *ref_pipeline_create(double voxel_size, double max_range, double min_range, int max_points_per_voxel, double min_motion_th,
                          double initial_threshold, int max_num_iterations, double convergence_criterion, int max_num_threads,
                          int deskew) {
    KISSConfig c;
    c.voxel_size = voxel_size;
    c.max_range = max_range;
    c.min_range = min_range;
    c.max_points_per_voxel = max_points_per_voxel;
    c.min_motion_th = min_motion_th;
    c.initial_threshold = initial_threshold;
    c.max_num_iterations = max_num_iterations;
    c.convergence_criterion = convergence_criterion;
    c.max_num_threads = max_num_threads;
    c.deskew = deskew != 0;
    return new KissICP(c);
}
void ref_pipeline_destroy(void *p) { delete static_cast<KissICP *>(p); }
// RegisterFrame (pipeline/KissICP.cpp:35-68) -> pose, sizes of the returned clouds
int ref_pipeline_register_frame(void *p, const double *xyz, size_t n, const double *timestamps, size_t n_timestamps, double pose_out[16],
                                size_t *n_preprocessed, size_t *n_source) {
    auto *icp = static_cast<KissICP *>(p);
    std::vector<Eigen::Vector3d> frame(n);
    if (n) std::memcpy(frame[0].data(), xyz, n * 3 * sizeof(double));
    std::vector<double> ts(timestamps, timestamps + n_timestamps);
    try {
        const auto [pre, src] = icp->RegisterFrame(frame, ts);
        if (n_preprocessed) *n_preprocessed = pre.size();
        if (n_source) *n_source = src.size();
    } catch (...) {
        return 1;
    }
    to_row_major(icp->pose(), pose_out);
    return 0;
}
size_t ref_pipeline_local_map_size(void *p) { return static_cast<KissICP *>(p)->LocalMap().size(); }
}
