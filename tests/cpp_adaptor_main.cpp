// Drives adaptors/kiss_icp/pipeline/KissICP.hpp the way ros/src/OdometryServer.cpp:80,162,165,222 does.
#include <cstdio>
#include <cstring>

#include "kiss_icp/pipeline/KissICP.hpp"

int main() {
    kiss_icp::pipeline::KISSConfig config;
    config.max_range = 50.0;
    config.deskew = false;
    try {
        kiss_icp::pipeline::KissICP icp(config);
        kb_config kc;
        kb_config_default(&kc);
        kc.max_range = config.max_range;
        kc.deskew = 0;
        kb_pipeline *twin = nullptr;
        if (kb_pipeline_create(&kc, &twin) != KB_OK) throw std::runtime_error(kb_last_error());
        std::vector<Eigen::Vector3d> points;
        for (int i = 0; i < 4000; ++i) {  // a coarse ring of walls
            const double a = 0.0015707963267948967 * i;
            points.emplace_back(10.0 * __builtin_cos(a), 10.0 * __builtin_sin(a), 0.01 * (i % 97) - 0.5);
        }
        for (int k = 0; k < 3; ++k) {
            const auto &[frame, keypoints] = icp.RegisterFrame(points, std::vector<double>{});
            const Sophus::SE3d pose = icp.pose();
            std::printf("frame %d: %zu preprocessed, %zu keypoints, t = (%.3g %.3g %.3g)\n", k, frame.size(), keypoints.size(),
                        pose.matrix()(0, 3), pose.matrix()(1, 3), pose.matrix()(2, 3));
            if (frame.size() != points.size() || keypoints.empty()) return 2;
            // the same frame through the plain C-ABI on a second pipeline: the adaptor adds nothing of its own
            if (kb_pipeline_register_frame(twin, points.front().data(), points.size(), nullptr, 0) != KB_OK) return 5;
            double M[16];
            if (kb_pipeline_pose(twin, M) != KB_OK) return 5;
            for (int i = 0; i < 4; ++i)
                for (int j = 0; j < 4; ++j)
                    if (pose.matrix()(i, j) != M[4 * i + j]) {
                        std::printf("pose(%d,%d) = %.17g, C-ABI %.17g\n", i, j, pose.matrix()(i, j), M[4 * i + j]);
                        return 3;
                    }
            if (__builtin_fabs(pose.matrix()(0, 3)) > 0.05) return 6;  // the sensor did not move (ICP stops at |dx| < 1e-4 per step)
        }
        kb_pipeline_destroy(twin);
        const auto [source, downsample] = icp.Voxelize(points);
        if (icp.LocalMap().empty() || source.empty() || downsample.size() < source.size()) return 4;
        std::puts("adaptor ok");
        return 0;
    } catch (const std::runtime_error &e) {
        if (std::strstr(e.what(), "no CUDA device")) {
            std::puts("adaptor ok (no CUDA device: the constructor reports it like every entry point)");
            return 0;
        }
        std::printf("unexpected: %s\n", e.what());
        return 1;
    }
}
