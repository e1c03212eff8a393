// Drives the header-only adaptors (adaptors/kiss_icp/{pipeline/KissICP,core/VoxelHashMap,core/Registration}.hpp) the way
// ros/src/OdometryServer.cpp:80,162,165,222 and the reference's own C++ callers do. Every result is also computed through
// the plain C-ABI (the adaptors add nothing of their own) and printed with 17 digits so that tests/test_cpp_adaptor.py
// can compare it with the oracle.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "kiss_icp/pipeline/KissICP.hpp"

static std::vector<Eigen::Vector3d> ring(double shift_x) {  // a coarse ring of walls seen from (shift_x, 0, 0)
    std::vector<Eigen::Vector3d> points;
    for (int i = 0; i < 4000; ++i) {
        const double a = 0.0015707963267948967 * i;
        points.emplace_back(10.0 * std::cos(a) - shift_x, 10.0 * std::sin(a), 0.01 * (i % 97) - 0.5);
    }
    return points;
}
static void print_pose(const char *tag, int k, const Sophus::SE3d &T) {
    std::printf("%s %d", tag, k);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) std::printf(" %.17g", T.matrix()(i, j));
    std::printf("\n");
}
static bool same(const Sophus::SE3d &T, const double M[16]) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (T.matrix()(i, j) != M[4 * i + j]) {
                std::printf("(%d,%d): adaptor %.17g, C-ABI %.17g\n", i, j, T.matrix()(i, j), M[4 * i + j]);
                return false;
            }
    return true;
}

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
        double M[16];
        // --- the ROS node's loop: RegisterFrame, pose(), LocalMap() ---------------------------------------------------
        for (int k = 0; k < 4; ++k) {
            const auto points = ring(0.05 * k);  // the sensor advances 5 cm per scan
            if (k == 3) {
                // OdometryServer.cpp:162,165 style: overwrite the state through the mutable accessors
                Eigen::Matrix4d P = icp.pose().matrix();
                P(0, 3) += 0.01;
                icp.pose() = Sophus::SE3d(P);
                for (int i = 0; i < 4; ++i)
                    for (int j = 0; j < 4; ++j) M[4 * i + j] = P(i, j);
                if (kb_pipeline_set_pose(twin, M) != KB_OK) return 7;
            }
            const auto &[frame, keypoints] = icp.RegisterFrame(points, std::vector<double>{});
            print_pose("POSE", k, icp.pose());
            if (frame.size() != points.size() || keypoints.empty()) return 2;
            if (kb_pipeline_register_frame(twin, points.front().data(), points.size(), nullptr, 0) != KB_OK) return 5;
            if (kb_pipeline_pose(twin, M) != KB_OK || !same(icp.pose(), M)) return 3;
            if (kb_pipeline_delta(twin, M) != KB_OK || !same(icp.delta(), M)) return 3;
            if (std::fabs(icp.pose().matrix()(0, 3) - 0.05 * k) > 0.05) return 6;  // it follows the motion
        }
        kb_pipeline_destroy(twin);
        const auto [source, downsample] = icp.Voxelize(ring(0.0));
        if (icp.LocalMap().empty() || source.empty() || downsample.size() < source.size()) return 4;
        const kiss_icp::VoxelHashMap &vm = std::as_const(icp).VoxelMap();
        if (vm.Empty() || vm.Pointcloud().size() != icp.LocalMap().size() || vm.voxel_size_ != config.voxel_size ||
            vm.max_distance_ != config.max_range || vm.max_points_per_voxel_ != 20u)
            return 8;
        // --- the core classes on their own: VoxelHashMap + Registration (Registration.hpp:33-45) -----------------------
        kiss_icp::VoxelHashMap map(1.0, 100.0, 20);
        if (!map.Empty()) return 9;
        const auto base = ring(0.0);
        map.Update(base, Eigen::Vector3d(0.0, 0.0, 0.0));
        if (map.Empty() || map.Pointcloud().empty()) return 9;
        const Eigen::Vector3d query(9.9, 0.3, 0.0);
        const auto [nn, dist] = map.GetClosestNeighbor(query);
        double p2[3], d2 = 0.0;
        if (kb_map_closest_neighbors(map.map_handle_, query.data(), 1, p2, &d2) != KB_OK) return 10;
        if (nn[0] != p2[0] || nn[1] != p2[1] || nn[2] != p2[2] || dist != d2 || !(dist < 1.0)) return 10;
        std::printf("NN %.17g %.17g %.17g %.17g\n", nn[0], nn[1], nn[2], dist);
        const auto [miss, dmiss] = map.GetClosestNeighbor(Eigen::Vector3d(500.0, 500.0, 500.0));
        if (miss[0] != 0.0 || miss[1] != 0.0 || miss[2] != 0.0 || !(dmiss > 1e300)) return 11;  // (Zero, DBL_MAX), VoxelHashMap.cpp:51-52
        kiss_icp::Registration reg(500, 1e-4, 0);
        const Sophus::SE3d aligned = reg.AlignPointsToMap(ring(0.07), map, Sophus::SE3d(), 3.0, 1.0);
        print_pose("ALIGN", 0, aligned);
        if (std::fabs(aligned.matrix()(0, 3) - 0.07) > 0.03) return 12;
        map.RemovePointsFarFromLocation(Eigen::Vector3d(1000.0, 0.0, 0.0));
        if (!map.Empty()) return 13;
        map.AddPoints(base);
        map.Clear();
        if (!map.Empty()) return 13;
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
