// Minimal stand-in for <sophus/se3.hpp> (see Eigen/Core beside it): stores the 4x4 it was built from.
#pragma once
#include <Eigen/Core>
namespace Sophus {
class SE3d {
public:
    SE3d() = default;
    explicit SE3d(const Eigen::Matrix4d &T) : T_(T) {}
    Eigen::Matrix4d matrix() const { return T_; }
private:
    Eigen::Matrix4d T_;
};
}  // namespace Sophus
