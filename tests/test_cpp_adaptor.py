"""adaptors/kiss_icp/pipeline/KissICP.hpp — the class ros/src/OdometryServer.cpp uses, forwarding to the C-ABI.
Eigen and Sophus are not in this image, so the header is compiled against two minimal stand-ins (tests/mock_deps)
and driven by tests/cpp_adaptor_main.cpp the way the ROS node drives the reference class."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_and_run(tmp_path):
    from kiss_icp_b200 import _native as N
    N.lib()  # makes sure the library is built
    exe = str(tmp_path / "cpp_adaptor_main")
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "tests", "mock_deps"),
           "-I" + os.path.join(ROOT, "adaptors"), "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp_adaptor_main.cpp"), "-L" + os.path.join(ROOT, "kiss-icp_b200"), "-lkiss_icp_b200",
           "-Wl,-rpath," + os.path.join(ROOT, "kiss-icp_b200"), "-o", exe]
    env = dict(os.environ)
    env.pop("CXX", None)
    subprocess.run(cmd, check=True, env=env)
    return subprocess.run([exe], capture_output=True, text=True, timeout=300)


def test_adaptor_compiles_links_and_reports_a_missing_device(tmp_path):
    from kiss_icp_b200 import _native as N
    r = build_and_run(tmp_path)
    assert r.returncode == 0, r.stdout + r.stderr
    if N.lib().kb_device_count() < 1:
        assert "no CUDA device" in r.stdout
    else:
        assert r.stdout.strip().endswith("adaptor ok")


@pytest.mark.gpu
def test_adaptor_registers_frames_like_the_ros_node(tmp_path):
    r = build_and_run(tmp_path)
    assert r.returncode == 0 and r.stdout.strip().endswith("adaptor ok"), r.stdout + r.stderr
