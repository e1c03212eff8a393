"""adaptors/kiss_icp/pipeline/KissICP.hpp — the class ros/src/OdometryServer.cpp uses, forwarding to the C-ABI.
Eigen and Sophus are not in this image, so the header is compiled against two minimal stand-ins (tests/mock_deps)
and driven by tests/cpp_adaptor_main.cpp the way the ROS node drives the reference class."""
import os
import subprocess

import numpy as np

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


def _ring(shift_x):
    i = np.arange(4000)
    a = 0.0015707963267948967 * i
    return np.stack([10.0 * np.cos(a) - shift_x, 10.0 * np.sin(a), 0.01 * (i % 97) - 0.5], axis=1)


@pytest.mark.gpu
def test_adaptors_match_the_c_abi_and_the_oracle(tmp_path, O):
    """the C++ driver checks adaptor == plain C-ABI bit for bit (incl. the write-back through the mutable pose());
    here its printed poses are compared with the oracle on the same frames"""
    r = build_and_run(tmp_path)
    assert r.returncode == 0 and r.stdout.strip().endswith("adaptor ok"), r.stdout + r.stderr
    lines = [l.split() for l in r.stdout.splitlines()]
    poses = {int(l[1]): np.array(l[2:], dtype=np.float64).reshape(4, 4) for l in lines if l and l[0] == "POSE"}
    align = [np.array(l[2:], dtype=np.float64).reshape(4, 4) for l in lines if l and l[0] == "ALIGN"][0]
    nn = [np.array(l[1:], dtype=np.float64) for l in lines if l and l[0] == "NN"][0]
    cfg = dict(max_range=50.0, deskew=False)
    icp = O.KissICP(**cfg)
    for k in range(3):  # (frame 3 follows a write through the mutable pose(): checked against the C-ABI by the driver)
        icp.register_frame(_ring(0.05 * k), np.empty(0), want_clouds=False)
        assert np.abs(np.array(icp.pose) - poses[k]).max() < 1e-9, k
    assert 3 in poses
    m = O.VoxelHashMap(1.0, 100.0, 20)
    m.update(_ring(0.0), np.zeros(3))
    p, d = m.closest_neighbors(np.array([[9.9, 0.3, 0.0]]))
    assert np.array_equal(p[0], nn[:3]) and d[0] == nn[3]
    ref, _ = O.align_points_to_map(m, _ring(0.07), np.eye(4), 3.0, 1.0)
    assert np.abs(ref - align).max() < 1e-9
