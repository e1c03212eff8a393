"""The restatement (oracle/) against the REAL reference, wherever `make -C oracle ref` could build it (oracle/_ref/, needs
Eigen3 + Sophus + oneTBB + tsl::robin_map, none of which is in this image: then these tests are skipped and parity stays
unpinned). The reference's own reduction order is nondeterministic (TBB), so the bar is 1e-9, not bits."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libkiss_icp_ref.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF_LIB), reason="oracle/_ref not built (the reference's dependencies are absent)")


def _ref():
    L = C.CDLL(REF_LIB)
    L.ref_pipeline_create.restype = C.c_void_p
    L.ref_pipeline_create.argtypes = [C.c_double] * 3 + [C.c_int] + [C.c_double] * 2 + [C.c_int, C.c_double, C.c_int, C.c_int]
    L.ref_pipeline_destroy.argtypes = [C.c_void_p]
    L.ref_pipeline_register_frame.restype = C.c_int
    L.ref_pipeline_register_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ref_pipeline_local_map_size.restype = C.c_size_t
    L.ref_pipeline_local_map_size.argtypes = [C.c_void_p]
    return L


@pytest.mark.parametrize("stamps", ["none", "column"])
def test_oracle_register_frame_stream_matches_the_reference(stamps):
    import sys
    sys.path.insert(0, ROOT)
    from kiss_icp_b200 import synthetic
    from oracle import oracle as O
    L = _ref()
    lidar = synthetic.small_shape(seed=3, beams=32, cols=512, stamps=stamps)
    ref = L.ref_pipeline_create(1.0, 100.0, 0.0, 20, 0.1, 2.0, 500, 1e-4, 0, 1)
    o = O.KissICP()
    try:
        for k in range(25):
            p, t = lidar.scan(k)
            p = np.ascontiguousarray(p, dtype=np.float64)
            t = np.ascontiguousarray(t, dtype=np.float64)
            pose = np.empty((4, 4))
            npre, nsrc = C.c_size_t(0), C.c_size_t(0)
            assert L.ref_pipeline_register_frame(ref, p.ctypes.data, len(p), t.ctypes.data if len(t) else None, len(t), pose.ctypes.data,
                                                 C.byref(npre), C.byref(nsrc)) == 0
            o.register_frame(p, t, want_clouds=False)
            assert np.abs(pose - o.pose).max() < 1e-9, (k, np.abs(pose - o.pose).max())
        assert L.ref_pipeline_local_map_size(ref) == o.local_map.num_points()
    finally:
        L.ref_pipeline_destroy(ref)
