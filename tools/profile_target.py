"""Small driver for ncu: registers a short KITTI-shape stream through the fused kernel."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kiss_icp_b200 as K
from kiss_icp_b200 import synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
L = synthetic.kitti_shape(seed=0, device="cuda")
scans = [L.scan(k) for k in range(n)]
g = K.KissICP(K.load_config())
g.set_profiling(len(sys.argv) > 2 and sys.argv[2] == "stamps")  # ncu captures the uninstrumented kernel
for p, t in scans:
    g.register_frame(p, t, return_clouds=False)
print("iters(last)", g.last_iterations, "phases", np.round(g.last_profile_us, 1))
