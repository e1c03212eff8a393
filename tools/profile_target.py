"""Small driver for ncu: registers a short KITTI-shape stream through the fused kernel.
usage: python tools/profile_target.py [scans=40] [blocking|queued]   (queued = kb_pipeline_register_frames: the launches
then also run the next scan's front end on the SMs the ICP team leaves idle, as in the bench)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kiss_icp_b200 as K
from kiss_icp_b200 import synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
mode = sys.argv[2] if len(sys.argv) > 2 else "queued"
L = synthetic.kitti_shape(seed=0)
scans = [L.scan(k) for k in range(n)]
g = K.KissICP(K.load_config())
g.start_history(n)
if mode == "queued":
    g.register_frames([p for p, _ in scans], None)
else:
    for p, t in scans:
        g.register_frame(p, t, return_clouds=False)
h = g.history()
print("mode", mode, "iters(last)", g.last_iterations, "team of the last scan", h[-1].team if h else None)
