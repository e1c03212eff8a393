"""Developer aid (GPU box): where a scan's time goes in the QUEUED path (kb_pipeline_register_frames, front-end prefetch on),
from the light in-kernel stamps (phase boundaries only). usage: python tools/queue_timeline.py [prime=100] [frames=40]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kiss_icp_b200 as K
from kiss_icp_b200 import _native as N, synthetic

prime = int(sys.argv[1]) if len(sys.argv) > 1 else 100
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 40
L = synthetic.kitti_shape(seed=0, device="cuda")
scans = [L.scan_torch(k) for k in range(prime + frames)]
g = K.KissICP(K.load_config())
raw = lambda ss: g._register_frames_raw([s[0].data_ptr() for s in ss], [s[0].shape[0] for s in ss], [None] * len(ss), [0] * len(ss), 2)
raw(scans[:prime])
g.set_profiling(2)
g.start_history(frames)
raw(scans[prime:])
n = N.sz(0)
N.check(N.lib().kb_pipeline_history_stamps(g._h, None, 0, C.byref(n)))
st = np.zeros((n.value, 20))
N.check(N.lib().kb_pipeline_history_stamps(g._h, N.ptr(st), n.value, C.byref(n)))
h = g.history()
t = st[2:-1]  # steady part
us = lambda a: np.round(np.median(a) * 1e-3, 1)
print("frames", len(t), "team", h[-1].team, "iterations (median)", np.median([x.iterations for x in h]))
print("kernel total", us(t[:, 6] - t[:, 0]), "| start->lists done", us(t[:, 7] - t[:, 0]), "| lists->iterations done", us(t[:, 10] - t[:, 7]),
      "| prefetch done after lists", us(t[:, 11] - t[:, 7]), "| iterations->result everywhere", us(t[:, 4] - t[:, 10]),
      "| map update", us(t[:, 5] - t[:, 4]), "| tail", us(t[:, 6] - t[:, 5]))
print("gap between kernels", us(st[3:, 0] - st[2:-1, 6]), "| period", us(np.diff(st[2:, 0])))
