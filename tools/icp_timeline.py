"""Developer aid (GPU box): per-phase and per-ICP-iteration times of k_register_frame on a KITTI-shape stream,
from the in-kernel %globaltimer stamps (they cost ~1 us each, so these are upper bounds).
usage: python tools/icp_timeline.py [prime=100] [frames=8]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kiss_icp_b200 as K
from kiss_icp_b200 import _native as N, synthetic

prime = int(sys.argv[1]) if len(sys.argv) > 1 else 100
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 8
L = synthetic.kitti_shape(seed=0, device="cuda")
scans = [L.scan(k) for k in range(prime + frames)]
g = K.KissICP(K.load_config())
for p, t in scans[:prime]:
    g.register_frame(p, t, return_clouds=False)
g.set_profiling(True)
g.start_history(frames)
def last_team(g):
    h = g.history()
    return h[-1].team if h else 0


def g_stats(g):
    out = np.zeros(3)
    N.check(N.lib().kb_pipeline_last_cache_stats(g._h, N.ptr(out)))
    return out


names = ["pre", "ds1", "ds2", "icp", "map", "epi"]
for p, t in scans[prime:]:
    g.register_frame(p, t, return_clouds=False)
    ns = np.zeros(64)
    N.check(N.lib().kb_pipeline_debug_stamps(g._h, N.ptr(ns), 64))
    it = g.last_iterations
    st = ns[41:41 + min(it, 20)]
    d = np.diff(st) * 1e-3
    f = ns[30:35]
    print("   us: fill pass (CTA 0)", round((f[1] - f[0]) * 1e-3, 1), "barrier", round((f[2] - f[1]) * 1e-3, 1), "stage lists", round((f[3] - f[2]) * 1e-3, 1),
          "iterations", round((f[4] - f[3]) * 1e-3, 1), "cache hits/refills/overflows", g_stats(g))
    T = int(last_team(g))
    if T > 0:
        nsb = np.zeros(64 + 4 * 148)
        N.check(N.lib().kb_pipeline_debug_stamps(g._h, N.ptr(nsb), len(nsb)))
        mm = nsb[64:64 + 4 * T].reshape(T, 4)
        comp = (mm[:, 1] - mm[:, 0]) * 1e-3
        wait = (mm[:, 2] - mm[:, 1]) * 1e-3
        skew = (mm[:, 1] - mm[:, 1].min()) * 1e-3
        print("   per member at iteration 4 [us]: compute min/med/max", np.round([comp.min(), np.median(comp), comp.max()], 2),
              "store-time skew med/max", np.round([np.median(skew), skew.max()], 2), "gather wait min/med/max", np.round([wait.min(), np.median(wait), wait.max()], 2), "T", T)
    c = ns[16:24]
    print("   cycles at iteration 4 (member 0, thread 0): transform+walk+terms -> barrier", c[1] - c[0], "refills", c[2] - c[1],
          "column sums", c[3] - c[2], "store partial", c[4] - c[3], "gather+sums", c[6] - c[4], "solve", c[7] - c[6])
    w = ns[24:29]
    print("   warp 0 inside the first phase: transform+validity", w[0] - c[0], "walk", w[1] - w[0], "merge+reload+sqrt", w[2] - w[1], "near check", w[3] - w[2],
          "terms+row", w[4] - w[3], "wait at barrier", c[1] - w[4])
    mp = np.zeros(3)
    N.check(N.lib().kb_pipeline_last_map_profile(g._h, N.ptr(mp)))
    print("   map update us: transform+claim+pending", round(mp[0], 1), "ordered insertion", round(mp[1], 1), "eviction scan", round(mp[2], 1))
    print("iters", it, "phases_us", dict(zip(names, np.round(g.last_profile_us, 1))),
          "iter_us first", np.round(d[:3], 2), "median", round(float(np.median(d)), 2) if len(d) else None,
          "env", os.environ.get("KB_ICP_TEAM_Q", "default"))
