"""BASELINE config 2 at full length: the KITTI-00-shape synthetic stream (4541 scans, 64 x 1024 rays, no stamps), free running,
GPU path (queued kb_pipeline_register_frames) against the oracle on the same scans. Records the largest pose difference, the
iteration-count mismatches, the throughput of both and the KITTI-style drift of both trajectories against the synthetic
ground truth. usage: python tools/config2_full.py [n_scans=4541] > profiles/r2_config2_full.json"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kiss_icp_b200 as K
from kiss_icp_b200 import metrics as M, synthetic
from oracle import oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4541
chunk = 250
L = synthetic.kitti_shape(seed=0, device="cuda")
g = K.KissICP(K.load_config())
o = O.KissICP(max_num_threads=int(os.environ.get("ORACLE_THREADS", "16")))
gp, op, git, oit = [], [], [], []
t_gpu = t_cpu = 0.0
worst_t = worst_r = 0.0
for lo in range(0, n, chunk):
    hi = min(n, lo + chunk)
    scans = [L.scan_torch(k)[0].contiguous() for k in range(lo, hi)]
    torch.cuda.synchronize()
    g.start_history(hi - lo)
    t0 = time.perf_counter()
    poses = g._register_frames_raw([s.data_ptr() for s in scans], [s.shape[0] for s in scans], [None] * len(scans), [0] * len(scans), 2)
    t_gpu += time.perf_counter() - t0
    git += [h.iterations for h in g.history()]
    gp.append(poses)
    host = [s.cpu().numpy() for s in scans]
    t0 = time.perf_counter()
    for p in host:
        o.register_frame(p, np.empty(0), want_clouds=False)
        op.append(np.array(o.pose))
        oit.append(o.last_iterations)
    t_cpu += time.perf_counter() - t0
    print("scans", hi, "gpu %.0f scans/s" % (hi / t_gpu), "cpu %.1f scans/s" % (hi / t_cpu), file=sys.stderr, flush=True)
gp = np.concatenate(gp)
op = np.array(op)
dt = np.linalg.norm(gp[:, :3, 3] - op[:, :3, 3], axis=1)
R = np.einsum("nij,nik->njk", gp[:, :3, :3], op[:, :3, :3])
dr = np.arccos(np.clip((np.trace(R, axis1=1, axis2=2) - 1) / 2, -1, 1))
gt = np.array([L.pose(k) for k in range(n)])
gt = np.array([np.linalg.inv(gt[0]) @ x for x in gt])
out = {"workload": "KITTI-00-shape synthetic stream, %d scans (BASELINE config 2), free running" % n,
       "max_translation_diff_m": float(dt.max()), "max_rotation_diff_rad": float(dr.max()), "tolerance": "1e-4 m / 1e-4 rad (BASELINE north_star)",
       "iteration_count_mismatches": int(np.sum(np.array(git) != np.array(oit))), "mean_iterations": float(np.mean(git)),
       "gpu_scans_per_s_queued_resident_incl_host_loop": n / t_gpu, "oracle_scans_per_s": n / t_cpu, "oracle_threads": int(os.environ.get("ORACLE_THREADS", "16")),
       "path_m": float(np.linalg.norm(np.diff(gt[:, :3, 3], axis=0), axis=1).sum()),
       "drift_gpu": dict(zip(["seq_trans_pct", "seq_rot_deg_per_m", "ate_rot_rad", "ate_trans_m"], list(M.sequence_error(gt, gp)) + list(M.absolute_trajectory_error(gt, gp)))),
       "drift_oracle": dict(zip(["seq_trans_pct", "seq_rot_deg_per_m", "ate_rot_rad", "ate_trans_m"], list(M.sequence_error(gt, op)) + list(M.absolute_trajectory_error(gt, op)))),
       "grow_retries": g.grow_retries()}
print(json.dumps(out))
