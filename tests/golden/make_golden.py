"""Generates tests/golden/golden_v1.npz — known-answer vectors for the registration hot path.

The reference ships NO golden vectors for this path (python/tests/test_kiss_icp.py:1-4 is an
import check) and cannot be imported or compiled offline, so these vectors are produced by this
project's oracle (oracle/, a CPU restatement of the reference) and frozen: PARITY UNPINNED.
They pin (a) the oracle against silent drift and (b) the CUDA path against the oracle on inputs
that travel with the repo.  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from kiss_icp_b200 import synthetic  # noqa: E402
from oracle import oracle as O  # noqa: E402

out = {}
rng = np.random.default_rng(20260922)

# ---- stream A: KITTI-shape (no stamps), reduced to 16 beams x 256 columns, 10 scans
for tag, stamps in (("a", "none"), ("b", "column")):
    lidar = synthetic.small_shape(seed=7, beams=16, cols=256, stamps=stamps)
    icp = O.KissICP(max_num_threads=1)
    poses, iters, sigmas, npre, nsrc = [], [], [], [], []
    for k in range(10):
        pts, ts = lidar.scan(k)
        out[f"{tag}_scan{k}"] = pts.astype(np.float32)  # coords are fp32-representable by construction
        if stamps == "column":
            out[f"{tag}_ts{k}"] = ts
        pre, src = icp.register_frame(pts, ts)
        poses.append(icp.pose)
        iters.append(icp.last_iterations)
        sigmas.append(icp.sigma)
        npre.append(len(pre))
        nsrc.append(len(src))
        if k == 3:
            out[f"{tag}_pre3"] = pre
            out[f"{tag}_src3"] = src
    out[f"{tag}_poses"] = np.array(poses)
    out[f"{tag}_iters"] = np.array(iters)
    out[f"{tag}_sigma_next"] = np.array(sigmas)
    out[f"{tag}_npre"] = np.array(npre)
    out[f"{tag}_nsrc"] = np.array(nsrc)
    v, c, p = icp.local_map.dump()
    order = np.lexsort((v[:, 2], v[:, 1], v[:, 0]))
    starts = np.concatenate([[0], np.cumsum(c)])[:-1]
    out[f"{tag}_map_vox"] = v[order]
    out[f"{tag}_map_cnt"] = c[order]
    out[f"{tag}_map_pts"] = np.concatenate([p[starts[i]:starts[i] + c[i]] for i in order])

# ---- VoxelDownsample KAT (order matters)
cloud = rng.normal(size=(3000, 3)) * 12.0
cloud[:6] = np.round(cloud[:6])       # points exactly on voxel faces
cloud[6:9] = -cloud[6:9] - 0.0        # negatives
out["ds_in"] = cloud
out["ds_out_05"] = O.voxel_down_sample(cloud, 0.5)
out["ds_out_15"] = O.voxel_down_sample(out["ds_out_05"], 1.5)
out["ds_out_03"] = O.voxel_down_sample(cloud, 0.3)

# ---- map + NN KAT
m = O.VoxelHashMap(1.0, 100.0, 20)
m.add_points(cloud)
q = cloud[::3] + rng.normal(size=cloud[::3].shape) * 0.4
q = np.concatenate([q, [[500.0, 500.0, 500.0]]])  # guaranteed miss
nn_p, nn_d = m.closest_neighbors(q, nthreads=1)
out["nn_q"] = q
out["nn_p"] = nn_p
out["nn_d"] = nn_d
JTJ, JTr, nc = O.build_system(m, q[:-1], 1.5, 0.5, nthreads=1)
out["sys_JTJ"], out["sys_JTr"], out["sys_nc"] = JTJ, JTr, np.array(nc)

# ---- config-1 style KAT: one scan against a pre-built map, perturbed guess
lidar = synthetic.small_shape(seed=11, beams=32, cols=512)
mm = O.VoxelHashMap(1.0, 100.0, 20)
T0 = lidar.pose(0)
for k in range(6):
    p, _ = lidar.scan(k)
    mm.update(O.voxel_down_sample(p, 0.5), np.linalg.inv(T0) @ lidar.pose(k))
p6, _ = lidar.scan(6)
src = O.voxel_down_sample(O.voxel_down_sample(p6, 0.5), 1.5)
gt = np.linalg.inv(T0) @ lidar.pose(6)
guess = gt @ O.se3_exp([0.3, -0.1, 0.02, 0.0, 0.0, np.deg2rad(1.0)])
pose, it = O.align_points_to_map(mm, src, guess, 3.0, 1.0, nthreads=1)
v, c, pp = mm.dump()
out["c1_map_pts"] = pp  # concatenated in reference iteration order; re-adding reproduces per-voxel order
out["c1_src"] = src
out["c1_guess"] = guess
out["c1_gt"] = gt
out["c1_pose"] = pose
out["c1_iters"] = np.array(it)

path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path) // 1024, "KiB;", "c1 iters", it, "c1 err vs gt", np.abs(pose - gt).max())
