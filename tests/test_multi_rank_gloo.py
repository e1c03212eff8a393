"""world_size-2 gloo test of the multi-GPU layout (independent sequences, pose all_gather).
No GPU compute: ranks carry stand-in trajectories produced by the oracle on tiny scans."""
import os
import socket
import subprocess
import sys

import numpy as np

from conftest import ROOT

WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from kiss_icp_b200 import sharding, synthetic
from oracle import oracle as O
rank, world, _ = sharding.rank_world()
dist.init_process_group("gloo", rank=rank, world_size=world)
seqs = sharding.sequences_of_rank(rank, world, world)
assert seqs == [rank]
lidar = synthetic.small_shape(seed=seqs[0], beams=8, cols=128)
icp = O.KissICP(max_num_threads=1)
poses = []
for k in range(4):
    p, t = lidar.scan(k)
    icp.register_frame(p, t, want_clouds=False)
    poses.append(icp.pose)
allp = sharding.gather_poses(np.array(poses))
assert allp.shape == (world, 4, 4, 4)
assert np.array_equal(allp[rank], np.array(poses))
t = sharding.max_over_ranks(float(rank + 1))
s = sharding.sum_over_ranks(float(rank + 1))
assert t == world and s == world * (world + 1) / 2
if rank == 0:
    np.save(sys.argv[2], allp)
dist.destroy_process_group()
'''


def test_two_ranks_gather_independent_trajectories(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    out = tmp_path / "poses.npy"
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script), ROOT, str(out)], env=env))
    for p in procs:
        assert p.wait(timeout=240) == 0
    allp = np.load(out)
    assert allp.shape == (2, 4, 4, 4)
    assert np.array_equal(allp[0, 0], np.eye(4)) and np.array_equal(allp[1, 0], np.eye(4))
    assert not np.array_equal(allp[0, 3], allp[1, 3])  # different seeds -> different trajectories


def test_sequence_assignment():
    from kiss_icp_b200 import sharding
    assert sharding.sequences_of_rank(0, 1, 1) == [0]
    assert [sharding.sequences_of_rank(r, 4, 8) for r in range(4)] == [[0, 4], [1, 5], [2, 6], [3, 7]]
    got = sorted(s for r in range(8) for s in sharding.sequences_of_rank(r, 8, 8))
    assert got == list(range(8))
