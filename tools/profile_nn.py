"""Driver for ncu: one batched GetClosestNeighbor launch (BASELINE config 5 scale)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kiss_icp_b200 as K
from kiss_icp_b200 import _native as N, synthetic
dev = torch.device("cuda", 0)
m = K.VoxelHashMap(1.0, 1e9, 20)
m.add_points(synthetic.surface_cloud(4_000_000, seed=5))
stored = torch.from_numpy(m.point_cloud())
g = torch.Generator(device="cpu"); g.manual_seed(5)
n_q = 1 << 20
sel = stored[torch.randint(0, stored.shape[0], (n_q,), generator=g)]
q = (sel + torch.randn(n_q, 3, generator=g, dtype=torch.float64) * 0.3).to(dev).contiguous()
outp = torch.empty_like(q); outd = torch.empty(n_q, dtype=torch.float64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for it in range(3):
    flush.fill_(it)
    N.check(N.lib().kb_map_closest_neighbors_dev(m._h, C.c_void_p(q.data_ptr()), n_q, C.c_void_p(outp.data_ptr()), C.c_void_p(outd.data_ptr())))
    N.check(N.lib().kb_map_sync(m._h))
print("map points", stored.shape[0], "voxels", m.num_voxels(), "hits", int((outd < 1e300).sum()))
