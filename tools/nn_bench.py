"""NN-query sweep (BASELINE config 5): batched GetClosestNeighbor GB/s vs map size."""
import ctypes as C, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kiss_icp_b200 as K
from kiss_icp_b200 import _native as N, synthetic
dev = torch.device("cuda", 0)
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
sizes = [int(x) for x in sys.argv[1:]] or [200_000, 1_000_000, 4_000_000, 8_000_000]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
L = N.lib()
for n_raw in sizes:
    m = K.VoxelHashMap(1.0, 1e9, 20)
    m.add_points(synthetic.surface_cloud(n_raw, seed=5))
    if os.environ.get('KB_COMPACT'): m.compact()
    stored = torch.from_numpy(m.point_cloud())
    g = torch.Generator(device="cpu"); g.manual_seed(5)
    n_q = 1 << 20
    sel = stored[torch.randint(0, stored.shape[0], (n_q,), generator=g)]
    q = (sel + torch.randn(n_q, 3, generator=g, dtype=torch.float64) * 0.3).to(dev).contiguous()
    outp = torch.empty_like(q); outd = torch.empty(n_q, dtype=torch.float64, device=dev)
    b = C.c_double(0)
    N.check(L.kb_map_query_bytes_dev(m._h, C.c_void_p(q.data_ptr()), n_q, C.byref(b)))
    times = []
    for it in range(9):
        flush.fill_(it)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # the library launches on its own stream here: bracket with host sync + library sync (coarse but fair)
        import time
        t0 = time.perf_counter()
        N.check(L.kb_map_closest_neighbors_dev(m._h, C.c_void_p(q.data_ptr()), n_q, C.c_void_p(outp.data_ptr()), C.c_void_p(outd.data_ptr())))
        N.check(L.kb_map_sync(m._h))
        if it >= 2: times.append((time.perf_counter() - t0) * 1e3)
    ms = float(np.median(times))
    print(json.dumps({"map_points": int(stored.shape[0]), "voxels": m.num_voxels(), "queries": n_q, "bytes_per_query": b.value / n_q,
                      "ms_wall": ms, "GBps": b.value / ms / 1e6, "frac_of_hbm_peak": b.value / ms / 1e6 / peak}))
    del m
