"""NN-query sweep (BASELINE config 5): batched GetClosestNeighbor GB/s vs map size and query count.
CUDA-event timing on the stream the library launches on, L2 flushed (256 MiB write) before every timed launch.
  python tools/nn_bench.py [raw map sizes ...] > gpurun_out/nn_sweep.jsonl"""
import ctypes as C, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kiss_icp_b200 as K
from kiss_icp_b200 import _native as N, synthetic
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
stream = torch.cuda.Stream(device=dev)  # torch's default stream has a NULL handle: share an explicit one
torch.cuda.set_stream(stream)
N.check(N.lib().kb_set_stream(C.c_void_p(stream.cuda_stream)))
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
sizes = [int(x) for x in sys.argv[1:]] or [100_000, 200_000, 500_000, 1_000_000, 2_000_000, 5_000_000]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
L = N.lib()
for n_raw in sizes:
    m = K.VoxelHashMap(1.0, 1e9, 20)
    m.add_points(synthetic.surface_cloud(n_raw, seed=5))
    stored = torch.from_numpy(m.point_cloud())
    for n_q in (1 << 16, 1 << 17, 1 << 20):
        g = torch.Generator(device="cpu"); g.manual_seed(5)
        sel = stored[torch.randint(0, stored.shape[0], (n_q,), generator=g)]
        q = (sel + torch.randn(n_q, 3, generator=g, dtype=torch.float64) * 0.3).to(dev).contiguous()
        outp = torch.empty_like(q); outd = torch.empty(n_q, dtype=torch.float64, device=dev)
        b = C.c_double(0)
        N.check(L.kb_map_query_bytes_dev(m._h, C.c_void_p(q.data_ptr()), n_q, C.byref(b)))
        times = []
        for it in range(8):
            flush.fill_(it)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(stream)
            N.check(L.kb_map_closest_neighbors_dev(m._h, C.c_void_p(q.data_ptr()), n_q, C.c_void_p(outp.data_ptr()), C.c_void_p(outd.data_ptr())))
            e1.record(stream)
            torch.cuda.synchronize()
            if it >= 3: times.append(e0.elapsed_time(e1))
        ms = float(np.median(times))
        print(json.dumps({"map_points": int(stored.shape[0]), "voxels": m.num_voxels(), "queries": n_q, "bytes_per_query": b.value / n_q,
                          "ms": ms, "GBps": b.value / ms / 1e6, "frac_of_hbm_peak": b.value / ms / 1e6 / peak}), flush=True)
    del m
