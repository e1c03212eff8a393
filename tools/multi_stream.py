"""Developer aid (GPU box): aggregate scans/s of S independent sequences on one GPU (bench.py's multi_stream leg) for several S."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import kiss_icp_b200 as K
from kiss_icp_b200 import _native as N
args = types.SimpleNamespace(steps=20, repeats=3, prime=40, workload="kitti")
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
for S in [int(x) for x in sys.argv[1:]] or [2, 4, 8]:
    r = bench.multi_stream_leg(args, K, N, N.lib(), torch, dev, S, bench.pipeline_config("kitti"))
    print(S, {k: (round(v, 2) if isinstance(v, float) else v) for k, v in r.items() if k != "note"}, "smem", os.environ.get("KB_ICP_SMEM_KB", "96"), flush=True)
