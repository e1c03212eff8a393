"""Target for `compute-sanitizer --tool initcheck`: a short free-running stream that starts from the minimal voxel table
(KB_MAP_RESERVE_SLOTS=0: vetoes, growth into fresh allocations, rebuilds), with and without timestamps, blocking and queued."""
import os, sys
os.environ.setdefault("KB_MAP_RESERVE_SLOTS", "0")
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kiss_icp_b200 as K
from kiss_icp_b200 import synthetic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
only_first = len(sys.argv) > 2 and sys.argv[2] == "long"  # the 300-scan stream of test_pipeline_long_stream_with_table_rebuilds alone
L = synthetic.small_shape(seed=21, beams=32, cols=512)
g = K.KissICP(K.load_config(max_range=40.0))
for k in range(n):
    p, t = L.scan(k)
    g.register_frame(p, t, return_clouds=False)
if only_first:
    print("done", g.last_iterations, g.local_map.num_points(), "grow retries", g.grow_retries())
    sys.exit(0)
L2 = synthetic.small_shape(seed=22, beams=32, cols=512)
h = K.KissICP(K.load_config(max_range=40.0))
h.register_frames([L2.scan(k)[0] for k in range(n)], None)
L3 = synthetic.small_shape(seed=23, beams=32, cols=512, stamps="column")
d = K.KissICP(K.load_config(max_range=40.0))
for k in range(max(3, n // 2)):
    p, t = L3.scan(k)
    d.register_frame(p, t, return_clouds=False)
m = K.VoxelHashMap(1.0, 100.0, 20)
pts = np.random.default_rng(0).normal(size=(20000, 3)) * 10
m.add_points(pts)
m.closest_neighbors(pts[:4096] + 0.05)
print("done", g.last_iterations, h.last_iterations, m.num_points())
