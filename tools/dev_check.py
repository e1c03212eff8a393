"""Developer parity sweep (GPU box): every C-ABI entry point against the oracle."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import kiss_icp_b200 as K
from kiss_icp_b200 import synthetic
from oracle import oracle as O

rng = np.random.default_rng(0)
def hdr(s): print("\n== " + s, flush=True)

hdr("voxel_down_sample")
for n, vs, scale in [(0, 1.0, 1), (1, 1.0, 1), (2, 0.5, 1), (1000, 0.5, 10), (65536, 0.5, 60), (20000, 1.5, 80), (5000, 0.3, 3)]:
    pts = rng.normal(size=(n, 3)) * scale
    if n > 10:
        pts[:5] = np.round(pts[:5])  # exact multiples
        pts[5:8] = pts[2:5]          # duplicates
    a = K.voxel_down_sample(pts, vs); b = O.voxel_down_sample(pts, vs)
    print(n, vs, "out", len(a), len(b), "exact-ordered-equal:", a.shape == b.shape and np.array_equal(a, b))

lidar = synthetic.small_shape(seed=1, beams=32, cols=512)
pts, _ = lidar.scan(0)
a = K.voxel_down_sample(pts, 0.5); b = O.voxel_down_sample(pts, 0.5)
print("scan", len(pts), "->", len(a), "equal", np.array_equal(a, b))

hdr("preprocess")
P = K.Preprocessor(100.0, 0.0, True, 0)
ts = np.linspace(0, 1, len(pts))
T = O.se3_exp([0.9, 0.05, 0.01, 0.002, -0.001, 0.03])
a = P.preprocess(pts, np.empty(0), T); b = O.preprocess(pts, np.empty(0), T, 100.0, 0.0, True)
print("no stamps: equal", np.array_equal(a, b), len(a))
a = P.preprocess(pts, ts, T); b = O.preprocess(pts, ts, T, 100.0, 0.0, True)
print("deskew: n", len(a), len(b), "max diff", np.abs(a - b).max() if len(a) == len(b) else None)
P2 = K.Preprocessor(30.0, 5.0, False, 0)
a = P2.preprocess(pts, ts, T); b = O.preprocess(pts, ts, T, 30.0, 5.0, False)
print("crop 5..30: equal", np.array_equal(a, b), len(a))

hdr("map add/remove/dump")
gm = K.VoxelHashMap(1.0, 100.0, 20); om = O.VoxelHashMap(1.0, 100.0, 20)
for k in range(4):
    p, _ = lidar.scan(k)
    Tk = np.linalg.inv(lidar.pose(0)) @ lidar.pose(k)
    ds = O.voxel_down_sample(p, 0.5)
    gm.update(ds, Tk); om.update(ds, Tk)
def canon(vox, cnt, pts):
    order = np.lexsort((vox[:, 2], vox[:, 1], vox[:, 0]))
    starts = np.concatenate([[0], np.cumsum(cnt)])[:-1]
    return vox[order], cnt[order], np.concatenate([pts[starts[i]:starts[i] + cnt[i]] for i in order]) if len(order) else pts
gv, gc, gp = gm.dump(); ov, oc, op = canon(*om.dump())
print("voxels", len(gv), len(ov), "points", len(gp), len(op), "equal:", np.array_equal(gv, ov), np.array_equal(gc, oc), np.array_equal(gp, op))
gm.remove_far_away_points([30., 0, 0]); om.remove_far_away_points([30., 0, 0])
gv, gc, gp = gm.dump(); ov, oc, op = canon(*om.dump())
print("after remove: voxels", len(gv), len(ov), "equal:", np.array_equal(gv, ov), np.array_equal(gp, op))
raw, _ = lidar.scan(5)
gm.add_points(raw); om.add_points(raw)   # many points per voxel
gv, gc, gp = gm.dump(); ov, oc, op = canon(*om.dump())
print("after add_points(raw): voxels", len(gv), len(ov), "equal:", np.array_equal(gv, ov), np.array_equal(gc, oc), np.array_equal(gp, op))

hdr("closest neighbours")
q = O.se3_act(np.linalg.inv(lidar.pose(0)) @ lidar.pose(3), lidar.scan(3)[0]) + rng.normal(size=(len(lidar.scan(3)[0]), 3)) * 0.2
ap, ad = gm.closest_neighbors(q); bp, bd = om.closest_neighbors(q)
print("n", len(q), "points equal", np.array_equal(ap, bp), "dist equal", np.array_equal(ad, bd), "misses", int((bd > 1e300).sum()))

hdr("build_system / align")
reg = K.Registration(500, 1e-4, 0)
src = O.voxel_down_sample(O.voxel_down_sample(lidar.scan(4)[0], 0.5), 1.5)
T4 = np.linalg.inv(lidar.pose(0)) @ lidar.pose(4)
srcm = O.se3_act(T4, src)
A, b_, n1 = reg.build_system(srcm, gm, 3.0, 1.0); A2, b2, n2 = O.build_system(om, srcm, 3.0, 1.0, nthreads=1)
print("ncorr", n1, n2, "JTJ rel", np.abs(A - A2).max() / np.abs(A2).max(), "JTr rel", np.abs(b_ - b2).max() / np.abs(b2).max())
guess = T4 @ O.se3_exp([0.3, -0.2, 0.05, 0.01, -0.01, 0.02])
Pg = reg.align_points_to_map(src, gm, guess, 3.0, 1.0); Po, ito = O.align_points_to_map(om, src, guess, 3.0, 1.0)
print("align iters", reg.last_iterations, ito, "pose diff", np.abs(Pg - Po).max(), "vs truth", np.abs(Pg - T4).max())

hdr("pipeline (fused) stream")
for stamps in ("none", "column"):
    L = synthetic.small_shape(seed=5, beams=32, cols=512, stamps=stamps)
    g = K.KissICP(K.load_config()); o = O.KissICP()
    worst = 0
    for k in range(12):
        p, t = L.scan(k)
        g.register_frame(p, t, return_clouds=False); o.register_frame(p, t, want_clouds=False)
        d = np.abs(g.last_pose - o.pose).max(); worst = max(worst, d)
    print(stamps, "max |pose diff| over 12 scans", worst, "iters", g.last_iterations, o.last_iterations, "map", g.local_map.num_points(), o.local_map.num_points())

hdr("timing: KITTI-shape fused pipeline")
L = synthetic.kitti_shape(seed=0, device="cuda")
g = K.KissICP(K.load_config()); o = O.KissICP()
g.set_profiling(not os.environ.get('KB_NO_PROFILE'))
tg = to = 0; worst = 0
for k in range(30):
    p, t = L.scan(k)
    t0 = time.perf_counter(); g.register_frame(p, t, return_clouds=False); t1 = time.perf_counter()
    o.register_frame(p, t, want_clouds=False); t2 = time.perf_counter()
    if k >= 5: tg += t1 - t0; to += t2 - t1
    if k in (0, 1, 2, 10, 29): print(k, "wall ms %.3f" % ((t1 - t0) * 1e3), "phases us", np.round(g.last_profile_us, 1), "iters", g.last_iterations, flush=True)
    worst = max(worst, np.abs(g.last_pose - o.pose).max())
import ctypes as C
from kiss_icp_b200 import _native as N
ns = np.zeros(64 + 4 * 148); N.check(N.lib().kb_pipeline_debug_stamps(g._h, N.ptr(ns), len(ns)))
dp = np.zeros(10); N.check(N.lib().kb_pipeline_last_ds_profile(g._h, N.ptr(dp))); print('downsample split [us] (clear, dedupe, count, prefix+rank, replay+emit) ds1:', np.round(dp[:5], 1), 'ds2:', np.round(dp[5:], 1))
mp = np.zeros(3); N.check(N.lib().kb_pipeline_last_map_profile(g._h, N.ptr(mp))); print('map update split [us]: claim+lists, ordered insert, evict scan:', np.round(mp, 1))
cs = np.zeros(3); N.check(N.lib().kb_pipeline_last_cache_stats(g._h, N.ptr(cs))); print('NN cache hits/fills/overflows (last frame, all iterations):', cs, 'iters', g.last_iterations)
cyc = ns[16:29] - ns[16]
names = ["start", "queries done", "block synced", "partial posted", "all arrived", "reduced", "rec loaded+expanded", "ldlt", "exp", "mul", "published", "epoch seen", "record in smem"]
print("CTA0 iteration-4 timeline [us @1.965GHz]:", ", ".join("%s %.2f" % (n_, c / 1965.0) for n_, c in zip(names, cyc)))
cta = ns[64:].reshape(-1, 4)
it = ns[41:41 + min(g.last_iterations, 20)]; print('ICP iteration durations [us]:', np.round(np.diff(it) * 1e-3, 1))
# experiment: the same alignment twice through the stand-alone API (second call: map data L2-warm, same kernel)
reg2 = K.Registration(500, 1e-4, 0)
src2 = O.voxel_down_sample(O.voxel_down_sample(p, 0.5), 1.5)
guess2 = g.last_pose @ O.se3_exp([0.4, 0.05, 0.0, 0.0, 0.0, 0.01])
for rep in range(3):
    reg2.align_points_to_map(src2, g.local_map, guess2, 3.0, 1.0)
    N.check(N.lib().kb_pipeline_debug_stamps(g._h, N.ptr(ns), len(ns)))
    it = ns[41:41 + min(reg2.last_iterations, 20)]
    print('  stand-alone align rep', rep, 'n_src', len(src2), 'iters', reg2.last_iterations, 'iteration durations [us]:', np.round(np.diff(it) * 1e-3, 1))
print('gather rounds in the last ICP iteration of the last frame:', ns[40])
dur = cta[:, 0] - cta[:, 3]
order = np.argsort(-dur)[:12]
print('slowest CTAs (cta, start->posted ns, fills this iteration):', [(int(i), float(dur[i]), float(cta[i, 2] + ns[0])) for i in order], 'median', np.median(dur), 'start spread', cta[:, 3].max() - cta[:, 3].min())
print('per-CTA posted (globaltimer, relative to the earliest) min/median/max [ns]', cta[:, 0].min() - cta[:, 0].min(), np.median(cta[:, 0]) - cta[:, 0].min(), cta[:, 0].max() - cta[:, 0].min(), 'argmax', cta[:, 0].argmax())
b = C.c_double(0)
for it in (1, 10, 100):
    N.check(N.lib().kb_debug_barrier_ns(it, C.byref(b))); print("grid barrier ns (avg over %d):" % it, b.value)
print("gpu ms/scan", tg / 25 * 1e3, "cpu ms/scan", to / 25 * 1e3, "threads", O.num_threads(), "max pose diff", worst, "iters", g.last_iterations)
