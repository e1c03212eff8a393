// ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the shipped product path.
// Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this.
//
// extern "C" surface over the CPU restatement (oracle_core.hpp) so that tests and the
// bench can drive it through ctypes. Shapes mirror include/kiss_icp_b200.h so a parity
// test reads "same call, two backends". 4x4 transforms are row-major double[16].
#include <cstring>
#include <vector>

#include "oracle_core.hpp"

using namespace oracle;

namespace {
Points to_points(const double *xyz, long n) {
    Points p(static_cast<size_t>(n));
    for (long i = 0; i < n; ++i) p[i] = Vec3{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    return p;
}
void from_points(const Points &p, double *xyz) {
    for (size_t i = 0; i < p.size(); ++i) {
        xyz[3 * i] = p[i].x;
        xyz[3 * i + 1] = p[i].y;
        xyz[3 * i + 2] = p[i].z;
    }
}
struct PipelineHandle {
    explicit PipelineHandle(const KISSConfig &c) : icp(c) {}
    KissICP icp;
    Points last_pre, last_src;
};
}  // namespace

extern "C" {

// ---------------------------------------------------------------- math
int oracle_se3_exp(const double a[6], double M[16]) {
    se3_to_matrix(se3_exp(a), M);
    return 0;
}
int oracle_se3_log(const double M[16], double a[6]) {
    SE3 T;
    if (!se3_from_matrix(M, &T)) return 1;
    se3_log(T, a);
    return 0;
}
int oracle_se3_mul(const double A[16], const double B[16], double C[16]) {
    SE3 a, b;
    if (!se3_from_matrix(A, &a) || !se3_from_matrix(B, &b)) return 1;
    se3_to_matrix(se3_mul(a, b), C);
    return 0;
}
int oracle_se3_inverse(const double A[16], double C[16]) {
    SE3 a;
    if (!se3_from_matrix(A, &a)) return 1;
    se3_to_matrix(se3_inverse(a), C);
    return 0;
}
int oracle_se3_act(const double A[16], const double *xyz, long n, double *out) {
    SE3 a;
    if (!se3_from_matrix(A, &a)) return 1;
    for (long i = 0; i < n; ++i) {
        const Vec3 r = se3_act(a, Vec3{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]});
        out[3 * i] = r.x;
        out[3 * i + 1] = r.y;
        out[3 * i + 2] = r.z;
    }
    return 0;
}
void oracle_ldlt6_solve(const double A[36], const double b[6], double x[6]) { ldlt6_solve(A, b, x); }

// ---------------------------------------------------------------- voxel utils
void oracle_point_to_voxel(const double *xyz, long n, double voxel_size, int *out) {
    for (long i = 0; i < n; ++i) {
        const Voxel v = PointToVoxel(Vec3{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, voxel_size);
        out[3 * i] = v.x;
        out[3 * i + 1] = v.y;
        out[3 * i + 2] = v.z;
    }
}
unsigned oracle_voxel_hash(int x, int y, int z) { return voxel_hash(Voxel{x, y, z}); }

// out must hold n points; returns number written
long oracle_voxel_downsample(const double *xyz, long n, double voxel_size, double *out) {
    const Points r = VoxelDownsample(to_points(xyz, n), voxel_size);
    from_points(r, out);
    return static_cast<long>(r.size());
}

// ---------------------------------------------------------------- VoxelHashMap
void *oracle_map_create(double voxel_size, double max_distance, unsigned max_points_per_voxel) {
    return new VoxelHashMap(voxel_size, max_distance, max_points_per_voxel);
}
void oracle_map_destroy(void *m) { delete static_cast<VoxelHashMap *>(m); }
void oracle_map_clear(void *m) { static_cast<VoxelHashMap *>(m)->Clear(); }
int oracle_map_empty(void *m) { return static_cast<VoxelHashMap *>(m)->Empty() ? 1 : 0; }
long oracle_map_num_voxels(void *m) { return static_cast<long>(static_cast<VoxelHashMap *>(m)->NumVoxels()); }
void oracle_map_add_points(void *m, const double *xyz, long n) {
    static_cast<VoxelHashMap *>(m)->AddPoints(to_points(xyz, n));
}
void oracle_map_remove_far(void *m, const double origin[3]) {
    static_cast<VoxelHashMap *>(m)->RemovePointsFarFromLocation(Vec3{origin[0], origin[1], origin[2]});
}
void oracle_map_update_origin(void *m, const double *xyz, long n, const double origin[3]) {
    static_cast<VoxelHashMap *>(m)->Update(to_points(xyz, n), Vec3{origin[0], origin[1], origin[2]});
}
int oracle_map_update_pose(void *m, const double *xyz, long n, const double pose[16]) {
    SE3 T;
    if (!se3_from_matrix(pose, &T)) return 1;
    static_cast<VoxelHashMap *>(m)->Update(to_points(xyz, n), T);
    return 0;
}
long oracle_map_num_points(void *m) {
    long c = 0;
    for (const auto &b : static_cast<VoxelHashMap *>(m)->map_.buckets())
        if (!b.empty()) c += static_cast<long>(b.value.size());
    return c;
}
// out must hold oracle_map_num_points(m) points; reference iteration order
long oracle_map_pointcloud(void *m, double *out) {
    const Points p = static_cast<VoxelHashMap *>(m)->Pointcloud();
    from_points(p, out);
    return static_cast<long>(p.size());
}
// per-voxel dump: voxels[3*v..], counts[v], points concatenated; returns voxel count
long oracle_map_dump(void *m, int *voxels, int *counts, double *points) {
    long v = 0, p = 0;
    for (const auto &b : static_cast<VoxelHashMap *>(m)->map_.buckets()) {
        if (b.empty()) continue;
        voxels[3 * v] = b.key.x;
        voxels[3 * v + 1] = b.key.y;
        voxels[3 * v + 2] = b.key.z;
        counts[v] = static_cast<int>(b.value.size());
        for (const auto &q : b.value) {
            points[3 * p] = q.x;
            points[3 * p + 1] = q.y;
            points[3 * p + 2] = q.z;
            ++p;
        }
        ++v;
    }
    return v;
}
void oracle_map_closest_neighbors(void *m, const double *queries, long n, double *out_points, double *out_dist,
                                  int nthreads) {
    const VoxelHashMap *map = static_cast<const VoxelHashMap *>(m);
    (void)nthreads;
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : omp_get_max_threads())
    for (long i = 0; i < n; ++i) {
        const auto [p, d] = map->GetClosestNeighbor(Vec3{queries[3 * i], queries[3 * i + 1], queries[3 * i + 2]});
        out_points[3 * i] = p.x;
        out_points[3 * i + 1] = p.y;
        out_points[3 * i + 2] = p.z;
        out_dist[i] = d;
    }
}

// ---------------------------------------------------------------- Registration
// points are expected ALREADY in the map frame (what BuildLinearSystem sees inside the loop)
void oracle_build_system(void *m, const double *xyz, long n, double max_dist, double kernel, int nthreads,
                         double JTJ[36], double JTr[6], int *n_corr) {
    Registration reg(1, 1e-4, nthreads);
    const auto ls = reg.BuildSystem(to_points(xyz, n), *static_cast<VoxelHashMap *>(m), max_dist, kernel, n_corr);
    std::memcpy(JTJ, ls.JTJ, sizeof(ls.JTJ));
    std::memcpy(JTr, ls.JTr, sizeof(ls.JTr));
}
int oracle_align_points_to_map(void *m, const double *xyz, long n, const double guess[16], double max_dist,
                               double kernel, int max_iter, double conv, int nthreads, double out_pose[16],
                               int *iterations) {
    SE3 g;
    if (!se3_from_matrix(guess, &g)) return 1;
    Registration reg(max_iter, conv, nthreads);
    const SE3 r = reg.AlignPointsToMap(to_points(xyz, n), *static_cast<VoxelHashMap *>(m), g, max_dist, kernel,
                                       iterations);
    se3_to_matrix(r, out_pose);
    return 0;
}

// ---------------------------------------------------------------- Preprocessor
// returns number of points written, -1 = bad motion matrix, -2 = std::out_of_range analogue
long oracle_preprocess(const double *xyz, long n, const double *timestamps, long n_ts, const double motion[16],
                       double max_range, double min_range, int deskew, double *out) {
    SE3 T;
    if (!se3_from_matrix(motion, &T)) return -1;
    Preprocessor pre(max_range, min_range, deskew != 0, 0);
    std::vector<double> ts(timestamps, timestamps + n_ts);
    Points r;
    if (!pre.Preprocess(to_points(xyz, n), ts, T, &r)) return -2;
    from_points(r, out);
    return static_cast<long>(r.size());
}

void oracle_correct_kitti_scan(const double *xyz, long n, double *out) { from_points(CorrectKITTIScan(to_points(xyz, n)), out); }

// ---------------------------------------------------------------- AdaptiveThreshold
int oracle_threshold_update(double *model_sse, int *num_samples, const double deviation[16], double min_motion_th,
                            double max_range) {
    SE3 T;
    if (!se3_from_matrix(deviation, &T)) return 1;
    AdaptiveThreshold th(1.0, min_motion_th, max_range);
    th.model_sse_ = *model_sse;
    th.num_samples_ = *num_samples;
    th.UpdateModelDeviation(T);
    *model_sse = th.model_sse_;
    *num_samples = th.num_samples_;
    return 0;
}

// ---------------------------------------------------------------- KissICP pipeline
void *oracle_pipeline_create(double voxel_size, double max_range, double min_range, int max_points_per_voxel,
                             double min_motion_th, double initial_threshold, int max_num_iterations,
                             double convergence_criterion, int max_num_threads, int deskew) {
    KISSConfig c;
    c.voxel_size = voxel_size;
    c.max_range = max_range;
    c.min_range = min_range;
    c.max_points_per_voxel = max_points_per_voxel;
    c.min_motion_th = min_motion_th;
    c.initial_threshold = initial_threshold;
    c.max_num_iterations = max_num_iterations;
    c.convergence_criterion = convergence_criterion;
    c.max_num_threads = max_num_threads;
    c.deskew = deskew != 0;
    return new PipelineHandle(c);
}
void oracle_pipeline_destroy(void *p) { delete static_cast<PipelineHandle *>(p); }
// returns 0 ok; fills n_pre / n_src (sizes of the clouds RegisterFrame returns)
int oracle_pipeline_register_frame(void *p, const double *xyz, long n, const double *timestamps, long n_ts,
                                   long *n_pre, long *n_src) {
    auto *h = static_cast<PipelineHandle *>(p);
    std::vector<double> ts(timestamps, timestamps + n_ts);
    if (!h->icp.RegisterFrame(to_points(xyz, n), ts, &h->last_pre, &h->last_src)) return 2;
    if (n_pre) *n_pre = static_cast<long>(h->last_pre.size());
    if (n_src) *n_src = static_cast<long>(h->last_src.size());
    return 0;
}
void oracle_pipeline_last_clouds(void *p, double *pre, double *src) {
    auto *h = static_cast<PipelineHandle *>(p);
    if (pre) from_points(h->last_pre, pre);
    if (src) from_points(h->last_src, src);
}
void oracle_pipeline_pose(void *p, double M[16]) { se3_to_matrix(static_cast<PipelineHandle *>(p)->icp.last_pose_, M); }
void oracle_pipeline_delta(void *p, double M[16]) { se3_to_matrix(static_cast<PipelineHandle *>(p)->icp.last_delta_, M); }
double oracle_pipeline_sigma(void *p) { return static_cast<PipelineHandle *>(p)->icp.adaptive_threshold_.ComputeThreshold(); }
int oracle_pipeline_last_iterations(void *p) { return static_cast<PipelineHandle *>(p)->icp.last_iterations_; }
void *oracle_pipeline_map(void *p) { return &static_cast<PipelineHandle *>(p)->icp.local_map_; }
int oracle_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}


// a trace of map operations on the robin_map emulation -> iteration order afterwards (tests/test_oracle_robin_fuzz.py
// replays the same trace on an independent Python emulation). ops[i] = {op, x, y, z}: 0 insert (value = i), 1 erase(find(key)),
// 2 clear(), 3 reserve(x). out_keys: [cap][3] keys in iteration order, out_vals: [cap] their values.
long oracle_robin_trace(const int *ops, long n_ops, int *out_keys, int *out_vals, long cap, long *bucket_count) {
    oracle::RobinMap<int> m;
    for (long i = 0; i < n_ops; ++i) {
        const int *o = ops + 4 * i;
        const oracle::Voxel key{o[1], o[2], o[3]};
        if (o[0] == 0) {
            m.insert(key, static_cast<int>(i));
        } else if (o[0] == 1) {
            const size_t ib = m.find_index(key);
            if (m.bucket_count() && ib < m.bucket_count()) m.erase_at(ib);
        } else if (o[0] == 2) {
            m.clear();
        } else if (o[0] == 3) {
            m.reserve(static_cast<size_t>(o[1]));
        }
    }
    long k = 0;
    for (size_t ib = m.next_occupied(0); ib < m.bucket_count(); ib = m.next_occupied(ib + 1)) {
        if (k < cap) {
            out_keys[3 * k] = m.buckets()[ib].key.x;
            out_keys[3 * k + 1] = m.buckets()[ib].key.y;
            out_keys[3 * k + 2] = m.buckets()[ib].key.z;
            out_vals[k] = m.buckets()[ib].value;
        }
        ++k;
    }
    if (bucket_count) *bucket_count = static_cast<long>(m.bucket_count());
    return k;
}

}  // extern "C"
