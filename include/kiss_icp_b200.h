/* kiss_icp_b200.h — C-ABI of the B200-native KISS-ICP registration hot path.
 *
 * This is the drop-in boundary: one entry point per function the reference's language
 * bindings reach on this path (python/kiss_icp/pybind/kiss_icp_pybind.cpp:48-144 and the C++
 * surface used by ros/src/OdometryServer.cpp:80,162,165,222). Each declaration cites the
 * reference interface it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - plain pointers + sizes, no C++/torch types; every call returns a kb_status (0 = ok) and
 *     never throws or aborts across the ABI; kb_last_error() gives a thread-local message.
 *   - point clouds are dense double[n][3] (== std::vector<Eigen::Vector3d>::data(), 24 B stride)
 *   - rigid transforms are 4x4 ROW-MAJOR double[16] (numpy layout). Eigen/Sophus callers pass
 *     T.matrix().transpose().data() or copy (Eigen default is column-major).
 *   - a matrix that is not in SE(3) returns KB_ERR_NOT_SE3 where Sophus would SOPHUS_ENSURE
 *     (abort) in the reference (kiss_icp_pybind.cpp:68,84,99,117).
 *   - "host" entry points take HOST buffers and do the H2D/D2H copies themselves;
 *     "_dev" entry points take DEVICE pointers resident in HBM on the handle's device.
 *   - handles are not thread-safe (same as the reference objects); distinct handles may be
 *     used from distinct threads / on distinct GPUs concurrently.
 *   - there is NO CPU fallback: every compute entry point fails with KB_ERR_NO_DEVICE when no
 *     CUDA device is present.
 */
#ifndef KISS_ICP_B200_H
#define KISS_ICP_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum kb_status {
    KB_OK = 0,
    KB_ERR_INVALID_ARG = 1,
    KB_ERR_NOT_SE3 = 2,      /* Sophus SE3d(Matrix4d) would have ENSURE-failed */
    KB_ERR_OUT_OF_RANGE = 3, /* timestamps.at(idx) would have thrown (core/Preprocessing.cpp:76-77) */
    KB_ERR_CUDA = 4,
    KB_ERR_CAPACITY = 5, /* caller buffer too small; *n_out holds the required count */
    KB_ERR_NO_DEVICE = 6
} kb_status;

typedef struct kb_map kb_map;                   /* kiss_icp::VoxelHashMap            */
typedef struct kb_registration kb_registration; /* kiss_icp::Registration            */
typedef struct kb_preprocessor kb_preprocessor; /* kiss_icp::Preprocessor            */
typedef struct kb_threshold kb_threshold;       /* kiss_icp::AdaptiveThreshold       */
typedef struct kb_pipeline kb_pipeline;         /* kiss_icp::pipeline::KissICP       */

/* ---- library -------------------------------------------------------------------------- */
const char *kb_last_error(void);
const char *kb_version(void);
/* number of visible CUDA devices (0 when none: every compute call then returns KB_ERR_NO_DEVICE) */
int kb_device_count(void);
/* device used by handles created afterwards on the calling thread (default 0) */
int kb_set_device(int device);
/* run the handle-less calls (kb_voxel_down_sample, kb_preprocess ...) and handles created
 * afterwards on this CUDA stream (a cudaStream_t; NULL = the library's own stream) */
int kb_set_stream(void *cuda_stream);
/* persistent-grid size (thread blocks) used by the cooperative kernels; 0 = one per SM */
int kb_set_grid_blocks(int blocks);

/* ---- kiss_icp::VoxelHashMap (cpp/kiss_icp/core/VoxelHashMap.hpp:38-57) ------------------ */
/* VoxelHashMap(double voxel_size, double max_distance, unsigned max_points_per_voxel); pybind
 * `_VoxelHashMap(voxel_size, max_distance, max_points_per_voxel)` kiss_icp_pybind.cpp:54-57 */
int kb_map_create(double voxel_size, double max_distance, unsigned max_points_per_voxel, kb_map **out);
int kb_map_destroy(kb_map *map);
/* Clear() VoxelHashMap.hpp:44 / `_clear` */
int kb_map_clear(kb_map *map);
/* Empty() VoxelHashMap.hpp:45 / `_empty` */
int kb_map_empty(const kb_map *map, int *out_is_empty);
/* Update(points, origin) VoxelHashMap.cpp:83-87 / `_update(points, origin)` */
int kb_map_update_origin(kb_map *map, const double *xyz, size_t n, const double origin[3]);
/* Update(points, pose) VoxelHashMap.cpp:89-95 / `_update(points, pose)` */
int kb_map_update_pose(kb_map *map, const double *xyz, size_t n, const double pose[16]);
/* AddPoints VoxelHashMap.cpp:97-119 / `_add_points` */
int kb_map_add_points(kb_map *map, const double *xyz, size_t n);
/* RemovePointsFarFromLocation VoxelHashMap.cpp:121-132 / `_remove_far_away_points` */
int kb_map_remove_far(kb_map *map, const double origin[3]);
/* Pointcloud() VoxelHashMap.cpp:72-81 / `_point_cloud`. Two-call: pass capacity 0 to query
 * *n_out. Per-voxel point order is the reference's (insertion order); the order of VOXELS is
 * ascending (x,y,z) instead of the reference's hash-table iteration order (documented in
 * that order is an artefact of robin_map history and feeds nothing on this path). */
int kb_map_pointcloud(const kb_map *map, double *out_xyz, size_t capacity, size_t *n_out);
int kb_map_num_points(const kb_map *map, size_t *out);
int kb_map_num_voxels(const kb_map *map, size_t *out);
/* per-voxel dump (checkpoint / parity tests): voxels int[nv][3] ascending, counts int[nv],
 * points double[np][3] concatenated in per-voxel insertion order */
int kb_map_dump(const kb_map *map, int *voxels, int *counts, double *points, size_t voxel_capacity,
                size_t point_capacity, size_t *n_voxels, size_t *n_points);
/* GetClosestNeighbor VoxelHashMap.cpp:46-70, batched over n queries (the reference calls it
 * from a TBB parallel_for, Registration.cpp:66-76). out_points double[n][3], out_dist
 * double[n]; a miss returns (0,0,0), DBL_MAX like VoxelHashMap.cpp:51-52. */
int kb_map_closest_neighbors(const kb_map *map, const double *queries, size_t n, double *out_points,
                             double *out_dist);
int kb_map_closest_neighbors_dev(const kb_map *map, const double *d_queries, size_t n, double *d_out_points,
                                 double *d_out_dist);
/* rebuild the table at load factor <= 0.5 without tombstones (smallest power-of-two capacity): shrinks the
 * address range the point blocks are scattered over — for query-heavy phases on a static map */
int kb_map_compact(kb_map *map);
/* public data members voxel_size_, max_distance_, max_points_per_voxel_ VoxelHashMap.hpp:53-55 */
int kb_map_params(const kb_map *map, double *voxel_size, double *max_distance, unsigned *max_points_per_voxel);
/* ALGORITHMIC bytes of kb_map_closest_neighbors_dev for these n queries on this map: sum over
 * queries of 24 + 27*16 + 24*candidates + 32, candidates counted on the device
 * by an instrumented (untimed) pass. */
int kb_map_query_bytes_dev(const kb_map *map, const double *d_queries, size_t n, double *bytes);
/* kb_map_closest_neighbors_dev is asynchronous on the map's stream; this waits for it */
int kb_map_sync(const kb_map *map);

/* ---- kiss_icp::Registration (cpp/kiss_icp/core/Registration.hpp:33-45) ------------------ */
/* Registration(int max_num_iteration, double convergence_criterion, int max_num_threads);
 * max_num_threads is accepted and ignored (it sizes the TBB pool in the reference). */
int kb_registration_create(int max_num_iterations, double convergence_criterion, int max_num_threads,
                           kb_registration **out);
int kb_registration_destroy(kb_registration *reg);
/* AlignPointsToMap(frame, voxel_map, initial_guess, max_correspondence_distance, kernel_scale)
 * Registration.cpp:138-167 / `_align_points_to_map` kiss_icp_pybind.cpp:94-106 */
int kb_registration_align_points_to_map(kb_registration *reg, const double *xyz, size_t n, const kb_map *map,
                                        const double initial_guess[16], double max_correspondence_distance,
                                        double kernel_scale, double out_pose[16]);
/* number of ICP iterations the last call ran (diagnostic; no reference counterpart) */
int kb_registration_last_iterations(const kb_registration *reg, int *out);
/* one DataAssociation + BuildLinearSystem pass (Registration.cpp:60-121) on points already in
 * the map frame: JTJ double[36] row-major (full symmetric), JTr double[6], correspondences */
int kb_registration_build_system(kb_registration *reg, const double *xyz, size_t n, const kb_map *map,
                                 double max_correspondence_distance, double kernel_scale, double JTJ[36],
                                 double JTr[6], int *n_correspondences);

/* ---- kiss_icp::Preprocessor (cpp/kiss_icp/core/Preprocessing.hpp:32-45) ----------------- */
int kb_preprocessor_create(double max_range, double min_range, int deskew, int max_num_threads,
                           kb_preprocessor **out);
int kb_preprocessor_destroy(kb_preprocessor *pre);
/* Preprocess(frame, timestamps, relative_motion) Preprocessing.cpp:55-95 / `_preprocess`.
 * out_xyz double[capacity][3]; *n_out = points kept (order preserved). */
int kb_preprocessor_preprocess(kb_preprocessor *pre, const double *xyz, size_t n, const double *timestamps,
                               size_t n_timestamps, const double relative_motion[16], double *out_xyz,
                               size_t capacity, size_t *n_out);

/* ---- kiss_icp::AdaptiveThreshold (cpp/kiss_icp/core/Threshold.hpp:29-47) ---------------- */
int kb_threshold_create(double initial_threshold, double min_motion_threshold, double max_range,
                        kb_threshold **out);
int kb_threshold_destroy(kb_threshold *th);
/* ComputeThreshold() Threshold.hpp:38 / `_compute_threshold` */
int kb_threshold_compute(const kb_threshold *th, double *out_sigma);
/* UpdateModelDeviation(SE3) Threshold.cpp:38-49 / `_update_model_deviation` */
int kb_threshold_update_model_deviation(kb_threshold *th, const double model_deviation[16]);

/* ---- kiss_icp::VoxelDownsample (cpp/kiss_icp/core/VoxelUtils.cpp:7-21) ------------------ */
/* `_voxel_down_sample(frame, voxel_size)` kiss_icp_pybind.cpp:123. Output ORDER is the
 * reference's (robin_map iteration order) — it is result-affecting downstream. */
int kb_voxel_down_sample(const double *xyz, size_t n, double voxel_size, double *out_xyz, size_t capacity,
                         size_t *n_out);

/* ---- `_correct_kitti_scan(frame)` (python/kiss_icp/pybind/kiss_icp_pybind.cpp:127-138) ---- */
/* The KITTI loader's per-point intrinsic correction (datasets/kitti.py:44-48,68): every point is rotated by
 * 0.205 deg about normalized(pt x e_z). out_xyz may alias xyz. The _dev form works on device buffers and is
 * asynchronous on the calling thread's stream (kb_set_stream), so it chains into kb_pipeline_register_frame_dev. */
int kb_correct_kitti_scan(const double *xyz, size_t n, double *out_xyz);
int kb_correct_kitti_scan_dev(const double *d_xyz, size_t n, double *d_out_xyz);

/* ---- kiss_icp::pipeline::KissICP (cpp/kiss_icp/pipeline/KissICP.hpp:36-96) -------------- */
/* KISSConfig KissICP.hpp:36-54, field for field */
typedef struct kb_config {
    double voxel_size;           /* 1.0   */
    double max_range;            /* 100.0 */
    double min_range;            /* 0.0   */
    int max_points_per_voxel;    /* 20    */
    double min_motion_th;        /* 0.1   */
    double initial_threshold;    /* 2.0   */
    int max_num_iterations;      /* 500   */
    double convergence_criterion; /* 1e-4 */
    int max_num_threads;         /* 0 (ignored on the GPU) */
    int deskew;                  /* 1     */
} kb_config;
void kb_config_default(kb_config *cfg);

int kb_pipeline_create(const kb_config *cfg, kb_pipeline **out);
int kb_pipeline_destroy(kb_pipeline *p);
/* RegisterFrame(frame, timestamps) KissICP.cpp:35-68. The two clouds the reference returns by
 * value (preprocessed_frame, source) stay on the device; fetch them with
 * kb_pipeline_last_clouds only when wanted (ROS publishes them only for debugging,
 * ros/src/OdometryServer.cpp:168-172). */
int kb_pipeline_register_frame(kb_pipeline *p, const double *xyz, size_t n, const double *timestamps,
                               size_t n_timestamps);
/* same with a float32 frame float[n][3] (KITTI .bin / PointCloud2 payloads are float32 and the reference widens
 * them on the host: python/kiss_icp/datasets/kitti.py:66, ros/src/Utils.hpp:198-208). Half the H2D bytes; the
 * exact float->double widening happens on the device, so results equal the f64 call on the widened array. */
int kb_pipeline_register_frame_f32(kb_pipeline *p, const float *xyz, size_t n, const double *timestamps,
                                   size_t n_timestamps);
/* same, frame (and stamps) already resident in HBM on the pipeline's device */
int kb_pipeline_register_frame_dev(kb_pipeline *p, const double *d_xyz, size_t n, const double *d_timestamps,
                                   size_t n_timestamps);
/* A whole sequence: the loop `for frame in dataset: RegisterFrame(frame, stamps)` of
 * python/kiss_icp/pipeline.py:106-112 as one call. Poses, motion model, threshold and map never leave the device
 * between frames, so the frames are QUEUED: frame k+1 is copied to the device while frame k is registered and
 * results are read back behind the queue (depth 3). Results are identical to `count` blocking calls.
 * xyz[k] is frame k in the given layout, timestamps / n_timestamps may be NULL (no stamps at all);
 * poses_out, if not NULL, receives the `count` row-major poses. An invalid frame (e.g. too few stamps) stops the
 * sequence with the error the blocking call would give, after the frames before it were registered. Host buffers
 * may be pageable (staged through pinned memory) or pinned (copied directly). */
enum kb_frame_layout { KB_FRAMES_HOST_F64 = 0, KB_FRAMES_HOST_F32 = 1, KB_FRAMES_DEVICE_F64 = 2 };
int kb_pipeline_register_frames(kb_pipeline *p, const void *const *xyz, const size_t *n,
                                const double *const *timestamps, const size_t *n_timestamps, size_t count,
                                int layout, double *poses_out);
/* how many frames were vetoed by the kernel (voxel table sized too optimistically), grown for and replayed */
int kb_pipeline_grow_retries(const kb_pipeline *p, unsigned long long *out);
/* sizes / contents of the (preprocessed_frame, source) tuple of the last RegisterFrame */
int kb_pipeline_last_cloud_sizes(const kb_pipeline *p, size_t *n_preprocessed, size_t *n_source);
int kb_pipeline_last_clouds(const kb_pipeline *p, double *preprocessed_xyz, size_t cap_preprocessed,
                            double *source_xyz, size_t cap_source);
/* Voxelize(frame) KissICP.cpp:70-75 -> (source, frame_downsample) */
int kb_pipeline_voxelize(kb_pipeline *p, const double *xyz, size_t n, double *source_xyz, size_t cap_source,
                         size_t *n_source, double *downsample_xyz, size_t cap_downsample, size_t *n_downsample);
/* pose() / delta() KissICP.hpp:81-85 (getters and setters) */
int kb_pipeline_pose(const kb_pipeline *p, double out[16]);
int kb_pipeline_delta(const kb_pipeline *p, double out[16]);
int kb_pipeline_set_pose(kb_pipeline *p, const double pose[16]);
int kb_pipeline_set_delta(kb_pipeline *p, const double delta[16]);
/* VoxelMap() KissICP.hpp:78-79 — borrowed handle, owned by the pipeline */
kb_map *kb_pipeline_voxel_map(kb_pipeline *p);
/* LocalMap() KissICP.hpp:76 == kb_map_pointcloud(kb_pipeline_voxel_map(p), ...) */
/* diagnostics: adaptive threshold sigma used by / ICP iterations of the last RegisterFrame */
int kb_pipeline_last_sigma(const kb_pipeline *p, double *out);
int kb_pipeline_last_iterations(const kb_pipeline *p, int *out);
/* adaptive_threshold_.ComputeThreshold() of the pipeline's own estimator (Threshold.hpp:38):
 * the sigma the NEXT RegisterFrame will use */
int kb_pipeline_threshold(const kb_pipeline *p, double *out_sigma);
/* (profiling and diagnostic entry points: kiss_icp_b200_debug.h) */
/* kernels launched by this pipeline so far (bench "gpu_launches") */
int kb_pipeline_launch_count(const kb_pipeline *p, unsigned long long *out);

#ifdef __cplusplus
}
#endif
#endif /* KISS_ICP_B200_H */
