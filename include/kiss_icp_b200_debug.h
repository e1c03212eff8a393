/* kiss_icp_b200_debug.h — profiling and diagnostic entry points of libkiss_icp_b200 (no reference counterpart).
 * Kept apart from the drop-in surface of kiss_icp_b200.h: a caller that replaces the reference never needs these; the
 * benchmark, the developer tools under tools/ and the numerics tests do. */
#ifndef KISS_ICP_B200_DEBUG_H
#define KISS_ICP_B200_DEBUG_H

#include "kiss_icp_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* profiling aids (no reference counterpart): raw %globaltimer stamps [ns, relative to stamp 0]
 * of the second ICP iteration of the last frame (CTA 0); cost of one grid barrier */
int kb_pipeline_debug_stamps(const kb_pipeline *p, double *ns, int n);
int kb_debug_barrier_ns(int iters, double *ns_per_barrier);
/* host evaluation of the ICP loop's structured (Schur complement) solve vs the pivoted LDL^T on the same 16
 * accumulators {sum w, sum w*s (3), lower triangle of sum w*hat(s)^T*hat(s) (6), JTr (6)} */
int kb_debug_icp_schur(const double acc[16], double x_schur[6], double x_ldlt[6], int *used_schur);
/* host evaluation of the device's two 6x6 LDLT code paths (loop form / register-resident form) */
int kb_debug_ldlt6(const double A[36], const double b[6], double x_loop[6], double x_unrolled[6]);
/* host evaluation of the exact vs latency-optimised (reciprocal / sincos) ICP solve step:
 * x = LDLT solve, T = exp(x) * exp(b) */
int kb_debug_icp_solve(const double A[36], const double b[6], double x_exact[6], double x_fast[6], double T_exact[16],
                       double T_fast[16]);
/* work done by the ICP loop of the last RegisterFrame: GetClosestNeighbor calls (iterations x
 * source points) and map points examined — the inputs of the algorithmic-bytes formula */
int kb_pipeline_last_icp_work(const kb_pipeline *p, double *queries, double *candidates);
/* split of the two downsample phases of the last frame [us]: clear, dedupe, count, prefix+rank, replay+emit (x2) */
int kb_pipeline_last_ds_profile(const kb_pipeline *p, double us[10]);
/* split of the map-update phase of the last frame [us]: claim + pending lists, ordered insertion, eviction */
int kb_pipeline_last_map_profile(const kb_pipeline *p, double us[3]);
/* shared-memory NN cache of the ICP loop, last frame: {hits, fills, overflows} over all iterations */
int kb_pipeline_last_cache_stats(const kb_pipeline *p, double out[3]);
/* device-side duration [us] of the phases of the last RegisterFrame, from %globaltimer stamps
 * inside the kernel: preprocess, downsample(0.5v), downsample(1.5v), ICP, map update, epilogue */
int kb_pipeline_last_profile(const kb_pipeline *p, double *us, int n);
/* in-kernel phase timestamps (kb_pipeline_last_profile & co.) are OFF by default: every %globaltimer read costs
 * ~1 us on the kernel's critical path. Enable them for profiling runs only. */
int kb_pipeline_set_profiling(kb_pipeline *p, int enabled);
/* per-frame statistics, recorded on the host after every RegisterFrame when enabled (so a
 * benchmark can read them AFTER its timed region): kb_pipeline_set_history(p, capacity) starts
 * a fresh log of up to `capacity` frames */
typedef struct kb_frame_stats {
    double pose[16];
    double phase_us[6];   /* preprocess, downsample 0.5v, downsample 1.5v, ICP, map update, epilogue */
    double icp_queries;   /* GetClosestNeighbor calls */
    double icp_candidates; /* map points examined */
    int iterations;
    int n_points_in, n_preprocessed, n_downsampled, n_source;
    int map_points, map_voxels;
    int team;             /* CTAs of the ICP team that ran the iterations (0: whole-grid loop) */
} kb_frame_stats;
/* 1 if a launch on some device of this process never ended (the call that waited for it returned KB_ERR_CUDA after
 * KB_SYNC_TIMEOUT_S seconds, default 30). Handles of that device then leak their device memory on destruction instead of
 * blocking in cudaFree; the process should exit without CUDA's teardown (the Python package does, through os._exit). */
int kb_any_device_stuck(void);
int kb_pipeline_set_history(kb_pipeline *p, size_t capacity);
int kb_pipeline_get_history(const kb_pipeline *p, kb_frame_stats *out, size_t capacity, size_t *n_out);
/* the logged frames' in-kernel %globaltimer stamps (profiling on), ns modulo 2^40, 20 per frame: [0] kernel start, [1..3] front
 * end phases, [7] candidate lists done, [10] ICP iterations done, [11] next frame's front end done (prefetch), [4] ICP result
 * everywhere, [5] map updated, [6] kernel end, [8,9] map phases, [12..19] downsample phases */
int kb_pipeline_history_stamps(const kb_pipeline *p, double *out, size_t capacity, size_t *n_out);

#ifdef __cplusplus
}
#endif
#endif /* KISS_ICP_B200_DEBUG_H */
