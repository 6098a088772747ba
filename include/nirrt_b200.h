/*
 * nirrt_b200.h -- C ABI of libnirrt_b200.so (sm_100a), the drop-in boundary of the NIRRT* hot path.
 *
 * The reference (tedhuang96/nirrt_star) has no native boundary: its hot path is Python/numpy
 * (SURVEY.md section 8b).  This header is the boundary a maintainer binds with ctypes (see
 * INTEGRATION.md); each entry point names the reference interface it replaces.  Plain pointers and
 * sizes only, no torch types.  All `host` pointers are caller-owned host memory (pinned memory makes
 * the copies asynchronous); `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *
 * Conventions: every function returns 0 on success or a negative code; nirrt_last_error() returns a
 * description for the calling thread.  A batch handle must not be used from two threads at once.
 * Entry points whose name ends in _sync wait for the stream; the others only enqueue work.
 */
#ifndef NIRRT_B200_H
#define NIRRT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NIRRT_OK 0
#define NIRRT_ERR_INVALID -1      /* bad argument */
#define NIRRT_ERR_CUDA -2         /* CUDA runtime error (message in nirrt_last_error) */
#define NIRRT_ERR_CAPACITY -3     /* a device-side buffer overflowed (near candidates, solutions, vertices) */
#define NIRRT_ERR_NO_DEVICE -4    /* no CUDA device / wrong architecture */

#define NIRRT_MAX_OBSTACLES 32    /* per obstacle type and problem */

/* planner families (path_planning_classes_3d/{rrt_star,irrt_star,nirrt_star_png}_3d.py) */
#define NIRRT_VARIANT_RRT_STAR 0
#define NIRRT_VARIANT_IRRT_STAR 1
#define NIRRT_VARIANT_NIRRT_STAR 2
#define NIRRT_VARIANT_NRRT_STAR 3   /* RRT* driver + fixed guidance cloud (nrrt_star_png_3d.py:52-56) */
/* loop drivers: planning() body (rrt_star_3d.py:36-55) / planning_random (rrt_star_3d.py:200-270,
 * irrt_star_3d.py:245-331) */
#define NIRRT_MODE_PLANNING 0
#define NIRRT_MODE_PLANNING_RANDOM 1

const char *nirrt_last_error(void);
int nirrt_version(void);
/* number of visible CUDA devices with compute capability 10.x; <= 0 means the library cannot run */
int nirrt_device_count(void);

/* Device blocks freed by batches / engines / temporaries are kept in a process-wide cache (exact size match per device,
 * at most NIRRT_CACHE_GB gigabytes, default 32; 0 disables) and handed out again zero-filled, so that repeated create /
 * destroy cycles of equal shape cost no driver allocation calls.  This returns the idle blocks to the driver. */
int nirrt_release_cached_memory(void);

typedef struct nirrt_batch nirrt_batch;

typedef struct nirrt_batch_desc {
    int dim;            /* 3, or 2 (problems then come from nirrt_batch_set_problems_2d; every [..][3] array
                           below is [..][2], edges are [m][2][2]) */
    int n_envs;         /* E: independent planning problems advanced in lock step */
    int capacity;       /* vertices per problem = 1 + iter_max (rrt_base_3d.py:25) */
    int record_capacity;/* per-problem path_len_list rows (>= iter_max + iter_after_initial + 2) */
    int near_capacity;  /* per-problem candidate buffer (entries) of the PRE-FILTER superset of Near
                           (speculative ball + mirror margin); 0 = default 1024.  An iteration handles at most 1024
                           candidates (shared-memory staging); larger values only enlarge the result buffer of
                           nirrt_within_sync.  Overflow is a hard error (NIRRT_ERR_CAPACITY) reported by
                           nirrt_batch_status_sync, never a silent truncation. */
    int device;         /* CUDA device ordinal */
} nirrt_batch_desc;

/* Allocates the flat HBM state for E problems: SoA vertex coordinates (scan layout), 32-byte
 * vertex records {x,y,z,parent} (cost-walk layout), obstacle tables, MT19937 streams. */
int nirrt_batch_create(const nirrt_batch_desc *desc, nirrt_batch **out);
int nirrt_batch_destroy(nirrt_batch *b);

/* Problem upload == the arguments of RRTStar3D.__init__ / get_path_planner
 * (rrt_star_3d.py:10-30,272-285) + Utils.__init__ (rrt_utils_3d.py:6-19), for all E problems.
 *   start, goal      [E][3]
 *   step_len, search_radius, clearance   [E]
 *   range            [E][6]  x0 x1 y0 y1 z0 z1   (Env.x_range..., rrt_env_3d.py:6-9)
 *   n_balls,n_boxes  [E]     each <= NIRRT_MAX_OBSTACLES
 *   balls            [E][NIRRT_MAX_OBSTACLES][4]  x y z r
 *   ball_r2          [E][NIRRT_MAX_OBSTACLES]     (r+clearance)**2 evaluated by numpy's scalar power
 *                    (collision_check_utils_3d.py:21,31-37; libm pow is not x*x, so the host passes it)
 *   boxes            [E][NIRRT_MAX_OBSTACLES][6]  x y z w h d
 *   near_table       [capacity+2]  t[n] = (math.log(n)/n)**(1/3.)  (rrt_star_3d.py:134; libm on host)
 *   rot_c            [E][9] IRRTStar3D.RotationToWorldFrame (irrt_star_3d.py:159-173, numpy SVD on host);
 *                    may be NULL for RRT*.
 * Resets every tree to the single start vertex. */
int nirrt_batch_set_problems(nirrt_batch *b, const double *start, const double *goal,
                             const double *step_len, const double *search_radius, const double *clearance,
                             const double *range, const int *n_balls, const double *balls,
                             const double *ball_r2, const int *n_boxes, const double *boxes,
                             const double *near_table, const double *rot_c, void *stream);

/* 2D problems == the arguments of RRTStar2D.__init__ / get_path_planner (rrt_star_2d.py:10-30,270-283) +
 * Utils.__init__ (rrt_utils_2d.py:5-17):
 *   start, goal [E][2]; step_len, search_radius, clearance [E]; range [E][4] x0 x1 y0 y1 (rrt_env.py:7-8)
 *   circles [E][NIRRT_MAX_OBSTACLES][3] x y r; rects [E][NIRRT_MAX_OBSTACLES][4] x y w h
 *   near_table [capacity+2]  t[n] = math.sqrt(math.log(n)/n)  (rrt_star_2d.py:133)
 *   rot_c [E][9] IRRTStar2D.RotationToWorldFrame (irrt_star_2d.py:153-161), may be NULL for RRT*. */
int nirrt_batch_set_problems_2d(nirrt_batch *b, const double *start, const double *goal,
                                const double *step_len, const double *search_radius, const double *clearance,
                                const double *range, const int *n_circles, const double *circles,
                                const int *n_rects, const double *rects, const double *near_table,
                                const double *rot_c, void *stream);
/* random.seed(s) state of each problem (2D informed sampling draws its unit disc from CPython's
 * `random`, irrt_star_2d.py:146-151): key [E][624], pos [E] (random.getstate()[1][:624], [624]) */
int nirrt_batch_set_py_rng(nirrt_batch *b, const uint32_t *key, const int *pos, void *stream);
int nirrt_batch_get_py_rng_sync(nirrt_batch *b, uint32_t *key, int *pos, void *stream);

/* np.random.seed(s) state of each problem: key [E][624], pos [E] (np.random.get_state()[1:3]) */
int nirrt_batch_set_rng(nirrt_batch *b, const uint32_t *key, const int *pos, void *stream);
int nirrt_batch_get_rng_sync(nirrt_batch *b, uint32_t *key, int *pos, void *stream);

/* NIRRT* knobs (nirrt_star_png_3d.py:39-45): pc_sample_rate, pc_update_cost_ratio, applied to all */
int nirrt_batch_set_guidance(nirrt_batch *b, double pc_sample_rate, double pc_update_cost_ratio);
/* path_point_cloud_pred of one problem (nirrt_star_png_3d.py:172): points [n][3] f64 host */
int nirrt_batch_set_cloud(nirrt_batch *b, int env, const double *points, int n, void *stream);

/* 2D worlds: the free-space image of every problem (binary_mask of datasets/planning_problem_utils_2d.py: 1 = free),
 * masks [E][height][width] u8 host; needed by nirrt_batch_sample_clouds_sync on a 2D batch
 * (datasets/point_cloud_mask_utils.py:52-66: a point is free iff the four pixels around it are). */
int nirrt_batch_set_free_masks(nirrt_batch *b, const uint8_t *masks, int height, int width, void *stream);

/* ---- guidance clouds on the device (3D; datasets_3d/point_cloud_mask_utils_3d.py:83-113,132-200 + the mask
 * helper datasets/point_cloud_mask_utils.py:20-31), for a list of `count` problems in one call:
 *   envs[count]    problem indices;  kind[count]  0 = generate_rectangle_point_cloud_3d, 1 = ellipsoid_point_cloud_sampling_3d
 *   params[count][12]  kind 1: M = C @ L (row major 3x3) and the ellipsoid centre, evaluated by the caller with numpy as
 *                      the reference does (c_max ** 2 is libm pow); ignored for kind 0
 *   n_points, n_raw    pc_n_points and pc_n_points * pc_over_sample_scale (nirrt_star_png_3d.py:141-156)
 * 2D batches (datasets/point_cloud_mask_utils.py:35-73,104-174): kind 0 = generate_rectangle_point_cloud, kind 1 =
 * ellipsoid_point_cloud_sampling (params: M = C @ L and the centre padded to 3), pc32 is [count][n_points][2].
 * Each problem's numpy MT19937 stream is consumed on the device exactly as the reference consumes it (dim * n_raw
 * doubles), candidates are filtered at clearance 0 and farthest-point down-sampled (open3d semantics) when more than
 * n_points survive.  Outputs, written to DEVICE buffers the caller owns (they are the PointNet++ engine's inputs):
 * pc32 [count][n_points][3] f32, start / goal masks [count][n_points] f32 (|p - x_start| < neighbor_radius, strict);
 * counts[count] (host) = points in each cloud (rows beyond it are zero; < n_points means a short cloud).  The three
 * device buffers may all be NULL (the batch then uses internal scratch; callers that only want the points).  Waits for
 * the stream.  The f64 clouds stay in the batch's workspace until the next call. */
int nirrt_batch_sample_clouds_sync(nirrt_batch *b, const int *envs, int count, const int *kind, const double *params,
                                   int n_points, int n_raw, double neighbor_radius, float *d_pc32, float *d_start_mask,
                                   float *d_goal_mask, int *counts, void *stream);
/* f64 clouds first .. first+count-1 of the last nirrt_batch_sample_clouds_sync call -> points [count][n_points][3] (host) */
int nirrt_batch_read_sampled_clouds_sync(nirrt_batch *b, int first, int count, double *points, void *stream);
/* path_point_cloud_pred = pc[path_pred.nonzero()[0]] (nirrt_star_png_3d.py:172) for the clouds of the last sample call:
 * d_pred device int64 [count][n_points] (the engine's path_pred); sel = host list of cloud positions to commit (NULL:
 * all).  Resumes the problems that were waiting for a cloud.  Enqueues only. */
int nirrt_batch_commit_clouds(nirrt_batch *b, const int64_t *d_pred, const int *sel, int n_sel, void *stream);

/* Tree snapshot in the reference's own layout: vertices [count][capacity][3] f64 (AoS),
 * parents [count][capacity] int64, n [count]  (RRTBase3D.vertices / vertex_parents / num_vertices,
 * rrt_base_3d.py:25-28) for problems env_begin .. env_begin+count-1. */
int nirrt_batch_load_trees(nirrt_batch *b, int env_begin, int count, const int *n,
                           const double *vertices, const int64_t *parents, void *stream);
int nirrt_batch_read_trees_sync(nirrt_batch *b, int env_begin, int count, int *n,
                                double *vertices, int64_t *parents, void *stream);

/* Starts a driver: variant/mode select the loop (see defines); iter_max = phase-1 cap,
 * iter_after_initial = phase-2 length (planning_random); for NIRRT_MODE_PLANNING only iter_max is
 * used.  Resets per-problem phase machines and record counters (not the trees). */
int nirrt_batch_begin(nirrt_batch *b, int variant, int mode, int iter_max, int iter_after_initial, void *stream);

/* planning_block_gap(path_len_threshold) (rrt_star_2d.py:159-196, irrt_star_3d.py:193-243): call after
 * nirrt_batch_begin(..., NIRRT_MODE_PLANNING_RANDOM, iter_max, 0); phase 1 then ends as soon as the
 * recorded value is < stop_below instead of < inf.  nirrt_batch_begin resets it to +inf. */
int nirrt_batch_set_stop_threshold(nirrt_batch *b, double stop_below);

/* Enqueues `iters` lock-step iterations of the loop body on every problem that is still running
 * (Sample -> Nearest scan -> Steer + collision -> Near scan -> ChooseParent/Rewire/goal work).
 * Problems that finished their driver idle.  No host synchronisation.  Blocks of 16 iterations replay from
 * the CUDA graph nirrt_batch_begin built for this variant/mode; a run never builds a graph it can find. */
int nirrt_batch_run(nirrt_batch *b, int iters, void *stream);

/* Benchmark pre-growth: while limit > 0, problems whose tree already holds `limit` vertices idle
 * (they neither sample nor consume iterations), so a batch can be grown to a common size. */
int nirrt_batch_set_vertex_limit(nirrt_batch *b, int limit);
/* nirrt_batch_run with a CUDA-event bracket around every kernel launch; ms5 receives the summed
 * device milliseconds of {top, nearest scan, steer, near scan, expand} over `iters` iterations. */
int nirrt_batch_run_profiled_sync(nirrt_batch *b, int iters, float *ms5, void *stream);

/* Waits for the stream; returns the number of problems whose driver has not finished in *running
 * and the number waiting for a guidance cloud in *need_cloud.  Returns NIRRT_ERR_CAPACITY if any
 * problem overflowed a device buffer. */
int nirrt_batch_status_sync(nirrt_batch *b, int *running, int *need_cloud, void *stream);
/* per problem: state[E] (0 done, 1 phase 1, 2 phase 2, 3 waiting for cloud), n_records[E], n_vertices[E] */
int nirrt_batch_env_state_sync(nirrt_batch *b, int *state, int *n_records, int *n_vertices, void *stream);

/* c_best as refreshed at the top of the last iteration (irrt_star_3d.py:253-254) and c_min = |goal - start|
 * of every problem: the arguments of NIRRTStarPNG3D.update_point_cloud(cmax, cmin) (nirrt_star_png_3d.py:113-115) */
int nirrt_batch_read_cbest_sync(nirrt_batch *b, double *c_best, double *c_min, void *stream);

/* Raw per-iteration records [count][record_capacity] f64 + lengths [count].  RRT* family: path
 * length after each iteration; IRRT* family: c_best at the top of each iteration plus the final
 * refresh -- path_len_list is records[1:]. */
int nirrt_batch_read_records_sync(nirrt_batch *b, int env_begin, int count, double *records, int *n_records, void *stream);

/* path_solutions of one problem (irrt_star_3d.py:29): returns count, fills out[<=cap] */
int nirrt_batch_read_solutions_sync(nirrt_batch *b, int env, int64_t *out, int cap, void *stream);
/* RRTStar3D.search_goal_parent (rrt_star_3d.py:101-117) / find_best_path_solution
 * (irrt_star_3d.py:80-93) for every problem: goal_parent[E] (-1 = None), cost[E] */
int nirrt_batch_goal_parent_sync(nirrt_batch *b, int64_t *goal_parent, double *cost, void *stream);

/* Per-iteration trace of the LAST executed iteration of every problem (parity tests):
 * nearest[E], new_index[E] (-1: steer edge collided), near_count[E], near [E][near_stride] int32 */
int nirrt_batch_read_trace_sync(nirrt_batch *b, int *nearest, int *new_index, int *near_count,
                                int *near, int near_stride, double *x_rand, void *stream);

/* ---- stand-alone batched predicates over a problem's obstacle table --------------------------
 * Utils.is_collision(start,end) for m edges: edges [m][2][3] f64 host -> out [m] u8
 * (rrt_utils_3d.py:22-36 -> collision_check_utils_3d.py:151-216) */
int nirrt_collide_edges_sync(nirrt_batch *b, int env, const double *edges, int64_t m, uint8_t *out, void *stream);
/* kind 0: Utils.is_inside_obs (rrt_utils_3d.py:39-51); kind 1: Utils.is_valid (:68-86);
 * points [m][3] -> out [m] u8 */
int nirrt_points_check_sync(nirrt_batch *b, int env, int kind, const double *points, int64_t m, uint8_t *out, void *stream);
/* RRTBase3D.nearest_neighbor (rrt_base_3d.py:100-113) for m queries against problem env's tree */
int nirrt_nearest_sync(nirrt_batch *b, int env, const double *queries, int64_t m, int64_t *out, void *stream);
/* np.where(np.linalg.norm(q - vertices, axis=-1) <= r)[0] (rrt_star_3d.py:135-137), ascending;
 * returns the count (may exceed cap; only cap entries are written) */
int64_t nirrt_within_sync(nirrt_batch *b, int env, const double *q, double r, int64_t *out, int64_t cap, void *stream);
/* RRTBase3D.cost (rrt_base_3d.py:60-67) for m vertex indices */
int nirrt_costs_sync(nirrt_batch *b, int env, const int64_t *idx, int64_t m, double *out, void *stream);

/* Farthest-point down-sampling of an f64 point set [n][3] (n <= 16384) to npoint indices, starting at
 * `start`: open3d's PointCloud.farthest_point_down_sample as the guidance-cloud generators call it
 * (datasets_3d/point_cloud_mask_utils_3d.py:49-52,196-199; datasets/point_cloud_mask_utils.py:69-72,170-173) */
int nirrt_fps_f64_sync(const double *points, int64_t n, int npoint, int start, int64_t *out_idx, void *stream);

/* math.sin / math.cos as the reference's runtime evaluates them (glibc 2.39 x86-64 FMA kernels, restated operation
 * by operation in csrc/glibc_trig.cuh -- NOT the correctly rounded values): what the device uses for the informed
 * sampler (irrt_star_3d.py:154-156) and the 2D steer (rrt_star_2d.py:77).  x, out_sin, out_cos: host [n]. */
int nirrt_sincos_sync(const double *x, int64_t n, double *out_sin, double *out_cos, void *stream);
/* math.atan2(y, x) likewise (2D steer, rrt_star_2d.py:74 / rrt_base_2d.py:120); y, x, out: host [n] */
int nirrt_atan2_sync(const double *y, const double *x, int64_t n, double *out, void *stream);

/* Device-resident benchmark hooks: bytes scanned per Nearest+Near pass and launch counters. */
int nirrt_batch_counters(nirrt_batch *b, int64_t *kernel_launches, int64_t *scan_bytes_per_vertex);
/* Work counters summed over all problems since nirrt_batch_set_problems: out8 = {expansions, sum |Near|, sum pre-filter
 * candidates, goal-tracking traversal rounds, full goal evaluations, refreshed goal candidates, re-parented subtree
 * roots, sum of the candidate-list length at full evaluations}.  Waits for the stream. */
int nirrt_batch_work_stats_sync(nirrt_batch *b, int64_t *out8, void *stream);
/* CUDA-graph bookkeeping of nirrt_batch_run: executables built (one per variant/mode, by nirrt_batch_begin),
 * graph replays launched, and capture failures (each one downgrades the batch to plain kernel launches --
 * a performance regression that is otherwise invisible). */
int nirrt_batch_graph_stats(nirrt_batch *b, int64_t *builds, int64_t *replays, int64_t *fallbacks);
/* Times `reps` back-to-back launches of ONE scan kernel (0 = Nearest, 1 = Near) on the batch's
 * current trees with CUDA events on `stream`; returns average milliseconds per launch in *ms and
 * the vertex-coordinate bytes one launch reads in *bytes. */
int nirrt_batch_time_scan_sync(nirrt_batch *b, int which, int reps, float *ms, int64_t *bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NIRRT_B200_H */
