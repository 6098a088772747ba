/*
 * nirrt_pointnet2.h -- C ABI of the PointNet++ guidance-state inference path in libnirrt_b200.so
 * (sm_100a).  Same conventions as nirrt_b200.h: plain pointers and sizes, 0 / negative return
 * codes, nirrt_last_error() for the message, `stream` is a cudaStream_t passed as void*.
 *
 * Replaces, for batches of B clouds of n_points points:
 *   PNGWrapper.classify_path_points     wrapper{,_3d}/pointnet_pointnet2/pointnet2_wrapper.py:28-64 / :28-59
 *   pc_normalize                        pointnet_pointnet2/models/pointnet2_utils.py:13-18
 *   get_model.forward                   pointnet_pointnet2/models/pointnet2.py:24-42
 *   farthest_point_sample               pointnet2_utils.py:65-86
 *   query_ball_point / square_distance  pointnet2_utils.py:89-109 / :21-42
 *   PointNetSetAbstractionMsg.forward   pointnet2_utils.py:226-264
 *   PointNetFeaturePropagation.forward  pointnet2_utils.py:278-317
 */
#ifndef NIRRT_POINTNET2_H
#define NIRRT_POINTNET2_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nirrt_pn2 nirrt_pn2;

/* One 1x1 convolution with its eval-mode BatchNorm (bn_* NULL for conv2, which has none).
 * weight [c_out][c_in] f32 host (PyTorch Conv1d/Conv2d weight flattened), bias [c_out]. */
typedef struct nirrt_pn2_layer {
    const float *weight, *bias;
    const float *bn_weight, *bn_bias, *bn_mean, *bn_var;
    int c_in, c_out;
} nirrt_pn2_layer;

#define NIRRT_PN2_NUM_LAYERS 35
/* Layer order of nirrt_pn2_create (state_dict prefixes, pointnet2.py:11-22):
 *   0..23  sa{1..4}.conv_blocks.{0,1}.{0,1,2}  (+ bn_blocks)      for l: for scale: for j
 *   24..32 fp4.mlp_convs.{0,1}, fp3.{0,1}, fp2.{0,1}, fp1.{0,1,2}  (+ mlp_bns)
 *   33     conv1 + bn1
 *   34     conv2 (no BatchNorm; c_out = num_classes = 2)
 * BatchNorm is folded on the host (eps 1e-5); weights are stored in HBM as fp16 [c_out][c_in] rows
 * (K-major), the layout both TMA and tcgen05.mma consume. */
int nirrt_pn2_create(const nirrt_pn2_layer *layers, int n_layers, int n_points, int max_batch, int device,
                     nirrt_pn2 **out);
int nirrt_pn2_destroy(nirrt_pn2 *h);

/* classify_path_points for `batch` clouds, host buffers (copies inside, synchronous):
 *   pc          [batch][n_points][dim] f32, dim 2 or 3 (2D clouds are z-padded with 0)
 *   start_mask  [batch][n_points] f32 0/1      goal_mask likewise
 *   fps_start   [batch][4] int32: first index of each of the four farthest_point_sample calls
 *               (the reference draws them with torch.randint on the CPU generator,
 *               pointnet2_utils.py:77; the host shim draws them the same way)
 *   path_pred   [batch][n_points] int64   argmax of the log-probabilities
 *   path_score  [batch][n_points] f32     softmax(...)[:, 1]
 *   logp        [batch][n_points][2] f32  log_softmax output (may be NULL) */
int nirrt_pn2_classify_sync(nirrt_pn2 *h, int batch, int dim, const float *pc, const float *start_mask,
                            const float *goal_mask, const int32_t *fps_start, int64_t *path_pred,
                            float *path_score, float *logp, void *stream);
/* Cloud size of the following classify calls: any 16 <= n_points <= the value given to nirrt_pn2_create (buffers are
 * sized for that).  The reference's samplers only down-sample `if len(point_cloud) > n_points`
 * (datasets_3d/point_cloud_mask_utils_3d.py:104-112), so clouds come in every size up to pc_n_points. */
int nirrt_pn2_set_n_points(nirrt_pn2 *h, int n_points);

/* Same with DEVICE pointers, asynchronous on `stream` (no host synchronisation). */
int nirrt_pn2_classify_device(nirrt_pn2 *h, int batch, int dim, const float *pc, const float *start_mask,
                              const float *goal_mask, const int32_t *fps_start, int64_t *path_pred,
                              float *path_score, float *logp, void *stream);

/* Parity-test taps of the LAST forward: name is one of
 *   "xyz0" f32 [B][N][3] (normalised cloud), "fps0".."fps3" int32 [B][npoint_l],
 *   "group0".."group7" int32 [B][S][K] (sa1 r0, sa1 r1, sa2 r0, ...),
 *   "feat1".."feat4" fp16 [B][npoint_l][C_l]  (SA outputs), "up3","up2","up1","up0" fp16 (FP outputs).
 * Copies min(bytes, size) bytes to `out` (host) and returns the buffer's size in bytes. */
int64_t nirrt_pn2_read_buffer_sync(nirrt_pn2 *h, const char *name, void *out, int64_t bytes, void *stream);

/* Device milliseconds of the last nirrt_pn2_classify_* call per stage (CUDA events on the stream):
 * ms[0] prep, [1] fps, [2] ball query, [3] gather, [4] SA MLP (tcgen05), [5] interpolation,
 * [6] FP MLP + conv1 (tcgen05), [7] head.  Only filled while profiling is enabled (costs syncs). */
int nirrt_pn2_set_profiling(nirrt_pn2 *h, int enabled);
int nirrt_pn2_last_stage_ms(nirrt_pn2 *h, float *ms8);
/* kernel launches issued so far */
int64_t nirrt_pn2_launch_count(nirrt_pn2 *h);

/* Neural Connect graph analysis (wrapper/utils/bfs_connect_heuristic.py:5-29,32-78; used by
 * PNGWrapper.generate_connected_path_points, pointnet2_wrapper_connect_bfs.py:76-240): over the r-disc
 * graph on [src, dst, pc[path_mask]] (float32 norms, strict <) returns has_path (dst reachable from
 * src), visited_mask [n] (path points in src's component) and boundary_mask [n] (visited path points
 * with a non-path point closer than radius).  Host buffers, synchronous; n <= 4096, dim 2 or 3. */
int nirrt_connect_analyse_sync(const float *pc, int n, int dim, const uint8_t *path_mask, const float *src,
                               const float *dst, float radius, int *has_path, uint8_t *visited_mask,
                               uint8_t *boundary_mask, void *stream);

/* The same analysis for `batch` (cloud, mask, src, dst) tuples in one launch, one CTA each: pc [batch][n_max][dim] with
 * n_pts[b] valid points, path_mask / visited_mask / boundary_mask [batch][n_max], src / dst [batch][3] (z ignored in
 * 2D), has_path [batch].  What a lock-step batch of NIRRT*-PNG(C) planners issues per Neural Connect trial (both
 * search directions of every waiting problem). */
int nirrt_connect_analyse_batch_sync(const float *pc, const int *n_pts, int n_max, int dim, int batch, const uint8_t *path_mask,
                                     const float *src, const float *dst, float radius, int *has_path,
                                     uint8_t *visited_mask, uint8_t *boundary_mask, void *stream);

/* The masks of the FIRST trial: start_mask[b][i] = |pc[b][i] - src[b]| < radius, goal_mask[b][i] = |pc[b][i] - src[batch + b]|
 * < radius in float32 (get_point_cloud_mask_around_points on pc.astype(float32), as pointnet2_wrapper_connect_bfs.py sees it).
 * Device pointers, asynchronous on `stream`. */
int nirrt_connect_masks_device(const float *pc, int n_points, int dim, int batch, const float *src, float radius,
                               float *start_mask, float *goal_mask, void *stream);
/* One Neural Connect trial for `batch` problems whose clouds all hold n_points points, everything in HBM
 * (generate_connected_path_points, pointnet2_wrapper_connect_bfs.py:76-240, what follows the network call; heuristic
 * wrapper/utils/bfs_connect_heuristic.py:142-181).  For every problem with active[b] != 0:
 *   path_mask[b] |= (pred[b] != 0);  the searches src[b] -> dst[b] and src[batch + b] -> dst[batch + b] (callers pass
 *   starts then goals in src and goals then starts in dst: [2 * batch][3], z ignored in 2D) over the r-disc graph;
 *   if neither reaches its target: the boundary point with the lowest rank sum (total cost ascending + cost from the
 *   source descending, first in index order among equals) replaces start_mask[b] (first search) / goal_mask[b] (second)
 *   by the points closer than radius to it; a search without boundary points leaves its mask as it is.
 * pc [batch][n_points][dim] f32, path_mask [batch][n_points] u8 (in/out), pred [batch][n_points] int64 (the network's
 * path_pred), start_mask / goal_mask [batch][n_points] f32 (in/out): device pointers.  Host outputs after the stream
 * has been synchronised: has_path [2 * batch]; ties [2 * batch] != 0 where two boundary points have EQUAL keys -- numpy's
 * argsort order among equal keys is an implementation detail, so that search's mask is left untouched and its boundary
 * mask is copied to tie_boundary [2 * batch][n_points] for the caller to decide with numpy itself. */
int nirrt_connect_trial_device(const float *pc, int n_points, int dim, int batch, const uint8_t *active, uint8_t *path_mask,
                               const int64_t *pred, const float *src, const float *dst, float radius, float *start_mask,
                               float *goal_mask, int32_t *has_path, int32_t *ties, uint8_t *tie_boundary, void *stream);

/* Stand-alone tensor-core GEMM (the kernel the network uses), host buffers, synchronous:
 *   A [m][k] fp16, W [n][k] fp16, bias [n] f32; k, n multiples of 16.
 *   mode 0: out [m][n]        = fp16(relu(A W^T + bias))
 *   mode 1: out [m/group][n]  = fp16(relu(max over each `group` (16|32) consecutive rows + bias)) */
int nirrt_gemm_f16_sync(const uint16_t *A, const uint16_t *W, const float *bias, int m, int n, int k,
                        int mode, int group, uint16_t *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NIRRT_POINTNET2_H */
