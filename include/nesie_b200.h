/*
 * nesie_b200.h -- C ABI of libnesie_b200.so, the B200 (sm_100a) implementation of the
 * Nesie / VoteNet data-parallel hot path (PointNet++ SA/FP operators, pseudo-label NMS,
 * side-uncertainty loss, teacher EMA).
 *
 * Every entry point replaces one launcher of the reference (OpenSpaceAI/Nesie, a fork of
 * mmdetection3d 0.15.0); the reference interface it stands in for is cited per function
 * (paths relative to the reference's mmdet3d/).  The conventions are the reference's:
 *   - plain device pointers + sizes, contiguous row-major tensors, int32 indices, fp32 data;
 *   - the CALLER allocates every output (and zero-fills where stated);
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*) and the call returns
 *     without synchronising;
 * with one deliberate difference: a failed launch returns a non-zero status (the CUDA error
 * code, or a NESIE_ERR_* value) and records a message for nesie_last_error() instead of
 * printing and calling exit(-1) like the reference launchers do
 * (e.g. ops/ball_query/src/ball_query_cuda.cu:73-77).
 *
 * There is no CPU path behind this ABI: without a CUDA device every compute entry point fails.
 */
#ifndef NESIE_B200_H_
#define NESIE_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define NESIE_OK 0
#define NESIE_ERR_INVALID_ARG 10001 /* bad sizes / null pointers / unsupported shape */
#define NESIE_ERR_UNSUPPORTED 10002 /* shape outside what the kernel family covers */

/* ABI version (bumped on any signature change) and last error text of the calling thread. */
int nesie_abi_version(void);
const char *nesie_last_error(void);

/* ---------------------------------------------------------------------------------------
 * Furthest point sampling.
 * Replaces furthest_point_sampling_kernel_launcher(b,n,m,dataset,temp,idxs,stream)
 *   ops/furthest_point_sample/src/furthest_point_sample_cuda.cu:143-209 (kernel :25-141),
 *   bound by furthest_point_sampling_wrapper, src/furthest_point_sample.cpp:32-43,59-65.
 * xyz (b,n,3) f32; idx (b,m) i32 out.  temp (b,n) f32 is OPTIONAL: the reference needs it as
 * scratch pre-filled with 1e10 (furthest_point_sample.py:30); here the running min-distances
 * live in registers.  temp == NULL means "start from 1e10"; a non-NULL temp is read as the
 * initial min-distances and receives the final ones, exactly like the reference's buffer.
 * Bit-exact contract: d = fma(dz,dz,fma(dx,dx,dy*dy)), start index 0, ties resolved the way
 * the reference's per-thread strided scan + shared-memory tree do.
 * nesie_fps_needs_temp() returns 1 when (b,n,m) falls back to the global-memory kernel that
 * requires a caller-provided temp.
 */
int nesie_fps(int b, int n, int m, const float *xyz, float *temp, int *idx, void *stream);
int nesie_fps_needs_temp(int b, int n, int m);

/* Replaces furthest_point_sampling_with_dist_kernel_launcher
 *   furthest_point_sample_cuda.cu:333-399 (kernel :213-331), wrapper .cpp:45-57,
 * dist (b,n,n) f32; temp (b,n) f32 REQUIRED, pre-filled with 1e10 by the caller
 * (furthest_point_sample.py:66); idx (b,m) i32 out. */
int nesie_fps_with_dist(int b, int n, int m, const float *dist, float *temp, int *idx,
                        void *stream);

/* ---------------------------------------------------------------------------------------
 * Ball query.  Replaces ball_query_kernel_launcher(b,n,m,min_radius,max_radius,nsample,
 *   new_xyz,xyz,idx,stream)  ops/ball_query/src/ball_query_cuda.cu:56-78 (kernel :11-54),
 *   bound by ball_query_wrapper, src/ball_query.cpp:30-47.  Note the reference's argument
 *   order: centres (new_xyz (b,m,3)) before points (xyz (b,n,3)).
 * idx (b,m,nsample) i32.  Unlike the reference, the kernel writes EVERY slot of idx (rows
 * without a hit are written as zeros), so the caller's zero-fill (ball_query.py:35) is
 * allowed but not required.
 */
int nesie_ball_query(int b, int n, int m, float min_radius, float max_radius, int nsample,
                     const float *new_xyz, const float *xyz, int *idx, void *stream);

/* Same contract and bit-identical output, computed through a uniform grid (cells >= max_radius,
 * 27-cell neighbourhoods, hits re-ranked by index): the formulation for large scenes, where the
 * brute-force kernel is fp32-issue-bound (SA1: 655 M distance tests at batch 8).  Needs
 * max_radius > 0 and a caller-provided, 16-byte aligned scratch buffer of
 * nesie_ball_query_grid_workspace(b, n, m) bytes (the reference launchers never allocate; neither
 * does this library).
 */
long long nesie_ball_query_grid_workspace(int b, int n, int m);
int nesie_ball_query_grid(int b, int n, int m, float min_radius, float max_radius, int nsample,
                          const float *new_xyz, const float *xyz, int *idx, void *workspace,
                          long long workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------
 * gather_points: out[b,c,j] = points[b,c,idx[b,j]].
 * Replaces gather_points_kernel_launcher / gather_points_grad_kernel_launcher
 *   ops/gather_points/src/gather_points_cuda.cu:28-49,72-95 (wrappers gather_points.cpp:28-59).
 * points (b,c,n), idx (b,npoints) i32, out (b,c,npoints).  grad: grad_out (b,c,npoints) is
 * scatter-added into grad_points (b,c,n), which the caller zero-fills (gather_points.py:44).
 */
int nesie_gather_points(int b, int c, int n, int npoints, const float *points, const int *idx,
                        float *out, void *stream);
int nesie_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out,
                             const int *idx, float *grad_points, void *stream);

/* ---------------------------------------------------------------------------------------
 * grouping_operation: out[b,c,j,k] = points[b,c,idx[b,j,k]].
 * Replaces group_points_kernel_launcher / group_points_grad_kernel_launcher
 *   ops/group_points/src/group_points_cuda.cu:81-105,33-54 (wrappers group_points.cpp:31-62).
 * points (b,c,n), idx (b,npoints,nsample) i32, out (b,c,npoints,nsample).  grad_points (b,c,n)
 * is zero-filled by the caller (group_points.py:219).
 */
int nesie_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                       const int *idx, float *out, void *stream);
int nesie_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                            const int *idx, float *grad_points, void *stream);

/* QueryAndGroup's whole body after the ball query in ONE pass (ops/group_points/
 * group_points.py:98-116): out (b, 3+c, npoints, nsample) with channels 0..2 =
 * (xyz[idx] - center) [/ radius when normalize_xyz] and channels 3.. = features[idx].
 * xyz (b,n,3), center_xyz (b,npoints,3), features (b,c,n) or NULL with c == 0.
 * radius: pass max_radius (> 0) when normalize_xyz=True, or 0 to skip the scaling.  The
 * scaling is x * (1.0f / radius) in fp32, which is what torch's CUDA `grouped_xyz /= radius`
 * with a python-scalar divisor executes (ATen: "a * reciprocal(b)").
 */
int nesie_query_group_concat(int b, int c, int n, int npoints, int nsample, const float *xyz,
                             const float *center_xyz, const float *features, const int *idx,
                             float radius, float *out, void *stream);

/* The same grouped tensor in ROW-major GEMM layout for the training path of the shared MLP:
 * rows ((b*npoints + j)*nsample + k, 3 + c) = [ (xyz[idx] - center) * (1/radius) | table[idx, :] ],
 * with `table_pm` the point-major (b, n, c) copy of the features (one contiguous read and write
 * per row).  _grad scatter-adds grad_rows into grad_table_pm (b, n, c), grad_xyz (b, n, 3) and
 * grad_center (b, npoints, 3); each may be NULL to skip it; all are zero-filled by the caller.
 * `ld` is the row stride in floats (3 + c <= ld <= 3 + c + 32): columns [3 + c, ld) are written as
 * zeros, so a caller can pad rows to a multiple of 4 floats for 16-byte aligned GEMM loads. */
int nesie_group_rows(int b, int c, int n, int npoints, int nsample, const float *xyz,
                     const float *center_xyz, const float *table_pm, const int *idx, float radius,
                     float *rows, int ld, void *stream);
int nesie_group_rows_grad(int b, int c, int n, int npoints, int nsample, const float *grad_rows,
                          const int *idx, float radius, float *grad_table_pm, float *grad_xyz,
                          float *grad_center, int ld, void *stream);

/* ---------------------------------------------------------------------------------------
 * three_nn.  Replaces three_nn_kernel_launcher(b,n,m,unknown,known,dist2,idx,stream)
 *   ops/interpolate/src/three_nn_cuda.cu:67-90 (kernel :11-65), wrapper interpolate.cpp:46-56.
 * unknown (b,n,3), known (b,m,3) -> dist2 (b,n,3) f32 SQUARED distances (python applies sqrt,
 * three_nn.py:38), idx (b,n,3) i32.  Earliest index wins ties; with m < 3 the missing
 * neighbours are reported as (inf, 0) exactly like the reference's (float)1e40 / 0.
 */
int nesie_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                   int *idx, void *stream);

/* nesie_three_nn through a uniform grid over the sources (bit-identical results): each target visits
 * the shells of cells around its own until nothing unvisited can enter its top three.  workspace:
 * nesie_ball_query_grid_workspace(b, m, 0) bytes, 16-byte aligned; build != 0 bins the sources first,
 * 0 reuses the grid a previous (stream-ordered) call built over the same sources. */
int nesie_three_nn_grid(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                        int *idx, void *workspace, long long workspace_bytes, int build, void *stream);

/* three_interpolate: out[b,c,j] = fma(w2,p2,fma(w0,p0,w1*p1)), p_i = points[b,c,idx[b,j,i]].
 * Replaces three_interpolate_kernel_launcher / three_interpolate_grad_kernel_launcher
 *   ops/interpolate/src/three_interpolate_cuda.cu:37-59,86-110, wrappers interpolate.cpp:58-93.
 * points (b,c,m), idx/weight (b,n,3), out (b,c,n).  grad_points (b,c,m) zero-filled by caller.
 */
int nesie_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                            const float *weight, float *out, void *stream);
int nesie_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out,
                                 const int *idx, const float *weight, float *grad_points,
                                 void *stream);

/* ---------------------------------------------------------------------------------------
 * Batched class-aware axis-aligned 3D NMS, one scene per CTA.
 * Replaces the per-scene python loop around aligned_3d_nms(boxes, scores, classes, thresh)
 *   core/post_processing/box3d_nms.py:129-176, called from
 *   models/dense_heads/nesie_head.py:715-724,759-762.
 * boxes (nscenes,max_n,6) f32 [x1,y1,z1,x2,y2,z2]; scores (nscenes,max_n) f32; classes
 * (nscenes,max_n) i32; counts (nscenes) i32 = valid boxes per scene (<= max_n <= 1024), or
 * NULL for "all max_n".  keep (nscenes,max_n) i64 receives the picked indices in pick order
 * (descending score), keep_cnt (nscenes) i32 how many.  IoU arithmetic is fp32 with one
 * rounding per torch op (no FMA); equal scores are ordered by index (stable ascending sort).
 */
int nesie_aligned_3d_nms_batched(int nscenes, int max_n, const float *boxes, const float *scores,
                                 const int *classes, const int *counts, float thresh,
                                 long long *keep, int *keep_cnt, void *stream);

/* Nesie's lenient train-time NMS (float64, volume + 1e-8, re-picks the top half of every
 * suppressed set).  Replaces lhs_3d_faster_samecls(boxes, overlap_threshold, old_type)
 *   models/detectors/votenet_nesie.py:733-779 and the numpy loop that feeds it (:238-258).
 * boxes (nscenes,max_n,8) f64 rows [x1,y1,z1,x2,y2,z2,score,cls]; pick (nscenes, 2*max_n) i32
 * receives the pick list in the reference's order, pick_cnt (nscenes) its length.
 */
int nesie_lhs_nms_batched(int nscenes, int max_n, const double *boxes, const int *counts,
                          double overlap_threshold, int old_type, int *pick, int *pick_cnt,
                          void *stream);

/* ---------------------------------------------------------------------------------------
 * Per-side uncertainty box-regression loss (fused elementwise + reduction).
 * Replaces models/dense_heads/nesie_head.py:332-349 (SurfaceLoss MSE branch,
 * models/losses/surface_loss.py:57-61,90-100, with mmdet's weighted MSELoss):
 *   tgt   = Bbox2Surface(box7);  L = loss_weight * w * (pred - tgt)^2
 *   s     = side_scores[row, side, argmax_cls(sem_scores[row])]
 *   sigma = 0.8 s^2 - 1.8 s + 1
 *   loss  = sum( exp(-sigma) * L + alpha * sigma * w )
 * surface_pred (rows,6); box_targets (rows,7); side_scores (rows,6,ncls); sem_scores
 * (rows,ncls); weight (rows,6).  Outputs: loss_out (1) f32 is ACCUMULATED into (caller
 * zero-fills), sigma_out (rows,6) f32 optional (NULL to skip).
 * _grad: given the upstream scalar grad (device pointer, 1 element) and, optionally, an
 * upstream grad for sigma_out (rows,6; NULL if sigma was not used), writes d/d surface_pred
 * (rows,6) and scatter-writes d/d side_scores (rows,6,ncls; caller zero-fills).
 */
int nesie_side_uncertainty_loss(int rows, int ncls, const float *surface_pred,
                                const float *box_targets, const float *side_scores,
                                const float *sem_scores, const float *weight, float loss_weight,
                                float alpha, float *loss_out, float *sigma_out, void *stream);
int nesie_side_uncertainty_loss_grad(int rows, int ncls, const float *surface_pred,
                                     const float *box_targets, const float *side_scores,
                                     const float *sem_scores, const float *weight,
                                     float loss_weight, float alpha, const float *grad_loss,
                                     const float *grad_sigma, float *grad_surface_pred,
                                     float *grad_side_scores, void *stream);

/* ---------------------------------------------------------------------------------------
 * Teacher EMA over one flat parameter buffer: ema = fma(momentum, param, ema * decay).
 * Replaces the per-tensor loop SimiTeacherHook.hooks_after_train_iter
 *   core/utils/simi_teacher_hook.py:54-64 (buffer.mul_(1-m).add_(param, alpha=m));
 * the caller passes decay = (float)(1 - m) and momentum = (float)m, both rounded from the
 * python doubles the way torch rounds its scalar arguments.
 */
int nesie_ema_update(long long count, float *ema, const float *param, float decay,
                     float momentum, void *stream);

/* ---------------------------------------------------------------------------------------
 * Fused set-abstraction forward for inference / folded BatchNorm (eval mode):
 *   grouped gather -> (xyz - centre) * (1/radius) -> 3 x (1x1 conv + folded BN + ReLU) -> max
 *   over the nsample axis, on tcgen05 tensor cores with TMEM accumulators (bf16 operands, fp32
 *   accumulate).  Replaces the chain QueryAndGroup.forward (ops/group_points/group_points.py:
 *   98-116) -> mlps[i] (ops/pointnet_modules/point_sa_module.py:272-289) -> _pool_features
 *   (:136-158) for `normalize_xyz`/`use_xyz` max-pool SA modules with three MLP layers.
 * features_pm_bf16: (b, n, c8) bf16 point-major table, c8 = c_in rounded up to 8 (0 -> 8), built
 *   by nesie_pack_features_bf16 from the reference's (b, c_in, n) fp32 layout.
 * idx (b, npoints, nsample) i32 from nesie_ball_query; npoints*nsample % 128 == 0.
 * w{1,2,3}_img: weights pre-packed as byte images of the UMMA K-major SWIZZLE_128B shared-memory
 *   layout, [K/64 slabs][C_out rows][128 B], 16-byte chunk c of row r stored at chunk c ^ (r & 7);
 *   layer-1 columns ordered [features (c8) | xyz (3) | zero pad to a multiple of 16].
 *   nsample must be 16, 32 or 64.
 * bias: fp32 [bias1 c1][bias2 c2][bias3 c3] = the folded BN shift; the BN scale must already be
 *   multiplied into the rows of w{1,2,3} (nesie_b200/sa_fused.py::fold_mlp).
 * out (b, c3, npoints) fp32.  nesie_sa_fused_supported() says whether a layer shape is covered.
 */
int nesie_sa_fused_supported(int nsample, int c_in, int c1, int c2, int c3);
int nesie_pack_features_bf16(int b, int c, int n, const float *features, void *table,
                             void *stream);
int nesie_sa_fused_forward(int b, int n, int npoints, int nsample, int c_in, int c1, int c2,
                           int c3, const float *xyz, const float *center_xyz,
                           const void *features_pm_bf16, const int *idx, float radius,
                           const void *w1_img, const void *w2_img, const void *w3_img,
                           const float *bias, float *out, void *stream);

/* ---------------------------------------------------------------------------------------
 * points_in_boxes.  Replace points_in_boxes_launcher / points_in_boxes_batch_launcher
 *   (batch_size, boxes_num, pts_num, boxes, pts, box_idx_of_points)
 *   ops/roiaware_pool3d/src/points_in_boxes_cuda.cu:107-155 (kernels :49-105), wrappers :157-200.
 * boxes (b, nbox, 7) = x, y, z (bottom centre), w, l, h, rz in LiDAR coordinates; pts (b, npts, 3).
 *   nesie_points_in_boxes       : out (b, npts) int32, index of the first containing box;
 *                                 caller-initialised to -1 (points_in_boxes.py:29-30)
 *   nesie_points_in_boxes_batch : out (b, npts, nbox) int32, 1 where inside, 0 elsewhere (every
 *                                 element is written: unlike the reference no zero fill is needed)
 * Bit-identical masks to the reference kernels (same float / double mix). */
int nesie_points_in_boxes(int b, int nbox, int npts, const float *boxes, const float *pts,
                          int *box_idx_of_points, void *stream);
int nesie_points_in_boxes_batch(int b, int nbox, int npts, const float *boxes, const float *pts,
                                int *box_idx_of_points, void *stream);

/* Max over groups of k consecutive rows of x (groups*k, c) + bias (nullable), the pooling steps of the
 * SidePooling MiniPointNets (models/dense_heads/side_pooling_module.py:360-370):
 *   concat == 0: out (groups, c) = group maxima (torch.max(feature, dim=-1).values)
 *   concat == 1: out (groups*k, 2c) = [ maxima broadcast to the group's rows | x + bias ]
 *                (torch.cat([feature_global.expand(..), feature], dim=1))
 * arg (groups, c) u8 = first maximising row.  _backward writes d_x (groups*k, c) from d_out of the
 * forward's output shape; the maximum's gradient goes to row arg (as torch.max(dim) routes it). */
int nesie_group_max_rows_forward(long long groups, int k, int c, const float *x, const float *bias,
                                 float *out, unsigned char *arg, int concat, void *stream);
/* d_bias_part (nullable): (nesie_group_max_bias_parts(groups, c), c) per-CTA column sums of d_x, i.e.
 * partial gradients of `bias`, to be summed over dim 0 by the caller (0 parts: not available for this c). */
int nesie_group_max_bias_parts(long long groups, int c);
int nesie_group_max_rows_backward(long long groups, int k, int c, const float *d_out,
                                  const unsigned char *arg, float *d_x, int concat,
                                  float *d_bias_part, void *stream);

/* A linear layer commuted with the (weighted) gather in front of it: the first convolution of a
 * SidePooling MiniPointNet (models/dense_heads/side_pooling_module.py:183-243,343-358: rows = [grid point -
 * centre | features interpolated from 3 seeds]) needs its GEMM only over the b * m seeds,
 *   y[r, :] = sum_{i < j} weight[r, i] * table[batch(r), idx[r, i], :] + head[r, 0:3] @ wx^T,
 * table (b, m, c) = seed features @ W_f^T (point-major), idx / weight (b, n, j) with j = 3 (contraction as
 * in three_interpolate) or j = 1 (weight nullable = 1); wx (c, 3) comes with head (b, n, 3) or, for the SA
 * grouping (ops/group_points/group_points.py:104-160: rows = [(neighbour - centre) / radius | features]),
 * with xyz (b, m, 3) + center (b, n / ns, 3): head[r] = (xyz[idx[r]] - center[r / ns]) * (1 / radius)
 * (radius 0: not normalised), j = 1.  y (b * n, c); c / 4 a power of two <= 256.  col_parts (nullable):
 * nesie_gather_linear_parts(b, c, n) blocks of [2][c] column sums of y and y^2 for
 * nesie_bn_rows_forward_fused. */
int nesie_gather_linear_parts(int b, int c, int n);
int nesie_gather_linear_forward(int b, int c, int m, int n, int j, const float *table, const int *idx,
                                const float *weight, const float *head, const float *wx,
                                const float *xyz, const float *center, int ns, float radius, float *y,
                                float *col_parts, void *stream);
/* d_table[batch(r), idx[r, i], :] += weight[r, i] * d_y[r, :] (vector reductions; d_table zeroed by the
 * caller) and, with head (or xyz + center), dwx_part: nesie_gather_linear_parts blocks of [c][4] partial
 * sums of d_y[r, o] * head[r, 0..2] (fourth entry 0) for nesie_gemm_sum_partials. */
int nesie_gather_linear_backward(int b, int c, int m, int n, int j, const float *d_y, const int *idx,
                                 const float *weight, const float *head, const float *xyz,
                                 const float *center, int ns, float radius, float *d_table,
                                 float *dwx_part, void *stream);

/* Inverse-distance interpolation into row-major GEMM rows: the grid features of the SidePooling
 * quality head (models/dense_heads/side_pooling_module.py:183-243, which builds them with a python
 * index_select loop over the batch).  table_pm is the POINT-major (b, m, c) copy of the seed features,
 * idx / weight (b, n, 3) the three neighbours and normalised weights of each of the n grid points,
 * head (b, n, 3) the grid point relative to its box centre.  rows (b*n, ld) receives
 * [ head | sum_j w_j * table[idx_j, :] | zero padding ], contraction as in three_interpolate. */
int nesie_interp_rows(int b, int c, int m, int n, const float *table_pm, const int *idx,
                      const float *weight, const float *head, float *rows, int ld, void *stream);

/* ---------------------------------------------------------------------------------------
 * fp32-accurate row GEMM on tcgen05 (kind::tf32, 3xTF32 split, fp32 TMEM accumulation):
 *   C[r x n] = A[r x k] * B[n x k]^T,  A/C row-major fp32 with leading dimensions lda/ldc.
 * The training-mode shared MLP of the SA / FP modules (the reference's cuDNN 1x1 Conv2d,
 * ops/pointnet_modules/point_sa_module.py:279-288): forward Y = X W^T and dX = dY W.
 * B (the small operand, n <= 256) is pre-split into TF32 hi/lo parts and pre-swizzled by
 * nesie_gemm_pack_b into an image of nesie_gemm_b_image_bytes(n, k) bytes; element (i, j) of B is
 * read from b[i*stride_n + j*stride_k], so W and W^T pack without a transpose.
 */
long long nesie_gemm_b_image_bytes(int n, int k);
int nesie_gemm_pack_b(int n, int k, long long stride_n, long long stride_k, const float *b,
                      void *image, void *stream);
int nesie_gemm_nt_3xtf32(long long r, int n, int k, const float *a, long long lda,
                         const void *b_image, float *c, long long ldc, void *stream);
/* The same GEMM with the neighbouring BatchNorm work fused in (TMA path only: 16-byte aligned rows,
 * n and k multiples of 4; nesie_gemm_fused_supported() tells):
 *   pro_scale / pro_shift (k floats each, both or neither): the A operand is
 *       relu(a * scale + shift), i.e. the training BatchNorm + ReLU of the previous layer applied on
 *       the fly from its statistics, so that the activation itself never goes to memory;
 *   col_stats (nullable): receives nesie_gemm_stats_parts(r) blocks of [2][n] floats, the partial
 *       column sums of C and of C^2 -- the batch statistics of this layer's BatchNorm -- for
 *       nesie_bn_rows_forward_fused. */
int nesie_gemm_fused_supported(long long r, int n, int k, const float *a, long long lda, long long ldc);
int nesie_gemm_stats_parts(long long r);
int nesie_gemm_nt_3xtf32_fused(long long r, int n, int k, const float *a, long long lda,
                               const void *b_image, float *c, long long ldc, const float *pro_scale,
                               const float *pro_shift, float *col_stats, void *stream);
/* Diagnostic only: per-role cycle counters of CTA 0 collected by launches made with the environment
 * variable NESIE_GEMM_DBG & 128 (16 values; reading resets them). */
int nesie_gemm_debug_profile(long long *out16);
/* Weight gradient of the same layer, W'[n x k] = sum over the r rows of A[r, n] * B[r, k]
 * (A = dY, B = X), split over the rows: the kernel writes nsplits partial [n x k] blocks
 * (nsplits = nesie_gemm_wgrad_splits(r, n, k)) into `partials` and the caller sums them, e.g. with
 * nesie_gemm_sum_partials (deterministic; no atomics).  n <= 256, k <= 512. */
int nesie_gemm_wgrad_splits(long long r, int n, int k);
int nesie_gemm_wgrad_3xtf32(long long r, int n, int k, const float *a, long long lda,
                            const float *b, long long ldb, float *partials, int nsplits,
                            void *stream);
/* out[count] = sum of the nparts partial blocks (ascending order: deterministic); count % 4 == 0. */
int nesie_gemm_sum_partials(int nparts, long long count, const float *partials, float *out,
                            void *stream);
/* Data gradient dA_prev = dY W with the BatchNorm-backward statistics of the PREVIOUS layer taken in the
 * epilogue: bn_y (r x n, row stride ldy) is that layer's pre-activation, bn_stats its 4 x n statistics
 * (mean | invstd | scale | shift from the forward); col_stats (nesie_gemm_stats_parts(r) blocks of
 * [2][n]) receives the column sums of g * [relu active] and g * [relu active] * xhat for
 * nesie_bn_relu_rows_backward_fused, which then only finalizes and applies. */
int nesie_gemm_nt_3xtf32_bnbwd(long long r, int n, int k, const float *a, long long lda,
                               const void *b_image, float *c, long long ldc, const float *bn_y,
                               long long ldy, const float *bn_stats, float *col_stats, void *stream);
/* Row GEMM of a MAX-POOLED or GROUP-BIASED layer (MiniPointNet, side_pooling_module.py:343-370; SA
 * pooling, point_sa_module.py:136-158), nesie_gemm_nt_3xtf32_fused plus, taken from the accumulator tile:
 *   pool_max / pool_amax [r / pool_u][n]  per-column maximum over every unit of pool_u (16 or 32, | r)
 *                                          consecutive rows and the first row of the unit that attains
 *                                          it; pool_min / pool_amin (nullable) the minimum likewise.
 *                                          With c == NULL the output itself is not written at all.
 *   grp_bias [r / grp_k][n] (nullable)     added to row i of the output as grp_bias[i / grp_k] (and seen by
 *                                          col_stats and the pooling): the part of the layer's input that
 *                                          is constant over a group's rows (torch.cat([global.expand, x])),
 *                                          pushed through the weights once per group; grp_k a power of
 *                                          two >= 16 that divides r. */
int nesie_gemm_nt_3xtf32_pool(long long r, int n, int k, const float *a, long long lda,
                              const void *b_image, float *c, long long ldc, const float *pro_scale,
                              const float *pro_shift, float *col_stats, int pool_u, float *pool_max,
                              unsigned char *pool_amax, float *pool_min, unsigned char *pool_amin,
                              const float *grp_bias, int grp_k, void *stream);
/* Unit maxima of nesie_gemm_nt_3xtf32_pool (u rows each) -> maxima over groups of k rows (u | k) plus
 * bias (n, nullable): out (groups, n), arg (groups, n) = first maximising row within the group. */
int nesie_pool_finalize(long long groups, int k, int u, int n, const float *pmax,
                        const unsigned char *amax, const float *bias, float *out, unsigned char *arg,
                        void *stream);
/* BatchNorm + ReLU + max-pool of an SA level's last layer (point_sa_module.py:136-158,279-288) from the
 * unit maxima AND minima of nesie_gemm_nt_3xtf32_pool: out[g, c] = relu(scale[c] * (max or min over the
 * group, by the sign of scale) + shift[c]), stats = mean | invstd | scale | shift (4 x n) as
 * nesie_bn_rows_forward_fused leaves them; arg = first row of the group attaining it, 255 when the pooled
 * value is not positive (what nesie_bn_relu_rows_backward expects). */
int nesie_bn_pool_finalize(long long groups, int k, int u, int n, const float *pmax,
                           const unsigned char *amax, const float *pmin, const unsigned char *amin,
                           const float *stats, float *out, unsigned char *arg, void *stream);
/* Weight gradient of a max-pooled convolution out[g, c] = max_j (a[g k + j, :] . w[c, :]):
 * d_w[c, :] = sum_g d_out[g, c] * a[g k + arg[g, c], :] with a = relu(y_prev * scale + shift) (or y_prev
 * itself when scale / shift are NULL), y_prev (groups * k, k_in) row-major, k_in % 4 == 0, k_in <= 256.
 * dw_part receives nesie_pool_wgrad_parts(groups) partial blocks of (n, k_in) for nesie_gemm_sum_partials. */
int nesie_pool_wgrad_parts(long long groups);
int nesie_pool_wgrad(long long groups, int k, int n, int k_in, const float *d_out,
                     const unsigned char *arg, const float *y_prev, const float *scale,
                     const float *shift, float *dw_part, void *stream);
/* Data gradient of the same layer (k <= 64): d_a[g k + j, :] = sum over the channels c with arg[g, c] == j of
 * d_out[g, c] * w[c, :] (w (n, k_in) row-major, n * k_in <= 55296: it is held in shared memory); every
 * row of d_a (groups * k, k_in) is written. */
int nesie_pool_dgrad(long long groups, int k, int n, int k_in, const float *d_out,
                     const unsigned char *arg, const float *w, float *d_a, void *stream);
/* out[g, :] = sum of rows g k .. g k + k - 1 of x (groups * k, n); n % 4 == 0. */
int nesie_group_sum_rows(long long groups, int k, int n, const float *x, float *out, void *stream);
/* out[n] = column sums of x (rows, n), one launch, fixed summation order; n % 4 == 0, n <= 1024.
 * work: nesie_colsum_rows_workspace(n) bytes, 256-byte aligned, zeroed once by the caller; launches that
 * share it must be stream-ordered. */
long long nesie_colsum_rows_workspace(int n);
int nesie_colsum_rows(long long rows, int n, const float *x, float *out, void *work, void *stream);
/* d_x[g k + arg[g, c], c] += d[g, c]: the gradient of a group maximum added in place. */
int nesie_scatter_rows_add(long long groups, int k, int n, const float *d, const unsigned char *arg,
                           float *d_x, void *stream);
/* ... with B = relu(b * scale + shift) applied on the fly (k floats each; TMA path only). */
int nesie_gemm_wgrad_3xtf32_fused(long long r, int n, int k, const float *a, long long lda,
                                  const float *b, long long ldb, const float *pro_scale,
                                  const float *pro_shift, float *partials, int nsplits, void *stream);

/* ---------------------------------------------------------------------------------------
 * Training-mode BatchNorm (batch statistics) + ReLU [+ max-pool over k consecutive rows] on
 * row-major activations y (r x c, c % 4 == 0): the BN2d / ReLU / max_pool2d tail of the SA shared
 * MLP (ops/pointnet_modules/point_sa_module.py:149-150,279-288) in the GEMM formulation.
 * forward : k == 0 -> a (r, c) = relu(bn(y));  k > 0 -> pooled (r/k, c) = max over each group's k
 *           rows, arg (r/k, c) u8 = first maximising row (255: non-positive maximum).  stats (4c)
 *           receives mean | invstd | scale | shift; running_mean/var (nullable) are updated with
 *           `momentum` (unbiased variance), like torch.nn.BatchNorm.
 * backward: d_a is (r, c) for k == 0 or the pooled gradient (r/k, c); writes d_y (r, c),
 *           d_gamma (c), d_beta (c).
 * workspace: nesie_bn_rows_workspace_bytes(c) bytes of scratch, caller-provided.
 */
long long nesie_bn_rows_workspace_bytes(int c);
int nesie_bn_relu_rows_forward(long long r, int c, int k, const float *y, const float *gamma,
                               const float *beta, float eps, float momentum, float *running_mean,
                               float *running_var, float *stats, float *a_or_pooled,
                               unsigned char *arg, void *workspace, void *stream);
/* forward with the options of the fused pipeline: col_partials (nullable) = nparts blocks of [2][c]
 * column sums of y and y^2 from nesie_gemm_nt_3xtf32_fused, replacing the statistics sweep over y;
 * a_or_pooled == NULL: statistics (and running stats) only -- the consumer GEMM applies
 * scale / shift + ReLU in its prologue.  workspace may be NULL when col_partials is given. */
int nesie_bn_rows_forward_fused(long long r, int c, int k, const float *y, const float *gamma,
                                const float *beta, float eps, float momentum, float *running_mean,
                                float *running_var, const float *col_partials, int nparts,
                                float *stats, float *a_or_pooled, unsigned char *arg,
                                void *workspace, void *stream);
int nesie_bn_relu_rows_backward_fused(long long r, int c, const float *y, const float *d_a,
                                      const float *stats, const float *col_partials, int nparts,
                                      float *d_y, float *d_gamma, float *d_beta, void *workspace,
                                      void *stream);
int nesie_bn_relu_rows_backward(long long r, int c, int k, const float *y, const float *d_a,
                                const unsigned char *arg, const float *stats, float *d_y,
                                float *d_gamma, float *d_beta, void *workspace, void *stream);

/* ---------------------------------------------------------------------------------------
 * Target assignment of the train step.
 *
 * nesie_vote_targets: vote part of NesieHead.get_targets_single
 *   (models/dense_heads/nesie_head.py:618-654) = DepthInstance3DBoxes.points_in_boxes
 *   (core/bbox/structures/depth_box3d.py:251-277 -> points_in_boxes_batch_launcher,
 *   ops/roiaware_pool3d/src/points_in_boxes_cuda.cu:128-155) + the per-box python loop.
 *   pts (b, n, pts_stride >= 3) depth frame; boxes (b, g, 7) depth frame, bottom centre, of which
 *   the first nvalid[b] (nullable: all g) are real; idx (b, rows) int64 (nullable: rows == n, every
 *   point) selects the points to evaluate (the vote loss only reads the seeds).  Writes
 *   vote_targets (b, rows, 9) = three (gravity centre - point) slots and vote_mask (b, rows) int64.
 * nesie_chamfer_assign: the argmins of chamfer_distance (models/losses/chamfer_distance.py:49-56),
 *   squared-L2 criterion: idx1 (b, n) = nearest of the first nvalid[b] dst points for every src
 *   point, idx2 (b, m) = nearest src point for every dst point (either may be NULL).
 * nesie_sort_vertices: replaces sort_vertices_wrapper(b, n, m, vertices, mask, num_valid, idx)
 *   (ops/rotated_iou/cuda_op/sort_vert_kernel.cu:136-139, kernel :42-134): vertices (b, n, 24, 2)
 *   normalised around their mean, mask (b, n, 24) bool, num_valid (b, n) int32 -> idx (b, n, 9).
 *   Runs on `stream` (the reference launches on the legacy default stream). */
int nesie_vote_targets(int b, int n, int g, int rows, const float *pts, int pts_stride,
                       const float *boxes, const int *nvalid, const long long *idx,
                       float *vote_targets, long long *vote_mask, void *stream);
int nesie_chamfer_assign(int b, int n, int m, const float *src, const float *dst,
                         const int *nvalid, long long *idx1, long long *idx2, void *stream);
int nesie_sort_vertices(int b, int n, int m, const float *vertices, const unsigned char *mask,
                        const int *num_valid, int *idx, void *stream);
/* nesie_iou3d: cal_iou_3d (ops/rotated_iou/oriented_iou_loss.py:86-109) fused: box1, box2 (n, 7) =
 *   x, y, z, w, h, l, alpha row-wise pairs -> iou (n) and, when jac_box1 != NULL, d iou / d box1 (n, 7)
 *   (the reference differentiates through its tensor formulation; targets carry no gradient). */
int nesie_iou3d(long long n, const float *box1, const float *box2, float *iou, float *jac_box1,
                void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NESIE_B200_H_ */
