/*
 * jmodt_b200.h — C ABI of libjmodt_b200.so, the B200 (sm_100a) implementation of the JMODT
 * per-frame hot path.  This is the drop-in boundary: plain pointers and sizes, no torch
 * types.  Every device pointer is a CUDA device pointer valid on the CURRENT device of the
 * calling thread (set it with jmb_set_device); `stream` is a cudaStream_t passed as void*.
 *
 * Conventions (differences from the reference shims are deliberate and listed here):
 *   - every entry point returns 0 on success or a negative JMB_ERR_* code; nothing ever
 *     calls exit() (the reference launchers do: e.g. ball_query_gpu.cu:62-66);
 *   - kernels are enqueued on `stream` and the call returns immediately; no entry point
 *     allocates, frees or synchronises (the reference does: roipool3d_kernel.cu:214-232,
 *     iou3d.cpp:87-96); scratch memory is passed in by the caller, sized by the
 *     matching *_workspace_bytes() query;
 *   - outputs are fully written by the kernels, so callers need not zero-initialise them
 *     (the reference requires it for ball_query idx and roipool3d outputs);
 *   - re-entrant: entry points may be called from several host threads and for several
 *     devices of one process (the reference's multi-GPU mode is single-process
 *     nn.DataParallel, tools/train.py:86-87).  Whatever is cached (SM counts, the opt-in
 *     shared-memory attribute of a kernel, the address of the tile-scheduler counters — a
 *     __device__ array, so every device has its own copy without an allocation) is kept per
 *     device, and the scheduler slot of a launch comes from an atomic sequence.
 *
 * Each declaration cites the reference interface it replaces (paths relative to the
 * reference repository root).
 */
#ifndef JMODT_B200_H
#define JMODT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define JMB_API __attribute__((visibility("default")))
#else
#define JMB_API
#endif

#define JMB_OK 0
#define JMB_ERR_INVALID_ARG (-1) /* null pointer, negative size, unsupported shape */
#define JMB_ERR_CUDA (-2)        /* a CUDA runtime call or kernel launch failed   */
#define JMB_ERR_WORKSPACE (-3)   /* workspace missing or too small               */

/* ABI version (bumped when a signature changes). */
JMB_API int jmb_version(void);
/* Message describing the last error on the calling thread ("" if none). */
JMB_API const char *jmb_last_error(void);
/* cudaSetDevice for this library's runtime instance; call once per thread before use. */
JMB_API int jmb_set_device(int device);
/* Number of SMs of the current device (used by callers to size batches). */
JMB_API int jmb_sm_count(void);

/* ---- pointnet2 (jmodt/ops/pointnet2/src/pointnet2_api.cpp:10-24) ------------------------ */

/* replaces ball_query_wrapper_fast (ball_query.cpp:14-25) -> ball_query_gpu.cu:9-67.
 * new_xyz (b,m,3), xyz (b,n,3) fp32; idx (b,m,nsample) int32, fully written
 * (rows with no neighbour are all 0). */
JMB_API int jmb_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                   const float *xyz, int *idx, void *stream);

/* Two ball queries over the same centres in one scan (multi-scale grouping, pointnet2_modules.py:41-42 calls
 * ball_query once per radius): same results as two jmb_ball_query calls. */
JMB_API int jmb_ball_query_msg2(int b, int n, int m, float radius_a, int nsample_a, float radius_b, int nsample_b,
                                const float *new_xyz, const float *xyz, int *idx_a, int *idx_b, void *stream);

/* Same contract through a hashed cell list built per call (for large clouds and radii small against their extent, e.g.
 * RPN level 0): a centre visits the 27 neighbouring cells instead of the whole cloud; hits are put in ascending index
 * order before they are written, so the result is IDENTICAL to jmb_ball_query / jmb_ball_query_msg2.  radius_b /
 * nsample_b / idx_b may be 0 / 0 / NULL.  workspace: jmb_ball_query_grid_workspace_bytes(b, n) bytes of device memory. */
JMB_API size_t jmb_ball_query_grid_workspace_bytes(int b, int n);
JMB_API int jmb_ball_query_msg2_grid(int b, int n, int m, float radius_a, int nsample_a, float radius_b, int nsample_b,
                                     const float *new_xyz, const float *xyz, int *idx_a, int *idx_b, void *workspace,
                                     size_t workspace_bytes, void *stream);

/* replaces group_points_wrapper_fast (group_points.cpp:24-35) -> group_points_gpu.cu:47-86.
 * points (b,c,n), idx (b,npoints,nsample) -> out (b,c,npoints,nsample). */
JMB_API int jmb_group_points(int b, int c, int n, int npoints, int nsample, const float *points,
                     const int *idx, float *out, void *stream);

/* replaces group_points_grad_wrapper_fast (group_points.cpp:10-21) -> group_points_gpu.cu:8-44.
 * grad_points (b,c,n) must be zero-initialised by the caller (it is accumulated into). */
JMB_API int jmb_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                          const int *idx, float *grad_points, void *stream);

/* replaces gather_points_wrapper_fast (sampling.cpp:11-21) -> sampling_gpu.cu:8-44. */
JMB_API int jmb_gather_points(int b, int c, int n, int npoints, const float *points, const int *idx,
                      float *out, void *stream);

/* replaces gather_points_grad_wrapper_fast (sampling.cpp:24-35) -> sampling_gpu.cu:46-83.
 * grad_points (b,c,n) must be zero-initialised by the caller. */
JMB_API int jmb_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out,
                           const int *idx, float *grad_points, void *stream);

/* replaces farthest_point_sampling_wrapper (sampling.cpp:38-46) -> sampling_gpu.cu:93-253.
 * dataset (b,n,3); temp (b,n) is an OUTPUT here (final min-distances, identical to what the
 * reference leaves in it) and may be NULL; it need not be pre-filled with 1e10.
 * idxs (b,m) int32.  Tie-breaking reproduces the reference block reduction exactly. */
JMB_API int jmb_furthest_point_sampling(int b, int n, int m, const float *dataset, float *temp,
                                int *idxs, void *stream);

/* Stream gate on sampling progress (no reference counterpart: the reference serialises FPS and its consumers).
 * Returns, in stream order, once idx[f*row_stride + k] >= 0 for every frame f < b and k0 <= k < k1.  The caller
 * fills idx with -1 before launching jmb_furthest_point_sampling on ANOTHER stream; work queued behind the gate can
 * then consume samples k0..k1 while the sampler continues.  Traps after timeout_ms if the producer never writes. */
JMB_API int jmb_wait_indices(const int *idx, int b, int row_stride, int k0, int k1, int timeout_ms, void *stream);

/* replaces three_nn_wrapper_fast (interpolate.cpp:15-25) -> interpolate_gpu.cu:9-74.
 * dist2 (b,n,3) SQUARED distances, idx (b,n,3). */
JMB_API int jmb_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2,
                 int *idx, void *stream);


/* replaces three_interpolate_wrapper_fast (interpolate.cpp:28-40) -> interpolate_gpu.cu:77-117. */
JMB_API int jmb_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                          const float *weight, float *out, void *stream);

/* replaces three_interpolate_grad_wrapper_fast (interpolate.cpp:42-53) -> interpolate_gpu.cu:120-160.
 * grad_points (b,c,m) must be zero-initialised by the caller. */
JMB_API int jmb_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx,
                               const float *weight, float *grad_points, void *stream);

/* ---- roipool3d (jmodt/ops/roipool3d/src/roipool3d.cpp:198-203) --------------------------- */

/* replaces roipool3d_gpu (roipool3d.cpp:48-79) -> roipool3dLauncher (roipool3d_kernel.cu:209-237).
 * xyz (batch,pts_num,3), boxes3d (batch,boxes_num,7) ALREADY ENLARGED, pts_feature
 * (batch,pts_num,feat_len) -> pooled_features (batch,boxes_num,sampled,3+feat_len),
 * pooled_empty_flag (batch,boxes_num) int32.  Both outputs are fully written (empty boxes
 * get an all-zero block and flag 1). */
JMB_API int jmb_roipool3d(int batch, int pts_num, int boxes_num, int feat_len, int sampled,
                  const float *xyz, const float *boxes3d, const float *pts_feature,
                  float *pooled_features, int *pooled_empty_flag, void *stream);

/* Eval-branch RoI pooling fused with the canonical transform
 * (proposal_target_layer.py:99-115 + kitti_utils.py:46-64,152-162): boxes3d are the RAW
 * rois (x, y_bottom, z, h, w, l, ry); they are enlarged by pool_extra_width in-kernel, and
 * the pooled xyz are centred on the roi and rotated by ry about y. */
JMB_API int jmb_roipool3d_canonical(int batch, int pts_num, int boxes_num, int feat_len, int sampled,
                            float pool_extra_width, const float *xyz, const float *boxes3d,
                            const float *pts_feature, float *pooled_features,
                            int *pooled_empty_flag, void *stream);

/* jmb_roipool3d_canonical writing the "head layout" consumed by jmb_rcnn_input_fused: a pooled row is
 * [feature lead..feat_len-1 | x, y, z | feature 0..lead-1 | zero padding], pitch round_up(3+feat_len, 8) floats
 * (lead = 2 for the reference's [mask, depth, 128 channels] feature vector, proposal_target_layer.py:17-34). */
JMB_API int jmb_roipool3d_canonical_head(int batch, int pts_num, int boxes_num, int feat_len, int sampled,
                                 float pool_extra_width, int lead, const float *xyz, const float *boxes3d,
                                 const float *pts_feature, float *pooled_features,
                                 int *pooled_empty_flag, void *stream);

/* ---- iou3d (jmodt/ops/iou3d/src/iou3d.cpp:170-175) --------------------------------------- */

/* replaces boxes_overlap_bev_gpu (iou3d.cpp:31-50) -> iou3d_kernel.cu:223-234,355-365.
 * boxes (n,5) [x1,y1,x2,y2,ry] -> ans (na,nb). */
JMB_API int jmb_boxes_overlap_bev(int na, const float *boxes_a, int nb, const float *boxes_b,
                          float *ans_overlap, void *stream);

/* replaces boxes_iou_bev_gpu (iou3d.cpp:52-71) -> iou3d_kernel.cu:236-248,367-373. */
JMB_API int jmb_boxes_iou_bev(int na, const float *boxes_a, int nb, const float *boxes_b, float *ans_iou,
                      void *stream);

/* boxes_iou3d_gpu (iou3d_utils.py:22-54) as one kernel: boxes (n,7) [x,y,z,h,w,l,ry]. */
JMB_API int jmb_boxes_iou3d(int na, const float *boxes_a, int nb, const float *boxes_b, float *ans_iou,
                    void *stream);

/* Scratch bytes needed by jmb_nms / jmb_nms_normal for n boxes. */
JMB_API size_t jmb_nms_workspace_bytes(int n);

/* replaces nms_gpu (iou3d.cpp:73-118) -> nms_kernel (iou3d_kernel.cu:250-292) + host sweep.
 * boxes (n,5) sorted by descending score.  Unlike the reference, the greedy sweep runs on
 * the device: keep (n) int64 and num_keep (1) int32 are DEVICE pointers; nothing is copied
 * to the host.  max_keep > 0 stops the sweep after that many boxes are kept (the callers
 * only use the first RPN_POST_NMS_TOP_N, proposal_layer.py:113); 0 means no limit. */
JMB_API int jmb_nms(int n, const float *boxes, float thresh, int64_t *keep, int *num_keep, int max_keep,
            void *workspace, size_t workspace_bytes, void *stream);

/* replaces nms_normal_gpu (iou3d.cpp:121-166) -> nms_normal_kernel (iou3d_kernel.cu:306-348). */
JMB_API int jmb_nms_normal(int n, const float *boxes, float thresh, int64_t *keep, int *num_keep,
                   int max_keep, void *workspace, size_t workspace_bytes, void *stream);

/* ---- dense MLP layers on tcgen05 tensor cores ------------------------------------------- */

/* One 1x1-conv layer  Y[g] = act(W . X[g] + bias)  of a SharedMLP / Conv1d stack (reference
 * jmodt/ops/pointnet2/pytorch_utils.py:6-33,127-198; run through cuDNN there).  Channel-first:
 * X[g] (K, N), Y[g] (M, N).  wpack holds W split into bf16 hi/lo chunk images (jmodt_b200/tc.py
 * pack_weights); products are accumulated in fp32 as W_hi.X_hi + W_lo.X_hi + W_hi.X_lo.
 *   mode 0: x is dense (G, K, N) with the given group / row strides (in elements).
 *   mode 1: fused QueryAndGroup / GroupAll (pointnet2_utils.py:231-290): column n of group g is
 *           point idx[g][n] (or n % n_pts if idx is NULL); rows 0-2 are xyz[g][point] - centres[g][n / nsample]
 *           (no centring if centres is NULL), rows 3.. are x[g][k-3][point] with x = feats (G, K-3, n_pts).
 *   out_mode 0: y (G, M, N);  out_mode 1: max over each `pool` consecutive columns -> y (G, M, N / pool)
 *           (the set-abstraction max-pool, pointnet2_modules.py:50-52).
 *   out_mode 2: point-major y (G, N, M) (what the fused set-abstraction kernel gathers from).
 *   y_group_stride: elements between groups of y (0 = dense); lets a layer write into a channel slice of a wider
 *           (G, C_total, N) tensor, replacing torch.cat. */
JMB_API int jmb_tc_mlp_layer(const void *wpack, const float *bias, int M, int K, int G, int N, int mode,
                             const float *x, long long x_group_stride, int x_row_stride, const int *idx,
                             const float *xyz, const float *centres, int nsample, int n_pts, int out_mode,
                             int pool, int relu, float *y, long long y_group_stride, void *stream);

/* one pointwise layer over POINT-MAJOR rows: x (G, N, C) -> y (G, N, M) = act(W x + bias) per row; C a power of two >= 32,
 * x 16-byte aligned.  Same kernel and packed weights as jmb_tc_mlp_layer; the layout sa_fused gathers from. */
JMB_API int jmb_tc_mlp_rows(const void *wpack, const float *bias, int M, int C, int G, int N, const float *x, int relu,
                            float *y, void *stream);

/* two pointwise layers in one launch when the second has ONE output channel (the heads end in a C -> 1 layer: rpn.py:40-47,
 * rcnn.py:91-111, tracker.py:86-109): layer 1 as jmb_tc_mlp_layer (dense x), its activated rows multiplied by dot_w
 * (ceil(M / 128) * 128 floats, zero padded: the second layer's weights) and reduced per 32-row block in fp32.
 * partial (G, 4 * ceil(M / 128), N): sum over dim 1 + the second layer's bias = the second layer's output. */
JMB_API int jmb_tc_mlp_layer_dot(const void *wpack, const float *bias, int M, int K, int G, int N, const float *x,
                                 long long x_group_stride, int x_row_stride, int relu, const float *dot_w,
                                 float *partial, void *stream);

/* out (G, N) = act(bias + sum over dim 1 of partial (G, rows, N)), rows added in index order (shape-independent result) */
JMB_API int jmb_tc_dot_finish(int G, int rows, int N, const float *partial, float bias, int relu, float *out, void *stream);

/* replaces the `conv3x3` layers of BasicBlock (jmodt/detection/modeling/backbone.py:9-30; cuDNN there): 3x3 convolution,
 * padding 1, stride 1 or 2, + bias (eval-mode BatchNorm folded in) + optional ReLU, as an implicit GEMM on the same tcgen05
 * kernel as jmb_tc_mlp_layer (fp32-grade three-term bf16 products).  Channels-last in and out: x (B, H, W, C) with C = 4
 * (RGB + one zero channel) or a power of two >= 32; wpack = the packed (Cout, 9 * C) matrix [cout][3 dy + dx][c] in the
 * layout of jmb_tc_mlp_layer's wpack; bias (ceil(Cout / 128) * 128) or null; y (B, OH, OW, Cout), OH = (H - 1) / stride + 1. */
JMB_API int jmb_tc_conv3x3(const void *wpack, const float *bias, int Cout, int C, int B, int H, int W, int stride,
                           const float *x, int relu, float *y, void *stream);

/* ---- LI-Fusion image feature sampling ---------------------------------------------------- */

/* replaces feature_gather (jmodt/detection/modeling/backbone.py:79-89): bilinear grid_sample with
 * align_corners=True and zero padding.  fmap (b,c,h,w), xy (b,n,2) in [-1,1] -> out (b,c,n). */
JMB_API int jmb_feature_gather(int b, int c, int h, int w, int n, const float *fmap, const float *xy,
                               float *out, void *stream);

/* the same on a channels-last map: fmap (b,h,w,c) — the layout the image convolutions emit (torch.channels_last);
 * c even, fmap 8-byte aligned. */
JMB_API int jmb_feature_gather_nhwc(int b, int c, int h, int w, int n, const float *fmap, const float *xy,
                                    float *out, void *stream);

/* replaces the decoder tail of PointNet2MSG.forward (jmodt/detection/modeling/backbone.py:187-196) for inference:
 *   grid_sample(relu(image_fusion_bn(image_fusion_conv(cat_i DeConv_i(img_i)))), xy)   -> out (b, 32, n)
 * evaluated only at the four bilinear taps of every point, so neither full-resolution map is written.
 * m0..m3: channels-last source maps (b, h >> (i+1), w >> (i+1), c_i), c_i multiples of 64, sum <= 1024; h, w multiples
 * of 16 (the padded image, 384 x 1280).  wexp (256, sum c_i / 32, 2, 16, 16) 32-bit words: for phase
 * (y mod 16) * 16 + (x mod 16) and 32-channel chunk k of the concatenated levels, the (16 out, 32 in) slice
 * DeConv_i.weight[chunk, :, y mod s_i, x mod s_i]^T as two planes hi + lo of bf16 PAIRS along the input channel (even
 * channel in the low half; hi = bf16(w), lo = bf16(w - hi): the products run on bf16 tensor-core instructions as
 * xh wh + xl wh + xh wl, fp32 accumulate);
 * w1 (32, 64) / b1 (32): image_fusion_conv with the BatchNorm (eval) affine and the DeConv biases folded in.
 * workspace: jmb_decode_workspace_bytes(b, n) bytes, 16-byte aligned.  Six launches on `stream`. */
JMB_API long long jmb_decode_workspace_bytes(int b, int n);
JMB_API int jmb_decode_gather(int b, int n, int h, int w, const float *xy, const float *m0, const float *m1,
                              const float *m2, const float *m3, int c0, int c1, int c2, int c3, const void *wexp,
                              const float *w1, const float *b1, void *workspace, float *out, void *stream);

/* HOST helpers of the reference's roipool3d module (roipool3d.cpp:97-195 `pts_in_boxes3d_cpu`, `roipool3d_cpu`; called by
 * the reference's dataset code on CPU tensors).  All pointers are HOST pointers; no CUDA call is made.  They complete the
 * operator API; they are not a fallback of the device path.
 * pts (n_pts,3), boxes3d (n_boxes,7) [x,y_bottom,z,h,w,l,ry] -> flags (n_boxes,n_pts) int64 0/1 */
JMB_API int jmb_pts_in_boxes3d_host(int n_pts, int n_boxes, const float *pts, const float *boxes3d, int64_t *flags);
/* first `sampled` in-box points per box in index order, wrap-around padding, empty flag; pooled arrays must be zeroed by the
 * caller for empty boxes as in the reference (roipool3d_utils.py:63-65) */
JMB_API int jmb_roipool3d_host(int n_pts, int n_boxes, int feat_len, int sampled, const float *pts, const float *boxes3d,
                               const float *pts_feature, float *pooled_pts, float *pooled_features, int64_t *empty_flag);

/* One whole single-scale set-abstraction layer (reference pointnet2_modules.py:20-63 with QueryAndGroup and a
 * 3-layer SharedMLP) in ONE kernel: grouped gather -> MLP -> max over nsample.  The first layer is linear up to its
 * ReLU and grouping only selects columns, so it is applied BEFORE the gather: the caller passes
 *     z   (G, n_pts, C1) point-major = W1[:, 3:] . features + b1 of the n_pts points (one dense layer with bias,
 *         jmb_tc_mlp_layer with out_mode 2; NULL when the layer has no input features — b1 is then taken from w1x), and
 *     w1x (C1, 4) fp32 rows [W1[k, 0], W1[k, 1], W1[k, 2], b1[k]] (the coordinate columns and the bias) in HOST memory:
 *         the 2 KB table is copied into the launch parameters (constant bank), so a captured launch keeps the values
 *         it was captured with,
 * and the kernel's gather finishes the layer, relu(z[idx] + W1x . (xyz[idx] - centre)), while it stages the operand.
 * Layer widths C1, C2 <= 128 (C1 % 8 == 0) and C3 <= 256; w2 / w3 are packed layers (jmodt_b200/tc.py) ZERO-PADDED to
 * 128 x 128 and (128 or 256) x 128; nsample in {8,16,32,64}, npoint*nsample a multiple of 128.
 * idx (G, npoint, nsample), xyz (G, n_pts, 3), centres (G, npoint, 3) -> out (G, C3, npoint), or point-major
 * (G, npoint, C3) if out_point_major != 0. */
JMB_API int jmb_sa_fused(const float *z, const float *w1x, const void *w2, const float *b2, const void *w3,
                         const float *b3, int C1, int C2, int C3, int G, int npoint, int nsample, int n_pts,
                         const int *idx, const float *xyz, const float *centres, float *out, int out_point_major,
                         void *stream);

/* LI-Fusion attention weight (reference backbone.py:33-58, IALayer: three Linear layers, tanh, sigmoid) as one fp32 kernel:
 * att[b,n] = sigmoid(w3 . tanh(W1 . img[b,:,n] + W2 . pt[b,:,n] + b12) + b3).  img (B, ic, N), pt (B, pc, N) channel-first;
 * w12 (rc, ic+pc) = [W1 | W2] row-major, b12 (rc) = b1 + b2, w3 (rc); rc <= 64, (ic+pc) * rc_padded * 4 <= 200 KB.  att (B, N). */
JMB_API int jmb_ia_attention(int B, int ic, int pc, int rc, int N, const float *img, const float *pt, const float *w12,
                             const float *b12, const float *w3, float b3, float *att, void *stream);

/* Deterministic accumulation for the three backward ops (the reference uses float atomicAdd: group_points_gpu.cu:8-25,
 * sampling_gpu.cu:46-63, interpolate_gpu.cu:120-142, so its gradients differ in the last bits from run to run).
 * order (B, Lq): positions of the batch-local flat index list sorted (stably) by target; seg_off (B, n_tgt+1): start of each
 * target's run; out[b][c][t] = sum over the run, in order, of src[b][c][q / rep] * (weight ? weight[b][q] : 1).
 * group_points_grad: Lq = npoint*nsample, rep 1; gather_points_grad: Lq = npoint, rep 1; three_interpolate_grad: Lq = 3n,
 * rep 3, weight = the interpolation weights.  out is fully written. */
JMB_API int jmb_segmented_scatter_add(int B, int C, int L_src, int Lq, int n_tgt, int rep, const float *src,
                                      const int *order, const int *seg_off, const float *weight, float *out,
                                      void *stream);

/* First SharedMLP layer of a set-abstraction level applied before the gather, for the layer-by-layer path (levels whose
 * later layers are too wide for jmb_sa_fused): out (G, C1, npoint*nsample) channel-first
 *   = relu(z[idx] + W1x . (xyz[idx] - centre)),  z (G, n_pts, C1) point-major = W1[:, 3:] . features + b1 (a dense layer over
 * the POINTS), w1x (C1, 4) DEVICE rows [W1[k,0], W1[k,1], W1[k,2], b1[k]].  Replaces QueryAndGroup + the first Conv2d
 * (pointnet2_utils.py:241-264, pytorch_utils.py:6-33) = a grouped GEMM over all npoint*nsample columns. */
JMB_API int jmb_sa_first_layer(const float *z, const float *w1x, int C1, int G, int npoint, int nsample, int n_pts,
                               const int *idx, const float *xyz, const float *centres, float *out, void *stream);

/* Input stage of the per-proposal network (reference rcnn.py:172-186: xyz_up_layer 5->128->128, cat with the 128
 * RPN channels, merge_down_layer 256->128) in ONE kernel over consecutive rows of the pooled tensor in the
 * "head layout" written by jmb_roipool3d_canonical_head: in (rows, 136) = [128 channels | x,y,z,mask,depth | 0,0,0]
 * -> out (rows, 128) point-major, or — rows_per_group > 0 (a multiple of 128) — channel-first
 * (rows / rows_per_group, 128, rows_per_group).  w1 is the packed 128 x 8 first layer, w2 128 x 128, w3 128 x 256. */
JMB_API int jmb_rcnn_input_fused(const void *w1, const float *b1, const void *w2, const float *b2, const void *w3,
                                 const float *b3, long long rows, int row_pitch, const float *in, float *out,
                                 int rows_per_group, void *stream);

/* Pair correlation features of the link / start-end heads (reference jmodt/tracking/tracker.py:81-112,
 * rcnn.py:239-258): pt (G, K, P) predecessor and dt (G, K, D) successor features, channel-first ->
 * cor (G, K, P*D) = |pt[.., i] - dt[.., j]| at column i*D + j, mean_over_p (G, K, D) = mean_i cor and
 * mean_over_d (G, K, P) = mean_j cor.  Any of the three outputs may be NULL. */
JMB_API int jmb_pair_corr(int G, int K, int P, int D, const float *pt, const float *dt, float *cor,
                          float *mean_over_p, float *mean_over_d, void *stream);

/* Per-point feature vector of the RoI-pooling stage (reference proposal_target_layer.py:17-34, point_rcnn.py:47):
 * feat (B, C, N) channel-first and E <= 2 per-point scalars extra0/extra1 (B, N) -> out (B, N, E + C) =
 * [extra0, extra1, feat[:, :, n]] — torch.cat((mask, depth, features.permute(0, 2, 1)), dim=2) as one tiled transpose. */
JMB_API int jmb_pack_point_features(int B, int C, int N, int E, const float *feat, const float *extra0,
                                    const float *extra1, float *out, void *stream);

/* Tracker association inputs (reference jmodt/tracking/data_association.py:10-28,42-45): boxes (n, 7)
 * [x, y, z, h, w, l, ry] -> dist (na, nb) = 1 - |centre_a - centre_b| / max corner-to-corner distance, and, when
 * link / iou (na, nb) are given, score = link * w_app + iou * w_iou + dist * w_dis.  dist or score may be NULL. */
JMB_API int jmb_boxes_dist(int na, const float *boxes_a, int nb, const float *boxes_b, float *dist,
                           const float *link, const float *iou, float w_app, float w_iou, float w_dis,
                           float *score, void *stream);

/* ---- proposal layer ---------------------------------------------------------------------- */

/* Scratch bytes for jmb_proposal_layer. */
JMB_API size_t jmb_proposal_workspace_bytes(int B, int N, int pre_top_n, int post_top_n);

/* replaces the per-frame loop of ProposalLayer.forward + distance_based_proposal (reference
 * jmodt/detection/layers/proposal_layer.py:36-121) for a whole batch, on the device: proposals (B,N,7) decoded
 * boxes [x, y_bottom, z, h, w, l, ry], scores (B,N), order (B,N) int64 = argsort(scores, descending) per frame.
 * Two distance bins (0,40] / (40,80] with 70 % / 30 % of the pre- and post-NMS quotas; NMS is axis-aligned
 * (rotated = 0, cfg.RPN.NMS_TYPE = 'normal') or rotated.  ret_boxes (B, post_top_n, 7) and ret_scores
 * (B, post_top_n) are zero padded like the reference's. */
JMB_API int jmb_proposal_layer(int B, int N, const float *proposals, const float *scores, const long long *order,
                               int pre_top_n, int post_top_n, float nms_thresh, int rotated, float *ret_boxes,
                               float *ret_scores, void *workspace, size_t workspace_bytes, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* JMODT_B200_H */
