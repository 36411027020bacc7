/*
 * mvs_b200.h -- C ABI of the B200-native MVSNet-family cost-volume hot path.
 *
 * The reference (doubleZ0108/MVS) has no FFI / plugin registry for this path: its boundary is a set
 * of Python callables in each project's models/module(s).py.  This header is the
 * C-ABI a maintainer would bind behind those callables; every entry point cites the reference
 * interface it replaces.  INTEGRATION.md shows the ctypes binding and the `patch_reference` hook.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer borrowed from the caller (torch owns the memory) unless
 *     the parameter name ends in `_host`; nothing is allocated or retained inside the library;
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *     all work is enqueued on it and the call returns without synchronising;
 *   - return value: MVS_OK (0) or a negative MVS_ERR_* code; mvs_last_error() gives the message
 *     for the calling thread.  No global mutable state => re-entrant across threads / streams;
 *   - "NCHW"/"NCDHW" are the reference's contiguous fp32 layouts.  "C8" is this library's fast
 *     layout: channels blocked by 8, [B][C/8][D][H][W][8] bf16 (16 B per voxel per block), see
 *     DESIGN.md "Data layout in HBM".
 *   - rot [B,nsrc,9] / trans [B,nsrc,3]: rows of proj[:, :3, :3] and proj[:, :3, 3] where
 *     proj = src_proj @ inverse(ref_proj) (MVSNet/models/module.py:63-65), fp32, computed by the
 *     caller with torch so that both sides of every parity test consume identical bits.
 */
#ifndef MVS_B200_H
#define MVS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVS_OK 0
#define MVS_ERR_INVALID (-1)      /* bad argument (shape, null pointer, unsupported combination) */
#define MVS_ERR_CUDA (-2)         /* a CUDA runtime / driver call failed                          */
#define MVS_ERR_UNSUPPORTED (-3)  /* valid request this build cannot serve                        */

/* flags (bit-or) */
#define MVS_ALIGN_CORNERS 1     /* grid_sample(align_corners=True): MVSNet_pl/models/modules.py:57-59 */
#define MVS_PL_ORDER 2          /* MVSNet_pl op order R@(xyz*d)+T: MVSNet_pl/models/modules.py:46-50  */
#define MVS_REF_SUM_SQUARED 4   /* CVP aliasing quirk: CVP-MVSNet/models/net.py:129-130, modules.py:228-229 */
#define MVS_RELU 8              /* conv epilogue: apply ReLU after the affine                    */
#define MVS_CLAMP_INDEX 16      /* CasMVSNet/models/cas_mvsnet.py:62 depth_index.clamp(0, D-1)    */
#define MVS_INPUT_IS_PROB 32    /* softargmin: input already soft-maxed (depth_regression(p, d)) */
#define MVS_BLEND_BF16 64       /* C8 builder: bilinear blend in packed bf16x2 (sums / variance stay fp32) */
#define MVS_FAST_COORDS 128     /* mvs_warp_taps: probe the C8 builder's division-free-call tap arithmetic */
#define MVS_FEAT_F16 256        /* C8 builder: feature maps are fp16 C8 (mvs_pack_c8h); blend in packed fp16  */
#define MVS_WARP_NO_TMA 512     /* C8 builder: force the L1-gather kernel                                       */
#define MVS_ACT_F16 2048        /* conv3d_c8: activations, packed weights and output are fp16 C8 instead of bf16 C8   */
#define MVS_X_DW 4096          /* conv3d_c8 (stride 2): x has W de-interleaved, column w at (w&1)*ceil(W/2) + (w>>1)      */
#define MVS_Y_DW 8192          /* conv3d_c8 (stride 1): write y W-de-interleaved (for a stride-2 consumer / a skip add)    */
#define MVS_SKIP_DW 16384      /* conv3d_c8: the skip tensor is W-de-interleaved                                            */
#define MVS_KD1 32768          /* conv3d_c8 (stride 1): the weights are zero outside the centre depth tap (kd = 1) -- a 2D
                                  convolution over D stacked images; the kernel skips the other depth taps' MMAs            */
#define MVS_FLAT2D 65536       /* conv3d_c8 + pack_weights_ex (stride 1, D = 1): plain 2D convolution with the centre depth slice of
                                  the weights; the kernel tiles the image rows instead of a depth axis (no step-tap partials)  */
#define MVS_SKIP_PS 131072     /* conv3d_c8 + MVS_FLAT2D: the skip operand is a half-resolution map [N][4*Cout/8][H/2][W/2][8] whose block
                                  ((h&1)*2 + (w&1)) * Cout/8 + cb holds the value for output pixel (h, w) ("pixel-shuffled")     */
#define MVS_WARP_TMA 1024       /* C8 builder: force the TMA-staged kernel (fp16 maps only); neither bit: library picks */

/* depth_mode */
#define MVS_DEPTH_PLANE 0       /* depth [B,D]       MVSNet/models/module.py:46                   */
#define MVS_DEPTH_PIXEL 1       /* depth [B,D,H,W]   CasMVSNet/models/module.py:245,267-268        */

/* dtype */
#define MVS_F32 0
#define MVS_BF16 1
#define MVS_F16 2
#define MVS_U8 3

#define MVS_MAX_SRC 8           /* source views the fused builder takes in one launch            */

int mvs_version(void);
int mvs_sm(void);                         /* architecture the kernels were compiled for (100)    */
const char *mvs_last_error(void);         /* thread-local, never NULL                            */
int64_t mvs_launch_count(void);           /* kernels launched by this library since load         */

/* ---- a1: homography warp --------------------------------------------------------------------
 * Replaces homo_warping(src_fea, src_proj, ref_proj, depth_values) -> [B,C,D,H,W]
 *   MVSNet/models/module.py:46-87, CasMVSNet/models/module.py:245-280,
 *   CVP-MVSNet/models/modules.py:81-119, and homo_warp: MVSNet_pl/models/modules.py:25-62
 *   (flags = MVS_ALIGN_CORNERS|MVS_PL_ORDER).  src_fea NCHW fp32, rot [B,9], trans [B,3],
 *   out NCDHW fp32.  Bit-exact with the reference's CPU path. */
int mvs_warp_fwd(const float *src_fea, const float *rot, const float *trans, const float *depth,
                 int depth_mode, float *out, int B, int C, int D, int H, int W, int flags,
                 void *stream);

/* Backward of mvs_warp_fwd w.r.t. src_fea (the grid is built under no_grad, module.py:62):
 * grad_src[b,c,tap] += grad_out[b,c,d,y,x] * w_tap.  grad_src is a caller-zeroed NCHW fp32 buffer. */
int mvs_warp_bwd(const float *grad_out, const float *rot, const float *trans, const float *depth,
                 int depth_mode, float *grad_src, int B, int C, int D, int H, int W, int flags,
                 void *stream);

/* Integer bilinear tap indices of the same warp (floor(ix), floor(iy) saturated to int32) and the
 * 4-bit in-bounds mask (bit0 nw, bit1 ne, bit2 sw, bit3 se); ixy (optional) receives the fp32
 * sample position.  All [B,D,H,W] (ixy: [B,D,H,W,2]).  Used to prove index parity. */
int mvs_warp_taps(const float *rot, const float *trans, const float *depth, int depth_mode,
                  int32_t *x0, int32_t *y0, uint8_t *mask, float *ixy, int B, int D, int H, int W,
                  int flags, void *stream);

/* ---- a1+a2 fused: warp of every source view + variance over views, volume written once -------
 * Replaces the builder loops MVSNet/models/mvsnet.py:152-170, CasMVSNet/models/cas_mvsnet.py:24-46,
 * CVP-MVSNet/models/net.py:127-152 and proj_cost modules.py:221-275 (MVS_REF_SUM_SQUARED).
 * srcs_host: HOST array of nsrc (<= MVS_MAX_SRC) device pointers, each [B,C,H,W] like ref.
 * Strict variant: fp32 NCHW in, fp32 NCDHW out, bit-exact with the reference's CPU path. */
int mvs_warp_variance_fwd(const float *ref, const float *const *srcs_host, int nsrc,
                          const float *rot, const float *trans, const float *depth, int depth_mode,
                          float *out, int B, int C, int D, int H, int W, int flags, void *stream);

/* Backward of the builder w.r.t. the feature maps (the grid is built under no_grad in the
 * reference, module.py:62).  grad_out NCDHW fp32; grad_ref / grad_srcs accumulate (+=) into
 * caller-zeroed NCHW fp32 buffers. */
int mvs_warp_variance_bwd(const float *grad_out, const float *ref, const float *const *srcs_host,
                          int nsrc, const float *rot, const float *trans, const float *depth,
                          int depth_mode, float *grad_ref, float *const *grad_srcs_host, int B, int C,
                          int D, int H, int W, int flags, void *stream);

/* Fast variant: C8 bf16 feature maps in ([B][C/8][H][W][8]), C8 bf16 volume out
 * ([B][C/8][D][H][W][8]); fp32 coordinates (same tap indices as the strict path), fp32 blend and
 * accumulation, one bf16 rounding at the store.  C % 8 == 0. */
int mvs_warp_variance_c8_fwd(const void *ref_c8, const void *const *srcs_c8_host, int nsrc,
                             const float *rot, const float *trans, const float *depth,
                             int depth_mode, void *out_c8, int B, int C, int D, int H, int W,
                             int flags, void *stream);

/* ---- layout hand-off ----------------------------------------------------
 * NC(D)HW (fp32 or bf16) <-> C8 bf16.  `inner` = D*H*W (or H*W).  C is zero-padded up to a
 * multiple of 8 on pack; unpack drops the padding. */
int mvs_pack_c8(const void *src, int src_dtype, void *dst_c8, int B, int C, int64_t inner, void *stream);
int mvs_unpack_c8(const void *src_c8, void *dst, int dst_dtype, int B, int C, int64_t inner, void *stream);
/* Same layout with fp16 elements ("C8H"), the feature-map format of the fast builder (MVS_FEAT_F16): 11-bit
 * significands instead of bf16's 8, and the bilinear blend runs on HFMA2 without per-tap unpacking.  Values are
 * clamped to +-65504 (fp16 max) instead of overflowing to infinity. */
int mvs_pack_c8h(const void *src, int src_dtype, void *dst_c8h, int B, int C, int64_t inner, void *stream);

/* ---- f3 (next row): the FeatureNet hand-off around the 3x3 convolutions (CasMVSNet/models/module.py:304-405) --------
 * The 2D FPN extractor's 3x3 layers run on mvs_conv3d_c8_fwd with D = 1 and MVS_ACT_F16; these three kernels are the rest:
 *   mvs_img_to_c8h     images [N,3,H,W] (MVS_F32, or MVS_U8 scaled by 1/255 like CasMVSNet/datasets/general_eval.py:81-86)
 *                      -> fp16 C8 [N,1,H,W,8], channels 3..7 zero
 *   mvs_s2d_c8         2x2 space-to-depth [N,CB,H,W,8] -> [N,4*CB,ceil(H/2),ceil(W/2),8], block order (py*2+px)*CB + cb:
 *                      turns the 5x5 stride-2 pad-2 layers (module.py:336,342) into 3x3 stride-1 layers over 4*Cin channels
 *   mvs_fpn_merge_c8h  out = nearest_up2(prev) + conv1x1(x) + bias (module.py:393-398): x [N,Cin/8,H,W,8], Cin in {8,16};
 *                      w_host [32,Cin] / bias_host [32] are HOST pointers (copied into the kernel parameter block);
 *                      prev [N,4,Hp,Wp,8] or NULL; out [N,4,H,W,8]; all maps fp16 C8. */
int mvs_img_to_c8h(const void *img, int src_dtype, void *dst_c8h, int N, int H, int W, void *stream);
/* Map layouts (flags): batch-major [N][CB][H][W][8] (default; what the builder takes) or "folded" [CB][N][H][W][8] = the
 * bytes mvs_conv3d_c8_fwd reads as ONE volume [1][CB][D=N][H][W][8], so the N images ride the row axis of the convolution.
 *   mvs_s2d_c8: MVS_MAP_SRC_FOLDED, MVS_MAP_DST_FOLDED;  mvs_fpn_merge_c8h: MVS_MAP_SRC_FOLDED (x), MVS_MAP_DST_FOLDED (out),
 *   MVS_MAP_PREV_FOLDED (prev). */
#define MVS_MAP_SRC_FOLDED 1
#define MVS_MAP_DST_FOLDED 2
#define MVS_MAP_PREV_FOLDED 4
int mvs_s2d_c8(const void *src_c8, void *dst_c8, int N, int CB, int H, int W, int flags, void *stream);
int mvs_fpn_merge_c8h(const void *x_c8h, const float *w_host, const float *bias_host, const void *prev_c8h,
                      void *out_c8h, int N, int Cin, int H, int W, int Hp, int Wp, int flags, void *stream);
/* y[n, c, h, w] += corr[(rc * 3 + cc) * C + c] on the one-pixel border of fp16 C8 maps [N][C/8][H][W][8] (rc / cc = 0 first, 1
 * inner, 2 last row / column; corr_host [9][C] is a HOST pointer, C <= 32).  Used when the last FPN stage is evaluated by
 * linearity -- out3(up2(intra) + inner2(conv0)) as two convolutions (featurenet.py) -- for the lateral bias, which a zero-padded
 * 3x3 convolution does not see outside the image (CasMVSNet/models/module.py:393-398). */
int mvs_border_add_c8h(void *y_c8h, const float *corr_host, int N, int C, int H, int W, void *stream);

/* ---- a3: 3x3x3 convolution + folded BatchNorm + ReLU + skip ------------------------------------
 * Replaces ConvBnReLU3D / Conv3d / Deconv3d blocks and the skip adds of CostRegNet.forward:
 *   MVSNet/models/module.py:26-33, mvsnet.py:55-93; CasMVSNet/models/module.py:115-200,407-438;
 *   CVP-MVSNet/models/net.py:52-89.
 * y = [skip +] act( conv(x, w) * scale[co] + shift[co] ),  kernel 3, padding 1.
 *   transposed = 0: weight [Cout,Cin,3,3,3], out extent = (n-1)/stride+1
 *   transposed = 1: ConvTranspose3d(stride, padding=1, output_padding=stride-1), weight
 *                   [Cin,Cout,3,3,3], out extent = n*stride
 * scale/shift may be NULL (identity / zero); skip has the output's shape.  D,H,W are INPUT extents.
 * Strict variant: fp32 NCDHW, fp32 FMA accumulation. */
int mvs_conv3d_fwd(const float *x, const float *w, const float *scale, const float *shift,
                   const float *skip, float *y, int B, int Cin, int Cout, int D, int H, int W,
                   int stride, int transposed, int flags, void *stream);

/* Weight gradient of the same layer (training: loss.backward(), CasMVSNet/train.py:165-170): strict fp32.
 * x [B,Cin,D,H,W] (the layer's input), grad_y = dL/dy (the layer's output extents) -> gw [Cout,Cin,3,3,3]
 * (transposed: [Cin,Cout,3,3,3]).  gw must be ZERO on entry (CTAs accumulate with atomicAdd).  The DATA gradient needs no
 * entry point of its own: d/dx of a (strided) convolution is mvs_conv3d_fwd(grad_y, w, transposed = !transposed). */
int mvs_conv3d_wgrad(const float *x, const float *grad_y, float *gw, int B, int Cin, int Cout, int D, int H, int W,
                     int stride, int transposed, void *stream);

/* Train-mode BatchNorm3d (+ ReLU) of the conv blocks under training (MVSNet/models/module.py:26-33,
 * CasMVSNet/models/module.py:139,182; replaces F.batch_norm(training=True) + F.relu and their autograd): full-grid streaming
 * passes over x [B,C,S] fp32 (S = D*H*W).  sums [2*C] double must be ZERO on entry (atomicAdd).
 *   mvs_bn_stats      sums[c] = sum x, sums[C+c] = sum x^2
 *   mvs_bn_apply      y = [relu](x * a[c] + k[c])                          a = invstd*gamma, k = beta - mean*a
 *   mvs_bn_bwd_stats  sums[c] = sum g, sums[C+c] = sum g*xhat              g = dy * [x*a + k > 0] (relu) | dy
 *   mvs_bn_bwd_apply  dx = ga[c] * (g - mg[c] - xhat * mgx[c])             ga = gamma*invstd, mg = sum_g/M, mgx = sum_gx/M */
int mvs_bn_stats(const float *x, double *sums, int B, int C, int64_t S, void *stream);
int mvs_bn_apply(const float *x, const float *a, const float *k, float *y, int B, int C, int64_t S, int relu, void *stream);
int mvs_bn_bwd_stats(const float *x, const float *dy, const float *a, const float *k, const float *mean, const float *invstd,
                     double *sums, int B, int C, int64_t S, int relu, void *stream);
int mvs_bn_bwd_apply(const float *x, const float *dy, const float *a, const float *k, const float *mean, const float *invstd,
                     const float *ga, const float *mg, const float *mgx, float *dx, int B, int C, int64_t S, int relu,
                     void *stream);

/* Fast variant: C8 bf16 activations, tcgen05 (UMMA) implicit GEMM, operands staged in shared memory by the TMA engine
 * (1-D bulk copies per staged line in the stride-1 layers, cp.async in the stride-2 / transposed layers), TMEM
 * accumulators; weights pre-packed by mvs_conv3d_c8_pack_weights.  y is C8 bf16, or fp32
 * [B,D,H,W] when Cout == 1 (the `prob` layer). */
int64_t mvs_conv3d_c8_packed_weight_bytes(int Cin, int Cout, int stride, int transposed);
int mvs_conv3d_c8_pack_weights(const float *w, void *packed, int Cin, int Cout, int stride,
                               int transposed, void *stream);
/* flags: MVS_ACT_F16 packs fp16 blocks for mvs_conv3d_c8_fwd(..., flags | MVS_ACT_F16) -- the 2D feature extractor
 * (CasMVSNet/models/module.py:304-405) runs on this kernel with D = 1 and fp16 C8 activations. */
int mvs_conv3d_c8_pack_weights_ex(const float *w, void *packed, int Cin, int Cout, int stride,
                                  int transposed, int flags, void *stream);
int mvs_conv3d_c8_fwd(const void *x_c8, const void *w_packed, const float *scale, const float *shift,
                      const void *skip_c8, void *y, int B, int Cin, int Cout, int D, int H, int W,
                      int stride, int transposed, int flags, void *stream);

/* Profiling hook (tools/prof_conv_trace.py): when dev_buf != NULL, the first n_ctas CTAs of every following
 * mvs_conv3d_c8_fwd launch from this host thread add their per-role clock64 timers to dev_buf[cta * 16 + k]
 * (int64; k documented at RoleTimer in csrc/conv3d_umma.cu).  NULL switches it off.  Not part of the data path. */
int mvs_conv3d_c8_set_trace(void *dev_buf, int n_ctas);

/* ---- a4+a5+a6: softmax over D + soft-argmin depth + photometric confidence ---------------------
 * Replaces F.softmax(cost_reg, 1) + depth_regression + the pad/avg_pool3d/gather confidence:
 *   MVSNet/models/mvsnet.py:183-191, module.py:91-103; CasMVSNet/models/cas_mvsnet.py:51-64
 *   (MVS_CLAMP_INDEX); CVP-MVSNet/models/net.py:162-163,185-199, modules.py:338-355.
 * logits [B,D,H,W] fp32; depth [B,D] or [B,D,H,W]; out_depth, out_conf [B,H,W] fp32;
 * out_prob [B,D,H,W] (optional), out_index int32 [B,H,W] (optional). */
int mvs_softargmin_conf_fwd(const float *logits, const float *depth, int depth_mode,
                            float *out_depth, float *out_conf, float *out_prob, int32_t *out_index,
                            int B, int D, int H, int W, int flags, void *stream);

/* ---- f2: depth-hypothesis generation (the step before the path) -------------------------------
 * Replaces get_depth_range_samples (per-pixel branch) CasMVSNet/models/module.py:485-524:
 * out[b,k,y,x] = (cur - D/2*interval) + k * ((cur + D/2*interval) - (cur - D/2*interval))/(D-1).
 * `interval` is the Python double depth_inteval_pixel; D/2*interval is formed in double and cast to
 * fp32 once, like the reference's tensor-minus-scalar.  cur [B,H,W] fp32 -> out [B,D,H,W] fp32,
 * bit-exact with the reference's CPU path. */
int mvs_depth_range_samples(const float *cur, double interval, int ndepth, float *out, int B, int H,
                            int W, void *stream);

/* Fused form of the whole inter-stage step of CascadeMVSNet.forward (cas_mvsnet.py:129-151): bilinear
 * up-sampling of the previous stage's depth [B,hp,wp] to the image extent [H,W] (align_corners=False),
 * get_depth_range_samples around it, and trilinear resampling of the [B,D,H,W] samples to the stage
 * extent [D,h,w] -- written directly as out [B,D,h,w] fp32; the full-resolution volume never exists. */
int mvs_cas_hypotheses(const float *prev_depth, int hp, int wp, int H, int W, int h, int w, int ndepth,
                       double interval, float *out, int B, void *stream);

/* Relative poses of a CasMVSNet projection block in one launch (fast path; replaces ~25 tiny ATen / cuSOLVER launches per
 * stage: cas_mvsnet.py:30-33 K[:3,:3] @ E[:3,:4], module.py:257-259 src @ inverse(ref)).  proj [n_sets,N,2,4,4] fp32
 * ([..,0] extrinsic, [..,1,:3,:3] intrinsic) -> rot [n_sets,N-1,9], trans [n_sets,N-1,3]; fp64 inside, rounded once. */
int mvs_cas_poses(const float *proj, float *rot, float *trans, int n_sets, int N, void *stream);

/* CVP-MVSNet's statistical hypothesis interval (calDepthHypo test branch, CVP-MVSNet/models/modules.py:146-209):
 * sum over the image of |delta_d| per batch element, float64.  ref_depth [B,H,W] fp32 (the up-sampled depth map);
 * cams [B,59] doubles = inv(K_ref)[9] | inv(E_ref)[16] | E_src[16] | K_src[9] | (K_ref R_ref) inv(K_src R_src)[9], derived by
 * the caller with the reference's own torch ops; sum_abs [B] doubles, ZERO on entry (the mean is sum / (H * W)). */
int mvs_cvp_depth_interval(const float *ref_depth, const double *cams, double *sum_abs, int B, int H, int W,
                           double pixel_interval, void *stream);

/* ---- f4: geometric-consistency filter of estimated depth maps (the step after the path) ------
 * Replaces reproject_with_depth + check_geometric_consistency (MVSNet/eval.py:138-208, CasMVSNet/test.py:237-294)
 * and the per-reference-view fusion loop of filter_depth (MVSNet/eval.py:240-263).  Depth / confidence maps fp32 [H,W]
 * (device).  `cam` (device, float64) holds, per (ref, src) pair, the 60 doubles the reference derives with
 * np.linalg.inv / np.matmul:  inv(K_ref)[9] | (E_src @ inv(E_ref))[:3][12] | K_src[9] | inv(K_src)[9] |
 * (E_ref @ inv(E_src))[:3][12] | K_ref[9].  float64 arithmetic where the reference's NumPy is float64, cv2.remap's
 * fixed-point bilinear sampling (1/32 px, constant border 0) reproduced exactly.  Any output pointer may be NULL.
 * mvs_geo_consistency: one pair; depth_reproj is zeroed where the mask fails iff apply_mask (check_* vs reproject_*).
 * mvs_geo_fuse: all sources of one reference view in one pass: geo_sum int32 = number of consistent sources,
 * depth_avg float64 = (sum of masked reprojected depths + ref depth) / (geo_sum + 1), final_mask = (conf > conf_thresh)
 * & (geo_sum >= min_views); optional per-source masks [nsrc,H,W] u8 and masked reprojected depths [nsrc,H,W]. */
int mvs_geo_consistency(const float *depth_ref, const float *depth_src, const double *cam, uint8_t *mask,
                        float *depth_reproj, float *x_src, float *y_src, float *x_rep, float *y_rep, int H, int W,
                        double dist_thresh, float rel_thresh, int apply_mask, void *stream);
/* Back-projection of a fused depth map (float64 [H,W], as mvs_geo_fuse writes it) to world points, MVSNet/eval.py:297-300:
 * cam = inv(K_ref)[9] | inv(E_ref)[:3][12] (device, float64).  xyz float32 [H,W,3]; NaN where mask (u8, optional) is 0. */
int mvs_geo_backproject(const double *depth, const uint8_t *mask, const double *cam, float *xyz, int H, int W,
                        void *stream);
int mvs_geo_fuse(const float *depth_ref, const float *conf, const void *const *depth_srcs_host, int nsrc,
                 const double *cams, int32_t *geo_sum, double *depth_avg, uint8_t *final_mask, uint8_t *geo_masks,
                 float *depth_reproj, int H, int W, double dist_thresh, float rel_thresh, float conf_thresh,
                 int min_views, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MVS_B200_H */
