/* svr_b200 -- C ABI of the B200-native IF-Net implicit query path.
 *
 * Drop-in boundary for the hot path of nihalsid/single-view-3d-reconstruction:
 * model/projection.py (depth -> point cloud -> voxel occupancy) and model/ifnet.py (multi-scale
 * trilinear sampling + pointwise MLP decoder, forward and backward).  The reference has no FFI of
 * its own for this path (it is pure PyTorch); each entry point below names the reference
 * function (file:line, relative to the reference root) whose device work it replaces.  The
 * reference-side binding is the ctypes stub in single-view-3d-reconstruction_b200/_abi.py
 * (see INTEGRATION.md).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host; sizes are element counts;
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *  - return value: 0 = ok, <0 = invalid argument, >0 = cudaError_t of a failed launch; the
 *    message is available from svr_last_error() (thread-local);
 *  - there is NO CPU fallback: a missing device, a non-sm_100 device or a failed launch is an error;
 *  - bf16 buffers are passed as uint16_t*.
 */
#ifndef SVR_B200_H
#define SVR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVR_ABI_VERSION 1
#define SVR_MAX_LEVELS 6
#define SVR_MAX_TAPS 15

int svr_abi_version(void);
const char *svr_last_error(void);
/* 0 when the current device is compute capability 10.x; fills name (<=256 bytes) and SM count */
int svr_device_info(char *name_host, int *sm_count_host, int *cc_major_host, int *cc_minor_host);

/* ---------------------------------------------------------------------------------------------
 * Projection (model/projection.py)
 * ------------------------------------------------------------------------------------------- */

/* project.depth_to_camera (projection.py:201-206) + depthmap_to_gridspace (:150-163) [+ optional
 * norm_grid_space (:124-132)] fused: depth (B,H,W) fp32 -> pts (B,H*W,3) fp32.
 * cam = ((u*d - cx*d)/f, -((v*d - cy*d)/f), d);  g_k = fl(fl(scale[k]*cam_k) + offset[k]);
 * if normalise: g_k = fl(fl(g_k - dims[k]/2) / dims[k]).  Bit-exact to the reference CPU path.
 * grad version: d depth from d pts (same map, linear in d).                                     */
int svr_unproject_fwd(const float *depth, int B, int H, int W, float f, float cx, float cy,
                      const float *scale3_host, const float *offset3_host, const int64_t *dims3_host,
                      int normalise, float *pts, void *stream);
int svr_unproject_bwd(const float *grad_pts, int B, int H, int W, float f, float cx, float cy,
                      const float *scale3_host, const int64_t *dims3_host, int normalise,
                      float *grad_depth, void *stream);

/* project.norm_grid_space (projection.py:124-132), in place on pts (n_points,3).               */
int svr_norm_grid_space(float *pts, int64_t n_points, const int64_t *dims3_host, void *stream);

/* project.pc_voxels (projection.py:39-80): trilinear splat of B*N normalised points into a dense
 * (B,D0,D1,D2) fp32 grid, x8 sequential self-sum, clamp(0,1).  Bit-exact to the reference's
 * serial (deterministic-mode) accumulation order: pass-major (k,j,i), then point index.
 * No floating-point atomics: points are bucketed by floor cell with integer counters, every
 * touched voxel is summed by exactly one thread in reference order.
 *   tail_start : flat output index from which the 8-fold self-sum uses ATen's row_sum
 *                association ((2a+2a)+2a)+2a instead of the sequential one (see
 *                oracle/svr_oracle.c); pass B*D0*D1*D2 (or <0) for the canonical order.
 *   sat_mask   : optional (may be NULL) bitmask, 1 bit per output voxel (ceil(B*V/32) words), set
 *                where the pre-clamp value exceeds 1 (needed by svr_voxelize_bwd).
 *   workspace  : svr_voxelize_workspace_bytes(B,N,dims) bytes of scratch.                        */
size_t svr_voxelize_workspace_bytes(int B, int N, const int64_t *dims3_host);
int svr_voxelize_fwd(const float *pts, int B, int N, const int64_t *dims3_host, double eps,
                     int64_t tail_start, float *grid, uint32_t *sat_mask, void *workspace,
                     size_t workspace_bytes, void *stream);
/* backward of pc_voxels w.r.t. the points: d update = 8 * grad_grid * [pre-clamp <= 1], product
 * rule over the three axis weights, times (dims[k]-1).  grad_pts (B,N,3) is overwritten.        */
int svr_voxelize_bwd(const float *pts, const float *grad_grid, const uint32_t *sat_mask, int B, int N,
                     const int64_t *dims3_host, double eps, float *grad_pts, void *stream);

/* project.voxels_smooth (projection.py:102-117): three zero-padded 1-D cross-correlations (taps_w
 * on the last axis, taps_h on the middle one, taps_d on the first), then clamp(0,1).
 * The taps are DEVICE arrays (no host sync).  tmp0/tmp1: two scratch grids of the same size as
 * `in`.  Odd tap counts <= SVR_MAX_TAPS.  `out` must not alias `in`.                             */
int svr_blur_fwd(const float *in, int B, int D, int H, int W, const float *taps_w, int kw,
                 const float *taps_h, int kh, const float *taps_d, int kd, float *out,
                 float *tmp0, float *tmp1, void *stream);
/* backward: grad_in (same shape) and grad_taps (kw+kh+kd floats, device, [w|h|d]) from grad_out.
 * tmp: FOUR scratch grids.                                                                      */
int svr_blur_bwd(const float *in, const float *grad_out, int B, int D, int H, int W,
                 const float *taps_w, int kw, const float *taps_h, int kh,
                 const float *taps_d, int kd, float *grad_in, float *grad_taps, float *tmp,
                 void *stream);

/* ---------------------------------------------------------------------------------------------
 * IF-Net sampling + decoder (model/ifnet.py)
 * ------------------------------------------------------------------------------------------- */

/* Static description of the sampled feature pyramid (IFNetFeatureExtractor128 ifnet.py:122-199,
 * IFNetFeatureExtractor ifnet.py:64-120).  Level 0 is the input grid itself (C = 1, fp32); levels
 * >= 1 are packed channel-last bf16 volumes (svr_pack_volume).                                  */
typedef struct svr_pyramid {
    int n_levels;                     /* 6 (128-net) or 4 (32-net)                               */
    int channels[SVR_MAX_LEVELS];     /* level 0 must be 1; others multiples of 8                */
    int dims[SVR_MAX_LEVELS][3];      /* (D,H,W) of each level                                   */
    int align_corners;                /* 0: 128-net (ifnet.py:162), 1: 32-net (ifnet.py:98)      */
    float displacement;               /* 0.0722 (ifnet.py:144) / 0.035 (ifnet.py:82)             */
} svr_pyramid;

/* Padded K' (multiple of 64) of the permuted feature axis for a pyramid.                        */
int svr_feature_kp(const svr_pyramid *pyr_host);

/* fp32 volume with arbitrary element strides (B,C,D,H,W) -> bf16 NDHWC contiguous.              */
int svr_pack_volume(const float *src, int B, int C, int D, int H, int W, int64_t sB, int64_t sC,
                    int64_t sD, int64_t sH, int64_t sW, uint16_t *dst, void *stream);
/* same source -> bf16 (B, D+2, H+2, W+2, C) with a one-voxel zero halo (C % 8 == 0): the zero padding of
 * F.grid_sample (ifnet.py:162-193, padding_mode='zeros') materialised, for the fused kernel's wide path */
int svr_pack_volume_halo(const float *src, int B, int C, int D, int H, int W, int64_t sB, int64_t sC,
                         int64_t sD, int64_t sH, int64_t sW, uint16_t *dst, void *stream);
/* fp32 NDHWC contiguous gradient -> accumulate (+=) into fp32 tensor with arbitrary strides.    */
int svr_unpack_volume_grad(const float *src_ndhwc, int B, int C, int D, int H, int W, int64_t sB,
                           int64_t sC, int64_t sD, int64_t sH, int64_t sW, float *dst, int accumulate,
                           void *stream);

/* fc_0 weight (H0, 7*sum C) fp32 with k = c*7+d (ifnet.py:43-45) -> bf16 (H0, KP) in the
 * permuted/padded K' order used by the kernels, and its transpose (KP, H0).  Either output may
 * be NULL.  svr_unpack_w0_grad maps a (H0,KP) fp32 gradient back to (H0, 7*sum C).              */
int svr_pack_w0(const float *w0, int H0, const svr_pyramid *pyr_host, uint16_t *w0p, uint16_t *w0pT,
                void *stream);
int svr_unpack_w0_grad(const float *gw0p, int H0, const svr_pyramid *pyr_host, float *gw0, void *stream);
/* generic fp32 (R,C) -> bf16 (R,C) and/or its transpose (C,R)                                   */
int svr_pack_matrix(const float *w, int R, int C, uint16_t *wp, uint16_t *wpT, void *stream);

/* Multi-scale trilinear stencil gather (ifnet.py:156-197 == 6x F.grid_sample + cat, fused):
 * points (B,N,3) fp32, x0 (B,D,H,W) fp32, vols[l] bf16 NDHWC for l>=1  ->  feat (B*N, KP) bf16 in
 * K' order.                                                                                     */
int svr_gather_fwd(const float *points, int B, int N, const float *x0, const uint16_t *const *vols_host,
                   const svr_pyramid *pyr_host, uint16_t *feat, void *stream);
/* Backward of the gather: dfeat (B*N, KP) bf16 -> scatter-add into gvols[l] (fp32 NDHWC, l>=1),
 * gx0 (fp32, may be NULL) and gpoints (B,N,3; may be NULL, needs x0/vols).  Outputs are
 * accumulated into (caller zero-fills).  perm (optional): row r of dfeat belongs to point perm[r];
 * cell_start (optional, with perm): the row ranges of the sort cells written by svr_sort_points -- the
 * tensor-core scatter of the coarse levels then cuts its row tiles at the boundaries of 2x2x2 cell
 * groups, which bounds the voxel box a tile touches. */
int svr_gather_bwd(const float *points, const int *perm, const int *cell_start, int B, int N, const float *x0,
                   const uint16_t *const *vols_host, const svr_pyramid *pyr_host, const uint16_t *dfeat,
                   float *gx0, float *const *gvols_host, float *gpoints, void *stream);

/* First encoder layer fused with its ReLU: relu(Conv3d(1 -> Co, 3x3x3, padding 1)(x)) (ifnet.py:126,164;
 * 32-net :68,100).  x (B,D,H,W) fp32, w (Co,27) = conv.weight, y (B,D,H,W,Co) fp32 NDHWC.  Co in {16,32}.
 * Backward: gw (Co,27), gb (Co), optional gx (B,D,H,W) from gy (NDHWC) and the saved output y.      */
int svr_conv1_relu_fwd(const float *x, const float *w, const float *bias, int B, int D, int H, int W, int Co, float *y, void *stream);
size_t svr_conv1_relu_bwd_workspace_bytes(int Co);
int svr_conv1_relu_bwd(const float *x, const float *y, const float *gy, const float *w, int B, int D, int H, int W, int Co,
                       float *gw, float *gb, float *gx, void *workspace, size_t workspace_bytes, void *stream);

/* nn.MaxPool3d(2) of the encoder (ifnet.py:133,169-190) on channels-last (NDHWC) fp32 activations,
 * forward (+ packed per-channel argmax, one byte per output element, 4 per uint32) and backward.
 * Keeps the torch/cuDNN encoder channels-last end to end (torch's max_pool3d would make an NCDHW
 * copy).  torch semantics: floor output size, first maximum wins, NaN propagates.  C % 4 == 0.  */
int svr_maxpool2_cl_fwd(const float *in, int B, int D, int H, int W, int C, float *out, uint32_t *idx, void *stream);
int svr_maxpool2_cl_bwd(const float *gout, const uint32_t *idx, int B, int D, int H, int W, int C, float *gin, void *stream);

/* Processing order for the query kernels: perm (B*N ints) lists point indices sorted by (scene,
 * Morton code of a 16^3 cell), so that consecutive rows are spatial neighbours (cache locality of
 * the gather; the reference has no counterpart -- every row is independent, results do not change).*/
size_t svr_sort_points_workspace_bytes(int B, int N);
int svr_sort_cells_per_scene(void);
/* cell_start (optional): B * svr_sort_cells_per_scene() + 1 ints, first sorted row of every (scene, cell)
 * in sort order, then B*N. */
int svr_sort_points(const float *points, int B, int N, int *perm, int *cell_start, void *workspace, size_t workspace_bytes,
                    void *stream);

/* tcgen05 GEMMs (Conv1d k=1 of ifnet.py:55-59 and their backward).
 * NT:  C[M,N] = epi( A[M,K] . B[N,K]^T + bias[N] ),  A,B bf16 row-major, K % 64 == 0.
 *   flags: bit0 relu, bit1 store bf16 to c_bf16, bit2 store fp32 to c_f32,
 *          bit3 multiply by (mask[M,N] > 0) (bf16 mask, ld = ldc) -- relu backward,
 *          bit4 row-dot: out_dot[m] = sum_n epi(..)[m,n]*dot_w[n] + dot_b[0] (fc_out fused; N<=256)
 *          bit5 accumulate: the product is added to the fp32 values already in c_f32 before bias/relu
 *               (running sum of the hi/lo split products of the fp32 tier, see "fp32-accurate tier" below)
 * TN:  C[M,N] (+)= A[P,M]^T . B[P,N], A,B bf16 row-major (contraction over rows), fp32 out;
 *      workspace >= svr_gemm_tn_workspace_bytes(M,N,P).                                          */
int svr_gemm_nt(const uint16_t *A, int64_t lda, const uint16_t *B, int64_t ldb, const float *bias, int M, int N,
                int K, int flags, uint16_t *c_bf16, float *c_f32, int64_t ldc, const uint16_t *mask,
                const float *dot_w, const float *dot_b, float *out_dot, void *stream);
size_t svr_gemm_tn_workspace_bytes(int M, int N, int P);
int svr_gemm_tn(const uint16_t *A, int64_t lda, const uint16_t *B, int64_t ldb, int M, int N, int P, float *C,
                int64_t ldc, int accumulate, void *workspace, size_t workspace_bytes, void *stream);

/* Small fused helpers of the decoder backward.
 * dz2[m,n] = dlogit[m] * wout[n] * (h2[m,n] > 0)  (bf16 out); gwout[n] += sum_m dlogit[m]*h2[m,n];
 * gbout += sum_m dlogit[m].  perm (optional): row m reads dlogit[perm[m]].                        */
int svr_decoder_head_bwd(const float *dlogit, const int *perm, const uint16_t *h2, const float *wout, int M,
                         int Hd, uint16_t *dz2, float *gwout, float *gbout, void *stream);
/* column sums of a bf16 (M,N) matrix into fp32 out[N] (bias gradients); accumulate != 0 adds.    */
int svr_colsum_bf16(const uint16_t *a, int M, int N, int64_t lda, float *out, int accumulate, void *stream);

/* Fused forward: gather -> smem -> tcgen05 fc_0 -> fc_1 -> fc_2 -> fc_out in ONE persistent kernel
 * (128-net and any pyramid whose decoder has hidden size 256).  Weights are passed as pre-swizzled
 * UMMA chunk images (svr_pack_decoder_image of the bf16 row-major matrices: fc_0 in K' order).
 * logits (B*N) fp32, indexed like `points`.  perm (optional, B*N ints): processing order, row r of
 * the kernel handles point perm[r] (spatially sorted points make the gather cache-friendly);
 * save_h (optional): post-ReLU hidden activations bf16 (3, B*N, 256) in ROW order; save_feat
 * (optional): gathered features bf16 (B*N, KP) in ROW order -- both feed the backward.           */
typedef struct svr_decoder_weights {
    const void *w0p;       /* image of (256, KP) bf16, K' order                                   */
    const void *w1;        /* image of (256, 256) bf16                                            */
    const void *w2;        /* image of (256, 256) bf16                                            */
    const float *b0, *b1, *b2;
    const float *wout;     /* (256) fp32                                                          */
    const float *bout;     /* (1) fp32, device                                                    */
    int h0, h1, h2;
} svr_decoder_weights;

/* bf16 row-major (R, K) -> K/64 chunks of (R x 128 B) in the 128B-swizzled K-major UMMA layout   */
int svr_pack_decoder_image(const uint16_t *w_rowmajor, int R, int K, uint8_t *image, void *stream);

/* debug only: SM-clock timeline of block 0 (4 roles x 1024 x (tag, clock) int64, device buffer; null = off) */
int svr_debug_fq_trace(void *buf);

/* halo_vols_host (nullable table, nullable entries): svr_pack_volume_halo copies of the levels with
 * C % 64 == 0; the trailing run of such levels is sampled through the kernel's bounds-check-free wide path.
 * cell_start (optional, with perm): the row ranges of the sort cells written by svr_sort_points; used by the box
 * kernel (svr_debug_fq_interp(2)), which cuts row tiles at the boundaries of sort-cell groups.                      */
int svr_query_fwd_fused(const float *points, const int *perm, const int *cell_start, int B, int N, const float *x0,
                        const uint16_t *const *vols_host, const uint16_t *const *halo_vols_host,
                        const svr_pyramid *pyr_host, const svr_decoder_weights *w_host, float *logits,
                        uint16_t *save_h, uint16_t *save_feat, int apply_sigmoid, void *stream);
/* Kernel choice.  0: every level gathered on the CUDA cores.  1 (default): svr_dense_eval runs the box kernel -- the
 * levels whose voxel box per 128-point brick is small (32^3 / 16^3 / 8^3 of the 128-net) are interpolated on the tensor
 * cores from a box staged in shared memory (bf16 trilinear weights) instead of gathered corner by corner.  2: explicit
 * points with a sort-cell table take the box kernel too (slower than the gather kernel at 50k points per scene: the
 * tiles cut at cell-group boundaries are 77 % full).  3: like 1 with the voxel boxes staged by cp.async instead of TMA
 * tensor copies (ablation); 4: like 2 with cp.async staging.                                                 */
int svr_debug_fq_interp(int mode);
/* debug: the block whose timeline svr_debug_fq_trace records (default 0) */
int svr_debug_fq_trace_block(int block);

/* First stage of the 128-net fused: y = BatchNorm3d(relu(Conv3d(1 -> 16, 3, padding 1)(x))), channels-last y
 * (model/ifnet.py:126,137,164 `net = self.actvn(self.conv_in(x)); net = self.conv_in_bn(net)`).  The pre-BN
 * activation is never stored: every pass recomputes it from the one-channel input.
 *   stats : training-mode batch mean / 1/sqrt(var+eps) (biased variance) into mean[16], invstd[16]; running_mean /
 *           running_var (nullable) get torch's momentum update (unbiased variance).
 *   apply : y (B,D,H,W,16) from given mean / invstd (batch statistics, or running statistics in eval mode); y_bf16
 *           (nullable) receives the bf16 copy the gather kernels sample (what svr_pack_volume would produce),
 *           relu_mask (nullable, one uint16 per voxel) the bits [relu(conv)_c > 0].
 *   bwd   : gw (16,27), gb (16), ggamma (16), gbeta (16), training-mode BN backward.  The gradient of y is
 *           gy (B,D,H,W,16, nullable) plus, when the stage's nn.MaxPool3d(2) (ifnet.py:169) is fused, the gradient
 *           g_pooled (B,D/2,H/2,W/2,16) of the pooled tensor routed through pool_idx (winner codes written by
 *           svr_maxpool2_cl_fwd; both nullable): neither the pooling backward nor the sum is materialised.
 *           With y and relu_mask (both nullable, as written by apply; beta then required when BN has one) the
 *           normalised activation is read back instead of recomputed.
 * x (B,D,H,W) fp32, w (16,27), bias / gamma / beta nullable.  Workspace: svr_conv1_bn_workspace_bytes().       */
size_t svr_conv1_bn_workspace_bytes(void);
int svr_conv1_relu_bn_stats(const float *x, const float *w, const float *bias, int B, int D, int H, int W, int Co, float eps,
                            float momentum, float *running_mean, float *running_var, float *mean, float *invstd, void *workspace,
                            size_t workspace_bytes, void *stream);
int svr_conv1_relu_bn_apply(const float *x, const float *w, const float *bias, const float *mean, const float *invstd,
                            const float *gamma, const float *beta, int B, int D, int H, int W, int Co, float *y, uint16_t *y_bf16,
                            uint16_t *relu_mask, void *stream);
int svr_conv1_relu_bn_bwd(const float *x, const float *w, const float *bias, const float *mean, const float *invstd,
                          const float *gamma, const float *beta, const float *y, const uint16_t *relu_mask, const float *gy,
                          const float *g_pooled, const uint32_t *pool_idx, int B, int D, int H, int W, int Co, float *gw,
                          float *gb, float *ggamma, float *gbeta, void *workspace, size_t workspace_bytes, void *stream);

/* Elementwise glue of the encoder's Conv3d -> ReLU layers (ifnet.py:127-135,165-183 `self.actvn(self.conv_x(net))`),
 * channels-last fp32 activations viewed as (rows = B*D*H*W, C), C % 4 == 0 and 256 % (C/4) == 0.
 *   svr_bias_relu_cl : y <- max(y + bias, 0) in place (bias nullable).
 *   svr_relu_bwd_cl  : g = gy * [y > 0] as fp32 (g_f32, nullable) and/or bf16 (g_bf16, nullable); gbias[c] = sum_rows g
 *                      (nullable) from the same pass.                                                             */
int svr_bias_relu_cl(float *y, const float *bias, int64_t rows, int C, void *stream);
/* dst[i] = (float) src[i] for a dense bf16 buffer of n elements (n % 8 == 0, 16-byte aligned): widens cuDNN's bf16
 * backward-data result for the fp32 gradient chain without torch's strided element-wise copy.                     */
int svr_widen_bf16(const uint16_t *src, int64_t n, float *dst, void *stream);
size_t svr_relu_bwd_cl_workspace_bytes(int C);
int svr_relu_bwd_cl(const float *gy, const float *y, int64_t rows, int C, float *g_f32, uint16_t *g_bf16, float *gbias,
                    void *workspace, size_t workspace_bytes, void *stream);

/* Fused decoder backward-data chain (Conv1d backward of ifnet.py:55-58, hidden size 256):
 *   dz1 = (dz2 . W2) * [h1 > 0],  dz0 = (dz1 . W1) * [h0 > 0],  dfeat = dz0 . W0'   (all bf16, row-major)
 * in one persistent tcgen05 kernel.  Weights are pre-swizzled chunk images (svr_pack_decoder_image) of
 * W2^T (256,256), W1^T (256,256) and W0'^T (KP,256).                                              */
int svr_decoder_bwd_fused(const uint16_t *dz2, const uint16_t *h1, const uint16_t *h0, const void *w2t_img,
                          const void *w1t_img, const void *w0pt_img, int64_t M, int kp, uint16_t *dz1, uint16_t *dz0,
                          uint16_t *dfeat, void *stream);

/* Debug: device buffer of 3*1024*2 int64 (tag, SM clock) records that block 0 of the following
 * svr_decoder_bwd_fused launches fills (row worker / MMA / loader roles); NULL switches tracing off.  */
int svr_debug_fb_trace(void *buf);

/* Dense evaluation (evaluate_network_on_grid, ifnet.py:215-229; make_3d_grid :202-212): evaluates
 * sigmoid(decoder(sample(x, lattice))) on the (sx,sy,sz) inclusive lattice over [-0.5,0.5]^3 for
 * z-slab [x_begin, x_end) of the FIRST lattice axis, generating the points on the fly; out is the
 * (sx,sy,sz) fp32 grid of ONE scene (only the slab is written).                                  */
int svr_dense_eval(int scene, int B, const float *x0, const uint16_t *const *vols_host,
                   const uint16_t *const *halo_vols_host, const svr_pyramid *pyr_host,
                   const svr_decoder_weights *w_host, int sx, int sy, int sz, int x_begin, int x_end, float *out,
                   void *stream);

/* ---------------------------------------------------------------------------------------------
 * fp32-accurate tier (`configure(precision=32)`): the reference's arithmetic for this path is fp32
 * (model/ifnet.py:38-61 Conv1d in fp32, :155-199 F.grid_sample on fp32 volumes).  Volumes, features, hidden
 * activations and gradients stay fp32 in HBM; the contractions run on the tensor cores as three bf16 GEMM passes
 * over hi/lo splits of both operands (A_hi.B_hi + A_lo.B_hi + A_hi.B_lo, fp32 accumulation; svr_gemm_nt bit5,
 * svr_gemm_tn accumulate).  csrc/precise.cu.
 * ------------------------------------------------------------------------------------------- */
/* x (n fp32, n % 8 == 0) -> hi = bf16(x), lo = bf16(x - hi)                                       */
int svr_split_bf16(const float *x, int64_t n, uint16_t *hi, uint16_t *lo, void *stream);
/* fc_0.weight (H0, C*7) in the reference's k = c*7+d order (ifnet.py:43-45) -> (H0, KP) fp32 in kernel order */
int svr_pack_w0_f32(const float *w0, int H0, const svr_pyramid *pyr, float *w0p, void *stream);
/* IFNetFeatureExtractor*.forward sampling (ifnet.py:156-197 / :93-118) on fp32 NDHWC volumes -> (B*N, KP) fp32 */
int svr_gather_fwd_f32(const float *points, int B, int N, const float *x0, const float *const *vols_host,
                       const svr_pyramid *pyr, float *feat, void *stream);
/* its backward (ATen grid_sampler_3d_backward semantics) with fp32 d-features and fp32 trilinear weights;
 * gx0 / gvols_host[l] / gpoints may be null (not needed)                                          */
int svr_gather_bwd_f32(const float *points, int B, int N, const float *x0, const float *const *vols_host,
                       const svr_pyramid *pyr, const float *dfeat, float *gx0, float *const *gvols_host,
                       float *gpoints, void *stream);
/* fc_out + relu(fc_2) backward in fp32: dz2 = dlogit (x) wout * [h2 > 0], gwout = sum dlogit*h2, gbout = sum dlogit */
int svr_decoder_head_bwd_f32(const float *dlogit, const float *h2, const float *wout, int M, int Hd, float *dz2,
                             float *gwout, float *gbout, void *stream);
/* column sums of an fp32 (M, N) matrix (bias gradients), deterministic                            */
int svr_colsum_f32(const float *a, int M, int N, int64_t lda, float *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SVR_B200_H */
