/*
 * b200unet.h — C ABI of libb200unet.so: the sm_100a (B200) kernels behind the U-Net hot path of
 * a-green-hand-jack/mdeical_image_segmentation (2D/3D encoder-decoder forward + backward, Dice/IoU
 * losses and metrics).
 *
 * The reference has no FFI of its own: every entry point below replaces a stock-PyTorch call site
 * of the reference (cited per function as file:line relative to the reference root). Host code
 * (mdeical_image_segmentation_b200/_lib.py) binds these with ctypes; see INTEGRATION.md.
 *
 * Conventions
 *   - activations are channels-last bf16: (N, D, H, W, C) with D = 1 for 2D; a tensor may be a
 *     channel slice of a wider buffer, described by (ctot = channel stride, coff = first channel);
 *   - parameters and their gradients are fp32 in the reference's own layouts (OIDHW, IODHW);
 *   - every buffer is caller-owned device memory; nothing is allocated, freed or retained;
 *   - all work is enqueued on the caller's stream, asynchronously, with no device sync;
 *   - return 0 on success, negative b200 error codes or a positive cudaError_t; no exceptions.
 */
#ifndef B200UNET_H
#define B200UNET_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* b200_stream_t; /* cudaStream_t */

#define B200_IMPL_AUTO 0
#define B200_IMPL_TCGEN05 1
#define B200_IMPL_DIRECT 2

int b200_version(void);
const char* b200_strerror(int code);
/* last tcgen05/direct decision taken by a conv-like entry point on this thread (1 or 2) */
int b200_last_impl(void);
/* tuning / A-B knobs (process-wide): key 0 = halo-reuse fwd/dgrad kernel,
 * key 1 = halo-reuse weight-gradient kernel (value 1 = on (default), 0 = off),
 * key 2 = max row-blocks per super-tile of the halo kernel (1, 2 or 4; default 4),
 * key 3 = weight-resident mode of the forward / dgrad kernels (1 = on (default), 0 = off),
 * key 4 = epilogue bring-up mode (0 = normal; 1 = no global stores, 2 = no epilogue work: timing only),
 * key 5 = epilogue warps of the forward / dgrad kernels (0 = heuristic (default), 4 or 8) */
int b200_set_tuning(int key, int value);

/* ---- 3x3 / 3x3x3 / 1x1 "same" convolution ------------------------------------------------------
 * replaces nn.Conv2d  model/unet2d/layers.py:20,122,125 and nn.Conv3d/2d
 * model/unet3d/buildingblocks.py:63-70 (and their autograd backward). */
typedef struct b200_conv_desc {
  int32_t n, d, h, w;          /* positions; d = 1 for 2D                                  */
  int32_t c_in, c_out;
  int32_t kd, kh, kw;          /* each 1 or 3; zero padding k/2, stride 1, dilation 1      */
  int32_t in_ctot, in_coff;    /* layout of x / dx                                        */
  int32_t out_ctot, out_coff;  /* layout of y / dy                                        */
  int32_t relu;                /* fwd epilogue: y = max(conv + bias, 0)                   */
  int32_t impl;                /* B200_IMPL_*                                             */
} b200_conv_desc;

/* w: fp32 (c_out, c_in, kd, kh, kw)  ->  wp_fwd bf16 [tap][c_out][c_in],
 * wp_bwd bf16 [flipped tap][c_in][c_out]  (either output may be NULL). */
int b200_conv_repack_weight(const b200_conv_desc* desc, const float* w, void* wp_fwd, void* wp_bwd,
                            b200_stream_t stream);
/* The same re-pack for ALL filters of a network in one launch (the fp32 masters change every optimizer
 * step).  c_in_pad >= c_in: the packs are written for c_in_pad input channels, the extra ones zero (narrow
 * network inputs are stored zero-padded to 16 channels).  Either pack pointer may be NULL. */
typedef struct b200_repack_item {
  const float* w;              /* fp32 (c_out, c_in, ntaps)                                */
  void* wp_fwd;                /* bf16 [tap][c_out][c_in_pad]                              */
  void* wp_bwd;                /* bf16 [ntaps-1-tap][c_in_pad][c_out]                      */
  int32_t c_out, c_in, c_in_pad, ntaps;
} b200_repack_item;
int b200_conv_repack_weights(int32_t count, const b200_repack_item* items, b200_stream_t stream);
int b200_conv_fwd(const b200_conv_desc* desc, const void* x, const void* wp_fwd, const float* bias,
                  void* y, b200_stream_t stream);
/* dx = conv_transpose(dy); if relu_mask != NULL (same layout as dx) dx is zeroed where mask <= 0 */
int b200_conv_bwd_data(const b200_conv_desc* desc, const void* dy, const void* wp_bwd,
                       const void* relu_mask, void* dx, b200_stream_t stream);
/* Fused column statistics.  The *_stats variants do the same work and additionally accumulate, in the
 * GEMM epilogue, per-sample per-channel sums of the STORED (bf16) result into colstats fp64
 * [n][2][channels] (caller zeroes): [.][0][c] += sum_pos r, [.][1][c] += sum_pos r*r — or
 * sum_pos r*dot when a dot operand (bf16, positions of the result, channel layout dot_ctot/dot_coff)
 * is given.  They replace the separate passes of the op that follows the convolution:
 * BatchNorm / GroupNorm statistics (buildingblocks.py:73-103) in forward; the conv bias gradient and
 * the (sum dy, sum dy*x) reductions of the norm backward in backward.  colstats == NULL: plain call. */
int b200_conv_fwd_stats(const b200_conv_desc* desc, const void* x, const void* wp_fwd,
                        const float* bias, void* y, double* colstats, b200_stream_t stream);
#define B200_STATS_NONE 0
#define B200_STATS_SUM 1      /* only [.][0][c] (e.g. a bias gradient); [.][1][c] is left untouched */
#define B200_STATS_SUM_SQ 2   /* sum r, sum r*r                                                    */
#define B200_STATS_SUM_DOT 3  /* sum r, sum r*dot                                                  */
int b200_conv_bwd_data_stats(const b200_conv_desc* desc, const void* dy, const void* wp_bwd,
                             const void* relu_mask, void* dx, int32_t stats_kind, const void* dot,
                             int32_t dot_ctot, int32_t dot_coff, double* colstats,
                             b200_stream_t stream);
/* dw fp32 (c_out, c_in, kd, kh, kw) is ACCUMULATED into (caller zeroes it) */
int b200_conv_bwd_weight(const b200_conv_desc* desc, const void* x, const void* dy, float* dw,
                         b200_stream_t stream);
/* same, plus the conv bias gradient dbias[c_out] += sum_pos dy (fp32, ACCUMULATED; NULL = skip): the
 * second half of aten::convolution_backward for nn.Conv2d(bias=True) layers.py:122-125, fused into
 * the weight-gradient kernel (the dy tiles are already in shared memory) */
int b200_conv_bwd_weight_bias(const b200_conv_desc* desc, const void* x, const void* dy, float* dw,
                              float* dbias, b200_stream_t stream);

/* ---- ConvTranspose2d(kernel 2, stride 2) -------------------------------------------------------
 * replaces nn.ConvTranspose2d  model/unet2d/layers.py:165. x: (n,1,h,w,c_in) -> y: (n,1,2h,2w,c_out) */
typedef struct b200_convt_desc {
  int32_t n, h, w;
  int32_t c_in, c_out;
  int32_t in_ctot, in_coff;
  int32_t out_ctot, out_coff;
  int32_t impl;
} b200_convt_desc;
/* w fp32 (c_in, c_out, 2, 2) -> wp_fwd bf16 [4*c_out][c_in] (row = (i*2+j)*c_out + co),
 * wp_bwd bf16 [4][c_in][c_out] */
int b200_convt2x2_repack_weight(const b200_convt_desc* desc, const float* w, void* wp_fwd,
                                void* wp_bwd, b200_stream_t stream);
int b200_convt2x2_fwd(const b200_convt_desc* desc, const void* x, const void* wp_fwd,
                      const float* bias, void* y, b200_stream_t stream);
int b200_convt2x2_bwd_data(const b200_convt_desc* desc, const void* dy, const void* wp_bwd,
                           const void* relu_mask, void* dx, b200_stream_t stream);
/* colstats fp64 [n][2][c_in] += column statistics of dx (B200_STATS_SUM or _SUM_SQ), see
 * b200_conv_bwd_data_stats */
int b200_convt2x2_bwd_data_stats(const b200_convt_desc* desc, const void* dy, const void* wp_bwd,
                                 const void* relu_mask, void* dx, int32_t stats_kind, double* colstats,
                                 b200_stream_t stream);
int b200_convt2x2_bwd_weight(const b200_convt_desc* desc, const void* x, const void* dy, float* dw,
                             b200_stream_t stream);

/* ---- ConvTranspose(kernel 3, stride 2, padding 1, no bias) -------------------------------------
 * replaces the conv_transposed of TransposeConvUpsampling  model/unet3d/buildingblocks.py:699-717.
 * The (2d-1, 2h-1, 2w-1) result is kept as PARITY PLANES: a dense bf16 buffer
 * (classes*n, d, h, w, c_out), classes = 8 (kd = 3) or 4 (kd = 1, 2D), where plane (pd,ph,pw) holds
 * ct[2m + p] at position m.  Each plane is a small tap-shift implicit GEMM on the tensor cores
 * (1/2/4/8 taps); backward-data and backward-weight are 27-tap (9-tap) GEMMs that read the planes.
 * The F.interpolate(size=skip) that follows (:703) is b200_deconv_planes_resize_*. */
typedef struct b200_deconv_desc {
  int32_t n, d, h, w;          /* input positions (d = 1 for 2D)      */
  int32_t c_in, c_out;
  int32_t kd;                  /* 3 (ConvTranspose3d) or 1 (2D)       */
  int32_t in_ctot, in_coff;    /* layout of x / dx                    */
  int32_t impl;
} b200_deconv_desc;
/* w fp32 (c_in, c_out, kd, 3, 3) -> wp_fwd bf16 [tap][c_out][c_in], wp_bwd bf16 [tap][c_in][c_out] */
int b200_deconv3s2_repack_weight(const b200_deconv_desc* desc, const float* w, void* wp_fwd,
                                 void* wp_bwd, b200_stream_t stream);
int b200_deconv3s2_fwd(const b200_deconv_desc* desc, const void* x, const void* wp_fwd, void* yp,
                       b200_stream_t stream);
int b200_deconv3s2_bwd_data(const b200_deconv_desc* desc, const void* dyp, const void* wp_bwd,
                            void* dx, b200_stream_t stream);
/* dw fp32 (c_in, c_out, kd, 3, 3) accumulated into */
int b200_deconv3s2_bwd_weight(const b200_deconv_desc* desc, const void* x, const void* dyp,
                              float* dw, b200_stream_t stream);

/* ---- memory-bound tensor ops -------------------------------------------------------------------
 * A "view" is a channels-last bf16 tensor slice. */
typedef struct b200_view {
  void* ptr;      /* base of the wider buffer */
  int32_t ctot;   /* channel stride           */
  int32_t coff;   /* first channel            */
} b200_view;

/* MaxPool(kernel 2, stride 2, floor) — nn.MaxPool2d layers.py:147; nn.MaxPool3d/2d
 * buildingblocks.py:409-413. pool_d = 1 keeps D (2D pooling). */
int b200_maxpool2_fwd(int32_t n, int32_t d, int32_t h, int32_t w, int32_t c, int32_t pool_d,
                      b200_view x, b200_view y, b200_stream_t stream);
/* the same pooling, which also accumulates the per-sample per-channel (sum y, sum y^2) of its output into
 * sums fp64 [n][2][c] (zeroed by the caller) — the statistics of the GroupNorm that opens the next encoder level
 * (buildingblocks.py:73-103, 409-425).  Even extents and 8-channel aligned views only (B200_ERR_UNSUPPORTED else). */
int b200_maxpool2_fwd_sums(int32_t n, int32_t d, int32_t h, int32_t w, int32_t c, int32_t pool_d, b200_view x,
                           b200_view y, double* sums, b200_stream_t stream);
/* dx fully written; ties go to the first maximum in (d,h,w) scan order, as ATen does.
 * add (ptr may be NULL): a second gradient of x's shape summed in (the skip branch of the U).
 * mask_relu != 0: x is a ReLU output; dx is zeroed where x <= 0 (its producer's ReLU backward). */
int b200_maxpool2_bwd(int32_t n, int32_t d, int32_t h, int32_t w, int32_t c, int32_t pool_d,
                      b200_view x, b200_view dy, b200_view add, b200_view dx, int32_t mask_relu,
                      b200_stream_t stream);
/* ---- UNet 3+ (model/unet2d/unet.py:136-446), 2D only -------------------------------------------
 * nn.MaxPool2d(k, k, ceil_mode=True) unet.py:175-187 on sizes divisible by k (the network only accepts
 * multiples of 16, where ceil mode = floor mode).  bwd: ties to the first maximum in scan order; dx fully
 * written; add (ptr may be NULL) = a gradient of x's shape summed in. */
int b200_maxpoolk_fwd(int32_t n, int32_t h, int32_t w, int32_t c, int32_t k, b200_view x, b200_view y,
                      b200_stream_t stream);
int b200_maxpoolk_bwd(int32_t n, int32_t h, int32_t w, int32_t c, int32_t k, b200_view x, b200_view dy,
                      b200_view add, b200_view dx, b200_stream_t stream);
/* nn.Upsample(scale_factor=scale, mode='bilinear') unet.py:198-315 (align_corners=False): (n, ih, iw, c)
 * -> (n, ih*scale, iw*scale, c); bwd is its exact adjoint, separable (along W into scratch, then along H),
 * deterministic, + add (ptr may be NULL).  scratch: n * ih*scale * iw * c floats, 16-byte aligned, caller-owned. */
int b200_bilinear2d_fwd(int32_t n, int32_t ih, int32_t iw, int32_t c, int32_t scale, b200_view x,
                        b200_view y, b200_stream_t stream);
int b200_bilinear2d_bwd(int32_t n, int32_t ih, int32_t iw, int32_t c, int32_t scale, b200_view dy,
                        b200_view add, b200_view dx, void* scratch, b200_stream_t stream);
/* ---- GPU-side volume augmentation ------------------------------------
 * replaces the numpy / scipy bodies of augment/unet3d_augment/transforms.py on dense fp32 (C, D, H, W)
 * device tensors; the random decisions stay with the caller's np.random.RandomState.
 * flip_rot90: np.flip along D / H / W (flip_bits 1 / 2 / 4; RandomFlip :25-50) followed by
 *   np.rot90(m, rot_k, axes=(1, 2)) (RandomRotate90 :53-80); dst is (C, D, W, H) for odd rot_k. */
int b200_aug_flip_rot90(int32_t c, int32_t d, int32_t h, int32_t w, int32_t flip_bits, int32_t rot_k,
                        const float* src, float* dst, b200_stream_t stream);
/* dst = clip(a[ch] * src + b[ch], lo, hi) (+-inf = no clip): RandomContrast :115-135, Standardize :495-523,
 * Normalize :547-605 once their scalars are folded; in place allowed. */
int b200_aug_affine_clip(int32_t c, int64_t per_channel, const float* src, const float* a, const float* b,
                         float lo, float hi, float* dst, b200_stream_t stream);
/* stats fp64 [c][4] accumulated into: sum, sum of squares, min, max (caller initialises 0, 0, +inf, -inf):
 * np.mean / np.std / np.min / np.max of Standardize and Normalize. */
int b200_aug_channel_stats(int32_t c, int64_t per_channel, const float* src, double* stats, b200_stream_t stream);
/* one axis of scipy.ndimage.gaussian_filter(x, sigma, mode='reflect') (truncate 4.0) over an (n0, n1, n2) field:
 * ElasticDeformation's displacement smoothing :165-175. */
int b200_aug_gaussian_axis(int32_t n0, int32_t n1, int32_t n2, int32_t axis, float sigma, const float* src,
                           float* dst, b200_stream_t stream);
/* scipy.ndimage.map_coordinates(src[ch], (z + dz, y + dy, x + dx), order = 0 | 1, mode='reflect') for every
 * channel (:178-190); dz may be NULL (apply_3d = False). */
int b200_aug_map_coordinates(int32_t c, int32_t d, int32_t h, int32_t w, const float* src, const float* dz,
                             const float* dy, const float* dx, int32_t order, float* dst, b200_stream_t stream);

/* ---- layer_order letters 'l' / 'e' / 'd' / 'D' and avg-pool encoders ----------------------------
 * replaces nn.LeakyReLU / nn.ELU (buildingblocks.py:55-59; ResNetBlock's non_linearity :306-311, applied to
 * out + residual :319-323), nn.Dropout / nn.Dropout2d (:104-107) and nn.AvgPool2d/3d (:414-418).
 * y = act(a [+ b]) (b.ptr may be NULL); param = negative slope (LEAKY_RELU) or alpha (ELU). */
#define B200_ACT_RELU 0
#define B200_ACT_LEAKY_RELU 1
#define B200_ACT_ELU 2
#define B200_ACT_IDENTITY 3 /* y = a + b: the channel-count-agnostic form of b200_add */
int b200_act_fwd(int64_t npos, int32_t c, int32_t kind, float param, b200_view a, b200_view b, b200_view y,
                 b200_stream_t stream);
/* dx = dy * act'(.) with the derivative recovered from the stored OUTPUT y (y > 0 ? 1 : 0 | slope | y + alpha) */
int b200_act_bwd(int64_t npos, int32_t c, int32_t kind, float param, b200_view dy, b200_view y, b200_view dx,
                 b200_stream_t stream);
/* y = x * scale * (keep ? keep[pos][c] != 0 : 1) * (chan ? chan[n][c] : 1): dropout forward and backward.
 * keep: dense uint8 [n * spatial][c] (element-wise dropout) or NULL; chan: fp32 [n][c] (channel-wise dropout,
 * 0 or 1) or NULL.  The mask itself comes from the caller's RNG. */
int b200_scale_mask(int32_t n, int64_t spatial, int32_t c, b200_view x, const uint8_t* keep, const float* chan,
                    float scale, b200_view y, b200_stream_t stream);
/* 2x2 (pool_d = 1) / 2x2x2 (pool_d = 2) average pool, stride 2, floor mode; bwd: dx = dy / window (+ add,
 * ptr may be NULL), zeroed where x <= 0 when mask_relu (x may be NULL otherwise). */
int b200_avgpool2_fwd(int32_t n, int32_t d, int32_t h, int32_t w, int32_t c, int32_t pool_d, b200_view x,
                      b200_view y, b200_stream_t stream);
int b200_avgpool2_bwd(int32_t n, int32_t d, int32_t h, int32_t w, int32_t c, int32_t pool_d, b200_view x,
                      b200_view dy, b200_view add, b200_view dx, int32_t mask_relu, b200_stream_t stream);
/* y = relu(a + b): residual join of ResNetBlock (model/unet3d/buildingblocks.py:319-323) */
int b200_add_relu(int64_t npos, int32_t c, b200_view a, b200_view b, b200_view y, b200_stream_t stream);
/* dst += src: gradient accumulation for tensors that feed several full-scale skip branches */
int b200_add(int64_t npos, int32_t c, b200_view src, b200_view dst, b200_stream_t stream);

/* F.interpolate(mode='nearest') buildingblocks.py:671-673: src = floor(dst * in / out) */
int b200_upsample_nearest_fwd(int32_t n, int32_t id, int32_t ih, int32_t iw, int32_t od,
                              int32_t oh, int32_t ow, int32_t c, b200_view x, b200_view y,
                              b200_stream_t stream);
int b200_upsample_nearest_bwd(int32_t n, int32_t id, int32_t ih, int32_t iw, int32_t od,
                              int32_t oh, int32_t ow, int32_t c, b200_view dy, b200_view dx,
                              b200_stream_t stream);
/* parity planes (see b200_deconv3s2_*) <-> nearest-resized (n, od, oh, ow, c) tensor; the backward
 * writes every plane element (zeros in the unused m = size-1 rows of odd planes) */
int b200_deconv_planes_resize_fwd(int32_t n, int32_t d, int32_t h, int32_t w, int32_t c, int32_t kd,
                                  int32_t od, int32_t oh, int32_t ow, b200_view yp, b200_view y,
                                  b200_stream_t stream);
int b200_deconv_planes_resize_bwd(int32_t n, int32_t d, int32_t h, int32_t w, int32_t c, int32_t kd,
                                  int32_t od, int32_t oh, int32_t ow, b200_view dy, b200_view dyp,
                                  b200_stream_t stream);
/* torch.cat halves — layers.py:188, buildingblocks.py:547: copy c channels of npos positions */
int b200_copy_channels(int64_t npos, int32_t c, b200_view src, b200_view dst, b200_stream_t stream);
/* dz = dy * (a > 0) (ReLU backward), optionally db[c] += sum_pos dz (conv bias gradient) */
int b200_relu_bwd(int64_t npos, int32_t c, b200_view dy, b200_view a, b200_view dz, float* db,
                  b200_stream_t stream);
/* db[c] += sum_pos dy */
int b200_bias_grad(int64_t npos, int32_t c, b200_view dy, float* db, b200_stream_t stream);
/* First convolution of a network (1-4 input channels; nn.Conv3d buildingblocks.py:63-70 with
 * in_channels = the volume's): taps concatenated along K — a[pos][tap * c_in + ci] bf16 (n*d*h*w, kp),
 * zero padding 1 (none in depth when kd = 1), zero beyond kd*9*c_in; kp a multiple of 8.  The weight
 * gradient of that layer is then b200_conv_bwd_weight of a 1x1x1 convolution with c_in = kp.
 * A NEGATIVE c_in = -(real channels) appends one synthesised column per tap, k' = tap * (c + 1) + c, that is 1
 * where the tap lies inside the volume: its weight gradient sum_q dy[q][co] [q + tap inside] yields the
 * parameter gradients of a norm in front of this first conv without a data-gradient pass. */
int b200_im2col3(int32_t n, int32_t d, int32_t h, int32_t w, int32_t c_in, int32_t kd, int32_t kp,
                 b200_view x, void* a, b200_stream_t stream);
/* layout / precision boundary: fp32 NC(spatial) <-> bf16 channels-last */
int b200_nchw_f32_to_cl_bf16(int32_t n, int32_t c, int64_t spatial, const float* src, b200_view dst,
                             b200_stream_t stream);
int b200_cl_bf16_to_nchw_f32(int32_t n, int32_t c, int64_t spatial, b200_view src, float* dst,
                             b200_stream_t stream);

/* ---- GroupNorm (nn.GroupNorm, buildingblocks.py:73-92), eps 1e-5, affine per channel ------------ */
/* stats: fp32 [n][groups][2] = (mean, rstd) over (c/groups x spatial);
 * sums: fp64 [n][2][c] scratch, zeroed by the caller */
int b200_gn_stats(int32_t n, int64_t spatial, int32_t c, int32_t groups, float eps, b200_view x,
                  double* sums, float* stats, b200_stream_t stream);
int b200_gn_apply(int32_t n, int64_t spatial, int32_t c, int32_t groups, b200_view x,
                  const float* stats, const float* gamma, const float* beta, int32_t relu,
                  b200_view y, b200_stream_t stream);
/* dgamma/dbeta fp32 [c] accumulated into (either may be NULL);
 * scratch fp32 [n][2][c] + [n][groups][2], zeroed by the caller */
int b200_gn_bwd(int32_t n, int64_t spatial, int32_t c, int32_t groups, b200_view x,
                const float* stats, const float* gamma, b200_view dy, b200_view dx, float* dgamma,
                float* dbeta, float* scratch, int32_t mask_relu, b200_stream_t stream);

/* b200_gn_bwd_fused for the GroupNorm in front of a decoder's first conv, whose input is the concat
 * (skip [0, cs) | nearest-upsampled [cs, cs + cx)) (buildingblocks.py:539-547, 671-673) with the up-sampling's
 * backward folded in: x / dy are the (n, ld*fd, lh*fh, lw*fw, cs + cx) concat tensors (f* = 1 or 2 per axis);
 * dskip receives the gradient of the skip half at full resolution, dlow (n, ld, lh, lw, cx) the gradient of the
 * up-sampled half summed over each source voxel's children — the full-resolution gradient of that half and the
 * separate F.interpolate backward pass do not exist.  colstats / scratch / dgamma / dbeta as b200_gn_bwd_fused.
 * xlow: {NULL,0,0}, or the low-resolution (n, ld, lh, lw, cx) source of the up-sampled half when the forward ran
 * b200_gn_apply_up and never wrote that half of x.
 * beta_if_y (needs xlow): NULL, or the norm's bias when colstats holds (sum dy, sum dy * y) with y the norm's OUTPUT
 * (a norm without fused activation: y = gamma * xhat + beta, so gamma * sum dy * xhat = sum dy * y - beta * sum dy);
 * channels whose gamma is 0 or below |beta| / 8 get their dgamma from a direct reduction over (dy, x) instead. */
int b200_gn_bwd_fused_up(int32_t n, int32_t ld, int32_t lh, int32_t lw, int32_t fd, int32_t fh, int32_t fw,
                         int32_t cs, int32_t cx, int32_t groups, b200_view x, const float* stats,
                         const float* gamma, b200_view dy, b200_view dskip, b200_view dlow, float* dgamma,
                         float* dbeta, const double* colstats, float* scratch, int32_t mask_relu,
                         b200_view xlow, const float* beta_if_y, b200_stream_t stream);

/* GroupNorm (+ ReLU) of a decoder concat (skip | nearest-upsampled) whose up-sampled half is NOT materialised
 * (replaces F.interpolate(mode="nearest") + torch.cat + nn.GroupNorm, buildingblocks.py:539-547, 671-673, 73-103):
 * xskip = the (n, ld*fd, lh*fh, lw*fw, cs) skip tensor, xlow = the (n, ld, lh, lw, cx) low-resolution tensor,
 * f* = 1 or 2; stats [n][groups][2] over the cs + cx channels; y = the full (…, cs + cx) normalised tensor. */
int b200_gn_apply_up(int32_t n, int32_t ld, int32_t lh, int32_t lw, int32_t fd, int32_t fh, int32_t fw,
                     int32_t cs, int32_t cx, int32_t groups, b200_view xskip, b200_view xlow, const float* stats,
                     const float* gamma, const float* beta, int32_t relu, b200_view y, b200_stream_t stream);
/* First layer of a 3-D U-Net on a ONE-channel volume, on the CUDA cores (replaces the forward of
 * nn.Conv3d(1, 32, 3, padding=1), buildingblocks.py:116-162 with in_channels = 1; c_out = 32 only, every extent >= 2,
 * B200_ERR_UNSUPPORTED otherwise — the caller then takes the zero-padded tensor-core path).
 * x: channel 0 of a channels-last bf16 buffer (any ctot), wgt fp32 [32][27] (the module's own (32, 1, 3, 3, 3)),
 * rounded to bf16 inside; y bf16 (n, d, h, w, 32) view; colstats fp64 [n][2][32] += (sum y, sum y^2) of the stored
 * output, or NULL. */
int b200_conv3d_c1_fwd(int32_t n, int32_t d, int32_t h, int32_t w, int32_t c_out, b200_view x, const float* wgt,
                       const float* bias, int32_t relu, b200_view y, double* colstats, b200_stream_t stream);
/* (mean, rstd) per (sample, group) from per-channel sums: sums fp64 [n][2][c] = (sum x, sum x^2) as
 * produced by b200_conv_fwd_stats (or b200_gn_stats' scratch) */
int b200_gn_stats_finalize(int32_t n, int64_t spatial, int32_t c, int32_t groups, float eps,
                           const double* sums, float* stats, b200_stream_t stream);
/* GroupNorm backward in ONE pass over (x, dy): the reductions come from the producer of dy —
 * colstats fp64 [n][2][c] = (sum dy, sum dy * x) (b200_conv_bwd_data_stats with dot = x).
 * scratch fp32 [n][groups][2]. */
int b200_gn_bwd_fused(int32_t n, int64_t spatial, int32_t c, int32_t groups, b200_view x,
                      const float* stats, const float* gamma, b200_view dy, b200_view dx,
                      float* dgamma, float* dbeta, const double* colstats, float* scratch,
                      int32_t mask_relu, b200_stream_t stream);

/* ---- BatchNorm (nn.BatchNorm2d/3d, layers.py:21, buildingblocks.py:93-103) ---------------------- */
/* sums: fp64 [2][c] = (sum x, sum x^2) accumulated into (caller zeroes) */
int b200_bn_stats(int64_t npos, int32_t c, b200_view x, double* sums, b200_stream_t stream);
/* y = (x - mean) * rstd * gamma + beta, optional ReLU; mean/rstd fp32 [c] */
int b200_bn_apply(int64_t npos, int32_t c, b200_view x, const float* mean, const float* rstd,
                  const float* gamma, const float* beta, int32_t relu, b200_view y,
                  b200_stream_t stream);
/* pass 1: red fp32 [2][c] += (sum dz, sum dz * xhat) with dz = dy * (y > 0 if relu_out given) */
int b200_bn_bwd_reduce(int64_t npos, int32_t c, b200_view x, b200_view dy, b200_view relu_out,
                       const float* mean, const float* rstd, float* red, b200_stream_t stream);
/* pass 1 replaced by fused sums: colstats fp64 [n][2][c] = (sum dz, sum dz * x) ->
 * red fp32 [2][c] = (sum dz, rstd * (sum dz*x - mean * sum dz)) */
int b200_bn_bwd_red_from_sums(int32_t n, int32_t c, const double* colstats, const float* mean,
                              const float* rstd, float* red, b200_stream_t stream);
/* pass 2: dx = gamma * rstd * (dz - red0/M - xhat * red1/M) */
int b200_bn_bwd_apply(int64_t npos, int32_t c, b200_view x, b200_view dy, b200_view relu_out,
                      const float* mean, const float* rstd, const float* gamma, const float* red,
                      b200_view dx, int32_t mask_relu, b200_stream_t stream);

/* ---- 1x1 head (nn.Conv2d(64,C,1) unet.py:89; nn.Conv3d(f0,C,1) model.py:110-113) ---------------- */
/* logits fp32 (n, c_out, spatial) = x (n, spatial, c_in) . w^T + b;  w fp32 [c_out][c_in] */
int b200_head_fwd(int32_t n, int64_t spatial, int32_t c_in, int32_t c_out, b200_view x,
                  const float* w, const float* b, float* logits, b200_stream_t stream);
/* dx bf16, dw fp32 [c_out][c_in] and db fp32 [c_out] accumulated into; mask_relu as above */
int b200_head_bwd(int32_t n, int64_t spatial, int32_t c_in, int32_t c_out, b200_view x,
                  const float* w, const float* dlogits, b200_view dx, float* dw, float* db,
                  int32_t mask_relu, b200_stream_t stream);

/* ---- losses ------------------------------------------------------------------------------------
 * One pass over fp32 logits (n, c, spatial) and a same-shape fp32 target produces the per-channel
 * sums every loss of the path needs:  sums[c][6] (fp64) =
 *   { sum p*t, sum p*p, sum t*t, sum p, sum t, sum bce(x,t) }  with p = sigmoid(x).
 * Serves compute_per_channel_dice / DiceLoss / BCEDiceLoss (model/unet3d/losses.py:7-33,119-129,
 * 167-178), nn.BCEWithLogitsLoss (model/unet2d/unet.py:1184-1188), IoULoss / F1Loss
 * (model/unet2d/loss.py:32-57) and DiceCoefficient (model/unet3d/metrics.py:15-29).
 * normalization: 0 = sigmoid, 1 = none (input already a probability; bce column is then 0). */
int b200_seg_sums(int32_t n, int32_t c, int64_t spatial, const float* logits, const float* target,
                  int32_t normalization, double* sums, b200_stream_t stream);
/* normalization 0: dlogits = ka*(p - t) + p*(1-p) * (k1[c]*t + k2[c]*p + k3[c]);
 * normalization 1: dlogits = k1[c]*t + k2[c]*x + k3[c].  coef fp32 [c][4] = (ka,k1,k2,k3) are
 * d(loss)/d(sums) read from device memory, so no host sync is needed. */
int b200_seg_loss_bwd(int32_t n, int32_t c, int64_t spatial, const float* logits,
                      const float* target, const float* coef, int32_t normalization, float* dlogits,
                      b200_stream_t stream);
/* nn.CrossEntropyLoss() (unet.py:1185): nll_sum is fp64 [3] accumulated into:
 * [0] += sum over valid positions of -log softmax(x)[label], [1] += number of valid positions (label !=
 * ignore_index = -100), [2] += number of labels outside [0, c) other than -100 (never dereferenced; the
 * host wrapper turns a non-zero count into a NaN loss).  b200_ce_bwd writes zero gradient at ignored /
 * invalid positions. */
int b200_ce_fwd(int32_t n, int32_t c, int64_t spatial, const float* logits, const int64_t* labels,
                double* nll_sum, b200_stream_t stream);
/* dlogits = scale[0] * (softmax(x) - onehot(label)) */
int b200_ce_bwd(int32_t n, int32_t c, int64_t spatial, const float* logits, const int64_t* labels,
                const float* scale, float* dlogits, b200_stream_t stream);

/* ---- MS-SSIM term of SegmentationLoss (model/unet2d/loss.py:21-28 -> pytorch-msssim==1.0.0 MS_SSIM(data_range,
 * channel=1), requirements.txt:127) on sigmoid(logits) vs targets, both fp32 (n, 1, h, w); min(h, w) > 160 as the
 * package asserts.  A pyramid buffer holds the 5 scales of one image stack packed back to back:
 * b200_msssim_pyramid_floats(n, h, w) floats (negative = unsupported shape); all buffers are caller-owned.
 * fwd: fills pyr_x / pyr_y and sums (fp64 [5][n], zeroed inside): per sample the SUM over valid positions of the
 *      cs map (scales 0-3) / ssim map (scale 4); valid positions of scale l = (h_l - 10) * (w_l - 10),
 *      h_{l+1} = ceil(h_l / 2).
 * bwd: gmeans fp32 [5][n] = d(loss)/d(per-sample MEAN of each scale's map) read from device memory;
 *      pyr_dx = scratch pyramid; dlogits (n, 1, h, w) fully written. */
int64_t b200_msssim_pyramid_floats(int32_t n, int32_t h, int32_t w);
int b200_msssim_fwd(int32_t n, int32_t h, int32_t w, const float* logits, const float* targets, float* pyr_x,
                    float* pyr_y, double* sums, float data_range, b200_stream_t stream);
int b200_msssim_bwd(int32_t n, int32_t h, int32_t w, const float* pyr_x, const float* pyr_y, const float* gmeans,
                    float* pyr_dx, float* dlogits, float data_range, b200_stream_t stream);

/* ---- class-guided module of UNet_3Plus_DeepSup_CGM (model/unet2d/unet.py:999-1017, 1036-1038, 1147-1153) ----
 * gate fp32 [n] = argmax over the 2 classes of sigmoid(max over positions of conv1x1(dropout(x)) + bias), x = hd5
 * (n, npos, c) channels-last bf16, w fp32 [2][c].  keep: channels-last uint8 [n][npos][c] dropout mask (NULL = eval
 * mode), kept values are scaled by keep_scale = 1 / (1 - p).  cls (may be NULL): fp32 [n][2] class probabilities. */
int b200_cgm_gate(int32_t n, int64_t npos, int32_t c, b200_view x, const float* w, const float* bias,
                  const uint8_t* keep, float keep_scale, float* gate, float* cls, b200_stream_t stream);
/* y = sigmoid(x * gate[sample]) (dotProduct + F.sigmoid, unet.py:1012-1017, 1153); dx = dy * y (1 - y) * gate */
int b200_gate_sigmoid_fwd(int32_t n, int64_t per_sample, const float* x, const float* gate, float* y,
                          b200_stream_t stream);
int b200_gate_sigmoid_bwd(int32_t n, int64_t per_sample, const float* y, const float* dy, const float* gate,
                          float* dx, b200_stream_t stream);

/* ---- concurrent spatial and channel squeeze-and-excitation (model/unet3d/se.py:18-113 ChannelSpatialSELayer3D as
 * applied by ResNetBlockSE, buildingblocks.py:328-362) on a channels-last bf16 ReLU output y (n, spatial, c), c <= 1024:
 *   z = max(y * a[n,c], y * sigmoid(s[n,pos])) = y * max(a, b);  a = cSE gate (fp32 [n][c], computed by the caller from
 *   the squeeze with two small FCs), s = sSE logits (fp32 [n][spatial], the 1x1x1 conv = b200_head_fwd).
 * channel_sums: sums fp64 [n][c] += sum_pos y (caller zeroes).   gate_fwd: z.
 * gate_bwd_reduce: da fp64 [n][c] += sum_pos g*y*[a >= b] (caller zeroes); ds fp32 [n][spatial] = d(loss)/d(s).
 * gate_bwd_apply: dy = g*max(a,b) + (y > 0 ? dmean[n][c] : 0) + dsse, with dmean = d(loss)/d(squeeze) / spatial and
 *   dsse = the sSE conv's data gradient (b200_head_bwd, masked by y > 0); g must already carry z's ReLU mask. */
int b200_se_channel_sums(int32_t n, int64_t spatial, int32_t c, b200_view y, double* sums, b200_stream_t stream);
int b200_se_gate_fwd(int32_t n, int64_t spatial, int32_t c, b200_view y, const float* a, const float* s, b200_view z,
                     b200_stream_t stream);
int b200_se_gate_bwd_reduce(int32_t n, int64_t spatial, int32_t c, b200_view y, const float* a, const float* s,
                            b200_view g, double* da, float* ds, b200_stream_t stream);
int b200_se_gate_bwd_apply(int32_t n, int64_t spatial, int32_t c, b200_view y, const float* a, const float* s,
                           b200_view g, const float* dmean, b200_view dsse, b200_view dy, b200_stream_t stream);

/* ---- sliding-window prediction on the device (model/unet3d/predictor.py:65-165; SliceBuilder / mirror_pad /
 * remove_padding of dataset/unet3d_dataset/utils.py:40-131, 314-361).  fp32 NCDHW, caller-owned buffers.
 * gather: out (c, pd, ph, pw) = the window of vol (c, d, h, w) that starts `halo` before the patch origin (oz, oy, ox),
 *         read with np.pad(mode='reflect') indexing — the padded volume is never materialised; pd = patch + 2 * hz ...
 * accumulate: map[ch, o + v] += pred[ch (or `channel` if >= 0), v + halo], norm[...] += 1 over the inner patch
 *         (predictor.py:147-150); map (c, d, h, w) fp32, norm uint8 of the same shape; one call per patch keeps the
 *         reference's summation order (deterministic, bit-identical).
 * finalize: map /= norm in place; seg (may be NULL) int32 (d*h*w) = arg-max over channels (predictor.py:161-165). */
int b200_patch_gather(int32_t c, int32_t d, int32_t h, int32_t w, const float* vol, int32_t oz, int32_t oy,
                      int32_t ox, int32_t hz, int32_t hy, int32_t hx, int32_t pd, int32_t ph, int32_t pw, float* out,
                      b200_stream_t stream);
int b200_patch_accumulate(int32_t c, int32_t d, int32_t h, int32_t w, int32_t pd, int32_t ph, int32_t pw, int32_t hz,
                          int32_t hy, int32_t hx, const float* pred, int32_t c_pred, int32_t channel, int32_t oz,
                          int32_t oy, int32_t ox, float* map, uint8_t* norm, b200_stream_t stream);
int b200_prediction_finalize(int32_t c, int64_t nvox, float* map, const uint8_t* norm, int32_t* seg,
                             b200_stream_t stream);

/* ---- metrics -----------------------------------------------------------------------------------
 * trainer/metrcis.py:61-109,153-168: p = 1/(1+exp(-x)+1e-6); psum (fp64 [1]) += sum p */
int b200_metric_prob_sum(int64_t numel, const float* logits, double* psum, b200_stream_t stream);
/* counts fp64 [n][3] += per-sample { |P & L|, |P|, |L| } with P = p > thr[0], L = label > thr[0].
 * mode 0: p = 1/(1+exp(-x)+1e-6) (compute_metrics); mode 1: p = sigmoid(x) (iou_score);
 * mode 2: p = x (already a probability map: compute_iou/compute_dice, MeanIoU C=1).
 * thr is read from device memory. */
int b200_metric_counts(int32_t n, int64_t per_sample, const float* logits, const float* labels,
                       const float* thr, int32_t mode, double* counts, b200_stream_t stream);
/* MeanIoU multi-class (model/unet3d/metrics.py:45-104): one-hot argmax over c of x (n,c,spatial)
 * against target (n,c,spatial); counts fp64 [n][c][2] += { |P & T|, |P | T| } */
int b200_metric_argmax_counts(int32_t n, int32_t c, int64_t spatial, const float* x,
                              const float* target, double* counts, b200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200UNET_H */
