/* nhvr.h — C-ABI of libnhvr_sm100.so: the B200 (sm_100a) rendering hot path of
 * Neural-Human-Video-Rendering.
 *
 * The reference ships no source; every entry point below therefore cites the
 * nearest pinned evidence (launch flag / README line) of the reference interface it stands in
 * for, and the torch.nn call it replaces inside the (absent) reference `models/networks.py`.
 *
 * Conventions
 *  - plain pointers and sizes only; all pointers are DEVICE pointers unless a name ends in _host;
 *  - the caller owns every byte of device memory; the library allocates no device memory and
 *    launches asynchronously on the cudaStream_t passed as `void* stream`;
 *  - every function returns 0 on success or a negative nhvr_status; nhvr_strerror() names it;
 *  - there is no CPU fallback: on a device that is not compute capability 10.x every compute
 *    entry point returns NHVR_ERR_ARCH.
 *
 * Activation layout "P8" (planar-by-8, 16-bit elements: fp16 or bf16, see nhvr_set_operand_dtype): [N][C8][Hp][Wp][8] where C8 = ceil(C/8) and one
 * 16-byte unit holds 8 consecutive channels of one pixel.  Hp/Wp include a halo that the WRITER
 * fills (zeros or mirror = ReflectionPad2d), so the conv kernel never pads.  `split`=1 stores the
 * padded image as four row/column-parity sub-images, which turns a stride-2 conv into unit shifts.
 */
#ifndef NHVR_H_
#define NHVR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum nhvr_status {
  NHVR_OK = 0,
  NHVR_ERR_ARCH = -1,      /* not an sm_100 device */
  NHVR_ERR_SHAPE = -2,     /* unsupported / inconsistent shape */
  NHVR_ERR_ALIGN = -3,     /* pointer not 16-byte aligned */
  NHVR_ERR_NULL = -4,      /* required pointer is NULL */
  NHVR_ERR_CUDA = -5,      /* a CUDA runtime call failed (see nhvr_last_cuda_error) */
  NHVR_ERR_SMEM = -6,      /* tile does not fit in shared / tensor memory */
  NHVR_ERR_UNSUPPORTED = -7
} nhvr_status;

enum { NHVR_HALO_ZERO = 0, NHVR_HALO_REFLECT = 1 };
enum { NHVR_ACT_NONE = 0, NHVR_ACT_RELU = 1, NHVR_ACT_LRELU02 = 2, NHVR_ACT_TANH = 3,
       NHVR_ACT_TANH_SIGMOID_LAST = 4 /* tanh on all channels but the last, sigmoid on the last */ };
enum { NHVR_CONV = 0, NHVR_CONV_TRANSPOSE = 1 /* stride 2: k3 p1 output_padding 1, or k4 p2 (input-gradient of a 4x4 s2 p2 conv) */,
       NHVR_CONV_DGRAD_S1 = 2 /* input-gradient of a stride-1 conv; the desc describes the FORWARD conv */ };
enum { NHVR_EPI_RAW_STATS = 0,   /* 16-bit P8 un-padded conv output + per-(n,c) sum / sum-of-squares */
       NHVR_EPI_BIAS_ACT_F32 = 1,/* bias + activation, fp32 NCHW output                            */
       NHVR_EPI_BIAS_ACT_P8 = 2, /* bias + activation, 16-bit P8 output in a consumer's format     */
       NHVR_EPI_RAW_P8 = 3,      /* P8 un-padded output, no statistics (gradient convs)            */
       NHVR_EPI_IN_FUSED = 4     /* internal: set by nhvr_conv_forward_in_fused on a RAW_STATS plan  */ };

/* P8 activation descriptor (see header comment). */
typedef struct nhvr_act_desc {
  int32_t N, C8, H, W;
  int32_t pad_t, pad_l, pad_b, pad_r;
  int32_t split;     /* 0 plain, 1 four-way parity split (Hp, Wp rounded up to even) */
  int32_t halo;      /* NHVR_HALO_* : how a writer must fill the halo */
  int32_t hilo;      /* 0: one 16-bit value per element.  1: split precision - every value v is stored as the pair
                        hi = rn16(v), lo = rn16(v - hi) (22 significant bits with fp16 operands).  C8 then counts PHYSICAL
                        planes = 2 x the logical plane count, in groups of four: [hi 2g, hi 2g+1, lo 2g, lo 2g+1]. */
} nhvr_act_desc;

/* One convolution layer.  Replaces nn.Conv2d / nn.ConvTranspose2d (+ the ReflectionPad2d in front)
 * of the pix2pixHD-shaped generators the reference builds from --n_downsample_global,
 * --n_blocks_global, --ngf_global (test_start/start.sh:15-17), --n_downsample_bg/--n_blocks_bg
 * (start.sh:20-21), --n_blocks_translate (pretrainTrans.sh:13). */
typedef struct nhvr_conv_desc {
  int32_t kind;          /* NHVR_CONV / NHVR_CONV_TRANSPOSE */
  int32_t Cin, Cout;
  int32_t kh, kw, stride, pad;
  int32_t N, H, W;       /* input logical size */
  int32_t halo;          /* NHVR_HALO_* the input must carry (reflect for ReflectionPad2d, zero for padding=) */
  int32_t epilogue;      /* NHVR_EPI_* */
  int32_t act;           /* NHVR_ACT_* (epilogues 1, 2) */
  int32_t in_extra_rows; /* extra zero rows below the input's bottom halo (gradient buffers shared with wgrad) */
  int32_t in_extra_cols; /* extra zero columns right of the input's right halo (same purpose)                 */
  int32_t out_h, out_w;  /* NHVR_CONV_TRANSPOSE only: output size override (0 = 2H x 2W for k3, 2H-2 for k4)  */
  int32_t flags;         /* bit 0: never use the row-mode lowering (wide kernels with few output channels)
                            bit 5: intended for nhvr_conv_forward_in_fused: the tile step shrinks so that an image has exactly
                                   one tile per SM (148), i.e. whole images fill the resident CTA slots
                            bit 4: RAW_STATS sums are centred on the shift the caller stored in slot 2 of the statistics record
                                   (nhvr_stem_stat_shift; first layers, where |mean| >> std on stick-figure pose maps)
                            bit 3: split precision ("3 x fp16"): the input is a hilo activation (nhvr_act_desc.hilo), the
                                   weights are packed as hi + lo blocks and every K step issues three MMAs
                                   (x_hi*w_hi + x_hi*w_lo + x_lo*w_hi) into the same fp32 accumulator; a RAW_STATS output
                                   is written as a hilo activation too.  fp32-class results at 3x the tensor work: used by
                                   the UV generator, whose output error is multiplied by the texture gradient in the lookup
                                   (inference only: no dgrad / wgrad plans for it)
                            bit 6: with bit 3: no w_lo blocks - two MMAs per K step (x_hi*w_hi + x_lo*w_hi): the activations keep
                                   ~22 bits, the weights are rounded to 16 (the temporal generator of the "strict2" preset)
                            bit 2: the input may use the single-plane tap-paired format (Cin <= 8 stride-1 convs:
                                   K group 1 of every MMA is the same plane one pixel to the right, so an MMA covers
                                   two filter columns).  Changes nhvr_conv_input_desc (C8 = 1): set it only when
                                   nothing else (e.g. a wgrad plan) reads the same input buffer                   */
} nhvr_conv_desc;

typedef struct nhvr_conv_plan nhvr_conv_plan;   /* opaque, host memory only */

/* ---- library ---- */
int nhvr_version(void);
const char* nhvr_strerror(int status);
const char* nhvr_last_cuda_error(void);
int nhvr_arch_ok(void);                 /* 0 iff the current device is compute capability 10.x */
/* 16-bit element type of P8 activations and packed weights (the tcgen05 kind::f16 operands):
 * 0 = bf16, 1 = IEEE fp16 (same tensor throughput, 3 more mantissa bits; the Python host selects fp16 by
 * default, see capi.DEFAULT_OPERAND).  fp16 conversions do NOT saturate: a value beyond 65504 becomes inf,
 * propagates, and is reported through nhvr_set_overflow_flag.  Process-wide; buffers and packed weights
 * written under one setting must be consumed under the same setting. */
int nhvr_set_operand_dtype(int is_f16);
int nhvr_get_operand_dtype(void);
/* Range guard of the 16-bit path: `flag` (device int32, caller-owned, may be NULL to disable) is OR-ed with 1 by
 * the InstanceNorm apply / backward kernels whenever they read a non-finite 16-bit value (an fp16 overflow of a conv
 * output or gradient).  Process-wide; the host reads it at a synchronisation point of its choice. */
int nhvr_set_overflow_flag(int32_t* flag);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
uint64_t nhvr_launch_count(void);

/* ---- P8 activations ---- */
size_t nhvr_act_bytes(const nhvr_act_desc* d);   /* includes the tail slack tiles may over-read */
/* NCHW fp32 -> P8 (16-bit operand type) with halo.  Up to 4 sources are concatenated along C (the reference's
 * torch.cat of texture / pose / Laplace / previous frame in front of the generator; evidence:
 * --input_nc 3, --use_laplace, --pose_plus_laplace, name *_Temporal, start.sh:7,11,19,24).
 * src[i] is [N][src_c[i]][H][W] fp32; channels beyond the sum are zero. */
int nhvr_pack_nchw(const float* const* src, const int32_t* src_c, int32_t nsrc,
                   void* dst, const nhvr_act_desc* dst_desc, void* stream);
/* P8 (interior) -> NCHW fp32, first C channels. */
int nhvr_unpack_nchw(const void* src, const nhvr_act_desc* src_desc, float* dst, int32_t C, void* stream);

/* ---- convolution (tcgen05 shift-GEMM) ---- */
int nhvr_conv_plan_create(const nhvr_conv_desc* d, nhvr_conv_plan** out);
void nhvr_conv_plan_destroy(nhvr_conv_plan* p);
int nhvr_conv_input_desc(const nhvr_conv_plan* p, nhvr_act_desc* in_desc);   /* format the input must have */
int nhvr_conv_output_dims(const nhvr_conv_plan* p, int32_t* Ho, int32_t* Wo, int32_t* Cout8);
/* replace the plan's input descriptor by a layout-compatible one with a taller bottom halo (gradient buffers
 * shared with a wgrad plan, see nhvr_wgrad_grad_desc) */
int nhvr_conv_plan_set_input_desc(nhvr_conv_plan* p, const nhvr_act_desc* desc);
size_t nhvr_conv_weight_bytes(const nhvr_conv_plan* p);
double nhvr_conv_flops(const nhvr_conv_plan* p);     /* algorithmic 2*k*k*Cin*Cout*Ho*Wo*N, un-padded */
/* tiling introspection: info[0..15] = kcp, nchunks, njobs, nruns, nacc, slab_units, Npad, bpb, nbstages,
 * SA, SB, tmem_cols, smem_bytes, tiles_per_img, nsplit, nblocks (used by tests and DESIGN.md tables) */
int nhvr_conv_plan_info(const nhvr_conv_plan* p, int32_t* info, int32_t n);
/* w: fp32, Conv2d layout [Cout][Cin][kh][kw] or ConvTranspose2d layout [Cin][Cout][kh][kw]. */
int nhvr_conv_pack_weights(const nhvr_conv_plan* p, const float* w, void* packed, void* stream);
/* The same packing for MANY layers in one launch (a training step re-packs every weight of every network after each optimiser
 * update).  nhvr_conv_pack_record_fill writes one layer's record (nhvr_conv_pack_record_bytes() bytes, HOST memory; *units = the
 * layer's packed 16-byte units) for the device pointers w / packed; the caller concatenates the records, copies them to the
 * device once and calls nhvr_conv_pack_weights_batched(records, n, max units over the layers) whenever the weights changed.
 * Not for split-precision plans (NHVR_ERR_UNSUPPORTED: their pack needs max|w| first). */
size_t nhvr_conv_pack_record_bytes(void);
int nhvr_conv_pack_record_fill(const nhvr_conv_plan* p, const float* w, void* packed, void* record_host, int64_t* units);
int nhvr_conv_pack_weights_batched(const void* records_dev, int32_t n, int64_t max_units, void* stream);
/* out / stats / bias meaning depends on the plan's epilogue:
 *  RAW_STATS    : out = P8 [N][Cout8][Ho][Wo] (no halo, Cout8 = ceil(Cout/8) rounded up to even); stats = double [N][Cout8*8][4]
 *                 = {sum (x-s), sum (x-s)^2, s, unused} per channel over H*W, zeroed by the caller (s = 0: plain sums) and
 *                 optionally centred with nhvr_stem_stat_shift before the conv; fp32 per-tile partial sums are merged with fp64
 *                 atomics, so that the variance keeps its digits when |mean| >> std; bias unused
 *                 (a bias in front of an affine-free InstanceNorm cancels exactly).
 *  BIAS_ACT_F32 : out = float [N][Cout][Ho][Wo]; bias = float [Cout] or NULL.
 *  BIAS_ACT_P8  : out = P8 in out_desc's format (interior only is written); bias as above. */
int nhvr_conv_forward(const nhvr_conv_plan* p, const void* in, const void* packed_w, const float* bias,
                      void* out, const nhvr_act_desc* out_desc, double* stats, void* stream);

/* Centring shift of a FIRST layer's statistics: writes s[n][co] = sum_ci wsum[co][ci] * mean(src[n][ci]) (mean over 256 sampled
 * pixels; wsum = the filter summed over its taps) into slot 2 of the zeroed statistics record, i.e. the conv output wherever
 * the input is flat.  Stick-figure pose maps
 * (keypoints/*.json rasterised, start.sh:9,24) are ~98 % background: the stem output is c + small with mean^2/var up to 250,
 * and un-centred sums lose 2-3 digits of the variance.  wsum: fp32 [Cout][Cin]; src as in nhvr_pack_nchw (Cin <= 32, H*W >= 256). */
int nhvr_stem_stat_shift(const float* wsum, int32_t Cout, int32_t Cin, const float* const* src, const int32_t* src_c,
                         int32_t nsrc, int32_t N, int32_t H, int32_t W, double* stats, void* stream);

/* ---- InstanceNorm2d(affine=False) apply + activation (+ residual) + halo write ----
 * Replaces nn.InstanceNorm2d + nn.ReLU/LeakyReLU (+ the ResnetBlock skip add, + the next layer's
 * ReflectionPad2d / zero padding).  raw: P8 un-padded; stats: the RAW_STATS record [N][C8*8][4]; residual
 * (nullable) is read at the interior of res_desc; dst is written completely, halo included. */
int nhvr_in_apply(const void* raw, const nhvr_act_desc* raw_desc, const double* stats, float eps, int32_t act,
                  const void* residual, const nhvr_act_desc* res_desc,
                  void* dst, const nhvr_act_desc* dst_desc, void* stream);

/* ---- conv + InstanceNorm2d + activation (+ residual) + halo write in ONE kernel (north_star: "InstanceNorm, ReLU and
 * reflection padding are fused into conv prologues and epilogues") ----
 * Same result as nhvr_conv_forward (RAW_STATS plan) followed by nhvr_in_apply, without the raw tensor's HBM round trip: the
 * accumulators of a tile stay in TMEM while the CTAs of its image meet at a per-image arrival counter (after their sum /
 * sum-of-squares atomics); then each CTA normalises its own tile from the fp32 accumulators and writes it into dst, mirrored
 * halo copies included (a zero halo is left untouched: the buffer must have been zeroed once).
 * Needs every CTA of an image resident at once: nhvr_conv_in_fused_supported() returns 1 when tiles-per-image x N-splits fits
 * half the GPU's resident CTA slots (so that two such kernels on concurrent streams cannot starve each other), the plan has one
 * M block / one accumulator per CTA and its statistics are not centred.  128 x 128 ResnetBlock layers: 130 tiles per image.
 * stats: zeroed by the caller; on return its four slots per channel hold two replicas of {sum, sum of squares} (even / odd
 * tiles; no centring shift on this path).  sync: device uint32 [N] arrival counters, zeroed by the caller before every launch
 * (the engines keep them behind the statistics so that one fill re-arms both).
 * A wait that exceeds ~2 s traps (sticky CUDA error) instead of hanging the GPU.  Consequently at most TWO such launches may be in
 * flight on a device at a time (the engines use one stream per network: two); more concurrent launches, or a GPU time-sliced
 * with another process for seconds, can exhaust the slots / the time limit and end in that trap. */
int nhvr_conv_in_fused_supported(const nhvr_conv_plan* p);
int nhvr_conv_forward_in_fused(const nhvr_conv_plan* p, const void* in, const void* packed_w, double* stats, float eps, int32_t act,
                               const void* residual, const nhvr_act_desc* res_desc, void* dst, const nhvr_act_desc* dst_desc,
                               uint32_t* sync, void* stream);

/* ---- keypoints -> pose maps (the step before the path: --pose_path ./keypoints of OpenPose BODY_25 JSONs, start.sh:9,24-25) ----
 * kps: device float [T][25][3] (x, y, confidence) in a src_size^2 frame; out: device float [T][pose_nc][size][size]: the
 * 3-channel stick figure of nhvr_b200/pose.py in [-1, 1] (bit-identical to the host rasteriser), further channels zero
 * (LaplaceProj input of --use_laplace: no data in the fixtures).  limb_colors_host: HOST uint8 [24][3]. */
int nhvr_pose_rasterize(const float* kps, int32_t T, int32_t size, float src_size, float thickness, float conf_thresh,
                        int32_t pose_nc, const uint8_t* limb_colors_host, float* out, void* stream);

/* ---- texture lookup ("--TexG part --use_mask_texture", start.sh:14,18; README.md:64) ----
 * uvp: float [N][73][H][W] = 25 part logits, 24 U, 24 V (raw UV-generator output).
 * atlas: float [24][S][S][Ct4] (channels-last, Ct4 = Ctex rounded up to 4).
 * tex_out: float [N][Ctex][H][W] = sum_k softmax(logits)_k * bilinear(atlas_k, u_k, v_k), k=1..24
 *          (divided by (1-P_0) when use_mask_texture == 0).
 * part_out (nullable): uint8 [N][H][W] argmax part (lowest index wins ties);
 * texel_out (nullable): int16 [N][H][W][2] = (x0, y0) integer texel corner of the argmax part
 *          (0,0 for background).  These two are the bit-exact integer contract. */
int nhvr_texture_sample(const float* uvp, const float* atlas, int32_t N, int32_t H, int32_t W,
                        int32_t S, int32_t Ctex, int32_t use_mask_texture,
                        float* tex_out, uint8_t* part_out, int16_t* texel_out, void* stream);
/* ---- unfold_texture (README.md:64: the initial texture.jpg built from the frames and their DensePose IUV) ----
 * The adjoint of the bilinear lookup: every pixel with part dp_i in 1..24 splats img into the four texels around
 * (u, v) * (S-1) of its part with the lookup's weights.  img float [N][C][H][W]; dp_i int32 [N][H][W]; dp_uv float [N][2][H][W]
 * in [0,1]; acc float [24][S][S][Q], Q = 4*ceil((C+1)/4) (C colour sums then the weight sum), zeroed by the caller and
 * accumulated over as many calls as there are frame batches; _finish divides: atlas float [24][C][S][S], 0 where the
 * weight is <= min_weight. */
int nhvr_texture_unfold(const float* img, const int32_t* dp_i, const float* dp_uv, int32_t N, int32_t H, int32_t W, int32_t S,
                        int32_t C, float* acc, void* stream);
int nhvr_texture_unfold_finish(const float* acc, int32_t S, int32_t C, float min_weight, float* atlas, void* stream);

/* ---- mask / background composite (README.md:15,52,60; --bg_path start.sh:12) ----
 * out = m*fg + (1-m)*bg.  fgm: float [N][4][H][W] (RGB in [-1,1], mask in [0,1]);
 * bg: float [3][H][W] (bg_batched==0, broadcast) or [N][3][H][W]; out: float [N][3][H][W]. */
int nhvr_composite(const float* fgm, const float* bg, int32_t bg_batched, int32_t N, int32_t H, int32_t W,
                   float* out, void* stream);

/* ---- weight gradient (tcgen05, MN-major operands straight from P8) ----
 * fwd describes the FORWARD conv (NHVR_CONV any stride, or NHVR_CONV_TRANSPOSE).  x is the forward conv's
 * P8 input (nhvr_conv_input_desc of the forward plan); g is the output gradient in the format
 * nhvr_wgrad_grad_desc() returns, which is also the (bottom-extended) input format of the matching dgrad
 * conv: NHVR_CONV_DGRAD_S1 for a stride-1 conv, NHVR_CONV_TRANSPOSE for a stride-2 conv, NHVR_CONV stride 2
 * for a transposed conv.  dw has the forward weight's layout and receives (accumulate ? dw : 0) + scale * dW. */
typedef struct nhvr_wgrad_plan nhvr_wgrad_plan;
int nhvr_wgrad_plan_create(const nhvr_conv_desc* fwd, nhvr_wgrad_plan** out);
void nhvr_wgrad_plan_destroy(nhvr_wgrad_plan* p);
int nhvr_wgrad_grad_desc(const nhvr_wgrad_plan* p, nhvr_act_desc* g_desc);
size_t nhvr_wgrad_workspace_bytes(const nhvr_wgrad_plan* p);
int nhvr_wgrad(const nhvr_wgrad_plan* p, const void* x, const void* g, void* workspace, float* dw, float scale,
               int32_t accumulate, void* stream);

/* ---- backward of the texture lookup (both blends: --use_mask_texture of start.sh:18, and the renormalised one
 * train_start/pretrain_start.sh runs with) and of the composite ----
 * grad_uvp float [N][73][H][W] is fully written; grad_atlas float [24][S][S][Ct4] (channels-last, zeroed by the
 * caller) receives vector reductions. */
int nhvr_texture_sample_bwd(const float* uvp, const float* atlas, const float* grad_tex, int32_t N, int32_t H, int32_t W,
                            int32_t S, int32_t Ctex, int32_t use_mask_texture, float* grad_uvp, float* grad_atlas, void* stream);
/* grad_fgm float [N][4][H][W]; grad_bg float [3][H][W] (summed over the batch) or [N][3][H][W] if bg_batched */
int nhvr_composite_bwd(const float* fgm, const float* bg, int32_t bg_batched, const float* grad_out, int32_t N, int32_t H,
                       int32_t W, float* grad_fgm, float* grad_bg, void* stream);

/* ---- backward of  x_next = pad(act(InstanceNorm(raw)) [+ residual]) ----
 * dx: gradient w.r.t. the PADDED x_next as an un-padded P8 buffer [N][C8][dx_H][dx_W] (the dgrad conv's
 * RAW_P8 output); the interior starts at (pad_t, pad_l); reflect != 0 folds the mirrored halo gradients
 * back (ReflectionPad2d), 0 drops them (zero padding).  skip (nullable): extra gradient of the un-padded
 * output (ResnetBlock skip path), P8 [N][C8][H][W].  raw / stats: the forward RAW_STATS outputs.
 * Writes g (gradient w.r.t. raw) in g_desc's format, halo zeroed; dy_out (nullable) receives the folded
 * total gradient of the un-padded output; sums is a 4-byte scratch of N*C8*17 elements (float [N][C8*8][2] sums + N*C8
 * arrival counters of the single-launch kernel, whose second pass re-reads its rows from L2). */
int nhvr_in_bwd(const void* dx, int32_t dx_H, int32_t dx_W, int32_t pad_t, int32_t pad_l, int32_t reflect,
                const void* skip, const void* raw, const nhvr_act_desc* raw_desc, const double* stats, float eps,
                int32_t act, float* sums, void* g, const nhvr_act_desc* g_desc, void* dy_out, void* stream);
/* same for a layer WITHOUT normalisation (conv + bias + activation): g = dY * act'(y), y read from the stored
 * activation yact (P8 in y_desc's format); dbias float [C8*8] (zeroed by the caller) += per-channel sums of g */
int nhvr_act_bwd(const void* dx, int32_t dx_H, int32_t dx_W, int32_t pad_t, int32_t pad_l, int32_t reflect,
                 const void* skip, const void* yact, const nhvr_act_desc* y_desc, int32_t act, void* g,
                 const nhvr_act_desc* g_desc, float* dbias, void* stream);
/* folded gradient of a chain input -> float [N][C][H][W] * scale (interior_desc: un-padded P8 of the input) */
int nhvr_fold_unpack(const void* dx, int32_t dx_H, int32_t dx_W, int32_t pad_t, int32_t pad_l, int32_t reflect,
                     const nhvr_act_desc* interior_desc, float* dst, int32_t C, float scale, void* stream);
/* output layer (bias + activation, no norm): g_pre = grad_out * act'(out) * scale, all float [N][C][H][W] */
int nhvr_head_bwd(const float* out, const float* grad_out, int32_t N, int32_t C, int32_t H, int32_t W, int32_t act,
                  float scale, float* g_pre, void* stream);
/* db[c] = (accumulate ? db[c] : 0) + scale * sum over n,h,w of g[n][c][h][w] */
int nhvr_bias_grad(const float* g, int32_t N, int32_t C, int32_t H, int32_t W, float scale, int32_t accumulate, float* db,
                   void* stream);

/* ---- training-side reductions (fp32 in, fp64 accumulate) ----
 * Each call ADDS partial sums into a caller-zeroed double accumulator; the host divides by the element
 * count to get the mean the reference's torch losses return (pretrain_start.sh:31-37 --lambda_L2 /
 * --lambda_UV / --lambda_Prob / --lambda_Temp; pix2pixHD GANLoss(use_lsgan) and feature matching). */
int nhvr_loss_sum_sq_diff(const float* a, const float* b, int64_t n, double* acc, void* stream);   /* L2 / MSE  */
int nhvr_loss_sum_abs_diff(const float* a, const float* b, int64_t n, double* acc, void* stream);  /* L1        */
int nhvr_loss_sum_sq_const(const float* a, float target, int64_t n, double* acc, void* stream);    /* LSGAN     */
/* uvp float [N][73][H][W]; dp_i int32 [N][H][W] (DensePose part, 0 = background); dp_uv float [N][2][H][W].
 * acc3[0] += sum over foreground of |u_k-U| + |v_k-V| (k = ground-truth part, u = clamp(.5*U+.5,0,1));
 * acc3[1] += foreground pixel count; acc3[2] += sum of 25-way cross-entropy of the part logits. */
int nhvr_loss_uv_prob(const float* uvp, const int32_t* dp_i, const float* dp_uv, int32_t N, int32_t H, int32_t W,
                      double* acc3, void* stream);
/* gradient w.r.t. a of  coef * mean(...)  for mode 0: (a-b)^2, 1: |a-b|, 2: (a-target)^2; times *grad_scale (device
 * scalar, nullable); grad_a = (accumulate ? grad_a : 0) + ... */
int nhvr_loss_pair_bwd(const float* a, const float* b, int64_t n, int32_t mode, float target, float coef,
                       const float* grad_scale, int32_t accumulate, float* grad_a, void* stream);
int nhvr_avgpool3s2_bwd(const float* grad_out, int32_t N, int32_t C, int32_t H, int32_t W, int32_t accumulate, float* grad_in,
                        void* stream);
/* d(w_uv*uv_loss + w_prob*prob_loss)/d uvp -> grad float [N][73][H][W], times *grad_scale (device scalar, nullable);
 * acc3 = the forward's sums (device): the foreground count is read on the device, no host round trip. */
int nhvr_loss_uv_prob_bwd(const float* uvp, const int32_t* dp_i, const float* dp_uv, int32_t N, int32_t H, int32_t W,
                          const double* acc3, float w_uv, float w_prob, const float* grad_scale, float* grad, void* stream);
/* acc += sum |cur - warp(prev, flow)|, flow float [N][2][H][W] in pixels (dx, dy), bilinear, border clamp. */
int nhvr_loss_temporal(const float* cur, const float* prev, const float* flow, int32_t N, int32_t C, int32_t H,
                       int32_t W, double* acc, void* stream);
/* gradient of coef * mean|cur - warp(prev, flow)| w.r.t. cur (prev = detached previous frame) */
int nhvr_loss_temporal_bwd(const float* cur, const float* prev, const float* flow, int32_t N, int32_t C, int32_t H,
                           int32_t W, float coef, const float* grad_scale, float* grad_cur, void* stream);
/* Adam over a flat fp32 bucket (pix2pixHD's torch.optim.Adam(lr 2e-4, betas (0.5, 0.999)); no weight decay, no amsgrad): one launch.
 * g is multiplied by grad_scale first (1 / world_size after a summing all-reduce); n % 4 == 0; step counts from 1. */
int nhvr_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                   float grad_scale, int32_t step, void* stream);
/* AvgPool2d(3, stride 2, padding 1, count_include_pad=False) between discriminator scales (pix2pixHD
 * MultiscaleDiscriminator.downsample): in float [N][C][H][W] -> out float [N][C][(H+1)/2][(W+1)/2]. */
int nhvr_avgpool3s2(const float* in, int32_t N, int32_t C, int32_t H, int32_t W, float* out, void* stream);
/* nn.MaxPool2d(2, 2) between the VGG19 stages of the perceptual loss (pix2pixHD VGGLoss, on unless --no_vgg_loss; README.md:101).
 * fp32 NCHW, output floor(H/2) x floor(W/2); the backward routes to the first maximum of each window (torch semantics). */
int nhvr_maxpool2(const float* in, int32_t N, int32_t C, int32_t H, int32_t W, float* out, void* stream);
int nhvr_maxpool2_bwd(const float* in, const float* grad_out, int32_t N, int32_t C, int32_t H, int32_t W, float* grad_in, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NHVR_H_ */
