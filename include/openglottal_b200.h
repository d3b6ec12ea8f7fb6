/* openglottal_b200 -- C ABI of the B200-native U-Net-only hot path.
 *
 * The reference (hari-krishnan/openglottal, pure Python) has no FFI of its own; its seam for
 * this path is Python duck typing. Each entry point below names the reference interface it
 * replaces (paths relative to the reference repository root):
 *
 *   ogl_unet_create / ogl_unet_load_state   openglottal/cli.py:61-65
 *        UNet(1,1,(32,64,128,256)).to(device); load_state_dict(torch.load(...)); eval()
 *        state-dict layout: openglottal/models/unet.py:18-33,50-72 (118 tensors)
 *   ogl_unet_forward                        openglottal/models/unet.py:74-88 (UNet.forward)
 *        + openglottal/utils.py:235-241 (u8/255, sigmoid, > threshold -> {0,255} mask)
 *        + openglottal/features.py:238 (area = count(mask > 0))
 *   ogl_features / ogl_features_f64         openglottal/features.py:38-68 (_kinematic_features)
 *   ogl_bgr_to_gray                         openglottal/features.py:235 (cv2.COLOR_BGR2GRAY)
 *
 * Conventions: every function returns 0 on success and non-zero on failure; the message is
 * available from ogl_last_error() (thread-local). All `*_dev` pointers are device pointers on
 * the handle's device; the caller owns inputs, outputs and workspaces. Kernels are enqueued
 * on `stream` (a cudaStream_t passed as void*) and never synchronise the device. There is no
 * CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef OPENGLOTTAL_B200_H
#define OPENGLOTTAL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OGL_VERSION 100 /* 0.1.0 */

enum { OGL_DTYPE_U8 = 0, OGL_DTYPE_F32 = 1 };
/* BF16: bf16 operands on tcgen05 tensor cores, fp32 accumulate (the product path).
 * F32 : fp32 weights/activations on CUDA cores -- validation mode (logits within 1e-4 of the
 *       reference); same folding, tiling-independent arithmetic. Not a timed path.
 * F16 : the BF16 kernels compiled for f16 operands (same tcgen05 kind::f16 rate, same bytes;
 *       3 more mantissa bits: logits within 2e-2 of the reference everywhere; activations saturate
 *       at +-65504). Needs ogl_unet_prepare(h, OGL_PRECISION_F16) once after loading weights. */
enum { OGL_PRECISION_BF16 = 0, OGL_PRECISION_F32 = 1, OGL_PRECISION_F16 = 2 };

typedef struct ogl_unet ogl_unet;

/* Conv2d(3x3, bias=False) + BatchNorm2d, host pointers into the state dict. */
typedef struct {
    const float* weight;       /* [cout][cin][3][3]  "<prefix>.net.{0,3}.weight" */
    const float* bn_weight;    /* [cout]             "<prefix>.net.{1,4}.weight" */
    const float* bn_bias;      /* [cout]             "....bias"                  */
    const float* running_mean; /* [cout]             "....running_mean"          */
    const float* running_var;  /* [cout]             "....running_var"           */
} ogl_conv_bn;

/* ConvTranspose2d(2f, f, kernel_size=2, stride=2). */
typedef struct {
    const float* weight; /* [cin][cout][2][2]  "ups.{0,2,4,6}.weight" */
    const float* bias;   /* [cout]             "ups.{0,2,4,6}.bias"   */
} ogl_convt;

/* The 118-entry state dict of UNet(1,1,(32,64,128,256)) (num_batches_tracked ignored). */
typedef struct {
    ogl_conv_bn downs[4][2];   /* downs.i.net.{0,1} / downs.i.net.{3,4}          */
    ogl_conv_bn bottleneck[2]; /* bottleneck.net.{0,1} / .net.{3,4}              */
    ogl_convt up_t[4];         /* ups.0, ups.2, ups.4, ups.6                     */
    ogl_conv_bn up_c[4][2];    /* ups.1, ups.3, ups.5, ups.7  (.net.{0,1}/{3,4}) */
    const float* head_weight;  /* [1][32][1][1] */
    const float* head_bias;    /* [1] */
    float bn_eps;              /* 1e-5 */
} ogl_unet_state;

int ogl_version(void);
const char* ogl_last_error(void);

/* Creates a handle bound to CUDA device `device` (fails when there is none). */
int ogl_unet_create(ogl_unet** out, int device);
int ogl_unet_destroy(ogl_unet* h);

/* Folds BatchNorm (eval mode) into the convolutions in fp64, keeps an fp32 copy for the
 * validation path, and packs bf16 tensor-core operands. Host pointers; copies synchronously. */
int ogl_unet_load_state(ogl_unet* h, const ogl_unet_state* state);

/* Packs the operands a precision mode needs that ogl_unet_load_state did not (the f16 set); a
 * no-op for the other modes. Allocates and copies synchronously -- ogl_unet_forward never does. */
int ogl_unet_prepare(ogl_unet* h, int precision);

/* Bytes of device workspace ogl_unet_forward needs for n frames of h x w (h, w % 16 == 0). */
size_t ogl_unet_workspace_bytes(const ogl_unet* h, int n, int height, int width, int precision);

/* UNet forward + sigmoid/threshold + per-frame area for n gray frames [n][height][width]
 * (u8 in 0..255, or f32 already scaled). Any of logits_dev [n][h][w] f32, mask_dev [n][h][w]
 * u8 {0,255}, area_dev [n] int32 may be NULL. `threshold` is on the probability (0.5 in the
 * reference); the kernel compares the fp32 logit against logit(threshold).
 * Concurrency: a handle and a workspace are single-stream objects. The handle's device must be
 * the current device (checked). Calls on one handle must be serialised by the caller; two
 * forwards may overlap only if they use different handles AND different workspaces (the
 * activation tensors of a forward live in the workspace until its last kernel has run). The
 * Python layer keeps one handle and one workspace per UNet module, i.e. a module is driven from
 * one stream at a time. */
int ogl_unet_forward(ogl_unet* h, const void* frames_dev, int in_dtype, int n, int height,
                     int width, void* workspace_dev, size_t workspace_bytes, float* logits_dev,
                     uint8_t* mask_dev, int32_t* area_dev, float threshold, int precision,
                     void* stream);

/* Optional per-launch timing of the bf16 path (CUDA events on the caller's stream between the
 * launches of every forward while enabled; used by bench.py for the roofline numbers).
 * ogl_unet_layer_times synchronises on and averages the (up to 16) most recent profiled forwards,
 * so a caller can time its own steady-state loop and read the per-launch times afterwards.
 * ogl_unet_launch_count / _name describe the launches of the most recent bf16 forward, each
 * named after the reference modules (openglottal/models/unet.py:50-72) it computes. */
int ogl_unet_set_profiling(ogl_unet* h, int enable);
int ogl_unet_layer_times(ogl_unet* h, float* ms_out, int capacity, int* count_out);
int ogl_unet_launch_count(const ogl_unet* h);
const char* ogl_unet_launch_name(const ogl_unet* h, int index);

/* Kernel schedule of the full-resolution level of the bf16 path. 1 (default): space-to-depth
 * GEMMs with ConvTranspose2d ups.6 composed into ups.7.net.0 (17 launches with the composed
 * decoder, else 20); 0: the direct per-tap form used at the other levels (2 more launches). Same
 * results within bf16 rounding. */
int ogl_unet_set_schedule(ogl_unet* h, int s2d_level0);

/* Decoder levels 1-3. 1 (default): every ConvTranspose2d (ups.0, ups.2, ups.4; unet.py:82) is
 * composed into the conv that follows it (ups.{1,3,5}.net.0), as the space-to-depth schedule does
 * for ups.6 -- 17 launches, no `up` tensor is ever written; 0: separate transposed-conv launches
 * and two-source convs (20 launches). Same results within bf16 rounding. */
int ogl_unet_set_compose(ogl_unet* h, int enable);

/* With u8 frames and the space-to-depth schedule, downs.0.net.0 (the Cin = 1 stem) can be computed
 * inside the downs.0.net.3 kernel, so that its output never touches HBM. 0: separate stem kernel;
 * 1: in-kernel on the CUDA cores in fp32 -- same results as 0 bit for bit; 2: in-kernel as a GEMM
 * on the tensor cores (u8 taps exact in bf16, weights / 255 and bias split hi + lo in bf16, fp32
 * accumulation) -- stem outputs within ~2^-17 relative of mode 1 before their rounding to bf16,
 * logits within the bf16 noise of the path; 3 (default): the same GEMM with 16 instead of 8 stem
 * warps and the im2col operand in f16 (u8 taps exact there too), built by byte permutes. */
int ogl_unet_set_fused_stem(ogl_unet* h, int enable);

/* CTA pairs for the conv3x3 layers with Cout >= 64: 1 = one CTA per tile; 2 = two CTAs of a
 * cluster share one 256-row tcgen05.mma.cta_group::2 and each stages half of the weights
 * (used when a launch has at least one tile per SM); 3 = pairs whenever a launch has two
 * tiles (unit tests). Same results bit for bit. */
int ogl_unet_set_cta_pairs(ogl_unet* h, int mode);

/* Measurement aid: launch `launch_index` (0-based position in ogl_unet_launch_name's list) of the
 * bf16 forward is enqueued `times` times (1..64) instead of once, with the same inputs and
 * outputs; launch_index -1 (default) repeats nothing. Used by scripts/layer_energy.py to take the
 * energy of one launch from the board's energy counter. The area vector is garbage when the
 * repeated launch is the head's (its atomics accumulate). */
int ogl_unet_set_repeat(ogl_unet* h, int launch_index, int times);

/* Kinematic features of an area waveform of n >= 2 samples.
 * out8_dev: {area_mean, area_std, area_range, open_quotient, f0, periodicity, cv, peak_bin};
 * flags2_dev: {is_silent (reference returns None), f0_is_none (peak in first bin)}. */
size_t ogl_features_workspace_bytes(int64_t n);
int ogl_features(const int32_t* area_dev, int64_t n, double* out8_dev, int32_t* flags2_dev,
                 void* workspace_dev, size_t workspace_bytes, void* stream);
int ogl_features_f64(const double* area_dev, int64_t n, double* out8_dev, int32_t* flags2_dev,
                     void* workspace_dev, size_t workspace_bytes, void* stream);

/* cv2.COLOR_BGR2GRAY on interleaved u8 BGR pixels: (3735 B + 19235 G + 9798 R + 16384) >> 15. */
int ogl_bgr_to_gray(const uint8_t* bgr_dev, uint8_t* gray_dev, int64_t pixels, void* stream);

/* ---- reference-resize mode of the per-frame wrapper, openglottal/utils.py:234-241 ----------------
 * ogl_resize_u8_linear: dst[i] = cv2.resize(src[i], (dst_w, dst_h), interpolation=cv2.INTER_LINEAR)
 *   for n u8 gray frames (utils.py:234 squashes every frame to 256 x 256); bit-exact with cv2
 *   (11-bit fixed-point weights; the 2x2 area mean cv2 substitutes when both axes halve; a copy
 *   when the sizes are equal).
 * ogl_prob_resize_mask: utils.py:237-241 after the forward pass: prob = sigmoid(logits) [n][src_h]
 *   [src_w]; when (dst_h, dst_w) differs it is resized with cv2's f32 INTER_LINEAR arithmetic;
 *   mask = (prob > threshold) * 255 [n][dst_h][dst_w] (or NULL); area[i] = count(mask[i] > 0)
 *   (features.py:238; or NULL). Only masks and areas leave the device. */
int ogl_resize_u8_linear(const uint8_t* src_dev, int n, int src_h, int src_w, uint8_t* dst_dev,
                         int dst_h, int dst_w, void* stream);
int ogl_prob_resize_mask(const float* logits_dev, int n, int src_h, int src_w, int dst_h,
                         int dst_w, float threshold, uint8_t* mask_dev, int32_t* area_dev,
                         void* stream);

/* ---- callers around the U-Net (all bit-exact, integer work) ---------------------------------
 * Detection-gated area, openglottal/features.py:240-245:
 *   area[i] = count(mask[i][y1:y2, x1:x2] > 0) with Python slice semantics for the box
 *   {x1, y1, x2, y2} = boxes_dev[4 i ..]; has_box_dev[i] == 0 (detector returned None) gives 0.
 *   has_box_dev may be NULL (every frame has a box). */
int ogl_mask_area_boxes(const uint8_t* mask_dev, int n, int height, int width,
                        const int32_t* boxes_dev, const uint8_t* has_box_dev, int32_t* area_dev,
                        void* stream);

/* yolo-crop+unet pipeline, scripts/infer.py:222-248. geom_dev holds 8 int32 per frame:
 * {x1, y1, x2, y2 (crop bounds, 0 <= x1 < x2 <= width ...), pad_top, pad_left, content_h,
 * content_w} as openglottal/utils.py:103-131 (letterbox_with_info) computes them; content_h == 0
 * marks a frame without a usable box.
 *   ogl_letterbox_crops : out[i] = letterbox(gray[i][y1:y2, x1:x2], size) (cv2 INTER_NEAREST,
 *                         zero padding), out_dev [n][size][size] u8.
 *   ogl_unletterbox_area: mask_orig = unletterbox(mask_cs[i]) (utils.py:170-186, INTER_NEAREST)
 *                         at the crop's size; area[i] = count(mask_orig > 0); full_mask_dev
 *                         [n][height][width] (or NULL) = zeros with mask_orig pasted at the box. */
int ogl_letterbox_crops(const uint8_t* gray_dev, int n, int height, int width,
                        const int32_t* geom_dev, int size, uint8_t* out_dev, void* stream);
int ogl_unletterbox_area(const uint8_t* mask_cs_dev, int n, int size, const int32_t* geom_dev,
                         int height, int width, uint8_t* full_mask_dev, int32_t* area_dev,
                         void* stream);

/* Batched evaluation, openglottal/utils.py:191-206: counts_dev[3 i ..] = {|pred & gt|, |pred|,
 * |gt|} over (value > 0); dice = 2 I / (P + G), iou = I / (P + G - I), 1.0 when empty. */
int ogl_mask_overlap_counts(const uint8_t* pred_dev, const uint8_t* gt_dev, int n, int64_t pixels,
                            int32_t* counts_dev, void* stream);

/* Unit-test hook: one tensor-core layer on fp32 NCHW device tensors (converted to the bf16
 * kernel layout internally). kind: 0 conv3x3+bias+ReLU, 1 same + 2x2 max-pool (out_pool_dev),
 * 3 ConvTranspose2d k2 s2 (+bias). src1_dev/c1 describe the second concat source (or NULL/0).
 * weight_host is [cout][c0+c1][3][3] (conv) or [c0][cout][2][2] (convT), bias_host [cout]. */
int ogl_debug_tc_layer(ogl_unet* h, int kind, const float* src0_dev, int c0, const float* src1_dev,
                       int c1, const float* weight_host, const float* bias_host, int cout, int n,
                       int height, int width, float* out_dev, float* out_pool_dev, void* stream);

/* Unit-test hook for the space-to-depth layers of the full-resolution level (Cout = 32).
 * src_dev [n][cin_s][H][W] f32; below_dev [n][64][H/2][W/2] f32 with wt_host [64][32][2][2] and
 * bt_host [32] (ConvTranspose2d composed in front of the conv; all three NULL for a plain
 * conv); w3_host [32][cin_s (+32)][3][3], b3_host [32]. kind: 0 conv+bias+ReLU, 1 + max-pool. */
int ogl_debug_s2d_layer(ogl_unet* h, int kind, const float* src_dev, int cin_s,
                        const float* below_dev, const float* w3_host, const float* b3_host,
                        const float* wt_host, const float* bt_host, int n, int height, int width,
                        float* out_dev, float* out_pool_dev, void* stream);


/* Unit-test hook for the composed decoder layer of levels 1-3 (upcat_tc.cu): out = relu(conv3x3(
 * cat([skip, conv_transpose2d(below, wt, bt, stride 2)]), w3) + b3) for skip_dev [n][f][H][W] f32,
 * below_dev [n][2f][H/2][W/2] f32, w3_host [f][2f][3][3], b3_host [f], wt_host [2f][f][2][2],
 * bt_host [f]; f in {64, 128, 256}; out_dev [n][f][H][W] f32. */
int ogl_debug_upcat_layer(ogl_unet* h, const float* skip_dev, const float* below_dev,
                          const float* w3_host, const float* b3_host, const float* wt_host,
                          const float* bt_host, int f, int n, int height, int width, float* out_dev,
                          void* stream);

/* Host-only (no device needed): the bf16 operands build_upcat_host packs for that layer, for CPU
 * emulation in the tests. wskip_out [f/N][f/32][9][4][N][8] and wbelow_out [f/N][2f/32][16][4][N][8]
 * (N = min(f, 128); pair index = (px * 2 + py) * 4 + oyi * 2 + oxi), their CTA-pair forms
 * [..][rank][..][N/2][8], btab_out [3][3][f]. Any output pointer may be NULL. */
int ogl_debug_upcat_program(const float* w3_host, const float* b3_host, const float* wt_host,
                            const float* bt_host, int f, uint16_t* wskip_out, uint16_t* wskip_pair_out,
                            uint16_t* wbelow_out, uint16_t* wbelow_pair_out, float* btab_out);

/* Host-only (no device needed): the MMA program build_s2d_host makes for a space-to-depth
 * layer, for CPU emulation in the tests. ops_out: 4 x uint32 per op {a_off | dcol << 16 |
 * src << 24 | accumulate << 25, b_off16 | N << 16, idesc, 0}; stages_out: 3 ints per stage
 * {source (0 S2D, 1 below), first 8-channel plane, end op}; btab_out [3][3][32]. Any output
 * pointer may be NULL to query sizes. */
int ogl_debug_s2d_program(const float* w3_host, const float* b3_host, int cin_s,
                          const float* wt_host, const float* bt_host, uint8_t* wblob_out,
                          size_t wblob_capacity, size_t* wblob_bytes, uint32_t* ops_out,
                          int ops_capacity, int* n_ops, int* stages_out, int* n_stages,
                          float* btab_out);

#ifdef __cplusplus
}
#endif
#endif /* OPENGLOTTAL_B200_H */
