// Internal (non-ABI) declarations shared by the translation units of libopenglottal_b200.so.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstddef>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <string>
#include <vector>

namespace ogl {

void set_error(const std::string& msg);
int fail(const std::string& msg);        // records msg, returns OGL_ERR (1)
int fail_cuda(cudaError_t e, const char* what);

#define OGL_CUDA(expr)                                        \
    do {                                                      \
        cudaError_t _e = (expr);                              \
        if (_e != cudaSuccess) return ::ogl::fail_cuda(_e, #expr); \
    } while (0)

// ------------------------------------------------------------------ layouts
// bf16 activations live in HBM as "C8-planar": [frame][C/8][H][W][8 channels], i.e. each
// 8-channel group is a plane whose pixels are 16 B apart. One TMA box
// (18 px * 8 ch, 18 rows, 4 planes) is then exactly the K-major SWIZZLE_NONE UMMA operand
// layout for a 16x16-pixel tile with its 3x3 halo, and a tap shift (dy,dx) is a plain
// 16 B-granular start-address offset.

enum Epilogue : int {
    EPI_RELU = 0,       // bias + ReLU -> bf16
    EPI_RELU_POOL = 1,  // bias + ReLU -> bf16 (skip tensor) and its 2x2 max-pool
    EPI_HEAD = 2,       // bias + ReLU, 1x1 head on the fp32 values, threshold, area popcount
    EPI_CONVT = 3       // ConvTranspose2d(k=2,s=2): bias, pixel-shuffle store, no ReLU
};

struct TcLayer {
    __nv_bfloat16* wpack = nullptr;  // [npass][Cin/32][taps][4][N][8] bf16 (device)
    __nv_bfloat16* wpack2 = nullptr; // CTA-pair form [npass][Cin/32][rank 2][taps][4][N/2][8], or null
    float* bias = nullptr;           // [cout] fp32 (device)
    int cin0 = 0, cin1 = 0;          // channels of source 0 (skip / only) and source 1 (up)
    int cout = 0;                    // channels of the output tensor
    int taps = 9;                    // 9 = conv3x3, 1 = convT (N_total = 4*cout)
    int N = 0;                       // MMA N per pass
    int npass = 0;                   // N_total / N
    int epi = EPI_RELU;
};

struct HeadParams {
    const float* w = nullptr;  // [32] device
    const float* w_host = nullptr;  // the same on the host (s2d kernels take it as a parameter)
    float b = 0.f;
    float logit_thr = 0.f;
    float* logits = nullptr;   // [B][H][W] or null
    uint8_t* mask = nullptr;   // [B][H][W] {0,255} or null
    int32_t* area = nullptr;   // [B] (must be zeroed by caller) or null
};

int conv_tc_init();  // resolves cuTensorMapEncodeTiled, sets smem attributes
int conv_tc_init_f16();   // the f16-operand twins (conv_tc_f16.cu, s2d_tc_f16.cu): same arguments,
                          // tensors and weights hold f16 instead of bf16
int launch_conv_tc_f16(const TcLayer& L, const __nv_bfloat16* src0, const __nv_bfloat16* src1, int B,
                       int H, int W, __nv_bfloat16* out, __nv_bfloat16* out_pool,
                       const HeadParams* head, int num_sms, cudaStream_t stream, int cta_group = 1,
                       bool reverse = false, bool out_s2d = false);
int launch_conv_tc(const TcLayer& L, const __nv_bfloat16* src0, const __nv_bfloat16* src1, int B,
                   int H, int W, __nv_bfloat16* out, __nv_bfloat16* out_pool,
                   const HeadParams* head, int num_sms, cudaStream_t stream, int cta_group = 1,
                   bool reverse = false,   // reverse: tiles from the last frame to the first
                   bool out_s2d = false);  // `out` (not the pooled tensor) is written space-to-depth:
                                           // [frame][C/8][phase][H/2][W/2][8], the layout upcat_tc reads

// stem: u8 / f32 gray -> conv3x3(1->32)+bias+ReLU in fp32 -> C8-planar bf16
struct StemWeights {  // passed by value: lives in the kernel-parameter constant bank
    float w[32 * 9];
    float b[32];
};
// Weights of the u8 stem as channel pairs: wp[co/2][tap] = (w[co][tap], w[co+1][tap]) / 255 and
// bp[co/2] = (b[co], b[co+1]), so that one packed FFMA2 (fma.rn.f32x2, two IEEE FMAs per
// instruction with the pair taken from uniform registers) advances two output channels of a pixel
// and its 64-bit result is exactly the bf16x2 the store packs. Bit-identical to scalar fmaf.
// The 1/255 of utils.py:235 is folded into the weights in fp64 and rounded once.
struct StemPairs {
    float2 wp[16 * 9];
    float2 bp[16];
};
inline StemPairs make_stem_pairs(const StemWeights& sw) {
    StemPairs sp;
    for (int cp = 0; cp < 16; ++cp) {
        for (int k = 0; k < 9; ++k) {
            sp.wp[cp * 9 + k].x = static_cast<float>(static_cast<double>(sw.w[(2 * cp) * 9 + k]) / 255.0);
            sp.wp[cp * 9 + k].y = static_cast<float>(static_cast<double>(sw.w[(2 * cp + 1) * 9 + k]) / 255.0);
        }
        sp.bp[cp].x = sw.b[2 * cp];
        sp.bp[cp].y = sw.b[2 * cp + 1];
    }
    return sp;
}
#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b,
                                                   unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long dup2(float v) {
    const unsigned long long u = __float_as_uint(v);
    return u | (u << 32);
}
__device__ __forceinline__ unsigned long long as_u64(float2 v) {
    return static_cast<unsigned long long>(__float_as_uint(v.x)) |
           (static_cast<unsigned long long>(__float_as_uint(v.y)) << 32);
}
#endif

// ------------------------------------------------------------------ full-resolution level
// The tensor-core convs at full resolution (Cout = 32) run on tensors stored "space-to-depth"
// (S2D): [frame][C/8][phase = (y&1)*2 + (x&1)][H/2][W/2][8]. A GEMM row is a half-resolution
// position and the accumulator columns are (output phase, cout) = 128, so one input phase at
// one half-resolution offset feeds 1, 2 or 4 output phases with ONE MMA of N = 32/64/96/128
// (16 MMAs per 16-channel slab instead of 36 N = 32 ones whose shared-memory operand reads
// bound the tensor pipe). The ConvTranspose2d(k2,s2) in front of the
// decoder's first conv is composed into it: its 2x2 phases are exactly the S2D phases, so
// conv3x3(up) becomes 9 half-resolution offsets of the 64-channel tensor below with weights
// multiplied at load time (fp64, rounded once to bf16).
struct S2dOp {        // one tcgen05.mma of the per-tile program, as the host packer lists it (the
                      // kernel unrolls the same table at compile time; tests emulate this list)
    uint32_t w0;      // a_off (16-B units inside the stage) | dcol << 16 | src << 24 | acc << 25
    uint32_t b_lo;    // B offset in the weight blob (16-B units) | (LBO >> 4) << 16
    uint32_t idesc;   // instruction descriptor (carries N)
    uint32_t pad;
};
// OGL_DBG switches parts of the tensor-core kernels OFF for timing experiments (results are
// garbage): honoured only together with OGL_EXPERIMENT=1, so that a stray variable in a production
// environment cannot silently corrupt masks and areas.
inline int experiment_dbg() {
    const char* dbg = getenv("OGL_DBG");
    const char* on = getenv("OGL_EXPERIMENT");
    return (dbg && on && atoi(on) != 0) ? atoi(dbg) : 0;
}

// Experiment knobs (ring depths, issuer counts, ...) come from the environment, read once per process.
// A value outside [lo, hi] is refused (the default is used and one line goes to stderr): a stray or
// mistyped variable must not be able to build a ring the kernels' barrier protocol cannot run.
inline int env_knob(const char* name, int dflt, int lo, int hi) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    char* end = nullptr;
    const long x = strtol(v, &end, 10);
    if (end == v || *end != '\0' || x < lo || x > hi) {
        fprintf(stderr, "openglottal_b200: %s=%s ignored (allowed %d..%d, default %d)\n", name, v, lo,
                hi, dflt);
        return dflt;
    }
    return static_cast<int>(x);
}

// grid.x of the per-frame byte kernels (grid.y = frame): `units` work items per frame (16-byte words,
// 4-pixel groups ...), `per_thread` of them per thread of a 256-thread block when the batch fills
// the GPU, fewer per thread (more blocks) when it does not.
inline int frame_chunks(long long units, int per_thread, int n) {
    long long c = (units + 256LL * per_thread - 1) / (256LL * per_thread);
    const long long cmax = (units + 255) / 256;
    while (c < cmax && c * n < 148 * 8) c *= 2;
    if (c > cmax) c = cmax;
    return static_cast<int>(c < 1 ? 1 : (c > 64 ? 64 : c));
}

constexpr int kS2dMaxStages = 8;
struct S2dLayer {
    uint8_t* wblob = nullptr;  // device: B tiles in op order
    uint8_t* wblob2 = nullptr; // device: the same for a CTA pair (rank 0's column halves, then rank 1's)
    uint32_t wbytes = 0;
    int n_stages = 0;          // activation stages (TMA boxes) per tile
    int stage_src[kS2dMaxStages] = {0};     // 0: S2D source, 1: plain half-resolution source
    int stage_plane0[kS2dMaxStages] = {0};  // first 8-channel plane of the box
    float* btab = nullptr;     // [3][3][32] device: bias per (row class, column class)
    float bias_host[32] = {0}; // bias of interior pixels (btab[1][1]), passed as kernel parameter
    int cin_s = 0, cin_b = 0;
    int epi = EPI_RELU;        // EPI_RELU / EPI_RELU_POOL / EPI_HEAD
};
struct S2dHost {               // host-side result of build_s2d_host, uploaded by the caller
    std::vector<uint8_t> wblob, wblob_pair;
    std::vector<S2dOp> ops;
    int n_stages = 0;
    int stage_src[kS2dMaxStages] = {0}, stage_plane0[kS2dMaxStages] = {0},
        stage_op_end[kS2dMaxStages] = {0};
    std::vector<float> btab;
};
// w3: folded conv weights [32][cin_s + (wt ? 32 : 0)][3][3], b3 [32]; wt: ConvTranspose2d weights
// [cin_b][32][2][2] with bias bt [32], or null.
int build_s2d_host(const float* w3, const float* b3, int cin_s, const float* wt, const float* bt,
                   int cin_b, S2dHost* out, bool f16 = false);   // f16: pack f16 instead of bf16
int s2d_tc_init();
int s2d_tc_init_f16();
// src_s2d: [B][cin_s/8][4][H/2][W/2][8]; below: [B][cin_b/8][H/2][W/2][8] or null. H, W = full res.
int launch_s2d_tc(const S2dLayer& L, const __nv_bfloat16* src_s2d, const __nv_bfloat16* below,
                  int B, int H, int W, __nv_bfloat16* out_s2d, __nv_bfloat16* out_pool,
                  const HeadParams* head, int num_sms, cudaStream_t stream, int cta_group = 1,
                  const uint8_t* stem_frames = nullptr, const StemWeights* stem = nullptr,
                  bool reverse = false, const uint8_t* stem_tc_blob = nullptr,
                  int stem_tc_warps = 8);
int launch_s2d_tc_f16(const S2dLayer& L, const __nv_bfloat16* src_s2d, const __nv_bfloat16* below,
                      int B, int H, int W, __nv_bfloat16* out_s2d, __nv_bfloat16* out_pool,
                      const HeadParams* head, int num_sms, cudaStream_t stream, int cta_group = 1,
                      const uint8_t* stem_frames = nullptr, const StemWeights* stem = nullptr,
                      bool reverse = false, const uint8_t* stem_tc_blob = nullptr,
                      int stem_tc_warps = 8);
// (stem_tc_blob != null, from build_stem_tc_blob: the fused stem runs on the tensor cores, with 8
//  stem warps and a bf16 im2col operand, or 16 warps and an f16 operand in the K order of
//  build_stem_tc_blob(.., true))
int build_stem_tc_blob(const StemWeights& sw, std::vector<uint8_t>* out, bool k_order3 = false,
                       bool f16 = false);
// (stem_frames != null: downs.0.net.3 with the stem fused in -- the A operand is computed from
//  the u8 frames [B][H][W] inside the kernel and src_s2d is not read)
// cuTensorMapEncodeTiled for a bf16 tensor (conv_tc.cu owns the driver entry point)
int encode_bf16_map(void* tensor_map, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box);
int encode_map(void* tensor_map, const void* base, int rank, const uint64_t* dims,
               const uint64_t* strides_bytes, const uint32_t* box, bool u8);

// ------------------------------------------------------------------ decoder levels 1-3
// ConvTranspose2d composed into the conv that follows it (upcat_tc.cu): the skip tensor is read in
// the space-to-depth layout, the tensor below plain; weights per K block: 9 taps of the skip half,
// 16 (output phase, half-resolution offset) pairs of the composed half.
struct UpcatLayer {
    uint8_t* wskip = nullptr;    // [pass][f/32][9][4][N][8]
    uint8_t* wskip2 = nullptr;   // CTA-pair form [pass][f/32][rank][9][4][N/2][8]
    uint8_t* wbelow = nullptr;   // [pass][2f/32][16][4][N][8]
    uint8_t* wbelow2 = nullptr;  // [pass][2f/32][rank][16][4][N/2][8]
    float* bias = nullptr;       // [f]: interior pixels (= btab[1][1])
    float* btab = nullptr;       // [3][3][f]: bias per (row class, column class)
    int f = 0, N = 0, npass = 0;
};
struct UpcatHost {
    std::vector<uint16_t> wskip, wskip_pair, wbelow, wbelow_pair;
    std::vector<float> btab;
    int f = 0, N = 0, npass = 0;
};
int build_upcat_host(const float* w3, const float* b3, const float* wt, const float* bt, int f,
                     bool f16, UpcatHost* out);
int upcat_tc_init();
int upcat_tc_init_f16();
int launch_upcat_tc(const UpcatLayer& L, const __nv_bfloat16* skip_s2d, const __nv_bfloat16* below,
                    int B, int H, int W, __nv_bfloat16* out, int num_sms, cudaStream_t stream,
                    int cta_group = 1);
int launch_upcat_tc_f16(const UpcatLayer& L, const __nv_bfloat16* skip_s2d, const __nv_bfloat16* below,
                        int B, int H, int W, __nv_bfloat16* out, int num_sms, cudaStream_t stream,
                        int cta_group = 1);

int launch_stem(const void* frames, int in_dtype, const StemWeights& sw, int B, int H, int W,
                __nv_bfloat16* out, bool s2d, cudaStream_t stream, bool f16 = false);

// fp32 validation path (NCHW fp32 activations, FFMA kernels)
int launch_f32_input(const void* frames, int in_dtype, int64_t count, float* out,
                     cudaStream_t stream);
int launch_f32_conv3x3(const float* src0, int c0, const float* src1, int c1, const float* w,
                       const float* b, float* out, int B, int cout, int H, int W, int relu,
                       cudaStream_t stream);
int launch_f32_maxpool(const float* in, float* out, int BC, int H, int W, cudaStream_t stream);
int launch_f32_convt(const float* in, const float* w, const float* b, float* out, int B, int cin,
                     int cout, int H, int W, cudaStream_t stream);
int launch_f32_head(const float* in, const float* w, float b, float thr, int B, int C, int H,
                    int W, float* logits, uint8_t* mask, int32_t* area, cudaStream_t stream);

// features on the int32 area vector
size_t features_workspace_bytes(int64_t n);
int launch_features(const int32_t* area, int64_t n, double* out8, int32_t* flags2, void* ws,
                    size_t ws_bytes, cudaStream_t stream);

// mask / frame operators around the U-Net (frame_ops.cu); n <= 65535 per launch
int launch_mask_area_boxes(const uint8_t* mask, int n, int H, int W, const int32_t* boxes,
                           const uint8_t* has_box, int32_t* area, cudaStream_t stream);
int launch_letterbox_crops(const uint8_t* gray, int n, int H, int W, const int32_t* geom, int size,
                           uint8_t* out, cudaStream_t stream);
int launch_unletterbox_area(const uint8_t* mask_cs, int n, int size, const int32_t* geom, int H,
                            int W, uint8_t* full, int32_t* area, cudaStream_t stream);
int launch_overlap_counts(const uint8_t* pred, const uint8_t* gt, int n, long long pixels,
                          int32_t* counts, cudaStream_t stream);

// the reference wrapper's two cv2.resize calls (resize.cu); n <= 65535 per launch
int launch_resize_u8_linear(const uint8_t* src, int n, int SH, int SW, uint8_t* dst, int DH, int DW,
                            cudaStream_t stream);
int launch_prob_resize_mask(const float* logits, int n, int SH, int SW, int DH, int DW,
                            float threshold, uint8_t* mask, int32_t* area, cudaStream_t stream);

// debugging / unit-test helpers: layout conversion NCHW f32 <-> C8-planar bf16
// (s2d = true: the space-to-depth form of the same tensor)
int launch_nchw_to_c8(const float* in, __nv_bfloat16* out, int B, int C, int H, int W,
                      cudaStream_t stream, bool s2d = false);
int launch_c8_to_nchw(const __nv_bfloat16* in, float* out, int B, int C, int H, int W,
                      cudaStream_t stream, bool s2d = false);

}  // namespace ogl
