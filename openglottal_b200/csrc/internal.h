// Internal (non-ABI) declarations shared by the translation units of libopenglottal_b200.so.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <string>

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
    float b = 0.f;
    float logit_thr = 0.f;
    float* logits = nullptr;   // [B][H][W] or null
    uint8_t* mask = nullptr;   // [B][H][W] {0,255} or null
    int32_t* area = nullptr;   // [B] (must be zeroed by caller) or null
};

int conv_tc_init();  // resolves cuTensorMapEncodeTiled, sets smem attributes
int launch_conv_tc(const TcLayer& L, const __nv_bfloat16* src0, const __nv_bfloat16* src1, int B,
                   int H, int W, __nv_bfloat16* out, __nv_bfloat16* out_pool,
                   const HeadParams* head, int num_sms, cudaStream_t stream);

// stem: u8 / f32 gray -> conv3x3(1->32)+bias+ReLU in fp32 -> C8-planar bf16
struct StemWeights {  // passed by value: lives in the kernel-parameter constant bank
    float w[32 * 9];
    float b[32];
};
int launch_stem(const void* frames, int in_dtype, const StemWeights& sw, int B, int H, int W,
                __nv_bfloat16* out, cudaStream_t stream);

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

// debugging / unit-test helpers: layout conversion NCHW f32 <-> C8-planar bf16
int launch_nchw_to_c8(const float* in, __nv_bfloat16* out, int B, int C, int H, int W,
                      cudaStream_t stream);
int launch_c8_to_nchw(const __nv_bfloat16* in, float* out, int B, int C, int H, int W,
                      cudaStream_t stream);

}  // namespace ogl
