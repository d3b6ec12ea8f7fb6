// The two cv2.resize calls of the reference's per-frame wrapper, on the device and batched
// (SURVEY.md section 2.2, K9):
//
//   ogl_resize_u8_linear   /root/reference/openglottal/utils.py:234
//                          cv2.resize(frame_gray, (256, 256), interpolation=cv2.INTER_LINEAR), u8
//   ogl_prob_resize_mask   /root/reference/openglottal/utils.py:237-241
//                          sigmoid(logits) -> cv2.resize(prob, (W, H), INTER_LINEAR) (f32, skipped
//                          for 256x256) -> (prob > threshold) * 255, + features.py:238 area
//
// cv2's INTER_LINEAR (OpenCV 4.x modules/imgproc/src/resize.cpp, the generic path; restated and
// pinned against cv2 itself in oracle/resize_oracle.py):
//   position   f = float((d + 0.5) * (src / dst) - 0.5)  (double arithmetic, one cast), s = floor(f),
//              f -= s
//   horizontal s < 0 -> (s, f) = (0, 0);  s >= src - 1 -> (s, f) = (src - 1, 0)
//   vertical   the weights keep the unclamped f; the two ROW INDICES s, s + 1 are clipped to
//              [0, src - 1]
//   u8         weights as short: cvRound(w * 2048); rows = S[s] * a0 + S[s + 1] * a1 (int);
//              out = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2
//              -- except src = 2 dst on both axes, which cv2 turns into the 2x2 area mean
//              (a + b + c + d + 2) >> 2
//   f32        r = S[s] * a0 + S[s + 1] * a1, out = r0 * b0 + r1 * b1, each product and sum rounded
//              to f32 (no fused multiply-add)
// The u8 path is bit-exact with cv2 for every size pair tested; the f32 path is bit-exact with
// OpenCV's own code and within 2e-5 of the IPP routine cv2 dispatches to by default for f32.
#include "internal.h"

namespace ogl {

namespace {

struct Lin {
    int s0, s1;
    float f;
};

__device__ __forceinline__ float lin_pos(int d, int src, int dst) {
    const double scale = __ddiv_rn(static_cast<double>(src), static_cast<double>(dst));
    return static_cast<float>(
        __dadd_rn(__dmul_rn(__dadd_rn(static_cast<double>(d), 0.5), scale), -0.5));
}
__device__ __forceinline__ Lin lin_x(int d, int src, int dst) {
    float f = lin_pos(d, src, dst);
    int s = static_cast<int>(floorf(f));
    f = __fsub_rn(f, static_cast<float>(s));
    if (s < 0) {
        s = 0;
        f = 0.f;
    }
    if (s >= src - 1) {
        s = src - 1;
        f = 0.f;
    }
    return Lin{s, s + 1 < src ? s + 1 : src - 1, f};
}
__device__ __forceinline__ Lin lin_y(int d, int src, int dst) {
    float f = lin_pos(d, src, dst);
    const int s = static_cast<int>(floorf(f));
    f = __fsub_rn(f, static_cast<float>(s));
    auto clip = [src](int v) { return v < 0 ? 0 : (v > src - 1 ? src - 1 : v); };
    return Lin{clip(s), clip(s + 1), f};
}
// saturate_cast<short>(w * 2048): round half to even
__device__ __forceinline__ int coef11(float w) { return __float2int_rn(__fmul_rn(w, 2048.f)); }

__global__ void __launch_bounds__(256)
resize_u8_linear_kernel(const uint8_t* __restrict__ src, int SH, int SW, uint8_t* __restrict__ dst,
                        int DH, int DW, int area2x2) {
    const int n = blockIdx.y;
    const uint8_t* s = src + static_cast<size_t>(n) * SH * SW;
    uint8_t* d = dst + static_cast<size_t>(n) * DH * DW;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < DH * DW; i += gridDim.x * blockDim.x) {
        const int y = i / DW, x = i - y * DW;
        if (area2x2) {
            const uint8_t* q = s + static_cast<size_t>(2 * y) * SW + 2 * x;
            d[i] = static_cast<uint8_t>((q[0] + q[1] + q[SW] + q[SW + 1] + 2) >> 2);
            continue;
        }
        const Lin lx = lin_x(x, SW, DW), ly = lin_y(y, SH, DH);
        const int a0 = coef11(__fsub_rn(1.f, lx.f)), a1 = coef11(lx.f);
        const int b0 = coef11(__fsub_rn(1.f, ly.f)), b1 = coef11(ly.f);
        const uint8_t* r0p = s + static_cast<size_t>(ly.s0) * SW;
        const uint8_t* r1p = s + static_cast<size_t>(ly.s1) * SW;
        const int r0 = r0p[lx.s0] * a0 + r0p[lx.s1] * a1;
        const int r1 = r1p[lx.s0] * a0 + r1p[lx.s1] * a1;
        const int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
        d[i] = static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
    }
}

__device__ __forceinline__ float sigmoidf(float z) { return 1.f / (1.f + expf(-z)); }

// grid = (chunks, n). identity (DH, DW) == (SH, SW): the reference skips the resize.
__global__ void __launch_bounds__(256)
prob_resize_mask_kernel(const float* __restrict__ logits, int SH, int SW, int DH, int DW,
                        float threshold, uint8_t* __restrict__ mask, int32_t* __restrict__ area) {
    __shared__ int scratch[8];
    const int n = blockIdx.y;
    const float* z = logits + static_cast<size_t>(n) * SH * SW;
    uint8_t* m = mask ? mask + static_cast<size_t>(n) * DH * DW : nullptr;
    const bool identity = SH == DH && SW == DW;
    int cnt = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < DH * DW; i += gridDim.x * blockDim.x) {
        const int y = i / DW, x = i - y * DW;
        float v;
        if (identity) {
            v = sigmoidf(z[i]);
        } else {
            const Lin lx = lin_x(x, SW, DW), ly = lin_y(y, SH, DH);
            const float a0 = __fsub_rn(1.f, lx.f), a1 = lx.f;
            const float b0 = __fsub_rn(1.f, ly.f), b1 = ly.f;
            const float* r0p = z + static_cast<size_t>(ly.s0) * SW;
            const float* r1p = z + static_cast<size_t>(ly.s1) * SW;
            const float r0 = __fadd_rn(__fmul_rn(sigmoidf(r0p[lx.s0]), a0), __fmul_rn(sigmoidf(r0p[lx.s1]), a1));
            const float r1 = __fadd_rn(__fmul_rn(sigmoidf(r1p[lx.s0]), a0), __fmul_rn(sigmoidf(r1p[lx.s1]), a1));
            v = __fadd_rn(__fmul_rn(r0, b0), __fmul_rn(r1, b1));
        }
        const bool on = v > threshold;
        if (m) m[i] = on ? 255 : 0;
        cnt += on ? 1 : 0;
    }
    // block sum -> one atomic per block
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0 && area) {
        int t = 0;
        for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) t += scratch[w];
        if (t) atomicAdd(area + n, t);
    }
}

inline int chunks_for(long long work) {
    long long c = (work + 256 * 8 - 1) / (256 * 8);
    return static_cast<int>(c < 1 ? 1 : (c > 64 ? 64 : c));
}

}  // namespace

int launch_resize_u8_linear(const uint8_t* src, int n, int SH, int SW, uint8_t* dst, int DH, int DW,
                            cudaStream_t stream) {
    if (SH == DH && SW == DW) {
        OGL_CUDA(cudaMemcpyAsync(dst, src, static_cast<size_t>(n) * SH * SW, cudaMemcpyDeviceToDevice,
                                 stream));
        return 0;
    }
    const int area2x2 = (SH == 2 * DH && SW == 2 * DW) ? 1 : 0;
    dim3 grid(chunks_for(static_cast<long long>(DH) * DW), n);
    resize_u8_linear_kernel<<<grid, 256, 0, stream>>>(src, SH, SW, dst, DH, DW, area2x2);
    OGL_CUDA(cudaGetLastError());
    return 0;
}

int launch_prob_resize_mask(const float* logits, int n, int SH, int SW, int DH, int DW,
                            float threshold, uint8_t* mask, int32_t* area, cudaStream_t stream) {
    if (area) OGL_CUDA(cudaMemsetAsync(area, 0, sizeof(int32_t) * n, stream));
    dim3 grid(chunks_for(static_cast<long long>(DH) * DW), n);
    prob_resize_mask_kernel<<<grid, 256, 0, stream>>>(logits, SH, SW, DH, DW, threshold, mask, area);
    OGL_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ogl
