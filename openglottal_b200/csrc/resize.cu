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

#include <cmath>

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

// How the threads of a block are laid over a band of output rows: `tcols` consecutive columns per
// row group (a whole number of warps, so that a warp's loads and stores are consecutive bytes or
// floats of one row), 256 / tcols row groups.
struct BandThreads {
    int tc, tr, tcols, trows;
};
__device__ __forceinline__ BandThreads band_threads(int DW) {
    BandThreads b;
    b.tcols = DW >= 256 ? 256 : ((DW + 31) & ~31);
    b.trows = 256 / b.tcols;
    b.tc = threadIdx.x % b.tcols;
    b.tr = threadIdx.x / b.tcols;      // >= trows: idle (tcols = 96, 160 ... do not divide 256)
    return b;
}

// The same arithmetic, one block per band of R output rows: the (index, coefficient) pairs of all
// columns and of the band's rows are computed once into shared memory (lin_pos is a double-precision
// division: two per pixel above), and a thread owns a COLUMN and walks the band's rows, so that the
// four byte gathers and the byte store of a warp are consecutive addresses of one row.
// tabx[0..DW) = {s0, s1, a0, a1}, taby[0..R) = {s0, s1, b0, b1}.
__global__ void __launch_bounds__(256)
resize_u8_linear_band_kernel(const uint8_t* __restrict__ src, int SH, int SW,
                             uint8_t* __restrict__ dst, int DH, int DW, int R) {
    extern __shared__ int4 band_u8[];
    const int4* tabx = band_u8;
    const int4* taby = band_u8 + DW;
    const int n = blockIdx.y, oy0 = blockIdx.x * R;
    const int nr = DH - oy0 < R ? DH - oy0 : R;
    for (int i = threadIdx.x; i < DW + nr; i += blockDim.x) {
        const Lin l = i < DW ? lin_x(i, SW, DW) : lin_y(oy0 + (i - DW), SH, DH);
        band_u8[i] = make_int4(l.s0, l.s1, coef11(__fsub_rn(1.f, l.f)), coef11(l.f));
    }
    __syncthreads();
    const BandThreads t = band_threads(DW);
    if (t.tr >= t.trows) return;
    const uint8_t* s = src + static_cast<size_t>(n) * SH * SW;
    uint8_t* d = dst + static_cast<size_t>(n) * DH * DW + static_cast<size_t>(oy0) * DW;
    for (int x = t.tc; x < DW; x += t.tcols) {
        const int4 tx = tabx[x];
        for (int r = t.tr; r < nr; r += t.trows) {
            const int4 ty = taby[r];
            const uint8_t* r0p = s + static_cast<size_t>(ty.x) * SW;
            const uint8_t* r1p = s + static_cast<size_t>(ty.y) * SW;
            const int r0 = r0p[tx.x] * tx.z + r0p[tx.y] * tx.w;
            const int r1 = r1p[tx.x] * tx.z + r1p[tx.y] * tx.w;
            const int v = (((ty.z * (r0 >> 4)) >> 16) + ((ty.w * (r1 >> 4)) >> 16) + 2) >> 2;
            d[static_cast<size_t>(r) * DW + x] = static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
    }
}

// src = 2 dst on both axes (cv2's 2x2 area mean): 8 source bytes of two rows per thread, 4 output
// bytes per store. SW % 8 == 0, src 8-byte and dst 4-byte aligned.
__global__ void __launch_bounds__(256)
resize_u8_area2x2_vec_kernel(const uint8_t* __restrict__ src, int SW, uint8_t* __restrict__ dst,
                             int DH, int DW) {
    const int n = blockIdx.y;
    const uint8_t* s = src + static_cast<size_t>(n) * (2 * DH) * SW;
    uint32_t* d = reinterpret_cast<uint32_t*>(dst + static_cast<size_t>(n) * DH * DW);
    const int quads = DW >> 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < DH * quads; i += gridDim.x * blockDim.x) {
        const int y = i / quads, xq = i - y * quads;
        const uint2 a = __ldg(reinterpret_cast<const uint2*>(s + static_cast<size_t>(2 * y) * SW) + xq);
        const uint2 b = __ldg(reinterpret_cast<const uint2*>(s + static_cast<size_t>(2 * y + 1) * SW) + xq);
        const uint32_t aw[2] = {a.x, a.y}, bw[2] = {b.x, b.y};
        uint32_t out = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = 2 * j;                           // first of the two source columns
            const uint32_t t = ((aw[k >> 2] >> (8 * (k & 3))) & 0xffu) +
                               ((aw[(k + 1) >> 2] >> (8 * ((k + 1) & 3))) & 0xffu) +
                               ((bw[k >> 2] >> (8 * (k & 3))) & 0xffu) +
                               ((bw[(k + 1) >> 2] >> (8 * ((k + 1) & 3))) & 0xffu);
            out |= ((t + 2u) >> 2) << (8 * j);
        }
        d[i] = out;
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

// The resizing case with the (index, fraction) pairs of the columns and rows in shared memory
// (tab[0..DW) columns, tab[DW..DW+DH) rows; .z holds the bits of the f32 fraction) and four mask
// bytes per store. Same products and sums in the same order as above. DW % 4 == 0, mask 4-byte
// aligned (or absent).
__global__ void __launch_bounds__(256)
prob_resize_mask_tab_kernel(const float* __restrict__ logits, int SH, int SW, int DH, int DW,
                            float threshold, uint8_t* __restrict__ mask, int32_t* __restrict__ area) {
    extern __shared__ int4 tab_f32[];
    __shared__ int scratch[8];
    for (int i = threadIdx.x; i < DW + DH; i += blockDim.x) {
        const Lin l = i < DW ? lin_x(i, SW, DW) : lin_y(i - DW, SH, DH);
        tab_f32[i] = make_int4(l.s0, l.s1, __float_as_int(l.f), 0);
    }
    __syncthreads();
    const int n = blockIdx.y;
    const float* z = logits + static_cast<size_t>(n) * SH * SW;
    uint32_t* m = mask ? reinterpret_cast<uint32_t*>(mask + static_cast<size_t>(n) * DH * DW) : nullptr;
    const int quads = DW >> 2;
    int cnt = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < DH * quads; i += gridDim.x * blockDim.x) {
        const int y = i / quads, x0 = (i - y * quads) << 2;
        const int4 ty = tab_f32[DW + y];
        const float fy = __int_as_float(ty.z);
        const float b0 = __fsub_rn(1.f, fy), b1 = fy;
        const float* r0p = z + static_cast<size_t>(ty.x) * SW;
        const float* r1p = z + static_cast<size_t>(ty.y) * SW;
        uint32_t out = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int4 tx = tab_f32[x0 + j];
            const float fx = __int_as_float(tx.z);
            const float a0 = __fsub_rn(1.f, fx), a1 = fx;
            const float r0 = __fadd_rn(__fmul_rn(sigmoidf(r0p[tx.x]), a0), __fmul_rn(sigmoidf(r0p[tx.y]), a1));
            const float r1 = __fadd_rn(__fmul_rn(sigmoidf(r1p[tx.x]), a0), __fmul_rn(sigmoidf(r1p[tx.y]), a1));
            const float v = __fadd_rn(__fmul_rn(r0, b0), __fmul_rn(r1, b1));
            if (v > threshold) {
                out |= 0xffu << (8 * j);
                ++cnt;
            }
        }
        if (m) m[i] = out;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0 && area) {
        int t = 0;
        for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) t += scratch[w];
        if (t) atomicAdd(area + n, t);
    }
}

// Up-scaling (the usual case: a 256 x 256 probability map back to the frame's size) reads every
// source pixel for several output pixels, and each read above costs a sigmoid. Here a block owns a
// band of R output rows: it evaluates the sigmoid ONCE per source pixel of the rows the band
// touches (a contiguous range: the row indices are monotone) into shared memory, then interpolates
// from there with the thread-per-column layout of resize_u8_linear_band_kernel. Same sigmoid
// values, same products and sums in the same order, hence the same masks.
// Shared memory: tabx[DW] {s0, s1, f}, taby[R] {s0, s1, f}, sig[rows_max * SW].
__global__ void __launch_bounds__(256)
prob_resize_mask_band_kernel(const float* __restrict__ logits, int SH, int SW, int DH, int DW, int R,
                             int rows_max, float threshold, uint8_t* __restrict__ mask,
                             int32_t* __restrict__ area) {
    extern __shared__ int4 band_f32[];
    __shared__ int scratch[8];
    const int4* tabx = band_f32;
    const int4* taby = band_f32 + DW;
    float* sig = reinterpret_cast<float*>(band_f32 + DW + R);
    const int n = blockIdx.y, oy0 = blockIdx.x * R;
    const int nr = DH - oy0 < R ? DH - oy0 : R;
    for (int i = threadIdx.x; i < DW + nr; i += blockDim.x) {
        const Lin l = i < DW ? lin_x(i, SW, DW) : lin_y(oy0 + (i - DW), SH, DH);
        band_f32[i] = make_int4(l.s0, l.s1, __float_as_int(l.f), 0);
    }
    __syncthreads();
    const int sy_lo = taby[0].x, rows = taby[nr - 1].y - sy_lo + 1;
    const bool staged = rows <= rows_max;     // holds by the host's choice of R; kept as a guard
    const float* z = logits + static_cast<size_t>(n) * SH * SW;
    if (staged) {
        const float* zb = z + static_cast<size_t>(sy_lo) * SW;
        for (int i = threadIdx.x; i < rows * SW; i += blockDim.x) sig[i] = sigmoidf(zb[i]);
    }
    __syncthreads();
    const BandThreads t = band_threads(DW);
    uint8_t* m = mask ? mask + static_cast<size_t>(n) * DH * DW + static_cast<size_t>(oy0) * DW : nullptr;
    int cnt = 0;
    if (t.tr < t.trows) {
        for (int x = t.tc; x < DW; x += t.tcols) {
            const int4 tx = tabx[x];
            const float fx = __int_as_float(tx.z);
            const float a0 = __fsub_rn(1.f, fx), a1 = fx;
            for (int r = t.tr; r < nr; r += t.trows) {
                const int4 ty = taby[r];
                const float fy = __int_as_float(ty.z);
                const float b0 = __fsub_rn(1.f, fy), b1 = fy;
                float p00, p01, p10, p11;
                if (staged) {
                    const float* q0 = sig + (ty.x - sy_lo) * SW;
                    const float* q1 = sig + (ty.y - sy_lo) * SW;
                    p00 = q0[tx.x], p01 = q0[tx.y], p10 = q1[tx.x], p11 = q1[tx.y];
                } else {
                    const float* q0 = z + static_cast<size_t>(ty.x) * SW;
                    const float* q1 = z + static_cast<size_t>(ty.y) * SW;
                    p00 = sigmoidf(q0[tx.x]), p01 = sigmoidf(q0[tx.y]);
                    p10 = sigmoidf(q1[tx.x]), p11 = sigmoidf(q1[tx.y]);
                }
                const float r0 = __fadd_rn(__fmul_rn(p00, a0), __fmul_rn(p01, a1));
                const float r1 = __fadd_rn(__fmul_rn(p10, a0), __fmul_rn(p11, a1));
                const float v = __fadd_rn(__fmul_rn(r0, b0), __fmul_rn(r1, b1));
                const bool on = v > threshold;
                if (m) m[static_cast<size_t>(r) * DW + x] = on ? 255 : 0;
                cnt += on ? 1 : 0;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0 && area) {
        int tot = 0;
        for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) tot += scratch[w];
        if (tot) atomicAdd(area + n, tot);
    }
}

constexpr size_t kMaxTabBytes = 40 * 1024;   // index tables in (default-limit) shared memory
constexpr int kBandRows = 32;                // output rows per block of the band kernels
inline bool aligned_to(const void* p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

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
    const long long quads = static_cast<long long>(DH) * (DW / 4);
    if (area2x2 && SW % 8 == 0 && aligned_to(src, 8) && aligned_to(dst, 4)) {
        grid.x = frame_chunks(quads, 4, n);
        resize_u8_area2x2_vec_kernel<<<grid, 256, 0, stream>>>(src, SW, dst, DH, DW);
    } else if (!area2x2 && sizeof(int4) * (static_cast<size_t>(DW) + kBandRows) <= kMaxTabBytes) {
        grid.x = (DH + kBandRows - 1) / kBandRows;
        resize_u8_linear_band_kernel<<<grid, 256, sizeof(int4) * (DW + kBandRows), stream>>>(
            src, SH, SW, dst, DH, DW, kBandRows);
    } else {
        resize_u8_linear_kernel<<<grid, 256, 0, stream>>>(src, SH, SW, dst, DH, DW, area2x2);
    }
    OGL_CUDA(cudaGetLastError());
    return 0;
}

int launch_prob_resize_mask(const float* logits, int n, int SH, int SW, int DH, int DW,
                            float threshold, uint8_t* mask, int32_t* area, cudaStream_t stream) {
    if (area) OGL_CUDA(cudaMemsetAsync(area, 0, sizeof(int32_t) * n, stream));
    dim3 grid(chunks_for(static_cast<long long>(DH) * DW), n);
    const size_t tab = sizeof(int4) * (static_cast<size_t>(DW) + DH);
    const bool identity = SH == DH && SW == DW;
    // band kernel: as many output rows per block (32, 16 ... 1) as keep the staged sigmoid rows within
    // 32 KB; taken when it evaluates at most half the sigmoids of the direct form (up-scaling)
    const double scale = static_cast<double>(SH) / DH;
    auto rows_for = [&](int R) { return static_cast<long long>(std::floor((R - 1) * scale)) + 4; };
    int R = kBandRows;
    while (R > 1 && rows_for(R) * SW * 4 > 32 * 1024) R >>= 1;
    const long long rows_max = rows_for(R);
    const size_t band_smem = sizeof(int4) * (static_cast<size_t>(DW) + R) + static_cast<size_t>(rows_max) * SW * 4;
    if (!identity && band_smem <= 47 * 1024 && rows_max * SW <= 2LL * R * DW) {
        grid.x = (DH + R - 1) / R;
        prob_resize_mask_band_kernel<<<grid, 256, band_smem, stream>>>(
            logits, SH, SW, DH, DW, R, static_cast<int>(rows_max), threshold, mask, area);
    } else if (!identity && DW % 4 == 0 && tab <= kMaxTabBytes && (!mask || aligned_to(mask, 4))) {
        grid.x = frame_chunks(static_cast<long long>(DH) * (DW / 4), 8, n);
        prob_resize_mask_tab_kernel<<<grid, 256, tab, stream>>>(logits, SH, SW, DH, DW, threshold,
                                                                mask, area);
    } else {
        prob_resize_mask_kernel<<<grid, 256, 0, stream>>>(logits, SH, SW, DH, DW, threshold, mask, area);
    }
    OGL_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ogl
