// CUDA-core kernels: the Cin=1 stem of the bf16 path, the fp32 validation path
// (precision_mode = OGL_PRECISION_F32: fp32 weights/activations, FFMA, NCHW), and the
// layout converters used by the unit tests.
//
// Reference semantics restated here:
//   utils.py:235       x = u8.astype(float32) / 255.0
//   unet.py:24-29      y = relu(bn(conv3x3(x)))  (BN folded into W', b' by the host)
//   unet.py:59,79      max_pool2d(2, 2)
//   unet.py:69,82      out[n,co,2y+dy,2x+dx] = b[co] + sum_ci x[n,ci,y,x] * W[ci,co,dy,dx]
//   unet.py:86         cat([skip, up], dim=1)  -> two source pointers
//   unet.py:72,88      logits = conv1x1(x) + b;  utils.py:237,241 sigmoid(z) > thr <=> z > logit(thr)
//   features.py:238    area = count(mask > 0)
#include "internal.h"
#include "ptx.cuh"

namespace ogl {

namespace {

__device__ __forceinline__ float load_gray(const void* frames, int in_dtype, size_t idx) {
    if (in_dtype == 0) {
        const float v = static_cast<float>(static_cast<const uint8_t*>(frames)[idx]);
        return __fdiv_rn(v, 255.0f);  // same rounding as numpy float32 / 255.0
    }
    return static_cast<const float*>(frames)[idx];
}

// Element offset of pixel (y, x) inside one 8-channel group of a [H][W][8] plane, or of its
// space-to-depth form [phase][H/2][W/2][8] (internal.h); the group stride is H*W*8 for both.
__device__ __forceinline__ size_t pix_off(int y, int x, int H, int W, int s2d) {
    if (s2d) {
        const size_t hw2 = static_cast<size_t>(H >> 1) * (W >> 1);
        return ((((y & 1) * 2 + (x & 1)) * hw2) + static_cast<size_t>(y >> 1) * (W >> 1) + (x >> 1)) * 8;
    }
    return (static_cast<size_t>(y) * W + x) * 8;
}

// ------------------------------------------------------------------- stem
// Generic form (f32 input): one thread per pixel, 9 taps -> 32 channels in fp32, written as
// 4 planes of 8 bf16. The 288 folded weights arrive as a by-value kernel parameter, so every
// FFMA takes its weight straight from the constant bank (no LDS / LDG in the inner loop).
template <bool F16>
__global__ void __launch_bounds__(256)
stem_kernel(const void* __restrict__ frames, int in_dtype, const __grid_constant__ StemWeights sw,
            int B, int H, int W, __nv_bfloat16* __restrict__ out, int s2d) {
    // u8 -> float32(v) / 255.0f (IEEE division, as numpy does) through a 256-entry table
    __shared__ float lut[256];
    lut[threadIdx.x] = __fdiv_rn(static_cast<float>(threadIdx.x), 255.0f);
    __syncthreads();
    const size_t total = static_cast<size_t>(B) * H * W;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(idx % W);
        const int y = static_cast<int>((idx / W) % H);
        const size_t n = idx / (static_cast<size_t>(W) * H);
        float in[9];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int yy = y + dy - 1, xx = x + dx - 1;
                const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
                const size_t at = (n * H + (ok ? yy : y)) * static_cast<size_t>(W) + (ok ? xx : x);
                float v;
                if (in_dtype == 0) v = lut[static_cast<const uint8_t*>(frames)[at]];
                else v = static_cast<const float*>(frames)[at];
                in[dy * 3 + dx] = ok ? v : 0.f;
            }
        const size_t plane = static_cast<size_t>(H) * W * 8;
        __nv_bfloat16* o = out + n * 4 * plane + pix_off(y, x, H, W, s2d);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            float v[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int co = g * 8 + c;
                float acc = 0.f;
#pragma unroll
                for (int t = 0; t < 9; ++t) acc = fmaf(in[t], sw.w[co * 9 + t], acc);
                v[c] = fmaxf(acc + sw.b[co], 0.f);
            }
            uint4 q;
            q.x = pack_x2<F16>(v[0], v[1]);
            q.y = pack_x2<F16>(v[2], v[3]);
            q.z = pack_x2<F16>(v[4], v[5]);
            q.w = pack_x2<F16>(v[6], v[7]);
            *reinterpret_cast<uint4*>(o + g * plane) = q;
        }
    }
}

// u8 form (the hot path): one warp per image row, one thread per 4 consecutive pixels.
// The three input rows arrive as one aligned 32-bit load each; the left/right neighbour bytes
// come from the adjacent lanes by shuffle (warp-edge lanes load them). sw.w holds the folded
// weights already divided by 255 (utils.py:235), so the bytes are used as integers-in-fp32.
// 1152 FFMAs (constant-bank weights) per 4 pixels; bias rides in as the first addend; ReLU is a
// packed bf16x2 max after rounding; each channel group stores 64 contiguous bytes per thread.
template <bool F16>
__global__ void __launch_bounds__(256)
stem_u8_kernel(const uint8_t* __restrict__ frames, const __grid_constant__ StemPairs sw,
               int rows_total, int H, int W, __nv_bfloat16* __restrict__ out, int s2d) {
    const int lane = threadIdx.x & 31;
    const int warps_per_grid = gridDim.x * (blockDim.x >> 5);
    const int Q = W >> 2;  // quads per row
    const size_t plane = static_cast<size_t>(H) * W * 8;
    for (int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows_total;
         row += warps_per_grid) {
        const int n = row / H;
        const int y = row - n * H;
        const uint8_t* base = frames + static_cast<size_t>(row) * W;
        for (int q0 = 0; q0 < Q; q0 += 32) {
            const int xq = q0 + lane;
            const bool act = xq < Q;
            unsigned long long in[3][6];   // every input value in both halves of a register pair
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const int yy = y + dy - 1;
                const bool rowok = yy >= 0 && yy < H;  // warp-uniform
                const uint8_t* rp = base + (dy - 1) * W;
                uint32_t w = 0;
                if (rowok && act) w = *reinterpret_cast<const uint32_t*>(rp + 4 * xq);
                uint32_t lw = __shfl_up_sync(0xffffffffu, w, 1);
                uint32_t rw = __shfl_down_sync(0xffffffffu, w, 1);
                uint32_t lb = lw >> 24, rb = rw & 0xffu;
                if (lane == 0) lb = (rowok && act && xq > 0) ? rp[4 * xq - 1] : 0u;
                if (lane == 31 || xq + 1 >= Q) rb = (rowok && act && xq + 1 < Q) ? rp[4 * xq + 4] : 0u;
                in[dy][0] = dup2(static_cast<float>(lb));
                in[dy][1] = dup2(static_cast<float>(w & 0xffu));
                in[dy][2] = dup2(static_cast<float>((w >> 8) & 0xffu));
                in[dy][3] = dup2(static_cast<float>((w >> 16) & 0xffu));
                in[dy][4] = dup2(static_cast<float>(w >> 24));
                in[dy][5] = dup2(static_cast<float>(rb));
            }
            if (!act) continue;
            // s2d: even pixels (4xq, 4xq+2) are neighbours in phase (y&1, 0), odd ones in (y&1, 1)
            __nv_bfloat16* o = out + static_cast<size_t>(n) * 4 * plane + pix_off(y, 4 * xq, H, W, s2d);
            const size_t odd = static_cast<size_t>(H >> 1) * (W >> 1) * 8;  // phase (., 1) - (., 0)
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint32_t pk[4][4];  // [pixel][channel pair]
#pragma unroll
                for (int c2 = 0; c2 < 4; ++c2) {
                    const int cp = g * 4 + c2;   // channel pair (2 cp, 2 cp + 1)
#pragma unroll
                    for (int px = 0; px < 4; ++px) {
                        unsigned long long a = as_u64(sw.bp[cp]);
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                            for (int dx = 0; dx < 3; ++dx)
                                a = fma2(in[dy][px + dx], as_u64(sw.wp[cp * 9 + dy * 3 + dx]), a);
                        const float lo = __uint_as_float(static_cast<uint32_t>(a));
                        const float hi = __uint_as_float(static_cast<uint32_t>(a >> 32));
                        pk[px][c2] = pack_relu_x2<F16>(lo, hi);   // = max(round(v), 0)
                    }
                }
                uint4* dst = reinterpret_cast<uint4*>(o + g * plane);
                if (s2d) {
                    uint4* dso = reinterpret_cast<uint4*>(o + g * plane + odd);
                    dst[0] = make_uint4(pk[0][0], pk[0][1], pk[0][2], pk[0][3]);
                    dst[1] = make_uint4(pk[2][0], pk[2][1], pk[2][2], pk[2][3]);
                    dso[0] = make_uint4(pk[1][0], pk[1][1], pk[1][2], pk[1][3]);
                    dso[1] = make_uint4(pk[3][0], pk[3][1], pk[3][2], pk[3][3]);
                } else {
#pragma unroll
                    for (int px = 0; px < 4; ++px)
                        dst[px] = make_uint4(pk[px][0], pk[px][1], pk[px][2], pk[px][3]);
                }
            }
        }
    }
}

// ---------------------------------------------------------- fp32 validation
__global__ void f32_input_kernel(const void* __restrict__ frames, int in_dtype, int64_t count,
                                 float* __restrict__ out) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < count;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        out[i] = load_gray(frames, in_dtype, static_cast<size_t>(i));
}

// Direct conv3x3, pad 1. Block = 16x16 output pixels x 8 output channels of one frame; the
// input tile (18x18) of each input channel is staged in shared memory.
constexpr int kFc = 8;
__global__ void __launch_bounds__(256)
f32_conv3x3_kernel(const float* __restrict__ src0, int c0, const float* __restrict__ src1, int c1,
                   const float* __restrict__ w, const float* __restrict__ b,
                   float* __restrict__ out, int cout, int H, int W, int relu) {
    __shared__ float tile[18][18 + 1];
    __shared__ float wsm[kFc][9];
    const int tiles_x = (W + 15) / 16;
    const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
    const int cog = blockIdx.y;  // group of kFc output channels
    const int n = blockIdx.z;
    const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
    const int x = tx * 16 + lx, y = ty * 16 + ly;
    const int cin = c0 + c1;
    float acc[kFc];
#pragma unroll
    for (int k = 0; k < kFc; ++k) acc[k] = 0.f;
    for (int ci = 0; ci < cin; ++ci) {
        const float* src = ci < c0 ? src0 + (static_cast<size_t>(n) * c0 + ci) * H * W
                                   : src1 + (static_cast<size_t>(n) * c1 + (ci - c0)) * H * W;
        __syncthreads();
        for (int i = threadIdx.x; i < 18 * 18; i += 256) {
            const int hy = i / 18, hx = i % 18;
            const int yy = ty * 16 + hy - 1, xx = tx * 16 + hx - 1;
            tile[hy][hx] =
                (yy >= 0 && yy < H && xx >= 0 && xx < W) ? src[static_cast<size_t>(yy) * W + xx] : 0.f;
        }
        if (threadIdx.x < kFc * 9) {
            const int k = threadIdx.x / 9, t = threadIdx.x % 9;
            const int co = cog * kFc + k;
            wsm[k][t] = co < cout ? w[(static_cast<size_t>(co) * cin + ci) * 9 + t] : 0.f;
        }
        __syncthreads();
        float in[9];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) in[dy * 3 + dx] = tile[ly + dy][lx + dx];
#pragma unroll
        for (int k = 0; k < kFc; ++k)
#pragma unroll
            for (int t = 0; t < 9; ++t) acc[k] = fmaf(in[t], wsm[k][t], acc[k]);
    }
    if (x < W && y < H) {
#pragma unroll
        for (int k = 0; k < kFc; ++k) {
            const int co = cog * kFc + k;
            if (co < cout) {
                float v = acc[k] + b[co];
                if (relu) v = fmaxf(v, 0.f);
                out[((static_cast<size_t>(n) * cout + co) * H + y) * W + x] = v;
            }
        }
    }
}

__global__ void f32_maxpool_kernel(const float* __restrict__ in, float* __restrict__ out, int BC,
                                   int H, int W) {
    const int OH = H / 2, OW = W / 2;
    const size_t total = static_cast<size_t>(BC) * OH * OW;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int ox = static_cast<int>(i % OW);
        const int oy = static_cast<int>((i / OW) % OH);
        const size_t bc = i / (static_cast<size_t>(OW) * OH);
        const float* p = in + (bc * H + 2 * oy) * W + 2 * ox;
        out[i] = fmaxf(fmaxf(p[0], p[1]), fmaxf(p[W], p[W + 1]));
    }
}

// ConvTranspose2d k=2 s=2: thread per output element.
__global__ void f32_convt_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                 const float* __restrict__ b, float* __restrict__ out, int B,
                                 int cin, int cout, int H, int W) {
    const int OH = 2 * H, OW = 2 * W;
    const size_t total = static_cast<size_t>(B) * cout * OH * OW;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int ox = static_cast<int>(i % OW);
        const int oy = static_cast<int>((i / OW) % OH);
        const int co = static_cast<int>((i / (static_cast<size_t>(OW) * OH)) % cout);
        const size_t n = i / (static_cast<size_t>(OW) * OH * cout);
        const int y = oy >> 1, x = ox >> 1, dy = oy & 1, dx = ox & 1;
        float acc = 0.f;
        const float* ip = in + (n * cin * H + y) * W + x;
        const float* wp = w + (static_cast<size_t>(co) * 2 + dy) * 2 + dx;
        for (int ci = 0; ci < cin; ++ci)
            acc = fmaf(ip[static_cast<size_t>(ci) * H * W], wp[static_cast<size_t>(ci) * cout * 4], acc);
        out[i] = acc + b[co];
    }
}

__global__ void f32_head_kernel(const float* __restrict__ in, const float* __restrict__ w, float b,
                                float thr, int C, int H, int W, float* __restrict__ logits,
                                uint8_t* __restrict__ mask, int32_t* __restrict__ area) {
    // grid.y = frame; block reduces its pixel count and does one atomic.
    const int n = blockIdx.y;
    const size_t hw = static_cast<size_t>(H) * W;
    int cnt = 0;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < hw;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float z = 0.f;
        for (int c = 0; c < C; ++c) z = fmaf(in[(static_cast<size_t>(n) * C + c) * hw + i], w[c], z);
        z += b;
        const bool on = z > thr;
        if (logits) logits[n * hw + i] = z;
        if (mask) mask[n * hw + i] = on ? 255 : 0;
        cnt += on ? 1 : 0;
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && area && cnt) atomicAdd(area + n, cnt);
}

// ------------------------------------------------------- layout converters
__global__ void nchw_to_c8_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                  int B, int C, int H, int W, int s2d) {
    const size_t total = static_cast<size_t>(B) * C * H * W;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(i % W);
        const int y = static_cast<int>((i / W) % H);
        const int c = static_cast<int>((i / (static_cast<size_t>(W) * H)) % C);
        const size_t n = i / (static_cast<size_t>(W) * H * C);
        out[(n * (C / 8) + c / 8) * H * W * 8 + pix_off(y, x, H, W, s2d) + (c & 7)] =
            __float2bfloat16_rn(in[i]);
    }
}
__global__ void c8_to_nchw_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out,
                                  int B, int C, int H, int W, int s2d) {
    const size_t total = static_cast<size_t>(B) * C * H * W;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(i % W);
        const int y = static_cast<int>((i / W) % H);
        const int c = static_cast<int>((i / (static_cast<size_t>(W) * H)) % C);
        const size_t n = i / (static_cast<size_t>(W) * H * C);
        out[i] = __bfloat162float(
            in[(n * (C / 8) + c / 8) * H * W * 8 + pix_off(y, x, H, W, s2d) + (c & 7)]);
    }
}

inline int grid_for(size_t total, int block = 256, int cap = 148 * 16) {
    size_t g = (total + block - 1) / block;
    if (g > static_cast<size_t>(cap)) g = cap;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}

}  // namespace

int launch_stem(const void* frames, int in_dtype, const StemWeights& sw, int B, int H, int W,
                __nv_bfloat16* out, bool s2d, cudaStream_t stream, bool f16) {
    const size_t total = static_cast<size_t>(B) * H * W;
    if (s2d && (H % 2 || W % 2)) return fail("space-to-depth stem output needs even H and W");
    if (in_dtype == 0 && W % 4 == 0) {
        // hot path: weights pre-scaled by 1/255 (in fp64, rounded once) so the u8 values are
        // used directly; differs from fl(v/255)*w by ~1 ulp of fp32, far below bf16 rounding
        const StemPairs scaled = make_stem_pairs(sw);
        const int rows = B * H;
        int grid = (rows + 7) / 8;
        if (grid > 148 * 6) grid = 148 * 6;
        if (f16)
            stem_u8_kernel<true><<<grid, 256, 0, stream>>>(static_cast<const uint8_t*>(frames), scaled,
                                                           rows, H, W, out, s2d ? 1 : 0);
        else
            stem_u8_kernel<false><<<grid, 256, 0, stream>>>(static_cast<const uint8_t*>(frames), scaled,
                                                            rows, H, W, out, s2d ? 1 : 0);
    } else if (f16) {
        stem_kernel<true><<<grid_for(total, 256, 148 * 8), 256, 0, stream>>>(frames, in_dtype, sw, B,
                                                                             H, W, out, s2d ? 1 : 0);
    } else {
        stem_kernel<false><<<grid_for(total, 256, 148 * 8), 256, 0, stream>>>(frames, in_dtype, sw, B,
                                                                              H, W, out, s2d ? 1 : 0);
    }
    OGL_CUDA(cudaGetLastError());
    return 0;
}

int launch_f32_input(const void* frames, int in_dtype, int64_t count, float* out,
                     cudaStream_t stream) {
    f32_input_kernel<<<grid_for(static_cast<size_t>(count)), 256, 0, stream>>>(frames, in_dtype,
                                                                              count, out);
    OGL_CUDA(cudaGetLastError());
    return 0;
}

int launch_f32_conv3x3(const float* src0, int c0, const float* src1, int c1, const float* w,
                       const float* b, float* out, int B, int cout, int H, int W, int relu,
                       cudaStream_t stream) {
    dim3 grid(((W + 15) / 16) * ((H + 15) / 16), (cout + kFc - 1) / kFc, B);
    f32_conv3x3_kernel<<<grid, 256, 0, stream>>>(src0, c0, src1, c1, w, b, out, cout, H, W, relu);
    OGL_CUDA(cudaGetLastError());
    return 0;
}

int launch_f32_maxpool(const float* in, float* out, int BC, int H, int W, cudaStream_t stream) {
    f32_maxpool_kernel<<<grid_for(static_cast<size_t>(BC) * (H / 2) * (W / 2)), 256, 0, stream>>>(
        in, out, BC, H, W);
    OGL_CUDA(cudaGetLastError());
    return 0;
}

int launch_f32_convt(const float* in, const float* w, const float* b, float* out, int B, int cin,
                     int cout, int H, int W, cudaStream_t stream) {
    f32_convt_kernel<<<grid_for(static_cast<size_t>(B) * cout * 4 * H * W), 256, 0, stream>>>(
        in, w, b, out, B, cin, cout, H, W);
    OGL_CUDA(cudaGetLastError());
    return 0;
}

int launch_f32_head(const float* in, const float* w, float b, float thr, int B, int C, int H,
                    int W, float* logits, uint8_t* mask, int32_t* area, cudaStream_t stream) {
    const size_t hw = static_cast<size_t>(H) * W;
    dim3 grid(grid_for(hw, 256, 64), B);
    f32_head_kernel<<<grid, 256, 0, stream>>>(in, w, b, thr, C, H, W, logits, mask, area);
    OGL_CUDA(cudaGetLastError());
    return 0;
}

int launch_nchw_to_c8(const float* in, __nv_bfloat16* out, int B, int C, int H, int W,
                      cudaStream_t stream, bool s2d) {
    nchw_to_c8_kernel<<<grid_for(static_cast<size_t>(B) * C * H * W), 256, 0, stream>>>(
        in, out, B, C, H, W, s2d ? 1 : 0);
    OGL_CUDA(cudaGetLastError());
    return 0;
}
int launch_c8_to_nchw(const __nv_bfloat16* in, float* out, int B, int C, int H, int W,
                      cudaStream_t stream, bool s2d) {
    c8_to_nchw_kernel<<<grid_for(static_cast<size_t>(B) * C * H * W), 256, 0, stream>>>(
        in, out, B, C, H, W, s2d ? 1 : 0);
    OGL_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ogl
