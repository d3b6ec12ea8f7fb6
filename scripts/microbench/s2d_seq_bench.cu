// Microbenchmark: the 16-MMA program of one 16-channel slab of the space-to-depth kernels
// (s2d_tc.cu: N = 128 / 96 / 64 / 32 by op, A = one halo stage at 16 start offsets, SBO = one halo
// row = 160 B, LBO = a plane pair) issued back to back from one warp, against the model
// sum(max(N/2, 32 + N/4)) = 832 cycles, and variants that isolate what costs more than the model:
//   0  the program as the kernel issues it
//   1  the same N sequence on a dense A tile (SBO = 128 B, LBO = 2 KB, 128-byte aligned starts)
//   2  the kernel's A addressing with N = 128 for every op
//   3  as 0 with every op on its own 128-column accumulator buffer offset 0 (dcol = 0)
//   4  as 0, ops sorted by N (128 x 4, 96 x 4, 64 x 4, 32 x 4)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I openglottal_b200/csrc
//        -o /tmp/s2d_seq_bench scripts/microbench/s2d_seq_bench.cu
#include "ptx.cuh"
#include <cstdio>
#include <cuda_runtime.h>
using namespace ogl;

constexpr int kHW = 10, kHH = 18, kPlane = kHW * kHH * 16;
__host__ __device__ constexpr int tap_of(int o, int q, int p) {
    const int d = 2 * o + q - p + 1;
    return (d >= 0 && d <= 2) ? d : -1;
}
__host__ __device__ constexpr int combo_o(int i) { return i == 2 ? -1 : (i == 3 ? 1 : 0); }
__host__ __device__ constexpr int combo_q(int i) { return (i == 1 || i == 2) ? 1 : 0; }
__host__ __device__ constexpr int set_s2d(int i) {
    return (tap_of(combo_o(i), combo_q(i), 0) >= 0 ? 1 : 0) | (tap_of(combo_o(i), combo_q(i), 1) >= 0 ? 2 : 0);
}
struct OpShape { int a_off, dcol, n; };
__host__ __device__ constexpr OpShape s2d_shape(int c) {
    const int cy = c >> 2, cx = c & 3, ys = set_s2d(cy), xs = set_s2d(cx);
    int pmin = 4, pmax = -1;
    for (int pp = 0; pp < 4; ++pp)
        if (((ys >> (pp >> 1)) & 1) && ((xs >> (pp & 1)) & 1)) {
            pmin = pp < pmin ? pp : pmin;
            pmax = pp > pmax ? pp : pmax;
        }
    return OpShape{(combo_q(cy) * 2 + combo_q(cx)) * (kPlane / 16) + (1 + combo_o(cy)) * kHW + (1 + combo_o(cx)),
                   pmin * 32, (pmax - pmin + 1) * 32};
}
__host__ __device__ constexpr int boff(int c) {
    int o = 0;
    for (int i = 0; i < c; ++i) o += 2 * s2d_shape(i).n;
    return o;
}
__host__ __device__ constexpr int sorted_op(int i) {   // ops ordered by decreasing N
    int k = 0;
    for (int n = 128; n >= 32; n -= 32)
        for (int c = 0; c < 16; ++c)
            if (s2d_shape(c).n == n) {
                if (k == i) return c;
                ++k;
            }
    return 0;
}

template <int MODE>
__global__ void __launch_bounds__(128, 1) seq_bench(int iters, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)))[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar), 1);
        fence_barrier_init();
    }
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(smem_u32(&tslot), 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&tslot);
    if (warp == 1) {
        const uint64_t a_s2d = (static_cast<uint64_t>(((kHW * 16u) >> 4) | (1u << 14)) << 32) |
                               ((base >> 4) | (((4u * kPlane) >> 4) << 16));
        const uint64_t a_dense = make_smem_desc(base, 2048, 128);
        const uint64_t b_hi = static_cast<uint64_t>((128u >> 4) | (1u << 14)) << 32;
        const uint32_t w_lo = (base + 64 * 1024) >> 4;
        const uint32_t bbar = smem_u32(&bar);
        uint32_t phase = 0;
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const int c = MODE == 4 ? sorted_op(k) : k;
                        const OpShape sh = s2d_shape(c);
                        const int n = MODE == 2 ? 128 : sh.n;
                        const uint32_t d = tmem + (MODE == 2 || MODE == 3 ? 0 : sh.dcol);
                        const uint64_t ad = MODE == 1 ? a_dense + k * 256 : a_s2d + sh.a_off;
                        const uint64_t bd = b_hi | (w_lo + boff(c) + (static_cast<uint32_t>(n) << 16));
                        umma_bf16(d, ad, bd, make_idesc_bf16(n), 1u);
                    }
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(bbar);
            __syncwarp();
            mbar_wait(bbar, phase);
            phase ^= 1u;
            const long long t1 = clock64();
            if (rep == 2 && (threadIdx.x & 31) == 0) out[blockIdx.x] = t1 - t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

template <int MODE>
void run(long long* out, const char* what) {
    const int iters = 512;
    cudaFuncSetAttribute(seq_bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
    for (int grid : {1, 148}) {
        seq_bench<MODE><<<grid, 128, 201 * 1024>>>(iters, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("error: %s\n", cudaGetErrorString(e));
            exit(1);
        }
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = out[i] > mx ? out[i] : mx;
        printf("mode %d grid %3d: %8.1f cycles per 16-op slab   (%s)\n", MODE, grid, double(mx) / iters, what);
    }
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 148 * sizeof(long long));
    int model = 0, math = 0;
    for (int c = 0; c < 16; ++c) {
        const int n = s2d_shape(c).n;
        model += n / 2 > 32 + n / 4 ? n / 2 : 32 + n / 4;
        math += n / 2;
    }
    printf("model: %d cycles per slab (math only %d)\n", model, math);
    run<0>(out, "kernel's program");
    run<1>(out, "same N sequence, dense A tile");
    run<2>(out, "kernel's A addressing, N = 128 everywhere");
    run<3>(out, "kernel's program, every op at accumulator column 0");
    run<4>(out, "kernel's program sorted by N");
    return 0;
}
