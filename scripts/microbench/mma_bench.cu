// Microbenchmark: throughput of tcgen05.mma (kind::f16, M=128, cta_group::1, SWIZZLE_NONE K-major
// operands in shared memory) as a function of N, of whether consecutive MMAs share the accumulator,
// and of whether A / B change from one MMA to the next. The issuing warp keeps every operand in
// uniform registers (unrolled block under one elect.sync, constant offsets), as the product
// kernels do. Prints cycles per MMA (clock64, commit + wait at the end).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I openglottal_b200/csrc
//        -o /tmp/mma_bench scripts/microbench/mma_bench.cu
#include "ptx.cuh"
#include <cstdio>
#include <cuda_runtime.h>
using namespace ogl;

template <int N, int ND, int VA, int VB>
__global__ void __launch_bounds__(128, 1) mma_bench(int iters8, long long* out) {
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
        const uint32_t idesc = make_idesc_bf16(N);
        const uint64_t ad0 = make_smem_desc(base, 2048, 128);                 // 128 x 16 tile: 4 KB
        const uint64_t bd0 = make_smem_desc(base + 64 * 1024, 16u * N, 128);  // N x 16
        const uint32_t bbar = smem_u32(&bar);
        uint32_t phase = 0;
        for (int rep = 0; rep < 3; ++rep) {   // rep 0, 1 warm up
            const long long t0 = clock64();
            for (int i = 0; i < iters8; ++i) {
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        umma_bf16(tmem + (k % ND) * N, ad0 + (VA ? k * 256 : 0),
                                  bd0 + (VB ? k * 2 * N : 0), idesc, 1u);
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

template <int N, int ND, int VA, int VB>
void run(long long* out) {
    const int iters8 = 512;
    cudaFuncSetAttribute(mma_bench<N, ND, VA, VB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         201 * 1024);
    for (int grid : {1, 148}) {
        mma_bench<N, ND, VA, VB><<<grid, 128, 201 * 1024>>>(iters8, out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("error: %s\n", cudaGetErrorString(e));
            exit(1);
        }
        long long mx = 0;
        for (int i = 0; i < grid; ++i) mx = out[i] > mx ? out[i] : mx;
        printf("%5d %4d %3d %3d %5d | %9.1f %9.1f %9.1f\n", N, ND, VA, VB, grid,
               double(mx) / (iters8 * 8), N / 2.0, 32 + N / 4.0);
    }
}
template <int N>
void run_n(long long* out) {
    run<N, 1, 0, 0>(out);
    run<N, 1, 1, 0>(out);
    run<N, 1, 0, 1>(out);
    run<N, 1, 1, 1>(out);
    if (4 * N <= 512) {
        run<N, 4, 0, 0>(out);
        run<N, 4, 1, 0>(out);
        run<N, 4, 1, 1>(out);
    } else if (2 * N <= 512) {
        run<N, 2, 1, 0>(out);
        run<N, 2, 1, 1>(out);
    }
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 148 * sizeof(long long));
    printf("%5s %4s %3s %3s %5s | %9s %9s %9s\n", "N", "nD", "vA", "vB", "grid", "cyc/mma", "math N/2",
           "smem wf");
    run_n<32>(out);
    run_n<64>(out);
    run_n<96>(out);
    run_n<128>(out);
    run_n<192>(out);
    run_n<256>(out);
    return 0;
}
