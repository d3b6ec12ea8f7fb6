// Microbenchmark: tcgen05.mma.cta_group::2 (kind::f16, M=256 across a CTA pair, SWIZZLE_NONE
// K-major operands) -- cycles per MMA on the issuing (leader) SM as a function of N.
// Each CTA holds its own 128 x 16 A tile and HALF of the N x 16 B tile.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I openglottal_b200/csrc
//        -o /tmp/mma2_bench scripts/microbench/mma2_bench.cu
#include "ptx.cuh"
#include <cstdio>
#include <cuda_runtime.h>
using namespace ogl;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_slot),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                           uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
        "[%0], %1;" ::"r"(bar),
        "h"(mask)
        : "memory");
}
__host__ __device__ inline uint32_t make_idesc_bf16_m256(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
           (static_cast<uint32_t>(256 >> 4) << 24);
}

template <int N, int ND, int VA, int VB>
__global__ void __launch_bounds__(128, 1) mma2_bench(int iters8, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const uint32_t rank = cluster_ctarank();
    for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)))[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar), 1);
        fence_barrier_init();
    }
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc2(smem_u32(&tslot), 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&tslot);
    const uint32_t bbar = smem_u32(&bar);
    if (warp == 1) {
        uint32_t phase = 0;
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
            if (rank == 0) {
                const uint32_t idesc = make_idesc_bf16_m256(N);
                const uint64_t ad0 = make_smem_desc(base, 2048, 128);                      // 128 x 16
                const uint64_t bd0 = make_smem_desc(base + 64 * 1024, 16u * (N / 2), 128); // N/2 x 16
                for (int i = 0; i < iters8; ++i) {
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            umma2_bf16(tmem + (k % ND) * N, ad0 + (VA ? k * 256 : 0),
                                       bd0 + (VB ? k * N : 0), idesc, 1u);
                    }
                    __syncwarp();
                }
                if (elect_one()) umma2_commit_mc(bbar, 3);
                __syncwarp();
            }
            mbar_wait(bbar, phase);   // both CTAs: the commit is multicast
            phase ^= 1u;
            const long long t1 = clock64();
            if (rep == 2 && (threadIdx.x & 31) == 0) out[blockIdx.x] = t1 - t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc2(tmem, 512);
    }
}

template <int N, int ND, int VA, int VB>
void run(long long* out) {
    const int iters8 = 512;
    auto kern = mma2_bench<N, ND, VA, VB>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
    for (int grid : {2, 148}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(128);
        cfg.dynamicSmemBytes = 201 * 1024;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, iters8, out);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("error: %s\n", cudaGetErrorString(e));
            exit(1);
        }
        long long mx = 0;
        for (int i = 0; i < grid; i += 2) mx = out[i] > mx ? out[i] : mx;
        printf("%5d %4d %3d %3d %5d | %9.1f %9.1f %9.1f\n", N, ND, VA, VB, grid,
               double(mx) / (iters8 * 8), N / 2.0, 32 + N / 8.0);
    }
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 148 * sizeof(long long));
    printf("%5s %4s %3s %3s %5s | %9s %9s %9s\n", "N", "nD", "vA", "vB", "grid", "cyc/mma", "math N/2",
           "smem wf");
    run<32, 1, 1, 1>(out);
    run<32, 4, 1, 1>(out);
    run<64, 1, 1, 1>(out);
    run<64, 4, 1, 1>(out);
    run<96, 1, 1, 1>(out);
    run<128, 1, 0, 0>(out);
    run<128, 1, 1, 1>(out);
    run<128, 4, 1, 1>(out);
    run<192, 2, 1, 1>(out);
    run<256, 1, 1, 1>(out);
    run<256, 2, 1, 1>(out);
    return 0;
}
