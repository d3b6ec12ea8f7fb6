// Microbenchmark: does a tcgen05.commit between groups of MMAs cost tensor-pipe time?
// One warp issues `groups` groups of G MMAs (M=128, N, K=16), each group followed by a commit to
// a scratch mbarrier (nobody waits on those), and a final commit that is waited on.
// COMMIT = 2 additionally executes tcgen05.fence::after_thread_sync before each group; COMMIT = 3
// waits on a (long completed) mbarrier phase + fence before each group, as the kernels' issuers do.
#include "ptx.cuh"
#include <cstdio>
#include <cuda_runtime.h>
using namespace ogl;

template <int N, int G, int COMMIT, int ISSUERS>
__global__ void __launch_bounds__(128, 1) bench(int groups, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    __shared__ uint64_t bar[2], scratch[16];
    __shared__ uint32_t tslot;
    for (int i = threadIdx.x; i < 200 * 1024 / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(smem_raw + (base - smem_u32(smem_raw)))[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar[0]), 1);
        mbar_init(smem_u32(&bar[1]), 1);
        for (int i = 0; i < 16; ++i) mbar_init(smem_u32(&scratch[i]), 1);
        fence_barrier_init();
    }
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(smem_u32(&tslot), 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&tslot);
    if (warp >= 1 && warp <= ISSUERS) {   // each issuer has its own accumulator and barriers
        const int w = warp - 1;
        const uint32_t idesc = make_idesc_bf16(N);
        const uint64_t ad0 = make_smem_desc(base, 2048, 128);
        const uint64_t bd0 = make_smem_desc(base + 64 * 1024, 16u * N, 128);
        const uint32_t bbar = smem_u32(&bar[w]);
        uint32_t phase = 0;
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
            for (int g = 0; g < groups; ++g) {
                if (COMMIT == 3) mbar_wait(smem_u32(&scratch[8 * w + 7]), 1);   // fresh barrier: parity 1 passes
                if (COMMIT >= 2) tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < G; ++k)
                        umma_bf16(tmem + w * 256, ad0 + (k & 7) * 256, bd0 + (k & 7) * 2 * N, idesc, 1u);
                    if (COMMIT) umma_commit(smem_u32(&scratch[8 * w + g % 7]));
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit(bbar);
            __syncwarp();
            mbar_wait(bbar, phase);
            phase ^= 1u;
            const long long t1 = clock64();
            if (rep == 2 && (threadIdx.x & 31) == 0) out[w] = t1 - t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

template <int N, int G, int COMMIT, int ISSUERS = 1>
void run(long long* out) {
    const int total = 4096;   // per issuer
    out[0] = out[1] = 0;
    cudaFuncSetAttribute(bench<N, G, COMMIT, ISSUERS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
    bench<N, G, COMMIT, ISSUERS><<<1, 128, 201 * 1024>>>(total / G, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
    const long long t = out[0] > out[1] ? out[0] : out[1];
    printf("N %3d  group %2d  mode %d  issuers %d : %7.1f cycles/MMA (pipe model %.0f)\n", N, G, COMMIT,
           ISSUERS, double(t) / (total * ISSUERS), N / 2.0 > 32 + N / 4.0 ? N / 2.0 : 32 + N / 4.0);
}

int main() {
    long long* out;
    cudaMallocManaged(&out, 8 * sizeof(long long));
    // mode 1: commit after each group; mode 3: wait on a completed mbarrier + fence before each group
    run<64, 16, 1>(out); run<64, 16, 3>(out); run<64, 16, 3, 2>(out);
    run<64, 4, 3>(out); run<64, 4, 3, 2>(out);
    run<128, 16, 3>(out); run<128, 16, 3, 2>(out);
    run<32, 16, 3>(out); run<32, 16, 3, 2>(out); run<32, 4, 3>(out); run<32, 4, 3, 2>(out);
    run<64, 16, 1, 2>(out);
    return 0;
}
