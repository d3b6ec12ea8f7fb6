// Isolated check of TMA tile loads of small-element tensors (used while building the fused stem).
// argv: dtype(0 u8, 1 bf16) rank(3|4) boxw  -- one configuration per process.
#include "ptx.cuh"
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <vector>
using namespace ogl;

__global__ void k(const __grid_constant__ CUtensorMap tm, int rank, int cx, int cy, int cn,
                  uint8_t* out, int bytes) {
    extern __shared__ __align__(128) uint8_t buf[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar), 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(smem_u32(&bar), bytes);
        if (rank == 3) tma_load_3d(smem_u32(buf), &tm, smem_u32(&bar), cx, cy, cn);
        else tma_load_4d(smem_u32(buf), &tm, smem_u32(&bar), cx, cy, 0, cn);
    }
    mbar_wait(smem_u32(&bar), 0);
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = buf[i];
}

int main(int argc, char** argv) {
    const int dt = atoi(argv[1]), rank = atoi(argv[2]), bw = atoi(argv[3]);
    const int es_bytes = dt == 0 ? 1 : 2;
    const int B = 3, H = 64, W = 48;
    std::vector<uint8_t> h(B * H * W * es_bytes);
    for (size_t i = 0; i < h.size(); ++i) h[i] = static_cast<uint8_t>(i * 7 + 1);
    uint8_t *d, *o;
    cudaMalloc(&d, h.size());
    cudaMalloc(&o, 16384);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto enc = reinterpret_cast<CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                             const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill)>(fn);
    CUtensorMap tm;
    cuuint64_t dims[4], str[3];
    cuuint32_t box[4], es[4] = {1, 1, 1, 1};
    if (rank == 3) {
        dims[0] = W; dims[1] = H; dims[2] = B;
        str[0] = (cuuint64_t)W * es_bytes; str[1] = (cuuint64_t)W * H * es_bytes;
        box[0] = bw; box[1] = 38; box[2] = 1;
    } else {
        dims[0] = W; dims[1] = H; dims[2] = 1; dims[3] = B;
        str[0] = (cuuint64_t)W * es_bytes; str[1] = (cuuint64_t)W * H * es_bytes; str[2] = str[1];
        box[0] = bw; box[1] = 38; box[2] = 1; box[3] = 1;
    }
    CUresult r = enc(&tm, dt == 0 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank,
                     d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("dt %d rank %d box %d: encode rc=%d\n", dt, rank, bw, (int)r); return 0; }
    const int cx = argc > 4 ? atoi(argv[4]) : -3, cy = -3, cn = 1, bytes = bw * 38 * es_bytes;
    k<<<1, 128, 16384 + 128>>>(tm, rank, cx, cy, cn, o, bytes);
    printf("launch: %s\n", cudaGetErrorString(cudaGetLastError()));
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("dt %d rank %d box %d: error: %s\n", dt, rank, bw, cudaGetErrorString(e)); return 0; }
    std::vector<uint8_t> g(bytes);
    cudaMemcpy(g.data(), o, g.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < 38; ++y)
        for (int x = 0; x < bw; ++x)
            for (int b = 0; b < es_bytes; ++b) {
                const int gy = cy + y, gx = cx + x;
                const uint8_t want = (gy >= 0 && gy < H && gx >= 0 && gx < W)
                                         ? h[((cn * H + gy) * W + gx) * es_bytes + b] : 0;
                bad += g[(y * bw + x) * es_bytes + b] != want;
            }
    printf("dt %d rank %d box %d: mismatches %d\n", dt, rank, bw, bad);
    return 0;
}
