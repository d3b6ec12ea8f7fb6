// Mask / frame operators of the callers around the U-Net:
//
//   ogl_mask_area_boxes      /root/reference/openglottal/features.py:240-245 (detection-gated area:
//                            np.sum(mask_full[y1:y2, x1:x2] > 0), box None -> 0)
//   ogl_letterbox_crops      /root/reference/scripts/infer.py:229-236 with
//                            openglottal/utils.py:103-131 (letterbox_with_info of the gray crop:
//                            cv2.resize INTER_NEAREST to the content size, zero padding)
//   ogl_unletterbox_area     openglottal/utils.py:170-186 (unletterbox: content region resized
//                            back with INTER_NEAREST) + scripts/infer.py:241-244 (paste, area)
//   ogl_mask_overlap_counts  openglottal/utils.py:191-206 (dice / iou are ratios of these counts)
//
// All byte/integer work, HBM-bound, bit-exact with the reference. cv2's INTER_NEAREST maps
// destination index d to source index min(floor(d * (1 / (dn / sn))), sn - 1) in double
// precision (checked against OpenCV 4.13 for all size pairs below 600 in tests/).
#include "internal.h"

namespace ogl {

namespace {

// Python slice semantics for [a:b] on an axis of length dim (negative indices wrap once).
__device__ __forceinline__ void slice_bounds(int a, int b, int dim, int* lo, int* hi) {
    if (a < 0) a += dim;
    if (b < 0) b += dim;
    a = a < 0 ? 0 : (a > dim ? dim : a);
    b = b < 0 ? 0 : (b > dim ? dim : b);
    *lo = a;
    *hi = b > a ? b : a;
}

__device__ __forceinline__ int nearest_src(int d, int sn, int dn) {
    const double fx = static_cast<double>(dn) / static_cast<double>(sn);
    const double ifx = 1.0 / fx;
    const int s = static_cast<int>(floor(__dmul_rn(static_cast<double>(d), ifx)));
    return s < sn - 1 ? s : sn - 1;
}

__device__ __forceinline__ int block_sum(int v, int* scratch) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    int t = 0;
    if (threadIdx.x < 32) {
        t = threadIdx.x < (blockDim.x >> 5) ? scratch[threadIdx.x] : 0;
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    __syncthreads();
    return t;  // valid in thread 0
}

// grid = (chunks, n): each block counts a slab of box rows and adds it to area[frame]
__global__ void __launch_bounds__(256)
mask_area_boxes_kernel(const uint8_t* __restrict__ mask, int H, int W,
                       const int32_t* __restrict__ boxes, const uint8_t* __restrict__ has_box,
                       int32_t* __restrict__ area) {
    __shared__ int scratch[8];
    const int n = blockIdx.y;
    if (has_box && !has_box[n]) return;
    int x1, x2, y1, y2;
    slice_bounds(boxes[4 * n + 0], boxes[4 * n + 2], W, &x1, &x2);
    slice_bounds(boxes[4 * n + 1], boxes[4 * n + 3], H, &y1, &y2);
    const int bw = x2 - x1, bh = y2 - y1;
    const long long total = static_cast<long long>(bw) * bh;
    const uint8_t* m = mask + static_cast<size_t>(n) * H * W;
    int cnt = 0;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int y = y1 + static_cast<int>(i / bw), x = x1 + static_cast<int>(i % bw);
        cnt += m[static_cast<size_t>(y) * W + x] > 0 ? 1 : 0;
    }
    const int t = block_sum(cnt, scratch);
    if (threadIdx.x == 0 && t) atomicAdd(area + n, t);
}

// geometry per frame: {x1, y1, x2, y2, pad_top, pad_left, content_h, content_w}; the crop is
// gray[y1:y2, x1:x2] with the bounds already normalised by the host (0 <= x1 < x2 <= W ...).
// content_h == 0 marks a frame without a crop: its output is all zeros.
__global__ void __launch_bounds__(256)
letterbox_crops_kernel(const uint8_t* __restrict__ gray, int H, int W,
                       const int32_t* __restrict__ geom, int size, uint8_t* __restrict__ out) {
    const int n = blockIdx.y;
    const int32_t* g = geom + 8 * n;
    const int x1 = g[0], y1 = g[1], cw = g[2] - g[0], ch = g[3] - g[1];
    const int pt = g[4], pl = g[5], nh = g[6], nw = g[7];
    const uint8_t* src = gray + static_cast<size_t>(n) * H * W;
    uint8_t* dst = out + static_cast<size_t>(n) * size * size;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < size * size; i += gridDim.x * blockDim.x) {
        const int oy = i / size, ox = i - oy * size;
        const int cy = oy - pt, cx = ox - pl;
        uint8_t v = 0;
        if (nh > 0 && cy >= 0 && cy < nh && cx >= 0 && cx < nw)
            v = src[static_cast<size_t>(y1 + nearest_src(cy, ch, nh)) * W + x1 + nearest_src(cx, cw, nw)];
        dst[i] = v;
    }
}

// mask_orig[y][x] = mask_cs[pad_top + nn(y), pad_left + nn(x)] for (y, x) in the crop;
// area = count(mask_orig > 0); full (optional) = zeros with mask_orig pasted at the box.
__global__ void __launch_bounds__(256)
unletterbox_area_kernel(const uint8_t* __restrict__ mask_cs, int size,
                        const int32_t* __restrict__ geom, int H, int W,
                        uint8_t* __restrict__ full, int32_t* __restrict__ area) {
    __shared__ int scratch[8];
    const int n = blockIdx.y;
    const int32_t* g = geom + 8 * n;
    const int x1 = g[0], y1 = g[1], cw = g[2] - g[0], ch = g[3] - g[1];
    const int pt = g[4], pl = g[5], nh = g[6], nw = g[7];
    if (nh <= 0) return;
    const uint8_t* src = mask_cs + static_cast<size_t>(n) * size * size;
    uint8_t* dst = full ? full + static_cast<size_t>(n) * H * W : nullptr;
    const int total = cw * ch;
    int cnt = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int y = i / cw, x = i - y * cw;
        const uint8_t v =
            src[static_cast<size_t>(pt + nearest_src(y, nh, ch)) * size + pl + nearest_src(x, nw, cw)];
        cnt += v > 0 ? 1 : 0;
        if (dst) dst[static_cast<size_t>(y1 + y) * W + x1 + x] = v;
    }
    const int t = block_sum(cnt, scratch);
    if (threadIdx.x == 0 && t) atomicAdd(area + n, t);
}

__global__ void __launch_bounds__(256)
overlap_counts_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ gt,
                      long long pixels, int32_t* __restrict__ counts) {
    __shared__ int scratch[8];
    const int n = blockIdx.y;
    const uint8_t* p = pred + static_cast<size_t>(n) * pixels;
    const uint8_t* q = gt + static_cast<size_t>(n) * pixels;
    int ci = 0, cp = 0, cg = 0;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < pixels;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int a = p[i] > 0, b = q[i] > 0;
        ci += a & b;
        cp += a;
        cg += b;
    }
    const int ti = block_sum(ci, scratch);
    const int tp = block_sum(cp, scratch);
    const int tg = block_sum(cg, scratch);
    if (threadIdx.x == 0) {
        if (ti) atomicAdd(counts + 3 * n + 0, ti);
        if (tp) atomicAdd(counts + 3 * n + 1, tp);
        if (tg) atomicAdd(counts + 3 * n + 2, tg);
    }
}

inline int chunks_for(long long work) {
    long long c = (work + 256 * 16 - 1) / (256 * 16);
    return static_cast<int>(c < 1 ? 1 : (c > 64 ? 64 : c));
}

}  // namespace

int launch_mask_area_boxes(const uint8_t* mask, int n, int H, int W, const int32_t* boxes,
                           const uint8_t* has_box, int32_t* area, cudaStream_t stream) {
    OGL_CUDA(cudaMemsetAsync(area, 0, sizeof(int32_t) * n, stream));
    dim3 grid(chunks_for(static_cast<long long>(H) * W), n);
    mask_area_boxes_kernel<<<grid, 256, 0, stream>>>(mask, H, W, boxes, has_box, area);
    OGL_CUDA(cudaGetLastError());
    return 0;
}

int launch_letterbox_crops(const uint8_t* gray, int n, int H, int W, const int32_t* geom, int size,
                           uint8_t* out, cudaStream_t stream) {
    dim3 grid(chunks_for(static_cast<long long>(size) * size), n);
    letterbox_crops_kernel<<<grid, 256, 0, stream>>>(gray, H, W, geom, size, out);
    OGL_CUDA(cudaGetLastError());
    return 0;
}

int launch_unletterbox_area(const uint8_t* mask_cs, int n, int size, const int32_t* geom, int H,
                            int W, uint8_t* full, int32_t* area, cudaStream_t stream) {
    OGL_CUDA(cudaMemsetAsync(area, 0, sizeof(int32_t) * n, stream));
    if (full) OGL_CUDA(cudaMemsetAsync(full, 0, static_cast<size_t>(n) * H * W, stream));
    dim3 grid(chunks_for(static_cast<long long>(H) * W), n);
    unletterbox_area_kernel<<<grid, 256, 0, stream>>>(mask_cs, size, geom, H, W, full, area);
    OGL_CUDA(cudaGetLastError());
    return 0;
}

int launch_overlap_counts(const uint8_t* pred, const uint8_t* gt, int n, long long pixels,
                          int32_t* counts, cudaStream_t stream) {
    OGL_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * 3 * n, stream));
    dim3 grid(chunks_for(pixels), n);
    overlap_counts_kernel<<<grid, 256, 0, stream>>>(pred, gt, pixels, counts);
    OGL_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ogl
