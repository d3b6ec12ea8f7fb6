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

// 0xff in every byte of v that is non-zero
__device__ __forceinline__ uint32_t nonzero_bytes(uint32_t v) { return __vcmpgtu4(v, 0u); }

// The same count over 16-byte words (mask 16-byte aligned, W % 16 == 0, so every row is a whole
// number of aligned words): a block is 16 rows x 16 lanes, a lane walks the words that overlap
// [x1, x2) and only the first / last word of a row needs a byte mask. No division per pixel and
// 16 bytes per load instead of 1.
__global__ void __launch_bounds__(256)
mask_area_boxes_vec16_kernel(const uint8_t* __restrict__ mask, int H, int W,
                             const int32_t* __restrict__ boxes, const uint8_t* __restrict__ has_box,
                             int32_t* __restrict__ area) {
    __shared__ int scratch[8];
    const int n = blockIdx.y;
    if (has_box && !has_box[n]) return;
    int x1, x2, y1, y2;
    slice_bounds(boxes[4 * n + 0], boxes[4 * n + 2], W, &x1, &x2);
    slice_bounds(boxes[4 * n + 1], boxes[4 * n + 3], H, &y1, &y2);
    const uint4* m = reinterpret_cast<const uint4*>(mask + static_cast<size_t>(n) * H * W);
    const int wpr = W >> 4;                                   // words per row
    const int w0 = x1 >> 4, w1 = (x2 + 15) >> 4;              // words overlapping [x1, x2)
    const int lane = threadIdx.x & 15, row = threadIdx.x >> 4;
    int bits = 0;                                             // 8 per counted byte
    if (x2 > x1) {
        for (int y = y1 + blockIdx.x * 16 + row; y < y2; y += gridDim.x * 16) {
            const uint4* r = m + static_cast<size_t>(y) * wpr;
            for (int w = w0 + lane; w < w1; w += 16) {
                const uint4 v = __ldg(r + w);
                uint32_t q[4] = {nonzero_bytes(v.x), nonzero_bytes(v.y), nonzero_bytes(v.z),
                                 nonzero_bytes(v.w)};
                const int xb = w << 4;
                if (xb < x1 || xb + 16 > x2) {                // edge word: keep bytes in [x1, x2)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        uint32_t keep = 0;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int x = xb + 4 * k + j;
                            if (x >= x1 && x < x2) keep |= 0xffu << (8 * j);
                        }
                        q[k] &= keep;
                    }
                }
                bits += __popc(q[0]) + __popc(q[1]) + __popc(q[2]) + __popc(q[3]);
            }
        }
    }
    const int t = block_sum(bits >> 3, scratch);
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

// The same image with the two nearest-neighbour index maps computed once per block into shared
// memory (nearest_src is double-precision arithmetic: 2 divisions per pixel before, 2 * size per
// block now) and four output bytes per store. tab[0..size) = source column (gray x) of output
// column ox or -1 for padding, tab[size..2 size) = source row. size % 4 == 0, out 4-byte aligned.
__global__ void __launch_bounds__(256)
letterbox_crops_tab_kernel(const uint8_t* __restrict__ gray, int H, int W,
                           const int32_t* __restrict__ geom, int size, uint8_t* __restrict__ out) {
    extern __shared__ int tab[];
    const int n = blockIdx.y;
    const int32_t* g = geom + 8 * n;
    const int x1 = g[0], y1 = g[1], cw = g[2] - g[0], ch = g[3] - g[1];
    const int pt = g[4], pl = g[5], nh = g[6], nw = g[7];
    for (int i = threadIdx.x; i < 2 * size; i += blockDim.x) {
        const bool is_x = i < size;
        const int c = is_x ? i - pl : i - size - pt;          // content coordinate
        const int cn = is_x ? nw : nh, sn = is_x ? cw : ch;
        tab[i] = (nh > 0 && c >= 0 && c < cn) ? (is_x ? x1 : y1) + nearest_src(c, sn, cn) : -1;
    }
    __syncthreads();
    const uint8_t* src = gray + static_cast<size_t>(n) * H * W;
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + static_cast<size_t>(n) * size * size);
    const int quads = size >> 2;                               // per row
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < size * quads; i += gridDim.x * blockDim.x) {
        const int oy = i / quads, ox = (i - oy * quads) << 2;
        const int sy = tab[size + oy];
        uint32_t v = 0;
        if (sy >= 0) {
            const uint8_t* row = src + static_cast<size_t>(sy) * W;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int sx = tab[ox + j];
                if (sx >= 0) v |= static_cast<uint32_t>(row[sx]) << (8 * j);
            }
        }
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

// The same with the index maps in shared memory: tab[0..cw) = crop-space column of crop column
// x, tab[W..W+ch) = crop-space row of crop row y (cw <= W, ch <= H; (W + H) ints of shared memory).
__global__ void __launch_bounds__(256)
unletterbox_area_tab_kernel(const uint8_t* __restrict__ mask_cs, int size,
                            const int32_t* __restrict__ geom, int H, int W,
                            uint8_t* __restrict__ full, int32_t* __restrict__ area) {
    extern __shared__ int tab[];
    __shared__ int scratch[8];
    const int n = blockIdx.y;
    const int32_t* g = geom + 8 * n;
    const int x1 = g[0], y1 = g[1], cw = g[2] - g[0], ch = g[3] - g[1];
    const int pt = g[4], pl = g[5], nh = g[6], nw = g[7];
    if (nh <= 0) return;
    for (int i = threadIdx.x; i < cw; i += blockDim.x) tab[i] = pl + nearest_src(i, nw, cw);
    for (int i = threadIdx.x; i < ch; i += blockDim.x) tab[W + i] = pt + nearest_src(i, nh, ch);
    __syncthreads();
    const uint8_t* src = mask_cs + static_cast<size_t>(n) * size * size;
    uint8_t* dst = full ? full + static_cast<size_t>(n) * H * W : nullptr;
    const int total = cw * ch;
    int cnt = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int y = i / cw, x = i - y * cw;
        const uint8_t v = src[static_cast<size_t>(tab[W + y]) * size + tab[x]];
        cnt += v > 0 ? 1 : 0;
        if (dst) dst[static_cast<size_t>(y1 + y) * W + x1 + x] = v;
    }
    const int t = block_sum(cnt, scratch);
    if (threadIdx.x == 0 && t) atomicAdd(area + n, t);
}

// 16 bytes of each mask per load (both 16-byte aligned, pixels % 16 == 0): a byte-wise "non-zero"
// compare and population counts instead of one byte per thread and iteration.
__global__ void __launch_bounds__(256)
overlap_counts_vec16_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ gt,
                            long long pixels, int32_t* __restrict__ counts) {
    __shared__ int scratch[8];
    const int n = blockIdx.y;
    const uint4* p = reinterpret_cast<const uint4*>(pred + static_cast<size_t>(n) * pixels);
    const uint4* q = reinterpret_cast<const uint4*>(gt + static_cast<size_t>(n) * pixels);
    const long long words = pixels >> 4;
    int bi = 0, bp = 0, bg = 0;                                // 8 bits per counted byte
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < words;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const uint4 a = __ldg(p + i), b = __ldg(q + i);
        const uint32_t ax = nonzero_bytes(a.x), ay = nonzero_bytes(a.y), az = nonzero_bytes(a.z),
                       aw = nonzero_bytes(a.w);
        const uint32_t bx = nonzero_bytes(b.x), by = nonzero_bytes(b.y), bz = nonzero_bytes(b.z),
                       bw = nonzero_bytes(b.w);
        bi += __popc(ax & bx) + __popc(ay & by) + __popc(az & bz) + __popc(aw & bw);
        bp += __popc(ax) + __popc(ay) + __popc(az) + __popc(aw);
        bg += __popc(bx) + __popc(by) + __popc(bz) + __popc(bw);
    }
    const int ti = block_sum(bi >> 3, scratch);
    const int tp = block_sum(bp >> 3, scratch);
    const int tg = block_sum(bg >> 3, scratch);
    if (threadIdx.x == 0) {
        if (ti) atomicAdd(counts + 3 * n + 0, ti);
        if (tp) atomicAdd(counts + 3 * n + 1, tp);
        if (tg) atomicAdd(counts + 3 * n + 2, tg);
    }
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
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
constexpr size_t kMaxTabBytes = 40 * 1024;   // index tables in (default-limit) shared memory

}  // namespace

int launch_mask_area_boxes(const uint8_t* mask, int n, int H, int W, const int32_t* boxes,
                           const uint8_t* has_box, int32_t* area, cudaStream_t stream) {
    OGL_CUDA(cudaMemsetAsync(area, 0, sizeof(int32_t) * n, stream));
    dim3 grid(chunks_for(static_cast<long long>(H) * W), n);
    if (W % 16 == 0 && aligned16(mask)) {        // (H * W) % 16 == 0 too: every frame is aligned
        grid.x = frame_chunks(static_cast<long long>(H) * (W / 16), 4, n);
        mask_area_boxes_vec16_kernel<<<grid, 256, 0, stream>>>(mask, H, W, boxes, has_box, area);
    } else {
        mask_area_boxes_kernel<<<grid, 256, 0, stream>>>(mask, H, W, boxes, has_box, area);
    }
    OGL_CUDA(cudaGetLastError());
    return 0;
}

int launch_letterbox_crops(const uint8_t* gray, int n, int H, int W, const int32_t* geom, int size,
                           uint8_t* out, cudaStream_t stream) {
    dim3 grid(chunks_for(static_cast<long long>(size) * size), n);
    const size_t tab = 2 * sizeof(int) * static_cast<size_t>(size);
    if (size % 4 == 0 && tab <= kMaxTabBytes && (reinterpret_cast<uintptr_t>(out) & 3) == 0) {
        grid.x = frame_chunks(static_cast<long long>(size) * (size / 4), 8, n);
        letterbox_crops_tab_kernel<<<grid, 256, tab, stream>>>(gray, H, W, geom, size, out);
    } else {
        letterbox_crops_kernel<<<grid, 256, 0, stream>>>(gray, H, W, geom, size, out);
    }
    OGL_CUDA(cudaGetLastError());
    return 0;
}

int launch_unletterbox_area(const uint8_t* mask_cs, int n, int size, const int32_t* geom, int H,
                            int W, uint8_t* full, int32_t* area, cudaStream_t stream) {
    OGL_CUDA(cudaMemsetAsync(area, 0, sizeof(int32_t) * n, stream));
    if (full) OGL_CUDA(cudaMemsetAsync(full, 0, static_cast<size_t>(n) * H * W, stream));
    dim3 grid(chunks_for(static_cast<long long>(H) * W), n);
    const size_t tab = sizeof(int) * (static_cast<size_t>(H) + W);
    if (tab <= kMaxTabBytes) {
        grid.x = frame_chunks(static_cast<long long>(H) * W, 32, n);
        unletterbox_area_tab_kernel<<<grid, 256, tab, stream>>>(mask_cs, size, geom, H, W, full, area);
    } else {
        unletterbox_area_kernel<<<grid, 256, 0, stream>>>(mask_cs, size, geom, H, W, full, area);
    }
    OGL_CUDA(cudaGetLastError());
    return 0;
}

int launch_overlap_counts(const uint8_t* pred, const uint8_t* gt, int n, long long pixels,
                          int32_t* counts, cudaStream_t stream) {
    OGL_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * 3 * n, stream));
    dim3 grid(chunks_for(pixels), n);
    if (pixels % 16 == 0 && aligned16(pred) && aligned16(gt)) {
        grid.x = frame_chunks(pixels / 16, 4, n);
        overlap_counts_vec16_kernel<<<grid, 256, 0, stream>>>(pred, gt, pixels, counts);
    } else {
        overlap_counts_kernel<<<grid, 256, 0, stream>>>(pred, gt, pixels, counts);
    }
    OGL_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ogl
