// Full-resolution conv3x3 layers (Cout = 32) as space-to-depth GEMMs on tcgen05 / TMEM.
//
// Replaces, at the first U-Net level, /root/reference/openglottal/models/unet.py:24-29
// (downs.0.net.3 + the MaxPool2d of :59,79; ups.7.net.0 with the ConvTranspose2d ups.6 of
// :69,82 and the torch.cat of :86 composed in; ups.7.net.3 + the 1x1 head of :72,88 +
// utils.py:237,241 sigmoid/threshold + features.py:238 area).
//
// Tensors are stored S2D (internal.h): [frame][C/8][phase][H/2][W/2][8]. A CTA tile is 128
// half-resolution positions (8 in x, 16 in y) = 512 pixels; its accumulator is 128 TMEM lanes x
// 128 columns = (output phase, cout). For input phase q at half-resolution offset o the taps
// that exist are those with 2*o + q - p in {-1,0,1}^2, so ONE MMA with A = that phase plane
// shifted by o serves every output phase p it reaches: N = 128 (4 phases), 64, 96 (two phases
// whose columns are not adjacent: the block between them is zero weights) or 32.
//   * The shape of every MMA of a tile (A offset, B offset, N, columns) is a compile-time
//     table (s2d_shape / below_shape) shared by the host packer and the issuer, whose loop is
//     fully unrolled so that all tcgen05.mma operands sit in uniform registers (measured:
//     85 cycles per MMA when they come through R2UR, max(N/2, 32 + N/4) when they do not).
//   * Weights (80 KB for Cin = 32, 152 KB with the composed transposed conv) are loaded into
//     shared memory ONCE per CTA and stay resident; only activations stream (TMA, ring of
//     23 KB stages = 8 planes of a 10x18 halo tile).
//   * 4 accumulator buffers in TMEM; two epilogue warp groups take alternate tiles, so the
//     epilogue of two tiles overlaps the MMAs of the next two.
//   * Epilogue per thread = one half-resolution position: 4 phases x 32 channels. The 2x2
//     max-pool is a max over the 4 phases held by the same thread (no shuffles); the head is
//     four dot products on the fp32 values.
#include "internal.h"
#include "ptx.cuh"

#include <cuda.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#ifdef OGL_F16   // s2d_tc_f16.cu: the same kernels with f16 operands, exported under other names
#define launch_s2d_tc launch_s2d_tc_f16
#define s2d_tc_init s2d_tc_init_f16
#endif

namespace ogl {

namespace {

constexpr int kTW = 8, kTH = 16;               // tile: 8 x 16 half-resolution positions
constexpr int kHW = kTW + 2, kHH = kTH + 2;    // with halo: 10 x 18
constexpr int kPlane = kHW * kHH * 16;         // one 8-channel plane of a halo tile (2880 B)
constexpr int kSlot = 8 * kPlane;              // one activation stage (23040 B)
constexpr int kThreads = 384;                  // 4 control warps + 2 epilogue groups of 4 warps
constexpr int kAccBufs = 4;
constexpr int kMaxSmem = 227 * 1024;
constexpr uint32_t kWChunk = 16384;            // weight bulk-copy granule

// ---- MMA shapes (compile-time, host + device) ------------------------------------------
// 1-D structure of a 3x3 stride-1 conv on phase-separated data: input phase q at
// half-resolution offset o sits at full-resolution coordinate 2*o + q relative to the output
// position's even pixel; output phase p uses it through tap d = 2*o + q - p + 1 if 0 <= d <= 2.
__host__ __device__ constexpr int tap_of(int o, int q, int p) {
    const int d = 2 * o + q - p + 1;
    return (d >= 0 && d <= 2) ? d : -1;
}
// the four (offset, phase) pairs of one axis; the one reaching both output phases comes first so
// that op 0 of a tile is the N = 128 one that initialises every accumulator column
__host__ __device__ constexpr int combo_o(int i) { return i == 2 ? -1 : (i == 3 ? 1 : 0); }
__host__ __device__ constexpr int combo_q(int i) { return (i == 1 || i == 2) ? 1 : 0; }
__host__ __device__ constexpr int below_o(int i) { return i == 0 ? 0 : (i == 1 ? -1 : 1); }
// bit p set: output phase p of this axis is reached
__host__ __device__ constexpr int set_s2d(int i) {
    return (tap_of(combo_o(i), combo_q(i), 0) >= 0 ? 1 : 0) |
           (tap_of(combo_o(i), combo_q(i), 1) >= 0 ? 2 : 0);
}
__host__ __device__ constexpr int set_below(int i) {  // through either phase of `up`
    return ((tap_of(below_o(i), 0, 0) >= 0 || tap_of(below_o(i), 1, 0) >= 0) ? 1 : 0) |
           ((tap_of(below_o(i), 0, 1) >= 0 || tap_of(below_o(i), 1, 1) >= 0) ? 2 : 0);
}
struct OpShape {
    int a_off;  // 16-byte units from the stage base (below: from the slab's first plane)
    int dcol;   // first accumulator column
    int n;      // MMA N
};
__host__ __device__ constexpr OpShape shape_from_sets(int ys, int xs, int a_off) {
    int pmin = 4, pmax = -1;
    for (int pp = 0; pp < 4; ++pp)
        if (((ys >> (pp >> 1)) & 1) && ((xs >> (pp & 1)) & 1)) {
            pmin = pp < pmin ? pp : pmin;
            pmax = pp > pmax ? pp : pmax;
        }
    return OpShape{a_off, pmin * 32, (pmax - pmin + 1) * 32};
}
// op c = cy * 4 + cx of an S2D slab (16 per 16-channel slab)
__host__ __device__ constexpr OpShape s2d_shape(int c) {
    const int cy = c >> 2, cx = c & 3;
    return shape_from_sets(set_s2d(cy), set_s2d(cx),
                           (combo_q(cy) * 2 + combo_q(cx)) * (kPlane / 16) +
                               (1 + combo_o(cy)) * kHW + (1 + combo_o(cx)));
}
// op i = iy * 3 + ix of a slab of the tensor below (9 per 16-channel slab)
__host__ __device__ constexpr OpShape below_shape(int i) {
    return shape_from_sets(set_below(i / 3), set_below(i % 3),
                           (1 + below_o(i / 3)) * kHW + (1 + below_o(i % 3)));
}
// B offsets inside a slab's weight block, 16-byte units (an N-column block is N * 32 bytes)
__host__ __device__ constexpr int s2d_boff(int c) {
    int o = 0;
    for (int i = 0; i < c; ++i) o += 2 * s2d_shape(i).n;
    return o;
}
__host__ __device__ constexpr int below_boff(int i) {
    int o = 0;
    for (int j = 0; j < i; ++j) o += 2 * below_shape(j).n;
    return o;
}
constexpr int kS2dSlabUnits = s2d_boff(16);      // 2560 (40 KB)
constexpr int kBelowSlabUnits = below_boff(9);   // 1152 (18 KB)
static_assert(kS2dSlabUnits == 2560 && kBelowSlabUnits == 1152, "weight block sizes");
static_assert(s2d_shape(0).n == 128 && s2d_shape(0).dcol == 0, "op 0 must initialise all columns");

constexpr int kStemThreads = 256;   // 8 warps computing the Cin = 1 stem into the A stages (STEM = 1)
constexpr int kStagePxH = 2 * kHH + 2;                // u8 region of a tile: 38 rows x 22 bytes
// ... loaded as a TMA box of 48 x 38 bytes starting 16 bytes left of the tile (TMA needs the
// innermost coordinate to give a 16-byte aligned address: measured, scripts/microbench/
// tma_u8_test.cu), so the region's first column sits at byte kU8Off of each row
constexpr int kU8Row = 48;
constexpr int kU8Off = 13;                            // (2 X0 - 3) - (2 X0 - 16)
static_assert(kU8Off + 2 * kHW + 2 <= kU8Row, "u8 box too narrow");
constexpr int kU8Bytes = kU8Row * kStagePxH;          // 1824
constexpr int kU8Slot = 1920;                         // 128-byte aligned slots
constexpr int kU8Slots = 4;
// STEM = 2: the stem itself is a GEMM on the tensor cores. Rows = the 720 positions of a stage
// (phase, halo row, halo column: exactly the order of a stage's 16-byte cells), K = 16 = the nine
// taps as bf16 (u8 values are exact), two ones for the bias and zeros, N = 32 channels.
constexpr int kStemRows = 4 * kHW * kHH;                 // 720
constexpr int kStemChunks = (kStemRows + 127) / 128;     // 6 MMAs of 128 rows (the last is partial)
constexpr int kStemABytes = 2 * kStemRows * 16;          // [K half][row][8 bf16] = 23040 B
constexpr int kStemBBytes = 2 * 2 * 32 * 16;             // B_hi, B_lo: [K half][32][8 bf16] each
constexpr int kStemDSlots = 4;                           // 32-column accumulators, TMEM cols 384..511
constexpr int kStemDCol = 384;
// STEM = 3: the same GEMM with 16 stem warps (four groups of four, one accumulator slot each) and
// an f16 im2col operand built without a single int -> float conversion (see the kernel).
constexpr uint32_t kTraceTiles = 96;
constexpr int kTraceRoles = 8;
constexpr int kStem3Threads = 512;
constexpr int kStem3BuildWarps = 4;                      // the other 12 stem warps drain
#ifdef OGL_STEM3_SLOTS8   // experiment build: eight stem accumulators above TWO main ones
constexpr bool kS8 = true;
#else
constexpr bool kS8 = false;
#endif
constexpr int kStem3DSlots = kS8 ? 8 : 4;                // 32-column accumulators, TMEM cols 384..511
constexpr int kStem3DCol = kS8 ? 256 : 384;              // (three main accumulators below them)
constexpr int kSdMax = 8;                                // entries of the sd_full / sd_empty arrays
constexpr int kStemItems = 2 * kHW * kHH;                // 360 build items: (y phase, halo row, halo column)
__host__ __device__ constexpr int stem_threads(int stem) {
    return stem == 3 ? kStem3Threads : (stem ? kStemThreads : 0);
}
// K order of the STEM = 3 operand: k = 0..5 taps (0,0) (0,1) (1,0) (1,1) (2,0) (2,1), 6 tap (0,2),
// 7 one (bias hi), 8 tap (1,2), 9 one (bias lo), 10 tap (2,2), 11..15 zero
__host__ __device__ constexpr int stem3_k_of_tap(int tap) {
    return tap % 3 < 2 ? 2 * (tap / 3) + tap % 3 : (tap == 2 ? 6 : (tap == 5 ? 8 : 10));
}
// register budget per role (setmaxnreg; the launch allocates 72 x 896 = 64512 registers)
constexpr int kRegsLaunch3 = 72, kRegsCtl3 = 56, kRegsEpi3 = 96, kRegsBuild3 = 40;   // drain: 72
static_assert(kRegsCtl3 * 128 + kRegsEpi3 * 256 + kRegsBuild3 * 128 + kRegsLaunch3 * 384 ==
                  kRegsLaunch3 * (384 + kStem3Threads), "register pool of the STEM = 3 kernel");

struct S2dParams {
    const uint8_t* frames;   // STEM: u8 gray frames [B][2 H2][2 W2]
    StemPairs stem;          // STEM = 1: folded stem weights / 255 and bias, as channel pairs
    const uint8_t* stem_b;   // STEM >= 2: B operands of the stem GEMM (build_stem_tc_blob), device
    uint32_t stem_idesc;     // ... and its instruction descriptor (operand formats)
    int stem_lo;             // second MMA per chunk with the low parts of the split weights
    unsigned long long* trace;   // debug (OGL_TRACE=file): clock64 of CTA 0's hand-offs, [role][tile][8]
    const uint8_t* wblob;
    const float* btab;     // [3][3][32]: bias per (row class, column class), border pixels only
    float bias[32];        // bias of interior pixels (= btab[1][1]); constant-bank operands
    float head_w[32];      // 1x1 head weights (EPI_HEAD)
    int border_bias;       // the bias of border pixels differs (composed transposed conv)
    float* logits;
    uint8_t* mask;
    int32_t* area;
    __nv_bfloat16* out;
    __nv_bfloat16* out_pool;
    uint32_t wbytes;
    int n_stages;     // activation stages per tile: n_s2d slabs, then (has_below) the tensor below
    int n_s2d, has_below;
    int stage_src[kS2dMaxStages], stage_plane0[kS2dMaxStages];
    float head_b, logit_thr;
    int H2, W2, B;
    int tiles_x, tiles_y, num_tiles;
    unsigned long long magic_tx, magic_tpf;
    int nslots;
    int reverse;  // walk the tiles from the last frame to the first (L2 reuse, see conv_tc.cu)
    int dual;   // two MMA issuer warps on alternate tiles (needs nslots >= 2 * n_stages)
    int dbg;  // 1 no MMA, 2 no stores, 4 no epilogue work, 8 no activation loads (1-CTA form)
};

struct Tile {
    int n, y0, x0;
};
__device__ __forceinline__ Tile decode_tile(const S2dParams& p, int tile) {
    Tile t;
    const unsigned u = static_cast<unsigned>(tile);
    const unsigned n = static_cast<unsigned>((u * p.magic_tpf) >> 40);
    const unsigned rem = u - n * static_cast<unsigned>(p.tiles_x * p.tiles_y);
    const unsigned ty = static_cast<unsigned>((rem * p.magic_tx) >> 40);
    const unsigned tx = rem - ty * static_cast<unsigned>(p.tiles_x);
    t.n = static_cast<int>(n);
    t.y0 = static_cast<int>(ty) * kTH;
    t.x0 = static_cast<int>(tx) * kTW;
    return t;
}

// always bf16: the im2col operand of the tensor-core stem (u8 values, exact)
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) { return pack_x2<false>(lo, hi); }
// this unit's operand type (bf16, or f16 in s2d_tc_f16.cu)
__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) { return max_x2<kF16>(a, b); }

// CG = 2: the two CTAs of a cluster take neighbouring tiles and run every MMA as one 256-row
// tcgen05.mma.cta_group::2; each CTA keeps only HALF of every weight block resident (columns
// [r N/2, (r+1) N/2) of each op for cluster rank r), which frees shared memory for a deeper
// activation ring (6 instead of 3 stages beside the composed transposed conv) and halves the
// B-operand shared-memory reads.
// STEM = 1 (downs.0.net.3): the A stages are not loaded by TMA but COMPUTED in place by four extra
// warps from the u8 frame -- utils.py:235 (/255) + downs.0.net.0 (conv3x3 1->32, BN, ReLU) in fp32
// on the CUDA cores, rounded to bf16 straight into the UMMA operand layout -- so the stem's
// 4.19 MB-per-frame output tensor is never written to or read from HBM.
// STEM = 2: the same, but the stem is computed on the tensor cores: the stem warps only build the
// 720 x 16 im2col operand of a tile from the u8 region (bf16, exact) and move the GEMM's result
// from TMEM (ReLU folded into the bf16 conversion) into the A stages; warp 3 issues the stem MMAs.
// The weights are split w/255 = hi + lo in bf16 (two MMAs per 128 rows, fp32 accumulation), the
// bias rides on two constant-one K columns the same way: within ~2^-17 relative of the fp32 stem.
template <int EPI, int CG, int STEM>
__global__ void __launch_bounds__(kThreads + stem_threads(STEM), 1)
s2d_tc_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmB,
              const S2dParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t smem_base = (raw + 127u) & ~127u;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
    const uint32_t wbytes = p.wbytes / CG;                            // resident in this CTA
    const uint32_t w_s = smem_base;                                   // resident weights
    const uint32_t a_ring = w_s + ((wbytes + 127u) & ~127u);          // activation stages
    const uint32_t btab_s = a_ring + static_cast<uint32_t>(p.nslots) * kSlot;
    const uint32_t bar_base = btab_s + 9u * 32u * 4u;
    const uint32_t w_full = bar_base;
    const uint32_t a_full = w_full + 8u;
    const uint32_t a_empty = a_full + 8u * p.nslots;
    const uint32_t acc_full = a_empty + 8u * p.nslots;
    const uint32_t acc_empty = acc_full + 8u * kAccBufs;
    const uint32_t tmem_slot = acc_empty + 8u * kAccBufs;
    const uint32_t w_peer = tmem_slot + 16u;   // leader: the peer's weights are resident (CG = 2)
    // STEM = 1: ring of u8 regions (one per tile, loaded by TMA ahead of the stem warps)
    const uint32_t u8_full = w_peer + 16u;
    const uint32_t u8_empty = u8_full + 8u * kU8Slots;
    // STEM = 2: im2col buffers (two tiles) and accumulator slots of the stem GEMM
    const uint32_t sa_full = u8_empty + 8u * kU8Slots;
    const uint32_t sa_empty = sa_full + 16u;
    const uint32_t sd_full = sa_empty + 16u;
    const uint32_t sd_empty = sd_full + 8u * kSdMax;
    const uint32_t u8_s = (sd_empty + 8u * kSdMax + 127u) & ~127u;
    const uint32_t sA = u8_s + kU8Slots * kU8Slot;
    const uint32_t sB = sA + 2u * kStemABytes;
    // main accumulators (128 columns each) and, above them, the stem GEMM's 32-column ones
    constexpr int kBufs = STEM >= 2 ? ((STEM == 3 && kS8) ? 2 : 3) : kAccBufs;
    constexpr uint32_t kDSlots = STEM == 3 ? kStem3DSlots : kStemDSlots;
    constexpr uint32_t kDCol = STEM == 3 ? kStem3DCol : kStemDCol;
    uint8_t* gen = smem_raw - raw;  // generic pointer = gen + shared address
    float* btab_sp = reinterpret_cast<float*>(gen + btab_s);
    volatile uint32_t* tmem_slot_p = reinterpret_cast<volatile uint32_t*>(gen + tmem_slot);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // tiles of this CTA: CG = 1: blockIdx.x, + gridDim.x, ...; CG = 2: cluster c takes the tile
    // pairs c, c + clusters, ... and rank r the r-th tile of each pair (the last pair may lack one)
    const int unit0 = CG == 2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int unit_step = CG == 2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    const int num_units = CG == 2 ? (p.num_tiles + 1) >> 1 : p.num_tiles;
    auto tile_of = [&](int unit) {
        if (p.reverse) unit = num_units - 1 - unit;
        return CG == 2 ? 2 * unit + static_cast<int>(rank) : unit;
    };

    // ---------------------------------------------------------------- setup
    for (int i = threadIdx.x; i < 9 * 32; i += kThreads + stem_threads(STEM)) btab_sp[i] = p.btab[i];
    if (STEM >= 2) {
        if (threadIdx.x < kStemBBytes / 16)
            *reinterpret_cast<uint4*>(gen + sB + 16u * threadIdx.x) =
                reinterpret_cast<const uint4*>(p.stem_b)[threadIdx.x];
        fence_proxy_async();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmS);
        tma_prefetch_desc(&tmB);
        mbar_init(w_full, 1);
        mbar_init(w_peer, 1);
        if (STEM)
            for (int i = 0; i < kU8Slots; ++i) {
                mbar_init(u8_full + 8u * i, 1);
                mbar_init(u8_empty + 8u * i, STEM == 3 ? kStem3BuildWarps : kStemThreads / 32);
            }
        if (STEM >= 2) {
            for (int i = 0; i < 2; ++i) {
                mbar_init(sa_full + 8u * i, STEM == 3 ? kStem3BuildWarps : kStemThreads / 32);
                mbar_init(sa_empty + 8u * i, 1);
            }
            for (int i = 0; i < static_cast<int>(kDSlots); ++i) {
                mbar_init(sd_full + 8u * i, 1);
                // STEM 2: slot i is free again (the four warps of the draining group). STEM 3:
                // entries 0 / 1: chunks 2..5 of an even / odd tile are in registers (4 x 4 warps),
                // entries 2 / 3: chunks 0, 1 (2 x 4 warps)
                mbar_init(sd_empty + 8u * i, STEM == 3 ? ((i < 2 || kS8) ? 16 : 8) : 4);
            }
        }
        for (int i = 0; i < p.nslots; ++i) {
            // one arrival per stem warp that writes the stages (STEM 3: the 12 drain warps)
            mbar_init(a_full + 8u * i, STEM == 3 ? kStem3Threads / 32 - kStem3BuildWarps
                                                 : (STEM ? kStemThreads / 32 : 1));
            mbar_init(a_empty + 8u * i, 1);
        }
        for (int i = 0; i < kBufs; ++i) {
            mbar_init(acc_full + 8u * i, 1);
            mbar_init(acc_empty + 8u * i, 128 * CG);   // CG = 2: the epilogues of both CTAs
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        if (CG == 2) tmem_alloc_pair(tmem_slot, 512);
        else tmem_alloc(tmem_slot, 512);
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();  // the peer's barriers exist before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_p;
    // debug timeline of CTA 0: role r, its tile number i < kTraceTiles, event e < 8
    auto trace = [&](int r, uint32_t i, int e) {
        if (p.trace && blockIdx.x == 0 && lane == 0 && i < kTraceTiles)
            p.trace[(static_cast<uint32_t>(r) * kTraceTiles + i) * 8u + e] = clock64();
    };
    // STEM = 3: 896 threads leave 72 registers each; the control warps and the stem warps hand
    // theirs to the two epilogue groups, whose 4 x 32 accumulator values per thread need 96. Each
    // warp group re-allocates at the top of its own branch (setmaxnreg is per warp group, and
    // ptxas budgets the code below each one accordingly).
    if (warp >= kThreads / 32) {
        if (STEM == 3) {
            // ============== stem on the tensor cores, 16 warps: f16 im2col in, A stages out
            // The stem is a chain of hand-offs per tile (u8 region -> im2col -> MMAs -> TMEM ->
            // A stages -> main MMAs), and what bounds the launch is how far the chains of consecutive
            // tiles overlap, not the work in them (measured: with every MMA, every TMEM read and the
            // epilogue switched off the 8-warp form still needed 2 800 cycles per tile). So the roles
            // are separated and every buffer between them holds more than one tile:
            //   * warps 0-3 BUILD only, one tile ahead (two im2col buffers);
            //   * warps 4-15 DRAIN only: three groups of four, group g takes the chunks c with
            //     c % 3 == g (two per tile) out of a ring of four 32-column accumulators;
            //   * the issuer (warp 3) waits for accumulator slots twice per tile, not per chunk.
            // Build: one item = the two x phases of one (y phase, halo row, halo column): the 3 x 4
            // bytes around it are read as aligned words (2 LDS.32 + 1 PRMT per window row, the
            // selector fixed per item), and every K pair of the operand is ONE byte permute that
            // puts each u8 under the f16 exponent of 1024 (0x64 b = 1024 + b exactly) followed by ONE
            // HFMA2, (1024 + b) * keep - 1024 * keep: no I2F, no F2FP, no separate masking (keep = 0
            // outside the image, where the stem's OUTPUT is zero: conv2's padding). The constant-one K
            // columns that carry the bias come out of the same permutes. 12 ALU instructions per
            // operand row instead of 9 LDS.U8 + 9 I2F + 5 F2FP + 5 LOP3.
            const int st = threadIdx.x - kThreads;
            const int sw = st >> 5;
            const int W = 2 * p.W2, H = 2 * p.H2;
            if (sw < kStem3BuildWarps) {
                setmaxnreg_dec<kRegsBuild3>();   // (the drain warps keep the launch value)
                constexpr int kRounds = (kStemItems + 32 * kStem3BuildWarps - 1) / (32 * kStem3BuildWarps);
                int woff[kRounds], row0[kRounds], ily[kRounds], ilx[kRounds];
                uint32_t wsel[kRounds];
#pragma unroll
                for (int rr = 0; rr < kRounds; ++rr) {
                    const int item = min(st + rr * 32 * kStem3BuildWarps, kStemItems - 1);
                    const int ipy = item / (kHW * kHH), irem = item - ipy * (kHW * kHH);
                    const int ihy = irem / kHW, ihx = irem - ihy * kHW;
                    ily[rr] = 2 * ihy + ipy;
                    ilx[rr] = 2 * ihx;   // region pixel of (x phase 0, tap (0, 0))
                    const int boff = ily[rr] * kU8Row + kU8Off + ilx[rr];   // its byte in the u8 box: odd
                    woff[rr] = boff & ~3;
                    wsel[rr] = (boff & 3) == 1 ? 0x4321u : 0x6543u;
                    row0[rr] = 2 * ipy * (kHW * kHH) + irem;   // operand row of x phase 0 (+180: phase 1)
                }
                uint32_t iu = 0;
                for (int unit = unit0; unit < num_units; unit += unit_step, ++iu) {
                    const Tile t = decode_tile(p, tile_of(unit));
                    const uint32_t us = iu % kU8Slots;
                    const uint8_t* u8p = gen + u8_s + us * kU8Slot;
                    uint8_t* abuf = gen + sA + (iu & 1u) * kStemABytes;
                    mbar_wait_relaxed(u8_full + 8u * us, (iu / kU8Slots) & 1u);
                    if (sw == 0) trace(0, iu, 0);
                    mbar_wait_relaxed(sa_empty + 8u * (iu & 1u), ((iu >> 1) & 1u) ^ 1u);
                    if (sw == 0) trace(0, iu, 1);
#pragma unroll
                    for (int rr = 0; rr < kRounds; ++rr) {
                        if (st + rr * 32 * kStem3BuildWarps >= kStemItems) break;
                        const int gy = 2 * t.y0 - 2 + ily[rr], gx = 2 * t.x0 - 2 + ilx[rr];
                        uint32_t lo[3], hi[3], wd[3];
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            lo[d] = *reinterpret_cast<const uint32_t*>(u8p + woff[rr] + d * kU8Row);
                            hi[d] = *reinterpret_cast<const uint32_t*>(u8p + woff[rr] + d * kU8Row + 4);
                        }
#pragma unroll
                        for (int d = 0; d < 3; ++d) wd[d] = __byte_perm(lo[d], hi[d], wsel[rr]);
                        const bool iny = gy >= 0 && gy < H && !(p.dbg & 64);
#pragma unroll
                        for (int px = 0; px < 2; ++px) {
                            const bool in = iny && gx + px >= 0 && gx + px < W;
                            const uint32_t keep = in ? 0x3c003c00u : 0u;        // (1, 1)
                            const uint32_t off2 = in ? 0xe400e400u : 0u;        // (-1024, -1024)
                            const uint32_t off1 = in ? 0x0000e400u : 0u;        // (-1024, 0)
                            // bytes 4..7 of the permutes: 0x64 (f16 exponent of 1024), 0x00, 0x00, 0x3c
                            const uint32_t pair_sel = px ? 0x4241u : 0x4140u;   // (b[px], b[px + 1])
                            const uint32_t one_sel = px ? 0x7543u : 0x7542u;    // (b[px + 2], 1.0)
                            const uint32_t zero_sel = px ? 0x5543u : 0x5542u;   // (b[px + 2], 0.0)
                            uint4 k0, k1;
                            k0.x = fma_f16x2(__byte_perm(wd[0], 0x64646464u, pair_sel), keep, off2);
                            k0.y = fma_f16x2(__byte_perm(wd[1], 0x64646464u, pair_sel), keep, off2);
                            k0.z = fma_f16x2(__byte_perm(wd[2], 0x64646464u, pair_sel), keep, off2);
                            k0.w = fma_f16x2(__byte_perm(wd[0], 0x3c000064u, one_sel), keep, off1);
                            k1.x = fma_f16x2(__byte_perm(wd[1], 0x3c000064u, one_sel), keep, off1);
                            k1.y = fma_f16x2(__byte_perm(wd[2], 0x3c000064u, zero_sel), keep, off1);
                            k1.z = 0u;
                            k1.w = 0u;
                            const int r = row0[rr] + px * (kHW * kHH);
                            if (p.dbg & 512) continue;
                            *reinterpret_cast<uint4*>(abuf + r * 16) = k0;
                            *reinterpret_cast<uint4*>(abuf + kStemRows * 16 + r * 16) = k1;
                        }
                    }
                    fence_proxy_async();   // visible to the tensor core's reads
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(sa_full + 8u * (iu & 1u));
                        mbar_arrive(u8_empty + 8u * us);
                    }
                    if (sw == 0) trace(0, iu, 2);
                }
            } else {
                // Drain, in two steps per tile so that the A stages are busy as briefly as possible:
                // (1) both chunks of this group leave TMEM as soon as their MMAs are done -- ReLU,
                // bf16, 2 x 4 packed 16-byte cells kept in registers -- which also frees the
                // accumulator slots for the issuer early; (2) only the eight stores wait for the two
                // stages of the ring (CTA 0's timeline, scripts/stem_trace.py: with load and store
                // both behind the stage wait a tile's stages were refilled 3 300 cycles after the main
                // MMAs had released them, and the ring holds two tiles).
                const uint32_t dg = static_cast<uint32_t>(sw - kStem3BuildWarps) >> 2;   // 0..2
                const uint32_t lane_sel = static_cast<uint32_t>((sw & 3) * 32) << 16;
                const bool trd = sw == kStem3BuildWarps;
                uint32_t iu = 0;
                Ring rs;
                for (int unit = unit0; unit < num_units; unit += unit_step, ++iu) {
                    uint4 q[2][4];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t j = dg + 3u * h;
                        const uint32_t c = static_cast<uint32_t>(kStemChunks) * iu + j;
                        const uint32_t slot = c % kStem3DSlots;
                        mbar_wait_relaxed(sd_full + 8u * slot, (c / kStem3DSlots) & 1u);
                        if (trd) trace(1, iu, h ? 4 : 1);
                        tc_fence_after();
                        uint32_t v[32];
                        if (p.dbg & 1024) {   // experiment: no TMEM read
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = 0u;
                        } else {
                            tmem_ld32(tmem_base + lane_sel + kStem3DCol + slot * 32u, v);
                            tmem_ld_wait();
                        }
                        tc_fence_before();
                        __syncwarp();
                        // the values are in registers. Slot (6 T + k) % 4 of tile T was last used by
                        // chunk 6 T + k - 4: chunks 2..5 of tile T - 1 for k = 0..3, chunks 0, 1 of
                        // tile T itself for k = 4, 5 -- TWO barrier waits per tile for the issuer
                        // instead of six.
                        // (kS8, eight slots: slot (6 T + k) % 8 was last used by chunks 4, 5 of T - 2
                        //  for k = 0, 1 and chunks 0..3 of T - 1 for k = 2..5; a group drains its chunk
                        //  of T - 2 before that of T - 1, so "chunks 0..3 of T - 1" frees all six: ONE wait)
                        if (kS8) {
                            if (lane == 0 && j < 4) mbar_arrive(sd_empty + 8u * (iu & 1u));
                        } else if (lane == 0) {
                            mbar_arrive(sd_empty + 8u * ((j < 2 ? 2u : 0u) + (iu & 1u)));
                        }
                        if (trd) trace(1, iu, h ? 5 : 2);
#pragma unroll
                        for (int g = 0; g < 4; ++g)   // 8-channel groups: (stage, plane group)
                            q[h][g] = make_uint4(
                                pack_relu_bf16x2(__uint_as_float(v[8 * g + 0]), __uint_as_float(v[8 * g + 1])),
                                pack_relu_bf16x2(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3])),
                                pack_relu_bf16x2(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5])),
                                pack_relu_bf16x2(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7])));
                    }
                    const uint32_t s0 = rs.slot, ph0 = rs.phase;
                    rs.next(p.nslots);
                    const uint32_t s1 = rs.slot, ph1 = rs.phase;
                    rs.next(p.nslots);
                    mbar_wait_relaxed(a_empty + 8u * s0, ph0 ^ 1u);
                    mbar_wait_relaxed(a_empty + 8u * s1, ph1 ^ 1u);
                    if (trd) trace(1, iu, 0);
                    uint8_t* stage0 = gen + a_ring + s0 * kSlot;
                    uint8_t* stage1 = gen + a_ring + s1 * kSlot;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int r = static_cast<int>(dg + 3u * h) * 128 + (sw & 3) * 32 + lane;
                        if (r < kStemRows && !(p.dbg & 256)) {
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                uint8_t* dst = ((g >> 1) ? stage1 : stage0) + (g & 1) * 4 * kPlane + r * 16;
                                *reinterpret_cast<uint4*>(dst) = q[h][g];
                            }
                        }
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(a_full + 8u * s0);
                        mbar_arrive(a_full + 8u * s1);
                    }
                    if (trd) trace(1, iu, 6);
                }
            }
        } else if (STEM == 2) {
            // ============================ stem on the tensor cores: im2col in, A stages out
            const int st = threadIdx.x - kThreads;
            const int sw = st >> 5;            // stem warp 0..7; TMEM lane quarter = sw & 3 (= warp & 3)
            const int grp = sw >> 2;           // drains the chunks j with (j & 1) == grp
            const uint32_t lane_sel = static_cast<uint32_t>((sw & 3) * 32) << 16;
            const int W = 2 * p.W2, H = 2 * p.H2;
            // The rows a thread builds (st, st + 256, st + 512) are the same in every tile: their
            // position inside the u8 region and the stage is computed once. Rows past the end (the
            // third row of threads >= 208) read row 719 and store nothing.
            constexpr int kRowsPerThread = (kStemRows + kStemThreads - 1) / kStemThreads;   // 3
            int row_src[kRowsPerThread], row_ly[kRowsPerThread], row_lx[kRowsPerThread];
#pragma unroll
            for (int rr = 0; rr < kRowsPerThread; ++rr) {
                const int r = min(st + rr * kStemThreads, kStemRows - 1);
                const int ph = r / (kHW * kHH), pos = r - ph * (kHW * kHH);
                const int hy = pos / kHW, hx = pos - hy * kHW;
                row_ly[rr] = 2 * hy + (ph >> 1);
                row_lx[rr] = 2 * hx + (ph & 1);
                row_src[rr] = row_ly[rr] * kU8Row + kU8Off + row_lx[rr];
            }
            // im2col of tile number iu of this CTA into buffer iu & 1
            auto build = [&](uint32_t iu, int unit) {
                const Tile t = decode_tile(p, tile_of(unit));
                const int gy0 = 2 * t.y0 - 2, gx0 = 2 * t.x0 - 2;   // frame coordinates of stage pixel (0, 0)
                const uint32_t us = iu % kU8Slots;
                const uint8_t* u8p = gen + u8_s + us * kU8Slot;      // rows of kU8Row bytes (TMA box)
                uint8_t* abuf = gen + sA + (iu & 1u) * kStemABytes;
                mbar_wait_relaxed(u8_full + 8u * us, (iu / kU8Slots) & 1u);
                mbar_wait_relaxed(sa_empty + 8u * (iu & 1u), ((iu >> 1) & 1u) ^ 1u);
                // all 27 byte loads first, then the conversions, then the stores: three independent
                // latency chains per thread instead of one after the other
                uint32_t b[kRowsPerThread][9];
#pragma unroll
                for (int rr = 0; rr < kRowsPerThread; ++rr)
#pragma unroll
                    for (int k = 0; k < 9; ++k) b[rr][k] = u8p[row_src[rr] + (k / 3) * kU8Row + (k % 3)];
#pragma unroll
                for (int rr = 0; rr < kRowsPerThread; ++rr) {
                    const int r = st + rr * kStemThreads;
                    const int gy = gy0 + row_ly[rr], gx = gx0 + row_lx[rr];
                    // outside the image the stem's OUTPUT is zero (conv2's padding): a zero row,
                    // bias columns included
                    const bool in = gy >= 0 && gy < H && gx >= 0 && gx < W && !(p.dbg & 64);
                    float v[9];
#pragma unroll
                    for (int k = 0; k < 9; ++k) v[k] = static_cast<float>(b[rr][k]);
                    const uint32_t keep = in ? 0xffffffffu : 0u;
                    const uint4 k0 = make_uint4(pack_bf16x2(v[0], v[1]) & keep, pack_bf16x2(v[2], v[3]) & keep,
                                                pack_bf16x2(v[4], v[5]) & keep, pack_bf16x2(v[6], v[7]) & keep);
                    const uint4 k1 = make_uint4(pack_bf16x2(v[8], 1.f) & keep, pack_bf16x2(1.f, 0.f) & keep, 0u, 0u);
                    if (r < kStemRows) {
                        *reinterpret_cast<uint4*>(abuf + r * 16) = k0;
                        *reinterpret_cast<uint4*>(abuf + kStemRows * 16 + r * 16) = k1;
                    }
                }
                fence_proxy_async();   // visible to the tensor core's reads
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(sa_full + 8u * (iu & 1u));
                    mbar_arrive(u8_empty + 8u * us);
                }
            };
            // result of the stem GEMM: TMEM -> ReLU -> bf16 -> the tile's two A stages (row r of the
            // GEMM is the 16-byte cell r of each 8-channel group of a stage). One chunk = 128 rows.
            auto drain_chunk = [&](uint32_t iu, int j, uint8_t* stage0, uint8_t* stage1) {
                const uint32_t c = static_cast<uint32_t>(kStemChunks) * iu + j;
                const uint32_t slot = c % kStemDSlots;
                mbar_wait_relaxed(sd_full + 8u * slot, (c / kStemDSlots) & 1u);
                tc_fence_after();
                uint32_t v[32];
                tmem_ld32(tmem_base + lane_sel + kStemDCol + slot * 32u, v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(sd_empty + 8u * slot);   // the values are in registers
                const int r = j * 128 + (sw & 3) * 32 + lane;
                if (r < kStemRows) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {   // 8-channel groups: (stage, plane group)
                        const uint4 q = make_uint4(
                            pack_relu_bf16x2(__uint_as_float(v[8 * g + 0]), __uint_as_float(v[8 * g + 1])),
                            pack_relu_bf16x2(__uint_as_float(v[8 * g + 2]), __uint_as_float(v[8 * g + 3])),
                            pack_relu_bf16x2(__uint_as_float(v[8 * g + 4]), __uint_as_float(v[8 * g + 5])),
                            pack_relu_bf16x2(__uint_as_float(v[8 * g + 6]), __uint_as_float(v[8 * g + 7])));
                        uint8_t* dst = ((g >> 1) ? stage1 : stage0) + (g & 1) * 4 * kPlane + r * 16;
                        *reinterpret_cast<uint4*>(dst) = q;
                    }
                }
            };
            // Per tile: first chunk, then the im2col of the NEXT tile, then the other two chunks. The
            // accumulator ring has 4 slots for 6 chunks, so chunks 2..5 of a tile are issued only when
            // earlier chunks have been drained: with this order their MMAs run while this warp builds.
            uint32_t iu = 0;
            int unit = unit0;
            if (unit < num_units) build(0, unit);
            for (; unit < num_units; unit += unit_step, ++iu) {
                const uint32_t it0 = 2u * iu, it1 = 2u * iu + 1u;
                const uint32_t s0 = it0 % p.nslots, s1 = it1 % p.nslots;
                mbar_wait_relaxed(a_empty + 8u * s0, ((it0 / p.nslots) & 1u) ^ 1u);
                mbar_wait_relaxed(a_empty + 8u * s1, ((it1 / p.nslots) & 1u) ^ 1u);
                uint8_t* stage0 = gen + a_ring + s0 * kSlot;
                uint8_t* stage1 = gen + a_ring + s1 * kSlot;
                drain_chunk(iu, grp, stage0, stage1);
                if (unit + unit_step < num_units) build(iu + 1, unit + unit_step);
                drain_chunk(iu, grp + 2, stage0, stage1);
                drain_chunk(iu, grp + 4, stage0, stage1);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(a_full + 8u * s0);
                    mbar_arrive(a_full + 8u * s1);
                }
            }
        } else if (STEM == 1) {
            // ====================================== stem: compute the A stages from the u8 frame
            // A tile needs the stem output on 10 x 18 half-resolution positions x 4 phases, of which
            // the outermost ring is read through one phase only: 34 x 18 full-resolution pixels.
            // Per pixel and stage 16 channels = 144 FFMAs with constant-bank weights, bias, ReLU,
            // bf16, two 16-byte stores into the operand layout.
            const int st = threadIdx.x - kThreads;
            const int W = 2 * p.W2, H = 2 * p.H2;
            constexpr int kNeedH = 2 * kHH - 2, kNeedW = 2 * kHW - 2;   // 34 x 18
            uint32_t it = 0, iu = 0;
            for (int unit = unit0; unit < num_units; unit += unit_step, ++iu) {
                const Tile t = decode_tile(p, tile_of(unit));
                const int gy0 = 2 * t.y0 - 3, gx0 = 2 * t.x0 - 3;   // frame coordinates of u8p[0][0]
                const uint32_t us = iu % kU8Slots;
                const uint8_t* u8p = gen + u8_s + us * kU8Slot;      // rows of kU8Row bytes (TMA box)
                mbar_wait_relaxed(u8_full + 8u * us, (iu / kU8Slots) & 1u);
                // One work item = 3 horizontally adjacent pixels of one row (34 rows x 6 triples = 204
                // items per tile, thread st takes item st): every weight fetched from the constant
                // bank feeds 3 FFMAs, the 3x5 input window is read once.
                constexpr int kItems = kNeedH * (kNeedW / 3);
                static_assert(kNeedW % 3 == 0 && kItems <= kStemThreads, "stem work split");
                const bool active = st < kItems;
                const int ly = 1 + st / (kNeedW / 3), lx0 = 1 + 3 * (st % (kNeedW / 3));
                const int gy = gy0 + 1 + ly;
                unsigned long long in[3][5];   // each input value in both halves of a register pair
                uint32_t inside[3];   // all-ones inside the image, 0 outside (a mask, not a branch:
                                      // the six accumulator chains of an item stay interleaved)
                if (active) {
#pragma unroll
                    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                        for (int dx = 0; dx < 5; ++dx)
                            in[dy][dx] =
                                dup2(static_cast<float>(u8p[(ly + dy) * kU8Row + kU8Off + lx0 + dx]));
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const int gx = gx0 + 1 + lx0 + q;
                        inside[q] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? 0xffffffffu : 0u;
                    }
                }
#pragma unroll   // `half` must be a compile-time constant: weights are constant-bank operands
                for (int half = 0; half < 2; ++half, ++it) {
                    const uint32_t slot = it % p.nslots;
                    mbar_wait_relaxed(a_empty + 8u * slot, ((it / p.nslots) & 1u) ^ 1u);
                    uint8_t* stage = gen + a_ring + slot * kSlot;
                    if (active && !(p.dbg & 64)) {
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            uint32_t pk[3][4];
#pragma unroll
                            for (int c2 = 0; c2 < 4; ++c2) {
                                // channel pair (2 cp, 2 cp + 1) of the 3 pixels: packed FFMA2s
                                const int cp = half * 8 + g * 4 + c2;
                                unsigned long long a[3];
#pragma unroll
                                for (int q = 0; q < 3; ++q) a[q] = as_u64(p.stem.bp[cp]);
#pragma unroll
                                for (int k = 0; k < 9; ++k) {
                                    const unsigned long long w = as_u64(p.stem.wp[cp * 9 + k]);
#pragma unroll
                                    for (int q = 0; q < 3; ++q) a[q] = fma2(in[k / 3][k % 3 + q], w, a[q]);
                                }
#pragma unroll
                                for (int q = 0; q < 3; ++q) {   // zero outside the image: conv2's padding
                                    const float lo = __uint_as_float(static_cast<uint32_t>(a[q]));
                                    const float hi = __uint_as_float(static_cast<uint32_t>(a[q] >> 32));
                                    pk[q][c2] = pack_relu_bf16x2(lo, hi) & inside[q];
                                }
                            }
#pragma unroll
                            for (int q = 0; q < 3; ++q) {
                                const int lx = lx0 + q;
                                uint8_t* dst = stage + ((ly & 1) * 2 + (lx & 1)) * kPlane +
                                               ((ly >> 1) * kHW + (lx >> 1)) * 16 + g * 4 * kPlane;
                                *reinterpret_cast<uint4*>(dst) =
                                    make_uint4(pk[q][0], pk[q][1], pk[q][2], pk[q][3]);
                            }
                        }
                    }
                    fence_proxy_async();   // this thread's stores visible to the tensor core's reads
                    __syncwarp();
                    if (lane == 0) mbar_arrive(a_full + 8u * slot);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(u8_empty + 8u * us);
            }
        }
    } else if (warp < 4) {
        if (STEM == 3) setmaxnreg_dec<kRegsCtl3>();
        if (warp == 0 && STEM) {
            // ================================ u8 regions of the tiles, by TMA, ahead of the stem warps
            // box = 48 x 38 bytes at (2 X0 - 16, 2 Y0 - 3): conv1's zero padding is the OOB fill
            if (lane == 0) {
                uint32_t iu = 0;
                for (int unit = unit0; unit < num_units; unit += unit_step, ++iu) {
                    const Tile t = decode_tile(p, tile_of(unit));
                    const uint32_t us = iu % kU8Slots;
                    mbar_wait_relaxed(u8_empty + 8u * us, ((iu / kU8Slots) & 1u) ^ 1u);
                    trace(7, iu, 0);
                    mbar_arrive_expect_tx(u8_full + 8u * us, kU8Bytes);
                    tma_load_3d(u8_s + us * kU8Slot, &tmS, u8_full + 8u * us, 2 * t.x0 - 16, 2 * t.y0 - 3,
                                t.n);
                }
            }
        } else if (warp == 0 && !STEM) {
            // ================================================ activation producer
            if (lane == 0) {
                Ring rs;
                uint32_t ti = 0;
                for (int unit = unit0; unit < num_units; unit += unit_step, ++ti) {
                    int tile = tile_of(unit);
                    if (tile >= p.num_tiles) tile = p.num_tiles - 1;   // tail of the last pair: a duplicate
                    const Tile t = decode_tile(p, tile);
                    for (int s = 0; s < p.n_stages; ++s, rs.next(p.nslots)) {
                        const uint32_t slot = rs.slot;
                        const uint32_t ph = rs.phase;
                        mbar_wait_relaxed(a_empty + 8u * slot, ph ^ 1u);
                        trace(7, ti, s);
                        const uint32_t dst = a_ring + slot * kSlot;
                        if (CG == 1 && (p.dbg & 8)) {   // experiment: no activation loads
                            mbar_arrive(a_full + 8u * slot);
                            continue;
                        }
                        if (CG == 1) {
                            const uint32_t full = a_full + 8u * slot;
                            mbar_arrive_expect_tx(full, kSlot);
                            if (p.stage_src[s] == 0)
                                tma_load_5d(dst, &tmS, full, (t.x0 - 1) * 8, t.y0 - 1, 0,
                                            p.stage_plane0[s], t.n);
                            else
                                tma_load_4d(dst, &tmB, full, (t.x0 - 1) * 8, t.y0 - 1,
                                            p.stage_plane0[s], t.n);
                        } else {
                            // the leader's barrier counts the bytes of both CTAs
                            if (rank == 0) mbar_arrive_expect_tx(a_full + 8u * slot, 2u * kSlot);
                            const uint32_t full = map_to_cta(a_full + 8u * slot, 0);
                            if (p.stage_src[s] == 0)
                                tma_load_5d_pair(dst, &tmS, full, (t.x0 - 1) * 8, t.y0 - 1, 0,
                                                 p.stage_plane0[s], t.n);
                            else
                                tma_load_4d_pair(dst, &tmB, full, (t.x0 - 1) * 8, t.y0 - 1,
                                                 p.stage_plane0[s], t.n);
                        }
                    }
                }
            }
        } else if (warp == 3) {
            // ================================================== weights, once per CTA
            // (CG = 2: the blob holds rank 0's halves, then rank 1's)
            if (lane == 0) {
                const uint8_t* src = p.wblob + static_cast<size_t>(rank) * wbytes;
                mbar_arrive_expect_tx(w_full, wbytes);
                for (uint32_t off = 0; off < wbytes; off += kWChunk) {
                    const uint32_t n = wbytes - off < kWChunk ? wbytes - off : kWChunk;
                    bulk_load(w_s + off, src + off, n, w_full);
                }
                if (CG == 2 && rank == 1) {
                    // tell the leader that this CTA's weights are resident too
                    mbar_wait(w_full, 0);
                    mbar_arrive_cluster(map_to_cta(w_peer, 0));
                }
            }
            if (STEM >= 2) {
                // ============================================== issuer of the stem GEMM
                // per tile 6 x (A[128 rows x 16] x B_hi, then x B_lo) into a ring of four 32-column
                // accumulators; the whole warp walks the loop, one elected lane issues
                __syncwarp();
                constexpr uint32_t a_hi_s = (128u >> 4) | (1u << 14);                 // SBO = 8 rows
                constexpr uint32_t a_lbo_s = (static_cast<uint32_t>(kStemRows * 16) >> 4) << 16;
                constexpr uint64_t b_hi_s = static_cast<uint64_t>((128u >> 4) | (1u << 14)) << 32;
                const uint64_t bd_hi = b_hi_s | ((sB >> 4) | (32u << 16));            // LBO = 16 N = 512 B
                const uint64_t bd_lo = bd_hi + (1024u >> 4);
                if (STEM == 3) {
                    // Two slot waits per tile (plus the im2col buffer) and the tile's 6 x 2 MMAs and 7
                    // commits as straight-line code: the slot of chunk j of tile T is (6 T + j) % 4,
                    // which depends on T % 2 only, so both variants are unrolled with compile-time
                    // descriptor offsets (operands in uniform registers, as in the main issuers).
                    // Measured before (CTA 0's timeline, scripts/stem_trace.py): with a wait + elect +
                    // issue round per chunk this warp needed ~3 800 cycles per tile and paced the
                    // whole kernel (period 4 300 cycles per tile).
                    const uint32_t idesc = p.stem_idesc;
                    const bool lo = p.stem_lo != 0, no_mma = (p.dbg & 1) != 0;
                    const uint64_t ad0 = (static_cast<uint64_t>(a_hi_s) << 32) | ((sA >> 4) | a_lbo_s);
                    const uint32_t d0 = tmem_base + kStem3DCol;
                    auto issue = [&](auto ph_tag, auto j0_tag, auto j1_tag) {
                        constexpr uint32_t PH = decltype(ph_tag)::value;   // T % 4
                        constexpr uint32_t kTileA = (PH & 1u) * (kStemABytes >> 4);
#pragma unroll
                        for (uint32_t j = decltype(j0_tag)::value; j < decltype(j1_tag)::value; ++j) {
                            const uint32_t slot = (6u * PH + j) & static_cast<uint32_t>(kStem3DSlots - 1);
                            const uint64_t ad = ad0 + (kTileA + j * (2048u >> 4));
                            if (!no_mma) {
                                umma_bf16(d0 + slot * 32u, ad, bd_hi, idesc, 0u);
                                if (lo) umma_bf16(d0 + slot * 32u, ad, bd_lo, idesc, 1u);
                            }
                            umma_commit(sd_full + 8u * slot);
                        }
                        if (decltype(j1_tag)::value == kStemChunks) umma_commit(sa_empty + 8u * (PH & 1u));
                    };
                    using U0 = std::integral_constant<uint32_t, 0>;
                    using U1 = std::integral_constant<uint32_t, 1>;
                    using U2 = std::integral_constant<uint32_t, 2>;
                    using U3 = std::integral_constant<uint32_t, 3>;
                    using U4 = std::integral_constant<uint32_t, 4>;
                    using U6 = std::integral_constant<uint32_t, 6>;
                    auto issue_ph = [&](uint32_t ph, auto j0_tag, auto j1_tag) {
                        switch (ph) {
                            case 0: issue(U0{}, j0_tag, j1_tag); break;
                            case 1: issue(U1{}, j0_tag, j1_tag); break;
                            case 2: issue(U2{}, j0_tag, j1_tag); break;
                            default: issue(U3{}, j0_tag, j1_tag); break;
                        }
                    };
                    uint32_t iu = 0;
                    for (int unit = unit0; unit < num_units; unit += unit_step, ++iu) {
                        const uint32_t par = iu & 1u;
                        mbar_wait(sa_full + 8u * par, (iu >> 1) & 1u);
                        trace(2, iu, 0);
                        if (iu) mbar_wait(sd_empty + 8u * (par ^ 1u), ((iu - 1u) >> 1) & 1u);
                        trace(2, iu, 1);
                        tc_fence_after();
                        if (kS8) {
                            if (elect_one()) issue_ph(iu & 3u, U0{}, U6{});
                            __syncwarp();
                            trace(2, iu, 4);
                            continue;
                        }
                        if (elect_one()) issue_ph(iu & 3u, U0{}, U4{});
                        __syncwarp();
                        trace(2, iu, 2);
                        mbar_wait(sd_empty + 8u * (2u + par), (iu >> 1) & 1u);
                        trace(2, iu, 3);
                        tc_fence_after();
                        if (elect_one()) issue_ph(iu & 3u, U4{}, U6{});
                        __syncwarp();
                        trace(2, iu, 4);
                    }
                } else {
                    uint32_t c = 0, iu = 0;
                    for (int unit = unit0; unit < num_units; unit += unit_step, ++iu) {
                        const uint32_t b = iu & 1u;
                        mbar_wait(sa_full + 8u * b, (iu >> 1) & 1u);
                        trace(2, iu, 0);
                        tc_fence_after();
#pragma unroll 1
                        for (int j = 0; j < kStemChunks; ++j, ++c) {
                            const uint32_t slot = c % kDSlots;
                            mbar_wait(sd_empty + 8u * slot, ((c / kDSlots) & 1u) ^ 1u);
                            if (j == 0 || j == kStemChunks - 1) trace(2, iu, j == 0 ? 1 : 3);
                            tc_fence_after();
                            if (elect_one()) {
                                const uint64_t ad =
                                    (static_cast<uint64_t>(a_hi_s) << 32) |
                                    (((sA + b * kStemABytes + static_cast<uint32_t>(j) * 2048u) >> 4) | a_lbo_s);
                                const uint32_t d = tmem_base + kDCol + slot * 32u;
                                // STEM 2: bf16 x bf16 (in either translation unit); STEM 3: f16 x f16
                                const uint32_t idesc = p.stem_idesc;
                                if (!(p.dbg & 1)) {
                                    umma_bf16(d, ad, bd_hi, idesc, 0u);
                                    if (p.stem_lo) umma_bf16(d, ad, bd_lo, idesc, 1u);
                                }
                                umma_commit(sd_full + 8u * slot);
                                if (j == kStemChunks - 1) umma_commit(sa_empty + 8u * b);
                            }
                            __syncwarp();
                            if (j == 0 || j == kStemChunks - 1) trace(2, iu, j == 0 ? 2 : 4);
                        }
                    }
                }
            }
        } else if (rank == 0) {
            // ================================================= MMA issuers (two warps)
            // The tensor pipe accepts an MMA only when the previous one is (nearly) done, and a wait
            // on an mbarrier costs the issuing warp ~120 cycles even when the phase completed long
            // ago, so with ONE issuer every barrier wait between two groups of MMAs is a bubble in the
            // pipe (measured: 82 instead of 48 cycles per N = 64 MMA with a wait every 4 MMAs;
            // profiles/microbench_mma_issuer_bubbles_r01.txt). Two warps issue alternate tiles (into
            // different accumulator buffers, so the summation order inside a tile is unchanged):
            // while one waits, the other's MMAs keep the pipe full -- 48.0 cycles in the same test.
            // An issuer may then wait for a ring slot's NEXT use while the other issuer has not yet
            // seen its current one; mbarrier parities tell phases apart only one step, so this needs
            // a ring of at least two tiles (p.dual is 0 otherwise and warp 2 idles).
            // The whole warp walks the loops (all values warp-uniform); one elected lane issues.
            const uint32_t me = static_cast<uint32_t>(warp - 1);
            // descriptor halves that never change: A (two LBOs: S2D stage / plain stage), B
            constexpr uint32_t a_hi = ((kHW * 16u) >> 4) | (1u << 14);           // SBO = one halo row
            constexpr uint32_t a_lbo_s2d = ((4u * kPlane) >> 4) << 16;          // plane pair of a phase
            constexpr uint32_t a_lbo_plain = (static_cast<uint32_t>(kPlane) >> 4) << 16;
            constexpr uint64_t b_hi = static_cast<uint64_t>((128u >> 4) | (1u << 14)) << 32;
            const uint32_t w_lo = w_s >> 4;
            // CG = 2: every B block is half as wide in this CTA, so offsets, slab sizes and LBO halve
            auto mma = [](uint32_t d, uint64_t a, uint64_t b, int n, uint32_t acc) {
                if (CG == 2) umma_bf16_pair(d, a, b, make_idesc_bf16_pair(n), acc);
                else umma_bf16(d, a, b, make_idesc_bf16(n), acc);
            };
            auto commit = [](uint32_t bar) {
                if (CG == 2) umma_commit_pair(bar);
                else umma_commit(bar);
            };
            mbar_wait(w_full, 0);
            if (CG == 2) mbar_wait_cluster(w_peer, 0);
            // ring positions without integer division (ptx.cuh: Ring); a tile of the other issuer
            // advances them too
            Ring rs, rb;
            const bool relaxed = STEM && c_wait_cfg[3] != 0;
            uint32_t li = 0;
            for (int unit = unit0; unit < num_units; unit += unit_step, ++li, rb.next(kBufs)) {
                if (p.dual ? (li & 1u) != me : me != 0u) {   // the other issuer's tile
                    rs.skip(p.nslots, p.n_stages);
                    continue;
                }
                const uint32_t buf = rb.slot;
                const uint32_t aph = rb.phase;
                if (CG == 2) mbar_wait_cluster(acc_empty + 8u * buf, aph ^ 1u);
                else mbar_wait(acc_empty + 8u * buf, aph ^ 1u);
                trace(3 + static_cast<int>(me), li, 0);
                tc_fence_after();
                const uint32_t d0 = tmem_base + buf * 128u;
                for (int s = 0; s < p.n_s2d; ++s, rs.next(p.nslots)) {
                    const uint32_t slot = rs.slot;
                    // fused stem: the CUDA-core stem warps are the critical path, not the issuers
                    if (relaxed) mbar_wait_relaxed(a_full + 8u * slot, rs.phase);
                    else mbar_wait(a_full + 8u * slot, rs.phase);
                    trace(3 + static_cast<int>(me), li, 1 + 2 * s);
                    tc_fence_after();
                    const uint64_t ad = (static_cast<uint64_t>(a_hi) << 32) |
                                        (((a_ring + slot * kSlot) >> 4) | a_lbo_s2d);
                    const uint64_t bd = b_hi | (w_lo + static_cast<uint32_t>(s) * (kS2dSlabUnits / CG));
                    const uint32_t first = s != 0 ? 1u : 0u;
                    const bool last = (s == p.n_s2d - 1) && !p.has_below;
                    if (elect_one()) {
#pragma unroll
                        for (int c = 0; c < 16; ++c) {
                            if (p.dbg & 1) break;
                            const OpShape sh = s2d_shape(c);
                            // LBO of B = 16 * N bytes -> (N) in the descriptor's bits 16..29
                            mma(d0 + sh.dcol, ad + sh.a_off,
                                bd + (s2d_boff(c) / CG + (static_cast<uint32_t>(sh.n / CG) << 16)), sh.n,
                                c ? 1u : first);
                        }
                        commit(a_empty + 8u * slot);
                        if (last) commit(acc_full + 8u * buf);
                    }
                    __syncwarp();
                    trace(3 + static_cast<int>(me), li, 2 + 2 * s);
                }
                if (p.has_below) {
                    const uint32_t slot = rs.slot;
                    mbar_wait(a_full + 8u * slot, rs.phase);
                    tc_fence_after();
                    const uint64_t ad = (static_cast<uint64_t>(a_hi) << 32) |
                                        (((a_ring + slot * kSlot) >> 4) | a_lbo_plain);
                    const uint64_t bd =
                        b_hi | (w_lo + static_cast<uint32_t>(p.n_s2d) * (kS2dSlabUnits / CG));
                    if (elect_one()) {
#pragma unroll
                        for (int k16 = 0; k16 < 4; ++k16) {
#pragma unroll
                            for (int i = 0; i < 9; ++i) {
                                if (p.dbg & 1) break;
                                const OpShape sh = below_shape(i);
                                mma(d0 + sh.dcol, ad + (2 * k16 * (kPlane / 16) + sh.a_off),
                                    bd + ((k16 * kBelowSlabUnits + below_boff(i)) / CG +
                                          (static_cast<uint32_t>(sh.n / CG) << 16)),
                                    sh.n, 1u);
                            }
                        }
                        commit(a_empty + 8u * slot);
                        commit(acc_full + 8u * buf);
                    }
                    __syncwarp();
                    rs.next(p.nslots);
                }
            }
        }
    } else {
        if (STEM == 3) setmaxnreg_inc<kRegsEpi3>();
        // =========================================================== epilogue
        const int et = (threadIdx.x - 128) & 127;   // TMEM lane = position inside the tile
        const int grp = (threadIdx.x - 128) >> 7;   // takes tiles with (li & 1) == grp
        const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const int py = et >> 3, px = et & 7;
        const int H = 2 * p.H2, W = 2 * p.W2;
        const size_t plane = static_cast<size_t>(p.H2) * p.W2 * 8;  // one phase of one 8-ch group
        uint32_t li = 0;
        Ring rb;
        for (int unit = unit0; unit < num_units; unit += unit_step, ++li, rb.next(kBufs)) {
            if (static_cast<int>(li & 1u) != grp) continue;
            const uint32_t buf = rb.slot;
            const uint32_t aph = rb.phase;
            const int tile = tile_of(unit);
            const bool in_range = tile < p.num_tiles;   // false only for the tail of the last pair
            const Tile t = decode_tile(p, in_range ? tile : p.num_tiles - 1);
            const int Y = t.y0 + py, X = t.x0 + px;
            const bool valid = in_range && Y < p.H2 && X < p.W2 && !(p.dbg & 2);
            mbar_wait_relaxed(acc_full + 8u * buf, aph);
            if ((warp & 3) == 0) trace(5 + grp, li, 0);
            tc_fence_after();
            const uint32_t tcol = tmem_base + lane_sel + buf * 128u;
            uint32_t mx[16];
            float zprev = 0.f;
            bool onprev = false;
            int cnt = 0;
#pragma unroll 1
            for (int ph = 0; ph < 4; ++ph) {
                if (p.dbg & 4) break;
                uint32_t r[32];
                tmem_ld32(tcol + ph * 32, r);
                tmem_ld_wait();
                if (ph == 3) {
                    // the last phase is in registers: the accumulator goes back to the issuers
                    // before this phase's arithmetic and stores, not after them
                    tc_fence_before();
                    if (CG == 2) mbar_arrive_cluster(map_to_cta(acc_empty + 8u * buf, 0));
                    else mbar_arrive(acc_empty + 8u * buf);
                }
                const int y = 2 * Y + (ph >> 1), x = 2 * X + (ph & 1);
                const int ry = y == 0 ? 0 : (y == H - 1 ? 2 : 1);
                const int rx = x == 0 ? 0 : (x == W - 1 ? 2 : 1);
                float v[32];
                // bias (+ ReLU for the head, whose dot product runs on the fp32 values; the other
                // epilogues fold the ReLU into the bf16 conversion). Interior pixels (every lane of
                // nearly every warp) take the bias straight from the constant bank; only warps
                // touching the image border of a layer with a composed transposed conv read
                // their per-class bias from smem.
                const bool plain = !p.border_bias || (ry == 1 && rx == 1);
                if (__all_sync(0xffffffffu, plain)) {
#pragma unroll
                    for (int k = 0; k < 32; ++k) {
                        v[k] = __uint_as_float(r[k]) + p.bias[k];
                        if (EPI == EPI_HEAD) v[k] = fmaxf(v[k], 0.f);
                    }
                } else {
                    const float* bp = btab_sp + (ry * 3 + rx) * 32;
#pragma unroll
                    for (int k = 0; k < 32; ++k) {
                        v[k] = __uint_as_float(r[k]) + bp[k];
                        if (EPI == EPI_HEAD) v[k] = fmaxf(v[k], 0.f);
                    }
                }
                if (EPI == EPI_HEAD) {
                    float z = 0.f;
#pragma unroll
                    for (int k = 0; k < 32; ++k) z = fmaf(v[k], p.head_w[k], z);
                    z += p.head_b;
                    const bool on = valid && (z > p.logit_thr);
                    const uint32_t bal = __ballot_sync(0xffffffffu, on);
                    cnt += __popc(bal);
                    if (ph & 1) {
                        // pixels (y, 2X) and (y, 2X+1) leave together
                        const size_t pix = (static_cast<size_t>(t.n) * H + y) * W + 2 * X;
                        if (valid && p.logits)
                            *reinterpret_cast<float2*>(p.logits + pix) = make_float2(zprev, z);
                        if (valid && p.mask)
                            *reinterpret_cast<uint16_t*>(p.mask + pix) =
                                static_cast<uint16_t>((onprev ? 0x00ffu : 0u) | (on ? 0xff00u : 0u));
                    }
                    zprev = z;
                    onprev = on;
                } else {
                    __nv_bfloat16* optr =
                        p.out + ((static_cast<size_t>(t.n) * 4) * 4 + ph) * plane +
                        (static_cast<size_t>(Y) * p.W2 + X) * 8;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint4 q4;
                        q4.x = pack_relu_bf16x2(v[g * 8 + 0], v[g * 8 + 1]);
                        q4.y = pack_relu_bf16x2(v[g * 8 + 2], v[g * 8 + 3]);
                        q4.z = pack_relu_bf16x2(v[g * 8 + 4], v[g * 8 + 5]);
                        q4.w = pack_relu_bf16x2(v[g * 8 + 6], v[g * 8 + 7]);
                        if (valid) *reinterpret_cast<uint4*>(optr + g * 4 * plane) = q4;
                        if (EPI == EPI_RELU_POOL) {
                            if (ph == 0) {
                                mx[g * 4 + 0] = q4.x;
                                mx[g * 4 + 1] = q4.y;
                                mx[g * 4 + 2] = q4.z;
                                mx[g * 4 + 3] = q4.w;
                            } else {
                                mx[g * 4 + 0] = max_bf16x2(mx[g * 4 + 0], q4.x);
                                mx[g * 4 + 1] = max_bf16x2(mx[g * 4 + 1], q4.y);
                                mx[g * 4 + 2] = max_bf16x2(mx[g * 4 + 2], q4.z);
                                mx[g * 4 + 3] = max_bf16x2(mx[g * 4 + 3], q4.w);
                            }
                        }
                    }
                }
            }
            if (EPI == EPI_RELU_POOL && valid && !(p.dbg & 4)) {
                // pooled tensor: plain C8-planar at half resolution
                __nv_bfloat16* pp = p.out_pool + (static_cast<size_t>(t.n) * 4) * plane +
                                    (static_cast<size_t>(Y) * p.W2 + X) * 8;
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    *reinterpret_cast<uint4*>(pp + g * plane) =
                        make_uint4(mx[g * 4 + 0], mx[g * 4 + 1], mx[g * 4 + 2], mx[g * 4 + 3]);
            }
            if (EPI == EPI_HEAD) {
                if (lane == 0 && p.area && cnt) atomicAdd(p.area + t.n, cnt);
            }
            if (p.dbg & 4) {   // experiment without epilogue work: nothing was loaded
                tc_fence_before();
                if (CG == 2) mbar_arrive_cluster(map_to_cta(acc_empty + 8u * buf, 0));
                else mbar_arrive(acc_empty + 8u * buf);
            }
            if ((warp & 3) == 0) trace(5 + grp, li, 1);
        }
    }

    // ------------------------------------------------------------- teardown
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();  // neither CTA leaves while the pair may still touch it
    if (warp == 2) {
        tc_fence_after();
        if (CG == 2) tmem_dealloc_pair(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
}

#ifndef OGL_F16
inline uint16_t operand_bits(double v, bool f16) {
    uint16_t b;
    if (f16) {
        float f = static_cast<float>(v);
        f = f > 65504.f ? 65504.f : (f < -65504.f ? -65504.f : f);
        const __half h = __float2half_rn(f);
        memcpy(&b, &h, 2);
    } else {
        const __nv_bfloat16 h = __float2bfloat16_rn(static_cast<float>(v));
        memcpy(&b, &h, 2);
    }
    return b;
}
#endif

}  // namespace

#ifndef OGL_F16

// B operands of the tensor-core stem: [B_hi, B_lo][K half][n = 32 channels][8 K elements] bf16.
// K rows 0..8 = the taps' weights / 255 (the fp32 value of the CUDA-core stem, split hi + lo),
// row 9 = bias (hi), row 10 = bias (lo) -- the im2col operand holds ones there -- rows 11..15 zero.
// `k_order3`: the K order of the STEM = 3 operand (stem3_k_of_tap; ones at K = 7 and 9).
int build_stem_tc_blob(const StemWeights& sw, std::vector<uint8_t>* out, bool k_order3, bool f16) {
    out->assign(kStemBBytes, 0);
    uint16_t* B = reinterpret_cast<uint16_t*>(out->data());
    auto split = [f16](float v, uint16_t* hi, uint16_t* lo) {
        if (f16) {   // 11 + 11 significant bits; the low part may be subnormal (spacing 2^-24)
            const __half h = __float2half_rn(v);
            const __half l = __float2half_rn(v - __half2float(h));
            memcpy(hi, &h, 2);
            memcpy(lo, &l, 2);
            return;
        }
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const float rest = v - __bfloat162float(h);
        const __nv_bfloat16 l = __float2bfloat16_rn(rest);
        memcpy(hi, &h, 2);
        memcpy(lo, &l, 2);
    };
    auto at = [&](int part, int k, int n) -> uint16_t& {
        return B[part * 512 + ((k >> 3) * 32 + n) * 8 + (k & 7)];
    };
    for (int n = 0; n < 32; ++n) {
        for (int k = 0; k < 9; ++k) {
            const float w = static_cast<float>(static_cast<double>(sw.w[n * 9 + k]) / 255.0);
            const int kk = k_order3 ? stem3_k_of_tap(k) : k;
            split(w, &at(0, kk, n), &at(1, kk, n));
        }
        split(sw.b[n], &at(0, k_order3 ? 7 : 9, n), &at(0, k_order3 ? 9 : 10, n));
    }
    return 0;
}

int build_s2d_host(const float* w3, const float* b3, int cin_s, const float* wt, const float* bt,
                   int cin_b, S2dHost* out, bool f16) {
    if (cin_s <= 0 || cin_s % 16) return fail("s2d layer: Cin of the S2D source must be a multiple of 16");
    if (wt && cin_b != 64) return fail("s2d layer: the composed ConvTranspose2d needs 64 input channels");
    const int cin3 = cin_s + (wt ? 32 : 0);
    auto W3 = [&](int co, int ci, int dy, int dx) -> double {
        return w3[((static_cast<size_t>(co) * cin3 + ci) * 3 + dy) * 3 + dx];
    };
    out->wblob.clear();
    out->ops.clear();
    int n_stages = 0;
    std::vector<uint8_t> pair_half[2];

    // One MMA: its shape comes from the table the kernel unrolls (s2d_shape / below_shape); the
    // B block [2][N][8] holds, for column n = (phase - first phase) * 32 + cout, the weight
    // through which this A reaches that output phase, or zero when it does not.
    auto emit = [&](const OpShape& sh, uint32_t a_extra, int src, int ys, int xs, size_t want_off,
                    auto&& weight /* (k, py, px, cout) -> double */) -> int {
        const int N = sh.n, pmin = sh.dcol / 32;
        const size_t b_off = out->wblob.size();
        if (b_off != want_off) return fail("s2d layer: weight block offset differs from the kernel's table");
        out->wblob.resize(b_off + static_cast<size_t>(N) * 32);
        uint16_t* B = reinterpret_cast<uint16_t*>(out->wblob.data() + b_off);
        for (int kh = 0; kh < 2; ++kh)
            for (int n = 0; n < N; ++n)
                for (int e = 0; e < 8; ++e) {
                    const int pp = pmin + n / 32, c = n % 32;
                    double v = 0.0;
                    if (((ys >> (pp >> 1)) & 1) && ((xs >> (pp & 1)) & 1))
                        v = weight(kh * 8 + e, pp >> 1, pp & 1, c);
                    B[(static_cast<size_t>(kh) * N + n) * 8 + e] = operand_bits(v, f16);
                }
        // CTA-pair form: rank r keeps columns [r N/2, (r+1) N/2) of this op as [2][N/2][8], at half
        // the offset inside its own half of the blob
        {
            const int nh = N / 2;
            for (int r = 0; r < 2; ++r) {
                std::vector<uint8_t>& dstv = pair_half[r];
                const size_t o = dstv.size();
                dstv.resize(o + static_cast<size_t>(nh) * 32);
                uint16_t* P = reinterpret_cast<uint16_t*>(dstv.data() + o);
                for (int kh = 0; kh < 2; ++kh)
                    for (int n = 0; n < nh; ++n)
                        for (int e = 0; e < 8; ++e)
                            P[(static_cast<size_t>(kh) * nh + n) * 8 + e] =
                                B[(static_cast<size_t>(kh) * N + r * nh + n) * 8 + e];
            }
        }
        S2dOp op;
        op.w0 = (static_cast<uint32_t>(sh.a_off) + a_extra) | (static_cast<uint32_t>(sh.dcol) << 16) |
                (static_cast<uint32_t>(src) << 24) | ((out->ops.empty() ? 0u : 1u) << 25);
        op.b_lo = static_cast<uint32_t>(b_off >> 4) | (static_cast<uint32_t>(N) << 16);  // LBO = 16 N
        op.idesc = make_idesc_fmt(N, f16 ? 0u : 1u);
        op.pad = 0;
        out->ops.push_back(op);
        return 0;
    };

    // ---- S2D source: one stage per 16-channel slab, 16 ops each
    for (int s = 0; s < cin_s / 16; ++s) {
        if (n_stages >= kS2dMaxStages - 1) return fail("s2d layer: too many stages");
        for (int c = 0; c < 16; ++c) {
            const int cy = c >> 2, cx = c & 3;
            const int oy = combo_o(cy), qy = combo_q(cy), ox = combo_o(cx), qx = combo_q(cx);
            const size_t want = (static_cast<size_t>(s) * kS2dSlabUnits + s2d_boff(c)) * 16;
            if (emit(s2d_shape(c), 0, 0, set_s2d(cy), set_s2d(cx), want,
                     [&](int k, int py, int px, int co) {
                         return W3(co, s * 16 + k, tap_of(oy, qy, py), tap_of(ox, qx, px));
                     }))
                return 1;
        }
        out->stage_src[n_stages] = 0;
        out->stage_plane0[n_stages] = 2 * s;
        out->stage_op_end[n_stages] = static_cast<int>(out->ops.size());
        ++n_stages;
    }
    // ---- composed ConvTranspose2d: one stage = all 64 channels of the tensor below,
    // 4 slabs x 9 half-resolution offsets. `up` phase q at offset o is wt[., ., q] applied to
    // below[Y + o], so the weight from below channel k to (output phase p, cout) is
    // sum over the q that p reaches at this offset, and over the 32 `up` channels.
    if (wt) {
        auto WT = [&](int ci, int co, int qy, int qx) -> double {
            return wt[((static_cast<size_t>(ci) * 32 + co) * 2 + qy) * 2 + qx];
        };
        const size_t base = static_cast<size_t>(cin_s / 16) * kS2dSlabUnits;
        for (int k16 = 0; k16 < cin_b / 16; ++k16)
            for (int i = 0; i < 9; ++i) {
                const int oy = below_o(i / 3), ox = below_o(i % 3);
                const size_t want = (base + static_cast<size_t>(k16) * kBelowSlabUnits + below_boff(i)) * 16;
                if (emit(below_shape(i), static_cast<uint32_t>(2 * k16 * (kPlane / 16)), 1,
                         set_below(i / 3), set_below(i % 3), want,
                         [&](int k, int py, int px, int co) {
                             double acc = 0.0;
                             for (int qy = 0; qy < 2; ++qy) {
                                 const int dy = tap_of(oy, qy, py);
                                 if (dy < 0) continue;
                                 for (int qx = 0; qx < 2; ++qx) {
                                     const int dx = tap_of(ox, qx, px);
                                     if (dx < 0) continue;
                                     for (int cm = 0; cm < 32; ++cm)
                                         acc += WT(k16 * 16 + k, cm, qy, qx) * W3(co, cin_s + cm, dy, dx);
                                 }
                             }
                             return acc;
                         }))
                    return 1;
            }
        out->stage_src[n_stages] = 1;
        out->stage_plane0[n_stages] = 0;
        out->stage_op_end[n_stages] = static_cast<int>(out->ops.size());
        ++n_stages;
    }
    out->n_stages = n_stages;
    out->wblob_pair = pair_half[0];
    out->wblob_pair.insert(out->wblob_pair.end(), pair_half[1].begin(), pair_half[1].end());
    // ---- bias per (row class, column class): the transposed conv's bias reaches an output
    // pixel only through the taps that fall inside the image (zero padding of `up`)
    out->btab.assign(9 * 32, 0.f);
    for (int ry = 0; ry < 3; ++ry)
        for (int rx = 0; rx < 3; ++rx)
            for (int c = 0; c < 32; ++c) {
                double v = b3[c];
                if (wt)
                    for (int dy = 0; dy < 3; ++dy) {
                        if ((ry == 0 && dy == 0) || (ry == 2 && dy == 2)) continue;
                        for (int dx = 0; dx < 3; ++dx) {
                            if ((rx == 0 && dx == 0) || (rx == 2 && dx == 2)) continue;
                            for (int cm = 0; cm < 32; ++cm) v += W3(c, cin_s + cm, dy, dx) * bt[cm];
                        }
                    }
                out->btab[(ry * 3 + rx) * 32 + c] = static_cast<float>(v);
            }
    return 0;
}
#endif  // !OGL_F16

namespace {
template <int EPI>
int set_attr() {
    OGL_CUDA(cudaFuncSetAttribute(s2d_tc_kernel<EPI, 1, 0>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    OGL_CUDA(cudaFuncSetAttribute(s2d_tc_kernel<EPI, 2, 0>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    return 0;
}
template <int EPI>
int launch_kernel(bool pair, int grid, size_t smem, cudaStream_t stream, const CUtensorMap& tmS,
                  const CUtensorMap& tmB, const S2dParams& p) {
    if (!pair) {
        s2d_tc_kernel<EPI, 1, 0><<<grid, kThreads, smem, stream>>>(tmS, tmB, p);
        OGL_CUDA(cudaGetLastError());
        return 0;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    OGL_CUDA(cudaLaunchKernelEx(&cfg, s2d_tc_kernel<EPI, 2, 0>, tmS, tmB, p));
    return 0;
}
}  // namespace

namespace {
// Debug timeline (OGL_TRACE=file, OGL_TRACE_LAUNCH=stem|head|up|down, default stem): CTA 0 of the
// chosen space-to-depth launch writes clock64 at its hand-offs; the launch then synchronises and the
// buffer is dumped. scripts/stem_trace.py reads it.
unsigned long long* g_trace_dev = nullptr;
int trace_begin(const char* which, cudaStream_t stream, S2dParams* p) {
    static const char* file = getenv("OGL_TRACE");
    static const char* want = getenv("OGL_TRACE_LAUNCH") ? getenv("OGL_TRACE_LAUNCH") : "stem";
    p->trace = nullptr;
    if (!file || strcmp(which, want) != 0) return 0;
    const size_t n = static_cast<size_t>(kTraceRoles) * kTraceTiles * 8;
    if (!g_trace_dev) OGL_CUDA(cudaMalloc(&g_trace_dev, n * 8));
    OGL_CUDA(cudaMemsetAsync(g_trace_dev, 0, n * 8, stream));
    p->trace = g_trace_dev;
    return 0;
}
int trace_end(cudaStream_t stream, const S2dParams& p) {
    if (!p.trace) return 0;
    const size_t n = static_cast<size_t>(kTraceRoles) * kTraceTiles * 8;
    std::vector<unsigned long long> host(n);
    OGL_CUDA(cudaStreamSynchronize(stream));
    OGL_CUDA(cudaMemcpy(host.data(), g_trace_dev, n * 8, cudaMemcpyDeviceToHost));
    if (FILE* f = fopen(getenv("OGL_TRACE"), "wb")) {
        fwrite(host.data(), 8, n, f);
        fclose(f);
    }
    return 0;
}
}  // namespace

int s2d_tc_init() {
    OGL_CUDA(set_wait_cfg());
    OGL_CUDA(cudaFuncSetAttribute(s2d_tc_kernel<EPI_RELU_POOL, 1, 1>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    OGL_CUDA(cudaFuncSetAttribute(s2d_tc_kernel<EPI_RELU_POOL, 1, 2>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    OGL_CUDA(cudaFuncSetAttribute(s2d_tc_kernel<EPI_RELU_POOL, 1, 3>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    return set_attr<EPI_RELU>() || set_attr<EPI_RELU_POOL>() || set_attr<EPI_HEAD>();
}

int launch_s2d_tc(const S2dLayer& L, const __nv_bfloat16* src_s2d, const __nv_bfloat16* below,
                  int B, int H, int W, __nv_bfloat16* out_s2d, __nv_bfloat16* out_pool,
                  const HeadParams* head, int num_sms, cudaStream_t stream, int cta_group,
                  const uint8_t* stem_frames, const StemWeights* stem, bool reverse,
                  const uint8_t* stem_tc_blob, int stem_tc_warps) {
    if (H < 2 || W < 2 || H % 2 || W % 2) return fail("s2d layer needs even, non-empty H and W");
    if ((W / 2) % 8) return fail("s2d layer needs W to be a multiple of 16");
    const bool fused_stem = stem_frames != nullptr;
    if (fused_stem && (!stem || L.cin_s != 32 || L.cin_b != 0 || L.epi != EPI_RELU_POOL))
        return fail("the fused stem feeds downs.0.net.3 only");
    if ((!src_s2d && !fused_stem) || L.n_stages < 1 || !L.wblob || !L.btab)
        return fail("s2d layer is not built");
    if (L.cin_b > 0 && !below) return fail("s2d layer: the tensor below is missing");
    if (L.epi == EPI_RELU_POOL && !out_pool) return fail("pool epilogue needs out_pool");
    if (L.epi == EPI_HEAD && !head) return fail("head epilogue needs head parameters");
    if (L.epi != EPI_HEAD && !out_s2d) return fail("s2d layer: no output tensor");

    S2dParams p;
    memset(&p, 0, sizeof p);
    p.wbytes = L.wbytes;
    p.n_stages = L.n_stages;
    p.has_below = L.stage_src[L.n_stages - 1] == 1 ? 1 : 0;
    p.n_s2d = L.n_stages - p.has_below;
    for (int i = 0; i < kS2dMaxStages; ++i) {
        p.stage_src[i] = L.stage_src[i];
        p.stage_plane0[i] = L.stage_plane0[i];
    }
    if (p.n_s2d != L.cin_s / 16 || (p.has_below != 0) != (L.cin_b > 0) ||
        L.wbytes != (static_cast<uint32_t>(p.n_s2d) * kS2dSlabUnits +
                     (p.has_below ? 4u * kBelowSlabUnits : 0u)) * 16u)
        return fail("s2d layer: program does not match the kernel's compile-time tables");
    if (fused_stem) {
        // weights pre-scaled by 1/255 (in fp64, rounded once) so that the u8 values are used as
        // they are: utils.py:235 folded into downs.0.net.0
        p.frames = stem_frames;
        p.stem = make_stem_pairs(*stem);
        p.stem_b = stem_tc_blob;   // non-null: the stem runs on the tensor cores (STEM >= 2)
        // 8 warps: bf16 im2col x bf16 weights; 16 warps: f16 im2col x f16 weights (a mixed
        // f16 x bf16 descriptor, OGL_STEM3_BFMT=1 with a bf16 blob, traps: illegal instruction)
        static const int bfmt_env = env_knob("OGL_STEM3_BFMT", 0, 0, 1);
        static const int lo_env = env_knob("OGL_STEM_LO", 1, 0, 1);
        p.stem_lo = lo_env;
        p.stem_idesc = stem_tc_warps == 16 ? make_idesc_ab(32, 0u, bfmt_env ? 1u : 0u)
                                           : make_idesc_fmt(32, 1u);
    }
    p.btab = L.btab;
    p.border_bias = L.cin_b > 0 ? 1 : 0;
    for (int i = 0; i < 32; ++i) p.bias[i] = L.bias_host[i];
    if (head) {
        if (L.epi == EPI_HEAD && !head->w_host) return fail("head epilogue needs the host weights");
        for (int i = 0; i < 32 && head->w_host; ++i) p.head_w[i] = head->w_host[i];
        p.head_b = head->b;
        p.logit_thr = head->logit_thr;
        p.logits = head->logits;
        p.mask = head->mask;
        p.area = head->area;
    }
    p.out = out_s2d;
    p.out_pool = out_pool;
    p.H2 = H / 2;
    p.W2 = W / 2;
    p.B = B;
    p.tiles_x = (p.W2 + kTW - 1) / kTW;
    p.tiles_y = (p.H2 + kTH - 1) / kTH;
    const long long total = static_cast<long long>(B) * p.tiles_x * p.tiles_y;
    if (total * (p.tiles_x * p.tiles_y) >= (1ll << 40) || total > 0x7fffffffll)
        return fail("batch too large for the tile decoder");
    p.num_tiles = static_cast<int>(total);
    p.magic_tx = ((1ull << 40) / static_cast<unsigned long long>(p.tiles_x)) + 1;
    p.magic_tpf = ((1ull << 40) / static_cast<unsigned long long>(p.tiles_x * p.tiles_y)) + 1;
    static const int dbg_env = experiment_dbg();
    p.dbg = dbg_env;
    p.reverse = reverse ? 1 : 0;

    // CTA pairs. cta_group 2: for the layer with the composed transposed conv, whose 152 KB of
    // weights leave one CTA only 3 activation stages (measured 1.57 -> 1.12 ms; the HBM-bound
    // downs.0.net.3 is slower paired, 0.85 -> 1.05 ms, and ups.7.net.3 unchanged), when there
    // is a tile per SM. cta_group 3 (unit tests): whenever there are two tiles.
    static const int pair_all_env = env_knob("OGL_S2D_PAIR_ALL", 0, 0, 1);
    const bool pair = !fused_stem && cta_group >= 2 && L.wblob2 && num_sms >= 2 &&
                      (cta_group == 3 ? p.num_tiles >= 2
                                      : (p.num_tiles >= num_sms && (L.cin_b > 0 || pair_all_env)));
    p.wblob = pair ? L.wblob2 : L.wblob;
    const size_t wres = L.wbytes / (pair ? 2 : 1);
    const bool tc_stem = fused_stem && stem_tc_blob != nullptr;
    const size_t fixed = 128 + ((wres + 127u) & ~static_cast<size_t>(127)) + 9 * 32 * 4 + 8 +
                         16 * kAccBufs + 16 + 16 + 16 * kU8Slots + 32 + 16 * kSdMax + 128 +
                         kU8Slots * kU8Slot + (tc_stem ? 2 * kStemABytes + kStemBBytes : 0) + 64;
    // Ring depth and issuers per layer (same-call A/Bs, profiles/exp_r02_s2d_rings.jsonl): a
    // two-stage tile runs best with a ring of two tiles (the head: 0.70 instead of 0.735 ms with
    // three) and two issuer warps; the three-stage tile of the composed layer (CTA pair, HBM-bound)
    // with ONE issuer over its six stages (1.16 instead of 1.25 ms).
    int nslots = L.cin_b > 0 ? 6 : 4;
    static const int ns_env = env_knob("OGL_S2D_SLOTS", 0, 2, kSdMax);   // 0: per layer
    static const int ns_head_env = env_knob("OGL_S2D_SLOTS_HEAD", 0, 2, kSdMax);
    if (ns_env > 0) nslots = ns_env;
    if (ns_head_env > 0 && L.cin_b == 0 && !fused_stem) nslots = ns_head_env;
    while (nslots > 2 && fixed + static_cast<size_t>(nslots) * (kSlot + 16) > kMaxSmem) --nslots;
    const size_t smem = fixed + static_cast<size_t>(nslots) * (kSlot + 16);
    if (smem > static_cast<size_t>(kMaxSmem)) return fail("s2d layer: shared memory budget exceeded");
    p.nslots = nslots;
    static const int dual_env = env_knob("OGL_DUAL", 1, 0, 1);
    static const int dual_below_env = env_knob("OGL_DUAL_BELOW", 0, 0, 1);
    p.dual = (dual_env && nslots >= 2 * p.n_stages && (L.cin_b == 0 || dual_below_env)) ? 1 : 0;

    CUtensorMap tmS, tmB;
    if (fused_stem) {
        // u8 frames [B][H][W] as a 3-D tensor; box = the 38 rows x 48 bytes around a tile
        const uint64_t dims[3] = {static_cast<uint64_t>(W), static_cast<uint64_t>(H),
                                  static_cast<uint64_t>(B)};
        const uint64_t str[2] = {static_cast<uint64_t>(W), static_cast<uint64_t>(W) * H};
        const uint32_t box[3] = {kU8Row, kStagePxH, 1};
        if (encode_map(&tmS, stem_frames, 3, dims, str, box, true)) return 1;
        memset(&tmB, 0, sizeof tmB);
        const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
        if (trace_begin(tc_stem && stem_tc_warps == 16 ? "stem" : "", stream, &p)) return 1;
        if (tc_stem) {
            if (nslots < 4) return fail("s2d layer: the tensor-core stem needs 4 activation stages");
            if (stem_tc_warps == 16)
                s2d_tc_kernel<EPI_RELU_POOL, 1, 3>
                    <<<grid, kThreads + kStem3Threads, smem, stream>>>(tmS, tmB, p);
            else
                s2d_tc_kernel<EPI_RELU_POOL, 1, 2>
                    <<<grid, kThreads + kStemThreads, smem, stream>>>(tmS, tmB, p);
        } else {
            s2d_tc_kernel<EPI_RELU_POOL, 1, 1>
                <<<grid, kThreads + kStemThreads, smem, stream>>>(tmS, tmB, p);
        }
        OGL_CUDA(cudaGetLastError());
        if (trace_end(stream, p)) return 1;
        return 0;
    }
    {
        const uint64_t W2 = p.W2, H2 = p.H2;
        const uint64_t dims[5] = {W2 * 8, H2, 4, static_cast<uint64_t>(L.cin_s / 8),
                                  static_cast<uint64_t>(B)};
        const uint64_t str[4] = {W2 * 16, W2 * 16 * H2, W2 * 16 * H2 * 4,
                                 W2 * 16 * H2 * 4 * (L.cin_s / 8)};
        const uint32_t box[5] = {kHW * 8, kHH, 4, 2, 1};
        if (encode_bf16_map(&tmS, src_s2d, 5, dims, str, box)) return 1;
    }
    if (L.cin_b > 0) {
        const uint64_t W2 = p.W2, H2 = p.H2;
        const uint64_t dims[4] = {W2 * 8, H2, static_cast<uint64_t>(L.cin_b / 8),
                                  static_cast<uint64_t>(B)};
        const uint64_t str[3] = {W2 * 16, W2 * 16 * H2, W2 * 16 * H2 * (L.cin_b / 8)};
        const uint32_t box[4] = {kHW * 8, kHH, 8, 1};
        if (encode_bf16_map(&tmB, below, 4, dims, str, box)) return 1;
    } else {
        tmB = tmS;
    }
    const int grid = pair ? (num_sms & ~1) : (p.num_tiles < num_sms ? p.num_tiles : num_sms);
    if (trace_begin(L.epi == EPI_HEAD ? "head" : (L.cin_b > 0 ? "up" : "down"), stream, &p)) return 1;
    int rc;
    if (L.epi == EPI_RELU) rc = launch_kernel<EPI_RELU>(pair, grid, smem, stream, tmS, tmB, p);
    else if (L.epi == EPI_RELU_POOL)
        rc = launch_kernel<EPI_RELU_POOL>(pair, grid, smem, stream, tmS, tmB, p);
    else if (L.epi == EPI_HEAD) rc = launch_kernel<EPI_HEAD>(pair, grid, smem, stream, tmS, tmB, p);
    else return fail("s2d layer: unknown epilogue");
    return rc ? rc : trace_end(stream, p);
}

}  // namespace ogl
