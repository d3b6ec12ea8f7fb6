// Decoder levels 1-3: ConvTranspose2d(2f -> f, k2, s2) + torch.cat([skip, up]) + conv3x3(2f -> f) + BN +
// ReLU in ONE tcgen05 launch; `up` is never written to or read from HBM.
//
// Replaces /root/reference/openglottal/models/unet.py:82 (x = self.ups[idx](x)), :86 (torch.cat) and
// the first Conv2d+BatchNorm2d+ReLU of the DoubleConv at :87 (DoubleConv :24-26), for
// ups.{0,2,4} / ups.{1,3,5}.net.0 (the full-resolution level does the same inside s2d_tc.cu).
//
// Composition. up[y, x] = bt + Wt[:, :, y & 1, x & 1]^T below[y >> 1, x >> 1], so for an output pixel
// of phase p = (y & 1, x & 1) at half-resolution position (Y, X) the 3x3 window of `up` touches only
// the 2x2 block of `below` positions (Y + py - 1 + {0, 1}, X + px - 1 + {0, 1}):
//     conv3x3(up)[2Y+py, 2X+px] = sum over the 4 offsets o of  Wc[p][o]^T below[Y + oy, X + ox]  + bias terms,
//     Wc[p][o][ci][co] = sum over the taps (dy, dx) that land on offset o, and over c, of
//                        Wt[ci][c][(py+dy) & 1][(px+dx) & 1] * W3[co][f + c][dy][dx]         (fp64, rounded once)
// 8 f^2 MACs per output pixel instead of 2 f^2 (ConvTranspose2d) + 9 f^2 (its half of the conv).
// The transposed conv's bias reaches a pixel only through the taps that fall inside the image (zero
// padding of `up`): 3 x 3 bias classes by row / column position, as in s2d_tc.cu.
//
// GEMM view. A GEMM row is a half-resolution position; a CTA tile is 8 (x) x 16 (y) positions = 128
// rows, with FOUR accumulators of N columns in TMEM, one per output phase (4 N <= 512 columns).
//   * skip tensor: stored space-to-depth [frame][C/8][phase][H/2][W/2][8] by its producer
//     (downs.l.net.3's epilogue), so that one 5-D TMA box (10 px x 8 ch, 18 rows, 4 phases, 4 planes)
//     is the K-major SWIZZLE_NONE operand tile of a 32-channel K block for all four output phases:
//     tap (dy, dx) of output phase p reads source phase ((py+dy) & 1, (px+dx) & 1) at half-resolution
//     offset ((py+dy) >> 1, (px+dx) >> 1) -- a start-address offset.
//   * below tensor: plain C8-planar at half resolution; one 4-D box (10 px x 8 ch, 18 rows, 16 planes)
//     brings 128 channels with their halo; every (phase, offset) pair is again a start-address offset.
//   * weights: streamed per K block, [pass][kb][9 taps][4][N][8] for the skip half and
//     [pass][kb][16 (phase, offset)][4][N][8] for the composed half, in stages of <= 36 KB.
//   * warp roles as in conv_tc.cu. Issuer warp `me` owns the two accumulators of output row phase
//     py = me, epilogue warp group `me` drains them.
//     N = 64: two accumulator sets (2 x 4 x 64 columns), so the epilogue of a tile overlaps the MMAs
//     of the next one; phases (py, 0) and (py, 1) of a position are two horizontally adjacent pixels
//     and leave as one 32-byte store per 8-channel group.
//     N = 128: the four accumulators fill TMEM. Each has its own full / empty barrier and a tile's K
//     loop starts and ends with a group of the composed half issued ACCUMULATOR-MAJOR (px = 0, then
//     px = 1; 32 MMAs = 2048 cycles each): accumulator (py, 0) is drained while the MMAs of (py, 1)
//     finish the tile, and (py, 1) while those of the next tile's (py, 0) start it.
//   * CG = 2: the two CTAs of a cluster take neighbouring tiles and run every MMA as one 256-row
//     tcgen05.mma.cta_group::2, each staging half of the weight columns (as conv_tc.cu).
#include "internal.h"
#include "ptx.cuh"

#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#ifdef OGL_F16   // upcat_tc_f16.cu: the same kernels with f16 operands, exported under other names
#define launch_upcat_tc launch_upcat_tc_f16
#define upcat_tc_init upcat_tc_init_f16
#endif

namespace ogl {

namespace {

constexpr int kTW = 8, kTH = 16;               // tile: 8 x 16 half-resolution positions
constexpr int kHW = kTW + 2, kHH = kTH + 2;    // with halo: 10 x 18
constexpr int kPlaneU = kHW * kHH;             // 16-byte cells of one 8-channel plane (180)
constexpr int kPlaneB = kPlaneU * 16;          // 2880 B
constexpr int kStage = 16 * kPlaneB;           // 46080 B: 4 planes x 4 phases (skip) or 16 planes (below)
constexpr int kThreads = 384;                  // 4 control warps + 8 epilogue warps
constexpr int kMaxSmem = 227 * 1024;

struct UpcatParams {
    const uint8_t* wskip;    // [pass][kbS][9][4][N][8]           (CG = 1; CG = 2 goes through tmWs)
    const uint8_t* wbelow;   // [pass][kbB][16][4][N][8]          (CG = 1; CG = 2 goes through tmWb)
    const float* bias;       // [cout]: interior pixels
    const float* btab;       // [3][3][cout]: per (row class, column class)
    __nv_bfloat16* out;      // [B][cout/8][2 H2][2 W2][8]
    int kbS, nG;             // 32-channel K blocks of the skip, 128-channel groups of the tensor below
    int npass, cout;
    int H2, W2, B;
    int tiles_x, tiles_y, num_tiles;
    unsigned long long magic_tx, magic_tpf;
    int na, nw;
    int pass_fast;
};

struct Tile {
    int n, y0, x0;
};
__device__ __forceinline__ Tile decode_tile(const UpcatParams& p, int tile) {
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

// shared-memory addresses every role needs
struct Smem {
    uint32_t a_ring, w_ring, w_slot;
    uint32_t a_full, a_empty, w_full, w_empty, acc_full, acc_empty;
};

// 1-D structure of conv3x3 on phase-separated data (see the header): source coordinate s = p + d - 1
// of output phase p and tap index d; source phase s & 1, half-resolution offset floor(s / 2) + 1 in
// halo-tile coordinates
__host__ __device__ constexpr int src_phase(int p, int d) { return (p + d - 1) & 1; }
__host__ __device__ constexpr int src_cell(int p, int d) { return (p + d + 1) >> 1; }

// One issuer warp: ME = the output row phase py whose two accumulators it owns.
template <int N, int CG, int ME>
__device__ __forceinline__ void issue_half(const UpcatParams& p, const Smem& s, uint32_t tmem_base,
                                           int item0, int items, int item_step) {
    constexpr int TPS = N == 64 ? 9 : 3;     // skip taps per weight stage
    constexpr bool kDB = N == 64;            // two accumulator sets
    constexpr uint32_t nb = N / CG;          // weight columns staged per CTA
    constexpr uint32_t btap = 4u * nb;       // one tap of B, 16-byte units
    constexpr uint32_t bstep = 2u * nb;      // second K = 16 half of a 32-channel block
    const uint32_t idesc = CG == 2 ? make_idesc_bf16_pair(N) : make_idesc_bf16(N);
    // A: SBO = one halo row (160 B); LBO = one channel plane: 4 phases apart in a skip stage
    const uint64_t adesc_s = make_smem_desc(0, 4u * kPlaneB, kHW * 16u);
    const uint64_t adesc_b = make_smem_desc(0, kPlaneB, kHW * 16u);
    const uint64_t bdesc0 = make_smem_desc(0, 16u * nb, 128u);
    auto mma = [](uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
        if (CG == 2) umma_bf16_pair(d, a, b, id, acc);
        else umma_bf16(d, a, b, id, acc);
    };
    auto commit = [](uint32_t bar) {
        if (CG == 2) umma_commit_pair(bar);
        else umma_commit(bar);
    };
    auto wait_acc_empty = [](uint32_t bar, uint32_t parity) {
        if (CG == 2) mbar_wait_cluster(bar, parity);   // the peer's epilogue arrives remotely
        else mbar_wait(bar, parity);
        tc_fence_after();
    };
    // ring positions advanced without integer division (ptx.cuh: Ring)
    Ring ra, rw;
    uint32_t li = 0;
    for (int item = item0; item < items; item += item_step, ++li) {
        const uint32_t buf = kDB ? (li & 1u) : 0u;
        const uint32_t aph = kDB ? ((li >> 1) & 1u) : (li & 1u);
        const uint32_t d_me = tmem_base + buf * 256u + static_cast<uint32_t>(2 * ME * N);
        // barrier index of accumulator (ME, px): kDB: one per (buffer, half); else one per accumulator
        auto acc_bar = [&](int px) { return 8u * (kDB ? buf * 2u + ME : ME * 2u + px); };

        // ---- composed half, one group of 4 K blocks. N = 64: per K block one stage per px with both
        // py (8 pairs); N = 128: accumulator-major -- px outer, K block, then one stage per py.
        auto below_group = [&](bool first_g, bool last_g) {
            const uint32_t sa = ra.slot;
            mbar_wait(s.a_full + 8u * sa, ra.phase);
            const uint64_t ad = adesc_b + ((s.a_ring + sa * kStage) >> 4);
            auto pair_mmas = [&](uint64_t bd, int px, int g4, int t0, bool first) {
#pragma unroll
                for (int o4 = 0; o4 < 4; ++o4) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const uint32_t aoff =
                            (g4 * 4 + 2 * j) * kPlaneU + (ME + (o4 >> 1)) * kHW + (px + (o4 & 1));
                        mma(d_me + px * N, ad + aoff, bd + ((t0 + o4) * btap + j * bstep), idesc,
                            (first && o4 == 0 && j == 0) ? 0u : 1u);
                    }
                }
            };
            if (kDB) {
#pragma unroll
                for (int g4 = 0; g4 < 4; ++g4) {
#pragma unroll
                    for (int px = 0; px < 2; ++px, rw.next(p.nw)) {
                        const uint32_t sw = rw.slot;
                        mbar_wait(s.w_full + 8u * sw, rw.phase);
                        tc_fence_after();
                        const uint64_t bd = bdesc0 + ((s.w_ring + sw * s.w_slot) >> 4);
                        if (elect_one()) {
                            pair_mmas(bd, px, g4, ME * 4, false);
                            commit(s.w_empty + 8u * sw);
                            if (g4 == 3 && px == 1) {
                                commit(s.a_empty + 8u * sa);
                                if (last_g) commit(s.acc_full + acc_bar(0));
                            }
                        }
                        __syncwarp();
                    }
                }
            } else {
#pragma unroll
                for (int px = 0; px < 2; ++px) {
                    if (first_g) wait_acc_empty(s.acc_empty + acc_bar(px), aph ^ 1u);
#pragma unroll
                    for (int g4 = 0; g4 < 4; ++g4) {
#pragma unroll
                        for (int py = 0; py < 2; ++py, rw.next(p.nw)) {
                            const uint32_t sw = rw.slot;
                            mbar_wait(s.w_full + 8u * sw, rw.phase);
                            tc_fence_after();
                            const uint64_t bd = bdesc0 + ((s.w_ring + sw * s.w_slot) >> 4);
                            if (elect_one()) {
                                if (py == ME) pair_mmas(bd, px, g4, 0, first_g && g4 == 0);
                                commit(s.w_empty + 8u * sw);
                                if (g4 == 3 && py == 1) {
                                    if (px == 1) commit(s.a_empty + 8u * sa);
                                    if (last_g) commit(s.acc_full + acc_bar(px));
                                }
                            }
                            __syncwarp();
                        }
                    }
                }
            }
            ra.next(p.na);
        };

        if (kDB) wait_acc_empty(s.acc_empty + acc_bar(0), aph ^ 1u);
        else below_group(true, false);         // N = 128: the tile starts accumulator-major
        // ---------------- skip half: kbS blocks x 9 taps, the same tap weights for all phases
        for (int kb = 0; kb < p.kbS; ++kb, ra.next(p.na)) {
            const uint32_t sa = ra.slot;
            mbar_wait(s.a_full + 8u * sa, ra.phase);
            const uint64_t ad = adesc_s + ((s.a_ring + sa * kStage) >> 4);
#pragma unroll
            for (int tg = 0; tg < 9 / TPS; ++tg, rw.next(p.nw)) {
                const uint32_t sw = rw.slot;
                mbar_wait(s.w_full + 8u * sw, rw.phase);
                tc_fence_after();
                const uint64_t bd = bdesc0 + ((s.w_ring + sw * s.w_slot) >> 4);
                const uint32_t first = (!kDB || (kb | tg) != 0) ? 1u : 0u;
                if (elect_one()) {
#pragma unroll
                    for (int t = 0; t < TPS; ++t) {
                        const int tap = tg * TPS + t, dy = tap / 3, dx = tap % 3;
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
#pragma unroll
                            for (int px = 0; px < 2; ++px) {
                                const uint32_t aoff =
                                    j * (8u * kPlaneU) +
                                    (src_phase(ME, dy) * 2 + src_phase(px, dx)) * kPlaneU +
                                    src_cell(ME, dy) * kHW + src_cell(px, dx);
                                mma(d_me + px * N, ad + aoff, bd + (t * btap + j * bstep), idesc,
                                    (t | j) ? 1u : first);
                            }
                        }
                    }
                    commit(s.w_empty + 8u * sw);
                    if (tg == 9 / TPS - 1) commit(s.a_empty + 8u * sa);
                }
                __syncwarp();
            }
        }
        // ---------------- the (other) groups of the composed half; the last one ends the tile
        for (int g = kDB ? 0 : 1; g < p.nG; ++g) below_group(false, g == p.nG - 1);
    }
}

template <int N, int CG>
__global__ void __launch_bounds__(kThreads, 1)
upcat_tc_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmWs, const __grid_constant__ CUtensorMap tmWb,
                const UpcatParams p) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int TPS = N == 64 ? 9 : 3;
    constexpr int TAPB = N == 64 ? 8 : 4;    // (phase, offset) pairs of the composed half per stage
    constexpr uint32_t nb = N / CG;
    constexpr uint32_t w_tap_bytes = 64u * nb;
    constexpr uint32_t skip_bytes = TPS * w_tap_bytes, below_bytes = TAPB * w_tap_bytes;
    constexpr uint32_t w_slot = skip_bytes > below_bytes ? skip_bytes : below_bytes;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
    Smem s;
    s.a_ring = (raw + 127u) & ~127u;
    s.w_ring = s.a_ring + p.na * kStage;
    s.w_slot = w_slot;
    s.a_full = s.w_ring + p.nw * w_slot;
    s.a_empty = s.a_full + 8u * p.na;
    s.w_full = s.a_empty + 8u * p.na;
    s.w_empty = s.w_full + 8u * p.nw;
    s.acc_full = s.w_empty + 8u * p.nw;
    s.acc_empty = s.acc_full + 32u;
    const uint32_t tmem_slot = s.acc_empty + 32u;
    const uint32_t bias_s = tmem_slot + 16u;
    uint8_t* gen = smem_raw - raw;   // generic pointer = gen + shared address
    float* bias_sp = reinterpret_cast<float*>(gen + bias_s);
    volatile uint32_t* tmem_slot_p = reinterpret_cast<volatile uint32_t*>(gen + tmem_slot);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int kbB = p.nG * 4;
    const int num_units = CG == 2 ? (p.num_tiles + 1) >> 1 : p.num_tiles;
    const int items = p.npass * num_units;
    const int item0 = CG == 2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int item_step = CG == 2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    auto tile_of = [&](int item) {
        const int unit = p.pass_fast ? item / p.npass : item % num_units;
        return CG == 2 ? 2 * unit + static_cast<int>(rank) : unit;   // may be >= num_tiles (tail)
    };
    auto pass_of = [&](int item) { return p.pass_fast ? item % p.npass : item / num_units; };

    // ---------------------------------------------------------------- setup
    for (int i = threadIdx.x; i < p.cout; i += kThreads) bias_sp[i] = p.bias[i];
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmS);
        tma_prefetch_desc(&tmB);
        if (CG == 2) {
            tma_prefetch_desc(&tmWs);
            tma_prefetch_desc(&tmWb);
        }
        for (int i = 0; i < p.na; ++i) {
            mbar_init(s.a_full + 8u * i, 1);
            mbar_init(s.a_empty + 8u * i, 2);   // both MMA issuers commit every stage
        }
        for (int i = 0; i < p.nw; ++i) {
            mbar_init(s.w_full + 8u * i, 1);
            mbar_init(s.w_empty + 8u * i, 2);
        }
        for (int i = 0; i < 4; ++i) {   // N = 64: (buffer, half); N = 128: (half, px)
            mbar_init(s.acc_full + 8u * i, 1);
            mbar_init(s.acc_empty + 8u * i, 128 * CG);   // one epilogue warp group (of both CTAs)
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

    if (warp == 0) {
        // ================================================ activation producer
        if (lane == 0) {
            Ring ra;
            for (int item = item0; item < items; item += item_step) {
                int tile = tile_of(item);
                if (tile >= p.num_tiles) tile = p.num_tiles - 1;   // tail of the last pair: a duplicate
                const Tile t = decode_tile(p, tile);
                for (int kk = 0; kk < p.kbS + p.nG; ++kk, ra.next(p.na)) {
                    // N = 64: skip blocks, then the group of the tensor below; N = 128: group 0 of
                    // the tensor below, the skip blocks, the other groups (k >= kbS: group k - kbS)
                    const int k = N == 64 ? kk : (kk == 0 ? p.kbS : (kk <= p.kbS ? kk - 1 : kk));
                    const uint32_t slot = ra.slot;
                    mbar_wait_relaxed(s.a_empty + 8u * slot, ra.phase ^ 1u);
                    const uint32_t dst = s.a_ring + slot * kStage;
                    uint32_t full = s.a_full + 8u * slot;
                    if (CG == 1) {
                        mbar_arrive_expect_tx(full, kStage);
                        if (k < p.kbS) tma_load_5d(dst, &tmS, full, (t.x0 - 1) * 8, t.y0 - 1, 0, k * 4, t.n);
                        else tma_load_4d(dst, &tmB, full, (t.x0 - 1) * 8, t.y0 - 1, (k - p.kbS) * 16, t.n);
                    } else {
                        // the leader's barrier counts the bytes of both CTAs
                        if (rank == 0) mbar_arrive_expect_tx(full, 2u * kStage);
                        full = map_to_cta(full, 0);
                        if (k < p.kbS)
                            tma_load_5d_pair(dst, &tmS, full, (t.x0 - 1) * 8, t.y0 - 1, 0, k * 4, t.n);
                        else
                            tma_load_4d_pair(dst, &tmB, full, (t.x0 - 1) * 8, t.y0 - 1, (k - p.kbS) * 16,
                                             t.n);
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ==================================================== weight producer
        if (lane == 0) {
            Ring rw;
            constexpr uint32_t rows_per_tap = w_tap_bytes >> 8;   // CG = 2: the blobs as rows of 256 B
            auto stage = [&](bool below, uint32_t tap0) {   // tap0: first tap of the stage in its blob
                const uint32_t slot = rw.slot;
                mbar_wait_relaxed(s.w_empty + 8u * slot, rw.phase ^ 1u);
                const uint32_t bytes = below ? below_bytes : skip_bytes;
                const uint32_t dst = s.w_ring + slot * w_slot;
                if (CG == 1) {
                    mbar_arrive_expect_tx(s.w_full + 8u * slot, bytes);
                    bulk_load(dst, (below ? p.wbelow : p.wskip) + static_cast<size_t>(tap0) * w_tap_bytes,
                              bytes, s.w_full + 8u * slot);
                } else {
                    if (rank == 0) mbar_arrive_expect_tx(s.w_full + 8u * slot, 2u * bytes);
                    tma_load_2d_pair(dst, below ? &tmWb : &tmWs, map_to_cta(s.w_full + 8u * slot, 0), 0,
                                     static_cast<int>(tap0 * rows_per_tap));
                }
                rw.next(p.nw);
            };
            for (int item = item0; item < items; item += item_step) {
                const uint32_t pass = static_cast<uint32_t>(pass_of(item));
                // CG = 2 layouts: [pass][kb][rank][taps][4][N/2][8]
                // the order the issuers consume (issue_half); pair index (px * 2 + py) * 4 + o4
                auto group = [&](int g) {
                    if (N == 64) {
                        for (int g4 = 0; g4 < 4; ++g4)
                            for (int px = 0; px < 2; ++px)
                                stage(true, ((pass * kbB + g * 4 + g4) * CG + rank) * 16u + px * 8);
                    } else {
                        for (int px = 0; px < 2; ++px)
                            for (int g4 = 0; g4 < 4; ++g4)
                                for (int py = 0; py < 2; ++py)
                                    stage(true, ((pass * kbB + g * 4 + g4) * CG + rank) * 16u +
                                                    (px * 2 + py) * 4);
                    }
                };
                if (N != 64) group(0);
                for (int kb = 0; kb < p.kbS; ++kb)
                    for (int tg = 0; tg < 9 / TPS; ++tg)
                        stage(false, ((pass * p.kbS + kb) * CG + rank) * 9u + tg * TPS);
                for (int g = N == 64 ? 0 : 1; g < p.nG; ++g) group(g);
            }
        }
    } else if ((warp == 1 || warp == 2) && rank == 0) {
        // ================================================= MMA issuers (two warps, one per row phase)
        // The whole warp walks the loops (all values warp-uniform); one elected lane issues.
        if (warp == 1) issue_half<N, CG, 0>(p, s, tmem_base, item0, items, item_step);
        else issue_half<N, CG, 1>(p, s, tmem_base, item0, items, item_step);
    } else if (warp >= 4) {
        // =========================================================== epilogue
        const int et = (threadIdx.x - 128) & 127;   // TMEM lane = position inside the tile
        const int py = (threadIdx.x - 128) >> 7;    // warp group = output row phase
        const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const int ty = et >> 3, tx = et & 7;
        const int H = 2 * p.H2, W = 2 * p.W2;
        const size_t plane = static_cast<size_t>(H) * W * 8;
        uint32_t li = 0;
        constexpr bool kDB = N == 64;
        for (int item = item0; item < items; item += item_step, ++li) {
            const int tile = tile_of(item);
            const int pass = pass_of(item);
            const bool in_range = tile < p.num_tiles;   // false only for the tail of the last pair
            const Tile t = decode_tile(p, in_range ? tile : p.num_tiles - 1);
            const int Y = t.y0 + ty, X = t.x0 + tx;
            const bool valid = in_range && Y < p.H2 && X < p.W2;
            const int y = 2 * Y + py;
            // bias class of this thread's two pixels (2X is never the last column, 2X + 1 never the
            // first): interior everywhere except on the image border
            const int ry = y == 0 ? 0 : (y == H - 1 ? 2 : 1);
            const int rx0 = X == 0 ? 0 : 1, rx1 = (2 * X + 1 == W - 1) ? 2 : 1;
            const bool interior = ry == 1 && rx0 == 1 && rx1 == 1;
            const bool plain = __all_sync(0xffffffffu, interior || !valid);
            const uint32_t buf = kDB ? (li & 1u) : 0u;
            const uint32_t aph = kDB ? ((li >> 1) & 1u) : (li & 1u);
            const uint32_t tcol = tmem_base + lane_sel + buf * 256u + static_cast<uint32_t>(2 * py * N);
            __nv_bfloat16* orow = p.out + (static_cast<size_t>(t.n) * (p.cout >> 3)) * plane +
                                  (static_cast<size_t>(y) * W + 2 * X) * 8;
            auto release = [&](uint32_t bar) {
                tc_fence_before();
                if (CG == 2) mbar_arrive_cluster(map_to_cta(bar, 0));   // the leader's barrier
                else mbar_arrive(bar);
            };
            if (kDB) {
                // both pixels of a position together: one 32-byte store per 8-channel group
                mbar_wait_relaxed(s.acc_full + 8u * (buf * 2u + py), aph);
                tc_fence_after();
#pragma unroll 1
                for (int c0 = 0; c0 < N; c0 += 32) {
                    uint32_t r0[32], r1[32];
                    tmem_ld32(tcol + c0, r0);
                    tmem_ld32(tcol + N + c0, r1);
                    tmem_ld_wait();
                    const int co0 = pass * N + c0;
                    const float* b0p = bias_sp + co0;
                    const float* b1p = b0p;
                    if (!plain) {   // border classes straight from global memory (rare)
                        b0p = p.btab + (ry * 3 + rx0) * p.cout + co0;
                        b1p = p.btab + (ry * 3 + rx1) * p.cout + co0;
                    }
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        uint32_t q[8];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 a = *reinterpret_cast<const float2*>(b0p + g * 8 + 2 * e);
                            const float2 b = *reinterpret_cast<const float2*>(b1p + g * 8 + 2 * e);
                            q[e] = pack_relu_bf16x2(__uint_as_float(r0[g * 8 + 2 * e]) + a.x,
                                                    __uint_as_float(r0[g * 8 + 2 * e + 1]) + a.y);
                            q[4 + e] = pack_relu_bf16x2(__uint_as_float(r1[g * 8 + 2 * e]) + b.x,
                                                        __uint_as_float(r1[g * 8 + 2 * e + 1]) + b.y);
                        }
                        if (valid) st_global_256(orow + ((co0 >> 3) + g) * plane, q);
                    }
                }
                release(s.acc_empty + 8u * (buf * 2u + py));
            } else {
                // one accumulator (= one pixel of the position) at a time, in the order the issuer
                // finishes them
#pragma unroll 1
                for (int px = 0; px < 2; ++px) {
                    mbar_wait_relaxed(s.acc_full + 8u * (py * 2 + px), aph);
                    tc_fence_after();
#pragma unroll 1
                    for (int c0 = 0; c0 < N; c0 += 32) {
                        uint32_t r[32];
                        tmem_ld32(tcol + px * N + c0, r);
                        tmem_ld_wait();
                        const int co0 = pass * N + c0;
                        const float* bp = plain ? bias_sp + co0
                                                : p.btab + (ry * 3 + (px ? rx1 : rx0)) * p.cout + co0;
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            uint4 q4;
                            const float4 a = *reinterpret_cast<const float4*>(bp + g * 8);
                            const float4 b = *reinterpret_cast<const float4*>(bp + g * 8 + 4);
                            q4.x = pack_relu_bf16x2(__uint_as_float(r[g * 8 + 0]) + a.x,
                                                    __uint_as_float(r[g * 8 + 1]) + a.y);
                            q4.y = pack_relu_bf16x2(__uint_as_float(r[g * 8 + 2]) + a.z,
                                                    __uint_as_float(r[g * 8 + 3]) + a.w);
                            q4.z = pack_relu_bf16x2(__uint_as_float(r[g * 8 + 4]) + b.x,
                                                    __uint_as_float(r[g * 8 + 5]) + b.y);
                            q4.w = pack_relu_bf16x2(__uint_as_float(r[g * 8 + 6]) + b.z,
                                                    __uint_as_float(r[g * 8 + 7]) + b.w);
                            if (valid)
                                *reinterpret_cast<uint4*>(orow + ((co0 >> 3) + g) * plane + px * 8) = q4;
                        }
                    }
                    release(s.acc_empty + 8u * (py * 2 + px));
                }
            }
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

template <int N, int CG>
size_t smem_need(int na, int nw, int cout) {
    constexpr size_t nb = N / CG;
    constexpr size_t skip = (N == 64 ? 9 : 3) * 64 * nb, below = (N == 64 ? 8 : 4) * 64 * nb;
    const size_t slot = skip > below ? skip : below;
    return 128 + static_cast<size_t>(na) * kStage + nw * slot + 16 * (na + nw) + 64 + 16 + 16 +
           sizeof(float) * cout + 64;
}

template <int N, int CG>
int launch_variant(const CUtensorMap& tmS, const CUtensorMap& tmB, const CUtensorMap& tmWs,
                   const CUtensorMap& tmWb, UpcatParams p, int num_sms, cudaStream_t stream) {
    // ring depths: 3 activation stages (2 when the unpaired weight slots are 36 KB), then as many
    // weight slots as fit (at most 8). OGL_UP_NA / OGL_UP_NW (experiments) override, clamped to what
    // fits.
    auto fit_nw = [&](int na, int want) {
        int nw = want;
        while (nw > 2 && smem_need<N, CG>(na, nw, p.cout) > static_cast<size_t>(kMaxSmem)) --nw;
        return nw;
    };
    static const int na_env = getenv("OGL_UP_NA") ? atoi(getenv("OGL_UP_NA")) : 0;
    static const int nw_env = getenv("OGL_UP_NW") ? atoi(getenv("OGL_UP_NW")) : 0;
    // measured (gpurun_out/r2_exp_upring.jsonl): N = 128 is faster with 2 activation stages and 8
    // weight slots (0.977 -> 0.935 ms, 0.843 -> 0.818 ms), N = 64 with 3 and 5 (0.95 vs 1.08 ms)
    p.na = (na_env >= 2 && na_env <= 4) ? na_env : (N == 128 ? 2 : 3);
    p.nw = fit_nw(p.na, (nw_env >= 2 && nw_env <= 12) ? nw_env : 8);
    if (smem_need<N, CG>(p.na, p.nw, p.cout) > static_cast<size_t>(kMaxSmem) || p.nw < 3) {
        p.na = 2;
        p.nw = fit_nw(2, (nw_env >= 2 && nw_env <= 12) ? nw_env : 8);
    }
    const size_t smem = smem_need<N, CG>(p.na, p.nw, p.cout);
    if (smem > static_cast<size_t>(kMaxSmem)) return fail("upcat layer: shared memory budget exceeded");
    if (CG == 1) {
        const int items = p.npass * p.num_tiles;
        const int grid = items < num_sms ? items : num_sms;
        p.pass_fast = (p.npass > 1 && grid % p.npass == 0) ? 1 : 0;
        upcat_tc_kernel<N, 1><<<grid, kThreads, smem, stream>>>(tmS, tmB, tmWs, tmWb, p);
        OGL_CUDA(cudaGetLastError());
        return 0;
    }
    const int grid = num_sms & ~1;
    // pass fastest: the clusters that re-read a tile pair for its other pass run at the same time, so
    // the re-reads hit L2; with an even number of clusters each keeps one pass
    p.pass_fast = (p.npass > 1 && (grid / 2) % p.npass == 0) ? 1 : 0;
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
    OGL_CUDA(cudaLaunchKernelEx(&cfg, upcat_tc_kernel<N, 2>, tmS, tmB, tmWs, tmWb, p));
    return 0;
}

}  // namespace

#ifndef OGL_F16
namespace {
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
}  // namespace

// Host packer. w3: folded conv weights [f][2f][3][3] (input channels [0, f) = skip, [f, 2f) = up,
// unet.py:86), b3 [f]; wt: ConvTranspose2d weights [2f][f][2][2], bt [f].
int build_upcat_host(const float* w3, const float* b3, const float* wt, const float* bt, int f,
                     bool f16, UpcatHost* out) {
    if (f != 64 && f != 128 && f != 256) return fail("upcat layer: f must be 64, 128 or 256");
    const int N = f < 128 ? f : 128, npass = f / N, kbS = f / 32, kbB = 2 * f / 32, cin3 = 2 * f;
    out->f = f;
    out->N = N;
    out->npass = npass;
    auto W3 = [&](int co, int ci, int dy, int dx) -> double {
        return w3[((static_cast<size_t>(co) * cin3 + ci) * 3 + dy) * 3 + dx];
    };
    // ---- skip half: [pass][kb][tap][4][N][8] and its CTA-pair form [pass][kb][rank][tap][4][N/2][8]
    out->wskip.assign(static_cast<size_t>(f) * f * 9, 0);
    out->wskip_pair.assign(out->wskip.size(), 0);
    for (int pass = 0; pass < npass; ++pass)
        for (int kb = 0; kb < kbS; ++kb)
            for (int tap = 0; tap < 9; ++tap)
                for (int c = 0; c < 4; ++c)
                    for (int n = 0; n < N; ++n)
                        for (int e = 0; e < 8; ++e) {
                            const int co = pass * N + n, ci = kb * 32 + c * 8 + e;
                            const uint16_t v = operand_bits(W3(co, ci, tap / 3, tap % 3), f16);
                            out->wskip[(((((static_cast<size_t>(pass) * kbS + kb) * 9 + tap) * 4 + c) * N + n) * 8) + e] = v;
                            const int r = n / (N / 2), nn = n % (N / 2);
                            out->wskip_pair[((((((static_cast<size_t>(pass) * kbS + kb) * 2 + r) * 9 + tap) * 4 + c) * (N / 2) + nn) * 8) + e] = v;
                        }
    // ---- composed half. comp[pr][ci][co], pr = (px * 2 + py) * 4 + (oyi * 2 + oxi): the order the
    // kernel's weight stages use. W3 of the `up` channels is transposed once so that the innermost
    // loop runs over contiguous output channels.
    std::vector<double> w3t(static_cast<size_t>(9) * f * f);   // [tap][c][co]
    for (int co = 0; co < f; ++co)
        for (int c = 0; c < f; ++c)
            for (int tap = 0; tap < 9; ++tap)
                w3t[(static_cast<size_t>(tap) * f + c) * f + co] = W3(co, f + c, tap / 3, tap % 3);
    std::vector<double> comp(static_cast<size_t>(16) * 2 * f * f, 0.0);
    for (int px = 0; px < 2; ++px)
        for (int py = 0; py < 2; ++py)
            for (int dy = 0; dy < 3; ++dy)
                for (int dx = 0; dx < 3; ++dx) {
                    // tap (dy, dx) of output phase (py, px): `up` phase q at half-resolution offset o
                    const int qy = src_phase(py, dy), qx = src_phase(px, dx);
                    const int oyi = src_cell(py, dy) - py, oxi = src_cell(px, dx) - px;   // 0 or 1
                    const int pr = (px * 2 + py) * 4 + oyi * 2 + oxi;
                    double* C = comp.data() + static_cast<size_t>(pr) * 2 * f * f;
                    const double* T = w3t.data() + static_cast<size_t>(dy * 3 + dx) * f * f;
                    for (int ci = 0; ci < 2 * f; ++ci) {
                        double* crow = C + static_cast<size_t>(ci) * f;
                        for (int c = 0; c < f; ++c) {
                            const double a = wt[((static_cast<size_t>(ci) * f + c) * 2 + qy) * 2 + qx];
                            const double* trow = T + static_cast<size_t>(c) * f;
                            for (int co = 0; co < f; ++co) crow[co] += a * trow[co];
                        }
                    }
                }
    out->wbelow.assign(static_cast<size_t>(16) * 2 * f * f, 0);
    out->wbelow_pair.assign(out->wbelow.size(), 0);
    for (int pass = 0; pass < npass; ++pass)
        for (int kb = 0; kb < kbB; ++kb)
            for (int pr = 0; pr < 16; ++pr)
                for (int c = 0; c < 4; ++c)
                    for (int n = 0; n < N; ++n)
                        for (int e = 0; e < 8; ++e) {
                            const int co = pass * N + n, ci = kb * 32 + c * 8 + e;
                            const uint16_t v = operand_bits(
                                comp[(static_cast<size_t>(pr) * 2 * f + ci) * f + co], f16);
                            out->wbelow[(((((static_cast<size_t>(pass) * kbB + kb) * 16 + pr) * 4 + c) * N + n) * 8) + e] = v;
                            const int r = n / (N / 2), nn = n % (N / 2);
                            out->wbelow_pair[((((((static_cast<size_t>(pass) * kbB + kb) * 2 + r) * 16 + pr) * 4 + c) * (N / 2) + nn) * 8) + e] = v;
                        }
    // ---- bias per (row class, column class): bt reaches a pixel only through in-image taps
    out->btab.assign(static_cast<size_t>(9) * f, 0.f);
    std::vector<double> tb(static_cast<size_t>(9) * f, 0.0);   // [tap][co] = sum_c W3[co][f+c][tap] bt[c]
    for (int tap = 0; tap < 9; ++tap)
        for (int c = 0; c < f; ++c)
            for (int co = 0; co < f; ++co)
                tb[static_cast<size_t>(tap) * f + co] += w3t[(static_cast<size_t>(tap) * f + c) * f + co] * bt[c];
    for (int ry = 0; ry < 3; ++ry)
        for (int rx = 0; rx < 3; ++rx)
            for (int co = 0; co < f; ++co) {
                double v = b3[co];
                for (int dy = 0; dy < 3; ++dy) {
                    if ((ry == 0 && dy == 0) || (ry == 2 && dy == 2)) continue;
                    for (int dx = 0; dx < 3; ++dx) {
                        if ((rx == 0 && dx == 0) || (rx == 2 && dx == 2)) continue;
                        v += tb[static_cast<size_t>(dy * 3 + dx) * f + co];
                    }
                }
                out->btab[(static_cast<size_t>(ry) * 3 + rx) * f + co] = static_cast<float>(v);
            }
    return 0;
}
#endif  // !OGL_F16

int upcat_tc_init() {
    OGL_CUDA(set_wait_cfg());
    OGL_CUDA(cudaFuncSetAttribute(upcat_tc_kernel<64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    OGL_CUDA(cudaFuncSetAttribute(upcat_tc_kernel<64, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    OGL_CUDA(cudaFuncSetAttribute(upcat_tc_kernel<128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    OGL_CUDA(cudaFuncSetAttribute(upcat_tc_kernel<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    return 0;
}

// skip_s2d: [B][f/8][4][H/2][W/2][8]; below: [B][2f/8][H/2][W/2][8]; out: [B][f/8][H][W][8]. H, W = the
// level's own resolution (even).
int launch_upcat_tc(const UpcatLayer& L, const __nv_bfloat16* skip_s2d, const __nv_bfloat16* below,
                    int B, int H, int W, __nv_bfloat16* out, int num_sms, cudaStream_t stream,
                    int cta_group) {
    if (H < 2 || W < 2 || H % 2 || W % 2) return fail("upcat layer needs even, non-empty H and W");
    if (!L.wskip || !L.wbelow || !L.bias || !L.btab || !skip_s2d || !below || !out)
        return fail("upcat layer is not built");
    if ((L.N != 64 && L.N != 128) || L.f % L.N) return fail("upcat layer: bad N");
    UpcatParams p;
    memset(&p, 0, sizeof p);
    p.bias = L.bias;
    p.btab = L.btab;
    p.out = out;
    p.kbS = L.f / 32;
    p.nG = 2 * L.f / 128;
    p.npass = L.npass;
    p.cout = L.f;
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
    // CTA pairs when there is a tile per SM (cta_group 2) or whenever there are two (3: unit tests)
    const bool pair = cta_group >= 2 && L.wskip2 && L.wbelow2 && num_sms >= 2 &&
                      p.num_tiles >= (cta_group == 3 ? 2 : num_sms);
    p.wskip = pair ? L.wskip2 : L.wskip;
    p.wbelow = pair ? L.wbelow2 : L.wbelow;

    CUtensorMap tmS, tmB, tmWs, tmWb;
    {
        const uint64_t W2 = p.W2, H2 = p.H2;
        const uint64_t dims[5] = {W2 * 8, H2, 4, static_cast<uint64_t>(L.f / 8), static_cast<uint64_t>(B)};
        const uint64_t str[4] = {W2 * 16, W2 * 16 * H2, W2 * 16 * H2 * 4, W2 * 16 * H2 * 4 * (L.f / 8)};
        const uint32_t box[5] = {kHW * 8, kHH, 4, 4, 1};
        if (encode_bf16_map(&tmS, skip_s2d, 5, dims, str, box)) return 1;
        const uint64_t dimb[4] = {W2 * 8, H2, static_cast<uint64_t>(2 * L.f / 8), static_cast<uint64_t>(B)};
        const uint64_t strb[3] = {W2 * 16, W2 * 16 * H2, W2 * 16 * H2 * (2 * L.f / 8)};
        const uint32_t boxb[4] = {kHW * 8, kHH, 16, 1};
        if (encode_bf16_map(&tmB, below, 4, dimb, strb, boxb)) return 1;
    }
    if (pair) {
        // the pair blobs as rows of 256 bytes; one box = this CTA's half of a weight stage
        const uint32_t nb = L.N / 2, tap_rows = 64 * nb / 256;
        const uint64_t rows_s = static_cast<uint64_t>(L.f) * L.f * 9 * 2 / 256;
        const uint64_t rows_b = static_cast<uint64_t>(16) * 2 * L.f * L.f * 2 / 256;
        const uint64_t str[1] = {256};
        const uint64_t dims_s[2] = {128, rows_s}, dims_b[2] = {128, rows_b};
        const uint32_t box_s[2] = {128, (L.N == 64 ? 9u : 3u) * tap_rows};
        const uint32_t box_b[2] = {128, (L.N == 64 ? 8u : 4u) * tap_rows};
        if (encode_bf16_map(&tmWs, p.wskip, 2, dims_s, str, box_s)) return 1;
        if (encode_bf16_map(&tmWb, p.wbelow, 2, dims_b, str, box_b)) return 1;
        return L.N == 64 ? launch_variant<64, 2>(tmS, tmB, tmWs, tmWb, p, num_sms, stream)
                         : launch_variant<128, 2>(tmS, tmB, tmWs, tmWb, p, num_sms, stream);
    }
    tmWs = tmS;
    tmWb = tmS;
    return L.N == 64 ? launch_variant<64, 1>(tmS, tmB, tmWs, tmWb, p, num_sms, stream)
                     : launch_variant<128, 1>(tmS, tmB, tmWs, tmWb, p, num_sms, stream);
}

}  // namespace ogl
