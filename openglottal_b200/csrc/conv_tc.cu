// Implicit-GEMM conv3x3 / convT2x2 on the sm_100a tensor cores (tcgen05 + TMEM), fed by TMA.
//
// Replaces, per fused group, the PyTorch calls of the reference U-Net
// (/root/reference/openglottal/models/unet.py:24-29 Conv2d+BatchNorm2d+ReLU, :59,79 MaxPool2d,
//  :69,82 ConvTranspose2d, :86 torch.cat, :72,88 1x1 head; utils.py:237,241 sigmoid+threshold;
//  features.py:238 np.sum of the mask).
//
// GEMM view: D[pixels, Cout] = sum over taps (dy,dx) and channels of
//            A_tap[pixels, Cin] * W_tap[Cin, Cout].
//   * M = 128 pixels = an 8(x) x 16(y) patch; a CTA tile is S sub-tiles of 16x16 px
//     (2 patches each), so up to 4 accumulators of N columns live in TMEM (double-buffered
//     when 2*S*N <= 256 so the epilogue of tile i overlaps the MMAs of tile i+1).
//   * A operand: one TMA box per (sub-tile, 32-channel block) brings the 18x18 halo tile of
//     4 channel planes; the 9 taps are 9 start-address offsets into that one tile
//     (zero padding = TMA out-of-bounds fill). Nothing is re-read per tap.
//   * B operand: host-packed weights [pass][32-ch block][tap][4][N][8], streamed with 1-D bulk
//     copies, TPS taps per stage (9 for N <= 64, 3 for N = 128, 1 for convT) so the issuer
//     pays one barrier wait per 8*TPS MMAs.
//   * torch.cat([skip, up]) is two tensor maps walked back to back in the K loop.
//   * Warp roles: w0 = activation TMA producer, w1 and w2 = MMA issuers (each issues half of the
//     tile's accumulators; w2 also allocates TMEM), w3 = weight producer, w4..11 = epilogue
//     (TMEM -> regs -> bias/ReLU/pool/head -> HBM); epilogue warp w reads TMEM lanes 32*(w%4)..
//     and handles sub-tile (w-4)/4 of each tile.
//   * CG = 2 (conv3x3, N >= 64): the two CTAs of a cluster run one 256-row tcgen05.mma.cta_group::2
//     per MMA and each stages half of the weight columns (see the kernel's comment).
#include "internal.h"
#include "ptx.cuh"

#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#ifdef OGL_F16   // conv_tc_f16.cu: the same kernels with f16 operands, exported under other names
#define launch_conv_tc launch_conv_tc_f16
#define conv_tc_init conv_tc_init_f16
#endif

namespace ogl {

namespace {

constexpr int kHalo = 18;                          // 16 + 2
// one 8-channel plane of a halo tile = 18 * 18 * 16 B; 32 channels of a sub-tile = 20736 B
constexpr int kThreads = 384;       // 4 control warps + 8 epilogue warps
constexpr int kEpiThreads = 256;
constexpr int kTmemCols = 512;

struct ConvParams {
    const __nv_bfloat16* wpack;
    const float* bias;
    __nv_bfloat16* out;
    __nv_bfloat16* out_pool;
    const float* head_w;
    float* logits;
    uint8_t* mask;
    int32_t* area;
    float head_b, logit_thr;
    int kb0, kb1;    // 32-channel K blocks taken from source 0 / source 1
    int taps;        // 9 or 1
    int N, npass;    // MMA N per pass, number of passes over N_total
    int cout;        // channels of the output tensor
    int H, W, B;     // resolution of the INPUT feature map
    int S;           // sub-tiles (16x16 px) per CTA tile
    int tiles_x, tiles_y, total_sub, num_tiles;
    unsigned long long magic_tx, magic_tpf;  // ceil(2^40 / d) for d = tiles_x, tiles_x*tiles_y
    int na, nw;      // ring depths
    int acc_bufs;    // 1 or 2 TMEM accumulator sets
    int pass_fast;   // items ordered (tile, pass) with the pass fastest: the CTAs that re-read a
                     // tile for its other passes run at the same time, so the re-reads hit L2;
                     // with gridDim % npass == 0 each CTA still keeps one pass (its weights)
    int reverse;     // walk the tiles from the last frame to the first (see launch_conv_tc)
    int split;       // N = 128, S = 2: per-accumulator barriers (see kSplit in the kernel)
    int acc_half;    // the other layers: one accumulator-empty barrier per (buffer, issuer half)
    int out_s2d;     // `out` is written space-to-depth [frame][C/8][phase][H/2][W/2][8] (skip tensors
                     // of levels 1-3, read by upcat_tc.cu); the pooled tensor stays plain
    int dbg;         // experiment switches (OGL_DBG): 1 no MMA, 2 no stores, 4 no epilogue math,
                     // 8 no activation TMA, 16 no weight copies. Results are garbage when set.
};

struct SubTile {
    int n, y0, x0;
};
// sub-tile index -> (frame, y0, x0); divisions by multiply-shift (exact for st * d < 2^40)
__device__ __forceinline__ SubTile decode_sub(const ConvParams& p, int st) {
    SubTile s;
    const unsigned u = static_cast<unsigned>(st);
    const unsigned n = static_cast<unsigned>((u * p.magic_tpf) >> 40);
    const unsigned rem = u - n * static_cast<unsigned>(p.tiles_x * p.tiles_y);
    const unsigned ty = static_cast<unsigned>((rem * p.magic_tx) >> 40);
    const unsigned tx = rem - ty * static_cast<unsigned>(p.tiles_x);
    s.n = static_cast<int>(n);
    s.y0 = static_cast<int>(ty) * 16;
    s.x0 = static_cast<int>(tx) * 16;
    return s;
}

// 16-bit operand pairs of this unit's type (bf16, or f16 in conv_tc_f16.cu)
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) { return pack_x2<kF16>(lo, hi); }
__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) { return max_x2<kF16>(a, b); }

// CG = 1: one CTA per tile. CG = 2: the two CTAs of a cluster (one TPC) take two neighbouring
// tiles of the same pass and run their MMAs as ONE 256-row tcgen05.mma.cta_group::2: each CTA
// stages its own activations but only HALF of the weight columns, so the weight traffic
// (L2 -> shared memory) and the B-operand shared-memory reads per SM are halved.
template <int EPI, int TPS, int S, int CG>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmW, const ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
    // activation tile of a sub-tile: 18x18 with the conv's halo; the transposed conv reads the
    // 16x16 pixels only (21 % fewer bytes L2 -> shared memory per K block)
    constexpr int kEdge = EPI == EPI_CONVT ? 16 : kHalo;
    constexpr int kPlaneB = kEdge * kEdge * 16;
    constexpr int kSubB = 4 * kPlaneB;

    constexpr uint32_t a_stage_bytes = static_cast<uint32_t>(S) * kSubB;
    const uint32_t nb = static_cast<uint32_t>(p.N) / CG;          // weight columns staged per CTA
    const uint32_t w_tap_bytes = 64u * nb;
    const uint32_t w_stage_bytes = TPS * w_tap_bytes;
    const uint32_t a_ring = smem_base;
    const uint32_t w_ring = a_ring + p.na * a_stage_bytes;
    const uint32_t bar_base = w_ring + p.nw * w_stage_bytes;  // 8 B aligned (sizes are x64)
    const uint32_t a_full = bar_base;
    const uint32_t a_empty = a_full + 8u * p.na;
    const uint32_t w_full = a_empty + 8u * p.na;
    const uint32_t w_empty = w_full + 8u * p.nw;
    const uint32_t acc_full = w_empty + 8u * p.nw;
    const uint32_t acc_empty = acc_full + 32u;
    const uint32_t tmem_slot = acc_empty + 32u;
    // N = 128 with 2 sub-tiles fills TMEM (4 x 128 columns), so the accumulator set cannot be
    // double-buffered. Instead each of the 4 accumulators has its own full/empty barrier, the
    // first and the last K block of a tile are issued accumulator-major, and all 8 epilogue
    // warps drain one accumulator at a time (half of its columns per warp group): accumulator
    // m is drained while the MMAs of accumulators m+1.. (last K block) and then those of the
    // next tile's first K block (accumulators ..m-1) run.
    const bool kSplit = (TPS == 3 && S == 2) && p.split;
    const uint32_t bias_s = tmem_slot + 16u;  // floats: bias[cout] then head_w[32]
    // generic pointers to the small fp32 tables
    float* bias_sp = reinterpret_cast<float*>(smem_raw + (bias_s - smem_u32(smem_raw)));
    volatile uint32_t* tmem_slot_p =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int kb_total = p.kb0 + p.kb1;
    // work items: (pass, tile) for CG = 1, (pass, pair of tiles) for CG = 2
    const int num_units = CG == 2 ? (p.num_tiles + 1) >> 1 : p.num_tiles;
    const int items = p.npass * num_units;
    const int item0 = CG == 2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int item_step = CG == 2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    auto tile_of = [&](int item) {
        int unit = p.pass_fast ? item / p.npass : item % num_units;
        if (p.reverse) unit = num_units - 1 - unit;
        return CG == 2 ? 2 * unit + static_cast<int>(rank) : unit;   // may be >= num_tiles (tail)
    };
    auto pass_of = [&](int item) { return p.pass_fast ? item % p.npass : item / num_units; };

    // ---------------------------------------------------------------- setup
    for (int i = threadIdx.x; i < p.cout; i += kThreads) bias_sp[i] = p.bias[i];
    if (EPI == EPI_HEAD) {
        if (threadIdx.x < 32) bias_sp[p.cout + threadIdx.x] = p.head_w[threadIdx.x];
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA0);
        tma_prefetch_desc(&tmA1);
        if (CG == 2) tma_prefetch_desc(&tmW);
        for (int i = 0; i < p.na; ++i) {
            mbar_init(a_full + 8u * i, 1);
            mbar_init(a_empty + 8u * i, 2);   // both MMA issuers commit every stage
        }
        for (int i = 0; i < p.nw; ++i) {
            mbar_init(w_full + 8u * i, 1);
            mbar_init(w_empty + 8u * i, 2);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(acc_full + 8u * i, 1);
            // released by all 8 epilogue warps, or (acc_half) by the warp group that drains the
            // half of the tile one issuer writes; CG = 2: the epilogues of both CTAs
            mbar_init(acc_empty + 8u * i, ((kSplit || !p.acc_half) ? kEpiThreads : kEpiThreads / 2) * CG);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        if (CG == 2) tmem_alloc_pair(tmem_slot, kTmemCols);
        else tmem_alloc(tmem_slot, kTmemCols);
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
                const int tile = tile_of(item);
                for (int kb = 0; kb < kb_total; ++kb, ra.next(p.na)) {
                    const uint32_t s = ra.slot;
                    const uint32_t ph = ra.phase;
                    mbar_wait_relaxed(a_empty + 8u * s, ph ^ 1u);
                    if (CG == 1 && (p.dbg & 8)) {
                        mbar_arrive(a_full + 8u * s);
                        continue;
                    }
                    // CG = 2: the leader's barrier counts the bytes of both CTAs
                    uint32_t full = a_full + 8u * s;
                    if (CG == 1) {
                        mbar_arrive_expect_tx(full, a_stage_bytes);
                    } else {
                        if (rank == 0) mbar_arrive_expect_tx(full, 2u * a_stage_bytes);
                        full = map_to_cta(full, 0);
                    }
                    const CUtensorMap* tm = kb < p.kb0 ? &tmA0 : &tmA1;
                    const int plane0 = (kb < p.kb0 ? kb : kb - p.kb0) * 4;
                    for (int sub = 0; sub < S; ++sub) {
                        int st = tile * S + sub;
                        if (st >= p.total_sub) st = p.total_sub - 1;  // tail: load a duplicate
                        const SubTile t = decode_sub(p, st);
                        const uint32_t dst = a_ring + s * a_stage_bytes + sub * kSubB;
                        constexpr int kPad = EPI == EPI_CONVT ? 0 : 1;
                        if (CG == 1)
                            tma_load_4d(dst, tm, full, (t.x0 - kPad) * 8, t.y0 - kPad, plane0, t.n);
                        else
                            tma_load_4d_pair(dst, tm, full, (t.x0 - kPad) * 8, t.y0 - kPad, plane0, t.n);
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ==================================================== weight producer
        if (lane == 0) {
            Ring rw;
            const uint8_t* wbytes = reinterpret_cast<const uint8_t*>(p.wpack);
            const int tgs = p.taps / TPS;
            for (int item = item0; item < items; item += item_step) {
                const int pass = pass_of(item);
                for (int kb = 0; kb < kb_total; ++kb) {
                    for (int tg = 0; tg < tgs; ++tg, rw.next(p.nw)) {
                        const uint32_t s = rw.slot;
                        const uint32_t ph = rw.phase;
                        mbar_wait_relaxed(w_empty + 8u * s, ph ^ 1u);
                        if (CG == 1 && (p.dbg & 16)) {
                            mbar_arrive(w_full + 8u * s);
                            continue;
                        }
                        if (CG == 1) {
                            mbar_arrive_expect_tx(w_full + 8u * s, w_stage_bytes);
                            const size_t off =
                                (static_cast<size_t>(pass * kb_total + kb) * p.taps + tg * TPS) *
                                w_tap_bytes;
                            bulk_load(w_ring + s * w_stage_bytes, wbytes + off, w_stage_bytes,
                                      w_full + 8u * s);
                        } else {
                            // pair layout [pass][kb][rank][tap][4][N/2][8], seen by TMA as rows
                            // of 256 bytes; this CTA's half of the stage is one box
                            if (rank == 0) mbar_arrive_expect_tx(w_full + 8u * s, 2u * w_stage_bytes);
                            const uint32_t rows_per_tap = w_tap_bytes >> 8;
                            const int row0 = static_cast<int>(
                                ((static_cast<uint32_t>(pass * kb_total + kb) * 2u + rank) * p.taps +
                                 tg * TPS) * rows_per_tap);
                            tma_load_2d_pair(w_ring + s * w_stage_bytes, &tmW,
                                             map_to_cta(w_full + 8u * s, 0), 0, row0);
                        }
                    }
                }
            }
        }
    } else if ((warp == 1 || warp == 2) && rank == 0) {
        // ================================================= MMA issuers (two warps)
        // The tensor pipe accepts an MMA only when the previous one is (nearly) done and a wait on
        // an mbarrier costs ~120 cycles even on a long-completed phase, so a single issuer's
        // barrier waits are bubbles in the pipe (profiles/microbench_mma_issuer_bubbles_r01.txt).
        // Two warps walk the SAME sequence of stages; warp 1 issues the MMAs of the tile's first
        // half of the accumulators (sub-tile 0), warp 2 those of the second half, each commits
        // every stage it has read (the empty barriers count two arrivals). While one waits, the
        // other's MMAs keep the pipe busy; per-accumulator summation order is unchanged.
        const uint32_t me = static_cast<uint32_t>(warp - 1);
        // The whole warp walks the loops (all values warp-uniform); one elected lane issues.
        // Descriptors differ only in their start-address field, so each MMA is one 32-bit add.
        // CG = 2: only the leader CTA issues; shared-memory offsets are the same in both CTAs.
        const uint32_t idesc = CG == 2 ? make_idesc_bf16_pair(p.N) : make_idesc_bf16(p.N);
        constexpr uint32_t lbo_a = kPlaneB, sbo_a = kEdge * 16;
        const uint32_t lbo_b = 16u * nb, sbo_b = 128u;
        const uint64_t adesc0 = make_smem_desc(0, lbo_a, sbo_a);
        const uint64_t bdesc0 = make_smem_desc(0, lbo_b, sbo_b);
        const uint32_t acc_cols = static_cast<uint32_t>(2 * S * p.N);
        const uint32_t bstep = (2u * lbo_b) >> 4;   // second K=16 half of a 32-channel block
        const uint32_t btap = 4u * nb;              // one tap of B, in 16-byte units
        // this issuer's accumulators: mt = me * S + m (S == 2: sub-tile me; S == 1: x-half me)
        const uint32_t a_me = S == 2 ? me * (kSubB >> 4) : me * 8u;
        const uint32_t d_me = me * S * static_cast<uint32_t>(p.N);
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
        };
        // ring positions advanced without integer division (ptx.cuh: Ring)
        Ring ra, rw, rb;
        const int tgs = p.taps / TPS;
        uint32_t li = 0;
        for (int item = item0; item < items; item += item_step, ++li, rb.next(p.acc_bufs)) {
            const uint32_t buf = kSplit ? 0u : rb.slot;
            const uint32_t aph = kSplit ? (li & 1u) : rb.phase;
            if (!kSplit) {
                // acc_half: only this issuer's half of the tile has to be drained, so the two
                // halves run as independent MMA -> epilogue chains that fall out of phase and keep
                // the pipe busy where TMEM has no room for a second accumulator set (convT)
                wait_acc_empty(acc_empty + 8u * (p.acc_half ? buf * 2u + me : buf), aph ^ 1u);
                tc_fence_after();
            }
            const uint32_t d0 = tmem_base + buf * acc_cols;
            for (int kb = 0; kb < kb_total; ++kb, ra.next(p.na)) {
                const uint32_t sa = ra.slot;
                mbar_wait(a_full + 8u * sa, ra.phase);
                const uint32_t abase = a_ring + sa * a_stage_bytes;
                if (kSplit && (kb == 0 || kb == kb_total - 1)) {
                    // accumulator-major K block: its three weight stages (one per tap row)
                    // are held together; needs nw >= 3
                    uint32_t wst[3];
#pragma unroll
                    for (int tg = 0; tg < 3; ++tg, rw.next(p.nw)) {
                        wst[tg] = rw.slot;
                        mbar_wait(w_full + 8u * wst[tg], rw.phase);
                    }
                    tc_fence_after();
                    for (int m = 0; m < 2; ++m) {
                        const uint32_t mt = me * 2u + m;
                        if (kb == 0) {
                            wait_acc_empty(acc_empty + 8u * mt, aph ^ 1u);
                            tc_fence_after();
                        }
                        if (elect_one()) {
#pragma unroll
                            for (int tg = 0; tg < 3; ++tg) {
                                const uint64_t ad =
                                    adesc0 + ((abase + static_cast<uint32_t>(tg) * kHalo * 16u) >> 4);
                                const uint64_t bd = bdesc0 + ((w_ring + wst[tg] * w_stage_bytes) >> 4);
#pragma unroll
                                for (int t = 0; t < 3; ++t) {
                                    if (p.dbg & 1) break;
#pragma unroll
                                    for (int j = 0; j < 2; ++j) {
                                        const uint32_t aoff =
                                            t + a_me + j * ((2u * lbo_a) >> 4) + m * 8u;
                                        mma(d0 + mt * p.N, ad + aoff, bd + (t * btap + j * bstep),
                                            idesc, (kb | tg | t | j) ? 1u : 0u);
                                    }
                                }
                            }
                            if (kb == kb_total - 1) commit(acc_full + 8u * mt);
                        }
                        __syncwarp();
                    }
                    if (elect_one()) {
#pragma unroll
                        for (int tg = 0; tg < 3; ++tg) commit(w_empty + 8u * wst[tg]);
                        commit(a_empty + 8u * sa);
                    }
                    __syncwarp();
                    continue;
                }
                for (int tg = 0; tg < tgs; ++tg, rw.next(p.nw)) {
                    const uint32_t sw = rw.slot;
                    mbar_wait(w_full + 8u * sw, rw.phase);
                    tc_fence_after();
                    // TPS == 9: all taps unrolled; TPS == 3: tg is the tap row dy; TPS == 1: convT
                    const uint32_t row_off =
                        TPS == 3 ? static_cast<uint32_t>(tg) * kHalo * 16u : 0u;
                    const uint64_t ad = adesc0 + ((abase + row_off) >> 4);
                    const uint64_t bd = bdesc0 + ((w_ring + sw * w_stage_bytes) >> 4);
                    const uint32_t first = (kb | tg) != 0 ? 1u : 0u;
                    const bool last_tg = tg == tgs - 1;
                    if (elect_one()) {
#pragma unroll
                        for (int t = 0; t < TPS; ++t) {
                            if (p.dbg & 1) break;
                            // tap offset inside the halo tile, in 16-byte units
                            constexpr int kCenter = EPI == EPI_CONVT ? 0 : kHalo + 1;
                            const uint32_t toff = TPS == 9   ? (t / 3) * kHalo + (t % 3)
                                                  : TPS == 3 ? t
                                                             : kCenter;
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
#pragma unroll
                                for (int m = 0; m < S; ++m) {
                                    const uint32_t aoff =
                                        toff + a_me + j * ((2u * lbo_a) >> 4) + m * 8u;
                                    mma(d0 + d_me + m * p.N, ad + aoff, bd + (t * btap + j * bstep),
                                        idesc, (t | j) ? 1u : first);
                                }
                            }
                        }
                        commit(w_empty + 8u * sw);
                        if (last_tg) commit(a_empty + 8u * sa);
                        if (!kSplit && last_tg && kb == kb_total - 1)
                            commit(acc_full + 8u * (buf * 2u + me));   // this issuer's half
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= 4) {
        // =========================================================== epilogue
        const int et = (threadIdx.x - 128) & 127;  // TMEM lane == pixel index inside a patch
        const int egrp = (threadIdx.x - 128) >> 7; // warp group: S == 2 -> one sub-tile (both
                                                   // x-halves); S == 1 -> one x-half of the tile
        const int esub = S == 2 ? egrp : 0;
        const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const int py = et >> 3, px = et & 7;
        const uint32_t acc_cols = static_cast<uint32_t>(2 * S * p.N);
        const int OH = EPI == EPI_CONVT ? 2 * p.H : p.H;
        const int OW = EPI == EPI_CONVT ? 2 * p.W : p.W;
        uint32_t li = 0;
        Ring rb;
        auto release_acc = [](uint32_t bar) {
            if (CG == 2) mbar_arrive_cluster(map_to_cta(bar, 0));  // the leader's barrier
            else mbar_arrive(bar);
        };
        for (int item = item0; item < items; item += item_step, ++li, rb.next(p.acc_bufs)) {
            const int tile = tile_of(item);
            const int pass = pass_of(item);
            const uint32_t buf = kSplit ? 0u : rb.slot;
            const uint32_t aph = kSplit ? (li & 1u) : rb.phase;
            if (!kSplit) {
                mbar_wait_relaxed(acc_full + 8u * (buf * 2u + egrp), aph);   // the issuer of this half
                tc_fence_after();
            }
#pragma unroll 1
            for (int u = 0; u < (kSplit ? 4 : S); ++u) {
                // split mode: every warp drains accumulator u (its group's half of the columns);
                // otherwise a warp group owns one sub-tile (S == 2) or one x-half (S == 1)
                const int sub = kSplit ? (u >> 1) : esub;
                const int half = kSplit ? (u & 1) : (S == 2 ? u : egrp);
                const int mt = sub * 2 + half;
                const int st = tile * S + sub;
                const bool in_range = st < p.total_sub;  // warp-uniform
                const SubTile t = decode_sub(p, in_range ? st : p.total_sub - 1);
                if (kSplit) {
                    mbar_wait_relaxed(acc_full + 8u * mt, aph);
                    tc_fence_after();
                }
                const int c_begin = kSplit ? egrp * 64 : 0;
                const int c_end = (p.dbg & 4) ? 0 : (kSplit ? c_begin + 64 : p.N);
                const int y = t.y0 + py;
                const int x = t.x0 + half * 8 + px;
                // partial tiles at the right / bottom edge: compute everything, store nothing
                const bool valid = in_range && y < p.H && x < p.W && !(p.dbg & 2);
                const uint32_t tcol = tmem_base + lane_sel + buf * acc_cols + mt * p.N;
                float zacc = 0.f;
                if (EPI == EPI_CONVT && c_end > 0) {
                    // pass = (dy, block of cb output channels); columns = [dx][cb]. One thread
                    // owns output pixels (2y+dy, 2x) and (2y+dy, 2x+1): 32 contiguous bytes.
                    const int cb = p.N >> 1;
                    const int nblk = p.cout / cb;
                    const int qy = pass / nblk;
                    const int blk = pass - qy * nblk;
                    const size_t plane = static_cast<size_t>(OH) * OW * 8;
                    for (int c0 = 0; c0 < cb; c0 += 32) {
                        uint32_t r0[32], r1[32];
                        tmem_ld32(tcol + c0, r0);
                        tmem_ld32(tcol + cb + c0, r1);
                        tmem_ld_wait();
                        const int co0 = blk * cb + c0;
                        __nv_bfloat16* optr =
                            p.out + (static_cast<size_t>(t.n) * (p.cout >> 3) + (co0 >> 3)) * plane +
                            (static_cast<size_t>(2 * y + qy) * OW + 2 * x) * 8;
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            uint32_t q[8];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 b2 =
                                    *reinterpret_cast<const float2*>(bias_sp + co0 + g * 8 + 2 * e);
                                const float b0 = b2.x, b1 = b2.y;
                                q[e] = pack_bf16x2(__uint_as_float(r0[g * 8 + 2 * e]) + b0,
                                                   __uint_as_float(r0[g * 8 + 2 * e + 1]) + b1);
                                q[4 + e] = pack_bf16x2(__uint_as_float(r1[g * 8 + 2 * e]) + b0,
                                                       __uint_as_float(r1[g * 8 + 2 * e + 1]) + b1);
                            }
                            if (valid) st_global_256(optr + g * plane, q);
                        }
                    }
                }
                for (int c0 = c_begin; EPI != EPI_CONVT && c0 < c_end; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld32(tcol + c0, r);
                    tmem_ld_wait();
                    const int co0 = pass * p.N + c0;  // first output channel of these columns
                    float v[32];   // + bias; ReLU in fp32 for the head only, else in the conversion
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(bias_sp + co0 + 4 * c4);
                        v[4 * c4 + 0] = __uint_as_float(r[4 * c4 + 0]) + b4.x;
                        v[4 * c4 + 1] = __uint_as_float(r[4 * c4 + 1]) + b4.y;
                        v[4 * c4 + 2] = __uint_as_float(r[4 * c4 + 2]) + b4.z;
                        v[4 * c4 + 3] = __uint_as_float(r[4 * c4 + 3]) + b4.w;
                        if (EPI == EPI_HEAD) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) v[4 * c4 + e] = fmaxf(v[4 * c4 + e], 0.f);
                        }
                    }
                    if (EPI == EPI_HEAD) {
#pragma unroll
                        for (int c4 = 0; c4 < 8; ++c4) {
                            const float4 h4 =
                                *reinterpret_cast<const float4*>(bias_sp + p.cout + c0 + 4 * c4);
                            zacc = fmaf(v[4 * c4 + 0], h4.x, zacc);
                            zacc = fmaf(v[4 * c4 + 1], h4.y, zacc);
                            zacc = fmaf(v[4 * c4 + 2], h4.z, zacc);
                            zacc = fmaf(v[4 * c4 + 3], h4.w, zacc);
                        }
                    } else {
                        const size_t plane = static_cast<size_t>(OH) * OW * 8;
                        // position inside an 8-channel plane: row-major, or space-to-depth (the four
                        // pixel phases as four quarter planes)
                        const size_t pix =
                            p.out_s2d ? (static_cast<size_t>((y & 1) * 2 + (x & 1)) * (OH >> 1) * (OW >> 1) +
                                         static_cast<size_t>(y >> 1) * (OW >> 1) + (x >> 1))
                                      : static_cast<size_t>(y) * OW + x;
                        __nv_bfloat16* optr =
                            p.out + (static_cast<size_t>(t.n) * (p.cout >> 3) + (co0 >> 3)) * plane + pix * 8;
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            uint4 q4;
                            q4.x = pack_relu_bf16x2(v[g * 8 + 0], v[g * 8 + 1]);
                            q4.y = pack_relu_bf16x2(v[g * 8 + 2], v[g * 8 + 3]);
                            q4.z = pack_relu_bf16x2(v[g * 8 + 4], v[g * 8 + 5]);
                            q4.w = pack_relu_bf16x2(v[g * 8 + 6], v[g * 8 + 7]);
                            if (valid) *reinterpret_cast<uint4*>(optr + g * plane) = q4;
                            if (EPI == EPI_RELU_POOL) {
                                // 2x2 max over (x^1, y^1): lanes ^1 and ^8 of this warp
                                uint4 m4;
                                m4.x = max_bf16x2(q4.x, __shfl_xor_sync(0xffffffffu, q4.x, 1));
                                m4.y = max_bf16x2(q4.y, __shfl_xor_sync(0xffffffffu, q4.y, 1));
                                m4.z = max_bf16x2(q4.z, __shfl_xor_sync(0xffffffffu, q4.z, 1));
                                m4.w = max_bf16x2(q4.w, __shfl_xor_sync(0xffffffffu, q4.w, 1));
                                m4.x = max_bf16x2(m4.x, __shfl_xor_sync(0xffffffffu, m4.x, 8));
                                m4.y = max_bf16x2(m4.y, __shfl_xor_sync(0xffffffffu, m4.y, 8));
                                m4.z = max_bf16x2(m4.z, __shfl_xor_sync(0xffffffffu, m4.z, 8));
                                m4.w = max_bf16x2(m4.w, __shfl_xor_sync(0xffffffffu, m4.w, 8));
                                if (valid && ((px & 1) == 0) && ((py & 1) == 0)) {
                                    const int PH = p.H >> 1, PW = p.W >> 1;
                                    const size_t pplane = static_cast<size_t>(PH) * PW * 8;
                                    __nv_bfloat16* pp =
                                        p.out_pool +
                                        (static_cast<size_t>(t.n) * (p.cout >> 3) + (co0 >> 3) + g) *
                                            pplane +
                                        (static_cast<size_t>(y >> 1) * PW + (x >> 1)) * 8;
                                    *reinterpret_cast<uint4*>(pp) = m4;
                                }
                            }
                        }
                    }
                }
                if (EPI == EPI_HEAD && c_end > 0) {
                    const float z = zacc + p.head_b;
                    const bool on = valid && (z > p.logit_thr);
                    const size_t pix = (static_cast<size_t>(t.n) * p.H + y) * p.W + x;
                    if (valid && p.logits) p.logits[pix] = z;
                    if (valid && p.mask) p.mask[pix] = on ? 255 : 0;
                    const uint32_t bal = __ballot_sync(0xffffffffu, on);
                    if (lane == 0 && p.area && bal) atomicAdd(p.area + t.n, __popc(bal));
                }
                if (kSplit) {
                    tc_fence_before();
                    release_acc(acc_empty + 8u * mt);
                }
            }
            if (!kSplit) {
                tc_fence_before();
                release_acc(acc_empty + 8u * (p.acc_half ? buf * 2u + static_cast<uint32_t>(egrp) : buf));
            }
        }
    }

    // ------------------------------------------------------------- teardown
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();  // neither CTA leaves while the pair may still touch it
    if (warp == 2) {
        tc_fence_after();
        if (CG == 2) tmem_dealloc_pair(tmem_base, kTmemCols);
        else tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
constexpr int kMaxSmem = 227 * 1024;

// Tensor map over a C8-planar activation tensor [B][C/8][H][W][8] bf16:
// dims (innermost first) = (W*8, H, C/8, B); box = (edge*8, edge, 4, 1), edge = 18 (conv3x3 with
// its halo) or 16 (transposed conv).
int make_act_map(CUtensorMap* tm, const __nv_bfloat16* base, int B, int C, int H, int W,
                 int edge = kHalo) {
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(W) * 8, static_cast<cuuint64_t>(H),
                          static_cast<cuuint64_t>(C / 8), static_cast<cuuint64_t>(B)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(W) * 16,
                             static_cast<cuuint64_t>(W) * 16 * H,
                             static_cast<cuuint64_t>(W) * 16 * H * (C / 8)};
    cuuint32_t box[4] = {static_cast<cuuint32_t>(edge) * 8, static_cast<cuuint32_t>(edge), 4, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                          const_cast<__nv_bfloat16*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[160];
        snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d) for B=%d C=%d H=%d W=%d",
                 static_cast<int>(r), B, C, H, W);
        return fail(buf);
    }
    return 0;
}

template <int EPI, int TPS>
int launch_epi(const CUtensorMap& tm0, const CUtensorMap& tm1, const ConvParams& p, int grid,
               size_t smem, cudaStream_t stream) {
    if (p.S == 1)
        conv_tc_kernel<EPI, TPS, 1, 1><<<grid, kThreads, smem, stream>>>(tm0, tm1, tm0, p);
    else
        conv_tc_kernel<EPI, TPS, 2, 1><<<grid, kThreads, smem, stream>>>(tm0, tm1, tm0, p);
    OGL_CUDA(cudaGetLastError());
    return 0;
}
// CTA-pair form (cta_group::2): clusters of 2 CTAs, S sub-tiles per CTA
template <int EPI, int TPS, int S = 2>
int launch_epi_pair(const CUtensorMap& tm0, const CUtensorMap& tm1, const CUtensorMap& tmw,
                    const ConvParams& p, int grid, size_t smem, cudaStream_t stream) {
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
    OGL_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<EPI, TPS, S, 2>, tm0, tm1, tmw, p));
    return 0;
}
template <int EPI, int TPS>
int set_smem_attr() {
    OGL_CUDA(cudaFuncSetAttribute(conv_tc_kernel<EPI, TPS, 1, 1>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    OGL_CUDA(cudaFuncSetAttribute(conv_tc_kernel<EPI, TPS, 2, 1>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    return 0;
}
template <int EPI, int TPS>
int set_smem_attr_pair() {
    OGL_CUDA(cudaFuncSetAttribute(conv_tc_kernel<EPI, TPS, 2, 2>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    return 0;
}
inline int taps_per_stage(const TcLayer& L) { return L.taps == 1 ? 1 : (L.N <= 64 ? 9 : 3); }

}  // namespace

#ifndef OGL_F16
int encode_bf16_map(void* tensor_map, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box) {
    return encode_map(tensor_map, base, rank, dims, strides_bytes, box, false);
}

int encode_map(void* tensor_map, const void* base, int rank, const uint64_t* dims,
               const uint64_t* strides_bytes, const uint32_t* box, bool u8) {
    if (!g_encode) return fail("conv_tc_init() was not called");
    cuuint64_t d[5], st[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        d[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        if (i < rank - 1) st[i] = strides_bytes[i];
    }
    CUresult r = g_encode(static_cast<CUtensorMap*>(tensor_map),
                          u8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                          static_cast<cuuint32_t>(rank), const_cast<void*>(base), d, st, bx, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[160];
        snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d) for a rank-%d map",
                 static_cast<int>(r), rank);
        return fail(buf);
    }
    return 0;
}
#endif  // !OGL_F16

int conv_tc_init() {
    if (!g_encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        OGL_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess)
            return fail("cuTensorMapEncodeTiled entry point not available");
        g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    OGL_CUDA(set_wait_cfg());
    if (set_smem_attr<EPI_RELU, 9>() || set_smem_attr<EPI_RELU, 3>() ||
        set_smem_attr<EPI_RELU_POOL, 9>() || set_smem_attr<EPI_RELU_POOL, 3>() ||
        set_smem_attr<EPI_HEAD, 9>() || set_smem_attr<EPI_CONVT, 1>() ||
        set_smem_attr_pair<EPI_RELU, 9>() || set_smem_attr_pair<EPI_RELU, 3>() ||
        set_smem_attr_pair<EPI_RELU_POOL, 9>() || set_smem_attr_pair<EPI_RELU_POOL, 3>())
        return 1;
    OGL_CUDA(cudaFuncSetAttribute(conv_tc_kernel<EPI_CONVT, 1, 1, 2>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    return 0;
}

int launch_conv_tc(const TcLayer& L, const __nv_bfloat16* src0, const __nv_bfloat16* src1, int B,
                   int H, int W, __nv_bfloat16* out, __nv_bfloat16* out_pool,
                   const HeadParams* head, int num_sms, cudaStream_t stream, int cta_group,
                   bool reverse, bool out_s2d) {
    if (!g_encode) return fail("conv_tc_init() was not called");
    if (out_s2d && (L.epi == EPI_CONVT || L.epi == EPI_HEAD || H % 2 || W % 2))
        return fail("space-to-depth output needs a conv3x3 layer with even H and W");
    if (H < 1 || W < 1) return fail("tensor-core conv needs a non-empty feature map");
    if (L.epi == EPI_RELU_POOL && (H % 2 || W % 2))
        return fail("fused 2x2 max-pool needs even H and W");
    if (L.cin0 % 32 || L.cin1 % 32 || (L.N != 32 && L.N != 64 && L.N != 128))
        return fail("tensor-core conv needs Cin % 32 == 0 and N in {32,64,128}");
    if (L.epi == EPI_CONVT && (L.taps != 1 || L.cout % (L.N / 2)))
        return fail("convT layer: N must be 2 * (a divisor block of Cout)");
    if (L.epi == EPI_RELU_POOL && !out_pool) return fail("pool epilogue needs out_pool");
    if (L.epi == EPI_HEAD && (!head || L.cout != 32 || L.npass != 1))
        return fail("head epilogue needs Cout == 32 in one pass");

    ConvParams p;
    memset(&p, 0, sizeof p);
    p.wpack = L.wpack;
    p.bias = L.bias;
    p.out = out;
    p.out_pool = out_pool;
    if (head) {
        p.head_w = head->w;
        p.head_b = head->b;
        p.logit_thr = head->logit_thr;
        p.logits = head->logits;
        p.mask = head->mask;
        p.area = head->area;
    }
    p.kb0 = L.cin0 / 32;
    p.kb1 = L.cin1 / 32;
    p.taps = L.taps;
    p.N = L.N;
    p.npass = L.npass;
    p.cout = L.cout;
    p.H = H;
    p.W = W;
    p.B = B;
    // Tile order (experiment, api.cu OGL_PINGPONG): a launch that starts with the frames its
    // producer wrote last could find them still in the 126 MB L2; measured, no gain at batch 512.
    p.reverse = reverse ? 1 : 0;
    p.out_s2d = out_s2d ? 1 : 0;
    // 2 sub-tiles (4 accumulators) per tile. Measured alternatives (experiment switches):
    // 1 sub-tile for N = 128 (OGL_S128=1) or for the transposed conv (OGL_ST=1) doubles the
    // weight traffic per pixel and is slower although TMEM could then be double-buffered.
    static const int s_env = env_knob("OGL_S128", 2, 1, 2);
    static const int st_env = env_knob("OGL_ST", 2, 1, 2);
    p.S = (L.taps == 1 && L.N == 128) ? st_env : ((L.N == 128) ? s_env : 2);
    p.tiles_x = (W + 15) / 16;
    p.tiles_y = (H + 15) / 16;
    p.total_sub = B * p.tiles_x * p.tiles_y;
    // Transposed conv on CTA pairs (OGL_CONVT_PAIR=1, experiment): ONE sub-tile per CTA, so that
    // TMEM holds two accumulator sets and the pixel-shuffle epilogue of a tile overlaps the MMAs
    // of the next one, while the pair still fetches each weight column once per 512 pixels.
    static const int convt_pair_env = env_knob("OGL_CONVT_PAIR", 0, 0, 1);
    const bool convt_pair = convt_pair_env && L.epi == EPI_CONVT && L.N == 128 && L.wpack2 &&
                            cta_group >= 2 && num_sms >= 2 &&
                            p.total_sub >= (cta_group == 3 ? 2 : num_sms);
    if (convt_pair) p.S = 1;
    p.num_tiles = (p.total_sub + p.S - 1) / p.S;
    p.magic_tx = ((1ull << 40) / static_cast<unsigned long long>(p.tiles_x)) + 1;
    p.magic_tpf = ((1ull << 40) / static_cast<unsigned long long>(p.tiles_x * p.tiles_y)) + 1;
    if (static_cast<unsigned long long>(p.total_sub) * (p.tiles_x * p.tiles_y) >= (1ull << 40))
        return fail("batch too large for the tile decoder");
    p.acc_bufs = (2 * p.S * p.N <= 256) ? 2 : 1;
    static const int split_env = env_knob("OGL_SPLIT", 1, 0, 1);
    p.split = split_env;
    static const int dbg_env = experiment_dbg();
    p.dbg = dbg_env;
    static const int half_env = env_knob("OGL_ACC_HALF", 1, 0, 1);
    p.acc_half = half_env;
    const int tps = taps_per_stage(L);
    // CTA pairs (cta_group::2) for the conv3x3 layers with N >= 64 when there is enough work
    // (cta_group 2), or whenever possible (cta_group 3: unit tests on small inputs)
    // Layers with a single 32-channel K block (downs.1.net.0) are faster unpaired (measured):
    // the pair's per-tile hand-shakes are not amortised over so short a K loop.
    static const int minkb_env = env_knob("OGL_CG_MINKB", 2, 1, 64);
    const bool pair = convt_pair ||
                      (cta_group >= 2 && L.wpack2 && L.taps == 9 && (L.N == 64 || L.N == 128) &&
                       (L.epi == EPI_RELU || L.epi == EPI_RELU_POOL) && p.S == 2 &&
                       num_sms >= 2 && p.num_tiles >= (cta_group == 3 ? 2 : num_sms) &&
                       (cta_group == 3 || p.kb0 + p.kb1 >= minkb_env));
    const size_t w_stage = static_cast<size_t>(tps) * 64u * p.N / (pair ? 2 : 1);
    const size_t tables = sizeof(float) * (L.cout + 32);
    const int edge = L.epi == EPI_CONVT ? 16 : kHalo;
    const int sub_bytes = 4 * edge * edge * 16;
    auto smem_need = [&](int na, int nw) {
        return static_cast<size_t>(128 /*align slack*/ + na * p.S * sub_bytes + nw * w_stage +
                                   16 * (na + nw) + 80 + tables + 64);
    };
    // ring depths: as many weight stages as fit beside 3 activation stages (2 when the
    // weight stages are large), at least 2 and at most 8
    p.na = 3;
    p.nw = 8;
    while (p.nw > 2 && smem_need(p.na, p.nw) > static_cast<size_t>(kMaxSmem)) --p.nw;
    if (p.nw < 3 && w_stage > 30000) {
        p.na = 2;
        p.nw = 3;
    }
    // transposed conv: a K block is only 2 x 16 KB and the two halves of a tile drift apart by at
    // most the ring's depth (acc_half), so its ring may be deeper (OGL_NA_CONVT)
    // (measured: ups.4 0.39 -> 0.36 ms with 4..6 stages)
    static const int na_convt_env = env_knob("OGL_NA_CONVT", 5, 2, 8);
    if (L.epi == EPI_CONVT && na_convt_env > 0) p.na = na_convt_env;
    static const int na_env = env_knob("OGL_NA", 0, 2, 8);   // 0: automatic
    static const int nw_env = env_knob("OGL_NW", 0, 2, 8);
    if (na_env > 0) p.na = na_env;
    if (nw_env > 0) p.nw = nw_env;
    while (p.nw > 2 && smem_need(p.na, p.nw) > static_cast<size_t>(kMaxSmem)) --p.nw;
    const size_t smem = smem_need(p.na, p.nw);
    if (smem > static_cast<size_t>(kMaxSmem)) return fail("shared memory budget exceeded");
    if (tps == 3 && p.S == 2 && p.nw < 3)
        return fail("N = 128 layers hold three weight stages at once: nw must be >= 3");

    CUtensorMap tm0, tm1;
    if (make_act_map(&tm0, src0, B, L.cin0, H, W, edge)) return 1;
    if (L.cin1 > 0) {
        if (make_act_map(&tm1, src1, B, L.cin1, H, W, edge)) return 1;
    } else {
        tm1 = tm0;
    }
    if (pair) {
        // weights as rows of 256 bytes; one box = this CTA's half of a stage
        CUtensorMap tmw;
        const uint64_t rows = static_cast<uint64_t>(L.npass) * (p.kb0 + p.kb1) * L.taps * 64u * p.N / 256;
        const uint64_t dims[2] = {128, rows};
        const uint64_t str[1] = {256};
        const uint32_t box[2] = {128, static_cast<uint32_t>(w_stage / 256)};
        if (encode_bf16_map(&tmw, L.wpack2, 2, dims, str, box)) return 1;
        p.pass_fast = convt_pair ? 1 : 0;   // transposed conv: the passes of a tile back to back,
                                            // so that their re-reads of the tile hit L2
        const int grid2 = num_sms & ~1;
        if (convt_pair) return launch_epi_pair<EPI_CONVT, 1, 1>(tm0, tm1, tmw, p, grid2, smem, stream);
        if (L.epi == EPI_RELU)
            return tps == 9 ? launch_epi_pair<EPI_RELU, 9>(tm0, tm1, tmw, p, grid2, smem, stream)
                            : launch_epi_pair<EPI_RELU, 3>(tm0, tm1, tmw, p, grid2, smem, stream);
        return tps == 9 ? launch_epi_pair<EPI_RELU_POOL, 9>(tm0, tm1, tmw, p, grid2, smem, stream)
                        : launch_epi_pair<EPI_RELU_POOL, 3>(tm0, tm1, tmw, p, grid2, smem, stream);
    }
    const int items = p.npass * p.num_tiles;
    const int grid = items < num_sms ? items : num_sms;
    static const int pf_env = env_knob("OGL_PASSFAST", 1, 0, 1);
    p.pass_fast = (pf_env && p.npass > 1 && grid % p.npass == 0) ? 1 : 0;
    if (L.epi == EPI_CONVT) return launch_epi<EPI_CONVT, 1>(tm0, tm1, p, grid, smem, stream);
    if (L.epi == EPI_HEAD) return launch_epi<EPI_HEAD, 9>(tm0, tm1, p, grid, smem, stream);
    if (L.epi == EPI_RELU)
        return tps == 9 ? launch_epi<EPI_RELU, 9>(tm0, tm1, p, grid, smem, stream)
                        : launch_epi<EPI_RELU, 3>(tm0, tm1, p, grid, smem, stream);
    if (L.epi == EPI_RELU_POOL)
        return tps == 9 ? launch_epi<EPI_RELU_POOL, 9>(tm0, tm1, p, grid, smem, stream)
                        : launch_epi<EPI_RELU_POOL, 3>(tm0, tm1, p, grid, smem, stream);
    return fail("unknown epilogue");
}

}  // namespace ogl
