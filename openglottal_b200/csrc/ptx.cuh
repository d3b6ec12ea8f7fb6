// Thin inline-PTX wrappers for the sm_100a features the conv kernels use:
// mbarrier, TMA (cp.async.bulk[.tensor]), tcgen05 (TMEM alloc / mma / commit / ld).
// Nothing here is generic: every wrapper is the exact form the kernels need.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

// Operand type of the tensor-core path, fixed per translation unit: the *_f16.cu units define
// OGL_F16 and include the bf16 source, so that every kernel exists once per operand type with the
// conversions resolved at compile time. tcgen05.mma.kind::f16 takes either type at the same rate;
// f16 has 3 more mantissa bits (logits within 2e-2 of the fp32 reference everywhere), bf16 the
// range of fp32 (the default; BASELINE.json's metric is quoted in bf16).
#ifdef OGL_F16
#define OGL_AB_FMT 0u
#else
#define OGL_AB_FMT 1u
#endif

namespace ogl {

#ifdef OGL_F16
constexpr bool kF16 = true;
#else
constexpr bool kF16 = false;
#endif

// Two f32 -> one packed pair of 16-bit operands {lo, hi}, round to nearest even. f16 saturates to
// +-65504 instead of overflowing to infinity (no NaN can appear downstream).
template <bool F16>
__device__ __forceinline__ uint32_t pack_x2(float lo, float hi) {
    uint32_t d;
    if (F16) asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
// the same of max(x, 0): ReLU folded into the conversion (one F2FP instead of two FMNMX + one F2FP;
// the values are those of fmaxf followed by the plain conversion)
template <bool F16>
__device__ __forceinline__ uint32_t pack_relu_x2(float lo, float hi) {
    uint32_t d;
    if (F16) asm("cvt.rn.satfinite.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
template <bool F16>
__device__ __forceinline__ uint32_t max_x2(uint32_t a, uint32_t b) {
    uint32_t d;
    if (F16) asm("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    else asm("max.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One lane of a converged warp (the lowest active one) gets `true`; lets the compiler keep
// the surrounding code warp-uniform (tcgen05.mma operands live in uniform registers).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .b32 rx;\n"
        ".reg .pred px;\n"
        "elect.sync rx|px, 0xFFFFFFFF;\n"
        "selp.u32 %0, 1, 0, px;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// Register re-allocation between the warp groups of a CTA (all four warps of a group execute it)
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
// a * b + c on packed f16 pairs (HFMA2)
__device__ __forceinline__ uint32_t fma_f16x2(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// Position in a ring of `n` stages and the mbarrier phase parity of its current pass, advanced
// without integer division: `it % n` / `(it / n) & 1` with a run-time n are ~40 dependent
// instructions (I2F, MUFU.RCP, F2I, IMADs), which in an MMA issuer's loop sit between two groups of
// MMAs (CTA 0's timeline, scripts/stem_trace.py: ~450 cycles per barrier wait of an issuer).
struct Ring {
    uint32_t slot = 0, phase = 0;
    __device__ __forceinline__ void next(uint32_t n) {
        if (++slot == n) {
            slot = 0;
            phase ^= 1u;
        }
    }
    __device__ __forceinline__ void skip(uint32_t n, int count) {
        for (int i = 0; i < count; ++i) next(n);
    }
};

// this translation unit's operand type
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float lo, float hi) {
    return pack_relu_x2<kF16>(lo, hi);
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// How the warps wait (per translation unit, set once from the environment by set_wait_cfg()):
//   [0] nanosleep between the polls of the non-critical waits (ns; 0 = none)
//   [1] suspend-time hint of the non-critical waits' mbarrier.try_wait (ns; 0 = the default limit)
//   [2] suspend-time hint of the critical waits (the MMA issuers)
//   [3] the issuers of the fused-stem kernel wait like the non-critical warps (0 / 1)
// With the default time limit a failed try_wait comes back after ~25 cycles, so a waiting warp
// issues SYNCS + BRA (+ NANOSLEEP) at that rate: 37 % of the instructions of the fused-stem
// kernel were such polls (profiles/ncu_full_tc_r01_v9_batch128.csv, source page), taking issue
// slots from its CUDA-core stem warps and energy from everything else.
static __constant__ uint32_t c_wait_cfg[4];
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(ns)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    const uint32_t hint = c_wait_cfg[2];
    if (hint) {
        while (!mbar_try_wait_hint(bar, parity, hint)) {
        }
    } else {
        while (!mbar_try_wait(bar, parity)) {
        }
    }
}
// For the warps that are NOT on the tensor pipe's critical path (producers waiting for a free
// slot, epilogue warps waiting for an accumulator): back off between polls, so that their spinning
// does not take issue slots from the warps that have work (measured: the in-kernel stem of
// downs.0.net.3 1.48 -> 1.25 ms).
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
    const uint32_t ns = c_wait_cfg[0], hint = c_wait_cfg[1];
    if (hint) {
        while (!mbar_try_wait_hint(bar, parity, hint)) {
            if (ns) __nanosleep(ns);
        }
    } else {
        while (!mbar_try_wait(bar, parity)) {
            if (ns) __nanosleep(ns);
        }
    }
}
// Host side: OGL_WAIT_SLEEP / OGL_WAIT_HINT / OGL_WAIT_HINT_CRIT (ns), OGL_WAIT_STEM_RELAXED
// -> this unit's c_wait_cfg
inline cudaError_t set_wait_cfg() {
    auto env = [](const char* name, uint32_t dflt) {
        const char* v = getenv(name);
        return v ? static_cast<uint32_t>(atoi(v)) : dflt;
    };
    const uint32_t cfg[4] = {env("OGL_WAIT_SLEEP", 64), env("OGL_WAIT_HINT", 0),
                             env("OGL_WAIT_HINT_CRIT", 0), env("OGL_WAIT_STEM_RELAXED", 0)};
    return cudaMemcpyToSymbol(c_wait_cfg, cfg, sizeof cfg);
}

// --------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
// 4-D tiled load, completion reported as bytes on an mbarrier.
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar,
                                            int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
        "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar,
                                            int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* tm, uint32_t bar,
                                            int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
        "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
// 1-D bulk copy global -> shared (contiguous, size multiple of 16 B).
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes,
                                          uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(dst),
        "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}

// ----------------------------------------------------------------- tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_slot),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> f32, issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrives once every MMA issued so far by this thread has completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::
                     "r"(bar)
                 : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp gets TMEM lane (base+i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 32-byte global store (one full sector per thread).
__device__ __forceinline__ void st_global_256(void* ptr, const uint32_t (&q)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "r"(q[0]),
                 "r"(q[1]), "r"(q[2]), "r"(q[3]), "r"(q[4]), "r"(q[5]), "r"(q[6]), "r"(q[7])
                 : "memory");
}

// ------------------------------------------------------- CTA pairs (cta_group::2)
// Two CTAs of a cluster (the two SMs of a TPC) execute one 256-row MMA: each provides its own
// 128 x 16 A tile and HALF of the N x 16 B tile; each receives its 128 rows of D in its own TMEM.
// Only the leader (cluster rank 0) issues; commits are multicast to the barriers of both CTAs.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}
// TMA loads issued by either CTA of the pair; `bar` is a shared::cluster address (the leader's
// barrier), the destination is the executing CTA's own shared memory.
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* tm, uint32_t bar,
                                                 int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
        "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d_pair(uint32_t dst, const CUtensorMap* tm, uint32_t bar,
                                                 int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
        "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* tm, uint32_t bar,
                                                 int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(tm), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_slot),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
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
// arrives (once all MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
        "[%0], %1;" ::"r"(bar),
        "h"(static_cast<uint16_t>(3))
        : "memory");
}

// Shared-memory matrix descriptor, K-major, SWIZZLE_NONE ("interleave") canonical layout:
// in 16-byte units ((8,n),2):((1,SBO),LBO) -- 8 rows x 16 B are one contiguous 128 B core
// matrix, SBO = byte step between 8-row groups, LBO = byte step between the two 8-element
// K halves of one K=16 MMA. Bits [46,48) = 1 is the sm_100 descriptor version.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    return d;
}
// Instruction descriptor for kind::f16: D = f32, A and B of format `fmt` (0 = f16, 1 = bf16), both
// K-major, M = 128 (256 for a CTA pair: 128 rows in each CTA), N runtime.
__host__ __device__ __forceinline__ uint32_t make_idesc_ab(int n, uint32_t a_fmt, uint32_t b_fmt,
                                                           int m = 128) {
    return (1u << 4) | (a_fmt << 7) | (b_fmt << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
           (static_cast<uint32_t>(m >> 4) << 24);
}
__host__ __device__ __forceinline__ uint32_t make_idesc_fmt(int n, uint32_t fmt, int m = 128) {
    return make_idesc_ab(n, fmt, fmt, m);
}
// ... with this translation unit's operand type
__host__ __device__ __forceinline__ uint32_t make_idesc_bf16(int n) {
    return make_idesc_fmt(n, OGL_AB_FMT);
}
__host__ __device__ __forceinline__ uint32_t make_idesc_bf16_pair(int n) {
    return make_idesc_fmt(n, OGL_AB_FMT, 256);
}

}  // namespace ogl
