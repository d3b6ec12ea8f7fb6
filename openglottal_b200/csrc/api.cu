// C ABI of libopenglottal_b200.so (see include/openglottal_b200.h): handle management,
// BatchNorm folding + operand packing, and the layer schedule of the U-Net forward
// (/root/reference/openglottal/models/unet.py:74-88).
#include "../../include/openglottal_b200.h"
#include "internal.h"

#include <cuda_fp16.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <nvtx3/nvToolsExt.h>   // header-only: ranges cost nothing unless a tool is attached
#include <string>
#include <vector>

namespace ogl {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
int fail(const std::string& msg) {
    g_err = msg;
    return 1;
}
int fail_cuda(cudaError_t e, const char* what) {
    g_err = std::string("CUDA error: ") + cudaGetErrorString(e) + " at " + what;
    return 1;
}
int launch_features_f64(const double* area, int64_t n, double* out8, int32_t* flags2, void* ws,
                        size_t ws_bytes, cudaStream_t stream);

namespace {

constexpr int kFeat[4] = {32, 64, 128, 256};

struct F32Conv {   // fp32 folded conv3x3 (validation path), PyTorch layout
    float* w = nullptr;  // [cout][cin][9]
    float* b = nullptr;  // [cout]
    int cin = 0, cout = 0;
};
struct F32ConvT {
    float* w = nullptr;  // [cin][cout][2][2]
    float* b = nullptr;
    int cin = 0, cout = 0;
};

__global__ void bgr_to_gray_kernel(const uint8_t* __restrict__ bgr, uint8_t* __restrict__ gray,
                                   long long pixels) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < pixels;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int b = bgr[3 * i], g = bgr[3 * i + 1], r = bgr[3 * i + 2];
        gray[i] = static_cast<uint8_t>((3735 * b + 19235 * g + 9798 * r + 16384) >> 15);
    }
}

// 16 pixels per thread: three 16-byte loads, one 16-byte store (both pointers 16-byte aligned);
// the same integer expression per pixel.
__global__ void __launch_bounds__(256)
bgr_to_gray_vec16_kernel(const uint4* __restrict__ bgr, uint4* __restrict__ gray, long long groups) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < groups;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const uint4 v0 = __ldg(bgr + 3 * i), v1 = __ldg(bgr + 3 * i + 1), v2 = __ldg(bgr + 3 * i + 2);
        const uint32_t w[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t acc = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = 3 * (4 * q + j);                 // byte offset of the pixel's B
                const uint32_t b = (w[k >> 2] >> (8 * (k & 3))) & 0xffu;
                const uint32_t g = (w[(k + 1) >> 2] >> (8 * ((k + 1) & 3))) & 0xffu;
                const uint32_t r = (w[(k + 2) >> 2] >> (8 * ((k + 2) & 3))) & 0xffu;
                acc |= ((3735u * b + 19235u * g + 9798u * r + 16384u) >> 15) << (8 * j);
            }
            o[q] = acc;
        }
        gray[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

}  // namespace
}  // namespace ogl

using namespace ogl;

struct ogl_unet {
    int device = 0;
    int num_sms = 148;
    bool loaded = false;
    // tensor-core path: one set of packed operands per operand type (0 = bf16, 1 = f16; the f16
    // set is built on demand by ogl_unet_prepare from the folded host weights)
    StemWeights stem;         // folded fp32 [32][9] + [32], passed by value to the stem kernel
    struct Pack {
        bool built = false;
        TcLayer down_c2[4];       // downs.i.net.3 (+pool)
        TcLayer down_c1[4];       // downs.i.net.0 for i = 1..3 (index 0 unused: stem)
        TcLayer bott[2];
        TcLayer up_t[4];
        TcLayer up_c[4][2];
        // full-resolution level as space-to-depth GEMMs (s2d_tc.cu): downs.0.net.3 (+pool),
        // ups.6 (convT) composed into ups.7.net.0, ups.7.net.3 (+head)
        S2dLayer s2d_down, s2d_up0, s2d_up1;
        // levels 3, 2, 1: ups.{0,2,4} (convT) composed into ups.{1,3,5}.net.0 (upcat_tc.cu)
        UpcatLayer upcat[3];
    } pack[2];
    struct Folded {           // BN-folded fp32 weights on the host (PyTorch layouts)
        std::vector<float> dw[4][2], db[4][2], bw[2], bb[2], tw[4], tb[4], uw[4][2], ub[4][2];
    } folded;
    float* head_w = nullptr;  // [32] device
    float head_w_host[32] = {0};
    float head_b = 0.f;
    bool use_s2d = true;
    // ConvTranspose2d composed into the conv that follows it at levels 1-3 as well (upcat_tc.cu):
    // 17 launches instead of 20, the `up` tensors never exist. 0: three transposed-conv launches +
    // two-source convs (the round-1 schedule; kept for A/B runs and as an independent implementation)
    bool compose_up = true;
    // u8 input: compute the stem inside the downs.0.net.3 kernel (its output never touches HBM:
    // 8.4 MB per frame less traffic). Mode 1 (fp32 on the CUDA cores) is bit-identical to the
    // stand-alone stem; modes 2 and 3 (a K = 16 GEMM on the tensor cores, weights split hi + lo)
    // agree with it to ~2^-17 relative before the rounding to bf16, i.e. NOT bit for bit.
    int fuse_stem = 3;                 // 0 separate kernel, 1 in-kernel on CUDA cores, 2 on tensor cores
                                       // (8 stem warps, bf16 im2col), 3 the same with 16 warps / f16 im2col
    uint8_t* stem_tc = nullptr;        // B operands of the tensor-core stem (device)
    uint8_t* stem_tc3 = nullptr;       // the same in the K order of the 16-warp form
    int cta_group = 2;  // 2: conv3x3 layers with N >= 64 run on CTA pairs (tcgen05 cta_group::2)
    // fp32 validation path
    F32Conv f_down[4][2], f_bott[2], f_up[4][2];
    F32ConvT f_upt[4];
    std::vector<void*> allocs;
    // optional per-launch timing (ogl_unet_set_profiling)
    // Events between the launches of the most recent kProfSets profiled forwards, so that a
    // caller can time its own steady-state loop and read the per-launch averages afterwards.
    bool profile = false;
    std::vector<cudaEvent_t> events;   // kProfSets x kMaxLaunches
    int n_events = 0;                  // events recorded by the forward in progress
    int prof_set = 0;                  // set the forward in progress writes
    int prof_forwards = 0;             // profiled forwards since profiling was enabled
    int prof_count[16] = {0};          // events recorded in each set
    std::vector<const char*> launch_names;  // of the most recent bf16 forward
    // measurement aid (ogl_unet_set_repeat): launch `rep_launch` of the bf16 forward runs rep_n times
    int rep_launch = -1, rep_n = 1;
};

namespace {

template <typename T>
int dev_upload(ogl_unet* h, const std::vector<T>& host, T** out) {
    void* p = nullptr;
    OGL_CUDA(cudaMalloc(&p, host.size() * sizeof(T)));
    h->allocs.push_back(p);
    OGL_CUDA(cudaMemcpy(p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = static_cast<T*>(p);
    return 0;
}

void free_all(ogl_unet* h) {
    for (void* p : h->allocs) cudaFree(p);
    h->allocs.clear();
    h->loaded = false;
}

// eval-mode BatchNorm folded into the preceding bias-free conv, in fp64:
//   s = gamma / sqrt(var + eps);  W' = W * s[co];  b' = beta - mean * s
void fold_conv_bn(const ogl_conv_bn& L, int cout, int cin, double eps, std::vector<float>* w,
                  std::vector<float>* b) {
    w->resize(static_cast<size_t>(cout) * cin * 9);
    b->resize(cout);
    for (int co = 0; co < cout; ++co) {
        const double s = static_cast<double>(L.bn_weight[co]) /
                         std::sqrt(static_cast<double>(L.running_var[co]) + eps);
        (*b)[co] = static_cast<float>(static_cast<double>(L.bn_bias[co]) -
                                      static_cast<double>(L.running_mean[co]) * s);
        const size_t base = static_cast<size_t>(co) * cin * 9;
        for (size_t k = 0; k < static_cast<size_t>(cin) * 9; ++k)
            (*w)[base + k] = static_cast<float>(static_cast<double>(L.weight[base + k]) * s);
    }
}

// one 16-bit tensor-core operand: bf16, or f16 saturated to +-65504 (no folded weight of a
// trained network comes near it)
__nv_bfloat16 to_op(float v, bool f16) {
    __nv_bfloat16 out;
    if (f16) {
        v = v > 65504.f ? 65504.f : (v < -65504.f ? -65504.f : v);
        const __half h = __float2half_rn(v);
        memcpy(&out, &h, 2);
    } else {
        out = __float2bfloat16_rn(v);
    }
    return out;
}

// conv3x3 weights [cout][cin][3][3] -> [pass][cin/32][tap][4][N][8] bf16: one contiguous
// 64*N-byte block per (32-channel block, tap), taps of a block adjacent so several can be
// fetched with one bulk copy.
std::vector<__nv_bfloat16> pack_conv(const std::vector<float>& w, int cout, int cin, int N,
                                     bool f16) {
    const int npass = cout / N, kb = cin / 32;
    std::vector<__nv_bfloat16> out(static_cast<size_t>(cout) * cin * 9);
    for (int pass = 0; pass < npass; ++pass)
        for (int b = 0; b < kb; ++b)
            for (int tap = 0; tap < 9; ++tap)
                for (int c = 0; c < 4; ++c)
                    for (int n = 0; n < N; ++n)
                        for (int e = 0; e < 8; ++e) {
                            const int co = pass * N + n, ci = b * 32 + c * 8 + e;
                            const size_t dst =
                                (((((static_cast<size_t>(pass) * kb + b) * 9 + tap) * 4 + c) * N + n) * 8) + e;
                            out[dst] = to_op(w[(static_cast<size_t>(co) * cin + ci) * 9 + tap], f16);
                        }
    return out;
}

// The same weights for a CTA pair (cta_group::2): rank r of the pair stages output channels
// [r*N/2, (r+1)*N/2) of every pass -- [pass][cin/32][rank][tap][4][N/2][8] bf16.
std::vector<__nv_bfloat16> pack_conv_pair(const std::vector<float>& w, int cout, int cin, int N,
                                          bool f16) {
    const int npass = cout / N, kb = cin / 32, nh = N / 2;
    std::vector<__nv_bfloat16> out(static_cast<size_t>(cout) * cin * 9);
    for (int pass = 0; pass < npass; ++pass)
        for (int b = 0; b < kb; ++b)
            for (int r = 0; r < 2; ++r)
                for (int tap = 0; tap < 9; ++tap)
                    for (int c = 0; c < 4; ++c)
                        for (int n = 0; n < nh; ++n)
                            for (int e = 0; e < 8; ++e) {
                                const int co = pass * N + r * nh + n, ci = b * 32 + c * 8 + e;
                                const size_t dst =
                                    ((((((static_cast<size_t>(pass) * kb + b) * 2 + r) * 9 + tap) * 4 + c) * nh + n) * 8) + e;
                                out[dst] = to_op(w[(static_cast<size_t>(co) * cin + ci) * 9 + tap], f16);
                            }
    return out;
}

// ConvTranspose2d weights [cin][cout][2][2]: pass = (dy, block of cb = N/2 output channels),
// GEMM column n of a pass = dx * cb + (co - blk*cb); layout [pass][cin/32][4][N][8] bf16.
int convt_n(int cout) { return 2 * (cout < 64 ? cout : 64); }
std::vector<__nv_bfloat16> pack_convt(const float* w, int cin, int cout, int N, bool f16) {
    const int cb = N / 2, nblk = cout / cb, npass = 2 * nblk, kb = cin / 32;
    std::vector<__nv_bfloat16> out(static_cast<size_t>(4) * cout * cin);
    for (int pass = 0; pass < npass; ++pass)
        for (int b = 0; b < kb; ++b)
            for (int c = 0; c < 4; ++c)
                for (int n = 0; n < N; ++n)
                    for (int e = 0; e < 8; ++e) {
                        const int dy = pass / nblk, blk = pass % nblk;
                        const int dx = n / cb, co = blk * cb + n % cb, ci = b * 32 + c * 8 + e;
                        const size_t dst =
                            ((((static_cast<size_t>(pass) * kb + b) * 4 + c) * N + n) * 8) + e;
                        out[dst] = to_op(w[(static_cast<size_t>(ci) * cout + co) * 4 + dy * 2 + dx], f16);
                    }
    return out;
}

// The same for a CTA pair: rank r stages columns [r*N/2, (r+1)*N/2) of every pass (that is dx = r
// when N = 2 * cb) -- [pass][cin/32][rank][4][N/2][8] bf16.
std::vector<__nv_bfloat16> pack_convt_pair(const float* w, int cin, int cout, int N, bool f16) {
    const std::vector<__nv_bfloat16> flat = pack_convt(w, cin, cout, N, f16);
    const int nh = N / 2, blocks = static_cast<int>(flat.size() / (static_cast<size_t>(4) * N * 8));
    std::vector<__nv_bfloat16> out(flat.size());
    for (int blk = 0; blk < blocks; ++blk)      // one (pass, 32-channel block)
        for (int r = 0; r < 2; ++r)
            for (int c = 0; c < 4; ++c)
                for (int n = 0; n < nh; ++n)
                    for (int e = 0; e < 8; ++e)
                        out[((((static_cast<size_t>(blk) * 2 + r) * 4 + c) * nh + n) * 8) + e] =
                            flat[(((static_cast<size_t>(blk) * 4 + c) * N + r * nh + n) * 8) + e];
    return out;
}

int build_tc_conv(ogl_unet* h, const std::vector<float>& w, const std::vector<float>& b, int cin0,
                  int cin1, int cout, int epi, TcLayer* L, bool f16) {
    L->cin0 = cin0;
    L->cin1 = cin1;
    L->cout = cout;
    L->taps = 9;
    L->N = cout < 128 ? cout : 128;
    L->npass = cout / L->N;
    L->epi = epi;
    if (dev_upload(h, pack_conv(w, cout, cin0 + cin1, L->N, f16), &L->wpack)) return 1;
    if (L->N >= 64 && (epi == EPI_RELU || epi == EPI_RELU_POOL) &&
        dev_upload(h, pack_conv_pair(w, cout, cin0 + cin1, L->N, f16), &L->wpack2))
        return 1;
    return dev_upload(h, b, &L->bias);
}

int build_s2d_layer(ogl_unet* h, const std::vector<float>& w3, const std::vector<float>& b3,
                    int cin_s, const float* wt, const float* bt, int cin_b, int epi, S2dLayer* L,
                    bool f16) {
    S2dHost hs;
    if (build_s2d_host(w3.data(), b3.data(), cin_s, wt, bt, cin_b, &hs, f16)) return 1;
    *L = S2dLayer();
    L->wbytes = static_cast<uint32_t>(hs.wblob.size());
    L->n_stages = hs.n_stages;
    for (int i = 0; i < kS2dMaxStages; ++i) {
        L->stage_src[i] = hs.stage_src[i];
        L->stage_plane0[i] = hs.stage_plane0[i];
    }
    L->cin_s = cin_s;
    L->cin_b = wt ? cin_b : 0;
    L->epi = epi;
    for (int i = 0; i < 32; ++i) L->bias_host[i] = hs.btab[4 * 32 + i];
    if (dev_upload(h, hs.wblob, &L->wblob)) return 1;
    if (dev_upload(h, hs.wblob_pair, &L->wblob2)) return 1;
    return dev_upload(h, hs.btab, &L->btab);
}

int build_upcat_layer(ogl_unet* h, const std::vector<float>& w3, const std::vector<float>& b3,
                      const std::vector<float>& wt, const std::vector<float>& bt, int f, UpcatLayer* L,
                      bool f16) {
    UpcatHost hs;
    if (build_upcat_host(w3.data(), b3.data(), wt.data(), bt.data(), f, f16, &hs)) return 1;
    *L = UpcatLayer();
    L->f = hs.f;
    L->N = hs.N;
    L->npass = hs.npass;
    auto up = [&](const std::vector<uint16_t>& v, uint8_t** out) {
        uint16_t* d = nullptr;
        if (dev_upload(h, v, &d)) return 1;
        *out = reinterpret_cast<uint8_t*>(d);
        return 0;
    };
    if (up(hs.wskip, &L->wskip) || up(hs.wskip_pair, &L->wskip2) || up(hs.wbelow, &L->wbelow) ||
        up(hs.wbelow_pair, &L->wbelow2))
        return 1;
    if (dev_upload(h, hs.btab, &L->btab)) return 1;
    std::vector<float> interior(hs.btab.begin() + 4 * f, hs.btab.begin() + 5 * f);   // class (1, 1)
    return dev_upload(h, interior, &L->bias);
}

int build_f32_conv(ogl_unet* h, const std::vector<float>& w, const std::vector<float>& b, int cin,
                   int cout, F32Conv* L) {
    L->cin = cin;
    L->cout = cout;
    if (dev_upload(h, w, &L->w)) return 1;
    return dev_upload(h, b, &L->b);
}

// Tensor-core operands of every layer from the folded host weights, for one operand type.
int build_pack(ogl_unet* h, bool f16) {
    ogl_unet::Pack& P = h->pack[f16 ? 1 : 0];
    const ogl_unet::Folded& F = h->folded;
    P = ogl_unet::Pack();
    int cin = 1;
    for (int i = 0; i < 4; ++i) {
        const int f = kFeat[i];
        if (i > 0 && build_tc_conv(h, F.dw[i][0], F.db[i][0], cin, 0, f, EPI_RELU, &P.down_c1[i], f16))
            return 1;
        if (build_tc_conv(h, F.dw[i][1], F.db[i][1], f, 0, f, EPI_RELU_POOL, &P.down_c2[i], f16))
            return 1;
        if (i == 0 && build_s2d_layer(h, F.dw[0][1], F.db[0][1], 32, nullptr, nullptr, 0,
                                      EPI_RELU_POOL, &P.s2d_down, f16))
            return 1;
        cin = f;
    }
    if (build_tc_conv(h, F.bw[0], F.bb[0], 256, 0, 512, EPI_RELU, &P.bott[0], f16)) return 1;
    if (build_tc_conv(h, F.bw[1], F.bb[1], 512, 0, 512, EPI_RELU, &P.bott[1], f16)) return 1;
    for (int k = 0; k < 4; ++k) {
        const int f = kFeat[3 - k];
        TcLayer* L = &P.up_t[k];
        L->cin0 = 2 * f;
        L->cin1 = 0;
        L->cout = f;
        L->taps = 1;
        L->N = convt_n(f);
        L->npass = 4 * f / L->N;
        L->epi = EPI_CONVT;
        if (dev_upload(h, pack_convt(F.tw[k].data(), 2 * f, f, L->N, f16), &L->wpack)) return 1;
        if (L->N == 128 &&
            dev_upload(h, pack_convt_pair(F.tw[k].data(), 2 * f, f, L->N, f16), &L->wpack2))
            return 1;
        if (dev_upload(h, F.tb[k], &L->bias)) return 1;
        if (build_tc_conv(h, F.uw[k][0], F.ub[k][0], f, f, f, EPI_RELU, &P.up_c[k][0], f16)) return 1;
        if (k < 3 && build_upcat_layer(h, F.uw[k][0], F.ub[k][0], F.tw[k], F.tb[k], f, &P.upcat[k], f16))
            return 1;
        if (k == 3 && build_s2d_layer(h, F.uw[k][0], F.ub[k][0], 32, F.tw[k].data(), F.tb[k].data(),
                                      64, EPI_RELU, &P.s2d_up0, f16))
            return 1;
        if (build_tc_conv(h, F.uw[k][1], F.ub[k][1], f, 0, f, k == 3 ? EPI_HEAD : EPI_RELU,
                          &P.up_c[k][1], f16))
            return 1;
        if (k == 3 && build_s2d_layer(h, F.uw[k][1], F.ub[k][1], 32, nullptr, nullptr, 0, EPI_HEAD,
                                      &P.s2d_up1, f16))
            return 1;
    }
    P.built = true;
    return 0;
}

int check_conv_bn(const ogl_conv_bn& L) {
    return (L.weight && L.bn_weight && L.bn_bias && L.running_mean && L.running_var) ? 0 : 1;
}

// ---- workspace plan: per level T (temp), S (skip), U (up / decoder out); P3, T4, B4 at the
// bottom. P_l (pooled level l) aliases U_{l+1}, which is only written later by the decoder.
struct Plan {
    size_t T[5], S[4], U[4], P3, B4, total;
};
Plan make_plan(int n, int H, int W, size_t elem) {
    Plan p;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t r = o;
        o += (bytes + 255) & ~static_cast<size_t>(255);
        return r;
    };
    for (int l = 0; l < 4; ++l) {
        const size_t bytes = static_cast<size_t>(n) * kFeat[l] * (H >> l) * (W >> l) * elem;
        p.T[l] = take(bytes);
        p.S[l] = take(bytes);
        p.U[l] = take(bytes);
    }
    const size_t b4 = static_cast<size_t>(n) * 512 * (H >> 4) * (W >> 4) * elem;
    p.T[4] = take(b4);
    p.B4 = take(b4);
    p.P3 = take(b4 / 2);
    p.total = o;
    return p;
}

constexpr int kMaxLaunches = 40;
constexpr int kProfSets = 16;
const char* const kDownC1[4] = {"stem", "downs.1.net.0", "downs.2.net.0", "downs.3.net.0"};
const char* const kDownC2[4] = {"downs.0.net.3+pool", "downs.1.net.3+pool", "downs.2.net.3+pool",
                                "downs.3.net.3+pool"};
const char* const kUpT[4] = {"ups.0(convT)", "ups.2(convT)", "ups.4(convT)", "ups.6(convT)"};
const char* const kUpC1[4] = {"ups.1.net.0(cat)", "ups.3.net.0(cat)", "ups.5.net.0(cat)",
                              "ups.7.net.0(cat)"};
const char* const kUpTC[3] = {"ups.0(convT)+ups.1.net.0(cat)", "ups.2(convT)+ups.3.net.0(cat)",
                              "ups.4(convT)+ups.5.net.0(cat)"};
const char* const kUpC2[4] = {"ups.1.net.3", "ups.3.net.3", "ups.5.net.3", "ups.7.net.3+head"};

// closes the launch that was just enqueued: its name, and (when profiling) an event after it
inline void mark(ogl_unet* h, cudaStream_t stream, const char* name) {
    if (name) h->launch_names.push_back(name);
    if (h->profile && h->n_events < kMaxLaunches) {
        cudaEventRecord(h->events[h->prof_set * kMaxLaunches + h->n_events++], stream);
        h->prof_count[h->prof_set] = h->n_events;
    }
}

}  // namespace

extern "C" {

int ogl_version(void) { return OGL_VERSION; }
const char* ogl_last_error(void) { return g_err.c_str(); }

int ogl_unet_create(ogl_unet** out, int device) {
    if (!out) return fail("ogl_unet_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail("no CUDA device available: openglottal_b200 has no CPU fallback");
    if (device < 0 || device >= count) return fail("ogl_unet_create: bad device index");
    OGL_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    OGL_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(std::string("openglottal_b200 is built for sm_100a (B200); found ") + prop.name);
    if (conv_tc_init() || s2d_tc_init() || upcat_tc_init() || conv_tc_init_f16() ||
        s2d_tc_init_f16() || upcat_tc_init_f16())
        return 1;
    ogl_unet* h = new ogl_unet();
    if (const char* e = getenv("OGL_S2D")) h->use_s2d = atoi(e) != 0;
    if (const char* e = getenv("OGL_CG")) h->cta_group = atoi(e);
    if (const char* e = getenv("OGL_FUSE_STEM")) h->fuse_stem = atoi(e);
    if (const char* e = getenv("OGL_COMPOSE")) h->compose_up = atoi(e) != 0;
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    *out = h;
    return 0;
}

int ogl_unet_destroy(ogl_unet* h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    free_all(h);
    for (auto& e : h->events) cudaEventDestroy(e);
    delete h;
    return 0;
}

int ogl_unet_load_state(ogl_unet* h, const ogl_unet_state* st) {
    if (!h || !st) return fail("ogl_unet_load_state: NULL argument");
    OGL_CUDA(cudaSetDevice(h->device));
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 2; ++j)
            if (check_conv_bn(st->downs[i][j]) || check_conv_bn(st->up_c[i][j]))
                return fail("ogl_unet_load_state: missing conv/bn tensor");
    if (check_conv_bn(st->bottleneck[0]) || check_conv_bn(st->bottleneck[1]) || !st->head_weight ||
        !st->head_bias)
        return fail("ogl_unet_load_state: missing tensor");
    for (int i = 0; i < 4; ++i)
        if (!st->up_t[i].weight || !st->up_t[i].bias)
            return fail("ogl_unet_load_state: missing convT tensor");
    free_all(h);
    h->pack[0] = ogl_unet::Pack();
    h->pack[1] = ogl_unet::Pack();
    const double eps = st->bn_eps;
    ogl_unet::Folded& F = h->folded;

    // ---- fold BatchNorm (fp64) into host copies; fp32 device copies for the validation path
    int cin = 1;
    for (int i = 0; i < 4; ++i) {
        const int f = kFeat[i];
        fold_conv_bn(st->downs[i][0], f, cin, eps, &F.dw[i][0], &F.db[i][0]);
        if (build_f32_conv(h, F.dw[i][0], F.db[i][0], cin, f, &h->f_down[i][0])) return 1;
        fold_conv_bn(st->downs[i][1], f, f, eps, &F.dw[i][1], &F.db[i][1]);
        if (build_f32_conv(h, F.dw[i][1], F.db[i][1], f, f, &h->f_down[i][1])) return 1;
        cin = f;
    }
    memcpy(h->stem.w, F.dw[0][0].data(), sizeof h->stem.w);
    memcpy(h->stem.b, F.db[0][0].data(), sizeof h->stem.b);
    {
        std::vector<uint8_t> blob;
        if (build_stem_tc_blob(h->stem, &blob) || dev_upload(h, blob, &h->stem_tc)) return 1;
        const bool b16 = getenv("OGL_STEM3_BFMT") && atoi(getenv("OGL_STEM3_BFMT"));   // experiment
        if (build_stem_tc_blob(h->stem, &blob, true, !b16) || dev_upload(h, blob, &h->stem_tc3)) return 1;
    }
    fold_conv_bn(st->bottleneck[0], 512, 256, eps, &F.bw[0], &F.bb[0]);
    if (build_f32_conv(h, F.bw[0], F.bb[0], 256, 512, &h->f_bott[0])) return 1;
    fold_conv_bn(st->bottleneck[1], 512, 512, eps, &F.bw[1], &F.bb[1]);
    if (build_f32_conv(h, F.bw[1], F.bb[1], 512, 512, &h->f_bott[1])) return 1;
    for (int k = 0; k < 4; ++k) {   // decoder: up_t[k] / up_c[k] act at level l = 3 - k
        const int f = kFeat[3 - k];
        const ogl_convt& T = st->up_t[k];
        F.tw[k].assign(T.weight, T.weight + static_cast<size_t>(2 * f) * f * 4);
        F.tb[k].assign(T.bias, T.bias + f);
        h->f_upt[k].cin = 2 * f;
        h->f_upt[k].cout = f;
        if (dev_upload(h, F.tw[k], &h->f_upt[k].w) || dev_upload(h, F.tb[k], &h->f_upt[k].b)) return 1;
        fold_conv_bn(st->up_c[k][0], f, 2 * f, eps, &F.uw[k][0], &F.ub[k][0]);
        if (build_f32_conv(h, F.uw[k][0], F.ub[k][0], 2 * f, f, &h->f_up[k][0])) return 1;
        fold_conv_bn(st->up_c[k][1], f, f, eps, &F.uw[k][1], &F.ub[k][1]);
        if (build_f32_conv(h, F.uw[k][1], F.ub[k][1], f, f, &h->f_up[k][1])) return 1;
    }
    std::vector<float> hw(st->head_weight, st->head_weight + 32);
    if (dev_upload(h, hw, &h->head_w)) return 1;
    memcpy(h->head_w_host, hw.data(), sizeof h->head_w_host);
    h->head_b = st->head_bias[0];
    h->loaded = true;
    return build_pack(h, false);
}

int ogl_unet_prepare(ogl_unet* h, int precision) {
    if (!h) return fail("ogl_unet_prepare: NULL handle");
    if (!h->loaded) return fail("ogl_unet_prepare: no weights loaded (ogl_unet_load_state)");
    if (precision != OGL_PRECISION_BF16 && precision != OGL_PRECISION_F32 &&
        precision != OGL_PRECISION_F16)
        return fail("ogl_unet_prepare: unknown precision mode");
    OGL_CUDA(cudaSetDevice(h->device));
    if (precision == OGL_PRECISION_F16 && !h->pack[1].built) return build_pack(h, true);
    return 0;
}

size_t ogl_unet_workspace_bytes(const ogl_unet* h, int n, int height, int width, int precision) {
    (void)h;
    if (n <= 0 || height <= 0 || width <= 0) return 0;
    const size_t elem = precision == OGL_PRECISION_F32 ? 4 : 2;
    Plan p = make_plan(n, height, width, elem);
    size_t extra = 0;
    if (precision == OGL_PRECISION_F32) {
        // fp32 path also needs the scaled input plane and pooled level-0..2 tensors
        extra = static_cast<size_t>(n) * height * width * 4 + 256;
    }
    return p.total + extra + 256;
}

int ogl_unet_forward(ogl_unet* h, const void* frames_dev, int in_dtype, int n, int height,
                     int width, void* workspace_dev, size_t workspace_bytes, float* logits_dev,
                     uint8_t* mask_dev, int32_t* area_dev, float threshold, int precision,
                     void* stream_v) {
    if (!h) return fail("ogl_unet_forward: NULL handle");
    if (!h->loaded) return fail("ogl_unet_forward: no weights loaded (ogl_unet_load_state)");
    if (!frames_dev || !workspace_dev) return fail("ogl_unet_forward: NULL frames/workspace");
    if (n <= 0) return fail("ogl_unet_forward: n must be positive");
    if (height % 16 || width % 16 || height <= 0 || width <= 0)
        return fail("ogl_unet_forward: height and width must be positive multiples of 16");
    if (in_dtype != OGL_DTYPE_U8 && in_dtype != OGL_DTYPE_F32)
        return fail("ogl_unet_forward: in_dtype must be OGL_DTYPE_U8 or OGL_DTYPE_F32");
    if (!(threshold > 0.f && threshold < 1.f))
        return fail("ogl_unet_forward: threshold must be in (0, 1)");
    if (precision != OGL_PRECISION_BF16 && precision != OGL_PRECISION_F32 &&
        precision != OGL_PRECISION_F16)
        return fail("ogl_unet_forward: unknown precision mode");
    if (workspace_bytes < ogl_unet_workspace_bytes(h, n, height, width, precision))
        return fail("ogl_unet_forward: workspace too small");
    if (reinterpret_cast<uintptr_t>(workspace_dev) & 255)
        return fail("ogl_unet_forward: workspace must be 256-byte aligned");
    {   // the handle's tensors, weights and kernels live on ONE device: it must be the current one
        int cur = -1;
        OGL_CUDA(cudaGetDevice(&cur));
        if (cur != h->device)
            return fail("ogl_unet_forward: the handle's device is not the current device");
    }
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    const float thr = static_cast<float>(
        std::log(static_cast<double>(threshold) / (1.0 - static_cast<double>(threshold))));
    const int H = height, W = width;
    uint8_t* ws = static_cast<uint8_t*>(workspace_dev);
    if (area_dev) OGL_CUDA(cudaMemsetAsync(area_dev, 0, sizeof(int32_t) * n, stream));

    if (precision != OGL_PRECISION_F32) {
        const bool f16 = precision == OGL_PRECISION_F16;
        const ogl_unet::Pack& K = h->pack[f16 ? 1 : 0];
        if (!K.built)
            return fail("ogl_unet_forward: f16 operands are not packed (call ogl_unet_prepare first)");
        // the two operand types are the same kernels compiled twice (conv_tc_f16.cu, s2d_tc_f16.cu)
        auto conv_tc = [&](auto&&... a) { return f16 ? launch_conv_tc_f16(a...) : launch_conv_tc(a...); };
        auto s2d_tc = [&](auto&&... a) { return f16 ? launch_s2d_tc_f16(a...) : launch_s2d_tc(a...); };
        auto upcat_tc = [&](auto&&... a) { return f16 ? launch_upcat_tc_f16(a...) : launch_upcat_tc(a...); };
        const bool compose = h->compose_up;
        const Plan p = make_plan(n, H, W, 2);
        auto B = [&](size_t off) { return reinterpret_cast<__nv_bfloat16*>(ws + off); };
        __nv_bfloat16* P[4] = {B(p.U[1]), B(p.U[2]), B(p.U[3]), B(p.P3)};
        h->n_events = 0;
        if (h->profile) {
            h->prof_set = h->prof_forwards % kProfSets;
            ++h->prof_forwards;
        }
        h->launch_names.clear();
        const bool s2d = h->use_s2d;
        mark(h, stream, nullptr);
        // one launch of the schedule: enqueued once (or rep_n times when it is the launch picked by
        // ogl_unet_set_repeat for an energy / time measurement), then named and time-stamped
        // OGL_PINGPONG=1: launches alternate their tile order (rev) so that each starts with the
        // frames the previous one wrote last, which are still in L2. Measured (batch 512,
        // gpurun_out/exp_pp.jsonl -> DESIGN.md section 6): 10.58 -> 10.67 ms, no gain; off.
        static const bool pingpong = getenv("OGL_PINGPONG") && atoi(getenv("OGL_PINGPONG")) != 0;
        int launch_idx = 0;
        bool rev = false;
        auto step = [&](const char* name, auto&& launch) {
            const int idx = launch_idx++;
            rev = pingpong && (idx & 1);
            // a repeated head launch would add every frame's popcount rep_n times (the area vector
            // is zeroed once per forward): a measurement aid must not return wrong areas
            if (idx == h->rep_launch && h->rep_n > 1 && area_dev && strstr(name, "head"))
                return fail("ogl_unet_forward: the head launch is repeated (ogl_unet_set_repeat); "
                            "pass area = NULL or reset the repetition");
            nvtxRangePushA(name);   // the reference layer names in nsys / ncu timelines
            for (int r = idx == h->rep_launch ? h->rep_n : 1; r > 0; --r)
                if (launch()) {
                    nvtxRangePop();
                    return 1;
                }
            nvtxRangePop();
            mark(h, stream, name);
            return 0;
        };
        const int sms = h->num_sms, cg = h->cta_group;
        const bool fused_stem = s2d && h->fuse_stem && in_dtype == OGL_DTYPE_U8;
        if (!fused_stem &&
            step(kDownC1[0], [&] {
                return launch_stem(frames_dev, in_dtype, h->stem, n, H, W, B(p.T[0]), s2d, stream, f16);
            }))
            return 1;
        for (int l = 0; l < 4; ++l) {
            const int hh = H >> l, ww = W >> l;
            if (l > 0 && step(kDownC1[l], [&] {
                    return conv_tc(K.down_c1[l], P[l - 1], nullptr, n, hh, ww, B(p.T[l]),
                                          nullptr, nullptr, sms, stream, cg, rev);
                }))
                return 1;
            if (l == 0 && fused_stem) {
                if (step("stem+downs.0.net.3+pool", [&] {
                        return s2d_tc(K.s2d_down, nullptr, nullptr, n, H, W, B(p.S[0]), P[0],
                                             nullptr, sms, stream, cg,
                                             static_cast<const uint8_t*>(frames_dev), &h->stem, rev,
                                             h->fuse_stem == 3   ? h->stem_tc3
                                             : h->fuse_stem == 2 ? h->stem_tc
                                                                 : nullptr,
                                             h->fuse_stem == 3 ? 16 : 8);
                    }))
                    return 1;
                continue;
            }
            if (step(kDownC2[l], [&] {
                    if (l == 0 && s2d)
                        return s2d_tc(K.s2d_down, B(p.T[0]), nullptr, n, H, W, B(p.S[0]), P[0],
                                             nullptr, sms, stream, cg, nullptr, nullptr, rev);
                    // composed decoder: the skip tensors of levels 1-3 are written space-to-depth
                    return conv_tc(K.down_c2[l], B(p.T[l]), nullptr, n, hh, ww, B(p.S[l]),
                                          P[l], nullptr, sms, stream, cg, rev, compose && l >= 1);
                }))
                return 1;
        }
        if (step("bottleneck.net.0", [&] {
                return conv_tc(K.bott[0], P[3], nullptr, n, H >> 4, W >> 4, B(p.T[4]), nullptr,
                                      nullptr, sms, stream, cg, rev);
            }))
            return 1;
        if (step("bottleneck.net.3", [&] {
                return conv_tc(K.bott[1], B(p.T[4]), nullptr, n, H >> 4, W >> 4, B(p.B4),
                                      nullptr, nullptr, sms, stream, cg, rev);
            }))
            return 1;
        const __nv_bfloat16* below = B(p.B4);
        HeadParams hp;
        hp.w = h->head_w;
        hp.w_host = h->head_w_host;
        hp.b = h->head_b;
        hp.logit_thr = thr;
        hp.logits = logits_dev;
        hp.mask = mask_dev;
        hp.area = area_dev;
        for (int k = 0; k < 4; ++k) {
            const int l = 3 - k;
            const int hh = H >> l, ww = W >> l;
            if (l == 0 && s2d) {
                // ups.6 is composed into ups.7.net.0: reads the skip (S2D) and the level-1 tensor
                if (step("ups.6(convT)+ups.7.net.0(cat)", [&] {
                        return s2d_tc(K.s2d_up0, B(p.S[0]), below, n, H, W, B(p.T[0]), nullptr,
                                             nullptr, sms, stream, cg, nullptr, nullptr, rev);
                    }))
                    return 1;
                if (step(kUpC2[k], [&] {
                        return s2d_tc(K.s2d_up1, B(p.T[0]), nullptr, n, H, W, nullptr, nullptr,
                                             &hp, sms, stream, cg, nullptr, nullptr, rev);
                    }))
                    return 1;
                break;
            }
            if (compose && l >= 1) {
                // ups.{2k} is composed into ups.{2k+1}.net.0: reads the skip (S2D) and the tensor below
                if (step(kUpTC[k], [&] {
                        return upcat_tc(K.upcat[k], B(p.S[l]), below, n, hh, ww, B(p.T[l]), sms, stream, cg);
                    }))
                    return 1;
            } else {
                if (step(kUpT[k], [&] {
                        return conv_tc(K.up_t[k], below, nullptr, n, hh / 2, ww / 2, B(p.U[l]),
                                              nullptr, nullptr, sms, stream, cg, rev);
                    }))
                    return 1;
                if (step(kUpC1[k], [&] {
                        return conv_tc(K.up_c[k][0], B(p.S[l]), B(p.U[l]), n, hh, ww, B(p.T[l]),
                                              nullptr, nullptr, sms, stream, cg, rev);
                    }))
                    return 1;
            }
            if (step(kUpC2[k], [&] {
                    return conv_tc(K.up_c[k][1], B(p.T[l]), nullptr, n, hh, ww, B(p.U[l]),
                                          nullptr, k == 3 ? &hp : nullptr, sms, stream, cg, rev);
                }))
                return 1;
            below = B(p.U[l]);
        }
        return 0;
    }

    // ------------------------------------------------ fp32 validation path (NCHW)
    const Plan p = make_plan(n, H, W, 4);
    auto F = [&](size_t off) { return reinterpret_cast<float*>(ws + off); };
    float* x0 = F(p.total);  // scaled input [n][1][H][W]
    float* P[4] = {F(p.U[1]), F(p.U[2]), F(p.U[3]), F(p.P3)};
    if (launch_f32_input(frames_dev, in_dtype, static_cast<int64_t>(n) * H * W, x0, stream))
        return 1;
    const float* cur = x0;
    int cin = 1;
    for (int l = 0; l < 4; ++l) {
        const int hh = H >> l, ww = W >> l, f = kFeat[l];
        if (launch_f32_conv3x3(cur, cin, nullptr, 0, h->f_down[l][0].w, h->f_down[l][0].b,
                               F(p.T[l]), n, f, hh, ww, 1, stream))
            return 1;
        if (launch_f32_conv3x3(F(p.T[l]), f, nullptr, 0, h->f_down[l][1].w, h->f_down[l][1].b,
                               F(p.S[l]), n, f, hh, ww, 1, stream))
            return 1;
        if (launch_f32_maxpool(F(p.S[l]), P[l], n * f, hh, ww, stream)) return 1;
        cur = P[l];
        cin = f;
    }
    if (launch_f32_conv3x3(P[3], 256, nullptr, 0, h->f_bott[0].w, h->f_bott[0].b, F(p.T[4]), n,
                           512, H >> 4, W >> 4, 1, stream))
        return 1;
    if (launch_f32_conv3x3(F(p.T[4]), 512, nullptr, 0, h->f_bott[1].w, h->f_bott[1].b, F(p.B4), n,
                           512, H >> 4, W >> 4, 1, stream))
        return 1;
    const float* below = F(p.B4);
    for (int k = 0; k < 4; ++k) {
        const int l = 3 - k;
        const int hh = H >> l, ww = W >> l, f = kFeat[l];
        if (launch_f32_convt(below, h->f_upt[k].w, h->f_upt[k].b, F(p.U[l]), n, 2 * f, f, hh / 2,
                             ww / 2, stream))
            return 1;
        if (launch_f32_conv3x3(F(p.S[l]), f, F(p.U[l]), f, h->f_up[k][0].w, h->f_up[k][0].b,
                               F(p.T[l]), n, f, hh, ww, 1, stream))
            return 1;
        if (launch_f32_conv3x3(F(p.T[l]), f, nullptr, 0, h->f_up[k][1].w, h->f_up[k][1].b,
                               F(p.U[l]), n, f, hh, ww, 1, stream))
            return 1;
        below = F(p.U[l]);
    }
    return launch_f32_head(below, h->head_w, h->head_b, thr, n, 32, H, W, logits_dev, mask_dev,
                           area_dev, stream);
}

int ogl_unet_set_profiling(ogl_unet* h, int enable) {
    if (!h) return fail("ogl_unet_set_profiling: NULL handle");
    OGL_CUDA(cudaSetDevice(h->device));
    if (enable && h->events.empty()) {
        h->events.resize(static_cast<size_t>(kProfSets) * kMaxLaunches);
        for (auto& e : h->events) OGL_CUDA(cudaEventCreate(&e));
    }
    h->profile = enable != 0;
    h->n_events = 0;
    h->prof_forwards = 0;
    for (int& c : h->prof_count) c = 0;
    return 0;
}

int ogl_unet_layer_times(ogl_unet* h, float* ms_out, int capacity, int* count_out) {
    if (!h || !ms_out || !count_out) return fail("ogl_unet_layer_times: NULL argument");
    *count_out = 0;
    if (!h->profile || h->prof_forwards < 1 || h->n_events < 2)
        return fail("ogl_unet_layer_times: no profiled forward");
    // average over the (up to kProfSets) most recent profiled forwards with the same launch list
    const int n = h->n_events - 1;
    const int sets = h->prof_forwards < kProfSets ? h->prof_forwards : kProfSets;
    for (int i = 0; i < n && i < capacity; ++i) ms_out[i] = 0.f;
    int used = 0;
    for (int sidx = 0; sidx < sets; ++sidx) {
        if (h->prof_count[sidx] != h->n_events) continue;
        cudaEvent_t* ev = h->events.data() + static_cast<size_t>(sidx) * kMaxLaunches;
        OGL_CUDA(cudaEventSynchronize(ev[n]));
        for (int i = 0; i < n && i < capacity; ++i) {
            float ms = 0.f;
            OGL_CUDA(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
            ms_out[i] += ms;
        }
        ++used;
    }
    if (used == 0) return fail("ogl_unet_layer_times: no complete profiled forward");
    for (int i = 0; i < n && i < capacity; ++i) ms_out[i] /= static_cast<float>(used);
    *count_out = n < capacity ? n : capacity;
    return 0;
}

int ogl_unet_launch_count(const ogl_unet* h) {
    return h ? static_cast<int>(h->launch_names.size()) : 0;
}

const char* ogl_unet_launch_name(const ogl_unet* h, int index) {
    if (!h || index < 0 || index >= static_cast<int>(h->launch_names.size())) return "";
    return h->launch_names[index];
}

int ogl_unet_set_schedule(ogl_unet* h, int s2d_level0) {
    if (!h) return fail("ogl_unet_set_schedule: NULL handle");
    h->use_s2d = s2d_level0 != 0;
    return 0;
}

int ogl_unet_set_compose(ogl_unet* h, int enable) {
    if (!h) return fail("ogl_unet_set_compose: NULL handle");
    h->compose_up = enable != 0;
    return 0;
}

int ogl_unet_set_fused_stem(ogl_unet* h, int enable) {
    if (!h) return fail("ogl_unet_set_fused_stem: NULL handle");
    if (enable < 0 || enable > 3) return fail("ogl_unet_set_fused_stem: mode must be 0, 1, 2 or 3");
    h->fuse_stem = enable;
    return 0;
}

int ogl_unet_set_cta_pairs(ogl_unet* h, int mode) {
    if (!h) return fail("ogl_unet_set_cta_pairs: NULL handle");
    if (mode < 1 || mode > 3) return fail("ogl_unet_set_cta_pairs: mode must be 1, 2 or 3");
    h->cta_group = mode;
    return 0;
}

int ogl_unet_set_repeat(ogl_unet* h, int launch_index, int times) {
    if (!h) return fail("ogl_unet_set_repeat: NULL handle");
    if (times < 1 || times > 64) return fail("ogl_unet_set_repeat: times must be in 1..64");
    h->rep_launch = launch_index;
    h->rep_n = times;
    return 0;
}

size_t ogl_features_workspace_bytes(int64_t n) { return features_workspace_bytes(n); }

int ogl_features(const int32_t* area_dev, int64_t n, double* out8_dev, int32_t* flags2_dev,
                 void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (!area_dev || !out8_dev || !flags2_dev || !workspace_dev)
        return fail("ogl_features: NULL argument");
    return launch_features(area_dev, n, out8_dev, flags2_dev, workspace_dev, workspace_bytes,
                           static_cast<cudaStream_t>(stream));
}

int ogl_features_f64(const double* area_dev, int64_t n, double* out8_dev, int32_t* flags2_dev,
                     void* workspace_dev, size_t workspace_bytes, void* stream) {
    if (!area_dev || !out8_dev || !flags2_dev || !workspace_dev)
        return fail("ogl_features_f64: NULL argument");
    return launch_features_f64(area_dev, n, out8_dev, flags2_dev, workspace_dev, workspace_bytes,
                               static_cast<cudaStream_t>(stream));
}

int ogl_bgr_to_gray(const uint8_t* bgr_dev, uint8_t* gray_dev, int64_t pixels, void* stream) {
    if (!bgr_dev || !gray_dev || pixels < 0) return fail("ogl_bgr_to_gray: bad argument");
    if (pixels == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    long long done = 0;
    if (pixels >= 16 && ((reinterpret_cast<uintptr_t>(bgr_dev) | reinterpret_cast<uintptr_t>(gray_dev)) & 15) == 0) {
        const long long groups = pixels / 16;
        long long g = (groups + 255) / 256;
        if (g > 148 * 16) g = 148 * 16;
        bgr_to_gray_vec16_kernel<<<static_cast<int>(g), 256, 0, st>>>(
            reinterpret_cast<const uint4*>(bgr_dev), reinterpret_cast<uint4*>(gray_dev), groups);
        OGL_CUDA(cudaGetLastError());
        done = groups * 16;
    }
    if (done < pixels) {        // unaligned buffers, or the last (pixels % 16) pixels
        long long g = (pixels - done + 255) / 256;
        if (g > 148 * 16) g = 148 * 16;
        bgr_to_gray_kernel<<<static_cast<int>(g), 256, 0, st>>>(bgr_dev + 3 * done, gray_dev + done,
                                                                 pixels - done);
        OGL_CUDA(cudaGetLastError());
    }
    return 0;
}

namespace {
constexpr int kFramesPerLaunch = 32768;  // grid.y limit
}

int ogl_mask_area_boxes(const uint8_t* mask_dev, int n, int height, int width,
                        const int32_t* boxes_dev, const uint8_t* has_box_dev, int32_t* area_dev,
                        void* stream) {
    if (!mask_dev || !boxes_dev || !area_dev || n < 0 || height <= 0 || width <= 0)
        return fail("ogl_mask_area_boxes: bad argument");
    const size_t hw = static_cast<size_t>(height) * width;
    for (int i = 0; i < n; i += kFramesPerLaunch) {
        const int m = n - i < kFramesPerLaunch ? n - i : kFramesPerLaunch;
        if (launch_mask_area_boxes(mask_dev + i * hw, m, height, width, boxes_dev + 4 * i,
                                   has_box_dev ? has_box_dev + i : nullptr, area_dev + i,
                                   static_cast<cudaStream_t>(stream)))
            return 1;
    }
    return 0;
}

int ogl_letterbox_crops(const uint8_t* gray_dev, int n, int height, int width,
                        const int32_t* geom_dev, int size, uint8_t* out_dev, void* stream) {
    if (!gray_dev || !geom_dev || !out_dev || n < 0 || height <= 0 || width <= 0 || size <= 0)
        return fail("ogl_letterbox_crops: bad argument");
    const size_t hw = static_cast<size_t>(height) * width, ss = static_cast<size_t>(size) * size;
    for (int i = 0; i < n; i += kFramesPerLaunch) {
        const int m = n - i < kFramesPerLaunch ? n - i : kFramesPerLaunch;
        if (launch_letterbox_crops(gray_dev + i * hw, m, height, width, geom_dev + 8 * i, size,
                                   out_dev + i * ss, static_cast<cudaStream_t>(stream)))
            return 1;
    }
    return 0;
}

int ogl_unletterbox_area(const uint8_t* mask_cs_dev, int n, int size, const int32_t* geom_dev,
                         int height, int width, uint8_t* full_mask_dev, int32_t* area_dev,
                         void* stream) {
    if (!mask_cs_dev || !geom_dev || !area_dev || n < 0 || height <= 0 || width <= 0 || size <= 0)
        return fail("ogl_unletterbox_area: bad argument");
    const size_t hw = static_cast<size_t>(height) * width, ss = static_cast<size_t>(size) * size;
    for (int i = 0; i < n; i += kFramesPerLaunch) {
        const int m = n - i < kFramesPerLaunch ? n - i : kFramesPerLaunch;
        if (launch_unletterbox_area(mask_cs_dev + i * ss, m, size, geom_dev + 8 * i, height, width,
                                    full_mask_dev ? full_mask_dev + i * hw : nullptr, area_dev + i,
                                    static_cast<cudaStream_t>(stream)))
            return 1;
    }
    return 0;
}

int ogl_resize_u8_linear(const uint8_t* src_dev, int n, int src_h, int src_w, uint8_t* dst_dev,
                         int dst_h, int dst_w, void* stream) {
    if (!src_dev || !dst_dev || n < 0 || src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0)
        return fail("ogl_resize_u8_linear: bad argument");
    const size_t ss = static_cast<size_t>(src_h) * src_w, ds = static_cast<size_t>(dst_h) * dst_w;
    for (int i = 0; i < n; i += kFramesPerLaunch) {
        const int m = n - i < kFramesPerLaunch ? n - i : kFramesPerLaunch;
        if (launch_resize_u8_linear(src_dev + i * ss, m, src_h, src_w, dst_dev + i * ds, dst_h, dst_w,
                                    static_cast<cudaStream_t>(stream)))
            return 1;
    }
    return 0;
}

int ogl_prob_resize_mask(const float* logits_dev, int n, int src_h, int src_w, int dst_h,
                         int dst_w, float threshold, uint8_t* mask_dev, int32_t* area_dev,
                         void* stream) {
    if (!logits_dev || n < 0 || src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0)
        return fail("ogl_prob_resize_mask: bad argument");
    if (!(threshold > 0.f && threshold < 1.f))
        return fail("ogl_prob_resize_mask: threshold must be in (0, 1)");
    const size_t ss = static_cast<size_t>(src_h) * src_w, ds = static_cast<size_t>(dst_h) * dst_w;
    for (int i = 0; i < n; i += kFramesPerLaunch) {
        const int m = n - i < kFramesPerLaunch ? n - i : kFramesPerLaunch;
        if (launch_prob_resize_mask(logits_dev + i * ss, m, src_h, src_w, dst_h, dst_w, threshold,
                                    mask_dev ? mask_dev + i * ds : nullptr,
                                    area_dev ? area_dev + i : nullptr,
                                    static_cast<cudaStream_t>(stream)))
            return 1;
    }
    return 0;
}

int ogl_mask_overlap_counts(const uint8_t* pred_dev, const uint8_t* gt_dev, int n, int64_t pixels,
                            int32_t* counts_dev, void* stream) {
    if (!pred_dev || !gt_dev || !counts_dev || n < 0 || pixels <= 0 || pixels > 0x7fffffffLL)
        return fail("ogl_mask_overlap_counts: bad argument");
    for (int i = 0; i < n; i += kFramesPerLaunch) {
        const int m = n - i < kFramesPerLaunch ? n - i : kFramesPerLaunch;
        if (launch_overlap_counts(pred_dev + i * pixels, gt_dev + i * pixels, m, pixels,
                                  counts_dev + 3 * i, static_cast<cudaStream_t>(stream)))
            return 1;
    }
    return 0;
}

int ogl_debug_tc_layer(ogl_unet* h, int kind, const float* src0_dev, int c0, const float* src1_dev,
                       int c1, const float* weight_host, const float* bias_host, int cout, int n,
                       int height, int width, float* out_dev, float* out_pool_dev,
                       void* stream_v) {
    g_err.clear();
    if (!h || !src0_dev || !weight_host || !bias_host || !out_dev)
        return fail("ogl_debug_tc_layer: NULL argument");
    if (kind != EPI_RELU && kind != EPI_RELU_POOL && kind != EPI_CONVT)
        return fail("ogl_debug_tc_layer: kind must be 0, 1 or 3");
    if (c0 % 32 || c1 % 32 || cout % 32) return fail("ogl_debug_tc_layer: channels % 32 != 0");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    OGL_CUDA(cudaSetDevice(h->device));
    const int cin = c0 + c1;
    TcLayer L;
    L.cin0 = c0;
    L.cin1 = c1;
    L.cout = cout;
    L.N = kind == EPI_CONVT ? convt_n(cout) : (cout < 128 ? cout : 128);
    L.epi = kind;
    std::vector<__nv_bfloat16> pk, pk2;
    if (kind == EPI_CONVT) {
        L.taps = 1;
        L.npass = 4 * cout / L.N;
        pk = pack_convt(weight_host, cin, cout, L.N, false);
        if (L.N == 128) pk2 = pack_convt_pair(weight_host, cin, cout, L.N, false);
    } else {
        L.taps = 9;
        L.npass = cout / L.N;
        std::vector<float> w(weight_host, weight_host + static_cast<size_t>(cout) * cin * 9);
        pk = pack_conv(w, cout, cin, L.N, false);
        if (L.N >= 64) pk2 = pack_conv_pair(w, cout, cin, L.N, false);
    }
    const size_t hw = static_cast<size_t>(height) * width;
    const int oh = kind == EPI_CONVT ? 2 * height : height;
    const int ow = kind == EPI_CONVT ? 2 * width : width;
    __nv_bfloat16 *d_w = nullptr, *d_w2 = nullptr, *d_s0 = nullptr, *d_s1 = nullptr, *d_o = nullptr,
                  *d_p = nullptr;
    float* d_b = nullptr;
    int rc = 1;
    do {
        if (cudaMalloc(&d_w, pk.size() * 2) != cudaSuccess) break;
        if (!pk2.empty() && cudaMalloc(&d_w2, pk2.size() * 2) != cudaSuccess) break;
        if (cudaMalloc(&d_b, cout * 4) != cudaSuccess) break;
        if (cudaMalloc(&d_s0, n * c0 * hw * 2) != cudaSuccess) break;
        if (c1 && cudaMalloc(&d_s1, n * c1 * hw * 2) != cudaSuccess) break;
        if (cudaMalloc(&d_o, static_cast<size_t>(n) * cout * oh * ow * 2) != cudaSuccess) break;
        if (kind == EPI_RELU_POOL && cudaMalloc(&d_p, n * cout * hw / 4 * 2) != cudaSuccess) break;
        cudaMemcpyAsync(d_w, pk.data(), pk.size() * 2, cudaMemcpyHostToDevice, stream);
        if (d_w2) cudaMemcpyAsync(d_w2, pk2.data(), pk2.size() * 2, cudaMemcpyHostToDevice, stream);
        cudaMemcpyAsync(d_b, bias_host, cout * 4, cudaMemcpyHostToDevice, stream);
        cudaStreamSynchronize(stream);
        L.wpack = d_w;
        L.wpack2 = d_w2;
        L.bias = d_b;
        if (launch_nchw_to_c8(src0_dev, d_s0, n, c0, height, width, stream)) break;
        if (c1 && launch_nchw_to_c8(src1_dev, d_s1, n, c1, height, width, stream)) break;
        if (launch_conv_tc(L, d_s0, d_s1, n, height, width, d_o, d_p, nullptr, h->num_sms, stream,
                           h->cta_group))
            break;
        if (launch_c8_to_nchw(d_o, out_dev, n, cout, oh, ow, stream)) break;
        if (kind == EPI_RELU_POOL && out_pool_dev &&
            launch_c8_to_nchw(d_p, out_pool_dev, n, cout, height / 2, width / 2, stream))
            break;
        cudaError_t e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) {
            fail_cuda(e, "ogl_debug_tc_layer");
            break;
        }
        rc = 0;
    } while (0);
    if (rc && g_err.empty()) fail("ogl_debug_tc_layer: allocation or launch failed");
    cudaFree(d_w);
    cudaFree(d_w2);
    cudaFree(d_b);
    cudaFree(d_s0);
    cudaFree(d_s1);
    cudaFree(d_o);
    cudaFree(d_p);
    return rc;
}

int ogl_debug_upcat_program(const float* w3_host, const float* b3_host, const float* wt_host,
                            const float* bt_host, int f, uint16_t* wskip_out, uint16_t* wskip_pair_out,
                            uint16_t* wbelow_out, uint16_t* wbelow_pair_out, float* btab_out) {
    g_err.clear();
    if (!w3_host || !b3_host || !wt_host || !bt_host)
        return fail("ogl_debug_upcat_program: NULL argument");
    UpcatHost hs;
    if (build_upcat_host(w3_host, b3_host, wt_host, bt_host, f, false, &hs)) return 1;
    if (wskip_out) memcpy(wskip_out, hs.wskip.data(), hs.wskip.size() * 2);
    if (wskip_pair_out) memcpy(wskip_pair_out, hs.wskip_pair.data(), hs.wskip_pair.size() * 2);
    if (wbelow_out) memcpy(wbelow_out, hs.wbelow.data(), hs.wbelow.size() * 2);
    if (wbelow_pair_out) memcpy(wbelow_pair_out, hs.wbelow_pair.data(), hs.wbelow_pair.size() * 2);
    if (btab_out) memcpy(btab_out, hs.btab.data(), hs.btab.size() * sizeof(float));
    return 0;
}

int ogl_debug_upcat_layer(ogl_unet* h, const float* skip_dev, const float* below_dev,
                          const float* w3_host, const float* b3_host, const float* wt_host,
                          const float* bt_host, int f, int n, int height, int width, float* out_dev,
                          void* stream_v) {
    g_err.clear();
    if (!h || !skip_dev || !below_dev || !w3_host || !b3_host || !wt_host || !bt_host || !out_dev)
        return fail("ogl_debug_upcat_layer: NULL argument");
    if (height % 2 || width % 2 || height < 2 || width < 2)
        return fail("ogl_debug_upcat_layer: H and W must be even");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    OGL_CUDA(cudaSetDevice(h->device));
    UpcatHost hs;
    if (build_upcat_host(w3_host, b3_host, wt_host, bt_host, f, false, &hs)) return 1;
    const size_t hw = static_cast<size_t>(height) * width;
    std::vector<float> interior(hs.btab.begin() + 4 * f, hs.btab.begin() + 5 * f);
    void *d_ws = nullptr, *d_ws2 = nullptr, *d_wb = nullptr, *d_wb2 = nullptr;
    float *d_bt = nullptr, *d_bi = nullptr;
    __nv_bfloat16 *d_s = nullptr, *d_b = nullptr, *d_o = nullptr;
    int rc = 1;
    do {
        if (cudaMalloc(&d_ws, hs.wskip.size() * 2) != cudaSuccess) break;
        if (cudaMalloc(&d_ws2, hs.wskip_pair.size() * 2) != cudaSuccess) break;
        if (cudaMalloc(&d_wb, hs.wbelow.size() * 2) != cudaSuccess) break;
        if (cudaMalloc(&d_wb2, hs.wbelow_pair.size() * 2) != cudaSuccess) break;
        if (cudaMalloc(&d_bt, hs.btab.size() * 4) != cudaSuccess) break;
        if (cudaMalloc(&d_bi, interior.size() * 4) != cudaSuccess) break;
        if (cudaMalloc(&d_s, n * f * hw * 2) != cudaSuccess) break;
        if (cudaMalloc(&d_b, n * 2 * f * hw / 4 * 2) != cudaSuccess) break;
        if (cudaMalloc(&d_o, n * f * hw * 2) != cudaSuccess) break;
        cudaMemcpyAsync(d_ws, hs.wskip.data(), hs.wskip.size() * 2, cudaMemcpyHostToDevice, stream);
        cudaMemcpyAsync(d_ws2, hs.wskip_pair.data(), hs.wskip_pair.size() * 2, cudaMemcpyHostToDevice, stream);
        cudaMemcpyAsync(d_wb, hs.wbelow.data(), hs.wbelow.size() * 2, cudaMemcpyHostToDevice, stream);
        cudaMemcpyAsync(d_wb2, hs.wbelow_pair.data(), hs.wbelow_pair.size() * 2, cudaMemcpyHostToDevice, stream);
        cudaMemcpyAsync(d_bt, hs.btab.data(), hs.btab.size() * 4, cudaMemcpyHostToDevice, stream);
        cudaMemcpyAsync(d_bi, interior.data(), interior.size() * 4, cudaMemcpyHostToDevice, stream);
        cudaStreamSynchronize(stream);
        UpcatLayer L;
        L.wskip = static_cast<uint8_t*>(d_ws);
        L.wskip2 = static_cast<uint8_t*>(d_ws2);
        L.wbelow = static_cast<uint8_t*>(d_wb);
        L.wbelow2 = static_cast<uint8_t*>(d_wb2);
        L.btab = d_bt;
        L.bias = d_bi;
        L.f = hs.f;
        L.N = hs.N;
        L.npass = hs.npass;
        if (launch_nchw_to_c8(skip_dev, d_s, n, f, height, width, stream, true)) break;
        if (launch_nchw_to_c8(below_dev, d_b, n, 2 * f, height / 2, width / 2, stream)) break;
        if (launch_upcat_tc(L, d_s, d_b, n, height, width, d_o, h->num_sms, stream, h->cta_group)) break;
        if (launch_c8_to_nchw(d_o, out_dev, n, f, height, width, stream)) break;
        cudaError_t e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) {
            fail_cuda(e, "ogl_debug_upcat_layer");
            break;
        }
        rc = 0;
    } while (0);
    if (rc && g_err.empty()) fail("ogl_debug_upcat_layer: allocation or launch failed");
    cudaFree(d_ws);
    cudaFree(d_ws2);
    cudaFree(d_wb);
    cudaFree(d_wb2);
    cudaFree(d_bt);
    cudaFree(d_bi);
    cudaFree(d_s);
    cudaFree(d_b);
    cudaFree(d_o);
    return rc;
}

int ogl_debug_s2d_program(const float* w3_host, const float* b3_host, int cin_s,
                          const float* wt_host, const float* bt_host, uint8_t* wblob_out,
                          size_t wblob_capacity, size_t* wblob_bytes, uint32_t* ops_out,
                          int ops_capacity, int* n_ops, int* stages_out, int* n_stages,
                          float* btab_out) {
    g_err.clear();
    if (!w3_host || !b3_host || !wblob_bytes || !n_ops || !n_stages)
        return fail("ogl_debug_s2d_program: NULL argument");
    S2dHost hs;
    if (build_s2d_host(w3_host, b3_host, cin_s, wt_host, bt_host, wt_host ? 64 : 0, &hs)) return 1;
    *wblob_bytes = hs.wblob.size();
    *n_ops = static_cast<int>(hs.ops.size());
    *n_stages = hs.n_stages;
    if (wblob_out) {
        if (wblob_capacity < hs.wblob.size()) return fail("ogl_debug_s2d_program: wblob too small");
        memcpy(wblob_out, hs.wblob.data(), hs.wblob.size());
    }
    if (ops_out) {
        if (ops_capacity < *n_ops) return fail("ogl_debug_s2d_program: ops buffer too small");
        memcpy(ops_out, hs.ops.data(), hs.ops.size() * sizeof(S2dOp));
    }
    if (stages_out)
        for (int i = 0; i < hs.n_stages; ++i) {
            stages_out[3 * i] = hs.stage_src[i];
            stages_out[3 * i + 1] = hs.stage_plane0[i];
            stages_out[3 * i + 2] = hs.stage_op_end[i];
        }
    if (btab_out) memcpy(btab_out, hs.btab.data(), hs.btab.size() * sizeof(float));
    return 0;
}

int ogl_debug_s2d_layer(ogl_unet* h, int kind, const float* src_dev, int cin_s,
                        const float* below_dev, const float* w3_host, const float* b3_host,
                        const float* wt_host, const float* bt_host, int n, int height, int width,
                        float* out_dev, float* out_pool_dev, void* stream_v) {
    g_err.clear();
    if (!h || !src_dev || !w3_host || !b3_host || !out_dev)
        return fail("ogl_debug_s2d_layer: NULL argument");
    if (kind != EPI_RELU && kind != EPI_RELU_POOL)
        return fail("ogl_debug_s2d_layer: kind must be 0 or 1");
    if ((below_dev != nullptr) != (wt_host != nullptr) || (wt_host && !bt_host))
        return fail("ogl_debug_s2d_layer: below, wt and bt come together");
    if (height % 16 || width % 16) return fail("ogl_debug_s2d_layer: H, W must be multiples of 16");
    cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
    OGL_CUDA(cudaSetDevice(h->device));
    const int cin_b = wt_host ? 64 : 0;
    const int cin3 = cin_s + (wt_host ? 32 : 0);
    std::vector<float> w3(w3_host, w3_host + static_cast<size_t>(32) * cin3 * 9);
    std::vector<float> b3(b3_host, b3_host + 32);
    S2dHost hs;
    if (build_s2d_host(w3.data(), b3.data(), cin_s, wt_host, bt_host, cin_b, &hs)) return 1;
    const size_t hw = static_cast<size_t>(height) * width;
    uint8_t *d_w = nullptr, *d_w2 = nullptr;
    float* d_bt = nullptr;
    __nv_bfloat16 *d_s = nullptr, *d_b = nullptr, *d_o = nullptr, *d_p = nullptr;
    int rc = 1;
    do {
        if (cudaMalloc(&d_w, hs.wblob.size()) != cudaSuccess) break;
        if (cudaMalloc(&d_w2, hs.wblob_pair.size()) != cudaSuccess) break;
        if (cudaMalloc(&d_bt, hs.btab.size() * 4) != cudaSuccess) break;
        if (cudaMalloc(&d_s, n * cin_s * hw * 2) != cudaSuccess) break;
        if (cin_b && cudaMalloc(&d_b, n * cin_b * hw / 4 * 2) != cudaSuccess) break;
        if (cudaMalloc(&d_o, n * 32 * hw * 2) != cudaSuccess) break;
        if (kind == EPI_RELU_POOL && cudaMalloc(&d_p, n * 32 * hw / 4 * 2) != cudaSuccess) break;
        cudaMemcpyAsync(d_w, hs.wblob.data(), hs.wblob.size(), cudaMemcpyHostToDevice, stream);
        cudaMemcpyAsync(d_w2, hs.wblob_pair.data(), hs.wblob_pair.size(), cudaMemcpyHostToDevice,
                        stream);
        cudaMemcpyAsync(d_bt, hs.btab.data(), hs.btab.size() * 4, cudaMemcpyHostToDevice, stream);
        cudaStreamSynchronize(stream);
        S2dLayer L;
        L.wblob = d_w;
        L.wblob2 = d_w2;
        L.wbytes = static_cast<uint32_t>(hs.wblob.size());
        L.n_stages = hs.n_stages;
        for (int i = 0; i < kS2dMaxStages; ++i) {
            L.stage_src[i] = hs.stage_src[i];
            L.stage_plane0[i] = hs.stage_plane0[i];
        }
        L.btab = d_bt;
        for (int i = 0; i < 32; ++i) L.bias_host[i] = hs.btab[4 * 32 + i];
        L.cin_s = cin_s;
        L.cin_b = cin_b;
        L.epi = kind;
        if (launch_nchw_to_c8(src_dev, d_s, n, cin_s, height, width, stream, true)) break;
        if (cin_b && launch_nchw_to_c8(below_dev, d_b, n, cin_b, height / 2, width / 2, stream))
            break;
        if (launch_s2d_tc(L, d_s, d_b, n, height, width, d_o, d_p, nullptr, h->num_sms, stream,
                          h->cta_group))
            break;
        if (launch_c8_to_nchw(d_o, out_dev, n, 32, height, width, stream, true)) break;
        if (kind == EPI_RELU_POOL && out_pool_dev &&
            launch_c8_to_nchw(d_p, out_pool_dev, n, 32, height / 2, width / 2, stream))
            break;
        cudaError_t e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) {
            fail_cuda(e, "ogl_debug_s2d_layer");
            break;
        }
        rc = 0;
    } while (0);
    if (rc && g_err.empty()) fail("ogl_debug_s2d_layer: allocation or launch failed");
    cudaFree(d_w);
    cudaFree(d_w2);
    cudaFree(d_bt);
    cudaFree(d_s);
    cudaFree(d_b);
    cudaFree(d_o);
    cudaFree(d_p);
    return rc;
}

}  // extern "C"
