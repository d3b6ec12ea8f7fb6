// Kinematic features of the glottal-area waveform as small fp64 CUDA reductions.
// Restates /root/reference/openglottal/features.py:38-68 (_kinematic_features):
//   mean, population std, range, open quotient = mean(x > 0.1*mean),
//   f0 = argmax_{k>=1} |rfft(x-mean)|[k] * (1/n)  (None when the peak is bin 1),
//   periodicity = max_{k=1..min(49,n-1)} r[k] / (r[0] + 1e-8), r = raw autocorrelation,
//   cv = std / (mean + 1e-8).
// np.correlate(...,"full") is O(n^2) in the reference but only lags 0..49 are used
// (features.py:55-58) -> 50 dot products here. The exact-length DFT magnitude is obtained with
// Bluestein's algorithm on a power-of-two Stockham FFT (no cuFFT), all in fp64.
#include "internal.h"

namespace ogl {

namespace {

constexpr int kLags = 50;     // lags 0..49
constexpr int kBlock = 256;
constexpr int kMaxBlocks = 1024;

struct Hdr {               // lives at the start of the workspace
    double mean;
    double sumsq;          // r[0]
    double vmin, vmax;
    double r[kLags];
    long long count_open;
    long long peak;        // argmax bin (>= 1)
    double peak_val;
};

struct Part1 {             // per-block partials, pass 1
    long long isum;
    double dsum, vmin, vmax;
};
struct Part2 {             // per-block partials, pass 2
    double r[kLags];
    long long count_open;
};
struct Part3 {
    double val;
    long long idx;
};

template <typename T>
__device__ __forceinline__ double as_double(T v) { return static_cast<double>(v); }

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename T, bool kInt>
__global__ void __launch_bounds__(kBlock)
stats1_kernel(const T* __restrict__ x, long long n, Part1* __restrict__ parts) {
    long long isum = 0;
    double dsum = 0.0, mn = INFINITY, mx = -INFINITY;
    for (long long i = blockIdx.x * static_cast<long long>(kBlock) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * kBlock) {
        const T v = x[i];
        if (kInt) isum += static_cast<long long>(v);
        else dsum += as_double(v);
        mn = fmin(mn, as_double(v));
        mx = fmax(mx, as_double(v));
    }
    __shared__ long long s_i[kBlock / 32];
    __shared__ double s_d[kBlock / 32], s_mn[kBlock / 32], s_mx[kBlock / 32];
    isum = warp_sum_ll(isum);
    dsum = warp_sum(dsum);
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_i[w] = isum; s_d[w] = dsum; s_mn[w] = mn; s_mx[w] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        Part1 p{0, 0.0, INFINITY, -INFINITY};
        for (int k = 0; k < kBlock / 32; ++k) {
            p.isum += s_i[k]; p.dsum += s_d[k];
            p.vmin = fmin(p.vmin, s_mn[k]); p.vmax = fmax(p.vmax, s_mx[k]);
        }
        parts[blockIdx.x] = p;
    }
}

template <bool kInt>
__global__ void reduce1_kernel(const Part1* __restrict__ parts, int nparts, long long n,
                               Hdr* __restrict__ h) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    long long isum = 0;
    double dsum = 0.0, mn = INFINITY, mx = -INFINITY;
    for (int k = 0; k < nparts; ++k) {
        isum += parts[k].isum; dsum += parts[k].dsum;
        mn = fmin(mn, parts[k].vmin); mx = fmax(mx, parts[k].vmax);
    }
    // integer areas: the sum is exact, so this equals numpy's mean bit for bit
    h->mean = (kInt ? static_cast<double>(isum) : dsum) / static_cast<double>(n);
    h->vmin = mn;
    h->vmax = mx;
}

// pass 2: d = x - mean (stored for the DFT), lagged products, open-quotient count
template <typename T>
__global__ void __launch_bounds__(kBlock)
stats2_kernel(const T* __restrict__ x, long long n, const Hdr* __restrict__ h,
              double* __restrict__ d, Part2* __restrict__ parts) {
    const double mean = h->mean;
    const double thr = mean * 0.1;
    double acc[kLags];
#pragma unroll
    for (int k = 0; k < kLags; ++k) acc[k] = 0.0;
    long long cnt = 0;
    for (long long i = blockIdx.x * static_cast<long long>(kBlock) + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * kBlock) {
        const double xi = as_double(x[i]);
        const double di = xi - mean;
        d[i] = di;
        cnt += xi > thr ? 1 : 0;
#pragma unroll
        for (int k = 0; k < kLags; ++k) {
            const long long j = i + k;
            if (j < n) acc[k] = fma(di, as_double(x[j]) - mean, acc[k]);
        }
    }
    __shared__ double s_r[kBlock / 32][kLags];
    __shared__ long long s_c[kBlock / 32];
    const int w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kLags; ++k) {
        const double v = warp_sum(acc[k]);
        if ((threadIdx.x & 31) == 0) s_r[w][k] = v;
    }
    cnt = warp_sum_ll(cnt);
    if ((threadIdx.x & 31) == 0) s_c[w] = cnt;
    __syncthreads();
    if (threadIdx.x < kLags) {
        double v = 0.0;
        for (int k = 0; k < kBlock / 32; ++k) v += s_r[k][threadIdx.x];
        parts[blockIdx.x].r[threadIdx.x] = v;
    }
    if (threadIdx.x == 0) {
        long long c = 0;
        for (int k = 0; k < kBlock / 32; ++k) c += s_c[k];
        parts[blockIdx.x].count_open = c;
    }
}

__global__ void reduce2_kernel(const Part2* __restrict__ parts, int nparts, Hdr* __restrict__ h) {
    const int k = threadIdx.x;
    if (k < kLags) {
        double v = 0.0;
        for (int b = 0; b < nparts; ++b) v += parts[b].r[k];
        h->r[k] = v;
        if (k == 0) h->sumsq = v;
    }
    if (k == kLags) {
        long long c = 0;
        for (int b = 0; b < nparts; ++b) c += parts[b].count_open;
        h->count_open = c;
    }
}

// ------------------------------------------------------------- Bluestein
// w[j] = exp(-i*pi*j^2/n); j^2 mod 2n is taken exactly in 64-bit integers.
__device__ __forceinline__ double2 chirp(long long j, long long n) {
    const unsigned long long m = static_cast<unsigned long long>(2 * n);
    const unsigned long long jj = (static_cast<unsigned long long>(j) % m);
    const unsigned long long q = (jj * jj) % m;  // j < 2^31 -> no overflow
    double s, c;
    sincospi(static_cast<double>(q) / static_cast<double>(n), &s, &c);
    return make_double2(c, -s);
}

__global__ void bluestein_init_kernel(const double* __restrict__ d, long long n, long long M,
                                      double2* __restrict__ a, double2* __restrict__ b) {
    for (long long j = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; j < M;
         j += static_cast<long long>(gridDim.x) * blockDim.x) {
        double2 av = make_double2(0.0, 0.0), bv = make_double2(0.0, 0.0);
        if (j < n) {
            const double2 w = chirp(j, n);
            av = make_double2(d[j] * w.x, d[j] * w.y);
            bv = make_double2(w.x, -w.y);
        } else if (M - j < n) {
            const double2 w = chirp(M - j, n);
            bv = make_double2(w.x, -w.y);
        }
        a[j] = av;
        b[j] = bv;
    }
}

// One radix-2 Stockham pass: p = current sub-transform length (1,2,4,...,M/2).
__global__ void fft_pass_kernel(const double2* __restrict__ in, double2* __restrict__ out,
                                long long half, long long p, double sign) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < half;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long k = i & (p - 1);
        const long long j = ((i - k) << 1) + k;
        double s, c;
        sincospi(sign * static_cast<double>(k) / static_cast<double>(p), &s, &c);
        const double2 u0 = in[i];
        const double2 v = in[i + half];
        const double2 u1 = make_double2(v.x * c - v.y * s, v.x * s + v.y * c);
        out[j] = make_double2(u0.x + u1.x, u0.y + u1.y);
        out[j + p] = make_double2(u0.x - u1.x, u0.y - u1.y);
    }
}

__global__ void cmul_kernel(double2* __restrict__ a, const double2* __restrict__ b, long long M) {
    for (long long j = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; j < M;
         j += static_cast<long long>(gridDim.x) * blockDim.x) {
        const double2 x = a[j], y = b[j];
        a[j] = make_double2(x.x * y.x - x.y * y.y, x.x * y.y + x.y * y.x);
    }
}

// argmax over bins 1..n/2 of |c[k]| (the chirp factor has unit modulus); first maximum wins.
__global__ void __launch_bounds__(kBlock)
argmax_kernel(const double2* __restrict__ c, long long nbins /* = n/2 */, Part3* __restrict__ parts) {
    double best = -1.0;
    long long bi = 0x7fffffffffffffffLL;
    for (long long k = 1 + blockIdx.x * static_cast<long long>(kBlock) + threadIdx.x; k <= nbins;
         k += static_cast<long long>(gridDim.x) * kBlock) {
        const double m = hypot(c[k].x, c[k].y);
        if (m > best || (m == best && k < bi)) { best = m; bi = k; }
    }
    __shared__ double s_v[kBlock];
    __shared__ long long s_i[kBlock];
    s_v[threadIdx.x] = best;
    s_i[threadIdx.x] = bi;
    __syncthreads();
    for (int o = kBlock / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double v2 = s_v[threadIdx.x + o];
            const long long i2 = s_i[threadIdx.x + o];
            if (v2 > s_v[threadIdx.x] || (v2 == s_v[threadIdx.x] && i2 < s_i[threadIdx.x])) {
                s_v[threadIdx.x] = v2;
                s_i[threadIdx.x] = i2;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { parts[blockIdx.x].val = s_v[0]; parts[blockIdx.x].idx = s_i[0]; }
}

__global__ void finalize_kernel(const Part3* __restrict__ parts, int nparts, long long n,
                                Hdr* __restrict__ h, double* __restrict__ out8,
                                int32_t* __restrict__ flags2) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double best = -1.0;
    long long bi = 1;
    for (int k = 0; k < nparts; ++k) {
        if (parts[k].val > best || (parts[k].val == best && parts[k].idx < bi)) {
            best = parts[k].val;
            bi = parts[k].idx;
        }
    }
    h->peak = bi;
    h->peak_val = best;
    const double nn = static_cast<double>(n);
    const double mean = h->mean;
    const double std = sqrt(h->sumsq / nn);
    double per = -INFINITY;
    const long long kmax = n - 1 < 49 ? n - 1 : 49;
    for (long long k = 1; k <= kmax; ++k) per = fmax(per, h->r[k] / (h->r[0] + 1e-8));
    out8[0] = mean;
    out8[1] = std;
    out8[2] = h->vmax - h->vmin;
    out8[3] = static_cast<double>(h->count_open) / nn;
    out8[4] = static_cast<double>(bi) * (1.0 / nn);  // np.fft.rfftfreq(n)[peak]
    out8[5] = per;
    out8[6] = std / (mean + 1e-8);
    out8[7] = static_cast<double>(bi);
    flags2[0] = h->vmax == 0.0 ? 1 : 0;  // silent waveform -> reference returns None
    flags2[1] = bi == 1 ? 1 : 0;         // peak in first bin -> f0 = None
}

inline long long next_pow2(long long v) {
    long long m = 2;
    while (m < v) m <<= 1;
    return m;
}
inline size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

struct WsLayout {
    size_t hdr, p1, p2, p3, d, a, b, t, total;
    long long M;
};
inline WsLayout ws_layout(long long n) {
    WsLayout L;
    L.M = next_pow2(2 * n - 1);
    size_t o = 0;
    L.hdr = o; o += align256(sizeof(Hdr));
    L.p1 = o; o += align256(sizeof(Part1) * kMaxBlocks);
    L.p2 = o; o += align256(sizeof(Part2) * kMaxBlocks);
    L.p3 = o; o += align256(sizeof(Part3) * kMaxBlocks);
    L.d = o; o += align256(sizeof(double) * static_cast<size_t>(n));
    L.a = o; o += align256(sizeof(double2) * static_cast<size_t>(L.M));
    L.b = o; o += align256(sizeof(double2) * static_cast<size_t>(L.M));
    L.t = o; o += align256(sizeof(double2) * static_cast<size_t>(L.M));
    L.total = o;
    return L;
}

inline int blocks_for(long long n, int per_thread = 4) {
    long long b = (n + static_cast<long long>(kBlock) * per_thread - 1) / (kBlock * per_thread);
    if (b < 1) b = 1;
    if (b > kMaxBlocks) b = kMaxBlocks;
    return static_cast<int>(b);
}

// in-place result: returns pointer to the buffer holding the transform
double2* fft_pow2(double2* x, double2* tmp, long long M, double sign, cudaStream_t stream) {
    double2* in = x;
    double2* out = tmp;
    const long long half = M / 2;
    const int grid = static_cast<int>(half / 256 > 0 ? (half / 256 > 4096 ? 4096 : half / 256) : 1);
    for (long long p = 1; p < M; p <<= 1) {
        fft_pass_kernel<<<grid, 256, 0, stream>>>(in, out, half, p, sign);
        double2* s = in; in = out; out = s;
    }
    return in;
}

template <typename T, bool kInt>
int run_features(const T* x, long long n, double* out8, int32_t* flags2, void* ws, size_t ws_bytes,
                 cudaStream_t stream) {
    if (n < 2) return fail("features need at least 2 samples (the reference raises at n == 1)");
    if (n > (1LL << 30)) return fail("features: n too large");
    const WsLayout L = ws_layout(n);
    if (ws_bytes < L.total) return fail("features workspace too small");
    uint8_t* base = static_cast<uint8_t*>(ws);
    Hdr* h = reinterpret_cast<Hdr*>(base + L.hdr);
    Part1* p1 = reinterpret_cast<Part1*>(base + L.p1);
    Part2* p2 = reinterpret_cast<Part2*>(base + L.p2);
    Part3* p3 = reinterpret_cast<Part3*>(base + L.p3);
    double* d = reinterpret_cast<double*>(base + L.d);
    double2* a = reinterpret_cast<double2*>(base + L.a);
    double2* b = reinterpret_cast<double2*>(base + L.b);
    double2* t = reinterpret_cast<double2*>(base + L.t);

    const int g1 = blocks_for(n);
    stats1_kernel<T, kInt><<<g1, kBlock, 0, stream>>>(x, n, p1);
    reduce1_kernel<kInt><<<1, 32, 0, stream>>>(p1, g1, n, h);
    stats2_kernel<T><<<g1, kBlock, 0, stream>>>(x, n, h, d, p2);
    reduce2_kernel<<<1, 64, 0, stream>>>(p2, g1, h);

    const long long M = L.M;
    const int gm = blocks_for(M, 1);
    bluestein_init_kernel<<<gm, kBlock, 0, stream>>>(d, n, M, a, b);
    double2* fa = fft_pow2(a, t, M, -1.0, stream);
    double2* ta = fa == a ? t : a;           // the free one of {a, t}
    // b's transform needs its own scratch: reuse the free buffer, then restore roles
    double2* fb = fft_pow2(b, ta, M, -1.0, stream);
    double2* free_buf = fb == b ? ta : b;
    cmul_kernel<<<gm, kBlock, 0, stream>>>(fa, fb, M);
    // fb is no longer needed after the product; fa must not alias the scratch
    double2* scratch = free_buf != fa ? free_buf : fb;
    double2* c = fft_pow2(fa, scratch, M, +1.0, stream);
    const long long nbins = n / 2;
    const int g3 = blocks_for(nbins);
    argmax_kernel<<<g3, kBlock, 0, stream>>>(c, nbins, p3);
    finalize_kernel<<<1, 32, 0, stream>>>(p3, g3, n, h, out8, flags2);
    OGL_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

size_t features_workspace_bytes(int64_t n) {
    if (n < 1) n = 1;
    return ws_layout(n).total;
}

int launch_features(const int32_t* area, int64_t n, double* out8, int32_t* flags2, void* ws,
                    size_t ws_bytes, cudaStream_t stream) {
    return run_features<int32_t, true>(area, n, out8, flags2, ws, ws_bytes, stream);
}

int launch_features_f64(const double* area, int64_t n, double* out8, int32_t* flags2, void* ws,
                        size_t ws_bytes, cudaStream_t stream) {
    return run_features<double, false>(area, n, out8, flags2, ws, ws_bytes, stream);
}

}  // namespace ogl
