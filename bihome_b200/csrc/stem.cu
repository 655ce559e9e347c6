// K7: the ResNet stem's BatchNorm2d (batch statistics) -> ReLU -> MaxPool2d(3, stride 2, padding 1) as ONE streaming stage,
// channels-last.  Row f3 of the hot-path table ("fusions around the frozen extractor"): the same three modules open the
// frozen perceptual extractor (torchvision resnet.bn1 / relu / maxpool, reference src/heads/PerceptualHead.py:56-58, four
// passes per training step) and the Zeng backbone (layer1, reference src/backbones/Rethinking.py:31-36, two passes).
//
// Through ATen the stage is six passes over a [B,64,64,64] tensor forward (BatchNorm statistics + normalise, ReLU, pooling
// with int64 indices: 2.1 GB of traffic at B = 256) and about as many backward.  Here:
//   forward   bn_stats_kernel        read x once: per-channel sum / sum of squares (float32 runs of 16 pixels, float64 above)
//             bn_finalize_kernel     fixed-order float64 reduction of the per-CTA partials -> scale, shift, mean, 1/std and
//                                    the running statistics (momentum update, unbiased variance) in place
//             stem_pool_fwd_kernel   read x once more, write the POOLED output (a quarter of the elements) and one byte per
//                                    output with the window position of its maximum (ATen keeps an int64 index)
//   backward  stem_bwd_reduce_kernel read x, the pooled upstream gradient and the codes: sum dy, sum dy * xhat per channel
//             stem_bwd_finalize_kernel  -> gamma / beta gradients and the three per-channel coefficients of the input gradient
//             stem_bwd_apply_kernel  gx = a dy + b x + c   (BatchNorm's batch-statistics backward folded into an affine map)
// Compulsory traffic per pass at B = 256 (x = 268 MB, pooled tensors 67 MB, codes 17 MB): forward 2 x 268 + 67 + 17 = 620 MB,
// backward 3 x 268 + 2 x (67 + 17) = 972 MB.  HBM bound; no re-read beyond the second pass that the statistics force.
//
// Semantics kept (torch.nn.functional.batch_norm / max_pool2d): biased variance for the normalisation, unbiased for
// running_var, eps inside the square root; the pooling maximum is the FIRST maximum in row-major window order (ATen's
// `val > maxval` scan), padding never wins; a window whose maximum is 0 (every pre-activation <= 0) passes no gradient,
// exactly as threshold_backward does after max_pool2d_backward.
#include <math.h>

#include "bh_common.cuh"

namespace bh {

constexpr int kStemThreads = 256;
constexpr int kStemRun = 16;        // pixels a thread accumulates in float32 before it adds to its float64 sums
constexpr int kStemRows = 8;        // pooled rows one forward work item walks

__device__ __forceinline__ float4 ld4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// a kernel launched as a programmatic dependent (host side of the same mechanism)
template <typename... KArgs, typename... Args>
inline int launch_dependent(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
    if (e != cudaSuccess) return static_cast<int>(e);
    return launch_status();
}

// ---- per-CTA partial sums -> partials[stat][C][cta] (float64), fixed order ------------------------------------------
// every thread holds 8 float64 sums (two statistics x its channel quad); pixel lanes are added in lane order
__device__ __forceinline__ void stem_cta_partials(const double (&s)[8], double* __restrict__ partials, int Q, int lq, int C) {
    __shared__ double red[8 * kStemThreads];
    const int q = threadIdx.x & (Q - 1), pl = threadIdx.x >> lq, PL = kStemThreads >> lq;
#pragma unroll
    for (int k = 0; k < 8; ++k) red[(k * PL + pl) * Q + q] = s[k];
    __syncthreads();
    for (int o = threadIdx.x; o < 8 * Q; o += kStemThreads) {
        const int k = o >> lq, qq = o & (Q - 1);
        double sum = 0.0;
        for (int p = 0; p < PL; ++p) sum += red[(k * PL + p) * Q + qq];
        partials[(static_cast<size_t>(k >> 2) * C + qq * 4 + (k & 3)) * gridDim.x + blockIdx.x] = sum;
    }
}

__global__ void __launch_bounds__(kStemThreads) bn_stats_kernel(const float* __restrict__ x, double* __restrict__ partials,
                                                                long long n_pix, int Q, int lq) {
    pdl_launch_dependents();
    const int q = threadIdx.x & (Q - 1), pl = threadIdx.x >> lq, PL = kStemThreads >> lq;
    const long long stride = static_cast<long long>(gridDim.x) * PL;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long p = static_cast<long long>(blockIdx.x) * PL + pl;
    while (p < n_pix) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
#pragma unroll 4
        for (int u = 0; u < kStemRun; ++u) {
            if (p < n_pix) {
                const float4 v = ldg_stream(x4 + p * Q + q);
                a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
                b.x = fmaf(v.x, v.x, b.x); b.y = fmaf(v.y, v.y, b.y); b.z = fmaf(v.z, v.z, b.z); b.w = fmaf(v.w, v.w, b.w);
            }
            p += stride;
        }
        s[0] += a.x; s[1] += a.y; s[2] += a.z; s[3] += a.w;
        s[4] += b.x; s[5] += b.y; s[6] += b.z; s[7] += b.w;
    }
    stem_cta_partials(s, partials, Q, lq, Q * 4);
}

// one warp per channel: lanes stride over the CTAs' partials, shuffle tree (fixed order => bit reproducible).
// The three kernels of a stage are chained by programmatic dependent launches: a dependent grid is scheduled while its
// predecessor still runs and blocks in griddepcontrol.wait until that grid has completed and its writes are visible -- the
// launch latency of the two small follow-up kernels (measured: 17-19 us for a finalize kernel launched the plain way, more
// than the 16 us statistics pass of a 67 MB tensor) disappears behind the predecessor.
__device__ __forceinline__ void stem_sum_partials(const double* __restrict__ partials, int G, int C, int c, double& s0, double& s1) {
    const int lane = threadIdx.x & 31;
    s0 = 0.0; s1 = 0.0;
    const double* p0 = partials + static_cast<size_t>(c) * G;
    const double* p1 = partials + (static_cast<size_t>(C) + c) * G;
#pragma unroll 4
    for (int g = lane; g < G; g += 32) {      // consecutive lanes, consecutive CTAs: 256-byte rows
        s0 += p0[g];
        s1 += p1[g];
    }
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
}

// stats [4][C]: scale = gamma / std, shift = beta - mean * scale, mean, 1/std
__global__ void __launch_bounds__(kStemThreads) bn_finalize_kernel(const double* __restrict__ partials, int G, int C, double n,
                                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                                                   float momentum, float eps, float* __restrict__ stats) {
    pdl_launch_dependents();
    pdl_wait();   // the statistics grid has completed
    const int c = blockIdx.x * (kStemThreads / 32) + (threadIdx.x >> 5);
    if (c >= C) return;
    double s, ss;
    stem_sum_partials(partials, G, C, c, s, ss);
    if ((threadIdx.x & 31) != 0) return;
    const double mean = s / n;
    const double var = fmax(ss / n - mean * mean, 0.0);
    const float mean_f = static_cast<float>(mean);
    const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    const float sc = (gamma ? gamma[c] : 1.0f) * invstd;
    stats[c] = sc;
    stats[C + c] = fmaf(-mean_f, sc, beta ? beta[c] : 0.0f);
    stats[2 * C + c] = mean_f;
    stats[3 * C + c] = invstd;
    if (running_mean) running_mean[c] = fmaf(momentum, mean_f - running_mean[c], running_mean[c]);
    if (running_var) {
        const float unbiased = static_cast<float>(n > 1.0 ? var * n / (n - 1.0) : var);
        running_var[c] = fmaf(momentum, unbiased - running_var[c], running_var[c]);
    }
}

// ---- forward: normalise, ReLU, 3x3 / stride-2 maximum ----------------------------------------------------------------
struct RowMax {
    float4 m;      // horizontal maximum of relu(scale x + shift) over the (up to) three columns of the window
    int kx[4];     // window column 0..2 of that maximum, per channel of the quad
};
__device__ __forceinline__ void take(float v, int k, float& m, int& km) {
    if (v > m) { m = v; km = k; }
}
__device__ __forceinline__ RowMax stem_row(const float* __restrict__ xn, int iy, int ox, int H, int W, int C, int q, const float4& sc,
                                           const float4& sh) {
    RowMax r;
    r.m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    r.kx[0] = r.kx[1] = r.kx[2] = r.kx[3] = 0;
    if (iy < 0 || iy >= H) return r;
    const float* row = xn + (static_cast<size_t>(iy) * W) * C + 4 * q;
    float4 v[3];
    bool in[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int ix = 2 * ox - 1 + k;
        in[k] = ix >= 0 && ix < W;
        if (in[k]) v[k] = ld4(row + static_cast<size_t>(ix) * C);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (!in[k]) continue;
        take(fmaxf(fmaf(v[k].x, sc.x, sh.x), 0.0f), k, r.m.x, r.kx[0]);
        take(fmaxf(fmaf(v[k].y, sc.y, sh.y), 0.0f), k, r.m.y, r.kx[1]);
        take(fmaxf(fmaf(v[k].z, sc.z, sh.z), 0.0f), k, r.m.z, r.kx[2]);
        take(fmaxf(fmaf(v[k].w, sc.w, sh.w), 0.0f), k, r.m.w, r.kx[3]);
    }
    return r;
}
__device__ __forceinline__ void stem_merge(float& m, int& code, float rm, int rk, int ky) {
    if (rm > m) { m = rm; code = 3 * ky + rk; }
}

// work item = (sample, strip of kStemRows pooled rows, PL pooled columns); a thread marches down its column: the bottom
// row of one window is the top row of the next, so every input row is loaded once per strip
template <bool kCode>
__global__ void __launch_bounds__(kStemThreads) stem_pool_fwd_kernel(const float* __restrict__ x, const float* __restrict__ stats,
                                                                     float* __restrict__ y, uint8_t* __restrict__ code, int H,
                                                                     int W, int Ho, int Wo, int Q, int lq, int strips, int xblocks) {
    const int C = Q * 4;
    const int q = threadIdx.x & (Q - 1), pl = threadIdx.x >> lq, PL = kStemThreads >> lq;
    // items in REVERSE order: the statistics pass swept x front to back, so the tail of the tensor is what the L2 still holds
    int item = gridDim.x - 1 - blockIdx.x;
    const int xb = item % xblocks;
    item /= xblocks;
    const int strip = item % strips, n = item / strips;
    const int ox = xb * PL + pl;
    if (ox >= Wo) return;
    pdl_wait();   // scale / shift come from the finalize kernel this grid was launched behind
    const float4 sc = ld4(stats + 4 * q), sh = ld4(stats + C + 4 * q);
    const float* xn = x + static_cast<size_t>(n) * H * W * C;
    const int oy0 = strip * kStemRows, oy1 = min(oy0 + kStemRows, Ho);
    RowMax top = stem_row(xn, 2 * oy0 - 1, ox, H, W, C, q, sc, sh);
    for (int oy = oy0; oy < oy1; ++oy) {
        const RowMax mid = stem_row(xn, 2 * oy, ox, H, W, C, q, sc, sh);
        const RowMax bot = stem_row(xn, 2 * oy + 1, ox, H, W, C, q, sc, sh);
        float4 m = top.m;
        int cd[4] = {top.kx[0], top.kx[1], top.kx[2], top.kx[3]};
        stem_merge(m.x, cd[0], mid.m.x, mid.kx[0], 1); stem_merge(m.y, cd[1], mid.m.y, mid.kx[1], 1);
        stem_merge(m.z, cd[2], mid.m.z, mid.kx[2], 1); stem_merge(m.w, cd[3], mid.m.w, mid.kx[3], 1);
        stem_merge(m.x, cd[0], bot.m.x, bot.kx[0], 2); stem_merge(m.y, cd[1], bot.m.y, bot.kx[1], 2);
        stem_merge(m.z, cd[2], bot.m.z, bot.kx[2], 2); stem_merge(m.w, cd[3], bot.m.w, bot.kx[3], 2);
        const size_t o = ((static_cast<size_t>(n) * Ho + oy) * Wo + ox) * C + 4 * q;
        stg_stream(reinterpret_cast<float4*>(y + o), m);
        if (kCode) *reinterpret_cast<uint32_t*>(code + o) = static_cast<uint32_t>(cd[0]) | (static_cast<uint32_t>(cd[1]) << 8) |
                                                             (static_cast<uint32_t>(cd[2]) << 16) | (static_cast<uint32_t>(cd[3]) << 24);
        top = bot;
    }
}

// ---- backward ---------------------------------------------------------------------------------------------------------
// A thread owns a 2x2 block of input pixels (rows 2k, 2k+1, columns 2m, 2m+1) of its channel quad.  The block is touched by
// exactly four windows -- (k,m), (k,m+1), (k+1,m), (k+1,m+1) -- at nine fixed (pixel, window position) pairs, so the
// upstream gradient is gathered with no divergence and no atomics:
//   pixel (2k,   2m)   : window (k,m) position (1,1)
//   pixel (2k,   2m+1) : (k,m) (1,2) ; (k,m+1) (1,0)
//   pixel (2k+1, 2m)   : (k,m) (2,1) ; (k+1,m) (0,1)
//   pixel (2k+1, 2m+1) : (k,m) (2,2) ; (k,m+1) (2,0) ; (k+1,m) (0,2) ; (k+1,m+1) (0,0)
struct StemBlock {
    float4 x[4];    // the four pixels, row-major
    float4 dy[4];   // gradient w.r.t. the BatchNorm output at those pixels (ReLU and pooling undone)
    bool in[4];
};
__device__ __forceinline__ float pick(uint32_t codes, int ch, int want, float g) {
    return ((codes >> (8 * ch)) & 0xffu) == static_cast<uint32_t>(want) ? g : 0.0f;
}
__device__ __forceinline__ float4 pick4(uint32_t codes, int want, const float4& g) {
    return make_float4(pick(codes, 0, want, g.x), pick(codes, 1, want, g.y), pick(codes, 2, want, g.z), pick(codes, 3, want, g.w));
}
__device__ __forceinline__ void add4(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ float4 gate4(const float4& d, const float4& x, const float4& sc, const float4& sh) {
    return make_float4(fmaf(x.x, sc.x, sh.x) > 0.0f ? d.x : 0.0f, fmaf(x.y, sc.y, sh.y) > 0.0f ? d.y : 0.0f,
                       fmaf(x.z, sc.z, sh.z) > 0.0f ? d.z : 0.0f, fmaf(x.w, sc.w, sh.w) > 0.0f ? d.w : 0.0f);
}
__device__ __forceinline__ void stem_block(StemBlock& b, const float* __restrict__ xn, const float* __restrict__ gn,
                                           const uint8_t* __restrict__ cn, int k, int m, int H, int W, int Ho, int Wo, int C, int q,
                                           const float4& sc, const float4& sh) {
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool wy = k + 1 < Ho, wx = m + 1 < Wo;
    const size_t w00 = (static_cast<size_t>(k) * Wo + m) * C + 4 * q;
    const size_t w01 = w00 + C, w10 = w00 + static_cast<size_t>(Wo) * C, w11 = w10 + C;
    // the four windows: gradient quad and the four position codes
    const float4 g00 = ld4(gn + w00);
    const uint32_t c00 = __ldg(reinterpret_cast<const uint32_t*>(cn + w00));
    const float4 g01 = wx ? ld4(gn + w01) : zero;
    const uint32_t c01 = wx ? __ldg(reinterpret_cast<const uint32_t*>(cn + w01)) : 0xffffffffu;
    const float4 g10 = wy ? ld4(gn + w10) : zero;
    const uint32_t c10 = wy ? __ldg(reinterpret_cast<const uint32_t*>(cn + w10)) : 0xffffffffu;
    const float4 g11 = (wx && wy) ? ld4(gn + w11) : zero;
    const uint32_t c11 = (wx && wy) ? __ldg(reinterpret_cast<const uint32_t*>(cn + w11)) : 0xffffffffu;
    const int iy = 2 * k, ix = 2 * m;
    b.in[0] = true;
    b.in[1] = ix + 1 < W;
    b.in[2] = iy + 1 < H;
    b.in[3] = b.in[1] && b.in[2];
    const float* p = xn + (static_cast<size_t>(iy) * W + ix) * C + 4 * q;
    b.x[0] = ldg_stream(reinterpret_cast<const float4*>(p));
    b.x[1] = b.in[1] ? ldg_stream(reinterpret_cast<const float4*>(p + C)) : zero;
    b.x[2] = b.in[2] ? ldg_stream(reinterpret_cast<const float4*>(p + static_cast<size_t>(W) * C)) : zero;
    b.x[3] = b.in[3] ? ldg_stream(reinterpret_cast<const float4*>(p + static_cast<size_t>(W) * C + C)) : zero;
    b.dy[0] = pick4(c00, 4, g00);
    b.dy[1] = pick4(c00, 5, g00); add4(b.dy[1], pick4(c01, 3, g01));
    b.dy[2] = pick4(c00, 7, g00); add4(b.dy[2], pick4(c10, 1, g10));
    b.dy[3] = pick4(c00, 8, g00); add4(b.dy[3], pick4(c01, 6, g01)); add4(b.dy[3], pick4(c10, 2, g10)); add4(b.dy[3], pick4(c11, 0, g11));
#pragma unroll
    for (int i = 0; i < 4; ++i) b.dy[i] = b.in[i] ? gate4(b.dy[i], b.x[i], sc, sh) : zero;
}

// item = (sample, block row k, xblock); partials[cta][0][C] = sum dy, [cta][1][C] = sum dy * xhat
__global__ void __launch_bounds__(kStemThreads, 2) stem_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ stats,
                                                                       const float* __restrict__ gy, const uint8_t* __restrict__ code,
                                                                       double* __restrict__ partials, int N, int H, int W, int Ho,
                                                                       int Wo, int Q, int lq, int xblocks) {
    pdl_launch_dependents();
    const int C = Q * 4;
    const int q = threadIdx.x & (Q - 1), pl = threadIdx.x >> lq, PL = kStemThreads >> lq;
    const float4 sc = ld4(stats + 4 * q), sh = ld4(stats + C + 4 * q), mu = ld4(stats + 2 * C + 4 * q), is = ld4(stats + 3 * C + 4 * q);
    const float4 off = make_float4(-mu.x * is.x, -mu.y * is.y, -mu.z * is.z, -mu.w * is.w);   // xhat = x * invstd + off
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long items = static_cast<long long>(N) * Ho * xblocks;
    long long item = blockIdx.x;
    while (item < items) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
#pragma unroll 1
        for (int u = 0; u < 4 && item < items; ++u, item += gridDim.x) {
            const int xb = static_cast<int>(item % xblocks);
            const long long t = item / xblocks;
            const int k = static_cast<int>(t % Ho), n = static_cast<int>(t / Ho);
            const int m = xb * PL + pl;
            if (m >= Wo) continue;
            StemBlock blk;
            stem_block(blk, x + static_cast<size_t>(n) * H * W * C, gy + static_cast<size_t>(n) * Ho * Wo * C,
                       code + static_cast<size_t>(n) * Ho * Wo * C, k, m, H, W, Ho, Wo, C, q, sc, sh);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 d = blk.dy[i], xv = blk.x[i];
                a.x += d.x; a.y += d.y; a.z += d.z; a.w += d.w;
                b.x = fmaf(d.x, fmaf(xv.x, is.x, off.x), b.x); b.y = fmaf(d.y, fmaf(xv.y, is.y, off.y), b.y);
                b.z = fmaf(d.z, fmaf(xv.z, is.z, off.z), b.z); b.w = fmaf(d.w, fmaf(xv.w, is.w, off.w), b.w);
            }
        }
        s[0] += a.x; s[1] += a.y; s[2] += a.z; s[3] += a.w;
        s[4] += b.x; s[5] += b.y; s[6] += b.z; s[7] += b.w;
    }
    stem_cta_partials(s, partials, Q, lq, C);
}

// coef [3][C]: gx = a dy + b x + c with a = scale, b = -scale * c2 / std, c = -scale * (c1 - c2 * mean / std),
// c1 = mean(dy), c2 = mean(dy * xhat): BatchNorm's batch-statistics backward.  ggamma = sum dy * xhat, gbeta = sum dy.
__global__ void __launch_bounds__(kStemThreads) stem_bwd_finalize_kernel(const double* __restrict__ partials, int G, int C, double n,
                                                                         const float* __restrict__ stats, float* __restrict__ coef,
                                                                         float* __restrict__ ggamma, float* __restrict__ gbeta) {
    pdl_launch_dependents();
    pdl_wait();
    const int c = blockIdx.x * (kStemThreads / 32) + (threadIdx.x >> 5);
    if (c >= C) return;
    double s1, s2;
    stem_sum_partials(partials, G, C, c, s1, s2);
    if ((threadIdx.x & 31) != 0) return;
    const double sc = stats[c], mean = stats[2 * C + c], invstd = stats[3 * C + c];
    const double c1 = s1 / n, c2 = s2 / n;
    coef[c] = static_cast<float>(sc);
    coef[C + c] = static_cast<float>(-sc * c2 * invstd);
    coef[2 * C + c] = static_cast<float>(-sc * (c1 - c2 * mean * invstd));
    if (ggamma) ggamma[c] = static_cast<float>(s2);
    if (gbeta) gbeta[c] = static_cast<float>(s1);
}

__global__ void __launch_bounds__(kStemThreads, 3) stem_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ stats,
                                                                      const float* __restrict__ coef, const float* __restrict__ gy,
                                                                      const uint8_t* __restrict__ code, float* __restrict__ gx, int H,
                                                                      int W, int Ho, int Wo, int Q, int lq, int xblocks) {
    const int C = Q * 4;
    const int q = threadIdx.x & (Q - 1), pl = threadIdx.x >> lq, PL = kStemThreads >> lq;
    int item = gridDim.x - 1 - blockIdx.x;      // reverse order: the reduce pass left the tail of x / gy in the L2
    const int xb = item % xblocks;
    item /= xblocks;
    const int k = item % Ho, n = item / Ho;
    const int m = xb * PL + pl;
    if (m >= Wo) return;
    pdl_wait();
    const float4 sc = ld4(stats + 4 * q), sh = ld4(stats + C + 4 * q);
    const float4 ca = ld4(coef + 4 * q), cb = ld4(coef + C + 4 * q), cc = ld4(coef + 2 * C + 4 * q);
    StemBlock blk;
    stem_block(blk, x + static_cast<size_t>(n) * H * W * C, gy + static_cast<size_t>(n) * Ho * Wo * C,
               code + static_cast<size_t>(n) * Ho * Wo * C, k, m, H, W, Ho, Wo, C, q, sc, sh);
    float* out = gx + (static_cast<size_t>(n) * H * W + static_cast<size_t>(2 * k) * W + 2 * m) * C + 4 * q;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (!blk.in[i]) continue;
        const float4 d = blk.dy[i], xv = blk.x[i];
        const float4 r = make_float4(fmaf(ca.x, d.x, fmaf(cb.x, xv.x, cc.x)), fmaf(ca.y, d.y, fmaf(cb.y, xv.y, cc.y)),
                                     fmaf(ca.z, d.z, fmaf(cb.z, xv.z, cc.z)), fmaf(ca.w, d.w, fmaf(cb.w, xv.w, cc.w)));
        stg_stream(reinterpret_cast<float4*>(out + (static_cast<size_t>(i >> 1) * W + (i & 1)) * C), r);
    }
}

// =======================================================================================================================
// K7b: BatchNorm2d (batch statistics) [+ residual] -> ReLU without pooling: the inner stages of the residual blocks
// (reference src/backbones/utils.py, torchvision BasicBlock / Bottleneck of the extractor).  Same statistics kernels; the
// element-wise passes see x as [n_pix, C] rows.
//   forward   y = relu(x * scale + shift [+ r])
//   backward  dy = gy where the ReLU was open: decided from y when there is a residual (y is the block's saved output),
//             recomputed from x otherwise (no extra read).  With a residual the reduce pass also WRITES dy -- it is the
//             residual's gradient -- and the apply pass reads that instead of gy and y.
// =======================================================================================================================
template <bool kRes>
__global__ void __launch_bounds__(kStemThreads) bnact_fwd_kernel(const float* __restrict__ x, const float* __restrict__ r,
                                                                 const float* __restrict__ stats, float* __restrict__ y,
                                                                 long long n_pix, int Q, int lq) {
    const int C = Q * 4;
    const int q = threadIdx.x & (Q - 1), pl = threadIdx.x >> lq, PL = kStemThreads >> lq;
    pdl_wait();
    const float4 sc = ld4(stats + 4 * q), sh = ld4(stats + C + 4 * q);
    const long long stride = static_cast<long long>(gridDim.x) * PL;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const float4* r4 = reinterpret_cast<const float4*>(r);
    float4* y4 = reinterpret_cast<float4*>(y);
    // pixels in REVERSE order (pixel n_pix-1-p): the statistics pass swept x front to back, so the tail of the tensor is
    // what the L2 still holds when this grid starts
    const long long last = n_pix - 1;
    long long p = static_cast<long long>(blockIdx.x) * PL + pl;
    for (; p + 3 * stride < n_pix; p += 4 * stride) {
        float4 v[4], w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = ldg_stream(x4 + (last - (p + u * stride)) * Q + q);
        if (kRes) {
#pragma unroll
            for (int u = 0; u < 4; ++u) w[u] = ldg_stream(r4 + (last - (p + u * stride)) * Q + q);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float4 o = make_float4(fmaf(v[u].x, sc.x, sh.x), fmaf(v[u].y, sc.y, sh.y), fmaf(v[u].z, sc.z, sh.z), fmaf(v[u].w, sc.w, sh.w));
            if (kRes) { o.x += w[u].x; o.y += w[u].y; o.z += w[u].z; o.w += w[u].w; }
            o.x = fmaxf(o.x, 0.0f); o.y = fmaxf(o.y, 0.0f); o.z = fmaxf(o.z, 0.0f); o.w = fmaxf(o.w, 0.0f);
            y4[(last - (p + u * stride)) * Q + q] = o;      // the next convolution reads it: leave it in L2
        }
    }
    for (; p < n_pix; p += stride) {
        const float4 v = ldg_stream(x4 + (last - p) * Q + q);
        float4 o = make_float4(fmaf(v.x, sc.x, sh.x), fmaf(v.y, sc.y, sh.y), fmaf(v.z, sc.z, sh.z), fmaf(v.w, sc.w, sh.w));
        if (kRes) {
            const float4 w = ldg_stream(r4 + (last - p) * Q + q);
            o.x += w.x; o.y += w.y; o.z += w.z; o.w += w.w;
        }
        o.x = fmaxf(o.x, 0.0f); o.y = fmaxf(o.y, 0.0f); o.z = fmaxf(o.z, 0.0f); o.w = fmaxf(o.w, 0.0f);
        y4[(last - p) * Q + q] = o;
    }
}

// partials[cta][0][C] = sum dy, [cta][1][C] = sum dy * xhat; kRes: gr = dy written
template <bool kRes>
__global__ void __launch_bounds__(kStemThreads) bnact_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                                        const float* __restrict__ stats, const float* __restrict__ gy,
                                                                        float* __restrict__ gr, double* __restrict__ partials,
                                                                        long long n_pix, int Q, int lq) {
    pdl_launch_dependents();
    const int C = Q * 4;
    const int q = threadIdx.x & (Q - 1), pl = threadIdx.x >> lq, PL = kStemThreads >> lq;
    const float4 sc = ld4(stats + 4 * q), sh = ld4(stats + C + 4 * q), mu = ld4(stats + 2 * C + 4 * q), is = ld4(stats + 3 * C + 4 * q);
    const float4 off = make_float4(-mu.x * is.x, -mu.y * is.y, -mu.z * is.z, -mu.w * is.w);
    const long long stride = static_cast<long long>(gridDim.x) * PL;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const float4* y4 = reinterpret_cast<const float4*>(y);
    const float4* g4 = reinterpret_cast<const float4*>(gy);
    float4* gr4 = reinterpret_cast<float4*>(gr);
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long p = static_cast<long long>(blockIdx.x) * PL + pl;
    while (p < n_pix) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
#pragma unroll 4
        for (int u = 0; u < kStemRun; ++u) {
            if (p < n_pix) {
                const float4 xv = ldg_stream(x4 + p * Q + q);
                float4 d = ldg_stream(g4 + p * Q + q);
                if (kRes) {
                    const float4 yv = ldg_stream(y4 + p * Q + q);
                    d.x = yv.x > 0.0f ? d.x : 0.0f; d.y = yv.y > 0.0f ? d.y : 0.0f;
                    d.z = yv.z > 0.0f ? d.z : 0.0f; d.w = yv.w > 0.0f ? d.w : 0.0f;
                    gr4[p * Q + q] = d;
                } else {
                    d = gate4(d, xv, sc, sh);
                }
                a.x += d.x; a.y += d.y; a.z += d.z; a.w += d.w;
                b.x = fmaf(d.x, fmaf(xv.x, is.x, off.x), b.x); b.y = fmaf(d.y, fmaf(xv.y, is.y, off.y), b.y);
                b.z = fmaf(d.z, fmaf(xv.z, is.z, off.z), b.z); b.w = fmaf(d.w, fmaf(xv.w, is.w, off.w), b.w);
            }
            p += stride;
        }
        s[0] += a.x; s[1] += a.y; s[2] += a.z; s[3] += a.w;
        s[4] += b.x; s[5] += b.y; s[6] += b.z; s[7] += b.w;
    }
    stem_cta_partials(s, partials, Q, lq, C);
}

// gx = a dy + b x + c; kRes: dy is read back from gr (written by the reduce pass), else gy gated by the recomputed ReLU
template <bool kRes>
__global__ void __launch_bounds__(kStemThreads) bnact_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ stats,
                                                                       const float* __restrict__ coef, const float* __restrict__ g,
                                                                       float* __restrict__ gx, long long n_pix, int Q, int lq) {
    const int C = Q * 4;
    const int q = threadIdx.x & (Q - 1), pl = threadIdx.x >> lq, PL = kStemThreads >> lq;
    pdl_wait();
    const float4 sc = ld4(stats + 4 * q), sh = ld4(stats + C + 4 * q);
    const float4 ca = ld4(coef + 4 * q), cb = ld4(coef + C + 4 * q), cc = ld4(coef + 2 * C + 4 * q);
    const long long stride = static_cast<long long>(gridDim.x) * PL;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* o4 = reinterpret_cast<float4*>(gx);
    const long long last = n_pix - 1;       // reverse order, as in the forward: the reduce pass left the tail in the L2
    long long p = static_cast<long long>(blockIdx.x) * PL + pl;
    for (; p < n_pix; p += 2 * stride) {
        const bool two = p + stride < n_pix;
        const long long i0 = (last - p) * Q + q, i1 = (last - p - stride) * Q + q;
        const float4 x0 = ldg_stream(x4 + i0), g0 = ldg_stream(g4 + i0);
        float4 x1 = x0, g1 = g0;
        if (two) { x1 = ldg_stream(x4 + i1); g1 = ldg_stream(g4 + i1); }
        const float4 d0 = kRes ? g0 : gate4(g0, x0, sc, sh), d1 = kRes ? g1 : gate4(g1, x1, sc, sh);
        stg_stream(o4 + i0, make_float4(fmaf(ca.x, d0.x, fmaf(cb.x, x0.x, cc.x)), fmaf(ca.y, d0.y, fmaf(cb.y, x0.y, cc.y)),
                                               fmaf(ca.z, d0.z, fmaf(cb.z, x0.z, cc.z)), fmaf(ca.w, d0.w, fmaf(cb.w, x0.w, cc.w))));
        if (two)
            stg_stream(o4 + i1, make_float4(fmaf(ca.x, d1.x, fmaf(cb.x, x1.x, cc.x)), fmaf(ca.y, d1.y, fmaf(cb.y, x1.y, cc.y)),
                                                              fmaf(ca.z, d1.z, fmaf(cb.z, x1.z, cc.z)), fmaf(ca.w, d1.w, fmaf(cb.w, x1.w, cc.w))));
    }
}

// ---- K7c: the end of a block whose skip path has its own BatchNorm: y = relu(bn_a(a) + bn_b(b)) ------------------------------
// (ResNet34ConvBlock with a projection, the up-sampling blocks: src/backbones/utils.py).  Through K7b this is the skip's BatchNorm
// in cuDNN (which writes r) plus relu(bn(a) + r); here neither r nor its gradient exists: the forward reads a and b and
// writes y, the backward reads a, b, gy, y twice and writes both input gradients.
__global__ void __launch_bounds__(kStemThreads) bnact2_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                  const float* __restrict__ stats_a, const float* __restrict__ stats_b,
                                                                  float* __restrict__ y, long long n_pix, int Q, int lq) {
    const int C = Q * 4;
    const int q = threadIdx.x & (Q - 1), pl = threadIdx.x >> lq, PL = kStemThreads >> lq;
    pdl_wait();
    const float4 sa = ld4(stats_a + 4 * q), ha = ld4(stats_a + C + 4 * q), sb = ld4(stats_b + 4 * q), hb = ld4(stats_b + C + 4 * q);
    const float4 sh = make_float4(ha.x + hb.x, ha.y + hb.y, ha.z + hb.z, ha.w + hb.w);
    const long long stride = static_cast<long long>(gridDim.x) * PL, last = n_pix - 1;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    float4* y4 = reinterpret_cast<float4*>(y);
    for (long long p = static_cast<long long>(blockIdx.x) * PL + pl; p < n_pix; p += 2 * stride) {
        const bool two = p + stride < n_pix;
        const long long i0 = (last - p) * Q + q, i1 = (last - p - stride) * Q + q;
        const float4 u0 = ldg_stream(a4 + i0), v0 = ldg_stream(b4 + i0);
        float4 u1 = u0, v1 = v0;
        if (two) { u1 = ldg_stream(a4 + i1); v1 = ldg_stream(b4 + i1); }
        y4[i0] = make_float4(fmaxf(fmaf(u0.x, sa.x, fmaf(v0.x, sb.x, sh.x)), 0.0f), fmaxf(fmaf(u0.y, sa.y, fmaf(v0.y, sb.y, sh.y)), 0.0f),
                             fmaxf(fmaf(u0.z, sa.z, fmaf(v0.z, sb.z, sh.z)), 0.0f), fmaxf(fmaf(u0.w, sa.w, fmaf(v0.w, sb.w, sh.w)), 0.0f));
        if (two)
            y4[i1] = make_float4(fmaxf(fmaf(u1.x, sa.x, fmaf(v1.x, sb.x, sh.x)), 0.0f), fmaxf(fmaf(u1.y, sa.y, fmaf(v1.y, sb.y, sh.y)), 0.0f),
                                 fmaxf(fmaf(u1.z, sa.z, fmaf(v1.z, sb.z, sh.z)), 0.0f), fmaxf(fmaf(u1.w, sa.w, fmaf(v1.w, sb.w, sh.w)), 0.0f));
    }
}

// both BatchNorms' sums in one pass: partials_a / partials_b [2][C][grid]
__global__ void __launch_bounds__(kStemThreads, 2) bnact2_bwd_reduce_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                            const float* __restrict__ y, const float* __restrict__ stats_a,
                                                                            const float* __restrict__ stats_b, const float* __restrict__ gy,
                                                                            double* __restrict__ partials_a, double* __restrict__ partials_b,
                                                                            long long n_pix, int Q, int lq) {
    pdl_launch_dependents();
    const int C = Q * 4;
    const int q = threadIdx.x & (Q - 1), pl = threadIdx.x >> lq, PL = kStemThreads >> lq;
    const float4 ma = ld4(stats_a + 2 * C + 4 * q), ia = ld4(stats_a + 3 * C + 4 * q), mb = ld4(stats_b + 2 * C + 4 * q), ib = ld4(stats_b + 3 * C + 4 * q);
    const float4 oa = make_float4(-ma.x * ia.x, -ma.y * ia.y, -ma.z * ia.z, -ma.w * ia.w);
    const float4 ob = make_float4(-mb.x * ib.x, -mb.y * ib.y, -mb.z * ib.z, -mb.w * ib.w);
    const long long stride = static_cast<long long>(gridDim.x) * PL;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    const float4* y4 = reinterpret_cast<const float4*>(y);
    const float4* g4 = reinterpret_cast<const float4*>(gy);
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t[4] = {0, 0, 0, 0};      // s: sum dy, sum dy ahat; t: sum dy bhat (sum dy is shared)
    long long p = static_cast<long long>(blockIdx.x) * PL + pl;
    while (p < n_pix) {
        float4 d0 = make_float4(0.f, 0.f, 0.f, 0.f), da = d0, db = d0;
#pragma unroll 4
        for (int u = 0; u < kStemRun; ++u) {
            if (p < n_pix) {
                const float4 av = ldg_stream(a4 + p * Q + q), bv = ldg_stream(b4 + p * Q + q), yv = ldg_stream(y4 + p * Q + q);
                float4 d = ldg_stream(g4 + p * Q + q);
                d.x = yv.x > 0.0f ? d.x : 0.0f; d.y = yv.y > 0.0f ? d.y : 0.0f;
                d.z = yv.z > 0.0f ? d.z : 0.0f; d.w = yv.w > 0.0f ? d.w : 0.0f;
                d0.x += d.x; d0.y += d.y; d0.z += d.z; d0.w += d.w;
                da.x = fmaf(d.x, fmaf(av.x, ia.x, oa.x), da.x); da.y = fmaf(d.y, fmaf(av.y, ia.y, oa.y), da.y);
                da.z = fmaf(d.z, fmaf(av.z, ia.z, oa.z), da.z); da.w = fmaf(d.w, fmaf(av.w, ia.w, oa.w), da.w);
                db.x = fmaf(d.x, fmaf(bv.x, ib.x, ob.x), db.x); db.y = fmaf(d.y, fmaf(bv.y, ib.y, ob.y), db.y);
                db.z = fmaf(d.z, fmaf(bv.z, ib.z, ob.z), db.z); db.w = fmaf(d.w, fmaf(bv.w, ib.w, ob.w), db.w);
            }
            p += stride;
        }
        s[0] += d0.x; s[1] += d0.y; s[2] += d0.z; s[3] += d0.w;
        s[4] += da.x; s[5] += da.y; s[6] += da.z; s[7] += da.w;
        t[0] += db.x; t[1] += db.y; t[2] += db.z; t[3] += db.w;
    }
    stem_cta_partials(s, partials_a, Q, lq, C);
    __syncthreads();       // the shared scratch of stem_cta_partials is reused
    const double sb2[8] = {s[0], s[1], s[2], s[3], t[0], t[1], t[2], t[3]};
    stem_cta_partials(sb2, partials_b, Q, lq, C);
}

// ga = ca0 dy + ca1 a + ca2, gb = cb0 dy + cb1 b + cb2, dy = gy where y > 0
__global__ void __launch_bounds__(kStemThreads) bnact2_bwd_apply_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                        const float* __restrict__ y, const float* __restrict__ gy,
                                                                        const float* __restrict__ coef_a, const float* __restrict__ coef_b,
                                                                        float* __restrict__ ga, float* __restrict__ gb, long long n_pix,
                                                                        int Q, int lq) {
    const int C = Q * 4;
    const int q = threadIdx.x & (Q - 1), pl = threadIdx.x >> lq, PL = kStemThreads >> lq;
    pdl_wait();
    const float4 a0 = ld4(coef_a + 4 * q), a1 = ld4(coef_a + C + 4 * q), a2 = ld4(coef_a + 2 * C + 4 * q);
    const float4 b0 = ld4(coef_b + 4 * q), b1 = ld4(coef_b + C + 4 * q), b2 = ld4(coef_b + 2 * C + 4 * q);
    const long long stride = static_cast<long long>(gridDim.x) * PL, last = n_pix - 1;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    const float4* y4 = reinterpret_cast<const float4*>(y);
    const float4* g4 = reinterpret_cast<const float4*>(gy);
    float4* oa4 = reinterpret_cast<float4*>(ga);
    float4* ob4 = reinterpret_cast<float4*>(gb);
    for (long long p = static_cast<long long>(blockIdx.x) * PL + pl; p < n_pix; p += stride) {
        const long long i = (last - p) * Q + q;
        const float4 av = ldg_stream(a4 + i), bv = ldg_stream(b4 + i), yv = ldg_stream(y4 + i);
        float4 d = ldg_stream(g4 + i);
        d.x = yv.x > 0.0f ? d.x : 0.0f; d.y = yv.y > 0.0f ? d.y : 0.0f;
        d.z = yv.z > 0.0f ? d.z : 0.0f; d.w = yv.w > 0.0f ? d.w : 0.0f;
        stg_stream(oa4 + i, make_float4(fmaf(a0.x, d.x, fmaf(a1.x, av.x, a2.x)), fmaf(a0.y, d.y, fmaf(a1.y, av.y, a2.y)),
                                        fmaf(a0.z, d.z, fmaf(a1.z, av.z, a2.z)), fmaf(a0.w, d.w, fmaf(a1.w, av.w, a2.w))));
        stg_stream(ob4 + i, make_float4(fmaf(b0.x, d.x, fmaf(b1.x, bv.x, b2.x)), fmaf(b0.y, d.y, fmaf(b1.y, bv.y, b2.y)),
                                        fmaf(b0.z, d.z, fmaf(b1.z, bv.z, b2.z)), fmaf(b0.w, d.w, fmaf(b1.w, bv.w, b2.w))));
    }
}

// ---- host side ----------------------------------------------------------------------------------------------------------
struct StemGeo {
    int Q, lq, PL, Ho, Wo, xblocks;
};
inline int stem_geo(int N, int H, int W, int C, StemGeo& g) {
    if (N <= 0 || H <= 0 || W <= 0 || C < 4) return BH_E_SHAPE;
    if (C % 4 != 0 || C > 4 * kStemThreads) return BH_E_UNSUPPORTED;
    g.Q = C / 4;
    if (g.Q & (g.Q - 1)) return BH_E_UNSUPPORTED;          // channel quads per pixel must divide the CTA: C = 4 .. 1024, a power of two
    g.lq = 0;
    while ((1 << g.lq) < g.Q) ++g.lq;
    g.PL = kStemThreads / g.Q;
    g.Ho = (H - 1) / 2 + 1;
    g.Wo = (W - 1) / 2 + 1;
    g.xblocks = (g.Wo + g.PL - 1) / g.PL;
    if (static_cast<long long>(N) * g.Ho * g.xblocks > 0x7fffffffll) return BH_E_SHAPE;
    return BH_OK;
}
inline int stem_reduce_grid() { return kNumSMs * 4; }   // one wave: every CTA resident when its dependents are scheduled
inline size_t stem_ws_doubles(int C) { return static_cast<size_t>(stem_reduce_grid()) * 2 * C; }

}  // namespace bh

extern "C" int bh_stem_supported(int C) {
    bh::StemGeo g;
    return bh::stem_geo(1, 2, 2, C, g) == BH_OK ? 1 : 0;
}

extern "C" size_t bh_stem_workspace_bytes(int C) {
    if (C <= 0) return 0;
    return bh::stem_ws_doubles(C) * sizeof(double) + static_cast<size_t>(3) * C * sizeof(float);
}

extern "C" int bh_stem_fwd(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var,
                           float momentum, float eps, float* y, uint8_t* code, float* stats, void* ws, size_t ws_bytes, int N,
                           int H, int W, int C, bh_stream_t stream) {
    using namespace bh;
    if (!x || !y || !stats || !ws) return BH_E_NULL;
    StemGeo g;
    const int rc = stem_geo(N, H, W, C, g);
    if (rc != BH_OK) return rc;
    if (!aligned16(x) || !aligned16(y) || !aligned16(stats) || !aligned16(ws) || (code && (reinterpret_cast<uintptr_t>(code) & 3u)))
        return BH_E_ALIGN;
    if (ws_bytes < bh_stem_workspace_bytes(C)) return BH_E_WORKSPACE;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const long long n_pix = static_cast<long long>(N) * H * W;
    const long long want = (n_pix + static_cast<long long>(g.PL) * kStemRun - 1) / (static_cast<long long>(g.PL) * kStemRun);
    const int G = static_cast<int>(want < stem_reduce_grid() ? want : stem_reduce_grid());
    double* partials = static_cast<double*>(ws);
    bn_stats_kernel<<<G, kStemThreads, 0, s>>>(x, partials, n_pix, g.Q, g.lq);
    int st = launch_status();
    if (st != BH_OK) return st;
    st = launch_dependent(bn_finalize_kernel, dim3((C + 7) / 8), dim3(kStemThreads), s, partials, G, C, static_cast<double>(n_pix), gamma,
                          beta, running_mean, running_var, momentum, eps, stats);
    if (st != BH_OK) return st;
    const int strips = (g.Ho + kStemRows - 1) / kStemRows;
    const long long items = static_cast<long long>(N) * strips * g.xblocks;
    const dim3 grid(static_cast<unsigned>(items)), block(kStemThreads);
    if (code) return launch_dependent(stem_pool_fwd_kernel<true>, grid, block, s, x, stats, y, code, H, W, g.Ho, g.Wo, g.Q, g.lq, strips, g.xblocks);
    return launch_dependent(stem_pool_fwd_kernel<false>, grid, block, s, x, stats, y, code, H, W, g.Ho, g.Wo, g.Q, g.lq, strips, g.xblocks);
}

extern "C" int bh_stem_bwd(const float* x, const float* stats, const uint8_t* code, const float* gy, float* gx, float* ggamma,
                           float* gbeta, void* ws, size_t ws_bytes, int N, int H, int W, int C, bh_stream_t stream) {
    using namespace bh;
    if (!x || !stats || !code || !gy || !gx || !ws) return BH_E_NULL;
    StemGeo g;
    const int rc = stem_geo(N, H, W, C, g);
    if (rc != BH_OK) return rc;
    if (!aligned16(x) || !aligned16(gy) || !aligned16(gx) || !aligned16(stats) || !aligned16(ws) || (reinterpret_cast<uintptr_t>(code) & 3u))
        return BH_E_ALIGN;
    if (ws_bytes < bh_stem_workspace_bytes(C)) return BH_E_WORKSPACE;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const long long items = static_cast<long long>(N) * g.Ho * g.xblocks;
    const long long want = (items + 3) / 4;
    const int G = static_cast<int>(want < stem_reduce_grid() ? want : stem_reduce_grid());
    double* partials = static_cast<double*>(ws);
    float* coef = reinterpret_cast<float*>(partials + stem_ws_doubles(C));
    stem_bwd_reduce_kernel<<<G, kStemThreads, 0, s>>>(x, stats, gy, code, partials, N, H, W, g.Ho, g.Wo, g.Q, g.lq, g.xblocks);
    int st = launch_status();
    if (st != BH_OK) return st;
    st = launch_dependent(stem_bwd_finalize_kernel, dim3((C + 7) / 8), dim3(kStemThreads), s, partials, G, C, static_cast<double>(N) * H * W,
                          stats, coef, ggamma, gbeta);
    if (st != BH_OK) return st;
    return launch_dependent(stem_bwd_apply_kernel, dim3(static_cast<unsigned>(items)), dim3(kStemThreads), s, x, stats, coef, gy, code, gx, H, W,
                            g.Ho, g.Wo, g.Q, g.lq, g.xblocks);
}

// ---- K7b entry points ---------------------------------------------------------------------------------------------------
namespace bh {
inline int bnact_geo(long long n_pix, int C, int& Q, int& lq, int& PL) {
    if (n_pix <= 0 || C < 4) return BH_E_SHAPE;
    if (C % 4 != 0 || C > 4 * kStemThreads) return BH_E_UNSUPPORTED;
    Q = C / 4;
    if (Q & (Q - 1)) return BH_E_UNSUPPORTED;
    lq = 0;
    while ((1 << lq) < Q) ++lq;
    PL = kStemThreads / Q;
    return BH_OK;
}
inline int bnact_stream_grid(long long n_pix, int PL, int per_thread) {
    const long long want = (n_pix + static_cast<long long>(PL) * per_thread - 1) / (static_cast<long long>(PL) * per_thread);
    const long long cap = static_cast<long long>(kNumSMs) * 16;
    return static_cast<int>(want < cap ? want : cap);
}
}  // namespace bh

extern "C" int bh_bnact_fwd(const float* x, const float* residual, const float* gamma, const float* beta, float* running_mean,
                            float* running_var, float momentum, float eps, float* y, float* stats, void* ws, size_t ws_bytes,
                            long long n_pix, int C, bh_stream_t stream) {
    using namespace bh;
    if (!x || !y || !stats || !ws) return BH_E_NULL;
    int Q, lq, PL;
    const int rc = bnact_geo(n_pix, C, Q, lq, PL);
    if (rc != BH_OK) return rc;
    if (!aligned16(x) || !aligned16(y) || !aligned16(stats) || !aligned16(ws) || (residual && !aligned16(residual))) return BH_E_ALIGN;
    if (ws_bytes < bh_stem_workspace_bytes(C)) return BH_E_WORKSPACE;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const long long want = (n_pix + static_cast<long long>(PL) * kStemRun - 1) / (static_cast<long long>(PL) * kStemRun);
    const int G = static_cast<int>(want < stem_reduce_grid() ? want : stem_reduce_grid());
    double* partials = static_cast<double*>(ws);
    bn_stats_kernel<<<G, kStemThreads, 0, s>>>(x, partials, n_pix, Q, lq);
    int st = launch_status();
    if (st != BH_OK) return st;
    st = launch_dependent(bn_finalize_kernel, dim3((C + 7) / 8), dim3(kStemThreads), s, partials, G, C, static_cast<double>(n_pix), gamma,
                          beta, running_mean, running_var, momentum, eps, stats);
    if (st != BH_OK) return st;
    const dim3 grid(bnact_stream_grid(n_pix, PL, 4)), block(kStemThreads);
    if (residual) return launch_dependent(bnact_fwd_kernel<true>, grid, block, s, x, residual, stats, y, n_pix, Q, lq);
    return launch_dependent(bnact_fwd_kernel<false>, grid, block, s, x, residual, stats, y, n_pix, Q, lq);
}

extern "C" int bh_bnact_bwd(const float* x, const float* y, const float* stats, const float* gy, float* gx, float* gresidual,
                            float* ggamma, float* gbeta, void* ws, size_t ws_bytes, long long n_pix, int C, bh_stream_t stream) {
    using namespace bh;
    if (!x || !stats || !gy || !gx || !ws) return BH_E_NULL;
    if (gresidual && !y) return BH_E_NULL;      // with a residual the ReLU decision comes from the saved output
    int Q, lq, PL;
    const int rc = bnact_geo(n_pix, C, Q, lq, PL);
    if (rc != BH_OK) return rc;
    if (!aligned16(x) || !aligned16(gy) || !aligned16(gx) || !aligned16(stats) || !aligned16(ws) || (y && !aligned16(y)) ||
        (gresidual && !aligned16(gresidual)))
        return BH_E_ALIGN;
    if (ws_bytes < bh_stem_workspace_bytes(C)) return BH_E_WORKSPACE;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const long long want = (n_pix + static_cast<long long>(PL) * kStemRun - 1) / (static_cast<long long>(PL) * kStemRun);
    const int G = static_cast<int>(want < stem_reduce_grid() ? want : stem_reduce_grid());
    double* partials = static_cast<double*>(ws);
    float* coef = reinterpret_cast<float*>(partials + stem_ws_doubles(C));
    if (gresidual) bnact_bwd_reduce_kernel<true><<<G, kStemThreads, 0, s>>>(x, y, stats, gy, gresidual, partials, n_pix, Q, lq);
    else bnact_bwd_reduce_kernel<false><<<G, kStemThreads, 0, s>>>(x, nullptr, stats, gy, nullptr, partials, n_pix, Q, lq);
    int st = launch_status();
    if (st != BH_OK) return st;
    st = launch_dependent(stem_bwd_finalize_kernel, dim3((C + 7) / 8), dim3(kStemThreads), s, partials, G, C, static_cast<double>(n_pix), stats,
                          coef, ggamma, gbeta);
    if (st != BH_OK) return st;
    const dim3 grid(bnact_stream_grid(n_pix, PL, 4)), block(kStemThreads);
    if (gresidual) return launch_dependent(bnact_bwd_apply_kernel<true>, grid, block, s, x, stats, coef, gresidual, gx, n_pix, Q, lq);
    return launch_dependent(bnact_bwd_apply_kernel<false>, grid, block, s, x, stats, coef, gy, gx, n_pix, Q, lq);
}

// ---- K7c entry points: y = relu(bn_a(a) + bn_b(b)) -------------------------------------------------------------------------
// ws: 2 x bh_stem_workspace_bytes(C) (one region per BatchNorm)
extern "C" int bh_bnact2_fwd(const float* a, const float* b, const float* gamma_a, const float* beta_a, float* rmean_a, float* rvar_a,
                             float momentum_a, float eps_a, const float* gamma_b, const float* beta_b, float* rmean_b, float* rvar_b,
                             float momentum_b, float eps_b, float* y, float* stats_a, float* stats_b, void* ws, size_t ws_bytes,
                             long long n_pix, int C, bh_stream_t stream) {
    using namespace bh;
    if (!a || !b || !y || !stats_a || !stats_b || !ws) return BH_E_NULL;
    int Q, lq, PL;
    const int rc = bnact_geo(n_pix, C, Q, lq, PL);
    if (rc != BH_OK) return rc;
    if (!aligned16(a) || !aligned16(b) || !aligned16(y) || !aligned16(stats_a) || !aligned16(stats_b) || !aligned16(ws)) return BH_E_ALIGN;
    const size_t region = bh_stem_workspace_bytes(C);
    if (ws_bytes < 2 * region) return BH_E_WORKSPACE;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const long long want = (n_pix + static_cast<long long>(PL) * kStemRun - 1) / (static_cast<long long>(PL) * kStemRun);
    const int G = static_cast<int>(want < stem_reduce_grid() ? want : stem_reduce_grid());
    double* pa = static_cast<double*>(ws);
    double* pb = reinterpret_cast<double*>(static_cast<char*>(ws) + region);
    bn_stats_kernel<<<G, kStemThreads, 0, s>>>(a, pa, n_pix, Q, lq);
    int st = launch_status();
    if (st != BH_OK) return st;
    st = launch_dependent(bn_finalize_kernel, dim3((C + 7) / 8), dim3(kStemThreads), s, pa, G, C, static_cast<double>(n_pix), gamma_a, beta_a,
                          rmean_a, rvar_a, momentum_a, eps_a, stats_a);
    if (st != BH_OK) return st;
    bn_stats_kernel<<<G, kStemThreads, 0, s>>>(b, pb, n_pix, Q, lq);
    st = launch_status();
    if (st != BH_OK) return st;
    st = launch_dependent(bn_finalize_kernel, dim3((C + 7) / 8), dim3(kStemThreads), s, pb, G, C, static_cast<double>(n_pix), gamma_b, beta_b,
                          rmean_b, rvar_b, momentum_b, eps_b, stats_b);
    if (st != BH_OK) return st;
    return launch_dependent(bnact2_fwd_kernel, dim3(bnact_stream_grid(n_pix, PL, 4)), dim3(kStemThreads), s, a, b, stats_a, stats_b, y, n_pix, Q, lq);
}

extern "C" int bh_bnact2_bwd(const float* a, const float* b, const float* y, const float* stats_a, const float* stats_b, const float* gy,
                             float* ga, float* gb, float* ggamma_a, float* gbeta_a, float* ggamma_b, float* gbeta_b, void* ws,
                             size_t ws_bytes, long long n_pix, int C, bh_stream_t stream) {
    using namespace bh;
    if (!a || !b || !y || !stats_a || !stats_b || !gy || !ga || !gb || !ws) return BH_E_NULL;
    int Q, lq, PL;
    const int rc = bnact_geo(n_pix, C, Q, lq, PL);
    if (rc != BH_OK) return rc;
    if (!aligned16(a) || !aligned16(b) || !aligned16(y) || !aligned16(gy) || !aligned16(ga) || !aligned16(gb) || !aligned16(stats_a) ||
        !aligned16(stats_b) || !aligned16(ws))
        return BH_E_ALIGN;
    const size_t region = bh_stem_workspace_bytes(C);
    if (ws_bytes < 2 * region) return BH_E_WORKSPACE;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const long long want = (n_pix + static_cast<long long>(PL) * kStemRun - 1) / (static_cast<long long>(PL) * kStemRun);
    const int cap = kNumSMs * 2;          // the reduce kernel holds two CTAs per SM: one wave
    const int G = static_cast<int>(want < cap ? want : cap);
    double* pa = static_cast<double*>(ws);
    double* pb = reinterpret_cast<double*>(static_cast<char*>(ws) + region);
    float* ca = reinterpret_cast<float*>(pa + stem_ws_doubles(C));
    float* cb = reinterpret_cast<float*>(pb + stem_ws_doubles(C));
    bnact2_bwd_reduce_kernel<<<G, kStemThreads, 0, s>>>(a, b, y, stats_a, stats_b, gy, pa, pb, n_pix, Q, lq);
    int st = launch_status();
    if (st != BH_OK) return st;
    st = launch_dependent(stem_bwd_finalize_kernel, dim3((C + 7) / 8), dim3(kStemThreads), s, pa, G, C, static_cast<double>(n_pix), stats_a, ca,
                          ggamma_a, gbeta_a);
    if (st != BH_OK) return st;
    st = launch_dependent(stem_bwd_finalize_kernel, dim3((C + 7) / 8), dim3(kStemThreads), s, pb, G, C, static_cast<double>(n_pix), stats_b, cb,
                          ggamma_b, gbeta_b);
    if (st != BH_OK) return st;
    return launch_dependent(bnact2_bwd_apply_kernel, dim3(bnact_stream_grid(n_pix, PL, 2)), dim3(kStemThreads), s, a, b, y, gy, ca, cb, ga, gb,
                            n_pix, Q, lq);
}

// =======================================================================================================================
// K8: the per-channel bias of a transposed convolution, channels-last.  cuDNN's transposed convolution (a data-gradient
// kernel) has no bias epilogue: ATen adds the bias with a strided broadcast kernel in a pass of its own and reduces its
// gradient with a generic reduction -- 2.4 ms of the B = 256 training step for the four up-sampling blocks of the Zeng
// backbone (reference src/backbones/utils.py:65-66, nn.ConvTranspose2d with its default bias=True).
//   bh_bias_add    y[p, c] += bias[c] in place: one read and one write of the tensor, 16-byte accesses
//   bh_bias_grad   gbias[c] = sum_p gy[p, c]: the statistics pass of K7 (float32 runs of 16 pixels, float64 above, fixed
//                  order => bit reproducible) followed by a one-warp-per-channel finish
// =======================================================================================================================
namespace bh {

__global__ void __launch_bounds__(kStemThreads) bias_add_kernel(float* __restrict__ y, const float* __restrict__ bias, long long n_pix,
                                                                int Q, int lq) {
    const int q = threadIdx.x & (Q - 1), pl = threadIdx.x >> lq, PL = kStemThreads >> lq;
    const float4 b = ld4(bias + 4 * q);
    const long long stride = static_cast<long long>(gridDim.x) * PL;
    float4* y4 = reinterpret_cast<float4*>(y);
    long long p = static_cast<long long>(blockIdx.x) * PL + pl;
    // every element is read and written by the same thread; plain (coherent) loads: the tensor is not read-only here
    for (; p + 3 * stride < n_pix; p += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = y4[(p + u * stride) * Q + q];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            v[u].x += b.x; v[u].y += b.y; v[u].z += b.z; v[u].w += b.w;
            y4[(p + u * stride) * Q + q] = v[u];      // the next convolution reads it: leave it in L2
        }
    }
    for (; p < n_pix; p += stride) {
        float4 v = y4[p * Q + q];
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        y4[p * Q + q] = v;
    }
}

// one warp per channel over the statistics pass' partials[0][C][G] (the sums; the sums of squares behind them are ignored)
__global__ void __launch_bounds__(kStemThreads) bias_grad_finalize_kernel(const double* __restrict__ partials, int G, int C,
                                                                          float* __restrict__ gbias) {
    pdl_wait();   // the statistics grid has completed
    const int c = blockIdx.x * (kStemThreads / 32) + (threadIdx.x >> 5);
    if (c >= C) return;
    double s, ss;
    stem_sum_partials(partials, G, C, c, s, ss);
    if ((threadIdx.x & 31) == 0) gbias[c] = static_cast<float>(s);
}

}  // namespace bh

extern "C" int bh_bias_add(float* y, const float* bias, long long n_pix, int C, bh_stream_t stream) {
    using namespace bh;
    if (!y || !bias) return BH_E_NULL;
    int Q, lq, PL;
    const int rc = bnact_geo(n_pix, C, Q, lq, PL);
    if (rc != BH_OK) return rc;
    if (!aligned16(y) || !aligned16(bias)) return BH_E_ALIGN;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    bias_add_kernel<<<bnact_stream_grid(n_pix, PL, 4), kStemThreads, 0, s>>>(y, bias, n_pix, Q, lq);
    return launch_status();
}

extern "C" int bh_bias_grad(const float* gy, float* gbias, void* ws, size_t ws_bytes, long long n_pix, int C, bh_stream_t stream) {
    using namespace bh;
    if (!gy || !gbias || !ws) return BH_E_NULL;
    int Q, lq, PL;
    const int rc = bnact_geo(n_pix, C, Q, lq, PL);
    if (rc != BH_OK) return rc;
    if (!aligned16(gy) || !aligned16(ws)) return BH_E_ALIGN;
    if (ws_bytes < bh_stem_workspace_bytes(C)) return BH_E_WORKSPACE;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const long long want = (n_pix + static_cast<long long>(PL) * kStemRun - 1) / (static_cast<long long>(PL) * kStemRun);
    const int G = static_cast<int>(want < stem_reduce_grid() ? want : stem_reduce_grid());
    double* partials = static_cast<double*>(ws);
    bn_stats_kernel<<<G, kStemThreads, 0, s>>>(gy, partials, n_pix, Q, lq);
    const int st = launch_status();
    if (st != BH_OK) return st;
    return launch_dependent(bias_grad_finalize_kernel, dim3((C + 7) / 8), dim3(kStemThreads), s, partials, G, C, gbias);
}
