// K6: the perspective-field head of the Zeng backbone as one per-pixel kernel.
//
// Reference: src/backbones/Rethinking.py:144-147 -- layer8 = Conv2d(16,128,1) -> BatchNorm2d(128) -> ReLU -> Conv2d(128,2,1)
// at full patch resolution.  Through ATen/cuDNN that is 21 passes over a [B,128,P,P] tensor per backbone pass (2.1 GB at
// B = 256, P = 128): the convolution writes it, BatchNorm reads it twice and rewrites it, ReLU rewrites it, the second
// convolution reads it, and the backward does all of that again with two operands.  None of it has to exist:
//   * the hidden activation is LINEAR in the 16 input channels, so the batch statistics BatchNorm needs are a function of
//     the first and second moments of the INPUT (16 + 16x16 numbers): mean_y = W1 m + b1, var_y = diag(W1 C W1^T);
//   * with the statistics known, BatchNorm folds into the first convolution (W1' = s W1, b1' = s (b1 - mean_y) + beta,
//     s = gamma / sqrt(var_y + eps)) and a pixel's two outputs are  W2 relu(W1' x + b1') + b2  -- 2.3 kFMA on 64 bytes.
// Kernels (all HBM traffic is the [N,16] input, its gradient, and the [B,2,P,P] field and its gradient):
//   moments_kernel      x -> per-CTA partial sums of x_i and x_i x_k (fp32 per 128-pixel tile, fp64 across tiles)
//   fieldhead_fwd       thread per pixel, folded weights broadcast from shared memory
//   fieldhead_bwd       32-pixel tiles: phase 1 (lane = pixel, warp = quarter of the hidden units) recomputes the hidden
//                       layer into a shared [hidden][pixel] tile and forms d/dx; phase 2 (thread = hidden unit) accumulates
//                       the weight gradients in registers across all tiles of the CTA; per-CTA partials are summed by the
//                       caller in a fixed order (no atomics: bit-reproducible)
//   affine_acc_kernel   gx += a + M x : the moments' adjoint (how the batch statistics feed back into the input)
// The fold itself and its adjoint are a few 128x16 operations on the host side (bihome_b200/functional.py, torch ops).
// x is channels-last ([N,16] pixel-major); the field and its gradient are planar [B,2,P,P] -- what K4 (dltn.cu) reads.
#ifdef BH_HOST_EMULATION   // tests/emu/fieldhead_emu.cpp: the kernels below compiled for the host, CTA threads as pthreads
#include "cuda_emu.h"
#else
#include "bh_common.cuh"
#endif

namespace bh {

constexpr int kFhThreads = 128;      // == hidden units of the shipped head: phase 2 gives every hidden unit one thread
constexpr int kFhTile = 32;          // pixels per backward tile (lane = pixel in phase 1)
constexpr int kFhPad = kFhTile + 1;  // row pitch of the [hidden][pixel] tiles: conflict-free in both phases
constexpr int kMomTile = 128;        // pixels per moments tile
constexpr int kMomThreads = 288;     // 16 sums + 16 x 16 second moments = 272 statistics, rounded up to whole warps

template <int CIN>
__device__ __forceinline__ void load_pixel(const float* __restrict__ x, long long p, float (&v)[CIN]) {
    const float4* src = reinterpret_cast<const float4*>(x + p * CIN);
#pragma unroll
    for (int q = 0; q < CIN / 4; ++q) {
        const float4 t = ldg_stream(src + q);
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
}

// ---- moments ------------------------------------------------------------------------------------------------------
// partials[cta][0..CIN) = sum_p x_i ; partials[cta][CIN + i*CIN + k] = sum_p x_i x_k (the full symmetric matrix: the
// redundant half costs nothing in a kernel that waits for memory, and the host side needs no index juggling)
template <int CIN>
__global__ void __launch_bounds__(kMomThreads) moments_kernel(const float* __restrict__ x, double* __restrict__ partials,
                                                             long long n_pix) {
    constexpr int kStats = CIN + CIN * CIN;
    static_assert(kStats <= kMomThreads, "one thread per statistic");
    __shared__ __align__(16) float sx[kMomTile * CIN];
    const int t = threadIdx.x;
    const bool is_sum = t < CIN, active = t < kStats;
    const int i = is_sum ? t : (t - CIN) / CIN, k = is_sum ? t : (t - CIN) % CIN;
    double acc = 0.0;
    const long long n_tiles = (n_pix + kMomTile - 1) / kMomTile;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long p0 = tile * kMomTile;
        const long long rem = n_pix - p0;
        const int np = rem < kMomTile ? static_cast<int>(rem) : kMomTile;
        __syncthreads();   // the previous tile is still being read
        for (int q = t; q < kMomTile * CIN / 4; q += kMomThreads) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q * 4 < np * CIN) v = ldg_stream(reinterpret_cast<const float4*>(x + p0 * CIN) + q);
            reinterpret_cast<float4*>(sx)[q] = v;
        }
        __syncthreads();
        if (active) {
            float s = 0.0f;
            if (is_sum) {
#pragma unroll 8
                for (int p = 0; p < kMomTile; ++p) s += sx[p * CIN + i];
            } else {
#pragma unroll 8
                for (int p = 0; p < kMomTile; ++p) s = fmaf(sx[p * CIN + i], sx[p * CIN + k], s);
            }
            acc += static_cast<double>(s);
        }
    }
    if (active) partials[static_cast<long long>(blockIdx.x) * kStats + t] = acc;
}

// ---- forward ------------------------------------------------------------------------------------------------------
template <int CIN, int HID>
__global__ void __launch_bounds__(256) fieldhead_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W1,
                                                           const float* __restrict__ b1, const float* __restrict__ W2,
                                                           const float* __restrict__ b2, float* __restrict__ out,
                                                           long long n_pix, int HW) {
    __shared__ __align__(16) float sW1[HID * CIN];
    __shared__ float sb1[HID];
    __shared__ float sW2[2 * HID];
    for (int q = threadIdx.x; q < HID * CIN; q += blockDim.x) sW1[q] = W1[q];
    for (int q = threadIdx.x; q < HID; q += blockDim.x) sb1[q] = b1[q];
    for (int q = threadIdx.x; q < 2 * HID; q += blockDim.x) sW2[q] = W2[q];
    __syncthreads();
    const float c0 = b2[0], c1 = b2[1];
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; p < n_pix; p += stride) {
        float xv[CIN];
        load_pixel<CIN>(x, p, xv);
        float o0 = c0, o1 = c1;
#pragma unroll 4
        for (int j = 0; j < HID; ++j) {
            const float4* w = reinterpret_cast<const float4*>(sW1 + j * CIN);
            float a = sb1[j];
#pragma unroll
            for (int q = 0; q < CIN / 4; ++q) {
                const float4 wq = w[q];
                a = fmaf(wq.x, xv[4 * q], a);
                a = fmaf(wq.y, xv[4 * q + 1], a);
                a = fmaf(wq.z, xv[4 * q + 2], a);
                a = fmaf(wq.w, xv[4 * q + 3], a);
            }
            const float h = fmaxf(a, 0.0f);
            o0 = fmaf(sW2[j], h, o0);
            o1 = fmaf(sW2[HID + j], h, o1);
        }
        const long long b = p / HW, s = p - b * HW;
        out[(2 * b) * HW + s] = o0;
        out[(2 * b + 1) * HW + s] = o1;
    }
}

// ---- backward -----------------------------------------------------------------------------------------------------
// partials[cta] = { gW1 [HID*CIN] | gb1 [HID] | gW2 [2*HID] | gb2 [2] }

template <int CIN, int HID>
__global__ void __launch_bounds__(kFhThreads) fieldhead_bwd_kernel(const float* __restrict__ x, const float* __restrict__ W1,
                                                                  const float* __restrict__ b1, const float* __restrict__ W2,
                                                                  const float* __restrict__ gOut, float* __restrict__ gx,
                                                                  float* __restrict__ partials, long long n_pix, int HW) {
    static_assert(HID == kFhThreads, "phase 2 maps one thread to one hidden unit");
    static_assert(kFhTile * CIN / 4 == kFhThreads, "one float4 of the input tile per thread");
    constexpr int kWarps = kFhThreads / 32, kPerWarp = HID / kWarps;
    __shared__ __align__(16) float sW1[HID * CIN];
    __shared__ float sb1[HID];
    __shared__ float sW2[2 * HID];
    __shared__ __align__(16) float sX[kFhTile * CIN];          // [pixel][channel]
    __shared__ float sG[2 * kFhTile];                          // [o][pixel]
    __shared__ float sH[HID * kFhPad];                         // [hidden][pixel], relu output (> 0 <=> the unit is active)
    __shared__ __align__(16) float sGX[kWarps * kFhTile * CIN];  // [warp][pixel][channel] partial d/dx
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int q = t; q < HID * CIN; q += kFhThreads) sW1[q] = W1[q];
    for (int q = t; q < HID; q += kFhThreads) sb1[q] = b1[q];
    for (int q = t; q < 2 * HID; q += kFhThreads) sW2[q] = W2[q];

    float aW1[CIN];
#pragma unroll
    for (int i = 0; i < CIN; ++i) aW1[i] = 0.0f;
    float ab1 = 0.0f, aW2a = 0.0f, aW2b = 0.0f, ab2 = 0.0f;

    const long long n_tiles = (n_pix + kFhTile - 1) / kFhTile;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long p0 = tile * kFhTile;
        const long long rem = n_pix - p0;
        const int np = rem < kFhTile ? static_cast<int>(rem) : kFhTile;
        __syncthreads();   // weights loaded (first tile) / previous tile fully consumed
        // ---- stage the tile: one float4 of x per thread, the two upstream gradients of the 32 pixels by threads 0..63
        {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t * 4 < np * CIN) v = ldg_stream(reinterpret_cast<const float4*>(x + p0 * CIN) + t);
            reinterpret_cast<float4*>(sX)[t] = v;
            if (t < 2 * kFhTile) {
                const int o = t / kFhTile, pp = t - o * kFhTile;
                float g = 0.0f;
                if (pp < np) {
                    const long long p = p0 + pp, b = p / HW, s = p - b * HW;
                    g = gOut[(2 * b + o) * HW + s];
                }
                sG[t] = g;
            }
        }
        __syncthreads();
        // ---- phase 1: lane = pixel, warp = hidden units [warp*kPerWarp, (warp+1)*kPerWarp)
        {
            float xv[CIN], gxv[CIN];
#pragma unroll
            for (int q = 0; q < CIN / 4; ++q) {
                const float4 v = reinterpret_cast<const float4*>(sX + lane * CIN)[q];
                xv[4 * q] = v.x; xv[4 * q + 1] = v.y; xv[4 * q + 2] = v.z; xv[4 * q + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < CIN; ++i) gxv[i] = 0.0f;
            const float g0 = sG[lane], g1 = sG[kFhTile + lane];
#pragma unroll 2
            for (int jj = 0; jj < kPerWarp; ++jj) {
                const int j = warp * kPerWarp + jj;
                const float4* w = reinterpret_cast<const float4*>(sW1 + j * CIN);
                float4 wq[CIN / 4];
                float a = sb1[j];
#pragma unroll
                for (int q = 0; q < CIN / 4; ++q) {
                    wq[q] = w[q];
                    a = fmaf(wq[q].x, xv[4 * q], a);
                    a = fmaf(wq[q].y, xv[4 * q + 1], a);
                    a = fmaf(wq[q].z, xv[4 * q + 2], a);
                    a = fmaf(wq[q].w, xv[4 * q + 3], a);
                }
                const float h = fmaxf(a, 0.0f);
                const float gh = a > 0.0f ? fmaf(sW2[j], g0, sW2[HID + j] * g1) : 0.0f;
                sH[j * kFhPad + lane] = h;
#pragma unroll
                for (int q = 0; q < CIN / 4; ++q) {
                    gxv[4 * q] = fmaf(wq[q].x, gh, gxv[4 * q]);
                    gxv[4 * q + 1] = fmaf(wq[q].y, gh, gxv[4 * q + 1]);
                    gxv[4 * q + 2] = fmaf(wq[q].z, gh, gxv[4 * q + 2]);
                    gxv[4 * q + 3] = fmaf(wq[q].w, gh, gxv[4 * q + 3]);
                }
            }
            float4* dst = reinterpret_cast<float4*>(sGX + (warp * kFhTile + lane) * CIN);
#pragma unroll
            for (int q = 0; q < CIN / 4; ++q) dst[q] = make_float4(gxv[4 * q], gxv[4 * q + 1], gxv[4 * q + 2], gxv[4 * q + 3]);
        }
        __syncthreads();
        // ---- d/dx of the tile: thread t owns float4 t of the [pixel][channel] tile, summed over the warps in order
        if (t * 4 < np * CIN) {
            float4 s = reinterpret_cast<const float4*>(sGX)[t];
#pragma unroll
            for (int w = 1; w < kWarps; ++w) {
                const float4 v = reinterpret_cast<const float4*>(sGX + w * kFhTile * CIN)[t];
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
            stg_stream(reinterpret_cast<float4*>(gx + p0 * CIN) + t, s);
        }
        // ---- phase 2: thread = hidden unit j; invalid pixels of a partial tile carry x = 0 and g = 0
        {
            const int j = t;
            const float w2a = sW2[j], w2b = sW2[HID + j];
            for (int p = 0; p < kFhTile; ++p) {
                const float h = sH[j * kFhPad + p], g0 = sG[p], g1 = sG[kFhTile + p];
                const float gh = h > 0.0f ? fmaf(w2a, g0, w2b * g1) : 0.0f;   // same expression as phase 1
                ab1 += gh;
                aW2a = fmaf(g0, h, aW2a);
                aW2b = fmaf(g1, h, aW2b);
#pragma unroll
                for (int q = 0; q < CIN / 4; ++q) {
                    const float4 v = reinterpret_cast<const float4*>(sX + p * CIN)[q];
                    aW1[4 * q] = fmaf(gh, v.x, aW1[4 * q]);
                    aW1[4 * q + 1] = fmaf(gh, v.y, aW1[4 * q + 1]);
                    aW1[4 * q + 2] = fmaf(gh, v.z, aW1[4 * q + 2]);
                    aW1[4 * q + 3] = fmaf(gh, v.w, aW1[4 * q + 3]);
                }
            }
            if (t < 2) {
                for (int p = 0; p < kFhTile; ++p) ab2 += sG[t * kFhTile + p];
            }
        }
    }
    constexpr int kPartial = HID * CIN + HID + 2 * HID + 2;
    float* dst = partials + static_cast<long long>(blockIdx.x) * kPartial;
#pragma unroll
    for (int i = 0; i < CIN; ++i) dst[t * CIN + i] = aW1[i];
    dst[HID * CIN + t] = ab1;
    dst[HID * CIN + HID + t] = aW2a;
    dst[HID * CIN + 2 * HID + t] = aW2b;
    if (t < 2) dst[HID * CIN + 3 * HID + t] = ab2;
}

// ---- adjoint of the moments: gx[p] (+)= a + M x[p] ------------------------------------------------------------------
template <int CIN>
__global__ void __launch_bounds__(256) affine_acc_kernel(const float* __restrict__ x, const float* __restrict__ a,
                                                         const float* __restrict__ M, float* __restrict__ gx,
                                                         long long n_pix, int accumulate) {
    __shared__ __align__(16) float sM[CIN * CIN];
    __shared__ float sa[CIN];
    for (int q = threadIdx.x; q < CIN * CIN; q += blockDim.x) sM[q] = M[q];
    for (int q = threadIdx.x; q < CIN; q += blockDim.x) sa[q] = a[q];
    __syncthreads();
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long p = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; p < n_pix; p += stride) {
        float xv[CIN];
        load_pixel<CIN>(x, p, xv);
        float4* dst = reinterpret_cast<float4*>(gx + p * CIN);
#pragma unroll
        for (int q = 0; q < CIN / 4; ++q) {
            float r[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = 4 * q + e;
                float s = sa[i];
#pragma unroll
                for (int k4 = 0; k4 < CIN / 4; ++k4) {
                    const float4 m = reinterpret_cast<const float4*>(sM + i * CIN)[k4];
                    s = fmaf(m.x, xv[4 * k4], s);
                    s = fmaf(m.y, xv[4 * k4 + 1], s);
                    s = fmaf(m.z, xv[4 * k4 + 2], s);
                    s = fmaf(m.w, xv[4 * k4 + 3], s);
                }
                r[e] = s;
            }
            float4 o = make_float4(r[0], r[1], r[2], r[3]);
            if (accumulate) {
                const float4 old = dst[q];
                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
            }
            dst[q] = o;
        }
    }
}

#ifndef BH_HOST_EMULATION
// the tensor-core kernels of the same entry points (fieldhead_mma.cu)
bool fieldhead_mma_ok(int HW, int cin, int hid);
int launch_fieldhead_fwd_mma(const float* x, const float* W1, const float* b1, const float* W2, const float* b2, float* out, long long n_pix,
                             int HW, int tf32, cudaStream_t stream);
int launch_fieldhead_bwd_mma(const float* x, const float* W1, const float* b1, const float* W2, const float* gOut, float* gx, float* partials,
                             long long n_pix, int HW, int tf32, cudaStream_t stream);
int launch_fieldhead_moments_mma(const float* x, double* partials, long long n_pix, int grid, cudaStream_t stream);
inline bool use_mma(int HW, int cin, int hid) { return g_tune[kTuneFieldheadVariant] != 1 && fieldhead_mma_ok(HW, cin, hid); }

inline int fh_grid(long long n_items, int per_sm) {
    const long long cap = static_cast<long long>(kNumSMs) * per_sm;
    return static_cast<int>(n_items < 1 ? 1 : (n_items < cap ? n_items : cap));
}

#endif
}  // namespace bh

#ifndef BH_HOST_EMULATION
// Only the shipped geometry is compiled: 16 input channels, 128 hidden units, 2 outputs (Rethinking.py:145-147, ResNet34
// blocks).  Other widths (the ResNet50 variant: 64 -> 512 -> 2) get BH_E_UNSUPPORTED and stay on the ATen modules.
extern "C" int bh_fieldhead_supported(int cin, int hid) { return cin == 16 && hid == 128; }

extern "C" int bh_fieldhead_grid(int what, long long n_pix) {
    using namespace bh;
    if (n_pix <= 0) return 0;
    if (what == 0) return fh_grid((n_pix + kMomTile - 1) / kMomTile, 7);   // moments: 7 x 288 threads per SM
    return fh_grid((n_pix + kFhTile - 1) / kFhTile, 5);                    // backward (either kernel): 32-pixel tiles, 5 CTAs per SM
}

extern "C" int bh_fieldhead_moments(const float* x, double* partials, long long n_pix, int cin, bh_stream_t stream) {
    using namespace bh;
    if (!x || !partials) return BH_E_NULL;
    if (n_pix <= 0) return BH_E_SHAPE;
    if (cin != 16) return BH_E_UNSUPPORTED;
    if (!aligned16(x)) return BH_E_ALIGN;
    if (g_tune[kTuneFieldheadVariant] != 1 && (n_pix % 8) == 0)
        return launch_fieldhead_moments_mma(x, partials, n_pix, bh_fieldhead_grid(0, n_pix), reinterpret_cast<cudaStream_t>(stream));
    moments_kernel<16><<<bh_fieldhead_grid(0, n_pix), kMomThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, partials, n_pix);
    return launch_status();
}

extern "C" int bh_fieldhead_fwd(const float* x, const float* W1, const float* b1, const float* W2, const float* b2, float* out,
                                int B, int HW, int cin, int hid, int tf32, bh_stream_t stream) {
    using namespace bh;
    if (!x || !W1 || !b1 || !W2 || !b2 || !out) return BH_E_NULL;
    if (B <= 0 || HW <= 0) return BH_E_SHAPE;
    if (!bh_fieldhead_supported(cin, hid)) return BH_E_UNSUPPORTED;
    if (!aligned16(x)) return BH_E_ALIGN;
    const long long n_pix = static_cast<long long>(B) * HW;
    if (use_mma(HW, cin, hid)) return launch_fieldhead_fwd_mma(x, W1, b1, W2, b2, out, n_pix, HW, tf32, reinterpret_cast<cudaStream_t>(stream));
    fieldhead_fwd_kernel<16, 128><<<fh_grid((n_pix + 255) / 256, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        x, W1, b1, W2, b2, out, n_pix, HW);
    return launch_status();
}

extern "C" int bh_fieldhead_bwd(const float* x, const float* W1, const float* b1, const float* W2, const float* gOut, float* gx,
                                float* partials, int B, int HW, int cin, int hid, int tf32, bh_stream_t stream) {
    using namespace bh;
    if (!x || !W1 || !b1 || !W2 || !gOut || !gx || !partials) return BH_E_NULL;
    if (B <= 0 || HW <= 0) return BH_E_SHAPE;
    if (!bh_fieldhead_supported(cin, hid)) return BH_E_UNSUPPORTED;
    if (!aligned16(x) || !aligned16(gx)) return BH_E_ALIGN;
    const long long n_pix = static_cast<long long>(B) * HW;
    if (use_mma(HW, cin, hid))
        return launch_fieldhead_bwd_mma(x, W1, b1, W2, gOut, gx, partials, n_pix, HW, tf32, reinterpret_cast<cudaStream_t>(stream));
    fieldhead_bwd_kernel<16, 128><<<bh_fieldhead_grid(1, n_pix), kFhThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        x, W1, b1, W2, gOut, gx, partials, n_pix, HW);
    return launch_status();
}

extern "C" int bh_fieldhead_affine(const float* x, const float* a, const float* M, float* gx, long long n_pix, int cin,
                                   int accumulate, bh_stream_t stream) {
    using namespace bh;
    if (!x || !a || !M || !gx) return BH_E_NULL;
    if (n_pix <= 0) return BH_E_SHAPE;
    if (cin != 16) return BH_E_UNSUPPORTED;
    if (!aligned16(x) || !aligned16(gx)) return BH_E_ALIGN;
    affine_acc_kernel<16><<<fh_grid((n_pix + 255) / 256, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, a, M, gx, n_pix,
                                                                                                               accumulate);
    return launch_status();
}
#endif  // BH_HOST_EMULATION
