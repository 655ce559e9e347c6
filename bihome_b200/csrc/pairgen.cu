// K5: synthetic (PD)S-COCO training-pair generation on the GPU.
//
// Reference semantics: HomographyNetPrep.__call__ (src/data/transforms.py:456-576,724-725) followed by
// DictToGrayscale (:344-354), DictStandardize (:369-378), DictToTensor (:728-743) and the .float() of
// train.py:308-309.  The CPU reference renders two full 240x320x3 float images per sample and crops; this
// kernel renders only the 2 x P x P patch pixels.  One CTA per 32x32 patch tile: the source pixels the tile's warped
// taps can touch (bounding box of its four projected corners) go through the photometric chain ONCE, into a float RGB
// window in shared memory, and the four bilinear taps of every patch-2 pixel read that window -- 1 + (box / tile) chain
// evaluations per pixel pair (about 2.3 at rho = 32) instead of 5.  The chain:
//   uint8 RGB -> float -> [+brightness] -> [*contrast] -> RGB2HSV -> [S *= a] -> [H += d, wrap] -> HSV2RGB
//             -> [*contrast] -> [channel permutation]                          (PhotometricDistortSimple :296-330)
//   patch_2 additionally goes through cv2.warpPerspective(image_2, inv(H)) (src/data/utils.py:61-64):
//   double-precision coordinates, rounded to 1/32 px, bilinear weights from OpenCV's 32x32 float table,
//   BORDER_CONSTANT 0 -- reproduced exactly (oracle check in tests/test_oracle_golden.py).
// The float32 colour math follows OpenCV 4.x cvtColor (RGB2HSV_f / HSV2RGB_f) including where its build fuses
// multiply-adds; this file is compiled with -fmad=false so that only the explicit fmaf below are fused.
#include <limits.h>

#include "bh_common.cuh"

namespace bh {

constexpr int kNP = BH_PAIR_NPARAM;

struct Photo {
    bool b_on, contrast_first, c_on, s_on, h_on, l_on;
    float b_delta, c_alpha, s_alpha, h_delta;
    int perm;
};

__device__ __forceinline__ Photo load_photo(const double* p) {
    Photo q;
    q.b_on = p[0] != 0.0; q.b_delta = static_cast<float>(p[1]);
    q.contrast_first = p[2] != 0.0;
    q.c_on = p[3] != 0.0; q.c_alpha = static_cast<float>(p[4]);
    q.s_on = p[5] != 0.0; q.s_alpha = static_cast<float>(p[6]);
    q.h_on = p[7] != 0.0; q.h_delta = static_cast<float>(p[8]);
    q.l_on = p[9] != 0.0; q.perm = static_cast<int>(p[10]);
    return q;
}

// one pixel through the photometric chain (float RGB out, never clipped -- as the reference)
__device__ __forceinline__ void photometric(const Photo& q, float& r, float& g, float& b) {
    if (q.b_on) { r += q.b_delta; g += q.b_delta; b += q.b_delta; }
    if (q.contrast_first && q.c_on) { r *= q.c_alpha; g *= q.c_alpha; b *= q.c_alpha; }
    // cv2.COLOR_RGB2HSV, float32: V = max, S = (V-min)/(|V|+eps), H = 60*(..)/(V-min+eps) (+120/+240), H<0 -> +360
    const float v = fmaxf(fmaxf(r, g), b), vmin = fminf(fminf(r, g), b);
    const float diff = v - vmin;
    float s = diff / (fabsf(v) + 1.1920929e-07f);
    const float d = 60.0f / (diff + 1.1920929e-07f);
    float h;
    if (v == r) {
        h = (g - b) * d;
        if (h < 0.0f) h = fmaf(g - b, d, 360.0f);
    } else if (v == g) {
        h = fmaf(b - r, d, 120.0f);
        if (h < 0.0f) h += 360.0f;
    } else {
        h = fmaf(r - g, d, 240.0f);
        if (h < 0.0f) h += 360.0f;
    }
    if (q.s_on) s *= q.s_alpha;
    if (q.h_on) {
        h += q.h_delta;
        if (h > 360.0f) h -= 360.0f;
        if (h < 0.0f) h += 360.0f;
    }
    // cv2.COLOR_HSV2RGB, float32
    if (s == 0.0f) {
        r = g = b = v;
    } else {
        float hh = h * 0.016666668f;  // 6/360 as float
        if (hh >= 6.0f) hh -= 6.0f;  // == fmodf(hh, 6) on [0, 12): h is in [0, 360] here
        int sector = static_cast<int>(floorf(hh));
        hh -= static_cast<float>(sector);
        if (static_cast<unsigned>(sector) >= 6u) { sector = 0; hh = 0.0f; }
        const float t0 = v, t1 = v * (1.0f - s), t2 = v * fmaf(-s, hh, 1.0f), t3 = v * fmaf(-s, 1.0f - hh, 1.0f);
        switch (sector) {  // (b, g, r) = tab[{1,3,0},{1,0,2},{3,0,1},{0,2,1},{0,1,3},{2,1,0}]
            case 0: b = t1; g = t3; r = t0; break;
            case 1: b = t1; g = t0; r = t2; break;
            case 2: b = t3; g = t0; r = t1; break;
            case 3: b = t0; g = t2; r = t1; break;
            case 4: b = t0; g = t1; r = t3; break;
            default: b = t2; g = t1; r = t0; break;
        }
    }
    if (!q.contrast_first && q.c_on) { r *= q.c_alpha; g *= q.c_alpha; b *= q.c_alpha; }
    if (q.l_on) {  // image[:, :, perm]: new channel k = old channel perm[k]
        const float c[3] = {r, g, b};
        const int P6[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
        const int k = q.perm < 0 ? 0 : (q.perm > 5 ? 5 : q.perm);
        r = c[P6[k][0]]; g = c[P6[k][1]]; b = c[P6[k][2]];
    }
}

__device__ __forceinline__ void fetch_rgb(const uint8_t* img, int Wi, int x, int y, float& r, float& g, float& b) {
    const uint8_t* p = img + (static_cast<size_t>(y) * Wi + x) * 3;
    r = static_cast<float>(p[0]); g = static_cast<float>(p[1]); b = static_cast<float>(p[2]);
}

// DictToGrayscale (float32, left to right) -> DictStandardize (float64) -> .float()
__device__ __forceinline__ float to_input(float r, float g, float b, double mean, double stdv) {
    const float gray = (r * 0.299f + g * 0.587f) + b * 0.114f;
    const float scaled = gray / 255.0f;
    return static_cast<float>((static_cast<double>(scaled) - mean) / stdv);
}

// cv2.getPerspectiveTransform twin, one warp: the 8x8 system in float64 with partial pivoting, rows spread over 8 lanes
__device__ __forceinline__ void pairgen_solve(const double* prm, int P, double* M) {
    const int x0 = static_cast<int>(prm[22]) - P / 2, y0 = static_cast<int>(prm[23]) - P / 2;
    const int sub = threadIdx.x & 7, i = sub >> 1;
    const double cx = (i == 1 || i == 2) ? double(x0 + P) : double(x0), cy = (i >= 2) ? double(y0 + P) : double(y0);
    const double ux = cx + prm[24 + 2 * i], uy = cy + prm[25 + 2 * i];
    double a[8], rhs, sol[8];
    if ((sub & 1) == 0) { a[0] = cx; a[1] = cy; a[2] = 1; a[3] = 0; a[4] = 0; a[5] = 0; a[6] = -cx * ux; a[7] = -cy * ux; rhs = ux; }
    else { a[0] = 0; a[1] = 0; a[2] = 0; a[3] = cx; a[4] = cy; a[5] = 1; a[6] = -cx * uy; a[7] = -cy * uy; rhs = uy; }
    solve8(a, rhs, sub, sol);
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) M[k] = sol[k];
        M[8] = 1.0;
    }
}
// cv2.warpPerspective's fixed-point source coordinate of image pixel (px, py): X, Y in 1/32 px
__device__ __forceinline__ void cv2_coords(const double* M, int px, int py, int& X, int& Y) {
    double W = M[6] * px + M[7] * py + M[8];
    W = (W != 0.0) ? 32.0 / W : 0.0;
    const double fX = fmax(-2147483648.0, fmin(2147483647.0, (M[0] * px + M[1] * py + M[2]) * W));
    const double fY = fmax(-2147483648.0, fmin(2147483647.0, (M[3] * px + M[4] * py + M[5]) * W));
    X = __double2int_rn(fX);
    Y = __double2int_rn(fY);
}

constexpr int kPgTile = 32;          // patch pixels per tile side
constexpr int kPgThreads = 128;      // lane = column, warp w = rows 8w .. 8w + 7 of the tile
constexpr int kPgBox = 52;           // largest staged source box (pixels per side): scale 1.5 (rho = 32 at P = 128) + slack
constexpr int kPgSmem = 3 * kPgBox * kPgBox * 4;

__global__ void __launch_bounds__(kPgThreads, 6)
    pairgen_apply_kernel(const uint8_t* __restrict__ images, const int32_t* __restrict__ index, const double* __restrict__ params,
                         float* __restrict__ patch1, float* __restrict__ patch2, float* __restrict__ delta, int n_img, int Hi,
                         int Wi, int P, int tiles_x, double mean, double stdv) {
    extern __shared__ __align__(16) float win[];     // [3][bh][bw] distorted RGB of the staged box
    __shared__ double M[9];
    __shared__ double prm[kNP];
    __shared__ int box[5];                           // bx0, by0, bw, bh, staged?
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < kNP) prm[threadIdx.x] = params[static_cast<size_t>(b) * kNP + threadIdx.x];
    __syncthreads();
    const int x0 = static_cast<int>(prm[22]) - P / 2, y0 = static_cast<int>(prm[23]) - P / 2;
    if (blockIdx.x == 0 && threadIdx.x < 8) delta[b * 8 + threadIdx.x] = static_cast<float>(prm[24 + threadIdx.x]);
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int tx_lo = tx * kPgTile, ty_lo = ty * kPgTile;
    const int tx_hi = min(P, tx_lo + kPgTile), ty_hi = min(P, ty_lo + kPgTile);
    int im = index[b];
    im = im < 0 ? 0 : (im >= n_img ? n_img - 1 : im);
    const uint8_t* img = images + static_cast<size_t>(im) * Hi * Wi * 3;
    const double kmean = mean, kstd = stdv;  // float64, as numpy promotes the YAML's list mean/std
    const int x = tx_lo + lane, px = x0 + x;
    const bool xin = x < tx_hi;
    if (warp == 0) {
        // warp 0: the homography, then the source box of the tile (a projective map sends the tile to a convex
        // quadrilateral when the denominator keeps its sign on the four corners; one pixel of slack covers the 1/32-px
        // rounding).  Meanwhile the other warps render their rows of patch 1, which needs neither.
        pairgen_solve(prm, P, M);
        __syncwarp();
        const int k = lane & 3;
        const int cx = x0 + ((k & 1) ? tx_hi - 1 : tx_lo), cy = y0 + ((k & 2) ? ty_hi - 1 : ty_lo);
        const double w = M[6] * cx + M[7] * cy + M[8];
        int X, Y;
        cv2_coords(M, cx, cy, X, Y);
        int xmin = X >> 5, xmax = X >> 5, ymin = Y >> 5, ymax = Y >> 5;
        bool pos = w > 0.0, neg = w < 0.0;
#pragma unroll
        for (int o = 1; o < 4; o <<= 1) {
            xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o)); xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
            ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o)); ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
            pos = __shfl_xor_sync(0xffffffffu, pos ? 1 : 0, o) != 0 && pos;
            neg = __shfl_xor_sync(0xffffffffu, neg ? 1 : 0, o) != 0 && neg;
        }
        if (lane == 0) {
            const int bx0 = xmin - 1, by0 = ymin - 1;
            const long long bw = static_cast<long long>(xmax) + 3 - bx0, bh = static_cast<long long>(ymax) + 3 - by0;
            const bool ok = (pos || neg) && bw <= kPgBox && bh <= kPgBox && abs(xmin) < (1 << 24) && abs(ymin) < (1 << 24);
            box[0] = bx0; box[1] = by0; box[2] = ok ? static_cast<int>(bw) : 0; box[3] = ok ? static_cast<int>(bh) : 0; box[4] = ok ? 1 : 0;
        }
    }
    // patch 1: plain crop of the distorted image (chain of image 1), rows 8 warp .. 8 warp + 7 of the tile
    const Photo q1 = load_photo(prm);
#pragma unroll 2
    for (int j = 0; j < 8; ++j) {
        const int y = ty_lo + 8 * warp + j, py = y0 + y;
        if (xin && y < ty_hi) {
            float r = 0.0f, g = 0.0f, bl = 0.0f;
            if (px >= 0 && px < Wi && py >= 0 && py < Hi) {
                fetch_rgb(img, Wi, px, py, r, g, bl);
                photometric(q1, r, g, bl);
            }
            patch1[(static_cast<size_t>(b) * P + y) * P + x] = to_input(r, g, bl, kmean, kstd);
        }
    }
    __syncthreads();   // M and the box are published
    const Photo q2 = load_photo(prm + 11);
    const int bx0 = box[0], by0 = box[1], bw = box[2], bh = box[3];
    const bool staged = box[4] != 0;
    const int plane = bw * bh;
    // stage: every source pixel of the box through image 2's chain, once (zeros outside the image: BORDER_CONSTANT)
    for (int r = warp; r < bh; r += kPgThreads / 32) {
        const int sy = by0 + r;
        for (int c = lane; c < bw; c += 32) {
            const int sx = bx0 + c;
            float tr = 0.0f, tg = 0.0f, tb = 0.0f;
            if (sx >= 0 && sx < Wi && sy >= 0 && sy < Hi) {
                fetch_rgb(img, Wi, sx, sy, tr, tg, tb);
                photometric(q2, tr, tg, tb);
            }
            const int i = r * bw + c;
            win[i] = tr; win[plane + i] = tg; win[2 * plane + i] = tb;
        }
    }
    __syncthreads();
    // patch 2: cv2.warpPerspective semantics at image pixel (px, py)
#pragma unroll 2
    for (int j = 0; j < 8; ++j) {
        const int y = ty_lo + 8 * warp + j, py = y0 + y;
        if (!(xin && y < ty_hi)) continue;
        int X, Y;
        cv2_coords(M, px, py, X, Y);
        const int sx = X >> 5, sy = Y >> 5;
        const float ax = static_cast<float>(X & 31) * 0.03125f, ay = static_cast<float>(Y & 31) * 0.03125f;
        const float ww[4] = {(1.0f - ay) * (1.0f - ax), (1.0f - ay) * ax, ay * (1.0f - ax), ay * ax};
        float acc[3] = {0.0f, 0.0f, 0.0f};
        if (!(sx >= Wi || sx + 1 < 0 || sy >= Hi || sy + 1 < 0)) {
            const int o = (sy - by0) * bw + (sx - bx0);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                float tr = 0.0f, tg = 0.0f, tb = 0.0f;
                if (staged) {
                    const int oo = o + (t >> 1) * bw + (t & 1);
                    tr = win[oo]; tg = win[plane + oo]; tb = win[2 * plane + oo];
                } else {
                    const int qx = sx + (t & 1), qy = sy + (t >> 1);
                    if (qx >= 0 && qx < Wi && qy >= 0 && qy < Hi) {
                        fetch_rgb(img, Wi, qx, qy, tr, tg, tb);
                        photometric(q2, tr, tg, tb);
                    }
                }
                if (t == 0) { acc[0] = tr * ww[0]; acc[1] = tg * ww[0]; acc[2] = tb * ww[0]; }
                else { acc[0] = acc[0] + tr * ww[t]; acc[1] = acc[1] + tg * ww[t]; acc[2] = acc[2] + tb * ww[t]; }
            }
        }
        patch2[(static_cast<size_t>(b) * P + y) * P + x] = to_input(acc[0], acc[1], acc[2], kmean, kstd);
    }
}

// the whole first image of the pair: photometric chain of image 1 -> grayscale -> standardise, [B,1,Hi,Wi]
// (HomographyNetPrep's 'image_1' after DictToGrayscale / DictStandardize / DictToTensor: the PhotometricHead input)
__global__ void __launch_bounds__(256)
    pairgen_image_kernel(const uint8_t* __restrict__ images, const int32_t* __restrict__ index, const double* __restrict__ params,
                         float* __restrict__ image1, int n_img, int Hi, int Wi, double mean, double stdv) {
    const int b = blockIdx.y;
    int im = index[b];
    im = im < 0 ? 0 : (im >= n_img ? n_img - 1 : im);
    const uint8_t* img = images + static_cast<size_t>(im) * Hi * Wi * 3;
    const Photo q1 = load_photo(params + static_cast<size_t>(b) * kNP);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Hi * Wi; i += gridDim.x * blockDim.x) {
        float r = static_cast<float>(img[3 * i]), g = static_cast<float>(img[3 * i + 1]), bl = static_cast<float>(img[3 * i + 2]);
        photometric(q1, r, g, bl);
        image1[static_cast<size_t>(b) * Hi * Wi + i] = to_input(r, g, bl, mean, stdv);
    }
}

// ---- counter-based draws: Philox4x32-10 keyed by seed, counter = (sample, step, draw) ------------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
    const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c[0], p1 = static_cast<uint64_t>(0xCD9E8D57u) * c[2];
    const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c[1] ^ k[0], n2 = static_cast<uint32_t>(p0 >> 32) ^ c[3] ^ k[1];
    c[1] = static_cast<uint32_t>(p1); c[3] = static_cast<uint32_t>(p0); c[0] = n0; c[2] = n2;
    k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
}
struct Draws {
    uint32_t key[2], ctr_hi[2], block, buf[4];
    int have;
    __device__ Draws(uint64_t seed, uint64_t step, uint32_t sample) : block(0), have(0) {
        key[0] = static_cast<uint32_t>(seed); key[1] = static_cast<uint32_t>(seed >> 32);
        ctr_hi[0] = static_cast<uint32_t>(step); ctr_hi[1] = static_cast<uint32_t>(step >> 32) ^ (sample * 0x9E3779B1u);
        sample_ = sample;
    }
    uint32_t sample_;
    __device__ uint32_t next() {
        if (have == 0) {
            uint32_t c[4] = {block++, sample_, ctr_hi[0], ctr_hi[1]};
            uint32_t k[2] = {key[0], key[1]};
#pragma unroll
            for (int r = 0; r < 10; ++r) philox_round(c, k);
            buf[0] = c[0]; buf[1] = c[1]; buf[2] = c[2]; buf[3] = c[3];
            have = 4;
        }
        return buf[--have];
    }
    __device__ double uniform() {  // [0, 1) with 53 random bits
        const uint64_t hi = next(), lo = next();
        return static_cast<double>(((hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
    }
    __device__ bool coin() { return (next() & 1u) != 0u; }
    __device__ int randint(int lo, int hi) {  // [lo, hi)
        const int n = hi - lo;
        if (n <= 1) return lo;
        int v = static_cast<int>(uniform() * n);
        return lo + (v >= n ? n - 1 : v);
    }
};

__device__ void draw_photo(Draws& d, double* p, double max_delta) {
    const double lo = 1.0 - max_delta / 32.0 * 0.5, hi = 1.0 + max_delta / 32.0 * 0.5;
    const bool b_on = d.coin();
    p[0] = b_on; p[1] = b_on ? (2.0 * d.uniform() - 1.0) * max_delta : 0.0;
    p[2] = d.coin();
    const bool c_on = d.coin();
    p[3] = c_on; p[4] = c_on ? lo + (hi - lo) * d.uniform() : 1.0;
    const bool s_on = d.coin();
    p[5] = s_on; p[6] = s_on ? lo + (hi - lo) * d.uniform() : 1.0;
    const bool h_on = d.coin();
    p[7] = h_on; p[8] = h_on ? (2.0 * d.uniform() - 1.0) * max_delta * 0.5 : 0.0;
    const bool l_on = max_delta > 0.0 ? d.coin() : false;
    p[9] = l_on; p[10] = l_on ? d.randint(0, 6) : 0;
}

__global__ void pairgen_draw_kernel(double* __restrict__ params, int32_t* __restrict__ index, int B, int n_img, int Hi, int Wi,
                                    int rho, int P, float max_delta, uint64_t seed, uint64_t step) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    Draws d(seed, step, static_cast<uint32_t>(b));
    double* p = params + static_cast<size_t>(b) * kNP;
    index[b] = d.randint(0, n_img);
    draw_photo(d, p, max_delta);
    draw_photo(d, p + 11, max_delta);
    if (P != Wi) {
        p[22] = d.randint(rho + P / 2, Wi - rho - P / 2 + 1);
        p[23] = d.randint(rho + P / 2, Hi - rho - P / 2 + 1);
    } else {
        p[22] = Wi / 2; p[23] = Hi / 2;
    }
    for (int i = 0; i < 8; ++i) p[24 + i] = d.randint(-rho, rho);
}

}  // namespace bh

extern "C" int bh_pairgen_draw(double* params, int32_t* index, int B, int n_img, int Hi, int Wi, int rho, int P,
                               float max_delta, uint64_t seed, uint64_t step, bh_stream_t stream) {
    if (!params || !index) return BH_E_NULL;
    if (B <= 0 || n_img <= 0 || Hi <= 0 || Wi <= 0 || P <= 0 || rho < 0 || max_delta < 0.0f) return BH_E_SHAPE;
    if (P != Wi && (Wi - rho - P / 2 + 1 <= rho + P / 2 || Hi - rho - P / 2 + 1 <= rho + P / 2)) return BH_E_SHAPE;
    bh::pairgen_draw_kernel<<<(B + 127) / 128, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(params, index, B, n_img, Hi, Wi,
                                                                                               rho, P, max_delta, seed, step);
    return bh::launch_status();
}

extern "C" int bh_pairgen_apply(const uint8_t* images, const int32_t* index, const double* params, float* patch1,
                                float* patch2, float* delta, int B, int n_img, int Hi, int Wi, int P, double mean, double std,
                                bh_stream_t stream) {
    if (!images || !index || !params || !patch1 || !patch2 || !delta) return BH_E_NULL;
    if (B <= 0 || n_img <= 0 || Hi <= 0 || Wi <= 0 || P <= 0 || std == 0.0) return BH_E_SHAPE;
    const int tiles_x = (P + bh::kPgTile - 1) / bh::kPgTile;
    if (static_cast<long long>(tiles_x) * tiles_x > 65535ll * 32 || B > 65535) return BH_E_SHAPE;
    dim3 grid(tiles_x * tiles_x, B);
    bh::pairgen_apply_kernel<<<grid, bh::kPgThreads, bh::kPgSmem, reinterpret_cast<cudaStream_t>(stream)>>>(
        images, index, params, patch1, patch2, delta, n_img, Hi, Wi, P, tiles_x, mean, std);
    return bh::launch_status();
}

extern "C" int bh_pairgen_image(const uint8_t* images, const int32_t* index, const double* params, float* image1, int B,
                                int n_img, int Hi, int Wi, double mean, double std, bh_stream_t stream) {
    if (!images || !index || !params || !image1) return BH_E_NULL;
    if (B <= 0 || B > 65535 || n_img <= 0 || Hi <= 0 || Wi <= 0 || std == 0.0) return BH_E_SHAPE;
    int gx = (Hi * Wi + 256 * 4 - 1) / (256 * 4);
    dim3 grid(gx, B);
    bh::pairgen_image_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(images, index, params, image1, n_img, Hi, Wi, mean, std);
    return bh::launch_status();
}
