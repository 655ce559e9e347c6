// K5: synthetic (PD)S-COCO training-pair generation on the GPU.
//
// Reference semantics: HomographyNetPrep.__call__ (src/data/transforms.py:456-576,724-725) followed by
// DictToGrayscale (:344-354), DictStandardize (:369-378), DictToTensor (:728-743) and the .float() of
// train.py:308-309.  The CPU reference renders two full 240x320x3 float images per sample and crops; this
// kernel renders only the 2 x P x P patch pixels, evaluating the photometric chain per tap:
//   uint8 RGB -> float -> [+brightness] -> [*contrast] -> RGB2HSV -> [S *= a] -> [H += d, wrap] -> HSV2RGB
//             -> [*contrast] -> [channel permutation]                          (PhotometricDistortSimple :296-330)
//   patch_2 additionally goes through cv2.warpPerspective(image_2, inv(H)) (src/data/utils.py:61-64):
//   double-precision coordinates, rounded to 1/32 px, bilinear weights from OpenCV's 32x32 float table,
//   BORDER_CONSTANT 0 -- reproduced exactly (oracle check in tests/test_oracle_golden.py).
// The float32 colour math follows OpenCV 4.x cvtColor (RGB2HSV_f / HSV2RGB_f) including where its build fuses
// multiply-adds; this file is compiled with -fmad=false so that only the explicit fmaf below are fused.
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

__global__ void __launch_bounds__(256)
    pairgen_apply_kernel(const uint8_t* __restrict__ images, const int32_t* __restrict__ index, const double* __restrict__ params,
                         float* __restrict__ patch1, float* __restrict__ patch2, float* __restrict__ delta, int n_img, int Hi,
                         int Wi, int P, double mean, double stdv) {
    __shared__ double M[9];
    __shared__ double prm[kNP];
    const int b = blockIdx.y;
    if (threadIdx.x < kNP) prm[threadIdx.x] = params[static_cast<size_t>(b) * kNP + threadIdx.x];
    __syncthreads();
    const int pos_x = static_cast<int>(prm[22]), pos_y = static_cast<int>(prm[23]);
    const int x0 = pos_x - P / 2, y0 = pos_y - P / 2;
    if (threadIdx.x < 32) {
        // cv2.getPerspectiveTransform twin: 8x8 system in float64, partial pivoting, rows spread over 8 lanes
        const int sub = threadIdx.x & 7, i = sub >> 1;
        const double cx = (i == 1 || i == 2) ? double(x0 + P) : double(x0), cy = (i >= 2) ? double(y0 + P) : double(y0);
        const double ux = cx + prm[24 + 2 * i], uy = cy + prm[25 + 2 * i];
        double a[8], rhs, sol[8];
        if ((sub & 1) == 0) { a[0] = cx; a[1] = cy; a[2] = 1; a[3] = 0; a[4] = 0; a[5] = 0; a[6] = -cx * ux; a[7] = -cy * ux; rhs = ux; }
        else { a[0] = 0; a[1] = 0; a[2] = 0; a[3] = cx; a[4] = cy; a[5] = 1; a[6] = -cx * uy; a[7] = -cy * uy; rhs = uy; }
        solve8(a, rhs, sub, sol);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) M[k] = sol[k];
            M[8] = 1.0;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < 8) delta[b * 8 + threadIdx.x] = static_cast<float>(prm[24 + threadIdx.x]);
    __syncthreads();
    int im = index[b];
    im = im < 0 ? 0 : (im >= n_img ? n_img - 1 : im);
    const uint8_t* img = images + static_cast<size_t>(im) * Hi * Wi * 3;
    const Photo q1 = load_photo(prm), q2 = load_photo(prm + 11);
    const double kmean = mean, kstd = stdv;  // float64, as numpy promotes the YAML's list mean/std
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P * P; i += gridDim.x * blockDim.x) {
        const int y = i / P, x = i - y * P;
        const int px = x0 + x, py = y0 + y;
        float r, g, bl;
        // patch 1: plain crop of the distorted image
        if (px >= 0 && px < Wi && py >= 0 && py < Hi) {
            fetch_rgb(img, Wi, px, py, r, g, bl);
            photometric(q1, r, g, bl);
        } else {
            r = g = bl = 0.0f;
        }
        patch1[(static_cast<size_t>(b) * P + y) * P + x] = to_input(r, g, bl, kmean, kstd);
        // patch 2: cv2.warpPerspective semantics at image pixel (px, py)
        double W = M[6] * px + M[7] * py + M[8];
        W = (W != 0.0) ? 32.0 / W : 0.0;
        const double fX = fmax(-2147483648.0, fmin(2147483647.0, (M[0] * px + M[1] * py + M[2]) * W));
        const double fY = fmax(-2147483648.0, fmin(2147483647.0, (M[3] * px + M[4] * py + M[5]) * W));
        const int X = __double2int_rn(fX), Y = __double2int_rn(fY);
        const int sx = X >> 5, sy = Y >> 5;
        const float ax = static_cast<float>(X & 31) * 0.03125f, ay = static_cast<float>(Y & 31) * 0.03125f;
        const float w00 = (1.0f - ay) * (1.0f - ax), w01 = (1.0f - ay) * ax, w10 = ay * (1.0f - ax), w11 = ay * ax;
        float acc[3] = {0.0f, 0.0f, 0.0f};
        if (!(sx >= Wi || sx + 1 < 0 || sy >= Hi || sy + 1 < 0)) {
            const float ww[4] = {w00, w01, w10, w11};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int tx = sx + (t & 1), ty = sy + (t >> 1);
                float tr = 0.0f, tg = 0.0f, tb = 0.0f;
                if (tx >= 0 && tx < Wi && ty >= 0 && ty < Hi) {
                    fetch_rgb(img, Wi, tx, ty, tr, tg, tb);
                    photometric(q2, tr, tg, tb);
                }
                if (t == 0) { acc[0] = tr * ww[0]; acc[1] = tg * ww[0]; acc[2] = tb * ww[0]; }
                else { acc[0] = acc[0] + tr * ww[t]; acc[1] = acc[1] + tg * ww[t]; acc[2] = acc[2] + tb * ww[t]; }
            }
        }
        patch2[(static_cast<size_t>(b) * P + y) * P + x] = to_input(acc[0], acc[1], acc[2], kmean, kstd);
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
    int gx = (P * P + 256 * 8 - 1) / (256 * 8);  // ~8 pixels per thread: the per-block homography solve is amortised
    if (gx > 64) gx = 64;
    dim3 grid(gx, B);
    bh::pairgen_apply_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(images, index, params, patch1, patch2,
                                                                                      delta, n_img, Hi, Wi, P, mean, std);
    return bh::launch_status();
}
