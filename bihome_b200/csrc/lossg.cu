// K3g: the masked triplet loss of the reference in every variant beside the north-star one (which stays on K3, loss.cu),
// forward and backward fused: each feature element is read once (twice, out of L1/L2, where a per-pixel gate over all
// channels must be known before the gradient can be written in a planar layout) and each gradient element written once.
//
// Reference semantics (paths relative to /root/reference):
//   src/heads/PerceptualHead.py:465-538   one-line: l1 / cosine distances summed over channels, max(l1 - l3 + margin, 0),
//                                         weights m1'*m2 or m1' alone (MASK_CRD, :533-537)
//   src/heads/PerceptualHead.py:555-665   double-line: l1 (per channel) / l2 (channel mean of squares) / cosine, margin
//                                         'inf' (signed difference) or numeric, channel-aware (hinge per channel, then
//                                         the channel sum, :622-623) or channel-agnostic (hinge of the sums, :625-626),
//                                         + mu * ||H12 H21 - I||^2
//   src/heads/TripletHead.py:78-153       the same algebra on the content-aware backbone's one-channel full-resolution
//                                         maps with learned masks (all four masks and all four features carry gradient)
//
// Per sample and line (line 1: x = f1w, y = f2, weights W = a1*b2; line 2: x = f2w, y = f1, W = a2*b1; anchor (f1, f2)):
//   S = sum_p W_p    N = sum_p W_p g_p    ln = scale * N / max(S, 1)
//   g_p = sum_c(phi(x-y) - phi(f1-f2))                       hinge 0
//       = sum_c max(phi(x-y) - phi(f1-f2) + margin, 0)       hinge 1 (l1 only)
//       = max(d(x,y) - d(f1,f2) + margin, 0)                 hinge 2 (d = channel sum of |.|, channel mean of squares,
//                                                                     or 1 - cosine similarity)
// Three stream-ordered launches: mask sums per sample -> the streaming pass (feature gradients, g_p parked in the
// mask-gradient buffers) -> per-sample finish (numerators in a fixed order, loss, parts, dH, mask gradients in place).
#include "bh_common.cuh"

namespace bh {

constexpr int kTripThreads = 256;
constexpr float kCosEps = 1e-8f;   // torch.cosine_similarity's default eps

struct TripArgs {
    const float *f1, *f2, *f1w, *f2w;    // features; f2w unused when lines == 1
    const float *a1, *b2, *a2, *b1;      // masks [B,hw]; b* may be NULL (ones); a2 / b1 unused when lines == 1
    const float *H12, *H21;              // [B,9], lines == 2 only
    float *loss, *parts;                 // [B], [B,5] = ln1, ln2, S1, S2, ln3
    float *g_f1w, *g_f2w, *g_f1, *g_f2;  // g_f2w NULL when lines == 1; g_f1 / g_f2 optional (both or none)
    float *g_a1, *g_b2, *g_a2, *g_b1;    // g_a1 (and g_a2 when lines == 2) required: they hold the parked g_p
    float *gH12, *gH21;
    int B, C, hw, lines, hinge, crd;
    long long sc, sp;                    // channel / pixel strides in elements
    float m1, m2, scale1, scale2, mu, inv_c;
};

__device__ __forceinline__ float sgnf(float x) { return static_cast<float>(x > 0.0f) - static_cast<float>(x < 0.0f); }
template <int kDist>
__device__ __forceinline__ float phi(float e, float inv_c) { return kDist == 0 ? fabsf(e) : e * e * inv_c; }
template <int kDist>
__device__ __forceinline__ float dphi(float e, float inv_c) { return kDist == 0 ? sgnf(e) : 2.0f * e * inv_c; }

// ---- per-pixel state -------------------------------------------------------------------------------------------------
// l1 / l2: two channel sums (v1, v2) decide everything; cosine: three dot products and four squared norms.
template <int kDist>
struct PixAcc {
    float v1 = 0.0f, v2 = 0.0f;                                                    // kDist 0 / 1
    float d1w2 = 0.0f, d12 = 0.0f, d2w1 = 0.0f, n1w = 0.0f, n2 = 0.0f, n1 = 0.0f, n2w = 0.0f;   // kDist 2
    __device__ __forceinline__ void add(float x1w, float x2, float x2w, float x1, int hinge, float m1, float m2, float inv_c) {
        if (kDist == 2) {
            d1w2 = fmaf(x1w, x2, d1w2); d12 = fmaf(x1, x2, d12); d2w1 = fmaf(x2w, x1, d2w1);
            n1w = fmaf(x1w, x1w, n1w); n2 = fmaf(x2, x2, n2); n1 = fmaf(x1, x1, n1); n2w = fmaf(x2w, x2w, n2w);
        } else {
            const float p1 = phi<kDist>(x1w - x2, inv_c), p2 = phi<kDist>(x2w - x1, inv_c), p3 = phi<kDist>(x1 - x2, inv_c);
            if (hinge == 1) {
                v1 += fmaxf(p1 - p3 + m1, 0.0f);
                v2 += fmaxf(p2 - p3 + m2, 0.0f);
            } else {
                v1 += p1 - p3;
                v2 += p2 - p3;
            }
        }
    }
    // sum over the `width` lanes (power of two) that share a pixel
    __device__ __forceinline__ void reduce(int width) {
        for (int o = width >> 1; o > 0; o >>= 1) {
            if (kDist == 2) {
                d1w2 += __shfl_xor_sync(0xffffffffu, d1w2, o); d12 += __shfl_xor_sync(0xffffffffu, d12, o);
                d2w1 += __shfl_xor_sync(0xffffffffu, d2w1, o); n1w += __shfl_xor_sync(0xffffffffu, n1w, o);
                n2 += __shfl_xor_sync(0xffffffffu, n2, o); n1 += __shfl_xor_sync(0xffffffffu, n1, o);
                n2w += __shfl_xor_sync(0xffffffffu, n2w, o);
            } else {
                v1 += __shfl_xor_sync(0xffffffffu, v1, o);
                v2 += __shfl_xor_sync(0xffffffffu, v2, o);
            }
        }
    }
};

// what the gradient pass needs per pixel
template <int kDist>
struct PixCoef {
    float k1, k2;                                   // scale * W / max(S,1), zero where the pixel hinge is closed
    float r1w2, r12, r2w1;                          // cosine: 1 / (Nx Ny)
    float s1w_a, s2_a, s1_b, s2_b, s2w_c, s1_c;     // cosine: cos_xy / (|x| Nx) for x in pair a = (1w,2), b = (1,2), c = (2w,1)
};

// g_p of the two lines from the accumulated sums; fills the gradient coefficients
template <int kDist>
__device__ __forceinline__ void pixel_finish(const PixAcc<kDist>& s, int hinge, float m1, float m2, float kw1, float kw2, float& g1,
                                             float& g2, PixCoef<kDist>& c) {
    float v1, v2;
    if (kDist == 2) {
        const float l1w = sqrtf(s.n1w), l2 = sqrtf(s.n2), l1 = sqrtf(s.n1), l2w = sqrtf(s.n2w);
        const float N1w = fmaxf(l1w, kCosEps), N2 = fmaxf(l2, kCosEps), N1 = fmaxf(l1, kCosEps), N2w = fmaxf(l2w, kCosEps);
        c.r1w2 = 1.0f / (N1w * N2);
        c.r12 = 1.0f / (N1 * N2);
        c.r2w1 = 1.0f / (N2w * N1);
        const float ca = s.d1w2 * c.r1w2, cb = s.d12 * c.r12, cc = s.d2w1 * c.r2w1;
        c.s1w_a = l1w > 0.0f ? ca / (l1w * N1w) : 0.0f;
        c.s2_a = l2 > 0.0f ? ca / (l2 * N2) : 0.0f;
        c.s1_b = l1 > 0.0f ? cb / (l1 * N1) : 0.0f;
        c.s2_b = l2 > 0.0f ? cb / (l2 * N2) : 0.0f;
        c.s2w_c = l2w > 0.0f ? cc / (l2w * N2w) : 0.0f;
        c.s1_c = l1 > 0.0f ? cc / (l1 * N1) : 0.0f;
        v1 = cb - ca;   // (1 - cos(f1w,f2)) - (1 - cos(f1,f2))
        v2 = cb - cc;
    } else {
        v1 = s.v1;
        v2 = s.v2;
    }
    c.k1 = kw1;
    c.k2 = kw2;
    if (hinge == 2) {
        v1 += m1;
        v2 += m2;
        if (!(v1 > 0.0f)) { v1 = 0.0f; c.k1 = 0.0f; }
        if (!(v2 > 0.0f)) { v2 = 0.0f; c.k2 = 0.0f; }
    }
    g1 = v1;
    g2 = v2;
}

// gradients of one feature element: ga = d/df1w, gb = d/df2w, gc = d/df1, gd = d/df2
template <int kDist>
__device__ __forceinline__ void grad_elem(float x1w, float x2, float x2w, float x1, const PixCoef<kDist>& c, int hinge, float m1, float m2,
                                          float inv_c, float& ga, float& gb, float& gc, float& gd) {
    if (kDist == 2) {
        const float K = c.k1 + c.k2;
        ga = -c.k1 * (x2 * c.r1w2 - x1w * c.s1w_a);
        gb = -c.k2 * (x1 * c.r2w1 - x2w * c.s2w_c);
        gc = K * (x2 * c.r12 - x1 * c.s1_b) - c.k2 * (x2w * c.r2w1 - x1 * c.s1_c);
        gd = K * (x1 * c.r12 - x2 * c.s2_b) - c.k1 * (x1w * c.r1w2 - x2 * c.s2_a);
    } else {
        const float e1 = x1w - x2, e2 = x2w - x1, e3 = x1 - x2;
        float c1 = c.k1, c2 = c.k2;
        if (hinge == 1) {
            const float p3 = phi<kDist>(e3, inv_c);
            if (!(phi<kDist>(e1, inv_c) - p3 + m1 > 0.0f)) c1 = 0.0f;
            if (!(phi<kDist>(e2, inv_c) - p3 + m2 > 0.0f)) c2 = 0.0f;
        }
        // d/df1 = -c2 phi'(e2) - (c1 + c2) phi'(e3), d/df2 = -c1 phi'(e1) + (c1 + c2) phi'(e3), grouped by weight: when one
        // line's weight is tiny and the signs oppose, the large terms cancel exactly instead of leaving their rounding error
        const float p1 = dphi<kDist>(e1, inv_c), p2 = dphi<kDist>(e2, inv_c), p3 = dphi<kDist>(e3, inv_c);
        ga = c1 * p1;
        gb = c2 * p2;
        gc = -fmaf(c2, p2 + p3, c1 * p3);
        gd = fmaf(c1, p3 - p1, c2 * p3);
    }
}

// ---- launch 1: mask sums per sample ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTripThreads) triplet_den_kernel(const TripArgs a) {
    __shared__ float red[2 * (kTripThreads / 32)];
    const int b = blockIdx.x;
    const long long mbase = static_cast<long long>(b) * a.hw;
    float s[2] = {0.0f, 0.0f};
    for (int p = threadIdx.x; p < a.hw; p += kTripThreads) {
        const float y2 = (a.b2 && !a.crd) ? __ldg(a.b2 + mbase + p) : 1.0f;
        s[0] = fmaf(__ldg(a.a1 + mbase + p), y2, s[0]);
        if (a.lines == 2) {
            const float y1 = (a.b1 && !a.crd) ? __ldg(a.b1 + mbase + p) : 1.0f;
            s[1] = fmaf(__ldg(a.a2 + mbase + p), y1, s[1]);
        }
    }
    block_sum<2>(s, red);
    if (threadIdx.x == 0) {
        a.parts[b * 5 + 2] = s[0];
        a.parts[b * 5 + 3] = s[1];
    }
}

// the weights of pixel P (flat over samples) of both lines, already divided by max(S,1) and multiplied by the line scale
__device__ __forceinline__ void pixel_weights(const TripArgs& a, long long P, int b, float& kw1, float& kw2) {
    const float S1 = a.parts[b * 5 + 2], S2 = a.parts[b * 5 + 3];   // written by the previous launch: plain loads
    const float y2 = (a.b2 && !a.crd) ? __ldg(a.b2 + P) : 1.0f;
    kw1 = a.scale1 * __ldg(a.a1 + P) * y2 / fmaxf(S1, 1.0f);
    kw2 = 0.0f;
    if (a.lines == 2) {
        const float y1 = (a.b1 && !a.crd) ? __ldg(a.b1 + P) : 1.0f;
        kw2 = a.scale2 * __ldg(a.a2 + P) * y1 / fmaxf(S2, 1.0f);
    }
}

__device__ __forceinline__ float4 ld4(const float* p) { return ldg_stream(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float x, float y, float z, float w) { stg_stream(reinterpret_cast<float4*>(p), make_float4(x, y, z, w)); }

// ---- launch 2, channels-last: min(32, C/4 / kPerLane) lanes share a pixel, its 4 x C floats live in registers ------------
template <int kDist, int kPerLane>
__global__ void __launch_bounds__(kTripThreads) triplet_nhwc_kernel(const TripArgs a) {
    const int tid = threadIdx.x;
    const int tpp = (a.C >> 2) / kPerLane;                                // lanes per pixel: power of two, <= 32
    const int gl = tid & (tpp - 1), pg = tid / tpp, gpp = kTripThreads / tpp;
    const long long npix = static_cast<long long>(a.B) * a.hw;
    const long long stride = static_cast<long long>(gridDim.x) * gpp;
    const bool two = a.lines == 2;
    for (long long base = static_cast<long long>(blockIdx.x) * gpp; base < npix; base += stride) {   // uniform per CTA
        const long long P = base + pg;
        const bool live = P < npix;
        float4 q1w[kPerLane], q2[kPerLane], q2w[kPerLane], q1[kPerLane];
        float kw1 = 0.0f, kw2 = 0.0f;
        PixAcc<kDist> acc;
        if (live) {
#pragma unroll
            for (int k = 0; k < kPerLane; ++k) {
                const long long off = P * a.C + (gl + k * tpp) * 4;
                q1w[k] = ld4(a.f1w + off);
                q2[k] = ld4(a.f2 + off);
                q1[k] = ld4(a.f1 + off);
                q2w[k] = two ? ld4(a.f2w + off) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            }
            pixel_weights(a, P, static_cast<int>(P / a.hw), kw1, kw2);
#pragma unroll
            for (int k = 0; k < kPerLane; ++k) {
                acc.add(q1w[k].x, q2[k].x, q2w[k].x, q1[k].x, a.hinge, a.m1, a.m2, a.inv_c);
                acc.add(q1w[k].y, q2[k].y, q2w[k].y, q1[k].y, a.hinge, a.m1, a.m2, a.inv_c);
                acc.add(q1w[k].z, q2[k].z, q2w[k].z, q1[k].z, a.hinge, a.m1, a.m2, a.inv_c);
                acc.add(q1w[k].w, q2[k].w, q2w[k].w, q1[k].w, a.hinge, a.m1, a.m2, a.inv_c);
            }
        }
        acc.reduce(tpp);
        if (live) {
            float g1, g2;
            PixCoef<kDist> c;
            pixel_finish<kDist>(acc, a.hinge, a.m1, a.m2, kw1, kw2, g1, g2, c);
            if (gl == 0) {
                a.g_a1[P] = g1;
                if (two) a.g_a2[P] = g2;
            }
#pragma unroll
            for (int k = 0; k < kPerLane; ++k) {
                const long long off = P * a.C + (gl + k * tpp) * 4;
                float ga[4], gb[4], gc[4], gd[4];
                grad_elem<kDist>(q1w[k].x, q2[k].x, q2w[k].x, q1[k].x, c, a.hinge, a.m1, a.m2, a.inv_c, ga[0], gb[0], gc[0], gd[0]);
                grad_elem<kDist>(q1w[k].y, q2[k].y, q2w[k].y, q1[k].y, c, a.hinge, a.m1, a.m2, a.inv_c, ga[1], gb[1], gc[1], gd[1]);
                grad_elem<kDist>(q1w[k].z, q2[k].z, q2w[k].z, q1[k].z, c, a.hinge, a.m1, a.m2, a.inv_c, ga[2], gb[2], gc[2], gd[2]);
                grad_elem<kDist>(q1w[k].w, q2[k].w, q2w[k].w, q1[k].w, c, a.hinge, a.m1, a.m2, a.inv_c, ga[3], gb[3], gc[3], gd[3]);
                st4(a.g_f1w + off, ga[0], ga[1], ga[2], ga[3]);
                if (two) st4(a.g_f2w + off, gb[0], gb[1], gb[2], gb[3]);
                if (a.g_f1) {
                    st4(a.g_f1 + off, gc[0], gc[1], gc[2], gc[3]);
                    st4(a.g_f2 + off, gd[0], gd[1], gd[2], gd[3]);
                }
            }
        }
    }
}

// ---- launch 2, any strides: a thread owns V consecutive pixels (V = 4: planar, hw % 4 == 0) and walks the channels; a second
// walk (the re-reads hit L1 / L2) writes the gradients when they depend on a per-pixel result of the first ----------------
template <int V>
__device__ __forceinline__ void ldp(const float* p, float (&r)[V]) {
    if (V == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w;
    } else {
        r[0] = __ldg(p);
    }
}
template <int V>
__device__ __forceinline__ void stp(float* p, const float (&r)[V]) {
    if (V == 4) *reinterpret_cast<float4*>(p) = make_float4(r[0], r[1], r[2], r[3]);
    else *p = r[0];
}

template <int kDist, int V>
__global__ void __launch_bounds__(kTripThreads) triplet_strided_kernel(const TripArgs a) {
    const int nvec = a.hw / V;
    const long long gv = static_cast<long long>(blockIdx.x) * kTripThreads + threadIdx.x;
    if (gv >= static_cast<long long>(a.B) * nvec) return;
    const int b = static_cast<int>(gv / nvec);
    const int p0 = static_cast<int>(gv - static_cast<long long>(b) * nvec) * V;
    const long long P0 = static_cast<long long>(b) * a.hw + p0;
    const long long fbase = static_cast<long long>(b) * a.C * a.hw + static_cast<long long>(p0) * a.sp;
    const bool two = a.lines == 2;
    const bool two_pass = kDist == 2 || a.hinge == 2;
    float kw1[V], kw2[V];
#pragma unroll
    for (int i = 0; i < V; ++i) pixel_weights(a, P0 + i, b, kw1[i], kw2[i]);
    PixAcc<kDist> acc[V];
    PixCoef<kDist> coef[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        coef[i].k1 = kw1[i];
        coef[i].k2 = kw2[i];
    }
#pragma unroll 2
    for (int c = 0; c < a.C; ++c) {
        const long long o = fbase + c * a.sc;
        float x1w[V], x2[V], x2w[V], x1[V];
        ldp<V>(a.f1w + o, x1w);
        ldp<V>(a.f2 + o, x2);
        ldp<V>(a.f1 + o, x1);
        if (two) {
            ldp<V>(a.f2w + o, x2w);
        } else {
#pragma unroll
            for (int i = 0; i < V; ++i) x2w[i] = 0.0f;
        }
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i].add(x1w[i], x2[i], x2w[i], x1[i], a.hinge, a.m1, a.m2, a.inv_c);
        if (!two_pass) {
            float ga[V], gb[V], gc[V], gd[V];
#pragma unroll
            for (int i = 0; i < V; ++i) grad_elem<kDist>(x1w[i], x2[i], x2w[i], x1[i], coef[i], a.hinge, a.m1, a.m2, a.inv_c, ga[i], gb[i], gc[i], gd[i]);
            stp<V>(a.g_f1w + o, ga);
            if (two) stp<V>(a.g_f2w + o, gb);
            if (a.g_f1) {
                stp<V>(a.g_f1 + o, gc);
                stp<V>(a.g_f2 + o, gd);
            }
        }
    }
    float g1[V], g2[V];
#pragma unroll
    for (int i = 0; i < V; ++i) pixel_finish<kDist>(acc[i], a.hinge, a.m1, a.m2, kw1[i], kw2[i], g1[i], g2[i], coef[i]);
    stp<V>(a.g_a1 + P0, g1);
    if (two) stp<V>(a.g_a2 + P0, g2);
    if (!two_pass) return;
#pragma unroll 2
    for (int c = 0; c < a.C; ++c) {
        const long long o = fbase + c * a.sc;
        float x1w[V], x2[V], x2w[V], x1[V];
        ldp<V>(a.f1w + o, x1w);
        ldp<V>(a.f2 + o, x2);
        ldp<V>(a.f1 + o, x1);
        if (two) {
            ldp<V>(a.f2w + o, x2w);
        } else {
#pragma unroll
            for (int i = 0; i < V; ++i) x2w[i] = 0.0f;
        }
        float ga[V], gb[V], gc[V], gd[V];
#pragma unroll
        for (int i = 0; i < V; ++i) grad_elem<kDist>(x1w[i], x2[i], x2w[i], x1[i], coef[i], a.hinge, a.m1, a.m2, a.inv_c, ga[i], gb[i], gc[i], gd[i]);
        stp<V>(a.g_f1w + o, ga);
        if (two) stp<V>(a.g_f2w + o, gb);
        if (a.g_f1) {
            stp<V>(a.g_f1 + o, gc);
            stp<V>(a.g_f2 + o, gd);
        }
    }
}

// ---- launch 3: per-sample finish -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTripThreads) triplet_finish_kernel(const TripArgs a) {
    __shared__ float red[2 * (kTripThreads / 32)];
    const int b = blockIdx.x, tid = threadIdx.x;
    const long long mbase = static_cast<long long>(b) * a.hw;
    const bool two = a.lines == 2;
    const bool use_b = !a.crd;
    float n[2] = {0.0f, 0.0f};
    for (int p = tid; p < a.hw; p += kTripThreads) {
        const float y2 = (a.b2 && use_b) ? __ldg(a.b2 + mbase + p) : 1.0f;
        n[0] = fmaf(__ldg(a.a1 + mbase + p) * y2, a.g_a1[mbase + p], n[0]);
        if (two) {
            const float y1 = (a.b1 && use_b) ? __ldg(a.b1 + mbase + p) : 1.0f;
            n[1] = fmaf(__ldg(a.a2 + mbase + p) * y1, a.g_a2[mbase + p], n[1]);
        }
    }
    block_sum<2>(n, red);
    const float S1 = a.parts[b * 5 + 2], S2 = a.parts[b * 5 + 3];
    const float inv1 = 1.0f / fmaxf(S1, 1.0f), inv2 = 1.0f / fmaxf(S2, 1.0f);
    const float ln1 = a.scale1 * n[0] * inv1, ln2 = two ? a.scale2 * n[1] * inv2 : 0.0f;
    if (tid == 0) {
        float ln3 = 0.0f;
        if (two) {
            float h1[9], h2[9], E[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) { h1[i] = __ldg(a.H12 + b * 9 + i); h2[i] = __ldg(a.H21 + b * 9 + i); }
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const float e = h1[i * 3] * h2[j] + h1[i * 3 + 1] * h2[3 + j] + h1[i * 3 + 2] * h2[6 + j] - (i == j ? 1.0f : 0.0f);
                    E[i * 3 + j] = e;
                    ln3 = fmaf(e, e, ln3);
                }
            const float k = 2.0f * a.mu;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    a.gH12[b * 9 + i * 3 + j] = k * (E[i * 3] * h2[j * 3] + E[i * 3 + 1] * h2[j * 3 + 1] + E[i * 3 + 2] * h2[j * 3 + 2]);
                    a.gH21[b * 9 + i * 3 + j] = k * (h1[i] * E[j] + h1[3 + i] * E[3 + j] + h1[6 + i] * E[6 + j]);
                }
        }
        a.loss[b] = ln1 + ln2 + a.mu * ln3;
        a.parts[b * 5 + 0] = ln1;
        a.parts[b * 5 + 1] = ln2;
        a.parts[b * 5 + 4] = ln3;
    }
    // d ln / d W_p = scale * (g_p / den - [S > 1] N / den^2); W = a * b
    const float c1 = (S1 > 1.0f) ? n[0] * inv1 * inv1 : 0.0f, c2 = (S2 > 1.0f) ? n[1] * inv2 * inv2 : 0.0f;
    for (int p = tid; p < a.hw; p += kTripThreads) {
        {
            const float dW = a.scale1 * (a.g_a1[mbase + p] * inv1 - c1);
            const float y2 = (a.b2 && use_b) ? __ldg(a.b2 + mbase + p) : 1.0f;
            if (a.g_b2) a.g_b2[mbase + p] = use_b ? __ldg(a.a1 + mbase + p) * dW : 0.0f;
            a.g_a1[mbase + p] = y2 * dW;
        }
        if (two) {
            const float dW = a.scale2 * (a.g_a2[mbase + p] * inv2 - c2);
            const float y1 = (a.b1 && use_b) ? __ldg(a.b1 + mbase + p) : 1.0f;
            if (a.g_b1) a.g_b1[mbase + p] = use_b ? __ldg(a.a2 + mbase + p) * dW : 0.0f;
            a.g_a2[mbase + p] = y1 * dW;
        }
    }
}

struct TripGrads {
    float* feat[4];
    float* mask[4];
    float* h[2];
};
__global__ void __launch_bounds__(256) triplet_rescale_kernel(const float* __restrict__ gscale, const TripGrads g, int C, int hw) {
    const int b = blockIdx.y;
    const float s = __ldg(gscale + b);
    if (s == 1.0f) return;   // loss.backward() of the training loop: nothing to do
    const long long nf = static_cast<long long>(C) * hw;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nf; i += static_cast<long long>(gridDim.x) * blockDim.x) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (g.feat[k]) g.feat[k][b * nf + i] *= s;
        if (i < hw) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (g.mask[k]) g.mask[k][static_cast<long long>(b) * hw + i] *= s;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < 18) {
        float* h = g.h[threadIdx.x / 9];
        if (h) h[b * 9 + threadIdx.x % 9] *= s;
    }
}

template <int kDist>
int launch_triplet_main(const TripArgs& a, bool nhwc_regs, int per_lane, bool vec, cudaStream_t stream) {
    if (nhwc_regs) {
        const int gpp = kTripThreads / ((a.C >> 2) / per_lane);
        const long long tiles = (static_cast<long long>(a.B) * a.hw + gpp - 1) / gpp;
        const long long cap = static_cast<long long>(kNumSMs) * 8;
        const unsigned grid = static_cast<unsigned>(tiles < cap ? tiles : cap);
        if (per_lane == 1) triplet_nhwc_kernel<kDist, 1><<<grid, kTripThreads, 0, stream>>>(a);
        else triplet_nhwc_kernel<kDist, 2><<<grid, kTripThreads, 0, stream>>>(a);
        return launch_status();
    }
    const int V = vec ? 4 : 1;
    const long long nthreads = static_cast<long long>(a.B) * (a.hw / V);
    const unsigned grid = static_cast<unsigned>((nthreads + kTripThreads - 1) / kTripThreads);
    if (vec) triplet_strided_kernel<kDist, 4><<<grid, kTripThreads, 0, stream>>>(a);
    else triplet_strided_kernel<kDist, 1><<<grid, kTripThreads, 0, stream>>>(a);
    return launch_status();
}

}  // namespace bh

extern "C" int bh_triplet_fwd_bwd(const float* f1, const float* f2, const float* f1w, const float* f2w, const float* a1, const float* b2,
                                  const float* a2, const float* b1, const float* H12, const float* H21, int lines, int distance,
                                  int hinge, int mask_crd, float margin1, float margin2, float scale1, float scale2, float mu, float* loss,
                                  float* parts, float* g_f1w, float* g_f2w, float* g_f1, float* g_f2, float* g_a1, float* g_b2, float* g_a2,
                                  float* g_b1, float* gH12, float* gH21, int B, int C, int h, int w, int channels_last,
                                  bh_stream_t stream_) {
    using namespace bh;
    if (!f1 || !f2 || !f1w || !a1 || !loss || !parts || !g_f1w || !g_a1) return BH_E_NULL;
    if (lines != 1 && lines != 2) return BH_E_UNSUPPORTED;
    if (lines == 2 && (!f2w || !a2 || !H12 || !H21 || !g_f2w || !g_a2 || !gH12 || !gH21)) return BH_E_NULL;
    if ((g_f1 == nullptr) != (g_f2 == nullptr)) return BH_E_NULL;
    if (distance < 0 || distance > 2 || hinge < 0 || hinge > 2) return BH_E_UNSUPPORTED;
    if (hinge == 1 && distance != 0) return BH_E_UNSUPPORTED;   // a per-channel margin needs per-channel distances
    if (B <= 0 || C <= 0 || h <= 0 || w <= 0) return BH_E_SHAPE;
    TripArgs a;
    a.f1 = f1; a.f2 = f2; a.f1w = f1w; a.f2w = f2w; a.a1 = a1; a.b2 = b2; a.a2 = a2; a.b1 = b1; a.H12 = H12; a.H21 = H21;
    a.loss = loss; a.parts = parts; a.g_f1w = g_f1w; a.g_f2w = g_f2w; a.g_f1 = g_f1; a.g_f2 = g_f2;
    a.g_a1 = g_a1; a.g_b2 = g_b2; a.g_a2 = g_a2; a.g_b1 = g_b1; a.gH12 = gH12; a.gH21 = gH21;
    a.B = B; a.C = C; a.hw = h * w; a.lines = lines; a.hinge = hinge; a.crd = mask_crd ? 1 : 0;
    const bool cl = channels_last && C > 1;
    a.sc = cl ? 1 : a.hw;
    a.sp = cl ? C : 1;
    a.m1 = margin1; a.m2 = margin2; a.scale1 = scale1; a.scale2 = scale2; a.mu = lines == 2 ? mu : 0.0f;
    a.inv_c = 1.0f / static_cast<float>(C);
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    const bool al = aligned16(f1) && aligned16(f2) && aligned16(f1w) && (!f2w || aligned16(f2w)) && aligned16(g_f1w) &&
                    (!g_f2w || aligned16(g_f2w)) && (!g_f1 || aligned16(g_f1)) && (!g_f2 || aligned16(g_f2)) && aligned16(g_a1) &&
                    (!g_a2 || aligned16(g_a2));
    const int quads = C / 4;
    const bool pow2 = (C % 4) == 0 && quads > 0 && (quads & (quads - 1)) == 0 && quads <= 64;
    const bool nhwc_regs = cl && pow2 && al;
    const int per_lane = quads > 32 ? 2 : 1;
    const bool vec = !cl && (a.hw % 4) == 0 && al;

    triplet_den_kernel<<<B, kTripThreads, 0, stream>>>(a);
    int rc = launch_status();
    if (rc != BH_OK) return rc;
    if (distance == 0) rc = launch_triplet_main<0>(a, nhwc_regs, per_lane, vec, stream);
    else if (distance == 1) rc = launch_triplet_main<1>(a, nhwc_regs, per_lane, vec, stream);
    else rc = launch_triplet_main<2>(a, nhwc_regs, per_lane, vec, stream);
    if (rc != BH_OK) return rc;
    triplet_finish_kernel<<<B, kTripThreads, 0, stream>>>(a);
    return launch_status();
}

extern "C" int bh_triplet_rescale(const float* gscale, float* g_f1w, float* g_f2w, float* g_f1, float* g_f2, float* g_a1, float* g_b2,
                                  float* g_a2, float* g_b1, float* gH12, float* gH21, int B, int C, int h, int w, bh_stream_t stream_) {
    using namespace bh;
    if (!gscale) return BH_E_NULL;
    if (B <= 0 || C <= 0 || h <= 0 || w <= 0) return BH_E_SHAPE;
    TripGrads g;
    g.feat[0] = g_f1w; g.feat[1] = g_f2w; g.feat[2] = g_f1; g.feat[3] = g_f2;
    g.mask[0] = g_a1; g.mask[1] = g_b2; g.mask[2] = g_a2; g.mask[3] = g_b1;
    g.h[0] = gH12; g.h[1] = gH21;
    const long long nf = static_cast<long long>(C) * h * w;
    long long gx = (nf + 256 * 8 - 1) / (256 * 8);
    if (gx > 64) gx = 64;
    if (gx < 1) gx = 1;
    dim3 grid(static_cast<unsigned>(gx), B);
    triplet_rescale_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(gscale, g, C, h * w);
    return launch_status();
}
