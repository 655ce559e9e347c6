// K3: bidirectional perceptual (biHomE) loss, forward and backward fused in ONE pass over the features.
//
// Reference semantics: src/heads/PerceptualHead.py:559-561 (l1 = |f1' - f2|, l2 = |f2' - f1|, l3 = |f1 - f2|),
// :609-665 (double-line, TRIPLET_MARGIN 'inf', channel-agnostic):
//   W1 = m1'*m2   S1 = sum_hw W1   ln1 = sum_hw W1*(sum_c l1 - sum_c l3) / max(S1, 1)        (same for 2)
//   ln3 = ||H12 H21 - I||_F^2      loss_b = ln1 + ln2 + mu*ln3
// and its autograd:
//   d/df1'   = W1 * sign(f1' - f2) / den1
//   d/dm1'   = m2 * D1 / den1 - [S1 > 1] * num1 * m2 / den1^2
//   d/dH12   = 2 mu E H21^T,  d/dH21 = 2 mu H12^T E,  E = H12 H21 - I
//
// Work decomposition: one thread-block CLUSTER per sample; the CTAs of a cluster split the h*w pixels and
// exchange their partial mask sums (before the pass, for den) and numerator sums (after it, for d/dm)
// through distributed shared memory.  Each feature element is read once and each gradient element written
// once: 4*C*h*w*4 B in, 2*C*h*w*4 B out per sample -- the compulsory traffic.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "bh_common.cuh"

namespace cg = cooperative_groups;

namespace bh {

struct LossArgs {
    const float *f1, *f2, *f1w, *f2w, *m1, *m2, *m1w, *m2w, *H12, *H21;
    float mu;
    float *loss, *parts, *g_f1w, *g_f2w, *g_f1, *g_f2, *g_m1w, *g_m2w, *gH12, *gH21;
    int B, C, hw;
    long long sc, sp;  // channel / pixel strides in elements (NCHW: hw, 1; NHWC: 1, C)
};

template <int V>
__device__ __forceinline__ void ldv(const float* p, float (&r)[V]) {
    if (V == 4) {
        const float4 t = ldg_stream(reinterpret_cast<const float4*>(p));
        r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w;
    } else {
        r[0] = __ldg(p);
    }
}
template <int V>
__device__ __forceinline__ void ldv_cached(const float* p, float (&r)[V]) {
    if (V == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        r[0] = t.x; r[1] = t.y; r[2] = t.z; r[3] = t.w;
    } else {
        r[0] = __ldg(p);
    }
}
template <int V>
__device__ __forceinline__ void stv(float* p, const float (&r)[V]) {
    if (V == 4) stg_stream(reinterpret_cast<float4*>(p), make_float4(r[0], r[1], r[2], r[3]));
    else *p = r[0];
}
__device__ __forceinline__ float sgn(float x) { return static_cast<float>(x > 0.0f) - static_cast<float>(x < 0.0f); }

constexpr int kLossThreads = 256;
constexpr int kLanes = 64;  // pixel-vectors per chunk; kLossThreads / kLanes channel groups
constexpr int kGroups = kLossThreads / kLanes;

// one feature element: distances, gradients
template <bool kInputGrads>
__device__ __forceinline__ void feat_elem(float p1w, float p2, float p2w, float p1, float k1, float k2, float& d1, float& d2,
                                          float& ga, float& gb, float& gc, float& gd) {
    const float e1 = p1w - p2, e2 = p2w - p1, e3 = p1 - p2;
    const float a3 = fabsf(e3);
    d1 += fabsf(e1) - a3;
    d2 += fabsf(e2) - a3;
    ga = k1 * sgn(e1);
    gb = k2 * sgn(e2);
    if (kInputGrads) {
        const float s3 = (k1 + k2) * sgn(e3);
        gc = -gb - s3;  // d/df1
        gd = -ga + s3;  // d/df2
    }
}

// loss, TensorBoard parts and the H12 / H21 gradients of one sample (one thread)
__device__ __forceinline__ void sample_tail(const LossArgs& a, int b, float N1, float N2, float S1, float S2, float inv1, float inv2) {
    float h1[9], h2[9], E[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { h1[i] = __ldg(a.H12 + b * 9 + i); h2[i] = __ldg(a.H21 + b * 9 + i); }
    float ln3 = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float e = h1[i * 3] * h2[j] + h1[i * 3 + 1] * h2[3 + j] + h1[i * 3 + 2] * h2[6 + j] - (i == j ? 1.0f : 0.0f);
            E[i * 3 + j] = e;
            ln3 = fmaf(e, e, ln3);
        }
    const float ln1 = N1 * inv1, ln2 = N2 * inv2;
    a.loss[b] = ln1 + ln2 + a.mu * ln3;
    a.parts[b * 5 + 0] = ln1; a.parts[b * 5 + 1] = ln2; a.parts[b * 5 + 2] = S1; a.parts[b * 5 + 3] = S2;
    a.parts[b * 5 + 4] = ln3;
    const float k = 2.0f * a.mu;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            // (E H21^T)[i][j] = sum_k E[i][k] H21[j][k] ; (H12^T E)[i][j] = sum_k H12[k][i] E[k][j]
            a.gH12[b * 9 + i * 3 + j] = k * (E[i * 3] * h2[j * 3] + E[i * 3 + 1] * h2[j * 3 + 1] + E[i * 3 + 2] * h2[j * 3 + 2]);
            a.gH21[b * 9 + i * 3 + j] = k * (h1[i] * E[j] + h1[3 + i] * E[3 + j] + h1[6 + i] * E[6 + j]);
        }
}

// V = pixels per thread-vector for the mask passes and the NCHW feature pass (4: hw % 4 == 0; 1: any strides).
// kNHWC: features are channels-last with C % 4 == 0 and C/4 a power of two: a group of min(32, C/4) lanes owns a pixel,
// every lane streams float4 channel quads (512 contiguous bytes per warp), the channel sum is a shuffle reduction.
template <int V, bool kInputGrads, bool kNHWC>
__global__ void __launch_bounds__(kLossThreads) bihome_kernel(const LossArgs a) {
    cg::cluster_group cluster = cg::this_cluster();
    const int CL = static_cast<int>(cluster.num_blocks());
    const int rank = static_cast<int>(cluster.block_rank());
    const int b = blockIdx.x / CL;
    const int tid = threadIdx.x, lane = tid & (kLanes - 1), grp = tid / kLanes;

    __shared__ float xchg_den[2];
    __shared__ float xchg_num[2];
    __shared__ float red[2 * (kLossThreads / 32)];
    __shared__ float dsm[kGroups][kLanes][2 * V];

    const int nvec = a.hw / V;
    const int v0 = static_cast<int>(static_cast<long long>(rank) * nvec / CL);
    const int v1 = static_cast<int>(static_cast<long long>(rank + 1) * nvec / CL);
    const long long mbase = static_cast<long long>(b) * a.hw;
    const long long fbase = static_cast<long long>(b) * a.C * a.hw;

    // ---- phase 0: mask sums of the whole sample (den1, den2) via DSMEM -------------------------------
    float s[2] = {0.0f, 0.0f};
    for (int pv = v0 + tid; pv < v1; pv += kLossThreads) {
        float x1[V], x2[V], y1[V], y2[V];
        ldv_cached<V>(a.m1w + mbase + pv * V, x1);
        ldv_cached<V>(a.m2w + mbase + pv * V, x2);
#pragma unroll
        for (int i = 0; i < V; ++i) y1[i] = y2[i] = 1.0f;
        if (a.m1) ldv_cached<V>(a.m1 + mbase + pv * V, y1);
        if (a.m2) ldv_cached<V>(a.m2 + mbase + pv * V, y2);
#pragma unroll
        for (int i = 0; i < V; ++i) {
            s[0] = fmaf(x1[i], y2[i], s[0]);
            s[1] = fmaf(x2[i], y1[i], s[1]);
        }
    }
    block_sum<2>(s, red);
    if (tid == 0) { xchg_den[0] = s[0]; xchg_den[1] = s[1]; }
    cluster.sync();
    float S1 = 0.0f, S2 = 0.0f;
    for (int r = 0; r < CL; ++r) {
        const float* remote = cluster.map_shared_rank(xchg_den, r);
        S1 += remote[0];
        S2 += remote[1];
    }
    const float inv1 = 1.0f / fmaxf(S1, 1.0f), inv2 = 1.0f / fmaxf(S2, 1.0f);

    // ---- phase 1: the streaming pass ---------------------------------------------------------------
    float num[2] = {0.0f, 0.0f};
    if (kNHWC) {
        const int quads = a.C >> 2;
        const int tpp = quads < 32 ? quads : 32;          // lanes per pixel (power of two)
        const int per_lane = quads / tpp;                 // channel quads per lane
        const int gl = tid & (tpp - 1), pg = tid / tpp, gpp = kLossThreads / tpp;
        const int p0 = v0 * V, p1e = v1 * V;
        const int iters = (p1e - p0 + gpp - 1) / gpp;
        for (int it = 0; it < iters; ++it) {
            const int p = p0 + it * gpp + pg;
            const bool live = p < p1e;
            float W1 = 0.0f, W2 = 0.0f, d1 = 0.0f, d2 = 0.0f;
            if (live) {
                const float x1 = __ldg(a.m1w + mbase + p), x2 = __ldg(a.m2w + mbase + p);
                const float y1 = a.m1 ? __ldg(a.m1 + mbase + p) : 1.0f, y2 = a.m2 ? __ldg(a.m2 + mbase + p) : 1.0f;
                W1 = x1 * y2;
                W2 = x2 * y1;
                const float k1 = W1 * inv1, k2 = W2 * inv2;
                const long long poff = fbase + static_cast<long long>(p) * a.C;
#pragma unroll 2
                for (int k = 0; k < per_lane; ++k) {
                    const long long o = poff + (gl + k * tpp) * 4;
                    float p1w[4], p2[4], p2w[4], p1[4], ga[4], gb[4], gc[4], gd[4];
                    ldv<4>(a.f1w + o, p1w);
                    ldv<4>(a.f2 + o, p2);
                    ldv<4>(a.f2w + o, p2w);
                    ldv<4>(a.f1 + o, p1);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        feat_elem<kInputGrads>(p1w[i], p2[i], p2w[i], p1[i], k1, k2, d1, d2, ga[i], gb[i], gc[i], gd[i]);
                    stv<4>(a.g_f1w + o, ga);
                    stv<4>(a.g_f2w + o, gb);
                    if (kInputGrads) {
                        stv<4>(a.g_f1 + o, gc);
                        stv<4>(a.g_f2 + o, gd);
                    }
                }
            }
            for (int o = tpp >> 1; o > 0; o >>= 1) {
                d1 += __shfl_xor_sync(0xffffffffu, d1, o);
                d2 += __shfl_xor_sync(0xffffffffu, d2, o);
            }
            if (live && gl == 0) {
                num[0] = fmaf(W1, d1, num[0]);
                num[1] = fmaf(W2, d2, num[1]);
                a.g_m1w[mbase + p] = d1;  // parked D, see phase 3
                a.g_m2w[mbase + p] = d2;
            }
        }
    } else {
        for (int cb = v0; cb < v1; cb += kLanes) {
            const int pv = cb + lane;
            const bool live = pv < v1;
            float W1[V], W2[V], d1[V], d2[V];
    #pragma unroll
            for (int i = 0; i < V; ++i) W1[i] = W2[i] = d1[i] = d2[i] = 0.0f;
            if (live) {
                float x1[V], x2[V], y1[V], y2[V];
                ldv_cached<V>(a.m1w + mbase + pv * V, x1);
                ldv_cached<V>(a.m2w + mbase + pv * V, x2);
    #pragma unroll
                for (int i = 0; i < V; ++i) y1[i] = y2[i] = 1.0f;
                if (a.m1) ldv_cached<V>(a.m1 + mbase + pv * V, y1);
                if (a.m2) ldv_cached<V>(a.m2 + mbase + pv * V, y2);
    #pragma unroll
                for (int i = 0; i < V; ++i) { W1[i] = x1[i] * y2[i]; W2[i] = x2[i] * y1[i]; }
                const long long poff = fbase + static_cast<long long>(pv) * V * a.sp;
    #pragma unroll 2
                for (int c = grp; c < a.C; c += kGroups) {
                    const long long o = poff + c * a.sc;
                    float p1w[V], p2[V], p2w[V], p1[V];
                    ldv<V>(a.f1w + o, p1w);
                    ldv<V>(a.f2 + o, p2);
                    ldv<V>(a.f2w + o, p2w);
                    ldv<V>(a.f1 + o, p1);
                    float ga[V], gb[V], gc[V], gd[V];
    #pragma unroll
                    for (int i = 0; i < V; ++i)
                        feat_elem<kInputGrads>(p1w[i], p2[i], p2w[i], p1[i], W1[i] * inv1, W2[i] * inv2, d1[i], d2[i], ga[i], gb[i],
                                               gc[i], gd[i]);
                    stv<V>(a.g_f1w + o, ga);
                    stv<V>(a.g_f2w + o, gb);
                    if (kInputGrads) {
                        stv<V>(a.g_f1 + o, gc);
                        stv<V>(a.g_f2 + o, gd);
                    }
                }
            }
    #pragma unroll
            for (int i = 0; i < V; ++i) { dsm[grp][lane][i] = d1[i]; dsm[grp][lane][V + i] = d2[i]; }
            __syncthreads();
            if (grp == 0 && live) {
                float D1[V], D2[V];
    #pragma unroll
                for (int i = 0; i < V; ++i) {
                    float t1 = 0.0f, t2 = 0.0f;
    #pragma unroll
                    for (int g = 0; g < kGroups; ++g) { t1 += dsm[g][lane][i]; t2 += dsm[g][lane][V + i]; }
                    D1[i] = t1; D2[i] = t2;
                    num[0] = fmaf(W1[i], t1, num[0]);
                    num[1] = fmaf(W2[i], t2, num[1]);
                }
                // park D in the mask-gradient buffers; phase 3 turns it into the gradient in place
                if (V == 4) {
                    *reinterpret_cast<float4*>(a.g_m1w + mbase + pv * V) = make_float4(D1[0], D1[1], D1[2], D1[3]);
                    *reinterpret_cast<float4*>(a.g_m2w + mbase + pv * V) = make_float4(D2[0], D2[1], D2[2], D2[3]);
                } else {
                    a.g_m1w[mbase + pv] = D1[0];
                    a.g_m2w[mbase + pv] = D2[0];
                }
            }
            __syncthreads();
        }
    }

    // ---- phase 2: numerators of the whole sample via DSMEM -------------------------------------------
    block_sum<2>(num, red);
    if (tid == 0) { xchg_num[0] = num[0]; xchg_num[1] = num[1]; }
    cluster.sync();
    float N1 = 0.0f, N2 = 0.0f;
    for (int r = 0; r < CL; ++r) {
        const float* remote = cluster.map_shared_rank(xchg_num, r);
        N1 += remote[0];
        N2 += remote[1];
    }
    if (rank == 0 && tid == 0) sample_tail(a, b, N1, N2, S1, S2, inv1, inv2);

    // ---- phase 3: mask gradients, in place over the parked D ------------------------------------------
    const float c1 = (S1 > 1.0f) ? N1 * inv1 * inv1 : 0.0f, c2 = (S2 > 1.0f) ? N2 * inv2 * inv2 : 0.0f;
    for (int pv = v0 + tid; pv < v1; pv += kLossThreads) {
        float y1[V], y2[V];
#pragma unroll
        for (int i = 0; i < V; ++i) y1[i] = y2[i] = 1.0f;
        if (a.m1) ldv_cached<V>(a.m1 + mbase + pv * V, y1);
        if (a.m2) ldv_cached<V>(a.m2 + mbase + pv * V, y2);
#pragma unroll
        for (int i = 0; i < V; ++i) {
            float* p1 = a.g_m1w + mbase + pv * V + i;
            float* p2 = a.g_m2w + mbase + pv * V + i;
            *p1 = y2[i] * (*p1 * inv1 - c1);
            *p2 = y1[i] * (*p2 * inv2 - c2);
        }
    }
    cluster.sync();  // keep this CTA's shared memory alive until every peer has read it
}

// ---- channels-last, TMA-staged variant ---------------------------------------------------------------------------
// The register file caps what the kernel above can keep in flight (every outstanding float4 holds four registers of a
// 63-register thread: ~64 KB per SM).  Here the four feature streams of a CTA's pixel range -- contiguous in NHWC -- are
// pulled into a shared-memory ring by 1-D bulk TMA copies (cp.async.bulk, SASS UBLKCP) issued by one thread, kLossStages
// tiles ahead, starting BEFORE the mask sums of phase 0 so that the first tiles land while the denominators are being
// exchanged.  A tile is one float4 per thread and tensor (kPerLane = 1) or two (kPerLane = 2, C = 256).
constexpr int kLossStages = 4;
template <bool kInputGrads, int kPerLane>
__global__ void __launch_bounds__(kLossThreads) bihome_nhwc_tma_kernel(const LossArgs a) {
    cg::cluster_group cluster = cg::this_cluster();
    const int CL = static_cast<int>(cluster.num_blocks());
    const int rank = static_cast<int>(cluster.block_rank());
    const int b = blockIdx.x / CL;
    const int tid = threadIdx.x;

    extern __shared__ __align__(128) float tiles[];   // [stage][tensor: f1w, f2, f2w, f1][tile_floats]
    __shared__ uint64_t full[kLossStages];
    __shared__ float xchg_den[2];
    __shared__ float xchg_num[2];
    __shared__ float red[2 * (kLossThreads / 32)];

    const int nvec = a.hw / 4;
    const int v0 = static_cast<int>(static_cast<long long>(rank) * nvec / CL);
    const int v1 = static_cast<int>(static_cast<long long>(rank + 1) * nvec / CL);
    const long long mbase = static_cast<long long>(b) * a.hw;
    const long long fbase = static_cast<long long>(b) * a.C * a.hw;

    const int quads = a.C >> 2;
    const int tpp = quads / kPerLane;                 // lanes per pixel (power of two, <= 32)
    const int gl = tid & (tpp - 1), pg = tid / tpp, gpp = kLossThreads / tpp;   // pixels per tile
    const int tile_floats = gpp * a.C;
    const int p0 = v0 * 4, p1e = v1 * 4;
    const int n_tiles = (p1e - p0 + gpp - 1) / gpp;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kLossStages; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    auto issue = [&](int i) {   // one thread: tile i -> stage i % kLossStages
        const int s = i % kLossStages;
        const int px0 = p0 + i * gpp;
        const int npx = min(gpp, p1e - px0);
        const uint32_t bytes = static_cast<uint32_t>(npx) * a.C * 4u;
        mbar_expect_tx(&full[s], 4u * bytes);
        float* dst = tiles + static_cast<size_t>(s) * 4 * tile_floats;
        const long long off = fbase + static_cast<long long>(px0) * a.C;
        bulk_g2s(dst, a.f1w + off, bytes, &full[s]);
        bulk_g2s(dst + tile_floats, a.f2 + off, bytes, &full[s]);
        bulk_g2s(dst + 2 * tile_floats, a.f2w + off, bytes, &full[s]);
        bulk_g2s(dst + 3 * tile_floats, a.f1 + off, bytes, &full[s]);
    };
    if (tid == 0)
        for (int i = 0; i < kLossStages && i < n_tiles; ++i) issue(i);

    // ---- phase 0: mask sums of the whole sample (den1, den2) via DSMEM; the first tiles are already on their way ----
    float s2[2] = {0.0f, 0.0f};
    for (int pv = v0 + tid; pv < v1; pv += kLossThreads) {
        float x1[4], x2[4], y1[4], y2[4];
        ldv_cached<4>(a.m1w + mbase + pv * 4, x1);
        ldv_cached<4>(a.m2w + mbase + pv * 4, x2);
#pragma unroll
        for (int i = 0; i < 4; ++i) y1[i] = y2[i] = 1.0f;
        if (a.m1) ldv_cached<4>(a.m1 + mbase + pv * 4, y1);
        if (a.m2) ldv_cached<4>(a.m2 + mbase + pv * 4, y2);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            s2[0] = fmaf(x1[i], y2[i], s2[0]);
            s2[1] = fmaf(x2[i], y1[i], s2[1]);
        }
    }
    block_sum<2>(s2, red);
    if (tid == 0) { xchg_den[0] = s2[0]; xchg_den[1] = s2[1]; }
    cluster.sync();
    float S1 = 0.0f, S2 = 0.0f;
    for (int r = 0; r < CL; ++r) {
        const float* remote = cluster.map_shared_rank(xchg_den, r);
        S1 += remote[0];
        S2 += remote[1];
    }
    const float inv1 = 1.0f / fmaxf(S1, 1.0f), inv2 = 1.0f / fmaxf(S2, 1.0f);

    // ---- phase 1: the streaming pass over the ring -------------------------------------------------------------
    float num[2] = {0.0f, 0.0f};
    for (int i = 0; i < n_tiles; ++i) {
        const int s = i % kLossStages;
        const int p = p0 + i * gpp + pg;
        const bool live = p < p1e;
        float W1 = 0.0f, W2 = 0.0f;
        if (live) {   // the pixel's mask weights: issued before the wait so that their latency overlaps it
            const float x1 = __ldg(a.m1w + mbase + p), x2 = __ldg(a.m2w + mbase + p);
            const float y1 = a.m1 ? __ldg(a.m1 + mbase + p) : 1.0f, y2 = a.m2 ? __ldg(a.m2 + mbase + p) : 1.0f;
            W1 = x1 * y2;
            W2 = x2 * y1;
        }
        mbar_wait(&full[s], static_cast<uint32_t>(i / kLossStages) & 1u);
        const float* st = tiles + static_cast<size_t>(s) * 4 * tile_floats + pg * a.C + gl * 4;
        float4 q1w[kPerLane], q2[kPerLane], q2w[kPerLane], q1[kPerLane];
#pragma unroll
        for (int k = 0; k < kPerLane; ++k) {
            const float* e = st + k * tpp * 4;
            q1w[k] = *reinterpret_cast<const float4*>(e);
            q2[k] = *reinterpret_cast<const float4*>(e + tile_floats);
            q2w[k] = *reinterpret_cast<const float4*>(e + 2 * tile_floats);
            q1[k] = *reinterpret_cast<const float4*>(e + 3 * tile_floats);
        }
        __syncthreads();   // every thread holds its part of the tile in registers: the stage can be refilled
        if (tid == 0 && i + kLossStages < n_tiles) {
            fence_proxy_async();
            issue(i + kLossStages);
        }
        float d1 = 0.0f, d2 = 0.0f;
        if (live) {
            const float k1 = W1 * inv1, k2 = W2 * inv2;
            const long long poff = fbase + static_cast<long long>(p) * a.C;
#pragma unroll
            for (int k = 0; k < kPerLane; ++k) {
                const long long o = poff + (gl + k * tpp) * 4;
                const float p1w[4] = {q1w[k].x, q1w[k].y, q1w[k].z, q1w[k].w}, p2[4] = {q2[k].x, q2[k].y, q2[k].z, q2[k].w};
                const float p2w[4] = {q2w[k].x, q2w[k].y, q2w[k].z, q2w[k].w}, p1[4] = {q1[k].x, q1[k].y, q1[k].z, q1[k].w};
                float ga[4], gb[4], gc[4], gd[4];
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    feat_elem<kInputGrads>(p1w[e], p2[e], p2w[e], p1[e], k1, k2, d1, d2, ga[e], gb[e], gc[e], gd[e]);
                stv<4>(a.g_f1w + o, ga);
                stv<4>(a.g_f2w + o, gb);
                if (kInputGrads) {
                    stv<4>(a.g_f1 + o, gc);
                    stv<4>(a.g_f2 + o, gd);
                }
            }
        }
        for (int o = tpp >> 1; o > 0; o >>= 1) {
            d1 += __shfl_xor_sync(0xffffffffu, d1, o);
            d2 += __shfl_xor_sync(0xffffffffu, d2, o);
        }
        if (live && gl == 0) {
            num[0] = fmaf(W1, d1, num[0]);
            num[1] = fmaf(W2, d2, num[1]);
            a.g_m1w[mbase + p] = d1;  // parked D, see phase 3
            a.g_m2w[mbase + p] = d2;
        }
    }

    // ---- phase 2: numerators of the whole sample via DSMEM, loss, dH ----------------------------------------------
    block_sum<2>(num, red);
    if (tid == 0) { xchg_num[0] = num[0]; xchg_num[1] = num[1]; }
    cluster.sync();
    float N1 = 0.0f, N2 = 0.0f;
    for (int r = 0; r < CL; ++r) {
        const float* remote = cluster.map_shared_rank(xchg_num, r);
        N1 += remote[0];
        N2 += remote[1];
    }
    if (rank == 0 && tid == 0) sample_tail(a, b, N1, N2, S1, S2, inv1, inv2);

    // ---- phase 3: mask gradients, in place over the parked D --------------------------------------------------------
    const float c1 = (S1 > 1.0f) ? N1 * inv1 * inv1 : 0.0f, c2 = (S2 > 1.0f) ? N2 * inv2 * inv2 : 0.0f;
    for (int pv = v0 + tid; pv < v1; pv += kLossThreads) {
        float y1[4], y2[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) y1[i] = y2[i] = 1.0f;
        if (a.m1) ldv_cached<4>(a.m1 + mbase + pv * 4, y1);
        if (a.m2) ldv_cached<4>(a.m2 + mbase + pv * 4, y2);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float* q1p = a.g_m1w + mbase + pv * 4 + i;
            float* q2p = a.g_m2w + mbase + pv * 4 + i;
            *q1p = y2[i] * (*q1p * inv1 - c1);
            *q2p = y1[i] * (*q2p * inv2 - c2);
        }
    }
    cluster.sync();  // keep this CTA's shared memory alive until every peer has read it
}

// ---- channels-last, persistent TMA-staged variant ("stream") ------------------------------------------------------
// No clusters: 3 CTAs per SM walk a CONTIGUOUS range of (sample, pixel-tile) work items each (balanced to one tile, one
// wave, no tail), the ring of bulk copies runs across tile and sample boundaries, and the only per-sample state a CTA
// needs up front -- the two mask sums -- it computes itself when it enters a sample (8 KB of L2 hits).  The numerators,
// the loss, dH and the mask gradients are left to bihome_finish_kernel (one CTA per sample, a fixed-order reduction over
// the D values this kernel parks in the mask-gradient buffers): nothing crosses CTAs here.
template <bool kInputGrads, int kPerLane>
__global__ void __launch_bounds__(kLossThreads) bihome_stream_kernel(const LossArgs a, int tiles_per_sample, long long n_tiles_total) {
    const int tid = threadIdx.x;
    extern __shared__ __align__(128) float tiles[];   // [stage][tensor: f1w, f2, f2w, f1][tile_floats]
    __shared__ uint64_t full[kLossStages];
    __shared__ float red[2 * (kLossThreads / 32)];

    const int quads = a.C >> 2;
    const int tpp = quads / kPerLane;                 // lanes per pixel (power of two, <= 32)
    const int gl = tid & (tpp - 1), pg = tid / tpp, gpp = kLossThreads / tpp;   // gpp = pixels per tile
    const int tile_floats = gpp * a.C;
    const long long t0 = n_tiles_total * blockIdx.x / gridDim.x, t1 = n_tiles_total * (blockIdx.x + 1) / gridDim.x;
    const int n_tiles = static_cast<int>(t1 - t0);

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kLossStages; ++s) mbar_init(&full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    auto issue = [&](int i) {   // one thread: local tile i -> stage i % kLossStages
        const int s = i % kLossStages;
        const long long t = t0 + i;
        const int b = static_cast<int>(t / tiles_per_sample);
        const int px0 = static_cast<int>(t - static_cast<long long>(b) * tiles_per_sample) * gpp;
        const int npx = min(gpp, a.hw - px0);
        const uint32_t bytes = static_cast<uint32_t>(npx) * a.C * 4u;
        mbar_expect_tx(&full[s], 4u * bytes);
        float* dst = tiles + static_cast<size_t>(s) * 4 * tile_floats;
        const long long off = (static_cast<long long>(b) * a.hw + px0) * a.C;
        bulk_g2s(dst, a.f1w + off, bytes, &full[s]);
        bulk_g2s(dst + tile_floats, a.f2 + off, bytes, &full[s]);
        bulk_g2s(dst + 2 * tile_floats, a.f2w + off, bytes, &full[s]);
        bulk_g2s(dst + 3 * tile_floats, a.f1 + off, bytes, &full[s]);
    };
    if (tid == 0)
        for (int i = 0; i < kLossStages && i < n_tiles; ++i) issue(i);

    int cur_b = -1;
    float inv1 = 0.0f, inv2 = 0.0f;
    for (int i = 0; i < n_tiles; ++i) {
        const int s = i % kLossStages;
        const long long t = t0 + i;
        const int b = static_cast<int>(t / tiles_per_sample);
        const long long mbase = static_cast<long long>(b) * a.hw;
        if (b != cur_b) {   // entering a sample (uniform across the CTA): its two mask sums
            cur_b = b;
            float s2[2] = {0.0f, 0.0f};
            for (int pv = tid; pv < a.hw / 4; pv += kLossThreads) {
                float x1[4], x2[4], y1[4], y2[4];
                ldv_cached<4>(a.m1w + mbase + pv * 4, x1);
                ldv_cached<4>(a.m2w + mbase + pv * 4, x2);
#pragma unroll
                for (int e = 0; e < 4; ++e) y1[e] = y2[e] = 1.0f;
                if (a.m1) ldv_cached<4>(a.m1 + mbase + pv * 4, y1);
                if (a.m2) ldv_cached<4>(a.m2 + mbase + pv * 4, y2);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    s2[0] = fmaf(x1[e], y2[e], s2[0]);
                    s2[1] = fmaf(x2[e], y1[e], s2[1]);
                }
            }
            block_sum<2>(s2, red);
            inv1 = 1.0f / fmaxf(s2[0], 1.0f);
            inv2 = 1.0f / fmaxf(s2[1], 1.0f);
        }
        const int p = static_cast<int>(t - static_cast<long long>(b) * tiles_per_sample) * gpp + pg;
        const bool live = p < a.hw;
        float W1 = 0.0f, W2 = 0.0f;
        if (live) {   // the pixel's mask weights: requested before the wait so that their latency overlaps it
            const float x1 = __ldg(a.m1w + mbase + p), x2 = __ldg(a.m2w + mbase + p);
            const float y1 = a.m1 ? __ldg(a.m1 + mbase + p) : 1.0f, y2 = a.m2 ? __ldg(a.m2 + mbase + p) : 1.0f;
            W1 = x1 * y2;
            W2 = x2 * y1;
        }
        mbar_wait(&full[s], static_cast<uint32_t>(i / kLossStages) & 1u);
        const float* st = tiles + static_cast<size_t>(s) * 4 * tile_floats + pg * a.C + gl * 4;
        float4 q1w[kPerLane], q2[kPerLane], q2w[kPerLane], q1[kPerLane];
#pragma unroll
        for (int k = 0; k < kPerLane; ++k) {
            const float* e = st + k * tpp * 4;
            q1w[k] = *reinterpret_cast<const float4*>(e);
            q2[k] = *reinterpret_cast<const float4*>(e + tile_floats);
            q2w[k] = *reinterpret_cast<const float4*>(e + 2 * tile_floats);
            q1[k] = *reinterpret_cast<const float4*>(e + 3 * tile_floats);
        }
        __syncthreads();   // every thread holds its part of the tile in registers: the stage can be refilled
        if (tid == 0 && i + kLossStages < n_tiles) {
            fence_proxy_async();
            issue(i + kLossStages);
        }
        float d1 = 0.0f, d2 = 0.0f;
        if (live) {
            const float k1 = W1 * inv1, k2 = W2 * inv2;
            const long long poff = (mbase + p) * a.C;
#pragma unroll
            for (int k = 0; k < kPerLane; ++k) {
                const long long o = poff + (gl + k * tpp) * 4;
                const float p1w[4] = {q1w[k].x, q1w[k].y, q1w[k].z, q1w[k].w}, p2[4] = {q2[k].x, q2[k].y, q2[k].z, q2[k].w};
                const float p2w[4] = {q2w[k].x, q2w[k].y, q2w[k].z, q2w[k].w}, p1[4] = {q1[k].x, q1[k].y, q1[k].z, q1[k].w};
                float ga[4], gb[4], gc[4], gd[4];
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    feat_elem<kInputGrads>(p1w[e], p2[e], p2w[e], p1[e], k1, k2, d1, d2, ga[e], gb[e], gc[e], gd[e]);
                stv<4>(a.g_f1w + o, ga);
                stv<4>(a.g_f2w + o, gb);
                if (kInputGrads) {
                    stv<4>(a.g_f1 + o, gc);
                    stv<4>(a.g_f2 + o, gd);
                }
            }
        }
        for (int o = tpp >> 1; o > 0; o >>= 1) {
            d1 += __shfl_xor_sync(0xffffffffu, d1, o);
            d2 += __shfl_xor_sync(0xffffffffu, d2, o);
        }
        if (live && gl == 0) {   // parked D: bihome_finish_kernel turns it into the mask gradient in place
            a.g_m1w[mbase + p] = d1;
            a.g_m2w[mbase + p] = d2;
        }
    }
}

// One CTA per sample, after bihome_stream_kernel: mask sums and numerators in a fixed order, loss / parts / dH, and the
// mask gradients in place over the parked D.
__global__ void __launch_bounds__(kLossThreads) bihome_finish_kernel(const LossArgs a) {
    __shared__ float red[2 * (kLossThreads / 32)];
    const int b = blockIdx.x, tid = threadIdx.x;
    const long long mbase = static_cast<long long>(b) * a.hw;
    // the mask sums exactly as bihome_stream_kernel forms them (same order): loss and gradients see the same denominators
    float s2[2] = {0.0f, 0.0f};
    for (int pv = tid; pv < a.hw / 4; pv += kLossThreads) {
        float x1[4], x2[4], y1[4], y2[4];
        ldv_cached<4>(a.m1w + mbase + pv * 4, x1);
        ldv_cached<4>(a.m2w + mbase + pv * 4, x2);
#pragma unroll
        for (int e = 0; e < 4; ++e) y1[e] = y2[e] = 1.0f;
        if (a.m1) ldv_cached<4>(a.m1 + mbase + pv * 4, y1);
        if (a.m2) ldv_cached<4>(a.m2 + mbase + pv * 4, y2);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            s2[0] = fmaf(x1[e], y2[e], s2[0]);
            s2[1] = fmaf(x2[e], y1[e], s2[1]);
        }
    }
    block_sum<2>(s2, red);
    float n2[2] = {0.0f, 0.0f};
    for (int p = tid; p < a.hw; p += kLossThreads) {
        const float x1 = __ldg(a.m1w + mbase + p), x2 = __ldg(a.m2w + mbase + p);
        const float y1 = a.m1 ? __ldg(a.m1 + mbase + p) : 1.0f, y2 = a.m2 ? __ldg(a.m2 + mbase + p) : 1.0f;
        n2[0] = fmaf(x1 * y2, a.g_m1w[mbase + p], n2[0]);
        n2[1] = fmaf(x2 * y1, a.g_m2w[mbase + p], n2[1]);
    }
    block_sum<2>(n2, red);
    const float S1 = s2[0], S2 = s2[1], N1 = n2[0], N2 = n2[1];
    const float inv1 = 1.0f / fmaxf(S1, 1.0f), inv2 = 1.0f / fmaxf(S2, 1.0f);
    if (tid == 0) sample_tail(a, b, N1, N2, S1, S2, inv1, inv2);
    const float c1 = (S1 > 1.0f) ? N1 * inv1 * inv1 : 0.0f, c2 = (S2 > 1.0f) ? N2 * inv2 * inv2 : 0.0f;
    for (int p = tid; p < a.hw; p += kLossThreads) {
        const float y1 = a.m1 ? __ldg(a.m1 + mbase + p) : 1.0f, y2 = a.m2 ? __ldg(a.m2 + mbase + p) : 1.0f;
        a.g_m1w[mbase + p] = y2 * (a.g_m1w[mbase + p] * inv1 - c1);
        a.g_m2w[mbase + p] = y1 * (a.g_m2w[mbase + p] * inv2 - c2);
    }
}

__global__ void __launch_bounds__(256)
    bihome_rescale_kernel(const float* __restrict__ gscale, float* g_f1w, float* g_f2w, float* g_f1, float* g_f2, float* g_m1w,
                          float* g_m2w, float* gH12, float* gH21, int C, int hw) {
    const int b = blockIdx.y;
    const float s = __ldg(gscale + b);
    if (s == 1.0f) return;  // the training loop's loss.backward(): nothing to do
    const long long nf = static_cast<long long>(C) * hw;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nf;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        g_f1w[b * nf + i] *= s;
        g_f2w[b * nf + i] *= s;
        if (g_f1) g_f1[b * nf + i] *= s;
        if (g_f2) g_f2[b * nf + i] *= s;
        if (i < hw) { g_m1w[static_cast<long long>(b) * hw + i] *= s; g_m2w[static_cast<long long>(b) * hw + i] *= s; }
        if (i < 9) { gH12[b * 9 + i] *= s; gH21[b * 9 + i] *= s; }
    }
}

template <int V, bool G, bool N>
int launch_bihome(const LossArgs& a, int CL, cudaStream_t stream) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(a.B) * CL);
    cfg.blockDim = dim3(kLossThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, bihome_kernel<V, G, N>, a);
    ++g_launch_count;
    if (e != cudaSuccess) return static_cast<int>(e);
    e = cudaGetLastError();
    return e == cudaSuccess ? BH_OK : static_cast<int>(e);
}

template <bool G, int kPerLane>
int launch_bihome_tma(const LossArgs& a, int CL, size_t smem, cudaStream_t stream) {
    cudaError_t e;
    static size_t granted = 0;   // per instantiation; the attribute only ever grows (racing callers set the same value)
    if (smem > granted) {
        e = cudaFuncSetAttribute(bihome_nhwc_tma_kernel<G, kPerLane>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return static_cast<int>(e);
        granted = smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(a.B) * CL);
    cfg.blockDim = dim3(kLossThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, bihome_nhwc_tma_kernel<G, kPerLane>, a);
    ++g_launch_count;
    if (e != cudaSuccess) return static_cast<int>(e);
    e = cudaGetLastError();
    return e == cudaSuccess ? BH_OK : static_cast<int>(e);
}

template <bool G, int kPerLane>
int launch_bihome_stream(const LossArgs& a, size_t smem, int gpp, cudaStream_t stream) {
    cudaError_t e;
    static size_t granted = 0;   // per instantiation; the attribute only ever grows (racing callers set the same value)
    if (smem > granted) {
        e = cudaFuncSetAttribute(bihome_stream_kernel<G, kPerLane>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return static_cast<int>(e);
        granted = smem;
    }
    const int tiles_per_sample = (a.hw + gpp - 1) / gpp;
    const long long n_tiles = static_cast<long long>(a.B) * tiles_per_sample;
    const long long resident = static_cast<long long>(kNumSMs) * (smem <= 72 * 1024 ? 3 : 1);
    const unsigned grid = static_cast<unsigned>(n_tiles < resident ? n_tiles : resident);
    bihome_stream_kernel<G, kPerLane><<<grid, kLossThreads, smem, stream>>>(a, tiles_per_sample, n_tiles);
    int rc = launch_status();
    if (rc != BH_OK) return rc;
    bihome_finish_kernel<<<a.B, kLossThreads, 0, stream>>>(a);
    return launch_status();
}

}  // namespace bh

extern "C" int bh_bihome_fwd_bwd(const float* f1, const float* f2, const float* f1w, const float* f2w, const float* m1,
                                 const float* m2, const float* m1w, const float* m2w, const float* H12, const float* H21,
                                 float mu, float* loss, float* parts, float* g_f1w, float* g_f2w, float* g_f1, float* g_f2,
                                 float* g_m1w, float* g_m2w, float* gH12, float* gH21, int B, int C, int h, int w,
                                 int channels_last, bh_stream_t stream_) {
    using namespace bh;
    if (!f1 || !f2 || !f1w || !f2w || !m1w || !m2w || !H12 || !H21 || !loss || !parts || !g_f1w || !g_f2w || !g_m1w ||
        !g_m2w || !gH12 || !gH21)
        return BH_E_NULL;
    if ((g_f1 == nullptr) != (g_f2 == nullptr)) return BH_E_NULL;
    if (B <= 0 || C <= 0 || h <= 0 || w <= 0) return BH_E_SHAPE;
    LossArgs a;
    a.f1 = f1; a.f2 = f2; a.f1w = f1w; a.f2w = f2w; a.m1 = m1; a.m2 = m2; a.m1w = m1w; a.m2w = m2w; a.H12 = H12; a.H21 = H21;
    a.mu = mu; a.loss = loss; a.parts = parts; a.g_f1w = g_f1w; a.g_f2w = g_f2w; a.g_f1 = g_f1; a.g_f2 = g_f2;
    a.g_m1w = g_m1w; a.g_m2w = g_m2w; a.gH12 = gH12; a.gH21 = gH21; a.B = B; a.C = C; a.hw = h * w;
    a.sc = channels_last ? 1 : a.hw;
    a.sp = channels_last ? C : 1;
    const bool vec = (a.hw % 4) == 0 && aligned16(f1) && aligned16(f2) && aligned16(f1w) && aligned16(f2w) &&
                     aligned16(g_f1w) && aligned16(g_f2w) && aligned16(m1w) && aligned16(m2w) && aligned16(g_m1w) &&
                     aligned16(g_m2w) && (!m1 || aligned16(m1)) && (!m2 || aligned16(m2)) && (!g_f1 || aligned16(g_f1)) &&
                     (!g_f2 || aligned16(g_f2));
    // cluster size: enough CTAs to cover the 148 SMs a few times over, never more than the pixel-vectors allow
    const int nvec = vec ? a.hw / 4 : a.hw;
    int CL = (B >= 2 * kNumSMs) ? 2 : (B >= kNumSMs ? 4 : 8);
    while (CL > 1 && nvec / CL < kLanes) CL >>= 1;
    {   // microbenchmark switch (bh_tune_set "loss_cluster"): cluster size 1, 2, 4 or 8; 0 = the rule above
        const int f = g_tune[kTuneLossCluster];
        if (f == 1 || f == 2 || f == 4 || f == 8) CL = f;
    }
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (channels_last) {
        const int quads = C / 4;
        const bool pow2 = (C % 4) == 0 && quads > 0 && (quads & (quads - 1)) == 0 && quads <= 256;
        // Measured on B200 (tools/microbench.py --loss-cl, C = 64, h = w = 32): the persistent stream wins while the whole
        // job is a few dozen tiles per CTA (B = 256: 84 % of the HBM peak against 72 % for the cluster kernels, which run
        // 2.3 waves there); from B = 1024 on the clustered TMA kernel does (94 - 100 % against 86 - 90 %).
        const int force = g_tune[kTuneLossVariant];   // microbenchmark switch (bh_tune_set "loss_variant"): 1 ldg, 2 cluster, 3 stream
        const bool want_stream = force ? force == 3 : B < 512;
        const bool want_tma = force ? force != 1 : true;
        if (vec && pow2 && quads <= 64 && want_stream) {
            // persistent TMA-staged stream + per-sample finish
            const int per_lane = quads > 32 ? 2 : 1;
            const int gpp = kLossThreads / (quads / per_lane);
            const size_t smem = static_cast<size_t>(kLossStages) * 4 * gpp * C * sizeof(float);
            if (per_lane == 1) return g_f1 ? launch_bihome_stream<true, 1>(a, smem, gpp, stream) : launch_bihome_stream<false, 1>(a, smem, gpp, stream);
            return g_f1 ? launch_bihome_stream<true, 2>(a, smem, gpp, stream) : launch_bihome_stream<false, 2>(a, smem, gpp, stream);
        }
        if (vec && pow2 && quads <= 64 && want_tma) {
            // TMA-staged ring: one float4 per thread and tensor per tile (two for C = 256)
            const int per_lane = quads > 32 ? 2 : 1;
            const size_t smem = static_cast<size_t>(kLossStages) * 4 * (kLossThreads / (quads / per_lane)) * C * sizeof(float);
            if (per_lane == 1) return g_f1 ? launch_bihome_tma<true, 1>(a, CL, smem, stream) : launch_bihome_tma<false, 1>(a, CL, smem, stream);
            return g_f1 ? launch_bihome_tma<true, 2>(a, CL, smem, stream) : launch_bihome_tma<false, 2>(a, CL, smem, stream);
        }
        if (vec && pow2) return g_f1 ? launch_bihome<4, true, true>(a, CL, stream) : launch_bihome<4, false, true>(a, CL, stream);
        return g_f1 ? launch_bihome<1, true, false>(a, CL, stream) : launch_bihome<1, false, false>(a, CL, stream);
    }
    if (vec) return g_f1 ? launch_bihome<4, true, false>(a, CL, stream) : launch_bihome<4, false, false>(a, CL, stream);
    return g_f1 ? launch_bihome<1, true, false>(a, CL, stream) : launch_bihome<1, false, false>(a, CL, stream);
}

extern "C" int bh_bihome_rescale(const float* gscale, float* g_f1w, float* g_f2w, float* g_f1, float* g_f2, float* g_m1w,
                                 float* g_m2w, float* gH12, float* gH21, int B, int C, int h, int w, bh_stream_t stream_) {
    using namespace bh;
    if (!gscale || !g_f1w || !g_f2w || !g_m1w || !g_m2w || !gH12 || !gH21) return BH_E_NULL;
    if (B <= 0 || C <= 0 || h <= 0 || w <= 0) return BH_E_SHAPE;
    const long long nf = static_cast<long long>(C) * h * w;
    long long gx = (nf + 256 * 8 - 1) / (256 * 8);
    if (gx > 64) gx = 64;
    if (nf < 9) return BH_E_SHAPE;
    dim3 grid(static_cast<unsigned>(gx), B);
    bihome_rescale_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(gscale, g_f1w, g_f2w, g_f1, g_f2, g_m1w,
                                                                                      g_m2w, gH12, gH21, C, h * w);
    return launch_status();
}
