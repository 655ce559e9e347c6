// K2 / K2b: fused homography-grid generation + bilinear sampling (forward) and its adjoint w.r.t. H.
//
// Reference semantics: src/data/utils.py:54-59 warp_image(inverse=True) ==
//   out[b,c,y,x] = sum_{4 taps} w_tap * src[b,c,tap],  (u,v) = proj(H_b [x,y,1]^T), zeros padding,
// i.e. F.grid_sample(bilinear, zeros, align_corners=True) on the grid kornia.warp_perspective builds.
// The sampling grid never exists in memory here: coordinates live in registers.
//
// Three code paths, chosen by shape (see bh_warp_fwd / bh_warp_bwd at the bottom):
//   plane  : NCHW, a whole source plane fits in shared memory.  Persistent CTAs; planes are staged by
//            1-D bulk TMA (cp.async.bulk -> UBLKCP) through an mbarrier ring so the copy of plane i+1
//            overlaps the sampling of plane i; taps are read from shared memory; each thread owns 4x4
//            output pixels (float4 row stores, pooled 4x4 coverage mask falls out in-thread).
//   nhwc   : channels-last, C % 4 == 0: one thread per (pixel, 4 channels), 128-bit coalesced tap loads.
//   generic: anything else (scalar, strided).
// Backward produces dH by a fixed-order per-sample reduction (bit-reproducible, no atomics); the image
// gradient (dead work on the biHomE path: the source never requires grad) is an optional red.global.add.
#include "bh_common.cuh"

namespace bh {

// analytic coverage of warp(ones): m = mx(u) * my(v)
__device__ __forceinline__ float cover(const Taps& t) {
    const float mx = (t.inx0 ? t.wx0 : 0.0f) + (t.inx1 ? t.wx1 : 0.0f);
    const float my = (t.iny0 ? t.wy0 : 0.0f) + (t.iny1 ? t.wy1 : 0.0f);
    return mx * my;
}
// d cover / du, d cover / dv
__device__ __forceinline__ void cover_grad(const Taps& t, float& du, float& dv) {
    const float mx = (t.inx0 ? t.wx0 : 0.0f) + (t.inx1 ? t.wx1 : 0.0f);
    const float my = (t.iny0 ? t.wy0 : 0.0f) + (t.iny1 ? t.wy1 : 0.0f);
    du = my * ((t.inx1 ? 1.0f : 0.0f) - (t.inx0 ? 1.0f : 0.0f));
    dv = mx * ((t.iny1 ? 1.0f : 0.0f) - (t.iny0 ? 1.0f : 0.0f));
}

// four tap values of a plane with row pitch `pitch` elements and element stride `es`
template <typename Ptr>
__device__ __forceinline__ void gather4(Ptr p, const Taps& t, int pitch, int es, float& nw, float& ne, float& sw,
                                        float& se) {
    const int o = t.y0 * pitch + t.x0 * es;
    nw = (t.inx0 && t.iny0) ? p[o] : 0.0f;
    ne = (t.inx1 && t.iny0) ? p[o + es] : 0.0f;
    sw = (t.inx0 && t.iny1) ? p[o + pitch] : 0.0f;
    se = (t.inx1 && t.iny1) ? p[o + pitch + es] : 0.0f;
}
__device__ __forceinline__ float blend(const Taps& t, float nw, float ne, float sw, float se) {
    return fmaf(fmaf(nw, t.wx0, ne * t.wx1), t.wy0, fmaf(sw, t.wx0, se * t.wx1) * t.wy1);
}
// d out / du, d out / dv for unit upstream (ATen grid_sampler_2d_backward, pixel units)
__device__ __forceinline__ void blend_grad(const Taps& t, float nw, float ne, float sw, float se, float& du, float& dv) {
    du = fmaf(ne - nw, t.wy0, (se - sw) * t.wy1);
    dv = fmaf(sw - nw, t.wx0, (se - ne) * t.wx1);
}
// accumulate the dH contribution of one pixel: (gu, gv) = d loss / d(u, v)
__device__ __forceinline__ void accum_gh(float (&acc)[9], float gu, float gv, float u, float v, float rw, float x,
                                         float y) {
    const float a = gu * rw, b = gv * rw, c = -fmaf(gu, u, gv * v) * rw;
    acc[0] = fmaf(a, x, acc[0]); acc[1] = fmaf(a, y, acc[1]); acc[2] += a;
    acc[3] = fmaf(b, x, acc[3]); acc[4] = fmaf(b, y, acc[4]); acc[5] += b;
    acc[6] = fmaf(c, x, acc[6]); acc[7] = fmaf(c, y, acc[7]); acc[8] += c;
}

// after block_sum every thread holds the 9 totals; thread k stores element k (no dynamic register indexing)
__device__ __forceinline__ void store9(const float (&acc)[9], float* dst) {
#pragma unroll
    for (int k = 0; k < 9; ++k)
        if (threadIdx.x == k) dst[k] = acc[k];
}

// =================================================================================================
// plane path (NCHW, TMA-staged source planes)
// =================================================================================================
constexpr int kPlaneThreads = 512;
constexpr int kMaxStages = 3;

struct PlaneRing {
    unsigned char* base;
    uint32_t plane_bytes;
    uint64_t* full;  // [stages]
    __device__ __forceinline__ float* stage(int s) const { return reinterpret_cast<float*>(base + static_cast<size_t>(s) * plane_bytes); }
};

__device__ __forceinline__ PlaneRing ring_setup(unsigned char* smem, uint32_t plane_bytes, int stages) {
    PlaneRing r;
    r.base = smem;
    r.plane_bytes = plane_bytes;
    r.full = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(stages) * plane_bytes);
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&r.full[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    return r;
}

// out tiles: thread owns a 4x4 block of output pixels.  kMask: also emit the 4x4-pooled coverage mask.
template <bool kMask>
__global__ void __launch_bounds__(kPlaneThreads, 1)
    warp_fwd_plane_kernel(const float* __restrict__ src, const float* __restrict__ H, float* __restrict__ out,
                          float* __restrict__ mask_pooled, int n_planes, int C, int Hs, int Ws, int Ho, int Wo,
                          int stages) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int plane_elems = Hs * Ws;
    const uint32_t plane_bytes = static_cast<uint32_t>(plane_elems) * 4u;
    PlaneRing ring = ring_setup(smem_raw, plane_bytes, stages);

    const int first = blockIdx.x, stride = gridDim.x;
    const int nloc = (n_planes - first + stride - 1) / stride;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages && i < nloc; ++i) {
            mbar_expect_tx(&ring.full[i], plane_bytes);
            bulk_g2s(ring.stage(i), src + static_cast<size_t>(first + i * stride) * plane_elems, plane_bytes, &ring.full[i]);
        }
    }
    const int tiles_x = Wo >> 2, n_tiles = tiles_x * (Ho >> 2);
    for (int it = 0; it < nloc; ++it) {
        const int s = it % stages;
        const int plane = first + it * stride;
        const int b = plane / C;
        const Hmat hm = load_h(H, b);
        mbar_wait(&ring.full[s], (it / stages) & 1);
        const float* sp = ring.stage(s);
        float* op = out + static_cast<size_t>(plane) * Ho * Wo;
        const bool emit_mask = kMask && (plane - b * C == 0);
        for (int tile = threadIdx.x; tile < n_tiles; tile += kPlaneThreads) {
            const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
            float msum = 0.0f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float y = static_cast<float>(ty * 4 + j);
                float o4[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float x = static_cast<float>(tx * 4 + i);
                    float u, v, rw;
                    project(hm, x, y, u, v, rw);
                    const Taps t = make_taps(u, v, Ws, Hs);
                    float nw, ne, sw, se;
                    gather4(sp, t, Ws, 1, nw, ne, sw, se);
                    o4[i] = blend(t, nw, ne, sw, se);
                    if (kMask) msum += cover(t);
                }
                stg_stream(reinterpret_cast<float4*>(op + (ty * 4 + j) * Wo + tx * 4), make_float4(o4[0], o4[1], o4[2], o4[3]));
            }
            if (emit_mask) mask_pooled[static_cast<size_t>(b) * n_tiles + tile] = msum * 0.0625f;
        }
        __syncthreads();  // every thread is done with stage s
        if (threadIdx.x == 0 && it + stages < nloc) {
            mbar_expect_tx(&ring.full[s], plane_bytes);
            bulk_g2s(ring.stage(s), src + static_cast<size_t>(first + (it + stages) * stride) * plane_elems, plane_bytes,
                     &ring.full[s]);
        }
    }
}

// Backward of the plane path.  Work unit = sample b (all C planes), so that dH[b] is reduced by one CTA
// in a fixed order.  kMask: add the pooled-mask term (pool == 4 tiles, upstream gMaskPooled).
template <bool kImage, bool kMask>
__global__ void __launch_bounds__(kPlaneThreads, 1)
    warp_bwd_plane_kernel(const float* __restrict__ src, const float* __restrict__ H, const float* __restrict__ gOut,
                          const float* __restrict__ gMaskPooled, float* __restrict__ gH, int B, int C, int Hs, int Ws,
                          int Ho, int Wo, int stages) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ float red[9 * (kPlaneThreads / 32)];
    const int plane_elems = Hs * Ws;
    const uint32_t plane_bytes = static_cast<uint32_t>(plane_elems) * 4u;
    const int first = blockIdx.x, stride = gridDim.x;
    const int nsamp = (B - first + stride - 1) / stride;
    const int nloc = kImage ? nsamp * C : 0;  // planes this CTA streams: sample-major, then channel
    PlaneRing ring = {nullptr, 0u, nullptr};
    if (kImage) {
        ring = ring_setup(smem_raw, plane_bytes, stages);
        if (threadIdx.x == 0) {
            for (int i = 0; i < stages && i < nloc; ++i) {
                const int pl = (first + (i / C) * stride) * C + (i % C);
                mbar_expect_tx(&ring.full[i], plane_bytes);
                bulk_g2s(ring.stage(i), src + static_cast<size_t>(pl) * plane_elems, plane_bytes, &ring.full[i]);
            }
        }
    }
    const int tiles_x = Wo >> 2, n_tiles = tiles_x * (Ho >> 2);
    int it = 0;
    for (int si = 0; si < nsamp; ++si) {
        const int b = first + si * stride;
        const Hmat hm = load_h(H, b);
        float acc[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) acc[k] = 0.0f;
        if (kMask) {
            for (int tile = threadIdx.x; tile < n_tiles; tile += kPlaneThreads) {
                const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
                const float gm = __ldg(gMaskPooled + static_cast<size_t>(b) * n_tiles + tile) * 0.0625f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float x = static_cast<float>(tx * 4 + i), y = static_cast<float>(ty * 4 + j);
                        float u, v, rw, du, dv;
                        project(hm, x, y, u, v, rw);
                        const Taps t = make_taps(u, v, Ws, Hs);
                        cover_grad(t, du, dv);
                        if (du != 0.0f || dv != 0.0f) accum_gh(acc, gm * du, gm * dv, u, v, rw, x, y);
                    }
                }
            }
        }
        if (kImage) {
            for (int c = 0; c < C; ++c, ++it) {
                const int s = it % stages;
                mbar_wait(&ring.full[s], (it / stages) & 1);
                const float* sp = ring.stage(s);
                const float* gp = gOut + (static_cast<size_t>(b) * C + c) * Ho * Wo;
                for (int tile = threadIdx.x; tile < n_tiles; tile += kPlaneThreads) {
                    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
                    float4 g4[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        g4[j] = ldg_stream(reinterpret_cast<const float4*>(gp + (ty * 4 + j) * Wo + tx * 4));
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float gj[4] = {g4[j].x, g4[j].y, g4[j].z, g4[j].w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float x = static_cast<float>(tx * 4 + i), y = static_cast<float>(ty * 4 + j);
                            float u, v, rw, du, dv, nw, ne, sw, se;
                            project(hm, x, y, u, v, rw);
                            const Taps t = make_taps(u, v, Ws, Hs);
                            gather4(sp, t, Ws, 1, nw, ne, sw, se);
                            blend_grad(t, nw, ne, sw, se, du, dv);
                            accum_gh(acc, gj[i] * du, gj[i] * dv, u, v, rw, x, y);
                        }
                    }
                }
                __syncthreads();
                if (threadIdx.x == 0 && it + stages < nloc) {
                    const int nx = it + stages;
                    const int pl = (first + (nx / C) * stride) * C + (nx % C);
                    mbar_expect_tx(&ring.full[s], plane_bytes);
                    bulk_g2s(ring.stage(s), src + static_cast<size_t>(pl) * plane_elems, plane_bytes, &ring.full[s]);
                }
            }
        }
        block_sum<9>(acc, red);
        store9(acc, gH + b * 9);
    }
}

// =================================================================================================
// generic + nhwc paths
// =================================================================================================
struct Layout {
    long long sb;   // batch stride
    int sc, sy, sx; // channel, row, column strides (elements)
};
__host__ __device__ inline Layout make_layout(int C, int Hh, int Ww, int channels_last) {
    Layout l;
    l.sb = static_cast<long long>(C) * Hh * Ww;
    if (channels_last) { l.sc = 1; l.sx = C; l.sy = Ww * C; }
    else { l.sc = Hh * Ww; l.sx = 1; l.sy = Ww; }
    return l;
}

// analytic pooled mask, any pool (one thread per pooled cell)
__global__ void mask_pooled_kernel(const float* __restrict__ H, float* __restrict__ mask_pooled, int B, int Hs, int Ws,
                                   int Ho, int Wo, int pool) {
    const int hp = Ho / pool, wp = Wo / pool;
    const long long n = static_cast<long long>(B) * hp * wp;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < n;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(idx / (hp * wp));
        const int r = static_cast<int>(idx - static_cast<long long>(b) * hp * wp);
        const int py = r / wp, px = r - py * wp;
        const Hmat hm = load_h(H, b);
        float s = 0.0f;
        for (int j = 0; j < pool; ++j)
            for (int i = 0; i < pool; ++i) {
                float u, v, rw;
                project(hm, static_cast<float>(px * pool + i), static_cast<float>(py * pool + j), u, v, rw);
                s += cover(make_taps(u, v, Ws, Hs));
            }
        mask_pooled[idx] = s / static_cast<float>(pool * pool);
    }
}

// one thread per output pixel, loop over channels
__global__ void warp_fwd_generic_kernel(const float* __restrict__ src, const float* __restrict__ H, float* __restrict__ out,
                                        int B, int C, int Hs, int Ws, int Ho, int Wo, Layout ls, Layout lo) {
    const long long n = static_cast<long long>(B) * Ho * Wo;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < n;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(idx / (Ho * Wo));
        const int r = static_cast<int>(idx - static_cast<long long>(b) * Ho * Wo);
        const int y = r / Wo, x = r - y * Wo;
        const Hmat hm = load_h(H, b);
        float u, v, rw;
        project(hm, static_cast<float>(x), static_cast<float>(y), u, v, rw);
        const Taps t = make_taps(u, v, Ws, Hs);
        const float* sp = src + b * ls.sb;
        float* op = out + b * lo.sb + static_cast<long long>(y) * lo.sy + static_cast<long long>(x) * lo.sx;
        for (int c = 0; c < C; ++c) {
            float nw, ne, sw, se;
            gather4(sp + static_cast<long long>(c) * ls.sc, t, ls.sy, ls.sx, nw, ne, sw, se);
            op[static_cast<long long>(c) * lo.sc] = blend(t, nw, ne, sw, se);
        }
    }
}

__device__ __forceinline__ float4 ld4_or_zero(const float* p, bool ok) {
    return ok ? __ldg(reinterpret_cast<const float4*>(p)) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// channels-last, C % 4 == 0: thread = (output pixel, channel quad); consecutive threads = consecutive quads
__global__ void __launch_bounds__(256)
    warp_fwd_nhwc_kernel(const float* __restrict__ src, const float* __restrict__ H, float* __restrict__ out, int B, int C,
                         int Hs, int Ws, int Ho, int Wo) {
    const int cq = C >> 2;
    const long long n = static_cast<long long>(B) * Ho * Wo * cq;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < n;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int q = static_cast<int>(idx % cq);
        const long long pix = idx / cq;
        const int b = static_cast<int>(pix / (Ho * Wo));
        const int r = static_cast<int>(pix - static_cast<long long>(b) * Ho * Wo);
        const int y = r / Wo, x = r - y * Wo;
        const Hmat hm = load_h(H, b);
        float u, v, rw;
        project(hm, static_cast<float>(x), static_cast<float>(y), u, v, rw);
        const Taps t = make_taps(u, v, Ws, Hs);
        const float* sp = src + static_cast<long long>(b) * Hs * Ws * C + (static_cast<long long>(t.y0) * Ws + t.x0) * C + q * 4;
        const float4 nw = ld4_or_zero(sp, t.inx0 && t.iny0);
        const float4 ne = ld4_or_zero(sp + C, t.inx1 && t.iny0);
        const float4 sw = ld4_or_zero(sp + static_cast<long long>(Ws) * C, t.inx0 && t.iny1);
        const float4 se = ld4_or_zero(sp + static_cast<long long>(Ws) * C + C, t.inx1 && t.iny1);
        float4 o;
        o.x = blend(t, nw.x, ne.x, sw.x, se.x);
        o.y = blend(t, nw.y, ne.y, sw.y, se.y);
        o.z = blend(t, nw.z, ne.z, sw.z, se.z);
        o.w = blend(t, nw.w, ne.w, sw.w, se.w);
        stg_stream(reinterpret_cast<float4*>(out + pix * C + q * 4), o);
    }
}

// Backward, generic layouts.  grid = (chunks, B); every block reduces its pixel chunk to 9 partial sums
// (partials[b][chunk][9]); warp_bwd_finish_kernel adds them in order.  Optional image gradient by red.add.
template <bool kVec4>
__global__ void __launch_bounds__(256)
    warp_bwd_generic_kernel(const float* __restrict__ src, const float* __restrict__ H, const float* __restrict__ gOut,
                            const float* __restrict__ gMaskPooled, float* __restrict__ partials, float* __restrict__ gSrc,
                            int C, int Hs, int Ws, int Ho, int Wo, int pool, Layout ls, Layout lo) {
    __shared__ float red[9 * 8];
    const int b = blockIdx.y;
    const Hmat hm = load_h(H, b);
    float acc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.0f;
    const int npix = Ho * Wo;
    // kVec4 (channels-last): a group of cq threads shares a pixel and splits the channels 4 by 4
    const int cq = kVec4 ? (C >> 2) : 1;
    const long long nwork = static_cast<long long>(npix) * cq;
    for (long long wi = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; wi < nwork;
         wi += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int q = kVec4 ? static_cast<int>(wi % cq) : 0;
        const int r = static_cast<int>(wi / cq);
        const int y = r / Wo, x = r - y * Wo;
        float u, v, rw;
        project(hm, static_cast<float>(x), static_cast<float>(y), u, v, rw);
        const Taps t = make_taps(u, v, Ws, Hs);
        float gu = 0.0f, gv = 0.0f;
        if (gMaskPooled != nullptr && q == 0) {
            float du, dv;
            cover_grad(t, du, dv);
            const float gm = __ldg(gMaskPooled + (static_cast<long long>(b) * (Ho / pool) + y / pool) * (Wo / pool) + x / pool) /
                             static_cast<float>(pool * pool);
            gu = gm * du;
            gv = gm * dv;
        }
        if (gOut != nullptr) {
            const float* sp = src + b * ls.sb;
            const float* gp = gOut + b * lo.sb + static_cast<long long>(y) * lo.sy + static_cast<long long>(x) * lo.sx;
            float* gs = gSrc ? gSrc + b * ls.sb : nullptr;
            if (kVec4) {
                const long long o = (static_cast<long long>(t.y0) * Ws + t.x0) * C + q * 4;
                const float4 g = __ldg(reinterpret_cast<const float4*>(gp + q * 4));
                const float4 nw = ld4_or_zero(sp + o, t.inx0 && t.iny0);
                const float4 ne = ld4_or_zero(sp + o + C, t.inx1 && t.iny0);
                const float4 sw = ld4_or_zero(sp + o + static_cast<long long>(Ws) * C, t.inx0 && t.iny1);
                const float4 se = ld4_or_zero(sp + o + static_cast<long long>(Ws) * C + C, t.inx1 && t.iny1);
                float du, dv;
                blend_grad(t, nw.x, ne.x, sw.x, se.x, du, dv); gu = fmaf(g.x, du, gu); gv = fmaf(g.x, dv, gv);
                blend_grad(t, nw.y, ne.y, sw.y, se.y, du, dv); gu = fmaf(g.y, du, gu); gv = fmaf(g.y, dv, gv);
                blend_grad(t, nw.z, ne.z, sw.z, se.z, du, dv); gu = fmaf(g.z, du, gu); gv = fmaf(g.z, dv, gv);
                blend_grad(t, nw.w, ne.w, sw.w, se.w, du, dv); gu = fmaf(g.w, du, gu); gv = fmaf(g.w, dv, gv);
                if (gs) {
                    const float w00 = t.wx0 * t.wy0, w10 = t.wx1 * t.wy0, w01 = t.wx0 * t.wy1, w11 = t.wx1 * t.wy1;
                    const float gg[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (t.inx0 && t.iny0) atomicAdd(gs + o + k, gg[k] * w00);
                        if (t.inx1 && t.iny0) atomicAdd(gs + o + C + k, gg[k] * w10);
                        if (t.inx0 && t.iny1) atomicAdd(gs + o + static_cast<long long>(Ws) * C + k, gg[k] * w01);
                        if (t.inx1 && t.iny1) atomicAdd(gs + o + static_cast<long long>(Ws) * C + C + k, gg[k] * w11);
                    }
                }
            } else {
                for (int c = 0; c < C; ++c) {
                    const float g = __ldg(gp + static_cast<long long>(c) * lo.sc);
                    float nw, ne, sw, se, du, dv;
                    gather4(sp + static_cast<long long>(c) * ls.sc, t, ls.sy, ls.sx, nw, ne, sw, se);
                    blend_grad(t, nw, ne, sw, se, du, dv);
                    gu = fmaf(g, du, gu);
                    gv = fmaf(g, dv, gv);
                    if (gs) {
                        float* gc = gs + static_cast<long long>(c) * ls.sc + static_cast<long long>(t.y0) * ls.sy +
                                    static_cast<long long>(t.x0) * ls.sx;
                        if (t.inx0 && t.iny0) atomicAdd(gc, g * t.wx0 * t.wy0);
                        if (t.inx1 && t.iny0) atomicAdd(gc + ls.sx, g * t.wx1 * t.wy0);
                        if (t.inx0 && t.iny1) atomicAdd(gc + ls.sy, g * t.wx0 * t.wy1);
                        if (t.inx1 && t.iny1) atomicAdd(gc + ls.sy + ls.sx, g * t.wx1 * t.wy1);
                    }
                }
            }
        }
        accum_gh(acc, gu, gv, u, v, rw, static_cast<float>(x), static_cast<float>(y));
    }
    block_sum<9>(acc, red);
    store9(acc, partials + (static_cast<long long>(b) * gridDim.x + blockIdx.x) * 9);
}

__global__ void warp_bwd_finish_kernel(const float* __restrict__ partials, float* __restrict__ gH, int B, int chunks) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * 9) return;
    const int b = i / 9, k = i - b * 9;
    float s = 0.0f;
    for (int c = 0; c < chunks; ++c) s += partials[(static_cast<long long>(b) * chunks + c) * 9 + k];
    gH[i] = s;
}

// ---- host-side path selection --------------------------------------------------------------------
constexpr int kSmemBudget = 227 * 1024 - 1024;

inline int plane_stages(int Hs, int Ws) {
    const long long bytes = static_cast<long long>(Hs) * Ws * 4;
    if ((bytes & 15) != 0 || bytes >= (1 << 20)) return 0;
    long long s = (kSmemBudget - 64) / bytes;
    return static_cast<int>(s > kMaxStages ? kMaxStages : s);
}
inline bool plane_ok(int Hs, int Ws, int Ho, int Wo, int channels_last) {
    return !channels_last && plane_stages(Hs, Ws) >= 1 && (Ho % 4) == 0 && (Wo % 4) == 0;
}
inline int bwd_chunks(int Ho, int Wo, int C, int vec) {
    const long long work = static_cast<long long>(Ho) * Wo * (vec ? C / 4 : 1);
    long long c = (work + 256 * 16 - 1) / (256 * 16);
    return static_cast<int>(c < 1 ? 1 : (c > 1024 ? 1024 : c));
}
inline int grid_for(long long n, int threads) {
    long long g = (n + threads - 1) / threads;
    const long long cap = static_cast<long long>(kNumSMs) * 16;
    return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace bh

extern "C" int bh_warp_fwd(const float* src, const float* H, float* out, float* mask_pooled, int B, int C, int Hs, int Ws,
                           int Ho, int Wo, int pool, int channels_last, bh_stream_t stream_) {
    using namespace bh;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (!H) return BH_E_NULL;
    if ((src == nullptr) != (out == nullptr)) return BH_E_NULL;
    if (!src && !mask_pooled) return BH_E_NULL;
    if (B <= 0 || Hs <= 0 || Ws <= 0 || Ho <= 0 || Wo <= 0 || (src && C <= 0)) return BH_E_SHAPE;
    if (mask_pooled && (pool <= 0 || Ho % pool || Wo % pool)) return BH_E_SHAPE;
    if (src && (!aligned16(src) || !aligned16(out))) return BH_E_ALIGN;
    int rc = BH_OK;
    bool mask_done = (mask_pooled == nullptr);
    if (src) {
        if (plane_ok(Hs, Ws, Ho, Wo, channels_last)) {
            const int stages = plane_stages(Hs, Ws);
            const int n_planes = B * C;
            const size_t smem = static_cast<size_t>(stages) * Hs * Ws * 4 + 64;
            const int grid = n_planes < kNumSMs ? n_planes : kNumSMs;
            const bool fuse_mask = mask_pooled && pool == 4;
            auto kern = fuse_mask ? warp_fwd_plane_kernel<true> : warp_fwd_plane_kernel<false>;
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
            if (e != cudaSuccess) return static_cast<int>(e);
            kern<<<grid, kPlaneThreads, smem, stream>>>(src, H, out, mask_pooled, n_planes, C, Hs, Ws, Ho, Wo, stages);
            mask_done = mask_done || fuse_mask;
        } else if (channels_last && (C % 4) == 0) {
            const long long n = static_cast<long long>(B) * Ho * Wo * (C / 4);
            warp_fwd_nhwc_kernel<<<grid_for(n, 256), 256, 0, stream>>>(src, H, out, B, C, Hs, Ws, Ho, Wo);
        } else {
            const long long n = static_cast<long long>(B) * Ho * Wo;
            warp_fwd_generic_kernel<<<grid_for(n, 256), 256, 0, stream>>>(src, H, out, B, C, Hs, Ws, Ho, Wo,
                                                                          make_layout(C, Hs, Ws, channels_last),
                                                                          make_layout(C, Ho, Wo, channels_last));
        }
        rc = launch_status();
        if (rc != BH_OK) return rc;
    }
    if (!mask_done) {
        const long long n = static_cast<long long>(B) * (Ho / pool) * (Wo / pool);
        mask_pooled_kernel<<<grid_for(n, 128), 128, 0, stream>>>(H, mask_pooled, B, Hs, Ws, Ho, Wo, pool);
        rc = launch_status();
    }
    return rc;
}

extern "C" size_t bh_warp_bwd_workspace_bytes(int B, int C, int Hs, int Ws, int Ho, int Wo, int channels_last) {
    using namespace bh;
    if (B <= 0 || Ho <= 0 || Wo <= 0) return 0;
    const int vec = channels_last && C > 0 && (C % 4) == 0;
    return static_cast<size_t>(B) * bwd_chunks(Ho, Wo, C > 0 ? C : 1, vec) * 9 * sizeof(float);
}

extern "C" int bh_warp_bwd(const float* src, const float* H, const float* gOut, const float* gMaskPooled, float* gH,
                           float* gSrc, int B, int C, int Hs, int Ws, int Ho, int Wo, int pool, int channels_last,
                           void* workspace, size_t workspace_bytes, bh_stream_t stream_) {
    using namespace bh;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (!H || !gH) return BH_E_NULL;
    if (!gOut && !gMaskPooled) return BH_E_NULL;
    if (gOut && !src) return BH_E_NULL;
    if (gSrc && !gOut) return BH_E_NULL;
    if (B <= 0 || Hs <= 0 || Ws <= 0 || Ho <= 0 || Wo <= 0 || (gOut && C <= 0)) return BH_E_SHAPE;
    if (gMaskPooled && (pool <= 0 || Ho % pool || Wo % pool)) return BH_E_SHAPE;
    if (gOut && (!aligned16(src) || !aligned16(gOut))) return BH_E_ALIGN;
    const bool mask4 = gMaskPooled == nullptr || pool == 4;
    if (!gSrc && mask4 && plane_ok(Hs, Ws, Ho, Wo, gOut ? channels_last : 0)) {
        const int stages = plane_stages(Hs, Ws);
        const size_t smem = gOut ? static_cast<size_t>(stages) * Hs * Ws * 4 + 64 : 0;
        const int grid = B < kNumSMs ? B : kNumSMs;
        void (*kern)(const float*, const float*, const float*, const float*, float*, int, int, int, int, int, int, int);
        if (gOut && gMaskPooled) kern = warp_bwd_plane_kernel<true, true>;
        else if (gOut) kern = warp_bwd_plane_kernel<true, false>;
        else kern = warp_bwd_plane_kernel<false, true>;
        if (smem) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
            if (e != cudaSuccess) return static_cast<int>(e);
        }
        kern<<<grid, kPlaneThreads, smem, stream>>>(src, H, gOut, gMaskPooled, gH, B, C, Hs, Ws, Ho, Wo, stages);
        return launch_status();
    }
    const int vec = gOut && channels_last && (C % 4) == 0;
    const int chunks = bwd_chunks(Ho, Wo, gOut ? C : 1, vec);
    const size_t need = static_cast<size_t>(B) * chunks * 9 * sizeof(float);
    if (!workspace || workspace_bytes < need) return BH_E_WORKSPACE;
    float* partials = static_cast<float*>(workspace);
    const Layout ls = make_layout(C > 0 ? C : 1, Hs, Ws, channels_last), lo = make_layout(C > 0 ? C : 1, Ho, Wo, channels_last);
    dim3 grid(chunks, B);
    if (vec)
        warp_bwd_generic_kernel<true><<<grid, 256, 0, stream>>>(src, H, gOut, gMaskPooled, partials, gSrc, C, Hs, Ws, Ho, Wo,
                                                                 pool, ls, lo);
    else
        warp_bwd_generic_kernel<false><<<grid, 256, 0, stream>>>(src, H, gOut, gMaskPooled, partials, gSrc, C, Hs, Ws, Ho,
                                                                  Wo, pool, ls, lo);
    int rc = launch_status();
    if (rc != BH_OK) return rc;
    warp_bwd_finish_kernel<<<(B * 9 + 127) / 128, 128, 0, stream>>>(partials, gH, B, chunks);
    return launch_status();
}
