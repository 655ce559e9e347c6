// K2 / K2b: fused homography-grid generation + bilinear sampling (forward) and its adjoint w.r.t. H.
//
// Reference semantics: src/data/utils.py:54-59 warp_image(inverse=True) ==
//   out[b,c,y,x] = sum_{4 taps} w_tap * src[b,c,tap],  (u,v) = proj(H_b [x,y,1]^T), zeros padding,
// i.e. F.grid_sample(bilinear, zeros, align_corners=True) on the grid kornia.warp_perspective builds.
// The sampling grid never exists in memory here: coordinates live in registers.
//
// Three code paths, chosen by shape (see bh_warp_fwd / bh_warp_bwd at the bottom):
//   block  : NCHW.  CTA = (plane, 64x64 block of output pixels); the source box that block can touch (bounding box
//            of its four projected corners) is staged row by row with 1-D bulk TMA copies (cp.async.bulk ->
//            UBLKCP) on one mbarrier while the threads classify their tiles; 4 CTAs per SM so the copies of one
//            block overlap the sampling of the others; taps are read from shared memory; each thread owns 4x4
//            output pixels (float4 row stores, the pooled 4x4 coverage mask falls out in-thread); tiles are
//            compacted into interior / border lists so that warps never diverge, and interior tiles skip every
//            bounds test.
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
// block path (NCHW): CTA = (plane, 64x64 block of output pixels); only the source box that block can touch is staged
// =================================================================================================
constexpr int kBlkThreads = 256;
constexpr int kBlk = 64;                      // output block side: 16 x 16 tiles of 4x4 pixels = one tile per thread
constexpr int kBlkSmemBytes = 48 * 1024;      // staging budget per CTA (4 CTAs per SM)

// Per output row: numerators/denominator at x = 0, so that a pixel costs 3 FMA + one reciprocal.
struct RowProj {
    float nx0, ny0, w0;
};
__device__ __forceinline__ RowProj row_proj(const Hmat& m, float y) {
    RowProj r;
    r.nx0 = fmaf(m.h[1], y, m.h[2]);
    r.ny0 = fmaf(m.h[4], y, m.h[5]);
    r.w0 = fmaf(m.h[7], y, m.h[8]);
    return r;
}
// (u, v) = (nx, ny) * rcp(w) with one residual correction per quotient: the textbook division sequence without its
// special-case branches (w > 0 on every valid pixel); error well below 1 ulp of a 128-px coordinate
__device__ __forceinline__ void project_fast(const Hmat& m, const RowProj& rp, float x, float& u, float& v, float& rw) {
    const float w = fmaf(m.h[6], x, rp.w0), nx = fmaf(m.h[0], x, rp.nx0), ny = fmaf(m.h[3], x, rp.ny0);
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(w));
    float q = nx * r;
    u = fmaf(fmaf(-q, w, nx), r, q);
    q = ny * r;
    v = fmaf(fmaf(-q, w, ny), r, q);
    rw = r;
}
__device__ __forceinline__ Taps make_taps_fast(float u, float v, int Ws, int Hs) {
    Taps t;
    t.x0 = __float2int_rd(u);  // floor with saturation: far-away coordinates stay out of range
    t.y0 = __float2int_rd(v);
    const float fx = __int2float_rn(t.x0), fy = __int2float_rn(t.y0);
    t.wx0 = (fx + 1.0f) - u;
    t.wx1 = u - fx;
    t.wy0 = (fy + 1.0f) - v;
    t.wy1 = v - fy;
    t.inx0 = static_cast<unsigned>(t.x0) < static_cast<unsigned>(Ws);
    t.inx1 = static_cast<unsigned>(t.x0) + 1u < static_cast<unsigned>(Ws);
    t.iny0 = static_cast<unsigned>(t.y0) < static_cast<unsigned>(Hs);
    t.iny1 = static_cast<unsigned>(t.y0) + 1u < static_cast<unsigned>(Hs);
    return t;
}

// Source box [c0, c1] x [r0, r1] that the output block [x_lo, x_hi) x [y_lo, y_hi) can sample (bilinear footprint
// included, columns aligned to 4 pixels for the 16-byte bulk copies).  A projective map with w > 0 sends the block to a
// convex quadrilateral, so the extremes sit at its four corners.  Returns false when the box is not bounded that way.
struct Box {
    int c0, c1, r0, r1;
};
__device__ __forceinline__ bool block_source_box(const Hmat& m, int x_lo, int x_hi, int y_lo, int y_hi, int Ws, int Hs, Box& bx) {
    float umin = 3.0e38f, umax = -3.0e38f, vmin = 3.0e38f, vmax = -3.0e38f;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float x = static_cast<float>((k & 1) ? x_hi - 1 : x_lo), y = static_cast<float>((k & 2) ? y_hi - 1 : y_lo);
        float u, v, rw;
        project(m, x, y, u, v, rw);
        ok = ok && (rw > 0.0f) && (fabsf(u) < 1.0e6f) && (fabsf(v) < 1.0e6f);
        umin = fminf(umin, u); umax = fmaxf(umax, u);
        vmin = fminf(vmin, v); vmax = fmaxf(vmax, v);
    }
    if (!ok) return false;
    bx.c0 = max(0, (static_cast<int>(floorf(umin)) - 1) & ~3);
    bx.c1 = min(Ws - 1, ((static_cast<int>(floorf(umax)) + 2) | 3));
    bx.r0 = max(0, static_cast<int>(floorf(vmin)) - 1);
    bx.r1 = min(Hs - 1, static_cast<int>(floorf(vmax)) + 2);
    return true;
}

// Where does the 4x4 output tile at (xt, yt) sample?  The tile maps to a convex quadrilateral (w > 0), so its four
// corners decide: kInside = every bilinear footprint lies strictly inside the source (no bounds tests, coverage 1),
// kOutside = no footprint touches the source (zeros, coverage 0), kBorder = anything else (predicated taps).
enum TileClass { kInside = 0, kBorder = 1, kOutside = 2 };
__device__ __forceinline__ int classify_tile(const Hmat& m, int xt, int yt, int Ws, int Hs) {
    bool in = true, pos = true;
    float umin = 3.0e38f, umax = -3.0e38f, vmin = 3.0e38f, vmax = -3.0e38f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float x = static_cast<float>(xt + ((k & 1) ? 3 : 0)), y = static_cast<float>(yt + ((k & 2) ? 3 : 0));
        const float w = fmaf(m.h[6], x, fmaf(m.h[7], y, m.h[8]));
        float r;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(w));
        const float u = fmaf(m.h[0], x, fmaf(m.h[1], y, m.h[2])) * r, v = fmaf(m.h[3], x, fmaf(m.h[4], y, m.h[5])) * r;
        pos = pos && (w > 0.0f);
        in = in && (u >= 0.001f) && (v >= 0.001f) && (u < static_cast<float>(Ws - 1) - 0.001f) && (v < static_cast<float>(Hs - 1) - 0.001f);
        umin = fminf(umin, u); umax = fmaxf(umax, u);
        vmin = fminf(vmin, v); vmax = fmaxf(vmax, v);
    }
    if (!pos) return kBorder;  // also catches NaN
    if (in) return kInside;
    // 0.001 px guard bands: the corner coordinates above carry ~2 ulp of error
    if (umax < -1.001f || vmax < -1.001f || umin > static_cast<float>(Ws) + 0.001f || vmin > static_cast<float>(Hs) + 0.001f) return kOutside;
    return kBorder;
}

// Staged source box: taps[y * pitch + x] addresses source pixel (x, y) for every pixel of the box (shared memory), or of
// the whole plane when the box does not fit / is not 16-byte friendly (global memory through the read-only cache).
struct Stage {
    const float* taps;
    int pitch;
    bool shared, pending;
};
// issue: one bulk TMA copy per source row of the box (cp.async.bulk -> UBLKCP), spread over the lanes of warp 0, all
// completing on one mbarrier.  Only warp 0 computes the box and issues; it publishes {r0, c0, ncols, shared} in `pub`
// and every thread picks the result up with stage_get() after the block's next __syncthreads.  Nothing waits here:
// the CTA classifies its tiles while the copies are in flight.
struct StagePub {
    int r0, c0, ncols, shared;
};
__device__ __forceinline__ void stage_issue_warp0(const Hmat& hm, const float* plane_ptr, float* smem, uint64_t* bar, int Ws, int Hs,
                                                  int x_lo, int x_hi, int y_lo, int y_hi, bool aligned, StagePub* pub) {
    Box bx;
    const bool have_box = block_source_box(hm, x_lo, x_hi, y_lo, y_hi, Ws, Hs, bx);
    const int ncols = bx.c1 - bx.c0 + 1, nrows = bx.r1 - bx.r0 + 1;
    const bool ok = have_box && aligned && ncols > 0 && nrows > 0 && ncols * nrows * 4 <= kBlkSmemBytes;
    if (ok) {
        if (threadIdx.x == 0) mbar_expect_tx(bar, static_cast<uint32_t>(ncols) * nrows * 4u);
        __syncwarp();
        for (int r = threadIdx.x; r < nrows; r += 32)
            bulk_g2s(smem + r * ncols, plane_ptr + static_cast<size_t>(bx.r0 + r) * Ws + bx.c0, static_cast<uint32_t>(ncols) * 4u, bar);
    }
    if (threadIdx.x == 0) {
        pub->r0 = bx.r0; pub->c0 = bx.c0; pub->ncols = ncols; pub->shared = ok ? 1 : 0;
    }
}
__device__ __forceinline__ Stage stage_get(const StagePub* pub, const float* plane_ptr, const float* smem, int Ws) {
    Stage st;
    st.shared = pub->shared != 0;
    st.pending = st.shared;
    if (st.shared) {
        st.pitch = pub->ncols;
        st.taps = smem - static_cast<ptrdiff_t>(pub->r0) * pub->ncols - pub->c0;
    } else {  // box too large / unaligned / block samples nothing: whole plane through the read-only cache
        st.pitch = Ws;
        st.taps = plane_ptr;
    }
    return st;
}
__device__ __forceinline__ void stage_wait(Stage& st, uint64_t* bar, uint32_t parity) {
    if (st.pending) mbar_wait(bar, parity);
    st.pending = false;
}

// Divergence-free scheduling of the block's tiles: every thread classifies one tile, interior and border tiles are
// compacted into two shared-memory lists, and the lists are then processed with all lanes of a warp on the same code
// path (a warp's 32 tiles otherwise nearly always straddle the warped image border somewhere).
struct TileLists {
    int n_in, n_bd;
    unsigned short in[kBlkThreads], bd[kBlkThreads];
};
__device__ __forceinline__ void lists_reset(TileLists& L) {
    __syncthreads();  // previous use fully consumed
    if (threadIdx.x == 0) { L.n_in = 0; L.n_bd = 0; }
    __syncthreads();
}
__device__ __forceinline__ void lists_push(TileLists& L, int cls, int tile) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned m_in = __ballot_sync(0xffffffffu, cls == kInside), m_bd = __ballot_sync(0xffffffffu, cls == kBorder);
    int base_in = 0, base_bd = 0;
    if (lane == 0) {
        if (m_in) base_in = atomicAdd(&L.n_in, __popc(m_in));
        if (m_bd) base_bd = atomicAdd(&L.n_bd, __popc(m_bd));
    }
    base_in = __shfl_sync(0xffffffffu, base_in, 0);
    base_bd = __shfl_sync(0xffffffffu, base_bd, 0);
    const unsigned below = (1u << lane) - 1u;
    if (cls == kInside) L.in[base_in + __popc(m_in & below)] = static_cast<unsigned short>(tile);
    if (cls == kBorder) L.bd[base_bd + __popc(m_bd & below)] = static_cast<unsigned short>(tile);
    __syncthreads();
}

template <bool kShared>
__device__ __forceinline__ float ld_tap(const float* p) {
    return kShared ? *p : __ldg(p);
}

// Interior pixel: coordinates, cell and the four taps with ~25 instructions.
//   cell by FRND.FLOOR, tap index y0 * pitch + x0 formed in float (exact below 2^24) and converted once.
struct Cell {
    float u, v, rw, fx, fy, nw, ne, sw, se;
};
template <bool kShared>
__device__ __forceinline__ Cell interior_cell(const Hmat& m, const RowProj& rp, float x, const float* taps, uint32_t taps_s32,
                                              int pitch, float pitchf) {
    Cell c;
    project_fast(m, rp, x, c.u, c.v, c.rw);
    const float x0f = floorf(c.u), y0f = floorf(c.v);
    c.fx = c.u - x0f;
    c.fy = c.v - y0f;
    const int idx = __float2int_rn(fmaf(y0f, pitchf, x0f));
    if (kShared) {
        // explicit shared-window addresses: LDS with immediate offsets instead of generic loads
        const uint32_t a0 = taps_s32 + static_cast<uint32_t>(idx) * 4u, a1 = a0 + static_cast<uint32_t>(pitch) * 4u;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(c.nw) : "r"(a0));
        asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(c.ne) : "r"(a0));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(c.sw) : "r"(a1));
        asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(c.se) : "r"(a1));
    } else {
        const float* p = taps + idx;
        c.nw = __ldg(p);
        c.ne = __ldg(p + 1);
        c.sw = __ldg(p + pitch);
        c.se = __ldg(p + pitch + 1);
    }
    return c;
}
// border pixel: predicated taps
template <bool kShared>
__device__ __forceinline__ void border_taps(const Taps& t, const float* taps, int pitch, float& nw, float& ne, float& sw, float& se) {
    const float* p = taps + t.y0 * pitch + t.x0;
    nw = (t.inx0 && t.iny0) ? ld_tap<kShared>(p) : 0.0f;
    ne = (t.inx1 && t.iny0) ? ld_tap<kShared>(p + 1) : 0.0f;
    sw = (t.inx0 && t.iny1) ? ld_tap<kShared>(p + pitch) : 0.0f;
    se = (t.inx1 && t.iny1) ? ld_tap<kShared>(p + pitch + 1) : 0.0f;
}

// geometry of the block a CTA owns
struct BlockGeom {
    int x_lo, y_lo, tiles_x, tiles_y;  // tiles of 4x4 pixels inside the block (edge blocks are partial)
};
__device__ __forceinline__ BlockGeom block_geom(int blk, int blocks_x, int Ho, int Wo) {
    BlockGeom g;
    const int by = blk / blocks_x, bxi = blk - by * blocks_x;
    g.x_lo = bxi * kBlk;
    g.y_lo = by * kBlk;
    g.tiles_x = (min(Wo, g.x_lo + kBlk) - g.x_lo) >> 2;
    g.tiles_y = (min(Ho, g.y_lo + kBlk) - g.y_lo) >> 2;
    return g;
}

template <bool kMask, bool kShared>
__device__ __forceinline__ void fwd_block_lists(const Hmat& hm, const Stage& st, float* __restrict__ op, float* __restrict__ mp,
                                                const BlockGeom& g, int Hs, int Ws, int Wo, const TileLists& L) {
    const float pitchf = static_cast<float>(st.pitch);
    // taps may point below the shared window (the box origin sits at its start): 32-bit arithmetic wraps back into it
    const uint32_t taps_s32 = kShared ? smem_u32(st.taps) : 0u;
    const int mask_w = Wo >> 2;
    for (int k = threadIdx.x; k < L.n_in; k += kBlkThreads) {
        const int tile = L.in[k];
        const int xt = g.x_lo + (tile & 15) * 4, yt = g.y_lo + (tile >> 4) * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const RowProj rp = row_proj(hm, static_cast<float>(yt + j));
            float o4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const Cell c = interior_cell<kShared>(hm, rp, static_cast<float>(xt + i), st.taps, taps_s32, st.pitch, pitchf);
                const float top = fmaf(c.fx, c.ne - c.nw, c.nw), bot = fmaf(c.fx, c.se - c.sw, c.sw);
                o4[i] = fmaf(c.fy, bot - top, top);
            }
            stg_stream(reinterpret_cast<float4*>(op + (yt + j) * Wo + xt), make_float4(o4[0], o4[1], o4[2], o4[3]));
        }
        if (kMask && mp != nullptr) mp[(yt >> 2) * mask_w + (xt >> 2)] = 1.0f;
    }
    for (int k = threadIdx.x; k < L.n_bd; k += kBlkThreads) {
        const int tile = L.bd[k];
        const int xt = g.x_lo + (tile & 15) * 4, yt = g.y_lo + (tile >> 4) * 4;
        float msum = 0.0f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const RowProj rp = row_proj(hm, static_cast<float>(yt + j));
            float o4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float u, v, rw, nw, ne, sw, se;
                project_fast(hm, rp, static_cast<float>(xt + i), u, v, rw);
                const Taps t = make_taps_fast(u, v, Ws, Hs);
                border_taps<kShared>(t, st.taps, st.pitch, nw, ne, sw, se);
                o4[i] = blend(t, nw, ne, sw, se);
                if (kMask) msum += cover(t);
            }
            stg_stream(reinterpret_cast<float4*>(op + (yt + j) * Wo + xt), make_float4(o4[0], o4[1], o4[2], o4[3]));
        }
        if (kMask && mp != nullptr) mp[(yt >> 2) * mask_w + (xt >> 2)] = msum * 0.0625f;
    }
}

// kMask: also emit the 4x4-pooled coverage mask (written by the CTAs of channel 0)
template <bool kMask>
__global__ void __launch_bounds__(kBlkThreads)
    warp_fwd_block_kernel(const float* __restrict__ src, const float* __restrict__ H, float* __restrict__ out,
                          float* __restrict__ mask_pooled, int C, int Hs, int Ws, int Ho, int Wo, int blocks_x, int n_blocks) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ TileLists lists;
    __shared__ StagePub pub;
    const int blk = blockIdx.x, b = blockIdx.z, plane = b * C + blockIdx.y;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
        lists.n_in = 0;
        lists.n_bd = 0;
    }
    __syncthreads();
    const Hmat hm = load_h(H, b);
    const BlockGeom g = block_geom(blk, blocks_x, Ho, Wo);
    const float* plane_ptr = src + static_cast<size_t>(plane) * Hs * Ws;
    if (threadIdx.x < 32)
        stage_issue_warp0(hm, plane_ptr, reinterpret_cast<float*>(smem_raw), &bar, Ws, Hs, g.x_lo, g.x_lo + g.tiles_x * 4, g.y_lo,
                          g.y_lo + g.tiles_y * 4, (Ws & 3) == 0, &pub);
    float* op = out + static_cast<size_t>(plane) * Ho * Wo;
    float* mp = (kMask && blockIdx.y == 0) ? mask_pooled + static_cast<size_t>(b) * (Ho >> 2) * (Wo >> 2) : nullptr;
    // classification (and the zero fill of tiles that sample nothing) overlaps the copies in flight
    {
        const int tile = threadIdx.x, tx = tile & 15, ty = tile >> 4;
        int cls = kOutside;
        if (tx < g.tiles_x && ty < g.tiles_y) {
            const int xt = g.x_lo + tx * 4, yt = g.y_lo + ty * 4;
            cls = classify_tile(hm, xt, yt, Ws, Hs);
            if (cls == kOutside) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    stg_stream(reinterpret_cast<float4*>(op + (yt + j) * Wo + xt), make_float4(0.0f, 0.0f, 0.0f, 0.0f));
                if (kMask && mp != nullptr) mp[(yt >> 2) * (Wo >> 2) + (xt >> 2)] = 0.0f;
            }
        }
        lists_push(lists, cls, tile);  // ends with __syncthreads: `pub` is visible
    }
    Stage st = stage_get(&pub, plane_ptr, reinterpret_cast<const float*>(smem_raw), Ws);
    stage_wait(st, &bar, 0u);
    if (st.shared) fwd_block_lists<kMask, true>(hm, st, op, mp, g, Hs, Ws, Wo, lists);
    else fwd_block_lists<kMask, false>(hm, st, op, mp, g, Hs, Ws, Wo, lists);
}

template <bool kImage, bool kMask, bool kShared>
__device__ __forceinline__ void bwd_block_lists(const Hmat& hm, const Stage& st, const float* __restrict__ gp,
                                                const float* __restrict__ gmp, float (&acc)[9], const BlockGeom& g, int Hs, int Ws,
                                                int Wo, const TileLists& L) {
    const float pitchf = static_cast<float>(st.pitch);
    const uint32_t taps_s32 = kShared ? smem_u32(st.taps) : 0u;
    const int mask_w = Wo >> 2;
    if (kImage) {
        for (int k = threadIdx.x; k < L.n_in; k += kBlkThreads) {
            const int tile = L.in[k];
            const int xt = g.x_lo + (tile & 15) * 4, yt = g.y_lo + (tile >> 4) * 4;
            float4 g4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) g4[j] = ldg_stream(reinterpret_cast<const float4*>(gp + (yt + j) * Wo + xt));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float y = static_cast<float>(yt + j);
                const RowProj rp = row_proj(hm, y);
                const float gj[4] = {g4[j].x, g4[j].y, g4[j].z, g4[j].w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float x = static_cast<float>(xt + i);
                    const Cell c = interior_cell<kShared>(hm, rp, x, st.taps, taps_s32, st.pitch, pitchf);
                    const float dt = c.ne - c.nw, db = c.se - c.sw, dl = c.sw - c.nw, dr = c.se - c.ne;
                    const float du = fmaf(c.fy, db - dt, dt), dv = fmaf(c.fx, dr - dl, dl);
                    accum_gh(acc, gj[i] * du, gj[i] * dv, c.u, c.v, c.rw, x, y);
                }
            }
        }
    }
    for (int k = threadIdx.x; k < L.n_bd; k += kBlkThreads) {
        const int tile = L.bd[k];
        const int xt = g.x_lo + (tile & 15) * 4, yt = g.y_lo + (tile >> 4) * 4;
        float4 g4[4];
        if (kImage) {
#pragma unroll
            for (int j = 0; j < 4; ++j) g4[j] = ldg_stream(reinterpret_cast<const float4*>(gp + (yt + j) * Wo + xt));
        }
        const float gm = (kMask && gmp != nullptr) ? __ldg(gmp + (yt >> 2) * mask_w + (xt >> 2)) * 0.0625f : 0.0f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float y = static_cast<float>(yt + j);
            const RowProj rp = row_proj(hm, y);
            const float gj[4] = {kImage ? g4[j].x : 0.0f, kImage ? g4[j].y : 0.0f, kImage ? g4[j].z : 0.0f, kImage ? g4[j].w : 0.0f};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float x = static_cast<float>(xt + i);
                float u, v, rw;
                project_fast(hm, rp, x, u, v, rw);
                const Taps t = make_taps_fast(u, v, Ws, Hs);
                float gu = 0.0f, gv = 0.0f;
                if (kImage) {
                    float nw, ne, sw, se, du, dv;
                    border_taps<kShared>(t, st.taps, st.pitch, nw, ne, sw, se);
                    blend_grad(t, nw, ne, sw, se, du, dv);
                    gu = gj[i] * du;
                    gv = gj[i] * dv;
                }
                if (kMask) {
                    float du, dv;
                    cover_grad(t, du, dv);
                    gu = fmaf(gm, du, gu);
                    gv = fmaf(gm, dv, gv);
                }
                accum_gh(acc, gu, gv, u, v, rw, x, y);
            }
        }
    }
}

// Backward: CTA = (sample, 64x64 output block), all C planes in turn; its 9 partial sums go to partials[b][block][9]
// and warp_bwd_finish_kernel adds the blocks in order (bit reproducible, no atomics).
// kMask: the pooled-mask upstream (pool == 4) is folded into the same pass (channel 0).
template <bool kImage, bool kMask>
__global__ void __launch_bounds__(kBlkThreads)
    warp_bwd_block_kernel(const float* __restrict__ src, const float* __restrict__ H, const float* __restrict__ gOut,
                          const float* __restrict__ gMaskPooled, float* __restrict__ partials, int C, int Hs, int Ws, int Ho,
                          int Wo, int blocks_x, int n_blocks) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ float red[9 * (kBlkThreads / 32)];
    __shared__ TileLists lists;
    __shared__ StagePub pub;
    const int blk = blockIdx.x, b = blockIdx.y;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
        lists.n_in = 0;
        lists.n_bd = 0;
    }
    __syncthreads();
    const Hmat hm = load_h(H, b);
    const BlockGeom g = block_geom(blk, blocks_x, Ho, Wo);
    const float* plane0 = kImage ? src + static_cast<size_t>(b) * C * Hs * Ws : nullptr;
    if (kImage && threadIdx.x < 32)
        stage_issue_warp0(hm, plane0, reinterpret_cast<float*>(smem_raw), &bar, Ws, Hs, g.x_lo, g.x_lo + g.tiles_x * 4, g.y_lo,
                          g.y_lo + g.tiles_y * 4, (Ws & 3) == 0, &pub);
    {
        const int tile = threadIdx.x, tx = tile & 15, ty = tile >> 4;
        int cls = kOutside;  // zero taps, zero coverage: no gradient
        if (tx < g.tiles_x && ty < g.tiles_y) {
            cls = classify_tile(hm, g.x_lo + tx * 4, g.y_lo + ty * 4, Ws, Hs);
            if (cls == kInside && !kImage) cls = kOutside;  // coverage is constant 1 inside
        }
        lists_push(lists, cls, tile);  // ends with __syncthreads: `pub` is visible
    }
    float acc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.0f;
    const float* gmp = kMask ? gMaskPooled + static_cast<size_t>(b) * (Ho >> 2) * (Wo >> 2) : nullptr;
    if (kImage) {
        for (int c = 0; c < C; ++c) {
            const size_t plane = static_cast<size_t>(b) * C + c;
            const float* plane_ptr = src + plane * Hs * Ws;
            if (c > 0) {
                __syncthreads();  // the staging buffer is reused
                if (threadIdx.x < 32)
                    stage_issue_warp0(hm, plane_ptr, reinterpret_cast<float*>(smem_raw), &bar, Ws, Hs, g.x_lo, g.x_lo + g.tiles_x * 4,
                                      g.y_lo, g.y_lo + g.tiles_y * 4, (Ws & 3) == 0, &pub);
                __syncthreads();
            }
            Stage st = stage_get(&pub, plane_ptr, reinterpret_cast<const float*>(smem_raw), Ws);
            stage_wait(st, &bar, static_cast<uint32_t>(c & 1));
            const float* gp = gOut + plane * Ho * Wo;
            const float* gm_c = (c == 0) ? gmp : nullptr;
            if (st.shared) bwd_block_lists<true, kMask, true>(hm, st, gp, gm_c, acc, g, Hs, Ws, Wo, lists);
            else bwd_block_lists<true, kMask, false>(hm, st, gp, gm_c, acc, g, Hs, Ws, Wo, lists);
        }
    } else {
        Stage none;
        none.taps = nullptr; none.pitch = Ws; none.shared = false; none.pending = false;
        bwd_block_lists<false, true, false>(hm, none, nullptr, gmp, acc, g, Hs, Ws, Wo, lists);
    }
    block_sum<9>(acc, red);
    store9(acc, partials + (static_cast<size_t>(b) * n_blocks + blk) * 9);
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
// block path applies to NCHW tensors whose output is made of whole 4x4 tiles; source boxes that do not fit the staging
// budget (or unaligned row pitches) are read through the read-only cache instead -- same kernel
inline bool block_ok(int Ho, int Wo, int channels_last) { return !channels_last && (Ho % 4) == 0 && (Wo % 4) == 0; }
constexpr int kMaxGridYZ = 65535;
inline int blocks_of(int n) { return (n + kBlk - 1) / kBlk; }
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
        if (block_ok(Ho, Wo, channels_last) && C <= kMaxGridYZ && B <= kMaxGridYZ) {
            const int blocks_x = blocks_of(Wo), n_blocks = blocks_x * blocks_of(Ho);
            const bool fuse_mask = mask_pooled && pool == 4;
            auto kern = fuse_mask ? warp_fwd_block_kernel<true> : warp_fwd_block_kernel<false>;
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kBlkSmemBytes);
            if (e != cudaSuccess) return static_cast<int>(e);
            kern<<<dim3(n_blocks, C, B), kBlkThreads, kBlkSmemBytes, stream>>>(src, H, out, mask_pooled, C, Hs, Ws, Ho, Wo,
                                                                               blocks_x, n_blocks);
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
    const size_t generic = static_cast<size_t>(B) * bwd_chunks(Ho, Wo, C > 0 ? C : 1, vec) * 9 * sizeof(float);
    const size_t blocks = static_cast<size_t>(B) * blocks_of(Wo) * blocks_of(Ho) * 9 * sizeof(float);
    return generic > blocks ? generic : blocks;
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
    if (!gSrc && mask4 && B <= kMaxGridYZ && block_ok(Ho, Wo, gOut ? channels_last : 0)) {
        const int blocks_x = blocks_of(Wo), n_blocks = blocks_x * blocks_of(Ho);
        const size_t need = static_cast<size_t>(B) * n_blocks * 9 * sizeof(float);
        if (!workspace || workspace_bytes < need) return BH_E_WORKSPACE;
        float* partials = static_cast<float*>(workspace);
        const size_t smem = gOut ? kBlkSmemBytes : 0;
        void (*kern)(const float*, const float*, const float*, const float*, float*, int, int, int, int, int, int, int);
        if (gOut && gMaskPooled) kern = warp_bwd_block_kernel<true, true>;
        else if (gOut) kern = warp_bwd_block_kernel<true, false>;
        else kern = warp_bwd_block_kernel<false, true>;
        if (smem) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
            if (e != cudaSuccess) return static_cast<int>(e);
        }
        kern<<<dim3(n_blocks, B), kBlkThreads, smem, stream>>>(src, H, gOut, gMaskPooled, partials, C, Hs, Ws, Ho, Wo, blocks_x, n_blocks);
        int rc = launch_status();
        if (rc != BH_OK) return rc;
        warp_bwd_finish_kernel<<<(B * 9 + 127) / 128, 128, 0, stream>>>(partials, gH, B, n_blocks);
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
