// K2 / K2b: fused homography-grid generation + bilinear sampling (forward) and its adjoint w.r.t. H.
//
// Reference semantics: src/data/utils.py:54-59 warp_image(inverse=True) ==
//   out[b,c,y,x] = sum_{4 taps} w_tap * src[b,c,tap],  (u,v) = proj(H_b [x,y,1]^T), zeros padding,
// i.e. F.grid_sample(bilinear, zeros, align_corners=True) on the grid kornia.warp_perspective builds.
// The sampling grid never exists in memory here: coordinates live in registers.
//
// Three code paths, chosen by shape (see bh_warp_fwd / bh_warp_bwd at the bottom):
//   ring   : NCHW (the biHomE path: 1-channel patches).  Persistent warp-specialised kernels: a producer warp stages
//            the source box of every 64x64 output block with bulk TMA row copies into a zero-framed shared-memory
//            window while eight consumer warps sample the previous block; see the section comment below.
//   nhwc   : channels-last, C % 4 == 0: one thread per (pixel, 4 channels), 128-bit coalesced tap loads.
//   generic: anything else (scalar, strided).
// Backward produces dH by a fixed-order per-sample reduction (bit-reproducible, no atomics); the image
// gradient (dead work on the biHomE path: the source never requires grad) is an optional red.global.add.
#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)

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
// ring path (NCHW, the default): persistent, warp-specialised, TMA-pipelined.
//
//   work item  = a whole source plane with its 128x128 (or smaller) output (kBlk = 128), or one 64x64 block of output
//                pixels (kBlk = 64); one (two) persistent CTA(s) per SM walk the items round-robin; the last round is
//                split into half-items when that evens out the tail (ring_split).
//   last warp  = PRODUCER.  For the items after the one being sampled it loads H, computes the source box the block can
//                touch (kBlk = 64: bounding box of its four projected corners; kBlk = 128: the plane), writes the ZERO
//                FRAME of the staging window, publishes the geometry as an item header in shared memory and issues the
//                bulk TMA copies (cp.async.bulk -> UBLKCP: 4 x 16 KB for a plane, one per source row for a box) into
//                the free stage, completing on full[stage].
//   the others = CONSUMERS.  Wait full[stage], read the header, classify their 32x4 groups (inside / border / outside,
//                convexity argument on the projected group corners), sample, arrive on empty[stage].  The lanes of a
//                warp sit on 32 CONSECUTIVE output columns and walk down 16 rows in groups of 4 (8 on the interior
//                fast path): with a window pitch that is a multiple of 32 floats the four tap loads of a warp are
//                conflict-free LDS, output stores are 128-byte rows.
//   zero frame = 1 row above, 2 below, 4 columns left and right of the staged box: grid_sample's "zeros" padding is a
//                clamp of (u, v) to [-1, W] x [-1, H] followed by the same four unpredicated LDS an interior pixel does.
//   floor      = add.rm.f32 with 1.5 * 2^23: cell index and fraction come out of the FMA pipe (no FRND / F2I).
//   arithmetic = packed fp32x2 (FFMA2 / FADD2 / FMUL2): rows (j, j + 1) of a lane's column share an instruction.
//   backward   = the upstream gradients of the NEXT strip travel by 4-byte cp.async into per-warp landing buffers while
//                the current strip is sampled (strip_prefetch); six packed partial sums per lane, fixed-order finish.
// The copies run two items (backward: one item) ahead of the consumers, so no warp waits on the DRAM latency of its own
// window.  Blocks whose box exceeds the stage (local scale > ~1.4, strong down-sampling) read the plane through the
// read-only cache with predicated taps inside the same kernel.
// =================================================================================================
constexpr int kStripRows = 16;    // rows walked by one consumer warp per strip (32 columns x 16 rows)
constexpr int kGBufBytes = (kStripRows + 4) * 32 * 4;   // one strip of upstream gradients + its 4 pooled-mask cells, per lane
constexpr int kPadL = 4, kPadT = 1, kPadB = 2;   // the 4 floats in front of a window row are also the right pad of the row above
// Two item sizes.  kBlk = 128: the source plane (<= 128x128) always fits one stage, so a whole plane is staged once and
// its 32 strips are sampled by 16 consumer warps, 1 CTA per SM, 3 stages (216 KB): the copies run two planes ahead.
// kBlk = 64: larger sources, 8 consumer warps, 3 stages, 2 CTAs per SM.
// (Measured dead ends, kept out of the code: folding the producer into the last consumer warp to finish -- 16 warps, 128
// registers -- is 7 % slower; fetching a whole strip of upstream gradients one strip ahead IN REGISTERS spills at the
// 96 registers a 17-warp CTA leaves per thread and is 40 % slower -- hence the cp.async landing buffers; issuing the
// copies before the header's H load (two arrivals on full[]) changes nothing measurable.)
template <int kBlk>
struct RingCfg {
    static constexpr int kStages = 3;
    static constexpr int kStageBytes = kBlk == 128 ? 72 * 1024 : 36 * 1024;
    static constexpr int kConsumers = kBlk == 128 ? 512 : 256;
    static constexpr int kThreads = kConsumers + 32;
    static constexpr int kCtasPerSm = kBlk == 128 ? 1 : 2;
    static constexpr int kStripsX = kBlk / 32, kStripsY = kBlk / kStripRows, kStrips = kStripsX * kStripsY;
    static constexpr int kStripsPerWarp = kStrips / (kConsumers / 32);
    static constexpr int kSmem = kStages * kStageBytes;
    // backward with an upstream image gradient: two window stages + a double-buffered per-warp landing zone for the
    // gradient stream (see strip_prefetch)
    static constexpr int kBwdStages = 2;
    static constexpr int kBwdSmem = kBwdStages * kStageBytes + (kConsumers / 32) * 2 * kGBufBytes;
};
constexpr float kMagic = 12582912.0f;            // 1.5 * 2^23: ulp == 1, add.rm.f32 leaves floor(t) in the mantissa
constexpr int kMagicBits = 0x4B400000;

enum GroupClass { kInside = 0, kBorder = 1, kOutside = 2, kSkip = 3 };

struct ItemHeader {
    float h[9];
    int c0, r0, ncols, nrows, pitch, ok;
    int plane, b, c, blk, x_lo, y_lo, x_hi, y_hi;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// wait with a hardware suspend hint (consumers) / with a sleep between polls (producer): spinning warps must not eat the
// issue slots of the warps that sample
__device__ __forceinline__ void mbar_wait_hint(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAITH_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONEH_%=;\n"
        "nanosleep.u32 64;\n"
        "bra WAITH_%=;\n"
        "DONEH_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(1000000u)
        : "memory");
}
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAITS_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONES_%=;\n"
        "nanosleep.u32 200;\n"
        "bra WAITS_%=;\n"
        "DONES_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(1000000u)
        : "memory");
}
__device__ __forceinline__ void prefetch_l2_bulk(const void* gsrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ float add_rm(float a, float b) {
    float r;
    asm("add.rm.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void st_stream1(float* p, float v) {
    asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ float ld_stream1(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FMUL2 / FADD2) ----------------------------------------------------------
// The ring kernels are bound by instruction issue, not by a pipe: the two rows a lane handles at a time share one
// instruction.  A value of type f2 is a 64-bit register pair (lo, hi) = (row j, row j + 1); pk() of two equal scalars
// costs nothing (the SASS forms take a broadcast 32-bit operand, negation folded in).
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float lo, float hi) {
    f2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk(f2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void upk_bits(f2 v, uint32_t& lo, uint32_t& hi) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
    f2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
    f2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
    f2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b) {
    f2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f2 add2_rm(f2 a, f2 b) {
    f2 d;
    asm("add.rm.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f2 dup2(float v) { return pk(v, v); }

// geometry of the block an item covers
struct BlockGeom {
    int x_lo, y_lo, x_hi, y_hi;  // [x_lo, x_hi) x [y_lo, y_hi), multiples of 4
};
__device__ __forceinline__ BlockGeom block_geom(int blk, int blocks_x, int Ho, int Wo, int side) {
    BlockGeom g;
    const int by = blk / blocks_x, bxi = blk - by * blocks_x;
    g.x_lo = bxi * side;
    g.y_lo = by * side;
    g.x_hi = min(Wo, g.x_lo + side);
    g.y_hi = min(Ho, g.y_lo + side);
    return g;
}

// (u, v) = proj(H [x, y, 1]) by reciprocal only: set-up quantities (boxes, classes) that carry their own guard bands
__device__ __forceinline__ void project_rcp(const Hmat& m, float x, float y, float& u, float& v, float& w) {
    w = fmaf(m.h[6], x, fmaf(m.h[7], y, m.h[8]));
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(w));
    u = fmaf(m.h[0], x, fmaf(m.h[1], y, m.h[2])) * r;
    v = fmaf(m.h[3], x, fmaf(m.h[4], y, m.h[5])) * r;
}
// min / max / and over the aligned group of 4 lanes
__device__ __forceinline__ float quad_min(float v) {
    v = fminf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fminf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ bool quad_all(bool p) {
    const unsigned m = __ballot_sync(0xffffffffu, p);
    return ((m >> ((threadIdx.x & 31u) & ~3u)) & 0xfu) == 0xfu;
}

// Per-column projection state of a consumer lane: numerators / denominator at y = 0.
struct ColProj {
    float ax, ay, aw;
};
__device__ __forceinline__ ColProj col_proj(const Hmat& m, float x) {
    ColProj c;
    c.ax = fmaf(m.h[0], x, m.h[2]);
    c.ay = fmaf(m.h[3], x, m.h[5]);
    c.aw = fmaf(m.h[6], x, m.h[8]);
    return c;
}
// (u, v) = (nx, ny) * rcp(w) with one residual correction per quotient: the textbook division sequence without its
// special-case branches (w > 0 on every valid pixel); error well below 1 ulp of a 128-px coordinate
__device__ __forceinline__ void project_col(const Hmat& m, const ColProj& cp, float y, float& u, float& v, float& rw) {
    const float w = fmaf(m.h[7], y, cp.aw), nx = fmaf(m.h[1], y, cp.ax), ny = fmaf(m.h[4], y, cp.ay);
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(w));
    float q = nx * r;
    u = fmaf(fmaf(-q, w, nx), r, q);
    q = ny * r;
    v = fmaf(fmaf(-q, w, ny), r, q);
    rw = r;
}
// the same for rows (y, y + 1) of a column, bit for bit: -w is carried instead of w so that every step is a plain FFMA2
__device__ __forceinline__ void project_col2(const Hmat& m, const ColProj& cp, f2 y2, f2& u2, f2& v2, f2& r2) {
    const f2 mw = fma2(dup2(-m.h[7]), y2, dup2(-cp.aw));
    const f2 nx = fma2(dup2(m.h[1]), y2, dup2(cp.ax)), ny = fma2(dup2(m.h[4]), y2, dup2(cp.ay));
    float m0, m1, r0, r1;
    upk(mw, m0, m1);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(-m0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(-m1));
    r2 = pk(r0, r1);
    f2 q = mul2(nx, r2);
    u2 = fma2(fma2(q, mw, nx), r2, q);
    q = mul2(ny, r2);
    v2 = fma2(fma2(q, mw, ny), r2, q);
}

// Where the taps of an item live: the framed shared window, or the whole plane in global memory.
struct Window {
    const float* taps;     // global: plane base.  shared: unused
    uint32_t base_s32;     // shared: byte address of source pixel (0, 0) minus the magic-number bias (may wrap)
    int pitch;             // floats
    bool shared;
    bool xpred;            // shared window without side pads: border pixels predicate their x taps
};
// bilinear cell of (u, v), |u|, |v| < 2^22: fractions and the four taps, no bounds tests (interior pixels, or any pixel of
// a framed window after the clamp)
struct Cell4 {
    float fx, fy, nw, ne, sw, se;
    int ix;   // floor(u)
};
template <bool kShared>
__device__ __forceinline__ Cell4 cell_at(float u, float v, const Window& wd) {
    Cell4 c;
    const float tx = add_rm(u, kMagic), ty = add_rm(v, kMagic);
    c.fx = u - (tx - kMagic);
    c.fy = v - (ty - kMagic);
    c.ix = __float_as_int(tx) - kMagicBits;
    // biased by kMagicBits * (pitch + 1); unsigned arithmetic: the bias wraps and is taken out again below / in base_s32
    const uint32_t lin = static_cast<uint32_t>(__float_as_int(ty)) * static_cast<uint32_t>(wd.pitch) + static_cast<uint32_t>(__float_as_int(tx));
    if (kShared) {
        const uint32_t a0 = wd.base_s32 + lin * 4u, a1 = a0 + static_cast<uint32_t>(wd.pitch) * 4u;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(c.nw) : "r"(a0));
        asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(c.ne) : "r"(a0));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(c.sw) : "r"(a1));
        asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(c.se) : "r"(a1));
    } else {
        const float* p = wd.taps + static_cast<int>(lin - static_cast<uint32_t>(kMagicBits) * static_cast<uint32_t>(wd.pitch + 1));
        c.nw = __ldg(p);
        c.ne = __ldg(p + 1);
        c.sw = __ldg(p + wd.pitch);
        c.se = __ldg(p + wd.pitch + 1);
    }
    return c;
}
// dense window (no side pads): taps left / right of the source read the neighbouring row -- replace them by zeros
__device__ __forceinline__ void zero_x_taps(Cell4& c, int Ws) {
    const bool l = static_cast<unsigned>(c.ix) < static_cast<unsigned>(Ws), r = static_cast<unsigned>(c.ix + 1) < static_cast<unsigned>(Ws);
    c.nw = l ? c.nw : 0.0f;
    c.sw = l ? c.sw : 0.0f;
    c.ne = r ? c.ne : 0.0f;
    c.se = r ? c.se : 0.0f;
}
// Two cells at once (rows j, j + 1 of a lane) out of a shared window; kPitch > 0: compile-time window pitch.
// xpred: dense window, see zero_x_taps.
struct Cell8 {
    f2 fx, fy, nw, ne, sw, se;
};
template <int kPitch>
__device__ __forceinline__ Cell8 cell_at2(f2 u2, f2 v2, const Window& wd, bool xpred, int Ws) {
    Cell8 c;
    const f2 tx = add2_rm(u2, dup2(kMagic)), ty = add2_rm(v2, dup2(kMagic));
    c.fx = sub2(u2, add2(tx, dup2(-kMagic)));
    c.fy = sub2(v2, add2(ty, dup2(-kMagic)));
    uint32_t tx0, tx1, ty0, ty1;
    upk_bits(tx, tx0, tx1);
    upk_bits(ty, ty0, ty1);
    const uint32_t pitch = kPitch > 0 ? static_cast<uint32_t>(kPitch) : static_cast<uint32_t>(wd.pitch);
    const uint32_t a0 = wd.base_s32 + (ty0 * pitch + tx0) * 4u, a1 = wd.base_s32 + (ty1 * pitch + tx1) * 4u;
    float n0, e0, s0, t0, n1, e1, s1, t1;   // nw, ne, sw, se of row j / row j + 1
    if (kPitch > 0) {
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(n0) : "r"(a0));
        asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(e0) : "r"(a0));
        asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(s0) : "r"(a0), "n"(kPitch * 4));
        asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(t0) : "r"(a0), "n"(kPitch * 4 + 4));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(n1) : "r"(a1));
        asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(e1) : "r"(a1));
        asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(s1) : "r"(a1), "n"(kPitch * 4));
        asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(t1) : "r"(a1), "n"(kPitch * 4 + 4));
    } else {
        const uint32_t b0 = a0 + pitch * 4u, b1 = a1 + pitch * 4u;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(n0) : "r"(a0));
        asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(e0) : "r"(a0));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(s0) : "r"(b0));
        asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(t0) : "r"(b0));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(n1) : "r"(a1));
        asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(e1) : "r"(a1));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(s1) : "r"(b1));
        asm volatile("ld.shared.f32 %0, [%1+4];" : "=f"(t1) : "r"(b1));
    }
    if (xpred) {
        const int i0 = static_cast<int>(tx0) - kMagicBits, i1 = static_cast<int>(tx1) - kMagicBits;
        const bool l0 = static_cast<unsigned>(i0) < static_cast<unsigned>(Ws), r0 = static_cast<unsigned>(i0 + 1) < static_cast<unsigned>(Ws);
        const bool l1 = static_cast<unsigned>(i1) < static_cast<unsigned>(Ws), r1 = static_cast<unsigned>(i1 + 1) < static_cast<unsigned>(Ws);
        n0 = l0 ? n0 : 0.0f; s0 = l0 ? s0 : 0.0f; e0 = r0 ? e0 : 0.0f; t0 = r0 ? t0 : 0.0f;
        n1 = l1 ? n1 : 0.0f; s1 = l1 ? s1 : 0.0f; e1 = r1 ? e1 : 0.0f; t1 = r1 ? t1 : 0.0f;
    }
    c.nw = pk(n0, n1); c.ne = pk(e0, e1); c.sw = pk(s0, s1); c.se = pk(t0, t1);
    return c;
}
// bilinear blend of two cells: the scalar sequence of fwd_group, two rows per instruction
__device__ __forceinline__ f2 lerp2(const Cell8& c) {
    const f2 top = fma2(c.fx, sub2(c.ne, c.nw), c.nw), bot = fma2(c.fx, sub2(c.se, c.sw), c.sw);
    return fma2(c.fy, sub2(bot, top), top);
}
// predicated taps for the global-memory fallback (no frame there)
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
__device__ __forceinline__ void global_taps(const Taps& t, const float* plane, int pitch, float& nw, float& ne, float& sw, float& se) {
    const float* p = plane + t.y0 * pitch + t.x0;
    nw = (t.inx0 && t.iny0) ? __ldg(p) : 0.0f;
    ne = (t.inx1 && t.iny0) ? __ldg(p + 1) : 0.0f;
    sw = (t.inx0 && t.iny1) ? __ldg(p + pitch) : 0.0f;
    se = (t.inx1 && t.iny1) ? __ldg(p + pitch + 1) : 0.0f;
}
// coverage of warp(ones) along one axis, and its derivative: m(t) = clamp(min(t + 1, n - t), 0, 1)
__device__ __forceinline__ float cover1(float t, float n) { return __saturatef(fminf(t + 1.0f, n - t)); }
__device__ __forceinline__ float cover1_grad(float t, float n) {
    const float a = t + 1.0f, b = n - t, m = fminf(a, b);
    return (m > 0.0f && m < 1.0f) ? (a < b ? 1.0f : -1.0f) : 0.0f;
}

// ---- producer ------------------------------------------------------------------------------------------------
// Window layout (floats): row pitch = ncols + 4 (rounded up to a multiple of 32 when the stage has room: conflict-free
// LDS); source pixel (c, r) of the box sits at [(r - r0 + 1) * pitch + 4 + (c - c0)].  Row 0, rows nrows + 1 and
// nrows + 2 and the floats between the data of consecutive rows are the zero frame.
// kBlk == 128 stages the whole source plane (constant geometry: the frame is written once per stage);
// kBlk == 64 stages the bounding box of the block's four projected corners.
// kStage = false: coverage-only items (mask backward), nothing to copy.
template <int kBlk, bool kStage>
__device__ __forceinline__ void ring_produce(const float* __restrict__ src, const float* __restrict__ H, int it, int n_blocks,
                                             int blocks_x, int C, ItemHeader* hd, float* stage, uint64_t* full, int Hs, int Ws,
                                             int Ho, int Wo, bool frame_written, const float* __restrict__ l2_prefetch) {
    using Cfg = RingCfg<kBlk>;
    const int lane = threadIdx.x & 31;
    const int plane = it / n_blocks, blk = it - plane * n_blocks;
    const int b = plane / C, c = plane - b * C;
    const BlockGeom g = block_geom(blk, blocks_x, Ho, Wo, kBlk);
    int c0 = 0, r0 = 0, ncols = 0, nrows = 0, pitch = 0;
    bool ok = false;
    if (kStage && kBlk == 128) {
        // dense plane: pitch == Ws, no side pads (the border path predicates its x taps instead)
        ncols = Ws; nrows = Hs; pitch = Ws;
        ok = (Ws & 3) == 0;   // the host picked this item size because the plane fits the stage
    } else if (kStage) {
        // source box of the block: lanes 4q..4q+3 each project one corner
        const Hmat hm = load_h(H, b);
        const int k = lane & 3;
        float u, v, w;
        project_rcp(hm, static_cast<float>((k & 1) ? g.x_hi - 1 : g.x_lo), static_cast<float>((k & 2) ? g.y_hi - 1 : g.y_lo), u, v, w);
        const bool good = quad_all(w > 0.0f && fabsf(u) < 1.0e6f && fabsf(v) < 1.0e6f);
        const float umin = quad_min(u), umax = quad_max(u), vmin = quad_min(v), vmax = quad_max(v);
        c0 = max(0, (static_cast<int>(floorf(umin)) - 1) & ~3);
        const int c1 = min(Ws - 1, ((static_cast<int>(floorf(umax)) + 2) | 3));
        r0 = max(0, static_cast<int>(floorf(vmin)) - 1);
        const int r1 = min(Hs - 1, static_cast<int>(floorf(vmax)) + 2);
        ncols = c1 - c0 + 1;
        nrows = r1 - r0 + 1;
        const int rows_all = nrows + kPadT + kPadB;
        pitch = (ncols + kPadL + 31) & ~31;
        if (rows_all * pitch * 4 > Cfg::kStageBytes) pitch = ncols + kPadL;   // large box: accept a few bank conflicts
        ok = good && (Ws & 3) == 0 && ncols > 0 && nrows > 0 && rows_all * pitch * 4 <= Cfg::kStageBytes;
    }
    if (lane < 9) hd->h[lane] = __ldg(H + 9 * b + lane);
    if (lane == 0) {
        hd->c0 = c0; hd->r0 = r0; hd->ncols = ncols; hd->nrows = nrows; hd->pitch = pitch; hd->ok = ok ? 1 : 0;
        hd->plane = plane; hd->b = b; hd->c = c; hd->blk = blk;
        hd->x_lo = g.x_lo; hd->y_lo = g.y_lo; hd->x_hi = g.x_hi; hd->y_hi = g.y_hi;
    }
    if (ok && kBlk == 128) {
        if (!frame_written) {
            // zero rows above / below the dense plane (+ 4 floats of slack in front), written once per stage
            const float4 z = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            float4* w4 = reinterpret_cast<float4*>(stage);
            const int p4 = pitch >> 2;
            for (int i = lane; i <= p4; i += 32) w4[i] = z;                                        // slack + row 0
            for (int i = lane; i < 2 * p4 + 1; i += 32) w4[1 + (nrows + 1) * p4 + i] = z;          // rows nrows+1, nrows+2 + slack
        }
    } else if (ok) {
        // zero frame, 16 bytes at a time (addresses disjoint from what the bulk copies write)
        const float4 z = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        float4* w4 = reinterpret_cast<float4*>(stage);
        const int p4 = pitch >> 2;
        for (int i = lane; i < p4; i += 32) {
            w4[i] = z;
            w4[(nrows + 1) * p4 + i] = z;
            w4[(nrows + 2) * p4 + i] = z;
        }
        // in front of every data row (= behind the previous one), and everything right of the data up to the pitch
        const int tail4 = (pitch - kPadL - ncols) >> 2;
        for (int r = lane; r < nrows; r += 32) {
            w4[(r + kPadT) * p4] = z;
            for (int t = 0; t < tail4; ++t) w4[(r + kPadT) * p4 + 1 + (ncols >> 2) + t] = z;
        }
    }
    __syncwarp();  // header and frame written by all lanes before lane 0 releases them
    if (ok) {
        const uint32_t total = static_cast<uint32_t>(ncols) * nrows * 4u;
        if (lane == 0) mbar_expect_tx(full, total);
        __syncwarp();
        const float* plane_ptr = src + static_cast<size_t>(plane) * Hs * Ws;
        if (kBlk == 128) {
            // the TMA unit spends ~50 cycles per request whatever its size: the contiguous plane goes in <= 4 requests
            const uint32_t chunk = ((total / 4u) + 15u) & ~15u;
            const uint32_t off = static_cast<uint32_t>(lane) * chunk;
            if (lane < 4 && off < total)
                bulk_g2s(reinterpret_cast<char*>(stage + kPadT * pitch + kPadL) + off, reinterpret_cast<const char*>(plane_ptr) + off,
                         min(chunk, total - off), full);
        } else {
            for (int r = lane; r < nrows; r += 32)
                bulk_g2s(stage + (r + kPadT) * pitch + kPadL, plane_ptr + static_cast<size_t>(r0 + r) * Ws + c0,
                         static_cast<uint32_t>(ncols) * 4u, full);
        }
    } else if (lane == 0) {
        mbar_arrive(full);
    }
    // backward: pull the plane of the upstream gradient towards L2 while the consumers still sample the previous item
    if (kBlk == 128 && l2_prefetch != nullptr && lane == 0 && n_blocks == 1 && ((Ho * Wo) & 3) == 0)
        prefetch_l2_bulk(l2_prefetch + static_cast<size_t>(plane) * Ho * Wo, static_cast<uint32_t>(Ho * Wo) * 4u);
}

// ---- consumers -----------------------------------------------------------------------------------------------
struct ItemView {
    Hmat hm;
    Window wd;
    int plane, b, c, blk;
    int x, y0;          // this lane's output column, first row of its warp's strip
    bool xin;
    unsigned cls4;
};
// Classes of the 32x4 groups this warp owns in the item (kStripsPerWarp strips x 4 groups), 4 bits each: every lane
// projects ONE corner of one group, quads combine them (convexity: a group maps to a convex quadrilateral when w > 0).
// kInsideIsSkip: groups that sample strictly inside carry no work (mask backward: coverage is constant 1 there).
template <int kBlk, bool kInsideIsSkip>
__device__ __forceinline__ unsigned ring_classify(const ItemHeader* hd, const Hmat& hm, int warp, int Hs, int Ws) {
    using Cfg = RingCfg<kBlk>;
    const int lane = threadIdx.x & 31;
    const int gidx = lane >> 2, k = lane & 3;                 // group 0..7 = (strip of this warp) * 4 + group in strip
    const int q = gidx >> 2, gq = gidx & 3;
    const int sidx = warp + (q < Cfg::kStripsPerWarp ? q : 0) * (Cfg::kConsumers / 32);
    const int x_hi = hd->x_hi, y_hi = hd->y_hi;
    const int xa = hd->x_lo + (sidx % Cfg::kStripsX) * 32, xb = min(xa + 31, x_hi - 1);
    const int ya = hd->y_lo + (sidx / Cfg::kStripsX) * kStripRows + gq * 4, yb = ya + 3;
    float u, v, w;
    project_rcp(hm, static_cast<float>((k & 1) ? xb : xa), static_cast<float>((k & 2) ? yb : ya), u, v, w);
    const bool pos = quad_all(w > 0.0f);
    // 0.001 px guard bands: the corner coordinates carry a few ulp of error
    const bool in = quad_all((u >= 0.001f) && (v >= 0.001f) && (u < static_cast<float>(Ws - 1) - 0.001f) &&
                             (v < static_cast<float>(Hs - 1) - 0.001f));
    const float umin = quad_min(u), umax = quad_max(u), vmin = quad_min(v), vmax = quad_max(v);
    int cls;
    if (!(xa < x_hi && ya < y_hi) || q >= Cfg::kStripsPerWarp) cls = kSkip;
    else if (!pos) cls = kBorder;  // also catches NaN
    else if (in) cls = kInsideIsSkip ? kSkip : kInside;
    else if (umax < -1.001f || vmax < -1.001f || umin > static_cast<float>(Ws) + 0.001f || vmin > static_cast<float>(Hs) + 0.001f)
        cls = kInsideIsSkip ? kSkip : kOutside;
    else cls = kBorder;
    return __reduce_or_sync(0xffffffffu, k == 0 ? static_cast<unsigned>(cls) << (4 * gidx) : 0u);
}
// the part of the view that changes from strip to strip of one item (q = which of this warp's strips)
template <int kBlk>
__device__ __forceinline__ void ring_view_strip(ItemView& v, const ItemHeader* hd, int warp, int q, unsigned classes, int Wo) {
    const int sidx = warp + q * (RingCfg<kBlk>::kConsumers / 32);
    v.x = hd->x_lo + (sidx % RingCfg<kBlk>::kStripsX) * 32 + static_cast<int>(threadIdx.x & 31u);
    v.y0 = hd->y_lo + (sidx / RingCfg<kBlk>::kStripsX) * kStripRows;
    v.xin = v.x < Wo;
    v.cls4 = (classes >> (16 * q)) & 0xffffu;
}
__device__ __forceinline__ ItemView ring_view(const ItemHeader* hd, const float* __restrict__ src, const float* stage, int Hs, int Ws) {
    ItemView v;
#pragma unroll
    for (int i = 0; i < 9; ++i) v.hm.h[i] = hd->h[i];
    v.plane = hd->plane; v.b = hd->b; v.c = hd->c; v.blk = hd->blk;
    v.x = 0; v.y0 = 0; v.xin = false; v.cls4 = 0u;
    v.wd.shared = hd->ok != 0;
    v.wd.xpred = hd->pitch == hd->ncols;
    if (v.wd.shared) {
        v.wd.pitch = hd->pitch;
        // byte address of source pixel (0, 0) inside the window, minus the bias the magic-number floor leaves in `lin`
        const uint32_t origin = static_cast<uint32_t>((kPadT - hd->r0) * v.wd.pitch + (kPadL - hd->c0)) -
                                static_cast<uint32_t>(kMagicBits) * static_cast<uint32_t>(v.wd.pitch + 1);
        v.wd.base_s32 = smem_u32(stage) + origin * 4u;
        v.wd.taps = nullptr;
    } else {
        v.wd.pitch = Ws;
        v.wd.base_s32 = 0u;
        v.wd.taps = src + static_cast<size_t>(v.plane) * Hs * Ws;
    }
    return v;
}

// one 32x4 group of the forward pass.  kWo > 0: compile-time output pitch (immediate store offsets); kPitch > 0:
// compile-time window pitch.  Shared windows go through the packed two-rows-per-instruction sequence.
template <bool kMask, bool kShared, int kWo, int kPitch>
__device__ __forceinline__ void fwd_group(const ItemView& iv, const ColProj& cp, int cls, float yg, float* __restrict__ og,
                                          float* __restrict__ mcell, int Hs, int Ws, int Wo_rt) {
    const int Wo = kWo > 0 ? kWo : Wo_rt;
    const float Wsf = static_cast<float>(Ws), Hsf = static_cast<float>(Hs);
    if (cls == kInside) {
        if (iv.xin) {
            float o4[4];
            if (kShared) {
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    f2 u2, v2, r2;
                    project_col2(iv.hm, cp, add2(dup2(yg), pk(2.0f * p, 2.0f * p + 1.0f)), u2, v2, r2);
                    upk(lerp2(cell_at2<kPitch>(u2, v2, iv.wd, false, Ws)), o4[2 * p], o4[2 * p + 1]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float u, v, rw;
                    project_col(iv.hm, cp, yg + static_cast<float>(j), u, v, rw);
                    const Cell4 c = cell_at<false>(u, v, iv.wd);
                    const float top = fmaf(c.fx, c.ne - c.nw, c.nw), bot = fmaf(c.fx, c.se - c.sw, c.sw);
                    o4[j] = fmaf(c.fy, bot - top, top);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) st_stream1(og + j * Wo, o4[j]);
            if (kMask && mcell != nullptr) *mcell = 1.0f;
        }
    } else if (cls == kBorder) {
        float msum = 0.0f;
        if (iv.xin) {  // lanes right of the output would sample outside the staged box
            float o4[4];
            if (kShared) {
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    f2 u2, v2, r2;
                    project_col2(iv.hm, cp, add2(dup2(yg), pk(2.0f * p, 2.0f * p + 1.0f)), u2, v2, r2);
                    // zero frame: clamp, then the unpredicated interior sequence (fmaxf / fminf also absorb NaN)
                    float u0, u1, v0, v1;
                    upk(u2, u0, u1);
                    upk(v2, v0, v1);
                    u0 = fminf(fmaxf(u0, -1.0f), Wsf); u1 = fminf(fmaxf(u1, -1.0f), Wsf);
                    v0 = fminf(fmaxf(v0, -1.0f), Hsf); v1 = fminf(fmaxf(v1, -1.0f), Hsf);
                    upk(lerp2(cell_at2<kPitch>(pk(u0, u1), pk(v0, v1), iv.wd, iv.wd.xpred, Ws)), o4[2 * p], o4[2 * p + 1]);
                    if (kMask) {
                        msum = fmaf(cover1(u0, Wsf), cover1(v0, Hsf), msum);
                        msum = fmaf(cover1(u1, Wsf), cover1(v1, Hsf), msum);
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float u, v, rw, nw, ne, sw, se;
                    project_col(iv.hm, cp, yg + static_cast<float>(j), u, v, rw);
                    const Taps t = make_taps_fast(u, v, Ws, Hs);
                    global_taps(t, iv.wd.taps, iv.wd.pitch, nw, ne, sw, se);
                    o4[j] = blend(t, nw, ne, sw, se);
                    if (kMask) msum += cover(t);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) st_stream1(og + j * Wo, o4[j]);
        }
        if (kMask) {
            msum += __shfl_xor_sync(0xffffffffu, msum, 1);
            msum += __shfl_xor_sync(0xffffffffu, msum, 2);
            if (mcell != nullptr && iv.xin) *mcell = msum * 0.0625f;
        }
    } else if (cls == kOutside) {
        if (iv.xin) {
#pragma unroll
            for (int j = 0; j < 4; ++j) st_stream1(og + j * Wo, 0.0f);
            if (kMask && mcell != nullptr) *mcell = 0.0f;
        }
    }
}

// two vertically adjacent interior groups of a shared window at once (32 x 8 pixels): four independent row pairs in
// flight per lane -- the consumer warps are few (the staging windows take the shared memory), so the latency of the
// reciprocal -> floor -> LDS -> blend chain has to be covered by instruction-level parallelism
template <int kWo, int kPitch>
__device__ __forceinline__ void fwd_inside8(const ItemView& iv, const ColProj& cp, float yg, float* __restrict__ og, int Ws, int Wo_rt) {
    const int Wo = kWo > 0 ? kWo : Wo_rt;
    f2 u2[4], v2[4], o2[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        f2 r2;
        project_col2(iv.hm, cp, add2(dup2(yg), pk(2.0f * p, 2.0f * p + 1.0f)), u2[p], v2[p], r2);
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) o2[p] = lerp2(cell_at2<kPitch>(u2[p], v2[p], iv.wd, false, Ws));
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        float a, b;
        upk(o2[p], a, b);
        st_stream1(og + (2 * p) * Wo, a);
        st_stream1(og + (2 * p + 1) * Wo, b);
    }
}

template <bool kMask, bool kShared, int kWo, int kPitch>
__device__ __forceinline__ void fwd_item(const ItemView& iv, float* __restrict__ out, float* __restrict__ mask_pooled, int Hs, int Ws,
                                         int Ho, int Wo_rt) {
    const int Wo = kWo > 0 ? kWo : Wo_rt;
    const ColProj cp = col_proj(iv.hm, static_cast<float>(iv.x));
    float* og = out + static_cast<size_t>(iv.plane) * Ho * Wo + iv.y0 * Wo + iv.x;
    // pooled-mask cell of group 0 (written by the lanes with (lane & 3) == 0 of channel 0)
    float* mcell = nullptr;
    if (kMask && iv.c == 0 && (threadIdx.x & 3) == 0)
        mcell = mask_pooled + static_cast<size_t>(iv.b) * (Ho >> 2) * (Wo >> 2) + (iv.y0 >> 2) * (Wo >> 2) + (iv.x >> 2);
    const float y0f = static_cast<float>(iv.y0);
#pragma unroll 1
    for (int gh = 0; gh < 2; ++gh) {
        const int cls_a = (iv.cls4 >> (8 * gh)) & 0xf, cls_b = (iv.cls4 >> (8 * gh + 4)) & 0xf;
        const float yg = y0f + static_cast<float>(8 * gh);
        float* ogg = og + gh * 8 * Wo;
        float* mc = mcell ? mcell + gh * 2 * (Wo >> 2) : nullptr;
        if (kShared && cls_a == kInside && cls_b == kInside) {
            if (iv.xin) {
                fwd_inside8<kWo, kPitch>(iv, cp, yg, ogg, Ws, Wo_rt);
                if (kMask && mc != nullptr) { mc[0] = 1.0f; mc[Wo >> 2] = 1.0f; }
            }
        } else {
            fwd_group<kMask, kShared, kWo, kPitch>(iv, cp, cls_a, yg, ogg, mc, Hs, Ws, Wo_rt);
            fwd_group<kMask, kShared, kWo, kPitch>(iv, cp, cls_b, yg + 4.0f, ogg + 4 * Wo, mc ? mc + (Wo >> 2) : nullptr, Hs, Ws, Wo_rt);
        }
    }
}

// kMask: also emit the 4x4-pooled coverage mask (written by the items of channel 0).
// kWo = 128: the north-star shape, Wo == 128 and (whole-plane items) Ws == 128 as well -- output and window pitches are
// immediates; kWo = 0: run-time sizes.
template <int kBlk, bool kMask, int kWo>
__global__ void __launch_bounds__(RingCfg<kBlk>::kThreads, RingCfg<kBlk>::kCtasPerSm)
    warp_fwd_ring_kernel(const float* __restrict__ src, const float* __restrict__ H, float* __restrict__ out,
                         float* __restrict__ mask_pooled, int C, int Hs, int Ws, int Ho, int Wo, int blocks_x, int n_blocks,
                         int n_items, int n_full) {
    using Cfg = RingCfg<kBlk>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full[Cfg::kStages], empty[Cfg::kStages];
    __shared__ ItemHeader header[Cfg::kStages];
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < Cfg::kStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], Cfg::kConsumers / 32);
        }
        fence_mbar_init();
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    const bool producer = warp == Cfg::kConsumers / 32;
    int k = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++k) {
        const int s = k % Cfg::kStages;
        const uint32_t round = static_cast<uint32_t>(k / Cfg::kStages);
        float* stage = reinterpret_cast<float*>(smem_raw + s * Cfg::kStageBytes);
        // launch items [0, n_full) are whole items; the rest are HALVES of the remaining items (see ring_split)
        const bool half = it >= n_full;
        const int item = half ? n_full + ((it - n_full) >> 1) : it;
        const int q_lo = half ? ((it - n_full) & 1) * (Cfg::kStripsPerWarp / 2) : 0;
        const int q_hi = half ? q_lo + Cfg::kStripsPerWarp / 2 : Cfg::kStripsPerWarp;
        if (producer) {
            if (k >= Cfg::kStages) mbar_wait_sleep(&empty[s], (round - 1u) & 1u);  // consumers released the stage
            ring_produce<kBlk, true>(src, H, item, n_blocks, blocks_x, C, &header[s], stage, &full[s], Hs, Ws, Ho, Wo, k >= Cfg::kStages,
                                     nullptr);
        } else {
            mbar_wait_hint(&full[s], round & 1u);
            ItemView iv = ring_view(&header[s], src, stage, Hs, Ws);
            const unsigned classes = ring_classify<kBlk, false>(&header[s], iv.hm, warp, Hs, Ws);
#pragma unroll 1
            for (int q = q_lo; q < q_hi; ++q) {
                ring_view_strip<kBlk>(iv, &header[s], warp, q, classes, Wo);
                constexpr int kPitch = (kBlk == 128 && kWo == 128) ? 128 : 0;
                if (iv.wd.shared) fwd_item<kMask, true, kWo, kPitch>(iv, out, mask_pooled, Hs, Ws, Ho, Wo);
                else fwd_item<kMask, false, kWo, 0>(iv, out, mask_pooled, Hs, Ws, Ho, Wo);
            }
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[s]);
        }
    }
}

// ---- backward ------------------------------------------------------------------------------------------------
// dH sums of one strip.  A lane keeps its column fixed, so the x factor of the nine sums is applied once at the end:
// per pixel only  sum a, sum a*y, sum b, sum b*y, sum c, sum c*y  are advanced -- packed, rows (j, j + 1) in the two
// halves of a register pair (added at the end of the strip); c is accumulated with its sign flipped.
struct StripSums2 {
    f2 sa, say, sb, sby, sc, scy;
};
__device__ __forceinline__ void sums_add2(StripSums2& t, f2 gu, f2 gv, f2 u, f2 v, f2 rw, f2 y) {
    const f2 a = mul2(gu, rw), b = mul2(gv, rw), c = mul2(fma2(gu, u, mul2(gv, v)), rw);
    t.sa = add2(t.sa, a); t.say = fma2(a, y, t.say);
    t.sb = add2(t.sb, b); t.sby = fma2(b, y, t.sby);
    t.sc = add2(t.sc, c); t.scy = fma2(c, y, t.scy);
}
// d out / du, d out / dv of two cells for unit upstream
__device__ __forceinline__ void cell_grad2(const Cell8& c, f2& du, f2& dv) {
    const f2 dt = sub2(c.ne, c.nw), db = sub2(c.se, c.sw), dl = sub2(c.sw, c.nw), dr = sub2(c.se, c.ne);
    du = fma2(c.fy, sub2(db, dt), dt);
    dv = fma2(c.fx, sub2(dr, dl), dl);
}

template <bool kImage, bool kMask, bool kShared, int kPitch>
__device__ __forceinline__ void bwd_group(const ItemView& iv, const ColProj& cp, int cls, float yg, const float (&g4)[4], float gm,
                                          StripSums2& t, int Hs, int Ws) {
    const float Wsf = static_cast<float>(Ws), Hsf = static_cast<float>(Hs);
    if (!iv.xin) return;
    if (cls == kInside) {
        if (kImage) {
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const f2 y2 = add2(dup2(yg), pk(2.0f * p, 2.0f * p + 1.0f)), g2 = pk(g4[2 * p], g4[2 * p + 1]);
                f2 u2, v2, r2, du, dv;
                project_col2(iv.hm, cp, y2, u2, v2, r2);
                if (kShared) {
                    cell_grad2(cell_at2<kPitch>(u2, v2, iv.wd, false, Ws), du, dv);
                } else {
                    float u0, u1, v0, v1;
                    upk(u2, u0, u1);
                    upk(v2, v0, v1);
                    const Cell4 c0 = cell_at<false>(u0, v0, iv.wd), c1 = cell_at<false>(u1, v1, iv.wd);
                    Cell8 c;
                    c.fx = pk(c0.fx, c1.fx); c.fy = pk(c0.fy, c1.fy);
                    c.nw = pk(c0.nw, c1.nw); c.ne = pk(c0.ne, c1.ne); c.sw = pk(c0.sw, c1.sw); c.se = pk(c0.se, c1.se);
                    cell_grad2(c, du, dv);
                }
                sums_add2(t, mul2(g2, du), mul2(g2, dv), u2, v2, r2, y2);
            }
        }
    } else if (cls == kBorder) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const f2 y2 = add2(dup2(yg), pk(2.0f * p, 2.0f * p + 1.0f)), g2 = pk(g4[2 * p], g4[2 * p + 1]);
            f2 u2, v2, r2;
            project_col2(iv.hm, cp, y2, u2, v2, r2);
            float u0, u1, v0, v1, r0, r1;
            upk(u2, u0, u1);
            upk(v2, v0, v1);
            upk(r2, r0, r1);
            f2 gu = dup2(0.0f), gv = dup2(0.0f);
            if (kShared) {
                // zero frame: clamp (fmaxf / fminf also absorb NaN), sample like an interior pixel
                const float uc0 = fminf(fmaxf(u0, -1.0f), Wsf), uc1 = fminf(fmaxf(u1, -1.0f), Wsf);
                const float vc0 = fminf(fmaxf(v0, -1.0f), Hsf), vc1 = fminf(fmaxf(v1, -1.0f), Hsf);
                const f2 uc = pk(uc0, uc1), vc = pk(vc0, vc1);
                if (kImage) {
                    f2 du, dv;
                    cell_grad2(cell_at2<kPitch>(uc, vc, iv.wd, iv.wd.xpred, Ws), du, dv);
                    gu = mul2(g2, du);
                    gv = mul2(g2, dv);
                }
                if (kMask) {
                    gu = fma2(dup2(gm), pk(cover1(vc0, Hsf) * cover1_grad(uc0, Wsf), cover1(vc1, Hsf) * cover1_grad(uc1, Wsf)), gu);
                    gv = fma2(dup2(gm), pk(cover1(uc0, Wsf) * cover1_grad(vc0, Hsf), cover1(uc1, Wsf) * cover1_grad(vc1, Hsf)), gv);
                }
                // a clamped coordinate does not move with H (also true for NaN: the comparison fails): no contribution
                const bool k0 = (uc0 == u0) && (vc0 == v0), k1 = (uc1 == u1) && (vc1 == v1);
                const f2 keep = pk(k0 ? 1.0f : 0.0f, k1 ? 1.0f : 0.0f);
                sums_add2(t, mul2(gu, keep), mul2(gv, keep), uc, vc, pk(k0 ? r0 : 0.0f, k1 ? r1 : 0.0f), y2);
            } else {
                float gus[2] = {0.0f, 0.0f}, gvs[2] = {0.0f, 0.0f};
                const float us[2] = {u0, u1}, vs[2] = {v0, v1};
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const Taps tp = make_taps_fast(us[e], vs[e], Ws, Hs);
                    if (kImage) {
                        float nw, ne, sw, se, du, dv;
                        global_taps(tp, iv.wd.taps, iv.wd.pitch, nw, ne, sw, se);
                        blend_grad(tp, nw, ne, sw, se, du, dv);
                        gus[e] = g4[2 * p + e] * du;
                        gvs[e] = g4[2 * p + e] * dv;
                    }
                    if (kMask) {
                        float du, dv;
                        cover_grad(tp, du, dv);
                        gus[e] = fmaf(gm, du, gus[e]);
                        gvs[e] = fmaf(gm, dv, gvs[e]);
                    }
                }
                sums_add2(t, pk(gus[0], gus[1]), pk(gvs[0], gvs[1]), u2, v2, r2, y2);
            }
        }
    }
}

// sum of nine per-lane values over the warp: butterfly that halves the number of live values at every step
// (9 + 5 shuffles instead of 45); on return lane 4*i (i < 8) holds total i in r, lane 0 also holds total 8 in r8
__device__ __forceinline__ void warp_reduce9(const float (&v)[9], float& r, float& r8) {
    const unsigned lane = threadIdx.x & 31u;
    const bool b4 = lane & 16u, b3 = lane & 8u, b2 = lane & 4u;
    float a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float keep = b4 ? v[i + 4] : v[i], send = b4 ? v[i] : v[i + 4];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    float c[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float keep = b3 ? a[i + 2] : a[i], send = b3 ? a[i] : a[i + 2];
        c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    const float keep = b2 ? c[1] : c[0], send = b2 ? c[0] : c[1];
    float d = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    r = d;   // lane L holds total index (b4 ? 4 : 0) + (b3 ? 2 : 0) + (b2 ? 1 : 0)
    r8 = warp_sum(v[8]);
}

// The upstream-gradient stream of the backward pass.  Lane l of a warp needs, for its strip, the 16 values of its own
// output column and (channel 0) the 4 pooled-mask gradients above it -- addresses that depend on the item index only, not
// on the staged window or on H.  Every lane therefore copies ITS values for the NEXT strip (of this item, or the first
// one of the next item) into a per-warp shared-memory buffer with 4-byte cp.async (LDGSTS: no registers held while in
// flight, unlike a register prefetch, which at the 96 registers of this CTA shape either spills or is too shallow to
// cover an L2 hit), double-buffered, and reads them back with LDS one strip later.  A lane only ever reads what it
// wrote: cp.async.wait_group is all the synchronisation there is.
__device__ __forceinline__ void cp_async4(uint32_t dst_s32, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_s32), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
template <int kBlk, bool kMask, int kWo>
__device__ __forceinline__ void strip_prefetch(const float* __restrict__ gOut, const float* __restrict__ gMaskPooled, int item, int q, int warp,
                                               int n_planes_items, int n_blocks, int blocks_x, int C, int Ho, int Wo_rt, uint32_t buf_s32) {
    using Cfg = RingCfg<kBlk>;
    const int Wo = kWo > 0 ? kWo : Wo_rt;
    if (item < n_planes_items) {
        const int plane = item / n_blocks, blk = item - plane * n_blocks;
        const int by = blk / blocks_x, bxi = blk - by * blocks_x;
        const int sidx = warp + q * (Cfg::kConsumers / 32);
        const int x = bxi * kBlk + (sidx % Cfg::kStripsX) * 32 + static_cast<int>(threadIdx.x & 31u);
        const int y0 = by * kBlk + (sidx / Cfg::kStripsX) * kStripRows;
        if (x < Wo) {
            const float* gp = gOut + static_cast<size_t>(plane) * Ho * Wo + static_cast<size_t>(y0) * Wo + x;
#pragma unroll
            for (int r = 0; r < kStripRows; ++r)
                if (y0 + r < Ho) cp_async4(buf_s32 + r * 128, gp + r * Wo);
            if (kMask) {
                const int b = plane / C;
                if (plane - b * C == 0) {
                    const float* mp = gMaskPooled + static_cast<size_t>(b) * (Ho >> 2) * (Wo >> 2) + (y0 >> 2) * (Wo >> 2) + (x >> 2);
#pragma unroll
                    for (int gq = 0; gq < 4; ++gq)
                        if (y0 + 4 * gq < Ho) cp_async4(buf_s32 + (kStripRows + gq) * 128, mp + gq * (Wo >> 2));
                }
            }
        }
    }
    cp_async_commit();   // always: the group count must advance uniformly
}

// gbuf_s32 (kImage): shared address of this lane's slot in the landing buffer of the current strip (row r at + r * 128 B,
// pooled-mask cell of group gq at + (16 + gq) * 128 B); kImage = false: mask gradients come straight from global memory.
template <bool kImage, bool kMask, bool kShared, int kWo, int kPitch>
__device__ __forceinline__ void bwd_item(const ItemView& iv, uint32_t gbuf_s32, const float* __restrict__ gMaskPooled,
                                         float (&acc)[9], int Hs, int Ws, int Ho, int Wo_rt) {
    const int Wo = kWo > 0 ? kWo : Wo_rt;
    const float xf = static_cast<float>(iv.x);
    const ColProj cp = col_proj(iv.hm, xf);
    const float* mcell = nullptr;
    if (!kImage && kMask && iv.c == 0)
        mcell = gMaskPooled + static_cast<size_t>(iv.b) * (Ho >> 2) * (Wo >> 2) + (iv.y0 >> 2) * (Wo >> 2) + (iv.x >> 2);
    const float y0f = static_cast<float>(iv.y0);
    StripSums2 t;
    t.sa = t.say = t.sb = t.sby = t.sc = t.scy = dup2(0.0f);
    // (loop deliberately not unrolled: the body is large and the instruction cache is the scarcer resource)
#pragma unroll 1
    for (int gq = 0; gq < 4; ++gq) {
        const int cls = (iv.cls4 >> (4 * gq)) & 0xf;
        float g4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        float gm = 0.0f;
        if (iv.xin && cls <= kBorder) {
            if (kImage) {
#pragma unroll
                for (int j = 0; j < 4; ++j) g4[j] = lds_f32(gbuf_s32 + (4 * gq + j) * 128);
            }
            if (kMask && cls == kBorder && iv.c == 0)
                gm = (kImage ? lds_f32(gbuf_s32 + (kStripRows + gq) * 128) : __ldg(mcell + gq * (Wo >> 2))) * 0.0625f;
        }
        bwd_group<kImage, kMask, kShared, kPitch>(iv, cp, cls, y0f + static_cast<float>(4 * gq), g4, gm, t, Hs, Ws);
    }
    float sa, say, sb, sby, sc, scy, hi;
    upk(t.sa, sa, hi); sa += hi;
    upk(t.say, say, hi); say += hi;
    upk(t.sb, sb, hi); sb += hi;
    upk(t.sby, sby, hi); sby += hi;
    upk(t.sc, sc, hi); sc += hi;
    upk(t.scy, scy, hi); scy += hi;
    acc[0] = sa * xf; acc[1] = say; acc[2] = sa;
    acc[3] = sb * xf; acc[4] = sby; acc[5] = sb;
    acc[6] = -(sc * xf); acc[7] = -scy; acc[8] = -sc;
}

// Every consumer warp writes the nine sums of each of its strips to partials[plane][block][strip][9] (no CTA-wide
// barrier), warp_bwd_finish_kernel adds them in a fixed order (bit reproducible, no atomics).
// kImage: gOut given (source staged); false = mask gradient only.  kMask: the pooled-mask upstream (pool == 4) is folded
// into the same pass (channel 0).
template <int kBlk, bool kImage, bool kMask, int kWo>
__global__ void __launch_bounds__(RingCfg<kBlk>::kThreads, RingCfg<kBlk>::kCtasPerSm)
    warp_bwd_ring_kernel(const float* __restrict__ src, const float* __restrict__ H, const float* __restrict__ gOut,
                         const float* __restrict__ gMaskPooled, float* __restrict__ partials, int C, int Hs, int Ws, int Ho,
                         int Wo, int blocks_x, int n_blocks, int n_items, int n_full) {
    using Cfg = RingCfg<kBlk>;
    constexpr int kS = Cfg::kBwdStages;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full[kS], empty[kS];
    __shared__ ItemHeader header[kS];
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kS; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], Cfg::kConsumers / 32);
        }
        fence_mbar_init();
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    const unsigned lane = threadIdx.x & 31u;
    const bool producer = warp == Cfg::kConsumers / 32;
    // launch items [0, n_full) are whole items; the rest are HALVES of the remaining items (see ring_split)
    const int n_planes_items = n_full + ((n_items - n_full) >> 1);
    auto decode = [&](int it, int& item, int& q_lo, int& q_hi) {
        const bool half = it >= n_full;
        item = half ? n_full + ((it - n_full) >> 1) : it;
        q_lo = half ? ((it - n_full) & 1) * (Cfg::kStripsPerWarp / 2) : 0;
        q_hi = half ? q_lo + Cfg::kStripsPerWarp / 2 : Cfg::kStripsPerWarp;
    };
    // this lane's slots in the two landing buffers of its warp
    const uint32_t gbuf0 = smem_u32(smem_raw + kS * Cfg::kStageBytes) + static_cast<uint32_t>(producer ? 0 : warp) * 2u * kGBufBytes + lane * 4u;
    uint32_t flip = 0u;
    if (kImage && !producer) {   // the first strip of this CTA
        int item, q_lo, q_hi;
        decode(blockIdx.x, item, q_lo, q_hi);
        strip_prefetch<kBlk, kMask, kWo>(gOut, gMaskPooled, blockIdx.x < n_items ? item : n_planes_items, q_lo, warp, n_planes_items, n_blocks,
                                         blocks_x, C, Ho, Wo, gbuf0);
    }
    int k = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++k) {
        const int s = k % kS;
        const uint32_t round = static_cast<uint32_t>(k / kS);
        float* stage = reinterpret_cast<float*>(smem_raw + s * Cfg::kStageBytes);
        int item, q_lo, q_hi;
        decode(it, item, q_lo, q_hi);
        if (producer) {
            if (k >= kS) mbar_wait_sleep(&empty[s], (round - 1u) & 1u);
            ring_produce<kBlk, kImage>(src, H, item, n_blocks, blocks_x, C, &header[s], stage, &full[s], Hs, Ws, Ho, Wo, k >= kS,
                                       kImage ? gOut : nullptr);
        } else {
            // the strip after the last one of this item: the first strip of this CTA's next launch item
            int nitem = n_planes_items, nq_lo = 0, nq_hi = 0;
            if (it + static_cast<int>(gridDim.x) < n_items) decode(it + static_cast<int>(gridDim.x), nitem, nq_lo, nq_hi);
            mbar_wait_hint(&full[s], round & 1u);
            ItemView iv = ring_view(&header[s], src, stage, Hs, Ws);
            const unsigned classes = ring_classify<kBlk, !kImage>(&header[s], iv.hm, warp, Hs, Ws);
#pragma unroll 1
            for (int q = q_lo; q < q_hi; ++q) {
                const bool last = q == q_hi - 1;
                if (kImage) {
                    strip_prefetch<kBlk, kMask, kWo>(gOut, gMaskPooled, last ? nitem : item, last ? nq_lo : q + 1, warp, n_planes_items, n_blocks,
                                                     blocks_x, C, Ho, Wo, gbuf0 + (flip ^ 1u) * kGBufBytes);
                    cp_async_wait_1();   // everything but the group just committed has landed: this strip's values
                }
                ring_view_strip<kBlk>(iv, &header[s], warp, q, classes, Wo);
                float acc[9];
                constexpr int kPitch = (kBlk == 128 && kWo == 128) ? 128 : 0;
                if (iv.wd.shared) bwd_item<kImage, kMask, true, kWo, kPitch>(iv, gbuf0 + flip * kGBufBytes, gMaskPooled, acc, Hs, Ws, Ho, Wo);
                else bwd_item<kImage, kMask, false, kWo, 0>(iv, gbuf0 + flip * kGBufBytes, gMaskPooled, acc, Hs, Ws, Ho, Wo);
                flip ^= 1u;
                if (last) {
                    __syncwarp();
                    if (lane == 0u) mbar_arrive(&empty[s]);   // the window is free while the last sums are reduced
                }
                float r, r8;
                warp_reduce9(acc, r, r8);
                const int sidx = warp + q * (Cfg::kConsumers / 32);
                float* dst = partials + ((static_cast<size_t>(iv.plane) * n_blocks + iv.blk) * Cfg::kStrips + sidx) * 9;
                if ((lane & 3u) == 0u) dst[((lane >> 4) & 1u) * 4u + ((lane >> 3) & 1u) * 2u + ((lane >> 2) & 1u)] = r;
                if (lane == 0u) dst[8] = r8;
            }
        }
    }
}

// =================================================================================================
// tile path (NCHW planes, the default): one CTA per 32x32 output tile, the source box staged by ONE tiled TMA copy.
//
//   work item  = a 32x32 tile of output pixels of one plane: 8192 CTAs of 4 warps for the north-star batch (512 planes of
//                128x128), eight or more of them resident per SM -- the hardware block scheduler balances the load (no
//                persistent rounds, no tail), and the DRAM latency of one CTA's box is covered by its neighbours' math.
//   source box = bounding box of the tile's four projected corners (a projective map with w > 0 sends the tile to a convex
//                quadrilateral), its left edge rounded down to a multiple of four floats (TMA wants the first byte of a box
//                on a 16-byte boundary: other offsets raise an illegal-instruction fault, tools/probes/tma_probe.cu),
//                fetched with cp.async.bulk.tensor.3d (SASS UTMALDG) through one of three tensor maps over
//                [planes, Hs, Ws] whose boxes are 40x36 and 56x52 floats.  The box may sit anywhere, also
//                partly or wholly outside the plane: TMA fills what is out of bounds with ZEROS, which is exactly
//                grid_sample's zeros padding -- the sampler has no border case at all (no clamp, no predicate, no zero
//                frame to write).  Tiles whose box exceeds 56x52 (local scale above ~1.5) or whose w changes sign read the
//                plane through the read-only cache with predicated taps.
//   sampling   = a warp owns 32 columns x 8 rows, lanes on consecutive columns (coalesced 128-byte row stores), four packed
//                row pairs in flight per lane (fp32x2 arithmetic, magic-number floor, see above).
//   pooled mask= analytic coverage, 4x4 mean; a warp whose 32x8 region maps strictly inside the source writes ones.
//   backward   = same staging; the upstream gradient rows of a lane's column are plain coalesced loads issued before the
//                wait on the box; nine sums per warp (butterfly), per-CTA partials in shared memory, one partial row per
//                tile, fixed-order finish kernel behind a programmatic dependent launch.
// =================================================================================================
struct alignas(64) TileMaps {
    CUtensorMap m[2];
};
constexpr int kTile = 32;
// box k: height 36 / 52 rows (local scale up to ~1.0 / ~1.5), width = height + 4 (the slack of the aligned left edge).
// 11.6 KB of shared memory per CTA: nineteen CTAs fit an SM, the register file decides the occupancy.
constexpr int kTileBoxes = 2;
__host__ __device__ constexpr int tile_box_h(int sel) { return sel == 0 ? 36 : 52; }
__host__ __device__ constexpr int tile_box_w(int sel) { return tile_box_h(sel) + 4; }
constexpr int kTileSmem = tile_box_w(kTileBoxes - 1) * tile_box_h(kTileBoxes - 1) * 4;

__device__ __forceinline__ void tma_load_box(void* smem_dst, const CUtensorMap* map, int c0, int r0, int plane, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(r0), "r"(plane), "r"(smem_u32(bar))
                 : "memory");
}

// Geometry of a tile.  Warp 0 computes it (every lane projects one corner, quads combine), its lane 0 initialises the
// barrier and issues the box copy straight away -- nothing between the H load and the TMA request but ~60 instructions --
// and publishes {c0, r0, sel, inside} in shared memory for the other warps.
struct TileHeader {
    int c0, r0, sel, inside;   // box origin, tensor-map index (< 0: no box, predicated global taps), tile sampled strictly inside
};
struct TileIndex {
    long long plane;
    int b, c, x_lo, y_lo, x_hi, y_hi;
};
// grid = (tiles per plane, planes low, planes high): no integer division on the way to the tile's coordinates except the
// exact multiply-high by tiles_x_magic = ceil(2^32 / tiles_x) (t < 65536)
__device__ __forceinline__ TileIndex tile_index(int tiles_x, unsigned tiles_x_magic, int C, int Ho, int Wo) {
    TileIndex g;
    g.plane = static_cast<long long>(blockIdx.z) * gridDim.y + blockIdx.y;
    const int t = blockIdx.x, ty = tiles_x == 1 ? t : static_cast<int>(__umulhi(static_cast<unsigned>(t), tiles_x_magic)), tx = t - ty * tiles_x;
    if (C == 1) { g.b = static_cast<int>(g.plane); g.c = 0; }
    else { g.b = static_cast<int>(g.plane / C); g.c = static_cast<int>(g.plane - static_cast<long long>(g.b) * C); }
    g.x_lo = tx * kTile; g.y_lo = ty * kTile;
    g.x_hi = min(Wo, g.x_lo + kTile); g.y_hi = min(Ho, g.y_lo + kTile);
    return g;
}
// warp 0, all lanes
__device__ __forceinline__ TileHeader tile_header(const Hmat& hm, const TileIndex& g, int Hs, int Ws, bool stage) {
    const int k = threadIdx.x & 3;
    float u, v, w;
    project_rcp(hm, static_cast<float>((k & 1) ? g.x_hi - 1 : g.x_lo), static_cast<float>((k & 2) ? g.y_hi - 1 : g.y_lo), u, v, w);
    // 1e5: the corner coordinates carry ~4e-7 relative error (reciprocal + 3 roundings), covered by the 0.05 px guard
    const bool good = quad_all(w > 0.0f && fabsf(u) < 1.0e5f && fabsf(v) < 1.0e5f);
    const float umin = quad_min(u), umax = quad_max(u), vmin = quad_min(v), vmax = quad_max(v);
    TileHeader h;
    h.c0 = static_cast<int>(floorf(umin - 0.05f)) & ~3;
    h.r0 = static_cast<int>(floorf(vmin - 0.05f));
    const int need_w = static_cast<int>(floorf(umax + 0.05f)) + 2 - h.c0, need_h = static_cast<int>(floorf(vmax + 0.05f)) + 2 - h.r0;
    h.sel = -1;
    if (good && stage) {
#pragma unroll
        for (int s = kTileBoxes - 1; s >= 0; --s)
            if (need_w <= tile_box_w(s) && need_h <= tile_box_h(s)) h.sel = s;
    }
    h.inside = (good && umin >= 0.001f && vmin >= 0.001f && umax < static_cast<float>(Ws - 1) - 0.001f &&
                vmax < static_cast<float>(Hs - 1) - 0.001f) ? 1 : 0;
    return h;
}
// is the 32x8 region of this warp sampled strictly inside the source (coverage == 1, no coverage gradient)?
__device__ __forceinline__ bool region_inside(const Hmat& hm, int xa, int xb, int ya, int yb, int Hs, int Ws) {
    const int k = threadIdx.x & 3;
    float u, v, w;
    project_rcp(hm, static_cast<float>((k & 1) ? xb : xa), static_cast<float>((k & 2) ? yb : ya), u, v, w);
    const bool in = quad_all(w > 0.0f && (u >= 0.001f) && (v >= 0.001f) && (u < static_cast<float>(Ws - 1) - 0.001f) &&
                             (v < static_cast<float>(Hs - 1) - 0.001f));
    return __shfl_sync(0xffffffffu, in ? 1 : 0, 0) != 0;
}
__device__ __forceinline__ Window tile_window(const void* stage, const TileHeader& h, const float* plane_ptr, int Ws) {
    Window wd;
    wd.shared = h.sel >= 0;
    wd.xpred = false;
    if (wd.shared) {
        wd.pitch = tile_box_w(h.sel);
        // byte address of source pixel (0, 0) relative to the box, minus the bias the magic-number floor leaves in `lin`
        const uint32_t origin = static_cast<uint32_t>(-h.r0 * wd.pitch - h.c0) - static_cast<uint32_t>(kMagicBits) * static_cast<uint32_t>(wd.pitch + 1);
        wd.base_s32 = smem_u32(stage) + origin * 4u;
        wd.taps = nullptr;
    } else {
        wd.pitch = Ws;
        wd.base_s32 = 0u;
        wd.taps = plane_ptr;
    }
    return wd;
}
// prologue shared by forward and backward: returns the header; on return the box copy (if any) is in flight
__device__ __forceinline__ TileHeader tile_prologue(const TileMaps& maps, const Hmat& hm, const TileIndex& g, int Hs, int Ws, bool stage_src,
                                                    void* stage, uint64_t* bar, TileHeader* shared_hdr) {
    if (threadIdx.x < 32) {
        const TileHeader h = tile_header(hm, g, Hs, Ws, stage_src);
        if (threadIdx.x == 0) {
            *shared_hdr = h;
            if (h.sel >= 0) {
                mbar_init(bar, 1);
                fence_mbar_init();
                mbar_expect_tx(bar, static_cast<uint32_t>(tile_box_w(h.sel) * tile_box_h(h.sel)) * 4u);
                tma_load_box(stage, &maps.m[h.sel], h.c0, h.r0, static_cast<int>(g.plane), bar);
            }
        }
    }
    __syncthreads();   // header and barrier visible
    return *shared_hdr;
}

// 8 rows of one column out of the staged box: values (and coordinates for the coverage)
// rows of a packed pair; partial tiles (kFull = false) clamp to the tile's last row: every tap stays inside the staged box
template <bool kFull>
__device__ __forceinline__ f2 tile_rows(float yg, int p, float ymax) {
    if (kFull) return add2(dup2(yg), pk(2.0f * p, 2.0f * p + 1.0f));
    return pk(fminf(yg + 2.0f * p, ymax), fminf(yg + (2.0f * p + 1.0f), ymax));
}
template <int kPitch, bool kFull>
__device__ __forceinline__ void tile_sample8(const Hmat& hm, const ColProj& cp, const Window& wd, float yg, float ymax, f2 (&u2)[4],
                                             f2 (&v2)[4], f2 (&o2)[4]) {
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        f2 r2;
        project_col2(hm, cp, tile_rows<kFull>(yg, p, ymax), u2[p], v2[p], r2);
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) o2[p] = lerp2(cell_at2<kPitch>(u2[p], v2[p], wd, false, 0));
}

// one pass of a warp: 32 columns x 8 rows starting at row y0.  kWo = 128: the north-star geometry (Wo == 128, whole tiles
// only) -- store offsets are immediates and the partial-tile clamps and predicates drop out; kWo = 0: run-time sizes.
template <bool kMask, int kWo>
__device__ __forceinline__ void tile_fwd_pass(const Hmat& hm, const TileIndex& g, const TileHeader& h, const Window& wd, const ColProj& cp,
                                              int x, bool xin, int y0, float* __restrict__ out, float* __restrict__ mask_pooled, int Hs,
                                              int Ws, int Ho, int Wo_rt) {
    constexpr bool kFull = kWo > 0;
    const int Wo = kFull ? kWo : Wo_rt;
    if (!kFull && y0 >= g.y_hi) return;
    const int lane = threadIdx.x & 31;
    const float yg = static_cast<float>(y0), ymax = static_cast<float>(g.y_hi - 1);
    const bool want_mask = kMask && g.c == 0;
    // coverage of a tile that is not wholly inside: per 32x8 region first
    bool inside = true;
    if (want_mask && !h.inside) inside = region_inside(hm, g.x_lo, g.x_hi - 1, y0, min(y0 + 7, g.y_hi - 1), Hs, Ws);
    float o8[8];
    float m_top = 1.0f, m_bot = 1.0f;
    if (h.sel >= 0) {
        f2 u2[4], v2[4], o2[4];
        if (h.sel == 0) tile_sample8<tile_box_w(0), kFull>(hm, cp, wd, yg, ymax, u2, v2, o2);
        else tile_sample8<tile_box_w(1), kFull>(hm, cp, wd, yg, ymax, u2, v2, o2);
#pragma unroll
        for (int p = 0; p < 4; ++p) upk(o2[p], o8[2 * p], o8[2 * p + 1]);
        if (kMask && !inside) {
            const float Wsf = static_cast<float>(Ws), Hsf = static_cast<float>(Hs);
            float cv[8];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                float u0, u1, v0, v1;
                upk(u2[p], u0, u1);
                upk(v2[p], v0, v1);
                cv[2 * p] = cover1(u0, Wsf) * cover1(v0, Hsf);
                cv[2 * p + 1] = cover1(u1, Wsf) * cover1(v1, Hsf);
            }
            m_top = (cv[0] + cv[1]) + (cv[2] + cv[3]);
            m_bot = (cv[4] + cv[5]) + (cv[6] + cv[7]);
        }
    } else {
        m_top = m_bot = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float u, v, rw, nw, ne, sw, se;
            project_col(hm, cp, fminf(yg + static_cast<float>(j), ymax), u, v, rw);
            const Taps t = make_taps_fast(u, v, Ws, Hs);
            global_taps(t, wd.taps, Ws, nw, ne, sw, se);
            o8[j] = blend(t, nw, ne, sw, se);
            if (kMask) {
                const float cvr = cover(t);
                if (j < 4) m_top += cvr; else m_bot += cvr;
            }
        }
        inside = false;
    }
    if (kFull || xin) {
        float* og = out + (g.plane * Ho + y0) * Wo + x;
        if (kFull || y0 + 8 <= g.y_hi) {
#pragma unroll
            for (int j = 0; j < 8; ++j) st_stream1(og + j * Wo, o8[j]);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (y0 + j < g.y_hi) st_stream1(og + j * Wo, o8[j]);
        }
    }
    if (want_mask) {
        if (!inside) {
            m_top += __shfl_xor_sync(0xffffffffu, m_top, 1); m_top += __shfl_xor_sync(0xffffffffu, m_top, 2);
            m_bot += __shfl_xor_sync(0xffffffffu, m_bot, 1); m_bot += __shfl_xor_sync(0xffffffffu, m_bot, 2);
            m_top *= 0.0625f; m_bot *= 0.0625f;
        }
        if ((kFull || xin) && (lane & 3) == 0) {
            float* mc = mask_pooled + (static_cast<size_t>(g.b) * (Ho >> 2) + (y0 >> 2)) * (Wo >> 2) + (x >> 2);
            mc[0] = m_top;
            if (kFull || y0 + 4 < g.y_hi) mc[Wo >> 2] = m_bot;
        }
    }
}

// kWarps = 4: a warp owns 8 rows of the tile; kWarps = 2: 16 rows in two passes (the per-thread prologue -- H, column
// projection, window set-up -- is paid once per 16 pixels instead of 8)
template <bool kMask, int kWarps, int kWo, int kMinBlocks>
__global__ void __launch_bounds__(32 * kWarps, kMinBlocks)
    warp_fwd_tile_kernel(const __grid_constant__ TileMaps maps, const float* __restrict__ src, const float* __restrict__ H,
                         float* __restrict__ out, float* __restrict__ mask_pooled, int C, int Hs, int Ws, int Ho, int Wo, int tiles_x,
                         unsigned tiles_x_magic, long long n_planes) {
    extern __shared__ __align__(128) unsigned char stage[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ TileHeader hdr;
    const TileIndex g = tile_index(tiles_x, tiles_x_magic, C, Ho, Wo);
    if (g.plane >= n_planes) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Hmat hm = load_h(H, g.b);
    const TileHeader h = tile_prologue(maps, hm, g, Hs, Ws, true, stage, &bar, &hdr);
    const int x = g.x_lo + lane;
    const bool xin = x < g.x_hi;
    // lanes / rows beyond a partial tile recompute its last column / row: every tap stays inside the staged box
    const ColProj cp = col_proj(hm, static_cast<float>(min(x, g.x_hi - 1)));
    const Window wd = tile_window(stage, h, h.sel >= 0 ? nullptr : src + g.plane * Hs * Ws, Ws);
    if (h.sel >= 0) mbar_wait(&bar, 0u);
    constexpr int kPasses = 4 / kWarps;
#pragma unroll 1
    for (int pass = 0; pass < kPasses; ++pass)
        tile_fwd_pass<kMask, kWo>(hm, g, h, wd, cp, x, xin, g.y_lo + 8 * (warp * kPasses + pass), out, mask_pooled, Hs, Ws, Ho, Wo);
}

// dH partial sums of one tile.  kImage: upstream image gradient given (source box staged); kMask: pooled-mask upstream
// (pool == 4) folded into the same pass on the tiles of channel 0.  partials[plane][tile][9].
template <bool kImage, int kWo>
__device__ __forceinline__ void tile_load_g8(float (&g8)[8], const float* __restrict__ gOut, const TileIndex& g, int x, bool xin, int y0, int Ho,
                                             int Wo_rt) {
    constexpr bool kFull = kWo > 0;
    const int Wo = kFull ? kWo : Wo_rt;
#pragma unroll
    for (int j = 0; j < 8; ++j) g8[j] = 0.0f;
    if (kImage && (kFull || (xin && y0 < g.y_hi))) {
        const float* gp = gOut + (g.plane * Ho + y0) * Wo + x;
        if (kFull || y0 + 8 <= g.y_hi) {
#pragma unroll
            for (int j = 0; j < 8; ++j) g8[j] = ld_stream1(gp + j * Wo);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (y0 + j < g.y_hi) g8[j] = ld_stream1(gp + j * Wo);
        }
    }
}
// one pass of a warp of the backward: 32 columns x 8 rows starting at y0, sums into t
template <bool kImage, bool kMask, int kWo>
__device__ __forceinline__ void tile_bwd_pass(const Hmat& hm, const TileIndex& g, const TileHeader& h, const Window& wd, const ColProj& cp,
                                              int x, bool xin, int y0, const float (&g8)[8], const float* __restrict__ gMaskPooled,
                                              StripSums2& t, int Hs, int Ws, int Ho, int Wo_rt) {
    constexpr bool kFull = kWo > 0;
    const int Wo = kFull ? kWo : Wo_rt;
    const bool live = kFull || (xin && y0 < g.y_hi);
    bool mask_live = false;
    float gm_top = 0.0f, gm_bot = 0.0f;
    if (kMask && g.c == 0 && (kFull || y0 < g.y_hi) && !h.inside) {
        mask_live = !region_inside(hm, g.x_lo, g.x_hi - 1, y0, min(y0 + 7, g.y_hi - 1), Hs, Ws);
        if (mask_live && (kFull || xin)) {
            const float* mc = gMaskPooled + (static_cast<size_t>(g.b) * (Ho >> 2) + (y0 >> 2)) * (Wo >> 2) + (x >> 2);
            gm_top = __ldg(mc) * 0.0625f;
            if (kFull || y0 + 4 < g.y_hi) gm_bot = __ldg(mc + (Wo >> 2)) * 0.0625f;
        }
    }
    if (!(live && (kImage || mask_live))) return;
    const float yg = static_cast<float>(y0), ymax = static_cast<float>(g.y_hi - 1);
    const float Wsf = static_cast<float>(Ws), Hsf = static_cast<float>(Hs);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const f2 y2 = tile_rows<kFull>(yg, p, ymax);
        f2 u2, v2, r2, gu = dup2(0.0f), gv = dup2(0.0f);
        project_col2(hm, cp, y2, u2, v2, r2);
        float u0, u1, v0, v1;
        upk(u2, u0, u1);
        upk(v2, v0, v1);
        if (kImage) {
            const f2 g2 = pk(g8[2 * p], g8[2 * p + 1]);
            f2 du, dv;
            if (h.sel >= 0) {
                Cell8 c;
                if (h.sel == 0) c = cell_at2<tile_box_w(0)>(u2, v2, wd, false, 0);
                else c = cell_at2<tile_box_w(1)>(u2, v2, wd, false, 0);
                cell_grad2(c, du, dv);
            } else {
                float d0, e0, d1, e1, nw, ne, sw, se;
                const Taps t0 = make_taps_fast(u0, v0, Ws, Hs), t1 = make_taps_fast(u1, v1, Ws, Hs);
                global_taps(t0, wd.taps, Ws, nw, ne, sw, se);
                blend_grad(t0, nw, ne, sw, se, d0, e0);
                global_taps(t1, wd.taps, Ws, nw, ne, sw, se);
                blend_grad(t1, nw, ne, sw, se, d1, e1);
                du = pk(d0, d1);
                dv = pk(e0, e1);
            }
            gu = mul2(g2, du);
            gv = mul2(g2, dv);
        }
        if (kMask && mask_live) {
            const float gm = p < 2 ? gm_top : gm_bot;
            gu = fma2(dup2(gm), pk(cover1(v0, Hsf) * cover1_grad(u0, Wsf), cover1(v1, Hsf) * cover1_grad(u1, Wsf)), gu);
            gv = fma2(dup2(gm), pk(cover1(u0, Wsf) * cover1_grad(v0, Hsf), cover1(u1, Wsf) * cover1_grad(v1, Hsf)), gv);
        }
        // rows below the output (partial tiles) carry g = 0 and gm = 0: no contribution
        sums_add2(t, gu, gv, u2, v2, r2, y2);
    }
}

// dH partial sums of one tile.  kImage: upstream image gradient given (source box staged); kMask: pooled-mask upstream
// (pool == 4) folded into the same pass on the tiles of channel 0.  partials[plane][tile][9].
template <bool kImage, bool kMask, int kWarps, int kWo, int kMinBlocks>
__global__ void __launch_bounds__(32 * kWarps, kMinBlocks)
    warp_bwd_tile_kernel(const __grid_constant__ TileMaps maps, const float* __restrict__ src, const float* __restrict__ H,
                         const float* __restrict__ gOut, const float* __restrict__ gMaskPooled, float* __restrict__ partials, int C,
                         int Hs, int Ws, int Ho, int Wo, int tiles_x, unsigned tiles_x_magic, long long n_planes) {
    extern __shared__ __align__(128) unsigned char stage[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ TileHeader hdr;
    __shared__ float red[kWarps][9];
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the finish kernel may be scheduled; it waits for this grid
    const TileIndex g = tile_index(tiles_x, tiles_x_magic, C, Ho, Wo);
    if (g.plane >= n_planes) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const Hmat hm = load_h(H, g.b);
    const int x = g.x_lo + lane;
    const bool xin = x < g.x_hi;
    constexpr int kPasses = 4 / kWarps;
    // the upstream gradients of this lane's column: the first pass' are issued before the box is even requested
    float g8[8];
    tile_load_g8<kImage, kWo>(g8, gOut, g, x, xin, g.y_lo + 8 * warp * kPasses, Ho, Wo);
    const TileHeader h = tile_prologue(maps, hm, g, Hs, Ws, kImage, stage, &bar, &hdr);
    const float xf = static_cast<float>(min(x, g.x_hi - 1));
    const ColProj cp = col_proj(hm, xf);
    const Window wd = tile_window(stage, h, (kImage && h.sel < 0) ? src + g.plane * Hs * Ws : nullptr, Ws);
    StripSums2 t;
    t.sa = t.say = t.sb = t.sby = t.sc = t.scy = dup2(0.0f);
    if (kImage && h.sel >= 0) mbar_wait(&bar, 0u);
#pragma unroll
    for (int pass = 0; pass < kPasses; ++pass) {
        const int y0 = g.y_lo + 8 * (warp * kPasses + pass);
        float gnext[8];
        if (pass + 1 < kPasses) tile_load_g8<kImage, kWo>(gnext, gOut, g, x, xin, y0 + 8, Ho, Wo);   // in flight while this pass computes
        tile_bwd_pass<kImage, kMask, kWo>(hm, g, h, wd, cp, x, xin, y0, g8, gMaskPooled, t, Hs, Ws, Ho, Wo);
        if (pass + 1 < kPasses) {
#pragma unroll
            for (int j = 0; j < 8; ++j) g8[j] = gnext[j];
        }
    }
    float acc[9];
    {
        float sa, say, sb, sby, sc, scy, hi;
        upk(t.sa, sa, hi); sa += hi;
        upk(t.say, say, hi); say += hi;
        upk(t.sb, sb, hi); sb += hi;
        upk(t.sby, sby, hi); sby += hi;
        upk(t.sc, sc, hi); sc += hi;
        upk(t.scy, scy, hi); scy += hi;
        acc[0] = sa * xf; acc[1] = say; acc[2] = sa;
        acc[3] = sb * xf; acc[4] = sby; acc[5] = sb;
        acc[6] = -(sc * xf); acc[7] = -scy; acc[8] = -sc;
    }
    float r, r8;
    warp_reduce9(acc, r, r8);
    if ((lane & 3) == 0) red[warp][((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)] = r;
    if (lane == 0) red[warp][8] = r8;
    __syncthreads();
    if (threadIdx.x < 9) {
        float sum = red[0][threadIdx.x];
#pragma unroll
        for (int k = 1; k < kWarps; ++k) sum += red[k][threadIdx.x];
        partials[(static_cast<size_t>(g.plane) * gridDim.x + blockIdx.x) * 9 + threadIdx.x] = sum;
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

// ---- channels-last feature maps, warp-cooperative geometry ---------------------------------------------------------------
// The thread-per-(pixel, channel quad) kernel above repeats the projection, the two divisions and the 64-bit index arithmetic of
// a pixel in every one of its C/4 threads: 235 instructions per 16 output bytes, 70 % SM utilisation at 34 % of the DRAM
// bandwidth (profiles/r01d_ncu_kernels.csv).  Here a warp takes 32 consecutive output pixels of one sample; lane L computes
// the tap geometry of pixel L ONCE, then the warp visits the pixels in groups of 32 / LPP, the geometry arrives by shuffle
// and every lane moves 16-byte channel quads: 512 contiguous bytes per tap and warp.
// LPP = lanes per pixel = min(32, C/4) (a power of two); a lane owns C/4/LPP quads of its pixel.
// Measured alternative (round 2, removed): one CTA per 8x8 output tile and 64-channel slab with the tile's source box
// (16 x 16 x 64 floats = 64 KB) staged by one cp.async.bulk.tensor.4d copy and the taps read from shared memory -- 66 % of the
// HBM peak at B = 64, C = 64 against 60 % for this kernel (68 % at B = 256, 73 % at C = 256): the explicit reuse buys little
// once the L1 path has enough warps in flight (ncu: 45 % L1 hit rate, long-scoreboard bound at 24 warps per SM -> 32), and
// the staged kernel never returned on a 2x up-sampling case whose boxes lie wholly outside a 32 x 32 source.
struct TapPack {
    int off;           // (y0 * Ws + x0) * C
    int bits;          // validity of nw / ne / sw / se (bit 0..3), bit 4 = the pixel exists
    float wx0, wx1, wy0, wy1;
};
__device__ __forceinline__ TapPack pack_taps(const Taps& t, int Ws, int C, bool live) {
    TapPack p;
    p.off = (t.y0 * Ws + t.x0) * C;
    p.bits = live ? (16 | (t.inx0 && t.iny0 ? 1 : 0) | (t.inx1 && t.iny0 ? 2 : 0) | (t.inx0 && t.iny1 ? 4 : 0) | (t.inx1 && t.iny1 ? 8 : 0)) : 0;
    p.wx0 = t.wx0; p.wx1 = t.wx1; p.wy0 = t.wy0; p.wy1 = t.wy1;
    return p;
}
__device__ __forceinline__ TapPack shfl_taps(const TapPack& p, int src_lane) {
    TapPack r;
    r.off = __shfl_sync(0xffffffffu, p.off, src_lane);
    r.bits = __shfl_sync(0xffffffffu, p.bits, src_lane);
    r.wx0 = __shfl_sync(0xffffffffu, p.wx0, src_lane);
    r.wx1 = __shfl_sync(0xffffffffu, p.wx1, src_lane);
    r.wy0 = __shfl_sync(0xffffffffu, p.wy0, src_lane);
    r.wy1 = __shfl_sync(0xffffffffu, p.wy1, src_lane);
    return r;
}
__device__ __forceinline__ float blend_w(const TapPack& t, float nw, float ne, float sw, float se) {
    return fmaf(fmaf(nw, t.wx0, ne * t.wx1), t.wy0, fmaf(sw, t.wx0, se * t.wx1) * t.wy1);   // == blend()
}

template <int LPP>
__global__ void __launch_bounds__(256, 4)
    warp_fwd_nhwc_coop_kernel(const float* __restrict__ src, const float* __restrict__ H, float* __restrict__ out, int C, int Hs,
                              int Ws, int Ho, int Wo) {
    constexpr int G = 32 / LPP;                       // pixels the warp handles at a time
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int sub = lane / LPP, ql = lane % LPP;
    const int qpl = (C >> 2) / LPP;
    const Hmat hm = load_h(H, b);
    const int npix = Ho * Wo;
    const float* sb = src + static_cast<long long>(b) * Hs * Ws * C;
    float* ob = out + static_cast<long long>(b) * npix * C;
    const int row = Ws * C;
    const int warps = gridDim.x * (blockDim.x >> 5);
    for (int r0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; r0 < npix; r0 += warps * 32) {
        const int r = min(r0 + lane, npix - 1);
        const int y = r / Wo, x = r - y * Wo;
        float u, v, rw;
        project(hm, static_cast<float>(x), static_cast<float>(y), u, v, rw);
        const TapPack mine = pack_taps(make_taps(u, v, Ws, Hs), Ws, C, r0 + lane < npix);
#pragma unroll 2
        for (int j = 0; j < 32; j += G) {
            const TapPack t = shfl_taps(mine, j + sub);
            if (!(t.bits & 16)) continue;
            const float* sp = sb + t.off + 4 * ql;
            float* op = ob + static_cast<long long>(r0 + j + sub) * C + 4 * ql;
            for (int k = 0; k < qpl; ++k, sp += 4 * LPP, op += 4 * LPP) {
                const float4 nw = ld4_or_zero(sp, t.bits & 1);
                const float4 ne = ld4_or_zero(sp + C, t.bits & 2);
                const float4 sw = ld4_or_zero(sp + row, t.bits & 4);
                const float4 se = ld4_or_zero(sp + row + C, t.bits & 8);
                float4 o;
                o.x = blend_w(t, nw.x, ne.x, sw.x, se.x);
                o.y = blend_w(t, nw.y, ne.y, sw.y, se.y);
                o.z = blend_w(t, nw.z, ne.z, sw.z, se.z);
                o.w = blend_w(t, nw.w, ne.w, sw.w, se.w);
                stg_stream(reinterpret_cast<float4*>(op), o);
            }
        }
    }
}

// backward -> dH (no image gradient): same walk; the channel sums of d out/du, d out/dv are reduced over the LPP lanes of a
// pixel and handed back to the lane that owns the pixel's geometry, which accumulates the nine dH terms.
// grid = (chunks, B), partials[b][chunk][9] as warp_bwd_generic_kernel.
template <int LPP>
__global__ void __launch_bounds__(256, 3)
    warp_bwd_nhwc_coop_kernel(const float* __restrict__ src, const float* __restrict__ H, const float* __restrict__ gOut,
                              const float* __restrict__ gMaskPooled, float* __restrict__ partials, int C, int Hs, int Ws, int Ho,
                              int Wo, int pool) {
    constexpr int G = 32 / LPP;
    __shared__ float red[9 * 8];
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int sub = lane / LPP, ql = lane % LPP;
    const int qpl = (C >> 2) / LPP;
    const Hmat hm = load_h(H, b);
    const int npix = Ho * Wo;
    const float* sb = src + static_cast<long long>(b) * Hs * Ws * C;
    const float* gb = gOut + static_cast<long long>(b) * npix * C;
    const int row = Ws * C;
    const int warps = gridDim.x * (blockDim.x >> 5);
    float acc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.0f;
    for (int r0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; r0 < npix; r0 += warps * 32) {
        const bool live = r0 + lane < npix;
        const int r = min(r0 + lane, npix - 1);
        const int y = r / Wo, x = r - y * Wo;
        float u, v, rw;
        project(hm, static_cast<float>(x), static_cast<float>(y), u, v, rw);
        float cdu = 0.0f, cdv = 0.0f;          // coverage gradient of this lane's pixel (the taps themselves need not stay live)
        TapPack mine;
        {
            const Taps taps = make_taps(u, v, Ws, Hs);
            mine = pack_taps(taps, Ws, C, live);
            if (gMaskPooled != nullptr) cover_grad(taps, cdu, cdv);
        }
        float my_gu = 0.0f, my_gv = 0.0f;
#pragma unroll 1
        for (int j = 0; j < 32; j += G) {
            const TapPack t = shfl_taps(mine, j + sub);
            float gu = 0.0f, gv = 0.0f;
            if (t.bits & 16) {
                const float* sp = sb + t.off + 4 * ql;
                const float* gp = gb + static_cast<long long>(r0 + j + sub) * C + 4 * ql;
                for (int k = 0; k < qpl; ++k, sp += 4 * LPP, gp += 4 * LPP) {
                    const float4 g = __ldg(reinterpret_cast<const float4*>(gp));
                    const float4 nw = ld4_or_zero(sp, t.bits & 1);
                    const float4 ne = ld4_or_zero(sp + C, t.bits & 2);
                    const float4 sw = ld4_or_zero(sp + row, t.bits & 4);
                    const float4 se = ld4_or_zero(sp + row + C, t.bits & 8);
                    // blend_grad per channel: du = (ne - nw) wy0 + (se - sw) wy1, dv = (sw - nw) wx0 + (se - ne) wx1
                    gu = fmaf(g.x, fmaf(ne.x - nw.x, t.wy0, (se.x - sw.x) * t.wy1), gu); gv = fmaf(g.x, fmaf(sw.x - nw.x, t.wx0, (se.x - ne.x) * t.wx1), gv);
                    gu = fmaf(g.y, fmaf(ne.y - nw.y, t.wy0, (se.y - sw.y) * t.wy1), gu); gv = fmaf(g.y, fmaf(sw.y - nw.y, t.wx0, (se.y - ne.y) * t.wx1), gv);
                    gu = fmaf(g.z, fmaf(ne.z - nw.z, t.wy0, (se.z - sw.z) * t.wy1), gu); gv = fmaf(g.z, fmaf(sw.z - nw.z, t.wx0, (se.z - ne.z) * t.wx1), gv);
                    gu = fmaf(g.w, fmaf(ne.w - nw.w, t.wy0, (se.w - sw.w) * t.wy1), gu); gv = fmaf(g.w, fmaf(sw.w - nw.w, t.wx0, (se.w - ne.w) * t.wx1), gv);
                }
            }
#pragma unroll
            for (int o = LPP >> 1; o > 0; o >>= 1) {
                gu += __shfl_xor_sync(0xffffffffu, gu, o);
                gv += __shfl_xor_sync(0xffffffffu, gv, o);
            }
            // pixel j + s lives in lanes [s LPP, (s+1) LPP): lane L fetches its own pixel's sums in the iteration that covers it
            const int from = ((lane - j) & (G - 1)) * LPP;
            const float ru = __shfl_sync(0xffffffffu, gu, from), rv = __shfl_sync(0xffffffffu, gv, from);
            if (lane >= j && lane < j + G) { my_gu = ru; my_gv = rv; }
        }
        if (live) {
            if (gMaskPooled != nullptr) {
                const float gm = __ldg(gMaskPooled + (static_cast<long long>(b) * (Ho / pool) + y / pool) * (Wo / pool) + x / pool) /
                                 static_cast<float>(pool * pool);
                my_gu = fmaf(gm, cdu, my_gu);
                my_gv = fmaf(gm, cdv, my_gv);
            }
            accum_gh(acc, my_gu, my_gv, u, v, rw, static_cast<float>(x), static_cast<float>(y));
        }
    }
    block_sum<9>(acc, red);
    store9(acc, partials + (static_cast<long long>(b) * gridDim.x + blockIdx.x) * 9);
}

// one 128-bit reduction (SASS REDG.E.ADD.F32x4) instead of four scalar atomics
__device__ __forceinline__ void red_add_v4(float* p, const float4& v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 scale4(const float4& g, float w) { return make_float4(g.x * w, g.y * w, g.z * w, g.w * w); }
__device__ __forceinline__ float4 add4(const float4& a, const float4& b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 shfl_up4(const float4& v, int d) {
    return make_float4(__shfl_up_sync(0xffffffffu, v.x, d), __shfl_up_sync(0xffffffffu, v.y, d), __shfl_up_sync(0xffffffffu, v.z, d),
                       __shfl_up_sync(0xffffffffu, v.w, d));
}

// Backward, generic layouts.  grid = (chunks, B); every block reduces its pixel chunk to 9 partial sums
// (partials[b][chunk][9]); warp_bwd_finish_kernel adds them in order.
// Optional image gradient (gSrc, accumulated): WARP-AGGREGATED reductions.  The lanes of a warp walk neighbouring output
// pixels of a row (`agg` lanes apart: 1 in the planar layouts, C/4 in channels-last), and at scales near one the right-hand
// taps of a pixel are the left-hand taps of its neighbour: the neighbour's two right-hand contributions arrive by shuffle
// and are added in registers, so a pixel issues two reductions instead of four (the warp's first pixel: four); in
// channels-last each of them is one 128-bit red.global.add.v4.f32 for four channels.  The merge is keyed on the tap
// addresses being EQUAL, nothing else, so it is exact whatever the homography does.
template <bool kVec4>
__global__ void __launch_bounds__(256)
    warp_bwd_generic_kernel(const float* __restrict__ src, const float* __restrict__ H, const float* __restrict__ gOut,
                            const float* __restrict__ gMaskPooled, float* __restrict__ partials, float* __restrict__ gSrc,
                            int C, int Hs, int Ws, int Ho, int Wo, int pool, Layout ls, Layout lo) {
    __shared__ float red[9 * 8];
    constexpr unsigned kFull = 0xffffffffu;
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const Hmat hm = load_h(H, b);
    float acc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k] = 0.0f;
    const int npix = Ho * Wo;
    // kVec4 (channels-last): a group of cq threads shares a pixel and splits the channels 4 by 4
    const int cq = kVec4 ? (C >> 2) : 1;
    const long long nwork = static_cast<long long>(npix) * cq;
    // lanes `agg` apart hold the same channels of x-neighbouring pixels (0: no such lane inside the warp)
    const int agg = (gSrc != nullptr && gOut != nullptr && cq < 32 && (cq & (cq - 1)) == 0) ? cq : 0;
    // warp-uniform trip count: the shuffles below need every lane of the warp
    for (long long w0 = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x - lane; w0 < nwork;
         w0 += static_cast<long long>(gridDim.x) * blockDim.x) {
        const bool live = w0 + lane < nwork;
        const long long wi = live ? w0 + lane : nwork - 1;
        const int q = kVec4 ? static_cast<int>(wi % cq) : 0;
        const int r = static_cast<int>(wi / cq);
        const int y = r / Wo, x = r - y * Wo;
        float u, v, rw;
        project(hm, static_cast<float>(x), static_cast<float>(y), u, v, rw);
        const Taps t = make_taps(u, v, Ws, Hs);
        // does this lane take over the right-hand taps of the lane `agg` below it / were its own taken by the lane above?
        bool absorb = false, absorbed = false;
        if (agg > 0) {
            const int px0 = __shfl_up_sync(kFull, t.x0, agg), py0 = __shfl_up_sync(kFull, t.y0, agg);
            const int plive = __shfl_up_sync(kFull, static_cast<int>(live), agg);
            absorb = lane >= agg && live && plive != 0 && px0 + 1 == t.x0 && py0 == t.y0;
            absorbed = __shfl_down_sync(kFull, static_cast<int>(absorb), agg) != 0 && lane + agg < 32;
        }
        float gu = 0.0f, gv = 0.0f;
        if (gMaskPooled != nullptr && q == 0 && live) {
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
            const float w00 = t.wx0 * t.wy0, w10 = t.wx1 * t.wy0, w01 = t.wx0 * t.wy1, w11 = t.wx1 * t.wy1;
            if (kVec4) {
                const long long o = (static_cast<long long>(t.y0) * Ws + t.x0) * C + q * 4;
                float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live) {
                    g = __ldg(reinterpret_cast<const float4*>(gp + q * 4));
                    const float4 nw = ld4_or_zero(sp + o, t.inx0 && t.iny0);
                    const float4 ne = ld4_or_zero(sp + o + C, t.inx1 && t.iny0);
                    const float4 sw = ld4_or_zero(sp + o + static_cast<long long>(Ws) * C, t.inx0 && t.iny1);
                    const float4 se = ld4_or_zero(sp + o + static_cast<long long>(Ws) * C + C, t.inx1 && t.iny1);
                    float du, dv;
                    blend_grad(t, nw.x, ne.x, sw.x, se.x, du, dv); gu = fmaf(g.x, du, gu); gv = fmaf(g.x, dv, gv);
                    blend_grad(t, nw.y, ne.y, sw.y, se.y, du, dv); gu = fmaf(g.y, du, gu); gv = fmaf(g.y, dv, gv);
                    blend_grad(t, nw.z, ne.z, sw.z, se.z, du, dv); gu = fmaf(g.z, du, gu); gv = fmaf(g.z, dv, gv);
                    blend_grad(t, nw.w, ne.w, sw.w, se.w, du, dv); gu = fmaf(g.w, du, gu); gv = fmaf(g.w, dv, gv);
                }
                if (gs) {
                    float4 c_nw = scale4(g, w00), c_sw = scale4(g, w01);
                    const float4 c_ne = scale4(g, w10), c_se = scale4(g, w11);
                    if (agg > 0) {
                        const float4 r_ne = shfl_up4(c_ne, agg), r_se = shfl_up4(c_se, agg);
                        if (absorb) { c_nw = add4(c_nw, r_ne); c_sw = add4(c_sw, r_se); }
                    }
                    if (live) {
                        if (t.inx0 && t.iny0) red_add_v4(gs + o, c_nw);
                        if (t.inx0 && t.iny1) red_add_v4(gs + o + static_cast<long long>(Ws) * C, c_sw);
                        if (!absorbed) {
                            if (t.inx1 && t.iny0) red_add_v4(gs + o + C, c_ne);
                            if (t.inx1 && t.iny1) red_add_v4(gs + o + static_cast<long long>(Ws) * C + C, c_se);
                        }
                    }
                }
            } else {
                for (int c = 0; c < C; ++c) {
                    float g = 0.0f;
                    if (live) {
                        g = __ldg(gp + static_cast<long long>(c) * lo.sc);
                        float nw, ne, sw, se, du, dv;
                        gather4(sp + static_cast<long long>(c) * ls.sc, t, ls.sy, ls.sx, nw, ne, sw, se);
                        blend_grad(t, nw, ne, sw, se, du, dv);
                        gu = fmaf(g, du, gu);
                        gv = fmaf(g, dv, gv);
                    }
                    if (gs) {
                        float c_nw = g * w00, c_sw = g * w01;
                        const float c_ne = g * w10, c_se = g * w11;
                        if (agg > 0) {
                            const float r_ne = __shfl_up_sync(kFull, c_ne, agg), r_se = __shfl_up_sync(kFull, c_se, agg);
                            if (absorb) { c_nw += r_ne; c_sw += r_se; }
                        }
                        if (live) {
                            float* gc = gs + static_cast<long long>(c) * ls.sc + static_cast<long long>(t.y0) * ls.sy +
                                        static_cast<long long>(t.x0) * ls.sx;
                            if (t.inx0 && t.iny0) atomicAdd(gc, c_nw);
                            if (t.inx0 && t.iny1) atomicAdd(gc + ls.sy, c_sw);
                            if (!absorbed) {
                                if (t.inx1 && t.iny0) atomicAdd(gc + ls.sx, c_ne);
                                if (t.inx1 && t.iny1) atomicAdd(gc + ls.sy + ls.sx, c_se);
                            }
                        }
                    }
                }
            }
        }
        if (live) accum_gh(acc, gu, gv, u, v, rw, static_cast<float>(x), static_cast<float>(y));
    }
    block_sum<9>(acc, red);
    store9(acc, partials + (static_cast<long long>(b) * gridDim.x + blockIdx.x) * 9);
}

__global__ void warp_bwd_finish_kernel(const float* __restrict__ partials, float* __restrict__ gH, int B, int chunks) {
    asm volatile("griddepcontrol.wait;" ::: "memory");   // programmatic dependent launch: the partials are complete and visible
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * 9) return;
    const int b = i / 9, k = i - b * 9;
    float s = 0.0f;
    for (int c = 0; c < chunks; ++c) s += partials[(static_cast<long long>(b) * chunks + c) * 9 + k];
    gH[i] = s;
}

// ---- host-side path selection --------------------------------------------------------------------
// ring path: NCHW tensors whose output is made of whole 4x4 tiles and whose coordinates stay below 2^22; everything else
// (channels-last, odd sizes) goes to the nhwc / generic kernels
inline bool ring_ok(int Hs, int Ws, int Ho, int Wo, int channels_last) {
    return !channels_last && (Ho % 4) == 0 && (Wo % 4) == 0 && Hs < (1 << 20) && Ws < (1 << 20) && Ho < (1 << 20) && Wo < (1 << 20) &&
           static_cast<long long>(Hs) * Ws < (1ll << 30) && static_cast<long long>(Ho) * Wo < (1ll << 30);
}
// item size: whole planes when the source fits one 72 KB stage with its frame, 64x64 blocks otherwise
inline int ring_blk(int Hs, int Ws) {
    return (static_cast<long long>(Hs + kPadT + kPadB) * Ws * 4 + 32 <= RingCfg<128>::kStageBytes && (Ws & 3) == 0) ? 128 : 64;
}
inline int blocks_of(int n, int side) { return (n + side - 1) / side; }
inline int bwd_chunks(int Ho, int Wo, int C, int vec) {
    const long long work = static_cast<long long>(Ho) * Wo * (vec ? C / 4 : 1);
    long long c = (work + 256 * 16 - 1) / (256 * 16);
    return static_cast<int>(c < 1 ? 1 : (c > 1024 ? 1024 : c));
}
template <int kBlk>
inline int ring_grid(long long n_items) {
    const long long cap = static_cast<long long>(kNumSMs) * RingCfg<kBlk>::kCtasPerSm;
    return static_cast<int>(n_items < cap ? n_items : cap);
}
// Tail balance.  Items are dealt round-robin to `cap` persistent CTAs, so n_items = r * cap + R costs r + 1 rounds.  When
// R <= cap / 2 (B = 256 pairs: 512 planes on 148 SMs, R = 68) the last R items are launched as 2 R half-items -- the two
// halves of an item stage the same window and take half of each warp's strips -- and the last round takes half the time.
// Returns the number of launch items; n_full = how many of them are whole.
template <int kBlk>
inline long long ring_split(long long n_items, int& n_full) {
    const long long cap = static_cast<long long>(kNumSMs) * RingCfg<kBlk>::kCtasPerSm;
    const long long R = n_items % cap;
    n_full = static_cast<int>(n_items);
    if (RingCfg<kBlk>::kStripsPerWarp < 2 || R == 0 || 2 * R > cap) return n_items;
    n_full = static_cast<int>(n_items - R);
    return n_items + R;
}
inline int grid_for(long long n, int threads) {
    long long g = (n + threads - 1) / threads;
    const long long cap = static_cast<long long>(kNumSMs) * 16;
    return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

template <int kBlk>
inline int launch_fwd_ring(const float* src, const float* H, float* out, float* mask_pooled, int B, int C, int Hs, int Ws, int Ho, int Wo,
                           bool fuse_mask, cudaStream_t stream) {
    using Cfg = RingCfg<kBlk>;
    const int blocks_x = blocks_of(Wo, kBlk), n_blocks = blocks_x * blocks_of(Ho, kBlk);
    const long long n_items = static_cast<long long>(B) * C * n_blocks;
    void (*kern)(const float*, const float*, float*, float*, int, int, int, int, int, int, int, int, int);
    if (Wo == 128 && (kBlk != 128 || Ws == 128)) kern = fuse_mask ? warp_fwd_ring_kernel<kBlk, true, 128> : warp_fwd_ring_kernel<kBlk, false, 128>;
    else kern = fuse_mask ? warp_fwd_ring_kernel<kBlk, true, 0> : warp_fwd_ring_kernel<kBlk, false, 0>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem);
    if (e != cudaSuccess) return static_cast<int>(e);
    int n_full;
    const long long n_launch = ring_split<kBlk>(n_items, n_full);
    kern<<<ring_grid<kBlk>(n_launch), Cfg::kThreads, Cfg::kSmem, stream>>>(src, H, out, mask_pooled, C, Hs, Ws, Ho, Wo, blocks_x, n_blocks,
                                                                           static_cast<int>(n_launch), n_full);
    return launch_status();
}

// chunks of nine partial sums per sample that the ring backward writes
template <int kBlk>
inline int ring_bwd_chunks(int Cw, int Ho, int Wo) {
    return Cw * blocks_of(Wo, kBlk) * blocks_of(Ho, kBlk) * RingCfg<kBlk>::kStrips;
}
template <int kBlk>
inline int launch_bwd_ring(const float* src, const float* H, const float* gOut, const float* gMaskPooled, float* partials, int B, int Cw,
                           int Hs, int Ws, int Ho, int Wo, cudaStream_t stream) {
    using Cfg = RingCfg<kBlk>;
    const int blocks_x = blocks_of(Wo, kBlk), n_blocks = blocks_x * blocks_of(Ho, kBlk);
    const long long n_items = static_cast<long long>(B) * Cw * n_blocks;
    void (*kern)(const float*, const float*, const float*, const float*, float*, int, int, int, int, int, int, int, int, int);
    const bool fixed = Wo == 128 && (kBlk != 128 || Ws == 128);   // immediate pitches, see warp_fwd_ring_kernel
    if (gOut && gMaskPooled) kern = fixed ? warp_bwd_ring_kernel<kBlk, true, true, 128> : warp_bwd_ring_kernel<kBlk, true, true, 0>;
    else if (gOut) kern = fixed ? warp_bwd_ring_kernel<kBlk, true, false, 128> : warp_bwd_ring_kernel<kBlk, true, false, 0>;
    else kern = fixed ? warp_bwd_ring_kernel<kBlk, false, true, 128> : warp_bwd_ring_kernel<kBlk, false, true, 0>;
    const int smem = gOut ? Cfg::kBwdSmem : 0;
    if (smem) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return static_cast<int>(e);
    }
    int n_full;
    const long long n_launch = ring_split<kBlk>(n_items, n_full);
    kern<<<ring_grid<kBlk>(n_launch), Cfg::kThreads, smem, stream>>>(src, H, gOut, gMaskPooled, partials, Cw, Hs, Ws, Ho, Wo, blocks_x, n_blocks,
                                                                     static_cast<int>(n_launch), n_full);
    return launch_status();
}

// ---- tile path: host side ----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// NCHW planes whose rows are a whole number of 16-byte units: what a tiled tensor map can describe
inline bool tile_ok(const float* src, int Hs, int Ws, int Ho, int Wo, long long planes, int channels_last) {
    const long long tpp = static_cast<long long>(blocks_of(Wo, kTile)) * blocks_of(Ho, kTile);
    return g_tune[kTuneWarpPath] == 0 && !channels_last && (Ws & 3) == 0 && aligned16(src) && Hs < (1 << 20) && Ws < (1 << 20) && Ho < (1 << 20) &&
           Wo < (1 << 20) && planes < (1ll << 30) && tpp < 65536 && tensor_map_encoder() != nullptr;
}
// resident CTAs per SM the tile kernels are compiled for (register budget 65536 / (128 threads x blocks))
// grid = (tiles per plane, planes low, planes high); ceil(2^32 / tiles_x) for the exact multiply-high division
inline dim3 tile_grid(long long planes, int tpp) {
    const long long py = planes < 32768 ? planes : 32768;
    return dim3(static_cast<unsigned>(tpp), static_cast<unsigned>(py), static_cast<unsigned>((planes + py - 1) / py));
}
inline unsigned tile_magic(int tiles_x) { return tiles_x <= 1 ? 0u : static_cast<unsigned>(((1ull << 32) + tiles_x - 1) / tiles_x); }
inline int make_tile_maps(TileMaps& maps, const float* src, long long planes, int Hs, int Ws) {
    const cuuint64_t gdim[3] = {static_cast<cuuint64_t>(Ws), static_cast<cuuint64_t>(Hs), static_cast<cuuint64_t>(planes)};
    const cuuint64_t gstride[2] = {static_cast<cuuint64_t>(Ws) * 4u, static_cast<cuuint64_t>(Ws) * Hs * 4u};
    const cuuint32_t estride[3] = {1u, 1u, 1u};
    for (int k = 0; k < kTileBoxes; ++k) {
        const cuuint32_t box[3] = {static_cast<cuuint32_t>(tile_box_w(k)), static_cast<cuuint32_t>(tile_box_h(k)), 1u};
        const CUresult r = tensor_map_encoder()(&maps.m[k], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3u, const_cast<float*>(src), gdim, gstride, box,
                                                estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return BH_E_UNSUPPORTED;
    }
    return BH_OK;
}
inline int launch_fwd_tile(const float* src, const float* H, float* out, float* mask_pooled, int B, int C, int Hs, int Ws, int Ho, int Wo,
                           bool fuse_mask, cudaStream_t stream) {
    TileMaps maps;
    const long long planes = static_cast<long long>(B) * C;
    int rc = make_tile_maps(maps, src, planes, Hs, Ws);
    if (rc != BH_OK) return rc;
    const int tiles_x = blocks_of(Wo, kTile), tpp = tiles_x * blocks_of(Ho, kTile);
    const bool four = (g_tune[kTuneWarpVariant] & 1) != 0;
    void (*kern)(const TileMaps, const float*, const float*, float*, float*, int, int, int, int, int, int, unsigned, long long);
    const bool star = Wo == 128 && (Ho % kTile) == 0;      // whole tiles of 128-pixel rows: the north-star geometry
    if (four) kern = star ? (fuse_mask ? warp_fwd_tile_kernel<true, 4, 128, 12> : warp_fwd_tile_kernel<false, 4, 128, 12>)
                          : (fuse_mask ? warp_fwd_tile_kernel<true, 4, 0, 12> : warp_fwd_tile_kernel<false, 4, 0, 12>);
    else kern = star ? (fuse_mask ? warp_fwd_tile_kernel<true, 2, 128, 16> : warp_fwd_tile_kernel<false, 2, 128, 16>)
                     : (fuse_mask ? warp_fwd_tile_kernel<true, 2, 0, 16> : warp_fwd_tile_kernel<false, 2, 0, 16>);
    kern<<<tile_grid(planes, tpp), four ? 128 : 64, kTileSmem, stream>>>(maps, src, H, out, mask_pooled, C, Hs, Ws, Ho, Wo, tiles_x,
                                                                         tile_magic(tiles_x), planes);
    return launch_status();
}
inline int tile_bwd_chunks(int Cw, int Ho, int Wo) { return Cw * blocks_of(Wo, kTile) * blocks_of(Ho, kTile); }
inline int launch_bwd_tile(const float* src, const float* H, const float* gOut, const float* gMaskPooled, float* partials, int B, int Cw,
                           int Hs, int Ws, int Ho, int Wo, cudaStream_t stream) {
    TileMaps maps;
    const long long planes = static_cast<long long>(B) * Cw;
    if (gOut) {
        int rc = make_tile_maps(maps, src, planes, Hs, Ws);
        if (rc != BH_OK) return rc;
    } else {
        memset(&maps, 0, sizeof(maps));
    }
    const int tiles_x = blocks_of(Wo, kTile), tpp = tiles_x * blocks_of(Ho, kTile);
    void (*kern)(const TileMaps, const float*, const float*, const float*, const float*, float*, int, int, int, int, int, int, unsigned,
                 long long);
    const bool four = (g_tune[kTuneWarpVariant] & 1) != 0;
    const bool star = Wo == 128 && (Ho % kTile) == 0;
    if (gOut && gMaskPooled) {
        if (star) kern = four ? warp_bwd_tile_kernel<true, true, 4, 128, 8> : warp_bwd_tile_kernel<true, true, 2, 128, 12>;
        else kern = four ? warp_bwd_tile_kernel<true, true, 4, 0, 8> : warp_bwd_tile_kernel<true, true, 2, 0, 12>;
    } else if (gOut) {
        if (star) kern = four ? warp_bwd_tile_kernel<true, false, 4, 128, 8> : warp_bwd_tile_kernel<true, false, 2, 128, 12>;
        else kern = four ? warp_bwd_tile_kernel<true, false, 4, 0, 8> : warp_bwd_tile_kernel<true, false, 2, 0, 12>;
    } else {
        kern = four ? warp_bwd_tile_kernel<false, true, 4, 0, 8> : warp_bwd_tile_kernel<false, true, 2, 0, 12>;
    }
    kern<<<tile_grid(planes, tpp), four ? 128 : 64, gOut ? kTileSmem : 0, stream>>>(maps, src, H, gOut, gMaskPooled, partials, Cw, Hs, Ws, Ho, Wo,
                                                                                 tiles_x, tile_magic(tiles_x), planes);
    return launch_status();
}
// warp-cooperative channels-last kernels: C/4 a power of two up to 32, or a multiple of 32; 32-bit element offsets per sample
inline bool nhwc_coop_ok(int B, int C, int Hs, int Ws, int Ho, int Wo) {
    const int cq = C / 4;
    const bool shape = cq >= 32 ? (cq % 32) == 0 : (cq & (cq - 1)) == 0;
    return g_tune[kTuneWarpPath] == 0 && B <= 65535 && shape && static_cast<long long>(Hs) * Ws * C < (1ll << 31) &&
           static_cast<long long>(Ho) * Wo * C < (1ll << 31);
}
inline int launch_fwd_nhwc_coop(const float* src, const float* H, float* out, int B, int C, int Hs, int Ws, int Ho, int Wo,
                                cudaStream_t stream) {
    const int cq = C / 4;
    // 8 warps of 32 pixels per CTA; enough CTAs per sample to fill the machine a few times over, never more than the pixels give
    const long long per_sample = (static_cast<long long>(Ho) * Wo + 255) / 256;
    long long gx = (static_cast<long long>(kNumSMs) * 16 + B - 1) / B;
    if (gx > per_sample) gx = per_sample;
    if (gx < 1) gx = 1;
    const dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(B));
    if (cq >= 32) warp_fwd_nhwc_coop_kernel<32><<<grid, 256, 0, stream>>>(src, H, out, C, Hs, Ws, Ho, Wo);
    else if (cq == 16) warp_fwd_nhwc_coop_kernel<16><<<grid, 256, 0, stream>>>(src, H, out, C, Hs, Ws, Ho, Wo);
    else if (cq == 8) warp_fwd_nhwc_coop_kernel<8><<<grid, 256, 0, stream>>>(src, H, out, C, Hs, Ws, Ho, Wo);
    else if (cq == 4) warp_fwd_nhwc_coop_kernel<4><<<grid, 256, 0, stream>>>(src, H, out, C, Hs, Ws, Ho, Wo);
    else if (cq == 2) warp_fwd_nhwc_coop_kernel<2><<<grid, 256, 0, stream>>>(src, H, out, C, Hs, Ws, Ho, Wo);
    else warp_fwd_nhwc_coop_kernel<1><<<grid, 256, 0, stream>>>(src, H, out, C, Hs, Ws, Ho, Wo);
    return launch_status();
}
// the fixed-order sum behind a programmatic dependent launch: its blocks are resident when the producer grid drains
inline int launch_bwd_finish(const float* partials, float* gH, int B, int chunks, cudaStream_t stream) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((B * 9 + 127) / 128);
    cfg.blockDim = dim3(128);
    cfg.stream = stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, warp_bwd_finish_kernel, partials, gH, B, chunks);
    if (e != cudaSuccess) return static_cast<int>(e);
    return launch_status();
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
        const long long n_items64 = static_cast<long long>(B) * C * blocks_of(Wo, 64) * blocks_of(Ho, 64);
        if (tile_ok(src, Hs, Ws, Ho, Wo, static_cast<long long>(B) * C, channels_last)) {
            const bool fuse_mask = mask_pooled && pool == 4 && (Ho % 4) == 0 && (Wo % 4) == 0;
            rc = launch_fwd_tile(src, H, out, mask_pooled, B, C, Hs, Ws, Ho, Wo, fuse_mask, stream);
            mask_done = mask_done || fuse_mask;
        } else if (ring_ok(Hs, Ws, Ho, Wo, channels_last) && n_items64 < (1ll << 31) - 1024) {
            const bool fuse_mask = mask_pooled && pool == 4;
            rc = ring_blk(Hs, Ws) == 128 ? launch_fwd_ring<128>(src, H, out, mask_pooled, B, C, Hs, Ws, Ho, Wo, fuse_mask, stream)
                                         : launch_fwd_ring<64>(src, H, out, mask_pooled, B, C, Hs, Ws, Ho, Wo, fuse_mask, stream);
            mask_done = mask_done || fuse_mask;
        } else if (channels_last && (C % 4) == 0 && nhwc_coop_ok(B, C, Hs, Ws, Ho, Wo)) {
            rc = launch_fwd_nhwc_coop(src, H, out, B, C, Hs, Ws, Ho, Wo, stream);
        } else if (channels_last && (C % 4) == 0) {
            const long long n = static_cast<long long>(B) * Ho * Wo * (C / 4);
            warp_fwd_nhwc_kernel<<<grid_for(n, 256), 256, 0, stream>>>(src, H, out, B, C, Hs, Ws, Ho, Wo);
            rc = launch_status();
        } else {
            const long long n = static_cast<long long>(B) * Ho * Wo;
            warp_fwd_generic_kernel<<<grid_for(n, 256), 256, 0, stream>>>(src, H, out, B, C, Hs, Ws, Ho, Wo,
                                                                          make_layout(C, Hs, Ws, channels_last),
                                                                          make_layout(C, Ho, Wo, channels_last));
            rc = launch_status();
        }
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
    const int Cw = C > 0 ? C : 1;
    const int rc64 = ring_bwd_chunks<64>(Cw, Ho, Wo), rc128 = ring_bwd_chunks<128>(Cw, Ho, Wo);
    const size_t ring = static_cast<size_t>(B) * (rc64 > rc128 ? rc64 : rc128) * 9 * sizeof(float);
    const size_t tile = static_cast<size_t>(B) * tile_bwd_chunks(Cw, Ho, Wo) * 9 * sizeof(float);
    const size_t m = generic > ring ? generic : ring;
    return m > tile ? m : tile;
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
    const int Cw = gOut ? C : 1;   // coverage-only work has one "plane" per sample
    const long long n_items64 = static_cast<long long>(B) * Cw * blocks_of(Wo, 64) * blocks_of(Ho, 64);
    const bool mask4_tiles = gMaskPooled == nullptr || (pool == 4 && (Ho % 4) == 0 && (Wo % 4) == 0);
    if (!gSrc && mask4_tiles && tile_ok(gOut ? src : H, Hs, gOut ? Ws : 4, Ho, Wo, static_cast<long long>(B) * Cw, gOut ? channels_last : 0)) {
        const int chunks = tile_bwd_chunks(Cw, Ho, Wo);
        const size_t need = static_cast<size_t>(B) * chunks * 9 * sizeof(float);
        if (!workspace || workspace_bytes < need) return BH_E_WORKSPACE;
        float* partials = static_cast<float*>(workspace);
        const int rc = launch_bwd_tile(src, H, gOut, gMaskPooled, partials, B, Cw, Hs, Ws, Ho, Wo, stream);
        if (rc != BH_OK) return rc;
        return launch_bwd_finish(partials, gH, B, chunks, stream);
    }
    if (!gSrc && mask4 && ring_ok(Hs, Ws, Ho, Wo, gOut ? channels_last : 0) && n_items64 < (1ll << 31) - 1024) {
        const bool planes = ring_blk(Hs, Ws) == 128;
        const int chunks = planes ? ring_bwd_chunks<128>(Cw, Ho, Wo) : ring_bwd_chunks<64>(Cw, Ho, Wo);
        const size_t need = static_cast<size_t>(B) * chunks * 9 * sizeof(float);
        if (!workspace || workspace_bytes < need) return BH_E_WORKSPACE;
        float* partials = static_cast<float*>(workspace);
        int rc = planes ? launch_bwd_ring<128>(src, H, gOut, gMaskPooled, partials, B, Cw, Hs, Ws, Ho, Wo, stream)
                        : launch_bwd_ring<64>(src, H, gOut, gMaskPooled, partials, B, Cw, Hs, Ws, Ho, Wo, stream);
        if (rc != BH_OK) return rc;
        warp_bwd_finish_kernel<<<(B * 9 + 127) / 128, 128, 0, stream>>>(partials, gH, B, chunks);
        return launch_status();
    }
    const int vec = gOut && channels_last && (C % 4) == 0;
    const int chunks = bwd_chunks(Ho, Wo, gOut ? C : 1, vec);
    const size_t need = static_cast<size_t>(B) * chunks * 9 * sizeof(float);
    if (!workspace || workspace_bytes < need) return BH_E_WORKSPACE;
    float* partials = static_cast<float*>(workspace);
    const Layout ls = make_layout(C > 0 ? C : 1, Hs, Ws, channels_last), lo = make_layout(C > 0 ? C : 1, Ho, Wo, channels_last);
    dim3 grid(chunks, B);
    if (vec && !gSrc && nhwc_coop_ok(B, C, Hs, Ws, Ho, Wo)) {
        const int cq = C / 4;
        if (cq >= 32) warp_bwd_nhwc_coop_kernel<32><<<grid, 256, 0, stream>>>(src, H, gOut, gMaskPooled, partials, C, Hs, Ws, Ho, Wo, pool);
        else if (cq == 16) warp_bwd_nhwc_coop_kernel<16><<<grid, 256, 0, stream>>>(src, H, gOut, gMaskPooled, partials, C, Hs, Ws, Ho, Wo, pool);
        else if (cq == 8) warp_bwd_nhwc_coop_kernel<8><<<grid, 256, 0, stream>>>(src, H, gOut, gMaskPooled, partials, C, Hs, Ws, Ho, Wo, pool);
        else if (cq == 4) warp_bwd_nhwc_coop_kernel<4><<<grid, 256, 0, stream>>>(src, H, gOut, gMaskPooled, partials, C, Hs, Ws, Ho, Wo, pool);
        else if (cq == 2) warp_bwd_nhwc_coop_kernel<2><<<grid, 256, 0, stream>>>(src, H, gOut, gMaskPooled, partials, C, Hs, Ws, Ho, Wo, pool);
        else warp_bwd_nhwc_coop_kernel<1><<<grid, 256, 0, stream>>>(src, H, gOut, gMaskPooled, partials, C, Hs, Ws, Ho, Wo, pool);
    } else if (vec)
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
