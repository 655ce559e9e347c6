// K6 on the tensor cores: the per-pixel 16 -> 128 -> 2 network of the Zeng backbone's field head (fieldhead.cu states the
// algebra and the reference lines) is GEMM shaped -- [pixels x 16] x [16 x 128] -- so its contractions run as
// mma.sync.m16n8k8 TF32 tensor-core instructions with the operands split into a TF32 head and a float32 remainder
// (x = hi + lo, three products per instruction group: lo*hi + hi*lo + hi*hi), which keeps the float32 accuracy of the
// scalar kernels (measured against them in tests/test_gpu_zzz_field_head.py) at a fifth of their instruction count.
// K = 16 is far too short for a tcgen05 pipeline (one 128 x 128 x 16 instruction per 128 pixels, then 128 columns of TMEM
// read back per pixel row for the ReLU/projection epilogue): the work per byte is the epilogue, not the contraction, and
// the register-resident mma.sync accumulators are exactly where the epilogue wants its operands.
//
//   fieldhead_fwd_mma   pixels on the M side: pre[16 px, 8 hid] tiles, ReLU and the 128 -> 2 projection on the accumulator
//                       registers, quad reduction, planar field out
//   fieldhead_gx_mma    the same recomputation, gh = relu'(pre) * (W2^T g) on the accumulators, which ARE the A fragments
//                       of gx[16 px, 16] += gh[16 px, 8 hid] W1[8 hid, 16] (no data movement between the two GEMMs)
//   moments_mma         X^T X over the pixels (BatchNorm's batch statistics from the input moments), three-product form
//   fieldhead_gw_mma    hidden units on the M side: pre^T[16 hid, 8 px] tiles, whose accumulators are the A fragments of
//                       gW1[16 hid, 16] += gh^T x and gW2^T[16 hid, 2] += h^T g; per-CTA partial sums (fixed order)
// Fragment layouts (PTX ISA, mma.m16n8k8 .tf32; g = lane >> 2, t = lane & 3):
//   A 16x8 row:  a0 (g, t)  a1 (g+8, t)  a2 (g, t+4)  a3 (g+8, t+4)
//   B 8x8 col:   b0 (k = t, n = g)       b1 (k = t+4, n = g)
//   C 16x8:      c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)
// The K index of a contraction may be permuted freely as long as A and B agree; the kernels use that to take fragments
// straight from 128-bit loads and from accumulator registers.
#include "bh_common.cuh"

namespace bh {

constexpr int kMmaThreads = 128;
constexpr int kMmaWarps = kMmaThreads / 32;
constexpr int kCin = 16, kHid = 128;
constexpr int kXPitch = 20;   // floats per pixel row of the staged input tile: conflict-free 32-bit fragment loads in the gw kernel

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// Two precisions (template parameter kExact of every kernel):
//   kExact = true   x = hi + lo exactly; hi has 10 mantissa bits (a TF32 number), the tensor core ignores the low 13 bits of
//                   lo; three products per contraction step (lo*hi + hi*lo + hi*hi): float32-faithful
//   kExact = false  x rounded to the nearest TF32 number (cvt.rna, unbiased), one product: the arithmetic cuDNN / cuBLAS use
//                   for the convolutions this kernel replaces when torch.backends.cudnn.allow_tf32 is on (torch's default)
template <bool kExact>
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    if (kExact) {
        hi = __float_as_uint(x) & 0xffffe000u;
        lo = __float_as_uint(x - __uint_as_float(hi));
    } else {
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
        lo = 0u;
    }
}
// c += A B with both operands as (hi, lo): the two small products first
template <bool kExact>
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4], uint32_t b0hi, uint32_t b1hi,
                                     uint32_t b0lo, uint32_t b1lo) {
    if (kExact) {
        mma_tf32(c, alo, b0hi, b1hi);
        mma_tf32(c, ahi, b0lo, b1lo);
    }
    mma_tf32(c, ahi, b0hi, b1hi);
}
template <bool kExact>
__device__ __forceinline__ float4 pack_split(float w0, float w1) {
    uint32_t h0, l0, h1, l1;
    split_tf32<kExact>(w0, h0, l0);
    split_tf32<kExact>(w1, h1, l1);
    return make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(l0), __uint_as_float(l1));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

// B fragments of W1^T for the pixel-major contraction pre[px, hid] = sum_ch x[px, ch] W1[hid, ch]:
// k-step s, slot t <-> channel 4t + 2s, slot t+4 <-> channel 4t + 2s + 1 (so that lane (g,t) takes its A fragments from the
// float4 x[px g][4t .. 4t+3] it loaded); table entry [hid tile j][s][lane] = (b0 hi, b1 hi, b0 lo, b1 lo)
template <bool kExact>
__device__ __forceinline__ void build_w1_b_table(float4* tab, const float* __restrict__ W1) {
    for (int q = threadIdx.x; q < 16 * 2 * 32; q += blockDim.x) {
        const int j = q >> 6, s = (q >> 5) & 1, l = q & 31, g = l >> 2, t = l & 3;
        const float* row = W1 + (8 * j + g) * kCin + 4 * t + 2 * s;
        tab[q] = pack_split<kExact>(row[0], row[1]);
    }
}

// the A fragments (hi / lo) of MT pixel tiles from the lanes' own 128-bit loads; rows g and g+8 of tile m
template <int MT, bool kExact>
__device__ __forceinline__ void load_x_fragments(const float* __restrict__ x, long long p0, int g, int t, uint32_t (&ahi)[MT][2][4],
                                                 uint32_t (&alo)[MT][2][4]) {
#pragma unroll
    for (int m = 0; m < MT; ++m) {
        const float4 xa = ldg_stream(reinterpret_cast<const float4*>(x + (p0 + 16 * m + g) * kCin) + t);
        const float4 xb = ldg_stream(reinterpret_cast<const float4*>(x + (p0 + 16 * m + g + 8) * kCin) + t);
        split_tf32<kExact>(xa.x, ahi[m][0][0], alo[m][0][0]);   // k-step 0: a0 = (g, ch 4t), a2 = (g, ch 4t+1)
        split_tf32<kExact>(xb.x, ahi[m][0][1], alo[m][0][1]);   //           a1 = (g+8, ch 4t), a3 = (g+8, ch 4t+1)
        split_tf32<kExact>(xa.y, ahi[m][0][2], alo[m][0][2]);
        split_tf32<kExact>(xb.y, ahi[m][0][3], alo[m][0][3]);
        split_tf32<kExact>(xa.z, ahi[m][1][0], alo[m][1][0]);   // k-step 1: channels 4t+2, 4t+3
        split_tf32<kExact>(xb.z, ahi[m][1][1], alo[m][1][1]);
        split_tf32<kExact>(xa.w, ahi[m][1][2], alo[m][1][2]);
        split_tf32<kExact>(xb.w, ahi[m][1][3], alo[m][1][3]);
    }
}

// ---- forward -------------------------------------------------------------------------------------------------------
// a warp owns groups of 16 * MT consecutive pixels (inside one sample: HW % (16 MT) == 0)
template <int MT, bool kExact>
__global__ void __launch_bounds__(kMmaThreads) fieldhead_fwd_mma_kernel(const float* __restrict__ x, const float* __restrict__ W1,
                                                                       const float* __restrict__ b1, const float* __restrict__ W2,
                                                                       const float* __restrict__ b2, float* __restrict__ out,
                                                                       long long n_pix, int HW) {
    __shared__ float4 sWB[16 * 2 * 32];
    __shared__ __align__(8) float sb1[kHid];
    __shared__ __align__(8) float sW2[2 * kHid];
    build_w1_b_table<kExact>(sWB, W1);
    for (int q = threadIdx.x; q < kHid; q += blockDim.x) sb1[q] = b1[q];
    for (int q = threadIdx.x; q < 2 * kHid; q += blockDim.x) sW2[q] = W2[q];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const float c0 = b2[0], c1 = b2[1];
    const long long n_groups = n_pix / (16 * MT);
    for (long long grp = static_cast<long long>(blockIdx.x) * kMmaWarps + warp; grp < n_groups; grp += static_cast<long long>(gridDim.x) * kMmaWarps) {
        const long long p0 = grp * (16 * MT);
        uint32_t ahi[MT][2][4], alo[MT][2][4];
        load_x_fragments<MT, kExact>(x, p0, g, t, ahi, alo);
        float o[MT][4];   // (row g, out 0) (row g+8, out 0) (row g, out 1) (row g+8, out 1): this lane's columns only
#pragma unroll
        for (int m = 0; m < MT; ++m) o[m][0] = o[m][1] = o[m][2] = o[m][3] = 0.0f;
#pragma unroll 2
        for (int j = 0; j < 16; ++j) {
            const float4 w0 = sWB[(j * 2 + 0) * 32 + lane], w1 = sWB[(j * 2 + 1) * 32 + lane];
            const float2 bj = *reinterpret_cast<const float2*>(sb1 + 8 * j + 2 * t);
            const float2 wa = *reinterpret_cast<const float2*>(sW2 + 8 * j + 2 * t);
            const float2 wb = *reinterpret_cast<const float2*>(sW2 + kHid + 8 * j + 2 * t);
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                float c[4] = {bj.x, bj.y, bj.x, bj.y};
                mma3<kExact>(c, ahi[m][0], alo[m][0], __float_as_uint(w0.x), __float_as_uint(w0.y), __float_as_uint(w0.z), __float_as_uint(w0.w));
                mma3<kExact>(c, ahi[m][1], alo[m][1], __float_as_uint(w1.x), __float_as_uint(w1.y), __float_as_uint(w1.z), __float_as_uint(w1.w));
                const float h0 = fmaxf(c[0], 0.0f), h1 = fmaxf(c[1], 0.0f), h2 = fmaxf(c[2], 0.0f), h3 = fmaxf(c[3], 0.0f);
                o[m][0] = fmaf(h0, wa.x, fmaf(h1, wa.y, o[m][0]));
                o[m][1] = fmaf(h2, wa.x, fmaf(h3, wa.y, o[m][1]));
                o[m][2] = fmaf(h0, wb.x, fmaf(h1, wb.y, o[m][2]));
                o[m][3] = fmaf(h2, wb.x, fmaf(h3, wb.y, o[m][3]));
            }
        }
        const long long b = p0 / HW, s0 = p0 - b * HW;
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            const float v0 = quad_sum(o[m][0]), v1 = quad_sum(o[m][1]), v2 = quad_sum(o[m][2]), v3 = quad_sum(o[m][3]);
            // lane t of the quad stores one of the four results of rows g / g+8
            const float v = (t == 0) ? v0 + c0 : (t == 1) ? v1 + c0 : (t == 2) ? v2 + c1 : v3 + c1;
            const int row = 16 * m + g + ((t & 1) ? 8 : 0);
            out[(2 * b + (t >> 1)) * HW + s0 + row] = v;
        }
    }
}

// ---- backward, part 1: d/dx ------------------------------------------------------------------------------------------
// B fragments of W1 for gx[px, ch] = sum_hid gh[px, hid] W1[hid, ch], K = the 8 hidden units of tile j with
// slot t <-> hidden 8j + 2t, slot t+4 <-> hidden 8j + 2t + 1 (the accumulator columns of the lane): [j][ch tile n][lane]
template <bool kExact>
__device__ __forceinline__ void build_w1_gx_table(float4* tab, const float* __restrict__ W1) {
    for (int q = threadIdx.x; q < 16 * 2 * 32; q += blockDim.x) {
        const int j = q >> 6, n = (q >> 5) & 1, l = q & 31, g = l >> 2, t = l & 3;
        tab[q] = pack_split<kExact>(W1[(8 * j + 2 * t) * kCin + 8 * n + g], W1[(8 * j + 2 * t + 1) * kCin + 8 * n + g]);
    }
}

template <int MT, bool kExact>
__global__ void __launch_bounds__(kMmaThreads) fieldhead_gx_mma_kernel(const float* __restrict__ x, const float* __restrict__ W1,
                                                                      const float* __restrict__ b1, const float* __restrict__ W2,
                                                                      const float* __restrict__ gOut, float* __restrict__ gx,
                                                                      long long n_pix, int HW) {
    __shared__ float4 sWB[16 * 2 * 32];
    __shared__ float4 sWG[16 * 2 * 32];
    __shared__ __align__(8) float sb1[kHid];
    __shared__ __align__(8) float sW2[2 * kHid];
    build_w1_b_table<kExact>(sWB, W1);
    build_w1_gx_table<kExact>(sWG, W1);
    for (int q = threadIdx.x; q < kHid; q += blockDim.x) sb1[q] = b1[q];
    for (int q = threadIdx.x; q < 2 * kHid; q += blockDim.x) sW2[q] = W2[q];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const long long n_groups = n_pix / (16 * MT);
    for (long long grp = static_cast<long long>(blockIdx.x) * kMmaWarps + warp; grp < n_groups; grp += static_cast<long long>(gridDim.x) * kMmaWarps) {
        const long long p0 = grp * (16 * MT);
        const long long b = p0 / HW, s0 = p0 - b * HW;
        uint32_t ahi[MT][2][4], alo[MT][2][4];
        load_x_fragments<MT, kExact>(x, p0, g, t, ahi, alo);
        float ga[MT][2], gb[MT][2];   // upstream gradients of rows g, g+8 (both outputs)
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            ga[m][0] = __ldg(gOut + (2 * b) * HW + s0 + 16 * m + g);
            ga[m][1] = __ldg(gOut + (2 * b) * HW + s0 + 16 * m + g + 8);
            gb[m][0] = __ldg(gOut + (2 * b + 1) * HW + s0 + 16 * m + g);
            gb[m][1] = __ldg(gOut + (2 * b + 1) * HW + s0 + 16 * m + g + 8);
        }
        float acc[MT][2][4];
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int n = 0; n < 2; ++n) acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.0f;
#pragma unroll 2
        for (int j = 0; j < 16; ++j) {
            const float4 w0 = sWB[(j * 2 + 0) * 32 + lane], w1 = sWB[(j * 2 + 1) * 32 + lane];
            const float4 v0 = sWG[(j * 2 + 0) * 32 + lane], v1 = sWG[(j * 2 + 1) * 32 + lane];
            const float2 bj = *reinterpret_cast<const float2*>(sb1 + 8 * j + 2 * t);
            const float2 wa = *reinterpret_cast<const float2*>(sW2 + 8 * j + 2 * t);
            const float2 wb = *reinterpret_cast<const float2*>(sW2 + kHid + 8 * j + 2 * t);
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                float c[4] = {bj.x, bj.y, bj.x, bj.y};
                mma3<kExact>(c, ahi[m][0], alo[m][0], __float_as_uint(w0.x), __float_as_uint(w0.y), __float_as_uint(w0.z), __float_as_uint(w0.w));
                mma3<kExact>(c, ahi[m][1], alo[m][1], __float_as_uint(w1.x), __float_as_uint(w1.y), __float_as_uint(w1.z), __float_as_uint(w1.w));
                // gh on the accumulator layout: rows g (c0, c1) / g+8 (c2, c3), hidden columns 8j+2t (c0, c2) / 8j+2t+1 (c1, c3)
                const float gh0 = c[0] > 0.0f ? fmaf(wa.x, ga[m][0], wb.x * gb[m][0]) : 0.0f;
                const float gh1 = c[1] > 0.0f ? fmaf(wa.y, ga[m][0], wb.y * gb[m][0]) : 0.0f;
                const float gh2 = c[2] > 0.0f ? fmaf(wa.x, ga[m][1], wb.x * gb[m][1]) : 0.0f;
                const float gh3 = c[3] > 0.0f ? fmaf(wa.y, ga[m][1], wb.y * gb[m][1]) : 0.0f;
                // as A fragments of the second GEMM: a0 (g, slot t) = gh0, a1 (g+8, slot t) = gh2, a2 (g, slot t+4) = gh1, a3 = gh3
                uint32_t hi[4], lo[4];
                split_tf32<kExact>(gh0, hi[0], lo[0]);
                split_tf32<kExact>(gh2, hi[1], lo[1]);
                split_tf32<kExact>(gh1, hi[2], lo[2]);
                split_tf32<kExact>(gh3, hi[3], lo[3]);
                mma3<kExact>(acc[m][0], hi, lo, __float_as_uint(v0.x), __float_as_uint(v0.y), __float_as_uint(v0.z), __float_as_uint(v0.w));
                mma3<kExact>(acc[m][1], hi, lo, __float_as_uint(v1.x), __float_as_uint(v1.y), __float_as_uint(v1.z), __float_as_uint(v1.w));
            }
        }
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int n = 0; n < 2; ++n) {
                float* r0 = gx + (p0 + 16 * m + g) * kCin + 8 * n + 2 * t;
                *reinterpret_cast<float2*>(r0) = make_float2(acc[m][n][0], acc[m][n][1]);
                *reinterpret_cast<float2*>(r0 + 8 * kCin) = make_float2(acc[m][n][2], acc[m][n][3]);
            }
    }
}

// ---- backward, part 2: the weight gradients ----------------------------------------------------------------------------
// Hidden units on the M side: warp w owns hidden rows 32w .. 32w+31 (two 16-row tiles), the CTA walks 32-pixel tiles staged
// in shared memory.  partials[cta] = { gW1 [HID*CIN] | gb1 [HID] | gW2 [2*HID] | gb2 [2] } as in fieldhead.cu.
template <bool kExact>
__global__ void __launch_bounds__(kMmaThreads, kExact ? 4 : 5) fieldhead_gw_mma_kernel(const float* __restrict__ x, const float* __restrict__ W1,
                                                                      const float* __restrict__ b1, const float* __restrict__ W2,
                                                                      const float* __restrict__ gOut, float* __restrict__ partials,
                                                                      long long n_pix, int HW) {
    __shared__ __align__(16) float sX[32 * kXPitch];   // [pixel][channel], pitch 20 floats
    __shared__ __align__(8) float sG[2 * 32];          // [out][pixel]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    // this warp's A fragments of W1 (pre^T = W1 x^T; k-step s: slot t <-> channel 4t + 2s, slot t+4 <-> channel 4t + 2s + 1, the
    // mapping of the pixel-major kernels, so that the recomputed pre-activations are the same sums of the same products and
    // the ReLU gates of the three kernels agree), bias and projection rows
    uint32_t whi[2][2][4], wlo[2][2][4];
    float bia[2][2], w2a[2][2], w2b[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int r0 = 32 * warp + 16 * i + g, r1 = r0 + 8;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            split_tf32<kExact>(__ldg(W1 + r0 * kCin + 4 * t + 2 * s), whi[i][s][0], wlo[i][s][0]);
            split_tf32<kExact>(__ldg(W1 + r1 * kCin + 4 * t + 2 * s), whi[i][s][1], wlo[i][s][1]);
            split_tf32<kExact>(__ldg(W1 + r0 * kCin + 4 * t + 2 * s + 1), whi[i][s][2], wlo[i][s][2]);
            split_tf32<kExact>(__ldg(W1 + r1 * kCin + 4 * t + 2 * s + 1), whi[i][s][3], wlo[i][s][3]);
        }
        bia[i][0] = __ldg(b1 + r0); bia[i][1] = __ldg(b1 + r1);
        w2a[i][0] = __ldg(W2 + r0); w2a[i][1] = __ldg(W2 + r1);
        w2b[i][0] = __ldg(W2 + kHid + r0); w2b[i][1] = __ldg(W2 + kHid + r1);
    }
    float aW1[2][2][4], aW2[2][4], ab1[2][2], ab2 = 0.0f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        ab1[i][0] = ab1[i][1] = 0.0f;
#pragma unroll
        for (int e = 0; e < 4; ++e) aW2[i][e] = aW1[i][0][e] = aW1[i][1][e] = 0.0f;
    }
    const long long n_tiles = n_pix / 32;
    // the next tile's share of this thread (one float4 of x, one upstream gradient for threads 0..63) is fetched into registers
    // while the current tile is consumed: the global-load latency hides behind a whole tile of tensor-core work
    float4 nx = make_float4(0.f, 0.f, 0.f, 0.f);
    float ng = 0.0f;
    auto fetch = [&](long long tile) {
        const long long p0 = tile * 32;
        const long long b = p0 / HW, s0 = p0 - b * HW;
        nx = ldg_stream(reinterpret_cast<const float4*>(x + p0 * kCin) + tid);      // pixel tid / 4, channels 4 (tid % 4) ..
        if (tid < 64) ng = __ldg(gOut + (2 * b + (tid >> 5)) * HW + s0 + (tid & 31));
    };
    if (blockIdx.x < n_tiles) fetch(blockIdx.x);
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncthreads();   // the previous tile is fully consumed
        *reinterpret_cast<float4*>(sX + (tid >> 2) * kXPitch + 4 * (tid & 3)) = nx;
        if (tid < 64) {
            sG[tid] = ng;
            ab2 += ng;
        }
        if (tile + gridDim.x < n_tiles) fetch(tile + gridDim.x);
        __syncthreads();
#pragma unroll 1
        for (int n = 0; n < 4; ++n) {   // 8-pixel column tiles
            // B fragments of x^T for the recomputation: b0 (slot t, px g) = x[px 8n+g][4t + 2s], b1 (slot t+4) = x[..][4t + 2s + 1]
            uint32_t xb_hi[2][2], xb_lo[2][2];
            {
                const float4 xv = *reinterpret_cast<const float4*>(sX + (8 * n + g) * kXPitch + 4 * t);
                split_tf32<kExact>(xv.x, xb_hi[0][0], xb_lo[0][0]);
                split_tf32<kExact>(xv.y, xb_hi[0][1], xb_lo[0][1]);
                split_tf32<kExact>(xv.z, xb_hi[1][0], xb_lo[1][0]);
                split_tf32<kExact>(xv.w, xb_hi[1][1], xb_lo[1][1]);
            }
            // B fragments of x for gW1 (K = pixels: slot t <-> px 8n+2t, slot t+4 <-> px 8n+2t+1; N = channels 8m + g)
            uint32_t xw_hi[2][2], xw_lo[2][2];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                split_tf32<kExact>(sX[(8 * n + 2 * t) * kXPitch + 8 * m + g], xw_hi[m][0], xw_lo[m][0]);
                split_tf32<kExact>(sX[(8 * n + 2 * t + 1) * kXPitch + 8 * m + g], xw_hi[m][1], xw_lo[m][1]);
            }
            // upstream gradients of the lane's two pixel columns, and as B fragments of gW2^T[hid, out] (N = 8, two used)
            const float2 g0 = *reinterpret_cast<const float2*>(sG + 8 * n + 2 * t);
            const float2 g1 = *reinterpret_cast<const float2*>(sG + 32 + 8 * n + 2 * t);
            uint32_t gb_hi[2], gb_lo[2];
            {
                const float2 gg = (g < 2) ? *reinterpret_cast<const float2*>(sG + 32 * g + 8 * n + 2 * t) : make_float2(0.0f, 0.0f);
                split_tf32<kExact>(gg.x, gb_hi[0], gb_lo[0]);
                split_tf32<kExact>(gg.y, gb_hi[1], gb_lo[1]);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                float c[4] = {bia[i][0], bia[i][0], bia[i][1], bia[i][1]};
                mma3<kExact>(c, whi[i][0], wlo[i][0], xb_hi[0][0], xb_hi[0][1], xb_lo[0][0], xb_lo[0][1]);
                mma3<kExact>(c, whi[i][1], wlo[i][1], xb_hi[1][0], xb_hi[1][1], xb_lo[1][0], xb_lo[1][1]);
                // rows: hidden r0 (c0, c1) / r1 (c2, c3); columns: pixels 8n+2t (c0, c2) / 8n+2t+1 (c1, c3)
                const float h0 = fmaxf(c[0], 0.0f), h1 = fmaxf(c[1], 0.0f), h2 = fmaxf(c[2], 0.0f), h3 = fmaxf(c[3], 0.0f);
                const float gh0 = c[0] > 0.0f ? fmaf(w2a[i][0], g0.x, w2b[i][0] * g1.x) : 0.0f;
                const float gh1 = c[1] > 0.0f ? fmaf(w2a[i][0], g0.y, w2b[i][0] * g1.y) : 0.0f;
                const float gh2 = c[2] > 0.0f ? fmaf(w2a[i][1], g0.x, w2b[i][1] * g1.x) : 0.0f;
                const float gh3 = c[3] > 0.0f ? fmaf(w2a[i][1], g0.y, w2b[i][1] * g1.y) : 0.0f;
                ab1[i][0] += gh0 + gh1;
                ab1[i][1] += gh2 + gh3;
                uint32_t hi[4], lo[4];
                split_tf32<kExact>(gh0, hi[0], lo[0]);
                split_tf32<kExact>(gh2, hi[1], lo[1]);
                split_tf32<kExact>(gh1, hi[2], lo[2]);
                split_tf32<kExact>(gh3, hi[3], lo[3]);
                mma3<kExact>(aW1[i][0], hi, lo, xw_hi[0][0], xw_hi[0][1], xw_lo[0][0], xw_lo[0][1]);
                mma3<kExact>(aW1[i][1], hi, lo, xw_hi[1][0], xw_hi[1][1], xw_lo[1][0], xw_lo[1][1]);
                split_tf32<kExact>(h0, hi[0], lo[0]);
                split_tf32<kExact>(h2, hi[1], lo[1]);
                split_tf32<kExact>(h1, hi[2], lo[2]);
                split_tf32<kExact>(h3, hi[3], lo[3]);
                mma3<kExact>(aW2[i], hi, lo, gb_hi[0], gb_hi[1], gb_lo[0], gb_lo[1]);
            }
        }
    }
    constexpr int kPartial = kHid * kCin + kHid + 2 * kHid + 2;
    float* dst = partials + static_cast<long long>(blockIdx.x) * kPartial;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int r0 = 32 * warp + 16 * i + g, r1 = r0 + 8;
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            *reinterpret_cast<float2*>(dst + r0 * kCin + 8 * m + 2 * t) = make_float2(aW1[i][m][0], aW1[i][m][1]);
            *reinterpret_cast<float2*>(dst + r1 * kCin + 8 * m + 2 * t) = make_float2(aW1[i][m][2], aW1[i][m][3]);
        }
        const float s0v = quad_sum(ab1[i][0]), s1v = quad_sum(ab1[i][1]);
        if (t == 0) {
            dst[kHid * kCin + r0] = s0v;
            dst[kHid * kCin + r1] = s1v;
            // gW2^T accumulator columns 2t, 2t+1 = outputs 0, 1 for t == 0
            dst[kHid * kCin + kHid + r0] = aW2[i][0];
            dst[kHid * kCin + 2 * kHid + r0] = aW2[i][1];
            dst[kHid * kCin + kHid + r1] = aW2[i][2];
            dst[kHid * kCin + 2 * kHid + r1] = aW2[i][3];
        }
    }
    // gb2: threads 0..31 hold output 0, 32..63 output 1
    ab2 = warp_sum(ab2);
    if (warp < 2 && lane == 0) dst[kHid * kCin + 3 * kHid + warp] = ab2;
}

// ---- input moments ---------------------------------------------------------------------------------------------------
// sum_p x_i and sum_p x_i x_k of the [N,16] input: X^T X with the pixels as the contraction index.  One 8-pixel block is
// one k-step; the four values a lane loads -- x[p+t][g], x[p+t][g+8], x[p+t+4][g], x[p+t+4][g+8] -- are at the same time its
// A fragment (channels on the rows) and its B fragments for the two 8-channel column tiles, so the kernel is four scalar
// loads, four splits and six tensor-core instructions per 8 pixels.  Always the float32-faithful three-product form: the
// second moments feed a variance (E[xx] - mm) and must not carry TF32 rounding.  float32 accumulation over 128 pixels, then
// float64 (as the scalar kernel: partials[cta][16 + 256] double).
__global__ void __launch_bounds__(kMmaThreads) moments_mma_kernel(const float* __restrict__ x, double* __restrict__ partials,
                                                                 long long n_pix) {
    constexpr int kStats = kCin + kCin * kCin;
    __shared__ double sAcc[kMmaWarps][kStats];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    double d2[2][4], d1[2] = {0.0, 0.0};
#pragma unroll
    for (int n = 0; n < 2; ++n) d2[n][0] = d2[n][1] = d2[n][2] = d2[n][3] = 0.0;
    const long long n_blocks = n_pix / 8;
    const long long stride = static_cast<long long>(gridDim.x) * kMmaWarps * 16;
    // a warp takes runs of 16 consecutive 8-pixel blocks (128 pixels = 8 KB contiguous)
    for (long long b0 = (static_cast<long long>(blockIdx.x) * kMmaWarps + warp) * 16; b0 < n_blocks; b0 += stride) {
        // every 8-pixel block starts from a zero accumulator and is added to the running float32 sums with a round-to-nearest
        // FADD: accumulating 48 instructions in the tensor core's own adder (which truncates) biased the sums by -1.7e-6
        float c[2][4], s0 = 0.0f, s1 = 0.0f;
#pragma unroll
        for (int n = 0; n < 2; ++n) c[n][0] = c[n][1] = c[n][2] = c[n][3] = 0.0f;
        const int nb = static_cast<int>(n_blocks - b0 < 16 ? n_blocks - b0 : 16);
#pragma unroll 4
        for (int k = 0; k < nb; ++k) {
            const float* p = x + ((b0 + k) * 8 + t) * kCin + g;
            const float v0 = __ldg(p), v1 = __ldg(p + 8), v2 = __ldg(p + 4 * kCin), v3 = __ldg(p + 4 * kCin + 8);
            s0 += v0 + v2;
            s1 += v1 + v3;
            uint32_t hi[4], lo[4];
            split_tf32<true>(v0, hi[0], lo[0]);
            split_tf32<true>(v1, hi[1], lo[1]);
            split_tf32<true>(v2, hi[2], lo[2]);
            split_tf32<true>(v3, hi[3], lo[3]);
            float c0[4] = {0.0f, 0.0f, 0.0f, 0.0f}, c1[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            mma3<true>(c0, hi, lo, hi[0], hi[2], lo[0], lo[2]);   // columns = channels 0..7: b0 = x[p+t][g], b1 = x[p+t+4][g]
            mma3<true>(c1, hi, lo, hi[1], hi[3], lo[1], lo[3]);   // columns = channels 8..15
#pragma unroll
            for (int e = 0; e < 4; ++e) { c[0][e] += c0[e]; c[1][e] += c1[e]; }
        }
#pragma unroll
        for (int n = 0; n < 2; ++n)
#pragma unroll
            for (int e = 0; e < 4; ++e) d2[n][e] += static_cast<double>(c[n][e]);
        d1[0] += static_cast<double>(s0);
        d1[1] += static_cast<double>(s1);
    }
    // sums over the four lanes that hold the same channel; c layout: rows g / g+8, columns 8n + 2t / 8n + 2t + 1
    for (int o = 1; o < 4; o <<= 1) {
        d1[0] += __shfl_xor_sync(0xffffffffu, d1[0], o);
        d1[1] += __shfl_xor_sync(0xffffffffu, d1[1], o);
    }
    if (t == 0) {
        sAcc[warp][g] = d1[0];
        sAcc[warp][g + 8] = d1[1];
    }
#pragma unroll
    for (int n = 0; n < 2; ++n) {
        sAcc[warp][kCin + g * kCin + 8 * n + 2 * t] = d2[n][0];
        sAcc[warp][kCin + g * kCin + 8 * n + 2 * t + 1] = d2[n][1];
        sAcc[warp][kCin + (g + 8) * kCin + 8 * n + 2 * t] = d2[n][2];
        sAcc[warp][kCin + (g + 8) * kCin + 8 * n + 2 * t + 1] = d2[n][3];
    }
    __syncthreads();
    for (int q = threadIdx.x; q < kStats; q += kMmaThreads) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kMmaWarps; ++w) v += sAcc[w][q];
        partials[static_cast<long long>(blockIdx.x) * kStats + q] = v;
    }
}

int launch_fieldhead_moments_mma(const float* x, double* partials, long long n_pix, int grid, cudaStream_t stream) {
    moments_mma_kernel<<<grid, kMmaThreads, 0, stream>>>(x, partials, n_pix);
    return launch_status();
}

inline int mma_grid(long long n_items, int per_sm) {
    const long long cap = static_cast<long long>(kNumSMs) * per_sm;
    return static_cast<int>(n_items < 1 ? 1 : (n_items < cap ? n_items : cap));
}

// geometry the tensor-core kernels cover: 16 -> 128 -> 2, whole 32-pixel groups inside a sample
bool fieldhead_mma_ok(int HW, int cin, int hid) { return cin == kCin && hid == kHid && HW > 0 && (HW % 32) == 0; }
int fieldhead_gw_mma_grid(long long n_pix) { return mma_grid(n_pix / 32, 5); }   // == bh_fieldhead_grid(1, n_pix): rows of `partials`

int launch_fieldhead_fwd_mma(const float* x, const float* W1, const float* b1, const float* W2, const float* b2, float* out, long long n_pix,
                             int HW, int tf32, cudaStream_t stream) {
    const int grid = mma_grid((n_pix / 32 + kMmaWarps - 1) / kMmaWarps, 8);   // 64 registers: eight 128-thread CTAs per SM
    if (tf32) fieldhead_fwd_mma_kernel<2, false><<<grid, kMmaThreads, 0, stream>>>(x, W1, b1, W2, b2, out, n_pix, HW);
    else fieldhead_fwd_mma_kernel<2, true><<<grid, kMmaThreads, 0, stream>>>(x, W1, b1, W2, b2, out, n_pix, HW);
    return launch_status();
}
int launch_fieldhead_bwd_mma(const float* x, const float* W1, const float* b1, const float* W2, const float* gOut, float* gx, float* partials,
                             long long n_pix, int HW, int tf32, cudaStream_t stream) {
    const int grid = mma_grid((n_pix / 32 + kMmaWarps - 1) / kMmaWarps, 6);   // 77-96 registers, 33 KB of tables: six CTAs per SM
    if (tf32) fieldhead_gx_mma_kernel<2, false><<<grid, kMmaThreads, 0, stream>>>(x, W1, b1, W2, gOut, gx, n_pix, HW);
    else fieldhead_gx_mma_kernel<2, true><<<grid, kMmaThreads, 0, stream>>>(x, W1, b1, W2, gOut, gx, n_pix, HW);
    int rc = launch_status();
    if (rc != BH_OK) return rc;
    if (tf32) fieldhead_gw_mma_kernel<false><<<fieldhead_gw_mma_grid(n_pix), kMmaThreads, 0, stream>>>(x, W1, b1, W2, gOut, partials, n_pix, HW);
    else fieldhead_gw_mma_kernel<true><<<fieldhead_gw_mma_grid(n_pix), kMmaThreads, 0, stream>>>(x, W1, b1, W2, gOut, partials, n_pix, HW);
    return launch_status();
}

}  // namespace bh
