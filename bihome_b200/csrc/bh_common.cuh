// Shared device/host helpers for libbihome_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bihome_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "bihome_b200 kernels are written for sm_100a (B200) only"
#endif

namespace bh {

// SM count of the current device (B200: 148 = 2 dies x 74); persistent grids are sized in multiples of it.  Queried once
// per device ordinal; 148 when no device is visible (the build container), so that grid arithmetic stays testable there.
inline int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
        cudaGetLastError();
        return 148;
    }
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
            cudaGetLastError();
            n = 148;
        }
        cached[dev] = n;
    }
    return cached[dev];
}
#define kNumSMs (::bh::num_sms())

// Tuning switches of tools/microbench.py (bh_tune_set, include/bihome_b200.h): process-global, never touched by the product
// path -- every entry point picks its kernel from its arguments alone while these hold their default 0.
enum TuneKey { kTuneWarpPath = 0, kTuneLossVariant = 1, kTuneLossCluster = 2, kTuneWarpVariant = 3, kTuneFieldheadVariant = 4, kTuneCount = 8 };
extern int g_tune[kTuneCount];

// Every kernel launch in the library goes through BH_LAUNCH_CHECK so the launch counter that
// bench.py reports (gpu_launches) cannot drift from what really ran.
extern unsigned long long g_launch_count;

inline int launch_status() {
    ++g_launch_count;
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? BH_OK : static_cast<int>(e);
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- warp / block reductions -------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum of N values per thread; result valid in every thread.  `scratch` holds
// N * (blockDim.x/32) elements.  Fixed summation order => bit-reproducible.
template <int N, typename T>
__device__ __forceinline__ void block_sum(T (&v)[N], T* scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();  // scratch may still be read by a previous call
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < N; ++i) scratch[i * nw + wid] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < N; ++i) {
        T s = 0;
        for (int k = 0; k < nw; ++k) s += scratch[i * nw + k];
        v[i] = s;
    }
}

// ---- 8x8 linear solve shared by K1 (DLT-4) and K5 (cv2.getPerspectiveTransform twin) ---------------------
// Solve the 8x8 system whose row `sub` = (a[0..7] | rhs) lives in lane `sub` of an aligned 8-lane group.
// On return x[0..7] holds the solution in every lane of the group.
__device__ __forceinline__ void solve8(double (&a)[8], double rhs, int sub, double (&x)[8]) {
    constexpr unsigned kFull = 0xffffffffu;
    bool used = false;
    int piv_lane[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        // partial pivoting: largest |a[r][k]| among rows not yet used as a pivot (lowest lane wins ties)
        double best = used ? -1.0 : fabs(a[k]);
        int who = sub;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const double ob = __shfl_xor_sync(kFull, best, o, 8);
            const int ow = __shfl_xor_sync(kFull, who, o, 8);
            if (ob > best || (ob == best && ow < who)) {
                best = ob;
                who = ow;
            }
        }
        piv_lane[k] = who;
        double prow[8];
#pragma unroll
        for (int j = k; j < 8; ++j) prow[j] = __shfl_sync(kFull, a[j], who, 8);
        const double prhs = __shfl_sync(kFull, rhs, who, 8);
        if (sub == who) {
            used = true;
        } else if (!used) {
            const double f = a[k] / prow[k];
#pragma unroll
            for (int j = k + 1; j < 8; ++j) a[j] = fma(-f, prow[j], a[j]);
            rhs = fma(-f, prhs, rhs);
            a[k] = 0.0;
        }
    }
    // back substitution: the row that was pivot k has zeros left of column k
#pragma unroll
    for (int k = 7; k >= 0; --k) {
        double acc = rhs;
#pragma unroll
        for (int j = k + 1; j < 8; ++j) acc = fma(-a[j], x[j], acc);
        x[k] = __shfl_sync(kFull, acc / a[k], piv_lane[k], 8);
    }
}

// ---- streaming global access (read-once / write-once data: keep it out of L1) --------------------
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

// ---- mbarrier + 1-D bulk TMA (cp.async.bulk, SASS: UBLKCP) ----------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy of `bytes` (multiple of 16, both addresses 16-B aligned), completion on `bar`
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- bilinear tap geometry shared by the warp kernels --------------------------------------------
// ATen grid_sampler_2d (bilinear, zeros, align_corners=True) in pixel units:
//   x0 = floor(u), weights wx0 = (x0+1)-u, wx1 = u-x0; a tap contributes iff 0 <= xi <= W-1.
struct Taps {
    int x0, y0;
    float wx0, wx1, wy0, wy1;
    bool inx0, inx1, iny0, iny1;
};
__device__ __forceinline__ Taps make_taps(float u, float v, int Ws, int Hs) {
    Taps t;
    const float fx = floorf(u), fy = floorf(v);
    t.wx0 = (fx + 1.0f) - u;
    t.wx1 = u - fx;
    t.wy0 = (fy + 1.0f) - v;
    t.wy1 = v - fy;
    // clamp before the float->int conversion so that far-away / non-finite coordinates stay out of range
    t.x0 = static_cast<int>(fminf(fmaxf(fx, -2.0f), static_cast<float>(Ws)));
    t.y0 = static_cast<int>(fminf(fmaxf(fy, -2.0f), static_cast<float>(Hs)));
    t.inx0 = static_cast<unsigned>(t.x0) < static_cast<unsigned>(Ws);
    t.inx1 = static_cast<unsigned>(t.x0 + 1) < static_cast<unsigned>(Ws);
    t.iny0 = static_cast<unsigned>(t.y0) < static_cast<unsigned>(Hs);
    t.iny1 = static_cast<unsigned>(t.y0 + 1) < static_cast<unsigned>(Hs);
    return t;
}

struct Hmat {
    float h[9];
};
__device__ __forceinline__ Hmat load_h(const float* __restrict__ H, int b) {
    Hmat m;
#pragma unroll
    for (int i = 0; i < 9; ++i) m.h[i] = __ldg(H + 9 * b + i);
    return m;
}
// source coordinates of output pixel (x,y); rw = 1/w is returned for the dH adjoint
__device__ __forceinline__ void project(const Hmat& m, float x, float y, float& u, float& v, float& rw) {
    const float w = fmaf(m.h[6], x, fmaf(m.h[7], y, m.h[8]));
    rw = 1.0f / w;
    // correctly rounded divisions: a coordinate near 128 has an ulp of 1.5e-5 px, do not add to it
    u = fmaf(m.h[0], x, fmaf(m.h[1], y, m.h[2])) / w;
    v = fmaf(m.h[3], x, fmaf(m.h[4], y, m.h[5])) / w;
}

}  // namespace bh
