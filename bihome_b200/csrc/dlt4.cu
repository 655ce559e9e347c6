// K1: 4-point offsets -> homography.  8x8 direct linear transform solved by Gaussian elimination with
// partial pivoting entirely in registers: 8 lanes per sample, lane r owns row r of [A | b], pivot search
// and row broadcast by warp shuffles (4 samples per warp, no shared memory, no global scratch).
// The elimination runs in float64 (A mixes 1, x ~ 1e2 and x*u ~ 1e4: in float32 the solve loses ~1e-4
// relative for off-origin corners); inputs and outputs are float32.  It is ~300 DFMA per sample.
//
// Reference semantics: kornia.get_perspective_transform as called from src/data/utils.py:20-24
//   row 2i   = [x, y, 1, 0, 0, 0, -x*u, -y*u | u]      (x,y) = corner_i, (u,v) = corner_i + delta_i
//   row 2i+1 = [0, 0, 0, x, y, 1, -x*v, -y*v | v]      H = [X; 1]
// Adjoint: solve A^T lambda = gH[0..7]; gDelta_i = (lambda_2i, lambda_2i+1) * (h6 x_i + h7 y_i + 1).
#include "bh_common.cuh"

namespace bh {

__device__ __forceinline__ void load_corner(const float* __restrict__ corners, int b, int i, float W, float Hh, float& x,
                                            float& y) {
    if (corners != nullptr) {
        x = __ldg(corners + (b * 4 + i) * 2 + 0);
        y = __ldg(corners + (b * 4 + i) * 2 + 1);
    } else {  // image_shape_to_corners (src/data/utils.py:42): (0,0) (W,0) (W,H) (0,H)
        x = (i == 1 || i == 2) ? W : 0.0f;
        y = (i >= 2) ? Hh : 0.0f;
    }
}

__global__ void __launch_bounds__(128) dlt4_fwd_kernel(const float* __restrict__ corners, const float* __restrict__ delta,
                                                       float* __restrict__ H, int B, float W, float Hh) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int sub = t & 7;
    const bool live = (t >> 3) < B;
    const int b = live ? (t >> 3) : (B - 1);  // idle groups shadow the last sample: shuffles stay full-warp
    const int i = sub >> 1;
    float x, y;
    load_corner(corners, b, i, W, Hh, x, y);
    const double xd = x, yd = y;
    const double u = xd + static_cast<double>(__ldg(delta + (b * 4 + i) * 2 + 0));
    const double v = yd + static_cast<double>(__ldg(delta + (b * 4 + i) * 2 + 1));
    double a[8], rhs, sol[8];
    if ((sub & 1) == 0) {
        a[0] = xd; a[1] = yd; a[2] = 1.0; a[3] = 0.0; a[4] = 0.0; a[5] = 0.0; a[6] = -xd * u; a[7] = -yd * u;
        rhs = u;
    } else {
        a[0] = 0.0; a[1] = 0.0; a[2] = 0.0; a[3] = xd; a[4] = yd; a[5] = 1.0; a[6] = -xd * v; a[7] = -yd * v;
        rhs = v;
    }
    solve8(a, rhs, sub, sol);
    if (live) {
        double mine = sol[0];
#pragma unroll
        for (int k = 1; k < 8; ++k) mine = (sub == k) ? sol[k] : mine;
        H[b * 9 + sub] = static_cast<float>(mine);
        if (sub == 0) H[b * 9 + 8] = 1.0f;
    }
}

__global__ void __launch_bounds__(128)
    dlt4_bwd_kernel(const float* __restrict__ corners, const float* __restrict__ delta, const float* __restrict__ H,
                    const float* __restrict__ gH, float* __restrict__ gDelta, float* __restrict__ gCorners, int B, float W,
                    float Hh) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int sub = t & 7;
    const bool live = (t >> 3) < B;
    const int b = live ? (t >> 3) : (B - 1);
    // lane `sub` owns row `sub` of A^T, i.e. column `sub` of A
    double a[8], lam[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float xf, yf;
        load_corner(corners, b, i, W, Hh, xf, yf);
        const double x = xf, y = yf;
        const double u = x + static_cast<double>(__ldg(delta + (b * 4 + i) * 2 + 0));
        const double v = y + static_cast<double>(__ldg(delta + (b * 4 + i) * 2 + 1));
        double ex, ey;  // A[2i][sub], A[2i+1][sub]
        switch (sub) {
            case 0: ex = x; ey = 0.0; break;
            case 1: ex = y; ey = 0.0; break;
            case 2: ex = 1.0; ey = 0.0; break;
            case 3: ex = 0.0; ey = x; break;
            case 4: ex = 0.0; ey = y; break;
            case 5: ex = 0.0; ey = 1.0; break;
            case 6: ex = -x * u; ey = -x * v; break;
            default: ex = -y * u; ey = -y * v; break;
        }
        a[2 * i] = ex;
        a[2 * i + 1] = ey;
    }
    solve8(a, static_cast<double>(__ldg(gH + b * 9 + sub)), sub, lam);
    if (!live) return;
    const int i = sub >> 1;
    float xf, yf;
    load_corner(corners, b, i, W, Hh, xf, yf);
    const double x = xf, y = yf;
    const double h6 = __ldg(H + b * 9 + 6), h7 = __ldg(H + b * 9 + 7);
    const double wi = fma(h6, x, fma(h7, y, 1.0));
    double lx = lam[0], ly = lam[1];
#pragma unroll
    for (int k = 1; k < 4; ++k) {
        lx = (i == k) ? lam[2 * k] : lx;
        ly = (i == k) ? lam[2 * k + 1] : ly;
    }
    const double gd = ((sub & 1) ? ly : lx) * wi;
    gDelta[b * 8 + sub] = static_cast<float>(gd);
    if (gCorners != nullptr) {
        const double u = x + static_cast<double>(__ldg(delta + (b * 4 + i) * 2 + 0));
        const double v = y + static_cast<double>(__ldg(delta + (b * 4 + i) * 2 + 1));
        const double h0 = __ldg(H + b * 9 + 0), h1 = __ldg(H + b * 9 + 1), h3 = __ldg(H + b * 9 + 3),
                     h4 = __ldg(H + b * 9 + 4);
        double gs;
        if ((sub & 1) == 0) gs = -lx * h0 + lx * h6 * u - ly * h3 + ly * h6 * v;
        else gs = -lx * h1 + lx * h7 * u - ly * h4 + ly * h7 * v;
        gCorners[b * 8 + sub] = static_cast<float>(gs + gd);
    }
}

}  // namespace bh

extern "C" int bh_dlt4_fwd(const float* corners, const float* delta, float* H, int B, float W, float Hh,
                           bh_stream_t stream) {
    if (!delta || !H) return BH_E_NULL;
    if (B <= 0) return BH_E_SHAPE;
    const int threads = 128, blocks = (B * 8 + threads - 1) / threads;
    bh::dlt4_fwd_kernel<<<blocks, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(corners, delta, H, B, W, Hh);
    return bh::launch_status();
}

extern "C" int bh_dlt4_bwd(const float* corners, const float* delta, const float* H, const float* gH, float* gDelta,
                           float* gCorners, int B, float W, float Hh, bh_stream_t stream) {
    if (!delta || !H || !gH || !gDelta) return BH_E_NULL;
    if (B <= 0) return BH_E_SHAPE;
    const int threads = 128, blocks = (B * 8 + threads - 1) / threads;
    bh::dlt4_bwd_kernel<<<blocks, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(corners, delta, H, gH, gDelta,
                                                                                        gCorners, B, W, Hh);
    return bh::launch_status();
}
