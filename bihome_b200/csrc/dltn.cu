// K4: N-point normalised DLT (the Zeng / DSAC branch), forward and adjoint, one warp per hypothesis.
//
// Reference semantics: kornia.find_homography_dlt as called from src/heads/ransac_utils.py:58-72, preceded by the
// gather of the sampled correspondences (:58-59) and followed by the corner projection
// kornia.transform_points(H, four_points) - four_points of src/heads/PerceptualHead.py:175-178:
//   normalize_points: m = mean(p), s = sqrt(2) / (mean|p - m| + 1e-8), pn = s (p - m)          (both point sets)
//   rows  ax = [0,0,0, -x1,-y1,-1, y2 x1, y2 y1, y2],  ay = [x1,y1,1, 0,0,0, -x2 x1, -x2 y1, -x2]
//   G = A^T A (9x9), h = eigenvector of the smallest eigenvalue (torch.svd(G).V[:, -1])
//   Hd = T2^-1 h T1,  Hn = Hd / (Hd[2,2] + 1e-8),  delta_i = proj(Hn, q_i) - q_i
// The Gram matrix has only four distinct 3x3 blocks (sum p p^T weighted by 1, x2, y2, x2^2+y2^2, p = (x1,y1,1)),
// accumulated per lane and reduced by shuffles; the eigen-decomposition is a cyclic Jacobi in shared memory.
// Everything runs in float64: the smallest eigenvector of a float32 Gram matrix is where the reference's own
// float32 result loses 3 digits (tests/golden/zeng_dsac_P32.npz: fp32 vs fp64 reference differ by ~1e-3).
// The adjoint recomputes the forward (cheaper than saving 100 doubles per sample) and back-propagates through
// the eigenvector (dv = sum_j v_j (v_j^T dG v) / (l_min - l_j)), the normalisations and the gather.
#include "bh_common.cuh"

namespace bh {

constexpr int kDltnWarps = 4;
constexpr int kMaxSweeps = 12;
constexpr double kSqrt2F32 = 1.41421353816986083984375;  // torch.sqrt(torch.tensor(2.)) is a float32 value

struct DltnArgs {
    const float* p1;       // [B,N,2] or null
    const float* p2;       // [B,N,2] or null
    const float* field;    // [B,2,N] or null (N = Hf*Wf, p1 = grid coords, p2 = p1 + field)
    const long long* choice;  // [B,M] or null (identity, M == N)
    const float* four;     // [4,2] or null
    int B, N, M, Wf;
};

__device__ __forceinline__ void fetch(const DltnArgs& a, int b, int k, double& x1, double& y1, double& x2, double& y2,
                                      long long& idx) {
    idx = a.choice ? a.choice[static_cast<long long>(b) * a.M + k] : k;
    if (a.field) {
        x1 = static_cast<double>(idx % a.Wf);
        y1 = static_cast<double>(idx / a.Wf);
        x2 = x1 + static_cast<double>(__ldg(a.field + (static_cast<long long>(b) * 2 + 0) * a.N + idx));
        y2 = y1 + static_cast<double>(__ldg(a.field + (static_cast<long long>(b) * 2 + 1) * a.N + idx));
    } else {
        const long long o = (static_cast<long long>(b) * a.N + idx) * 2;
        x1 = __ldg(a.p1 + o); y1 = __ldg(a.p1 + o + 1);
        x2 = __ldg(a.p2 + o); y2 = __ldg(a.p2 + o + 1);
    }
}

struct Norm {
    double m1x, m1y, s1, sig1, m2x, m2y, s2, sig2;
};

// everything the forward produces that the adjoint needs; G/V live in shared memory
struct Fwd {
    Norm n;
    int imin;
    double Hd[9], Hn[9], den;  // den = Hd[8] + 1e-8
};

// Parallel-ordered Jacobi on the warp's 9x9 symmetric matrix A (destroyed: eigenvalues end up on its diagonal),
// V = eigenvectors.  A round-robin tournament over 10 slots (9 indices + a bye) gives 9 rounds of 4 disjoint (p,q)
// pairs per sweep; the 4 rotations of a round commute, so their parameters are computed by 4 lanes at once and
// applied as A <- J^T (A J), V <- V J with the 72 + 36 element updates spread over the warp.
__device__ __forceinline__ void jacobi9(double* A, double* V, int lane) {
    __shared__ double rot_cs[kDltnWarps][4][2];
    __shared__ int rot_pq[kDltnWarps][4][2];
    double(*cs)[2] = rot_cs[threadIdx.x >> 5];
    int(*pq)[2] = rot_pq[threadIdx.x >> 5];
    for (int i = lane; i < 81; i += 32) V[i] = (i / 9 == i % 9) ? 1.0 : 0.0;
    __syncwarp();
    for (int sweep = 0; sweep < kMaxSweeps; ++sweep) {
        double off = 0.0, dia = 0.0;
        for (int i = lane; i < 81; i += 32) {
            const double v = A[i];
            if (i / 9 == i % 9) dia += v * v; else off += v * v;
        }
        off = warp_sum(off);
        dia = warp_sum(dia);
        if (off <= 1e-26 * dia) break;  // off-diagonal mass below 1e-13 relative (the next sweep would square it)
        for (int round = 0; round < 9; ++round) {
            if (lane < 4) {
                // circle method: slot 0 is fixed (the bye, index 9), slots 1..9 hold indices rotated by `round`
                // pairs: (slot k, slot 9-k) for k = 1..4; slot 0 pairs with slot 9... use 10 slots: 0..9
                // slot s (1..9) holds index (s - 1 + round) % 9; slot 0 holds the bye
                const int k = lane + 1;                       // 1..4
                int p = (k - 1 + round) % 9, q = (9 - k - 1 + round + 9) % 9;   // slots k and 9-k
                if (p > q) { const int t = p; p = q; q = t; }
                const double apq = A[p * 9 + q], app = A[p * 9 + p], aqq = A[q * 9 + q];
                // rotation that annihilates A[p][q], |angle| <= pi/4, without a division: with al = aqq - app, be = 2 apq,
                // cos 2phi = |al| / rho, c = sqrt((1 + cos 2phi) / 2), s = sign(al) be / (2 rho c) -- two rsqrt instead of the
                // textbook's div, sqrt, div, sqrt, div chain (the rotation parameters are the serial part of every round)
                double c = 1.0, s = 0.0;
                const double al = aqq - app, be = 2.0 * apq, r2 = al * al + be * be;
                if (fabs(apq) > 1e-300 && r2 > 1e-300) {
                    const double ir = rsqrt(r2);
                    const double hc = fma(0.5 * fabs(al), ir, 0.5);
                    const double ih = rsqrt(hc);
                    c = hc * ih;
                    s = (al >= 0.0 ? be : -be) * (0.5 * ir) * ih;
                }
                cs[lane][0] = c; cs[lane][1] = s;
                pq[lane][0] = p; pq[lane][1] = q;
            }
            __syncwarp();
            // phase 1: columns of A and of V:  X[:, p], X[:, q] <- X[:, p] c - X[:, q] s,  X[:, p] s + X[:, q] c
            for (int w = lane; w < 72; w += 32) {
                double* X = (w < 36) ? A : V;
                const int e = (w < 36) ? w : w - 36;
                const int row = e >> 2, k = e & 3;
                const int p = pq[k][0], q = pq[k][1];
                const double c = cs[k][0], s = cs[k][1];
                const double xp = X[row * 9 + p], xq = X[row * 9 + q];
                X[row * 9 + p] = c * xp - s * xq;
                X[row * 9 + q] = s * xp + c * xq;
            }
            __syncwarp();
            // phase 2: rows of A:  A[p, :], A[q, :] <- c A[p, :] - s A[q, :],  s A[p, :] + c A[q, :]
            for (int w = lane; w < 36; w += 32) {
                const int col = w >> 2, k = w & 3;
                const int p = pq[k][0], q = pq[k][1];
                const double c = cs[k][0], s = cs[k][1];
                const double xp = A[p * 9 + col], xq = A[q * 9 + col];
                A[p * 9 + col] = c * xp - s * xq;
                A[q * 9 + col] = s * xp + c * xq;
            }
            __syncwarp();
        }
    }
    __syncwarp();
}

__device__ __forceinline__ void mat3(const double* a, const double* b, double* c) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}

// forward for sample b by one warp; G, V: this warp's shared 81-double scratch
__device__ void dltn_forward(const DltnArgs& a, int b, int lane, double* G, double* V, Fwd& f) {
    const double invM = 1.0 / static_cast<double>(a.M);
    double x1, y1, x2, y2;
    long long idx;
    // means
    double s[4] = {0, 0, 0, 0};
    for (int k = lane; k < a.M; k += 32) {
        fetch(a, b, k, x1, y1, x2, y2, idx);
        s[0] += x1; s[1] += y1; s[2] += x2; s[3] += y2;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i] = warp_sum(s[i]) * invM;
    Norm& n = f.n;
    n.m1x = s[0]; n.m1y = s[1]; n.m2x = s[2]; n.m2y = s[3];
    // mean distances
    double d1 = 0, d2 = 0;
    for (int k = lane; k < a.M; k += 32) {
        fetch(a, b, k, x1, y1, x2, y2, idx);
        d1 += sqrt((x1 - n.m1x) * (x1 - n.m1x) + (y1 - n.m1y) * (y1 - n.m1y));
        d2 += sqrt((x2 - n.m2x) * (x2 - n.m2x) + (y2 - n.m2y) * (y2 - n.m2y));
    }
    n.sig1 = warp_sum(d1) * invM;
    n.sig2 = warp_sum(d2) * invM;
    n.s1 = kSqrt2F32 / (n.sig1 + 1e-8);
    n.s2 = kSqrt2F32 / (n.sig2 + 1e-8);
    // Gram blocks: S[t][6] for weights 1, x2, y2, x2^2+y2^2; entries (xx, xy, x, yy, y, 1)
    double S[4][6];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int e = 0; e < 6; ++e) S[t][e] = 0.0;
    for (int k = lane; k < a.M; k += 32) {
        fetch(a, b, k, x1, y1, x2, y2, idx);
        const double u1 = n.s1 * (x1 - n.m1x), v1 = n.s1 * (y1 - n.m1y);
        const double u2 = n.s2 * (x2 - n.m2x), v2 = n.s2 * (y2 - n.m2y);
        const double pp[6] = {u1 * u1, u1 * v1, u1, v1 * v1, v1, 1.0};
        const double wt[4] = {1.0, u2, v2, u2 * u2 + v2 * v2};
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int e = 0; e < 6; ++e) S[t][e] = fma(wt[t], pp[e], S[t][e]);
    }
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int e = 0; e < 6; ++e) S[t][e] = warp_sum(S[t][e]);
    // assemble G = [[S0, 0, -Sx], [0, S0, -Sy], [-Sx, -Sy, Sr]]
    if (lane == 0) {
        const int map[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                const int e = map[i][j];
                G[i * 9 + j] = S[0][e];            G[i * 9 + 3 + j] = 0.0;             G[i * 9 + 6 + j] = -S[1][e];
                G[(3 + i) * 9 + j] = 0.0;          G[(3 + i) * 9 + 3 + j] = S[0][e];   G[(3 + i) * 9 + 6 + j] = -S[2][e];
                G[(6 + i) * 9 + j] = -S[1][e];     G[(6 + i) * 9 + 3 + j] = -S[2][e];  G[(6 + i) * 9 + 6 + j] = S[3][e];
            }
    }
    __syncwarp();
    jacobi9(G, V, lane);
    int imin = 0;
    double lmin = G[0];
    for (int i = 1; i < 9; ++i)
        if (G[i * 10] < lmin) { lmin = G[i * 10]; imin = i; }
    f.imin = imin;
    double h[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) h[i] = V[i * 9 + imin];
    const double T1[9] = {n.s1, 0, -n.s1 * n.m1x, 0, n.s1, -n.s1 * n.m1y, 0, 0, 1};
    const double T2i[9] = {1.0 / n.s2, 0, n.m2x, 0, 1.0 / n.s2, n.m2y, 0, 0, 1};
    double hT1[9];
    mat3(h, T1, hT1);
    mat3(T2i, hT1, f.Hd);
    f.den = f.Hd[8] + 1e-8;
#pragma unroll
    for (int i = 0; i < 9; ++i) f.Hn[i] = f.Hd[i] / f.den;
}

__global__ void __launch_bounds__(kDltnWarps * 32) dltn_fwd_kernel(const DltnArgs a, float* __restrict__ Hn, float* __restrict__ delta) {
    __shared__ double Gs[kDltnWarps][81];
    __shared__ double Vs[kDltnWarps][81];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * kDltnWarps + wid;
    if (b >= a.B) return;
    Fwd f;
    dltn_forward(a, b, lane, Gs[wid], Vs[wid], f);
    if (lane < 9) {
        double mine = f.Hn[0];
#pragma unroll
        for (int i = 1; i < 9; ++i) mine = (lane == i) ? f.Hn[i] : mine;
        Hn[b * 9 + lane] = static_cast<float>(mine);
    }
    if (delta != nullptr && lane < 4) {
        const double qx = __ldg(a.four + lane * 2), qy = __ldg(a.four + lane * 2 + 1);
        const double z = f.Hn[6] * qx + f.Hn[7] * qy + f.Hn[8];
        const double sc = fabs(z) > 1e-8 ? 1.0 / z : 1.0;
        delta[(b * 4 + lane) * 2 + 0] = static_cast<float>((f.Hn[0] * qx + f.Hn[1] * qy + f.Hn[2]) * sc - qx);
        delta[(b * 4 + lane) * 2 + 1] = static_cast<float>((f.Hn[3] * qx + f.Hn[4] * qy + f.Hn[5]) * sc - qy);
    }
}

__global__ void __launch_bounds__(kDltnWarps * 32)
    dltn_bwd_kernel(const DltnArgs a, const float* __restrict__ gHn_in, const float* __restrict__ gDelta, float* gP2,
                    float* gField) {
    __shared__ double Gs[kDltnWarps][81];
    __shared__ double Vs[kDltnWarps][81];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * kDltnWarps + wid;
    if (b >= a.B) return;
    Fwd f;
    double* G = Gs[wid];
    double* V = Vs[wid];
    dltn_forward(a, b, lane, G, V, f);
    const Norm& n = f.n;
    // ---- upstream into gHn (all lanes redundantly: 9 values) ----
    double gHn[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) gHn[i] = gHn_in ? static_cast<double>(__ldg(gHn_in + b * 9 + i)) : 0.0;
    if (gDelta != nullptr) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const double qx = __ldg(a.four + c * 2), qy = __ldg(a.four + c * 2 + 1);
            const double gx = __ldg(gDelta + (b * 4 + c) * 2), gy = __ldg(gDelta + (b * 4 + c) * 2 + 1);
            const double z = f.Hn[6] * qx + f.Hn[7] * qy + f.Hn[8];
            if (fabs(z) > 1e-8) {
                const double u = (f.Hn[0] * qx + f.Hn[1] * qy + f.Hn[2]) / z, v = (f.Hn[3] * qx + f.Hn[4] * qy + f.Hn[5]) / z;
                const double gX = gx / z, gY = gy / z, gz = -(gx * u + gy * v) / z;
                gHn[0] += gX * qx; gHn[1] += gX * qy; gHn[2] += gX;
                gHn[3] += gY * qx; gHn[4] += gY * qy; gHn[5] += gY;
                gHn[6] += gz * qx; gHn[7] += gz * qy; gHn[8] += gz;
            } else {  // scale == 1 branch of convert_points_from_homogeneous
                gHn[0] += gx * qx; gHn[1] += gx * qy; gHn[2] += gx;
                gHn[3] += gy * qx; gHn[4] += gy * qy; gHn[5] += gy;
            }
        }
    }
    // ---- Hn = Hd / den ----
    double gHd[9], dot = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) { gHd[i] = gHn[i] / f.den; dot += gHn[i] * f.Hd[i]; }
    gHd[8] -= dot / (f.den * f.den);
    // ---- Hd = T2i h T1 ----
    double h[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) h[i] = V[i * 9 + f.imin];
    const double T1[9] = {n.s1, 0, -n.s1 * n.m1x, 0, n.s1, -n.s1 * n.m1y, 0, 0, 1};
    const double T2i[9] = {1.0 / n.s2, 0, n.m2x, 0, 1.0 / n.s2, n.m2y, 0, 0, 1};
    const double T1t[9] = {T1[0], T1[3], T1[6], T1[1], T1[4], T1[7], T1[2], T1[5], T1[8]};
    const double T2it[9] = {T2i[0], T2i[3], T2i[6], T2i[1], T2i[4], T2i[7], T2i[2], T2i[5], T2i[8]};
    double tmp[9], gh[9], hT1[9], hT1t[9], gT2i[9];
    mat3(T2it, gHd, tmp);
    mat3(tmp, T1t, gh);  // g_h = T2i^T gHd T1^T
    mat3(h, T1, hT1);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) hT1t[i * 3 + j] = hT1[j * 3 + i];
    mat3(gHd, hT1t, gT2i);  // g_T2i = gHd (h T1)^T
    // (T1 and the first point set carry no gradient on this path: the coordinates are constants)
    double g_s2 = -(gT2i[0] + gT2i[4]) / (n.s2 * n.s2);
    double g_m2x = gT2i[2], g_m2y = gT2i[5];
    // ---- eigenvector adjoint: w = sum_{j != min} v_j (v_j . g_h) / (l_min - l_j) ----
    double w[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) w[i] = 0.0;
    const double lmin = G[f.imin * 10];
    for (int j = 0; j < 9; ++j) {
        if (j == f.imin) continue;
        double dj = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) dj += V[i * 9 + j] * gh[i];
        const double gap = lmin - G[j * 10];
        if (fabs(gap) < 1e-300) continue;
        const double cj = dj / gap;
#pragma unroll
        for (int i = 0; i < 9; ++i) w[i] += cj * V[i * 9 + j];
    }
    // ---- per point: g_row = (row.w) v + (row.v) w  ->  g(u2, v2);  sums for the normalisation adjoint ----
    double sum_gu = 0, sum_gv = 0, sum_s = 0;
    for (int k = lane; k < a.M; k += 32) {
        double x1, y1, x2, y2;
        long long idx;
        fetch(a, b, k, x1, y1, x2, y2, idx);
        const double u1 = n.s1 * (x1 - n.m1x), v1 = n.s1 * (y1 - n.m1y);
        const double u2 = n.s2 * (x2 - n.m2x), v2 = n.s2 * (y2 - n.m2y);
        // ax = [0,0,0,-u1,-v1,-1, v2 u1, v2 v1, v2], ay = [u1,v1,1,0,0,0,-u2 u1,-u2 v1,-u2]
        const double axw = -u1 * w[3] - v1 * w[4] - w[5] + v2 * (u1 * w[6] + v1 * w[7] + w[8]);
        const double axv = -u1 * h[3] - v1 * h[4] - h[5] + v2 * (u1 * h[6] + v1 * h[7] + h[8]);
        const double ayw = u1 * w[0] + v1 * w[1] + w[2] - u2 * (u1 * w[6] + v1 * w[7] + w[8]);
        const double ayv = u1 * h[0] + v1 * h[1] + h[2] - u2 * (u1 * h[6] + v1 * h[7] + h[8]);
        // d/dv2 of ax: coefficients (u1, v1, 1) at columns 6..8;  d/du2 of ay: -(u1, v1, 1) at columns 6..8
        const double gv2 = axw * (u1 * h[6] + v1 * h[7] + h[8]) + axv * (u1 * w[6] + v1 * w[7] + w[8]);
        const double gu2 = -(ayw * (u1 * h[6] + v1 * h[7] + h[8]) + ayv * (u1 * w[6] + v1 * w[7] + w[8]));
        sum_gu += gu2;
        sum_gv += gv2;
        sum_s += gu2 * (x2 - n.m2x) + gv2 * (y2 - n.m2y);
    }
    sum_gu = warp_sum(sum_gu);
    sum_gv = warp_sum(sum_gv);
    sum_s = warp_sum(sum_s);
    g_s2 += sum_s;
    g_m2x += -n.s2 * sum_gu;
    g_m2y += -n.s2 * sum_gv;
    const double g_sig2 = -g_s2 * n.s2 / (n.sig2 + 1e-8);
    const double invM = 1.0 / static_cast<double>(a.M);
    // sigma2 = mean |p - m|: its gradient also reaches m2
    double sx = 0, sy = 0;
    for (int k = lane; k < a.M; k += 32) {
        double x1, y1, x2, y2;
        long long idx;
        fetch(a, b, k, x1, y1, x2, y2, idx);
        const double dx = x2 - n.m2x, dy = y2 - n.m2y, r = sqrt(dx * dx + dy * dy);
        if (r > 0.0) { sx += dx / r; sy += dy / r; }
    }
    sx = warp_sum(sx);
    sy = warp_sum(sy);
    g_m2x += -g_sig2 * invM * sx;
    g_m2y += -g_sig2 * invM * sy;
    // ---- final per-point gradient and scatter ----
    for (int k = lane; k < a.M; k += 32) {
        double x1, y1, x2, y2;
        long long idx;
        fetch(a, b, k, x1, y1, x2, y2, idx);
        const double u1 = n.s1 * (x1 - n.m1x), v1 = n.s1 * (y1 - n.m1y);
        const double u2 = n.s2 * (x2 - n.m2x), v2 = n.s2 * (y2 - n.m2y);
        const double rw_ = u1 * w[6] + v1 * w[7] + w[8], rh = u1 * h[6] + v1 * h[7] + h[8];
        const double axw = -u1 * w[3] - v1 * w[4] - w[5] + v2 * rw_;
        const double axv = -u1 * h[3] - v1 * h[4] - h[5] + v2 * rh;
        const double ayw = u1 * w[0] + v1 * w[1] + w[2] - u2 * rw_;
        const double ayv = u1 * h[0] + v1 * h[1] + h[2] - u2 * rh;
        const double gv2 = axw * rh + axv * rw_;
        const double gu2 = -(ayw * rh + ayv * rw_);
        const double dx = x2 - n.m2x, dy = y2 - n.m2y, r = sqrt(dx * dx + dy * dy);
        double gx = n.s2 * gu2 + g_m2x * invM, gy = n.s2 * gv2 + g_m2y * invM;
        if (r > 0.0) { gx += g_sig2 * invM * dx / r; gy += g_sig2 * invM * dy / r; }
        if (gField != nullptr) {
            atomicAdd(gField + (static_cast<long long>(b) * 2 + 0) * a.N + idx, static_cast<float>(gx));
            atomicAdd(gField + (static_cast<long long>(b) * 2 + 1) * a.N + idx, static_cast<float>(gy));
        } else {
            atomicAdd(gP2 + (static_cast<long long>(b) * a.N + idx) * 2 + 0, static_cast<float>(gx));
            atomicAdd(gP2 + (static_cast<long long>(b) * a.N + idx) * 2 + 1, static_cast<float>(gy));
        }
    }
}

inline int check_dltn(const float* p1, const float* p2, const float* field, const long long* choice, int B, int N, int M,
                      int Wf) {
    if (field) {
        if (p1 || p2) return BH_E_UNSUPPORTED;
        if (Wf <= 0 || N % Wf) return BH_E_SHAPE;
    } else if (!p1 || !p2) {
        return BH_E_NULL;
    }
    if (B <= 0 || N <= 0 || M < 4) return BH_E_SHAPE;
    if (!choice && M != N) return BH_E_SHAPE;
    return BH_OK;
}

}  // namespace bh

extern "C" int bh_dltn_fwd(const float* p1, const float* p2, const float* field, const int64_t* choice, const float* four,
                           float* Hn, float* delta, int B, int N, int M, int Wf, bh_stream_t stream) {
    using namespace bh;
    const int rc = check_dltn(p1, p2, field, reinterpret_cast<const long long*>(choice), B, N, M, Wf);
    if (rc != BH_OK) return rc;
    if (!Hn) return BH_E_NULL;
    if (delta && !four) return BH_E_NULL;
    DltnArgs a{p1, p2, field, reinterpret_cast<const long long*>(choice), four, B, N, M, Wf};
    dltn_fwd_kernel<<<(B + kDltnWarps - 1) / kDltnWarps, kDltnWarps * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a, Hn, delta);
    return launch_status();
}

extern "C" int bh_dltn_bwd(const float* p1, const float* p2, const float* field, const int64_t* choice, const float* four,
                           const float* gHn, const float* gDelta, float* gP2, float* gField, int B, int N, int M, int Wf,
                           bh_stream_t stream) {
    using namespace bh;
    const int rc = check_dltn(p1, p2, field, reinterpret_cast<const long long*>(choice), B, N, M, Wf);
    if (rc != BH_OK) return rc;
    if (!gHn && !gDelta) return BH_E_NULL;
    if (gDelta && !four) return BH_E_NULL;
    if (field ? !gField : !gP2) return BH_E_NULL;
    DltnArgs a{p1, p2, field, reinterpret_cast<const long long*>(choice), four, B, N, M, Wf};
    dltn_bwd_kernel<<<(B + kDltnWarps - 1) / kDltnWarps, kDltnWarps * 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        a, gHn, gDelta, gP2, gField);
    return launch_status();
}
