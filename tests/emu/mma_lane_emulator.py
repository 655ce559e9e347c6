"""Lane-level emulation of the tensor-core field-head kernels (bihome_b200/csrc/fieldhead_mma.cu) in numpy.

Each function below restates one kernel's index arithmetic lane by lane -- which element a lane loads into which fragment
register, which accumulator register it stores where -- on top of `mma()`, which implements the PTX ISA fragment layout of
mma.sync.m16n8k8 (.tf32) for a whole warp.  tests/test_field_head.py compares the results with dense float64 algebra, so a
wrong fragment mapping in the kernels' design shows up on a machine without a GPU.  (The emulation is float64: it checks
the data movement, not the TF32 head/remainder split.)"""
import numpy as np

LANES = np.arange(32)
G, T = LANES >> 2, LANES & 3


def mma(c, a, b0, b1):
    """c [32,4], a [32,4], b0/b1 [32] per-lane registers -> c + A @ B in the same per-lane layout"""
    A = np.zeros((16, 8))
    B = np.zeros((8, 8))
    C = np.zeros((16, 8))
    for l in range(32):
        g, t = l >> 2, l & 3
        A[g, t], A[g + 8, t], A[g, t + 4], A[g + 8, t + 4] = a[l]
        B[t, g], B[t + 4, g] = b0[l], b1[l]
        C[g, 2 * t], C[g, 2 * t + 1], C[g + 8, 2 * t], C[g + 8, 2 * t + 1] = c[l]
    D = C + A @ B
    out = np.zeros((32, 4))
    for l in range(32):
        g, t = l >> 2, l & 3
        out[l] = D[g, 2 * t], D[g, 2 * t + 1], D[g + 8, 2 * t], D[g + 8, 2 * t + 1]
    return out


def quad_sum(v):
    return v.reshape(8, 4).sum(1).repeat(4)


def _x_fragments(x, p0, m):
    """load_x_fragments: a[s] [32,4] for k-step s of pixel tile m"""
    xa = np.stack([x[p0 + 16 * m + G[l], 4 * T[l]:4 * T[l] + 4] for l in range(32)])
    xb = np.stack([x[p0 + 16 * m + G[l] + 8, 4 * T[l]:4 * T[l] + 4] for l in range(32)])
    return [np.stack([xa[:, 0], xb[:, 0], xa[:, 1], xb[:, 1]], 1), np.stack([xa[:, 2], xb[:, 2], xa[:, 3], xb[:, 3]], 1)]


def _w1_b(W1, j, s):
    """build_w1_b_table entry [j][s]: (b0, b1) per lane"""
    b0 = np.array([W1[8 * j + G[l], 4 * T[l] + 2 * s] for l in range(32)])
    b1 = np.array([W1[8 * j + G[l], 4 * T[l] + 2 * s + 1] for l in range(32)])
    return b0, b1


def fwd(x, W1, b1, W2, b2, HW, MT=2):
    n_pix = x.shape[0]
    B = n_pix // HW
    out = np.zeros((B, 2, HW))
    for grp in range(n_pix // (16 * MT)):
        p0 = grp * 16 * MT
        b, s0 = p0 // HW, p0 % HW
        for m in range(MT):
            a = _x_fragments(x, p0, m)
            o = np.zeros((32, 4))
            for j in range(16):
                bj = np.stack([b1[8 * j + 2 * T], b1[8 * j + 2 * T + 1]], 1)
                wa = np.stack([W2[0, 8 * j + 2 * T], W2[0, 8 * j + 2 * T + 1]], 1)
                wb = np.stack([W2[1, 8 * j + 2 * T], W2[1, 8 * j + 2 * T + 1]], 1)
                c = np.stack([bj[:, 0], bj[:, 1], bj[:, 0], bj[:, 1]], 1)
                for s in range(2):
                    c = mma(c, a[s], *_w1_b(W1, j, s))
                h = np.maximum(c, 0)
                o[:, 0] += h[:, 0] * wa[:, 0] + h[:, 1] * wa[:, 1]
                o[:, 1] += h[:, 2] * wa[:, 0] + h[:, 3] * wa[:, 1]
                o[:, 2] += h[:, 0] * wb[:, 0] + h[:, 1] * wb[:, 1]
                o[:, 3] += h[:, 2] * wb[:, 0] + h[:, 3] * wb[:, 1]
            v = [quad_sum(o[:, k]) for k in range(4)]
            for l in range(32):
                g, t = l >> 2, l & 3
                val = v[t][l] + (b2[0] if t < 2 else b2[1])
                row = 16 * m + g + (8 if (t & 1) else 0)
                out[b, t >> 1, s0 + row] = val
    return out


def gx(x, W1, b1, W2, gOut, HW, MT=2):
    n_pix = x.shape[0]
    res = np.zeros_like(x)
    for grp in range(n_pix // (16 * MT)):
        p0 = grp * 16 * MT
        b, s0 = p0 // HW, p0 % HW
        for m in range(MT):
            a = _x_fragments(x, p0, m)
            ga = np.stack([gOut[b, 0, s0 + 16 * m + G], gOut[b, 0, s0 + 16 * m + G + 8]], 1)
            gb = np.stack([gOut[b, 1, s0 + 16 * m + G], gOut[b, 1, s0 + 16 * m + G + 8]], 1)
            acc = [np.zeros((32, 4)), np.zeros((32, 4))]
            for j in range(16):
                bj = np.stack([b1[8 * j + 2 * T], b1[8 * j + 2 * T + 1]], 1)
                wa = np.stack([W2[0, 8 * j + 2 * T], W2[0, 8 * j + 2 * T + 1]], 1)
                wb = np.stack([W2[1, 8 * j + 2 * T], W2[1, 8 * j + 2 * T + 1]], 1)
                c = np.stack([bj[:, 0], bj[:, 1], bj[:, 0], bj[:, 1]], 1)
                for s in range(2):
                    c = mma(c, a[s], *_w1_b(W1, j, s))
                gh0 = np.where(c[:, 0] > 0, wa[:, 0] * ga[:, 0] + wb[:, 0] * gb[:, 0], 0)
                gh1 = np.where(c[:, 1] > 0, wa[:, 1] * ga[:, 0] + wb[:, 1] * gb[:, 0], 0)
                gh2 = np.where(c[:, 2] > 0, wa[:, 0] * ga[:, 1] + wb[:, 0] * gb[:, 1], 0)
                gh3 = np.where(c[:, 3] > 0, wa[:, 1] * ga[:, 1] + wb[:, 1] * gb[:, 1], 0)
                afrag = np.stack([gh0, gh2, gh1, gh3], 1)
                for n in range(2):      # build_w1_gx_table entry [j][n]
                    v0 = np.array([W1[8 * j + 2 * T[l], 8 * n + G[l]] for l in range(32)])
                    v1 = np.array([W1[8 * j + 2 * T[l] + 1, 8 * n + G[l]] for l in range(32)])
                    acc[n] = mma(acc[n], afrag, v0, v1)
            for n in range(2):
                for l in range(32):
                    g, t = l >> 2, l & 3
                    r0 = p0 + 16 * m + g
                    res[r0, 8 * n + 2 * t], res[r0, 8 * n + 2 * t + 1] = acc[n][l, 0], acc[n][l, 1]
                    res[r0 + 8, 8 * n + 2 * t], res[r0 + 8, 8 * n + 2 * t + 1] = acc[n][l, 2], acc[n][l, 3]
    return res


def gw(x, W1, b1, W2, gOut, HW):
    """one CTA walking every 32-pixel tile -> (gW1 [128,16], gb1 [128], gW2 [2,128], gb2 [2])"""
    n_pix = x.shape[0]
    gW1, gb1, gW2, gb2 = np.zeros((128, 16)), np.zeros(128), np.zeros((2, 128)), np.zeros(2)
    for warp in range(4):
        aW1 = [[np.zeros((32, 4)) for _ in range(2)] for _ in range(2)]
        aW2 = [np.zeros((32, 4)) for _ in range(2)]
        ab1 = [np.zeros((32, 2)) for _ in range(2)]
        for tile in range(n_pix // 32):
            p0 = tile * 32
            b, s0 = p0 // HW, p0 % HW
            sX = x[p0:p0 + 32]
            sG = gOut[b, :, s0:s0 + 32]
            for n in range(4):
                xb = [(sX[8 * n + G, 4 * T + 2 * s], sX[8 * n + G, 4 * T + 2 * s + 1]) for s in range(2)]
                xw = [(sX[8 * n + 2 * T, 8 * m + G], sX[8 * n + 2 * T + 1, 8 * m + G]) for m in range(2)]
                g0 = np.stack([sG[0, 8 * n + 2 * T], sG[0, 8 * n + 2 * T + 1]], 1)
                g1 = np.stack([sG[1, 8 * n + 2 * T], sG[1, 8 * n + 2 * T + 1]], 1)
                gbf = (np.where(G < 2, sG[np.minimum(G, 1), 8 * n + 2 * T], 0), np.where(G < 2, sG[np.minimum(G, 1), 8 * n + 2 * T + 1], 0))
                for i in range(2):
                    r0 = 32 * warp + 16 * i + G
                    r1 = r0 + 8
                    c = np.stack([b1[r0], b1[r0], b1[r1], b1[r1]], 1)
                    for s in range(2):
                        afr = np.stack([W1[r0, 4 * T + 2 * s], W1[r1, 4 * T + 2 * s], W1[r0, 4 * T + 2 * s + 1], W1[r1, 4 * T + 2 * s + 1]], 1)
                        c = mma(c, afr, *xb[s])
                    h = np.maximum(c, 0)
                    gh0 = np.where(c[:, 0] > 0, W2[0, r0] * g0[:, 0] + W2[1, r0] * g1[:, 0], 0)
                    gh1 = np.where(c[:, 1] > 0, W2[0, r0] * g0[:, 1] + W2[1, r0] * g1[:, 1], 0)
                    gh2 = np.where(c[:, 2] > 0, W2[0, r1] * g0[:, 0] + W2[1, r1] * g1[:, 0], 0)
                    gh3 = np.where(c[:, 3] > 0, W2[0, r1] * g0[:, 1] + W2[1, r1] * g1[:, 1], 0)
                    ab1[i][:, 0] += gh0 + gh1
                    ab1[i][:, 1] += gh2 + gh3
                    afrag = np.stack([gh0, gh2, gh1, gh3], 1)
                    for m in range(2):
                        aW1[i][m] = mma(aW1[i][m], afrag, *xw[m])
                    hfrag = np.stack([h[:, 0], h[:, 2], h[:, 1], h[:, 3]], 1)
                    aW2[i] = mma(aW2[i], hfrag, *gbf)
        for i in range(2):
            s0v, s1v = quad_sum(ab1[i][:, 0]), quad_sum(ab1[i][:, 1])
            for l in range(32):
                g, t = l >> 2, l & 3
                r0 = 32 * warp + 16 * i + g
                r1 = r0 + 8
                for m in range(2):
                    gW1[r0, 8 * m + 2 * t], gW1[r0, 8 * m + 2 * t + 1] = aW1[i][m][l, 0], aW1[i][m][l, 1]
                    gW1[r1, 8 * m + 2 * t], gW1[r1, 8 * m + 2 * t + 1] = aW1[i][m][l, 2], aW1[i][m][l, 3]
                if t == 0:
                    gb1[r0], gb1[r1] = s0v[l], s1v[l]
                    gW2[0, r0], gW2[1, r0], gW2[0, r1], gW2[1, r1] = aW2[i][l]
    gb2[:] = gOut.sum((0, 2))
    return gW1, gb1, gW2, gb2


def moments(x):
    """moments_mma_kernel: (sum_p x_i [16], sum_p x_i x_k [16,16]); one 8-pixel block per k-step, the lane's four loads serve
    as A fragment and as both B fragments"""
    c = [np.zeros((32, 4)), np.zeros((32, 4))]
    s0, s1 = np.zeros(32), np.zeros(32)
    for k in range(x.shape[0] // 8):
        v0, v1 = x[k * 8 + T, G], x[k * 8 + T, G + 8]
        v2, v3 = x[k * 8 + T + 4, G], x[k * 8 + T + 4, G + 8]
        s0 += v0 + v2
        s1 += v1 + v3
        a = np.stack([v0, v1, v2, v3], 1)
        c[0] = mma(c[0], a, v0, v2)
        c[1] = mma(c[1], a, v1, v3)
    M, S = np.zeros((16, 16)), np.zeros(16)
    q0, q1 = quad_sum(s0), quad_sum(s1)
    for l in range(32):
        g, t = l >> 2, l & 3
        S[g], S[g + 8] = q0[l], q1[l]
        for n in range(2):
            M[g, 8 * n + 2 * t], M[g, 8 * n + 2 * t + 1] = c[n][l, 0], c[n][l, 1]
            M[g + 8, 8 * n + 2 * t], M[g + 8, 8 * n + 2 * t + 1] = c[n][l, 2], c[n][l, 3]
    return S, M
