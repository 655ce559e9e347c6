"""GPU parity: every CUDA kernel, called through the C ABI (bihome_b200.functional -> ctypes), against
 (a) the committed golden vectors of the unmodified reference (tests/golden, fp32 and fp64),
 (b) the CPU oracle (oracle/ref_path.py) evaluated in float64 on the same seeded inputs.

Tolerances (BASELINE.json north_star): 1e-5 norm-wise relative for H, warped images, loss and gradients,
1e-4 for gradients that pass through the DLT adjoint.  Norm-wise because the reference's own fp32 pipeline
is only reproducible to ~6e-6 relative against its fp64 evaluation."""
import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu

TOL = 1e-5
TOL_DLT_ADJOINT = 1e-4


@pytest.fixture(scope='module')
def F():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    import bihome_b200.functional as fn
    return fn


def cu(a):
    return torch.as_tensor(np.asarray(a), dtype=torch.float32).cuda()


def R():
    from oracle import ref_path
    return ref_path


# ----------------------------------------------------------------------------------------------- K1
def test_dlt4_golden(F, golden):
    g = golden('warp_P64.npz')
    P = g['img'].shape[-1]
    H = F.dlt4(cu(g['delta']), size=(P, P))
    assert rel_l2(H.cpu().numpy(), g['H64']) < TOL
    assert rel_l2(H.cpu().numpy(), g['H32']) < TOL
    assert (H[:, 2, 2] == 1).all()


@pytest.mark.parametrize('B', [1, 3, 257, 4096])
def test_dlt4_random_and_adjoint(F, B):
    gen = torch.Generator().manual_seed(B)
    P = 128.0
    # the oracle sees the float32-rounded inputs the kernel gets (cond(A) ~ 1e2 amplifies input rounding)
    delta = ((torch.rand(B, 4, 2, generator=gen, dtype=torch.float64) * 2 - 1) * 32).float().double()
    corners = torch.tensor([[0, 0], [P, 0], [P, P], [0, P]], dtype=torch.float64).repeat(B, 1, 1)
    gH = torch.randn(B, 3, 3, generator=gen, dtype=torch.float64).float().double()
    d64 = delta.clone().requires_grad_(True)
    H64 = R().four_point_to_homography(corners, d64)
    gd64, = torch.autograd.grad((H64 * gH).sum(), d64)
    d = delta.float().cuda().requires_grad_(True)
    H = F.dlt4(d, size=(P, P))
    assert rel_l2(H.detach().cpu().numpy(), H64.detach().numpy()) < TOL
    gd, = torch.autograd.grad((H * gH.float().cuda()).sum(), d)
    assert rel_l2(gd.cpu().numpy(), gd64.numpy()) < TOL_DLT_ADJOINT
    # explicit (non canonical) corners incl. their gradient
    shift = torch.rand(B, 1, 2, generator=gen, dtype=torch.float64) * 100
    c64 = (corners + shift).float().double().requires_grad_(True)
    d64 = delta.clone().requires_grad_(True)
    H64 = R().four_point_to_homography(c64, d64)
    gc64, gd64 = torch.autograd.grad((H64 * gH).sum(), (c64, d64))
    c = c64.detach().float().cuda().requires_grad_(True)
    d = delta.float().cuda().requires_grad_(True)
    H = F.dlt4(d, corners=c)
    assert rel_l2(H.detach().cpu().numpy(), H64.detach().numpy()) < TOL
    gc, gd = torch.autograd.grad((H * gH.float().cuda()).sum(), (c, d))
    assert rel_l2(gd.cpu().numpy(), gd64.numpy()) < TOL_DLT_ADJOINT
    assert rel_l2(gc.cpu().numpy(), gc64.numpy()) < TOL_DLT_ADJOINT


def test_dlt4_identity(F):
    H = F.dlt4(torch.zeros(5, 4, 2, device='cuda'), size=(128, 128))
    assert torch.allclose(H, torch.eye(3, device='cuda').expand(5, 3, 3), atol=1e-7)


# ----------------------------------------------------------------------------------------------- K2
def test_warp_golden(F, golden):
    g = golden('warp_P64.npz')
    P = g['img'].shape[-1]
    out, mask = F.warp(cu(g['img']), cu(g['H64']), P, P, pool=4)
    assert rel_l2(out.cpu().numpy(), g['warped64']) < TOL
    assert rel_l2(out.cpu().numpy(), g['warped32']) < 2e-5          # reference fp32 noise floor
    pooled64 = torch.nn.functional.avg_pool2d(torch.from_numpy(g['mask64']), 4).squeeze(1).numpy()
    assert np.abs(mask.cpu().numpy() - pooled64).max() < 1e-5
    full = F.coverage_mask(cu(g['H64']), (P, P), (P, P), pool=1)
    assert np.abs(full.cpu().numpy() - g['mask64'][:, 0]).max() < 3e-5


def test_warp_golden_gradients(F, golden):
    g = golden('warp_P64.npz')
    P = g['img'].shape[-1]
    # dH at the H node for a given dOut (reference: autograd through inverse + warp_perspective)
    H = cu(g['H64']).requires_grad_(True)
    out = F.warp(cu(g['img']), H, P, P)
    gH, = torch.autograd.grad((out * cu(g['g_out'])).sum(), H)
    assert rel_l2(gH.cpu().numpy(), g['gH_img64']) < 5e-5
    # d/d(delta) through DLT adjoint, image and mask separately
    d = cu(g['delta']).requires_grad_(True)
    Hd = F.dlt4(d, size=(P, P))
    out = F.warp(cu(g['img']), Hd, P, P)
    gd, = torch.autograd.grad((out * cu(g['g_out'])).sum(), d)
    assert rel_l2(gd.cpu().numpy(), g['gdelta_img64']) < TOL_DLT_ADJOINT
    d = cu(g['delta']).requires_grad_(True)
    m = F.coverage_mask(F.dlt4(d, size=(P, P)), (P, P), (P, P), pool=1)
    gd, = torch.autograd.grad((m * cu(g['g_mask'][:, 0])).sum(), d)
    assert rel_l2(gd.cpu().numpy(), g['gdelta_mask64']) < TOL_DLT_ADJOINT


def _rand_h(B, P, gen, scale=0.25):
    delta = (torch.rand(B, 4, 2, generator=gen, dtype=torch.float64) * 2 - 1) * P * scale
    corners = torch.tensor([[0, 0], [P, 0], [P, P], [0, P]], dtype=torch.float64).repeat(B, 1, 1)
    return R().four_point_to_homography(corners, delta)


WARP_SHAPES = [
    (5, 1, 128, 128, 128, 128, False),     # ring path, north-star shape
    (300, 1, 128, 128, 128, 128, False),   # more items than resident CTAs: every CTA walks several items through both stages
    (2, 3, 64, 96, 32, 48, False),         # ring path, C>1, rectangular, resampling
    (2, 1, 240, 320, 240, 320, False),     # large plane: every block stages its own source box
    (2, 2, 256, 256, 64, 64, False),       # 4x down-sampling: the box of a block exceeds the stage -> read-only cache taps
    (3, 1, 128, 128, 192, 160, False),     # 1.5x up-sampling, several blocks per plane, partial blocks
    (3, 1, 128, 128, 36, 44, False),       # partial strips (36 = 2 x 16 + 4 rows, 44 = 32 + 12 columns)
    (3, 2, 37, 53, 29, 31, False),         # odd sizes -> generic
    (2, 8, 40, 40, 40, 40, True),          # channels-last vec4
    (2, 6, 33, 20, 17, 24, True),          # channels-last generic
    (3, 4, 33, 31, 25, 27, True),          # warp-cooperative channels-last kernels: one lane per pixel ...
    (2, 64, 40, 48, 40, 48, True),         # ... 16 lanes per pixel (the sweep's C = 64)
    (2, 128, 24, 24, 30, 26, True),        # ... 32 lanes per pixel, partial last group of 32 pixels
    (1, 256, 20, 20, 24, 16, True),        # ... two channel quads per lane
    (2, 64, 32, 32, 64, 64, True),         # 2x up-sampling: most taps outside the source
    (2, 64, 96, 96, 32, 32, True),         # 3x down-sampling
]


def _kernel_cells(H32, Ho, Wo):
    """floor(u), floor(v) exactly as the kernels compute them (bh_common.cuh project(): float32 fma chains and
    correctly rounded divisions), emulated in float64 + rounding.  d out/dH is discontinuous where a coordinate
    crosses an integer; fixing the cell to the kernel's choice lets values AND gradients be compared at 1e-5 on
    arbitrary projective H and white-noise images (either one-sided derivative is a valid answer there)."""
    r32 = lambda a: a.astype(np.float32).astype(np.float64)
    h = H32.reshape(-1, 9).astype(np.float64)[:, :, None, None]
    ys, xs = np.meshgrid(np.arange(Ho, dtype=np.float64), np.arange(Wo, dtype=np.float64), indexing='ij')
    fma = lambda a, b, c: r32(a * b + c)
    w = fma(h[:, 6], xs, fma(h[:, 7], ys, h[:, 8]))
    u = r32(fma(h[:, 0], xs, fma(h[:, 1], ys, h[:, 2])) / w)
    v = r32(fma(h[:, 3], xs, fma(h[:, 4], ys, h[:, 5])) / w)
    return torch.from_numpy(np.floor(u)), torch.from_numpy(np.floor(v))


@pytest.fixture(params=['tile', 'tile4', 'ring'])
def warp_path(request, F):
    """the implementations behind bh_warp_fwd / bh_warp_bwd: the tile kernels (default: two warps per tile; 'tile4': the
    four-warp build) and the persistent ring kernels for planar tensors; channels-last feature maps take the
    warp-cooperative kernels by default and the round-1 thread-per-quad kernels under 'ring'"""
    F.tune('warp_path', 1 if request.param == 'ring' else 0)
    F.tune('warp_variant', 1 if request.param == 'tile4' else 0)
    yield request.param
    F.tune('warp_path', 0)
    F.tune('warp_variant', 0)


@pytest.mark.parametrize('B,C,Hs,Ws,Ho,Wo,nhwc', WARP_SHAPES)
def test_warp_paths_vs_oracle(F, warp_path, B, C, Hs, Ws, Ho, Wo, nhwc):
    gen = torch.Generator().manual_seed(B * 1000 + C)
    img = torch.rand(B, C, Hs, Ws, generator=gen, dtype=torch.float64).float().double()
    H = _rand_h(B, min(Hs, Ws), gen).float().double()
    gO = torch.randn(B, C, Ho, Wo, generator=gen, dtype=torch.float64).float().double()
    cells = _kernel_cells(H.numpy().astype(np.float32), Ho, Wo)
    H64 = H.clone().requires_grad_(True)
    i64 = img.clone().requires_grad_(True)
    ref = _warp_direct_autograd(i64, H64, Ho, Wo, cells)
    gi64, gH64 = torch.autograd.grad((ref * gO).sum(), (i64, H64))
    # the fixed-cell evaluation is the true warp (bilinear interpolation is continuous across cells)
    assert rel_l2(ref.detach().numpy(), R().warp_direct(img, H, Ho, Wo).numpy()) < 1e-6
    x = img.float().cuda()
    if nhwc:
        x = x.contiguous(memory_format=torch.channels_last)
    x.requires_grad_(True)
    Hc = H.float().cuda().requires_grad_(True)
    out = F.warp(x, Hc, Ho, Wo)
    assert rel_l2(out.detach().cpu().numpy(), ref.detach().numpy()) < TOL
    gi, gH = torch.autograd.grad((out * gO.float().cuda()).sum(), (x, Hc))
    # coordinates beyond 256 px have twice the ulp of the 128-px north-star patches
    assert rel_l2(gH.cpu().numpy(), gH64.numpy()) < (TOL if max(Hs, Ws) <= 128 else 3e-5)
    assert rel_l2(gi.cpu().numpy(), gi64.numpy()) < 2e-5     # tap weights carry the fp32 coordinate ulp (7.6e-6 px at 64..128)
    if nhwc:
        # without an image gradient the channels-last backward takes the warp-cooperative dH-only kernel (same project(),
        # hence the same bilinear cells as the oracle's; the planar dH-only kernels round their reciprocal differently and
        # are compared on smooth images in test_warp_golden_gradients / test_gpu_fullsize)
        Hd = H.float().cuda().requires_grad_(True)
        gH2, = torch.autograd.grad((F.warp(x.detach(), Hd, Ho, Wo) * gO.float().cuda()).sum(), Hd)
        assert rel_l2(gH2.cpu().numpy(), gH64.numpy()) < TOL


def _warp_direct_autograd(img, H, Ho, Wo, cells=None):
    """differentiable float64 closed form (same math as oracle.ref_path.warp_direct); `cells` fixes floor(u), floor(v)"""
    B, C, Hs, Ws = img.shape
    ys, xs = torch.meshgrid(torch.arange(Ho, dtype=H.dtype), torch.arange(Wo, dtype=H.dtype), indexing='ij')
    h = H.reshape(-1, 9)
    gg = lambda i: h[:, i].view(-1, 1, 1)
    w = gg(6) * xs + gg(7) * ys + gg(8)
    u = (gg(0) * xs + gg(1) * ys + gg(2)) / w
    v = (gg(3) * xs + gg(4) * ys + gg(5)) / w
    x0, y0 = (torch.floor(u.detach()), torch.floor(v.detach())) if cells is None else cells
    out = 0
    flat = img.reshape(B, C, -1)
    for dy in (0, 1):
        for dx in (0, 1):
            xi, yi = x0 + dx, y0 + dy
            wx = (u - x0) if dx else (x0 + 1 - u)
            wy = (v - y0) if dy else (y0 + 1 - v)
            ok = ((xi >= 0) & (xi <= Ws - 1) & (yi >= 0) & (yi <= Hs - 1)).to(img.dtype)
            idx = (yi.clamp(0, Hs - 1) * Ws + xi.clamp(0, Ws - 1)).long().view(B, 1, -1).expand(B, C, -1)
            val = torch.gather(flat, 2, idx).view(B, C, Ho, Wo)
            out = out + val * (wx * wy * ok).unsqueeze(1)
    return out


@pytest.mark.parametrize('pool', [1, 2, 4, 8])
def test_pooled_mask_and_gradient(F, warp_path, pool):
    gen = torch.Generator().manual_seed(pool)
    B, P = 6, 64
    H = _rand_h(B, P, gen, scale=0.4)
    gM = torch.randn(B, P // pool, P // pool, generator=gen, dtype=torch.float64)
    H64 = H.clone().requires_grad_(True)
    m64 = torch.nn.functional.avg_pool2d(R().analytic_mask(H64, P, P, P, P), pool).squeeze(1)
    gH64, = torch.autograd.grad((m64 * gM).sum(), H64)
    Hc = H.float().cuda().requires_grad_(True)
    img = torch.rand(B, 1, P, P, device='cuda')
    _, m = F.warp(img, Hc, P, P, pool=pool)
    assert np.abs(m.detach().cpu().numpy() - m64.detach().numpy()).max() < 2e-5
    gH, = torch.autograd.grad((m * gM.float().cuda()).sum(), Hc)
    assert rel_l2(gH.cpu().numpy(), gH64.numpy()) < 5e-5


def test_warp_identity_and_translation(F):
    img = torch.rand(3, 1, 128, 128, device='cuda')
    eye = torch.eye(3, device='cuda').repeat(3, 1, 1)
    out, mask = F.warp(img, eye, 128, 128, pool=4)
    assert torch.equal(out, img) and torch.equal(mask, torch.ones_like(mask))
    T = eye.clone()
    T[:, 0, 2], T[:, 1, 2] = 5.0, -3.0                      # out[y,x] = img[y-3, x+5]
    out = F.warp(img, T, 128, 128)
    assert torch.equal(out[:, :, 3:, :123], img[:, :, :125, 5:])
    assert (out[:, :, :3] == 0).all() and (out[:, :, :, 123:] == 0).all()


def test_rejects_cpu_tensors(F):
    with pytest.raises(RuntimeError):
        F.warp(torch.rand(1, 1, 8, 8), torch.eye(3).unsqueeze(0), 8, 8)
    with pytest.raises(RuntimeError):
        F.dlt4(torch.zeros(1, 4, 2), size=(8, 8))


# ----------------------------------------------------------------------------------------------- K3
LOSS_VARIANTS = {'ldg': 1, 'cluster': 2, 'stream': 3}      # bh_tune_set("loss_variant", .)

def _loss_inputs(B, C, h, w, seed, user_masks=False):
    gen = torch.Generator().manual_seed(seed)
    f = [torch.relu(torch.randn(B, C, h, w, generator=gen, dtype=torch.float64)) for _ in range(4)]
    m1w = torch.rand(B, h, w, generator=gen, dtype=torch.float64)
    m2w = torch.rand(B, h, w, generator=gen, dtype=torch.float64)
    m1w[0] *= 0.5 / m1w[0].sum()            # S1 < 1 : clamp branch
    m2w[-1] = 0                              # empty overlap
    m1 = torch.rand(B, h, w, generator=gen, dtype=torch.float64) if user_masks else None
    m2 = torch.rand(B, h, w, generator=gen, dtype=torch.float64) if user_masks else None
    H12 = torch.eye(3, dtype=torch.float64) + 0.1 * torch.randn(B, 3, 3, generator=gen, dtype=torch.float64)
    H21 = torch.eye(3, dtype=torch.float64) + 0.1 * torch.randn(B, 3, 3, generator=gen, dtype=torch.float64)
    return f, m1w, m2w, m1, m2, H12, H21


def _loss_oracle(f, m1w, m2w, m1, m2, H12, H21, mu):
    ones = torch.ones_like(m1w)
    a = lambda t: (ones if t is None else t).unsqueeze(1)
    loss, parts = R().bihome_double_line(f[0], f[1], f[2], f[3], a(m1), a(m2), m1w.unsqueeze(1), m2w.unsqueeze(1), H12, H21, mu)
    return loss, parts


@pytest.mark.parametrize('B,C,h,w,nhwc,user_masks,in_grads', [
    (6, 64, 32, 32, False, False, False),     # north-star shape (vec4, cluster of 8)
    (150, 16, 16, 16, False, False, False),   # cluster of 4
    (300, 8, 16, 16, False, True, False),     # cluster of 2, user masks
    (4, 5, 7, 9, False, False, True),         # odd: scalar path, gradients w.r.t. f1/f2 too
    (4, 16, 8, 8, True, True, True),          # channels-last vec4, 4 lanes per pixel
    (5, 64, 32, 32, True, False, False),      # channels-last, north-star shape (16 lanes per pixel)
    (3, 256, 8, 8, True, False, False),       # channels-last, 2 quads per lane
    (3, 24, 6, 6, True, False, True),         # channels-last, C/4 not a power of two -> scalar strides
])
def test_bihome_loss_vs_oracle(F, B, C, h, w, nhwc, user_masks, in_grads):
    mu = 0.01
    f, m1w, m2w, m1, m2, H12, H21 = _loss_inputs(B, C, h, w, seed=B + C, user_masks=user_masks)
    leaves64 = [t.clone().requires_grad_(True) for t in (f[0], f[1], f[2], f[3], m1w, m2w, H12, H21)]
    loss64, parts64 = _loss_oracle(leaves64[:4], leaves64[4], leaves64[5], m1, m2, leaves64[6], leaves64[7], mu)
    g64 = torch.autograd.grad(loss64, leaves64)

    def dev(t, feat=False):
        if t is None:
            return None
        t = t.float().cuda()
        if feat and nhwc:
            t = t.contiguous(memory_format=torch.channels_last)
        return t
    leaves = [dev(t, feat=i < 4).requires_grad_(i >= 2 or in_grads) for i, t in enumerate((f[0], f[1], f[2], f[3], m1w, m2w, H12, H21))]
    loss_b, parts = F.bihome_loss(leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], leaves[5], leaves[6], leaves[7], mu,
                                  m1=dev(m1), m2=dev(m2))
    ref_b = parts64['ln1'] + parts64['ln2'] + mu * parts64['ln3']
    assert rel_l2(loss_b.detach().cpu().numpy(), ref_b.detach().numpy()) < TOL
    assert abs(loss_b.sum().item() - loss64.item()) < TOL * abs(loss64.item()) + 1e-6
    assert rel_l2(parts[:, 2].cpu().numpy(), parts64['den1'].detach().numpy()) < TOL
    assert rel_l2(parts[:, 4].cpu().numpy(), parts64['ln3'].detach().numpy()) < TOL
    wanted = [i for i in range(8) if leaves[i].requires_grad]
    g = torch.autograd.grad(loss_b.sum(), [leaves[i] for i in wanted])
    for gi, i in zip(g, wanted):
        assert rel_l2(gi.cpu().numpy(), g64[i].numpy()) < TOL, 'gradient %d' % i


@pytest.mark.parametrize('variant', ['ldg', 'cluster', 'stream'])
@pytest.mark.parametrize('B,C,h,w,user_masks,in_grads', [
    (5, 64, 32, 32, False, False),     # north-star shape: 16 lanes per pixel, 64 tiles per sample
    (3, 256, 8, 8, False, True),       # two float4 per thread and tensor (128 KB ring), gradients w.r.t. f1/f2 too
    (7, 16, 6, 6, True, False),        # 36 pixels per sample: one PARTIAL tile of 64 pixels, user masks
    (9, 4, 20, 20, False, False),      # one lane per pixel, 400 pixels = 1 full + 1 partial tile of 256
    (600, 8, 8, 8, False, False),      # more samples than resident CTAs (the batch size that selects 'cluster' by default)
])
def test_bihome_loss_channels_last_variants(F, monkeypatch, variant, B, C, h, w, user_masks, in_grads):
    """the three channels-last kernels (cluster + LDG, cluster + TMA ring, persistent TMA stream + finish) against the
    float64 oracle on the same inputs; bh_tune_set("loss_variant") is the library's microbenchmark switch"""
    F.tune('loss_variant', LOSS_VARIANTS[variant])
    try:
        test_bihome_loss_vs_oracle(F, B, C, h, w, True, user_masks, in_grads)
    finally:
        F.tune('loss_variant', 0)


def test_bihome_loss_variants_agree_on_gradients(F, monkeypatch):
    """feature gradients are sign * W / den; the variants differ only in the order the mask sums are added (per-CTA
    partial sums in the cluster kernels, one pass in the stream kernel): agreement to a few ulp"""
    f, m1w, m2w, _, _, H12, H21 = _loss_inputs(6, 64, 32, 32, seed=11)
    cl = lambda t: t.float().cuda().contiguous(memory_format=torch.channels_last)
    outs = []
    for variant in ('ldg', 'cluster', 'stream'):
        F.tune('loss_variant', LOSS_VARIANTS[variant])
        a = [cl(f[2]).requires_grad_(True), cl(f[3]).requires_grad_(True)]
        loss_b, _ = F.bihome_loss(cl(f[0]), cl(f[1]), a[0], a[1], m1w.float().cuda(), m2w.float().cuda(), H12.float().cuda(),
                                  H21.float().cuda(), 0.01)
        outs.append(torch.autograd.grad(loss_b.sum(), a))
    F.tune('loss_variant', 0)
    for other in outs[1:]:
        assert torch.allclose(outs[0][0], other[0], rtol=2e-6, atol=0) and torch.allclose(outs[0][1], other[1], rtol=2e-6, atol=0)


def test_bihome_loss_upstream_scale(F):
    """backward with a non-unit upstream gradient goes through bh_bihome_rescale"""
    f, m1w, m2w, _, _, H12, H21 = _loss_inputs(3, 8, 8, 8, seed=1)
    mk = lambda t: t.float().cuda().requires_grad_(True)
    a = [mk(t) for t in (f[2], f[3], m1w, m2w, H12, H21)]
    loss_b, _ = F.bihome_loss(f[0].float().cuda(), f[1].float().cuda(), a[0], a[1], a[2], a[3], a[4], a[5], 0.01)
    g1 = torch.autograd.grad(loss_b.sum(), a)
    b = [mk(t) for t in (f[2], f[3], m1w, m2w, H12, H21)]
    loss_b, _ = F.bihome_loss(f[0].float().cuda(), f[1].float().cuda(), b[0], b[1], b[2], b[3], b[4], b[5], 0.01)
    wts = torch.tensor([2.0, 1.0, -0.5], device='cuda')
    g2 = torch.autograd.grad((loss_b * wts).sum(), b)
    for x, y in zip(g1, g2):
        shape = [3] + [1] * (x.dim() - 1)
        assert torch.allclose(x * wts.view(shape), y, rtol=1e-6, atol=1e-7)


def test_mace(F):
    a, b = torch.randn(9, 4, 2), torch.randn(9, 4, 2)
    assert abs(F.mace(a.cuda(), b.cuda()).item() - R().mace(a.numpy(), b.numpy())) < 1e-6


def test_feature_warp_fuzz_shapes(F):
    """channels-last feature warp (warp-cooperative kernels, the thread-per-quad kernels for other channel counts, the generic
    kernel for C % 4 != 0) on ragged shapes: 24 deterministic draws, forward and dH against the float64 closed form.
    Tolerances 1e-4 / 5e-4: a 5 x 4 output of a white-noise image has no averaging over pixels, and the float32 coordinate of
    a single far-away tap (2e-6 px per unit of |u|) shows through; a wiring error would be O(1)."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=24, deadline=None, derandomize=True)
    @given(B=st.integers(1, 3), C=st.sampled_from([4, 8, 12, 16, 24, 32, 64, 128]), Hs=st.integers(5, 40), Ws=st.integers(5, 40),
           Ho=st.integers(1, 37), Wo=st.integers(1, 37))
    def run(B, C, Hs, Ws, Ho, Wo):
        gen = torch.Generator().manual_seed(B * 7919 + C * 31 + Hs * 13 + Wo)
        img = torch.rand(B, C, Hs, Ws, generator=gen, dtype=torch.float64).float().double()
        H = _rand_h(B, min(Hs, Ws), gen).float().double()
        gO = torch.randn(B, C, Ho, Wo, generator=gen, dtype=torch.float64).float().double()
        cells = _kernel_cells(H.numpy().astype(np.float32), Ho, Wo)
        H64 = H.clone().requires_grad_(True)
        ref = _warp_direct_autograd(img, H64, Ho, Wo, cells)
        gH64, = torch.autograd.grad((ref * gO).sum(), H64)
        x = img.float().cuda().contiguous(memory_format=torch.channels_last)
        Hc = H.float().cuda().requires_grad_(True)
        out = F.warp(x, Hc, Ho, Wo)
        assert rel_l2(out.detach().cpu().numpy(), ref.detach().numpy()) < 1e-4
        gH, = torch.autograd.grad((out * gO.float().cuda()).sum(), Hc)
        assert rel_l2(gH.cpu().numpy(), gH64.numpy()) < 5e-4
    run()
