"""BASELINE.json configs[3] (the micro-benchmark sweep: batch 64-4096, patch 128-512, C = 64-256): parity at the far
corners of the sweep, by the same two means as tests/test_gpu_fullsize.py -- the float64 oracle on a few samples of the
full-size launch, and size-independent properties (identity and integer translation exact, linearity, agreement of
the channels-last and the planar kernels on the same data, run-to-run bit reproducibility)."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def F():
    assert torch.cuda.is_available()
    import bihome_b200.functional as fn
    return fn


def _smooth(n, c, size, gen):
    lo = torch.rand(n, c, size // 8 + 2, size // 8 + 2, generator=gen)
    return torch.nn.functional.interpolate(lo, size=(size, size), mode='bicubic', align_corners=True)


def test_warp_512_patches(F):
    """P = 512 (64x64 output blocks with staged source boxes), B = 16 pairs -> 32 planes"""
    from oracle import ref_path as R
    n, P = 32, 512
    gen = torch.Generator().manual_seed(21)
    img = _smooth(n, 1, P, gen)
    x = img.cuda()
    zero = torch.zeros(n, 4, 2, device='cuda')
    assert torch.equal(F.warp(x, F.dlt4(zero, size=(P, P)), P, P), x)
    tx, ty = -7, 11
    shift = zero.clone()
    shift[..., 0] += tx
    shift[..., 1] += ty
    out = F.warp(x, F.dlt4(shift, size=(P, P)), P, P)
    ref = torch.zeros_like(x)
    ref[:, :, :P - ty, -tx:] = x[:, :, ty:, :P + tx]
    assert (out - ref).abs().max().item() < 1e-4
    delta = (torch.rand(n, 4, 2, generator=gen) * 2 - 1) * 128
    gout = torch.randn(n, 1, P, P, generator=gen)
    d = delta.cuda().requires_grad_(True)
    H = F.dlt4(d, size=(P, P))
    out, mask = F.warp(x, H, P, P, pool=4)
    gd, = torch.autograd.grad((out * gout.cuda()).sum(), d)
    for i in (0, 13, n - 1):
        d64 = delta[i:i + 1].double().requires_grad_(True)
        H64 = R.four_point_to_homography(R.image_shape_to_corners(img[i:i + 1].double()), d64)
        o64 = R.warp_direct(img[i:i + 1].double(), H64, P, P)
        m64 = torch.nn.functional.avg_pool2d(R.analytic_mask(H64, P, P, P, P), 4)[:, 0]
        assert rel_l2(out[i].detach().cpu().numpy(), o64[0].detach().numpy()) < 2e-5, i      # ulp of a coordinate near 512
        assert rel_l2(mask[i].detach().cpu().numpy(), m64[0].detach().numpy()) < 2e-5, i
        g64, = torch.autograd.grad((o64 * gout[i:i + 1].double()).sum(), d64)
        assert rel_l2(gd[i].cpu().numpy(), g64[0].numpy()) < 5e-3, i
    y = torch.rand(n, 1, P, P, generator=gen).cuda()
    a = F.warp(2.0 * x + 3.0 * y, H.detach(), P, P)
    b = 2.0 * F.warp(x, H.detach(), P, P) + 3.0 * F.warp(y, H.detach(), P, P)
    assert rel_l2(a.cpu().numpy(), b.cpu().numpy()) < 1e-6
    assert torch.equal(F.warp(x, H.detach(), P, P), out.detach())


def test_feature_warp_channels_last_agrees_with_planar(F):
    """the C-channel NHWC warp of the sweep (C = 64, P = 128) against the planar ring kernel on the same data"""
    B, C, P = 24, 64, 128
    gen = torch.Generator().manual_seed(22)
    feat = _smooth(B, C, P, gen).cuda()        # smooth maps: d out/dH is then continuous across bilinear cells to 1e-6
    delta = ((torch.rand(B, 4, 2, generator=gen) * 2 - 1) * 32).cuda()
    gout = torch.randn(B, C, P, P, generator=gen).cuda()
    res = []
    for nhwc in (False, True):
        x = feat.contiguous(memory_format=torch.channels_last) if nhwc else feat
        H = F.dlt4(delta, size=(P, P)).detach().requires_grad_(True)
        out = F.warp(x, H, P, P)
        assert out.is_contiguous(memory_format=torch.channels_last if nhwc else torch.contiguous_format)
        gH, = torch.autograd.grad((out * gout).sum(), H)
        res.append((out.detach(), gH))
    assert rel_l2(res[1][0].cpu().numpy(), res[0][0].cpu().numpy()) < 1e-6
    assert rel_l2(res[1][1].cpu().numpy(), res[0][1].cpu().numpy()) < 1e-4
    ident = F.dlt4(torch.zeros(B, 4, 2, device='cuda'), size=(P, P))
    assert torch.equal(F.warp(feat.contiguous(memory_format=torch.channels_last), ident, P, P), feat)


@pytest.mark.parametrize('B,C,h', [(256, 256, 32), (64, 64, 128), (1024, 64, 32)])
def test_loss_sweep_shapes(F, B, C, h):
    """channels-last loss at the sweep's widest (C = 256), largest-map (P = 512 -> h = 128) and B = 1024 shapes"""
    from oracle import ref_path as R
    mu = 0.01
    gen = torch.Generator().manual_seed(B + C + h)
    mk = lambda: torch.relu(torch.randn(B, C, h, h, generator=gen))
    f1, f2, f1w, f2w = mk(), mk(), mk(), mk()
    m1w, m2w = torch.rand(B, h, h, generator=gen), torch.rand(B, h, h, generator=gen)
    H12 = torch.eye(3) + 0.05 * torch.randn(B, 3, 3, generator=gen)
    H21 = torch.eye(3) + 0.05 * torch.randn(B, 3, 3, generator=gen)
    cl = lambda t: t.cuda().contiguous(memory_format=torch.channels_last)
    a = [cl(f1), cl(f2), cl(f1w).requires_grad_(True), cl(f2w).requires_grad_(True), m1w.cuda().requires_grad_(True),
         m2w.cuda().requires_grad_(True), H12.cuda().requires_grad_(True), H21.cuda().requires_grad_(True)]
    loss_b, parts = F.bihome_loss(*a, mu)
    g = torch.autograd.grad(loss_b.sum(), a[2:])
    ones = torch.ones(1, 1, h, h, dtype=torch.float64)
    for i in (0, B // 2 + 1, B - 1):
        lv = [t[i:i + 1].double().requires_grad_(True) for t in (f1w, f2w, m1w, m2w, H12, H21)]
        ref, _ = R.bihome_double_line(f1[i:i + 1].double(), f2[i:i + 1].double(), lv[0], lv[1], ones, ones,
                                      lv[2].unsqueeze(1), lv[3].unsqueeze(1), lv[4], lv[5], mu)
        g64 = torch.autograd.grad(ref, lv)
        assert abs(loss_b[i].item() - ref.item()) < 1e-5 * abs(ref.item()) + 1e-6, i
        for k in range(6):
            assert rel_l2(g[k][i].cpu().numpy(), g64[k][0].numpy()) < 1e-5, (i, k)
    again, _ = F.bihome_loss(*[t.detach() for t in a], mu)
    assert torch.equal(again, loss_b.detach())
