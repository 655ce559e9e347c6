"""Parity at BASELINE.json's full size (B = 256 pairs -> 512 planes of 128x128, C = 64, h = w = 32).

The float64 oracle is too slow for whole batches of this size, so the full-size runs are checked
 (a) against the oracle on a random SUBSET of the samples (the kernels' work decomposition -- persistent CTAs, half-items
     in the last round, tile ranges that straddle samples -- depends on the batch size, the arithmetic of a sample does not),
 (b) through size-independent properties: identity / integer translation are exact, the warp is linear in the image, the
     analytic pooled mask equals the pooled warp of ones, the loss is symmetric under exchanging the two directions and
     1-homogeneous in the features, and every result is bit-reproducible from run to run."""
import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu

B, P, C, MU = 256, 128, 64, 0.01


@pytest.fixture(scope='module')
def F():
    assert torch.cuda.is_available()
    import bihome_b200.functional as fn
    return fn


def _smooth(n, gen):
    lo = torch.rand(n, 1, 18, 18, generator=gen)
    return torch.nn.functional.interpolate(lo, size=(P, P), mode='bicubic', align_corners=True)


def test_warp_full_batch_matches_oracle_on_a_subset(F):
    from oracle import ref_path as R
    gen = torch.Generator().manual_seed(5)
    img = _smooth(2 * B, gen)
    delta = (torch.rand(2 * B, 4, 2, generator=gen) * 2 - 1) * 32
    gout = torch.randn(2 * B, 1, P, P, generator=gen)
    gmask = torch.randn(2 * B, P // 4, P // 4, generator=gen)
    d = delta.cuda().requires_grad_(True)
    H = F.dlt4(d, size=(P, P))
    out, mask = F.warp(img.cuda(), H, P, P, pool=4)
    gd, = torch.autograd.grad((out * gout.cuda()).sum() + (mask * gmask.cuda()).sum(), d)
    pick = torch.randperm(2 * B, generator=gen)[:6].tolist() + [0, 2 * B - 1, 443, 444, 511 - 67]   # incl. the half-item tail
    for i in pick:
        d64 = delta[i:i + 1].double().requires_grad_(True)
        o64, H64 = R.warp(img[i:i + 1].double(), d64)
        m64, _ = R.warp(torch.ones(1, 1, P, P, dtype=torch.float64), d64)
        mp64 = torch.nn.functional.avg_pool2d(m64, 4)[:, 0]
        assert rel_l2(H[i].detach().cpu().numpy(), H64[0].detach().numpy()) < 1e-5
        assert rel_l2(out[i].detach().cpu().numpy(), o64[0].detach().numpy()) < 1e-5, i
        assert rel_l2(mask[i].detach().cpu().numpy(), mp64[0].detach().numpy()) < 1e-5, i
        g64, = torch.autograd.grad((o64 * gout[i:i + 1].double()).sum() + (mp64 * gmask[i:i + 1].double()).sum(), d64)
        # smooth images: the bilinear-cell ambiguity of dH (test_gpu_kernels._kernel_cells) is below the tolerance
        assert rel_l2(gd[i].cpu().numpy(), g64[0].numpy()) < 2e-3, i


def test_warp_identity_translation_linearity_and_mask_consistency(F):
    gen = torch.Generator().manual_seed(6)
    x, y = torch.rand(2 * B, 1, P, P, generator=gen).cuda(), torch.rand(2 * B, 1, P, P, generator=gen).cuda()
    zero = torch.zeros(2 * B, 4, 2, device='cuda')
    out = F.warp(x, F.dlt4(zero, size=(P, P)), P, P)
    assert torch.equal(out, x)                                           # identity: bit exact
    # a pure integer translation of the corners by (tx, ty): out[y, x] = in[y + ty, x + tx], zeros outside
    tx, ty = 5, -3
    shift = zero.clone()
    shift[..., 0] += tx
    shift[..., 1] += ty
    out = F.warp(x, F.dlt4(shift, size=(P, P)), P, P)
    ref = torch.zeros_like(x)
    ref[:, :, -ty:, :P - tx] = x[:, :, :P + ty, tx:]
    assert (out - ref).abs().max().item() < 2e-5
    # linearity in the image on arbitrary projective H
    delta = ((torch.rand(2 * B, 4, 2, generator=gen) * 2 - 1) * 32).cuda()
    H = F.dlt4(delta, size=(P, P))
    a = F.warp(2.0 * x + 3.0 * y, H, P, P)
    b = 2.0 * F.warp(x, H, P, P) + 3.0 * F.warp(y, H, P, P)
    assert rel_l2(a.cpu().numpy(), b.cpu().numpy()) < 1e-6
    # the analytic pooled mask is the pooled warp of ones
    ones, mask = F.warp(torch.ones_like(x), H, P, P, pool=4)
    assert (torch.nn.functional.avg_pool2d(ones, 4)[:, 0] - mask).abs().max().item() < 1e-5
    # run-to-run bit reproducibility, forward and dH (fixed-order reductions, no atomics)
    Hg = H.detach().requires_grad_(True)
    g1, = torch.autograd.grad((F.warp(x, Hg, P, P) * y).sum(), Hg)
    g2, = torch.autograd.grad((F.warp(x, Hg, P, P) * y).sum(), Hg)
    assert torch.equal(g1, g2) and torch.equal(F.warp(x, H, P, P), F.warp(x, H, P, P))


@pytest.mark.parametrize('nhwc', [True, False])
def test_loss_full_batch_subset_symmetry_and_homogeneity(F, nhwc):
    from oracle import ref_path as R
    gen = torch.Generator().manual_seed(8)
    h = P // 4
    mk = lambda: torch.relu(torch.randn(B, C, h, h, generator=gen))
    f1, f2, f1w, f2w = mk(), mk(), mk(), mk()
    m1w, m2w = torch.rand(B, h, h, generator=gen), torch.rand(B, h, h, generator=gen)
    H12 = torch.eye(3) + 0.05 * torch.randn(B, 3, 3, generator=gen)
    H21 = torch.eye(3) + 0.05 * torch.randn(B, 3, 3, generator=gen)
    dev = lambda t, feat=False: (t.cuda().contiguous(memory_format=torch.channels_last) if (feat and nhwc) else t.cuda())
    a = [dev(f1, True), dev(f2, True), dev(f1w, True).requires_grad_(True), dev(f2w, True).requires_grad_(True),
         dev(m1w).requires_grad_(True), dev(m2w).requires_grad_(True), dev(H12).requires_grad_(True), dev(H21).requires_grad_(True)]
    loss_b, parts = F.bihome_loss(*a, MU)
    g = torch.autograd.grad(loss_b.sum(), a[2:])
    # (a) oracle on a subset
    ones = torch.ones(1, 1, h, h, dtype=torch.float64)
    for i in [0, 1, 37, 128, 255]:
        lv = [t[i:i + 1].double().requires_grad_(True) for t in (f1w, f2w, m1w, m2w, H12, H21)]
        ref, _ = R.bihome_double_line(f1[i:i + 1].double(), f2[i:i + 1].double(), lv[0], lv[1], ones, ones, lv[2].unsqueeze(1),
                                      lv[3].unsqueeze(1), lv[4], lv[5], MU)
        g64 = torch.autograd.grad(ref, lv)
        assert abs(loss_b[i].item() - ref.item()) < 1e-5 * abs(ref.item()) + 1e-6
        for k in range(6):
            assert rel_l2(g[k][i].cpu().numpy(), g64[k][0].numpy()) < 1e-5, (i, k)
    # (b) exchanging the two directions leaves the loss unchanged and exchanges the gradients
    s = [a[1], a[0], a[3].detach().requires_grad_(True), a[2].detach().requires_grad_(True), a[5].detach().requires_grad_(True),
         a[4].detach().requires_grad_(True), a[7].detach().requires_grad_(True), a[6].detach().requires_grad_(True)]
    loss_s, _ = F.bihome_loss(*s, MU)
    gs = torch.autograd.grad(loss_s.sum(), s[2:])
    # ln3 = ||H12 H21 - I||^2 is not symmetric in general (H12 H21 != H21 H12): compare ln1 + ln2
    _, parts_s = F.bihome_loss(*[t.detach() for t in s], MU)
    assert rel_l2((parts_s[:, 0] + parts_s[:, 1]).cpu().numpy(), (parts[:, 0] + parts[:, 1]).detach().cpu().numpy()) < 1e-6
    assert rel_l2(gs[0].cpu().numpy(), g[1].cpu().numpy()) < 1e-6 and rel_l2(gs[1].cpu().numpy(), g[0].cpu().numpy()) < 1e-6
    # (c) 1-homogeneous in the features: scaling all four by 4 (exact in binary) scales ln1, ln2 by 4, feature gradients by 1
    b4 = [4.0 * a[0], 4.0 * a[1], (4.0 * a[2]).detach().requires_grad_(True), (4.0 * a[3]).detach().requires_grad_(True)] + \
         [t.detach() for t in a[4:]]
    loss4, parts4 = F.bihome_loss(*b4, MU)
    g4 = torch.autograd.grad(loss4.sum(), b4[2:4])
    assert torch.equal(parts4[:, :2], 4.0 * parts[:, :2].detach())
    assert torch.equal(g4[0], g[0]) and torch.equal(g4[1], g[1])
    # (d) bit reproducible
    loss_r, _ = F.bihome_loss(*[t.detach() for t in a], MU)
    assert torch.equal(loss_r, loss_b.detach())
