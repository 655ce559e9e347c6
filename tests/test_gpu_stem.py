"""K7 (csrc/stem.cu) on the device: BatchNorm2d (batch statistics) -> ReLU -> MaxPool2d(3, 2, 1) as one channels-last stage
against the three ATen modules it replaces (reference src/heads/PerceptualHead.py:56-58, src/backbones/Rethinking.py:31-36),
evaluated in float64 on the same inputs: output, running statistics, num_batches_tracked, input / weight / bias gradients.

ReLU and the pooling maximum are discontinuous in their gradients: an input whose pre-activation is within round-off of zero,
or a window whose two largest values are within round-off of each other, may route its gradient differently in two correct
float32 evaluations.  The exact-parity cases therefore build inputs on a lattice -- per channel a random permutation of
equidistant levels, the bias chosen so that zero sits midway between two levels -- so that every decision has a margin of
>= 1e-4 and the comparison can be element-wise; the north-star sized case uses Gaussian data and bounds the few knife-edge
elements instead."""
import math

import numpy as np
import pytest
import torch

TOL = 1e-5


def _modules(C, affine=True, track=True, momentum=0.1, seed=0):
    torch.manual_seed(seed)
    bn = torch.nn.BatchNorm2d(C, affine=affine, track_running_stats=track, momentum=momentum)
    if track:
        with torch.no_grad():
            bn.running_mean.normal_()
            bn.running_var.uniform_(0.5, 2.0)
    return bn, torch.nn.MaxPool2d(kernel_size=3, stride=2, padding=1)


def _lattice_input(N, C, H, W, bn, seed):
    """x [N,C,H,W] float32 whose channel c holds a random permutation of n = N*H*W equidistant levels (shifted and scaled per
    channel); bn.weight random with both signs, bn.bias such that the ReLU threshold lies midway between two levels"""
    rng = np.random.RandomState(seed)
    n = N * H * W
    x = np.empty((N, C, H, W), np.float32)
    for c in range(C):
        levels = ((np.arange(n) + 0.5) / n * 4.0 - 2.0) * rng.uniform(0.5, 3.0) + rng.uniform(-3.0, 3.0)
        x[:, c] = rng.permutation(levels).reshape(N, H, W).astype(np.float32)
    if bn.affine:
        xd = x.astype(np.float64)
        mean, var = xd.mean(axis=(0, 2, 3)), xd.var(axis=(0, 2, 3))
        gamma = rng.uniform(0.5, 1.5, C) * rng.choice([-1.0, 1.0], C)
        beta = np.empty(C)
        for c in range(C):
            lv = np.sort(xd[:, c].ravel())
            j = rng.randint(n // 4, 3 * n // 4)
            cross = 0.5 * (lv[j] + lv[j + 1])           # relu threshold between two levels
            beta[c] = -(cross - mean[c]) / math.sqrt(var[c] + bn.eps) * gamma[c]
        with torch.no_grad():
            bn.weight.copy_(torch.from_numpy(gamma).float())
            bn.bias.copy_(torch.from_numpy(beta).float())
    return torch.from_numpy(x)


def _reference(bn, pool, x, g):
    """the three modules in float64 (on the tensor's device) -> y, gx, gweight, gbias, running_mean, running_var, tracked"""
    import copy
    bn64 = copy.deepcopy(bn).double()
    bn64.train()
    x64 = x.double().requires_grad_(True)
    y = pool(torch.relu(bn64(x64)))
    (y * g.double()).sum().backward()
    return (y.detach(), x64.grad, bn64.weight.grad if bn64.affine else None, bn64.bias.grad if bn64.affine else None,
            bn64.running_mean, bn64.running_var, bn64.num_batches_tracked)


def _close(a, b, tol, what):
    a, b = a.double(), b.double()
    scale = float(b.abs().max().clamp_min(1e-30))
    err = float((a - b).abs().max()) / scale
    assert err <= tol, '%s: max error %.3e of the largest magnitude (%.3e)' % (what, err, scale)


@pytest.mark.gpu
@pytest.mark.parametrize('N,C,H,W', [(2, 8, 12, 12), (3, 64, 16, 20), (2, 16, 9, 7), (2, 4, 5, 6), (1, 128, 8, 8), (2, 256, 6, 6),
                                     (1, 1024, 4, 6), (8, 64, 32, 32), (1, 32, 2, 3)])
def test_stem_matches_the_modules(N, C, H, W):
    import bihome_b200.functional as F
    dev = torch.device('cuda', 0)
    bn, pool = _modules(C, seed=C + H)
    x = _lattice_input(N, C, H, W, bn, seed=N * 1000 + W).to(dev).contiguous(memory_format=torch.channels_last)
    bn = bn.to(dev).train()
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    g = torch.randn(N, C, Ho, Wo, generator=torch.Generator().manual_seed(5)).to(dev)
    ref = _reference(bn, pool, x, g)
    if C == 1 or not F.stem_supported(bn, pool, x):
        pytest.fail('K7 refuses a geometry it is documented for: %r' % ((N, C, H, W),))
    xs = x.clone().requires_grad_(True)
    y = F.stem(bn, xs)
    assert y.shape == (N, C, Ho, Wo) and y.is_contiguous(memory_format=torch.channels_last)
    (y * g).sum().backward()
    _close(y.detach(), ref[0], TOL, 'output')
    _close(bn.running_mean, ref[4], TOL, 'running_mean')
    _close(bn.running_var, ref[5], TOL, 'running_var')
    assert int(bn.num_batches_tracked) == int(ref[6])
    _close(xs.grad, ref[1], 5 * TOL, 'input gradient')
    _close(bn.weight.grad, ref[2], 5 * TOL, 'weight gradient')
    _close(bn.bias.grad, ref[3], 5 * TOL, 'bias gradient')


@pytest.mark.gpu
@pytest.mark.parametrize('affine,track', [(False, True), (True, False), (False, False)])
def test_stem_without_affine_or_running_statistics(affine, track):
    import bihome_b200.functional as F
    dev = torch.device('cuda', 0)
    N, C, H, W = 3, 16, 10, 14
    bn, pool = _modules(C, affine=affine, track=track, momentum=0.3)
    x = _lattice_input(N, C, H, W, bn, seed=11).to(dev).contiguous(memory_format=torch.channels_last)
    if not affine:   # no bias to place the threshold: shift the data instead so that zero is not a level
        x = x + 1e-3
    bn = bn.to(dev).train()
    g = torch.randn(N, C, 5, 7, generator=torch.Generator().manual_seed(6)).to(dev)
    y64, gx64, gw64, gb64, rm64, rv64, _ = _reference(bn, pool, x, g)
    assert F.stem_supported(bn, pool, x)
    xs = x.clone().requires_grad_(True)
    y = F.stem(bn, xs)
    (y * g).sum().backward()
    _close(y.detach(), y64, TOL, 'output')
    _close(xs.grad, gx64, 5 * TOL, 'input gradient')
    if track:
        _close(bn.running_mean, rm64, TOL, 'running_mean')
        _close(bn.running_var, rv64, TOL, 'running_var')
    if affine:
        _close(bn.weight.grad, gw64, 5 * TOL, 'weight gradient')
        _close(bn.bias.grad, gb64, 5 * TOL, 'bias gradient')


@pytest.mark.gpu
def test_stem_frozen_and_no_grad_paths():
    """the frozen extractor: parameters without gradient (input gradient only), and the no_grad passes (no codes kept)"""
    import bihome_b200.functional as F
    dev = torch.device('cuda', 0)
    N, C, H, W = 4, 64, 16, 16
    bn, pool = _modules(C)
    x = _lattice_input(N, C, H, W, bn, seed=3).to(dev).contiguous(memory_format=torch.channels_last)
    bn = bn.to(dev).train()
    for p in bn.parameters():
        p.requires_grad = False
    g = torch.randn(N, C, 8, 8, generator=torch.Generator().manual_seed(7)).to(dev)
    y64, gx64 = _reference(bn, pool, x, g)[:2]
    xs = x.clone().requires_grad_(True)
    y = F.stem(bn, xs)
    (y * g).sum().backward()
    _close(y.detach(), y64, TOL, 'output')
    _close(xs.grad, gx64, 5 * TOL, 'input gradient')
    assert bn.weight.grad is None and bn.bias.grad is None
    with torch.no_grad():
        y2 = F.stem(bn, x)
    _close(y2, y64, TOL, 'output under no_grad')
    assert not y2.requires_grad


@pytest.mark.gpu
def test_stem_refuses_what_it_does_not_cover():
    import bihome_b200.functional as F
    dev = torch.device('cuda', 0)
    bn, pool = _modules(64)
    bn = bn.to(dev)
    x = torch.randn(2, 64, 8, 8, device=dev)
    assert not F.stem_supported(bn, pool, x)                                            # NCHW
    xl = x.contiguous(memory_format=torch.channels_last)
    assert F.stem_supported(bn, pool, xl)
    assert not F.stem_supported(bn.eval(), pool, xl)                                    # running statistics: ATen modules
    bn.train()
    assert not F.stem_supported(bn, torch.nn.MaxPool2d(2, 2), xl)
    assert not F.stem_supported(bn, torch.nn.MaxPool2d(3, 2, 1, ceil_mode=True), xl)
    bn48, _ = _modules(48)
    assert not F.stem_supported(bn48.to(dev), pool, torch.randn(2, 48, 8, 8, device=dev).contiguous(memory_format=torch.channels_last))
    bn_cma = torch.nn.BatchNorm2d(64, momentum=None).to(dev)
    assert not F.stem_supported(bn_cma, pool, xl)                                       # cumulative average: ATen modules


@pytest.mark.gpu
def test_stem_north_star_size_and_speed():
    """[256,64,64,64] (the stem of the B = 256 step): output and statistics element-wise; gradients with Gaussian data differ
    from a float64 evaluation only at knife-edge decisions -- at most a handful of the 67 M elements -- and the stage beats the
    three ATen modules it replaces (forward + backward) by a wide margin"""
    import bihome_b200.functional as F
    dev = torch.device('cuda', 0)
    N, C, H, W = 256, 64, 64, 64
    bn, pool = _modules(C)
    bn = bn.to(dev).train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.3)
    gen = torch.Generator(device=dev).manual_seed(1)
    x = (torch.randn(N, C, H, W, device=dev, generator=gen) * 1.7 + 0.4).contiguous(memory_format=torch.channels_last)
    g = torch.randn(N, C, 32, 32, device=dev, generator=gen).contiguous(memory_format=torch.channels_last)
    import copy
    bn_ref = copy.deepcopy(bn)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya = F.stem(bn, xa)
    (ya * g).sum().backward()
    yb = pool(torch.relu(bn_ref(xb)))
    (yb * g).sum().backward()
    _close(ya.detach(), yb.detach(), 2e-5, 'output')
    _close(bn.running_mean, bn_ref.running_mean, 1e-5, 'running_mean')
    _close(bn.running_var, bn_ref.running_var, 1e-5, 'running_var')
    scale = float(xb.grad.abs().max())
    bad = int(((xa.grad - xb.grad).abs() > 1e-4 * scale).sum())
    assert bad <= 2000, '%d of %d input-gradient elements differ' % (bad, xa.grad.numel())
    werr = float((bn.weight.grad - bn_ref.weight.grad).abs().max() / bn_ref.weight.grad.abs().max())
    berr = float((bn.bias.grad - bn_ref.bias.grad).abs().max() / bn_ref.bias.grad.abs().max())
    assert werr < 2e-2 and berr < 2e-2, (werr, berr)      # ~1 knife-edge element per channel against a sum of ~500

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 5

    def run_fused():
        xs = x.detach().requires_grad_(True)
        (F.stem(bn, xs) * g).sum().backward()

    def run_aten():
        xs = x.detach().requires_grad_(True)
        (pool(torch.relu(bn_ref(xs))) * g).sum().backward()
    t_fused, t_aten = timed(run_fused), timed(run_aten)
    print('K7 stem fwd+bwd at [256,64,64,64]: fused %.3f ms, ATen modules %.3f ms' % (t_fused, t_aten))
    assert t_fused < 0.7 * t_aten, (t_fused, t_aten)


# ---- K7b: BatchNorm2d [+ residual] -> ReLU ------------------------------------------------------------------------------
def _lattice_residual(x, bn, seed):
    """r on the lattice of z = bn(x): integer multiples (-3 .. 3) of the per-channel level spacing of z, so that z + r keeps
    the margin of half a level to the ReLU threshold"""
    xd = x.double()
    n = xd.numel() // xd.shape[1]
    flat = xd.permute(1, 0, 2, 3).reshape(xd.shape[1], -1)
    dx = (flat.max(1).values - flat.min(1).values) / (n - 1)
    sigma = torch.sqrt(flat.var(1, unbiased=False) + bn.eps)
    sz = bn.weight.detach().double().cpu().abs() * dx / sigma
    k = torch.randint(-3, 4, x.shape, generator=torch.Generator().manual_seed(seed)).double()
    return (k * sz.view(1, -1, 1, 1)).float()


@pytest.mark.gpu
@pytest.mark.parametrize('residual', [False, True])
@pytest.mark.parametrize('N,C,H,W', [(2, 16, 8, 8), (3, 64, 16, 20), (2, 32, 9, 7), (1, 4, 5, 6), (1, 128, 8, 8), (2, 256, 6, 6),
                                     (1, 1024, 3, 2), (4, 64, 32, 32), (2, 16, 40, 24)])
def test_bn_relu_matches_the_modules(N, C, H, W, residual):
    import copy
    import bihome_b200.functional as F
    dev = torch.device('cuda', 0)
    bn, _ = _modules(C, seed=C + W)
    x = _lattice_input(N, C, H, W, bn, seed=N * 100 + H)
    r = _lattice_residual(x, bn, seed=9) if residual else None
    x = x.to(dev).contiguous(memory_format=torch.channels_last)
    bn = bn.to(dev).train()
    g = torch.randn(N, C, H, W, generator=torch.Generator().manual_seed(5)).to(dev)
    bn64 = copy.deepcopy(bn).double()
    x64 = x.double().requires_grad_(True)
    r64 = r.to(dev).double().requires_grad_(True) if residual else None
    y64 = torch.relu(bn64(x64) + r64) if residual else torch.relu(bn64(x64))
    (y64 * g.double()).sum().backward()
    rs = r.to(dev).contiguous(memory_format=torch.channels_last).requires_grad_(True) if residual else None
    assert F.bnact_supported(bn, x, rs)
    xs = x.clone().requires_grad_(True)
    y = F.bn_relu(bn, xs, residual=rs)
    assert y.shape == x.shape and y.stride() == x.stride()
    (y * g).sum().backward()
    _close(y.detach(), y64.detach(), TOL, 'output')
    _close(bn.running_mean, bn64.running_mean, TOL, 'running_mean')
    _close(bn.running_var, bn64.running_var, TOL, 'running_var')
    assert int(bn.num_batches_tracked) == int(bn64.num_batches_tracked)
    _close(xs.grad, x64.grad, 5 * TOL, 'input gradient')
    _close(bn.weight.grad, bn64.weight.grad, 5 * TOL, 'weight gradient')
    _close(bn.bias.grad, bn64.bias.grad, 5 * TOL, 'bias gradient')
    if residual:
        _close(rs.grad, r64.grad, TOL, 'residual gradient')


@pytest.mark.gpu
def test_residual_blocks_fused_vs_modules():
    """the backbone's residual blocks and the extractor's torchvision blocks with K7b / K8 on and off (BH_BNACT=aten,
    BH_CONVT_BIAS=aten): same output,
    same parameter gradients, same running statistics -- up to float32 round-off through two convolutions"""
    import copy
    import os
    import torchvision
    from bihome_b200.backbones import blocks
    from bihome_b200.heads.PerceptualHead import _run_layer
    dev = torch.device('cuda', 0)
    saved = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(3)
        cases = [(blocks.ResNet34IdentityBlock(64), (4, 64, 16, 16)), (blocks.ResNet34ConvBlock(64, 128, 2), (4, 64, 16, 16)),
                 (blocks.ResNet50DeconvBlock(32), (2, 32, 12, 12)), (torchvision.models.resnet34(weights=None).layer1, (4, 64, 16, 16))]
        for mod, shape in cases:
            fused = mod.to(dev).to(memory_format=torch.channels_last).train()
            plain = copy.deepcopy(fused)
            x = torch.randn(*shape, device=dev).contiguous(memory_format=torch.channels_last)
            run = (lambda m, t: _run_layer(m, t)) if isinstance(mod, torch.nn.Sequential) else (lambda m, t: m(t))
            xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
            ya = run(fused, xa)
            os.environ['BH_BNACT'] = 'aten'
            os.environ['BH_CONVT_BIAS'] = 'aten'
            try:
                yb = run(plain, xb)
            finally:
                del os.environ['BH_BNACT'], os.environ['BH_CONVT_BIAS']
            g = torch.randn_like(yb)
            (ya * g).sum().backward()
            (yb * g).sum().backward()
            from conftest import rel_l2
            assert rel_l2(ya.detach().cpu(), yb.detach().cpu()) < 1e-5
            assert rel_l2(xa.grad.cpu(), xb.grad.cpu()) < 1e-3          # a ReLU within round-off of zero may switch
            for (na, pa), (nb, pb) in zip(fused.named_parameters(), plain.named_parameters()):
                assert rel_l2(pa.grad.cpu(), pb.grad.cpu()) < 2e-3, na
            for (na, ba), (nb, bb) in zip(fused.named_buffers(), plain.named_buffers()):
                assert rel_l2(ba.double().cpu(), bb.double().cpu()) < 1e-5, na
    finally:
        torch.backends.cudnn.allow_tf32 = saved


@pytest.mark.gpu
def test_bn_relu_speed_at_backbone_shapes():
    """forward + backward of relu(bn(x)) and relu(bn(x) + r) against the ATen modules at three shapes of the B = 256 step"""
    import copy
    import bihome_b200.functional as F
    dev = torch.device('cuda', 0)

    def timed(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 5
    for shape in ((256, 64, 32, 32), (256, 128, 16, 16), (256, 32, 128, 128)):
        bn = torch.nn.BatchNorm2d(shape[1]).to(dev).train()
        ref = copy.deepcopy(bn)
        x = torch.randn(*shape, device=dev).contiguous(memory_format=torch.channels_last)
        r = torch.randn(*shape, device=dev).contiguous(memory_format=torch.channels_last)
        g = torch.randn(*shape, device=dev).contiguous(memory_format=torch.channels_last)
        for res in (False, True):
            def fused():
                xs, rs = x.detach().requires_grad_(True), r.detach().requires_grad_(True)
                (F.bn_relu(bn, xs, residual=rs if res else None) * g).sum().backward()

            def aten():
                xs, rs = x.detach().requires_grad_(True), r.detach().requires_grad_(True)
                (torch.relu(ref(xs) + rs if res else ref(xs)) * g).sum().backward()
            tf, ta = timed(fused), timed(aten)
            print('K7b %s residual=%s: fused %.3f ms, ATen %.3f ms' % (shape, res, tf, ta))


# ---- K7c: relu(bn_a(a) + bn_b(b)) ---------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize('N,C,H,W', [(2, 16, 8, 8), (3, 64, 16, 20), (2, 32, 9, 7), (1, 128, 8, 8), (4, 64, 32, 32), (2, 16, 40, 24)])
def test_bn_bn_relu_matches_the_modules(N, C, H, W):
    """a on the lattice of the K7b test; b takes seven integer levels and bn_b's weight / bias are chosen so that bn_b(b) is an
    integer multiple of the level spacing of bn_a(a): every ReLU decision keeps a margin of half a level"""
    import copy
    import bihome_b200.functional as F
    dev = torch.device('cuda', 0)
    bn_a, _ = _modules(C, seed=C + W)
    bn_b, _ = _modules(C, seed=C + W + 1, momentum=0.2)
    a = _lattice_input(N, C, H, W, bn_a, seed=N * 100 + H)
    n = N * H * W
    flat = a.double().permute(1, 0, 2, 3).reshape(C, -1)
    spacing = bn_a.weight.detach().double().abs() * ((flat.max(1).values - flat.min(1).values) / (n - 1)) / torch.sqrt(flat.var(1, unbiased=False) + bn_a.eps)
    k = torch.randint(-3, 4, (N, C, H, W), generator=torch.Generator().manual_seed(4)).double()
    b = (0.37 * k + 1.1).float()
    bflat = b.double().permute(1, 0, 2, 3).reshape(C, -1)
    sigma_b, mu_b = torch.sqrt(bflat.var(1, unbiased=False) + bn_b.eps), bflat.mean(1)
    gamma_b = spacing * sigma_b / 0.37                      # bn_b(b) = spacing * k + const
    beta_b = 2 * spacing + gamma_b * (mu_b - 1.1) / sigma_b   # const = 2 * spacing
    with torch.no_grad():
        bn_b.weight.copy_(gamma_b.float())
        bn_b.bias.copy_(beta_b.float())
    a = a.to(dev).contiguous(memory_format=torch.channels_last)
    b = b.to(dev).contiguous(memory_format=torch.channels_last)
    bn_a, bn_b = bn_a.to(dev).train(), bn_b.to(dev).train()
    g = torch.randn(N, C, H, W, generator=torch.Generator().manual_seed(5)).to(dev)
    ra, rb = copy.deepcopy(bn_a).double(), copy.deepcopy(bn_b).double()
    a64, b64 = a.double().requires_grad_(True), b.double().requires_grad_(True)
    y64 = torch.relu(ra(a64) + rb(b64))
    (y64 * g.double()).sum().backward()
    assert F.bnact_supported(bn_a, a) and F.bnact_supported(bn_b, b)
    xa, xb = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y = F.bn_bn_relu(bn_a, xa, bn_b, xb)
    (y * g).sum().backward()
    _close(y.detach(), y64.detach(), TOL, 'output')
    for mine, ref in ((bn_a, ra), (bn_b, rb)):
        _close(mine.running_mean, ref.running_mean, TOL, 'running_mean')
        _close(mine.running_var, ref.running_var, TOL, 'running_var')
        assert int(mine.num_batches_tracked) == int(ref.num_batches_tracked)
        _close(mine.weight.grad, ref.weight.grad, 5 * TOL, 'weight gradient')
        _close(mine.bias.grad, ref.bias.grad, 5 * TOL, 'bias gradient')
    _close(xa.grad, a64.grad, 5 * TOL, 'gradient of a')
    _close(xb.grad, b64.grad, 5 * TOL, 'gradient of b')


# ---- K8: the bias of a transposed convolution ---------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize('N,C,H,W', [(1, 4, 3, 5), (2, 8, 7, 9), (3, 32, 16, 20), (2, 64, 33, 31), (5, 256, 6, 6), (1, 1024, 4, 3),
                                     (8, 32, 128, 128), (16, 128, 32, 32)])
def test_channel_bias_kernels(N, C, H, W):
    """bh_bias_add / bh_bias_grad through the autograd function: the in-place add is ONE float32 addition per element, so it
    equals ATen's broadcast add bit for bit; the gradient is a float64 column sum"""
    import bihome_b200.functional as F
    dev = torch.device('cuda', 0)
    g = torch.Generator().manual_seed(N * 1000 + C)
    y0 = torch.randn(N, C, H, W, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
    bias = torch.randn(C, generator=g).to(dev).requires_grad_(True)
    gy = torch.randn(N, C, H, W, generator=g).to(dev).contiguous(memory_format=torch.channels_last)
    src = y0.clone().requires_grad_(True)
    y = F._ChannelBias.apply(src * 1.0, bias)            # a non-leaf, like a convolution's output
    assert y.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(y.detach(), y0 + bias.detach().view(1, -1, 1, 1))
    y.backward(gy)
    assert torch.equal(src.grad, gy)
    ref = gy.double().sum((0, 2, 3))
    _close(bias.grad, ref, 1e-6, 'bias gradient')
    # an upstream gradient that is not channels-last is laid out first
    bias.grad = None
    y = F._ChannelBias.apply(y0.clone(), bias)
    y.backward(gy.contiguous())
    _close(bias.grad, ref, 1e-6, 'bias gradient (NCHW upstream gradient)')


@pytest.mark.gpu
@pytest.mark.parametrize('N,cin,cout,H,W', [(2, 32, 32, 6, 6), (3, 64, 64, 5, 7), (2, 16, 8, 9, 4), (4, 256, 256, 4, 4)])
def test_conv_transpose_bias_matches_the_module(N, cin, cout, H, W):
    """F.conv_transpose_bias(m, x) against m(x) evaluated in float64: output, input / weight / bias gradients"""
    import copy
    import bihome_b200.functional as F
    dev = torch.device('cuda', 0)
    saved = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.manual_seed(cin + H)
        m = torch.nn.ConvTranspose2d(cin, cout, kernel_size=2, stride=2).to(dev).to(memory_format=torch.channels_last)
        with torch.no_grad():
            m.bias.normal_()
        m64 = copy.deepcopy(m).double()
        x = torch.randn(N, cin, H, W, device=dev).contiguous(memory_format=torch.channels_last)
        assert F.convt_bias_supported(m, x)
        xa, xb = x.clone().requires_grad_(True), x.double().requires_grad_(True)
        ya, yb = F.conv_transpose_bias(m, xa), m64(xb)
        g = torch.randn_like(ya)
        (ya * g).sum().backward()
        (yb * g.double()).sum().backward()
        _close(ya.detach(), yb.detach(), TOL, 'output')
        _close(xa.grad, xb.grad, TOL, 'input gradient')
        _close(m.weight.grad, m64.weight.grad, TOL, 'weight gradient')
        _close(m.bias.grad, m64.bias.grad, TOL, 'bias gradient')
        # frozen bias / no_grad: nothing to save, same output
        with torch.no_grad():
            _close(F.conv_transpose_bias(m, x), yb.detach(), TOL, 'output under no_grad')
    finally:
        torch.backends.cudnn.allow_tf32 = saved


@pytest.mark.gpu
def test_conv_transpose_bias_refuses_what_it_does_not_cover(monkeypatch):
    import bihome_b200.functional as F
    dev = torch.device('cuda', 0)
    x = torch.randn(2, 8, 4, 4, device=dev).contiguous(memory_format=torch.channels_last)
    assert F.convt_bias_supported(torch.nn.ConvTranspose2d(8, 8, 2, stride=2).to(dev), x)
    assert not F.convt_bias_supported(torch.nn.ConvTranspose2d(8, 8, 2, stride=2, bias=False).to(dev), x)
    assert not F.convt_bias_supported(torch.nn.ConvTranspose2d(8, 6, 2, stride=2).to(dev), x)       # 6 channels: not a power of two
    assert not F.convt_bias_supported(torch.nn.Conv2d(8, 8, 1).to(dev), x)
    assert not F.convt_bias_supported(torch.nn.ConvTranspose2d(8, 8, 2, stride=2), x.cpu())
    monkeypatch.setenv('BH_CONVT_BIAS', 'aten')
    assert not F.convt_bias_supported(torch.nn.ConvTranspose2d(8, 8, 2, stride=2).to(dev), x)
