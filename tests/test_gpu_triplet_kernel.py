"""GPU parity of K3g (bh_triplet_fwd_bwd): every variant of the reference's masked triplet loss beside the north-star one,
through the C ABI, against the float64 oracle (oracle/ref_path.py triplet_general, which restates
src/heads/PerceptualHead.py:465-665 and src/heads/TripletHead.py:78-153) on the same seeded inputs -- loss, parts and the
gradients with respect to all four features, all four masks and both homographies.

Tolerance: 1e-5 norm-wise relative (BASELINE.json north_star).  The hinged variants are discontinuous where a hinge argument
crosses zero; the inputs are drawn so that no argument lies within 1e-4 of its hinge in the float64 evaluation (the draw is
a deterministic search over seeds), so that float32 rounding cannot take the other branch."""
import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(scope='module')
def F():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    import bihome_b200.functional as fn
    return fn


def R():
    from oracle import ref_path
    return ref_path


def _hinge_arguments(f, distance, hinge, margin):
    """the quantities whose sign the hinges test, float64"""
    f1, f2, f1w, f2w = f
    if distance == 'l1':
        d = lambda x, y: (x - y).abs()
    elif distance == 'l2':
        d = lambda x, y: ((x - y) ** 2).mean(1)
    else:
        d = lambda x, y: 1 - torch.cosine_similarity(x, y, dim=1)
    l3 = d(f1, f2)
    out = []
    for x, y in ((f1w, f2), (f2w, f1)):
        la = d(x, y)
        if hinge == 'channel':
            out.append(la - l3 + margin)
        else:
            if la.dim() == 4:
                out.append(la.sum(1) - l3.sum(1) + margin)
            else:
                out.append(la - l3 + margin)
    return out


def _inputs(B, C, h, w, distance, hinge, margin, seed0):
    for seed in range(seed0, seed0 + 200):
        g = torch.Generator().manual_seed(seed)
        rnd = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64).float().double()
        base = rnd(B, C, h, w)
        f = [torch.relu(base + 0.6 * rnd(B, C, h, w)) + 0.05 for _ in range(4)]     # correlated, positive like post-ReLU maps
        masks = [torch.rand(B, h, w, generator=g, dtype=torch.float64).float().double() for _ in range(4)]
        masks[0][0] *= 0.0005      # sample 0: sum of weights below 1 -> the max(den, 1) branch
        H12 = (torch.eye(3, dtype=torch.float64) + 0.05 * rnd(B, 3, 3)).float().double()
        H21 = (torch.linalg.inv(H12) + 0.02 * rnd(B, 3, 3)).float().double()
        if hinge is None:
            return f, masks, H12, H21
        if all((v.abs() > 1e-4).all() for v in _hinge_arguments(f, distance, hinge, margin)):
            return f, masks, H12, H21
    raise AssertionError('no safe seed found')


CASES = [
    # lines, distance, hinge, margin, crd
    (2, 'l1', None, 0.0, False),
    (2, 'l1', 'channel', 0.05, False),
    (2, 'l1', 'pixel', 1.0, False),
    (2, 'l2', None, 0.0, False),
    (2, 'l2', 'pixel', 0.02, False),
    (2, 'cosine', None, 0.0, False),
    (2, 'cosine', 'pixel', 0.01, False),
    (1, 'l1', 'pixel', 1.0, False),
    (1, 'l1', 'pixel', 0.5, True),
    (1, 'cosine', 'pixel', 0.02, True),
    (1, 'l1', 'channel', 0.1, False),
    (1, 'l1', None, 0.0, False),
]
SHAPES = [
    # B, C, h, w, channels_last
    (3, 64, 8, 8, True),       # 16 lanes per pixel, features in registers
    (3, 64, 8, 8, False),      # planar, four pixels per thread, two walks over the channels for the gated variants
    (2, 256, 4, 4, True),      # two float4 per lane and tensor
    (4, 1, 32, 32, False),     # TripletHead: one-channel full-resolution maps
    (2, 6, 5, 7, False),       # odd sizes: one pixel per thread
    (2, 6, 5, 7, True),        # channels-last without a power-of-two channel count: strided walk
]


@pytest.mark.parametrize('B,C,h,w,nhwc', SHAPES)
@pytest.mark.parametrize('lines,distance,hinge,margin,crd', CASES)
def test_triplet_loss_vs_oracle(F, lines, distance, hinge, margin, crd, B, C, h, w, nhwc):
    if C == 1 and distance == 'cosine':
        pytest.skip('cosine similarity of one-channel maps is +-1: no gradient to compare')
    _compare_with_oracle(F, lines, distance, hinge, margin, crd, B, C, h, w, nhwc, TOL)


def _compare_with_oracle(F, lines, distance, hinge, margin, crd, B, C, h, w, nhwc, TOL):
    mu = 0.01
    scale = (1.0, 1.0) if hinge != 'pixel' else (float(B), 0.5)
    f, masks, H12, H21 = _inputs(B, C, h, w, distance, hinge, margin, seed0=1000 * lines + 10 * C + h)
    two = lines == 2
    leaves64 = [t.clone().requires_grad_(True) for t in (*f, *masks, H12, H21)]
    f64, m64 = leaves64[:4], leaves64[4:8]
    loss64, p64 = R().triplet_general(f64[0], f64[1], f64[2], f64[3] if two else None, m64[0], m64[1], m64[2] if two else None,
                                      m64[3] if two else None, leaves64[8], leaves64[9], lines, distance, hinge, margin, mask_crd=crd,
                                      mu=mu, scale=scale)
    g64 = torch.autograd.grad(loss64.sum(), leaves64, allow_unused=True)

    def dev(t, feat):
        t = t.float().cuda()
        if feat and nhwc:
            t = t.contiguous(memory_format=torch.channels_last)
        return t.requires_grad_(True)
    leaves = [dev(t, i < 4) for i, t in enumerate((*f, *masks, H12, H21))]
    loss_b, parts = F.triplet_loss(leaves[0], leaves[1], leaves[2], leaves[3] if two else None, leaves[4], leaves[5],
                                   leaves[6] if two else None, leaves[7] if two else None, leaves[8] if two else None,
                                   leaves[9] if two else None, lines=lines, distance=distance, hinge=hinge, margin=margin,
                                   mask_crd=crd, mu=mu, scale=scale)
    assert rel_l2(loss_b.detach().cpu().numpy(), loss64.detach().numpy()) < TOL
    assert rel_l2(parts[:, 0].cpu().numpy(), p64['ln1'].detach().numpy()) < TOL
    assert rel_l2(parts[:, 2].cpu().numpy(), p64['den1'].detach().numpy()) < TOL
    if two:
        assert rel_l2(parts[:, 1].cpu().numpy(), p64['ln2'].detach().numpy()) < TOL
        assert rel_l2(parts[:, 3].cpu().numpy(), p64['den2'].detach().numpy()) < TOL
        assert rel_l2(parts[:, 4].cpu().numpy(), p64['ln3'].detach().numpy()) < TOL
    used = [i for i in range(10) if g64[i] is not None and float(g64[i].abs().max()) > 0]
    if not used:          # every hinge closed (tiny maps): nothing carries a gradient
        return
    g = torch.autograd.grad(loss_b.sum(), [leaves[i] for i in used], allow_unused=True)
    names = ['f1', 'f2', 'f1w', 'f2w', 'a1', 'b2', 'a2', 'b1', 'H12', 'H21']
    for gi, i in zip(g, used):
        assert gi is not None, names[i]
        assert rel_l2(gi.cpu().numpy(), g64[i].numpy()) < TOL, 'gradient of %s' % names[i]


def test_triplet_loss_upstream_scale_and_optional_gradients(F):
    """a non-unit upstream gradient goes through bh_triplet_rescale; inputs that need no gradient get none"""
    f, masks, H12, H21 = _inputs(3, 8, 8, 8, 'l1', 'pixel', 0.3, seed0=5)
    mk = lambda t, rg=True: t.float().cuda().requires_grad_(rg)

    def run(weights):
        a = [mk(f[0], False), mk(f[1], False), mk(f[2]), mk(f[3]), mk(masks[0]), mk(masks[1], False), mk(masks[2]), mk(masks[3], False),
             mk(H12), mk(H21)]
        loss_b, _ = F.triplet_loss(*a, lines=2, distance='l1', hinge='pixel', margin=0.3, mu=0.01)
        wanted = [a[i] for i in (2, 3, 4, 6, 8, 9)]
        return torch.autograd.grad((loss_b * weights).sum(), wanted)
    ones = torch.ones(3, device='cuda')
    wts = torch.tensor([2.0, 1.0, -0.5], device='cuda')
    for x, y in zip(run(ones), run(wts)):
        shape = [3] + [1] * (x.dim() - 1)
        assert torch.allclose(x * wts.view(shape), y, rtol=1e-6, atol=1e-7)


def test_triplet_loss_rejects_cpu_tensors_and_bad_variants(F):
    t = torch.zeros(1, 4, 4, 4)
    with pytest.raises(RuntimeError):
        F.triplet_loss(t, t, t, t, t[:, 0], None)
    c = t.cuda()
    with pytest.raises(ValueError):
        F.triplet_loss(c, c, c, c, c[:, 0], None, distance='l3')
    with pytest.raises(ValueError):
        F.triplet_loss(c, c, c, c, c[:, 0], None, distance='cosine', hinge='channel', margin=0.1)


def test_triplet_loss_full_size_properties(F):
    """B = 256, C = 64, 32 x 32 channels-last (the north-star feature shape) -- size-independent properties instead of the
    oracle: (i) hinge None / l1 equals K3 (bihome_loss) on the same inputs; (ii) a margin so large that every hinge is open
    equals the un-hinged loss plus margin * sum(W) / den; (iii) swapping the two directions swaps ln1 and ln2."""
    B, C, h, w = 256, 64, 32, 32
    g = torch.Generator(device='cuda').manual_seed(3)
    cl = lambda: torch.relu(torch.randn(B, C, h, w, device='cuda', generator=g)).contiguous(memory_format=torch.channels_last)
    f1, f2, f1w, f2w = cl(), cl(), cl(), cl()
    a1, a2 = torch.rand(B, h, w, device='cuda', generator=g), torch.rand(B, h, w, device='cuda', generator=g)
    H12 = torch.eye(3, device='cuda').repeat(B, 1, 1) + 0.01 * torch.randn(B, 3, 3, device='cuda', generator=g)
    H21 = torch.linalg.inv(H12)
    ref_b, ref_parts = F.bihome_loss(f1, f2, f1w, f2w, a1, a2, H12, H21, 0.01)
    x = f1w.clone().requires_grad_(True)
    loss_b, parts = F.triplet_loss(f1, f2, x, f2w, a1, None, a2, None, H12, H21, lines=2, distance='l1', hinge=None, mu=0.01)
    assert torch.allclose(loss_b, ref_b, rtol=2e-5, atol=1e-5)
    assert torch.allclose(parts[:, 2:4], ref_parts[:, 2:4], rtol=1e-6)
    y = f1w.clone().requires_grad_(True)
    ref2, _ = F.bihome_loss(f1, f2, y, f2w, a1, a2, H12, H21, 0.01)
    gx, = torch.autograd.grad(loss_b.sum(), x)
    gy, = torch.autograd.grad(ref2.sum(), y)
    assert torch.allclose(gx, gy, rtol=1e-5, atol=1e-9)
    big = 1.0e4
    open_b, open_parts = F.triplet_loss(f1, f2, f1w, f2w, a1, None, a2, None, H12, H21, lines=2, distance='l1', hinge='pixel',
                                        margin=big, mu=0.01)
    # every hinge open: ln = (N + margin * S) / max(S, 1); S > 1 for these masks
    assert torch.allclose(open_parts[:, 0], parts[:, 0] + big, rtol=1e-5)
    assert torch.allclose(open_parts[:, 1], parts[:, 1] + big, rtol=1e-5)
    swap_b, swap_parts = F.triplet_loss(f2, f1, f2w, f1w, a2, None, a1, None, H21, H12, lines=2, distance='l1', hinge=None, mu=0.01)
    assert torch.allclose(swap_parts[:, 0], parts[:, 1], rtol=1e-6, atol=1e-6) and torch.allclose(swap_parts[:, 1], parts[:, 0], rtol=1e-6, atol=1e-6)


def test_triplet_loss_fuzz_shapes(F):
    """ragged shapes the fixed cases do not reach: channel counts that are not multiples of four, pixel counts that are not
    multiples of four or of the tile sizes, one-pixel maps, both layouts -- 30 deterministic draws against the oracle.
    Tolerance 1e-4: a map of a few pixels has no averaging, a loss that is a difference of nearly equal cosines (or a gradient
    left over after two weights cancel) carries the float32 rounding of its terms."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=30, deadline=None, derandomize=True)
    @given(B=st.integers(1, 4), C=st.sampled_from([1, 2, 3, 4, 5, 8, 12, 16, 20, 32]), h=st.integers(1, 9), w=st.integers(1, 9),
           nhwc=st.booleans(), case=st.sampled_from(CASES))
    def run(B, C, h, w, nhwc, case):
        lines, distance, hinge, margin, crd = case
        if distance == 'cosine' and C == 1:
            return
        _compare_with_oracle(F, lines, distance, hinge, margin, crd, B, C, h, w, nhwc, 1e-4)
    run()
