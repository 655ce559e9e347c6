"""K6 on the device: every entry point of csrc/fieldhead.cu / fieldhead_mma.cu (scalar kernels, tensor-core kernels in their
float32-faithful and TF32 modes) against plain torch ops evaluated in float64 on the GPU, the whole stage and the whole Zeng
backbone against the ATen modules.  Companion of tests/test_field_head.py (host algebra, host emulation of the scalar kernels,
lane-level emulation of the tensor-core fragment mapping, ThreadSanitizer -- all on CPU).  K6 runs on a device only after its
self-test passed there (bihome_b200/autotune.py); these tests follow that verdict (BH_TEST_UNVERIFIED=1 forces them)."""
import copy
import os

import pytest
import torch

import cpu_kernels
from conftest import rel_l2
from test_field_head import compare_stage, make_stage, mute_knife_edge_pixels


@pytest.fixture
def k6_on_this_device():
    """K6 runs on a device only after its self-test (bihome_b200/autotune.py, in a child process) passed there; the GPU
    tests below follow that verdict, or BH_TEST_UNVERIFIED=1 forces them (bring-up on new hardware)"""
    if os.environ.get('BH_TEST_UNVERIFIED') == '1':
        return
    import bihome_b200.functional as F
    if not (torch.cuda.is_available() and F.field_head_enabled(torch.device('cuda', 0))):
        pytest.skip("K6 did not pass this device's self-test (or BH_FIELD_HEAD=aten): the backbone uses the ATen modules")


def tf32_rna(t):
    """round a float32 tensor to the nearest TF32 number, ties away from zero (PTX cvt.rna.tf32.f32), as float64"""
    bits = t.float().contiguous().view(torch.int32)
    return ((bits + 0x1000) & -8192).view(torch.float32).double()


@pytest.mark.gpu
@pytest.mark.parametrize('variant', ['auto', 'scalar', 'tf32'])
@pytest.mark.parametrize('B,H,W', [(1, 4, 8), (3, 9, 7), (2, 128, 128), (5, 33, 65), (7, 16, 24), (300, 8, 8)])
def test_device_entry_points_vs_float64(k6_on_this_device, B, H, W, variant):
    """'auto' runs the tensor-core kernels (csrc/fieldhead_mma.cu: mma.sync TF32 with float32 head + remainder operands)
    wherever H*W is a multiple of 32 and the scalar kernels elsewhere; 'scalar' forces the per-pixel kernels; 'tf32' is the
    tensor-core path with torch.backends.cudnn.allow_tf32 on (operands rounded to TF32, one product), compared with the
    float64 evaluation of the SAME rounded operands"""
    import bihome_b200.functional as F
    if variant == 'tf32' and (H * W) % 32:
        pytest.skip('the scalar kernels have no TF32 mode')
    saved = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = variant == 'tf32'
    F.tune('fieldhead_variant', 1 if variant == 'scalar' else 0)
    try:
        _entry_points_vs_float64(F, B, H, W, variant == 'tf32')
    finally:
        F.tune('fieldhead_variant', 0)
        torch.backends.cudnn.allow_tf32 = saved


def _entry_points_vs_float64(F, B, H, W, tf32):
    gen = torch.Generator().manual_seed(B * 100 + H)
    x = torch.relu(torch.randn(B, 16, H, W, generator=gen) + 0.3).cuda().contiguous(memory_format=torch.channels_last)
    W1 = (torch.randn(128, 16, generator=gen) * 0.3).cuda()
    b1 = torch.randn(128, generator=gen).cuda()
    W2 = (torch.randn(2, 128, generator=gen) * 0.2).cuda()
    b2 = torch.randn(2, generator=gen).cuda()
    g = torch.randn(B, 2, H, W, generator=gen).cuda()
    d = lambda t: t.double()
    rnd = tf32_rna if tf32 else d                          # what the contraction operands look like inside the kernel
    pre = torch.nn.functional.conv2d(rnd(x), rnd(W1).view(128, 16, 1, 1), d(b1))
    g = mute_knife_edge_pixels(g, pre, 1e-4)             # float32 kernel against float64 reference: see the helper
    s1, s2 = F._fh_moments(x)
    r1, r2 = cpu_kernels.fh_moments(x)
    assert rel_l2(s1.cpu().numpy(), r1.cpu().numpy()) < 1e-6 and rel_l2(s2.cpu().numpy(), r2.cpu().numpy()) < 1e-6
    out = F._fh_fwd(x, W1, b1, W2, b2)
    ref = cpu_kernels.fh_fwd(rnd(x), rnd(W1), d(b1), d(W2), d(b2))
    assert rel_l2(out.cpu().numpy(), ref.cpu().numpy()) < 1e-5
    got = F._fh_bwd(x, W1, b1, W2, g)
    if tf32:
        # every contraction with its operands rounded the way the kernels round them: gh and h (activations), x, W1, g
        X = rnd(x).permute(0, 2, 3, 1).reshape(-1, 16)
        P = X @ rnd(W1).t() + d(b1)
        G = d(g).permute(0, 2, 3, 1).reshape(-1, 2)
        gh = (G @ d(W2)) * (P > 0)
        ghr, hr = tf32_rna(gh.float()), tf32_rna(torch.relu(P).float())
        gx = torch.empty_like(x, dtype=torch.float64)
        gx.copy_((ghr @ rnd(W1)).reshape(B, H, W, 16).permute(0, 3, 1, 2))
        want = (gx, ghr.t() @ X, gh.sum(0), tf32_rna(G.float()).t() @ hr, G.sum(0))
        tol = 5e-5          # gh / h are rounded from float32 values that differ in the last bit between kernel and reference
    else:
        want = cpu_kernels.fh_bwd(d(x), d(W1), d(b1), d(W2), d(g))
        tol = 2e-5
    for a, b, name in zip(got, want, ('gx', 'gW1', 'gb1', 'gW2', 'gb2')):
        assert rel_l2(a.cpu().numpy(), b.cpu().numpy()) < tol, name
    a, M = torch.randn(16, generator=gen).cuda(), torch.randn(16, 16, generator=gen).cuda()
    gx = got[0].clone()
    F._fh_affine(x, a, M, gx)
    want_gx = cpu_kernels.fh_affine(d(x), d(a), d(M), d(got[0]).clone())
    assert rel_l2(gx.cpu().numpy(), want_gx.cpu().numpy()) < 1e-5
    again = F._fh_bwd(x, W1, b1, W2, g)
    assert all(torch.equal(p, q) for p, q in zip(got, again))           # fixed-order partial sums: bit reproducible


@pytest.mark.gpu
def test_stage_and_backbone_vs_aten(k6_on_this_device, monkeypatch):
    import bihome_b200.functional as F
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    compare_stage(F, make_stage(torch.float32, 'cuda'), torch.float32, 'cuda', (4, 16, 64, 64), 2e-5)
    from bihome_b200.backbones import Rethinking
    kw = dict(IMAGE_SIZE=128, PATCH_KEYS=['patch_1', 'patch_2'], TARGET_KEYS=['pf_hat_12', 'pf_hat_21'], RESNET_BLOCK='ResNet34',
              PRETRAINED_RESNET=False, VARIANT='DoubleLine')
    torch.manual_seed(1)
    net = Rethinking.Model(**kw).cuda().to(memory_format=torch.channels_last).train()
    twin = copy.deepcopy(net)
    p1, p2 = torch.rand(4, 1, 128, 128).cuda(), torch.rand(4, 1, 128, 128).cuda()
    g = torch.randn(4, 2, 128, 128).cuda()
    res = []
    for model, mode in ((net, 'fused'), (twin, 'aten')):
        monkeypatch.setenv('BH_FIELD_HEAD', mode)
        out = model({'patch_1': p1, 'patch_2': p2})
        loss = (out['pf_hat_12'] * g).sum() + (out['pf_hat_21'] * g.flip(0)).sum()
        grads = torch.autograd.grad(loss, [p for p in model.parameters()], allow_unused=True)
        res.append((out['pf_hat_12'].detach(), grads))
    assert rel_l2(res[0][0].cpu().numpy(), res[1][0].cpu().numpy()) < 1e-4
    num = sum(float(((a - b).double() ** 2).sum()) for a, b in zip(*[r[1] for r in res]) if a is not None)
    den = sum(float((b.double() ** 2).sum()) for b in res[1][1] if b is not None)
    # layer8's ReLU units within 1e-6 of their threshold (a handful among 8 M) may switch between the two evaluations;
    # each moves the gradient of one pixel by ~10 %: a few 1e-4 of the whole -- a wiring error would show as O(1)
    assert (num / den) ** 0.5 < 1e-2


@pytest.mark.gpu
def test_training_entry_point_with_the_fused_head(k6_on_this_device, monkeypatch, tmp_path):
    """train.py / eval.py on the shipped Zeng config with layer8 on K6: steps run, weights stay finite, the checkpoint
    carries the BatchNorm buffers K6 maintains, evaluation (K6's eval-mode fold) gives a finite MACE"""
    import numpy as np
    from conftest import load_entry
    monkeypatch.setenv('BH_FIELD_HEAD', 'fused')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    config = os.path.join(root, 'config', 'pds-coco', 'zeng-bihome-lr-1e-3.yaml')
    train, ev = load_entry('train'), load_entry('eval')
    log_dir = os.path.join(str(tmp_path), 'log')
    train.main(config, batch_size=8, max_steps=3, synthetic_pool=8, log_dir=log_dir)
    blob = torch.load(os.path.join(log_dir, 'model_000003.pth'), map_location='cpu', weights_only=False)
    assert blob['step'] == 3
    assert all(torch.isfinite(v).all() for v in blob['model'].values() if torch.is_floating_point(v))
    assert int(blob['model']['0.layer8.1.num_batches_tracked']) == 6          # two backbone passes per step
    assert float(blob['model']['0.layer8.1.running_var'].min()) > 0
    assert np.isfinite(ev.main(config, os.path.join(log_dir, 'model_000003.pth'), batch_size=8, samples=16))
