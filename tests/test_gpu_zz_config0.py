"""BASELINE.json configs[0] -- the reference's own CPU-runnable case: s-coco DeTone-style ResNet-34 regressor + biHomE
loss (ResNet-34 stem extractor), 128x128 patches, batch 8, forward + backward.  The whole model on the GPU path against
the CPU oracle model (oracle/ref_train.py, float64) with the same weights on the same batch: loss, predicted offsets and
the gradient of every learnable parameter."""
import os

import pytest
import torch

from conftest import rel_l2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIG = os.path.join(ROOT, 'config', 's-coco', 'detone-bihome-lr-5e-3.yaml')


def build_pair(dtype):
    """(ours, oracle) with identical weights; `ours` is the product model, `oracle` the CPU restatement in `dtype`"""
    from bihome_b200 import engine
    from bihome_b200.backbones import ResNet34
    from oracle.ref_train import OracleModel
    cfg = engine.load_config(CONFIG)
    torch.manual_seed(0)
    ours = engine.build_model(cfg, pretrained=False)
    bcfg = dict(cfg['MODEL']['BACKBONE'])
    bcfg['PRETRAINED_RESNET'] = False
    oracle = OracleModel(ResNet34.Model(**bcfg), cfg['MODEL']['HEAD'])
    oracle.backbone.load_state_dict(ours[0].state_dict())
    oracle.extractor.resnet.load_state_dict(ours[1].auxiliary_resnet.resnet.state_dict(), strict=False)
    return ours, oracle.to(dtype)


def batch(B=8, P=128):
    g = torch.Generator().manual_seed(4)
    lo = torch.rand(B, 1, P // 8 + 1, P // 8 + 1, generator=g)
    p1 = torch.nn.functional.interpolate(lo, size=(P, P), mode='bicubic', align_corners=True)
    p2 = torch.roll(p1, shifts=(3, -2), dims=(2, 3)) + 0.02 * torch.randn(B, 1, P, P, generator=g)
    return p1, p2


def step(model, p1, p2):
    model.train()
    loss, _, delta_hat = model({'patch_1': p1, 'patch_2': p2})
    params = [p for p in model.parameters() if p.requires_grad]
    return loss.detach(), delta_hat.detach(), torch.autograd.grad(loss, params)


def compare(ours_out, ref_out, loss_abs, grad_tol):
    (lo, dh_o, go), (lr, dh_r, gr) = ours_out, ref_out
    assert torch.isfinite(lr) and abs(float(lo) - float(lr)) <= 5e-4 * abs(float(lr)) + loss_abs, (float(lo), float(lr))
    assert rel_l2(dh_o.cpu().numpy(), dh_r.numpy()) < 1e-3
    assert len(go) == len(gr)
    num = sum(float(((a.cpu().double() - b.double()) ** 2).sum()) for a, b in zip(go, gr))
    den = sum(float((b.double() ** 2).sum()) for b in gr)
    assert den > 0 and (num / den) ** 0.5 < grad_tol, (num / den) ** 0.5


@pytest.mark.gpu
def test_config0_training_step_matches_cpu_oracle():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ours, oracle = build_pair(torch.float64)
    p1, p2 = batch()
    ref_out = step(oracle, p1.double(), p2.double())
    # a float32 evaluation of this 34-layer model differs from the float64 one by 7e-3 in the parameter gradient
    # (ReLU / bilinear-cell flips; measured on CPU with the oracle ops) -- a wiring error would show as O(1)
    ours = ours.cuda()
    compare(step(ours, p1.cuda(), p2.cuda()), ref_out, loss_abs=1e-3, grad_tol=5e-2)
    ours_cl = ours.to(memory_format=torch.channels_last)
    compare(step(ours_cl, p1.cuda(), p2.cuda()), ref_out, loss_abs=1e-3, grad_tol=5e-2)


def test_config0_host_logic_matches_cpu_oracle(monkeypatch):
    """the same comparison on CPU with the CUDA ops replaced by the oracle's closed forms (tests/cpu_kernels.py):
    pins the weight mapping and the wiring of the product model for this configuration"""
    import cpu_kernels
    cpu_kernels.install(monkeypatch)
    ours, oracle = build_pair(torch.float32)
    p1, p2 = batch(B=4)
    compare(step(ours, p1, p2), step(oracle, p1, p2), loss_abs=1e-4, grad_tol=5e-3)
