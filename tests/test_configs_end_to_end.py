"""Every shipped YAML of the reference, end to end on CPU: the reference's nn.Sequential(backbone, head) and this
package's, same weights (state dict copied across), same batch, same torch seed -> same loss and same gradients.

Build container only (needs the reference tree; skipped elsewhere).  The CUDA ops are replaced by the oracle's closed
forms (tests/cpu_kernels.py), so this pins the HOST logic of all seven backbone x head combinations -- dict protocol,
loss dispatch of train.py, RNG consumption of the multinomial draw, parameter names -- not the kernels.
"""
import glob
import os

import pytest
import torch

import cpu_kernels
from conftest import load_entry, rel_l2


def _configs():
    from oracle import ref_import
    if not ref_import.available():
        return []
    return sorted(glob.glob(os.path.join(ref_import.REFERENCE_ROOT, 'config', 'pds-coco', '*.yaml')) +
                  glob.glob(os.path.join(ref_import.REFERENCE_ROOT, 'config', 's-coco', '*.yaml')))


def _batch(B, P, gen):
    lo = torch.rand(B, 1, P // 8 + 1, P // 8 + 1, generator=gen)
    p1 = torch.nn.functional.interpolate(lo, size=(P, P), mode='bicubic', align_corners=True)
    p2 = torch.roll(p1, shifts=(2, -3), dims=(2, 3)) + 0.02 * torch.randn(B, 1, P, P, generator=gen)
    delta = torch.randint(-32, 32, (B, 4, 2), generator=gen).float()
    base = torch.tensor([[0, 0], [P, 0], [P, P], [0, P]]).float()
    corners = base.unsqueeze(0) + torch.tensor([[40., 30.], [100., 60.]])[:B].unsqueeze(1)
    return {'patch_1': p1, 'patch_2': p2, 'delta': delta, 'corners': corners}


@pytest.mark.parametrize('path', _configs(), ids=lambda p: os.path.basename(os.path.dirname(p)) + '/' + os.path.basename(p)[:-5])
def test_training_step_matches_reference(monkeypatch, path):
    from oracle import ref_import
    from bihome_b200 import engine
    from bihome_b200.data import gpu_pairs
    cpu_kernels.install(monkeypatch)
    train = load_entry('train')
    cfg = engine.load_config(path)
    bcfg, hcfg = dict(cfg['MODEL']['BACKBONE']), dict(cfg['MODEL']['HEAD'])
    bcfg['PRETRAINED_RESNET'] = False
    torch.manual_seed(3)
    rb = ref_import.load('src.backbones.' + bcfg['NAME']).Model(**bcfg)
    ref = torch.nn.Sequential(rb, ref_import.load('src.heads.' + hcfg['NAME']).Model(rb, **hcfg))
    ours = engine.build_model(cfg, pretrained=False)
    ours.load_state_dict(ref.state_dict())            # strict: identical names and shapes
    try:
        loss_fn = getattr(torch.nn, cfg['SOLVER']['LOSS'])()
    except AttributeError:
        loss_fn = cfg['SOLVER']['LOSS']
    B, P = 2, 128
    data = _batch(B, P, torch.Generator().manual_seed(11))
    t = gpu_pairs.transform_args(cfg['DATA']['TRANSFORMS'])
    data['target'] = data['delta'] if t['target_gen'] == '4_points' else \
        gpu_pairs.perspective_field_target(data['corners'], data['delta'], P)
    if 'image_1' in t['image_keys']:                    # s-coco/nguyen-orig: PhotometricHead warps the whole first image
        lo = torch.rand(B, 1, 31, 41, generator=torch.Generator().manual_seed(5))
        data['image_1'] = torch.nn.functional.interpolate(lo, size=(240, 320), mode='bicubic', align_corners=True)
    ref.train()
    ours.train()

    def run(model, ref_side):
        torch.manual_seed(99)                          # the Zeng heads draw their correspondences with torch.multinomial
        d = {k: v.clone() for k, v in data.items()}
        if ref_side:
            if isinstance(loss_fn, torch.nn.Module):
                gt, out, _, delta_hat = model(d)
                loss = loss_fn(gt, out)
            else:
                loss, _, delta_hat = model(d)
        else:
            loss, _, delta_hat = train.forward_loss(model, d, loss_fn)
        params = [p for p in model.parameters() if p.requires_grad]
        grads = torch.autograd.grad(loss, params, allow_unused=True)
        return loss.detach(), delta_hat.detach(), grads

    loss_r, dh_r, g_r = run(ref, True)
    loss_o, dh_o, g_o = run(ours, False)
    # float32 on both sides; the biHomE loss is a difference of O(10) masked feature distances that nearly cancel for
    # a random-init backbone (detone-bihome: -0.005), hence the absolute term
    assert torch.isfinite(loss_r) and abs(float(loss_o) - float(loss_r)) <= 2e-4 * abs(float(loss_r)) + 1e-4, (loss_o, loss_r)
    # the Zeng head's DLT on a random-init field amplifies the backbone's float32 round-off (1e-6) a hundredfold
    assert rel_l2(dh_o.numpy(), dh_r.numpy()) < 1e-3
    assert len(g_r) == len(g_o)
    num = sum(float(((a - b).double() ** 2).sum()) for a, b in zip(g_o, g_r) if a is not None and b is not None)
    den = sum(float((b.double() ** 2).sum()) for b in g_r if b is not None)
    assert all((a is None) == (b is None) for a, b in zip(g_o, g_r))
    assert den > 0 and (num / den) ** 0.5 < 5e-3, (num / den) ** 0.5
