#!/usr/bin/env python
"""GPU-vs-GPU baseline of the whole training step (BASELINE.md 5.1 "again with the frozen ResNet-34 extractor included",
reference train.py:299-387): the reference's own torch-CUDA op sequence against this repo's step, same device, same
configuration (pds-coco/zeng-bihome-lr-1e-3, B = 256, 128x128, fp32 with torch's default TF32 convolutions).

    python tools/ref_step_cuda.py --steps 20 --warmup 5 --out gpurun_out/ref_step_cuda.json

reference arm = the backbone's plain torch modules (NCHW, ATen / cuDNN, the reference's op order incl. layer8's bias) +
the oracle restatement of the reference's head on the GPU (oracle/ref_train.py -> oracle/ref_path.py over
oracle/kornia050.py: torch.multinomial draw, gather, normalised DLT by batched SVD, get_perspective_transform, torch.inverse,
materialised sampling grids, F.grid_sample x4, AvgPool2d x4, the ~110-op loss and its autograd) + Adam over all parameters.
The reference cannot be imported on the GPU box (no /root/reference there, kornia 0.5.0 does not run on torch 2.11); the
restatement reproduces its modules bit for bit on CPU (tests/test_oracle_golden.py).
our arm = engine.train_step on the same synthetic patches (pair generation excluded on both sides).
Test infrastructure: imports oracle/.  Timing: CUDA events around `steps` steps after `warmup`, one synchronize per side.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CONFIG = os.path.join(ROOT, 'config', 'pds-coco', 'zeng-bihome-lr-1e-3.yaml')


def timed(step, steps, warmup):
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, float(loss)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--batch', type=int, default=256)
    ap.add_argument('--out', default=None)
    a = ap.parse_args()
    from bihome_b200 import engine
    from bihome_b200.backbones import Rethinking
    from oracle.ref_train import OracleModel
    dev = torch.device('cuda')
    cfg = engine.load_config(CONFIG)
    B = a.batch
    g = torch.Generator().manual_seed(1)
    lo = torch.rand(B, 1, 18, 18, generator=g)
    p1 = torch.nn.functional.interpolate(lo, size=(128, 128), mode='bicubic', align_corners=True)
    p2 = torch.roll(p1, shifts=(3, -2), dims=(2, 3)) + 0.05 * torch.randn(B, 1, 128, 128, generator=g)
    p1, p2 = p1.to(dev), p2.to(dev)

    # ---- reference arm -------------------------------------------------------------------------------------------------
    os.environ['BH_FIELD_HEAD'] = 'aten'
    bcfg = dict(cfg['MODEL']['BACKBONE'])
    bcfg['PRETRAINED_RESNET'] = False
    torch.manual_seed(0)
    ref = OracleModel(Rethinking.Model(**bcfg), cfg['MODEL']['HEAD']).to(dev)
    ref.backbone.skip_cancelled_bias = False
    ref.train()
    ref_opt = torch.optim.Adam(ref.parameters(), lr=cfg['SOLVER']['LR'])

    def ref_step():
        ref_opt.zero_grad()
        loss, _, _ = ref({'patch_1': p1, 'patch_2': p2})
        loss.backward()
        ref_opt.step()
        return loss.detach()
    torch.backends.cudnn.benchmark = False          # the reference sets no cuDNN flag
    torch.cuda.reset_peak_memory_stats()
    ref_ms, ref_loss = timed(ref_step, a.steps, a.warmup)
    ref_mem = torch.cuda.max_memory_allocated() / 1e9
    torch.backends.cudnn.benchmark = True           # ... and with the flag this repo's train.py sets, for a like-for-like row
    ref_ms_tuned, _ = timed(ref_step, a.steps, a.warmup + 3)
    del ref, ref_opt
    torch.cuda.empty_cache()

    # ---- our arm -------------------------------------------------------------------------------------------------------
    out = {'config': 'pds-coco/zeng-bihome-lr-1e-3', 'B': B, 'steps': a.steps, 'warmup': a.warmup,
           'timing': 'cuda events around the timed steps, fp32, TF32 convolutions (torch default) on both sides',
           'reference_torch_cuda': {'ms_per_step': ref_ms, 'pairs_per_s': B / (ref_ms * 1e-3), 'peak_memory_GB': ref_mem,
                                    'final_loss': ref_loss, 'ms_per_step_with_cudnn_benchmark': ref_ms_tuned}}
    for side in ('aten', 'fused'):
        os.environ['BH_FIELD_HEAD'] = side
        torch.manual_seed(0)
        model = engine.build_model(cfg, pretrained=False).to(dev).to(memory_format=torch.channels_last)
        model.train()
        opt, sched = engine.build_optimizer(cfg, model)
        batch = {'patch_1': p1, 'patch_2': p2}

        def our_step():
            loss, _, _ = engine.train_step(model, dict(batch), opt, sched)
            return loss.detach()
        torch.cuda.reset_peak_memory_stats()
        ms, loss = timed(our_step, a.steps, a.warmup + 3)          # cudnn.benchmark (train.py's default) is on from here
        out['bihome_b200_field_head_' + side] = {'ms_per_step': ms, 'pairs_per_s': B / (ms * 1e-3),
                                                 'peak_memory_GB': torch.cuda.max_memory_allocated() / 1e9, 'final_loss': loss,
                                                 'speedup_vs_reference_torch_cuda': ref_ms / ms,
                                                 'speedup_vs_reference_with_cudnn_benchmark': ref_ms_tuned / ms}
        del model, opt, sched
        torch.cuda.empty_cache()
    text = json.dumps(out, indent=1)
    print(text)
    if a.out:
        with open(a.out, 'w') as f:
            f.write(text + '\n')


if __name__ == '__main__':
    main()
