"""K7 / K7b per shape: CUDA-event time of bh_stem_* / bh_bnact_* at the tensor shapes of the B = 256 zeng-bihome step, against
the ATen modules they replace (forward and backward timed separately, L2 flushed before every launch unless --warm).
One JSON line per case.  Under `ncu --metrics gpu__time_duration.sum` the same script gives the per-kernel split (--once)."""
import argparse
import copy
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bihome_b200.functional as F  # noqa: E402

SHAPES = [(256, 64, 32, 32), (256, 128, 16, 16), (256, 256, 8, 8), (256, 256, 16, 16), (256, 128, 32, 32), (256, 64, 64, 64),
          (256, 32, 64, 64), (256, 32, 128, 128), (256, 16, 128, 128)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--once', action='store_true', help='one launch per case (for ncu)')
    ap.add_argument('--warm', action='store_true', help='do not flush the L2 between launches')
    ap.add_argument('--iters', type=int, default=10)
    ap.add_argument('--only-stem', action='store_true', help='skip the K7b shapes (short ncu captures of the K7 kernels)')
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    iters = 1 if a.once else a.iters

    def timed(fn, prep=None):
        ms = []
        for i in range(iters + (0 if a.once else 2)):
            if prep:
                prep()
            if not a.warm:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        return sum(ms[-iters:]) / iters

    for shape in ([] if a.only_stem else SHAPES):
        N, C, H, W = shape
        T = N * C * H * W * 4
        bn = torch.nn.BatchNorm2d(C).to(dev).train()
        ref = copy.deepcopy(bn)
        x = torch.randn(*shape, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        r = torch.randn(*shape, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        g = torch.randn(*shape, device=dev).contiguous(memory_format=torch.channels_last)
        for res in (False, True):
            state = {}

            def f_fwd():
                state['y'] = F.bn_relu(bn, x, residual=r if res else None)

            def f_bwd():
                torch.autograd.grad(state['y'], [x] + ([r] if res else []) + [bn.weight, bn.bias], g)

            def a_bwd():
                torch.autograd.grad(state['y'], [x] + ([r] if res else []) + [ref.weight, ref.bias], g)

            def a_fwd():
                state['y'] = torch.relu(ref(x) + r) if res else torch.relu(ref(x))
            row = {'case': 'bn_relu', 'shape': shape, 'residual': res, 'MB': T / 1e6}
            row['fused_fwd_us'] = timed(f_fwd) * 1e3
            row['fused_bwd_us'] = timed(f_bwd, prep=f_fwd) * 1e3
            row['aten_fwd_us'] = timed(a_fwd) * 1e3
            row['aten_bwd_us'] = timed(a_bwd, prep=a_fwd) * 1e3
            row['fwd_frac_hbm'] = T * (3 if res else 2) / (row['fused_fwd_us'] * 1e-6) / 6556.8e9
            row['bwd_frac_hbm'] = T * (5 if res else 3) / (row['fused_bwd_us'] * 1e-6) / 6556.8e9
            print(json.dumps(row), flush=True)
    # the stem
    shape = (256, 64, 64, 64)
    bn = torch.nn.BatchNorm2d(64).to(dev).train()
    ref = copy.deepcopy(bn)
    pool = torch.nn.MaxPool2d(3, 2, 1)
    x = torch.randn(*shape, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    g = torch.randn(256, 64, 32, 32, device=dev).contiguous(memory_format=torch.channels_last)
    state = {}

    def s_fwd():
        state['y'] = F.stem(bn, x)

    def s_bwd():
        torch.autograd.grad(state['y'], [x, bn.weight, bn.bias], g)

    def a_fwd():
        state['y'] = pool(torch.relu(ref(x)))

    def a_bwd():
        torch.autograd.grad(state['y'], [x, ref.weight, ref.bias], g)
    row = {'case': 'stem', 'shape': shape}
    row['fused_fwd_us'] = timed(s_fwd) * 1e3
    row['fused_bwd_us'] = timed(s_bwd, prep=s_fwd) * 1e3
    row['aten_fwd_us'] = timed(a_fwd) * 1e3
    row['aten_bwd_us'] = timed(a_bwd, prep=a_fwd) * 1e3
    print(json.dumps(row), flush=True)


if __name__ == '__main__':
    main()
