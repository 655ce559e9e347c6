"""north_star target: >= 10x the reference's own torch-CUDA warp + biHomE-loss step on one B200 at B=256, 128x128.

SURVEY.md 8(d): "time 4 x _warp + pooling + loss fwd+bwd with features supplied (extractor excluded) -- CUDA events,
20 warm-up + 100 timed iterations, median".  The reference arm is the oracle restatement of the reference's torch code
(oracle/ref_path.py: get_perspective_transform -> torch.inverse -> warp_perspective/grid_sample x4 -> AvgPool2d x4 ->
double-line loss) executed on the GPU; our arm is K1 -> K2(+pooled masks) -> K3 -> K2b -> K1 adjoint through the
public Python API.  Both arms consume the same inputs, produce the same loss and the same gradients (checked here);
the extractor's backward is stood in for by a fixed upstream gradient on the warped patches.
The measured numbers are written to gpurun_out/speedup_warp_loss.json (copied to profiles/ by hand).
"""
import json
import os

import pytest
import torch

from conftest import ROOT, rel_l2

pytestmark = pytest.mark.gpu

B, P, C, MU = 256, 128, 64, 0.01


def _inputs():
    g = torch.Generator().manual_seed(7)
    lo = torch.rand(2 * B, 1, 18, 18, generator=g)
    patches = torch.nn.functional.interpolate(lo, size=(P, P), mode='bicubic', align_corners=True).cuda()
    h = P // 4
    feats = [torch.rand(B, C, h, h, generator=g).cuda() for _ in range(4)]
    d12 = ((torch.rand(B, 4, 2, generator=g) * 2 - 1) * 24).cuda()
    d21 = (-d12.cpu() + torch.rand(B, 4, 2, generator=g) * 3).cuda()
    up = [torch.randn(B, 1, P, P, generator=g).cuda() * 1e-3 for _ in range(2)]
    return patches[:B].contiguous(), patches[B:].contiguous(), feats, d12, d21, up


def _time(fn, warm=20, iters=100):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def test_warp_loss_step_is_10x_the_reference_torch_cuda_path():
    import bihome_b200.functional as F
    from oracle import ref_path as R
    p1, p2, (f1, f2, f1w0, f2w0), d12, d21, (u1, u2) = _inputs()
    ones = torch.ones_like(p1)

    def ref_step():
        a, b = d12.clone().requires_grad_(True), d21.clone().requires_grad_(True)
        f1w, f2w = f1w0.clone().requires_grad_(True), f2w0.clone().requires_grad_(True)
        p1w, _ = R.warp(p1, a)
        m1w, H12 = R.warp(ones, a)
        p2w, _ = R.warp(p2, b)
        m2w, H21 = R.warp(ones, b)
        loss, _ = R.bihome_double_line(f1, f2, f1w, f2w, ones, ones, m1w, m2w, H12, H21, MU)
        (loss + (p1w * u1).sum() + (p2w * u2).sum()).backward()
        return loss.detach(), a.grad, b.grad, f1w.grad, f2w.grad

    def our_step():
        a, b = d12.clone().requires_grad_(True), d21.clone().requires_grad_(True)
        f1w, f2w = f1w0.clone().requires_grad_(True), f2w0.clone().requires_grad_(True)
        H = F.dlt4(torch.cat([a, b]), size=(P, P))
        pw, mw = F.warp(torch.cat([p1, p2]), H, P, P, pool=4)
        loss_b, _ = F.bihome_loss(f1, f2, f1w, f2w, mw[:B], mw[B:], H[:B], H[B:], MU)
        loss = loss_b.sum()
        (loss + (pw * torch.cat([u1, u2])).sum()).backward()
        return loss.detach(), a.grad, b.grad, f1w.grad, f2w.grad

    r, o = ref_step(), our_step()
    # same step: loss and feature gradients to 1e-5, corner gradients through warp + DLT adjoint to 1e-3 against the
    # reference's own FLOAT32 evaluation (its grid noise flips bilinear cells; the float64 parity is in test_gpu_kernels)
    assert abs(float(o[0]) - float(r[0])) <= 1e-5 * abs(float(r[0]))
    assert rel_l2(o[3].cpu().numpy(), r[3].cpu().numpy()) < 1e-5
    assert rel_l2(o[4].cpu().numpy(), r[4].cpu().numpy()) < 1e-5
    assert rel_l2(o[1].cpu().numpy(), r[1].cpu().numpy()) < 5e-3
    assert rel_l2(o[2].cpu().numpy(), r[2].cpu().numpy()) < 5e-3

    t_ref, t_our = _time(ref_step), _time(our_step)
    # the same step with the Python/ATen glue removed: the five C-ABI launches captured once in a CUDA graph
    a, b = d12.clone().requires_grad_(True), d21.clone().requires_grad_(True)
    f1w, f2w = f1w0.clone().requires_grad_(True), f2w0.clone().requires_grad_(True)
    pcat, ucat = torch.cat([p1, p2]), torch.cat([u1, u2])

    def graph_body():
        a.grad = b.grad = f1w.grad = f2w.grad = None
        H = F.dlt4(torch.cat([a, b]), size=(P, P))
        pw, mw = F.warp(pcat, H, P, P, pool=4)
        loss_b, _ = F.bihome_loss(f1, f2, f1w, f2w, mw[:B], mw[B:], H[:B], H[B:], MU)
        torch.autograd.backward([loss_b.sum(), pw], [torch.ones((), device='cuda'), ucat])
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            graph_body()
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        graph_body()
    t_graph = _time(graph.replay)
    assert rel_l2(a.grad.cpu().numpy(), r[1].cpu().numpy()) < 5e-3

    out = {'B': B, 'P': P, 'C': C, 'h': P // 4, 'timing': 'cuda events, 20 warm-up + 100 timed, median',
           'reference_torch_cuda_ms': t_ref, 'ours_python_api_ms': t_our, 'ours_cuda_graph_ms': t_graph,
           'speedup_python_api': t_ref / t_our, 'speedup_cuda_graph': t_ref / t_graph,
           'pairs_per_s_reference': B / (t_ref * 1e-3), 'pairs_per_s_ours': B / (t_our * 1e-3),
           'pairs_per_s_ours_graph': B / (t_graph * 1e-3)}
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'speedup_warp_loss.json'), 'w') as f:
        json.dump(out, f, indent=1)
    assert t_ref / t_graph >= 10.0, out
