"""Warp + biHomE-loss kernel microbenchmark (BASELINE.json configs[3]): batch 64-4096, patch 128-512, C = 64-256.

    python tools/microbench.py                 # north-star shape + a small sweep, JSON lines
    python tools/microbench.py --sweep         # the full sweep (shapes over ~60 GB footprint are skipped)
    python tools/microbench.py --once          # one launch of every kernel at B=256, P=128 (the command ncu profiles)

Every timing: CUDA events on the launching stream, >= 3 warm-up launches, the L2 flushed (256 MB written, then 256 MB read) before each
timed launch, mean of `--iters` launches.  Algorithmic bytes are those of SURVEY.md 8(d) / DESIGN.md section 4.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes

import torch
import bihome_b200.functional as F
from bihome_b200 import cabi

LIB = cabi.lib()


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

PEAK = 6460.5
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:  # noqa: BLE001
    pass


class Timer:
    """L2 flush before every timed launch: a 256 MB buffer is WRITTEN (evicts everything the kernel under test could hit),
    then a second 256 MB buffer is READ so that the L2 is left full of clean lines.  Without the read the kernel inherits
    ~126 MB of dirty lines from the memset and pays their write-back (~20 us at HBM speed) inside its own timed region --
    measured as a constant 18-30 us offset on every kernel of this file, whatever its size."""
    def __init__(self, iters, flush=True, dirty=False):
        self.iters = iters
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda') if flush else None
        self.clean = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device='cuda') if flush and not dirty else None
        self.sink = None

    def __call__(self, fn, warm=3):
        for _ in range(warm):
            fn()
        tot = 0.0
        for _ in range(self.iters):
            if self.flush is not None:
                self.flush.zero_()
                if self.clean is not None:
                    self.sink = self.clean.sum()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / self.iters


def report(name, shape, ms, nbytes):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({'kernel': name, **shape, 'ms': round(ms, 5), 'algorithmic_MB': round(nbytes / 1e6, 2), 'GB/s': round(gbs, 1),
                      'frac_of_measured_peak': round(gbs / PEAK, 3)}), flush=True)


def rand_h(B, P, scale=0.25):
    d = (torch.rand(B, 4, 2, device='cuda') * 2 - 1) * (P * scale)
    return F.dlt4(d, size=(P, P)), d


def bench_image_warp(B, P, t, scale=0.25, tag=''):
    """north-star path: 1-channel patches, both directions batched (2B planes), pooled 4x4 mask fused.
    Timed as raw C-ABI calls on preallocated buffers (the autograd wrapper costs ~50 us of CPU per call).
    scale: corner offsets are U(-scale P, scale P) -- 0.25 is the data set's worst case (rho = 32 at P = 128), 0.02 what an
    untrained network predicts (bench.py's in-situ launches)."""
    img = torch.rand(2 * B, 1, P, P, device='cuda')
    H, _ = rand_h(2 * B, P, scale)
    H = H.detach().contiguous()
    out, mask = torch.empty_like(img), torch.empty(2 * B, P // 4, P // 4, device='cuda')
    gO, gM = torch.randn_like(out), torch.randn_like(mask)
    gH = torch.empty(2 * B, 9, device='cuda')
    nws = LIB.bh_warp_bwd_workspace_bytes(2 * B, 1, P, P, P, P, 0)
    ws = torch.empty(max(nws, 16), dtype=torch.uint8, device='cuda')
    st = stream()

    def fwd():
        cabi.check(LIB.bh_warp_fwd(ptr(img), ptr(H), ptr(out), ptr(mask), 2 * B, 1, P, P, P, P, 4, 0, st), 'fwd')

    def bwd():
        cabi.check(LIB.bh_warp_bwd(ptr(img), ptr(H), ptr(gO), ptr(gM), ptr(gH), None, 2 * B, 1, P, P, P, P, 4, 0, ptr(ws), nws, st), 'bwd')
    pair = 2 * (4 * P * P + 4 * P * P + 4 * (P // 4) ** 2)            # read src + write out + write pooled mask, 2 directions
    report('warp_fwd(image+mask)' + tag, {'B': B, 'P': P, 'offsets': scale}, t(fwd), pair * B)
    pairb = 2 * (4 * P * P + 4 * P * P + 4 * (P // 4) ** 2 + 36)
    report('warp_bwd(dH)' + tag, {'B': B, 'P': P, 'offsets': scale}, t(bwd), pairb * B)


def bench_copy(B, P, t):
    """calibration: a device-to-device copy of the same 2B planes (read + write = the warp's traffic minus the masks) --
    what fraction of the 2 GiB-copy peak of MEASURED_PEAKS.json a transfer of this size can reach at all"""
    a = torch.rand(2 * B, 1, P, P, device='cuda')
    b = torch.empty_like(a)
    report('copy (same planes, torch copy_)', {'B': B, 'P': P}, t(lambda: b.copy_(a)), 2 * a.numel() * 4)


def bench_feature_warp(B, P, C, t):
    x = torch.rand(B, C, P, P, device='cuda').contiguous(memory_format=torch.channels_last)
    H, _ = rand_h(B, P)
    H = H.detach().requires_grad_(True)
    out = F.warp(x, H, P, P)
    gO = torch.randn_like(out)
    report('warp_fwd(nhwc features)', {'B': B, 'P': P, 'C': C}, t(lambda: F.warp(x, H, P, P)), 2 * 4 * C * P * P * B)
    report('warp_bwd(nhwc, dH)', {'B': B, 'P': P, 'C': C}, t(lambda: torch.autograd.grad(out, H, gO, retain_graph=True)),
           2 * 4 * C * P * P * B)


def bench_loss(B, P, C, t, nhwc):
    h = P // 4
    mk = lambda: (torch.rand(B, C, h, h, device='cuda').contiguous(memory_format=torch.channels_last) if nhwc
                  else torch.rand(B, C, h, h, device='cuda'))
    f1, f2, f1w, f2w = mk(), mk(), mk(), mk()
    m1w, m2w = torch.rand(B, h, h, device='cuda'), torch.rand(B, h, h, device='cuda')
    H12, _ = rand_h(B, P)
    H21, _ = rand_h(B, P)
    nbytes = ((4 + 2) * C * h * h * 4 + 4 * h * h * 4) * B
    report('bihome_loss fwd+bwd (%s)' % ('nhwc' if nhwc else 'nchw'), {'B': B, 'P': P, 'C': C, 'h': h},
           t(lambda: F.bihome_loss(f1, f2, f1w, f2w, m1w, m2w, H12, H21, 0.01)), nbytes)


def _torch_triplet(f1, f2, f1w, f2w, a1, a2, h12, h21, hinge, margin, mu):
    """the ATen op chain K3g replaces (the reference's algebra, PerceptualHead.py:555-665, l1 distances) incl. autograd"""
    l1, l2, l3 = (f1w - f2).abs(), (f2w - f1).abs(), (f1 - f2).abs()

    def line(la, lb):
        if hinge is None:
            return la.sum(1) - lb.sum(1)
        if hinge == 'channel':
            return torch.clamp(la - lb + margin, min=0).sum(1)
        return torch.clamp(la.sum(1) - lb.sum(1) + margin, min=0)
    d1, d2 = a1.sum(-1).sum(-1), a2.sum(-1).sum(-1)
    ln1 = ((a1 * line(l1, l3)).sum(-1).sum(-1) / torch.max(d1, torch.ones_like(d1))).sum()
    ln2 = ((a2 * line(l2, l3)).sum(-1).sum(-1) / torch.max(d2, torch.ones_like(d2))).sum()
    eye = torch.eye(3, device=h12.device).unsqueeze(0)
    return ln1 + ln2 + mu * ((torch.matmul(h12, h21) - eye) ** 2).sum()


def bench_triplet(B, C, h, nhwc, t):
    """K3g (bh_triplet_fwd_bwd: three launches) against the torch element-wise ops + autograd it replaces, forward +
    backward to the warped features, masks and homographies (the extractor is frozen: no gradient for f1 / f2)"""
    mk = lambda: (torch.rand(B, C, h, h, device='cuda').contiguous(memory_format=torch.channels_last) if nhwc and C > 1
                  else torch.rand(B, C, h, h, device='cuda'))
    f1, f2 = mk(), mk()
    f1w, f2w = mk().requires_grad_(True), mk().requires_grad_(True)
    a1, a2 = torch.rand(B, h, h, device='cuda', requires_grad=True), torch.rand(B, h, h, device='cuda', requires_grad=True)
    H12 = (torch.eye(3, device='cuda').repeat(B, 1, 1) + 0.01 * torch.randn(B, 3, 3, device='cuda')).requires_grad_(True)
    H21 = torch.linalg.inv(H12.detach()).requires_grad_(True)
    leaves = [f1w, f2w, a1, a2, H12, H21]
    shape = {'B': B, 'C': C, 'h': h, 'layout': 'nhwc' if nhwc and C > 1 else 'nchw'}
    for hinge, margin in ((None, 0.0), ('channel', 0.05), ('pixel', 1.0)):
        nbytes = ((4 + 2) * C * h * h * 4 + 4 * h * h * 4) * B
        if hinge == 'pixel' and not (nhwc and C > 1):
            nbytes += 4 * C * h * h * 4 * B      # the planar kernel walks the channels twice (second walk out of L2)

        def fused():
            loss_b, _ = F.triplet_loss(f1, f2, f1w, f2w, a1, None, a2, None, H12, H21, lines=2, distance='l1', hinge=hinge,
                                       margin=margin, mu=0.01)
            torch.autograd.grad(loss_b.sum(), leaves)

        def aten():
            torch.autograd.grad(_torch_triplet(f1, f2, f1w, f2w, a1, a2, H12, H21, hinge, margin, 0.01), leaves)
        name = 'hinge=%s' % hinge
        report('triplet fwd+bwd K3g (%s)' % name, shape, t(fused), nbytes)
        report('triplet fwd+bwd ATen ops (%s)' % name, shape, t(aten), nbytes)


def bench_small(B, P, t):
    H, d = rand_h(B, P)
    d = d.requires_grad_(True)
    gH = torch.randn(B, 3, 3, device='cuda')
    report('dlt4_fwd', {'B': B}, t(lambda: F.dlt4(d, size=(P, P))), 68 * B)
    Hd = F.dlt4(d, size=(P, P))
    report('dlt4_bwd', {'B': B}, t(lambda: torch.autograd.grad(Hd, d, gH, retain_graph=True)), 68 * B)
    field = torch.randn(B, 2, P, P, device='cuda', requires_grad=True)
    choice = torch.randint(1, P * P, (B, 128), device='cuda')
    four = torch.tensor([[0., 0.], [P, 0.], [P, P], [0., P]], device='cuda')
    report('dltn_fwd(field, 128 pts)', {'B': B}, t(lambda: F.dltn_field(field, choice, four)), (128 * 8 + 128 * 8 + 68) * B)
    _, dh = F.dltn_field(field, choice, four)
    gd = torch.randn_like(dh)
    report('dltn_bwd', {'B': B}, t(lambda: torch.autograd.grad(dh, field, gd, retain_graph=True)), (128 * 16 + 128 * 8) * B)
    pool = torch.randint(0, 256, (64, 240, 320, 3), dtype=torch.uint8, device='cuda')
    params, index = F.pairgen_draw(B, 64, (240, 320), 32, P if P <= 128 else 128, 32.0, 1, 0, 'cuda')
    report('pairgen_apply', {'B': B}, t(lambda: F.pairgen_apply(pool, index, params, 128)),
           (2 * 3 * (128 + 64) ** 2 + 2 * 4 * 128 * 128) * B)


def bench_field_head(B, P, t):
    """K6 (csrc/fieldhead.cu) entry points and the whole stage, fused against the four ATen modules, at [B,16,P,P]"""
    from bihome_b200 import autotune
    n = B * P * P
    x = torch.relu(torch.randn(B, 16, P, P, device='cuda') + 0.3).contiguous(memory_format=torch.channels_last)
    W1, b1 = torch.randn(128, 16, device='cuda') * 0.3, torch.randn(128, device='cuda')
    W2, b2 = torch.randn(2, 128, device='cuda') * 0.2, torch.randn(2, device='cuda')
    g = torch.randn(B, 2, P, P, device='cuda')
    gx = torch.empty_like(x)
    a, M = torch.randn(16, device='cuda'), torch.randn(16, 16, device='cuda')
    shape = {'B': B, 'P': P}
    report('fieldhead_moments', shape, t(lambda: F._fh_moments(x)), 64 * n)
    report('fieldhead_fwd', shape, t(lambda: F._fh_fwd(x, W1, b1, W2, b2)), (64 + 8) * n)
    report('fieldhead_bwd', shape, t(lambda: F._fh_bwd(x, W1, b1, W2, g)), (64 + 8 + 64) * n)
    report('fieldhead_affine', shape, t(lambda: F._fh_affine(x, a, M, gx)), (64 + 64 + 64) * n)
    fused, aten = autotune._stage(torch.device('cuda')), autotune._stage(torch.device('cuda'))

    def step(fn):
        xa = x.detach().requires_grad_(True)
        (fn(xa) * g).sum().backward()
    total = (64 + 64 + 8 + 64 + 8 + 64 + 192) * n      # moments + fwd + bwd + affine
    report('field head fwd+bwd, K6', shape, t(lambda: step(lambda v: F.field_head(fused, v))), total)
    report('field head fwd+bwd, ATen modules', shape, t(lambda: step(aten)), total)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sweep', action='store_true')
    ap.add_argument('--once', action='store_true')
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--warp-only', action='store_true', help='time only the 1-channel image warp (forward + backward)')
    ap.add_argument('--dirty', action='store_true', help='flush by memset only (leaves the L2 full of dirty lines: adds their write-back to every timing)')
    ap.add_argument('--loss-cl', action='store_true', help='time the fused loss for every cluster size (BH_LOSS_CL knob)')
    ap.add_argument('--feature-warp', action='store_true', help='time the channels-last feature-map warp (forward, backward to dH) over the sweep')
    ap.add_argument('--triplet', action='store_true', help='time K3g (the other loss variants) against the ATen op chain it replaces')
    ap.add_argument('--field-head', action='store_true', help='time K6 (the Zeng field head) against the ATen modules')
    a = ap.parse_args()
    torch.manual_seed(0)
    if a.once:
        t = lambda fn, warm=1: (fn(), fn(), torch.cuda.synchronize(), 1.0)[-1]
        bench_image_warp(256, 128, t)
        bench_loss(256, 128, 64, t, True)
        bench_loss(256, 128, 64, t, False)
        bench_loss(1024, 128, 64, t, True)      # B >= 512: the clustered TMA-ring kernel
        bench_feature_warp(64, 128, 64, t)
        bench_small(256, 128, t)
        F.mace(torch.randn(256, 4, 2, device='cuda'), torch.randn(256, 4, 2, device='cuda'))
        F.coverage_mask(rand_h(256, 128)[0].detach(), (128, 128), (128, 128), pool=8)      # the stand-alone analytic mask kernel
        bench_field_head(256, 128, t)
        bench_triplet(256, 64, 32, True, t)
        bench_triplet(64, 1, 128, False, t)
        torch.cuda.synchronize()
        return
    t = Timer(a.iters, dirty=a.dirty)
    if a.feature_warp:
        for B, P, C in ((64, 128, 64), (256, 128, 64), (64, 128, 256), (16, 256, 64), (256, 64, 128)):
            bench_feature_warp(B, P, C, t)
        return
    if a.triplet:
        for B, C, h, nhwc in ((256, 64, 32, True), (256, 64, 32, False), (64, 1, 128, False), (1024, 64, 32, True)):
            bench_triplet(B, C, h, nhwc, t)
        return
    if a.field_head:
        for B in (64, 256):
            bench_field_head(B, 128, t)
        return
    if a.loss_cl:
        for cl in (1, 2, 4, 8):
            F.tune('loss_cluster', cl)
            print(json.dumps({'loss_cluster': cl}))
            for B in (256, 1024):
                bench_loss(B, 128, 64, t, True)
                bench_loss(B, 128, 64, t, False)
        F.tune('loss_cluster', 0)
        for k, var in enumerate(('ldg', 'cluster', 'stream')):
            F.tune('loss_variant', k + 1)
            print(json.dumps({'loss_variant': var}))
            for B in (256, 1024, 4096):
                bench_loss(B, 128, 64, t, True)
        F.tune('loss_variant', 0)
        return
    if a.warp_only:
        for B, P in ((256, 128), (1024, 128), (4096, 128), (256, 256), (64, 512)):
            bench_copy(B, P, t)
            for variant in ((0, 1) if P == 128 and B != 1024 else (0,)):
                F.tune('warp_variant', variant)
                tag = ' [tile%s]' % (', 4 warps x 8 rows' if variant & 1 else ', 2 warps x 16 rows')
                for scale in ((0.02, 0.25) if variant == 0 else (0.02,)):
                    bench_image_warp(B, P, t, scale, tag)
            F.tune('warp_variant', 0)
            F.tune('warp_path', 1)
            bench_image_warp(B, P, t, 0.02, ' [ring]')
            F.tune('warp_path', 0)
        return
    print(json.dumps({'peak_GBps_measured': PEAK, 'timing': 'cuda events, L2 flushed (256 MB written, then 256 MB read: clean lines) before every launch, mean of %d' % a.iters}))
    shapes = [(256, 128, 64)]
    if a.sweep:
        shapes = [(B, P, C) for B in (64, 256, 1024, 4096) for P in (128, 256, 512) for C in (64, 128, 256)]
    else:
        shapes += [(1024, 128, 64), (4096, 128, 64), (256, 256, 64), (64, 512, 64), (256, 128, 256)]
    done_img, done_small = set(), set()
    for B, P, C in shapes:
        h = P // 4
        if (B, P) not in done_img and 2 * B * P * P * 4 * 4 < 60e9:
            done_img.add((B, P))
            bench_image_warp(B, P, t)
        if 6 * B * C * h * h * 4 < 60e9:
            bench_loss(B, P, C, t, True)
            bench_loss(B, P, C, t, False)
        if 4 * B * C * P * P * 4 < 60e9:
            bench_feature_warp(B, P, C, t)
        if B not in done_small and P == 128:
            done_small.add(B)
            bench_small(B, P, t)
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
