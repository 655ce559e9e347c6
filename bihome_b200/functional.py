"""PyTorch autograd front-end of libbihome_b200.so.

Every function here takes CUDA float32 tensors, hands raw device pointers and the current CUDA stream to the
C ABI (include/bihome_b200.h) and returns torch tensors.  Nothing falls back to ATen: a CPU tensor raises.

  dlt4(delta, corners=None, size=(W, H))                   K1   four_point_to_homography   (src/data/utils.py:20-24)
  warp(src, H, out_h, out_w, pool=None)                    K2   warp_image(inverse=True)   (src/data/utils.py:54-59)
  coverage_mask(H, src_hw, out_hw, pool)                   K2   warp(ones) + AvgPool2d     (PerceptualHead.py:380-382,447-459)
  bihome_loss(f1, f2, f1w, f2w, m1w, m2w, H12, H21, mu)    K3   double-line biHomE loss    (PerceptualHead.py:559-561,609-665)
  triplet_loss(f1, f2, f1w, f2w, a1, b2, a2, b1, ...)      K3g  every other loss variant   (PerceptualHead.py:465-665, TripletHead.py:78-153)
  dltn(p1, p2, choice) / dltn_field(field, choice, four)   K4   find_homography_dlt        (ransac_utils.py:58-72, PerceptualHead.py:164-178)
  pairgen_draw(...) / pairgen_apply(...)                   K5   HomographyNetPrep pipeline (src/data/transforms.py:456-576)
  mace(delta_gt, delta_hat)                                     train.py:401-404
  field_head(stage, x)                                     K6   Zeng backbone layer8       (src/backbones/Rethinking.py:144-147)
  stem(bn, x)                                              K7   bn1 -> relu -> maxpool     (PerceptualHead.py:56-58, Rethinking.py:31-36)
  bn_relu(bn, x, residual=None)                            K7b  bn -> [+ skip] -> relu     (src/backbones/utils.py, torchvision BasicBlock)
  bn_bn_relu(bn_a, a, bn_b, b)                             K7c  relu(bn_a(a) + bn_b(b))    (src/backbones/utils.py: blocks with a projection skip)
  conv_transpose_bias(m, x)                                K8   ConvTranspose2d's bias     (src/backbones/utils.py:65-66, up-sampling blocks)
"""
import ctypes
import os

import torch

from . import cabi


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)


def _stream():
    """cudaStream_t of torch's current stream on the current device (the raw query: a tenth of the cost of building a
    torch.cuda.Stream object, and these wrappers run ~250 times per training step)"""
    if _raw_stream is not None:
        return ctypes.c_void_p(_raw_stream(torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class _NoSwitch:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NO_SWITCH = _NoSwitch()


def _on(device):
    """context that makes `device` current -- nothing at all when it already is (the usual case: one process per GPU)"""
    if device.index is None or device.index == torch.cuda.current_device():
        return _NO_SWITCH
    return torch.cuda.device(device)


# In-situ kernel timing for bench.py: when enabled, every C-ABI launch is bracketed by CUDA events recorded on the
# launching stream; timings() returns {entry point: [ms, ...]} after a synchronize.  Off by default (zero overhead).
_TIMING = None


def enable_timing(on=True):
    global _TIMING
    _TIMING = [] if on else None


def timings():
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1, _ in (_TIMING or []):
        out.setdefault(name, []).append(e0.elapsed_time(e1))
    return out


def timing_bytes():
    """{entry point: algorithmic bytes summed over its timed calls} for the entry points whose traffic depends on the call's
    shapes (K7b works on tensors of many sizes within one step); call before enable_timing(False)"""
    out = {}
    for name, _, _, nbytes in (_TIMING or []):
        if nbytes:
            out[name] = out.get(name, 0) + nbytes
    return out


# BH_NVTX=1: every C-ABI call sits inside an NVTX range named after its entry point (nsys / ncu --nvtx timelines show
# bh_warp_fwd, bh_bihome_fwd_bwd, ... between the cuDNN kernels of the backbone); off by default
_NVTX = os.environ.get('BH_NVTX', '0') == '1'


class _timed:
    def __init__(self, name, nbytes=0):
        self.name = name
        self.nbytes = nbytes

    def __enter__(self):
        if _NVTX:
            torch.cuda.nvtx.range_push(self.name)
        if _TIMING is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if _TIMING is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            _TIMING.append((self.name, self.e0, e1, self.nbytes))
        if _NVTX:
            torch.cuda.nvtx.range_pop()
        return False


def tune(key, value):
    """microbenchmark / test switch (bh_tune_set in include/bihome_b200.h); 0 restores the default choice"""
    cabi.check(cabi.lib().bh_tune_set(str(key).encode(), int(value)), 'bh_tune_set')


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _need_cuda_f32(name, t):
    if not torch.is_tensor(t):
        raise TypeError('%s must be a tensor, got %r' % (name, type(t)))
    if not t.is_cuda:
        raise RuntimeError('bihome_b200: %s is on %s; the hot path runs on CUDA only (no CPU fallback)' % (name, t.device))
    if t.dtype != torch.float32:
        raise TypeError('bihome_b200: %s must be float32, got %s' % (name, t.dtype))


def _is_nhwc(t):
    return t.dim() == 4 and t.shape[1] > 1 and not t.is_contiguous() and t.is_contiguous(memory_format=torch.channels_last)


def _dense(t):
    """contiguous NCHW or channels-last as is; anything else -> contiguous copy"""
    if t.is_contiguous() or _is_nhwc(t):
        return t
    return t.contiguous()


# ------------------------------------------------------------------------------------------------
# K1
# ------------------------------------------------------------------------------------------------
class _Dlt4(torch.autograd.Function):
    @staticmethod
    def forward(ctx, delta, corners, W, Hh):
        _need_cuda_f32('delta', delta)
        delta = delta.contiguous()
        if corners is not None:
            _need_cuda_f32('corners', corners)
            corners = corners.contiguous()
        B = delta.shape[0]
        H = torch.empty(B, 3, 3, device=delta.device, dtype=torch.float32)
        with _on(delta.device), _timed('bh_dlt4_fwd'):
            cabi.check(cabi.lib().bh_dlt4_fwd(_ptr(corners), _ptr(delta), _ptr(H), B, float(W), float(Hh), _stream()),
                       'bh_dlt4_fwd')
        ctx.save_for_backward(delta, corners, H)
        ctx.size = (float(W), float(Hh))
        return H

    @staticmethod
    def backward(ctx, gH):
        delta, corners, H = ctx.saved_tensors
        gH = gH.contiguous()
        B = delta.shape[0]
        g_delta = torch.empty_like(delta)
        want_c = corners is not None and ctx.needs_input_grad[1]
        g_corners = torch.empty_like(corners) if want_c else None
        with _on(delta.device), _timed('bh_dlt4_bwd'):
            cabi.check(cabi.lib().bh_dlt4_bwd(_ptr(corners), _ptr(delta), _ptr(H), _ptr(gH), _ptr(g_delta), _ptr(g_corners),
                                              B, ctx.size[0], ctx.size[1], _stream()), 'bh_dlt4_bwd')
        return g_delta, g_corners, None, None


def dlt4(delta, corners=None, size=None):
    """delta [B,4,2] (+ corners [B,4,2] or the canonical rectangle of `size`=(W,H)) -> H [B,3,3], h33 = 1."""
    if delta.dim() != 3 or delta.shape[1:] != (4, 2):
        raise ValueError('deltas should be of size B, 4, 2, but got: {}'.format(tuple(delta.shape)))
    if corners is None:
        if size is None:
            raise ValueError('dlt4 needs either corners or size=(W, H)')
        return _Dlt4.apply(delta, None, size[0], size[1])
    if corners.shape != delta.shape:
        raise ValueError('corners should be of size B, 4, 2, but got: {}'.format(tuple(corners.shape)))
    return _Dlt4.apply(delta, corners, 0.0, 0.0)


# ------------------------------------------------------------------------------------------------
# K2
# ------------------------------------------------------------------------------------------------
class _Warp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, H, out_h, out_w, pool, src_hw):
        ctx.set_materialize_grads(False)
        _need_cuda_f32('homography', H)
        Hc = H.contiguous().view(-1, 9)
        B = Hc.shape[0]
        lib = cabi.lib()
        out = mask = None
        C = 0
        nhwc = 0
        if src is not None:
            _need_cuda_f32('image', src)
            src = _dense(src)
            if src.shape[0] != B:
                raise ValueError('image batch %d != homography batch %d' % (src.shape[0], B))
            nhwc = int(_is_nhwc(src))
            C, Hs, Ws = src.shape[1], src.shape[2], src.shape[3]
            out = torch.empty((B, C, out_h, out_w), device=src.device, dtype=torch.float32,
                              memory_format=torch.channels_last if nhwc else torch.contiguous_format)
        else:
            Hs, Ws = src_hw
        if pool:
            mask = torch.empty((B, out_h // pool, out_w // pool), device=Hc.device, dtype=torch.float32)
        with _on(Hc.device), _timed('bh_warp_fwd'):
            cabi.check(lib.bh_warp_fwd(_ptr(src), _ptr(Hc), _ptr(out), _ptr(mask), B, C, Hs, Ws, out_h, out_w,
                                       int(pool or 0), nhwc, _stream()), 'bh_warp_fwd')
        ctx.save_for_backward(src, Hc)
        ctx.geom = (B, C, Hs, Ws, out_h, out_w, int(pool or 0), nhwc)
        if src is None:
            return mask
        if pool:
            return out, mask
        return out

    @staticmethod
    def backward(ctx, *grads):
        src, Hc = ctx.saved_tensors
        B, C, Hs, Ws, Ho, Wo, pool, nhwc = ctx.geom
        if src is None:
            g_out, g_mask = None, grads[0]
        elif pool:
            g_out, g_mask = grads
        else:
            g_out, g_mask = grads[0], None
        lib = cabi.lib()
        if g_out is not None:
            g_out = g_out.contiguous(memory_format=torch.channels_last) if nhwc else g_out.contiguous()
        if g_mask is not None:
            g_mask = g_mask.contiguous()
        gH = torch.empty(B, 9, device=Hc.device, dtype=torch.float32)
        if g_out is None and g_mask is None:
            return None, gH.zero_().view(B, 3, 3), None, None, None, None
        g_src = None
        if src is not None and ctx.needs_input_grad[0] and g_out is not None:
            g_src = torch.zeros_like(src)
        nbytes = lib.bh_warp_bwd_workspace_bytes(B, C, Hs, Ws, Ho, Wo, nhwc)
        ws = torch.empty(max(int(nbytes), 16), device=Hc.device, dtype=torch.uint8)
        with _on(Hc.device), _timed('bh_warp_bwd'):
            cabi.check(lib.bh_warp_bwd(_ptr(src if g_out is not None else None), _ptr(Hc), _ptr(g_out), _ptr(g_mask), _ptr(gH),
                                       _ptr(g_src), B, C, Hs, Ws, Ho, Wo, pool, nhwc, _ptr(ws), int(nbytes), _stream()),
                       'bh_warp_bwd')
        return g_src, gH.view(B, 3, 3), None, None, None, None


def warp(src, H, out_h, out_w, pool=None):
    """out[b,c,y,x] = bilinear(src[b,c], proj(H_b [x,y,1]^T)), zeros outside.

    With ``pool`` also returns the analytic coverage mask warp(ones) average-pooled by ``pool``:
    ``(out, mask [B, out_h/pool, out_w/pool])``.
    """
    if src.dim() != 4:
        raise ValueError('image should be of size B, C, H, W, got {}'.format(tuple(src.shape)))
    return _Warp.apply(src, H, int(out_h), int(out_w), pool, None)


def coverage_mask(H, src_hw, out_hw, pool=1):
    """warp(ones_like(image)) followed by AvgPool2d(pool), computed analytically from H: [B, out_h/pool, out_w/pool]."""
    return _Warp.apply(None, H, int(out_hw[0]), int(out_hw[1]), int(pool), (int(src_hw[0]), int(src_hw[1])))


# ------------------------------------------------------------------------------------------------
# K3
# ------------------------------------------------------------------------------------------------
class _BihomeLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f1, f2, f1w, f2w, m1, m2, m1w, m2w, H12, H21, mu):
        ctx.set_materialize_grads(False)
        for n, t in (('f1', f1), ('f2', f2), ('f1w', f1w), ('f2w', f2w), ('m1w', m1w), ('m2w', m2w), ('H12', H12), ('H21', H21)):
            _need_cuda_f32(n, t)
        nhwc = all(_is_nhwc(t) for t in (f1, f2, f1w, f2w))
        if nhwc:
            feats = [f1, f2, f1w, f2w]
        else:
            feats = [t.contiguous() for t in (f1, f2, f1w, f2w)]
        f1, f2, f1w, f2w = feats
        B, C, h, w = f1w.shape
        m1 = None if m1 is None else m1.contiguous()
        m2 = None if m2 is None else m2.contiguous()
        m1w, m2w = m1w.contiguous(), m2w.contiguous()
        H12c, H21c = H12.contiguous().view(B, 9), H21.contiguous().view(B, 9)
        dev = f1w.device
        loss = torch.empty(B, device=dev, dtype=torch.float32)
        parts = torch.empty(B, 5, device=dev, dtype=torch.float32)
        g_f1w, g_f2w = torch.empty_like(f1w), torch.empty_like(f2w)
        want_in = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        g_f1 = torch.empty_like(f1) if want_in else None
        g_f2 = torch.empty_like(f2) if want_in else None
        g_m1w, g_m2w = torch.empty_like(m1w), torch.empty_like(m2w)
        gH12 = torch.empty(B, 9, device=dev, dtype=torch.float32)
        gH21 = torch.empty(B, 9, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev), _timed('bh_bihome_fwd_bwd'):
            cabi.check(cabi.lib().bh_bihome_fwd_bwd(
                _ptr(f1), _ptr(f2), _ptr(f1w), _ptr(f2w), _ptr(m1), _ptr(m2), _ptr(m1w), _ptr(m2w), _ptr(H12c), _ptr(H21c),
                float(mu), _ptr(loss), _ptr(parts), _ptr(g_f1w), _ptr(g_f2w), _ptr(g_f1), _ptr(g_f2), _ptr(g_m1w),
                _ptr(g_m2w), _ptr(gH12), _ptr(gH21), B, C, h, w, int(nhwc), _stream()), 'bh_bihome_fwd_bwd')
        ctx.grads = (g_f1w, g_f2w, g_f1, g_f2, g_m1w, g_m2w, gH12, gH21)
        ctx.dims = (B, C, h, w)
        ctx.mark_non_differentiable(parts)
        return loss, parts

    @staticmethod
    def backward(ctx, g_loss, _g_parts):
        if g_loss is None:
            return (None,) * 11
        if ctx.grads is None:
            raise RuntimeError('bihome_loss: the fused loss stores its gradients once; backward twice is not supported')
        g_f1w, g_f2w, g_f1, g_f2, g_m1w, g_m2w, gH12, gH21 = ctx.grads
        ctx.grads = None
        B, C, h, w = ctx.dims
        g_loss = g_loss.contiguous().float()
        with _on(g_f1w.device), _timed('bh_bihome_rescale'):
            cabi.check(cabi.lib().bh_bihome_rescale(_ptr(g_loss), _ptr(g_f1w), _ptr(g_f2w), _ptr(g_f1), _ptr(g_f2), _ptr(g_m1w),
                                                    _ptr(g_m2w), _ptr(gH12), _ptr(gH21), B, C, h, w, _stream()),
                       'bh_bihome_rescale')
        return (g_f1, g_f2, g_f1w, g_f2w, None, None, g_m1w, g_m2w, gH12.view(B, 3, 3), gH21.view(B, 3, 3), None)


def bihome_loss(f1, f2, f1w, f2w, m1w, m2w, H12, H21, mu, m1=None, m2=None):
    """Per-sample double-line biHomE loss (margin 'inf', l1, channel-agnostic).

    features [B,C,h,w]; pooled masks [B,h,w] (m1/m2 None == ones); H12,H21 [B,3,3].
    Returns (loss_b [B], parts [B,5] = ln1, ln2, den1 (unclamped), den2, ln3); the reference's scalar is loss_b.sum().
    """
    return _BihomeLoss.apply(f1, f2, f1w, f2w, m1, m2, m1w, m2w, H12, H21, float(mu))


# ------------------------------------------------------------------------------------------------
# K3g
# ------------------------------------------------------------------------------------------------
_DISTANCES = {'l1': 0, 'l2': 1, 'cosine': 2}
_HINGES = {None: 0, 'channel': 1, 'pixel': 2}


class _TripletLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f1, f2, f1w, f2w, a1, b2, a2, b1, H12, H21, cfg):
        ctx.set_materialize_grads(False)
        lines, distance, hinge, margins, mask_crd, mu, scale = cfg
        two = lines == 2
        feats = [f1, f2, f1w] + ([f2w] if two else [])
        for n, t in zip(('f1', 'f2', 'f1w', 'f2w'), feats):
            _need_cuda_f32(n, t)
        _need_cuda_f32('a1', a1)
        nhwc = all(_is_nhwc(t) for t in feats)
        if not nhwc:
            feats = [t.contiguous() for t in feats]
        f1, f2, f1w = feats[:3]
        f2w = feats[3] if two else None
        B, C, h, w = f1w.shape
        ctx.mask_shapes = tuple(None if t is None else tuple(t.shape) for t in (a1, b2, a2, b1))
        dense = lambda t: None if t is None else t.reshape(B, h, w).contiguous()
        a1, b2, a2, b1 = dense(a1), dense(b2), dense(a2) if two else None, dense(b1) if two else None
        dev = f1w.device
        loss = torch.empty(B, device=dev, dtype=torch.float32)
        parts = torch.empty(B, 5, device=dev, dtype=torch.float32)
        g_f1w = torch.empty_like(f1w)
        g_f2w = torch.empty_like(f2w) if two else None
        want_in = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        g_f1 = torch.empty_like(f1) if want_in else None
        g_f2 = torch.empty_like(f2) if want_in else None
        g_a1 = torch.empty_like(a1)
        g_a2 = torch.empty_like(a2) if two else None
        g_b2 = torch.empty_like(b2) if (b2 is not None and ctx.needs_input_grad[5]) else None
        g_b1 = torch.empty_like(b1) if (b1 is not None and ctx.needs_input_grad[7]) else None
        if two:
            H12c, H21c = H12.contiguous().view(B, 9), H21.contiguous().view(B, 9)
            gH12 = torch.empty(B, 9, device=dev, dtype=torch.float32)
            gH21 = torch.empty(B, 9, device=dev, dtype=torch.float32)
        else:
            H12c = H21c = gH12 = gH21 = None
        with torch.cuda.device(dev), _timed('bh_triplet_fwd_bwd'):
            cabi.check(cabi.lib().bh_triplet_fwd_bwd(
                _ptr(f1), _ptr(f2), _ptr(f1w), _ptr(f2w), _ptr(a1), _ptr(b2), _ptr(a2), _ptr(b1), _ptr(H12c), _ptr(H21c),
                lines, _DISTANCES[distance], _HINGES[hinge], int(bool(mask_crd)), float(margins[0]), float(margins[1]),
                float(scale[0]), float(scale[1]), float(mu), _ptr(loss), _ptr(parts), _ptr(g_f1w), _ptr(g_f2w), _ptr(g_f1),
                _ptr(g_f2), _ptr(g_a1), _ptr(g_b2), _ptr(g_a2), _ptr(g_b1), _ptr(gH12), _ptr(gH21), B, C, h, w, int(nhwc),
                _stream()), 'bh_triplet_fwd_bwd')
        ctx.grads = (g_f1w, g_f2w, g_f1, g_f2, g_a1, g_b2, g_a2, g_b1, gH12, gH21)
        ctx.dims = (B, C, h, w)
        ctx.mark_non_differentiable(parts)
        return loss, parts

    @staticmethod
    def backward(ctx, g_loss, _g_parts):
        if g_loss is None:
            return (None,) * 11
        if ctx.grads is None:
            raise RuntimeError('triplet_loss: the fused loss stores its gradients once; backward twice is not supported')
        g_f1w, g_f2w, g_f1, g_f2, g_a1, g_b2, g_a2, g_b1, gH12, gH21 = ctx.grads
        ctx.grads = None
        B, C, h, w = ctx.dims
        g_loss = g_loss.contiguous().float()
        with _on(g_f1w.device), _timed('bh_triplet_rescale'):
            cabi.check(cabi.lib().bh_triplet_rescale(_ptr(g_loss), _ptr(g_f1w), _ptr(g_f2w), _ptr(g_f1), _ptr(g_f2), _ptr(g_a1),
                                                     _ptr(g_b2), _ptr(g_a2), _ptr(g_b1), _ptr(gH12), _ptr(gH21), B, C, h, w, _stream()),
                       'bh_triplet_rescale')
        shaped = lambda g, like: None if g is None else g.view(like)
        ms = ctx.mask_shapes
        return (g_f1, g_f2, g_f1w, g_f2w, shaped(g_a1, ms[0]), shaped(g_b2, ms[1]), shaped(g_a2, ms[2]), shaped(g_b1, ms[3]),
                None if gH12 is None else gH12.view(B, 3, 3), None if gH21 is None else gH21.view(B, 3, 3), None)


def triplet_loss(f1, f2, f1w, f2w, a1, b2, a2=None, b1=None, H12=None, H21=None, lines=2, distance='l1', hinge=None, margin=0.0,
                 mask_crd=False, mu=0.0, scale=(1.0, 1.0)):
    """Per-sample masked triplet loss in the reference's other variants (one-/double-line; l1 / l2 / cosine; margin 'inf'
    (hinge None), channel-aware ('channel') or channel-agnostic / one-line ('pixel') numeric margin; MASK_CRD weights).

    features [B,C,h,w]; masks [B,h,w] or [B,1,h,w] at the feature resolution (a = warped mask of the line's source patch,
    b = mask of its target patch, None == ones); ``margin`` a number or a pair (line 1, line 2); ``scale`` multiplies the
    two lines.  Returns (loss_b [B], parts [B,5] = ln1, ln2, den1 (unclamped), den2, ln3); gradients flow to all four
    features, all four masks and both homographies."""
    if distance not in _DISTANCES:
        raise ValueError('Do not know this distance metric: %r' % (distance,))
    if hinge not in _HINGES:
        raise ValueError('hinge must be None, \'channel\' or \'pixel\', got %r' % (hinge,))
    if hinge == 'channel' and distance != 'l1':
        raise ValueError('a per-channel margin needs per-channel distances (l1)')
    margins = tuple(margin) if isinstance(margin, (tuple, list)) else (margin, margin)
    cfg = (int(lines), distance, hinge, margins, bool(mask_crd), float(mu), (float(scale[0]), float(scale[1])))
    return _TripletLoss.apply(f1, f2, f1w, f2w, a1, b2, a2, b1, H12, H21, cfg)


# ------------------------------------------------------------------------------------------------
# K4
# ------------------------------------------------------------------------------------------------
class _DltN(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p1, p2, field, choice, four):
        ctx.set_materialize_grads(False)
        src = field if field is not None else p2
        _need_cuda_f32('correspondences', src)
        if field is not None:
            field = field.contiguous()
            B, _, Hf, Wf = field.shape
            N = Hf * Wf
        else:
            _need_cuda_f32('points1', p1)
            p1, p2 = p1.detach().contiguous(), p2.contiguous()
            B, N = p2.shape[0], p2.shape[1]
            Wf = 0
        if choice is not None:
            choice = choice.to(torch.int64).contiguous().view(B, -1)
            M = choice.shape[1]
        else:
            M = N
        Hn = torch.empty(B, 3, 3, device=src.device, dtype=torch.float32)
        delta = None
        if four is not None:
            four = four.detach().to(torch.float32).contiguous().view(4, 2)
            delta = torch.empty(B, 4, 2, device=src.device, dtype=torch.float32)
        with _on(src.device), _timed('bh_dltn_fwd'):
            cabi.check(cabi.lib().bh_dltn_fwd(_ptr(p1), _ptr(p2), _ptr(field), _ptr(choice), _ptr(four), _ptr(Hn), _ptr(delta),
                                              B, N, M, Wf, _stream()), 'bh_dltn_fwd')
        ctx.save_for_backward(p1, p2, field, choice, four)
        ctx.dims = (B, N, M, Wf)
        if four is None:
            return Hn
        return Hn, delta

    @staticmethod
    def backward(ctx, gHn, gDelta=None):
        p1, p2, field, choice, four = ctx.saved_tensors
        B, N, M, Wf = ctx.dims
        gHn = None if gHn is None else gHn.contiguous()
        gDelta = None if gDelta is None else gDelta.contiguous()
        g_p2 = g_field = None
        if field is not None:
            g_field = torch.zeros_like(field)
        else:
            g_p2 = torch.zeros_like(p2)
        if gHn is not None or gDelta is not None:
            with torch.cuda.device((field if field is not None else p2).device), _timed('bh_dltn_bwd'):
                cabi.check(cabi.lib().bh_dltn_bwd(_ptr(p1), _ptr(p2), _ptr(field), _ptr(choice), _ptr(four), _ptr(gHn),
                                                  _ptr(gDelta), _ptr(g_p2), _ptr(g_field), B, N, M, Wf, _stream()),
                           'bh_dltn_bwd')
        return None, g_p2, g_field, None, None


def dltn(points1, points2, choice=None):
    """Normalised N-point DLT: points [B,N,2], choice [B,M] sampled indices (None = all) -> H [B,3,3] with
    H[2,2] = 1/(1+1e-8) normalisation as kornia.find_homography_dlt.  Gradient flows to points2 only."""
    return _DltN.apply(points1, points2, None, choice, None)


def dltn_field(field, choice, four_points):
    """Perspective field [B,2,Hf,Wf] + sampled indices [B,M] -> (H [B,3,3], delta_hat [B,4,2]) in one launch
    (forward_map_field + gather + find_homography_dlt + corner projection of the reference head)."""
    return _DltN.apply(None, None, field, choice, four_points)


# ------------------------------------------------------------------------------------------------
# K5
# ------------------------------------------------------------------------------------------------
PAIR_NPARAM = 32


def pairgen_draw(batch, n_img, image_hw, rho, patch_size, max_delta, seed, step, device):
    """Counter-based draws of HomographyNetPrep's random parameters for `batch` samples:
    (params float64 [B,32] in the oracle/pairgen.py pack_params layout, image index int32 [B])."""
    params = torch.empty(batch, PAIR_NPARAM, device=device, dtype=torch.float64)
    index = torch.empty(batch, device=device, dtype=torch.int32)
    with torch.cuda.device(device), _timed('bh_pairgen_draw'):
        cabi.check(cabi.lib().bh_pairgen_draw(_ptr(params), _ptr(index), batch, n_img, image_hw[0], image_hw[1], rho, patch_size,
                                              float(max_delta), int(seed) & (2 ** 64 - 1), int(step), _stream()), 'bh_pairgen_draw')
    return params, index


def pairgen_apply(images, index, params, patch_size, mean=0.443, std=0.129):
    """images uint8 [n,Hi,Wi,3] (RGB, on the GPU), index int32 [B], params float64 [B,32] ->
    (patch_1, patch_2 [B,1,P,P] float32 standardised grayscale, delta [B,4,2]).  The two patch tensors are the two
    halves of one buffer, so the head batches them through the warp without a copy."""
    if not images.is_cuda or images.dtype != torch.uint8 or images.dim() != 4 or images.shape[-1] != 3:
        raise TypeError('images must be a CUDA uint8 tensor [n, H, W, 3]')
    images = images.contiguous()
    B, P = index.shape[0], int(patch_size)
    buf = torch.empty(2 * B, 1, P, P, device=images.device, dtype=torch.float32)
    delta = torch.empty(B, 4, 2, device=images.device, dtype=torch.float32)
    p1, p2 = buf[:B], buf[B:]
    with _on(images.device), _timed('bh_pairgen_apply'):
        cabi.check(cabi.lib().bh_pairgen_apply(_ptr(images), _ptr(index.contiguous()), _ptr(params.contiguous()), _ptr(p1), _ptr(p2),
                                               _ptr(delta), B, images.shape[0], images.shape[1], images.shape[2], P,
                                               float(mean), float(std), _stream()), 'bh_pairgen_apply')
    return p1, p2, delta


def pairgen_image(images, index, params, mean=0.443, std=0.129):
    """The whole first image of every pair through its photometric chain, grayscale, standardised: [B,1,Hi,Wi] -- the
    'image_1' entry of the reference's batch (HomographyNetPrep + DictToGrayscale + DictStandardize + DictToTensor) that
    PhotometricHead warps (src/heads/PhotometricHead.py:24)."""
    if not images.is_cuda or images.dtype != torch.uint8:
        raise RuntimeError('bihome_b200: the image pool must be a CUDA uint8 tensor [n,H,W,3] (no CPU fallback)')
    n, hi, wi, _ = images.shape
    B = index.shape[0]
    out = torch.empty(B, 1, hi, wi, device=images.device, dtype=torch.float32)
    with _on(images.device), _timed('bh_pairgen_image'):
        cabi.check(cabi.lib().bh_pairgen_image(_ptr(images), _ptr(index.contiguous()), _ptr(params.contiguous()), _ptr(out), B, n, hi, wi,
                                               float(mean), float(std), _stream()), 'bh_pairgen_image')
    return out


# ------------------------------------------------------------------------------------------------
def mace(delta_gt, delta_hat):
    """mean corner error over B*4 corners, as a 0-dim CUDA tensor (no host sync)."""
    _need_cuda_f32('delta_gt', delta_gt)
    _need_cuda_f32('delta_hat', delta_hat)
    a, b = delta_gt.detach().contiguous(), delta_hat.detach().contiguous()
    out = torch.empty(1, device=a.device, dtype=torch.float32)
    with _on(a.device), _timed('bh_mace'):
        cabi.check(cabi.lib().bh_mace(_ptr(a), _ptr(b), _ptr(out), a.numel() // 8, _stream()), 'bh_mace')
    return out[0]


# ------------------------------------------------------------------------------------------------
# K6: perspective-field head (Conv2d(16,128,1) -> BatchNorm2d -> ReLU -> Conv2d(128,2,1)) as a per-pixel kernel
# ------------------------------------------------------------------------------------------------
def _fh_moments(x):
    """x [B,C,H,W] channels-last -> (sum_p x_i [C], sum_p x_i x_k [C,C]) in float64"""
    _need_cuda_f32('x', x)
    B, C, H, W = x.shape
    n = B * H * W
    lib = cabi.lib()
    grid = lib.bh_fieldhead_grid(0, n)
    parts = torch.empty(grid, C + C * C, device=x.device, dtype=torch.float64)
    with _on(x.device), _timed('bh_fieldhead_moments'):
        cabi.check(lib.bh_fieldhead_moments(_ptr(x), _ptr(parts), n, C, _stream()), 'bh_fieldhead_moments')
    s = parts.sum(0)
    return s[:C], s[C:].view(C, C)


def _fh_tf32():
    """K6 replaces two cuDNN convolutions: it multiplies in TF32 exactly when they would (torch.backends.cudnn.allow_tf32,
    torch's default) and float32-faithfully otherwise"""
    return int(bool(torch.backends.cudnn.allow_tf32))


def _fh_fwd(x, W1, b1, W2, b2):
    """x [B,C,H,W] channels-last, folded W1 [hid,C], b1 [hid], W2 [2,hid], b2 [2] -> field [B,2,H,W] (planar)"""
    B, C, H, W = x.shape
    out = torch.empty(B, 2, H, W, device=x.device, dtype=torch.float32)
    with _on(x.device), _timed('bh_fieldhead_fwd'):
        cabi.check(cabi.lib().bh_fieldhead_fwd(_ptr(x), _ptr(W1), _ptr(b1), _ptr(W2), _ptr(b2), _ptr(out), B, H * W, C,
                                               W1.shape[0], _fh_tf32(), _stream()), 'bh_fieldhead_fwd')
    return out


def _fh_bwd(x, W1, b1, W2, g_out):
    """-> (gx like x, gW1 [hid,C], gb1 [hid], gW2 [2,hid], gb2 [2]) for the folded weights"""
    B, C, H, W = x.shape
    hid = W1.shape[0]
    lib = cabi.lib()
    grid = lib.bh_fieldhead_grid(1, B * H * W)
    parts = torch.empty(grid, hid * C + 3 * hid + 2, device=x.device, dtype=torch.float32)
    gx = torch.empty_like(x)
    with _on(x.device), _timed('bh_fieldhead_bwd'):
        cabi.check(lib.bh_fieldhead_bwd(_ptr(x), _ptr(W1), _ptr(b1), _ptr(W2), _ptr(g_out), _ptr(gx), _ptr(parts), B, H * W, C,
                                        hid, _fh_tf32(), _stream()), 'bh_fieldhead_bwd')
    s = parts.sum(0)
    return (gx, s[:hid * C].view(hid, C), s[hid * C:hid * C + hid], s[hid * C + hid:hid * C + 3 * hid].view(2, hid),
            s[hid * C + 3 * hid:])


def _fh_affine(x, a, M, gx):
    """gx += a + M x per pixel, in place"""
    B, C, H, W = x.shape
    with _on(x.device), _timed('bh_fieldhead_affine'):
        cabi.check(cabi.lib().bh_fieldhead_affine(_ptr(x), _ptr(a), _ptr(M), _ptr(gx), B * H * W, C, 1, _stream()),
                   'bh_fieldhead_affine')
    return gx


def _fold_batchnorm(W1, b1, gamma, beta, mean_y, var_y, eps):
    """BatchNorm(y) with y = W1 x + b1 as a rescaled first layer: (W1', b1')"""
    scale = gamma / torch.sqrt(var_y + eps)
    return W1 * scale[:, None], (b1 - mean_y) * scale + beta


def _hidden_statistics(W1, b1, s1, s2, n):
    """mean and (biased) variance over all pixels of y = W1 x + b1, from the input's first and second moments"""
    mean_x = s1 / n
    cov_x = s2 / n - torch.outer(mean_x, mean_x)
    return W1 @ mean_x + b1, ((W1 @ cov_x) * W1).sum(1)


class _FieldHead(torch.autograd.Function):
    """out = conv2(relu(bn(conv1(x)))) for 1x1 convolutions, without the hidden tensor (csrc/fieldhead.cu).
    The tiny algebra (fold, statistics and their adjoints) runs in float64 torch ops."""

    @staticmethod
    def forward(ctx, x, W1, b1, gamma, beta, W2, b2, running_mean, running_var, training, eps):
        B, C, H, W = x.shape
        n = B * H * W
        hid = W1.shape[0]
        d = torch.float64
        W1d, b1d, gd, bd = W1.reshape(hid, C).to(d), b1.to(d), gamma.to(d), beta.to(d)
        if training:
            s1, s2 = _fh_moments(x)
            mean_y, var_y = _hidden_statistics(W1d, b1d, s1, s2, n)
            stats = (s1, s2)
        else:
            mean_y, var_y = running_mean.to(d), running_var.to(d)
            stats = (mean_y, var_y)
        W1f, b1f = _fold_batchnorm(W1d, b1d, gd, bd, mean_y, var_y, eps)
        out = _fh_fwd(x, W1f.to(x.dtype).contiguous(), b1f.to(x.dtype).contiguous(), W2.reshape(2, hid).contiguous(), b2.contiguous())
        ctx.save_for_backward(x, W1, b1, gamma, beta, W2, *stats)
        ctx.cfg = (training, eps, n)
        ctx.mark_non_differentiable(mean_y, var_y)
        return out, mean_y, var_y

    @staticmethod
    def backward(ctx, g_out, _g_mean, _g_var):
        x, W1, b1, gamma, beta, W2, st0, st1 = ctx.saved_tensors
        training, eps, n = ctx.cfg
        B, C, H, W = x.shape
        hid = W1.shape[0]
        d = torch.float64
        with torch.enable_grad():
            leaves = [t.detach().to(d).requires_grad_(True) for t in (W1.reshape(hid, C), b1, gamma, beta)]
            if training:
                s1, s2 = st0.detach().requires_grad_(True), st1.detach().requires_grad_(True)
                mean_y, var_y = _hidden_statistics(leaves[0], leaves[1], s1, s2, n)
                leaves += [s1, s2]
            else:
                mean_y, var_y = st0, st1
            W1f, b1f = _fold_batchnorm(leaves[0], leaves[1], leaves[2], leaves[3], mean_y, var_y, eps)
        gx, gW1f, gb1f, gW2, gb2 = _fh_bwd(x, W1f.detach().to(x.dtype).contiguous(), b1f.detach().to(x.dtype).contiguous(),
                                           W2.reshape(2, hid).contiguous(), g_out.contiguous())
        grads = torch.autograd.grad([W1f, b1f], leaves, [gW1f.to(d), gb1f.to(d)])
        if training:
            g1, g2 = grads[4], grads[5]
            gx = _fh_affine(x, g1.to(x.dtype).contiguous(), (g2 + g2.t()).to(x.dtype).contiguous(), gx)
        # gradients in the parameters' own strides (a channels-last model keeps its 1x1 kernels as [hid,cin,1,1] with strides
        # [cin,1,cin,cin]): DDP's bucket views expect exactly that layout
        like = lambda g, p: torch.empty_like(p).copy_(g.to(p.dtype).reshape(p.shape))
        return (gx, like(grads[0], W1), grads[1].to(b1.dtype), grads[2].to(gamma.dtype),
                grads[3].to(beta.dtype), like(gW2, W2), gb2.to(W2.dtype), None, None, None, None)


def field_head_enabled(device=None):
    """does the Zeng backbone's last stage run on K6 on this device?  BH_FIELD_HEAD=fused / aten force it; otherwise a
    one-off self-test in a child process decides (bihome_b200/autotune.py: parity with the ATen modules, then speed)"""
    forced = os.environ.get('BH_FIELD_HEAD', 'auto')
    if forced in ('fused', 'aten'):
        return forced == 'fused'
    if device is None or torch.device(device).type != 'cuda':
        return False
    from . import autotune
    return autotune.field_head_choice(torch.device(device)) == 'fused'


def field_head_supported(stage, x):
    """stage = nn.Sequential(Conv2d 1x1, BatchNorm2d, ReLU, Conv2d 1x1 -> 2): is this the geometry K6 is compiled for?"""
    if len(stage) != 4 or not x.is_cuda or x.dtype != torch.float32 or x.dim() != 4:
        return False
    c1, bn, act, c2 = stage[0], stage[1], stage[2], stage[3]
    ok = (isinstance(c1, torch.nn.Conv2d) and isinstance(c2, torch.nn.Conv2d) and isinstance(bn, torch.nn.BatchNorm2d)
          and isinstance(act, torch.nn.ReLU) and c1.kernel_size == (1, 1) and c2.kernel_size == (1, 1)
          and c1.stride == (1, 1) and c2.stride == (1, 1) and c1.bias is not None and c2.bias is not None
          and c2.out_channels == 2 and bn.affine and bn.momentum is not None
          and (bn.training or bn.running_mean is not None))
    return bool(ok and cabi.lib().bh_fieldhead_supported(c1.in_channels, c1.out_channels))


def field_head(stage, x):
    """the four modules of ``stage`` applied to x [B,16,H,W] -> field [B,2,H,W] (planar), same parameters, same
    running-statistics update as nn.BatchNorm2d, no [B,128,H,W] tensor"""
    c1, bn, c2 = stage[0], stage[1], stage[3]
    x = x.contiguous(memory_format=torch.channels_last)
    training = bn.training or bn.running_mean is None
    out, mean_y, var_y = _FieldHead.apply(x, c1.weight, c1.bias, bn.weight, bn.bias, c2.weight, c2.bias, bn.running_mean,
                                          bn.running_var, training, float(bn.eps))
    if bn.training and bn.track_running_stats and bn.running_mean is not None:
        n = x.shape[0] * x.shape[2] * x.shape[3]
        with torch.no_grad():       # nn.BatchNorm2d's update: biased variance normalises, unbiased variance is tracked
            bn.num_batches_tracked.add_(1)
            m = float(bn.momentum)
            bn.running_mean.mul_(1 - m).add_(mean_y.to(bn.running_mean.dtype), alpha=m)
            bn.running_var.mul_(1 - m).add_((var_y * (n / max(n - 1, 1))).to(bn.running_var.dtype), alpha=m)
    return out


# ------------------------------------------------------------------------------------------------
# K7: ResNet stem, BatchNorm2d (batch statistics) -> ReLU -> MaxPool2d(3, 2, 1) as one channels-last stage
# ------------------------------------------------------------------------------------------------
def _stem_ws(C, device):
    return torch.empty(int(cabi.lib().bh_stem_workspace_bytes(C)), device=device, dtype=torch.uint8)


class _Stem(torch.autograd.Function):
    """y = maxpool3x3/2(relu(batch_norm(x))) for channels-last x (csrc/stem.cu); the running statistics are updated in place"""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, momentum, eps):
        N, C, H, W = x.shape
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        y = torch.empty((N, C, Ho, Wo), device=x.device, dtype=torch.float32, memory_format=torch.channels_last)
        want_bwd = any(ctx.needs_input_grad[:3])
        code = torch.empty((N, Ho, Wo, C), device=x.device, dtype=torch.uint8) if want_bwd else None
        stats = torch.empty((4, C), device=x.device, dtype=torch.float32)
        ws = _stem_ws(C, x.device)
        with _on(x.device), _timed('bh_stem_fwd'):
            cabi.check(cabi.lib().bh_stem_fwd(_ptr(x), _ptr(gamma), _ptr(beta), _ptr(running_mean), _ptr(running_var), float(momentum),
                                              float(eps), _ptr(y), _ptr(code), _ptr(stats), _ptr(ws), ws.numel(), N, H, W, C,
                                              _stream()), 'bh_stem_fwd')
        if want_bwd:
            ctx.save_for_backward(x, stats, code)
            ctx.affine = (gamma is not None, beta is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, stats, code = ctx.saved_tensors
        N, C, H, W = x.shape
        if not _is_nhwc(gy) and not (C == 1 and gy.is_contiguous()):
            gy = gy.contiguous(memory_format=torch.channels_last)
        gx = torch.empty_like(x)
        has_g, has_b = ctx.affine
        gg = torch.empty(C, device=x.device, dtype=torch.float32) if has_g and ctx.needs_input_grad[1] else None
        gb = torch.empty(C, device=x.device, dtype=torch.float32) if has_b and ctx.needs_input_grad[2] else None
        ws = _stem_ws(C, x.device)
        with _on(x.device), _timed('bh_stem_bwd'):
            cabi.check(cabi.lib().bh_stem_bwd(_ptr(x), _ptr(stats), _ptr(code), _ptr(gy), _ptr(gx), _ptr(gg), _ptr(gb), _ptr(ws),
                                              ws.numel(), N, H, W, C, _stream()), 'bh_stem_bwd')
        return gx, gg, gb, None, None, None, None


def stem_supported(bn, pool, x):
    """can K7 stand in for ``pool(relu(bn(x)))``?  BatchNorm2d on batch statistics with a float momentum, MaxPool2d(3, 2, 1),
    a channels-last float32 CUDA tensor whose channel count the library is compiled for.  BH_STEM=aten keeps the modules."""
    if os.environ.get('BH_STEM', 'fused') == 'aten':
        return False
    if not (torch.is_tensor(x) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and _is_nhwc(x)):
        return False
    nn = torch.nn
    if not (isinstance(bn, nn.BatchNorm2d) and isinstance(pool, nn.MaxPool2d)):
        return False
    if not (bn.training or bn.running_mean is None) or bn.momentum is None or bn.num_features != x.shape[1]:
        return False
    if bn.weight is not None and bn.weight.dtype != torch.float32:
        return False
    as2 = lambda v: (v, v) if isinstance(v, int) else tuple(v)
    if (as2(pool.kernel_size) != (3, 3) or as2(pool.stride) != (2, 2) or as2(pool.padding) != (1, 1) or as2(pool.dilation) != (1, 1)
            or pool.ceil_mode or pool.return_indices):
        return False
    return bool(cabi.lib().bh_stem_supported(int(x.shape[1])))


def stem(bn, x):
    """``max_pool2d(relu(bn(x)), 3, 2, 1)`` in training mode: same output, same gradients (input, weight, bias), same
    running-statistics update (incl. num_batches_tracked) as the three modules; call only when stem_supported() said yes"""
    track = bn.track_running_stats and bn.running_mean is not None
    if track and bn.num_batches_tracked is not None:
        with torch.no_grad():
            bn.num_batches_tracked.add_(1)
    return _Stem.apply(x, bn.weight, bn.bias, bn.running_mean if track else None, bn.running_var if track else None,
                       float(bn.momentum), float(bn.eps))


# ------------------------------------------------------------------------------------------------
# K7b: BatchNorm2d (batch statistics) [+ residual] -> ReLU, the inner stages of the residual blocks
# ------------------------------------------------------------------------------------------------
_L2_BYTES = 126e6      # B200: what a second pass can still find on chip


def _bn_bytes(x, reread, reads, writes):
    """algorithmic bytes of a two-pass BatchNorm stage over tensors of x's size (bench.py's per-call accounting): `reads` input
    tensors and `writes` output tensors once each, plus whatever part of the `reread` tensors the second pass needs again
    cannot have stayed in the L2 (the batch statistics need the whole tensor before the first output element, so the second
    pass is compulsory; its traffic is not, up to the L2's capacity)"""
    t = x.numel() * 4
    return t * (reads + writes) + max(0.0, reread * t - _L2_BYTES)


class _BnAct(torch.autograd.Function):
    """y = relu(batch_norm(x) [+ residual]) for channels-last tensors (csrc/stem.cu, K7b)"""

    @staticmethod
    def forward(ctx, x, residual, gamma, beta, running_mean, running_var, momentum, eps):
        N, C, H, W = x.shape
        n_pix = N * H * W
        y = torch.empty_like(x)
        stats = torch.empty((4, C), device=x.device, dtype=torch.float32)
        ws = _stem_ws(C, x.device)
        with _on(x.device), _timed('bh_bnact_fwd', _bn_bytes(x, 1, 1 if residual is None else 2, 1)):
            cabi.check(cabi.lib().bh_bnact_fwd(_ptr(x), _ptr(residual), _ptr(gamma), _ptr(beta), _ptr(running_mean), _ptr(running_var),
                                               float(momentum), float(eps), _ptr(y), _ptr(stats), _ptr(ws), ws.numel(), n_pix, C,
                                               _stream()), 'bh_bnact_fwd')
        if any(ctx.needs_input_grad[:4]):
            if residual is None:
                ctx.save_for_backward(x, stats)
            else:
                ctx.save_for_backward(x, stats, y)
            ctx.has_residual = residual is not None
            ctx.affine = (gamma is not None, beta is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        saved = ctx.saved_tensors
        x, stats = saved[0], saved[1]
        y = saved[2] if ctx.has_residual else None
        N, C, H, W = x.shape
        if gy.stride() != x.stride():
            gy = gy.contiguous(memory_format=torch.channels_last)
        gx = torch.empty_like(x)
        gr = torch.empty_like(x) if ctx.has_residual else None
        has_g, has_b = ctx.affine
        gg = torch.empty(C, device=x.device, dtype=torch.float32) if has_g and ctx.needs_input_grad[2] else None
        gb = torch.empty(C, device=x.device, dtype=torch.float32) if has_b and ctx.needs_input_grad[3] else None
        ws = _stem_ws(C, x.device)
        with _on(x.device), _timed('bh_bnact_bwd', _bn_bytes(x, 2, 3 if ctx.has_residual else 2, 2 if ctx.has_residual else 1)):
            cabi.check(cabi.lib().bh_bnact_bwd(_ptr(x), _ptr(y), _ptr(stats), _ptr(gy), _ptr(gx), _ptr(gr), _ptr(gg), _ptr(gb), _ptr(ws),
                                               ws.numel(), N * H * W, C, _stream()), 'bh_bnact_bwd')
        return gx, (gr if ctx.needs_input_grad[1] else None), gg, gb, None, None, None, None


def bnact_supported(bn, x, residual=None):
    """can K7b stand in for ``relu(bn(x))`` / ``relu(bn(x) + residual)``?  BatchNorm2d on batch statistics with a float momentum,
    channels-last float32 CUDA tensors, a channel count the library is compiled for.  BH_BNACT=aten keeps the modules."""
    if os.environ.get('BH_BNACT', 'fused') == 'aten':
        return False
    if not (torch.is_tensor(x) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and _is_nhwc(x)):
        return False
    if not isinstance(bn, torch.nn.BatchNorm2d) or bn.momentum is None or bn.num_features != x.shape[1]:
        return False
    if not (bn.training or bn.running_mean is None):
        return False
    if bn.weight is not None and bn.weight.dtype != torch.float32:
        return False
    if residual is not None and not (torch.is_tensor(residual) and residual.shape == x.shape and residual.dtype == x.dtype
                                     and residual.device == x.device):
        return False
    return bool(cabi.lib().bh_stem_supported(int(x.shape[1])))


def bn_relu(bn, x, residual=None):
    """``relu(bn(x))`` or ``relu(bn(x) + residual)`` in training mode: same output, gradients (x, residual, weight, bias) and
    running-statistics update as the modules; call only when bnact_supported() said yes"""
    track = bn.track_running_stats and bn.running_mean is not None
    if track and bn.num_batches_tracked is not None:
        with torch.no_grad():
            bn.num_batches_tracked.add_(1)
    if residual is not None and residual.stride() != x.stride():
        residual = residual.contiguous(memory_format=torch.channels_last)
    return _BnAct.apply(x, residual, bn.weight, bn.bias, bn.running_mean if track else None, bn.running_var if track else None,
                        float(bn.momentum), float(bn.eps))


class _BnAct2(torch.autograd.Function):
    """y = relu(batch_norm_a(a) + batch_norm_b(b)) for channels-last tensors (csrc/stem.cu, K7c)"""

    @staticmethod
    def forward(ctx, a, b, ga, ba, rma, rva, mom_a, eps_a, gb, bb, rmb, rvb, mom_b, eps_b):
        N, C, H, W = a.shape
        n_pix = N * H * W
        y = torch.empty_like(a)
        stats = torch.empty((2, 4, C), device=a.device, dtype=torch.float32)
        ws = torch.empty(2 * int(cabi.lib().bh_stem_workspace_bytes(C)), device=a.device, dtype=torch.uint8)
        with _on(a.device), _timed('bh_bnact2_fwd', _bn_bytes(a, 2, 2, 1)):
            cabi.check(cabi.lib().bh_bnact2_fwd(_ptr(a), _ptr(b), _ptr(ga), _ptr(ba), _ptr(rma), _ptr(rva), float(mom_a), float(eps_a),
                                                _ptr(gb), _ptr(bb), _ptr(rmb), _ptr(rvb), float(mom_b), float(eps_b), _ptr(y),
                                                _ptr(stats[0]), _ptr(stats[1]), _ptr(ws), ws.numel(), n_pix, C, _stream()), 'bh_bnact2_fwd')
        if any(ctx.needs_input_grad):
            ctx.save_for_backward(a, b, y, stats)
            ctx.affine = (ga is not None, ba is not None, gb is not None, bb is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        a, b, y, stats = ctx.saved_tensors
        N, C, H, W = a.shape
        if gy.stride() != a.stride():
            gy = gy.contiguous(memory_format=torch.channels_last)
        g_a, g_b = torch.empty_like(a), torch.empty_like(a)
        need = ctx.needs_input_grad
        new = lambda have, k: torch.empty(C, device=a.device, dtype=torch.float32) if have and need[k] else None
        gga, gba, ggb, gbb = new(ctx.affine[0], 2), new(ctx.affine[1], 3), new(ctx.affine[2], 8), new(ctx.affine[3], 9)
        ws = torch.empty(2 * int(cabi.lib().bh_stem_workspace_bytes(C)), device=a.device, dtype=torch.uint8)
        with _on(a.device), _timed('bh_bnact2_bwd', _bn_bytes(a, 4, 4, 2)):
            cabi.check(cabi.lib().bh_bnact2_bwd(_ptr(a), _ptr(b), _ptr(y), _ptr(stats[0]), _ptr(stats[1]), _ptr(gy), _ptr(g_a), _ptr(g_b),
                                                _ptr(gga), _ptr(gba), _ptr(ggb), _ptr(gbb), _ptr(ws), ws.numel(), N * H * W, C, _stream()),
                       'bh_bnact2_bwd')
        return (g_a, g_b, gga, gba, None, None, None, None, ggb, gbb, None, None, None, None)


def bnact2_enabled():
    """BH_BNACT2=aten keeps the skip path's BatchNorm on cuDNN (K7b then takes its output as the residual)"""
    return os.environ.get('BH_BNACT2', 'fused') != 'aten' and os.environ.get('BH_BNACT', 'fused') != 'aten'


def bn_bn_relu(bn_a, a, bn_b, b):
    """``relu(bn_a(a) + bn_b(b))`` in training mode (the end of a residual block whose skip has its own BatchNorm); call only
    when bnact_supported(bn_a, a) and bnact_supported(bn_b, b) said yes and the two tensors have the same shape and strides"""
    args = []
    for bn in (bn_a, bn_b):
        track = bn.track_running_stats and bn.running_mean is not None
        if track and bn.num_batches_tracked is not None:
            with torch.no_grad():
                bn.num_batches_tracked.add_(1)
        args += [bn.weight, bn.bias, bn.running_mean if track else None, bn.running_var if track else None, float(bn.momentum), float(bn.eps)]
    return _BnAct2.apply(a, b, *args)


# ------------------------------------------------------------------------------------------------
# K8: the bias of a transposed convolution (cuDNN has no bias epilogue for it), channels-last
# ------------------------------------------------------------------------------------------------
class _ChannelBias(torch.autograd.Function):
    """y += bias[c] in place on a channels-last tensor; the bias gradient is the per-channel sum of the upstream gradient,
    which passes through to y unchanged (csrc/stem.cu, K8)"""

    @staticmethod
    def forward(ctx, y, bias):
        N, C, H, W = y.shape
        ctx.mark_dirty(y)
        with _on(y.device), _timed('bh_bias_add', 2 * y.numel() * 4):
            cabi.check(cabi.lib().bh_bias_add(_ptr(y), _ptr(bias), N * H * W, C, _stream()), 'bh_bias_add')
        return y

    @staticmethod
    def backward(ctx, gy):
        gb = None
        if ctx.needs_input_grad[1]:
            N, C, H, W = gy.shape
            g = gy if _is_nhwc(gy) else gy.contiguous(memory_format=torch.channels_last)
            gb = torch.empty(C, device=gy.device, dtype=torch.float32)
            ws = _stem_ws(C, gy.device)
            with _on(gy.device), _timed('bh_bias_grad', g.numel() * 4):
                cabi.check(cabi.lib().bh_bias_grad(_ptr(g), _ptr(gb), _ptr(ws), ws.numel(), N * H * W, C, _stream()), 'bh_bias_grad')
        return gy, gb


def convt_bias_supported(m, x):
    """can K8 add the bias of ``m`` = nn.ConvTranspose2d?  a float32 CUDA input, a float32 bias, zero padding mode, an output
    channel count the library is compiled for.  BH_CONVT_BIAS=aten keeps the module as it is."""
    if os.environ.get('BH_CONVT_BIAS', 'fused') == 'aten':
        return False
    if not (isinstance(m, torch.nn.ConvTranspose2d) and m.bias is not None and m.bias.dtype == torch.float32
            and m.padding_mode == 'zeros'):
        return False
    if not (torch.is_tensor(x) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and m.bias.device == x.device):
        return False
    if m.bias.data_ptr() % 16 != 0 or not m.bias.is_contiguous():     # a bias that is a view into a flat parameter buffer
        return False
    return bool(cabi.lib().bh_stem_supported(int(m.out_channels)))


def conv_transpose_bias(m, x):
    """``m(x)`` for m = nn.ConvTranspose2d with a bias: the transposed convolution itself stays on cuDNN (without its bias),
    the bias is added in place by K8 and its gradient reduced by K8.  Same output and gradients as the module; call only
    when convt_bias_supported() said yes.  An output that cuDNN did not leave channels-last takes ATen's broadcast add."""
    y = torch.nn.functional.conv_transpose2d(x, m.weight, None, m.stride, m.padding, m.output_padding, m.groups, m.dilation)
    if not _is_nhwc(y) or y.dtype != torch.float32:          # NCHW output, or a reduced-precision one under autocast
        return y + m.bias.view(1, -1, 1, 1).to(y.dtype)
    return _ChannelBias.apply(y, m.bias)
