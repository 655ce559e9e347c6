"""K6 -- the Zeng backbone's last stage (Conv2d(16,128,1) -> BatchNorm2d -> ReLU -> Conv2d(128,2,1), reference
src/backbones/Rethinking.py:144-147) computed per pixel with BatchNorm's batch statistics derived from the moments of
the 16-channel input.

CPU part: the host algebra of ``functional._FieldHead`` (statistics from moments, fold, their adjoints, running-statistics
update) against the four ATen modules in float64, with the device entry points replaced by torch ops
(tests/cpu_kernels.py); the kernels themselves on a host emulator (the .cu file compiled for the host, checked against
torch ops and under ThreadSanitizer); the per-device choice between K6 and the ATen modules (bihome_b200/autotune.py).
The kernels were written after the last GPU minutes of round 1 were spent: this file is what pins them without a GPU,
tests/test_gpu_zzz_field_head.py is the device side.
"""
import copy
import os

import pytest
import torch

import cpu_kernels
from conftest import rel_l2

def make_stage(dtype, device='cpu', seed=0):
    torch.manual_seed(seed)
    nn = torch.nn
    stage = nn.Sequential(nn.Conv2d(16, 128, 1), nn.BatchNorm2d(128), nn.ReLU(), nn.Conv2d(128, 2, 1))
    with torch.no_grad():
        stage[1].weight.uniform_(0.5, 1.5)
        stage[1].bias.normal_()
    return stage.to(dtype).to(device)


def mute_knife_edge_pixels(g, pre, eps):
    """zero the upstream gradient of every pixel that has a hidden unit within `eps` of its ReLU threshold: whether such
    a unit counts as active depends on the last bit of the pre-activation (summation order, float32 vs float64), and
    a single switched unit would move the compared gradients by far more than the tolerances below"""
    near = (pre.abs() < eps).any(dim=1, keepdim=True)              # pre [B,hid,H,W] -> [B,1,H,W]
    return g * (~near).to(g.dtype)


def hidden_preactivation(stage, x):
    """BatchNorm output of `stage` in float64 without touching its buffers (batch statistics in training)"""
    conv, bn = stage[0], stage[1]
    y = torch.nn.functional.conv2d(x.double(), conv.weight.double(), conv.bias.double())
    if bn.training:
        mean, var = y.mean(dim=(0, 2, 3)), y.var(dim=(0, 2, 3), unbiased=False)
    else:
        mean, var = bn.running_mean.double(), bn.running_var.double()
    scale = bn.weight.double() / torch.sqrt(var + bn.eps)
    return (y - mean.view(1, -1, 1, 1)) * scale.view(1, -1, 1, 1) + bn.bias.double().view(1, -1, 1, 1)


def compare_stage(F, stage, dtype, device, shape, tol, modes=('train', 'train', 'eval')):
    ref = copy.deepcopy(stage)
    gen = torch.Generator().manual_seed(3)
    for mode in modes:
        getattr(stage, mode)()
        getattr(ref, mode)()
        x = torch.relu(torch.randn(*shape, generator=gen) + 0.3).to(dtype).to(device).contiguous(memory_format=torch.channels_last)
        g = torch.randn(shape[0], 2, shape[2], shape[3], generator=gen).to(dtype).to(device)
        g = mute_knife_edge_pixels(g, hidden_preactivation(ref, x), 1e-9 if dtype == torch.float64 else 1e-4)
        xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        oa, ob = F.field_head(stage, xa), ref(xb)
        assert oa.shape == ob.shape and oa.is_contiguous()
        assert rel_l2(oa.detach().cpu().numpy(), ob.detach().cpu().numpy()) < tol, mode
        (oa * g).sum().backward()
        (ob * g).sum().backward()
        assert rel_l2(xa.grad.cpu().numpy(), xb.grad.cpu().numpy()) < 10 * tol, mode
        scale = max(float(q.grad.abs().max()) for q in ref.parameters())
        for (name, p), (_, q) in zip(stage.named_parameters(), ref.named_parameters()):
            # conv1's bias is removed by a batch-statistics BatchNorm: its gradient is exactly zero, compare absolutely
            err = float((p.grad - q.grad).abs().max()) / max(float(q.grad.abs().max()), 1e-3 * scale)
            assert err < 30 * tol, (mode, name, err)
            p.grad = q.grad = None
        for (name, p), (_, q) in zip(stage.named_buffers(), ref.named_buffers()):
            assert float((p.double() - q.double()).abs().max()) < 10 * tol, (mode, name)


def test_host_algebra_matches_the_aten_modules(monkeypatch):
    F = cpu_kernels.install(monkeypatch)
    compare_stage(F, make_stage(torch.float64), torch.float64, 'cpu', (3, 16, 12, 10), 1e-12)


@pytest.mark.parametrize('dtype,tol,gtol', [(torch.float64, 1e-11, 1e-11), (torch.float32, 2e-5, 1e-2)])
def test_aten_path_drops_the_cancelled_bias_exactly(dtype, tol, gtol):
    """default path: in training layer8's first convolution runs without its bias (a constant in front of a batch-
    statistics BatchNorm cancels) -- outputs, gradients and running statistics equal the plain nn.Sequential, the
    bias' gradient is exactly zero instead of round-off noise, eval mode still applies the bias"""
    from bihome_b200.backbones import Rethinking
    kw = dict(IMAGE_SIZE=128, PATCH_KEYS=['patch_1', 'patch_2'], TARGET_KEYS=['pf_hat_12', 'pf_hat_21'], RESNET_BLOCK='ResNet34',
              PRETRAINED_RESNET=False, VARIANT='OneLine')
    torch.manual_seed(0)
    net = Rethinking.Model(**kw).to(dtype).train()
    with torch.no_grad():
        net.layer8[0].bias.normal_()
    ref = copy.deepcopy(net)
    gen = torch.Generator().manual_seed(2)
    for _ in range(2):
        x = torch.rand(2, 16, 32, 32, generator=gen).to(dtype)
        xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        a, b = net._layer8_aten(xa), ref.layer8(xb)
        g = torch.randn(a.shape, generator=gen).to(dtype)
        (a * g).sum().backward()
        (b * g).sum().backward()
        assert rel_l2(a.detach().numpy(), b.detach().numpy()) < tol
        # float32: a pre-activation within round-off of zero may switch its ReLU between the two evaluations; one switch
        # moves the gradient of one pixel by ~10 % (3e-3 of the whole map) -- the float64 case is the exactness check
        assert rel_l2(xa.grad.numpy(), xb.grad.numpy()) < gtol
        for (name, p), (_, q) in zip(net.layer8.named_parameters(), ref.layer8.named_parameters()):
            if name == '0.bias':
                assert float(p.grad.abs().max()) == 0.0
            else:
                assert rel_l2(p.grad.numpy(), q.grad.numpy()) < gtol, name
            p.grad = q.grad = None
        for (name, p), (_, q) in zip(net.layer8.named_buffers(), ref.layer8.named_buffers()):
            assert float((p.double() - q.double()).abs().max()) < tol, name
    net.eval()
    ref.eval()
    assert rel_l2(net._layer8_aten(x).detach().numpy(), ref.layer8(x).detach().numpy()) < tol


def test_backbone_switch(monkeypatch, tmp_path):
    """BH_FIELD_HEAD forces the choice; unset, a CUDA device's cached self-test verdict decides (parity AND speed)"""
    import tempfile

    import bihome_b200.functional as F
    from bihome_b200 import autotune
    monkeypatch.setenv('BH_FIELD_HEAD', 'fused')
    assert F.field_head_enabled()
    monkeypatch.setenv('BH_FIELD_HEAD', 'aten')
    assert not F.field_head_enabled(torch.device('cuda', 0))
    monkeypatch.delenv('BH_FIELD_HEAD', raising=False)
    assert not F.field_head_enabled() and not F.field_head_enabled(torch.device('cpu'))
    assert not F.field_head_supported(make_stage(torch.float32), torch.zeros(1, 16, 4, 4))      # CPU tensor: ATen modules
    monkeypatch.setenv('BH_CACHE_DIR', str(tmp_path))
    monkeypatch.setattr(torch.cuda, 'get_device_name', lambda i: 'emulated B200')
    calls = []
    for verdict, want in (({'ok': True, 'fused_ms': 1.0, 'aten_ms': 5.0}, True), ({'ok': True, 'fused_ms': 6.0, 'aten_ms': 5.0}, False),
                          ({'ok': False, 'err': 'train gx mismatch'}, False)):
        monkeypatch.setattr(autotune, '_choice', {})
        monkeypatch.setattr(autotune, '_probe_in_child', lambda index, v=verdict: calls.append(index) or v)
        for f in tmp_path.iterdir():
            f.unlink()
        assert F.field_head_enabled(torch.device('cuda', 1)) is want
        assert F.field_head_enabled(torch.device('cuda', 1)) is want          # per-process cache
        monkeypatch.setattr(autotune, '_choice', {})
        assert F.field_head_enabled(torch.device('cuda', 1)) is want          # on-disk cache: no second self-test ...
    assert calls == [1, 1, 1, 1]      # ... except after a self-test that did not finish ('err'): that is not a verdict to keep


def test_self_test_procedure_on_the_stand_in_kernels(monkeypatch):
    """the child's parity + timing procedure itself, run on CPU with the device entry points replaced by torch ops"""
    from bihome_b200 import autotune
    cpu_kernels.install(monkeypatch)
    verdict = autotune.compare_and_time(torch.device('cpu'), timing_batch=1)
    assert verdict['ok'] and verdict['worst'] < 0.5 and verdict['fused_ms'] > 0 and verdict['aten_ms'] > 0, verdict
    import bihome_b200.functional as F
    monkeypatch.setattr(F, '_fh_fwd', lambda *a: cpu_kernels.fh_fwd(*a) * 1.01)               # a wrong kernel is caught
    assert not autotune.compare_and_time(torch.device('cpu'), timing_batch=1)['ok']


@pytest.mark.parametrize('name', ['zeng-bihome-lr-1e-3', 'zeng-orig-lr-1e-3'])
def test_zeng_configs_with_the_fused_head_match_the_reference(monkeypatch, name):
    """the whole shipped Zeng models with layer8 routed through functional.field_head (stand-in device ops): loss and
    parameter gradients of the unmodified reference model for the same weights and batch (build container only)"""
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip('reference tree not present')
    import bihome_b200.functional as F
    import test_configs_end_to_end as E
    monkeypatch.setenv('BH_FIELD_HEAD', 'fused')
    real = F.field_head_supported
    calls = []

    class OnDevice:                              # the geometry check minus "is on a CUDA device"
        def __init__(self, t):
            self.is_cuda, self.dtype, self._t = True, t.dtype, t

        def dim(self):
            return self._t.dim()

    def supported(stage, x):
        calls.append(1)
        return real(stage, OnDevice(x))
    monkeypatch.setattr(F, 'field_head_supported', supported)
    monkeypatch.setattr(torch.Tensor, 'is_cuda', property(lambda self: True))
    path = os.path.join(ref_import.REFERENCE_ROOT, 'config', 'pds-coco', name + '.yaml')
    E.test_training_step_matches_reference(monkeypatch, path)
    assert calls


# ------------------------------------------------------------------------------------------------ host emulation
EMU_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emu')


def _build_emu(flags, out):
    import shutil
    import subprocess
    if shutil.which('g++') is None:
        pytest.skip('g++ not available')
    os.makedirs(os.path.join(EMU_DIR, '_build'), exist_ok=True)
    path = os.path.join(EMU_DIR, '_build', out)
    src = [os.path.join(EMU_DIR, 'fieldhead_emu.cpp'), os.path.join(EMU_DIR, 'cuda_emu.h'),
           os.path.join(os.path.dirname(EMU_DIR), '..', 'bihome_b200', 'csrc', 'fieldhead.cu')]
    if not os.path.isfile(path) or any(os.path.getmtime(f) > os.path.getmtime(path) for f in src):
        subprocess.run(['g++', '-std=c++17', '-O1', '-g', '-pthread', '-I', EMU_DIR] + flags + [src[0], '-o', path], check=True)
    return path


@pytest.fixture(scope='module')
def emu():
    import ctypes
    lib = ctypes.CDLL(_build_emu(['-fPIC', '-shared'], 'fieldhead_emu.so'))
    vp, i, ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong
    lib.emu_moments.argtypes = [vp, vp, ll, i]
    lib.emu_fwd.argtypes = [vp] * 6 + [i, i, i]
    lib.emu_bwd.argtypes = [vp] * 7 + [i, i, i]
    lib.emu_affine.argtypes = [vp] * 4 + [ll, i, i]
    return lib


@pytest.mark.parametrize('B,H,W,grid', [(1, 4, 8, 1), (3, 9, 7, 2), (2, 16, 24, 3), (5, 11, 13, 4)])
def test_kernels_on_the_host_emulator(emu, B, H, W, grid):
    """csrc/fieldhead.cu itself (not a restatement) compiled for the host with CTA threads as pthreads: the tile loops,
    the sample-straddling index arithmetic, partial tiles and the per-CTA partial layout against plain torch ops"""
    gen = torch.Generator().manual_seed(B * 100 + H)
    x = torch.relu(torch.randn(B, 16, H, W, generator=gen) + 0.3).contiguous(memory_format=torch.channels_last)
    W1, b1 = (torch.randn(128, 16, generator=gen) * 0.3).contiguous(), torch.randn(128, generator=gen)
    W2, b2 = (torch.randn(2, 128, generator=gen) * 0.2).contiguous(), torch.randn(2, generator=gen)
    g = torch.randn(B, 2, H, W, generator=gen)
    n, d = B * H * W, (lambda t: t.double())
    ptr = lambda t: t.data_ptr()
    mom = torch.full((grid, 16 + 256), float('nan'), dtype=torch.float64)
    emu.emu_moments(ptr(x), ptr(mom), n, grid)
    s = mom.sum(0)
    r1, r2 = cpu_kernels.fh_moments(x)
    assert rel_l2(s[:16].numpy(), r1.numpy()) < 1e-6 and rel_l2(s[16:].view(16, 16).numpy(), r2.numpy()) < 1e-6
    out = torch.full((B, 2, H, W), float('nan'))
    emu.emu_fwd(ptr(x), ptr(W1), ptr(b1), ptr(W2), ptr(b2), ptr(out), B, H * W, grid)
    assert rel_l2(out.numpy(), cpu_kernels.fh_fwd(d(x), d(W1), d(b1), d(W2), d(b2)).numpy()) < 1e-6
    gx = torch.full_like(x, float('nan'))
    parts = torch.full((grid, 128 * 16 + 3 * 128 + 2), float('nan'))
    emu.emu_bwd(ptr(x), ptr(W1), ptr(b1), ptr(W2), ptr(g), ptr(gx), ptr(parts), B, H * W, grid)
    ps = parts.sum(0)
    got = (gx, ps[:2048].view(128, 16), ps[2048:2176], ps[2176:2432].view(2, 128), ps[2432:])
    want = cpu_kernels.fh_bwd(d(x), d(W1), d(b1), d(W2), d(g))
    for a, b, name in zip(got, want, ('gx', 'gW1', 'gb1', 'gW2', 'gb2')):
        assert rel_l2(a.numpy(), b.numpy()) < 1e-5, name
    a, M = torch.randn(16, generator=gen), torch.randn(16, 16, generator=gen).contiguous()
    acc = gx.clone()
    emu.emu_affine(ptr(x), ptr(a), ptr(M), ptr(acc), n, 1, grid)
    assert rel_l2(acc.numpy(), cpu_kernels.fh_affine(d(x), d(a), d(M), d(gx).clone()).numpy()) < 1e-6
    fresh = torch.full_like(x, float('nan'))
    emu.emu_affine(ptr(x), ptr(a), ptr(M), ptr(fresh), n, 0, grid)
    assert rel_l2(fresh.numpy(), (acc - gx).numpy()) < 1e-5


def test_no_data_race_under_thread_sanitizer():
    """the same host build under -fsanitize=thread: a missing or misplaced __syncthreads() is a reported race on the
    shared arrays (checked by hand once: removing the barrier after phase 1 produces five reports)"""
    import subprocess
    exe = _build_emu(['-fsanitize=thread', '-DBH_EMU_MAIN'], 'fieldhead_tsan')
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    if 'FATAL: ThreadSanitizer' in r.stderr and 'data race' not in r.stderr:
        pytest.skip('ThreadSanitizer cannot run here: ' + r.stderr.strip().splitlines()[0])
    assert r.returncode == 0 and 'data race' not in r.stderr and r.stdout.startswith('ok'), r.stderr[-2000:]


def test_tensor_core_kernels_fragment_mapping_on_the_lane_emulator():
    """csrc/fieldhead_mma.cu takes its mma.sync fragments straight from 128-bit loads and from accumulator registers; the
    lane-by-lane restatement of that data movement (tests/emu/mma_lane_emulator.py, on top of the PTX ISA fragment layout of
    mma.m16n8k8) must reproduce the dense algebra: forward field, d/dx and the four weight gradients"""
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'emu'))
    import mma_lane_emulator as E
    rs = np.random.RandomState(0)
    HW, B = 64, 2
    n = B * HW
    x, W1, b1 = rs.randn(n, 16), rs.randn(128, 16) * 0.3, rs.randn(128) * 0.2
    W2, b2, g_out = rs.randn(2, 128) * 0.2, rs.randn(2), rs.randn(B, 2, HW)
    pre = x @ W1.T + b1
    h = np.maximum(pre, 0)
    out = (h @ W2.T + b2).reshape(B, HW, 2).transpose(0, 2, 1)
    G = g_out.transpose(0, 2, 1).reshape(n, 2)
    gh = (G @ W2) * (pre > 0)
    assert np.abs(E.fwd(x, W1, b1, W2, b2, HW) - out).max() < 1e-12
    assert np.abs(E.gx(x, W1, b1, W2, g_out, HW) - gh @ W1).max() < 1e-12
    gW1, gb1, gW2, gb2 = E.gw(x, W1, b1, W2, g_out, HW)
    assert np.abs(gW1 - gh.T @ x).max() < 1e-12 and np.abs(gb1 - gh.sum(0)).max() < 1e-12
    assert np.abs(gW2 - G.T @ h).max() < 1e-12 and np.abs(gb2 - G.sum(0)).max() < 1e-12
    S, M = E.moments(x)
    assert np.abs(S - x.sum(0)).max() < 1e-12 and np.abs(M - x.T @ x).max() < 1e-12
