"""Which implementation of the Zeng backbone's last stage runs on this device: K6 (csrc/fieldhead.cu) or the ATen modules?

``BH_FIELD_HEAD=fused`` / ``=aten`` force the answer.  Unset (``auto``), the answer comes from a self-test in a CHILD
process, once per (library build, device type) and cached under the temp directory: the child runs both
implementations on the device, checks K6's output, input gradient, parameter gradients and running statistics against
the four ATen modules in training and in eval mode, and times one forward + backward of each at a production-sized
input.  K6 is chosen only if every check passes and it is faster.  Both candidates are GPU paths -- there is no CPU
fallback anywhere -- and a crash of the child (a CUDA error is sticky for its process) cannot take the caller down:
the caller then stays on the ATen modules and says so on stderr.

The same mechanism as cuDNN's benchmark mode.  On the B200s measured so far the verdict is "fused": K6 (tensor-core kernels,
csrc/fieldhead_mma.cu) takes 1.95 ms for the stage at 256 x 128 x 128 pixels against 9.7 ms through the ATen modules
(profiles/r02u_microbench_fieldhead.jsonl); the self-test stays as the guard for other devices and library builds.
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

_choice = {}
_verdict = {}


def _cache_path(device_name):
    from . import cabi
    import torch
    try:
        stamp = '%d' % os.path.getmtime(cabi.LIB_PATH)
    except OSError:
        stamp = 'nolib'
    key = hashlib.sha1(('%s|%s|%s' % (stamp, device_name, torch.__version__)).encode()).hexdigest()[:16]
    # a per-user directory (0700), not the shared temp directory: the verdict is trusted when it is read back
    root = os.environ.get('BH_CACHE_DIR') or os.path.join(os.path.expanduser('~'), '.cache', 'bihome_b200')
    try:
        os.makedirs(root, mode=0o700, exist_ok=True)
    except OSError:
        root = tempfile.gettempdir()
    return os.path.join(root, 'fieldhead_%s.json' % key)


def field_head_choice(device):
    """'fused' or 'aten' for a CUDA torch.device (cached per process and on disk)"""
    forced = os.environ.get('BH_FIELD_HEAD', 'auto')
    if forced in ('fused', 'aten'):
        return forced
    import torch
    index = device.index if device.index is not None else torch.cuda.current_device()
    if index in _choice:
        return _choice[index]
    name = torch.cuda.get_device_name(index)
    path = _cache_path(name)
    verdict = None
    try:
        with open(path) as f:
            verdict = json.load(f)
    except (OSError, ValueError):
        verdict = _probe_in_child(index)
        if 'err' not in verdict:        # a child that crashed or timed out (transient: memory, a busy device) is not a verdict to keep
            try:
                fd, tmp = tempfile.mkstemp(dir=os.path.dirname(path))
                with os.fdopen(fd, 'w') as f:
                    json.dump(verdict, f)
                os.replace(tmp, path)
            except OSError:
                pass
        if os.environ.get('RANK', '0') == '0':
            print('bihome_b200: field-head self-test on %s: %s' % (name, json.dumps(verdict)), file=sys.stderr)
    _verdict[index] = verdict
    _choice[index] = 'fused' if verdict.get('ok') and verdict.get('fused_ms', 1e9) < verdict.get('aten_ms', 0) else 'aten'
    return _choice[index]


def last_verdict(device):
    """the self-test result behind field_head_choice(device), or None when the choice was forced / never needed"""
    import torch
    index = device.index if device.index is not None else torch.cuda.current_device()
    return _verdict.get(index)


def _probe_in_child(index):
    env = dict(os.environ)
    env['BH_FIELD_HEAD'] = 'aten'           # the child compares explicitly; it must not recurse into the self-test
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env['PYTHONPATH'] = root + os.pathsep + env.get('PYTHONPATH', '')
    for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK', 'MASTER_ADDR', 'MASTER_PORT', 'TORCHELASTIC_RUN_ID'):
        env.pop(k, None)
    try:
        r = subprocess.run([sys.executable, '-m', 'bihome_b200.autotune', str(index)], env=env, capture_output=True, text=True,
                           timeout=300)
    except (OSError, subprocess.TimeoutExpired) as e:
        return {'ok': False, 'err': 'self-test did not finish: %r' % (e,)}
    for line in reversed(r.stdout.strip().splitlines()):
        if line.startswith('{'):
            try:
                return json.loads(line)
            except ValueError:
                break
    return {'ok': False, 'err': 'self-test exited %d: %s' % (r.returncode, r.stderr.strip()[-300:])}


# ------------------------------------------------------------------------------------------------
# the child
# ------------------------------------------------------------------------------------------------
def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _stage(device, seed=0):
    import torch
    nn = torch.nn
    torch.manual_seed(seed)
    stage = nn.Sequential(nn.Conv2d(16, 128, 1), nn.BatchNorm2d(128), nn.ReLU(), nn.Conv2d(128, 2, 1))
    with torch.no_grad():
        stage[1].weight.uniform_(0.5, 1.5)
        stage[1].bias.normal_()
    return stage.to(device).to(memory_format=torch.channels_last)


def self_test(index):
    """runs in the child: {'ok', 'fused_ms', 'aten_ms', 'worst'} (or 'err')"""
    import torch
    from . import functional as F
    dev = torch.device('cuda', index)
    torch.cuda.set_device(dev)
    if not F.field_head_supported(_stage(dev), torch.zeros(1, 16, 4, 4, device=dev)):
        return {'ok': False, 'err': 'geometry not supported by the library'}
    return compare_and_time(dev)


def compare_and_time(dev, timing_batch=64):
    """K6 against the four ATen modules on `dev`: parity in training and eval mode, then ms per forward + backward"""
    import copy
    import time

    import torch
    from . import functional as F
    cuda = dev.type == 'cuda'
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    fused = _stage(dev)
    aten = copy.deepcopy(fused)
    worst = 0.0
    gen = torch.Generator().manual_seed(1)
    # float32 on both sides: a ReLU unit within round-off of zero may switch between the two evaluations and move the
    # input gradient of one pixel by ~10 %, and a library may still pick a TF32 kernel for the ATen side's 1x1
    # convolutions (1e-3) -- the bounds leave room for both; a wrong kernel is off by O(0.1 - 1)
    bounds = {'out': 5e-3, 'gx': 5e-2, 'param': 3e-2, 'buffer': 5e-3}
    for mode, shape in (('train', (3, 16, 40, 56)), ('train', (2, 16, 33, 17)), ('eval', (2, 16, 32, 32))):
        getattr(fused, mode)()
        getattr(aten, mode)()
        x = torch.relu(torch.randn(*shape, generator=gen) + 0.3).to(dev).contiguous(memory_format=torch.channels_last)
        g = torch.randn(shape[0], 2, shape[2], shape[3], generator=gen).to(dev)
        xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        oa, ob = F.field_head(fused, xa), aten(xb)
        (oa * g).sum().backward()
        (ob * g).sum().backward()
        errs = [('out', _rel(oa.detach(), ob.detach())), ('gx', _rel(xa.grad, xb.grad))]
        scale = max(float(q.grad.abs().max()) for q in aten.parameters())
        for p, q in zip(fused.parameters(), aten.parameters()):
            errs.append(('param', float((p.grad - q.grad).abs().max()) / max(float(q.grad.abs().max()), 1e-3 * scale)))
            p.grad = q.grad = None
        for p, q in zip(fused.buffers(), aten.buffers()):
            errs.append(('buffer', float((p.double() - q.double()).abs().max()) / max(float(q.double().abs().max()), 1.0)))
        for kind, e in errs:
            if not (e == e) or e > bounds[kind]:
                return {'ok': False, 'err': '%s %s mismatch %.3e' % (mode, kind, e)}
            worst = max(worst, e / bounds[kind])
    # timing at a production-sized input: 64 x 128 x 128 pixels (a quarter of the north-star batch)
    fused.train()
    aten.train()
    x = torch.relu(torch.randn(timing_batch, 16, 128, 128, generator=gen) + 0.3).to(dev).contiguous(memory_format=torch.channels_last)
    g = torch.randn(timing_batch, 2, 128, 128, generator=gen).to(dev)

    def once(fn):
        xa = x.clone().requires_grad_(True)
        (fn(xa) * g).sum().backward()

    def time_ms(fn):
        for _ in range(3):
            once(fn)
        if not cuda:
            t0 = time.perf_counter()
            once(fn)
            return (time.perf_counter() - t0) * 1e3
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            once(fn)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 5
    torch.backends.cudnn.allow_tf32 = True          # the training default: time what the step would really run
    return {'ok': True, 'fused_ms': time_ms(lambda t: F.field_head(fused, t)), 'aten_ms': time_ms(aten), 'worst': worst}


if __name__ == '__main__':
    try:
        out = self_test(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
    except Exception as e:  # noqa: BLE001 -- the parent only needs a verdict
        out = {'ok': False, 'err': '%s: %s' % (type(e).__name__, str(e)[:300])}
    print(json.dumps(out))
