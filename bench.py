#!/usr/bin/env python
"""bench.py -- training throughput of the biHomE hot path on N B200s (BASELINE.json metric / configs[1]).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun, one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    (the reference's CPU path, oracle port, host cores)

A step = one training iteration of pds-coco/zeng-bihome at 128x128, B=256 per GPU: GPU pair generation (K5) ->
Zeng backbone (cuDNN) -> N-point DLT (K4) -> DLT-4 (K1) -> warp + pooled masks (K2) -> frozen ResNet-34 stem x4 (cuDNN)
-> fused biHomE loss fwd+bwd (K3) -> backward (K2b, K1/K4 adjoints, cuDNN) -> Adam.  Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

CONFIG = os.path.join(ROOT, 'config', 'pds-coco', 'zeng-bihome-lr-1e-3.yaml')
METRIC = 'train_image_pairs_per_sec'
UNIT = 'image-pairs/s'
# algorithmic bytes per image pair of the fused loss kernel (SURVEY.md 8d / DESIGN.md 4): 4 feature reads + 2 gradient
# writes of C*h*w fp32 + 2 pooled masks in + 2 mask gradients out, C=64, h=w=32
LOSS_BYTES_PER_PAIR = (4 + 2) * 64 * 32 * 32 * 4 + 4 * 32 * 32 * 4


def kernel_traffic(entry_point):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of an entry point at B=256, from the committed ncu capture
    (profiles/kernel_traffic.json: {entry point: bytes}, written by tools/summarize_profiles.py; the round-1 file
    profiles/loss_traffic.json for the loss); None when the capture is absent"""
    try:
        with open(os.path.join(ROOT, 'profiles', 'kernel_traffic.json')) as f:
            return float(json.load(f)[entry_point])
    except Exception:  # noqa: BLE001
        pass
    if entry_point == 'bh_bihome_fwd_bwd':
        try:
            with open(os.path.join(ROOT, 'profiles', 'loss_traffic.json')) as f:
                return float(json.load(f)['dram_bytes_per_launch'])
        except Exception:  # noqa: BLE001
            pass
    return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured'
    except Exception:  # noqa: BLE001
        return 6650.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)"""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(sm)}


def workload_config(batch, world, pool):
    """the `config` object of the JSON line -- the same for both arms (the reference arm runs the same workload on the
    host cores; what differs is said in its cpu_baseline.sample)"""
    return {'workload': 'pds-coco zeng-bihome-lr-1e-3 training step (Zeng ResNet34 perspective-field backbone + biHomE '
                        'loss, ResNet-34 stem extractor), 128x128 patches, B=%d per GPU, fp32 (cuDNN TF32 convs = torch default), '
                        'random-init weights, synthetic uint8 image pool (%d x 240x320) resident in HBM' % (batch, pool),
            'global_batch': batch * world, 'parallelism': 'dp%d' % world,
            'l2': 'per-step working set (activations of B=256) >> 126 MB L2: no flush needed'}


def cpu_oracle_run(steps, warmup, batch, threads, budget_s=None):
    """the reference's CPU path for the same config: torch backbone on CPU + oracle head (oracle/ref_train.py).
    Returns (image-pairs/s, seconds per step, image pairs per step).  With a time budget the first (untimed) step is also a
    probe: when `steps + warmup` steps of `batch` pairs would not fit, the steps become a smaller sample of the same workload
    (a quarter of the pairs, down to 16) -- the driver runs this arm with the GPU arm's step count."""
    from bihome_b200 import engine
    from bihome_b200.backbones import Rethinking
    from oracle.ref_train import OracleModel
    torch.set_num_threads(threads)
    cfg = engine.load_config(CONFIG)
    bcfg = dict(cfg['MODEL']['BACKBONE'])
    bcfg['PRETRAINED_RESNET'] = False
    torch.manual_seed(0)
    model = OracleModel(Rethinking.Model(**bcfg), cfg['MODEL']['HEAD'])
    model.backbone.skip_cancelled_bias = False      # the reference's op sequence, literally
    model.train()
    opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=cfg['SOLVER']['LR'])

    def inputs(n):
        g = torch.Generator().manual_seed(1)
        lo = torch.rand(n, 1, 18, 18, generator=g)
        a = torch.nn.functional.interpolate(lo, size=(128, 128), mode='bicubic', align_corners=True)
        return a, torch.roll(a, shifts=(3, -2), dims=(2, 3)) + 0.05 * torch.randn(n, 1, 128, 128, generator=g)
    p1, p2 = inputs(batch)

    def step():
        opt.zero_grad()
        loss, _, _ = model({'patch_1': p1, 'patch_2': p2})
        loss.backward()
        opt.step()
        return float(loss.detach())
    done = 0
    if budget_s is not None:
        while True:
            t0 = time.perf_counter()
            step()
            probe = time.perf_counter() - t0
            done = 1
            if probe * (steps + max(warmup - 1, 0)) <= budget_s or batch <= 16:
                break
            batch = max(batch // 4, 16)
            p1, p2 = inputs(batch)
            done = 0
    for _ in range(max(warmup - done, 0)):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, batch


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # the same B = 256 step as the GPU arm (BatchNorm statistics and the per-step overheads are those of the real workload);
    # a B = 256 step of the Zeng backbone keeps ~35 GB of activations on the host, so a box with less RAM runs B = 64 steps;
    # a box whose cores would need more than --ref-budget seconds for steps + warmup such steps runs smaller ones
    batch = args.ref_batch
    if batch is None:
        try:
            import psutil
            avail = psutil.virtual_memory().available
        except Exception:  # noqa: BLE001
            avail = 0
        batch = args.batch if avail >= 56e9 else 64
    value, spp, used = cpu_oracle_run(args.steps, args.warmup, batch, threads, budget_s=args.ref_budget if args.ref_batch is None else None)
    sample = ('%d steps of B=%d after %d warm-up (%.1f s/step): torch CPU backbone (plain ATen modules, NCHW) + the oracle\'s '
              'kornia-0.5.0 head and its autograd, Adam, %d host threads' % (args.steps, used, args.warmup, spp, threads))
    if used != args.batch:
        sample += '; bounded sample: B=%d steps of the B=%d workload (%s)' % (used, args.batch, 'host RAM' if used == batch else
                                                                              'time budget of %d s for the run' % args.ref_budget)
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': spp * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args.batch, max(args.gpus, 1), args.pool),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def run_ours(args):
    import torch.distributed as dist
    from bihome_b200 import autotune, cabi, engine
    from bihome_b200 import functional as F
    from bihome_b200.data import gpu_pairs

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    cabi.lib()  # fail loudly if the extension is missing
    torch.backends.cudnn.benchmark = bool(args.cudnn_benchmark)
    cfg = engine.load_config(CONFIG)
    B = args.batch
    torch.manual_seed(0)
    model = engine.build_model(cfg, pretrained=False).to(dev)
    if args.channels_last:
        model = model.to(memory_format=torch.channels_last)
    model.train()
    use_graph = not args.eager
    net = model
    if world > 1 and not use_graph:
        net = engine.data_parallel(model, local)
    opt, sched = engine.build_optimizer(cfg, model)
    t = gpu_pairs.transform_args(cfg['DATA']['TRANSFORMS'])
    pool = gpu_pairs.synthetic_pool(args.pool, device=dev)
    loader = gpu_pairs.GpuPairLoader(pool, B, B * (3 * args.steps + args.warmup + 16) * 4, seed=cfg['DATA']['SAMPLER']['TRAIN_SEED'],
                                     rank=rank, **t)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        x = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(x, op=dist.ReduceOp.MAX)
        return float(x.item())

    # The step, through the repo's public API.  Default: forward + backward replayed from ONE CUDA graph per rank
    # (engine.GraphedStep = `train.py --cuda_graph`), K5 pair generation, the data-parallel gradient mean (one NCCL all-reduce of
    # the flat gradient buffer) and Adam eager.  --eager: engine.train_step, DDP's bucketed all-reduce at N > 1.
    graphed = engine.GraphedStep(model, loader.next_batch()) if use_graph else None

    def one_step(batch):
        if graphed is not None:
            return engine.graphed_train_step(graphed, batch, opt, sched)
        return engine.train_step(net, batch, opt, sched)

    # ---------------- device-resident throughput: `value` ----------------
    for _ in range(args.warmup):
        one_step(loader.next_batch())
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    launches0 = cabi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss, _, _ = one_step(loader.next_batch())
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    # kernels of this library launched inside the timed region: the eager ones (K5) are counted as they launch, the ones inside
    # the graph were counted once, at capture
    launches = cabi.launch_count() - launches0 + (graphed.launches_per_replay * args.steps if graphed is not None else 0)
    clk = clocks.stop() if rank == 0 else None
    value = B * world * args.steps / (ms * 1e-3)
    final_loss = float(loss.item())

    # ---------------- end to end through the public API with HOST buffers: `e2e` ----------------
    host = []
    for _ in range(0 if args.no_e2e else 4):
        b = loader.next_batch()
        host.append({k: v.cpu().pin_memory() for k, v in b.items()})
    stage = {k: torch.empty_like(v, device=dev) for k, v in host[0].items()} if host and graphed is None else {}
    h2d = sum(v.numel() * v.element_size() for v in host[0].values()) if host else 0

    def e2e_step(i):
        hb = host[i % len(host)]
        if graphed is not None:
            l, _, _ = one_step(hb)          # GraphedStep copies the pinned host batch straight into its static device buffers
        else:
            for k in stage:
                stage[k].copy_(hb[k], non_blocking=True)
            l, _, _ = one_step(dict(stage))
        return float(l.item())          # device -> host read of the step's loss
    ms_e2e, e2e_value = None, None
    if host:
        for i in range(3):
            e2e_step(i)
        barrier()
        e0.record()
        for i in range(args.steps):
            e2e_step(i)
        e1.record()
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1))
        e2e_value = B * world * args.steps / (ms_e2e * 1e-3)

    # ---------------- per-kernel in-situ durations: the same K steps, eager, a CUDA-event pair around every C-ABI call ----------
    # (on the launching stream).  Kept out of the regions `value` / `e2e` are timed over: ~250 bracketed calls per step slow the
    # host's launch rate (ms_per_step_instrumented says by how much), and events cannot be recorded inside a graph replay.
    def eager_step(batch):
        opt.zero_grad(set_to_none=graphed is None)       # a GraphedStep owns the gradient buffers
        l, _, _ = net(batch)
        l.backward()
        opt.step()
    F.enable_timing(True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        eager_step(loader.next_batch())
    e1.record()
    barrier()
    ms_instr = e0.elapsed_time(e1)
    kt = F.timings()
    kbytes = F.timing_bytes()
    F.enable_timing(False)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- rooflines of the custom kernels, timed in situ (CUDA events on the launching stream) ----------------
    peak, peak_kind = measured_peak()
    kernels = {k: {'launches': len(v), 'avg_ms': sum(v) / len(v), 'ms_per_step': sum(v) / args.steps} for k, v in sorted(kt.items())}
    # algorithmic bytes per image pair and launch (SURVEY.md 8d / DESIGN.md section 4); K6 works per pixel of one backbone
    # pass (16 384 pixels per pair): 64 B in, 8 B field, 64 B input gradient
    px = 128 * 128
    stem_t = B * 64 * 64 * 64 * 4.0
    ALG = {'bh_warp_fwd': 270336, 'bh_warp_bwd': 270408, 'bh_bihome_fwd_bwd': LOSS_BYTES_PER_PAIR,
           'bh_pairgen_apply': 2 * 3 * (128 + 64) ** 2 + 2 * 4 * 128 * 128,
           'bh_fieldhead_moments': 64 * px, 'bh_fieldhead_fwd': 72 * px, 'bh_fieldhead_bwd': 136 * px, 'bh_fieldhead_affine': 192 * px,
           # K7 per pass over the [B,64,64,64] stem tensor (T bytes, pooled quarter T/4, 1-byte codes T/16): every tensor once, plus
           # the part of the second pass' inputs that cannot have stayed in the 126 MB L2 (the batch statistics force two passes)
           'bh_stem_fwd': (stem_t + stem_t / 4 + max(0.0, stem_t - 126e6)) / B,
           'bh_stem_bwd': (2 * stem_t + stem_t / 4 + stem_t / 16 + max(0.0, stem_t + stem_t / 4 + stem_t / 16 - 126e6)) / B}
    NAMES = {'bh_bihome_fwd_bwd': 'bihome_stream_kernel<false,1> + bihome_finish_kernel (bh_bihome_fwd_bwd, channels-last C=64, B<512)',
             'bh_fieldhead_bwd': 'fieldhead_gx_mma_kernel + fieldhead_gw_mma_kernel (bh_fieldhead_bwd: mma.sync TF32 tensor-core kernels, '
                                 'bound by the mma.sync issue rate and the ReLU/projection epilogue, not by HBM)',
             'bh_fieldhead_fwd': 'fieldhead_fwd_mma_kernel (bh_fieldhead_fwd, mma.sync TF32)',
             'bh_warp_fwd': 'warp_fwd_tile_kernel (bh_warp_fwd: TMA box per 32x32 tile)', 'bh_warp_bwd': 'warp_bwd_tile_kernel + finish (bh_warp_bwd)',
             'bh_pairgen_apply': 'pairgen_apply_kernel (bh_pairgen_apply: instruction bound, cv2-exact colour math)',
             'bh_stem_fwd': 'bn_stats_kernel + bn_finalize_kernel + stem_pool_fwd_kernel (bh_stem_fwd: BatchNorm -> ReLU -> MaxPool, channels-last)',
             'bh_stem_bwd': 'stem_bwd_reduce_kernel + stem_bwd_finalize_kernel + stem_bwd_apply_kernel (bh_stem_bwd)',
             'bh_bnact_fwd': 'bn_stats_kernel + bn_finalize_kernel + bnact_fwd_kernel (bh_bnact_fwd: BatchNorm [+ residual] -> ReLU of the residual blocks, all shapes of the step)',
             'bh_bnact_bwd': 'bnact_bwd_reduce_kernel + stem_bwd_finalize_kernel + bnact_bwd_apply_kernel (bh_bnact_bwd, all shapes of the step)',
             'bh_bnact2_fwd': '2 x (bn_stats_kernel + bn_finalize_kernel) + bnact2_fwd_kernel (bh_bnact2_fwd: relu(bn(a) + bn(b)), all shapes of the step)',
             'bh_bnact2_bwd': 'bnact2_bwd_reduce_kernel + 2 x stem_bwd_finalize_kernel + bnact2_bwd_apply_kernel (bh_bnact2_bwd, all shapes of the step)',
             'bh_bias_add': 'bias_add_kernel (bh_bias_add: in-place bias of the transposed convolutions, all shapes of the step)',
             'bh_bias_grad': 'bn_stats_kernel + bias_grad_finalize_kernel (bh_bias_grad, all shapes of the step)'}
    for name, per_pair in ALG.items():
        if name in kernels and kernels[name]['avg_ms'] > 0:
            gbs = per_pair * B / (kernels[name]['avg_ms'] * 1e-3) / 1e9
            kernels[name].update({'algorithmic_GBps': gbs, 'frac_of_hbm_peak': gbs / peak})
    # entry points whose tensors change from call to call (K7b / K7c): bytes summed over the timed calls -- every input and
    # output tensor once, plus the second read of what the second pass needs again when that does not fit in the 126 MB L2
    # (functional._bn_bytes; the batch statistics force two passes)
    for name, total in kbytes.items():
        if name in kernels and name not in ALG and kernels[name]['avg_ms'] > 0:
            gbs = total / (kernels[name]['avg_ms'] * kernels[name]['launches'] * 1e-3) / 1e9
            kernels[name].update({'algorithmic_GBps': gbs, 'frac_of_hbm_peak': gbs / peak, 'algorithmic_bytes_per_step': total / args.steps})
            ALG[name] = total / kernels[name]['launches'] / B
    # the block the contract asks for: the custom kernel that takes the most time per step
    dominant = max((k for k in kernels if k in ALG), key=lambda k: kernels[k]['ms_per_step'], default=None)
    roofline = None
    if dominant is not None:
        d = kernels[dominant]
        roofline = {'bound': 'hbm', 'kernel': NAMES.get(dominant, dominant), 'entry_point': dominant, 'achieved': d['algorithmic_GBps'],
                    'peak': peak, 'peak_source': peak_kind, 'unit': 'GB/s', 'frac': d['frac_of_hbm_peak'],
                    'traffic': kernel_traffic(dominant) if args.loss_traffic is None or dominant != 'bh_bihome_fwd_bwd' else args.loss_traffic,
                    'launch_ms': d['avg_ms'], 'launches_per_step': d['launches'] / args.steps,
                    'algorithmic_bytes_per_launch': ALG[dominant] * B}

    # BASELINE.json's second metric, "warp+loss HBM GB/s": the three bandwidth kernels of the path together
    wl = [(n, kernels[n]['avg_ms'] * kernels[n]['launches'] / args.steps) for n in ('bh_warp_fwd', 'bh_warp_bwd', 'bh_bihome_fwd_bwd') if n in kernels]
    wl_ms = sum(ms_ for _, ms_ in wl)
    wl_bytes = (270336 + 270408 + LOSS_BYTES_PER_PAIR) * B
    warp_loss = {'kernels': [n for n, _ in wl], 'ms_per_step': wl_ms, 'algorithmic_bytes_per_step': wl_bytes,
                 'achieved': wl_bytes / (wl_ms * 1e-3) / 1e9 if wl_ms > 0 else None, 'peak': peak, 'unit': 'GB/s',
                 'frac': (wl_bytes / (wl_ms * 1e-3) / 1e9 / peak) if wl_ms > 0 else None}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, spp, _ = cpu_oracle_run(steps=2, warmup=1, batch=64, threads=threads)
        cpu_baseline = {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                        'sample': '2 steps of B=64 after 1 warm-up (%.1f s/step; bounded sample of the B=256 workload, which '
                                  '`--impl reference` runs in full): torch CPU backbone + oracle kornia-0.5.0 head' % spp}

    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'ms_per_step_instrumented': ms_instr / args.steps, 'cuda_graph': graphed is not None,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            'config': workload_config(B, world, args.pool),
            'layout': {'channels_last': bool(args.channels_last), 'cudnn_benchmark': bool(args.cudnn_benchmark), 'field_head': 'fused (K6)' if F.field_head_enabled(dev) else 'aten',
                       'stem': 'fused (K7)' if 'bh_stem_fwd' in kernels else 'aten',
                       'bn_relu': 'fused (K7b)' if 'bh_bnact_fwd' in kernels else 'aten',
                       'bn_bn_relu': 'fused (K7c)' if 'bh_bnact2_fwd' in kernels else 'aten',
                       'conv_transpose_bias': 'fused (K8)' if 'bh_bias_add' in kernels else 'aten'},
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4,
                    'ms_per_step': (ms_e2e / args.steps) if ms_e2e else None},
            'gpu_launches': launches, 'roofline': roofline, 'warp_loss_roofline': warp_loss, 'cpu_baseline': cpu_baseline, 'clocks': clk,
            'kernels': kernels, 'final_loss': final_loss, 'field_head_self_test': autotune.last_verdict(dev)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    # ONE JSON line on stdout: libraries (NCCL prints its version banner there) get stderr for the rest of the run
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, 'w', buffering=1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=256, help='image pairs per GPU per step')
    ap.add_argument('--pool', type=int, default=256, help='synthetic uint8 images resident on each GPU')
    ap.add_argument('--nchw', dest='channels_last', action='store_false',
                    help='keep the conv stack in NCHW (default: channels-last, 2.5x faster cuDNN path on B200)')
    ap.add_argument('--field-head', default=None, choices=['aten', 'fused'],
                    help="Zeng backbone's last stage: 'aten' = the four torch modules (default), 'fused' = K6 (csrc/fieldhead.cu); "
                         "unset = BH_FIELD_HEAD, else the device's self-test decides (bihome_b200/autotune.py)")
    ap.add_argument('--ref-budget', type=int, default=600, help='--impl reference: seconds the whole run (steps + warmup) may take before the steps '
                    'become smaller samples of the workload')
    ap.add_argument('--ref-batch', type=int, default=None, help='--impl reference: image pairs per CPU step (default: --batch, 64 on a box with < 56 GB of free RAM)')
    ap.add_argument('--no-cudnn-benchmark', dest='cudnn_benchmark', action='store_false', default=True,
                    help='leave torch.backends.cudnn.benchmark off (default on, as train.py: the shapes of a training run are '
                         'fixed, cuDNN times its convolution algorithms once per shape during the warm-up: 72.2 -> 69.6 ms/step)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--eager', action='store_true',
                    help='engine.train_step (DDP at N > 1) instead of the default: forward + backward replayed from one CUDA graph per '
                         'rank (engine.GraphedStep, train.py --cuda_graph)')
    ap.add_argument('--no-e2e', action='store_true', help='skip the host-buffer phase (profiling runs)')
    ap.add_argument('--loss-traffic', type=float, default=None, help='dram bytes per launch of the loss kernel from ncu (profiles/)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    if args.field_head is not None:
        os.environ['BH_FIELD_HEAD'] = args.field_head
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
