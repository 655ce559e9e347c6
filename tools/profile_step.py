"""Where does a training step spend its GPU time?  (torch.profiler kernel table; dev tool, not part of the bench)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bihome_b200 import engine
from bihome_b200.data import gpu_pairs

CONFIG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'config', 'pds-coco', 'zeng-bihome-lr-1e-3.yaml')


def run(channels_last, benchmark, B=256, steps=5, profile=False):
    torch.backends.cudnn.benchmark = benchmark
    cfg = engine.load_config(CONFIG)
    torch.manual_seed(0)
    model = engine.build_model(cfg, pretrained=False).cuda()
    if channels_last:
        model = model.to(memory_format=torch.channels_last)
    model.train()
    opt, sched = engine.build_optimizer(cfg, model)
    pool = gpu_pairs.synthetic_pool(64)
    loader = gpu_pairs.GpuPairLoader(pool, B, B * 1000, **gpu_pairs.transform_args(cfg['DATA']['TRANSFORMS']))
    for _ in range(3):
        engine.train_step(model, loader.next_batch(), opt, sched)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        engine.train_step(model, loader.next_batch(), opt, sched)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    print('channels_last=%s benchmark=%s: %.1f ms/step, %.0f pairs/s, mem %.1f GB' % (channels_last, benchmark, dt * 1e3, B / dt,
          torch.cuda.max_memory_allocated() / 1e9), flush=True)
    if profile:
        from torch.profiler import profile as prof, ProfilerActivity
        with prof(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as p:
            for _ in range(2):
                engine.train_step(model, loader.next_batch(), opt, sched)
            torch.cuda.synchronize()
        print(p.key_averages().table(sort_by='cuda_time_total', row_limit=45, max_name_column_width=90))


if __name__ == '__main__':
    run(False, False)
    run(False, True)
    run(True, True, profile=True)
