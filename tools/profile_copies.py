"""Which ATen ops of the training step still move whole tensors around the custom kernels?  (dev tool, not part of the bench)

One eager step of bench.py's workload under torch.profiler with input shapes; prints the copy / layout / add /
reduction ops (everything that is neither a convolution nor one of this library's entry points) grouped by input shapes and
the chain of ops that enclose them (the convolution or autograd node a copy belongs to), sorted by device time.  usage: python tools/profile_copies.py [B] > gpurun_out/copies.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from bihome_b200 import engine
from bihome_b200.data import gpu_pairs

CONFIG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'config', 'pds-coco', 'zeng-bihome-lr-1e-3.yaml')
WATCH = ('aten::copy_', 'aten::clone', 'aten::contiguous', 'aten::add', 'aten::add_', 'aten::sum', 'aten::mul', 'aten::cat',
         'aten::fill_', 'aten::zero_', 'aten::_to_copy', 'aten::threshold_backward', 'aten::relu', 'aten::cudnn_batch_norm',
         'aten::cudnn_batch_norm_backward', 'aten::native_batch_norm', 'aten::native_batch_norm_backward')


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    torch.backends.cudnn.benchmark = True
    cfg = engine.load_config(CONFIG)
    torch.manual_seed(0)
    model = engine.build_model(cfg, pretrained=False).cuda().to(memory_format=torch.channels_last).train()
    opt, sched = engine.build_optimizer(cfg, model)
    loader = gpu_pairs.GpuPairLoader(gpu_pairs.synthetic_pool(64), B, B * 1000, **gpu_pairs.transform_args(cfg['DATA']['TRANSFORMS']))
    for _ in range(3):
        engine.train_step(model, loader.next_batch(), opt, sched)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as p:
        engine.train_step(model, loader.next_batch(), opt, sched)
        torch.cuda.synchronize()
    # group by (op, input shapes, chain of enclosing ops): the chain names the convolution / autograd node a copy belongs to
    groups = {}
    for e in p.events():
        if e.name not in WATCH:
            continue
        dev_us = getattr(e, 'self_device_time_total', None)
        if dev_us is None:
            dev_us = getattr(e, 'self_cuda_time_total', 0)
        if dev_us <= 0:
            continue
        chain, q = [], e.cpu_parent
        while q is not None and len(chain) < 6:
            shapes = [s_ for s_ in (q.input_shapes or []) if s_]
            chain.append(q.name.replace('autograd::engine::evaluate_function: ', 'node ') + (' ' + str(shapes[:3]) if shapes and len(chain) < 5 else ''))
            q = q.cpu_parent
        key = (e.name, str(e.input_shapes)[:110], ' <- '.join(chain)[:400])
        g = groups.setdefault(key, [0.0, 0])
        g[0] += dev_us
        g[1] += 1
    rows = sorted(((us, n) + key for key, (us, n) in groups.items() if us >= 20), reverse=True)
    total = sum(r[0] for r in rows)
    print('# one eager step at B = %d: %d groups of watched ATen ops, %.2f ms of device time' % (B, len(rows), total / 1e3))
    for us, n, key, shapes, site in rows[:60]:
        print('%9.1f us  n=%-3d %-34s %s\n              %s' % (us, n, key, shapes, site))


if __name__ == '__main__':
    main()
