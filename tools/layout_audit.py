"""Which modules of the training model emit tensors that are NOT channels-last?  (dev tool: layout mismatches make ATen
fall back to strided element-wise kernels and cuDNN to its NCHW BatchNorm)  Prints one line per offending module."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bihome_b200 import engine
from bihome_b200.data import gpu_pairs

CONFIG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'config', 'pds-coco', 'zeng-bihome-lr-1e-3.yaml')
cfg = engine.load_config(CONFIG)
torch.manual_seed(0)
model = engine.build_model(cfg, pretrained=False).cuda().to(memory_format=torch.channels_last)
model.train()
seen = {}


def hook(name):
    def f(mod, inp, out):
        for kind, ts in (('in', inp), ('out', (out,) if torch.is_tensor(out) else ())):
            for t in ts:
                if torch.is_tensor(t) and t.dim() == 4 and t.shape[1] > 1 and not t.is_contiguous(memory_format=torch.channels_last):
                    key = (name, kind)
                    if key not in seen:
                        seen[key] = (type(mod).__name__, tuple(t.shape), t.stride())
    return f


for n, m in model.named_modules():
    if not list(m.children()):
        m.register_forward_hook(hook(n))
pool = gpu_pairs.synthetic_pool(16)
loader = gpu_pairs.GpuPairLoader(pool, 32, 32 * 100, **gpu_pairs.transform_args(cfg['DATA']['TRANSFORMS']))
opt, sched = engine.build_optimizer(cfg, model)
engine.train_step(model, loader.next_batch(), opt, sched)
torch.cuda.synchronize()
for (name, kind), (typ, shape, stride) in seen.items():
    print('%-50s %-4s %-18s %s stride %s' % (name, kind, typ, shape, stride))
print('%d offending (module, direction) pairs' % len(seen))

from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as p:
    engine.train_step(model, loader.next_batch(), opt, sched)
    torch.cuda.synchronize()
rows = [e for e in p.key_averages(group_by_input_shape=True) if e.key in ('aten::add', 'aten::add_', 'aten::copy_', 'aten::contiguous', 'aten::cudnn_batch_norm', 'aten::native_batch_norm', 'aten::max_pool2d_with_indices', 'aten::cat')]
rows.sort(key=lambda e: -e.device_time_total)
for e in rows[:25]:
    print('%-34s n=%3d cuda %8.1f us  shapes %s' % (e.key, e.count, e.device_time_total, str(e.input_shapes)[:120]))
