"""One launch of the image-warp kernels (tile path, both builds) and of the pair generator: the command ncu profiles.
usage: python tools/once_warp.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch

import bihome_b200.functional as F
import microbench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
once = lambda fn, warm=0: (fn(), torch.cuda.synchronize(), 1.0)[-1]
for variant in (0, 1):
    F.tune('warp_variant', variant)
    microbench.bench_image_warp(B, 128, once, 0.02, ' [variant %d]' % variant)
F.tune('warp_variant', 0)
pool = torch.randint(0, 256, (64, 240, 320, 3), dtype=torch.uint8, device='cuda')
params, index = F.pairgen_draw(256, 64, (240, 320), 32, 128, 32.0, 1, 0, 'cuda')
F.pairgen_apply(pool, index, params, 128)
torch.cuda.synchronize()
