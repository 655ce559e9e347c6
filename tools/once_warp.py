"""One launch of the image warp (forward + backward) at two batch sizes -- the command ncu profiles."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import microbench as mb

sizes = [int(v) for v in sys.argv[1].split(',')] if len(sys.argv) > 1 else [256, 4096]
once = lambda fn, warm=1: (fn(), torch.cuda.synchronize(), 1.0)[-1]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
for B in sizes:
    flush.zero_()
    mb.bench_image_warp(B, 128, once)
