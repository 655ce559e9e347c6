"""GPU-resident synthetic (PD)S-COCO pair source: the B200 replacement for the reference's CPU input pipeline
(DataLoader workers running ``HomographyNetPrep`` etc., reference train.py:80-137, src/data/transforms.py:456-743).

A pool of uint8 RGB images lives in HBM; every batch is drawn and rendered by the K5 kernels
(``bh_pairgen_draw`` + ``bh_pairgen_apply``): no worker processes, no pickling, no host->device copy of
240x320x3 float images the model never reads.  The batch dict has the reference's keys
that the model consumes: patch_1, patch_2 [B,1,P,P], delta [B,4,2] (ground truth for MACE).
"""
import os

import numpy as np
import torch

from .. import functional as F


def transform_args(transforms, name='HomographyNetPrep'):
    """pull [rho, patch_size, distort_keys, max_delta, target_gen] and the standardisation out of a DATA.TRANSFORMS list"""
    out = {'rho': 32, 'patch_size': 128, 'max_delta': 32, 'mean': 0.443, 'std': 0.129}
    for t in transforms:
        (k, v), = t.items()
        if k == name:
            out['rho'], out['patch_size'] = int(v[0]), int(v[1])
            if len(v) > 3:
                out['max_delta'] = float(v[3])
            if len(v) > 4 and v[4] != '4_points':
                raise NotImplementedError("GPU pair generation implements target_gen '4_points' only")
        elif k == 'DictStandardize':
            m, s = v[0], v[1]
            out['mean'] = float(m[0] if isinstance(m, (list, tuple)) else m)
            out['std'] = float(s[0] if isinstance(s, (list, tuple)) else s)
    return out


def synthetic_pool(n_images, height=240, width=320, seed=1234, device='cuda'):
    """COCO-like smooth uint8 RGB images generated on the device (no dataset in the container / on the GPU box)."""
    g = torch.Generator(device=device).manual_seed(seed)
    lo = torch.rand(n_images, 3, height // 8 + 2, width // 8 + 2, device=device, generator=g) * 255
    im = torch.nn.functional.interpolate(lo, size=(height, width), mode='bicubic', align_corners=True)
    im = im + 6.0 * torch.randn(n_images, 3, height, width, device=device, generator=g)
    return im.clamp_(0, 255).round_().to(torch.uint8).permute(0, 2, 3, 1).contiguous()


def load_pool(directory, device='cuda', limit=None):
    """uint8 [n,H,W,3] pool from a directory of preprocessed ``.npy`` images (reference preprocess_offline.py output)."""
    names = sorted(f for f in os.listdir(directory) if f.endswith('.npy'))
    if limit:
        names = names[:limit]
    if not names:
        raise FileNotFoundError('no .npy images under %s' % directory)
    arrs = [np.load(os.path.join(directory, f), allow_pickle=True) for f in names]
    return torch.from_numpy(np.stack(arrs)).to(device)


def rank_seed(seed, rank):
    """rank-seeded generator key: ranks of a data-parallel job draw disjoint streams"""
    return (int(seed) * 1000003 + int(rank)) & (2 ** 63 - 1)


class GpuPairLoader:
    """Iterable of batch dicts; ``len`` = steps per epoch (reference DatasetSampler.__len__)."""

    def __init__(self, pool, batch_size, samples_per_epoch, rho=32, patch_size=128, max_delta=32.0, mean=0.443, std=0.129,
                 seed=42, rank=0):
        assert pool.is_cuda and pool.dtype == torch.uint8 and pool.dim() == 4
        self.pool = pool
        self.batch_size = int(batch_size)
        self.steps = int(samples_per_epoch) // self.batch_size
        self.cfg = dict(rho=int(rho), patch_size=int(patch_size), max_delta=float(max_delta), mean=float(mean), std=float(std))
        self.seed = rank_seed(seed, rank)
        self.step = 0

    def __len__(self):
        return self.steps

    def next_batch(self):
        c = self.cfg
        n, hi, wi, _ = self.pool.shape
        params, index = F.pairgen_draw(self.batch_size, n, (hi, wi), c['rho'], c['patch_size'], c['max_delta'], self.seed,
                                       self.step, self.pool.device)
        p1, p2, delta = F.pairgen_apply(self.pool, index, params, c['patch_size'], c['mean'], c['std'])
        self.step += 1
        return {'patch_1': p1, 'patch_2': p2, 'delta': delta}

    def __iter__(self):
        for _ in range(self.steps):
            yield self.next_batch()
