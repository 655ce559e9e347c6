"""GPU-resident synthetic (PD)S-COCO pair source: the B200 replacement for the reference's CPU input pipeline
(DataLoader workers running ``HomographyNetPrep`` etc., reference train.py:80-137, src/data/transforms.py:456-743).

A pool of uint8 RGB images lives in HBM; every batch is drawn and rendered by the K5 kernels
(``bh_pairgen_draw`` + ``bh_pairgen_apply``): no worker processes, no pickling, no host->device copy of
240x320x3 float images the model never reads.  The batch dict has the reference's keys
that the model consumes: patch_1, patch_2 [B,1,P,P], delta [B,4,2] (ground truth for MACE), corners [B,4,2]
(image-frame patch corners) and target -- delta again for target_gen '4_points', the dense perspective field
[B,2,P,P] for 'all_points' (the supervised zeng-orig configs).
"""
import os

import numpy as np
import torch

from .. import functional as F


def transform_args(transforms, name='HomographyNetPrep'):
    """pull [rho, patch_size, distort_keys, max_delta, target_gen] and the standardisation out of a DATA.TRANSFORMS list"""
    out = {'rho': 32, 'patch_size': 128, 'max_delta': 32, 'mean': 0.443, 'std': 0.129, 'target_gen': '4_points', 'image_keys': ()}
    for t in transforms:
        (k, v), = t.items()
        if k == 'DictToTensor':
            # whole images the model reads besides the patches (s-coco/nguyen-orig: PhotometricHead warps 'image_1')
            out['image_keys'] = tuple(key for key in v[0] if key.startswith('image'))
        if k == name:
            out['rho'], out['patch_size'] = int(v[0]), int(v[1])
            if len(v) > 3:
                out['max_delta'] = float(v[3])
            if len(v) > 4:
                assert v[4] in ('4_points', 'all_points'), 'I do not know this, it should be either \'4_points\' ar \'all_points\''
                out['target_gen'] = v[4]
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


def rescale_center_crop(image, size=(320, 240)):
    """the reference's offline preprocessing (``Rescale((320, 240))`` + ``CenterCrop((320, 240))``,
    src/data/transforms.py:11-46, 87-122): the shorter side is matched, the longer one is cropped around the centre."""
    import cv2
    h, w = image.shape[:2]
    tw, th = size
    ratio = h / w
    if ratio < th / tw:
        nw, nh = int(np.round(th / ratio)), th
    else:
        nw, nh = tw, int(np.round(tw * ratio))
    image = cv2.resize(image, (nw, nh))
    top = (nh - th) // 2 if nh != th else 0
    left = (nw - tw) // 2 if nw != tw else 0
    return image[top:top + th, left:left + tw]


def load_pool(directory, device='cuda', limit=None, size=(320, 240)):
    """uint8 [n,H,W,3] RGB pool from a directory of ``.npy`` images (reference preprocess_offline.py output) and / or
    ``.jpg`` files (read like ``Dataset.load_image``, src/data/coco/dataset.py:95-103, then rescaled and cropped)."""
    names = sorted(f for f in os.listdir(directory) if f.endswith(('.npy', '.jpg')))
    if limit:
        names = names[:limit]
    if not names:
        raise FileNotFoundError('no .npy / .jpg images under %s' % directory)
    arrs = []
    for f in names:
        path = os.path.join(directory, f)
        if f.endswith('.npy'):
            a = np.load(path, allow_pickle=True)
        else:
            import cv2
            a = cv2.cvtColor(cv2.imread(path), cv2.COLOR_BGR2RGB)
        if a.shape[:2] != (size[1], size[0]):
            a = rescale_center_crop(a, size)
        arrs.append(np.ascontiguousarray(a, dtype=np.uint8))
    return torch.from_numpy(np.stack(arrs)).to(device)


def patch_corners(pos_xy, patch_size):
    """centre positions [B,2] -> image-frame corners [B,4,2], clockwise from the top left (transforms.py:527-531)"""
    s = patch_size // 2
    off = torch.tensor([[-s, -s], [s, -s], [s, s], [-s, s]], device=pos_xy.device, dtype=pos_xy.dtype)
    return pos_xy.unsqueeze(1) + off


def perspective_field_target(corners, delta, patch_size):
    """HomographyNetPrep's 'all_points' target (reference src/data/transforms.py:634-687): the displacement
    proj(H, p) - p of every patch pixel p under the image-frame homography corners -> corners + delta, [B,2,P,P].

    Arithmetic as the reference's: the 8x8 system of cv2.getPerspectiveTransform and the projection in float64, the
    projected point rounded to float32 (cv2.perspectiveTransform's output type), the difference to the integer pixel
    coordinate taken exactly, the result cast to float32 (train.py:308-309).  Plain torch ops on the pool's device:
    33 MB of writes per 256 pairs, off the north-star path (only the supervised *-orig configs ask for it)."""
    B = corners.shape[0]
    src = corners.to(torch.float64)
    dst = src + delta.to(torch.float64)
    x, y, u, v = src[..., 0], src[..., 1], dst[..., 0], dst[..., 1]
    zero, one = torch.zeros_like(x), torch.ones_like(x)
    rows_u = torch.stack([x, y, one, zero, zero, zero, -x * u, -y * u], dim=-1)
    rows_v = torch.stack([zero, zero, zero, x, y, one, -x * v, -y * v], dim=-1)
    A = torch.cat([rows_u, rows_v], dim=1)                                   # [B,8,8]
    h = torch.linalg.solve_ex(A, torch.cat([u, v], dim=1).unsqueeze(-1))[0].squeeze(-1)     # _ex: no host sync
    P = int(patch_size)
    r = torch.arange(P, device=src.device, dtype=torch.float64)
    px = (src[:, 0, 0].view(B, 1, 1) + r.view(1, 1, P)).expand(B, P, P)
    py = (src[:, 0, 1].view(B, 1, 1) + r.view(1, P, 1)).expand(B, P, P)
    g = lambda i: h[:, i].view(B, 1, 1)
    w = g(6) * px + g(7) * py + 1.0
    qx = ((g(0) * px + g(1) * py + g(2)) / w).float().double() - px
    qy = ((g(3) * px + g(4) * py + g(5)) / w).float().double() - py
    return torch.stack([qx, qy], dim=1).float()


def rank_seed(seed, rank):
    """rank-seeded generator key: ranks of a data-parallel job draw disjoint streams"""
    return (int(seed) * 1000003 + int(rank)) & (2 ** 63 - 1)


class GpuPairLoader:
    """Iterable of batch dicts; ``len`` = steps per epoch (reference DatasetSampler.__len__)."""

    def __init__(self, pool, batch_size, samples_per_epoch, rho=32, patch_size=128, max_delta=32.0, mean=0.443, std=0.129,
                 seed=42, rank=0, target_gen='4_points', image_keys=()):
        assert pool.is_cuda and pool.dtype == torch.uint8 and pool.dim() == 4
        self.pool = pool
        self.batch_size = int(batch_size)
        self.steps = int(samples_per_epoch) // self.batch_size
        self.cfg = dict(rho=int(rho), patch_size=int(patch_size), max_delta=float(max_delta), mean=float(mean), std=float(std))
        self.target_gen = target_gen
        self.image_keys = tuple(image_keys)
        unknown = [k for k in self.image_keys if k != 'image_1']
        if unknown:
            raise NotImplementedError('GpuPairLoader renders the patches and image_1; the config also asks for %s as tensors (no '
                                      'shipped configuration reads them)' % unknown)
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
        corners = patch_corners(params[:, 22:24], c['patch_size'])           # pos_x, pos_y of the params table
        target = delta if self.target_gen == '4_points' else perspective_field_target(corners, delta, c['patch_size'])
        batch = {'patch_1': p1, 'patch_2': p2, 'delta': delta, 'corners': corners.float(), 'target': target}
        if 'image_1' in self.image_keys:
            batch['image_1'] = F.pairgen_image(self.pool, index, params, c['mean'], c['std'])
        return batch

    def __iter__(self):
        return self.batches(self.steps)

    def batches(self, n):
        """the next n batches of the stream (a resumed run finishes a partially trained epoch with fewer than len(self))"""
        for _ in range(int(n)):
            yield self.next_batch()

    def shard(self, rank, world):
        """batches rank, rank + world, ... of one epoch of the stream that starts at step 0 (data-parallel evaluation)"""
        for i in range(int(rank), self.steps, int(world)):
            self.step = i
            yield self.next_batch()


def loader_from_config(config, split, device, rank=0, batch_size=None, synthetic_images=256, pool_limit=None):
    """GPU pair source for DATA.TRAIN_SPLIT / DATA.TEST_SPLIT (reference make_coco_dataloader, train.py:80-137)."""
    data = config['DATA']
    train = split == 'train'
    directory = data['TRAIN_SPLIT' if train else 'TEST_SPLIT']
    transforms = data['TRANSFORMS'] if train or 'TEST_TRANSFORM' not in data else data['TEST_TRANSFORM']
    sampler = data['SAMPLER']
    bs = int(batch_size or sampler['BATCH_SIZE'])
    samples = int(sampler['TRAIN_SAMPLES_PER_EPOCH' if train else 'TEST_SAMPLES_PER_EPOCH'])
    seed = int(sampler['TRAIN_SEED' if train else 'TEST_SEED'])
    if os.path.isdir(directory) and any(f.endswith(('.npy', '.jpg')) for f in os.listdir(directory)):
        pool = load_pool(directory, device=device, limit=pool_limit)
    else:
        if rank == 0:
            print('bihome_b200: %s holds no images; using a synthetic pool of %d images' % (directory, synthetic_images))
        pool = synthetic_pool(synthetic_images, seed=1234 if train else 4321, device=device)
    return GpuPairLoader(pool, bs, samples, seed=seed, rank=rank, **transform_args(transforms))
