#!/usr/bin/env python
"""MACE evaluation entry point -- the reference's ``eval.py`` contract on the B200 path.

    python eval.py --config_file config/pds-coco/zeng-bihome-lr-1e-3.yaml --ckpt log/.../model_090000.pth
    python eval.py --config_file ... --random_init --samples 10000      (BASELINE.json configs[4])

As reference ``eval.py:349-513`` (main) / ``:60-346`` (evaluate): builds backbone + head by NAME, wraps them in a
``ModelWrapper`` whose ``predict_homography`` chains the two (``eval.py:21-28``), loads ``--ckpt``, runs the test
transforms of the config and reports the mean of the per-batch MACE plus the mean model time from CUDA events with
the first iteration dropped (``eval.py:83-88,128-134,334-341``).  ``--ckpt`` may be replaced by ``--random_init``
because there are no trained weights offline.  Pairs come from the GPU generator (K5) with ``DATA.SAMPLER.TEST_SEED``;
``--pairs file.npz`` evaluates a fixed set (patch_1, patch_2, delta) instead -- that is how reference and new
implementation are compared on identical tensors (tests/test_gpu_eval.py).
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from bihome_b200 import engine  # noqa: E402
from bihome_b200 import functional as F  # noqa: E402
from bihome_b200.data.gpu_pairs import loader_from_config as make_pair_loader  # noqa: E402


class ModelWrapper(torch.nn.Sequential):
    """Sequential(backbone, head) with the reference's chained ``predict_homography`` (eval.py:21-28)."""

    def __init__(self, backbone, head):
        super().__init__(backbone, head)

    def predict_homography(self, data):
        for module in self:
            data = module.predict_homography(data)
        return data


def evaluate(model, batches, log_filepath=None, seed=None):
    """batches: iterable of dicts with patch_1, patch_2, delta (CUDA).  Returns (mean MACE, per-batch MACE list, mean ms)."""
    model.eval()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    maces, times = [], []
    if seed is not None:
        torch.manual_seed(seed)          # the Zeng head draws its 128 points with torch.multinomial
    with torch.no_grad():
        for it, data in enumerate(batches):
            e0.record()
            delta_hat, _ = model.predict_homography(data)
            e1.record()
            if not torch.is_tensor(delta_hat):      # NoOpHead 'all_points': cv2 RANSAC post-processing returns numpy
                delta_hat = torch.as_tensor(np.asarray(delta_hat), device=data['delta'].device)
            mace = F.mace(data['delta'].float(), delta_hat.float())
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
            maces.append(float(mace))
            if log_filepath is not None:
                with open(log_filepath, 'a') as f:
                    f.write(str(it) + ',' + str(maces[-1]) + '\n')
    mean_ms = float(np.mean(times[1:])) if len(times) > 1 else float('nan')
    return float(np.mean(maces)), maces, mean_ms


def fixed_batches(path, batch_size, device):
    z = np.load(path)
    keys = ['patch_1', 'patch_2', 'delta'] + [k for k in ('corners', 'target') if k in z.files]
    arrays = {k: torch.from_numpy(z[k]).float() for k in keys}
    for i in range(0, arrays['patch_1'].shape[0], batch_size):
        yield {k: v[i:i + batch_size].to(device) for k, v in arrays.items()}


def main(config_file_path, ckpt_file_path=None, batch_size=None, visualize=False, log_filepath=None, samples=None,
         pairs=None, random_init=False, seed=None):
    config = engine.load_config(config_file_path)
    if not torch.cuda.is_available():
        raise SystemExit('eval.py: no CUDA device -- the hot path runs on sm_100a kernels only (no CPU fallback)')
    if visualize:
        raise NotImplementedError('--vis (matplotlib figures of the reference, eval.py:140-330) is out of scope')
    device = torch.device('cuda', 0)
    np.random.seed(config['DATA']['SAMPLER']['TEST_SEED'])
    torch.manual_seed(0)
    seq = engine.build_model(config, pretrained=False if random_init else None)
    model = ModelWrapper(seq[0], seq[1]).to(device).to(memory_format=torch.channels_last)
    if ckpt_file_path:
        blob = torch.load(ckpt_file_path, map_location='cpu', weights_only=False)
        model.load_state_dict(blob['model'])
    elif not random_init:
        raise SystemExit('eval.py: --ckpt is required (or pass --random_init)')
    bs = int(batch_size or config['DATA']['SAMPLER']['BATCH_SIZE'])
    if pairs:
        batches = fixed_batches(pairs, bs, device)
    else:
        if samples:
            config['DATA']['SAMPLER']['TEST_SAMPLES_PER_EPOCH'] = int(samples)
        batches = make_pair_loader(config, 'test', device, 0, bs)
    n_params = sum(p.numel() for p in seq[0].parameters())
    mean_mace, maces, mean_ms = evaluate(model, batches, log_filepath, seed)
    print('Number of backbone parameters: {}'.format(n_params))
    print('MACE: {:.4f} over {} batches of {}'.format(mean_mace, len(maces), bs))
    print('Mean model time: {:.3f} ms per batch'.format(mean_ms))
    return mean_mace


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--config_file', type=str, required=True)
    ap.add_argument('--ckpt', type=str, default=None)
    ap.add_argument('--batch_size', type=int, default=None)
    ap.add_argument('--vis', action='store_true')
    ap.add_argument('--log', type=str, default=None, help='per-batch MACE csv')
    ap.add_argument('--samples', type=int, default=None, help='override DATA.SAMPLER.TEST_SAMPLES_PER_EPOCH')
    ap.add_argument('--pairs', type=str, default=None, help='.npz with patch_1, patch_2, delta: evaluate this fixed set')
    ap.add_argument('--random_init', action='store_true')
    ap.add_argument('--seed', type=int, default=None, help='torch seed for the multinomial point draw')
    a = ap.parse_args()
    main(a.config_file, a.ckpt, a.batch_size, a.vis, a.log, a.samples, a.pairs, a.random_init, a.seed)
