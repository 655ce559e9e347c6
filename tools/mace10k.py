#!/usr/bin/env python
"""BASELINE.json configs[4] for real: MACE on 10 000 synthetic PDS-COCO pairs, random-init Zeng backbone, reference path
against the B200 path (reference eval.py:60-134,334-341).

    python tools/mace10k.py --pairs 10000 --batch 250 --out gpurun_out/mace10k.json

1. the pairs come from the CPU restatement of the reference's transform pipeline (oracle/pairgen.py, pinned bit for bit
   against the unmodified reference transforms by tests/test_oracle_golden.py), numpy RandomState(42) = DATA.SAMPLER.TEST_SEED,
   rendered by a pool of host processes;
2. both sides use the same random-init weights (torch.manual_seed(0)) and the same multinomial point draws (the reference
   samples with torch.multinomial on whatever generator its device has; identical draws are the only way to compare two
   implementations pair by pair);
3. reference path = the torch modules of the backbone run through plain ATen (NCHW, TF32 off, no custom kernel anywhere)
   + the oracle's DSAC / kornia find_homography_dlt / corner projection in float64 (oracle/ref_path.py): on the GPU for all
   pairs, and with the backbone on the host cores for the first --cpu-ref pairs (the reference's CPU path);
4. B200 path = eval.py's own evaluate() (channels-last backbone, K4 N-point DLT, bh_mace), once per field head
   (BH_FIELD_HEAD = aten | fused).
Writes one JSON object: the MACE of every arm, the largest per-batch and per-pair differences.  Test infrastructure: imports
oracle/.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CONFIG = os.path.join(ROOT, 'config', 'pds-coco', 'zeng-bihome-lr-1e-3.yaml')


def _render(args):
    from oracle import pairgen
    i, q = args
    out = pairgen.make_pair(pairgen.synthetic_image(i % 64), q, 128)
    return (pairgen.to_network_input(out['patch_1']).astype(np.float32), pairgen.to_network_input(out['patch_2']).astype(np.float32),
            q['delta'].astype(np.float32))


def make_pairs(n, seed=42, workers=None):
    from oracle import pairgen
    rs = np.random.RandomState(seed)
    qs = [pairgen.draw_params(rs, 240, 320, 32, 128, 32) for _ in range(n)]
    workers = workers or min(os.cpu_count() or 1, 32)
    with mp.get_context('fork').Pool(workers) as pool:
        res = pool.map(_render, list(enumerate(qs)), chunksize=32)
    p1, p2, d = zip(*res)
    return np.stack(p1), np.stack(p2), np.stack(d)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--pairs', type=int, default=10000)
    ap.add_argument('--batch', type=int, default=250)
    ap.add_argument('--cpu-ref', type=int, default=500, help='pairs also run with the reference backbone on the host cores')
    ap.add_argument('--out', type=str, default=None)
    a = ap.parse_args()
    t0 = time.time()
    p1, p2, delta = make_pairs(a.pairs)
    t_pairs = time.time() - t0

    import torch
    from bihome_b200 import engine
    from oracle import ref_path as R
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from conftest import load_entry
    ev = load_entry('eval')
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    n, bs = a.pairs, a.batch
    nb = n // bs
    cfg = engine.load_config(CONFIG)
    torch.manual_seed(0)
    seq = engine.build_model(cfg, pretrained=False)
    seq.eval()
    M = cfg['MODEL']['HEAD']['POINTS_PER_HYPOTHESIS']
    g = torch.Generator().manual_seed(7)
    choices = [R.multinomial_choice(128 * 128, bs * M, generator=g) for _ in range(nb)]

    def reference(device, batches):
        """plain torch modules + float64 oracle head; returns per-pair delta_hat and per-batch MACE"""
        os.environ['BH_FIELD_HEAD'] = 'aten'
        bb = seq[0].to(device)
        hats, maces = [], []
        with torch.no_grad():
            for i in range(batches):
                sl = slice(i * bs, (i + 1) * bs)
                data = bb({'patch_1': torch.from_numpy(p1[sl]).to(device), 'patch_2': torch.from_numpy(p2[sl]).to(device)})
                dh, _, _ = R.zeng_delta_hat(data['pf_hat_12'].double().cpu(), M, 1, choice=choices[i])
                hats.append(dh.reshape(bs, 4, 2).numpy())
                maces.append(R.mace(delta[sl], hats[-1]))
        return np.concatenate(hats), maces

    t0 = time.time()
    ref_hat, ref_maces = reference('cuda', nb)
    t_ref = time.time() - t0
    cpu_batches = min(nb, max(a.cpu_ref // bs, 0))
    cpu_hat, cpu_maces = (reference('cpu', cpu_batches) if cpu_batches else (None, []))
    seq.to('cuda')

    results = {}
    path = os.path.join(os.environ.get('TMPDIR', '/tmp'), 'mace10k_pairs.npz')
    np.savez(path, patch_1=p1[:nb * bs], patch_2=p2[:nb * bs], delta=delta[:nb * bs])
    for side in ('aten', 'fused'):
        os.environ['BH_FIELD_HEAD'] = side
        model = ev.ModelWrapper(seq[0], seq[1]).cuda().to(memory_format=torch.channels_last)
        model.eval()
        head = model[1]
        orig = head._field_to_delta
        it = iter(choices)
        hats = []

        def forced(pf, which, _orig=orig, _it=it, _head=head, _hats=hats):
            _head.forced_choice = [next(_it).cuda(), None]
            out = _orig(pf, which)
            _hats.append(out[0].reshape(-1, 4, 2).float().cpu().numpy())
            return out
        head._field_to_delta = forced
        mean_mace, maces, ms = ev.evaluate(model, ev.fixed_batches(path, bs, 'cuda'))
        del head._field_to_delta
        head.forced_choice = None
        hat = np.concatenate(hats)
        per_pair = np.linalg.norm((hat - ref_hat).reshape(-1, 2), axis=-1).reshape(-1, 4)
        results[side] = {'mace': mean_mace, 'max_abs_batch_diff_vs_reference': float(np.max(np.abs(np.array(maces) - np.array(ref_maces)))),
                         'abs_mean_diff_vs_reference': abs(mean_mace - float(np.mean(ref_maces))),
                         'max_corner_distance_to_reference_px': float(per_pair.max()),
                         'mean_corner_distance_to_reference_px': float(per_pair.mean()), 'ms_per_batch': ms}
    os.remove(path)
    out = {'pairs': nb * bs, 'batch': bs, 'seed': 42, 'config': 'pds-coco/zeng-bihome-lr-1e-3.yaml', 'weights': 'random init, torch.manual_seed(0)',
           'reference_gpu_aten_plus_float64_oracle_head': {'mace': float(np.mean(ref_maces)), 'seconds': t_ref},
           'reference_cpu_backbone': ({'pairs': cpu_batches * bs, 'mace': float(np.mean(cpu_maces)),
                                       'max_abs_batch_diff_vs_gpu_reference': float(np.max(np.abs(np.array(cpu_maces) - np.array(ref_maces[:cpu_batches])))),
                                       'b200_aten_head_mace_on_the_same_pairs': None} if cpu_batches else None),
           'b200_path': results, 'pair_generation_seconds': t_pairs, 'host_workers': min(os.cpu_count() or 1, 32),
           'tolerance_px': 0.01,
           'within_tolerance': all(r['max_abs_batch_diff_vs_reference'] <= 0.01 and r['abs_mean_diff_vs_reference'] <= 0.01 for r in results.values())}
    text = json.dumps(out, indent=1)
    print(text)
    if a.out:
        with open(a.out, 'w') as f:
            f.write(text + '\n')
    return 0 if out['within_tolerance'] else 1


if __name__ == '__main__':
    sys.exit(main())
