"""BASELINE.json configs[4]: MACE of the reference path vs the B200 path on one fixed synthetic PDS-COCO set, same
random-init weights, same multinomial draws (north_star: |MACE difference| <= 0.01 px)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _aten_field_head(monkeypatch):
    """these tests pin the K1-K5 path and the entry points; the Zeng backbone's last stage stays on the ATen modules here
    whatever the device's K6 self-test says (K6 has its own file, tests/test_gpu_zzz_field_head.py)"""
    monkeypatch.setenv('BH_FIELD_HEAD', 'aten')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONFIG = os.path.join(ROOT, 'config', 'pds-coco', 'zeng-bihome-lr-1e-3.yaml')


def reference_pairs(n, seed=42):
    """pairs made by the CPU restatement of the reference transform pipeline (oracle/pairgen.py)"""
    from oracle import pairgen
    rs = np.random.RandomState(seed)
    p1, p2, d = [], [], []
    for i in range(n):
        q = pairgen.draw_params(rs, 240, 320, 32, 128, 32)
        out = pairgen.make_pair(pairgen.synthetic_image(i % 8), q, 128)
        p1.append(pairgen.to_network_input(out['patch_1']))
        p2.append(pairgen.to_network_input(out['patch_2']))
        d.append(q['delta'].astype(np.float32))
    return np.stack(p1), np.stack(p2), np.stack(d)


def test_mace_matches_reference_path(tmp_path):
    from conftest import load_entry
    ev = load_entry('eval')
    from bihome_b200 import engine
    from oracle import ref_path as R
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    n, bs = 48, 16
    p1, p2, delta = reference_pairs(n)
    cfg = engine.load_config(CONFIG)
    torch.manual_seed(0)
    seq = engine.build_model(cfg, pretrained=False)
    # a random-init field is pure noise: give the last layer a gain so the field carries a few pixels of structure
    with torch.no_grad():
        seq[0].layer8[-1].weight.mul_(20.0)
    seq.eval()
    M = cfg['MODEL']['HEAD']['POINTS_PER_HYPOTHESIS']
    g = torch.Generator().manual_seed(7)
    choices = [R.multinomial_choice(128 * 128, bs * M, generator=g) for _ in range(n // bs)]

    # reference path on the CPU: torch backbone + oracle DSAC/DLT/corner projection (PerceptualHead.py:716-767)
    ref_maces, ref_hats = [], []
    with torch.no_grad():
        for i in range(n // bs):
            sl = slice(i * bs, (i + 1) * bs)
            data = seq[0]({'patch_1': torch.from_numpy(p1[sl]), 'patch_2': torch.from_numpy(p2[sl])})
            dh, _, _ = R.zeng_delta_hat(data['pf_hat_12'].double(), M, 1, choice=choices[i])
            ref_hats.append(dh.reshape(bs, 4, 2).numpy())
            ref_maces.append(R.mace(delta[sl], ref_hats[-1]))

    # B200 path through eval.py's own evaluate()
    model = ev.ModelWrapper(seq[0], seq[1]).cuda()
    model.eval()
    it = iter(choices)
    head = model[1]
    orig = head._field_to_delta

    def forced(pf, which):
        head.forced_choice = [next(it).cuda(), None]
        return orig(pf, which)
    head._field_to_delta = forced
    path = os.path.join(str(tmp_path), 'pairs.npz')
    np.savez(path, patch_1=p1, patch_2=p2, delta=delta)
    mean_mace, maces, _ = ev.evaluate(model, ev.fixed_batches(path, bs, 'cuda'))
    assert len(maces) == n // bs
    assert abs(mean_mace - float(np.mean(ref_maces))) <= 0.01, (mean_mace, np.mean(ref_maces))
    for a, b in zip(maces, ref_maces):
        assert abs(a - b) <= 0.01, (a, b)
    assert mean_mace > 1.0          # a random-init model is far from the ground truth: the comparison is not vacuous


def test_train_entry_point_runs_and_resumes(tmp_path):
    """train.py main(): a few steps of the shipped Zeng config, checkpoint, resume from last_checkpoint.txt"""
    from conftest import load_entry
    train = load_entry('train')
    log_dir = os.path.join(str(tmp_path), 'log')
    train.main(CONFIG, batch_size=8, max_steps=3, synthetic_pool=8, log_dir=log_dir)
    assert os.path.isfile(os.path.join(log_dir, 'model_000003.pth'))
    blob = torch.load(os.path.join(log_dir, 'model_000003.pth'), map_location='cpu', weights_only=False)
    assert blob['step'] == 3 and blob['scheduler']['last_epoch'] == 3
    train.main(CONFIG, batch_size=8, max_steps=5, synthetic_pool=8, log_dir=log_dir)
    assert os.path.isfile(os.path.join(log_dir, 'model_000005.pth'))
    ev = load_entry('eval')
    mace = ev.main(CONFIG, os.path.join(log_dir, 'model_000005.pth'), batch_size=8, samples=32)
    assert np.isfinite(mace)


def test_config5_mace_subsample_both_field_heads(tmp_path, monkeypatch):
    """BASELINE.json configs[4] on a 500-pair sub-sample (the 10 000-pair run is tools/mace10k.py under tools/gpu_round.sh,
    its record is profiles/r02_mace10k.json): unmodified random-init weights -- no hand-scaled layer --, reference path =
    plain ATen backbone + float64 oracle head on the GPU and on the host cores, B200 path = eval.py's evaluate() with the ATen
    AND the fused (K6) field head; every per-batch MACE and the mean within 0.01 px (north_star)."""
    import json
    import subprocess
    import sys
    monkeypatch.delenv('BH_FIELD_HEAD', raising=False)
    out = os.path.join(str(tmp_path), 'mace.json')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'mace10k.py'), '--pairs', '500', '--batch', '250', '--cpu-ref', '250',
                        '--out', out], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.load(open(out))
    assert res['within_tolerance'] and res['pairs'] == 500
    for side in ('aten', 'fused'):
        assert res['b200_path'][side]['max_abs_batch_diff_vs_reference'] <= 0.01
        assert res['b200_path'][side]['mace'] > 1.0          # not vacuous: a random-init model is far from the ground truth
    assert res['reference_cpu_backbone']['max_abs_batch_diff_vs_gpu_reference'] <= 0.01
