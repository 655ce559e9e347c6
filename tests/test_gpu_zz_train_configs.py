"""train.py / eval.py on the other backbone x head combinations the reference ships (zhang-bihome, zhang-orig,
zeng-orig, detone-orig): a few optimisation steps on the GPU pair source, a checkpoint, an evaluation pass.
The configs are derived from the committed zeng-bihome YAML by swapping MODEL / SOLVER.LOSS / the target generator,
exactly the sections in which the reference's own files differ.  Sorted after the kernel and north-star tests."""
import copy
import os

import numpy as np
import pytest
import torch
import yaml

from conftest import load_entry

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _aten_field_head(monkeypatch):
    """these tests pin the K1-K5 path and the entry points; the Zeng backbone's last stage stays on the ATen modules here
    whatever the device's K6 self-test says (K6 has its own file, tests/test_gpu_zzz_field_head.py)"""
    monkeypatch.setenv('BH_FIELD_HEAD', 'aten')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = os.path.join(ROOT, 'config', 'pds-coco', 'zeng-bihome-lr-1e-3.yaml')

KEYS = dict(PATCH_KEYS=['patch_1', 'patch_2'])
CONTENT_AWARE = dict(NAME='ContentAware', VARIANT='DoubleLine', IMAGE_SIZE=128, PRETRAINED_RESNET=False, IMAGE_KEY=['image'],
                     MASK_KEYS=['mask_1', 'mask_2'], FIX_MASK=True, FEATURE_KEYS=['feature_1', 'feature_2'],
                     TARGET_KEYS=['delta_hat_12', 'delta_hat_21'], **KEYS)
BIHOME_HEAD = dict(NAME='PerceptualHead', PATCH_SIZE=128, DELTA_HAT_KEYS=['delta_hat_12', 'delta_hat_21'], PF_KEYS=[],
                   RANSAC_HYPOTHESIS_NO=-1, POINTS_PER_HYPOTHESIS=-1, AUXILIARY_RESNET='resnet34', AUXILIARY_RESNET_OUTPUT_LAYER=1,
                   AUXILIARY_RESNET_PRETRAINED=False, TRIPLET_LOSS='double-line', TRIPLET_AGGREGATION='channel-agnostic',
                   TRIPLET_MARGIN='inf', TRIPLET_DISTANCE='l1', TRIPLET_MU=0.01, MASK_KEYS=[], SAMPLING_STRATEGY='downsample-mask',
                   **KEYS)
COMBOS = {
    'zhang-bihome': (CONTENT_AWARE, BIHOME_HEAD, 'biHomE', '4_points'),
    'zhang-orig': (CONTENT_AWARE,
                   dict(NAME='TripletHead', VARIANT='DoubleLine', PATCH_SIZE=128, MASK_KEYS=['mask_1', 'mask_2'],
                        FEATURE_KEYS=['feature_1', 'feature_2'], TARGET_KEYS=['delta_hat_12', 'delta_hat_21'], LD=2, MU=0.01,
                        TRIPLET_MARGIN=1.0, TRIPLET_AGGREGATION='channel-agnostic', **KEYS), 'TripletLoss', '4_points'),
    'zeng-orig': (dict(NAME='Rethinking', VARIANT='OneLine', IMAGE_SIZE=128, RESNET_BLOCK='ResNet34', PRETRAINED_RESNET=False,
                       IMAGE_KEY=['image'], TARGET_KEYS=['pf_hat_12'], **KEYS),
                  dict(NAME='NoOpHead', TARGET_GEN='all_points', LEARNING_KEYS=['target', 'pf_hat_12', 'delta', 'pf_hat_12']),
                  'SmoothL1Loss', 'all_points'),
    'detone-orig': (dict(NAME='ResNet34', VARIANT='OneLine', IMAGE_SIZE=128, PRETRAINED_RESNET=False, IMAGE_KEY=['image'],
                         TARGET_KEYS=['delta_hat_12'], **KEYS),
                    dict(NAME='NoOpHead', TARGET_GEN='4_points', LEARNING_KEYS=['delta', 'delta_hat_12', 'delta', 'delta_hat_12']),
                    'MSELoss', '4_points'),
}


def write_config(tmp_path, name):
    backbone, head, loss, target_gen = COMBOS[name]
    with open(BASE) as f:
        cfg = yaml.full_load(f)
    cfg = copy.deepcopy(cfg)
    cfg['MODEL']['BACKBONE'], cfg['MODEL']['HEAD'] = dict(backbone), dict(head)
    cfg['SOLVER']['LOSS'] = loss
    for key in ('TRANSFORMS', 'TEST_TRANSFORM'):
        for t in cfg['DATA'].get(key, []):
            if 'HomographyNetPrep' in t:
                t['HomographyNetPrep'][4] = target_gen
    path = os.path.join(str(tmp_path), name + '.yaml')
    with open(path, 'w') as f:
        yaml.safe_dump(cfg, f)
    return path


@pytest.mark.parametrize('name', sorted(COMBOS))
def test_train_and_eval_entry_points(tmp_path, name):
    train, ev = load_entry('train'), load_entry('eval')
    path = write_config(tmp_path, name)
    log_dir = os.path.join(str(tmp_path), 'log')
    train.main(path, batch_size=4, max_steps=3, synthetic_pool=8, log_dir=log_dir)
    ckpt = os.path.join(log_dir, 'model_000003.pth')
    assert os.path.isfile(ckpt)
    blob = torch.load(ckpt, map_location='cpu', weights_only=False)
    assert blob['step'] == 3
    assert all(torch.isfinite(v).all() for v in blob['model'].values() if torch.is_floating_point(v))
    mace = ev.main(path, ckpt, batch_size=4, samples=8)
    assert np.isfinite(mace)
