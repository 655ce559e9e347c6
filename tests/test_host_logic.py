"""CPU: host-side logic around the kernels -- config factory, checkpoint-compatible module names, numpy twins,
refusal to run the hot path on CPU tensors, GPU pair-loader argument plumbing."""
import os

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cfg(rel):
    from bihome_b200 import engine
    return engine.load_config(os.path.join(ROOT, 'config', rel))


@pytest.mark.parametrize('rel', ['pds-coco/zeng-bihome-lr-1e-3.yaml', 's-coco/detone-bihome-lr-5e-3.yaml'])
def test_build_model_from_yaml(rel):
    from bihome_b200 import engine
    c = cfg(rel)
    model = engine.build_model(c, pretrained=False)
    assert isinstance(model, torch.nn.Sequential) and model[1].backbone is model[0]
    keys = list(model.state_dict().keys())
    assert any(k.startswith('0.') for k in keys) and any(k.startswith('1.backbone.') for k in keys)
    assert any(k.startswith('1.auxiliary_resnet.resnet.') for k in keys)          # reference checkpoint layout
    trainable = sum(p.numel() for p in model.parameters() if p.requires_grad)
    assert trainable == (10574178 if 'zeng' in rel else 21285640)
    opt, sched = engine.build_optimizer(c, model)
    assert opt.defaults['lr'] == c['SOLVER']['LR'] and sched.milestones


def test_reference_state_dict_compatibility():
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip('reference tree not present')
    from bihome_b200.backbones import Rethinking
    kw = dict(cfg('pds-coco/zeng-bihome-lr-1e-3.yaml')['MODEL']['BACKBONE'], PRETRAINED_RESNET=False)
    ours, ref = Rethinking.Model(**kw), ref_import.load('src.backbones.Rethinking').Model(**kw)
    ours.load_state_dict(ref.state_dict())
    x = {'patch_1': torch.randn(2, 1, 128, 128), 'patch_2': torch.randn(2, 1, 128, 128)}
    ours.eval(), ref.eval()
    with torch.no_grad():
        a, b = ours(dict(x)), ref(dict(x))
    assert torch.equal(a['pf_hat_12'], b['pf_hat_12']) and torch.equal(a['pf_hat_21'], b['pf_hat_21'])


def test_numpy_twins_follow_opencv():
    import cv2
    from bihome_b200.data import utils as U
    c = np.float32([[[10, 20], [138, 20], [138, 148], [10, 148]]])
    d = np.float32([[[3, -2], [-5, 7], [1, 1], [0, -8]]])
    H = U.four_point_to_homography(c, d)
    assert np.allclose(H, cv2.getPerspectiveTransform(c[0], c[0] + d[0]))
    img = np.random.RandomState(0).rand(240, 320).astype(np.float32)
    assert np.array_equal(U.warp_image(img, H, 240, 320), cv2.warpPerspective(img, np.linalg.inv(H), dsize=(320, 240)))
    assert np.array_equal(U.image_shape_to_corners(np.zeros((2, 1, 128, 128), np.float32))[1],
                          np.float32([[0, 0], [128, 0], [128, 128], [0, 128]]))


def test_hot_path_refuses_cpu_tensors():
    from bihome_b200.data import utils as U
    from bihome_b200.heads import PerceptualHead as PH
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        U.four_point_to_homography(torch.zeros(1, 4, 2), torch.zeros(1, 4, 2))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        PH.Model._warp(torch.zeros(1, 1, 8, 8), torch.zeros(1, 4, 2))


def test_transform_args_from_yaml():
    from bihome_b200.data import gpu_pairs
    t = gpu_pairs.transform_args(cfg('pds-coco/zeng-bihome-lr-1e-3.yaml')['DATA']['TRANSFORMS'])
    assert t == {'rho': 32, 'patch_size': 128, 'max_delta': 32.0, 'mean': 0.443, 'std': 0.129}
    t = gpu_pairs.transform_args(cfg('s-coco/detone-bihome-lr-5e-3.yaml')['DATA']['TRANSFORMS'])
    assert t['max_delta'] == 0.0


def test_dsac_scores_single_hypothesis_are_ones():
    from bihome_b200.heads.ransac_utils import DSACSoftmax
    d = DSACSoftmax()
    s = d.score_hypotheses(torch.zeros(3, 10, 2), torch.zeros(3, 10, 2), torch.eye(3).repeat(3, 1, 1, 1))
    assert torch.equal(s, torch.ones(3, 1))
