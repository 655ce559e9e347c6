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


@pytest.mark.parametrize('rel', ['pds-coco/zeng-bihome-lr-1e-3.yaml', 's-coco/detone-bihome-lr-5e-3.yaml'])
def test_committed_configs_equal_the_reference_files(rel):
    """config/ holds normalised YAML: every key and value must equal the reference's file of the same name"""
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip('reference tree not present')
    from bihome_b200 import engine
    assert cfg(rel) == engine.load_config(os.path.join(ref_import.REFERENCE_ROOT, 'config', rel))


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
    # K3g, the fused loss of every other variant, and the CUDA-graph step: same rule
    import bihome_b200.functional as F
    from bihome_b200 import engine
    f = torch.zeros(1, 4, 4, 4)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        F.triplet_loss(f, f, f, f, torch.ones(1, 4, 4), None, torch.ones(1, 4, 4), None, torch.eye(3)[None], torch.eye(3)[None])
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match='no CPU fallback'):
            engine.GraphedStep(torch.nn.Identity(), {'patch_1': f})


def test_loss_variant_mapping_of_the_mirrors():
    """TRIPLET_* settings -> (distance, hinge, margin) of F.triplet_loss, the way the reference's branches read them
    (PerceptualHead.py:465-538, 609-665; TripletHead.py:79-95)"""
    from bihome_b200.heads import PerceptualHead as PH, TripletHead as TH
    base = dict(PATCH_SIZE=32, PATCH_KEYS=['patch_1', 'patch_2'], DELTA_HAT_KEYS=['a', 'b'], PF_KEYS=[], RANSAC_HYPOTHESIS_NO=-1,
                POINTS_PER_HYPOTHESIS=-1, AUXILIARY_RESNET='resnet18', AUXILIARY_RESNET_OUTPUT_LAYER=1, AUXILIARY_RESNET_PRETRAINED=False,
                TRIPLET_MU=0.01, MASK_KEYS=[], SAMPLING_STRATEGY='downsample-mask', TRIPLET_DISTANCE='l1')

    def variant(double, **kw):
        cfg = dict(base)
        cfg.update(kw)
        return PH.Model(backbone=torch.nn.Identity(), **cfg)._loss_variant(double)
    assert variant(True, TRIPLET_LOSS='double-line', TRIPLET_MARGIN='inf', TRIPLET_AGGREGATION='channel-agnostic') == ('l1', None, 0.0)
    assert variant(True, TRIPLET_LOSS='double-line', TRIPLET_MARGIN=0.05, TRIPLET_AGGREGATION='channel-aware') == ('l1', 'channel', 0.05)
    assert variant(True, TRIPLET_LOSS='double-line', TRIPLET_MARGIN=1.0, TRIPLET_AGGREGATION='channel-agnostic') == ('l1', 'pixel', 1.0)
    assert variant(True, TRIPLET_LOSS='double-line', TRIPLET_MARGIN=0.1, TRIPLET_AGGREGATION='channel-aware',
                   TRIPLET_DISTANCE='cosine') == ('cosine', 'pixel', 0.1)
    assert variant(False, TRIPLET_LOSS='one-line', TRIPLET_MARGIN=0.2, TRIPLET_AGGREGATION='channel-agnostic',
                   TRIPLET_DISTANCE='cosine') == ('cosine', 'pixel', 0.2)
    with pytest.raises(AssertionError, match='distance metric'):
        variant(False, TRIPLET_LOSS='one-line', TRIPLET_MARGIN=0.2, TRIPLET_AGGREGATION='channel-agnostic', TRIPLET_DISTANCE='l2')
    kw = dict(PATCH_KEYS=['patch_1', 'patch_2'], MASK_KEYS=['m1', 'm2'], FEATURE_KEYS=['f1', 'f2'], TARGET_KEYS=['a', 'b'], LD=2, MU=0.01,
              VARIANT='DoubleLine')
    head = TH.Model(None, TRIPLET_MARGIN=1.0, TRIPLET_AGGREGATION='channel-agnostic', **kw)
    assert head._loss_variant(64, 1) == ('pixel', 1.0, 64.0)          # the B-fold broadcast of TripletHead.py:91-92
    with pytest.raises(RuntimeError, match='only defined for'):
        head._loss_variant(4, 8)
    assert TH.Model(None, TRIPLET_MARGIN='inf', TRIPLET_AGGREGATION='channel-aware', **kw)._loss_variant(4, 8) == (None, 0.0, 1.0)
    assert TH.Model(None, TRIPLET_MARGIN=0.05, TRIPLET_AGGREGATION='channel-aware', **kw)._loss_variant(4, 8) == ('channel', 0.05, 1.0)


def test_transform_args_from_yaml():
    from bihome_b200.data import gpu_pairs
    t = gpu_pairs.transform_args(cfg('pds-coco/zeng-bihome-lr-1e-3.yaml')['DATA']['TRANSFORMS'])
    assert t == {'rho': 32, 'patch_size': 128, 'max_delta': 32.0, 'mean': 0.443, 'std': 0.129, 'target_gen': '4_points',
                 'image_keys': ()}
    t = gpu_pairs.transform_args(cfg('s-coco/detone-bihome-lr-5e-3.yaml')['DATA']['TRANSFORMS'])
    assert t['max_delta'] == 0.0
    # the PhotometricHead config reads the whole first image besides the patches
    nguyen = [{'HomographyNetPrep': [32, 128, [], 0]}, {'DictToTensor': [['image_1', 'patch_1', 'patch_2']]}]
    assert gpu_pairs.transform_args(nguyen)['image_keys'] == ('image_1',)
    zeng_orig = [{'HomographyNetPrep': [32, 128, ['image_1', 'image_2'], 32, 'all_points']},
                 {'DictStandardize': [[0.443], [0.129], ['patch_1', 'patch_2']]}]
    assert gpu_pairs.transform_args(zeng_orig)['target_gen'] == 'all_points'


def test_all_points_target_is_bit_exact(golden):
    """the dense perspective-field target of the zeng-orig configs against HomographyNetPrep(target_gen='all_points')
    of the unmodified reference (oracle/make_golden_heads.py): every one of the 2 x 2 x 128 x 128 floats equal"""
    from bihome_b200.data import gpu_pairs
    g = golden('all_points_target.npz')
    corners, delta = torch.as_tensor(g['corners']), torch.as_tensor(g['delta'])
    target = gpu_pairs.perspective_field_target(corners, delta, 128)
    assert target.dtype == torch.float32 and np.array_equal(target.numpy(), g['target'])
    centre = (corners[:, 0] + 64).double()
    assert torch.equal(gpu_pairs.patch_corners(centre, 128), corners.double())


def test_dsac_scores_single_hypothesis_are_ones():
    from bihome_b200.heads.ransac_utils import DSACSoftmax
    d = DSACSoftmax()
    s = d.score_hypotheses(torch.zeros(3, 10, 2), torch.zeros(3, 10, 2), torch.eye(3).repeat(3, 1, 1, 1))
    assert torch.equal(s, torch.ones(3, 1))


def test_checkpointer_round_trip_in_reference_format(tmp_path):
    """reference src/utils/checkpoint.py:31-85: model_%06d.pth + last_checkpoint.txt, {'model','optimizer','scheduler','step'}"""
    from bihome_b200 import engine
    from bihome_b200.utils.checkpoint import CheckPointer
    c = cfg('s-coco/detone-bihome-lr-5e-3.yaml')
    torch.manual_seed(1)
    model = engine.build_model(c, pretrained=False)
    opt, sched = engine.build_optimizer(c, model)
    for _ in range(3):
        sched.step()
    ck = CheckPointer(model, opt, sched, str(tmp_path), save_to_disk=True)
    assert ck.load() == {} and not ck.has_checkpoint()
    path = ck.save('model_{:06d}'.format(3), step=3)
    assert os.path.basename(path) == 'model_000003.pth'
    assert open(os.path.join(str(tmp_path), 'last_checkpoint.txt')).read() == path
    blob = torch.load(path, map_location='cpu', weights_only=False)
    assert set(blob) == {'model', 'optimizer', 'scheduler', 'step'}
    assert any(k.startswith('1.auxiliary_resnet.resnet.') for k in blob['model'])
    torch.manual_seed(2)
    model2 = engine.build_model(c, pretrained=False)
    opt2, sched2 = engine.build_optimizer(c, model2)
    extra = CheckPointer(model2, opt2, sched2, str(tmp_path)).load()
    assert extra == {'step': 3} and sched2.last_epoch == 3
    for a, b in zip(model.state_dict().values(), model2.state_dict().values()):
        assert torch.equal(a, b)
    # rank > 0 never writes
    assert CheckPointer(model, save_dir=str(tmp_path), save_to_disk=False).save('x') is None


def test_entry_points_keep_the_reference_contract():
    import inspect
    from conftest import load_entry
    train, ev = load_entry('train'), load_entry('eval')
    assert list(inspect.signature(train.main).parameters)[0] == 'config_file_path'
    assert list(inspect.signature(ev.main).parameters)[:5] == ['config_file_path', 'ckpt_file_path', 'batch_size', 'visualize',
                                                              'log_filepath']
    for name in ('do_train', 'train_one_epoch', 'eval_one_epoch'):
        assert callable(getattr(train, name))
    assert callable(ev.evaluate) and hasattr(ev.ModelWrapper, 'predict_homography')
    with pytest.raises(SystemExit, match='no CPU fallback'):
        train.main(os.path.join(ROOT, 'config', 'pds-coco', 'zeng-bihome-lr-1e-3.yaml'))


def test_rescale_center_crop_matches_reference_rule():
    from bihome_b200.data.gpu_pairs import rescale_center_crop
    for h, w in ((480, 640), (640, 480), (240, 320), (500, 333), (427, 640)):
        out = rescale_center_crop(np.zeros((h, w, 3), np.uint8))
        assert out.shape == (240, 320, 3)


def test_checkpoints_are_interchangeable_with_the_reference_checkpointer(tmp_path):
    """files written by the reference's CheckPointer (src/utils/checkpoint.py:32-53) resume here, and the other way
    round: same file names, same dictionary, same last_checkpoint.txt protocol (build container only)"""
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip('reference tree not present')
    RefCheckPointer = ref_import.load('src.utils.checkpoint').CheckPointer
    from bihome_b200.utils.checkpoint import CheckPointer

    def trio(seed):
        torch.manual_seed(seed)
        model = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.BatchNorm1d(3))
        opt = torch.optim.Adam(model.parameters(), lr=1e-3)
        sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[2, 4], gamma=0.1)
        model(torch.randn(5, 4)).sum().backward()
        opt.step()
        sched.step()
        return model, opt, sched

    for writer, reader, sub in ((RefCheckPointer, CheckPointer, 'ref_to_ours'), (CheckPointer, RefCheckPointer, 'ours_to_ref')):
        d = str(tmp_path / sub)
        os.makedirs(d, exist_ok=True)
        m1, o1, s1 = trio(1)
        writer(m1, o1, s1, d, save_to_disk=True, device='cpu').save('model_000007', step=7)
        m2, o2, s2 = trio(2)
        extra = reader(m2, o2, s2, d, save_to_disk=False, device='cpu').load()
        assert extra['step'] == 7
        for (k, a), (_, b) in zip(m1.state_dict().items(), m2.state_dict().items()):
            assert torch.equal(a, b), k
        assert s2.state_dict() == s1.state_dict()
        assert o2.state_dict()['param_groups'] == o1.state_dict()['param_groups']


def test_reference_optimizer_state_loads_into_the_engine(tmp_path):
    """the 'optimizer' entry of a reference-written model_%06d.pth (Adam over ALL model.parameters(), frozen extractor
    included: reference train.py:703-707) must fit engine.build_optimizer's param groups, and the other way round --
    a real biHomE config, not a toy model (build container only)"""
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip('reference tree not present')
    RefCheckPointer = ref_import.load('src.utils.checkpoint').CheckPointer
    from bihome_b200 import engine
    from bihome_b200.utils.checkpoint import CheckPointer
    cfg = engine.load_config(os.path.join(ROOT, 'config', 's-coco', 'detone-bihome-lr-5e-3.yaml'))
    bcfg, hcfg = dict(cfg['MODEL']['BACKBONE']), dict(cfg['MODEL']['HEAD'])
    bcfg['PRETRAINED_RESNET'] = False

    def ref_trio():
        rb = ref_import.load('src.backbones.' + bcfg['NAME']).Model(**bcfg)
        model = torch.nn.Sequential(rb, ref_import.load('src.heads.' + hcfg['NAME']).Model(rb, **hcfg))
        s = cfg['SOLVER']
        opt = torch.optim.Adam(model.parameters(), lr=s['LR'], betas=(s['MOMENTUM_1'], s['MOMENTUM_2']), weight_decay=0)
        return model, opt, torch.optim.lr_scheduler.MultiStepLR(opt, milestones=s['MILESTONES'], gamma=s['LR_DECAY'])

    def our_trio():
        model = engine.build_model(cfg, pretrained=False)
        return (model,) + engine.build_optimizer(cfg, model)

    def one_step(model, opt, sched):
        for p in model.parameters():
            if p.requires_grad:
                p.grad = torch.full_like(p, 1e-3)
        opt.step()
        sched.step()

    for make_w, W, make_r, R, sub in ((ref_trio, RefCheckPointer, our_trio, CheckPointer, 'ref_to_ours'),
                                      (our_trio, CheckPointer, ref_trio, RefCheckPointer, 'ours_to_ref')):
        d = str(tmp_path / sub)
        os.makedirs(d, exist_ok=True)
        m1, o1, s1 = make_w()
        one_step(m1, o1, s1)
        W(m1, o1, s1, d, save_to_disk=True, device='cpu').save('model_000001', step=1)
        m2, o2, s2 = make_r()
        assert R(m2, o2, s2, d, save_to_disk=False, device='cpu').load()['step'] == 1
        sd1, sd2 = o1.state_dict(), o2.state_dict()
        assert sd1['param_groups'] == sd2['param_groups']
        assert sorted(sd1['state']) == sorted(sd2['state']) and len(sd1['state']) > 0
        for k in sd1['state']:
            assert torch.equal(sd1['state'][k]['exp_avg'], sd2['state'][k]['exp_avg'])
        one_step(m2, o2, s2)          # and training goes on


def test_bench_reference_arm_line_follows_the_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): ONE JSON line with the GPU arm's metric,
    unit and `config` object, `impl`, a `cpu_baseline` describing the run and an `e2e` block without copies"""
    import json
    import subprocess
    import sys
    bench = os.path.join(ROOT, 'bench.py')
    r = subprocess.run([sys.executable, bench, '--impl', 'reference', '--steps', '1', '--warmup', '0', '--ref-batch', '2'],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    sys.path.insert(0, ROOT)
    import bench as B
    assert line['impl'] == 'reference' and line['metric'] == B.METRIC and line['unit'] == B.UNIT and line['higher_is_better'] is True
    assert line['config'] == B.workload_config(256, 1, 256)          # the same object the GPU arm prints
    assert line['steps'] == 1 and line['warmup'] == 0 and line['n_gpus'] == 1 and line['value'] > 0
    cb = line['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == line['value'] and 'B=2' in cb['sample']
    assert line['e2e'] == {'value': line['value'], 'unit': B.UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    # under torchrun only rank 0 works; the other ranks exit 0 without output
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    r = subprocess.run([sys.executable, bench, '--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0'], capture_output=True,
                       text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_layout_helpers_keep_the_cpu_path_and_the_values():
    """the channels-last shortcuts of the GPU step (K8's bias split, the backbone's input pair, the extractor's one-channel
    input view) are off for CPU tensors, and the input pair holds the values of the reference's torch.cat either way"""
    import bihome_b200.functional as F
    from bihome_b200 import engine
    m = torch.nn.ConvTranspose2d(8, 8, 2, stride=2)
    assert not F.convt_bias_supported(m, torch.zeros(1, 8, 4, 4))                   # CPU tensor: the module runs as it is
    assert not F.convt_bias_supported(torch.nn.Conv2d(8, 8, 1), torch.zeros(1, 8, 4, 4))
    backbone = engine.build_model(cfg('pds-coco/zeng-bihome-lr-1e-3.yaml'), pretrained=False)[0]
    a, b = torch.rand(3, 1, 16, 16), torch.rand(3, 1, 16, 16)
    for mod in (backbone, backbone.to(memory_format=torch.channels_last)):
        x = mod._pair(a, b)
        assert torch.equal(x, torch.cat([a, b], dim=1)) and x.is_contiguous()      # CPU: the reference's NCHW cat
    # the construction the CUDA branch uses: the same values, channels-last strides, one kernel
    y = torch.stack([a[:, 0], b[:, 0]], dim=-1).permute(0, 3, 1, 2)
    assert torch.equal(y, torch.cat([a, b], dim=1)) and y.is_contiguous(memory_format=torch.channels_last) and not y.is_contiguous()
    # and the one-channel view of AuxiliaryResnet.forward: same values, strides that read as channels-last
    v = a.view(3, 16, 16, 1).permute(0, 3, 1, 2)
    assert torch.equal(v, a) and v.stride() == (256, 1, 16, 1)
