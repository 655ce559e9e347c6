"""Host logic of the head / backbone mirrors against the unmodified reference, on CPU.

The reference's outputs come from tests/golden/*.npz (oracle/make_golden_heads.py ran the reference modules) and,
where the whole network is too large for a fixture, from the reference tree itself when it is present.  The CUDA ops
are replaced by the oracle's closed forms (tests/cpu_kernels.py) -- this file checks the wiring around the kernels;
tests/test_gpu_zz_head_mirrors.py repeats the golden comparisons on the real kernels.
"""
import os

import numpy as np
import pytest
import torch

import cpu_kernels
from conftest import rel_l2

DT = torch.float64


def t64(a):
    return torch.as_tensor(np.asarray(a)).to(DT)


@pytest.fixture
def F(monkeypatch):
    return cpu_kernels.install(monkeypatch)


def tiny_extractor():
    from oracle.make_golden import TinyExtractor
    return TinyExtractor().to(DT)


def perceptual_kwargs(P, **over):
    import yaml
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'detone-bihome.yaml')) as f:
        kw = dict(yaml.full_load(f)['MODEL']['HEAD'])
    kw.update(PATCH_SIZE=P, AUXILIARY_RESNET_PRETRAINED=False, AUXILIARY_RESNET='resnet18')
    kw.update(over)
    return kw


PERCEPTUAL_CASES = {
    'one_l1': dict(TRIPLET_LOSS='one-line', TRIPLET_MARGIN=1.0, TRIPLET_DISTANCE='l1'),
    'one_cos_crd': dict(TRIPLET_LOSS='one-line', TRIPLET_MARGIN=0.2, TRIPLET_DISTANCE='cosine', MASK_CRD=True),
    'one_l1_masks': dict(TRIPLET_LOSS='one-line', TRIPLET_MARGIN=0.5, TRIPLET_DISTANCE='l1', MASK_KEYS=['mask_1', 'mask_2']),
    'double_aware_margin': dict(TRIPLET_LOSS='double-line', TRIPLET_MARGIN=0.05, TRIPLET_AGGREGATION='channel-aware'),
    'double_aware_inf_masks': dict(TRIPLET_LOSS='double-line', TRIPLET_MARGIN='inf', TRIPLET_AGGREGATION='channel-aware',
                                   MASK_KEYS=['mask_1', 'mask_2']),
    'double_dual': dict(TRIPLET_LOSS='double-line-dual', TRIPLET_MARGIN='inf', TRIPLET_AGGREGATION='channel-agnostic'),
}


def perceptual_backbone(g, name, dtype=torch.float64, device='cpu'):
    """Identity, except for the 'dual' variants, whose extra term runs the content-aware backbone's feature extractor"""
    if 'dual' not in name:
        return torch.nn.Identity()
    return content_backbone(g, 'dual', True).to(dtype).to(device)


@pytest.mark.parametrize('name', sorted(PERCEPTUAL_CASES))
def test_perceptual_variants_match_reference(F, golden, name):
    """reference PerceptualHead.triplet_resnet_loss variants (:320-714) -- values and gradients w.r.t. the offsets"""
    from bihome_b200.heads import PerceptualHead as PH
    g = golden('perceptual_variants_P64.npz')
    model = PH.Model(backbone=perceptual_backbone(g, name), **perceptual_kwargs(64, **PERCEPTUAL_CASES[name]))
    model.auxiliary_resnet = tiny_extractor()
    a, b = t64(g['delta_12']).requires_grad_(True), t64(g['delta_21']).requires_grad_(True)
    data = {'patch_1': t64(g['patch_1']), 'patch_2': t64(g['patch_2']), 'delta_hat_12': a, 'delta_hat_21': b,
            'mask_1': t64(g['mask_1']), 'mask_2': t64(g['mask_2'])}
    loss, delta_gt, delta_hat = model(data)
    assert delta_gt is None and delta_hat.shape == (3, 4, 2)
    # 1e-8: the coverage masks are analytic here, the reference warps a plane of ones over a sampling grid that is born in
    # float32 even in its float64 evaluation (1e-8 px of coordinate noise on every border pixel of the full-size masks)
    assert abs(loss.item() - float(g[name + '_loss64'])) < 1e-8 * abs(float(g[name + '_loss64']))
    double = 'double' in name
    grads = torch.autograd.grad(loss, (a, b) if double else (a,))
    # 1e-6: the same 1e-8 of mask noise, amplified by the cancellations of the DLT adjoint (north_star allows 1e-4 there)
    assert rel_l2(grads[0].numpy(), g[name + '_g12_64']) < 1e-6
    if double:
        assert rel_l2(grads[1].numpy(), g[name + '_g21_64']) < 1e-6


def content_backbone(g, name, fix_mask):
    from bihome_b200.backbones import ContentAware as CA

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.mask_predictor = CA.MaskPredictor(fix_mask=fix_mask)
            self.feature_extractor = CA.FeatureExtractor()
    bb = Tiny()
    prefix = name + '_w_'
    state = {k[len(prefix):].replace('__', '.'): torch.as_tensor(v) for k, v in g.items() if k.startswith(prefix)}
    bb.load_state_dict(state)           # strict: the mirror's parameter names are the reference's
    return bb


TRIPLET_CASES = {
    'zhang': dict(VARIANT='DoubleLine', TRIPLET_MARGIN=1.0, TRIPLET_AGGREGATION='channel-agnostic', fix_mask=True),
    'aware': dict(VARIANT='DoubleLine', TRIPLET_MARGIN=0.05, TRIPLET_AGGREGATION='channel-aware', fix_mask=False),
    'inf': dict(VARIANT='DoubleLine', TRIPLET_MARGIN='inf', TRIPLET_AGGREGATION='channel-agnostic', fix_mask=False),
    'one': dict(VARIANT='OneLine', TRIPLET_MARGIN=0.1, TRIPLET_AGGREGATION='channel-aware', fix_mask=True),
}


def triplet_forward(g, name, dtype, device):
    from bihome_b200.heads import TripletHead as TH
    case = dict(TRIPLET_CASES[name])
    bb = content_backbone(g, name, case.pop('fix_mask')).to(dtype).to(device)
    head = TH.Model(bb, PATCH_KEYS=['patch_1', 'patch_2'], MASK_KEYS=['mask_1', 'mask_2'], FEATURE_KEYS=['feature_1', 'feature_2'],
                    TARGET_KEYS=['delta_hat_12', 'delta_hat_21'], LD=2, MU=0.01, PATCH_SIZE=32, **case)
    to = lambda k: torch.as_tensor(g[k]).to(dtype).to(device)
    a, b = to('delta_12').requires_grad_(True), to('delta_21').requires_grad_(True)
    x1, x2 = to('patch_1'), to('patch_2')
    data = {'patch_1': x1, 'patch_2': x2, 'delta_hat_12': a, 'delta_hat_21': b,
            'mask_1': bb.mask_predictor(x1), 'mask_2': bb.mask_predictor(x2),
            'feature_1': bb.feature_extractor(x1), 'feature_2': bb.feature_extractor(x2)}
    loss, delta_gt, delta_hat = head(data)
    assert delta_gt is None and delta_hat is a
    double = 'Double' in case['VARIANT']
    names = [n for n, _ in bb.named_parameters()]
    grads = torch.autograd.grad(loss, ([a, b] if double else [a]) + list(bb.parameters()), allow_unused=True)
    gp = grads[2 if double else 1:]
    gnorm = float(torch.sqrt(sum((x.double() ** 2).sum() for x in gp if x is not None)))
    conv1 = gp[names.index('feature_extractor.layer1.0.weight')]
    return loss, grads[0], (grads[1] if double else None), gnorm, conv1


@pytest.mark.parametrize('name', sorted(TRIPLET_CASES))
def test_triplet_head_matches_reference(F, golden, name):
    """reference TripletHead.forward (:44-192), including the B-fold broadcast of the numeric channel-agnostic margin
    ('zhang' = the shipped zhang-orig configuration) and the gradients into the backbone's small networks"""
    g = golden('triplet_head_P32.npz')
    loss, g12, g21, gnorm, conv1 = triplet_forward(g, name, DT, 'cpu')
    ref = float(g[name + '_loss64'])
    # FIX_MASK: the mirror takes the exact coverage mask where the reference warps a ones image through its
    # float32-born sampling grid (1e-8 px of coordinate noise even in float64) -- like for like otherwise
    tol = 1e-6 if TRIPLET_CASES[name]['fix_mask'] else 1e-8
    assert abs(loss.item() - ref) < 0.1 * tol * abs(ref)
    assert rel_l2(g12.numpy(), g[name + '_g12_64']) < tol
    if g21 is not None:
        assert rel_l2(g21.numpy(), g[name + '_g21_64']) < tol
    assert abs(gnorm - float(g[name + '_gparam_norm64'])) < tol * float(g[name + '_gparam_norm64'])
    assert rel_l2(conv1.numpy(), g[name + '_gfe_conv1_64']) < tol


def test_photometric_head_matches_reference(F, golden):
    """reference PhotometricHead.forward (:19-46): warp of the whole image + per-sample crop == one P x P render"""
    from bihome_b200.heads import PhotometricHead as PHO
    g = golden('photometric_noop_P32.npz')
    head = PHO.Model(None, LEARNING_KEYS=['patch_2', 'image_1', 'delta', 'delta_hat_12'])
    d = t64(g['delta_hat']).requires_grad_(True)
    data = {'patch_2': t64(g['patch_gt']), 'image_1': t64(g['image']), 'delta': t64(g['delta_gt']), 'delta_hat_12': d,
            'corners': t64(g['corners'])}
    gt, hat, dg, dh = head(data)
    assert gt is data['patch_2'] and dg is data['delta'] and dh is d
    # exact closed form here against the reference's float32-born grid: 1e-7, not 1e-12
    assert rel_l2(hat.detach().numpy(), g['photo_patch_hat64']) < 1e-6
    gd, = torch.autograd.grad((hat * t64(g['g_out'])).sum(), d)
    assert rel_l2(gd.numpy(), g['photo_gdelta64']) < 1e-5
    dh2, hom = head.predict_homography(data)
    assert dh2 is d and rel_l2(hom.detach().numpy(), g['photo_H64']) < 1e-10


def test_noop_head_matches_reference(F, golden):
    """reference NoOpHead (:22-109): pass-through, corner read-out of a dense field, both predict_homography modes"""
    from bihome_b200.heads import NoOpHead as NO
    g = golden('photometric_noop_P32.npz')
    head = NO.Model(None, TARGET_GEN='all_points', LEARNING_KEYS=['target', 'pf_hat_12', 'delta', 'pf_hat_12'])
    field = torch.as_tensor(g['noop_field']).requires_grad_(True)
    target, delta = torch.as_tensor(g['noop_target']), torch.as_tensor(g['delta_gt'])
    ret = head({'target': target, 'pf_hat_12': field, 'delta': delta})
    assert ret[0] is target and ret[1] is field and ret[2] is delta
    assert np.array_equal(ret[3].detach().numpy(), g['noop_delta_hat'])
    ret[3].sum().backward()
    assert float(field.grad.sum()) == 3 * 8        # the read-out is differentiable: one unit per corner and axis
    pd, ph = head.predict_homography({'pf_hat_12': torch.as_tensor(g['noop_flow'])})
    assert pd.shape == (3, 4, 2) and ph.shape == (3, 3, 3)
    assert np.abs(pd - g['noop_post_delta']).max() < 1e-3
    head4 = NO.Model(None, TARGET_GEN='4_points', LEARNING_KEYS=['delta', 'delta_hat_12', 'delta', 'delta_hat_12'])
    d = t64(g['delta_hat'])
    out = head4({'delta': delta, 'delta_hat_12': d})
    assert out[0] is delta and out[1] is d and out[2] is delta and out[3] is d
    dh, hom = head4.predict_homography({'delta_hat_12': d, 'corners': t64(g['corners'])})
    assert dh is d and rel_l2(hom.numpy(), g['noop_H64']) < 1e-10
    with pytest.raises(AssertionError):
        head4.predict_homography({'delta_hat_12': d})


@pytest.mark.parametrize('version,n', [('one-line', 4), ('double-line', 1), ('', 3)])
def test_dsac_heads_match_the_reference_module(F, version, n):
    """perspective-field heads with several hypotheses (DSAC scoring, score-weighted loss and offsets) and the feature-
    returning mode (TRIPLET_LOSS '' -> external nn loss), against the reference module itself with the same torch seed"""
    PH_ref = _reference().load('src.heads.PerceptualHead')
    from bihome_b200.heads import PerceptualHead as PH
    P, B, M = 32, 2, 24
    kw = perceptual_kwargs(P, TRIPLET_LOSS=version, TRIPLET_MARGIN=1.0 if version == 'one-line' else 'inf', DELTA_HAT_KEYS=[],
                           PF_KEYS=['pf_hat_12', 'pf_hat_21'], RANSAC_HYPOTHESIS_NO=n, POINTS_PER_HYPOTHESIS=M)
    gen = torch.Generator().manual_seed(13)
    ys, xs = torch.meshgrid(torch.arange(P, dtype=DT), torch.arange(P, dtype=DT), indexing='ij')
    flow = torch.stack([0.05 * xs - 0.02 * ys + 1.5, 0.03 * ys + 0.01 * xs - 2.0]).unsqueeze(0).repeat(B, 1, 1, 1)
    pf12 = flow + 0.2 * torch.randn(B, 2, P, P, generator=gen, dtype=DT)
    pf21 = -flow + 0.2 * torch.randn(B, 2, P, P, generator=gen, dtype=DT)
    p1, p2 = torch.rand(B, 1, P, P, generator=gen, dtype=DT), torch.rand(B, 1, P, P, generator=gen, dtype=DT)
    res = []
    for mod in (PH_ref, PH):
        model = mod.Model(backbone=torch.nn.Identity(), **kw)
        model.auxiliary_resnet = tiny_extractor()
        model.auxiliary_resnet.with_projection_head = None
        if True:
            # both cache their coordinate field in float32 (reference :139): pre-populate the caches in float64 like the
            # golden generator does, so that both sides evaluate the same float64 function
            with torch.no_grad():
                _, cf, fp = model.forward_map_field(pf12.float(), None, None)
            model.coordinate_field_12 = model.coordinate_field_21 = cf.double()
            model.four_points_12 = model.four_points_21 = fp.double()
        a, b = pf12.clone().requires_grad_(True), pf21.clone().requires_grad_(True)
        torch.manual_seed(5)
        out = model({'patch_1': p1, 'patch_2': p2, 'pf_hat_12': a, 'pf_hat_21': b})
        if version == '':
            f2, f1w, _, delta_hat = out
            loss = (f2 - f1w).abs().sum()
        else:
            loss, _, delta_hat = out
        grads = torch.autograd.grad(loss, (a, b), allow_unused=True)
        res.append((loss.detach(), delta_hat.detach(), grads))
    (lr, dr, gr), (lo, do, go) = res
    assert abs(float(lo) - float(lr)) < 1e-8 * abs(float(lr))
    assert rel_l2(do.numpy(), dr.numpy()) < 1e-8
    for x, y in zip(go, gr):
        assert (x is None) == (y is None)
        if y is not None:
            assert rel_l2(x.numpy(), y.numpy()) < 1e-6


def test_score_cnn_scoring_matches_the_reference_module(F):
    """SCORING_METHOD 'score_cnn' (reference ransac_utils.py:10-23,113-121): same parameter names (state dict loads
    strictly), same hypotheses and scores for the same draw"""
    ref = _reference().load('src.heads.ransac_utils')
    from bihome_b200.heads import ransac_utils as ours
    kw = dict(SCORING_METHOD='score_cnn', SCORE_CNN_PRETRAINED=False)
    torch.manual_seed(3)
    a = ref.DSACSoftmax(**kw).double().eval()
    b = ours.DSACSoftmax(**kw).double().eval()
    b.load_state_dict(a.state_dict())          # strict
    B, side, n, M = 2, 16, 3, 12
    gen = torch.Generator().manual_seed(4)
    ys, xs = torch.meshgrid(torch.arange(side, dtype=DT), torch.arange(side, dtype=DT), indexing='ij')
    p1 = torch.stack((xs.reshape(-1), ys.reshape(-1)), dim=-1).unsqueeze(0).repeat(B, 1, 1)
    p2 = p1 * 1.05 + 0.7 + 0.1 * torch.randn(B, side * side, 2, generator=gen, dtype=DT)
    torch.manual_seed(9)
    Hr, sr = a(p1, p2, points_per_hypothesis=M, hypothesis_no=n)
    torch.manual_seed(9)
    Ho, so = b(p1, p2, points_per_hypothesis=M, hypothesis_no=n)
    assert rel_l2(Ho.detach().numpy(), Hr.detach().numpy()) < 1e-8
    assert rel_l2(so.detach().numpy(), sr.detach().numpy()) < 1e-8 and abs(float(so.detach().sum()) - B) < 1e-9


@pytest.mark.parametrize('strategy,version', [('upsample-patch-2x', 'double-line'), ('upsample-patch-4x', 'one-line'),
                                              ('upsample-patch-2x', 'one-line')])
def test_upsampling_strategies_match_the_reference_module(F, strategy, version):
    """SAMPLING_STRATEGY 'upsample-patch-2x/4x' (reference :353-396): features of the bilinearly enlarged patches, masks
    pooled by what is left of the extractor's stride (2 or 1)"""
    PH_ref = _reference().load('src.heads.PerceptualHead')
    from bihome_b200.heads import PerceptualHead as PH
    P, B = 32, 3
    kw = perceptual_kwargs(P, SAMPLING_STRATEGY=strategy, TRIPLET_LOSS=version,
                           TRIPLET_MARGIN='inf' if version == 'double-line' else 0.7)
    gen = torch.Generator().manual_seed(17)
    lo = torch.rand(B, 1, 9, 9, generator=gen, dtype=DT)
    p1 = torch.nn.functional.interpolate(lo, size=(P, P), mode='bicubic', align_corners=True)
    p2 = 0.5 * p1 + 0.5 * torch.rand(B, 1, P, P, generator=gen, dtype=DT)
    d12 = (torch.rand(B, 4, 2, generator=gen, dtype=DT) * 2 - 1) * 6
    d21 = -d12 + torch.rand(B, 4, 2, generator=gen, dtype=DT)
    res = []
    for mod in (PH_ref, PH):
        model = mod.Model(backbone=torch.nn.Identity(), **kw)
        model.auxiliary_resnet = tiny_extractor()
        model.auxiliary_resnet.with_projection_head = None
        a, b = d12.clone().requires_grad_(True), d21.clone().requires_grad_(True)
        loss, _, _ = model({'patch_1': p1, 'patch_2': p2, 'delta_hat_12': a, 'delta_hat_21': b})
        res.append((loss.detach(), torch.autograd.grad(loss, (a, b), allow_unused=True)))
    (lr, gr), (lo_, go) = res
    # FIX: the fused path takes the exact coverage mask, the reference warps ones through its float32-born grid
    assert abs(float(lo_) - float(lr)) < 1e-6 * abs(float(lr))
    for x, y in zip(go, gr):
        assert (x is None) == (y is None)
        if y is not None:
            assert rel_l2(x.numpy(), y.numpy()) < 1e-5


@pytest.mark.parametrize('layer,projection', [(1, None), (2, None), (1, [[64, 32], [32, 16]])])
def test_auxiliary_resnet_matches_the_reference_module(layer, projection):
    """the frozen feature extractor (reference :15-76): parameter names, the 1 -> 3 channel repeat folded into conv1,
    deeper output layers, the optional projection head; train-mode BatchNorm as in the reference"""
    PH_ref = _reference().load('src.heads.PerceptualHead')
    from bihome_b200.heads import PerceptualHead as PH
    kw = dict(AUXILIARY_RESNET='resnet18', AUXILIARY_RESNET_OUTPUT_LAYER=layer, AUXILIARY_RESNET_FREEZE=True)
    if projection is not None:
        kw['WITH_PROJECTION_HEAD'] = projection
    torch.manual_seed(4)
    ref = PH_ref.AuxiliaryResnet(**kw).double()
    ours = PH.AuxiliaryResnet(AUXILIARY_RESNET_PRETRAINED=False, **kw).double()
    assert list(ref.state_dict().keys()) == list(ours.state_dict().keys())
    ours.load_state_dict(ref.state_dict())
    assert not any(p.requires_grad for p in ours.resnet.parameters())
    gen = torch.Generator().manual_seed(8)
    for channels in (1, 3):
        x = torch.rand(2, channels, 64, 64, generator=gen, dtype=DT)
        for mode in ('train', 'eval'):
            getattr(ref, mode)()
            getattr(ours, mode)()
            a, b = ref(x), ours(x)
            assert a.shape == b.shape and rel_l2(b.detach().numpy(), a.detach().numpy()) < 1e-12, (channels, mode)
    for (name, p), (_, q) in zip(ours.named_buffers(), ref.named_buffers()):
        assert torch.allclose(p, q, rtol=1e-12, atol=1e-14), name          # the running statistics moved together


# ------------------------------------------------------------------------------------------------
# whole networks: compared with the reference tree itself (build container only)
# ------------------------------------------------------------------------------------------------
def _reference():
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip('reference tree not present')
    return ref_import


@pytest.mark.parametrize('module,kwargs', [
    ('ContentAware', dict(VARIANT='DoubleLine', IMAGE_SIZE=128, PRETRAINED_RESNET=False, PATCH_KEYS=['patch_1', 'patch_2'],
                          MASK_KEYS=['mask_1', 'mask_2'], FIX_MASK=True, FEATURE_KEYS=['feature_1', 'feature_2'],
                          TARGET_KEYS=['delta_hat_12', 'delta_hat_21'])),
    ('ContentAware', dict(VARIANT='OneLine', IMAGE_SIZE=128, PRETRAINED_RESNET=False, PATCH_KEYS=['patch_1', 'patch_2'],
                          MASK_KEYS=['mask_1', 'mask_2'], FIX_MASK=False, MASK_NORMALIZATION_STRENGTH=0.5,
                          FEATURE_KEYS=['feature_1', 'feature_2'], TARGET_KEYS=['delta_hat_12', 'delta_hat_21'])),
    ('HomographyNet', dict(IMAGE_SIZE=128, PATCH_KEYS=['patch_1', 'patch_2'], TARGET_KEYS=['delta_hat_12'])),
    ('ResNet34', dict(VARIANT='DoubleLine', IMAGE_SIZE=128, PRETRAINED_RESNET=False, PATCH_KEYS=['patch_1', 'patch_2'],
                      TARGET_KEYS=['delta_hat_12', 'delta_hat_21'])),
    ('Rethinking', dict(VARIANT='DoubleLine', IMAGE_SIZE=128, RESNET_BLOCK='ResNet34', PRETRAINED_RESNET=False,
                        PATCH_KEYS=['patch_1', 'patch_2'], TARGET_KEYS=['pf_hat_12', 'pf_hat_21'])),
    ('Rethinking', dict(VARIANT='OneLine', IMAGE_SIZE=128, RESNET_BLOCK='ResNet50', PRETRAINED_RESNET=False,
                        PATCH_KEYS=['patch_1', 'patch_2'], TARGET_KEYS=['pf_hat_12'])),
])
def test_backbone_mirrors_match_reference_modules(module, kwargs):
    """same parameter names and shapes, same initialisation stream, same outputs for the same weights and inputs"""
    import importlib
    ref_mod = _reference().load('src.backbones.' + module)
    torch.manual_seed(5)
    ref = ref_mod.Model(**kwargs)
    torch.manual_seed(5)
    ours = importlib.import_module('bihome_b200.backbones.' + module).Model(**kwargs)
    rs, os_ = ref.state_dict(), ours.state_dict()
    assert list(rs.keys()) == list(os_.keys())
    assert all(rs[k].shape == os_[k].shape for k in rs)
    ours.load_state_dict(rs)
    g = torch.Generator().manual_seed(1)
    n = 2 if module == 'Rethinking' else 4
    p1, p2 = torch.rand(n, 1, 128, 128, generator=g), torch.rand(n, 1, 128, 128, generator=g)
    for mode in ('train', 'eval'):
        getattr(ref, mode)()
        getattr(ours, mode)()
        a = ref({'patch_1': p1, 'patch_2': p2})
        b = ours({'patch_1': p1, 'patch_2': p2})
        assert sorted(a.keys()) == sorted(b.keys())
        for k in a:
            assert torch.allclose(a[k], b[k], rtol=1e-4, atol=2e-5), (mode, k, float((a[k] - b[k]).abs().max()))
    a = ref.predict_homography({'patch_1': p1, 'patch_2': p2})
    b = ours.predict_homography({'patch_1': p1, 'patch_2': p2})
    assert sorted(a.keys()) == sorted(b.keys())


def test_every_shipped_config_builds_with_reference_parameter_names():
    """all 14 YAML files of the reference go through engine.build_model unchanged, and the state dict of
    nn.Sequential(backbone, head) has the reference's keys (checkpoints are interchangeable)"""
    import glob
    ref_import = _reference()
    from bihome_b200 import engine
    paths = sorted(glob.glob(os.path.join(ref_import.REFERENCE_ROOT, 'config', '*', '*.yaml')))
    assert len(paths) == 14
    seen = {}
    for path in paths:
        cfg = engine.load_config(path)
        combo = (cfg['MODEL']['BACKBONE']['NAME'], cfg['MODEL']['HEAD']['NAME'],
                 str(sorted(cfg['MODEL']['BACKBONE'].items())), str(sorted(cfg['MODEL']['HEAD'].items())))
        if combo in seen:
            continue
        model = engine.build_model(cfg, pretrained=False)
        bcfg, hcfg = dict(cfg['MODEL']['BACKBONE']), dict(cfg['MODEL']['HEAD'])
        bcfg['PRETRAINED_RESNET'] = False
        rb = ref_import.load('src.backbones.' + bcfg['NAME']).Model(**bcfg)
        rh = ref_import.load('src.heads.' + hcfg['NAME']).Model(rb, **hcfg)
        ref = torch.nn.Sequential(rb, rh)
        ours_keys, ref_keys = set(model.state_dict().keys()), set(ref.state_dict().keys())
        assert ours_keys == ref_keys, (path, sorted(ours_keys ^ ref_keys)[:6])
        seen[combo] = path
    assert len(seen) >= 7
