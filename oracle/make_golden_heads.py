"""Golden vectors for the heads and loss variants beside the north-star configuration -- TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_heads          (build container only: needs /root/reference)

Every array is an output of the UNMODIFIED reference modules (imported through oracle/ref_import.py) in float32 and
float64, stored with its inputs and the few hundred weights of the small networks involved:

  triplet_*      src/heads/TripletHead.py          with src/backbones/ContentAware.py's MaskPredictor / FeatureExtractor
                 (zhang-orig: DoubleLine, margin 1.0, channel-agnostic, FIX_MASK; and OneLine / 'inf' / learned masks)
  photometric    src/heads/PhotometricHead.py      (image-frame corners, crop of the warped image)
  noop_*         src/heads/NoOpHead.py             ('all_points' corner read-out; '4_points' predict_homography)
  all_points     src/data/transforms.py            HomographyNetPrep's dense perspective-field target (zeng-orig)
  perceptual_*   src/heads/PerceptualHead.py       variants of triplet_resnet_loss that are well defined in the
                 reference: one-line l1 / cosine with a numeric margin, MASK_CRD, user masks, double-line
                 channel-aware numeric margin.  (double-line 'l2' / 'cosine' and the numeric channel-agnostic margin
                 are shape-inconsistent there -- :627-628,647-649 -- and have no reference output to pin.)
"""
import os
import sys

import numpy as np
import torch

from . import ref_import
from .make_golden import OUT, TinyExtractor, head_kwargs, smooth_images

DTYPES = (('32', torch.float32), ('64', torch.float64))


class TinyContentBackbone(torch.nn.Module):
    """the two small networks of the content-aware backbone, without its ResNet-34 trunk (the TripletHead only touches
    ``backbone.feature_extractor``; masks and features arrive through the batch dict)"""

    def __init__(self, CA, fix_mask, seed=3):
        super().__init__()
        state = torch.random.get_rng_state()
        torch.manual_seed(seed)
        self.mask_predictor = CA.MaskPredictor(fix_mask=fix_mask)
        self.feature_extractor = CA.FeatureExtractor()
        torch.random.set_rng_state(state)


def _state(module, prefix):
    return {prefix + k.replace('.', '__'): v.detach().cpu().numpy() for k, v in module.state_dict().items()}


def golden_triplet(TH, CA):
    B, P = 3, 32
    g = torch.Generator().manual_seed(41)
    p1 = smooth_images(B, P, 51)
    p2 = smooth_images(B, P, 52) * 0.5 + p1 * 0.5
    d12 = ((torch.rand(B, 4, 2, generator=g) * 2 - 1) * (P / 5)).float()
    d21 = (-d12 + (torch.rand(B, 4, 2, generator=g) * 2 - 1)).float()
    d12[2] *= 2.0
    out = dict(patch_1=p1.numpy(), patch_2=p2.numpy(), delta_12=d12.numpy(), delta_21=d21.numpy())
    cases = {
        'zhang': dict(VARIANT='DoubleLine', TRIPLET_MARGIN=1.0, TRIPLET_AGGREGATION='channel-agnostic', fix_mask=True),
        'aware': dict(VARIANT='DoubleLine', TRIPLET_MARGIN=0.05, TRIPLET_AGGREGATION='channel-aware', fix_mask=False),
        'inf': dict(VARIANT='DoubleLine', TRIPLET_MARGIN='inf', TRIPLET_AGGREGATION='channel-agnostic', fix_mask=False),
        'one': dict(VARIANT='OneLine', TRIPLET_MARGIN=0.1, TRIPLET_AGGREGATION='channel-aware', fix_mask=True),
    }
    for name, case in cases.items():
        case = dict(case)
        fix_mask = case.pop('fix_mask')
        kw = dict(PATCH_KEYS=['patch_1', 'patch_2'], MASK_KEYS=['mask_1', 'mask_2'], FEATURE_KEYS=['feature_1', 'feature_2'],
                  TARGET_KEYS=['delta_hat_12', 'delta_hat_21'], LD=2, MU=0.01, PATCH_SIZE=P, **case)
        for tag, dt in DTYPES:
            bb = TinyContentBackbone(CA, fix_mask).to(dt)
            if tag == '32':
                out.update(_state(bb, name + '_w_'))
            head = TH.Model(bb, **kw)
            a = d12.to(dt).requires_grad_(True)
            b = d21.to(dt).requires_grad_(True)
            x1, x2 = p1.to(dt), p2.to(dt)
            data = {'patch_1': x1, 'patch_2': x2, 'delta_hat_12': a, 'delta_hat_21': b,
                    'mask_1': bb.mask_predictor(x1), 'mask_2': bb.mask_predictor(x2),
                    'feature_1': bb.feature_extractor(x1), 'feature_2': bb.feature_extractor(x2)}
            loss, _, delta_hat = head(data)
            params = [p for p in bb.parameters()]
            wanted = [a, b] if 'Double' in case['VARIANT'] else [a]
            grads = torch.autograd.grad(loss, wanted + params, allow_unused=True)
            out.update({'%s_loss%s' % (name, tag): loss.detach().numpy(), '%s_g12_%s' % (name, tag): grads[0].numpy()})
            if len(wanted) == 2:
                out['%s_g21_%s' % (name, tag)] = grads[1].numpy()
            gp = [x for x in grads[len(wanted):]]
            gnorm = torch.sqrt(sum((x.double() ** 2).sum() for x in gp if x is not None))
            out['%s_gparam_norm%s' % (name, tag)] = gnorm.numpy()
            # one parameter gradient in full: the first convolution of the feature extractor
            names = [n for n, _ in bb.named_parameters()]
            out['%s_gfe_conv1_%s' % (name, tag)] = gp[names.index('feature_extractor.layer1.0.weight')].numpy()
        print('triplet', name, out[name + '_loss32'], out[name + '_loss64'])
    np.savez_compressed(os.path.join(OUT, 'triplet_head_P32.npz'), **out)


def golden_photometric_noop(PHO, NO):
    B, P, Hi, Wi = 3, 32, 60, 80
    g = torch.Generator().manual_seed(43)
    lo = torch.rand(B, 1, 9, 12, generator=g)
    image = torch.nn.functional.interpolate(lo, size=(Hi, Wi), mode='bicubic', align_corners=True).float()
    cx = torch.tensor([8, 20, 40])
    cy = torch.tensor([8, 14, 20])
    base = torch.tensor([[0, 0], [P, 0], [P, P], [0, P]])
    corners = (base.unsqueeze(0) + torch.stack([cx, cy], -1).unsqueeze(1)).float()
    delta_hat = ((torch.rand(B, 4, 2, generator=g) * 2 - 1) * 6).float()
    delta_gt = ((torch.rand(B, 4, 2, generator=g) * 2 - 1) * 6).float()
    patch_gt = smooth_images(B, P, 61)
    g_out = torch.randn(B, 1, P, P, generator=g).float()
    out = dict(image=image.numpy(), corners=corners.numpy(), delta_hat=delta_hat.numpy(), delta_gt=delta_gt.numpy(),
               patch_gt=patch_gt.numpy(), g_out=g_out.numpy())
    keys = ['patch_2', 'image_1', 'delta', 'delta_hat_12']
    for tag, dt in DTYPES:
        head = PHO.Model(None, LEARNING_KEYS=keys)
        d = delta_hat.to(dt).requires_grad_(True)
        data = {'patch_2': patch_gt.to(dt), 'image_1': image.to(dt), 'delta': delta_gt.to(dt), 'delta_hat_12': d,
                'corners': corners.to(dt)}
        gt, hat, dg, dh = head(data)
        gd, = torch.autograd.grad((hat * g_out.to(dt)).sum(), d)
        _, hom = head.predict_homography(data)
        out.update({'photo_patch_hat' + tag: hat.detach().numpy(), 'photo_gdelta' + tag: gd.numpy(),
                    'photo_H' + tag: hom.detach().numpy()})
        assert gt is data['patch_2'] and dg is data['delta'] and dh is d
    # NoOpHead: all_points corner read-out and the cv2 RANSAC post-processing; 4_points predict_homography
    field = (torch.randn(B, 2, P, P, generator=g) * 3).float()
    target = (torch.randn(B, 2, P, P, generator=g) * 3).float()
    head = NO.Model(None, TARGET_GEN='all_points', LEARNING_KEYS=['target', 'pf_hat_12', 'delta', 'pf_hat_12'])
    ret = head({'target': target, 'pf_hat_12': field, 'delta': delta_gt})
    assert ret[0] is target and ret[1] is field and ret[2] is delta_gt
    out.update(noop_field=field.numpy(), noop_target=target.numpy(), noop_delta_hat=ret[3].numpy())
    # a consistent field (exact homography flow + small noise) for the RANSAC fit
    U = ref_import.load('src.data.utils')
    c0 = base.unsqueeze(0).repeat(B, 1, 1).double()
    Hgt = U.four_point_to_homography(c0, delta_gt.double())
    ys, xs = torch.meshgrid(torch.arange(P, dtype=torch.float64), torch.arange(P, dtype=torch.float64), indexing='ij')
    pts = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(P * P, dtype=torch.float64)], 0)
    q = Hgt @ pts
    flow = (q[:, :2] / q[:, 2:3] - pts[:2]).reshape(B, 2, P, P)
    flow = (flow + 0.05 * torch.randn(B, 2, P, P, generator=g, dtype=torch.float64)).float()
    np.random.seed(0)
    import cv2
    cv2.setRNGSeed(0)
    pd, ph = head.predict_homography({'pf_hat_12': flow})
    out.update(noop_flow=flow.numpy(), noop_post_delta=np.asarray(pd), noop_post_H=np.asarray(ph))
    head4 = NO.Model(None, TARGET_GEN='4_points', LEARNING_KEYS=['delta', 'delta_hat_12', 'delta', 'delta_hat_12'])
    dh, hom = head4.predict_homography({'delta_hat_12': delta_hat.double(), 'corners': corners.double()})
    out.update(noop_H64=hom.detach().numpy())
    np.savez_compressed(os.path.join(OUT, 'photometric_noop_P32.npz'), **out)
    print('photometric/noop', {k: v.shape for k, v in out.items() if k.startswith(('photo', 'noop'))})


def golden_perceptual_variants(PH, CA=None):
    B, P = 3, 64
    base_kw = head_kwargs('config/s-coco/detone-bihome-lr-5e-3.yaml', P)
    g = torch.Generator().manual_seed(45)
    p1 = smooth_images(B, P, 71)
    p2 = smooth_images(B, P, 72) * 0.5 + p1 * 0.5
    d12 = ((torch.rand(B, 4, 2, generator=g) * 2 - 1) * (P / 5)).float()
    d21 = (-d12 + (torch.rand(B, 4, 2, generator=g) * 2 - 1) * 2).float()
    d12[2] *= 2.5
    d21[2] *= 2.5
    um1 = (torch.rand(B, 1, P, P, generator=g) > 0.3).float()
    um2 = (torch.rand(B, 1, P, P, generator=g) > 0.3).float()
    out = dict(patch_1=p1.numpy(), patch_2=p2.numpy(), delta_12=d12.numpy(), delta_21=d21.numpy(),
               mask_1=um1.numpy(), mask_2=um2.numpy())
    ext = TinyExtractor()
    out.update(ext_w=ext.conv.weight.detach().numpy(), ext_b=ext.conv.bias.detach().numpy())
    cases = {
        'one_l1': dict(TRIPLET_LOSS='one-line', TRIPLET_MARGIN=1.0, TRIPLET_DISTANCE='l1'),
        'one_cos_crd': dict(TRIPLET_LOSS='one-line', TRIPLET_MARGIN=0.2, TRIPLET_DISTANCE='cosine', MASK_CRD=True),
        'one_l1_masks': dict(TRIPLET_LOSS='one-line', TRIPLET_MARGIN=0.5, TRIPLET_DISTANCE='l1', MASK_KEYS=['mask_1', 'mask_2']),
        'double_aware_margin': dict(TRIPLET_LOSS='double-line', TRIPLET_MARGIN=0.05, TRIPLET_AGGREGATION='channel-aware'),
        'double_aware_inf_masks': dict(TRIPLET_LOSS='double-line', TRIPLET_MARGIN='inf', TRIPLET_AGGREGATION='channel-aware',
                                       MASK_KEYS=['mask_1', 'mask_2']),
        # 'dual': the content-aware backbone's own feature extractor adds a second, full-resolution triplet term (:407-441)
        'double_dual': dict(TRIPLET_LOSS='double-line-dual', TRIPLET_MARGIN='inf', TRIPLET_AGGREGATION='channel-agnostic'),
    }
    out.update(_state(TinyContentBackbone(CA, True), 'dual_w_'))
    for name, over in cases.items():
        kw = dict(base_kw)
        kw.update(over)
        double = 'double' in kw['TRIPLET_LOSS']
        for tag, dt in DTYPES:
            backbone = TinyContentBackbone(CA, True).to(dt) if 'dual' in name else torch.nn.Identity()
            model = PH.Model(backbone=backbone, **kw)
            model.auxiliary_resnet = TinyExtractor().to(dt)
            model.auxiliary_resnet.with_projection_head = None
            a = d12.to(dt).requires_grad_(True)
            b = d21.to(dt).requires_grad_(True)
            data = {'patch_1': p1.to(dt), 'patch_2': p2.to(dt), 'delta_hat_12': a, 'delta_hat_21': b,
                    'mask_1': um1.to(dt), 'mask_2': um2.to(dt)}
            loss, _, _ = model(data)
            grads = torch.autograd.grad(loss, (a, b) if double else (a,))
            out.update({'%s_loss%s' % (name, tag): loss.detach().numpy(), '%s_g12_%s' % (name, tag): grads[0].numpy()})
            if double:
                out['%s_g21_%s' % (name, tag)] = grads[1].numpy()
        print('perceptual', name, out[name + '_loss32'], out[name + '_loss64'])
    np.savez_compressed(os.path.join(OUT, 'perceptual_variants_P64.npz'), **out)


def golden_all_points(T):
    """HomographyNetPrep(target_gen='all_points') from the shipped zeng-orig YAML: the dense perspective-field target
    (src/data/transforms.py:634-687) with the corners and offsets it was made from"""
    import yaml
    from . import pairgen
    with open(os.path.join(ref_import.REFERENCE_ROOT, 'config/pds-coco/zeng-orig-lr-1e-3.yaml')) as f:
        cfg = yaml.full_load(f)
    seed = cfg['DATA']['SAMPLER']['TRAIN_SEED']
    tfs = []
    for t in cfg['DATA']['TRANSFORMS']:                 # train.py:111-120: the seed is appended to every ctor
        t_name = list(t.keys())[0]
        tfs.append(getattr(T, t_name)(*(t[t_name] + [seed])))
    targets, deltas, corners = [], [], []
    for i in range(2):
        data = ([pairgen.synthetic_image(i)], None)
        for t in tfs:
            data = t(data)
        assert data['target'].dtype == torch.float64 and tuple(data['target'].shape) == (2, 128, 128)
        targets.append(data['target'].float().numpy())   # train.py:308-309 casts the batch to float
        deltas.append(data['delta'].numpy())
        corners.append(data['corners'].numpy())
    out = dict(target=np.stack(targets), delta=np.stack(deltas), corners=np.stack(corners))
    np.savez_compressed(os.path.join(OUT, 'all_points_target.npz'), **out)
    print('all_points', {k: (v.shape, v.dtype) for k, v in out.items()})


def main():
    if not ref_import.available():
        sys.exit('needs the reference tree at ' + ref_import.REFERENCE_ROOT)
    import warnings
    warnings.filterwarnings('ignore')
    golden_triplet(ref_import.load('src.heads.TripletHead'), ref_import.load('src.backbones.ContentAware'))
    golden_photometric_noop(ref_import.load('src.heads.PhotometricHead'), ref_import.load('src.heads.NoOpHead'))
    golden_perceptual_variants(ref_import.load('src.heads.PerceptualHead'), ref_import.load('src.backbones.ContentAware'))
    golden_all_points(ref_import.load('src.data.transforms'))


if __name__ == '__main__':
    main()
