"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules -- TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference):

    python -m oracle.make_golden

The reference is imported through oracle/ref_import.py (kornia 0.5.0 restatement, matplotlib stub,
no-network torchvision).  Every array that ends up in a fixture is an output of the reference's own
code (``src.heads.PerceptualHead.Model``, ``src.heads.ransac_utils.DSACSoftmax``,
``src.data.transforms.*``) evaluated in float32 (like for like) and float64 (ground truth), on
inputs that are stored next to it.  The fixtures are small on purpose (P=64 / P=32 patches).
"""
import os
import sys

import numpy as np
import torch
import yaml

from . import pairgen, ref_import

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


class TinyExtractor(torch.nn.Module):
    """Stand-in for the frozen ResNet stem (stride 4, ReLU features): 1 -> C channels.

    Same shape contract as AuxiliaryResnet (PerceptualHead.py:50-76) at OUTPUT_LAYER 1, but with a
    few hundred weights so they fit in the fixture.
    """

    def __init__(self, channels=8, seed=7):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.conv = torch.nn.Conv2d(1, channels, kernel_size=7, stride=2, padding=3, bias=True)
        with torch.no_grad():
            self.conv.weight.copy_(torch.randn(self.conv.weight.shape, generator=g) * 0.2)
            self.conv.bias.copy_(torch.randn(self.conv.bias.shape, generator=g) * 0.1)
        self.pool = torch.nn.AvgPool2d(kernel_size=2, stride=2)
        for p in self.parameters():
            p.requires_grad = False

    def forward(self, x):
        return self.pool(torch.relu(self.conv(x)))


def head_kwargs(cfg_rel, patch_size):
    with open(os.path.join(ref_import.REFERENCE_ROOT, cfg_rel)) as f:
        cfg = yaml.full_load(f)
    kw = dict(cfg['MODEL']['HEAD'])
    kw['PATCH_SIZE'] = patch_size
    return kw


def smooth_images(B, P, seed):
    g = torch.Generator().manual_seed(seed)
    lo = torch.rand(B, 1, P // 4 + 1, P // 4 + 1, generator=g)
    im = torch.nn.functional.interpolate(lo, size=(P, P), mode='bicubic', align_corners=True)
    return (im + 0.05 * torch.randn(B, 1, P, P, generator=g)).float()


def golden_warp(PH):
    """Model._warp (PerceptualHead.py:237-243) on images and on the ones mask, values and d/d(delta)."""
    B, P = 4, 64
    g = torch.Generator().manual_seed(11)
    img = smooth_images(B, P, 3)
    delta = ((torch.rand(B, 4, 2, generator=g) * 2 - 1) * (P / 4)).float()
    delta[0] = 0.37 * delta[0]                      # one mild sample
    g_out = torch.randn(B, 1, P, P, generator=g).float()
    g_mask = torch.randn(B, 1, P, P, generator=g).float()
    out = dict(img=img.numpy(), delta=delta.numpy(), g_out=g_out.numpy(), g_mask=g_mask.numpy())
    for tag, dt in (('32', torch.float32), ('64', torch.float64)):
        d = delta.to(dt).requires_grad_(True)
        w, H = PH.Model._warp(img.to(dt), d)
        m, _ = PH.Model._warp(torch.ones_like(img).to(dt), d)
        gd_img, = torch.autograd.grad((w * g_out.to(dt)).sum(), d, retain_graph=True)
        gd_msk, = torch.autograd.grad((m * g_mask.to(dt)).sum(), d, retain_graph=True)
        # dH for a given dOut, taken at the H node
        Hleaf = H.detach().clone().requires_grad_(True)
        U = ref_import.load('src.data.utils')
        w2 = U.warp_image(img.to(dt), Hleaf, P, P)
        gH, = torch.autograd.grad((w2 * g_out.to(dt)).sum(), Hleaf)
        out.update({'warped' + tag: w.detach().numpy(), 'H' + tag: H.detach().numpy(),
                    'mask' + tag: m.detach().numpy(), 'gdelta_img' + tag: gd_img.numpy(),
                    'gdelta_mask' + tag: gd_msk.numpy(), 'gH_img' + tag: gH.numpy()})
    np.savez_compressed(os.path.join(OUT, 'warp_P64.npz'), **out)
    print('warp_P64', {k: v.shape for k, v in out.items()})


def golden_head(PH):
    """PerceptualHead.Model.forward, detone-bihome config (double-line / l1 / 'inf' / channel-agnostic)."""
    B, P = 4, 64
    kw = head_kwargs('config/s-coco/detone-bihome-lr-5e-3.yaml', P)
    g = torch.Generator().manual_seed(5)
    p1 = smooth_images(B, P, 21)
    p2 = smooth_images(B, P, 22) * 0.5 + p1 * 0.5
    d12 = ((torch.rand(B, 4, 2, generator=g) * 2 - 1) * (P / 5)).float()
    d21 = (-d12 + (torch.rand(B, 4, 2, generator=g) * 2 - 1) * 2).float()
    # sample 3: near-pure translation by ~61 px of 64 => ~3x3 px overlap, pooled den < 1 (max(den,1) clamp)
    d12[3] = torch.tensor([61.3, 60.6]) + 0.3 * d12[3] / (P / 5)
    d21[3] = -torch.tensor([60.9, 61.2]) + 0.3 * d21[3] / (P / 5)
    d12[2] *= 2.5          # large warp, partial overlap
    d21[2] *= 2.5
    ext = TinyExtractor()
    out = dict(patch_1=p1.numpy(), patch_2=p2.numpy(), delta_12=d12.numpy(), delta_21=d21.numpy(),
               ext_w=ext.conv.weight.detach().numpy(), ext_b=ext.conv.bias.detach().numpy(),
               mu=np.float64(kw['TRIPLET_MU']))
    for tag, dt in (('32', torch.float32), ('64', torch.float64)):
        model = PH.Model(backbone=torch.nn.Identity(), **kw)
        model.auxiliary_resnet = TinyExtractor().to(dt)
        a = d12.to(dt).requires_grad_(True)
        b = d21.to(dt).requires_grad_(True)
        data = {'patch_1': p1.to(dt), 'patch_2': p2.to(dt), 'delta_hat_12': a, 'delta_hat_21': b,
                'delta': torch.zeros(B, 4, 2, dtype=dt)}
        loss, delta_gt, delta_hat = model(data)
        ga, gb = torch.autograd.grad(loss, (a, b))
        out.update({'loss' + tag: loss.detach().numpy(), 'g12_' + tag: ga.numpy(), 'g21_' + tag: gb.numpy(),
                    'delta_hat' + tag: delta_hat.detach().numpy()})
    np.savez_compressed(os.path.join(OUT, 'head_doubleline_P64.npz'), **out)
    print('head_doubleline_P64 loss32=%r loss64=%r' % (out['loss32'], out['loss64']))


def golden_zeng(PH, RU):
    """DSAC branch (PerceptualHead.py:154-205 + ransac_utils.py:48-128) with the multinomial draw recorded."""
    B, P = 3, 32
    kw = head_kwargs('config/pds-coco/zeng-bihome-lr-1e-3.yaml', P)
    g = torch.Generator().manual_seed(9)
    # perspective fields = exact homography flow + noise (sample 2: pure noise, ill-conditioned on purpose)
    U = ref_import.load('src.data.utils')
    corners = torch.tensor([[0, 0], [P, 0], [P, P], [0, P]], dtype=torch.float64).repeat(B, 1, 1)
    dgt = (torch.rand(B, 4, 2, generator=g, dtype=torch.float64) * 2 - 1) * (P / 4)
    Hgt = U.four_point_to_homography(corners, dgt)
    ys, xs = torch.meshgrid(torch.arange(P, dtype=torch.float64), torch.arange(P, dtype=torch.float64), indexing='ij')
    pts = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(P * P, dtype=torch.float64)], 0)
    q = Hgt @ pts
    flow = (q[:, :2] / q[:, 2:3] - pts[:2]).reshape(B, 2, P, P)

    def fields(seed):
        gg = torch.Generator().manual_seed(seed)
        f = flow + 0.3 * torch.randn(B, 2, P, P, generator=gg, dtype=torch.float64)
        f[2] = 2.0 * torch.randn(2, P, P, generator=gg, dtype=torch.float64)
        return f.float()
    pf12, pf21 = fields(1), -fields(2)
    p1 = smooth_images(B, P, 31)
    p2 = smooth_images(B, P, 32)
    out = dict(pf12=pf12.numpy(), pf21=pf21.numpy(), patch_1=p1.numpy(), patch_2=p2.numpy(),
               mu=np.float64(kw['TRIPLET_MU']), points_per_hypothesis=np.int64(kw['POINTS_PER_HYPOTHESIS']))
    ext0 = TinyExtractor()
    out.update(ext_w=ext0.conv.weight.detach().numpy(), ext_b=ext0.conv.bias.detach().numpy())
    real_multinomial = torch.multinomial
    for tag, dt in (('32', torch.float32), ('64', torch.float64)):
        draws = []

        def recording(*a, **k):
            r = real_multinomial(*a, **k)
            draws.append(r.clone())
            return r
        torch.manual_seed(77)
        torch.multinomial = recording
        try:
            model = PH.Model(backbone=torch.nn.Identity(), **kw)
            model.auxiliary_resnet = TinyExtractor().to(dt)
            a = pf12.to(dt).requires_grad_(True)
            b = pf21.to(dt).requires_grad_(True)
            data = {'patch_1': p1.to(dt), 'patch_2': p2.to(dt), 'pf_hat_12': a, 'pf_hat_21': b}
            if dt == torch.float64:
                # the reference caches its coordinate field as float32 (PerceptualHead.py:139); for the
                # float64 ground truth pre-populate the caches in float64 (shape check :135 then passes)
                with torch.no_grad():
                    _, cf, fp = model.forward_map_field(pf12, None, None)
                model.coordinate_field_12 = cf.double()
                model.coordinate_field_21 = cf.double()
                model.four_points_12 = fp.double()
                model.four_points_21 = fp.double()
            loss, _, delta_hat = model(data)
            ga, gb = torch.autograd.grad(loss, (a, b))
            # the sampler on its own: homographies + scores
            dsac = RU.DSACSoftmax(**kw)
            coords = model.coordinate_field_12
            Hs, scores = dsac(coords, coords + a.detach().reshape(B, 2, -1).permute(0, 2, 1),
                              points_per_hypothesis=kw['POINTS_PER_HYPOTHESIS'], hypothesis_no=1)
        finally:
            torch.multinomial = real_multinomial
        assert len(draws) == 3
        if tag == '32':
            out.update(choice12=draws[0].numpy(), choice21=draws[1].numpy(), choice_dsac=draws[2].numpy())
        else:
            assert (out['choice12'] == draws[0].numpy()).all()
        out.update({'loss' + tag: loss.detach().numpy(), 'delta_hat12_' + tag: delta_hat.detach().numpy(),
                    'gpf12_' + tag: ga.numpy(), 'gpf21_' + tag: gb.numpy(),
                    'dsac_H' + tag: Hs.detach().numpy(), 'dsac_scores' + tag: scores.detach().numpy()})
    np.savez_compressed(os.path.join(OUT, 'zeng_dsac_P32.npz'), **out)
    print('zeng_dsac_P32 loss32=%r loss64=%r' % (out['loss32'], out['loss64']))


def golden_pairgen(T):
    """HomographyNetPrep -> DictToGrayscale -> DictStandardize -> DictToTensor from the shipped YAMLs."""
    res = {}
    for name, cfg_rel in (('pds', 'config/pds-coco/zeng-bihome-lr-1e-3.yaml'),
                          ('s', 'config/s-coco/detone-bihome-lr-5e-3.yaml')):
        with open(os.path.join(ref_import.REFERENCE_ROOT, cfg_rel)) as f:
            cfg = yaml.full_load(f)
        seed = cfg['DATA']['SAMPLER']['TRAIN_SEED']
        # train.py:111-120: every transform gets the seed appended as the last ctor argument
        tfs = []
        for t in cfg['DATA']['TRANSFORMS']:
            t_name = list(t.keys())[0]
            tfs.append(getattr(T, t_name)(*(t[t_name] + [seed])))
        n = 3
        p1s, p2s, deltas, homs, idx = [], [], [], [], []
        for i in range(n):
            image = pairgen.synthetic_image(i)
            data = ([image], None)
            for t in tfs:
                data = t(data)
            p1s.append(data['patch_1'].float().numpy())          # train.py:308-309 casts to float
            p2s.append(data['patch_2'].float().numpy())
            deltas.append(data['delta'].numpy())
            homs.append(data['homography'].numpy())
            idx.append(i)
        res.update({name + '_patch_1': np.stack(p1s), name + '_patch_2': np.stack(p2s),
                    name + '_delta': np.stack(deltas), name + '_homography': np.stack(homs),
                    name + '_seed': np.int64(seed), name + '_image_index': np.asarray(idx)})
    np.savez_compressed(os.path.join(OUT, 'pairgen.npz'), **res)
    print('pairgen', {k: v.shape for k, v in res.items()})


def main():
    if not ref_import.available():
        sys.exit('needs the reference tree at ' + ref_import.REFERENCE_ROOT)
    os.makedirs(OUT, exist_ok=True)
    import warnings
    warnings.filterwarnings('ignore')
    PH = ref_import.load('src.heads.PerceptualHead')
    RU = ref_import.load('src.heads.ransac_utils')
    T = ref_import.load('src.data.transforms')
    golden_warp(PH)
    golden_head(PH)
    golden_zeng(PH, RU)
    golden_pairgen(T)


if __name__ == '__main__':
    main()
