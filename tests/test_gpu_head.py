"""GPU parity of the head-level path: K4 (N-point DLT), K5 (pair generation) and the PerceptualHead mirror,
against the golden vectors of the unmodified reference and the CPU oracle."""
import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module', autouse=True)
def _exact_convs():
    assert torch.cuda.is_available()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def cu(a, dt=torch.float32):
    return torch.as_tensor(np.asarray(a)).to(dt).cuda()


def tiny_extractor():
    from oracle.make_golden import TinyExtractor
    return TinyExtractor().cuda()


def head_kwargs(name, P):
    import yaml, os
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, 'golden', name)) as f:
        kw = dict(yaml.full_load(f)['MODEL']['HEAD'])
    kw['PATCH_SIZE'] = P
    kw['AUXILIARY_RESNET_PRETRAINED'] = False
    return kw


# ------------------------------------------------------------------------------------------------ K4
def test_dltn_golden(golden):
    import bihome_b200.functional as F
    g = golden('zeng_dsac_P32.npz')
    B, _, P, _ = g['pf12'].shape
    M = int(g['points_per_hypothesis'])
    pf = cu(g['pf12'])
    four = torch.tensor([[0, 0], [P, 0], [P, P], [0, P]], dtype=torch.float32).cuda()
    H, _ = F.dltn_field(pf, cu(g['choice_dsac'], torch.int64).view(B, M), four)
    # float64 reference to 1e-5; the reference's own float32 result is only good to ~1e-3 here
    assert rel_l2(H.cpu().numpy(), g['dsac_H64'][:, 0]) < 1e-5
    assert rel_l2(g['dsac_H32'], g['dsac_H64']) > 1e-6
    _, delta = F.dltn_field(pf, cu(g['choice12'], torch.int64).view(B, M), four)
    assert rel_l2(delta.cpu().numpy(), g['delta_hat12_64']) < 1e-5


@pytest.mark.parametrize('B,N,M', [(3, 400, 128), (70, 256, 40), (2, 64, 64)])
def test_dltn_points_vs_oracle_with_gradient(B, N, M):
    import bihome_b200.functional as F
    from oracle import kornia050 as K
    gen = torch.Generator().manual_seed(N)
    p1 = (torch.rand(B, N, 2, generator=gen, dtype=torch.float64) * 100).float().double()
    Ht = torch.eye(3, dtype=torch.float64) + 0.01 * torch.randn(B, 3, 3, generator=gen, dtype=torch.float64)
    Ht[:, 2, :2] *= 0.01
    q = torch.nn.functional.pad(p1, (0, 1), value=1.0) @ Ht.transpose(1, 2)
    p2 = (q[..., :2] / q[..., 2:] + 0.5 * torch.randn(B, N, 2, generator=gen, dtype=torch.float64)).float().double()
    choice = torch.randint(0, N, (B, M), generator=gen) if M != N else None
    gH = torch.randn(B, 3, 3, generator=gen, dtype=torch.float64)
    p2r = p2.clone().requires_grad_(True)
    idx = choice if choice is not None else torch.arange(N).repeat(B, 1)
    s1 = torch.gather(p1, 1, idx.unsqueeze(-1).repeat(1, 1, 2))
    s2 = torch.gather(p2r, 1, idx.unsqueeze(-1).repeat(1, 1, 2))
    H64 = K.find_homography_dlt(s1, s2)
    g64, = torch.autograd.grad((H64 * gH).sum(), p2r)
    p2c = p2.float().cuda().requires_grad_(True)
    H = F.dltn(p1.float().cuda(), p2c, None if choice is None else choice.cuda())
    assert rel_l2(H.detach().cpu().numpy(), H64.detach().numpy()) < 1e-5
    g, = torch.autograd.grad((H * gH.float().cuda()).sum(), p2c)
    assert rel_l2(g.cpu().numpy(), g64.numpy()) < 1e-4


# ------------------------------------------------------------------------------------------------ K5
@pytest.mark.parametrize('max_delta', [32, 0])
def test_pairgen_vs_oracle(max_delta):
    import bihome_b200.functional as F
    from oracle import pairgen
    n_img, B, P, rho = 4, 12, 128, 32
    images = np.stack([pairgen.synthetic_image(i) for i in range(n_img)])
    rs = np.random.RandomState(5 + max_delta)
    qs, idx = [], []
    for b in range(B):
        idx.append(b % n_img)
        qs.append(pairgen.draw_params(rs, 240, 320, rho, P, max_delta))
    params = torch.from_numpy(pairgen.pack_params(qs)).cuda()
    p1, p2, delta = F.pairgen_apply(torch.from_numpy(images).cuda(), torch.tensor(idx, dtype=torch.int32).cuda(), params, P)
    for b in range(B):
        out = pairgen.make_pair(images[idx[b]], qs[b], P)
        r1, r2 = pairgen.to_network_input(out['patch_1']), pairgen.to_network_input(out['patch_2'])
        assert np.array_equal(delta[b].cpu().numpy(), qs[b]['delta'].astype(np.float32))
        # float32 colour chain re-implemented op for op; residual = 1-ulp differences in OpenCV's SIMD hue path
        assert np.abs(p1[b].cpu().numpy() - r1).max() < 2e-4, b
        ties = pairgen.rounding_ties(qs[b], P)
        assert ties.mean() < 0.01
        assert np.abs(p2[b].cpu().numpy() - r2)[0][~ties].max() < 2e-4, b
        assert (p1[b].cpu().numpy() == r1).mean() > 0.9
    assert p2.data_ptr() == p1.data_ptr() + p1.numel() * 4


@pytest.mark.parametrize('max_delta', [32, 0])
def test_pairgen_image_1_vs_oracle(max_delta):
    """the whole first image as the reference's batch carries it ('image_1' after DictToGrayscale / DictStandardize /
    DictToTensor, the tensor PhotometricHead warps in s-coco/nguyen-orig), and the loader's batch for such a config"""
    import bihome_b200.functional as F
    from bihome_b200.data import gpu_pairs
    from oracle import pairgen
    n_img, B = 3, 5
    images = np.stack([pairgen.synthetic_image(i) for i in range(n_img)])
    rs = np.random.RandomState(77 + max_delta)
    qs = [pairgen.draw_params(rs, 240, 320, 32, 128, max_delta) for _ in range(B)]
    idx = [b % n_img for b in range(B)]
    im1 = F.pairgen_image(torch.from_numpy(images).cuda(), torch.tensor(idx, dtype=torch.int32).cuda(),
                          torch.from_numpy(pairgen.pack_params(qs)).cuda())
    assert im1.shape == (B, 1, 240, 320)
    for b in range(B):
        ref = pairgen.to_network_input(pairgen.apply_photometric(images[idx[b]], qs[b]['photo_1']))
        assert np.abs(im1[b].cpu().numpy() - ref).max() < 2e-4, b
        assert (im1[b].cpu().numpy() == ref).mean() > 0.9
    # patch_1 is the crop of image_1 at the batch's corners, bit for bit (same chain, same pixels)
    loader = gpu_pairs.GpuPairLoader(torch.from_numpy(images).cuda(), 4, 8, max_delta=float(max_delta), image_keys=('image_1',))
    batch = loader.next_batch()
    c = batch['corners'].long().cpu()
    for b in range(4):
        x0, y0 = int(c[b, 0, 0]), int(c[b, 0, 1])
        assert torch.equal(batch['image_1'][b, :, y0:y0 + 128, x0:x0 + 128], batch['patch_1'][b])


def test_pairgen_golden_reference_pipeline(golden):
    """explicit parameters replayed from the seeded RandomState the reference transforms consumed"""
    import bihome_b200.functional as F
    from oracle import pairgen
    g = golden('pairgen.npz')
    for name, max_delta in (('pds', 32), ('s', 0)):
        rs = np.random.RandomState(int(g[name + '_seed']))
        ids = [int(i) for i in g[name + '_image_index']]
        images = np.stack([pairgen.synthetic_image(i) for i in ids])
        qs = [pairgen.draw_params(rs, 240, 320, 32, 128, max_delta) for _ in ids]
        p1, p2, delta = F.pairgen_apply(torch.from_numpy(images).cuda(), torch.arange(len(ids), dtype=torch.int32).cuda(),
                                        torch.from_numpy(pairgen.pack_params(qs)).cuda(), 128)
        assert np.abs(p1.cpu().numpy() - g[name + '_patch_1']).max() < 2e-4
        ties = np.stack([pairgen.rounding_ties(q, 128) for q in qs])
        assert np.abs(p2.cpu().numpy() - g[name + '_patch_2'])[:, 0][~ties].max() < 2e-4
        assert np.array_equal(delta.cpu().numpy(), g[name + '_delta'].astype(np.float32))


def test_pairgen_draw_distributions():
    import bihome_b200.functional as F
    B = 20000
    params, index = F.pairgen_draw(B, 7, (240, 320), 32, 128, 32.0, seed=42, step=3, device='cuda')
    p = params.cpu().numpy()
    assert index.min() >= 0 and index.max() == 6
    assert p[:, 22].min() == 96 and p[:, 22].max() == 224 and p[:, 23].min() == 96 and p[:, 23].max() == 144
    d = p[:, 24:]
    assert d.min() == -32 and d.max() == 31 and abs(d.mean() + 0.5) < 0.3
    assert abs(p[:, 0].mean() - 0.5) < 0.02 and np.abs(p[:, 1]).max() <= 32 and (p[p[:, 0] == 0, 1] == 0).all()
    assert p[:, 4].min() >= 0.5 and p[:, 4].max() <= 1.5 and np.abs(p[:, 8]).max() <= 16
    assert set(np.unique(p[:, 10])) <= set(range(6))
    again, _ = F.pairgen_draw(B, 7, (240, 320), 32, 128, 32.0, seed=42, step=3, device='cuda')
    other, _ = F.pairgen_draw(B, 7, (240, 320), 32, 128, 32.0, seed=42, step=4, device='cuda')
    assert torch.equal(params, again) and not torch.equal(params, other)


# ------------------------------------------------------------------------------------------------ head
def test_head_double_line_golden(golden):
    from bihome_b200.heads import PerceptualHead as PH
    g = golden('head_doubleline_P64.npz')
    P = g['patch_1'].shape[-1]
    model = PH.Model(backbone=torch.nn.Identity(), **head_kwargs('detone-bihome.yaml', P)).cuda()
    model.auxiliary_resnet = tiny_extractor()
    a, b = cu(g['delta_12']).requires_grad_(True), cu(g['delta_21']).requires_grad_(True)
    data = {'patch_1': cu(g['patch_1']), 'patch_2': cu(g['patch_2']), 'delta_hat_12': a, 'delta_hat_21': b,
            'delta': torch.zeros(4, 4, 2).cuda()}
    loss, delta_gt, delta_hat = model(data)
    assert abs(loss.item() - float(g['loss64'])) < 1e-5 * abs(float(g['loss64']))
    ga, gb = torch.autograd.grad(loss, (a, b))
    assert rel_l2(ga.cpu().numpy(), g['g12_64']) < 1e-4
    assert rel_l2(gb.cpu().numpy(), g['g21_64']) < 1e-4
    assert np.array_equal(delta_hat.detach().cpu().numpy(), g['delta_12'])


def test_head_zeng_golden(golden):
    from bihome_b200.heads import PerceptualHead as PH
    g = golden('zeng_dsac_P32.npz')
    P = g['patch_1'].shape[-1]
    model = PH.Model(backbone=torch.nn.Identity(), **head_kwargs('zeng-bihome.yaml', P)).cuda()
    model.auxiliary_resnet = tiny_extractor()
    model.forced_choice = [cu(g['choice12'], torch.int64), cu(g['choice21'], torch.int64)]
    a, b = cu(g['pf12']).requires_grad_(True), cu(g['pf21']).requires_grad_(True)
    loss, _, delta_hat = model({'patch_1': cu(g['patch_1']), 'patch_2': cu(g['patch_2']), 'pf_hat_12': a, 'pf_hat_21': b})
    assert rel_l2(delta_hat.detach().cpu().numpy(), g['delta_hat12_64']) < 1e-5
    assert abs(loss.item() - float(g['loss64'])) < 2e-5 * max(1.0, abs(float(g['loss64'])))
    ga, gb = torch.autograd.grad(loss, (a, b))
    # sample 2 of the fixture is a pure-noise field (ill-conditioned on purpose): compare the well-posed ones tightly
    assert rel_l2(ga[:2].cpu().numpy(), g['gpf12_64'][:2]) < 1e-4
    assert rel_l2(gb[:2].cpu().numpy(), g['gpf21_64'][:2]) < 1e-4
    assert rel_l2(ga.cpu().numpy(), g['gpf12_64']) < 1e-3


def test_head_generic_variants_match_fused():
    """the torch-algebra path (one-line / margins / ...) and the fused kernel agree on the configuration both cover"""
    from bihome_b200.heads import PerceptualHead as PH
    kw = head_kwargs('detone-bihome.yaml', 64)
    torch.manual_seed(0)
    model = PH.Model(backbone=torch.nn.Identity(), **kw).cuda()
    model.auxiliary_resnet = tiny_extractor()
    p1, p2 = torch.rand(3, 1, 64, 64).cuda(), torch.rand(3, 1, 64, 64).cuda()
    d12 = ((torch.rand(3, 4, 2) * 2 - 1) * 10).cuda().requires_grad_(True)
    d21 = ((torch.rand(3, 4, 2) * 2 - 1) * 10).cuda().requires_grad_(True)
    data = lambda: {'patch_1': p1, 'patch_2': p2, 'delta_hat_12': d12, 'delta_hat_21': d21}
    loss_f, _, _ = model(data())
    gf = torch.autograd.grad(loss_f, (d12, d21))
    model._fused_path = lambda: False
    loss_g, _, _ = model(data())
    gg = torch.autograd.grad(loss_g, (d12, d21))
    assert abs(loss_f.item() - loss_g.item()) < 1e-5 * abs(loss_g.item())
    for x, y in zip(gf, gg):
        assert rel_l2(x.cpu().numpy(), y.cpu().numpy()) < 1e-4
