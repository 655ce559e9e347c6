"""CPU: the oracle restatement (oracle/ref_path.py, oracle/pairgen.py) against the committed golden vectors
(made by oracle/make_golden.py from the UNMODIFIED reference modules) and against OpenCV."""
import cv2
import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import kornia050 as K
from oracle import pairgen, ref_path as R
from oracle.make_golden import TinyExtractor


def T(a, dt):
    return torch.from_numpy(np.asarray(a)).to(dt)


@pytest.mark.parametrize('tag,dt', [('32', torch.float32), ('64', torch.float64)])
def test_warp_matches_reference(golden, tag, dt):
    g = golden('warp_P64.npz')
    img, d = T(g['img'], dt), T(g['delta'], dt).requires_grad_(True)
    w, H = R.warp(img, d)
    m, _ = R.warp(torch.ones_like(img), d)
    assert np.array_equal(w.detach().numpy(), g['warped' + tag])
    assert np.array_equal(H.detach().numpy(), g['H' + tag])
    assert np.array_equal(m.detach().numpy(), g['mask' + tag])
    gd, = torch.autograd.grad((w * T(g['g_out'], dt)).sum(), d, retain_graph=True)
    assert rel_l2(gd.numpy(), g['gdelta_img' + tag]) < 1e-6
    gm, = torch.autograd.grad((m * T(g['g_mask'], dt)).sum(), d)
    assert rel_l2(gm.numpy(), g['gdelta_mask' + tag]) < 1e-6


def test_closed_forms_match_reference_fp64(golden):
    """warp_direct / analytic_mask (the semantics the CUDA kernels implement) vs the reference's fp64 output.
    The reference builds its sampling grid from a float32 linspace even in float64 (kornia create_meshgrid),
    which bounds the agreement at ~1e-6 relative."""
    g = golden('warp_P64.npz')
    img, H = T(g['img'], torch.float64), T(g['H64'], torch.float64)
    assert rel_l2(R.warp_direct(img, H).numpy(), g['warped64']) < 5e-6
    P = img.shape[-1]
    assert np.abs(R.analytic_mask(H, P, P, P, P).numpy() - g['mask64']).max() < 2e-5


def test_reference_fp32_noise_floor(golden):
    g = golden('warp_P64.npz')
    assert rel_l2(g['warped32'], g['warped64']) < 2e-5
    assert rel_l2(g['H32'], g['H64']) < 1e-6


def test_dlt4_against_opencv(golden):
    g = golden('warp_P64.npz')
    P = g['img'].shape[-1]
    c = np.float32([[0, 0], [P, 0], [P, P], [0, P]])
    for b in range(g['delta'].shape[0]):
        Hc = cv2.getPerspectiveTransform(c, c + g['delta'][b])
        assert rel_l2(g['H64'][b], Hc) < 1e-6


def test_warp_against_opencv():
    """cv2.warpPerspective quantises taps to 1/32 px: coarse semantic pin; exact on an integer translation."""
    torch.manual_seed(0)
    P = 64
    img = torch.rand(1, 1, P, P, dtype=torch.float64)
    d = torch.tensor([[[3., -2.], [3., -2.], [3., -2.], [3., -2.]]], dtype=torch.float64)
    w, H = R.warp(img, d)
    ref = cv2.warpPerspective(img[0, 0].numpy(), np.linalg.inv(H[0].numpy()), dsize=(P, P))
    assert np.abs(w[0, 0].numpy() - ref).max() < 1e-5      # fp32 linspace grid inside kornia
    d = (torch.rand(1, 4, 2, dtype=torch.float64) * 2 - 1) * 12
    w, H = R.warp(img, d)
    ref = cv2.warpPerspective(img[0, 0].numpy(), np.linalg.inv(H[0].numpy()), dsize=(P, P))
    assert np.abs(w[0, 0].numpy() - ref).max() < 0.05


@pytest.mark.parametrize('tag,dt', [('32', torch.float32), ('64', torch.float64)])
def test_head_double_line_matches_reference(golden, tag, dt):
    g = golden('head_doubleline_P64.npz')
    ext = TinyExtractor().to(dt)
    a = T(g['delta_12'], dt).requires_grad_(True)
    b = T(g['delta_21'], dt).requires_grad_(True)
    loss, parts = R.head_double_line(T(g['patch_1'], dt), T(g['patch_2'], dt), a, b, ext, float(g['mu']))
    assert abs(loss.item() - float(g['loss' + tag])) <= 2e-6 * abs(float(g['loss' + tag]))
    ga, gb = torch.autograd.grad(loss, (a, b))
    tol = 1e-4 if tag == '32' else 1e-9
    assert rel_l2(ga.numpy(), g['g12_' + tag]) < tol
    assert rel_l2(gb.numpy(), g['g21_' + tag]) < tol
    # the large-warp sample exercises the max(den, 1) clamp
    assert 0.0 < float(parts['den1'].detach().min()) < 1.0


@pytest.mark.parametrize('tag,dt', [('32', torch.float32), ('64', torch.float64)])
def test_zeng_branch_matches_reference(golden, tag, dt):
    g = golden('zeng_dsac_P32.npz')
    pph = int(g['points_per_hypothesis'])
    pf = T(g['pf12'], dt)
    delta, H, scores = R.zeng_delta_hat(pf, pph, 1, choice=torch.from_numpy(g['choice_dsac']))
    assert rel_l2(H.numpy(), g['dsac_H' + tag]) < (1e-3 if tag == '32' else 1e-9)
    assert np.allclose(scores.numpy(), g['dsac_scores' + tag])
    delta, _, _ = R.zeng_delta_hat(pf, pph, 1, choice=torch.from_numpy(g['choice12']))
    tol = 5e-3 if tag == '32' else 1e-9       # fp32 SVD of a Gram matrix: ill-conditioned sample included
    assert rel_l2(delta.reshape(-1, 4, 2).numpy(), g['delta_hat12_' + tag]) < tol


def test_zeng_dlt_against_opencv():
    """find_homography_dlt on exactly consistent correspondences == cv2.findHomography(method=0)."""
    rs = np.random.RandomState(0)
    P = 64
    c = np.float32([[0, 0], [P, 0], [P, P], [0, P]])
    Ht = cv2.getPerspectiveTransform(c, c + rs.uniform(-12, 12, (4, 2)).astype(np.float32))
    pts = rs.uniform(0, P, (128, 2))
    q = cv2.perspectiveTransform(pts[None].astype(np.float64), Ht)[0]
    Hk = K.find_homography_dlt(torch.from_numpy(pts)[None], torch.from_numpy(q)[None])[0].numpy()
    Hc, _ = cv2.findHomography(pts, q, method=0)
    assert rel_l2(Hk, Hc) < 1e-6 and rel_l2(Hk, Ht) < 1e-6


@pytest.mark.parametrize('name,max_delta', [('pds', 32), ('s', 0)])
def test_pairgen_replay_is_bit_exact(golden, name, max_delta):
    g = golden('pairgen.npz')
    rs = np.random.RandomState(int(g[name + '_seed']))
    for i, idx in enumerate(g[name + '_image_index']):
        image = pairgen.synthetic_image(int(idx))
        q = pairgen.draw_params(rs, image.shape[0], image.shape[1], 32, 128, max_delta)
        out = pairgen.make_pair(image, q, 128)
        assert np.array_equal(q['delta'], g[name + '_delta'][i])
        assert np.array_equal(out['homography'], g[name + '_homography'][i])
        assert np.array_equal(pairgen.to_network_input(out['patch_1']), g[name + '_patch_1'][i])
        assert np.array_equal(pairgen.to_network_input(out['patch_2']), g[name + '_patch_2'][i])


def test_mace():
    a = np.zeros((2, 4, 2)); b = np.ones((2, 4, 2)) * 3
    assert abs(R.mace(a, b) - 3 * np.sqrt(2)) < 1e-12
