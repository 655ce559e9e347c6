"""CPU restatement of the reference's hot path -- TEST INFRASTRUCTURE ONLY.

Follows, function by function (paths relative to /root/reference):

  image_shape_to_corners     src/data/utils.py:36-51
  four_point_to_homography   src/data/utils.py:7-33   (torch branch -> kornia.get_perspective_transform)
  warp_image                 src/data/utils.py:54-59  (torch branch -> torch.inverse + kornia.warp_perspective)
  warp                       src/heads/PerceptualHead.py:237-243   (Model._warp)
  bihome_double_line         src/heads/PerceptualHead.py:447-459 (mask pooling), 559-561 (l1 distances),
                             609-665 (ln1, ln2, ln3, loss; margin 'inf', channel-agnostic)
  head_double_line           src/heads/PerceptualHead.py:320-402 (+ the loss above): the whole biHomE head
  triplet_general            src/heads/PerceptualHead.py:465-538 (one-line: l1 / cosine, numeric margin, MASK_CRD),
                             :555-665 (double-line: l1 / l2 / cosine, 'inf' or numeric margin, channel-aware /
                             channel-agnostic) and src/heads/TripletHead.py:78-153 (the same algebra on the content-aware
                             backbone's full-resolution maps) -- every variant of the masked triplet loss beside the
                             north-star one, on already pooled masks
  field_to_points            src/heads/PerceptualHead.py:125-146   (forward_map_field)
  dsac_hypotheses            src/heads/ransac_utils.py:48-74       (__sample_hypotheses)
  dsac_scores                src/heads/ransac_utils.py:76-128      ('repr_error' mode)
  zeng_delta_hat             src/heads/PerceptualHead.py:164-178   (DSAC branch of forward)
  mace                       train.py:401-404 / eval.py:133-134

plus two independent closed forms used as fp64 ground truth:

  warp_direct                out[y,x] = bilinear(src, H.[x,y,1]) written tap by tap
  analytic_mask              warp(ones) = mx(u) * my(v)

All functions are dtype generic (float32 like-for-like, float64 truth) and run
on CPU tensors.
"""
import numpy as np
import torch

from . import kornia050 as K


# ---------------------------------------------------------------------------------------------
# src/data/utils.py
# ---------------------------------------------------------------------------------------------
def image_shape_to_corners(patch):
    assert patch.dim() == 4
    # the reference names shape[-2] "width" and shape[-1] "height" (utils.py:39-40): keep the swap
    a, b = patch.shape[-2], patch.shape[-1]
    c = torch.tensor([[0, 0], [a, 0], [a, b], [0, b]], dtype=patch.dtype, device=patch.device)
    return c.repeat(patch.shape[0], 1, 1)


def four_point_to_homography(corners, deltas, crop=False):
    assert corners.dim() == 3 and deltas.dim() == 3
    if crop:
        corners = corners - corners[:, 0].view(-1, 1, 2)
    return K.get_perspective_transform(corners, corners + deltas)


def warp_image(image, homography, target_h, target_w, inverse=True):
    if inverse:
        homography = torch.inverse(homography)
    return K.warp_perspective(image, homography, (target_h, target_w))


# ---------------------------------------------------------------------------------------------
# src/heads/PerceptualHead.py
# ---------------------------------------------------------------------------------------------
def warp(image, delta_hat, corners=None):
    if corners is None:
        corners = image_shape_to_corners(image)
    H = four_point_to_homography(corners, delta_hat, crop=False)
    return warp_image(image, H, image.shape[-2], image.shape[-1]), H


def bihome_double_line(f1, f2, f1w, f2w, m1, m2, m1w, m2w, H12, H21, mu):
    """Masks are full resolution [B,1,P,P]; features [B,C,h,w].  Returns (loss, dict of parts)."""
    k = m1.shape[-1] // f1w.shape[-2]
    pool = torch.nn.AvgPool2d(kernel_size=k, stride=k, padding=0)
    a1, b2 = pool(m1w).squeeze(1), pool(m2).squeeze(1)
    b1, a2 = pool(m1).squeeze(1), pool(m2w).squeeze(1)
    l1 = torch.abs(f1w - f2)
    l2 = torch.abs(f2w - f1)
    l3 = torch.abs(f1 - f2)
    den1 = (a1 * b2).sum(-1).sum(-1)
    mat1 = l1.sum(1) - l3.sum(1)
    ln1 = (a1 * b2 * mat1).sum(-1).sum(-1) / torch.max(den1, torch.ones_like(den1))
    den2 = (a2 * b1).sum(-1).sum(-1)
    mat2 = l2.sum(1) - l3.sum(1)
    ln2 = (a2 * b1 * mat2).sum(-1).sum(-1) / torch.max(den2, torch.ones_like(den2))
    eye = torch.eye(3, dtype=H12.dtype, device=H12.device).unsqueeze(0)
    ln3_b = ((torch.matmul(H12, H21) - eye) ** 2).sum(-1).sum(-1)
    loss = ln1.sum() + ln2.sum() + mu * ln3_b.sum()
    return loss, dict(ln1=ln1, ln2=ln2, den1=den1, den2=den2, ln3=ln3_b,
                      m1w_pooled=a1, m2w_pooled=a2)


def triplet_general(f1, f2, f1w, f2w, a1, b2, a2, b1, H12, H21, lines, distance, hinge, margin, mask_crd=False, mu=0.0,
                    scale=(1.0, 1.0)):
    """The masked triplet loss of the reference in all its well-defined variants, per sample.

    features [B,C,h,w]; masks [B,h,w] at the feature resolution (a = warped mask of the line's source patch, b = mask of
    its target patch; b None == ones); ``lines`` 1 (one-line, PerceptualHead.py:465-538; TripletHead 'oneline') or 2;
    ``distance`` 'l1' | 'l2' | 'cosine' (:559-603); ``hinge`` None (margin 'inf': the signed difference, :614-620),
    'channel' (channel-aware numeric margin: max(. + margin, 0) per channel, then the channel sum, :622-623) or 'pixel'
    (max(. + margin, 0) of the channel-aggregated distances, :512 and :625-626); ``mask_crd`` weighs with a alone
    (:533-537).  ``scale`` multiplies the two lines (the B-fold broadcast of TripletHead.py:91-92 for one-channel
    features).  Returns (loss_b [B], dict(ln1, ln2, den1, den2, ln3))."""
    def dist(x, y):
        if distance == 'l1':
            return (x - y).abs()                                            # per channel
        if distance == 'l2':
            return ((x - y) ** 2).mean(1)                                   # per pixel
        return 1 - torch.cosine_similarity(x, y, dim=1)                     # per pixel

    def line(la, lb, m):
        per_channel = la.dim() == 4
        if hinge is None:
            return (la - lb).sum(1) if per_channel else la - lb
        if hinge == 'channel':
            assert per_channel, 'a per-channel margin needs per-channel distances (l1)'
            return torch.clamp(la - lb + m, min=0).sum(1)
        if per_channel:
            la, lb = la.sum(1), lb.sum(1)
        return torch.clamp(la - lb + m, min=0)

    def masked(a, b, mat):
        w = a if (mask_crd or b is None) else a * b
        den = w.sum(-1).sum(-1)
        return (w * mat).sum(-1).sum(-1) / torch.max(den, torch.ones_like(den)), den

    l3 = dist(f1, f2)
    ln1, den1 = masked(a1, b2, line(dist(f1w, f2), l3, margin))
    ln1 = ln1 * scale[0]
    zero = torch.zeros_like(ln1)
    if lines == 1:
        return ln1, dict(ln1=ln1, ln2=zero, den1=den1, den2=zero, ln3=zero)
    ln2, den2 = masked(a2, b1, line(dist(f2w, f1), l3, margin))
    ln2 = ln2 * scale[1]
    eye = torch.eye(3, dtype=H12.dtype, device=H12.device).unsqueeze(0)
    ln3 = ((torch.matmul(H12, H21) - eye) ** 2).sum(-1).sum(-1)
    return ln1 + ln2 + mu * ln3, dict(ln1=ln1, ln2=ln2, den1=den1, den2=den2, ln3=ln3)


def head_double_line(patch_1, patch_2, delta_12, delta_21, extractor, mu, mask_1=None, mask_2=None):
    """The whole biHomE head for given 4-point offsets (DeTone/Zhang configs and the tail of Zeng)."""
    m1 = torch.ones_like(patch_1) if mask_1 is None else mask_1
    m2 = torch.ones_like(patch_2) if mask_2 is None else mask_2
    f1 = extractor(patch_1)
    f2 = extractor(patch_2)
    p1w, _ = warp(patch_1, delta_12)
    f1w = extractor(p1w)
    m1w, H12 = warp(m1, delta_12)
    p2w, _ = warp(patch_2, delta_21)
    f2w = extractor(p2w)
    m2w, H21 = warp(m2, delta_21)
    loss, parts = bihome_double_line(f1, f2, f1w, f2w, m1, m2, m1w, m2w, H12, H21, mu)
    parts.update(p1w=p1w, p2w=p2w, m1w=m1w, m2w=m2w, H12=H12, H21=H21, f1=f1, f2=f2, f1w=f1w, f2w=f2w)
    return loss, parts


def field_to_points(pf):
    """pf [B,2,P,Q] -> (coords [B,PQ,2] as (x,y), coords + field, four_points [4,2])."""
    B, _, P, Q = pf.shape
    yy, xx = np.mgrid[0:P, 0:Q]
    coords = torch.from_numpy(np.stack((xx.reshape(-1), yy.reshape(-1)), axis=-1)).to(pf.dtype).to(pf.device)
    coords = coords.unsqueeze(0).repeat(B, 1, 1)
    four = torch.tensor([[0, 0], [Q, 0], [Q, P], [0, P]], dtype=pf.dtype, device=pf.device)
    return coords, coords + pf.reshape(B, 2, -1).permute(0, 2, 1), four


def multinomial_choice(n_points, count, generator=None):
    """The reference's sampler: weights = arange(N) (index-proportional), with replacement."""
    w = torch.arange(0, n_points, dtype=torch.float32)
    return torch.multinomial(w, count, replacement=True, generator=generator)


def dsac_hypotheses(points1, points2, points_per_hypothesis, hypothesis_no, choice=None):
    B = points1.shape[0]
    if choice is None:
        choice = multinomial_choice(points1.shape[1], B * points_per_hypothesis * hypothesis_no)
    idx = choice.to(points1.device).reshape(B, -1, 1).repeat(1, 1, 2)
    s1 = torch.gather(points1, 1, idx).reshape(B * hypothesis_no, points_per_hypothesis, 2)
    s2 = torch.gather(points2, 1, idx).reshape(B * hypothesis_no, points_per_hypothesis, 2)
    return K.find_homography_dlt(s1, s2).reshape(B, hypothesis_no, 3, 3)


def dsac_scores(points1, points2, homographies):
    B, n = homographies.shape[:2]
    N = points1.shape[1]
    p1 = points1.reshape(B, 1, N, 2).repeat(1, n, 1, 1).reshape(B * n, N, 2)
    p2 = points2.reshape(B, 1, N, 2).repeat(1, n, 1, 1).reshape(B * n, N, 2)
    proj = K.transform_points(homographies.reshape(B * n, 3, 3), p1)
    err = torch.abs(proj - p2).sum(-1).sum(-1).reshape(B, n)
    return torch.softmax(-err, dim=-1)


def zeng_delta_hat(pf, points_per_hypothesis, hypothesis_no=1, choice=None):
    """Perspective field -> (delta_hat [B,n,4,2], H [B,n,3,3], scores [B,n])."""
    B = pf.shape[0]
    coords, mapped, four = field_to_points(pf)
    H = dsac_hypotheses(coords, mapped, points_per_hypothesis, hypothesis_no, choice)
    scores = dsac_scores(coords, mapped, H)
    four = four.unsqueeze(0).repeat(B * hypothesis_no, 1, 1)
    delta = (K.transform_points(H.reshape(-1, 3, 3), four) - four).reshape(B, hypothesis_no, 4, 2)
    return delta, H, scores


def mace(delta_gt, delta_hat):
    d = np.asarray(delta_gt, dtype=np.float64).reshape(-1, 2) - np.asarray(delta_hat, dtype=np.float64).reshape(-1, 2)
    return float(np.mean(np.linalg.norm(d, axis=-1)))


# ---------------------------------------------------------------------------------------------
# Independent closed forms -- fp64 ground truth
# ---------------------------------------------------------------------------------------------
def _pixel_coords(H, h_out, w_out):
    ys, xs = torch.meshgrid(torch.arange(h_out, dtype=H.dtype), torch.arange(w_out, dtype=H.dtype), indexing='ij')
    h = H.reshape(-1, 9)
    g = lambda i: h[:, i].view(-1, 1, 1)
    w = g(6) * xs + g(7) * ys + g(8)
    u = (g(0) * xs + g(1) * ys + g(2)) / w
    v = (g(3) * xs + g(4) * ys + g(5)) / w
    return u, v


def warp_direct(image, H, h_out=None, w_out=None):
    """out[b,c,y,x] = sum over the 4 bilinear taps of image[b,c] at (u,v) = proj(H_b [x,y,1]); zeros outside."""
    B, C, Hs, Ws = image.shape
    h_out = Hs if h_out is None else h_out
    w_out = Ws if w_out is None else w_out
    u, v = _pixel_coords(H, h_out, w_out)
    x0 = torch.floor(u)
    y0 = torch.floor(v)
    out = torch.zeros(B, C, h_out, w_out, dtype=image.dtype)
    flat = image.reshape(B, C, -1)
    for dy in (0, 1):
        for dx in (0, 1):
            xi = x0 + dx
            yi = y0 + dy
            wx = (u - x0) if dx else (x0 + 1 - u)
            wy = (v - y0) if dy else (y0 + 1 - v)
            ok = (xi >= 0) & (xi <= Ws - 1) & (yi >= 0) & (yi <= Hs - 1)
            idx = (yi.clamp(0, Hs - 1) * Ws + xi.clamp(0, Ws - 1)).long().view(B, 1, -1).expand(B, C, -1)
            val = torch.gather(flat, 2, idx).view(B, C, h_out, w_out)
            out = out + val * (wx * wy * ok.to(image.dtype)).unsqueeze(1)
    return out


def analytic_mask(H, Hs, Ws, h_out, w_out):
    """warp(ones): separable coverage mx(u)*my(v), m(t) = clamp(min(t+1, size-t), 0, 1)."""
    u, v = _pixel_coords(H, h_out, w_out)
    mx = torch.minimum(u + 1, Ws - u).clamp(0, 1)
    my = torch.minimum(v + 1, Hs - v).clamp(0, 1)
    return (mx * my).unsqueeze(1)
