"""Oracle-backed stand-ins for the CUDA ops of ``bihome_b200.functional`` -- for the CPU tests of the HOST LOGIC only.

The product has no CPU path (every op raises on a non-CUDA tensor).  To check on a machine without a GPU that the
Python mirrors of the reference's heads wire the ops together the way the reference does, the tests patch the
entry points the mirrors call with the oracle's closed forms (float64 capable, autograd through plain torch ops).
Nothing outside tests/ imports this module.
"""
import torch

from oracle import ref_path as R


def dlt4(delta, corners=None, size=None):
    if corners is None:
        w, h = size
        corners = torch.tensor([[0, 0], [w, 0], [w, h], [0, h]], dtype=delta.dtype).repeat(delta.shape[0], 1, 1)
    return R.four_point_to_homography(corners.to(delta.dtype), delta)


def coverage_mask(H, src_hw, out_hw, pool=1):
    m = R.analytic_mask(H, src_hw[0], src_hw[1], out_hw[0], out_hw[1])
    return torch.nn.functional.avg_pool2d(m, pool).squeeze(1)


def warp(src, H, out_h, out_w, pool=None):
    if tuple(src.shape[-2:]) == (out_h, out_w):
        # like for like with the golden vectors: the kornia route of the reference, whose sampling grid is born in
        # float32 even in a float64 evaluation (1e-8 px of coordinate noise, enough to flip bilinear cells)
        out = R.warp_image(src, H, out_h, out_w)
    else:
        out = R.warp_direct(src, H, out_h, out_w)
    if pool:
        return out, coverage_mask(H, src.shape[-2:], (out_h, out_w), pool)
    return out


def bihome_loss(f1, f2, f1w, f2w, m1w, m2w, H12, H21, mu, m1=None, m2=None):
    ones = torch.ones_like(m1w)
    u = lambda t: (ones if t is None else t).unsqueeze(1)
    _, p = R.bihome_double_line(f1, f2, f1w, f2w, u(m1), u(m2), u(m1w), u(m2w), H12, H21, mu)
    loss_b = p['ln1'] + p['ln2'] + mu * p['ln3']
    parts = torch.stack([p['ln1'], p['ln2'], p['den1'], p['den2'], p['ln3']], dim=1).detach()
    return loss_b, parts


def triplet_loss(f1, f2, f1w, f2w, a1, b2, a2=None, b1=None, H12=None, H21=None, lines=2, distance='l1', hinge=None, margin=0.0,
                 mask_crd=False, mu=0.0, scale=(1.0, 1.0)):
    assert not isinstance(margin, (tuple, list))
    sq = lambda t: None if t is None else t.reshape(t.shape[0], t.shape[-2], t.shape[-1])
    loss_b, p = R.triplet_general(f1, f2, f1w, f2w, sq(a1), sq(b2), sq(a2), sq(b1), H12, H21, lines, distance, hinge, margin,
                                  mask_crd=mask_crd, mu=mu, scale=scale)
    parts = torch.stack([p['ln1'], p['ln2'], p['den1'], p['den2'], p['ln3']], dim=1).detach()
    return loss_b, parts


def dltn(points1, points2, choice=None):
    from oracle import kornia050 as K
    if choice is not None:
        idx = choice.reshape(points1.shape[0], -1, 1).repeat(1, 1, 2)
        points1, points2 = torch.gather(points1, 1, idx), torch.gather(points2, 1, idx)
    return K.find_homography_dlt(points1, points2)


def dltn_field(field, choice, four_points):
    delta, H, _ = R.zeng_delta_hat(field, choice.shape[1], 1, choice.reshape(-1))
    return H.reshape(-1, 3, 3), delta.reshape(-1, 4, 2)


def mace(delta_gt, delta_hat):
    return (delta_gt.reshape(-1, 2) - delta_hat.reshape(-1, 2)).norm(dim=-1).mean()


# ---- K6 (field head): the four device entry points as plain torch ops on [N, C] views ---------------------------------
def _rows(x):
    return x.permute(0, 2, 3, 1).reshape(-1, x.shape[1])


def fh_moments(x):
    X = _rows(x).double()
    return X.sum(0), X.t() @ X


def fh_fwd(x, W1, b1, W2, b2):
    B, C, H, W = x.shape
    out = torch.relu(_rows(x) @ W1.t() + b1) @ W2.t() + b2
    return out.reshape(B, H, W, 2).permute(0, 3, 1, 2).contiguous()


def fh_bwd(x, W1, b1, W2, g_out):
    B, C, H, W = x.shape
    X = _rows(x)
    pre = X @ W1.t() + b1
    h = torch.relu(pre)
    G = g_out.permute(0, 2, 3, 1).reshape(-1, 2)
    gh = (G @ W2) * (pre > 0).to(X.dtype)
    gx = torch.empty_like(x)
    gx.copy_((gh @ W1).reshape(B, H, W, C).permute(0, 3, 1, 2))
    return gx, gh.t() @ X, gh.sum(0), G.t() @ h, G.sum(0)


def fh_affine(x, a, M, gx):
    B, C, H, W = x.shape
    gx.add_((a + _rows(x) @ M.t()).reshape(B, H, W, C).permute(0, 3, 1, 2))
    return gx


def install(monkeypatch):
    import bihome_b200.functional as F
    for name, fn in (('dlt4', dlt4), ('warp', warp), ('coverage_mask', coverage_mask), ('bihome_loss', bihome_loss), ('triplet_loss', triplet_loss),
                     ('dltn', dltn), ('dltn_field', dltn_field), ('mace', mace), ('_fh_moments', fh_moments), ('_fh_fwd', fh_fwd),
                     ('_fh_bwd', fh_bwd), ('_fh_affine', fh_affine)):
        monkeypatch.setattr(F, name, fn)
    return F
