"""Restatement of the kornia==0.5.0 functions used by the biHomE reference.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

kornia 0.5.0 is the reference's pinned dependency (``requirements.txt:1``); it is
not vendored under /root/reference and cannot run on torch>=2.0 because
``get_perspective_transform`` calls the removed ``torch.solve``.  The four
functions below restate the published 0.5.0 algorithms in plain torch so that
the reference's own modules can be executed (oracle/ref_import.py installs this
module as ``kornia``).  Reference call sites:

  get_perspective_transform  src/data/utils.py:24
  warp_perspective           src/data/utils.py:59
  transform_points           src/heads/PerceptualHead.py:175,201,765; src/heads/ransac_utils.py:90
  find_homography_dlt        src/heads/ransac_utils.py:72,143

Everything is dtype-generic: run it in float32 for a like-for-like comparison
and in float64 for the ground truth.
"""
import warnings

import torch
import torch.nn.functional as F

__version__ = '0.5.0-restated'


# --------------------------------------------------------------------------------------
# kornia/geometry/conversions.py
# --------------------------------------------------------------------------------------
def convert_points_to_homogeneous(points):
    return F.pad(points, [0, 1], 'constant', 1.0)


def convert_points_from_homogeneous(points, eps=1e-8):
    # 0.5.0: scale = 1/z where |z| > eps else 1
    z = points[..., -1:]
    mask = torch.abs(z) > eps
    scale = torch.where(mask, 1.0 / torch.where(mask, z, torch.ones_like(z)), torch.ones_like(z))
    return scale * points[..., :-1]


# --------------------------------------------------------------------------------------
# kornia/geometry/linalg.py
# --------------------------------------------------------------------------------------
def transform_points(trans_01, points_1):
    shape_inp = list(points_1.shape)
    points_1 = points_1.reshape(-1, points_1.shape[-2], points_1.shape[-1])
    trans_01 = trans_01.reshape(-1, trans_01.shape[-2], trans_01.shape[-1])
    trans_01 = torch.repeat_interleave(trans_01, repeats=points_1.shape[0] // trans_01.shape[0], dim=0)
    points_1_h = convert_points_to_homogeneous(points_1)
    points_0_h = torch.bmm(points_1_h, trans_01.permute(0, 2, 1))
    points_0 = convert_points_from_homogeneous(points_0_h)
    shape_inp[-2] = points_0.shape[-2]
    shape_inp[-1] = points_0.shape[-1]
    return points_0.reshape(shape_inp)


# --------------------------------------------------------------------------------------
# kornia/geometry/transform/imgwarp.py
# --------------------------------------------------------------------------------------
def _perspective_row(p, q, axis):
    ones = torch.ones_like(p)[..., 0:1]
    zeros = torch.zeros_like(p)[..., 0:1]
    if axis == 'x':
        return torch.cat([p[:, 0:1], p[:, 1:2], ones, zeros, zeros, zeros,
                          -p[:, 0:1] * q[:, 0:1], -p[:, 1:2] * q[:, 0:1]], dim=1)
    return torch.cat([zeros, zeros, zeros, p[:, 0:1], p[:, 1:2], ones,
                      -p[:, 0:1] * q[:, 1:2], -p[:, 1:2] * q[:, 1:2]], dim=1)


def get_perspective_transform(src, dst):
    """src, dst [B,4,2] -> [B,3,3] with h33 == 1 (LU with partial pivoting)."""
    rows = []
    for i in range(4):
        rows.append(_perspective_row(src[:, i], dst[:, i], 'x'))
        rows.append(_perspective_row(src[:, i], dst[:, i], 'y'))
    A = torch.stack(rows, dim=1)                                    # [B,8,8]
    b = torch.stack([dst[:, 0:1, 0], dst[:, 0:1, 1], dst[:, 1:2, 0], dst[:, 1:2, 1],
                     dst[:, 2:3, 0], dst[:, 2:3, 1], dst[:, 3:4, 0], dst[:, 3:4, 1]], dim=1)   # [B,8,1]
    # 0.5.0: X, LU = torch.solve(b, A).  torch.linalg.solve is the same LAPACK gesv.
    X = torch.linalg.solve(A, b)
    M = torch.ones(src.shape[0], 9, device=src.device, dtype=src.dtype)
    M[..., :8] = torch.squeeze(X, dim=-1)
    return M.view(-1, 3, 3)


def _torch_inverse_cast(x):
    dtype = x.dtype
    if dtype not in (torch.float32, torch.float64):
        dtype = torch.float32
    return torch.inverse(x.to(dtype)).to(x.dtype)


def normal_transform_pixel(height, width):
    tr = torch.tensor([[1.0, 0.0, -1.0], [0.0, 1.0, -1.0], [0.0, 0.0, 1.0]], dtype=torch.float64)
    w_den = 1e-14 if width == 1 else width - 1.0
    h_den = 1e-14 if height == 1 else height - 1.0
    tr[0, 0] = tr[0, 0] * 2.0 / w_den
    tr[1, 1] = tr[1, 1] * 2.0 / h_den
    return tr.unsqueeze(0)


def normalize_homography(dst_pix_trans_src_pix, dsize_src, dsize_dst):
    src_h, src_w = dsize_src
    dst_h, dst_w = dsize_dst
    src_norm_trans_src_pix = normal_transform_pixel(src_h, src_w).to(dst_pix_trans_src_pix)
    src_pix_trans_src_norm = _torch_inverse_cast(src_norm_trans_src_pix)
    dst_norm_trans_dst_pix = normal_transform_pixel(dst_h, dst_w).to(dst_pix_trans_src_pix)
    return dst_norm_trans_dst_pix @ (dst_pix_trans_src_pix @ src_pix_trans_src_norm)


def create_meshgrid(height, width, normalized_coordinates=True, device=None):
    xs = torch.linspace(0, width - 1, width, device=device, dtype=torch.float)
    ys = torch.linspace(0, height - 1, height, device=device, dtype=torch.float)
    if normalized_coordinates:
        xs = (xs / (width - 1) - 0.5) * 2
        ys = (ys / (height - 1) - 0.5) * 2
    base = torch.stack(torch.meshgrid([xs, ys], indexing='ij')).transpose(1, 2)   # 2 x H x W
    return torch.unsqueeze(base, dim=0).permute(0, 2, 3, 1)                       # 1 x H x W x 2


def warp_perspective(src, M, dsize, mode='bilinear', padding_mode='zeros', align_corners=None):
    if align_corners is None:
        # 0.5.0 warns here on every call and falls back to True; the reference never
        # passes the flag (src/data/utils.py:59) so True is the behaviour to match.
        align_corners = True
    B, C, H, W = src.size()
    h_out, w_out = dsize
    dst_norm_trans_src_norm = normalize_homography(M, (H, W), (h_out, w_out))
    src_norm_trans_dst_norm = _torch_inverse_cast(dst_norm_trans_src_norm)
    grid = create_meshgrid(h_out, w_out, normalized_coordinates=True, device=src.device).to(src.dtype)
    grid = grid.repeat(B, 1, 1, 1)
    grid = transform_points(src_norm_trans_dst_norm[:, None, None], grid)
    return F.grid_sample(src, grid, align_corners=align_corners, mode=mode, padding_mode=padding_mode)


# --------------------------------------------------------------------------------------
# kornia/geometry/epipolar/fundamental.py + kornia/geometry/homography.py
# --------------------------------------------------------------------------------------
def normalize_points(points, eps=1e-8):
    x_mean = torch.mean(points, dim=1, keepdim=True)
    scale = (points - x_mean).norm(dim=-1).mean(dim=-1)
    scale = torch.sqrt(torch.tensor(2.0)) / (scale + eps)
    ones, zeros = torch.ones_like(scale), torch.zeros_like(scale)
    transform = torch.stack([scale, zeros, -scale * x_mean[..., 0, 0],
                             zeros, scale, -scale * x_mean[..., 0, 1],
                             zeros, zeros, ones], dim=-1).view(-1, 3, 3)
    return transform_points(transform, points), transform


def find_homography_dlt(points1, points2, weights=None):
    assert points1.shape == points2.shape, points1.shape
    eps = 1e-8
    points1_norm, transform1 = normalize_points(points1)
    points2_norm, transform2 = normalize_points(points2)
    x1, y1 = torch.chunk(points1_norm, dim=-1, chunks=2)
    x2, y2 = torch.chunk(points2_norm, dim=-1, chunks=2)
    ones, zeros = torch.ones_like(x1), torch.zeros_like(x1)
    ax = torch.cat([zeros, zeros, zeros, -x1, -y1, -ones, y2 * x1, y2 * y1, y2], dim=-1)
    ay = torch.cat([x1, y1, ones, zeros, zeros, zeros, -x2 * x1, -x2 * y1, -x2], dim=-1)
    A = torch.cat((ax, ay), dim=-1).reshape(ax.shape[0], -1, ax.shape[-1])
    if weights is None:
        A = A.transpose(-2, -1) @ A
    else:
        w_diag = torch.diag_embed(weights.unsqueeze(dim=-1).repeat(1, 1, 2).reshape(weights.shape[0], -1))
        A = A.transpose(-2, -1) @ w_diag @ A
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            U, S, V = torch.svd(A)
    except Exception:   # noqa: BLE001 - 0.5.0 swallows everything here
        warnings.warn('SVD did not converge')
        return torch.empty((points1_norm.size(0), 3, 3), device=points1.device)
    H = V[..., -1].view(-1, 3, 3)
    H = transform2.inverse() @ (H @ transform1)
    return H / (H[..., -1:, -1:] + eps)
