"""Geometry helpers with the reference's signatures (reference: src/data/utils.py:7-67).

torch inputs run on the sm_100a kernels (CUDA float32 only, no ATen fallback); numpy inputs keep the
reference's OpenCV twins, which only the CPU DataLoader path uses (src/data/transforms.py:571-573).
"""
import numpy as np
import torch

from .. import functional as F


def _is_torch(x):
    return torch.is_tensor(x)


def four_point_to_homography(corners, deltas, crop=False):
    """corners, deltas [B,4,2] -> H [B,3,3] mapping corners -> corners + deltas (reference utils.py:7-33).

    crop=True first moves corner 0 to the origin (utils.py:21-22).
    """
    assert len(corners.shape) == 3, 'corners should be of size B, 4, 2, but got: {}'.format(corners.shape)
    assert len(deltas.shape) == 3, 'deltas should be of size B, 4, 2, but got: {}'.format(deltas.shape)
    if _is_torch(corners):
        if crop:
            corners = corners - corners[:, 0].view(-1, 1, 2)
        return F.dlt4(deltas.to(corners.dtype), corners=corners)
    if isinstance(corners, np.ndarray):
        import cv2
        if crop:
            corners = corners - corners[:, 0].reshape(-1, 1, 2)
        return cv2.getPerspectiveTransform(np.float32(corners), np.float32(corners + deltas))
    assert False, 'Wrong type?'


def image_shape_to_corners(patch):
    """[B,C,H,W] -> [B,4,2] corners [[0,0],[a,0],[a,b],[0,b]], a = shape[-2], b = shape[-1]
    (the reference's width/height swap, utils.py:39-40, harmless for square patches)."""
    assert len(patch.shape) == 4, 'patch should be of size B, C, H, W'
    a, b = patch.shape[-2], patch.shape[-1]
    pts = [[0, 0], [a, 0], [a, b], [0, b]]
    if _is_torch(patch):
        c = torch.tensor(pts, device=patch.device, dtype=patch.dtype, requires_grad=False)
        return c.repeat(patch.shape[0], 1, 1)
    if isinstance(patch, np.ndarray):
        return np.tile(np.float32(pts)[None], (patch.shape[0], 1, 1))
    assert False, 'Wrong type?'


def warp_image(image, homography, target_h, target_w, inverse=True):
    """out[y,x] = bilinear(image, H [x,y,1]) for inverse=True (the only mode the reference's heads use,
    utils.py:54-59); inverse=False samples through H^-1."""
    if _is_torch(homography):
        if not inverse:
            homography = torch.inverse(homography)
        return F.warp(image, homography, target_h, target_w)
    if isinstance(homography, np.ndarray):
        import cv2
        if inverse:
            homography = np.linalg.inv(homography)
        return cv2.warpPerspective(image, homography, dsize=(target_w, target_h))
    assert False, 'Wrong type?'


def perspectiveTransformBatched(points, homography):
    """points [B,N,2], homography [B,3,3] -> projected points (reference utils.py:108-136, torch branch)."""
    assert points.dim() == 3 and points.shape[2] == 2, points.shape
    assert homography.shape[1:] == (3, 3), homography.shape
    ph = torch.nn.functional.pad(points, (0, 1), 'constant', 1.0)
    q = ph @ homography.transpose(1, 2)
    return q[:, :, :2] / q[:, :, 2:]
