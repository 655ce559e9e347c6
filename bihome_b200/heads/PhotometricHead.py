"""Photometric head (Nguyen et al., reference ``src/heads/PhotometricHead.py``): warps the second image with the
predicted homography and returns the patch at the ground-truth corners for an external pixel loss.
Same kwargs (LEARNING_KEYS = patch_gt, image, delta_gt, delta_hat), same return tuple.

The reference warps the whole image and then slices ``corners[0] .. corners[2]`` out of it sample by sample
(:27-39).  Here the crop offset is folded into the homography, ``H' = H * T(corner_0)``, and K2 renders the P x P
patch directly: the same samples of the same bilinear surface, 1 / (image area / patch area) of the work and no
per-sample Python loop.  Gradients reach ``delta_hat`` through K2b and the K1 adjoint.
"""
import torch
import torch.nn as nn

from .. import functional as F
from ..data.utils import four_point_to_homography


class Model(nn.Module):

    def __init__(self, backbone, **kwargs):
        super().__init__()
        self.learning_keys = kwargs['LEARNING_KEYS']

    def forward(self, data):
        assert 'corners' in data, 'Check this twice!'
        corners = data['corners']
        image = data[self.learning_keys[1]]
        delta_hat = data[self.learning_keys[3]]
        homography_hat = four_point_to_homography(corners=corners, deltas=delta_hat, crop=False)
        c = corners.int()
        patch_gt = data[self.learning_keys[0]]
        out_h, out_w = patch_gt.shape[-2], patch_gt.shape[-1]
        # all crops of a batch have the patch's size (the reference stacks them, :39), so one launch renders them
        shift = torch.zeros(corners.shape[0], 3, 3, device=corners.device, dtype=homography_hat.dtype)
        shift[:, 0, 0] = shift[:, 1, 1] = shift[:, 2, 2] = 1
        shift[:, 0, 2] = c[:, 0, 0].to(shift.dtype)
        shift[:, 1, 2] = c[:, 0, 1].to(shift.dtype)
        patch_hat = F.warp(image, torch.bmm(homography_hat, shift), out_h, out_w)
        return patch_gt, patch_hat, data[self.learning_keys[2]], delta_hat

    def predict_homography(self, data):
        assert 'corners' in data, 'How to handle it?'
        delta_hat = data[self.learning_keys[3]]
        return delta_hat, four_point_to_homography(corners=data['corners'], deltas=delta_hat, crop=False)
