"""DSAC-style hypothesis sampling for the perspective-field (Zeng) configs.

Mirror of the reference's ``src/heads/ransac_utils.py`` (DSACSoftmax :26-160): same constructor kwargs, same
``forward(points1, points2, points_per_hypothesis, hypothesis_no) -> (homographies [B,n,3,3], scores [B,n])``.

  * sampling  : ``torch.multinomial(arange(N), B*M*n, replacement=True)`` -- index-proportional weights, the
                reference's quirk (ransac_utils.py:55-56), drawn with the same call so that a seeded torch RNG
                yields the same correspondences;
  * solve     : N-point normalised DLT (kornia.find_homography_dlt, reference :72) on the K4 kernel
                (bihome_b200.functional.dltn), gather fused;
  * scoring   : 'repr_error' / 'inliers_ratio' / 'soft_inliers_ratio' / 'score_cnn' softmax (:97-126).  With one
                hypothesis the softmax is identically 1 and the 16384-point reprojection pass is skipped (bit-identical
                output).
"""
import math

import torch
import torch.nn as nn
import torchvision.models as models

from .. import functional as F


class ScoreCNN(nn.Module):
    """ResNet-18 with a 2-channel stem and one output: scores a hypothesis from the [2, sqrt(N), sqrt(N)] image of its
    signed reprojection errors (reference ransac_utils.py:10-23; same attribute names, so checkpoints interchange)."""

    def __init__(self, pretrained):
        super().__init__()
        weights = None
        if pretrained:
            try:
                self.resnet18 = models.resnet18(weights='DEFAULT', progress=True)
            except Exception:  # noqa: BLE001 -- no network / no cached checkpoint
                self.resnet18 = models.resnet18(weights=weights)
        else:
            self.resnet18 = models.resnet18(weights=weights)
        self.resnet18.conv1 = nn.Conv2d(2, 64, kernel_size=(7, 7), stride=(2, 2), padding=(3, 3), bias=False)
        self.resnet18.fc = nn.Linear(512, 1, bias=True)

    def forward(self, x):
        return self.resnet18(x)


def sample_choice(n_points, count, device):
    """The reference's draw (ransac_utils.py:54-56)."""
    weights = torch.arange(start=0, end=n_points, dtype=torch.float32, device=device)
    return torch.multinomial(weights, count, replacement=True)


def transform_points(H, pts):
    """kornia.transform_points semantics for H [B,3,3], pts [B,N,2] (eps = 1e-8 guard on the homogeneous scale)."""
    ph = torch.nn.functional.pad(pts, (0, 1), 'constant', 1.0)
    q = torch.bmm(ph, H.transpose(1, 2))
    z = q[..., 2:]
    ok = z.abs() > 1e-8
    scale = torch.where(ok, 1.0 / torch.where(ok, z, torch.ones_like(z)), torch.ones_like(z))
    return q[..., :2] * scale


class DSACSoftmax(torch.nn.Module):

    def __init__(self, **kwargs):
        super().__init__()
        self.scoring_method = kwargs['SCORING_METHOD'] if 'SCORING_METHOD' in kwargs else 'repr_error'
        if self.scoring_method == 'inliers_ratio':
            self.scoring_distance_threshold = kwargs['SCORING_DISTANCE_THRESHOLD']
        if self.scoring_method == 'soft_inliers_ratio':
            self.scoring_distance_beta = kwargs['SCORING_DISTANCE_BETA']
            self.scoring_distance_threshold = kwargs['SCORING_DISTANCE_THRESHOLD']
        if self.scoring_method == 'score_cnn':
            self.score_cnn = ScoreCNN(kwargs['SCORE_CNN_PRETRAINED'])

    def sample_hypotheses(self, points1, points2, points_per_hypothesis, hypothesis_no, choice=None):
        B, N = points1.shape[0], points1.shape[1]
        if choice is None:
            choice = sample_choice(N, B * points_per_hypothesis * hypothesis_no, points1.device)
        choice = choice.reshape(B * hypothesis_no, points_per_hypothesis)
        if hypothesis_no != 1:
            points1 = points1.repeat_interleave(hypothesis_no, dim=0)
            points2 = points2.repeat_interleave(hypothesis_no, dim=0)
        H = F.dltn(points1, points2, choice)
        return H.reshape(B, hypothesis_no, 3, 3)

    def score_hypotheses(self, points1, points2, homographies):
        B, n = homographies.shape[:2]
        if n == 1 and self.scoring_method in ('repr_error', 'inliers_ratio', 'soft_inliers_ratio'):
            return torch.ones(B, 1, device=homographies.device, dtype=homographies.dtype)
        p1 = points1.repeat_interleave(n, dim=0)
        p2 = points2.repeat_interleave(n, dim=0)
        proj = transform_points(homographies.reshape(B * n, 3, 3), p1)
        if self.scoring_method == 'repr_error':
            scores = (proj - p2).abs().sum(-1).sum(-1)
        elif self.scoring_method == 'inliers_ratio':
            err = torch.norm(proj - p2, dim=-1)
            scores = (err < self.scoring_distance_threshold).float().mean(-1)
        elif self.scoring_method == 'soft_inliers_ratio':
            err = torch.norm(proj - p2, dim=-1)
            scores = torch.sigmoid(self.scoring_distance_beta * (err - self.scoring_distance_threshold)).sum(-1)
        elif self.scoring_method == 'score_cnn':
            # the signed reprojection errors of all N = side^2 points as a 2-channel image (reference :113-121)
            side = int(math.sqrt(p1.shape[1]))
            scores = self.score_cnn((proj - p2).permute(0, 2, 1).reshape(B * n, 2, side, side))
        else:
            assert False, 'I do not know this scoring method'
        return torch.softmax(-scores.reshape(B, n), dim=-1)

    def forward(self, points1, points2, points_per_hypothesis=4, hypothesis_no=128, choice=None):
        homographies = self.sample_hypotheses(points1, points2, points_per_hypothesis, hypothesis_no, choice)
        scores = self.score_hypotheses(points1, points2, homographies)
        return homographies, scores
