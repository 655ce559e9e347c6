"""Content-aware triplet head (Zhang et al., reference ``src/heads/TripletHead.py``), the loss of the shipped
``zhang-orig`` configs: full-resolution one-channel features of the backbone's own (trainable) feature extractor,
the backbone's masks, one- or double-line margin loss.  Same kwargs, ``forward(data) -> (loss, delta_gt, delta_hat)``
and ``predict_homography`` as the reference.

On the B200 path both directions go through K1 (4-point DLT) and K2 (warp) in one launch each; when the backbone's
masks are the constant ones of ``FIX_MASK`` the warped masks come out of K2's analytic coverage computation instead
of two more full warps.  The margin algebra on the [B,1,P,P] maps, ln3 and every gradient come out of K3g
(``F.triplet_loss``) in one call.

Reference quirk kept on purpose: with a numeric margin and 'channel-agnostic' aggregation the
reference clamps ``[B,H,W]`` against ``zeros_like`` of a ``[B,1,H,W]`` tensor (:91-92,134-135); the broadcast makes a
``[B,B,H,W]`` tensor whose rows are all equal, so its loss is B times the per-sample sum (and needs B^2 maps of
memory).  Same value here, without the B^2 tensor.
"""
import torch

from .. import functional as F
from ..data.utils import four_point_to_homography, image_shape_to_corners, warp_image


class Model(torch.nn.Module):

    def __init__(self, backbone, **kwargs):
        super().__init__()
        self.backbone = backbone
        self.patch_keys = kwargs['PATCH_KEYS']
        self.mask_keys = kwargs['MASK_KEYS']
        self.feature_keys = kwargs['FEATURE_KEYS']
        self.target_keys = kwargs['TARGET_KEYS']
        self.ld = kwargs['LD']
        self.mu = kwargs['MU']
        assert self.ld == 2, 'Only ld==2 is supported at the moment'
        self.variant = str.lower(kwargs['VARIANT'])
        assert self.variant == 'oneline' or self.variant == 'doubleline', 'Supported variants: OneLine or DoubleLine'
        self.triplet_margin = kwargs['TRIPLET_MARGIN']
        self.triplet_channel_aggregation = kwargs['TRIPLET_AGGREGATION']
        assert self.triplet_channel_aggregation in ('channel-aware', 'channel-agnostic'), \
            'Do not know this aggregation technique'

    @staticmethod
    def _warp(image, delta_hat, corners=None):
        if corners is None:
            corners = image_shape_to_corners(patch=image)
        homography = four_point_to_homography(corners=corners, deltas=delta_hat, crop=False)
        return warp_image(image, homography, target_h=image.shape[-2], target_w=image.shape[-1]), homography

    def _constant_masks(self):
        predictor = getattr(self.backbone, 'mask_predictor', None)
        return predictor is not None and bool(getattr(predictor, 'fix_mask', False))

    def _loss_variant(self, b, c):
        """(hinge, margin, scale) of F.triplet_loss for TRIPLET_MARGIN / TRIPLET_AGGREGATION (reference :79-95)"""
        if isinstance(self.triplet_margin, str):
            return None, 0.0, 1.0
        if self.triplet_channel_aggregation == 'channel-aware':
            return 'channel', float(self.triplet_margin), 1.0
        if c != 1 and b != 1:
            raise RuntimeError('TripletHead: a numeric margin with channel-agnostic aggregation is only defined for '
                               'one-channel features (the reference broadcast fails for B = %d, C = %d)' % (b, c))
        return 'pixel', float(self.triplet_margin), float(b if c == 1 else c)

    def forward(self, data):
        e1, e2 = self.patch_keys
        k1, k2 = self.mask_keys
        q1, q2 = self.feature_keys
        o1, o2 = self.target_keys
        patch_1, patch_2 = data[e1], data[e2]
        m1, m2 = data[k1], data[k2]
        f1, f2 = data[q1], data[q2]
        double = self.variant == 'doubleline'
        B = patch_1.shape[0]
        size = (patch_1.shape[-2], patch_1.shape[-1])
        out_h, out_w = size

        if double:
            H = F.dlt4(torch.cat([data[o1], data[o2]], dim=0), size=size)
            src = torch.cat([patch_1, patch_2], dim=0)
            masks = None if self._constant_masks() else torch.cat([m1, m2], dim=0)
        else:
            H = F.dlt4(data[o1], size=size)
            src = patch_1
            masks = None if self._constant_masks() else m1
        if masks is None:
            warped, mw = F.warp(src, H, out_h, out_w, pool=1)
        else:
            warped = F.warp(src, H, out_h, out_w)
            mw = F.warp(masks, H, out_h, out_w).squeeze(1)
        h1 = H[:B]
        h2 = H[B:] if double else None
        f1w = self.backbone.feature_extractor(warped[:B])
        f2w = self.backbone.feature_extractor(warped[B:]) if double else None
        hinge, margin, scale = self._loss_variant(B, f1.shape[1])
        # both lines, ln3 and every gradient (features, masks, homographies) in one K3g call (reference :78-153)
        loss_b, parts = F.triplet_loss(f1, f2, f1w, f2w, mw[:B], m2.squeeze(1), mw[B:] if double else None,
                                       m1.squeeze(1) if double else None, h1, h2, lines=2 if double else 1, distance='l1',
                                       hinge=hinge, margin=margin, mu=self.mu if double else 0.0, scale=(scale, scale))
        loss = loss_b.sum()

        if 'summary_writer' in data:
            step, sw = data['summary_writer_step'], data['summary_writer']
            sw.add_scalars('feature_space', {'patch_2_f': f2.mean().item()}, step)
            sw.add_scalars('feature_space', {'patch_1_f_prime': f1w.mean().item()}, step)
            sw.add_scalars('feature_space', {'patch_1_f': f1.mean().item()}, step)
            sw.add_scalars('loss_comp', {'l1': (f2 - f1w).abs().mean().item()}, step)
            sw.add_scalars('loss_comp', {'l3': (f1 - f2).abs().mean().item()}, step)
            eye = torch.eye(3, dtype=h1.dtype, device=h1.device).unsqueeze(0)
            sw.add_scalars('h', {'h1': ((h1 - eye) ** 2).sum().item()}, step)
            if double:
                sw.add_scalars('feature_space', {'patch_2_f_prime': f2w.mean().item()}, step)
                sw.add_scalars('loss_comp', {'l2': (f1 - f2w).abs().mean().item()}, step)
                sw.add_scalars('loss_comp', {'ln1': parts[:, 0].sum().item()}, step)
                sw.add_scalars('loss_comp', {'ln2': parts[:, 1].sum().item()}, step)
                sw.add_scalars('loss_comp', {'ln3': self.mu * parts[:, 4].sum().item()}, step)
                sw.add_scalars('h', {'h2': ((h2 - eye) ** 2).sum().item()}, step)

        delta_gt = data['delta'] if 'delta' in data else None
        delta_hat = data[o1] if o1 in data else None
        return loss, delta_gt, delta_hat

    def predict_homography(self, data):
        delta_hat = data[self.target_keys[0]]
        homography = F.dlt4(delta_hat, size=(data[self.patch_keys[0]].shape[-2], data[self.patch_keys[0]].shape[-1]))
        return delta_hat, homography
