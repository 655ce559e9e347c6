"""Perceptual (biHomE) head -- same constructor kwargs, methods and return tuples as the reference's
``src/heads/PerceptualHead.py`` (Model :79-767, AuxiliaryResnet :15-76), computed on the sm_100a kernels.

Hot path of the shipped ``*-bihome-*`` configs (TRIPLET_LOSS 'double-line', TRIPLET_MARGIN 'inf', l1):

    delta_hat --K1 dlt4--> H --K2 warp (+ analytic pooled mask)--> p' --frozen ResNet stem (cuDNN)--> f'
    (f1, f2, f1', f2', masks, H12, H21) --K3 fused loss fwd+bwd--> loss

Both directions are batched through K1/K2 in one launch each (2B samples), the coverage masks never exist at
full resolution, and the loss kernel reads every feature once.  Zeng configs (PF_KEYS) get their delta_hat from
K4 (dltn_field) after the reference's index-proportional multinomial draw.  Everything else the reference's
head offers (one-line, numeric margins, l2/cosine, 'dual', MASK_CRD, user masks, up-sampling strategies) runs on the same
K1/K2 warps with the loss algebra and its gradients in K3g (F.triplet_loss); the multihead loss returns features.
"""
import warnings

import numpy as np
import torch
import torch.nn as nn
import torchvision.models as models

from .. import functional as F
from ..data.utils import four_point_to_homography, image_shape_to_corners, warp_image
from .ransac_utils import DSACSoftmax, sample_choice, transform_points


def _make_resnet(name, pretrained):
    fn = getattr(models, name)
    if pretrained:
        try:
            return fn(weights='DEFAULT', progress=True)
        except Exception as e:  # noqa: BLE001 -- no network / no cached checkpoint
            warnings.warn('bihome_b200: pretrained %s weights unavailable (%s); using random init' % (name, e))
    return fn(weights=None)


def _tv_block(blk, x):
    """torchvision BasicBlock / Bottleneck forward (conv -> bn -> relu ... conv -> bn, += identity, relu) with the BatchNorm /
    residual / ReLU stages fused (K7b) where the tensors allow it; same modules, same parameters, same statistics updates"""
    identity = x if blk.downsample is None else blk.downsample(x)
    if isinstance(blk, models.resnet.Bottleneck):
        stages = [(blk.conv1, blk.bn1), (blk.conv2, blk.bn2), (blk.conv3, blk.bn3)]
    else:
        stages = [(blk.conv1, blk.bn1), (blk.conv2, blk.bn2)]
    h = x
    for k, (conv, bn) in enumerate(stages):
        h = conv(h)
        if k + 1 < len(stages):
            h = F.bn_relu(bn, h) if F.bnact_supported(bn, h) else nn.functional.relu(bn(h))
        elif F.bnact_supported(bn, h, identity):
            h = F.bn_relu(bn, h, residual=identity)
        else:
            h = nn.functional.relu(bn(h) + identity)
    return h


def _run_layer(layer, x):
    if not isinstance(layer, nn.Sequential):
        return layer(x)
    for blk in layer:
        x = _tv_block(blk, x) if isinstance(blk, (models.resnet.BasicBlock, models.resnet.Bottleneck)) else blk(x)
    return x


class AuxiliaryResnet(nn.Module):
    """Frozen torchvision ResNet stem + layer1..k used as the perceptual feature extractor (reference :15-76)."""

    def __init__(self, **kwargs):
        super().__init__()
        self.resnet = _make_resnet(kwargs['AUXILIARY_RESNET'], kwargs.get('AUXILIARY_RESNET_PRETRAINED', True))
        self.auxiliary_resnet_output_layer = kwargs['AUXILIARY_RESNET_OUTPUT_LAYER']
        for i in (2, 3, 4):
            if self.auxiliary_resnet_output_layer < i:
                setattr(self.resnet, 'layer%d' % i, nn.Identity())
        self.resnet.avgpool = nn.Identity()
        self.resnet.fc = nn.Identity()
        self.freeze = kwargs['AUXILIARY_RESNET_FREEZE'] if 'AUXILIARY_RESNET_FREEZE' in kwargs else True
        if self.freeze:
            for p in self.resnet.parameters():
                p.requires_grad = False
        self.with_projection_head = kwargs['WITH_PROJECTION_HEAD'] if 'WITH_PROJECTION_HEAD' in kwargs else None
        self.projection_head = nn.ModuleList()
        if self.with_projection_head is not None:
            for idx, layer in enumerate(self.with_projection_head):
                self.projection_head.append(nn.Linear(layer[0], layer[1]))
                if idx != len(self.with_projection_head) - 1:
                    self.projection_head.append(nn.ReLU())

    def forward(self, x):
        r = self.resnet
        if x.shape[1] == 1:
            # reference: x.repeat(1, 3, 1, 1) then conv1 (:52-55).  conv1 over three identical channels equals a
            # one-channel conv with the kernel summed over its input channels: a third of the bytes and FLOPs.
            c = r.conv1
            w = c.weight.sum(dim=1, keepdim=True)
            if c.weight.is_contiguous(memory_format=torch.channels_last) and not c.weight.is_contiguous():
                # keep the extractor in NHWC when the model was converted: a [64,1,k,k] tensor is layout-ambiguous,
                # explicit channels-last strides make cuDNN pick its NHWC kernels (and bn1/maxpool follow)
                k_h, k_w = w.shape[-2], w.shape[-1]
                w = w.as_strided(w.shape, (k_h * k_w, 1, k_w, 1))
                if x.is_cuda and x.is_contiguous():
                    # and the same for the one-channel input: ATen's cuDNN backward first lays the upstream gradient out in
                    # the INPUT's suggested format (NCHW for plain [B,1,H,W] strides) and then back in the weight's
                    # (channels-last) -- two strided copies of the [B,64,H/2,W/2] gradient per pass, 1.07 ms of the
                    # B = 256 step (profiles/r05b_copies.txt).  A view, no data moves.
                    x = x.view(x.shape[0], x.shape[2], x.shape[3], 1).permute(0, 3, 1, 2)
            x = nn.functional.conv2d(x, w, c.bias, c.stride, c.padding, c.dilation)
        else:
            x = r.conv1(x)
        # bn1 -> relu -> maxpool (reference :56-58) as one stage (K7) on channels-last tensors; the extractor's BatchNorm runs
        # on batch statistics like the reference's (the head never puts it in eval mode during training)
        x = F.stem(r.bn1, x) if F.stem_supported(r.bn1, r.maxpool, x) else r.maxpool(r.relu(r.bn1(x)))
        x = _run_layer(r.layer1, x)
        if self.auxiliary_resnet_output_layer > 1:
            x = _run_layer(r.layer2, x)
        if self.auxiliary_resnet_output_layer > 2:
            x = _run_layer(r.layer3, x)
        if self.auxiliary_resnet_output_layer > 3:
            x = _run_layer(r.layer4, x)
        if self.with_projection_head is not None:
            x = x.permute(0, 2, 3, 1)
            for layer in self.projection_head:
                x = layer(x)
            x = x.permute(0, 3, 1, 2)
        return x


def _stack(a, b):
    """[a; b] along the batch; free when the two tensors already sit back to back in one buffer."""
    if (a.is_contiguous() and b.is_contiguous() and a.shape == b.shape and not a.requires_grad and not b.requires_grad
            and a.untyped_storage().data_ptr() == b.untyped_storage().data_ptr()
            and b.storage_offset() == a.storage_offset() + a.numel()):
        return torch.as_strided(a, (2 * a.shape[0],) + tuple(a.shape[1:]), a.stride())
    return torch.cat([a, b], dim=0)


class Model(nn.Module):

    def __init__(self, backbone, **kwargs):
        super().__init__()
        self.backbone = backbone
        self.four_points_12 = None
        self.four_points_21 = None
        self.patch_size = kwargs['PATCH_SIZE']
        self.patch_keys = kwargs['PATCH_KEYS']
        self.delta_hat_keys = kwargs['DELTA_HAT_KEYS']
        if len(self.delta_hat_keys):
            self.hypothesis_no = 1
        else:
            self.coordinate_field_12 = None
            self.coordinate_field_21 = None
            self.pf_keys = kwargs['PF_KEYS']
            self.hypothesis_no = kwargs['RANSAC_HYPOTHESIS_NO']
            self.point_per_hypothesis = kwargs['POINTS_PER_HYPOTHESIS']
            self.dsac = DSACSoftmax(**kwargs)
        self.triplet_version = kwargs['TRIPLET_LOSS']
        if self.triplet_version != '':
            self.mask_keys = kwargs['MASK_KEYS']
            self.change_detection_mask = kwargs['MASK_CRD'] if 'MASK_CRD' in kwargs else False
            self.triplet_margin = kwargs['TRIPLET_MARGIN']
            self.triplet_channel_aggregation = kwargs['TRIPLET_AGGREGATION']
            self.sampling_strategy = kwargs['SAMPLING_STRATEGY']
            self.triplet_distance = kwargs['TRIPLET_DISTANCE']
            if 'one-line' in self.triplet_version:
                self.triplet_loss = nn.TripletMarginLoss(margin=self.triplet_margin, p=1, reduction='none')
            elif 'double-line' in self.triplet_version:
                self.triplet_mu = kwargs['TRIPLET_MU']
        self.auxiliary_resnet = AuxiliaryResnet(**kwargs)
        # parity hook: tests inject the multinomial draw here ([choice_12, choice_21]); None = draw like the reference
        self.forced_choice = None
        self.last_parts = None

    # ------------------------------------------------------------------------------------------------
    # perspective field -> correspondences (reference :125-146)
    # ------------------------------------------------------------------------------------------------
    def forward_map_field(self, perspective_field, self_coordinate_field, self_four_points):
        B, _, Hf, Wf = perspective_field.shape
        want = (B, Hf * Wf, 2)
        if self_coordinate_field is None or tuple(self_coordinate_field.shape) != want:
            ys, xs = torch.meshgrid(torch.arange(Hf, device=perspective_field.device, dtype=torch.float32),
                                    torch.arange(Wf, device=perspective_field.device, dtype=torch.float32), indexing='ij')
            self_coordinate_field = torch.stack((xs.reshape(-1), ys.reshape(-1)), dim=-1).unsqueeze(0).repeat(B, 1, 1)
            four = torch.tensor([[0, 0], [Wf, 0], [Wf, Hf], [0, Hf]], device=perspective_field.device, dtype=torch.float32)
            self_four_points = four.unsqueeze(0).repeat(B * self.hypothesis_no, 1, 1)
        field = perspective_field.reshape(B, 2, -1).permute(0, 2, 1)
        return self_coordinate_field + field, self_coordinate_field, self_four_points

    def _corner_points(self, Wf, Hf, device):
        """[[0,0],[Wf,0],[Wf,Hf],[0,Hf]] on the device, built once per geometry: a host-to-device copy per step would
        synchronise the host and cannot be captured in a CUDA graph (engine.GraphedStep)"""
        key = (int(Wf), int(Hf), str(device))
        cache = self.__dict__.setdefault('_corner_cache', {})
        if key not in cache:
            cache[key] = torch.tensor([[0, 0], [Wf, 0], [Wf, Hf], [0, Hf]], device=device, dtype=torch.float32)
        return cache[key]

    def _field_to_delta(self, pf, which):
        """DSAC branch of the reference's forward (:164-178 / :187-205) -> (delta_hats [B,n,4,2], scores or None)."""
        B, _, Hf, Wf = pf.shape
        n, M = self.hypothesis_no, self.point_per_hypothesis
        choice = None
        if self.forced_choice is not None:
            choice = self.forced_choice[which]
        if n == 1 and self.dsac.scoring_method in ('repr_error', 'inliers_ratio', 'soft_inliers_ratio'):
            # one hypothesis: softmax score == 1, nothing but H is needed -> fused K4 straight from the field
            if choice is None:
                choice = sample_choice(Hf * Wf, B * M, pf.device)
            four = self._corner_points(Wf, Hf, pf.device)
            _, delta = F.dltn_field(pf, choice.reshape(B, M), four)
            return delta.reshape(B, 1, 4, 2), torch.ones(B, 1, device=pf.device, dtype=pf.dtype)
        cf = self.coordinate_field_12 if which == 0 else self.coordinate_field_21
        fp = self.four_points_12 if which == 0 else self.four_points_21
        map_field, cf, fp = self.forward_map_field(pf, cf, fp)
        if which == 0:
            self.coordinate_field_12, self.four_points_12 = cf, fp
        else:
            self.coordinate_field_21, self.four_points_21 = cf, fp
        H, scores = self.dsac(cf, map_field, hypothesis_no=n, points_per_hypothesis=M, choice=choice)
        proj = transform_points(H.reshape(-1, 3, 3), fp)
        return (proj - fp).reshape(B, n, 4, 2), scores

    def forward(self, data):
        delta_hats_21 = None
        if not len(self.delta_hat_keys):
            delta_hats_12, scores_12 = self._field_to_delta(data[self.pf_keys[0]], 0)
            if 'double-line' in self.triplet_version:
                delta_hats_21, _ = self._field_to_delta(data[self.pf_keys[1]], 1)
        else:
            delta_hats_12 = data[self.delta_hat_keys[0]]
            scores_12 = None
            if 'double-line' in self.triplet_version:
                delta_hats_21 = data[self.delta_hat_keys[1]]
        if 'one-line' in self.triplet_version:
            return self.triplet_resnet_loss(data, delta_hats_12, scores=scores_12)
        elif 'double-line' in self.triplet_version:
            return self.triplet_resnet_loss(data, delta_hats_12, delta_hats_21=delta_hats_21)
        return self.multihead_resnet_loss(data, delta_hats_12, scores=scores_12)

    # ------------------------------------------------------------------------------------------------
    @staticmethod
    def _warp(image, delta_hat, corners=None):
        """(image_warped, homography) -- reference :237-243."""
        if corners is None:
            homography = F.dlt4(delta_hat, size=(image.shape[-2], image.shape[-1]))
        else:
            homography = four_point_to_homography(corners=corners, deltas=delta_hat, crop=False)
        image_warped = warp_image(image, homography, target_h=image.shape[-2], target_w=image.shape[-1])
        return image_warped, homography

    def _upsample(self, img, scale_factor):
        return nn.functional.interpolate(img, scale_factor=scale_factor, mode='bilinear', align_corners=True)

    def _features(self, patch):
        if self.sampling_strategy == 'upsample-patch-4x':
            patch = self._upsample(patch, 4)
        elif self.sampling_strategy == 'upsample-patch-2x':
            patch = self._upsample(patch, 2)
        return self.auxiliary_resnet(patch)

    def _fused_path(self):
        return ('double-line' in self.triplet_version and 'dual' not in self.triplet_version
                and self.triplet_distance == 'l1' and isinstance(self.triplet_margin, str)
                and self.triplet_channel_aggregation in ('channel-agnostic', 'channel-aware')
                and getattr(self.auxiliary_resnet, 'with_projection_head', None) is None)

    # ------------------------------------------------------------------------------------------------
    def multihead_resnet_loss(self, data, delta_hats, scores=None):
        """iHomE-style: returns features for an external nn loss (reference :245-315)."""
        e1, e2 = self.patch_keys
        b, n, i = delta_hats.shape[0], self.hypothesis_no, self.patch_size
        patch_1 = data[e1].reshape(b, 1, i, i).repeat(1, n, 1, 1).reshape(b * n, 1, i, i)
        patch_2 = data[e2].reshape(b, 1, i, i).repeat(1, n, 1, 1).reshape(b * n, 1, i, i)
        patch_2_f = self.auxiliary_resnet(patch_2)
        delta_hats = delta_hats.reshape(b * n, 4, 2)
        patch_1_prime, h1 = self._warp(patch_1, delta_hat=delta_hats)
        patch_1_f_prime = self.auxiliary_resnet(patch_1_prime)
        if scores is not None:
            s = scores.reshape(b * n, 1, 1, 1)
            patch_1_f_prime = patch_1_f_prime * s
            patch_2_f = patch_2_f * s
        if 'summary_writer' in data:
            step, sw = data['summary_writer_step'], data['summary_writer']
            sw.add_scalars('feature_space', {'patch_2_f': patch_2_f.mean().item()}, step)
            sw.add_scalars('feature_space', {'patch_1_f_prime': patch_1_f_prime.mean().item()}, step)
            sw.add_scalars('loss_comp', {'l1': (patch_2_f - patch_1_f_prime).abs().mean().item()}, step)
            eye = torch.eye(3, dtype=h1.dtype, device=h1.device).unsqueeze(0)
            sw.add_scalars('h', {'h1': ((h1 - eye) ** 2).sum().item()}, step)
        delta_gt = data['delta'] if 'delta' in data else None
        if scores is not None:
            delta_hats = (delta_hats * scores.reshape(b * n, 1, 1)).reshape(b, n, 4, 2).sum(dim=1)
        return patch_2_f, patch_1_f_prime, delta_gt, delta_hats

    # ------------------------------------------------------------------------------------------------
    def triplet_resnet_loss(self, data, delta_hats, delta_hats_21=None, scores=None):
        assert (delta_hats_21 is not None and scores is None) or delta_hats_21 is None, \
            'They should not be on at the same time - at least its not implemented yet'
        e1, e2 = self.patch_keys
        b, n, i = delta_hats.shape[0], self.hypothesis_no, self.patch_size
        patch_1 = data[e1].reshape(b, 1, i, i)
        patch_2 = data[e2].reshape(b, 1, i, i)
        if n != 1:
            patch_1 = patch_1.repeat(1, n, 1, 1).reshape(b * n, 1, i, i)
            patch_2 = patch_2.repeat(1, n, 1, 1).reshape(b * n, 1, i, i)
        delta_hats = delta_hats.reshape(b * n, 4, 2)
        if delta_hats_21 is not None:
            delta_hats_21 = delta_hats_21.reshape(b * n, 4, 2)
        if self._fused_path():
            loss, aux = self._double_line_fused(data, patch_1, patch_2, delta_hats, delta_hats_21)
        else:
            loss, aux = self._triplet_generic(data, patch_1, patch_2, delta_hats, delta_hats_21, scores)
        if 'summary_writer' in data:
            self._log(data, aux)
        delta_gt = data['delta'] if 'delta' in data else None
        if scores is not None:
            delta_hats = (delta_hats * scores.reshape(b * n, 1, 1)).reshape(b, n, 4, 2).sum(dim=1)
        return loss, delta_gt, delta_hats

    def _double_line_fused(self, data, patch_1, patch_2, d12, d21):
        """The shipped biHomE configuration on K1 + K2 + K3 (reference :352-402, 447-459, 559-561, 609-665)."""
        P = patch_1.shape[-1]
        trainable_extractor = any(p.requires_grad for p in self.auxiliary_resnet.parameters())
        with torch.set_grad_enabled(trainable_extractor and torch.is_grad_enabled()):
            f1 = self._features(patch_1)
            f2 = self._features(patch_2)
        pool = P // f1.shape[-2] if self.sampling_strategy not in ('upsample-patch-4x', 'upsample-patch-2x') \
            else patch_1.shape[-1] // f1.shape[-2]
        B = patch_1.shape[0]
        H = F.dlt4(torch.cat([d12, d21], dim=0), size=(patch_1.shape[-2], patch_1.shape[-1]))
        user_masks = len(self.mask_keys) > 0
        if user_masks:
            m1 = data[self.mask_keys[0]].reshape(B, 1, P, P)
            m2 = data[self.mask_keys[1]].reshape(B, 1, P, P)
            warped = F.warp(_stack(patch_1, patch_2), H, P, P)
            mw = nn.functional.avg_pool2d(F.warp(_stack(m1, m2), H, P, P), pool).squeeze(1)
            m1p = nn.functional.avg_pool2d(m1, pool).squeeze(1)
            m2p = nn.functional.avg_pool2d(m2, pool).squeeze(1)
        else:
            warped, mw = F.warp(_stack(patch_1, patch_2), H, P, P, pool=pool)
            m1p = m2p = None
        f1w = self._features(warped[:B])
        f2w = self._features(warped[B:])
        loss_b, parts = F.bihome_loss(f1, f2, f1w, f2w, mw[:B], mw[B:], H[:B], H[B:], self.triplet_mu, m1=m1p, m2=m2p)
        self.last_parts = parts
        return loss_b.sum(), dict(f1=f1, f2=f2, f1w=f1w, h1=H[:B], parts=parts)

    def _loss_variant(self, double):
        """(distance, hinge, margin) of F.triplet_loss for this head's TRIPLET_* settings (reference :465-538, :555-665)"""
        distance = self.triplet_distance
        if not double:
            # one-line: max(l1 - l3 + margin, 0) on channel-aggregated distances (:512)
            assert distance in ('l1', 'cosine'), 'Do not know this distance metric'
            return distance, 'pixel', float(self.triplet_margin)
        assert distance in ('l1', 'l2', 'cosine'), 'Do not know this distance metric'
        if isinstance(self.triplet_margin, str):
            return distance, None, 0.0
        assert self.triplet_channel_aggregation in ('channel-aware', 'channel-agnostic'), 'Do not know this aggregation technique'
        if self.triplet_channel_aggregation == 'channel-aware' and distance == 'l1':
            return distance, 'channel', float(self.triplet_margin)
        # channel-agnostic numeric margin (and the per-pixel distances, which have no channels left to be aware of): the
        # reference's branch is shape-inconsistent (:627-628,647-649); this is its evident intent, the margin applied to
        # the channel-aggregated distance
        return distance, 'pixel', float(self.triplet_margin)

    def _triplet_generic(self, data, patch_1, patch_2, d12, d21, scores):
        """Every other variant of the reference's triplet_resnet_loss (:320-714): both directions through K1 / K2 in one
        launch each, the loss algebra and its gradients in K3g (F.triplet_loss)."""
        B, P = patch_1.shape[0], patch_1.shape[-1]
        double = 'double-line' in self.triplet_version
        user_masks = len(self.mask_keys) > 0
        if user_masks:
            m1 = data[self.mask_keys[0]].reshape(-1, 1, P, P)
            m2 = data[self.mask_keys[1]].reshape(-1, 1, P, P)
            if m1.shape[0] != B:
                m1 = m1.repeat_interleave(B // m1.shape[0], dim=0)
                m2 = m2.repeat_interleave(B // m2.shape[0], dim=0)
        f1 = self._features(patch_1)
        f2 = self._features(patch_2)
        k = P // f1.shape[-2]
        H = F.dlt4(torch.cat([d12, d21], dim=0) if double else d12, size=(patch_1.shape[-2], patch_1.shape[-1]))
        src = _stack(patch_1, patch_2) if double else patch_1
        if user_masks:
            warped = F.warp(src, H, P, P)
            mw_full = F.warp(_stack(m1, m2) if double else m1, H, P, P).squeeze(1)
            mw = nn.functional.avg_pool2d(mw_full.unsqueeze(1), k).squeeze(1)
            m1p = nn.functional.avg_pool2d(m1, k).squeeze(1)
            m2p = nn.functional.avg_pool2d(m2, k).squeeze(1)
        else:
            warped, mw = F.warp(src, H, P, P, pool=k)
            m1p = m2p = None
        h1 = H[:B]
        h2 = H[B:] if double else None
        f1w = self._features(warped[:B])
        f2w = self._features(warped[B:]) if double else None
        aux = dict(f1=f1, f2=f2, f1w=f1w, h1=h1, parts=None)
        n1, n2, n1w = f1, f2, f1w
        if not double and getattr(self.auxiliary_resnet, 'with_projection_head', None) is not None:
            n1w = f1w / f1w.norm(p=2, dim=1, keepdim=True)
            n2 = f2 / f2.norm(p=2, dim=1, keepdim=True)
            n1 = f1 / f1.norm(p=2, dim=1, keepdim=True)
        distance, hinge, margin = self._loss_variant(double)
        loss_b, parts = F.triplet_loss(n1, n2, n1w, f2w, mw[:B], m2p, mw[B:] if double else None, m1p, h1, h2,
                                       lines=2 if double else 1, distance=distance, hinge=hinge, margin=margin,
                                       mask_crd=bool(self.change_detection_mask) and not double,
                                       mu=self.triplet_mu if double else 0.0)
        if scores is not None:
            loss_b = loss_b * scores.reshape(B)
        loss = loss_b.sum()
        aux['parts'] = parts
        if 'dual' in self.triplet_version:
            # the content-aware backbone's own full-resolution feature extractor adds a second triplet term (:407-441)
            fe = self.backbone.feature_extractor
            g1, g2, g1w = fe(patch_1), fe(patch_2), fe(warped[:B])
            g2w = fe(warped[B:]) if double else None
            if user_masks:
                a_full, m1f, m2f = mw_full, m1.squeeze(1), m2.squeeze(1)
            else:
                a_full, m1f, m2f = F.coverage_mask(H, (P, P), (P, P), 1), None, None
            dual_b, _ = F.triplet_loss(g1, g2, g1w, g2w, a_full[:B], m2f, a_full[B:] if double else None, m1f, h1, h2,
                                       lines=2 if double else 1, distance='l1', hinge=None, mu=0.0)
            loss = loss + dual_b.sum()
        return loss, aux

    def _log(self, data, aux):
        """TensorBoard scalars of the reference (:678-697); host syncs only on logging steps."""
        step, sw = data['summary_writer_step'], data['summary_writer']
        f1, f2, f1w, h1 = aux['f1'], aux['f2'], aux['f1w'], aux['h1']
        sw.add_scalars('feature_space', {'patch_1_f': f1.mean().item()}, step)
        sw.add_scalars('feature_space', {'patch_2_f': f2.mean().item()}, step)
        sw.add_scalars('feature_space', {'patch_1_f_prime': f1w.mean().item()}, step)
        sw.add_scalars('loss_comp', {'l1': (f2 - f1w).abs().mean().item()}, step)
        sw.add_scalars('loss_comp', {'l3': (f2 - f1).abs().mean().item()}, step)
        eye = torch.eye(3, dtype=h1.dtype, device=h1.device).unsqueeze(0)
        sw.add_scalars('h', {'h1': ((h1 - eye) ** 2).sum().item()}, step)
        if aux.get('parts') is not None:
            sw.add_scalars('loss_den', {'l1_den': aux['parts'][:, 2].min().item()}, step)
            sw.add_scalars('loss_den', {'l2_den': aux['parts'][:, 3].min().item()}, step)
        elif 'dens' in aux:
            sw.add_scalars('loss_den', {'l1_den': aux['dens'][0].min().item()}, step)
            sw.add_scalars('loss_den', {'l2_den': aux['dens'][1].min().item()}, step)

    # ------------------------------------------------------------------------------------------------
    def predict_homography(self, data):
        """(delta_hat [B,4,2], None) -- reference :716-767."""
        if len(self.delta_hat_keys):
            return data[self.delta_hat_keys[0]], None
        pf = data[self.pf_keys[0]]
        B, _, Hf, Wf = pf.shape
        n, M = self.hypothesis_no, self.point_per_hypothesis
        if n == 1:
            delta, _ = self._field_to_delta(pf, 0)
            return delta.reshape(B, 4, 2), None
        map_field, self.coordinate_field_12, self.four_points_12 = self.forward_map_field(
            pf, self.coordinate_field_12, self.four_points_12)
        H, scores = self.dsac(self.coordinate_field_12, map_field, hypothesis_no=n, points_per_hypothesis=M)
        best = torch.argmax(scores, dim=-1).reshape(-1, 1, 1, 1).repeat(1, 1, 3, 3)
        H = torch.gather(H, dim=1, index=best).reshape(-1, 3, 3)
        four = torch.tensor([[0, 0], [Wf, 0], [Wf, Hf], [0, Hf]], device=pf.device, dtype=torch.float32).unsqueeze(0).repeat(B, 1, 1)
        return (transform_points(H, four) - four).reshape(B, 4, 2), None
