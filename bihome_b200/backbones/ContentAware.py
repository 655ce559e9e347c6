"""Content-aware backbone ("Zhang", reference ``src/backbones/ContentAware.py``): a small mask predictor and a small
feature extractor run on each patch, their product G = M * F goes through a 2-channel ResNet-34 that regresses the
4-point offsets.  Same kwargs, same ``forward(data) -> data`` dict protocol (it adds the mask / feature / delta_hat
keys), same parameter names as the reference (``mask_predictor.layer<i>.{0,1}``, ``feature_extractor.layer<i>.{0,1}``,
``resnet34.*``), so the shipped ``zhang-*`` YAML files and reference checkpoints load unchanged.

Dense convolutions: cuDNN through PyTorch (BASELINE.json north_star).  Each patch goes through the small nets in its
own call, and the trunk runs once per direction, because every BatchNorm here normalises with the statistics of the
call it sits in -- batching the two patches would change the numbers the reference produces.
"""
import torch
import torch.nn as nn

from .blocks import offset_regressor


def _stage(cin, cout, last=None):
    """3x3 conv (no bias) + BatchNorm + activation, indices 0 / 1 / 2 as in the reference's Sequentials"""
    return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=3, stride=1, padding=1, bias=False), nn.BatchNorm2d(cout),
                         last if last is not None else nn.ReLU())


class MaskPredictor(nn.Module):
    """1 -> 4 -> 8 -> 16 -> 32 -> 1 channels, sigmoid output (reference :6-54).  ``fix_mask`` short-circuits to ones."""

    def __init__(self, fix_mask=False, normalization_strength=-1):
        super().__init__()
        self.fix_mask = fix_mask
        self.normalization_strength = normalization_strength
        widths = (1, 4, 8, 16, 32)
        for i in range(4):
            setattr(self, 'layer%d' % (i + 1), _stage(widths[i], widths[i + 1]))
        self.layer5 = _stage(32, 1, nn.Sigmoid())

    @staticmethod
    def _normalize(mask, strength):
        peak = mask.flatten(1).amax(dim=1).reshape(-1, 1, 1, 1)
        return torch.clamp(mask / (peak * strength), 0, 1)

    def forward(self, x):
        if self.fix_mask:
            return torch.ones_like(x)
        out = x
        for i in range(1, 6):
            out = getattr(self, 'layer%d' % i)(out)
        assert out.shape[-2:] == x.shape[-2:], 'Mask and input image should have the same w/h'
        if self.normalization_strength > 0:
            out = self._normalize(out, self.normalization_strength)
        return out


class FeatureExtractor(nn.Module):
    """1 -> 4 -> 8 -> 1 channels at full resolution (reference :57-81); the TripletHead calls it on warped patches."""

    def __init__(self):
        super().__init__()
        self.layer1 = _stage(1, 4)
        self.layer2 = _stage(4, 8)
        self.layer3 = _stage(8, 1)

    def forward(self, x):
        out = self.layer3(self.layer2(self.layer1(x)))
        assert out.shape[-2:] == x.shape[-2:], 'Feature map and input image should have the same w/h'
        return out

    def retrieve_weights(self):
        return {name: p.data for name, p in self.named_parameters()}


class Model(nn.Module):

    def __init__(self, **kwargs):
        super().__init__()
        self.patch_keys = kwargs['PATCH_KEYS']
        self.mask_keys = kwargs['MASK_KEYS']
        self.feature_keys = kwargs['FEATURE_KEYS']
        self.target_keys = kwargs['TARGET_KEYS']
        strength = kwargs['MASK_NORMALIZATION_STRENGTH'] if 'MASK_NORMALIZATION_STRENGTH' in kwargs else -1
        self.mask_predictor = MaskPredictor(fix_mask=kwargs['FIX_MASK'], normalization_strength=strength)
        self.feature_extractor = FeatureExtractor()
        self.variant = str.lower(kwargs['VARIANT'])
        assert 'oneline' in self.variant or 'doubleline' in self.variant, 'Only OneLine or DoubleLine variant is supported'

        pretrained = kwargs['PRETRAINED_RESNET']
        if pretrained:
            self.init()                 # reference order (:104-117): small nets first, then the pretrained trunk
        self.resnet34 = offset_regressor(pretrained)
        if not pretrained:
            self.init()

    def init(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _content(self, x):
        """(mask, features, G = mask * features); the product is skipped for the constant-ones mask (bit-identical)"""
        m = self.mask_predictor(x)
        f = self.feature_extractor(x)
        g = f if self.mask_predictor.fix_mask else m * f
        assert g.shape == x.shape, 'G feature map should have the same size as input image'
        return m, f, g

    def _regress(self, ga, gb):
        g = torch.cat([ga, gb], dim=1)
        assert g.shape[1] == 2, 'G feature map should have 2 channels'
        return self.resnet34(g).reshape(-1, 4, 2)

    def _forward(self, input_1, input_2):
        m1, f1, g1 = self._content(input_1)
        m2, f2, g2 = self._content(input_2)
        return m1, f1, m2, f2, g1, g2, self._regress(g1, g2)

    def forward(self, data):
        e1, e2 = self.patch_keys
        m1, m2 = self.mask_keys
        f1, f2 = self.feature_keys
        data[m1], data[f1], data[m2], data[f2], g1, g2, data[self.target_keys[0]] = self._forward(data[e1], data[e2])
        if self.variant == 'doubleline':
            data[self.target_keys[1]] = self._regress(g2, g1)
        return data

    def predict_homography(self, data):
        e1, e2 = self.patch_keys
        m1, m2 = self.mask_keys
        data[m1], _, data[m2], _, _, _, data[self.target_keys[0]] = self._forward(data[e1], data[e2])
        return data

    def retrieve_weights(self):
        return {name: p.data for name, p in self.resnet34.named_parameters()}
