"""Perspective-field backbone ("Zeng", reference ``src/backbones/Rethinking.py``): an encoder/decoder ResNet that
maps the two stacked patches [B,2,P,P] to a dense displacement field [B,2,P,P].  Same kwargs, same
``forward(data) -> data`` dict protocol, same parameter names as the reference (checkpoint compatible)."""
import warnings

import torch
import torch.nn as nn

from .. import functional as F
from .blocks import (ResNet34ConvBlock, ResNet34IdentityBlock, ResNet50ConvBlock, ResNet50DeconvBlock,
                     ResNet50IdentityBlock)

# per stage: ('conv', cin, cout, stride) | ('id', c) | ('up', c); widths for ResNet34; ResNet50 multiplies by 4
_STAGES = {
    'layer2': [('conv', 64, 64, 1), ('id', 64), ('id', 64)],
    'layer3': [('conv', 64, 128, 2)] + [('id', 128)] * 3,
    'layer4': [('conv', 128, 256, 2)] + [('id', 256)] * 5 + [('up', 256)],
    'layer5': [('id', 128)] * 3 + [('up', 128)],
    'layer6': [('id', 64)] * 2 + [('up', 64)],
    'layer7': [('id', 32), ('up', 32)],
}
_RESNET_URLS = {'ResNet50': 'https://download.pytorch.org/models/resnet50-19c8e357.pth',
                'ResNet34': 'https://download.pytorch.org/models/resnet34-333f7ec4.pth'}


class Model(nn.Module):
    # training-time shortcut of _layer8_aten (exact); bench.py's reference arm switches it off so that the CPU baseline
    # runs the reference's op sequence literally
    skip_cancelled_bias = True

    def __init__(self, **kwargs):
        super().__init__()
        self.image_size = kwargs['IMAGE_SIZE']
        self.patch_keys = kwargs['PATCH_KEYS']
        self.target_keys = kwargs['TARGET_KEYS']
        self.resnet_block = kwargs['RESNET_BLOCK']
        self.pretrained_resnet = kwargs['PRETRAINED_RESNET']
        self.variant = str.lower(kwargs['VARIANT']) if 'VARIANT' in kwargs else 'oneline'
        assert 'oneline' in self.variant or 'doubleline' in self.variant, 'Only OneLine or DoubleLine variant is supported'
        assert self.resnet_block in ('ResNet34', 'ResNet50'), 'I know only ResNet50 and ResNet34'
        wide = self.resnet_block == 'ResNet50'
        m = 4 if wide else 1

        self.layer1 = nn.Sequential(nn.Conv2d(2, 64, kernel_size=7, padding=3, stride=2, bias=False), nn.BatchNorm2d(64),
                                    nn.ReLU(), nn.MaxPool2d(kernel_size=3, stride=2, padding=1))
        for name, stage in _STAGES.items():
            blocks = []
            for spec in stage:
                if spec[0] == 'conv':
                    cin = spec[1] if (name == 'layer2') else spec[1] * m
                    if wide:
                        blocks.append(ResNet50ConvBlock(cin, spec[2] * m, spec[3]))
                    else:
                        blocks.append(ResNet34ConvBlock(cin, spec[2], spec[3]))
                elif spec[0] == 'id':
                    blocks.append(ResNet50IdentityBlock(spec[1] * m) if wide else ResNet34IdentityBlock(spec[1]))
                else:
                    blocks.append(ResNet50DeconvBlock(spec[1] * m))
            setattr(self, name, nn.Sequential(*blocks))
        c = 16 * m
        self.layer8 = nn.Sequential(nn.Conv2d(c, 8 * c, kernel_size=1), nn.BatchNorm2d(8 * c), nn.ReLU(),
                                    nn.Conv2d(8 * c, 2, kernel_size=1))
        if self.pretrained_resnet:
            self._load_pretrained_weights()

    def _load_pretrained_weights(self):
        """ImageNet weights of torchvision layer1..3 -> layer2..4 (reference :161-289); skipped with a warning offline."""
        try:
            from torch.hub import load_state_dict_from_url
            state = load_state_dict_from_url(_RESNET_URLS[self.resnet_block], progress=True)
        except Exception as e:  # noqa: BLE001
            warnings.warn('bihome_b200: pretrained %s weights unavailable (%s); backbone keeps its random init'
                          % (self.resnet_block, e))
            return
        rename = {'conv1': 'upper_branch.0', 'bn1': 'upper_branch.1', 'conv2': 'upper_branch.3', 'bn2': 'upper_branch.4',
                  'conv3': 'upper_branch.6', 'bn3': 'upper_branch.7', 'downsample': 'lower_branch'}
        own = self.state_dict()
        picked = {}
        for key, value in state.items():
            for i in (1, 2, 3):
                if key.startswith('layer%d.' % i):
                    new = key.replace('layer%d' % i, 'layer%d' % (i + 1))
                    for a, b in rename.items():
                        new = new.replace(a, b)
                    if new in own and own[new].shape == value.shape:
                        picked[new] = value
        self.load_state_dict(picked, strict=False)

    def _forward(self, x):
        # layer1 = Conv2d 7x7/2 -> BatchNorm2d -> ReLU -> MaxPool2d(3, 2, 1): after the convolution (cuDNN) the three
        # memory-bound modules run as one stage (K7, csrc/stem.cu) when the tensor is channels-last and BatchNorm trains
        conv, bn, act, pool = self.layer1
        x = conv(x)
        x = F.stem(bn, x) if F.stem_supported(bn, pool, x) else pool(act(bn(x)))
        for i in range(2, 8):
            x = getattr(self, 'layer%d' % i)(x)
        # layer8 (1x1 conv -> BatchNorm -> ReLU -> 1x1 conv at full resolution) is 21 passes over a [B,128,P,P] tensor
        # through ATen; K6 computes it per pixel from the 16-channel input when the device's self-test picked it
        # (functional.field_head_enabled -> bihome_b200/autotune.py)
        if x.is_cuda and F.field_head_supported(self.layer8, x) and F.field_head_enabled(x.device):
            return F.field_head(self.layer8, x)
        return self._layer8_aten(x)

    def _layer8_aten(self, x):
        """layer8 through the ATen modules.  In training its first convolution runs WITHOUT its bias: a per-channel
        constant in front of a batch-statistics BatchNorm cancels exactly (output and every gradient unchanged, the
        bias' own gradient is exactly zero), while adding it costs a broadcast pass over the [B,128,P,P] tensor
        (2.1 GB read + written at B = 256: PyTorch adds cuDNN convolution biases in a separate strided kernel) and
        a 2.1 GB reduction for its gradient.  Only BatchNorm's running mean sees the bias; it is added there."""
        conv, bn, act, head = self.layer8
        if not (self.skip_cancelled_bias and bn.training and conv.bias is not None and isinstance(bn, nn.BatchNorm2d) and bn.momentum is not None
                and 0 < bn.momentum < 1):
            return self.layer8(x)
        # keep the bias in the graph (DDP / optimizers see a used parameter) with its exact zero gradient
        weight = conv.weight * (1 + 0 * conv.bias.sum())
        y = nn.functional.conv2d(x, weight, None, conv.stride, conv.padding, conv.dilation, conv.groups)
        if bn.track_running_stats and bn.running_mean is not None:
            # running_mean <- (1 - m) running_mean + m (mean(y) + bias): shift it before the module's own update.  The
            # buffer is among the tensors every BatchNorm call saves for backward (unused there in training), so the
            # shift goes through .data, like the op's own in-place update it does not advance the version counter.
            m = float(bn.momentum)
            bn.running_mean.data.add_(conv.bias.detach().to(bn.running_mean.dtype), alpha=m / (1 - m))
        return head(act(bn(y)))

    def _pair(self, a, b):
        """the two patches stacked along the channels (reference :304,310 `torch.cat`).  When the model is channels-last the
        [B,2,P,P] tensor is built in that layout right away: handed an NCHW tensor, cuDNN's channels-last convolution
        re-lays it out in the forward pass and once more for the weight gradient."""
        w = self.layer1[0].weight
        if (a.is_cuda and a.dim() == 4 and a.shape[1] == 1 and a.shape == b.shape and a.dtype == b.dtype
                and w.is_contiguous(memory_format=torch.channels_last) and not w.is_contiguous()):
            return torch.stack([a[:, 0], b[:, 0]], dim=-1).permute(0, 3, 1, 2)
        return torch.cat([a, b], dim=1)

    def forward(self, data):
        e1, e2 = self.patch_keys
        p1, p2 = data[e1], data[e2]
        data[self.target_keys[0]] = self._forward(self._pair(p1, p2))
        if self.variant == 'doubleline':
            data[self.target_keys[1]] = self._forward(self._pair(p2, p1))
        return data

    def predict_homography(self, data):
        return self.forward(data)
