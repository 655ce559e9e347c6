"""DeTone et al.'s original VGG-style regressor (reference ``src/backbones/HomographyNet.py``): eight (twelve for
512-pixel inputs) conv3x3 -> ReLU -> BatchNorm stages with max-pooling, two fully connected layers, 4-point offsets
out.  Same kwargs, dict protocol and parameter names (``layer<i>.{0,2}``, ``fc1.0``, ``fc2``).  cuDNN / cuBLAS work."""
import torch
import torch.nn as nn

# (in, out, pool after) per stage
_PLAN_128 = [(2, 64, 0), (64, 64, 1), (64, 64, 0), (64, 64, 1), (64, 128, 0), (128, 128, 1), (128, 128, 0), (128, 128, 0)]
_PLAN_512 = _PLAN_128[:7] + [(128, 128, 1), (128, 128, 0), (128, 128, 1), (128, 128, 0), (128, 128, 0)]


def _stage(cin, cout, pool):
    layers = [nn.Conv2d(cin, cout, 3, padding=1), nn.ReLU(), nn.BatchNorm2d(cout)]
    if pool:
        layers.append(nn.MaxPool2d(2))
    return nn.Sequential(*layers)


class Model(nn.Module):

    def __init__(self, **kwargs):
        super().__init__()
        self.image_size = kwargs['IMAGE_SIZE']
        self.patch_keys = kwargs['PATCH_KEYS']
        self.target_keys = kwargs['TARGET_KEYS']
        assert self.image_size in (128, 512), 'HomographyNet is defined for 128- and 512-pixel patches'
        plan = _PLAN_128 if self.image_size == 128 else _PLAN_512
        self.depth = len(plan)
        for i, spec in enumerate(plan):
            setattr(self, 'layer%d' % (i + 1), _stage(*spec))
        self.fc1 = nn.Sequential(nn.Linear(128 * 16 * 16, 1024), nn.ReLU())
        self.fc2 = nn.Linear(1024, 8)

    def _forward(self, x):
        for i in range(self.depth):
            x = getattr(self, 'layer%d' % (i + 1))(x)
        x = x.reshape(-1, 128 * 16 * 16)      # NCHW element order, as the reference's .view on a contiguous tensor
        return self.fc2(self.fc1(x)).reshape(-1, 4, 2)

    def forward(self, data):
        e1, e2 = self.patch_keys
        data[self.target_keys[0]] = self._forward(torch.cat([data[e1], data[e2]], dim=1))
        return data

    def predict_homography(self, data):
        return self.forward(data)
