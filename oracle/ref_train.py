"""CPU training step of the reference's zeng-bihome / detone-bihome models -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.

The head is the oracle restatement (oracle/ref_path.py over oracle/kornia050.py: the reference's own op chain,
torch.solve-style 8x8 solves, materialised sampling grids, F.grid_sample, ~110-op loss, batched SVD); the backbone
and the frozen extractor are plain torch modules (they are cuDNN/ATen work in the reference too).  Follows
src/heads/PerceptualHead.py:148-235 (forward), :320-714 (triplet_resnet_loss), train.py:305-387 (step).
"""
import torch
import torchvision.models as models

from . import ref_path as R


class OracleExtractor(torch.nn.Module):
    """AuxiliaryResnet (PerceptualHead.py:15-76) at OUTPUT_LAYER 1 with random-init weights, frozen."""

    def __init__(self, name='resnet34'):
        super().__init__()
        self.resnet = getattr(models, name)(weights=None)
        for p in self.resnet.parameters():
            p.requires_grad = False

    def forward(self, x):
        r = self.resnet
        if x.shape[1] == 1:
            x = x.repeat(1, 3, 1, 1)
        return r.layer1(r.maxpool(r.relu(r.bn1(r.conv1(x)))))


class OracleModel(torch.nn.Module):
    """backbone (any module following the reference's dict protocol) + oracle biHomE head"""

    def __init__(self, backbone, head_cfg):
        super().__init__()
        self.backbone = backbone
        self.cfg = dict(head_cfg)
        self.extractor = OracleExtractor(self.cfg['AUXILIARY_RESNET'])

    def forward(self, data):
        data = self.backbone(data)
        c = self.cfg
        p1, p2 = data[c['PATCH_KEYS'][0]], data[c['PATCH_KEYS'][1]]
        if len(c['DELTA_HAT_KEYS']):
            d12, d21 = data[c['DELTA_HAT_KEYS'][0]], data[c['DELTA_HAT_KEYS'][1]]
        else:
            n, m = c['RANSAC_HYPOTHESIS_NO'], c['POINTS_PER_HYPOTHESIS']
            d12, _, _ = R.zeng_delta_hat(data[c['PF_KEYS'][0]], m, n)
            d21, _, _ = R.zeng_delta_hat(data[c['PF_KEYS'][1]], m, n)
            d12, d21 = d12.reshape(-1, 4, 2), d21.reshape(-1, 4, 2)
        loss, _ = R.head_double_line(p1, p2, d12, d21, self.extractor, c['TRIPLET_MU'])
        return loss, data.get('delta'), d12
