"""Stage-1/2/3 DSRL network, written for this harness with the SAME parameter / buffer names as the reference's
``models/DSRL.py`` (+ ``models/modules/ASPP.py``, ``models/modules/backbone/ResNet101.py``) so that reference
``.weights`` / ``.checkpoint`` state dicts load unchanged: ``feature_extractor.{backbone,aspp.branches.0-5,shortcut_conv}``,
``SSSR_decoder.{cat_conv,cls_conv,upsample16_pred}``, ``SISR_decoder``, ``SSSR_feature_transformer``,
``SISR_feature_transformer``.  tests/test_harness_model.py checks key-for-key / value-for-value agreement with the
reference model (in the build container, where /root/reference is mounted).

Architecture (DSRL.py:11-186): ResNet-101 with the last stage dilated (output stride 16) -> ASPP (rates 6/12/18 + image
pooling) -> x4 bilinear -> concat with a 48-channel projection of the stride-4 features -> SSSR head (two 3x3 convs,
classifier, x2 bilinear, two stride-2 transposed convs = x8) ; SISR head (3x3 conv -> PixelShuffle(8)) ; stage 3 adds the
two feature transformers (1x1 conv stride 8 -> BN -> ReLU, one output channel) whose outputs feed the FA loss.
"""
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F
import torchvision

NUM_RGB = 3


def _cbr(cin, cout, k, padding=0, dilation=1, stride=1):
    return nn.Sequential(nn.Conv2d(cin, cout, k, stride=stride, padding=padding, dilation=dilation, bias=False),
                         nn.BatchNorm2d(cout), nn.ReLU())


class _Backbone(nn.Module):
    """torchvision ResNet-101 trunk (no avgpool/fc), stride-16 variant; returns (stride-16 features, stride-4 features)."""

    def __init__(self):
        super().__init__()
        trunk = torchvision.models.resnet101(weights=None, replace_stride_with_dilation=[False, False, True])
        for name in ("conv1", "bn1", "relu", "maxpool", "layer1", "layer2", "layer3", "layer4"):
            setattr(self, name, getattr(trunk, name))

    def forward(self, x):
        low = self.layer1(self.maxpool(self.relu(self.bn1(self.conv1(x)))))
        return self.layer4(self.layer3(self.layer2(low))), low


class _ASPP(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        specs = [(cin, 1, 0, 1), (cin, 3, 6, 6), (cin, 3, 12, 12), (cin, 3, 18, 18), (cin, 1, 0, 1), (5 * cout, 1, 0, 1)]
        self.branches = nn.ModuleList(_cbr(ci, cout, k, padding=p, dilation=d) for ci, k, p, d in specs)

    def forward(self, x):
        outs = [self.branches[i](x) for i in range(4)]
        pooled = self.branches[4](F.adaptive_avg_pool2d(x, 1))
        outs.append(F.interpolate(pooled, size=x.shape[-2:], mode="bilinear", align_corners=True))
        return self.branches[5](torch.cat(outs, dim=1))


def _kaiming(*modules):
    for mod in modules:
        for m in mod.modules():
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)


class DSRL(nn.Module):
    def __init__(self, stage=3, num_classes=19):
        super().__init__()
        assert stage in (1, 2, 3)
        self.stage = stage
        self.feature_extractor = nn.ModuleDict(OrderedDict(
            backbone=_Backbone(), aspp=_ASPP(2048, 256), shortcut_conv=_cbr(256, 48, 1)))
        nc = num_classes
        self.SSSR_decoder = nn.ModuleDict(OrderedDict(
            cat_conv=nn.Sequential(nn.Conv2d(304, 256, 3, padding=1, bias=False), nn.BatchNorm2d(256), nn.ReLU(), nn.Dropout(0.2),
                                   nn.Conv2d(256, 256, 3, padding=1, bias=False), nn.BatchNorm2d(256), nn.ReLU(), nn.Dropout(0.2)),
            cls_conv=nn.Conv2d(256, nc, 1),
            upsample16_pred=nn.Sequential(nn.UpsamplingBilinear2d(scale_factor=2.0), nn.Dropout(0.2),
                                          nn.ConvTranspose2d(nc, nc, 2, stride=2, bias=False), nn.BatchNorm2d(nc), nn.ReLU(),
                                          nn.Dropout(0.2), nn.ConvTranspose2d(nc, nc, 2, stride=2, bias=True))))
        _kaiming(self.feature_extractor["backbone"], self.feature_extractor["aspp"], self.feature_extractor["shortcut_conv"],
                 self.SSSR_decoder)
        for m in self.feature_extractor["backbone"].modules():          # zero-init the last BN of every residual branch
            if isinstance(m, torchvision.models.resnet.Bottleneck):
                nn.init.zeros_(m.bn3.weight)
        if stage > 1:
            self.SISR_decoder = nn.Sequential(nn.Conv2d(304, NUM_RGB * 64, 3, padding=1), nn.PixelShuffle(8))
            _kaiming(self.SISR_decoder)
        if stage > 2:
            self.SSSR_feature_transformer = _cbr(nc, 1, 1, stride=8)
            self.SISR_feature_transformer = _cbr(NUM_RGB, 1, 1, stride=8)
            _kaiming(self.SSSR_feature_transformer, self.SISR_feature_transformer)

    def forward(self, x, apply_transformers=True):
        """apply_transformers=False: the two stage-3 feature transformers are left to the loss (dsrl-b200 Stage3Loss evaluates
        them inside the CE / MSE / FA passes); the last two outputs are then None."""
        fe = self.feature_extractor
        deep, low = fe["backbone"](x)
        deep = F.interpolate(fe["aspp"](deep), scale_factor=4.0, mode="bilinear", align_corners=True)
        cat = torch.cat([deep, fe["shortcut_conv"](low)], dim=1)
        sssr = self.SSSR_decoder["upsample16_pred"](self.SSSR_decoder["cls_conv"](self.SSSR_decoder["cat_conv"](cat)))
        sisr = sssr_t = sisr_t = torch.zeros(1)
        if self.stage > 1:
            sisr = self.SISR_decoder(cat)
        if self.stage > 2:
            if not apply_transformers:
                return sssr, sisr, None, None
            sssr_t = self.SSSR_feature_transformer(sssr)
            sisr_t = self.SISR_feature_transformer(sisr)
        return sssr, sisr, sssr_t, sisr_t
