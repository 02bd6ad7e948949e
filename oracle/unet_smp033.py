"""Plain-torch restatement of segmentation_models_pytorch==0.3.3 `Unet(encoder_name="resnet34")`.

TEST INFRASTRUCTURE (see oracle/__init__.py). The reference builds the network with
`smp.create_model(arch="unet", encoder_name="resnet34", classes=n, in_channels=c)`
(src/zone_detect/model.py:30-39, src/flair/model.py:35-41); smp is an un-vendored dependency pinned
at 0.3.3 (setup.py:36), so its published architecture is restated here with IDENTICAL state_dict key
names so that a FLAIR `.pth` loads with strict=True:

  encoder  = torchvision ResNet-34 without avgpool/fc (conv1 7x7/2, bn1, relu, maxpool 3x3/2,
             layer1..4 = [3,4,6,3] BasicBlocks of 64/128/256/512 channels)
  decoder  = 5 DecoderBlocks: nearest x2 upsample, concat skip, 2 x [conv3x3 (no bias), BN, ReLU],
             out channels (256,128,64,32,16), skips = layer3, layer2, layer1, stem(relu), none
  head     = Conv2d(16, classes, 3, padding=1) with bias, no activation
"""
from __future__ import annotations

from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F


class BasicBlock(nn.Module):
    def __init__(self, cin: int, cout: int, stride: int):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return self.relu(out + idt)


def _make_layer(cin: int, cout: int, blocks: int, stride: int) -> nn.Sequential:
    layers = [BasicBlock(cin, cout, stride)]
    layers += [BasicBlock(cout, cout, 1) for _ in range(blocks - 1)]
    return nn.Sequential(*layers)


class ResNet34Encoder(nn.Module):
    def __init__(self, in_channels: int):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.layer1 = _make_layer(64, 64, 3, 1)
        self.layer2 = _make_layer(64, 128, 4, 2)
        self.layer3 = _make_layer(128, 256, 6, 2)
        self.layer4 = _make_layer(256, 512, 3, 2)

    def forward(self, x) -> List[torch.Tensor]:
        feats = [x]
        x = self.relu(self.bn1(self.conv1(x)))
        feats.append(x)
        x = self.layer1(self.maxpool(x))
        feats.append(x)
        x = self.layer2(x)
        feats.append(x)
        x = self.layer3(x)
        feats.append(x)
        x = self.layer4(x)
        feats.append(x)
        return feats


class _ConvBnRelu(nn.Sequential):  # smp Conv2dReLU: keys "0" (conv) and "1" (bn)
    def __init__(self, cin: int, cout: int):
        super().__init__(nn.Conv2d(cin, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class DecoderBlock(nn.Module):
    def __init__(self, cin: int, cskip: int, cout: int):
        super().__init__()
        self.conv1 = _ConvBnRelu(cin + cskip, cout)
        self.conv2 = _ConvBnRelu(cout, cout)

    def forward(self, x, skip=None):
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        if skip is not None:
            x = torch.cat([x, skip], dim=1)
        return self.conv2(self.conv1(x))


class UnetDecoder(nn.Module):
    def __init__(self):
        super().__init__()
        ins, skips, outs = [512, 256, 128, 64, 32], [256, 128, 64, 64, 0], [256, 128, 64, 32, 16]
        self.blocks = nn.ModuleList([DecoderBlock(i, s, o) for i, s, o in zip(ins, skips, outs)])

    def forward(self, *features):
        features = features[1:][::-1]  # drop the identity feature, deepest first
        x, skips = features[0], features[1:]
        for i, blk in enumerate(self.blocks):
            x = blk(x, skips[i] if i < len(skips) else None)
        return x


class Unet(nn.Module):
    """`smp.Unet("resnet34", in_channels=c, classes=n)`-compatible module (eval-mode inference)."""

    def __init__(self, in_channels: int = 3, classes: int = 15):
        super().__init__()
        self.encoder = ResNet34Encoder(in_channels)
        self.decoder = UnetDecoder()
        self.segmentation_head = nn.Sequential(nn.Conv2d(16, classes, 3, padding=1))

    def forward(self, x):
        h, w = x.shape[-2:]
        if h % 32 or w % 32:  # smp check_input_shape
            raise RuntimeError(f"Wrong input shape height={h}, width={w}. Expected divisible by 32.")
        return self.segmentation_head(self.decoder(*self.encoder(x)))


class MetadataMLP(nn.Module):
    """src/flair/model.py:74-96 (dropout is inactive in eval mode)."""

    def __init__(self):
        super().__init__()
        self.enc_mlp = nn.Sequential(nn.Linear(45, 64), nn.Dropout(0.4), nn.ReLU(), nn.Linear(64, 32), nn.Dropout(0.4),
                                     nn.ReLU(), nn.Linear(32, 16), nn.Dropout(0.4), nn.ReLU())

    def forward(self, x):
        return self.enc_mlp(x)


class FlairModel(nn.Module):
    """src/flair/model.py:7-70 restated (smp branch). The reference constructor tests the bare name
    `model_provider` (model.py:32, a NameError whenever use_metadata is True); the evident intent
    `self.model_provider` is used here."""

    def __init__(self, in_channels: int, classes: int, use_metadata: bool):
        super().__init__()
        self.use_metadata = use_metadata
        if use_metadata:
            self.enc = MetadataMLP()
        self.seg_model = Unet(in_channels, classes)

    def forward(self, x, met=None):
        if self.use_metadata:
            feats = self.seg_model.encoder(x)
            x_enc = self.enc(met)
            x_enc = x_enc.unsqueeze(1).unsqueeze(-1).repeat(1, 512, 1, 16)  # model.py:59: value depends on row h only
            feats[-1] = torch.add(feats[-1], x_enc)
            return self.seg_model.segmentation_head(self.seg_model.decoder(*feats))
        return self.seg_model(x)


def layer_activations(model: Unet, x: torch.Tensor, met_enc: torch.Tensor | None = None) -> dict:
    """Named fp32 NCHW activations matching the names of fb_debug_activation (csrc/api.cu)."""
    acts = {}
    e = model.encoder
    with torch.no_grad():
        f1 = e.relu(e.bn1(e.conv1(x)))
        acts["f1"] = f1
        cur = e.maxpool(f1)
        acts["pool"] = cur
        for li, layer in enumerate([e.layer1, e.layer2, e.layer3, e.layer4]):
            for bi, blk in enumerate(layer):
                cur = blk(cur)
                acts[f"layer{li + 1}.{bi}.out"] = cur
        if met_enc is not None:
            cur = cur + met_enc.unsqueeze(1).unsqueeze(-1).repeat(1, 512, 1, 16)
            acts["layer4.2.out"] = cur
        skips = [acts["layer3.5.out"], acts["layer2.3.out"], acts["layer1.2.out"], f1, None]
        for di, blk in enumerate(model.decoder.blocks):
            cur = blk(cur, skips[di])
            acts[f"dec{di}"] = cur
        acts["logits"] = model.segmentation_head(cur)
    return acts
