"""Model factory and checkpoint loading for zone detection (src/zone_detect/model.py).

The reference builds `smp.create_model("unet", "resnet34", classes, in_channels)` and calls
load_state_dict(strict=True) (model.py:30-39, 79-88). Here the "model" is a libflairb200 context that
owns BN-folded bf16 weights on the GPU; the checkpoint layout (.pth = bare state_dict, .ckpt =
Lightning dict with "state_dict", optional "model.seg_model." key prefix) is unchanged.
"""
from __future__ import annotations

import os
from pathlib import Path
from typing import Mapping

import torch

from .. import _native

# exactly the float tensors smp-0.3.3 Unet(resnet34) holds; strict=True parity
_STAGES = ((1, 3), (2, 4), (3, 6), (4, 3))


def expected_keys(use_metadata: bool = False) -> set:
    ks = {"encoder.conv1.weight"}
    bn = ("weight", "bias", "running_mean", "running_var")
    ks |= {f"encoder.bn1.{s}" for s in bn}
    for layer, blocks in _STAGES:
        for b in range(blocks):
            p = f"encoder.layer{layer}.{b}"
            ks |= {f"{p}.conv1.weight", f"{p}.conv2.weight"}
            ks |= {f"{p}.bn{i}.{s}" for i in (1, 2) for s in bn}
            if layer > 1 and b == 0:
                ks.add(f"{p}.downsample.0.weight")
                ks |= {f"{p}.downsample.1.{s}" for s in bn}
    for i in range(5):
        for c in (1, 2):
            p = f"decoder.blocks.{i}.conv{c}"
            ks.add(f"{p}.0.weight")
            ks |= {f"{p}.1.{s}" for s in bn}
    ks |= {"segmentation_head.0.weight", "segmentation_head.0.bias"}
    if use_metadata:
        ks |= {f"enc.enc_mlp.{i}.{s}" for i in (0, 3, 6) for s in ("weight", "bias")}
    return ks


def get_module(checkpoint: str | Path) -> Mapping:
    """src/zone_detect/model.py:61-76."""
    if checkpoint is not None and os.path.isfile(checkpoint):
        weights = torch.load(checkpoint, map_location="cpu")
        if str(checkpoint).endswith(".ckpt"):
            weights = weights["state_dict"]
    else:
        print('Error with checkpoint provided: either a .ckpt with a "state_dict" key or an OrderedDict pt/pth file')
        return {}
    if "model.seg_model" in list(weights.keys())[0]:
        weights = {k.partition("model.seg_model.")[2]: v for k, v in weights.items()}
        weights = {k: v for k, v in weights.items() if k != ""}
    return weights


def check_strict(state_dict: Mapping, use_metadata: bool = False) -> None:
    """load_state_dict(strict=True) semantics (model.py:86): missing or unexpected keys are errors
    (num_batches_tracked buffers belong to the module and are accepted)."""
    have = {k for k in state_dict if not k.endswith("num_batches_tracked")}
    want = expected_keys(use_metadata)
    missing, unexpected = sorted(want - have), sorted(have - want)
    if missing or unexpected:
        raise RuntimeError("Error(s) in loading state_dict for Unet:\n"
                           + (f"\tMissing key(s) in state_dict: {missing[:8]}{'...' if len(missing) > 8 else ''}\n" if missing else "")
                           + (f"\tUnexpected key(s) in state_dict: {unexpected[:8]}{'...' if len(unexpected) > 8 else ''}\n" if unexpected else ""))


class FLAIR_ModelFactory:
    """Config -> GPU model (model.py:12-58, SegmentationModelsPytorch branch only; the HuggingFace
    Swin-UperNet branch needs `from_pretrained` downloads and is out of scope)."""

    def __init__(self, config: Mapping, device: int | torch.device = 0):
        self.config = config
        self.model_provider = config["model_framework"]["model_provider"]
        if self.model_provider != "SegmentationModelsPytorch":
            raise NotImplementedError(f"model_provider {self.model_provider!r}: only SegmentationModelsPytorch/resnet34_unet is built for B200")
        encoder, architecture = config["model_framework"]["SegmentationModelsPytorch"]["encoder_decoder"].split("_")
        if (encoder, architecture) != ("resnet34", "unet"):
            raise NotImplementedError(f"encoder_decoder {encoder}_{architecture}: only resnet34_unet is built for B200")
        self.n_channels = int(len(config["channels"]))
        self.n_classes = config["n_classes"]
        self.seg_model = _native.Context(device)

    def load_state_dict(self, state_dict: Mapping, strict: bool = True) -> None:
        if strict:
            check_strict(state_dict)
        self.seg_model.load_weights(state_dict, self.n_channels, self.n_classes, use_metadata=False)


def load_model(config: dict, device: int | torch.device = 0) -> _native.Context:
    """src/zone_detect/model.py:79-88: build, read the checkpoint, strict load. Returns the context
    (the object the hot loop calls instead of `model(imgs)`)."""
    factory = FLAIR_ModelFactory(config, device)
    state_dict = get_module(checkpoint=config["model_weights"])
    factory.load_state_dict(state_dict, strict=True)
    return factory.seg_model
