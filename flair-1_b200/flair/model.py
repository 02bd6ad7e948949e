"""Model factory for the patch path (mirrors src/flair/model.py + load_checkpoint of src/flair/main.py).

`FLAIR_ModelFactory(config)` holds a libflairb200 context instead of an smp module; the metadata MLP
(model.py:74-96) and its broadcast-add onto the bottleneck feature (model.py:57-60) run inside the
library (K4 + the layer4 epilogue). The reference constructor tests an undefined name when
use_metadata is True (model.py:32); the evident intent is implemented.
"""
from __future__ import annotations

import os
from typing import Mapping

import numpy as np
import torch

from .. import _native
from ..zone_detect.model import expected_keys


def _strip_prefixes(state_dict: Mapping) -> dict:
    """Lightning checkpoints of the flair CLI prefix everything with "model." (seg_module.model =
    FLAIR_ModelFactory): model.seg_model.* -> *, model.enc.enc_mlp.* -> enc.enc_mlp.*."""
    out = {}
    for k, v in state_dict.items():
        if k.startswith("model.seg_model."):
            out[k[len("model.seg_model."):]] = v
        elif k.startswith("seg_model."):
            out[k[len("seg_model."):]] = v
        elif k.startswith("model.enc."):
            out[k[len("model."):]] = v
        else:
            out[k] = v
    return out


class FLAIR_ModelFactory:
    def __init__(self, config: Mapping, device: int | torch.device = 0):
        self.model_provider = config["model_framework"]["model_provider"]
        self.use_metadata = bool(config["use_metadata"])
        if self.model_provider != "SegmentationModelsPytorch":
            raise NotImplementedError(f"model_provider {self.model_provider!r}: only SegmentationModelsPytorch/resnet34_unet is built for B200")
        encoder, architecture = config["model_framework"]["SegmentationModelsPytorch"]["encoder_decoder"].split("_")
        if (encoder, architecture) != ("resnet34", "unet"):
            raise NotImplementedError(f"encoder_decoder {encoder}_{architecture}: only resnet34_unet is built for B200")
        self.n_channels = int(len(config["channels"]))
        self.n_classes = int(len(config["classes"]))
        self.classes = config["classes"]
        self.seg_model = _native.Context(device)
        self.loaded = False

    def state_dict_shapes(self) -> dict:
        """Shapes load_checkpoint compares against (only the class-dependent tensors matter)."""
        return {"segmentation_head.0.weight": (self.n_classes, 16, 3, 3), "segmentation_head.0.bias": (self.n_classes,),
                "criterion.weight": (self.n_classes,)}

    def load_state_dict(self, state_dict: Mapping, strict: bool = False) -> None:
        sd = _strip_prefixes(state_dict)
        want = expected_keys(self.use_metadata)
        missing = sorted(want - set(sd))
        if missing:
            # strict=False in the reference silently keeps random weights for missing keys; a GPU model with
            # uninitialised layers is never what anybody wants, so this is an error here.
            raise RuntimeError(f"checkpoint misses {len(missing)} tensors of the U-Net, e.g. {missing[:4]}")
        self.seg_model.load_weights(sd, self.n_channels, self.n_classes, use_metadata=self.use_metadata)
        self.loaded = True


def load_checkpoint(conf, seg_module: FLAIR_ModelFactory, exit_on_fail: bool = False) -> None:
    """src/flair/main.py:77-146: read .ckpt/.pth, probe the class count on 'classifier.weight' /
    'criterion.weight', on mismatch truncate-and-zero the mis-shaped `head` tensors (and rebuild
    `criterion.weight` from the config), then load."""
    print()
    print("###############################################################")
    ckpt_file_path = conf["paths"]["ckpt_model_path"]
    num_classes = len(conf["classes"])
    if ckpt_file_path and os.path.isfile(ckpt_file_path):
        checkpoint = torch.load(ckpt_file_path, map_location="cpu")
        if ckpt_file_path.endswith(".ckpt"):
            state_dict = checkpoint.get("state_dict", checkpoint)
        elif ckpt_file_path.endswith(".pth") or ckpt_file_path.endswith(".pt"):
            state_dict = checkpoint
        else:
            print("Invalid file extension.")
            if exit_on_fail:
                raise SystemExit()
            return
        ckpt_num_classes = None
        for k, v in state_dict.items():
            if "classifier.weight" in k or "criterion.weight" in k:
                ckpt_num_classes = v.shape[0]
                break
        if ckpt_num_classes is not None and ckpt_num_classes == num_classes:
            seg_module.load_state_dict(state_dict, strict=False)
            print("--------------- Loaded model weights from checkpoint with matching number of classes. ---------------")
        else:
            print(f"Number of classes in checkpoint ({ckpt_num_classes}) does not match the current number of classes ({num_classes}). Proceeding with modifications.")
            state_dict = dict(state_dict)
            shapes = seg_module.state_dict_shapes()
            stripped = {k: kk for k in state_dict for kk in [next(iter(_strip_prefixes({k: 0})))]}
            ignored = [k for k, v in state_dict.items() if stripped[k] in shapes and tuple(v.shape) != shapes[stripped[k]]]
            ignored = [i for i in ignored if any(x in i for x in ["head", "criterion"])]
            for k in ignored:
                print("-", k, "has been modified.")
                print(state_dict[k].shape, "  ->  ", flush=True, end="")
                if "criterion" in k:
                    state_dict[k] = torch.FloatTensor([conf["classes"][i][0] for i in conf["classes"]])
                else:
                    state_dict[k] = 0 * np.abs(state_dict[k][0:num_classes])
                print(state_dict[k].shape)
            seg_module.load_state_dict(state_dict, strict=False)
        print("###############################################################")
    else:
        print("Invalid checkpoint file path.")
        if exit_on_fail:
            raise SystemExit()
        print("###############################################################")
    print()
