"""CPU restatement of the patch-level predict path. TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows src/flair/data_loader.py:9-30, src/flair/tasks_utils.py:158-213, src/flair/task_module.py:206-213,
src/zone_detect/model.py:61-76 and src/flair/main.py:77-146.
"""
from __future__ import annotations

from typing import Mapping, Sequence

import numpy as np
import torch


def norm(in_img: np.ndarray, norm_type: str = None, means: Sequence[float] = (), stds: Sequence[float] = ()):
    """src/flair/data_loader.py:9-30 ('scaling' = skimage.img_as_float(uint8) = x / 255 float64)."""
    if norm_type not in ["scaling", "custom", "without"]:
        print("Normalization argument should be 'scaling', 'custom' or 'without'.")
        raise SystemExit()
    if norm_type == "custom":
        if len(means) != len(stds):
            print("If custom, provided normalization means and stds should be of same lenght.")
            raise SystemExit()
        in_img = in_img.astype(np.float64)
        for i in range(in_img.shape[0]):
            in_img[i] -= means[i]
            in_img[i] /= stds[i]
    elif norm_type == "scaling":
        in_img = in_img.astype(np.float64) / 255.0
    return in_img


def encode_metadata(entry: Mapping) -> list:
    """One 45-float vector of parsing_metadata (src/flair/tasks_utils.py:160-210), including the
    `int(month)-1/12` operator-precedence quirk at lines 189-190 (trained weights depend on it)."""
    def coordenc_opt(coords, enc_size=32):
        d = int(enc_size / 2)
        d_i = np.arange(0, d / 2)
        freq = 1 / (10e7 ** (2 * d_i / d))
        x, y = coords[0] / 10e7, coords[1] / 10e7
        enc = np.zeros(d * 2)
        enc[0:d:2] = np.sin(x * freq)
        enc[1:d:2] = np.cos(x * freq)
        enc[d::2] = np.sin(y * freq)
        enc[d + 1::2] = np.cos(y * freq)
        return list(enc)

    def norm_alti(alti):
        return [(alti - 0) / (3164.9099121094 - 0)]

    def format_cam(cam):
        return [1, 0] if "UCE" in cam else [0, 1]

    def cyclical_enc_datetime(date, time):
        def nrm(num):
            return (num - (-1)) / (1 - (-1))
        year, month, day = date.split("-")
        enc_y = {"2018": [1, 0, 0, 0], "2019": [0, 1, 0, 0], "2020": [0, 0, 1, 0], "2021": [0, 0, 0, 1]}[year]
        sin_month = np.sin(2 * np.pi * (int(month) - 1 / 12))
        cos_month = np.cos(2 * np.pi * (int(month) - 1 / 12))
        sin_day = np.sin(2 * np.pi * (int(day) / 31))
        cos_day = np.cos(2 * np.pi * (int(day) / 31))
        h, m = time.split("h")
        sec_day = int(h) * 3600 + int(m) * 60
        sin_time = np.sin(2 * np.pi * (sec_day / 86400))
        cos_time = np.cos(2 * np.pi * (sec_day / 86400))
        return enc_y + [nrm(sin_month), nrm(cos_month), nrm(sin_day), nrm(cos_day), nrm(sin_time), nrm(cos_time)]

    return (coordenc_opt([entry["patch_centroid_x"], entry["patch_centroid_y"]]) + norm_alti(entry["patch_centroid_z"])
            + format_cam(entry["camera"]) + cyclical_enc_datetime(entry["date"], entry["time"]))


def parsing_metadata(image_path_list: Sequence[str], metadata_dict: Mapping) -> list:
    """src/flair/tasks_utils.py:203-213: key = basename without the 4-char extension."""
    return [encode_metadata(metadata_dict[img.split("/")[-1][:-4]]) for img in image_path_list]


def get_module(weights: Mapping, is_ckpt: bool) -> Mapping:
    """src/zone_detect/model.py:61-76 after torch.load: unwrap "state_dict" for .ckpt, strip the
    "model.seg_model." prefix when the FIRST key carries it, drop keys that become empty."""
    if is_ckpt:
        weights = weights["state_dict"]
    if "model.seg_model" in list(weights.keys())[0]:
        weights = {k.partition("model.seg_model.")[2]: v for k, v in weights.items()}
        weights = {k: v for k, v in weights.items() if k != ""}
    return weights


def load_checkpoint_state(state_dict: dict, model_state: Mapping, classes: Mapping) -> dict:
    """The state_dict surgery of src/flair/main.py:106-138 (before load_state_dict(strict=False)):
    when the class count probed on 'classifier.weight' / 'criterion.weight' differs (or is absent),
    mis-shaped `head` tensors are truncated to num_classes and zeroed, `criterion` gets the config
    weights."""
    num_classes = len(classes)
    ckpt_num_classes = None
    for k, v in state_dict.items():
        if "classifier.weight" in k or "criterion.weight" in k:
            ckpt_num_classes = v.shape[0]
            break
    if ckpt_num_classes is not None and ckpt_num_classes == num_classes:
        return state_dict
    state_dict = dict(state_dict)
    ignored = [k for k, v in state_dict.items() if k in model_state and v.shape != model_state[k].shape]
    ignored = [i for i in ignored if any(x in i for x in ["head", "criterion"])]
    for k in ignored:
        if "criterion" in k:
            state_dict[k] = torch.FloatTensor([classes[i][0] for i in classes])
        else:
            state_dict[k] = 0 * np.abs(state_dict[k][0:num_classes])
    return state_dict


def predict_step(model: torch.nn.Module, img: torch.Tensor, mtd=None) -> torch.Tensor:
    """src/flair/task_module.py:206-213: softmax then argmax over classes -> int64 [B,H,W]."""
    with torch.no_grad():
        logits = model(img, mtd) if mtd is not None else model(img)
    return torch.argmax(torch.softmax(logits, dim=1), dim=1)
