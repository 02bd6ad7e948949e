"""CSV -> path dictionaries and the 45-float metadata encoding (mirrors src/flair/tasks_utils.py)."""
from __future__ import annotations

import json
import os

import numpy as np
import pandas as pd


def gather_paths(config, split="train"):
    """src/flair/tasks_utils.py:127-155 (same messages, same SystemExit)."""
    key = {"train": "train_csv", "val": "val_csv", "test": "test_csv"}[split]
    path = config["paths"][key]
    if path is not None and os.path.isfile(path) and path.endswith(".csv"):
        paths = pd.read_csv(path, header=None)
    else:
        print({"train": "Invalid .csv training file path.", "val": "Invalid .csv val file path.",
               "test": "Invalid .csv test file path."}[split])
        raise SystemExit()
    images = paths.iloc[:, 0].tolist()
    labels = paths.iloc[:, 1].tolist()
    metadata = parsing_metadata(images, config) if config["use_metadata"] == True else []  # noqa: E712
    return {"IMG": images, "MSK": labels, "MTD": metadata}


def parsing_metadata(image_path_list, config):
    """src/flair/tasks_utils.py:158-213: positional encoding of the patch centroid (32), normalised
    altitude (1), camera one-hot (2), year one-hot (4) and cyclical month/day/time (6) = 45 floats.
    The month term keeps the reference's operator precedence, `int(month) - 1/12` (lines 189-190):
    checkpoints were trained with it."""
    def coordenc_opt(coords, enc_size=32) -> list:
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

    def norm_alti(alti) -> list:
        min_alti, max_alti = 0, 3164.9099121094
        return [(alti - min_alti) / (max_alti - min_alti)]

    def format_cam(cam: str) -> list:
        return [[1, 0] if "UCE" in cam else [0, 1]][0]

    def cyclical_enc_datetime(date: str, time: str) -> list:
        def norm(num: float) -> float:
            return (num - (-1)) / (1 - (-1))
        year, month, day = date.split("-")
        if year == "2018":
            enc_y = [1, 0, 0, 0]
        elif year == "2019":
            enc_y = [0, 1, 0, 0]
        elif year == "2020":
            enc_y = [0, 0, 1, 0]
        elif year == "2021":
            enc_y = [0, 0, 0, 1]
        sin_month = np.sin(2 * np.pi * (int(month) - 1 / 12))
        cos_month = np.cos(2 * np.pi * (int(month) - 1 / 12))
        sin_day = np.sin(2 * np.pi * (int(day) / 31))
        cos_day = np.cos(2 * np.pi * (int(day) / 31))
        h, m = time.split("h")
        sec_day = int(h) * 3600 + int(m) * 60
        sin_time = np.sin(2 * np.pi * (sec_day / 86400))
        cos_time = np.cos(2 * np.pi * (sec_day / 86400))
        return enc_y + [norm(sin_month), norm(cos_month), norm(sin_day), norm(cos_day), norm(sin_time), norm(cos_time)]

    with open(config["paths"]["path_metadata_aerial"], "r") as f:
        metadata_dict = json.load(f)
    MTD = []
    for img in image_path_list:
        curr_img = img.split("/")[-1][:-4]
        enc_coords = coordenc_opt([metadata_dict[curr_img]["patch_centroid_x"], metadata_dict[curr_img]["patch_centroid_y"]])
        enc_alti = norm_alti(metadata_dict[curr_img]["patch_centroid_z"])
        enc_camera = format_cam(metadata_dict[curr_img]["camera"])
        enc_temporal = cyclical_enc_datetime(metadata_dict[curr_img]["date"], metadata_dict[curr_img]["time"])
        MTD.append(enc_coords + enc_alti + enc_camera + enc_temporal)
    return MTD
