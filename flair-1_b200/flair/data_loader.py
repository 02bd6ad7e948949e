"""Patch reading and normalisation semantics (mirrors src/flair/data_loader.py).

The reference normalises on the host in float64 (`norm`, data_loader.py:9-30). Here `norm` only
validates its arguments -- the normalisation itself is the bf16 look-up table built by fb_set_norm from
the same float64 formula and applied on the GPU while the patch is copied (csrc/elementwise.cu, K1).
"""
from __future__ import annotations

from typing import Sequence

import numpy as np

from .. import geotiff


def norm(in_img: np.ndarray, norm_type: str = None, means: Sequence[float] = (), stds: Sequence[float] = ()):
    """Argument checks of src/flair/data_loader.py:9-30 (same messages, same SystemExit); returns the
    image untouched (uint8) because the arithmetic happens on the device."""
    if norm_type not in ["scaling", "custom", "without"]:
        print("Normalization argument should be 'scaling', 'custom' or 'without'.")
        raise SystemExit()
    if norm_type == "custom" and len(means) != len(stds):
        print("If custom, provided normalization means and stds should be of same lenght.")
        raise SystemExit()
    return in_img


class predict_dataset:
    """src/flair/data_loader.py:100-144: whole-patch read of the selected bands (+ metadata vector)."""

    def __init__(self, dict_files: dict, channels: list = [1, 2, 3, 4, 5], num_classes: int = 13, use_metadata: bool = True,
                 norm_type: str = "scaling", means: list = [], stds: list = []):
        self.list_imgs = np.array(dict_files["IMG"])
        self.num_classes = num_classes
        self.use_metadata = use_metadata
        if use_metadata:
            self.list_metadata = np.array(dict_files["MTD"])
        self.channels = channels
        self.norm_type = norm_type
        self.means = means
        self.stds = stds

    def read_img(self, raster_file: str) -> np.ndarray:
        return geotiff.read(raster_file, bands=self.channels)

    def __len__(self):
        return len(self.list_imgs)

    def __getitem__(self, index):
        image_file = self.list_imgs[index]
        img = norm(self.read_img(image_file), norm_type=self.norm_type, means=self.means, stds=self.stds)
        item = {"img": img, "id": image_file}
        if self.use_metadata:
            item["mtd"] = np.asarray(self.list_metadata[index], dtype=np.float32)
        return item
