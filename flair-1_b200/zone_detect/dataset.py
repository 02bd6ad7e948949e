"""Raster access for zone detection (mirrors src/zone_detect/dataset.py).

The reference's Sliced_Dataset re-opens the raster in DataLoader workers and reads one boundless
window per tile through GDAL, normalising in float64 on the host (dataset.py:68-113). Here the
selected bands of the raster (or of the strip a rank owns) are read ONCE into pinned host memory and
uploaded; tile extraction, zero fill outside the raster and normalisation happen on the GPU
(csrc/elementwise.cu, K1), so `__getitem__` only exists as a description of what a tile is.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from .. import geotiff


def convert(img: np.ndarray, img_type: str) -> np.ndarray:
    """src/zone_detect/dataset.py:11-34 (host version, used for class_prob output and by tests)."""
    if img_type == "class_prob":
        if img.max() > 1:
            info = np.iinfo(img.dtype)
            img = img.astype(np.float32) / info.max
        return (img * 255).astype(np.uint8)
    elif img_type == "argmax":
        img_arg = np.expand_dims(np.argmax(img, axis=0).astype(np.uint8), axis=0)
        img_max = np.expand_dims(np.max(img, axis=0).astype(np.float32), axis=0)
        return np.concatenate([img_arg, img_max], axis=0)
    print("The output type has not been interpreted.")
    return img


class Sliced_Dataset:
    """Holds the raster rows a rank needs, band-planar uint8 in pinned host memory.

    dataframe: int32 [n, 6] tile table from slicing_job.tile_table (x0, y0, wx0, wy0, wx1, wy1).
    bands: 1-based band list (config "channels"), norma_dict: config "norma_task"."""

    def __init__(self, dataframe: np.ndarray, img_path, resolution, bands: Sequence[int], patch_detection_size: int,
                 norma_dict: list, row_range: tuple[int, int] | None = None) -> None:
        self.dataframe = dataframe
        self.img_path = img_path
        self.resolution = resolution
        self.bands = list(bands)
        self.num_bands = len(bands)
        self.height = self.width = patch_detection_size
        self.norma_dict = norma_dict[0]
        self.norm_type = self.norma_dict["norm_type"]
        self.norm_means = self.norma_dict.get("norm_means", [])
        self.norm_stds = self.norma_dict.get("norm_stds", [])
        if self.norm_type not in ("custom", "scaling"):
            print("Invalid normalization type: should be custom or scaling. Going with scaling.")
            self.norm_type = "scaling"
        if self.norm_type == "custom" and len(self.norm_means) != len(self.norm_stds):
            print("If custom, provided normalization means and stds should be of the same length. Going with scaling.")
            self.norm_type = "scaling"
        info = geotiff.read_info(img_path)
        self.raster_width, self.raster_height = info.width, info.height
        r0, r1 = row_range if row_range is not None else (0, info.height)
        self.row0, self.rows = max(r0, 0), min(r1, info.height) - max(r0, 0)
        self.big_image = torch.empty((self.num_bands, self.rows, info.width), dtype=torch.uint8,
                                     pin_memory=torch.cuda.is_available())
        geotiff.read(img_path, bands=self.bands, window=(0, self.row0, info.width, self.rows), out=self.big_image.numpy())

    def __len__(self) -> int:
        return len(self.dataframe)

    def close_raster(self) -> None:
        self.big_image = None

    def __getitem__(self, index: int) -> dict:
        """Raw (un-normalised) tile `index` with zero fill outside the raster -- for inspection only."""
        x0, y0 = int(self.dataframe[index, 0]), int(self.dataframe[index, 1])
        T = self.height
        patch = np.zeros((self.num_bands, T, T), np.uint8)
        ra = self.big_image.numpy()
        r0, r1 = max(y0, self.row0), min(y0 + T, self.row0 + self.rows)
        c0, c1 = max(x0, 0), min(x0 + T, self.raster_width)
        if r1 > r0 and c1 > c0:
            patch[:, r0 - y0:r1 - y0, c0 - x0:c1 - x0] = ra[:, r0 - self.row0:r1 - self.row0, c0:c1]
        return {"image": torch.from_numpy(patch), "index": torch.tensor([index], dtype=torch.int32)}
