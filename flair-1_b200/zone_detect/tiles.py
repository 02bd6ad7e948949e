"""Stride and overlap-weight helpers (src/zone_detect/test/tiles.py, a runtime module of the reference
despite its directory name)."""
from __future__ import annotations

import numpy as np


def get_stride(config: dict) -> list:
    """src/zone_detect/test/tiles.py:4-14: default stride = img_pixels_detection - 2*margin; with
    `overlap_strat` the strides are fractions of the tile size from strategies.tiling.stride_range."""
    img_size = config["img_pixels_detection"]
    if not config.get("overlap_strat"):
        return [int(img_size - 2 * config["margin"])]
    return [int(i * img_size) for i in config["strategies"]["tiling"]["stride_range"]]


def patch_weights(patch_size: int, sigma: float = 0.5, mode: str = "exp") -> np.ndarray:
    """tiles.py:97-108, mode "exp" only (the one stitching() uses, compare.py:126):
    exp(-sigma * chebyshev_distance_to_centre / max_distance)."""
    if mode != "exp":
        raise ValueError("only mode='exp' is supported (the reference's 'gaussian' branch is not a Gaussian)")
    center = patch_size // 2
    y, x = np.ogrid[:patch_size, :patch_size]
    dist = np.maximum(np.abs(y - center), np.abs(x - center))
    return np.exp(-dist / dist.max() * sigma)
