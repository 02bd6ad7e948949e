"""Restatement of the overlap / weight-map helpers. TEST INFRASTRUCTURE (see oracle/__init__.py).

src/zone_detect/test/tiles.py:30-169, pixel space. Only mode="exp" of patch_weights is restated (the
one the pipeline uses, compare.py:126; the "gaussian" branch is not a Gaussian).
`blend_zone` is OUR restatement of the *intent* of the weighted branch of stitching()
(compare.py:84-138), which is not executable as written (Appendix C): fp32 soft-max probabilities
weighted by patch_weights, summed over every tile covering a pixel, argmax of the sum (the division
by total_weights does not change the argmax and is applied only to the returned probabilities).
"""
from __future__ import annotations

import numpy as np


def get_tile_coord(start: int, end: int, limit: int, patch_size: int, stride: int) -> list:
    """tiles.py:30-51."""
    max_coord = limit - patch_size
    if max_coord < 0:
        return []
    tile_starts = set()
    for i in range(0, end, stride):
        if i + patch_size > limit:
            i = max_coord
        tile_starts.add(i)
    return [t for t in tile_starts if t + patch_size > start and t < end]


def patch_overlap(image_size, patch_size: int, query_bounds, stride: int) -> np.ndarray:
    """tiles.py:54-94."""
    x_min, x_max, y_min, y_max = query_bounds
    overlap_map = np.zeros((y_max - y_min, x_max - x_min), dtype=np.uint8)
    sx, sy = image_size
    for ty in get_tile_coord(y_min, y_max, sy, patch_size, stride):
        for tx in get_tile_coord(x_min, x_max, sx, patch_size, stride):
            ty_, tx_ = min(ty, sy - patch_size), min(tx, sx - patch_size)
            iy0, iy1 = max(ty_, y_min), min(ty_ + patch_size, y_max)
            ix0, ix1 = max(tx_, x_min), min(tx_ + patch_size, x_max)
            if iy1 > iy0 and ix1 > ix0:
                overlap_map[iy0 - y_min:iy1 - y_min, ix0 - x_min:ix1 - x_min] += 1
    return overlap_map


def patch_weights(patch_size: int, sigma: float = 0.5) -> np.ndarray:
    """tiles.py:97-108 with mode="exp": exp(-cheb/cheb.max() * sigma)."""
    center = patch_size // 2
    y, x = np.ogrid[:patch_size, :patch_size]
    dist = np.maximum(np.abs(y - center), np.abs(x - center))
    return np.exp(-dist / dist.max() * sigma)


def total_weights(image_size, patch_size: int, query_bounds, stride: int) -> np.ndarray:
    """tiles.py:111-169 (the map; the second tuple member `steps` is debugging output)."""
    x_min, x_max, y_min, y_max = query_bounds
    sx, sy = image_size
    out = np.zeros((y_max - y_min, x_max - x_min), dtype=np.float32)
    w = patch_weights(patch_size, 0.5)
    for ty in get_tile_coord(y_min, y_max, sy, patch_size, stride):
        for tx in get_tile_coord(x_min, x_max, sx, patch_size, stride):
            ty_, tx_ = min(ty, sy - patch_size), min(tx, sx - patch_size)
            iy0, iy1 = max(ty_, y_min), min(ty_ + patch_size, y_max)
            ix0, ix1 = max(tx_, x_min), min(tx_ + patch_size, x_max)
            if iy1 > iy0 and ix1 > ix0:
                out[iy0 - y_min:iy1 - y_min, ix0 - x_min:ix1 - x_min] += w[iy0 - ty_:iy1 - ty_, ix0 - tx_:ix1 - tx_]
    return out
