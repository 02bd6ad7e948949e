"""Inference and stitching for zone detection (mirrors src/zone_detect/compare.py).

Reference: inference() uploads a batch, runs the model, soft-maxes and copies the whole probability
tensor back (15.7 MB per tile, compare.py:27-36); stitching() crops the margin, takes argmax / max on
the host and computes a rasterio window per tile (compare.py:66-82). Here both steps are one call into
libflairb200 per zone strip: forward, soft-max maximum, argmax, margin clipping and the write into the
class map all stay on the GPU and only the uint8 maps come back.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _native


def inference(model: _native.Context, tiles: np.ndarray, tile: int) -> torch.Tensor:
    """Soft-max probabilities [n, n_classes, tile, tile] (device, fp32) for tiles of the raster currently
    attached to `model` -- the tensor the reference's inference() returns (compare.py:35), kept for the
    class_prob output type and for tests."""
    logits = model.forward_tiles(np.ascontiguousarray(tiles[:, :2]), tile)
    return torch.softmax(logits[..., :model.n_classes].permute(0, 3, 1, 2), dim=1)


def stitching(model: _native.Context, tiles: np.ndarray, tile: int, batch: int, cls_map: torch.Tensor,
              conf_map: torch.Tensor | None, map_w: int, map_row0: int = 0, stitch: str = "exact-clipping") -> None:
    """exact-clipping (compare.py:68-82): every tile writes the part of its interior it owns under the
    reference's write order (slicing_job.tile_table), band 1 = class index, band 2 = max probability
    cast to uint8."""
    if stitch != "exact-clipping":
        raise ValueError(f"stitching(): {stitch!r} is a blended method, use stitching_blend()")
    model.detect_strip(tiles, tile, batch, cls_map, conf_map, map_w, map_row0)


STITCH_METHODS = ("exact-clipping", "average", "average_weights", "max")


def stitching_blend(model: _native.Context, tiles: np.ndarray, tile: int, batch: int, cls_map: torch.Tensor,
                    conf_map: torch.Tensor | None, map_w: int, map_row0: int, stitch: str) -> None:
    """The weighted branches of stitching() (compare.py:84-138). As written they cannot run (they push
    n_classes-channel float products through the 2-band uint8 output raster and compare class indices
    where confidences are meant), so this implements what they state: every tile contributes its whole
    soft-max, clipped to the raster, with weight 1 ("average", normalised by the overlap count of
    test/tiles.py:54-94), with patch_weights(size, 0.5, "exp") ("average_weights", normalised by
    total_weights, tiles.py:111-169), or the most confident tile wins ("max"). `tiles` must hold every
    tile that touches map rows [map_row0, map_row0 + cls_map.shape[0]); accumulation stays on the GPU."""
    if stitch not in ("average", "average_weights", "max"):
        raise ValueError(f"unknown stitching method {stitch!r}")
    acc, wsum = model.blend_buffers(stitch, cls_map.shape[0], map_w)
    model.blend_strip(tiles, tile, batch, stitch, acc, wsum, map_w, map_row0)
    model.blend_finalize(stitch, acc, wsum, cls_map, conf_map)


def stitching_class_prob(model: _native.Context, tiles: np.ndarray, tile: int, batch: int, prob_map: torch.Tensor,
                         map_w: int, map_row0: int = 0) -> None:
    """output_type "class_prob" (compare.py:68-76 + dataset.py:15-21): exact clipping, every class
    probability as uint8(p * 255) into prob_map [n_classes, rows, map_w]."""
    model.detect_strip_prob(tiles, tile, batch, prob_map, map_w, map_row0)
