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
        raise NotImplementedError(f"stitching method {stitch!r}: the reference's weighted branches are not executable "
                                  "(SURVEY.md Appendix C); only exact-clipping is implemented")
    model.detect_strip(tiles, tile, batch, cls_map, conf_map, map_w, map_row0)
