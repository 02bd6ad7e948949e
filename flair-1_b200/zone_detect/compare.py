"""Inference and stitching for zone detection (mirrors src/zone_detect/compare.py): one rank's share of the hot loop.

Reference: inference() uploads a batch, runs the model, soft-maxes and copies the whole probability
tensor back (15.7 MB per tile, compare.py:27-36); stitching() crops the margin, takes argmax / max on
the host and computes a rasterio window per tile (compare.py:66-82); main.py:398-426 loops over the batches.
Here the loop over a rank's tiles is one call into libflairb200: forward, soft-max maximum, argmax, margin
clipping and the write into the class map all stay on the GPU and only the uint8 maps come back --
host to host and pipelined for the default run (detect_zone_pipelined -> fb_detect_zone_shard), with device-resident
maps for the runs that need every logit of a tile (detect_zone: blended stitching, class_prob, per-patch metrics).
"""
from __future__ import annotations

from typing import TYPE_CHECKING

import numpy as np
import torch

from .. import _native

if TYPE_CHECKING:
    from .dataset import Sliced_Dataset


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


def tiles_per_pass(config: dict) -> int:
    """Tiles per forward pass: one 512^2 tile per SM by default, scaled with the tile area so that the activation
    arena (~70 MB per 512^2 tile) stays near 10 GB for the 128 .. 2048 px tiles of the compare grid."""
    size = config["img_pixels_detection"]
    per_launch = int(config.get("tiles_per_launch", max(1, min(1024, round(148 * (512 / size) ** 2)))))
    return min(1024, max(int(config.get("batch_size", 4)), per_launch))


def detect_zone_pipelined(config: dict, model, dataset: Sliced_Dataset, my_tiles: np.ndarray, out: "OutputMap",
                          truth_rows: np.ndarray | None, truth_row0: int, cm: torch.Tensor | None) -> None:
    """One rank's share of the hot loop, host to host (fb_detect_zone_shard): the raster rows of `dataset` go up in
    chunks, the tiles run as their rows land, every finished row band of this rank's write rectangles goes straight
    to its place in `out`; with `truth_rows` (this rank's truth rows, already minus 1) the confusion matrix of the
    same rectangles is added to `cm` on the GPU (test/metrics.py:229-231 for the whole raster)."""
    if len(my_tiles) == 0:
        return
    size = config["img_pixels_detection"]
    H, W = dataset.raster_height, dataset.raster_width
    out.pin_rows(int(my_tiles[:, 3].min()), int(my_tiles[:, 5].max()))
    model.detect_zone_shard(dataset.big_image, list(range(dataset.num_bands)), W, H, dataset.row0, 0, my_tiles, size,
                            tiles_per_pass(config), out.array[0], out.array[1], W, 0, H,
                            truth=truth_rows, truth_row0=truth_row0, truth_sub=0, cm=cm)


def detect_zone(config: dict, model, dataset: Sliced_Dataset, my_tiles: np.ndarray, device: torch.device,
                stitch: str = "exact-clipping", truth_dev: torch.Tensor | None = None, map_rows: tuple[int, int] | None = None):
    """One rank's share of the runs that need whole logits (blended stitching, class_prob, per-patch metrics).
    Returns (class strip, confidence strip, first row, rows) with the strips on the device. `map_rows`: rows the
    device maps cover (default: the rows this rank owns). With `truth_dev` (truth rows of the same span, already
    minus 1) and exact clipping, config["_patch_cm"] receives the per-tile confusion matrices of
    compute_metrics_patch (main.py:349-366): each tile's own prediction over its whole margin-cropped window."""
    config["_patch_cm"] = None
    size = config["img_pixels_detection"]
    W = dataset.raster_width
    if len(my_tiles) == 0:
        e = torch.empty((0, W), dtype=torch.uint8, device=device)
        return e, e.clone(), 0, 0
    my0, my1 = map_rows or config.get("_own_rows", (int(my_tiles[:, 3].min()), int(my_tiles[:, 5].max())))
    raster_dev = dataset.big_image.to(device, non_blocking=True)
    model.set_raster(raster_dev, list(range(dataset.num_bands)), W, dataset.raster_height, row0=dataset.row0)
    batch = tiles_per_pass(config)
    if config["output_type"] == "class_prob":
        # main.py:409-426 always clips exactly for this output type (compare.py:68): n_classes planes, no band 2
        prob = torch.zeros((config["n_classes"], my1 - my0, W), dtype=torch.uint8, device=device)
        stitching_class_prob(model, my_tiles, size, batch, prob, W, my0)
        return prob, None, my0, my1 - my0
    cls = torch.zeros((my1 - my0, W), dtype=torch.uint8, device=device)
    conf = torch.zeros((my1 - my0, W), dtype=torch.uint8, device=device)
    if stitch == "exact-clipping" and truth_dev is not None:
        config["_patch_cm"] = model.detect_strip_metrics(my_tiles, np.asarray(config["_my_windows"], dtype=np.int32), size, batch,
                                                         cls, conf, W, my0, truth_dev)
    elif stitch == "exact-clipping":
        stitching(model, my_tiles, size, batch, cls, conf, W, my0, stitch)
    else:
        stitching_blend(model, my_tiles, size, batch, cls, conf, W, my0, stitch)
    return cls, conf, my0, my1 - my0
