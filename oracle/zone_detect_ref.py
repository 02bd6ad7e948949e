"""CPU restatement of the reference's zone_detect path. TEST INFRASTRUCTURE (see oracle/__init__.py).

Each function cites the reference lines it follows. rasterio / geopandas / shapely are absent, so the
raster is a numpy array plus (min_x, max_y, resolution) georeferencing, the "GeoDataFrame" is a list
of dicts with the same columns, and windows are computed with the same affine arithmetic rasterio
applies to a north-up raster (row = (max_y - y) / res, col = (x - min_x) / res).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch


@dataclass
class GeoRaster:
    """A north-up georeferenced uint8 raster held in memory: data[band, row, col]."""
    data: np.ndarray
    min_x: float = 0.0
    max_y: float = 0.0
    res: float = 1.0

    @property
    def height(self) -> int:
        return self.data.shape[1]

    @property
    def width(self) -> int:
        return self.data.shape[2]

    @property
    def bounds(self) -> Tuple[float, float, float, float]:  # rasterio order: left, bottom, right, top
        return (self.min_x, self.max_y - self.height * self.res, self.min_x + self.width * self.res, self.max_y)


# --------------------------------------------------------------------------------------------- a1
def get_stride(config: dict) -> list:
    """src/zone_detect/test/tiles.py:4-14."""
    img_size = config["img_pixels_detection"]
    if not config.get("overlap_strat"):
        return [int(img_size - 2 * config["margin"])]
    return [int(i * img_size) for i in config["strategies"]["tiling"]["stride_range"]]


def slice_extent(bounds: Sequence[float], res: Tuple[float, float], patch_size: int, margin: int, stride: int) -> List[dict]:
    """src/zone_detect/slicing_job.py:36-106, same float arithmetic, same iteration and de-duplication
    order. `geometry` is the (x_min, y_min, x_max, y_max) bounds of the margin-expanded square
    (what shapely's box(...).bounds returns at dataset.py:92)."""
    min_x, min_y, max_x, max_y = bounds
    resolution_x, resolution_y = map(lambda r: abs(round(r, 5)), res)                # :33
    geo_output_w, geo_output_h = patch_size * resolution_x, patch_size * resolution_y  # :36
    geo_margin_x, geo_margin_y = margin * resolution_x, margin * resolution_y          # :37
    if stride:
        geo_step = [stride * resolution_x, stride * resolution_y]                       # :40
    else:
        geo_step = [geo_output_w - (2 * geo_margin_x), geo_output_h - (2 * geo_margin_y)]
    rows, seen = [], set()
    X = np.arange(min_x - geo_margin_x, max_x + geo_margin_x, geo_step[0])             # :51
    Y = np.arange(min_y - geo_margin_y, max_y + geo_margin_y, geo_step[1])             # :52
    for x_coord in X:
        if x_coord + geo_output_w > max_x + geo_margin_x:                               # :57
            x_coord = max_x + geo_margin_x - geo_output_w
        for y_coord in Y:
            if y_coord + geo_output_h > max_y + geo_margin_y:                           # :62
                y_coord = max_y + geo_margin_y - geo_output_h
            left = x_coord + geo_margin_x                                               # :67-70
            right = min(x_coord + geo_output_w - geo_margin_x, max_x)
            bottom = y_coord + geo_margin_y
            top = min(y_coord + geo_output_h - geo_margin_y, max_y)
            key = (round(left, 6), round(bottom, 6), round(right, 6), round(top, 6))    # :78-83
            if key in seen:
                continue
            seen.add(key)
            rows.append({"left": left, "bottom": bottom, "right": right, "top": top,
                         "left_o": min_x, "bottom_o": min_y, "right_o": max_x, "top_o": max_y,
                         "geometry": (x_coord, y_coord, x_coord + geo_output_w, y_coord + geo_output_h)})
    return rows


# --------------------------------------------------------------------------------------------- a2
def normalization(in_img: np.ndarray, norm_type: str, means: Sequence[float], stds: Sequence[float]) -> np.ndarray:
    """src/zone_detect/dataset.py:68-88. `custom` is computed in float64; anything else is
    skimage.img_as_float(uint8) == x / 255 in float64."""
    if norm_type == "custom":
        if len(means) != len(stds):
            return in_img.astype(np.float64) / 255.0
        img = in_img.astype(np.float64)
        for i in range(in_img.shape[0]):
            img[i] = (img[i] - means[i]) / stds[i]
        return img
    return in_img.astype(np.float64) / 255.0


def read_window_boundless(r: GeoRaster, bands: Sequence[int], geom_bounds: Sequence[float], size: int) -> np.ndarray:
    """rasterio.windows.from_bounds + src.read(indexes, window, out_shape, boundless=True)
    (dataset.py:92-104) for a pixel-aligned window whose shape equals out_shape (no resampling):
    pixels outside the raster are filled with 0 *before* normalisation. bands are 1-based."""
    x_min, y_min, x_max, y_max = geom_bounds
    col0 = int(round((x_min - r.min_x) / r.res))
    row0 = int(round((r.max_y - y_max) / r.res))
    out = np.zeros((len(bands), size, size), dtype=r.data.dtype)
    r0, r1 = max(row0, 0), min(row0 + size, r.height)
    c0, c1 = max(col0, 0), min(col0 + size, r.width)
    if r1 > r0 and c1 > c0:
        out[:, r0 - row0:r1 - row0, c0 - col0:c1 - col0] = r.data[[b - 1 for b in bands], r0:r1, c0:c1]
    return out


# --------------------------------------------------------------------------------------------- a6
def convert(img: np.ndarray, img_type: str) -> np.ndarray:
    """src/zone_detect/dataset.py:11-34."""
    if img_type == "class_prob":
        if img.max() > 1:
            info = np.iinfo(img.dtype)
            img = img.astype(np.float32) / info.max
        return (img * 255).astype(np.uint8)
    if img_type == "argmax":
        img_arg = np.expand_dims(np.argmax(img, axis=0).astype(np.uint8), axis=0)
        img_max = np.expand_dims(np.max(img, axis=0).astype(np.float32), axis=0)
        return np.concatenate([img_arg, img_max], axis=0)
    return img


def window_of_box(r: GeoRaster, left: float, right: float, bottom: float, top: float) -> Tuple[int, int, int, int]:
    """rasterio.features.geometry_window(out, [box], pixel_precision=6).round_shape(op="ceil")
    (compare.py:79-81) for a north-up raster: returns (col_off, row_off, width, height)."""
    left, right, bottom, top = (round(c, 3) for c in (left, right, bottom, top))       # compare.py:66
    c0 = round((left - r.min_x) / r.res, 6)
    c1 = round((right - r.min_x) / r.res, 6)
    r0 = round((r.max_y - top) / r.res, 6)
    r1 = round((r.max_y - bottom) / r.res, 6)
    col_off, row_off = int(np.floor(c0)), int(np.floor(r0))
    return col_off, row_off, int(np.ceil(round(c1 - col_off, 4))), int(np.ceil(round(r1 - row_off, 4)))


def stitching_exact_clipping(prediction: np.ndarray, margin: int, img_size: int, output_type: str) -> np.ndarray:
    """compare.py:68-76: crop the margins then convert()."""
    prediction = prediction[:, margin:img_size - margin, margin:img_size - margin]
    return convert(prediction, output_type)


# --------------------------------------------------------------------------------------------- loop
def run_zone(model: torch.nn.Module, raster: GeoRaster, config: dict, batch_size: int = 4,
             tile_indices: Optional[Sequence[int]] = None, return_probs: bool = False):
    """The default branch of run_pipeline (src/zone_detect/main.py:386-433): slice, read + normalise,
    forward, softmax (compare.py:27-36), exact clipping, write band 1 (class) and band 2 (max
    probability cast to uint8 the way GDAL casts float32 -> Byte: round half up, clamp) in tile order,
    later tiles overwriting earlier ones. Returns (class_map uint8 [H,W], conf_map uint8 [H,W], rows).
    tile_indices restricts the loop to a subset (untouched pixels stay 0)."""
    size, margin = config["img_pixels_detection"], config["margin"]
    stride = get_stride(config)[0]
    rows = slice_extent((raster.bounds[0], raster.bounds[1], raster.bounds[2], raster.bounds[3]),
                        (raster.res, raster.res), size, margin, stride)
    nt = config["norma_task"][0]
    cls_map = np.zeros((raster.height, raster.width), np.uint8)
    conf_map = np.zeros((raster.height, raster.width), np.uint8)
    probs_out: Dict[int, np.ndarray] = {}
    order = list(range(len(rows))) if tile_indices is None else list(tile_indices)
    model.eval()
    for s in range(0, len(order), batch_size):
        idx = order[s:s + batch_size]
        imgs = []
        for i in idx:
            patch = read_window_boundless(raster, config["channels"], rows[i]["geometry"], size)
            imgs.append(torch.as_tensor(normalization(patch, nt["norm_type"], nt["norm_means"], nt["norm_stds"]),
                                        dtype=torch.float))                              # dataset.py:111
        with torch.no_grad():
            logits = model(torch.stack(imgs))
        predictions = torch.softmax(logits, dim=1).cpu().numpy()                          # compare.py:35-36
        for i, prediction in zip(idx, predictions):
            if return_probs:
                probs_out[i] = prediction
            pred = stitching_exact_clipping(prediction, margin, size, "argmax")
            col, row, w, h = window_of_box(raster, rows[i]["left"], rows[i]["right"], rows[i]["bottom"], rows[i]["top"])
            # out.write_band([1, 2], prediction, window): the array is (2, 256, 256); a window smaller than
            # the array never happens on this path (interior boxes are always size-2*margin wide)
            cls_map[row:row + h, col:col + w] = pred[0][:h, :w].astype(np.uint8)
            conf_map[row:row + h, col:col + w] = np.clip(np.floor(pred[1][:h, :w] + 0.5), 0, 255).astype(np.uint8)
    if return_probs:
        return cls_map, conf_map, rows, probs_out
    return cls_map, conf_map, rows


# --------------------------------------------------------------------------------------------- f2
def run_zone_class_prob(model: torch.nn.Module, raster: GeoRaster, config: dict, batch_size: int = 4) -> np.ndarray:
    """output_type "class_prob" of the default branch (main.py:229, 409-426 with dataset.py:15-21): the
    margin-cropped soft-max of every tile as uint8(p * 255) (truncation), band k+1 = class k.
    Returns uint8 [n_classes, H, W]."""
    size, margin = config["img_pixels_detection"], config["margin"]
    stride = get_stride(config)[0]
    rows = slice_extent(raster.bounds, (raster.res, raster.res), size, margin, stride)
    nt = config["norma_task"][0]
    out = np.zeros((config["n_classes"], raster.height, raster.width), np.uint8)
    model.eval()
    for s in range(0, len(rows), batch_size):
        idx = list(range(s, min(s + batch_size, len(rows))))
        imgs = [torch.as_tensor(normalization(read_window_boundless(raster, config["channels"], rows[i]["geometry"], size),
                                              nt["norm_type"], nt["norm_means"], nt["norm_stds"]), dtype=torch.float)
                for i in idx]
        with torch.no_grad():
            predictions = torch.softmax(model(torch.stack(imgs)), dim=1).cpu().numpy()
        for i, prediction in zip(idx, predictions):
            pred = stitching_exact_clipping(prediction, margin, size, "class_prob")
            col, row, w, h = window_of_box(raster, rows[i]["left"], rows[i]["right"], rows[i]["bottom"], rows[i]["top"])
            out[:, row:row + h, col:col + w] = pred[:, :h, :w]
    return out


# --------------------------------------------------------------------------------------------- a8
def run_zone_blend(model: torch.nn.Module, raster: GeoRaster, config: dict, method: str, batch_size: int = 4):
    """What the weighted branches of `stitching` (compare.py:84-138) are after. They are not executable as
    written (they accumulate ncls-channel float products through the 2-band uint8 output raster and compare
    class indices instead of confidences), so this follows their stated intent:

      average          out = sum_t p_t / count          count = patch_overlap   (tiles.py:54-94)
      average_weights  out = sum_t p_t * w / sum_t w    w = patch_weights(size, 0.5, "exp"), sum_t w = total_weights
                                                        (tiles.py:97-108, 111-169)
      max              keep the class of the more confident tile; a later tile replaces an earlier one unless the
                       earlier one is strictly more confident (compare.py:133-136)

    over the whole tile clipped to the raster (compare.py:97-104), tiles in write order; then `convert` (argmax
    + max probability). The tile squares are the ones slice_extent produces (origin -margin, clamped last
    row / column), not the origin-0 grid test/tiles.py:get_tile_coord assumes -- one more way in which the
    reference's branch is inconsistent with its own tiling. patch_weights is the pinned restatement in
    oracle/tiles_ref.py. Returns (class_map uint8 [H,W], conf float32 [H,W])."""
    from .tiles_ref import patch_weights
    size, margin = config["img_pixels_detection"], config["margin"]
    stride = get_stride(config)[0]
    rows = slice_extent(raster.bounds, (raster.res, raster.res), size, margin, stride)
    nt = config["norma_task"][0]
    H, W, ncls = raster.height, raster.width, config["n_classes"]
    acc = np.zeros((ncls, H, W), np.float32)
    wsum = np.zeros((H, W), np.float32)
    best_conf = np.full((H, W), -1.0, np.float32)
    best_cls = np.zeros((H, W), np.uint8)
    w_tile = {"average": np.ones((size, size), np.float32),
              "average_weights": patch_weights(size, 0.5).astype(np.float32),
              "max": None}[method]
    model.eval()
    for s in range(0, len(rows), batch_size):
        idx = list(range(s, min(s + batch_size, len(rows))))
        imgs = [torch.as_tensor(normalization(read_window_boundless(raster, config["channels"], rows[i]["geometry"], size),
                                              nt["norm_type"], nt["norm_means"], nt["norm_stds"]), dtype=torch.float)
                for i in idx]
        with torch.no_grad():
            predictions = torch.softmax(model(torch.stack(imgs)), dim=1).cpu().numpy()
        for i, p in zip(idx, predictions):
            # tile origin in raster pixels (top-left of the margin-expanded square), then clip to the raster
            gx0, _, _, gy1 = rows[i]["geometry"]
            x0 = int(round((gx0 - raster.min_x) / raster.res))
            y0 = int(round((raster.max_y - gy1) / raster.res))
            ys, ye, xs, xe = max(y0, 0), min(y0 + size, H), max(x0, 0), min(x0 + size, W)
            pt = p[:, ys - y0:ye - y0, xs - x0:xe - x0]
            if method == "max":
                conf, cls = pt.max(axis=0), pt.argmax(axis=0).astype(np.uint8)
                take = ~(best_conf[ys:ye, xs:xe] > conf)
                best_conf[ys:ye, xs:xe] = np.where(take, conf, best_conf[ys:ye, xs:xe])
                best_cls[ys:ye, xs:xe] = np.where(take, cls, best_cls[ys:ye, xs:xe])
            else:
                wt = w_tile[ys - y0:ye - y0, xs - x0:xe - x0]
                acc[:, ys:ye, xs:xe] += pt * wt
                wsum[ys:ye, xs:xe] += wt
    if method == "max":
        return best_cls, best_conf
    norm = acc / np.maximum(wsum, 1e-30)
    return norm.argmax(axis=0).astype(np.uint8), norm.max(axis=0)
