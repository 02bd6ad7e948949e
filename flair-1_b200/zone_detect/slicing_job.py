"""Tile grid of the sliding-window job, in integer pixel space.

Replaces `slice_extent` (src/zone_detect/slicing_job.py:19-118). The reference works in geographic
floats and builds a GeoDataFrame of shapely boxes; for a north-up, pixel-aligned raster the same grid
is exact integer arithmetic, which is what the GPU tile table needs:

  x_k  = -margin + k*stride  for k = 0.. while x_k < W + margin      (np.arange at :51)
  x_k := W + margin - size   when x_k + size > W + margin            (:57-58, last column clamp)
  y is measured from the raster's BOTTOM edge the same way           (:52, :62-63)
  interior = [x+m, min(x+size-m, W)) x [yb+m, min(yb+size-m, H))     (:67-70)
  rows are appended x-major / y-minor, duplicates (same interior) dropped (:78-86)

The write order of the reference's hot loop is the row order and later writes overwrite earlier ones
(main.py:409-426), so every pixel belongs to the LAST tile whose interior covers it. The ownership is
separable per axis and is resolved here once, giving each tile a half-open write rectangle.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np

TILE_FIELDS = ("x0", "y0", "wx0", "wy0", "wx1", "wy1")


def _axis_origins(extent: int, size: int, margin: int, stride: int) -> List[int]:
    """Unique tile origins along one axis in first-appearance order (slicing_job.py:51-63, 78-86)."""
    out: List[int] = []
    k = -margin
    while k < extent + margin:
        o = extent + margin - size if k + size > extent + margin else k
        if o not in out:
            out.append(o)
        k += stride
    return out


def _axis_ownership(origins: List[int], extent: int, size: int, margin: int) -> List[Tuple[int, int]]:
    """For each origin (in write order) the half-open interval of [0, extent) it owns: the interior
    [o+m, min(o+size-m, extent)) minus everything a later origin's interior covers."""
    owner = np.full(extent, -1, dtype=np.int64)
    for i, o in enumerate(origins):
        a, b = max(o + margin, 0), min(o + size - margin, extent)
        if b > a:
            owner[a:b] = i
    res = []
    for i in range(len(origins)):
        idx = np.flatnonzero(owner == i)
        if idx.size == 0:
            res.append((0, 0))
            continue
        if idx[-1] - idx[0] + 1 != idx.size:
            raise ValueError("non-contiguous tile ownership (tile grid with stride > interior is not supported)")
        res.append((int(idx[0]), int(idx[-1]) + 1))
    return res


def tile_table(width: int, height: int, size: int, margin: int, stride: int = 0) -> np.ndarray:
    """int32 [n, 6] rows (x0, y0, wx0, wy0, wx1, wy1) in the reference's write order; pixel
    coordinates with y growing downwards; (x0, y0) = top-left of the margin-expanded tile."""
    if stride <= 0:
        stride = size - 2 * margin
    if stride <= 0:
        raise ValueError("stride must be positive (2*margin < img_pixels_detection)")
    xs = _axis_origins(width, size, margin, stride)
    ybs = _axis_origins(height, size, margin, stride)       # measured from the bottom edge
    own_x = _axis_ownership(xs, width, size, margin)
    own_yb = _axis_ownership(ybs, height, size, margin)
    rows = []
    for ix, x in enumerate(xs):
        for iy, yb in enumerate(ybs):
            wx0, wx1 = own_x[ix]
            b0, b1 = own_yb[iy]                              # from-bottom interval
            rows.append((x, height - (yb + size), wx0, height - b1, wx1, height - b0))
    return np.asarray(rows, dtype=np.int32).reshape(-1, 6)


def tile_interiors(width: int, height: int, size: int, margin: int, stride: int = 0) -> np.ndarray:
    """int64 [n, 4] (left, bottom, right, top) interior boxes in pixel units from the raster's
    bottom-left corner, same order as tile_table -- the `left/bottom/right/top` columns of the
    reference dataframe divided by the resolution."""
    if stride <= 0:
        stride = size - 2 * margin
    xs = _axis_origins(width, size, margin, stride)
    ybs = _axis_origins(height, size, margin, stride)
    return np.asarray([(x + margin, yb + margin, min(x + size - margin, width), min(yb + size - margin, height))
                       for x in xs for yb in ybs], dtype=np.int64).reshape(-1, 4)


def tile_windows(width: int, height: int, size: int, margin: int, stride: int = 0) -> np.ndarray:
    """int32 [n, 6] rows (x0, y0, cx0, cy0, cx1, cy1), same order as tile_table: the margin-cropped window each
    tile's prediction is written to and scored on (compare.py:66-82 -> `window`; test/metrics.py:146-149), in
    pixel coordinates with y growing downwards. Unlike the write rectangles these overlap where a clamped last
    row / column re-covers pixels."""
    if stride <= 0:
        stride = size - 2 * margin
    t = tile_table(width, height, size, margin, stride)
    ints = tile_interiors(width, height, size, margin, stride)
    out = t.copy()
    out[:, 2] = ints[:, 0]
    out[:, 3] = height - ints[:, 3]
    out[:, 4] = ints[:, 2]
    out[:, 5] = height - ints[:, 1]
    return out


def slice_extent(in_img, patch_size: int, margin: int, output_path, output_name: str, write_dataframe: bool, stride: int):
    """Same call and return shape as the reference's slice_extent (slicing_job.py:19-118):
    (tile table, profile, (res_x, res_y), [n_rows, n_cols]). The table is the int32 pixel-space array of
    tile_table() instead of a GeoDataFrame; the profile carries what the output raster needs (size and
    the GeoTIFF tags to copy). The reference reads the whole first band just to learn the shape (:30) and
    names (rows, cols) "width, height"; the header is enough and the order is kept."""
    from .. import geotiff
    info = geotiff.read_info(in_img)
    res = info.res
    resolution = (abs(round(res[0], 5)), abs(round(res[1], 5)))
    tiles = tile_table(info.width, info.height, patch_size, margin, stride)
    profile = {"width": info.width, "height": info.height, "count": info.count, "dtype": "uint8",
               "geo_tags": dict(info.geo_tags), "transform": info.transform}
    if write_dataframe:
        import os
        path = os.path.join(str(output_path), output_name.split(".tif")[0] + "_slicing_job.csv")
        ints = tile_interiors(info.width, info.height, patch_size, margin, stride)
        t = info.transform or (1.0, 0.0, 0.0, 0.0, -1.0, float(info.height))
        with open(path, "w") as f:  # the reference writes a GeoPackage through geopandas; CSV carries the same columns
            f.write("id,left,bottom,right,top,x0_px,y0_px\n")
            for i, ((l, b, r, tp), row) in enumerate(zip(ints, tiles)):
                f.write(f"{i},{t[2] + l * t[0]},{t[5] + (info.height - b) * t[4]},{t[2] + r * t[0]},"
                        f"{t[5] + (info.height - tp) * t[4]},{row[0]},{row[1]}\n")
    return tiles, profile, resolution, [info.height, info.width]


def split_rows_across_ranks(tiles: np.ndarray, world_size: int) -> List[np.ndarray]:
    """Shard the tile table by tile *rows* (equal y0) into `world_size` contiguous groups balanced by
    tile count; each rank then needs raster rows [min y0, max y0 + size) only. Returns index arrays
    into `tiles` that keep the original (write) order inside each shard."""
    y_vals = np.unique(tiles[:, 1])
    groups = np.array_split(y_vals, world_size)
    return [np.flatnonzero(np.isin(tiles[:, 1], g)) for g in groups]


def split_tiles_across_ranks(tiles: np.ndarray, world_size: int) -> List[np.ndarray]:
    """Shard the tile table into `world_size` contiguous ranges of the ROW-MAJOR tile order (tiles sorted by y0,
    ties in write order, i.e. ascending x), balanced to within one tile (24 649 tiles over 8 ranks: 3082 against
    3081.1 on average, where whole tile rows give 20 rows against 19.6). A rank's first and last tile row may be
    shared with its neighbours; it still needs raster rows [min y0, max y0 + size) only, and what it writes is a
    set of row bands (owned_rects) that is disjoint from every other rank's. Returns index arrays into `tiles`."""
    order = np.argsort(tiles[:, 1], kind="stable")
    return [np.asarray(part, dtype=np.int64) for part in np.array_split(order, world_size)]


def owned_rects(tiles: np.ndarray) -> np.ndarray:
    """int64 [k, 4] rows (y0, y1, x0, x1): the write rectangles of `tiles` merged into maximal runs that share
    their rows and own adjacent columns, in y-sorted order -- the pieces of the class map one shard produces
    (fb_detect_zone_shard sends them back, and scores them, in exactly these units)."""
    t = np.asarray(tiles, dtype=np.int64).reshape(-1, 6)
    t = t[np.argsort(t[:, 1], kind="stable")]
    rects: List[List[int]] = []
    for x0, y0, wx0, wy0, wx1, wy1 in t:
        if wx1 <= wx0 or wy1 <= wy0:
            continue
        if rects and rects[-1][0] == wy0 and rects[-1][1] == wy1 and rects[-1][3] == wx0:
            rects[-1][3] = int(wx1)
        else:
            rects.append([int(wy0), int(wy1), int(wx0), int(wx1)])
    return np.asarray(rects, dtype=np.int64).reshape(-1, 4)
