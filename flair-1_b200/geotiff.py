"""Minimal GeoTIFF / BigTIFF reader and writer for uint8 rasters (stands in for rasterio/GDAL).

The reference does all raster I/O through rasterio (src/zone_detect/dataset.py:95-104,
src/zone_detect/main.py:206-232,421-426, src/zone_detect/utils.py:282-297,
src/flair/data_loader.py:122-125, src/flair/writer.py:38-50). GDAL is not available here and nothing
in the image reads 5-band TIFFs or writes tiled BigTIFF, so this module implements the subset the hot
path needs:

  read : classic TIFF and BigTIFF, little/big endian, strips or tiles, chunky or planar, 8-bit samples,
         compression none / LZW (5) / Deflate (8, 32946) / PackBits (32773), horizontal predictor,
         window reads; GeoTIFF tags are returned verbatim so they can be copied to the output.
  write: uint8 [bands, H, W], tiled (default 512, the reference's block size) or striped, chunky,
         LZW (the reference's choice) / Deflate / none, classic or BigTIFF, GeoTIFF tags pass-through.

LZW blocks are coded by the host codec in libflairb200 (csrc/host_codec.cu) on a thread pool.
"""
from __future__ import annotations

import struct
import zlib
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

# tag ids
T_WIDTH, T_LENGTH, T_BITS, T_COMPRESSION, T_PHOTOMETRIC = 256, 257, 258, 259, 262
T_STRIP_OFFSETS, T_SPP, T_ROWS_PER_STRIP, T_STRIP_COUNTS = 273, 277, 278, 279
T_PLANAR, T_PREDICTOR, T_TILE_W, T_TILE_L, T_TILE_OFFSETS, T_TILE_COUNTS = 284, 317, 322, 323, 324, 325
T_EXTRA_SAMPLES, T_SAMPLE_FORMAT = 338, 339
GEO_TAGS = (33550, 33922, 34264, 34735, 34736, 34737, 42112, 42113)  # pixel scale, tiepoint, transform, geokeys x3, GDAL metadata / nodata

_TYPE_FMT = {1: "B", 2: "c", 3: "H", 4: "I", 5: "II", 6: "b", 7: "B", 8: "h", 9: "i", 10: "ii", 11: "f", 12: "d",
             16: "Q", 17: "q", 18: "Q"}
_TYPE_SIZE = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 6: 1, 7: 1, 8: 2, 9: 4, 10: 8, 11: 4, 12: 8, 16: 8, 17: 8, 18: 8}


@dataclass
class TiffInfo:
    width: int
    height: int
    count: int
    compression: int
    planar: int
    predictor: int
    tiled: bool
    block_w: int
    block_h: int
    offsets: Tuple[int, ...]
    byte_counts: Tuple[int, ...]
    bigtiff: bool
    byteorder: str
    geo_tags: Dict[int, Tuple[int, tuple]] = field(default_factory=dict)  # tag -> (tiff type, values)

    @property
    def transform(self) -> Optional[Tuple[float, float, float, float, float, float]]:
        """Affine (a, b, c, d, e, f) with x = a*col + b*row + c, y = d*col + e*row + f (rasterio order)."""
        if 34264 in self.geo_tags:
            m = self.geo_tags[34264][1]
            return (m[0], m[1], m[3], m[4], m[5], m[7])
        if 33550 in self.geo_tags and 33922 in self.geo_tags:
            sx, sy = self.geo_tags[33550][1][:2]
            tp = self.geo_tags[33922][1]
            i, j, x, y = tp[0], tp[1], tp[3], tp[4]
            return (sx, 0.0, x - i * sx, 0.0, -sy, y + j * sy)
        return None

    @property
    def bounds(self) -> Tuple[float, float, float, float]:
        """(left, bottom, right, top) like rasterio's src.bounds; pixel units when not georeferenced."""
        t = self.transform or (1.0, 0.0, 0.0, 0.0, -1.0, float(self.height))
        left, top = t[2], t[5]
        return (left, top + t[4] * self.height, left + t[0] * self.width, top)

    @property
    def res(self) -> Tuple[float, float]:
        t = self.transform or (1.0, 0.0, 0.0, 0.0, -1.0, 0.0)
        return (abs(t[0]), abs(t[4]))


def _read_ifd(f, bo: str, big: bool, pos: int) -> Dict[int, Tuple[int, tuple]]:
    f.seek(pos)
    if big:
        (n,) = struct.unpack(bo + "Q", f.read(8))
        esz, cfmt, inline = 20, "Q", 8
    else:
        (n,) = struct.unpack(bo + "H", f.read(2))
        esz, cfmt, inline = 12, "I", 4
    raw = f.read(n * esz)
    tags = {}
    for i in range(n):
        e = raw[i * esz:(i + 1) * esz]
        tag, typ = struct.unpack(bo + "HH", e[:4])
        (cnt,) = struct.unpack(bo + cfmt, e[4:4 + inline])
        if typ not in _TYPE_SIZE:
            continue
        nbytes = cnt * _TYPE_SIZE[typ]
        if nbytes <= inline:
            data = e[4 + inline:4 + inline + nbytes]
        else:
            (off,) = struct.unpack(bo + cfmt, e[4 + inline:4 + 2 * inline])
            here = f.tell()
            f.seek(off)
            data = f.read(nbytes)
            f.seek(here)
        if typ == 2:
            vals = (data.rstrip(b"\0").decode("latin-1"),)
        elif typ in (5, 10):
            flat = struct.unpack(bo + _TYPE_FMT[typ][0] * (2 * cnt), data)
            vals = tuple(flat[2 * k] / flat[2 * k + 1] if flat[2 * k + 1] else 0.0 for k in range(cnt))
        else:
            vals = struct.unpack(bo + _TYPE_FMT[typ] * cnt, data)
        tags[tag] = (typ, vals)
    return tags


def read_info(path) -> TiffInfo:
    with open(path, "rb") as f:
        hdr = f.read(16)
        bo = {b"II": "<", b"MM": ">"}.get(hdr[:2])
        if bo is None:
            raise ValueError(f"{path}: not a TIFF file")
        (magic,) = struct.unpack(bo + "H", hdr[2:4])
        if magic == 42:
            big = False
            (ifd,) = struct.unpack(bo + "I", hdr[4:8])
        elif magic == 43:
            big = True
            (ifd,) = struct.unpack(bo + "Q", hdr[8:16])
        else:
            raise ValueError(f"{path}: bad TIFF magic {magic}")
        tags = _read_ifd(f, bo, big, ifd)

    def one(t, default=None):
        return tags[t][1][0] if t in tags else default

    bits = tags.get(T_BITS, (3, (1,)))[1]
    if any(b != 8 for b in bits) or one(T_SAMPLE_FORMAT, 1) != 1:
        raise ValueError(f"{path}: only 8-bit unsigned samples are supported (got bits={bits})")
    width, height, count = one(T_WIDTH), one(T_LENGTH), one(T_SPP, 1)
    tiled = T_TILE_OFFSETS in tags
    if tiled:
        bw, bh = one(T_TILE_W), one(T_TILE_L)
        offs, cnts = tags[T_TILE_OFFSETS][1], tags[T_TILE_COUNTS][1]
    else:
        bw, bh = width, min(one(T_ROWS_PER_STRIP, height), height)
        offs, cnts = tags[T_STRIP_OFFSETS][1], tags[T_STRIP_COUNTS][1]
    return TiffInfo(width=width, height=height, count=count, compression=one(T_COMPRESSION, 1), planar=one(T_PLANAR, 1),
                    predictor=one(T_PREDICTOR, 1), tiled=tiled, block_w=bw, block_h=bh, offsets=tuple(offs),
                    byte_counts=tuple(cnts), bigtiff=big, byteorder=bo,
                    geo_tags={t: tags[t] for t in GEO_TAGS if t in tags})


def _default_threads() -> int:
    """Block codec threads: the host cores this process may use (torchrun pins OMP threads, not these), at most 32."""
    import os
    try:
        n = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        n = os.cpu_count() or 1
    return max(1, min(32, n))


def _packbits_decode(data: bytes, expected: int) -> np.ndarray:
    out = bytearray()
    i, n = 0, len(data)
    while i < n and len(out) < expected:
        h = data[i]
        i += 1
        if h < 128:
            out += data[i:i + h + 1]
            i += h + 1
        elif h > 128:
            out += data[i:i + 1] * (257 - h)
            i += 1
    out = out[:expected] + bytes(max(0, expected - len(out)))
    return np.frombuffer(bytes(out), np.uint8)


def _decode_block(raw: bytes, compression: int, expected: int) -> np.ndarray:
    if compression == 1:
        a = np.frombuffer(raw, np.uint8)
    elif compression in (8, 32946):
        a = np.frombuffer(zlib.decompress(raw), np.uint8)
    elif compression == 5:
        from . import _native
        a = _native.lzw_decode(raw, expected)
    elif compression == 32773:
        a = _packbits_decode(raw, expected)
    else:
        raise ValueError(f"unsupported TIFF compression {compression}")
    if a.size < expected:
        a = np.concatenate([a, np.zeros(expected - a.size, np.uint8)])
    return a[:expected]


def read(path, bands: Optional[Sequence[int]] = None, window: Optional[Tuple[int, int, int, int]] = None,
         out: Optional[np.ndarray] = None) -> np.ndarray:
    """uint8 [len(bands), h, w]. bands are 1-based like rasterio's `indexes` (default: all);
    window = (col_off, row_off, width, height) must lie inside the raster."""
    info = read_info(path)
    bands = list(range(1, info.count + 1)) if bands is None else list(bands)
    if any(b < 1 or b > info.count for b in bands):
        raise IndexError(f"band index out of range (raster has {info.count} bands)")
    c0, r0, w, h = window if window is not None else (0, 0, info.width, info.height)
    if c0 < 0 or r0 < 0 or c0 + w > info.width or r0 + h > info.height:
        raise ValueError("window outside the raster")
    if out is None:
        out = np.empty((len(bands), h, w), np.uint8)
    bw, bh = info.block_w, info.block_h
    nbx, nby = (info.width + bw - 1) // bw, (info.height + bh - 1) // bh
    planes = info.count if info.planar == 2 else 1
    spp = 1 if info.planar == 2 else info.count
    jobs = []
    for by in range(r0 // bh, (r0 + h - 1) // bh + 1):
        for bx in range(c0 // bw, (c0 + w - 1) // bw + 1):
            for pl in (range(planes) if info.planar == 2 else [0]):
                if info.planar == 2 and (pl + 1) not in bands:
                    continue
                jobs.append((by, bx, pl))
    with open(path, "rb") as f:
        raws = {}
        for job in jobs:
            by, bx, pl = job
            idx = (pl * nby + by) * nbx + bx
            f.seek(info.offsets[idx])
            raws[job] = f.read(info.byte_counts[idx])

    def decode(job):
        by, bx, pl = job
        rows = bh if info.tiled else min(bh, info.height - by * bh)
        blk = _decode_block(raws[job], info.compression, rows * bw * spp).reshape(rows, bw, spp)
        if info.predictor == 2:
            blk = np.cumsum(blk, axis=1, dtype=np.uint8)
        return job, blk

    with ThreadPoolExecutor(max_workers=_default_threads()) as ex:   # the codecs release the GIL
        for (by, bx, pl), blk in ex.map(decode, jobs):
            y0, x0 = by * bh, bx * bw
            ys, ye = max(y0, r0), min(y0 + blk.shape[0], r0 + h)
            xs, xe = max(x0, c0), min(x0 + bw, c0 + w, info.width)
            if ye <= ys or xe <= xs:
                continue
            sub = blk[ys - y0:ye - y0, xs - x0:xe - x0]
            if info.planar == 2:
                for k, b in enumerate(bands):
                    if b - 1 == pl:
                        out[k, ys - r0:ye - r0, xs - c0:xe - c0] = sub[:, :, 0]
            else:
                for k, b in enumerate(bands):
                    out[k, ys - r0:ye - r0, xs - c0:xe - c0] = sub[:, :, b - 1]
    return out


# ------------------------------------------------------------------------------------------ writer
def _encode_block(blk: np.ndarray, compress: str) -> bytes:
    if compress == "none":
        return blk.tobytes()
    if compress == "deflate":
        return zlib.compress(blk.tobytes(), 6)
    if compress == "lzw":
        from . import _native
        return _native.lzw_encode(np.ascontiguousarray(blk))
    raise ValueError(f"unknown compression {compress!r}")


def write(path, data: np.ndarray, geo_tags: Optional[Dict[int, Tuple[int, tuple]]] = None, compress: str = "lzw",
          tiled: bool = True, blocksize: int = 512, bigtiff: Optional[bool] = None, threads: int = 0) -> None:
    """Write uint8 [bands, H, W] (or [H, W]) as a chunky (pixel-interleaved) TIFF.

    compress in {"lzw", "deflate", "none"}; tiled blocks are blocksize x blocksize (the reference uses
    img_pixels_detection, main.py:224-226), striped output uses blocksize rows per strip."""
    if data.ndim == 2:
        data = data[None]
    if data.dtype != np.uint8 or data.ndim != 3:
        raise ValueError("write() takes uint8 [bands, H, W]")
    compress = compress.lower()
    bands, H, W = data.shape
    if bigtiff is None:
        bigtiff = data.nbytes > 3_500_000_000
    bw, bh = (blocksize, blocksize) if tiled else (W, min(blocksize, H))
    if tiled and (bw % 16 or bh % 16):
        raise ValueError("tile size must be a multiple of 16")
    nbx, nby = (W + bw - 1) // bw, (H + bh - 1) // bh

    def make(job):
        by, bx = job
        y0, x0 = by * bh, bx * bw
        rows = bh if tiled else min(bh, H - y0)
        blk = np.zeros((rows, bw, bands), np.uint8)
        ye, xe = min(y0 + rows, H), min(x0 + bw, W)
        blk[:ye - y0, :xe - x0] = np.moveaxis(data[:, y0:ye, x0:xe], 0, 2)
        return _encode_block(blk, compress)

    jobs = [(by, bx) for by in range(nby) for bx in range(nbx)]
    with ThreadPoolExecutor(max_workers=threads if threads > 0 else _default_threads()) as ex:
        blocks = list(ex.map(make, jobs))

    comp_code = {"none": 1, "lzw": 5, "deflate": 8}[compress]
    entries = [(T_WIDTH, 4, (W,)), (T_LENGTH, 4, (H,)), (T_BITS, 3, (8,) * bands), (T_COMPRESSION, 3, (comp_code,)),
               (T_PHOTOMETRIC, 3, (2 if bands == 3 else 1,)), (T_SPP, 3, (bands,)), (T_PLANAR, 3, (1,)),
               (T_SAMPLE_FORMAT, 3, (1,) * bands)]
    if bands not in (1, 3):
        entries.append((T_EXTRA_SAMPLES, 3, (0,) * (bands - 1)))
    otyp = 16 if bigtiff else 4
    if tiled:
        entries += [(T_TILE_W, 3, (bw,)), (T_TILE_L, 3, (bh,)), (T_TILE_OFFSETS, otyp, None), (T_TILE_COUNTS, otyp, None)]
    else:
        entries += [(T_ROWS_PER_STRIP, 4, (bh,)), (T_STRIP_OFFSETS, otyp, None), (T_STRIP_COUNTS, otyp, None)]
    for t, (typ, vals) in (geo_tags or {}).items():
        entries.append((t, typ, vals))
    entries.sort(key=lambda e: e[0])

    hdr_size = 16 if bigtiff else 8
    offsets, pos = [], hdr_size
    for b in blocks:
        offsets.append(pos)
        pos += len(b) + (len(b) & 1)
    counts = [len(b) for b in blocks]
    ifd_pos = pos
    bo = "<"
    inline, cfmt, esz = (8, "Q", 20) if bigtiff else (4, "I", 12)
    ifd_head = struct.pack(bo + ("Q" if bigtiff else "H"), len(entries))
    extra_pos = ifd_pos + len(ifd_head) + len(entries) * esz + inline
    body, extra = b"", b""
    for tag, typ, vals in entries:
        if vals is None:
            vals = tuple(offsets) if tag in (T_TILE_OFFSETS, T_STRIP_OFFSETS) else tuple(counts)
        if typ == 2:
            payload = vals[0].encode("latin-1") + b"\0"
            cnt = len(payload)
        elif typ in (5, 10):
            flat = []
            for v in vals:
                flat += [int(round(v * 10000)), 10000]
            payload = struct.pack(bo + _TYPE_FMT[typ][0] * len(flat), *flat)
            cnt = len(vals)
        else:
            payload = struct.pack(bo + _TYPE_FMT[typ] * len(vals), *vals)
            cnt = len(vals)
        e = struct.pack(bo + "HH" + cfmt, tag, typ, cnt)
        if len(payload) <= inline:
            e += payload + b"\0" * (inline - len(payload))
        else:
            e += struct.pack(bo + cfmt, extra_pos + len(extra))
            extra += payload + (b"\0" if len(payload) & 1 else b"")
        body += e
    if not bigtiff and extra_pos + len(extra) >= 2 ** 32:
        raise ValueError("file too large for classic TIFF; pass bigtiff=True")
    with open(path, "wb") as f:
        if bigtiff:
            f.write(b"II" + struct.pack("<HHHQ", 43, 8, 0, ifd_pos))
        else:
            f.write(b"II" + struct.pack("<HI", 42, ifd_pos))
        for b in blocks:
            f.write(b)
            if len(b) & 1:
                f.write(b"\0")
        f.write(ifd_head + body + struct.pack(bo + cfmt, 0) + extra)


def georef_tags(min_x: float, max_y: float, res_x: float, res_y: float, epsg: int = 2154) -> Dict[int, Tuple[int, tuple]]:
    """GeoTIFF tags of a north-up raster in a projected CRS (default Lambert-93, the FLAIR CRS)."""
    return {33550: (12, (float(res_x), float(res_y), 0.0)),
            33922: (12, (0.0, 0.0, 0.0, float(min_x), float(max_y), 0.0)),
            34735: (3, (1, 1, 0, 3, 1024, 0, 1, 1, 1025, 0, 1, 1, 3072, 0, 1, int(epsg)))}
