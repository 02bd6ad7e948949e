"""ctypes binding of libflairb200.so (include/flair_b200.h).

PyTorch is used here only for device memory and streams: tensors are allocated by torch and handed to
the C ABI as raw device pointers. There is no fallback: if the library is missing or no sm_100 device
is present, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Mapping, Optional, Sequence

import numpy as np
import torch

from . import build as _build

FB_NORM = {"custom": 0, "scaling": 1, "without": 2}
FB_LAYOUT_CHW, FB_LAYOUT_HWC = 0, 1
LOGIT_STRIDE = 16
METADATA_DIM = 45


class fb_tensor_desc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("ndim", C.c_int32), ("shape", C.c_int64 * 4)]


class fb_tile(C.Structure):
    _fields_ = [("x0", C.c_int32), ("y0", C.c_int32), ("wx0", C.c_int32), ("wy0", C.c_int32),
                ("wx1", C.c_int32), ("wy1", C.c_int32)]


TILE_DTYPE = np.dtype([("x0", "<i4"), ("y0", "<i4"), ("wx0", "<i4"), ("wy0", "<i4"), ("wx1", "<i4"), ("wy1", "<i4")])


class NativeError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libflairb200 error {code}: {message}")
        self.code = code


_lib: Optional[C.CDLL] = None

_SIGNATURES = {
    "fb_api_version": (C.c_int, []),
    "fb_create": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "fb_destroy": (None, [C.c_void_p]),
    "fb_last_error": (C.c_char_p, [C.c_void_p]),
    "fb_synchronize": (C.c_int, [C.c_void_p]),
    "fb_load_weights": (C.c_int, [C.c_void_p, C.POINTER(fb_tensor_desc), C.c_int, C.c_int, C.c_int, C.c_int]),
    "fb_set_norm": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]),
    "fb_set_raster": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.c_int, C.c_int64,
                                C.c_int64, C.c_int64, C.c_int64, C.c_int]),
    "fb_upload_raster": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.c_int, C.c_int64,
                                   C.c_int64, C.c_int64, C.c_int64, C.c_int]),
    "fb_forward_tiles": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "fb_detect_strip": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_int64, C.c_int64]),
    "fb_detect_strip_metrics": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                          C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int, C.c_void_p]),
    "fb_detect_strip_prob": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int64,
                                       C.c_int64]),
    "fb_blend_strip": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                 C.c_int64, C.c_int64, C.c_int64, C.c_int]),
    "fb_blend_finalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    "fb_detect_zone_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.c_int, C.c_int64,
                                      C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                      C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64]),
    "fb_detect_zone_shard": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.c_int, C.c_int64,
                                       C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                       C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                       C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "fb_predict_patches": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "fb_confusion": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p]),
    "fb_confusion_rect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int,
                                    C.c_int, C.c_void_p]),
    "fb_host_register": (C.c_int, [C.c_void_p, C.c_int64]),
    "fb_host_unregister": (C.c_int, [C.c_void_p]),
    "fb_comm_unique_id": (C.c_int, [C.c_void_p]),
    "fb_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "fb_comm_destroy": (C.c_int, [C.c_void_p]),
    "fb_allreduce_confusion": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "fb_gather_bytes": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                  C.c_int]),
    "fb_conv2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                            C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    "fb_conv2d_halo": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "fb_debug_activation": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "fb_profile_forward": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]),
    "fb_profile_begin": (C.c_int, [C.c_void_p]),
    "fb_profile_end": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "fb_launch_count": (C.c_int64, [C.c_void_p]),
    "fb_flop_count": (C.c_double, [C.c_void_p]),
    "fb_logit_stride": (C.c_int, [C.c_void_p]),
    "fb_debug_need_rect": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32)]),
    "fb_debug_tile_cover": (C.c_int, [C.c_int] * 9 + [C.POINTER(C.c_int32)]),
    "fb_lzw_bound": (C.c_int64, [C.c_int64]),
    "fb_lzw_encode": (C.c_int64, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]),
    "fb_lzw_decode": (C.c_int64, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def library_path() -> Path:
    return _build.LIB_PATH


def load_library(build_if_missing: bool = True) -> C.CDLL:
    """dlopen libflairb200.so (building it with nvcc first if it is absent or stale)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if build_if_missing and _build.needs_build():
        _build.build_library()
    if not path.exists():
        raise RuntimeError(f"{path} is missing: run `python __graft_entry__.py build` (no CPU fallback exists)")
    lib = C.CDLL(str(path))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library out of sync
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _np_ptr(a: Optional[np.ndarray]) -> Optional[int]:
    return None if a is None else a.ctypes.data


def make_tiles(rows: Sequence[Sequence[int]]) -> np.ndarray:
    """[(x0, y0, wx0, wy0, wx1, wy1), ...] -> packed fb_tile array."""
    arr = np.asarray(rows, dtype=np.int32).reshape(-1, 6)
    return np.ascontiguousarray(arr)


class Context:
    """One fb_ctx: a device, a stream, a loaded model, a normalisation table and a resident raster."""

    def __init__(self, device: int | torch.device = 0, stream: Optional[torch.cuda.Stream] = None):
        self._h = C.c_void_p()
        self._lib = load_library()
        if not torch.cuda.is_available():
            raise RuntimeError("flair1_b200 needs a CUDA device (sm_100a); there is no CPU path")
        dev = torch.device(device) if not isinstance(device, int) else torch.device("cuda", device)
        self.device = dev
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        torch.cuda.set_device(idx)
        self.stream = stream if stream is not None else torch.cuda.current_stream(idx)
        rc = self._lib.fb_create(idx, C.c_void_p(self.stream.cuda_stream), C.byref(self._h))
        if rc != 0:
            msg = self._lib.fb_last_error(None)
            raise NativeError(rc, msg.decode() if msg else "fb_create failed")
        self.in_channels = 0
        self.n_classes = 0
        self.use_metadata = False
        self._keepalive = {}

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc: int) -> None:
        if rc != 0:
            msg = self._lib.fb_last_error(self._h)
            raise NativeError(rc, msg.decode() if msg else "unknown")

    def close(self) -> None:
        if getattr(self, "_h", None) and self._h.value:
            self._lib.fb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self) -> None:
        self._check(self._lib.fb_synchronize(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.fb_launch_count(self._h))

    @property
    def logit_stride(self) -> int:
        """Floats per pixel of logits / blend accumulators for the loaded model (16, or 32 above 16 classes)."""
        return int(self._lib.fb_logit_stride(self._h))

    @property
    def flop_count(self) -> float:
        """Algorithmic conv FLOPs of the outputs computed so far (fb_flop_count)."""
        return float(self._lib.fb_flop_count(self._h))

    # ------------------------------------------------------------------ model
    def load_weights(self, state_dict: Mapping[str, torch.Tensor], in_channels: int, n_classes: int,
                     use_metadata: bool = False) -> None:
        keep = []
        descs = []
        for k, v in state_dict.items():
            if not torch.is_tensor(v) or not v.is_floating_point():
                continue  # num_batches_tracked etc.
            a = np.ascontiguousarray(v.detach().to("cpu", torch.float32).numpy())
            if a.ndim > 4:
                continue
            keep.append(a)
            d = fb_tensor_desc()
            d.name = k.encode()
            d.data = a.ctypes.data
            d.ndim = a.ndim
            for i, s in enumerate(a.shape):
                d.shape[i] = s
            descs.append(d)
        arr = (fb_tensor_desc * len(descs))(*descs)
        self._check(self._lib.fb_load_weights(self._h, arr, len(descs), in_channels, n_classes, int(use_metadata)))
        self.in_channels, self.n_classes, self.use_metadata = in_channels, n_classes, bool(use_metadata)

    def set_norm(self, norm_type: str, means: Sequence[float] = (), stds: Sequence[float] = (), channels: Optional[int] = None) -> None:
        c = channels if channels is not None else (len(means) if norm_type == "custom" else self.in_channels)
        m = (C.c_double * 8)(*([float(x) for x in means][:8] + [0.0] * (8 - min(8, len(means)))))
        s = (C.c_double * 8)(*([float(x) for x in stds][:8] + [1.0] * (8 - min(8, len(stds)))))
        self._check(self._lib.fb_set_norm(self._h, FB_NORM[norm_type], m, s, c))

    # ------------------------------------------------------------------ raster
    def set_raster(self, raster: torch.Tensor, band_idx: Sequence[int], W: int, H: int, row0: int = 0,
                   layout: int = FB_LAYOUT_CHW) -> None:
        """raster: uint8 device tensor [bands, rows, W] (CHW) or [rows, W, bands] (HWC)."""
        assert raster.dtype == torch.uint8 and raster.is_cuda and raster.is_contiguous()
        bands_total = raster.shape[0] if layout == FB_LAYOUT_CHW else raster.shape[2]
        rows = raster.shape[1] if layout == FB_LAYOUT_CHW else raster.shape[0]
        bi = (C.c_int32 * len(band_idx))(*band_idx)
        self._keepalive["raster"] = raster
        self._check(self._lib.fb_set_raster(self._h, raster.data_ptr(), bands_total, bi, len(band_idx), W, H, row0, rows, layout))

    def upload_raster(self, raster: np.ndarray | torch.Tensor, band_idx: Sequence[int], W: int, H: int,
                      row0: int = 0, layout: int = FB_LAYOUT_CHW) -> None:
        """raster: uint8 host array/tensor (pinned or pageable), copied into a context-owned buffer."""
        if torch.is_tensor(raster):
            assert not raster.is_cuda and raster.dtype == torch.uint8 and raster.is_contiguous()
            ptr, shape = raster.data_ptr(), tuple(raster.shape)
        else:
            raster = np.ascontiguousarray(raster, dtype=np.uint8)
            ptr, shape = raster.ctypes.data, raster.shape
        bands_total = shape[0] if layout == FB_LAYOUT_CHW else shape[2]
        rows = shape[1] if layout == FB_LAYOUT_CHW else shape[0]
        bi = (C.c_int32 * len(band_idx))(*band_idx)
        self._keepalive["raster_host"] = raster
        self._check(self._lib.fb_upload_raster(self._h, ptr, bands_total, bi, len(band_idx), W, H, row0, rows, layout))

    # ------------------------------------------------------------------ compute
    def forward_tiles(self, tile_xy: np.ndarray, tile: int, metadata: Optional[np.ndarray] = None) -> torch.Tensor:
        """logits [n, tile, tile, logit_stride] fp32 (device) for tiles at (x0, y0) of the current raster."""
        xy = np.ascontiguousarray(tile_xy, dtype=np.int32).reshape(-1, 2)
        n = xy.shape[0]
        md = None if metadata is None else np.ascontiguousarray(metadata, dtype=np.float32).reshape(n, METADATA_DIM)
        out = torch.empty((n, tile, tile, self.logit_stride), dtype=torch.float32, device=self.device)
        self._check(self._lib.fb_forward_tiles(self._h, xy.ctypes.data, n, tile, _np_ptr(md), out.data_ptr()))
        return out

    def detect_strip(self, tiles: np.ndarray, tile: int, batch: int, cls_map: torch.Tensor,
                     conf_map: Optional[torch.Tensor], map_w: int, map_row0: int = 0) -> None:
        t = make_tiles(tiles)
        assert cls_map.dtype == torch.uint8 and cls_map.is_cuda
        self._check(self._lib.fb_detect_strip(self._h, t.ctypes.data, t.shape[0], tile, batch, cls_map.data_ptr(),
                                              _ptr(conf_map), map_w, map_row0))

    def detect_strip_metrics(self, tiles: np.ndarray, windows: np.ndarray, tile: int, batch: int, cls_map: torch.Tensor,
                             conf_map: Optional[torch.Tensor], map_w: int, map_row0: int, truth: torch.Tensor,
                             truth_sub: int = 0) -> torch.Tensor:
        """detect_strip plus the per-patch confusion matrices of the compare loop (fb_detect_strip_metrics):
        returns int64 [n, n_classes, n_classes] (device), tile i's own prediction over windows[i] against `truth`
        (uint8, same geometry as cls_map)."""
        t, w = make_tiles(tiles), make_tiles(windows)
        assert t.shape == w.shape
        assert cls_map.dtype == torch.uint8 and cls_map.is_cuda and truth.dtype == torch.uint8 and truth.is_cuda
        assert truth.is_contiguous() and truth.shape == cls_map.shape
        cm = torch.zeros((t.shape[0], self.n_classes, self.n_classes), dtype=torch.int64, device=self.device)
        self._check(self._lib.fb_detect_strip_metrics(self._h, t.ctypes.data, w.ctypes.data, t.shape[0], tile, batch,
                                                      cls_map.data_ptr(), _ptr(conf_map), map_w, map_row0,
                                                      truth.data_ptr(), truth_sub, cm.data_ptr()))
        return cm

    def detect_strip_prob(self, tiles: np.ndarray, tile: int, batch: int, prob_map: torch.Tensor, map_w: int,
                          map_row0: int = 0) -> None:
        """output_type class_prob: prob_map uint8 [n_classes, map_rows, map_w] (device)."""
        t = make_tiles(tiles)
        assert prob_map.dtype == torch.uint8 and prob_map.is_cuda and prob_map.dim() == 3 and prob_map.is_contiguous()
        self._check(self._lib.fb_detect_strip_prob(self._h, t.ctypes.data, t.shape[0], tile, batch, prob_map.data_ptr(),
                                                   map_w, map_row0, prob_map.shape[1]))

    STITCH_METHODS = {"average": 0, "average_weights": 1, "max": 2}

    def blend_buffers(self, method: str, map_rows: int, map_w: int):
        """Zeroed accumulators for blend_strip: (acc, wsum); wsum is None for 'max'."""
        if self.STITCH_METHODS[method] == 2:
            return torch.zeros((map_rows, map_w), dtype=torch.int64, device=self.device), None
        return (torch.zeros((map_rows, map_w, self.logit_stride), dtype=torch.float32, device=self.device),
                torch.zeros((map_rows, map_w), dtype=torch.float32, device=self.device))

    def blend_strip(self, tiles: np.ndarray, tile: int, batch: int, method: str, acc: torch.Tensor,
                    wsum: Optional[torch.Tensor], map_w: int, map_row0: int = 0, tile_seq0: int = 0) -> None:
        """Accumulate the whole tiles (clipped to the raster) into the blend accumulators."""
        t = make_tiles(tiles)
        self._check(self._lib.fb_blend_strip(self._h, t.ctypes.data, t.shape[0], tile, batch, self.STITCH_METHODS[method],
                                             acc.data_ptr(), _ptr(wsum), map_w, map_row0, acc.shape[0], tile_seq0))

    def blend_finalize(self, method: str, acc: torch.Tensor, wsum: Optional[torch.Tensor], cls_map: torch.Tensor,
                       conf_map: Optional[torch.Tensor]) -> None:
        npx = acc.shape[0] * acc.shape[1]
        assert cls_map.numel() == npx and cls_map.dtype == torch.uint8
        self._check(self._lib.fb_blend_finalize(self._h, acc.data_ptr(), _ptr(wsum), self.STITCH_METHODS[method], npx,
                                                cls_map.data_ptr(), _ptr(conf_map)))

    def detect_zone_host(self, raster, band_idx: Sequence[int], W: int, H: int, row0: int, layout: int,
                         tiles: np.ndarray, tile: int, batch: int, out_cls, out_conf, map_w: int,
                         map_row0: int, map_rows: int) -> None:
        """Host in / host out: upload raster rows, detect, download the class map (synchronous)."""
        def hp(x):
            if x is None:
                return None, None
            if torch.is_tensor(x):
                assert not x.is_cuda and x.dtype == torch.uint8 and x.is_contiguous()
                return x.data_ptr(), tuple(x.shape)
            assert x.dtype == np.uint8 and x.flags["C_CONTIGUOUS"]
            return x.ctypes.data, x.shape
        rptr, rshape = hp(raster)
        bands_total = rshape[0] if layout == FB_LAYOUT_CHW else rshape[2]
        rows = rshape[1] if layout == FB_LAYOUT_CHW else rshape[0]
        bi = (C.c_int32 * len(band_idx))(*band_idx)
        t = make_tiles(tiles)
        self._check(self._lib.fb_detect_zone_host(self._h, rptr, bands_total, bi, len(band_idx), W, H, row0, rows,
                                                  layout, t.ctypes.data, t.shape[0], tile, batch, hp(out_cls)[0],
                                                  hp(out_conf)[0], map_w, map_row0, map_rows))

    def detect_zone_shard(self, raster, band_idx: Sequence[int], W: int, H: int, row0: int, layout: int,
                          tiles: np.ndarray, tile: int, batch: int, out_cls, out_conf, map_w: int, map_row0: int,
                          map_rows: int, truth=None, truth_row0: int = 0, truth_sub: int = 0,
                          cm: Optional[torch.Tensor] = None) -> None:
        """One shard of a zone (fb_detect_zone_shard): host raster rows in, ONLY this shard's write rectangles out
        into the (possibly shared) host maps out_cls / out_conf [map_rows, map_w]. With `truth` (host uint8 rows
        starting at truth_row0, pitch map_w) the shard's confusion matrix is added to `cm` (device int64
        [ncls, ncls])."""
        def hp(x):
            if x is None:
                return None, None
            if torch.is_tensor(x):
                assert not x.is_cuda and x.dtype == torch.uint8 and x.is_contiguous()
                return x.data_ptr(), tuple(x.shape)
            assert x.dtype == np.uint8 and x.flags["C_CONTIGUOUS"]
            return x.ctypes.data, x.shape
        rptr, rshape = hp(raster)
        bands_total = rshape[0] if layout == FB_LAYOUT_CHW else rshape[2]
        rows = rshape[1] if layout == FB_LAYOUT_CHW else rshape[0]
        bi = (C.c_int32 * len(band_idx))(*band_idx)
        t = make_tiles(tiles)
        ncls = 0
        if truth is not None:
            assert cm is not None and cm.dtype == torch.int64 and cm.is_cuda and cm.is_contiguous() and cm.shape[0] == cm.shape[1]
            ncls = int(cm.shape[0])
        self._check(self._lib.fb_detect_zone_shard(self._h, rptr, bands_total, bi, len(band_idx), W, H, row0, rows, layout,
                                                   t.ctypes.data, t.shape[0], tile, batch, hp(out_cls)[0], hp(out_conf)[0],
                                                   map_w, map_row0, map_rows, hp(truth)[0], truth_row0, truth_sub, ncls,
                                                   _ptr(cm)))

    def predict_patches(self, patches: torch.Tensor, tile: int, batch: int, metadata: Optional[np.ndarray] = None) -> torch.Tensor:
        """patches: uint8 device [n, c, tile, tile]; returns uint8 device [n, tile, tile] class ids."""
        assert patches.dtype == torch.uint8 and patches.is_cuda and patches.is_contiguous()
        n = patches.shape[0]
        md = None if metadata is None else np.ascontiguousarray(metadata, dtype=np.float32).reshape(n, METADATA_DIM)
        out = torch.empty((n, tile, tile), dtype=torch.uint8, device=self.device)
        self._check(self._lib.fb_predict_patches(self._h, patches.data_ptr(), _np_ptr(md), n, tile, batch, out.data_ptr()))
        return out

    def confusion(self, pred: torch.Tensor, truth: torch.Tensor, ncls: int, truth_sub: int = 0,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """int64 [ncls, ncls] += histogram of (truth - truth_sub, pred); rows = truth."""
        assert pred.dtype == torch.uint8 and truth.dtype == torch.uint8 and pred.is_cuda and truth.is_cuda
        assert pred.numel() == truth.numel() and pred.is_contiguous() and truth.is_contiguous()
        if out is None:
            out = torch.zeros((ncls, ncls), dtype=torch.int64, device=self.device)
        self._check(self._lib.fb_confusion(self._h, pred.data_ptr(), truth.data_ptr(), pred.numel(), ncls, truth_sub, out.data_ptr()))
        return out

    def confusion_rect(self, pred: torch.Tensor, truth: torch.Tensor, ncls: int, truth_sub: int = 0,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """confusion() over two 2-D uint8 device views of the same shape whose rows may be strided (a rectangle
        cut out of a wider map): fb_confusion_rect."""
        assert pred.dtype == torch.uint8 and truth.dtype == torch.uint8 and pred.is_cuda and truth.is_cuda
        assert pred.dim() == 2 and pred.shape == truth.shape and pred.stride(1) == 1 and truth.stride(1) == 1
        if out is None:
            out = torch.zeros((ncls, ncls), dtype=torch.int64, device=self.device)
        if pred.numel():
            self._check(self._lib.fb_confusion_rect(self._h, pred.data_ptr(), truth.data_ptr(), pred.shape[0], pred.shape[1],
                                                    pred.stride(0) if pred.shape[0] > 1 else pred.shape[1],
                                                    truth.stride(0) if truth.shape[0] > 1 else truth.shape[1],
                                                    ncls, truth_sub, out.data_ptr()))
        return out

    # ------------------------------------------------------------------ multi-GPU (NCCL through the C ABI)
    def comm_init(self, rank: int, world: int, exchange) -> None:
        """Join the library's own NCCL communicator. `exchange(bytes_or_None) -> bytes` hands rank 0's 128-byte id
        to every rank (e.g. a torch.distributed broadcast_object_list wrapper)."""
        ident = None
        if rank == 0:
            buf = (C.c_uint8 * 128)()
            rc = self._lib.fb_comm_unique_id(buf)
            if rc != 0:
                msg = self._lib.fb_last_error(None)
                raise NativeError(rc, msg.decode() if msg else "fb_comm_unique_id failed")
            ident = bytes(buf)
        ident = exchange(ident)
        assert isinstance(ident, (bytes, bytearray)) and len(ident) == 128
        arr = (C.c_uint8 * 128).from_buffer_copy(bytes(ident))
        self._check(self._lib.fb_comm_init(self._h, arr, rank, world))
        self._comm_rank = rank

    def comm_destroy(self) -> None:
        self._check(self._lib.fb_comm_destroy(self._h))

    def allreduce_confusion(self, cm: torch.Tensor) -> torch.Tensor:
        """In-place sum over ranks of the int64 [ncls, ncls] device matrix (fb_allreduce_confusion)."""
        assert cm.dtype == torch.int64 and cm.is_cuda and cm.is_contiguous() and cm.dim() == 2 and cm.shape[0] == cm.shape[1]
        self._check(self._lib.fb_allreduce_confusion(self._h, cm.data_ptr(), cm.shape[0]))
        return cm

    def gather_bytes(self, send: torch.Tensor, counts: Sequence[int], root: int = 0) -> Optional[torch.Tensor]:
        """Every rank's uint8 device tensor `send` (counts[rank] bytes) concatenated in rank order on `root`
        (fb_gather_bytes); returns the uint8 device buffer on root, None elsewhere."""
        assert send.dtype == torch.uint8 and send.is_cuda and send.is_contiguous()
        world = len(counts)
        cnt = (C.c_int64 * world)(*[int(x) for x in counts])
        offs = (C.c_int64 * world)(*np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64).tolist())
        rank_is_root = self._comm_rank == root if hasattr(self, "_comm_rank") else None
        recv = torch.empty(int(sum(counts)), dtype=torch.uint8, device=self.device) if rank_is_root in (True, None) else None
        self._check(self._lib.fb_gather_bytes(self._h, send.data_ptr(), send.numel(), _ptr(recv), offs, cnt, root))
        return recv if rank_is_root in (True, None) else None

    def conv2d(self, x1: torch.Tensor, weights: torch.Tensor, bias: torch.Tensor, KH: int, KW: int, stride: int,
               pad: int, x2: Optional[torch.Tensor] = None, up1: bool = False, residual: Optional[torch.Tensor] = None,
               rowbias: Optional[torch.Tensor] = None, relu: bool = False, out_f32: bool = False, mode: int = -1) -> torch.Tensor:
        """Test hook: NHWC bf16 conv. weights: bf16 [Cout, Kpad] packed (see pack_conv_weight)."""
        B, H1, W1, C1 = x1.shape
        Hin, Win = (H1 * 2, W1 * 2) if up1 else (H1, W1)
        C2 = 0 if x2 is None else x2.shape[3]
        Cout, Kpad = weights.shape
        if mode == 2:  # phase form: weights are [4*Cout, Kpad] (pack_phase_weight), x1 at half resolution
            Cout //= 4
        Hout = (Hin + 2 * pad - KH) // stride + 1
        Wout = (Win + 2 * pad - KW) // stride + 1
        out = torch.empty((B, Hout, Wout, Cout), dtype=torch.float32 if out_f32 else torch.bfloat16, device=self.device)
        self._check(self._lib.fb_conv2d(self._h, x1.data_ptr(), _ptr(x2), C1, C2, int(up1), B, Hin, Win, KH, KW, stride, pad,
                                        Cout, weights.data_ptr(), Kpad, bias.data_ptr(), _ptr(residual), _ptr(rowbias),
                                        int(relu), None if out_f32 else out.data_ptr(), out.data_ptr() if out_f32 else None, mode))
        return out

    def conv2d_halo(self, x1: torch.Tensor, w_oihw: torch.Tensor, bias: torch.Tensor, KH: int, stride: int,
                    x2: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None, relu: bool = False,
                    up2_out: bool = False, out_f32: bool = False) -> torch.Tensor:
        """Test hook: halo-staged conv. w_oihw: host fp32 [Cout, C1+C2, KH, KH] (folded weights)."""
        B, H, W, C1 = x1.shape
        C2 = 0 if x2 is None else x2.shape[3]
        Cout = w_oihw.shape[0]
        pad = KH // 2
        Hout, Wout = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KH) // stride + 1
        # up2_out == 2: phase form (conv of the upsampled x1), also twice the input size; 3 / 4: the depth-to-space
        # forms (3: same size, 4: conv of the upsampled x1); the output always has Cout channels (<= 16 there)
        m = 1 if up2_out == 3 else 2 if up2_out else 1
        out = torch.empty((B, Hout * m, Wout * m, Cout), dtype=torch.float32 if out_f32 else torch.bfloat16, device=self.device)
        w = np.ascontiguousarray(w_oihw.detach().cpu().float().numpy())
        self._check(self._lib.fb_conv2d_halo(self._h, x1.data_ptr(), _ptr(x2), C1, C2, B, H, W, KH, stride, Cout, w.ctypes.data,
                                             bias.data_ptr(), _ptr(residual), int(relu), int(up2_out),
                                             None if out_f32 else out.data_ptr(), out.data_ptr() if out_f32 else None))
        return out

    def debug_activation(self, name: str) -> torch.Tensor:
        cnt = C.c_int64()
        dims = (C.c_int32 * 4)()
        self._check(self._lib.fb_debug_activation(self._h, name.encode(), None, C.byref(cnt), dims))
        dt = torch.float32 if name == "logits" else torch.bfloat16
        out = torch.empty(tuple(dims), dtype=dt, device=self.device)
        self._check(self._lib.fb_debug_activation(self._h, name.encode(), out.data_ptr(), C.byref(cnt), dims))
        return out

    def debug_input_tiles(self) -> torch.Tensor:
        """The normalised tiles of the last forward pass as bf16 [n, T, T, in_channels], whichever way the library
        stored them: [n, T, T, 8] (channel-padded) or, for <= 4 bands, the 2x2 space-to-depth form
        [n, T/2, T/2, 16] with channel (py*2 + px)*in_channels + band (csrc/conv_halo.cuh). Padding lanes must be 0."""
        x0 = self.debug_activation("x0")
        c = self.in_channels
        if x0.shape[-1] == 8:
            assert bool((x0[..., c:] == 0).all())
            return x0[..., :c].contiguous()
        n, t2 = x0.shape[0], x0.shape[1]
        assert x0.shape[-1] == 16 and bool((x0[..., 4 * c:] == 0).all())
        g = x0[..., :4 * c].reshape(n, t2, t2, 2, 2, c)            # Y, X, py, px, band
        return g.permute(0, 1, 3, 2, 4, 5).reshape(n, 2 * t2, 2 * t2, c).contiguous()

    def profile_begin(self) -> None:
        self._check(self._lib.fb_profile_begin(self._h))

    def profile_end(self) -> dict:
        ms = (C.c_float * 4)()
        self._check(self._lib.fb_profile_end(self._h, ms))
        return {"extract_ms": ms[0], "conv_ms": ms[1], "pool_mlp_ms": ms[2], "stitch_ms": ms[3]}

    def profile_forward(self, n: int, tile: int, iters: int = 3) -> dict:
        ms = (C.c_float * 5)()
        self._check(self._lib.fb_profile_forward(self._h, n, tile, iters, ms))
        return {"extract_ms": ms[0], "conv_ms": ms[1], "pool_mlp_ms": ms[2], "stitch_ms": ms[3], "total_ms": ms[4]}


def pack_conv_weight(w: torch.Tensor, cin_pad: Optional[int] = None, cout_pad: Optional[int] = None) -> torch.Tensor:
    """OIHW fp32 -> bf16 [CoutPad, Kpad] with k = (kh*KW + kw)*CinPad + cin (test helper mirroring api.cu)."""
    Cout, Cin, KH, KW = w.shape
    cin_pad = cin_pad or (Cin + 7) // 8 * 8
    cout_pad = cout_pad or (Cout + 15) // 16 * 16
    ktot = KH * KW * cin_pad
    kpad = (ktot + 63) // 64 * 64
    p = torch.zeros((cout_pad, KH, KW, cin_pad), dtype=torch.float32)
    p[:Cout, :, :, :Cin] = w.permute(0, 2, 3, 1)
    out = torch.zeros((cout_pad, kpad), dtype=torch.float32)
    out[:, :ktot] = p.reshape(cout_pad, ktot)
    return out.to(torch.bfloat16)


def lzw_encode(data: bytes | np.ndarray) -> bytes:
    """TIFF-LZW encode a byte block (host codec in libflairb200, GIL released during the call)."""
    lib = load_library()
    src = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else np.ascontiguousarray(data, dtype=np.uint8).ravel()
    cap = int(lib.fb_lzw_bound(src.size))
    dst = np.empty(cap, np.uint8)
    n = lib.fb_lzw_encode(src.ctypes.data, src.size, dst.ctypes.data, cap)
    if n < 0:
        raise RuntimeError("fb_lzw_encode failed")
    return dst[:n].tobytes()


def lzw_decode(data: bytes, expected: int) -> np.ndarray:
    """TIFF-LZW decode into exactly `expected` bytes (short streams are zero-padded like libtiff)."""
    lib = load_library()
    src = np.frombuffer(data, dtype=np.uint8)
    dst = np.zeros(expected, np.uint8)
    n = lib.fb_lzw_decode(src.ctypes.data, src.size, dst.ctypes.data, expected)
    if n < 0:
        raise RuntimeError("corrupt LZW stream")
    return dst


def pack_phase_weight(w: torch.Tensor, C1: int, C2: int) -> torch.Tensor:
    """OIHW fp32 [Cout, C1+C2, 3, 3] -> bf16 [4*Cout, 4*C1 + 9*C2]: sub-pixel phase form of a 3x3 conv on
    [nearest-x2-upsampled x1 (+) x2] (test helper mirroring api.cu::build_conv). Phase p = 2*(oh%2) + (ow%2);
    columns: 4 low-res taps (di, dj) x C1 (taps of the original kernel that read the same low-res pixel are
    summed), then the 9 original taps x C2."""
    Cout = w.shape[0]
    taps = {(0, 0): [0], (0, 1): [1, 2], (1, 0): [0, 1], (1, 1): [2]}   # (parity, low-res tap) -> original taps
    out = torch.zeros((4, Cout, 4 * C1 + 9 * C2), dtype=torch.float32)
    w64 = w.double()
    for pa in range(2):
        for pb in range(2):
            for di in range(2):
                for dj in range(2):
                    s = w64[:, :C1][:, :, taps[(pa, di)]][:, :, :, taps[(pb, dj)]].sum(dim=(2, 3))
                    out[pa * 2 + pb, :, (di * 2 + dj) * C1:(di * 2 + dj + 1) * C1] = s.float()
            if C2:
                out[pa * 2 + pb, :, 4 * C1:] = w[:, C1:].permute(0, 2, 3, 1).reshape(Cout, 9 * C2)
    return out.reshape(4 * Cout, -1).to(torch.bfloat16)
