"""`flair-detect --conf x.yaml [-c] [-m] [-b]` on B200 (mirrors src/zone_detect/main.py).

Same CLI flags, same YAML keys, same log tee, same output raster (uint8, LZW, tiled at
img_pixels_detection, 2 bands for `argmax`: class index and max probability), same "never overwrite"
file naming. The hot loop of the reference (main.py:398-426: DataLoader -> H2D -> forward -> softmax ->
D2H of all probabilities -> numpy argmax -> per-tile GDAL write) becomes: read the raster rows once into pinned
memory, one pipelined `fb_detect_zone_shard` call per rank (chunked upload, compute and row-band download
overlapped; with `-m` the truth rows ride along and the confusion matrix is accumulated on the GPU), write the
GeoTIFF.

Multi-GPU: launched under `torchrun`, every rank takes a contiguous range of the row-major tile order (halo rows
are re-read, no exchange) and copies its write rectangles straight into ONE output map in shared memory
(shared_map.SharedHostMap) -- the reference's single output raster (main.py:421-426) -- which rank 0 writes to
disk after a barrier; with `-m` the per-rank confusion matrices are summed with one NCCL all-reduce
(fb_allreduce_confusion). torch.distributed only carries the rendezvous: barriers and a few small objects.
"""
from __future__ import annotations

import argparse
import datetime
import json
import os
import sys
import warnings
from pathlib import Path

import numpy as np
import torch

from .. import geotiff
from .compare import STITCH_METHODS, detect_zone, detect_zone_pipelined
from .dataset import Sliced_Dataset
from .metrics import batch_metrics, confusion_matrix_gpu, metrics_from_confmat
from .model import load_model
from .shared_map import SharedHostMap
from .slicing_job import owned_rects, slice_extent, split_rows_across_ranks, split_tiles_across_ranks, tile_windows
from .tiles import get_stride
from .utils import (gen_param_combination, metrics_json_path, open_images, setup, setup_device, setup_indiv_path,
                    setup_out_path)

warnings.simplefilter(action="ignore", category=FutureWarning)

argParser = argparse.ArgumentParser()
argParser.add_argument("--conf", help="Path to the .yaml config file")
argParser.add_argument("-c", "--compare", help="Compare different methods", action="store_true")
argParser.add_argument("-m", "--metrics", help="Compute metrics", action="store_true")
argParser.add_argument("-b", "--batch_mode", help="Run the pipeline for a batch of images", action="store_true")


def _rank_world() -> tuple[int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


class Logger(object):
    """main.py:52-64: tee stdout/stderr to a log file (rank 0 only writes the file)."""

    def __init__(self, filename: str = "Default.log") -> None:
        self.terminal = sys.stdout
        self.log = open(filename, "w", encoding="utf-8") if _rank_world()[0] == 0 else None
        self.encoding = self.terminal.encoding

    def write(self, message: str) -> None:
        self.terminal.write(message)
        if self.log:
            self.log.write(message)

    def flush(self) -> None:
        self.terminal.flush()
        if self.log:
            self.log.flush()


def conf_log(config: dict, resolution: tuple[float, float], img_size: list[int]) -> None:
    """main.py:68-119 (the `strategies` block is optional here: the reference raises KeyError on the
    plain detect YAML)."""
    strategies = config.get("strategies", {})
    print(f"""
    |- output path: {config['output_path']}
    |- output raster name: {config['output_name']}

    |- input image path: {config['input_img_path']}
    |- channels: {config['channels']}
    |- resolution: {resolution}
    |- image size: {img_size}

    |- output type: {config['output_type']}
    |- normalization: {config['norma_task'][0]['norm_type']}
    |- number of classes: {config['n_classes']}

    |- model weights path: {config['model_weights']}
    |- model template: {config['model_framework']['model_provider']}

    |- device: cuda (B200, libflairb200)
    |- batch size: {config['batch_size']}

    |- tiling: size {config['img_pixels_detection']}, margin {config['margin']}, strategies {bool(strategies)}
    \n\n""")


# __________Prepare objects___________#
def prepare_tiles(config: dict, stride: int):
    """main.py:123-148."""
    tiles, profile, resolution, img_size = slice_extent(
        in_img=Path(config["input_img_path"]), patch_size=config["img_pixels_detection"], margin=config["margin"],
        output_name=config["output_name"], output_path=Path(config["local_out"]),
        write_dataframe=config["write_dataframe"] and _rank_world()[0] == 0, stride=stride)
    if _rank_world()[0] == 0:
        conf_log(config, resolution, img_size)
        print(f"""    [x] sliced input raster to {len(tiles)} squares...""")
    return tiles, profile, resolution


def _pipelined(config: dict) -> bool:
    """True for the runs that go through fb_detect_zone_shard: exact clipping to a class map without per-patch
    metrics. The blended stitchings, class_prob and the per-patch metrics of `-c -m` need every logit of a tile
    (or their neighbours' accumulators) and keep whole tile rows per rank."""
    per_patch = config["metrics"] and config["compare"] and "classes" in config and len(config["classes"]) == config["n_classes"]
    return config["output_type"] == "argmax" and config.get("stitching", "exact-clipping") == "exact-clipping" and not per_patch


def prepare_data(config: dict, stride: int):
    """main.py:151-183: tile table + the raster rows this rank needs (no DataLoader: tiles are cut on
    the GPU)."""
    tiles, profile, resolution = prepare_tiles(config, stride)
    rank, world = _rank_world()
    shard = (split_tiles_across_ranks if _pipelined(config) else split_rows_across_ranks)(tiles, world)[rank]
    my_tiles = tiles[shard]
    size = config["img_pixels_detection"]
    # the margin-cropped window of every tile of this rank (per-patch metrics of the compare loop) and its
    # position in the write order
    config["_my_windows"] = tile_windows(profile["width"], profile["height"], size, config["margin"], stride)[shard]
    config["_my_index"] = shard
    if len(my_tiles) and config.get("stitching", "exact-clipping") != "exact-clipping" and config["output_type"] == "argmax":
        # blended stitching: every tile that touches the rows this rank owns contributes to them, so the
        # neighbouring tile rows are recomputed here instead of exchanged (halo recompute, like the raster halo)
        own0, own1 = int(my_tiles[:, 3].min()), int(my_tiles[:, 5].max())
        my_tiles = tiles[(tiles[:, 1] < own1) & (tiles[:, 1] + size > own0)]
        config["_own_rows"] = (own0, own1)
    row_range = (int(my_tiles[:, 1].min()), int(my_tiles[:, 1].max()) + size) if len(my_tiles) else (0, 0)
    dataset = Sliced_Dataset(dataframe=my_tiles, img_path=config["input_img_path"], resolution=resolution,
                             bands=config["channels"], patch_detection_size=size, norma_dict=config["norma_task"],
                             row_range=row_range)
    return dataset, my_tiles, tiles, profile


def prepare_model(config: dict, device: torch.device):
    """main.py:186-203."""
    if _rank_world()[0] == 0:
        print(f"""
    ##############################################
    ZONE DETECTION
    ##############################################

    CUDA available? {torch.cuda.is_available()}""")
    model = load_model(config, device)
    nt = config["norma_task"][0]
    model.set_norm(nt["norm_type"], nt.get("norm_means", []), nt.get("norm_stds", []), channels=len(config["channels"]))
    if _rank_world()[0] == 0:
        print("""    [x] loaded model and weights...""")
    return model


def prepare_output(config: dict, profile: dict, identifier: str = "") -> tuple[dict, str]:
    """main.py:206-232: output profile (uint8, LZW, BIGTIFF, tiled at img_pixels_detection, 2 bands for
    argmax / n_classes bands for class_prob) and a fresh path."""
    config, path_out = setup_indiv_path(config, identifier)
    out_profile = dict(profile)
    out_profile.update({"dtype": "uint8", "compress": "LZW", "driver": "GTiff", "BIGTIFF": "YES", "tiled": True,
                        "blockxsize": config["img_pixels_detection"], "blockysize": config["img_pixels_detection"]})
    out_profile["count"] = 2 if config["output_type"] == "argmax" else config["n_classes"]
    return out_profile, path_out


def _write_output(path_out: str, bands: np.ndarray, profile: dict) -> None:
    block = int(profile["blockxsize"])
    block = block if block % 16 == 0 else 512
    geotiff.write(path_out, bands, geo_tags=profile.get("geo_tags"), compress="lzw", tiled=True, blocksize=block, bigtiff=True)


# _________PIPELINES__________#
def run_from_config(config: dict) -> None:
    device, use_gpu = setup_device(config)
    run_pipeline(config, device, use_gpu)


class OutputMap:
    """The output raster of one run while it is being produced: uint8 [bands, H, W] on the host, filled by every
    rank (shared memory when there are several, see shared_map.py), written to disk by rank 0."""

    def __init__(self, bands: int, H: int, W: int) -> None:
        rank, world = _rank_world()
        self.shared = None
        if world == 1:
            self._tensor = torch.zeros((bands, H, W), dtype=torch.uint8, pin_memory=torch.cuda.is_available())
            self.array = self._tensor.numpy()
            return
        import torch.distributed as dist
        box = [SharedHostMap.fresh_path(bands * H * W) if rank == 0 else None]
        if rank == 0:
            self.shared = SharedHostMap(box[0], bands, H, W, create=True)   # a fresh tmpfs file reads as zeros
        dist.broadcast_object_list(box, src=0)
        if rank != 0:
            self.shared = SharedHostMap(box[0], bands, H, W, create=False)
        self.array = self.shared.array

    def pin_rows(self, row0: int, row1: int) -> None:
        if self.shared is not None and torch.cuda.is_available() and row1 > row0:
            self.shared.pin_rows(row0, row1)

    def put_rows(self, band: int, row0: int, strip: torch.Tensor) -> None:
        """Rows [row0, row0 + len(strip)) of `band` from a device (or host) tensor."""
        if strip.shape[0]:
            torch.from_numpy(self.array[band, row0:row0 + strip.shape[0]]).copy_(strip)

    def finish(self) -> np.ndarray | None:
        """Barrier; the complete map on rank 0 (None elsewhere)."""
        rank, world = _rank_world()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        return self.array if rank == 0 else None

    def close(self) -> None:
        rank, world = _rank_world()
        self.array = None
        if self.shared is not None:
            if world > 1:
                import torch.distributed as dist
                dist.barrier()   # nobody unmaps / unlinks while the writer still reads
            self.shared.close()
            if rank == 0:
                self.shared.unlink()
            self.shared = None


def _gather_patch_metrics(cfg: dict, method: str) -> list | None:
    """Per-patch metric dictionaries (test/metrics.py:165-192) of every tile in write order; rank 0 receives the
    other ranks' confusion matrices (they are tiny: n_tiles x n_classes^2 int64)."""
    rank, world = _rank_world()
    cm = cfg.get("_patch_cm")
    mine = (np.asarray(cfg["_my_index"]), np.asarray(cfg["_my_windows"]),
            cm.cpu().numpy() if cm is not None else np.zeros((0, cfg["n_classes"], cfg["n_classes"]), np.int64))
    parts = [mine]
    if world > 1:
        import torch.distributed as dist
        parts = [None] * world if rank == 0 else None
        dist.gather_object(mine, parts, dst=0)
    if rank != 0:
        return None
    rows = []
    for idx, win, cms in parts:
        for i, w, c in zip(idx, win, cms):
            rows.append((int(i), metrics_from_confmat(c, cfg, f"{method}_{int(w[2])}_{int(w[3])}")))
    return [m for _, m in sorted(rows, key=lambda r: r[0])]


def batch_metrics_pipeline(config: dict, truth_dpt: Path, device: torch.device, use_gpu: bool) -> str | None:
    """main.py:440-497: every zone directory of the department that holds a `*<data_type>.tif` image and has a
    ground-truth raster is run through run_pipeline (with `-c` the predictions land in time-stamped folders, which
    is what batch_metrics collects), then the per-method metrics over all zones are written to metrics_out."""
    out_json = Path(config["metrics_out"])
    file_pattern = f"*{config['data_type']}.tif"
    assert out_json, "Please provide an output path for the metrics"
    inputs_dpt = Path(config["input_path"])
    for full_zone in sorted(p for p in inputs_dpt.iterdir() if p.is_dir()):
        irc_path = next(full_zone.glob(file_pattern), None)
        if irc_path is None:
            continue
        zone = irc_path.parts[-2]
        truth_path = next(Path(truth_dpt / zone).glob("*.tif"), None)
        if truth_path is None:
            print(f"No ground truth found for zone: {zone}")
            continue
        config.update({"input_img_path": str(irc_path), "truth_path": str(truth_path),
                       "output_name": f"{irc_path.stem}-ARGMAX-S"})
        run_pipeline(config, device, use_gpu)
    out = out_json.with_suffix(".json")
    if _rank_world()[0] != 0:
        return None
    metrics_file = batch_metrics(config, truth_dpt)
    with open(out, "w") as f:
        json.dump(metrics_file, f)
    print(f"Metrics saved to {out}")
    return str(out)


def run_pipeline(config: dict, device: torch.device, use_gpu: bool) -> dict:
    """main.py:244-437, default branch (exact clipping, default tiling) plus `-m` whole-raster metrics;
    with `-c` the same loop runs once per entry of the strategy grid (tile size, stride, margin, stitching
    method: exact-clipping / average / average_weights / max). output_type "class_prob" writes n_classes
    probability bands instead of class + confidence."""
    rank, world = _rank_world()
    stamp = datetime.datetime.now().strftime("%Y%m%d_%H%M%S")
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo", device_id=device if torch.cuda.is_available() else None)
        box = [stamp]
        dist.broadcast_object_list(box, src=0)   # one time stamp for all ranks: one output folder, one log name
        stamp = box[0]
    config = setup_out_path(config, stamp)
    local_out = Path(config["local_out"])
    log_filename = local_out / Path(f"{config['output_name']}_{stamp}.log")
    old_out, old_err = sys.stdout, sys.stderr
    sys.stdout = Logger(filename=str(log_filename))
    sys.stderr = sys.stdout
    result = {}
    model = None
    try:
        if rank == 0:
            print(f"    [LOGGER] Writing logs to: {log_filename}")
        model = prepare_model(config, device)
        if world > 1:
            def exchange(ident):
                import torch.distributed as dist
                box = [ident]
                dist.broadcast_object_list(box, src=0)
                return box[0]
            model.comm_init(rank, world, exchange)   # the library's own NCCL communicator (confusion-matrix all-reduce)
        # a single process scores against the whole truth raster; the ranks of a sharded run read their rows only
        truth_array = open_images(config, local_out, True)[0] if (config["metrics"] and world == 1) else None
        metrics_json = metrics_json_path(config, local_out) if config["metrics"] else Path()

        if config["compare"]:
            # the weighted stitchings (average / average_weights / max) run as stitching_blend() implements
            # them; anything else in the grid is reported and skipped
            grid = gen_param_combination(config)
            settings = [c for c in grid if c["stitching"] in STITCH_METHODS]
            if len(settings) != len(grid) and rank == 0:
                print(f"    [x] {len(grid) - len(settings)} strategy combination(s) name an unknown stitching method; skipped")
        else:
            settings = [{"img_pixels_detection": config["img_pixels_detection"], "margin": config["margin"],
                         "padding": "no-padding", "stitching": config.get("stitching", "exact-clipping"),
                         "stride": get_stride(config)[0]}]
            if settings[0]["stitching"] not in STITCH_METHODS:
                raise ValueError(f"stitching must be one of {STITCH_METHODS}")

        def truth_rows_of(r0: int, r1: int) -> np.ndarray:
            if truth_array is not None:
                return truth_array[r0:r1]
            return open_images(config, local_out, True, rows=(r0, r1))[0]

        method_metrics = []
        method_times = {}
        patch_metrics = {}
        for combi in settings:
            cfg = dict(config)
            cfg.update({"img_pixels_detection": combi["img_pixels_detection"], "margin": combi["margin"],
                        "padding": combi["padding"], "stride": combi["stride"], "stitching": combi["stitching"]})
            method = (f"size={combi['img_pixels_detection']}_stride={combi['stride']}_margin={combi['margin']}"
                      f"_padding={combi['padding']}_stitching={combi['stitching']}")
            identifier = "_" + method if config["compare"] else ""
            start_time = datetime.datetime.now()
            dataset, my_tiles, tiles, profile = prepare_data(cfg, combi["stride"])
            out_profile, path_out = prepare_output(cfg, profile, identifier)
            if rank == 0:
                print("""    [ ] starting inference...\n""")
            H, W = dataset.raster_height, dataset.raster_width
            want_metrics = config["metrics"] and cfg["output_type"] == "argmax"
            n_classes = len(config["classes"]) if "classes" in config else config["n_classes"]
            cm = torch.zeros((n_classes, n_classes), dtype=torch.int64, device=device) if want_metrics else None
            per_patch = want_metrics and config["compare"] and "classes" in config and len(config["classes"]) == config["n_classes"]
            out = OutputMap(out_profile["count"], H, W)
            own0, own1 = cfg.get("_own_rows", (int(my_tiles[:, 3].min()), int(my_tiles[:, 5].max()))) if len(my_tiles) else (0, 0)
            if _pipelined(cfg):
                truth_rows = None
                if want_metrics and len(my_tiles):
                    truth_rows = torch.from_numpy(np.ascontiguousarray(truth_rows_of(own0, own1)))
                    truth_rows = truth_rows.pin_memory() if torch.cuda.is_available() else truth_rows
                detect_zone_pipelined(cfg, model, dataset, my_tiles, out, truth_rows, own0, cm)
            else:
                # rows of the device maps: what this rank owns, widened to its tiles' whole metric windows for the
                # per-patch metrics (a clamped or overlapping window reaches into a neighbouring rank's rows)
                span0, span1 = own0, own1
                if per_patch and len(my_tiles):
                    win = np.asarray(cfg["_my_windows"])
                    span0, span1 = min(own0, int(win[:, 3].min())), max(own1, int(win[:, 5].max()))
                truth_dev = None
                if want_metrics and len(my_tiles):
                    truth_dev = torch.from_numpy(np.ascontiguousarray(truth_rows_of(span0, span1))).to(device)
                cls, conf, row0, rows = detect_zone(cfg, model, dataset, my_tiles, device, combi["stitching"],
                                                    truth_dev if per_patch else None, (span0, span1) if len(my_tiles) else None)
                if want_metrics and len(my_tiles):
                    confusion_matrix_gpu(model, cls[own0 - row0:own1 - row0], truth_dev[own0 - span0:own1 - span0], n_classes, out=cm)
                out.pin_rows(own0, own1)
                if cfg["output_type"] == "argmax":
                    out.put_rows(0, own0, cls[own0 - row0:own1 - row0])
                    out.put_rows(1, own0, conf[own0 - row0:own1 - row0])
                else:
                    # class_prob: `cls` holds the n_classes probability planes (band k + 1 = class k, main.py:424-426)
                    for k in range(cls.shape[0]):
                        out.put_rows(k, own0, cls[k, own0 - row0:own1 - row0])
            if want_metrics:
                if world > 1:
                    model.allreduce_confusion(cm)
                result["confmat"] = cm.cpu().numpy()
            t_detect = (datetime.datetime.now() - start_time).total_seconds()
            full = out.finish()
            if rank == 0:
                _write_output(path_out, full, out_profile)
            out.close()
            dataset.close_raster()
            elapsed = (datetime.datetime.now() - start_time).total_seconds()
            method_times.setdefault(method, []).append(elapsed * 1000)   # main.py:352-358 (per zone here, not per patch)
            if per_patch and combi["stitching"] == "exact-clipping":
                got = _gather_patch_metrics(cfg, method)
                if rank == 0:
                    patch_metrics[method] = got
            if rank == 0:
                print(f"""    [X] done writing to {path_out.split('/')[-1]} raster file ({elapsed:.2f} s, {len(tiles)} tiles, """
                      f"""{H * W / 1e6 / max(elapsed, 1e-9):.1f} Mpx/s incl. I/O; read + detect {t_detect:.2f} s = """
                      f"""{H * W / 1e6 / max(t_detect, 1e-9):.1f} Mpx/s).\n""")
                result.setdefault("outputs", []).append(path_out)
                result.setdefault("mpx_per_s_incl_io", []).append(H * W / 1e6 / max(elapsed, 1e-9))
                result.setdefault("mpx_per_s_read_detect", []).append(H * W / 1e6 / max(t_detect, 1e-9))
                if config["metrics"] and "classes" in config:
                    method_metrics.append(metrics_from_confmat(result["confmat"], config, method))
        config["times"] = method_times   # main.py:378, read by batch_metrics
        if rank == 0 and config["metrics"] and method_metrics:
            if patch_metrics:
                # the reference dumps the per-patch list of every method to the same file, the last one stays
                # (main.py:379-384); the whole-raster figures per method go to a file of their own
                with open(metrics_json, "w") as f:
                    json.dump(list(patch_metrics.values())[-1], f, indent=2)
                per_method = metrics_json.with_name(metrics_json.name.replace("metrics_per-patch", "metrics_per-method"))
                with open(per_method, "w") as f:
                    json.dump(method_metrics, f, indent=2)
                result["patch_metrics"] = patch_metrics
                result["metrics_per_method_json"] = str(per_method)
            else:
                with open(metrics_json, "w") as f:
                    json.dump(method_metrics, f, indent=2)
            print(f"""    [X] done writing metrics to {metrics_json.name} file.\n""")
            result["metrics_json"] = str(metrics_json)
            result["metrics"] = method_metrics
    finally:
        sys.stdout, sys.stderr = old_out, old_err
        if model is not None:
            model.close()   # frees the context (and its NCCL communicator) now rather than at garbage collection
    return result


def main() -> None:
    """main.py:501-515."""
    args = argParser.parse_args()
    config, device, use_gpu = setup(args)
    if args.batch_mode:
        gt_dir = Path(config["truth_root"])
        gt_dpt = gt_dir / Path(config["truth_path"]).parts[-3]
        batch_metrics_pipeline(config, gt_dpt, device, use_gpu)
    else:
        run_pipeline(config, device, use_gpu)


if __name__ == "__main__":
    main()
