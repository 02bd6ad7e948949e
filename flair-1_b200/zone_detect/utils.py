"""Config validation, strategy grid, output naming and device set-up for `flair-detect`.

Same public names, YAML keys, error types and messages as src/zone_detect/utils.py of the reference (they are the
CLI contract); the implementation is table-driven: one schema of (check, message) rows, one
itertools grid, one key table for the method grammar of the output file names.
"""
from __future__ import annotations

import datetime
import itertools
import os
from pathlib import Path
from typing import Callable, Iterator

import numpy as np
import torch
import yaml

from .. import geotiff
from .tiles import get_stride

# ---------------------------------------------------------------------------------------------- config schema
_INT = lambda v: type(v) == int  # noqa: E731  (bool is rejected, like the reference's `type(x) == int`)

# (applies(config), check(config), message): evaluated in this order, first failure raises AssertionError(message)
# -- the order and wording of src/zone_detect/utils.py:33-64.
_SCHEMA: list[tuple[Callable[[dict], bool], Callable[[dict], bool], str]] = [
    (lambda c: True, lambda c: os.path.exists(c["input_img_path"]), "Input image path does not exist."),
    (lambda c: c["metrics"], lambda c: os.path.exists(c["truth_path"]), "Ground truth path does not exist."),
    (lambda c: True, lambda c: isinstance(c["channels"], list) and all(isinstance(x, int) for x in c["channels"]),
     "Channels should be a list of integers"),
    (lambda c: True, lambda c: _INT(c["img_pixels_detection"]), "img_pixels_detection should be an integer"),
    (lambda c: True, lambda c: _INT(c["margin"]) and 2 * c["margin"] < c["img_pixels_detection"],
     "Margin should be an integer and less than half of img_pixels_detection"),
    (lambda c: True, lambda c: c["output_type"] in ("class_prob", "argmax"), "Invalid output type: should be argmax or class_prob."),
    (lambda c: True, lambda c: _INT(c["n_classes"]), "n_classes should be an integer"),
    (lambda c: True, lambda c: c["norma_task"][0]["norm_type"] in ("custom", "scaling"),
     "Invalid normalization type: should be custom or scaling."),
    (lambda c: True, lambda c: os.path.isfile(c["model_weights"]), "Model weights file does not exist."),
]

# strategy lists of the compare grid: (section, key, element type, "percentage" message or None) -- utils.py:73-92
_STRATEGY_LISTS = [
    ("tiling", "size_range", int, None),
    ("tiling", "stride_range", float, "Stride should be a percentage"),
    ("stitching", "methods", str, None),
    ("stitching", "margin", float, "Margin should be a percentage"),
]


def read_config(args) -> dict:
    """utils.py:13-23: the YAML file plus the three CLI switches as keys."""
    with open(args.conf, "r") as f:
        config = yaml.safe_load(f)
    config.update(metrics=args.metrics, batch_mode=args.batch_mode, compare=args.compare)
    return preprocess_config(config)


def preprocess_config(config: dict) -> dict:
    """utils.py:26-94."""
    Path(config["output_path"]).mkdir(parents=True, exist_ok=True)
    for applies, check, message in _SCHEMA:
        if applies(config):
            assert check(config), message
    config["input_img_path"] = Path(config["input_img_path"]).with_suffix(".tif")
    if config["metrics"]:
        config["metrics_out"] = config["output_path"] + "/metrics.json"
        config["truth_path"] = Path(config["truth_path"]).with_suffix(".tif")
    ext = os.path.splitext(config["model_weights"])[1]
    if ext not in (".pth", ".ckpt"):
        raise ValueError(f"Model weights should be a .pth or .ckpt file. Got {ext}")
    if config["compare"]:
        for section, key, kind, percentage in _STRATEGY_LISTS:
            values = check_list_type(config["strategies"][section][key], kind)
            config["strategies"][section][key] = values
            if percentage:
                assert all(0 <= v <= 1 for v in values), percentage
    return config


def check_list_type(lst, expected_type: type) -> list:
    """utils.py:97-107: a scalar becomes a one-element list, an iterable keeps its elements of the expected type."""
    if isinstance(lst, expected_type):
        return [lst]
    res = [v for v in lst if isinstance(v, expected_type)] if hasattr(lst, "__iter__") else lst
    assert all(isinstance(v, expected_type) for v in res), f"List should be of type {expected_type}"
    return res


# ---------------------------------------------------------------------------------------------- strategy grid
def _grid_axes(config: dict) -> tuple[list, list, list, list]:
    """(paddings, tile sizes, margins, stitching methods) of utils.py:112-133: a disabled section falls back to
    the plain detect keys."""
    strategies = config.get("strategies", {})
    tiling, stitching = strategies.get("tiling", {}), strategies.get("stitching", {})
    paddings = strategies.get("padding_overall", []) or ["no-padding"]
    sizes = tiling.get("size_range", [config["img_pixels_detection"]]) if tiling.get("enabled", False) else [config["img_pixels_detection"]]
    on = stitching.get("enabled", False)
    margins = stitching.get("margin", [config["margin"]]) if on else [config["margin"]]
    methods = stitching.get("methods", ["exact-clipping"]) if on else ["exact-clipping"]
    return paddings, sizes, margins, methods


def _grid(config: dict) -> Iterator[dict]:
    paddings, sizes, margins, methods = _grid_axes(config)
    for padding, size, margin in itertools.product(paddings, sizes, margins):
        margin = int(margin * size) if margin < 1 else margin          # a fraction of the tile (utils.py:140-141)
        if size <= 2 * margin:
            print(f"""    [x] skipping {size} pixels detection size with {margin} margin...""")
            continue
        strides = get_stride({**config, "margin": margin, "img_pixels_detection": size})
        for stride, method in itertools.product(strides, methods):
            yield {"img_pixels_detection": size, "margin": margin, "padding": padding, "stitching": method, "stride": stride}


def gen_param_combination(config: dict) -> list:
    """utils.py:110-167: the (size, margin, padding, stitching, stride) runs of `-c`, padding outermost, then tile
    size, margin, stride, stitching method."""
    return list(_grid(config))


# ---------------------------------------------------------------------------------------------- file-name grammar
# "<key>=<value>" fields of a method string: key -> (name in the result, parser); anything else passes through
_METHOD_FIELDS = {"size": ("patch_size", int), "stride": ("stride", int), "margin": ("margin", int),
                  "padding": ("padding", str), "stitching": ("stitching", str)}


def extract_method(method: str, info: dict | None = None) -> dict:
    """utils.py:170-188: "size=512_stride=256_..." -> parameters. A value cannot contain "_": the field after it
    has no "=", which is an IndexError here as in the reference ("stitching=average_weights")."""
    info = {} if info is None else info
    for field in method.split("_"):
        parts = field.split("=")
        name, parse = _METHOD_FIELDS.get(parts[0], (parts[0], str))
        info[name] = parse(parts[1])
    return info


def info_extract(file: Path) -> dict:
    """utils.py:191-217: "<dpt>_<year>_<zone...>_<type>-ARGMAX-S_<method>.tif" -> dpt, zone, method and the method's
    parameters ("dpt" only when the name does not already start with "D", like the reference)."""
    filename = str(file)
    if not filename.endswith(".tif"):
        raise ValueError("Filename should end with .tif what are you doing ?")
    stem = filename.rsplit("/", 1)[-1].split(".")[0]
    region, method = stem.split("-ARGMAX-S_")
    words = region.split("_")
    info = {}
    if not words[0].startswith("D"):
        info["dpt"] = "D" + "_".join(words[:2])
    info["zone"] = "_".join(words[2:-1])
    info["method"] = method
    return extract_method(method, info)


# ---------------------------------------------------------------------------------------------- set-up
def setup_out_path(config: dict, stamp: str | None = None) -> dict:
    """utils.py:221-236: `-c` runs get a time-stamped folder of their own. `stamp`: the folder name chosen by
    rank 0 of a multi-process run, so that all ranks agree on it."""
    out = Path(config["output_path"])
    if config["compare"]:
        out = out / (stamp or datetime.datetime.now().strftime("%Y%m%d_%H%M%S"))
        print(f"Creating output directory: {out}")
    out.mkdir(parents=True, exist_ok=True)
    config["local_out"] = out
    return config


def setup_device(config: dict) -> tuple[torch.device, bool]:
    """utils.py:239-245 -- except that this implementation has no CPU path: asking for the CPU, or running
    without a CUDA device, is an error instead of a silent slow run."""
    if not torch.cuda.is_available():
        raise RuntimeError("flair1_b200 runs on a B200 (sm_100a) only: no CUDA device is visible and there is no CPU fallback")
    if not config["use_gpu"]:
        raise RuntimeError("use_gpu: false is not supported by flair1_b200 (no CPU fallback); use the reference implementation for CPU runs")
    return torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))), True


def setup(args) -> tuple[dict, torch.device, bool]:
    """utils.py:248-253."""
    config = read_config(args)
    return (config, *setup_device(config))


def setup_indiv_path(config: dict, identifier: str) -> tuple[dict, str]:
    """utils.py:256-279: never overwrite -- the first free name of <name>.tif, <name>_1.tif, <name>_2.tif, ..."""
    name = config["output_name"] + identifier
    name += "" if name.endswith(".tif") else ".tif"
    stem, ext = os.path.splitext(name)
    candidates = itertools.chain([name], (f"{stem}_{k}{ext}" for k in itertools.count(1)))
    return config, next(p for p in (os.path.join(config["local_out"], c) for c in candidates) if not os.path.exists(p))


def metrics_json_path(config: dict, local_out: Path) -> Path:
    """utils.py:290-294: the per-patch metrics file is named after the two parent directories of the input image."""
    parts = Path(config["input_img_path"]).parts[-3:-1]
    dpt, zone = (tuple(parts) + ("", ""))[:2]
    return Path(local_out) / f"metrics_per-patch_{dpt}_{zone}.json"


def open_images(config: dict, local_out: Path, get_truth: bool, rows: tuple[int, int] | None = None):
    """utils.py:282-297: truth = band 1 minus 1 in uint8 (0 wraps to 255 and is dropped by the confusion matrix).
    `rows`: read only truth rows [r0, r1) (what one rank of a sharded zone scores)."""
    if not get_truth:
        return np.zeros((1, 1), dtype=np.uint8), Path()
    path = Path(config["truth_path"])
    if rows is None:
        truth = geotiff.read(path, bands=[1])[0]
    else:
        info = geotiff.read_info(path)
        truth = geotiff.read(path, bands=[1], window=(0, rows[0], info.width, rows[1] - rows[0]))[0]
    return truth - np.uint8(1), metrics_json_path(config, local_out)
