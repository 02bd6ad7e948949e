"""Config reading / validation, output-path and device set-up for `flair-detect`
(mirrors src/zone_detect/utils.py of the reference: same function names, same keys, same errors)."""
from __future__ import annotations

import datetime
import os
from pathlib import Path

import numpy as np
import torch
import yaml

from .. import geotiff
from .tiles import get_stride


#### CONFIG ####
def read_config(args) -> dict:
    """src/zone_detect/utils.py:13-23: YAML + the three CLI flags injected as keys."""
    with open(args.conf, "r") as f:
        config = yaml.safe_load(f)
    config["metrics"] = args.metrics
    config["batch_mode"] = args.batch_mode
    config["compare"] = args.compare
    return preprocess_config(config)


def preprocess_config(config: dict) -> dict:
    """src/zone_detect/utils.py:26-94: same assertions and messages."""
    Path(config["output_path"]).mkdir(parents=True, exist_ok=True)
    assert os.path.exists(config["input_img_path"]), "Input image path does not exist."
    config["input_img_path"] = Path(config["input_img_path"]).with_suffix(".tif")

    if config["metrics"]:
        config["metrics_out"] = config["output_path"] + "/metrics.json"
        assert os.path.exists(config["truth_path"]), "Ground truth path does not exist."
        config["truth_path"] = Path(config["truth_path"]).with_suffix(".tif")

    assert isinstance(config["channels"], list) and all(
        isinstance(c, int) for c in config["channels"]
    ), "Channels should be a list of integers"

    assert type(config["img_pixels_detection"]) == int, "img_pixels_detection should be an integer"
    assert (
        type(config["margin"]) == int and 2 * config["margin"] < config["img_pixels_detection"]
    ), "Margin should be an integer and less than half of img_pixels_detection"
    assert config["output_type"] in ["class_prob", "argmax"], "Invalid output type: should be argmax or class_prob."
    assert type(config["n_classes"]) == int, "n_classes should be an integer"
    assert config["norma_task"][0]["norm_type"] in [
        "custom",
        "scaling",
    ], "Invalid normalization type: should be custom or scaling."

    assert os.path.isfile(config["model_weights"]), "Model weights file does not exist."
    if os.path.splitext(config["model_weights"])[1] not in [".pth", ".ckpt"]:
        raise ValueError(
            "Model weights should be a .pth or .ckpt file. " f"Got {os.path.splitext(config['model_weights'])[1]}"
        )

    if config["compare"]:
        config["strategies"]["tiling"]["size_range"] = check_list_type(config["strategies"]["tiling"]["size_range"], int)
        config["strategies"]["tiling"]["stride_range"] = check_list_type(config["strategies"]["tiling"]["stride_range"], float)
        assert all(i >= 0 and i <= 1 for i in config["strategies"]["tiling"]["stride_range"]), "Stride should be a percentage"
        config["strategies"]["stitching"]["methods"] = check_list_type(config["strategies"]["stitching"]["methods"], str)
        config["strategies"]["stitching"]["margin"] = check_list_type(config["strategies"]["stitching"]["margin"], float)
        assert all(i >= 0 and i <= 1 for i in config["strategies"]["stitching"]["margin"]), "Margin should be a percentage"
    return config


def check_list_type(lst: list, expected_type: type) -> list:
    """src/zone_detect/utils.py:97-107."""
    res = lst
    if isinstance(lst, expected_type):
        res = [lst]
    elif hasattr(lst, "__iter__"):
        res = [i for i in lst if isinstance(i, expected_type)]
    assert all(isinstance(i, expected_type) for i in res), f"List should be of type {expected_type}"
    return res


def gen_param_combination(config: dict) -> list:
    """src/zone_detect/utils.py:110-167: the (size, margin, padding, stitching, stride) grid of `-c`."""
    combi = []
    padding_list = config.get("strategies", {}).get("padding_overall", []) or ["no-padding"]
    tiling_cfg = config.get("strategies", {}).get("tiling", {})
    if tiling_cfg.get("enabled", False):
        tile_size_list = tiling_cfg.get("size_range", [config["img_pixels_detection"]])
    else:
        tile_size_list = [config["img_pixels_detection"]]
    stitching_cfg = config.get("strategies", {}).get("stitching", {})
    if stitching_cfg.get("enabled", False):
        margin_list = stitching_cfg.get("margin", [config["margin"]])
        stitching_methods = stitching_cfg.get("methods", ["exact-clipping"])
    else:
        margin_list = [config["margin"]]
        stitching_methods = ["exact-clipping"]
    for padding in padding_list:
        for img_pixels_detection in tile_size_list:
            for margin in margin_list:
                if margin < 1:
                    margin = int(margin * img_pixels_detection)
                if img_pixels_detection <= 2 * margin:
                    print(f"""    [x] skipping {img_pixels_detection} pixels detection size with {margin} margin...""")
                    continue
                tmp_config = config.copy()
                tmp_config["margin"] = margin
                tmp_config["img_pixels_detection"] = img_pixels_detection
                for stride in get_stride(tmp_config):
                    for stitch in stitching_methods:
                        combi.append({"img_pixels_detection": img_pixels_detection, "margin": margin, "padding": padding,
                                      "stitching": stitch, "stride": stride})
    return combi


def extract_method(method: str, info: dict | None = None) -> dict:
    """src/zone_detect/utils.py:170-188: "size=512_stride=256_..." -> parameters. Values cannot contain "_"
    ("stitching=average_weights" raises IndexError there as well)."""
    info = {} if info is None else info
    for param in method.split("_"):
        if param.startswith("size="):
            info["patch_size"] = int(param.split("=")[1])
        elif param.startswith("stride="):
            info["stride"] = int(param.split("=")[1])
        elif param.startswith("margin="):
            info["margin"] = int(param.split("=")[1])
        elif param.startswith("padding="):
            info["padding"] = param.split("=")[1]
        elif param.startswith("stitching="):
            info["stitching"] = param.split("=")[1]
        else:
            kv = param.split("=")
            info[kv[0]] = kv[1]
    return info


def info_extract(file: Path) -> dict:
    """src/zone_detect/utils.py:191-217: "<dpt>_<year>_<zone...>_<type>-ARGMAX-S_<method>.tif" -> dpt, zone and
    the method parameters. As in the reference, "dpt" is only set when the name does not already start with "D"."""
    filename = str(file)
    if not filename.endswith(".tif"):
        raise ValueError("Filename should end with .tif what are you doing ?")
    name = filename.split("/")[-1].split(".")[0]
    info = {}
    region_type, method = name.split("-ARGMAX-S_")
    region_type = region_type.split("_")
    dpt, zone = region_type[:2], region_type[2:-1]
    if not dpt[0].startswith("D"):
        info["dpt"] = "D" + "_".join(dpt)
    info["zone"] = "_".join(zone)
    info["method"] = method
    return extract_method(method, info)


#### SETUP ####
def setup_out_path(config: dict) -> dict:
    """src/zone_detect/utils.py:221-236."""
    output = Path(config["output_path"])
    output.mkdir(parents=True, exist_ok=True)
    child_dir = output
    if config["compare"]:
        current_time = datetime.datetime.now().strftime("%Y%m%d_%H%M%S")
        child_dir = child_dir / Path(current_time)
        os.makedirs(child_dir, exist_ok=True)
        print(f"Creating output directory: {child_dir}")
    config["local_out"] = child_dir
    return config


def setup_device(config: dict) -> tuple[torch.device, bool]:
    """src/zone_detect/utils.py:239-245 -- except that this implementation has no CPU path: asking for
    the CPU, or running without a CUDA device, is an error instead of a silent slow run."""
    if not torch.cuda.is_available():
        raise RuntimeError("flair1_b200 runs on a B200 (sm_100a) only: no CUDA device is visible and there is no CPU fallback")
    if not config["use_gpu"]:
        raise RuntimeError("use_gpu: false is not supported by flair1_b200 (no CPU fallback); use the reference implementation for CPU runs")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return torch.device("cuda", local), True


def setup(args) -> tuple[dict, torch.device, bool]:
    """src/zone_detect/utils.py:248-253."""
    config = read_config(args)
    device, use_gpu = setup_device(config)
    return config, device, use_gpu


def setup_indiv_path(config: dict, identifier: str) -> tuple[dict, str]:
    """src/zone_detect/utils.py:256-279: never overwrite, append _1, _2, ..."""
    out_name = config["output_name"] + identifier
    if not out_name.endswith(".tif"):
        out_name += ".tif"
    base_name = out_name
    path_out = os.path.join(config["local_out"], base_name)
    filename, ext = os.path.splitext(base_name)
    counter = 1
    while os.path.exists(path_out):
        path_out = os.path.join(config["local_out"], f"{filename}_{counter}{ext}")
        counter += 1
    return config, path_out


def open_images(config: dict, local_out: Path, get_truth: bool):
    """src/zone_detect/utils.py:282-297: truth = band 1 - 1 (uint8 wrap: 0 -> 255, dropped by the
    confusion matrix); metrics file named after the two parent directories of the input image."""
    if get_truth:
        truth_array = geotiff.read(Path(config["truth_path"]), bands=[1])[0] - np.uint8(1)
        parts = Path(config["input_img_path"]).parts[-3:-1]
        dpt, zone = (parts + ("", ""))[:2] if len(parts) < 2 else parts
        metrics_json = local_out / Path(f"metrics_per-patch_{dpt}_{zone}.json")
    else:
        truth_array = np.zeros((1, 1), dtype=np.uint8)
        metrics_json = Path()
    return truth_array, metrics_json
