"""Whole-raster and per-window metrics (mirrors src/zone_detect/test/metrics.py, a runtime module of the
reference): confusion matrix on the GPU (K9), ratios on the host in float64."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from .. import _native, geotiff
from .utils import extract_method, info_extract


def clean_confmat(confmat: np.ndarray, config: dict) -> np.ndarray:
    """test/metrics.py:18-29: drop rows/columns of classes whose weight is 0."""
    weights = np.array([class_info[0] for class_info in config["classes"].values()])
    unused_classes = np.where(weights == 0)[0]
    if unused_classes.size > 0:
        confmat_cleaned = np.delete(confmat, unused_classes, axis=0)
        return np.delete(confmat_cleaned, unused_classes, axis=1)
    return confmat


def overall_accuracy(npcm):  # test/metrics.py:88-90
    oa = np.trace(npcm) / npcm.sum()   # the ratio first, then the percentage: the order fixes the last ulp
    return 100 * oa


def class_IoU(npcm):  # test/metrics.py:93-100
    ious = 100 * np.diag(npcm) / (np.sum(npcm, axis=1) + np.sum(npcm, axis=0) - np.diag(npcm))
    ious[np.isnan(ious)] = 0
    return ious, np.mean(ious)


def class_precision(npcm):  # test/metrics.py:103-106
    precision = 100 * np.diag(npcm) / np.sum(npcm, axis=0)
    precision[np.isnan(precision)] = 0
    return precision, np.mean(precision)


def class_recall(npcm):  # test/metrics.py:109-112
    recall = 100 * np.diag(npcm) / np.sum(npcm, axis=1)
    recall[np.isnan(recall)] = 0
    return recall, np.mean(recall)


def class_fscore(npcm):  # test/metrics.py:115-120
    precision = class_precision(npcm)[0]
    recall = class_recall(npcm)[0]
    fscore = 2 * (precision * recall) / (precision + recall)
    fscore[np.isnan(fscore)] = 0
    return fscore, np.mean(fscore)


def confusion_matrix_gpu(model: _native.Context, pred: torch.Tensor, truth_minus1: torch.Tensor, n_classes: int,
                         out: torch.Tensor | None = None) -> torch.Tensor:
    """sklearn.metrics.confusion_matrix(truth.flatten(), pred.flatten(), labels=range(n)) (test/metrics.py:161-163,
    229-231) as an int64 [n, n] device tensor; truth_minus1 already holds `band - 1` (utils.open_images)."""
    return model.confusion(pred.contiguous(), truth_minus1.contiguous(), n_classes, truth_sub=0, out=out)


def metrics_from_confmat(confmat: np.ndarray, config: dict, key: str) -> dict:
    """The per-key dictionary of compute_metrics_patch (test/metrics.py:165-192)."""
    classes = config["classes"]
    n_classes = len(classes)
    confmat_cleaned = clean_confmat(confmat, config)
    with np.errstate(divide="ignore", invalid="ignore"):
        per_c_ious, avg_ious = class_IoU(confmat_cleaned)
        ovr_acc = overall_accuracy(confmat_cleaned)
        per_c_fscore, avg_fscore = class_fscore(confmat_cleaned)
    return {key: {"Avg_metrics_name": ["mIoU", "Overall Accuracy", "Fscore"],
                  "Avg_metrics": [float(avg_ious), float(ovr_acc), float(avg_fscore)],
                  "classes": [classes[i][1] for i in range(1, n_classes + 1)],
                  "per_class_iou": [float(v) for v in per_c_ious],
                  "per_class_fscore": [float(v) for v in per_c_fscore]}}


def compute_metrics_patch(model: _native.Context, pred_map: torch.Tensor, truth_map: torch.Tensor, window, config: dict,
                          method: str) -> dict:
    """test/metrics.py:124-192 for one (col_off, row_off, width, height) window of device-resident maps."""
    col, row, w, h = window
    n_classes = len(config["classes"])
    cm = confusion_matrix_gpu(model, pred_map[row:row + h, col:col + w], truth_map[row:row + h, col:col + w], n_classes)
    return metrics_from_confmat(cm.cpu().numpy(), config, f"{method}_{col}_{row}")


#### BATCH MODE ####
def valid_truth(config: dict) -> Path:
    """test/metrics.py:32-45: the truth raster must sit under the same <dpt>/<zone> directories as the input."""
    truth_path = Path(config["truth_path"])
    sanity_check = str(config["input_img_path"]).split("/")[-3:-1]
    truth_check = list(truth_path.parts[-3:-1])
    if truth_check != sanity_check:
        raise ValueError(f"Ground truth path {truth_path} does not match input path {config['input_img_path']}")
    return Path(truth_path)


def get_truth_path(pred_path: Path, truth_dir: Path) -> Path:
    """test/metrics.py:48-60."""
    info = info_extract(pred_path)
    _, zone_name = info["dpt"], info["zone"]
    truth_subdir = truth_dir / zone_name
    truth_path = next(truth_subdir.glob("*.tif"), None)
    if truth_path is None:
        raise FileNotFoundError(f"Ground truth file not found in {truth_subdir}. Please check the folder.")
    return truth_path


def collect_paths_truth(config: dict, truth_dir: Path):
    """test/metrics.py:63-86: one row (pred_path, truth_path, method) per prediction raster under the time-stamped
    folders of output_path. The reference looks the truth up once per folder (from its first file); here it is
    looked up per file, which is the same whenever a folder holds one zone and still right when two zones
    finished within the same second and share a folder."""
    import pandas as pd
    path_collection = []
    pred_dir = Path(config["output_path"])
    for timestamp in sorted(p for p in pred_dir.iterdir() if p.is_dir()):
        for pred_path in timestamp.rglob("*.tif"):
            path_collection.append({"pred_path": str(pred_path), "truth_path": str(get_truth_path(pred_path, truth_dir)),
                                    "method": info_extract(pred_path)["method"]})
    return pd.DataFrame(path_collection)


def confmat_of_rasters(model: _native.Context, pred_path: str, truth_path: str, n_classes: int) -> np.ndarray:
    """confusion_matrix((truth band 1) - 1, pred band 1, labels=range(n)) (test/metrics.py:223-231) on the GPU."""
    preds = torch.from_numpy(np.ascontiguousarray(geotiff.read(Path(pred_path), bands=[1])[0])).to(model.device)
    target = torch.from_numpy(np.ascontiguousarray(geotiff.read(Path(truth_path), bands=[1])[0])).to(model.device)
    if preds.shape != target.shape:
        raise ValueError(f"prediction {tuple(preds.shape)} and truth {tuple(target.shape)} differ in size")
    return model.confusion(preds, target, n_classes, truth_sub=1).cpu().numpy()


def batch_metrics(config: dict, truth_dir: Path, model: _native.Context | None = None) -> list:
    """test/metrics.py:195-287: per method, the summed confusion matrix of every (prediction, truth) pair and the
    averaged metrics + parameters parsed back from the file name."""
    metrics_file = []
    df = collect_paths_truth(config, truth_dir)
    classes = config["classes"]
    n_classes = len(classes)
    own = model is None
    if own:
        model = _native.Context(int(__import__("os").environ.get("LOCAL_RANK", "0")))
    print("Computing metrics...")
    try:
        for method, group in df.groupby("method"):
            sum_confmat = np.zeros((n_classes, n_classes))
            for pred_path, truth_path in zip(group["pred_path"].tolist(), group["truth_path"].tolist()):
                try:
                    sum_confmat += confmat_of_rasters(model, pred_path, truth_path, n_classes)
                except Exception as e:  # the reference reports and carries on (test/metrics.py:232-233)
                    print(f"Error processing {pred_path} and {truth_path}: {e}")
            confmat_cleaned = clean_confmat(sum_confmat, config)
            with np.errstate(divide="ignore", invalid="ignore"):
                per_c_ious, avg_ious = class_IoU(confmat_cleaned)
                ovr_acc = overall_accuracy(confmat_cleaned)
                per_c_fscore, avg_fscore = class_fscore(confmat_cleaned)
                method_times = config.get("times", {}).get(method, [])
                avg_time = np.mean(method_times) if method_times else 0
            info = extract_method(str(method))
            metrics_file.append({
                "Method parameters": ["model name", "patch size", "stride", "margin", "padding", "stitching method"],
                "Parameters values": [config["model_name"], info["patch_size"], info["stride"], info["margin"],
                                      info["padding"], info["stitching"]],
                "Avg_metrics_name": ["mIoU", "Overall Accuracy", "Fscore", "Time in ms"],
                "Avg_metrics": [float(avg_ious), float(ovr_acc), float(avg_fscore), float(avg_time)],
                "classes": [classes[i][1] for i in range(1, n_classes + 1)],
                "per_class_iou": [float(v) for v in per_c_ious],
                "per_class_fscore": [float(v) for v in per_c_fscore],
            })
    finally:
        if own:
            model.close()
    return metrics_file
