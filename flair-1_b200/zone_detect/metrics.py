"""Whole-raster and per-window metrics (mirrors src/zone_detect/test/metrics.py, a runtime module of the
reference): confusion matrix on the GPU (K9), ratios on the host in float64."""
from __future__ import annotations

import numpy as np
import torch

from .. import _native


def clean_confmat(confmat: np.ndarray, config: dict) -> np.ndarray:
    """test/metrics.py:18-29: drop rows/columns of classes whose weight is 0."""
    weights = np.array([class_info[0] for class_info in config["classes"].values()])
    unused_classes = np.where(weights == 0)[0]
    if unused_classes.size > 0:
        confmat_cleaned = np.delete(confmat, unused_classes, axis=0)
        return np.delete(confmat_cleaned, unused_classes, axis=1)
    return confmat


def overall_accuracy(npcm):  # test/metrics.py:88-90
    return 100 * np.trace(npcm) / npcm.sum()


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
