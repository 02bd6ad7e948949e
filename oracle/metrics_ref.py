"""CPU restatement of the metric code on the path. TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows src/flair/metrics.py:10-40,60-88 and src/zone_detect/test/metrics.py:18-29,88-120,146-171.
sklearn is present in this image, so the confusion matrix is the reference's own library call.
"""
from __future__ import annotations

import numpy as np
from sklearn.metrics import confusion_matrix


def patch_confusion(target_minus1: np.ndarray, pred: np.ndarray, n_classes: int) -> np.ndarray:
    """confusion_matrix(target.flatten(), preds.flatten(), labels=range(n)) -- flair/metrics.py:67-71,
    zone_detect/test/metrics.py:161-163. int64, rows = truth, out-of-range pairs dropped."""
    return confusion_matrix(target_minus1.flatten(), pred.flatten(), labels=list(range(n_classes)))


def confusion_numpy(truth_raw: np.ndarray, pred: np.ndarray, n_classes: int, truth_sub: int = 1) -> np.ndarray:
    """Same histogram without sklearn (never raises on an all-out-of-range patch): uint8 wrap of
    `mask - 1` as at flair/metrics.py:62-64 / zone_detect/utils.py:288."""
    t = (truth_raw.astype(np.uint8).ravel() - np.uint8(truth_sub)).astype(np.int64)
    p = pred.astype(np.uint8).ravel().astype(np.int64)
    m = (t < n_classes) & (p < n_classes)
    return np.bincount(t[m] * n_classes + p[m], minlength=n_classes * n_classes).reshape(n_classes, n_classes)


def clean_confmat(confmat: np.ndarray, classes: dict) -> np.ndarray:
    """zone_detect/test/metrics.py:18-29; flair/metrics.py:77-82 (drop rows/cols of weight-0 classes)."""
    weights = np.array([info[0] for info in classes.values()])
    unused = np.where(weights == 0)[0]
    if unused.size > 0:
        return np.delete(np.delete(confmat, unused, axis=0), unused, axis=1)
    return confmat


def overall_accuracy(npcm):  # flair/metrics.py:10-12
    oa = np.trace(npcm) / npcm.sum()   # the ratio first, then the percentage: the order fixes the last ulp
    return 100 * oa


def class_IoU(npcm):  # flair/metrics.py:15-22
    with np.errstate(divide="ignore", invalid="ignore"):
        ious = 100 * np.diag(npcm) / (np.sum(npcm, axis=1) + np.sum(npcm, axis=0) - np.diag(npcm))
    ious[np.isnan(ious)] = 0
    return ious, np.mean(ious)


def class_precision(npcm):  # flair/metrics.py:25-28
    with np.errstate(divide="ignore", invalid="ignore"):
        precision = 100 * np.diag(npcm) / np.sum(npcm, axis=0)
    precision[np.isnan(precision)] = 0
    return precision, np.mean(precision)


def class_recall(npcm):  # flair/metrics.py:31-34
    with np.errstate(divide="ignore", invalid="ignore"):
        recall = 100 * np.diag(npcm) / np.sum(npcm, axis=1)
    recall[np.isnan(recall)] = 0
    return recall, np.mean(recall)


def class_fscore(precision, recall):  # flair/metrics.py:37-40
    with np.errstate(divide="ignore", invalid="ignore"):
        fscore = 2 * (precision * recall) / (precision + recall)
    fscore[np.isnan(fscore)] = 0
    return fscore, np.mean(fscore)


def flair_metrics(sum_confmat: np.ndarray, classes: dict) -> dict:
    """The dictionary written to metrics.json (flair/metrics.py:75-108)."""
    weights = np.array([classes[i][0] for i in classes])
    cleaned = clean_confmat(sum_confmat, classes) if (weights == 0).any() else sum_confmat
    per_c_ious, avg_ious = class_IoU(cleaned)
    ovr_acc = overall_accuracy(cleaned)
    per_c_precision, avg_precision = class_precision(cleaned)
    per_c_recall, avg_recall = class_recall(cleaned)
    per_c_fscore, avg_fscore = class_fscore(per_c_precision, per_c_recall)
    return {
        "Avg_metrics_name": ["mIoU", "Overall Accuracy", "Fscore", "Precision", "Recall"],
        "Avg_metrics": [avg_ious, ovr_acc, avg_fscore, avg_precision, avg_recall],
        "classes": list(np.array([classes[i][1] for i in classes])[np.nonzero(weights)[0]]),
        "per_class_iou": list(per_c_ious),
        "per_class_fscore": list(per_c_fscore),
        "per_class_precision": list(per_c_precision),
        "per_class_recall": list(per_c_recall),
    }
