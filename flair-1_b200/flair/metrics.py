"""Patch-level metrics (mirrors src/flair/metrics.py): per-patch confusion matrix of (mask - 1, PRED)
summed over the test CSV, weight-0 classes dropped, IoU / OA / F-score / precision / recall written to
metrics/metrics.json and the un-cleaned matrix to metrics/confmat.npy. The histogram itself runs on the
GPU (K9) when a context is passed, in one accumulating matrix."""
from __future__ import annotations

import json
import shutil
from pathlib import Path

import numpy as np
import pandas as pd
import torch

from .. import geotiff


def overall_accuracy(npcm):
    oa = np.trace(npcm) / npcm.sum()   # the ratio first, then the percentage: the order fixes the last ulp
    return 100 * oa


def class_IoU(npcm, n_class=None):
    ious = 100 * np.diag(npcm) / (np.sum(npcm, axis=1) + np.sum(npcm, axis=0) - np.diag(npcm))
    ious[np.isnan(ious)] = 0
    return ious, np.mean(ious)


def class_precision(npcm):
    precision = 100 * np.diag(npcm) / np.sum(npcm, axis=0)
    precision[np.isnan(precision)] = 0
    return precision, np.mean(precision)


def class_recall(npcm):
    recall = 100 * np.diag(npcm) / np.sum(npcm, axis=1)
    recall[np.isnan(recall)] = 0
    return recall, np.mean(recall)


def class_fscore(precision, recall):
    fscore = 2 * (precision * recall) / (precision + recall)
    fscore[np.isnan(fscore)] = 0
    return fscore, np.mean(fscore)


def _report(out: dict, classes: dict) -> None:
    """Console report of metrics() (same text as the reference's, src/flair/metrics.py:118-160): the five averages,
    one line per scored class, then the classes with weight 0."""
    rule = "-" * 90
    lines = ["", "Global Metrics: ", rule]
    lines += [f"{name:<20s} {value:<20.4f}" for name, value in zip(out["Avg_metrics_name"], out["Avg_metrics"])]
    lines += [rule + "\n\n", "{:<25} {:<15} {:<10} {:<10} {:<10} {:<10}".format("Class", "Weight", "IoU", "F-score", "Precision", "Recall"),
              "-" * 65]
    per_class = {name: i for i, name in enumerate(out["classes"])}
    cols = ("per_class_iou", "per_class_fscore", "per_class_precision", "per_class_recall")
    scored = [(w, n) for w, n in classes.values() if w != 0]
    lines += ["{:<25} {:<15} ".format(n, w) + " ".join("{:<10.4f}".format(out[c][per_class[n]]) for c in cols) for w, n in scored]
    lines += ["\nNot learned Classes:"] + ["{:<25} {:<15}".format(n, w) for w, n in classes.values() if w == 0] + ["\n\n"]
    print("\n".join(lines))


def metrics(config: dict, path_preds, remove_preds: bool = False, context=None) -> dict:
    """src/flair/metrics.py:43-164. `context`: a libflairb200 context (required: there is no CPU path)."""
    if context is None:
        raise RuntimeError("metrics() needs the libflairb200 context that produced the predictions")
    path_preds = Path(path_preds)
    gt_csv = pd.read_csv(config["paths"]["test_csv"], header=None)
    truth_images = gt_csv.iloc[:, 0].to_list()
    truth_msks = gt_csv.iloc[:, 1].to_list()
    preds_msks = [Path(path_preds.as_posix(), "PRED_" + i.split("/")[-1]).as_posix() for i in truth_images]
    assert len(truth_msks) == len(preds_msks), "[WARNING !] mismatch number of predictions and test files."
    print("-- Calculating metrics --")
    n_classes = int(len(config["classes"]))
    cm_dev = torch.zeros((n_classes, n_classes), dtype=torch.int64, device=context.device)
    for u in range(len(truth_msks)):
        try:
            target = geotiff.read(truth_msks[u], bands=[1])[0]     # raw mask; "- 1" with uint8 wrap happens in the kernel
            preds = geotiff.read(preds_msks[u], bands=[1])[0]
            if target.shape != preds.shape:
                raise ValueError(f"shape mismatch {target.shape} vs {preds.shape}")
            context.confusion(torch.from_numpy(preds).to(context.device), torch.from_numpy(target).to(context.device),
                              n_classes, truth_sub=1, out=cm_dev)
        except Exception as e:  # noqa: BLE001  (metrics.py:73-74)
            print(f"Error at index {u}: {e}")
    sum_confmat = cm_dev.cpu().numpy()
    weights = np.array([config["classes"][i][0] for i in config["classes"]])
    unused_classes = np.where(weights == 0)[0]
    confmat_cleaned = np.delete(np.delete(sum_confmat, unused_classes, axis=0), unused_classes, axis=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        per_c_ious, avg_ious = class_IoU(confmat_cleaned, len(np.nonzero(weights)[0]))
        ovr_acc = overall_accuracy(confmat_cleaned)
        per_c_precision, avg_precison = class_precision(confmat_cleaned)
        per_c_recall, avg_recall = class_recall(confmat_cleaned)
        per_c_fscore, avg_fscore = class_fscore(per_c_precision, per_c_recall)
    out = {
        "Avg_metrics_name": ["mIoU", "Overall Accuracy", "Fscore", "Precision", "Recall"],
        "Avg_metrics": [float(avg_ious), float(ovr_acc), float(avg_fscore), float(avg_precison), float(avg_recall)],
        "classes": list(np.array([config["classes"][i][1] for i in config["classes"]])[np.nonzero(weights)[0]]),
        "per_class_iou": [float(v) for v in per_c_ious],
        "per_class_fscore": [float(v) for v in per_c_fscore],
        "per_class_precision": [float(v) for v in per_c_precision],
        "per_class_recall": [float(v) for v in per_c_recall],
    }
    out_folder_metrics = Path("/".join(path_preds.as_posix().split("/")[:-1]), "metrics")
    out_folder_metrics.mkdir(exist_ok=True, parents=True)
    np.save(out_folder_metrics.as_posix() + "/confmat.npy", sum_confmat)
    json.dump(out, open(out_folder_metrics / Path("metrics.json"), "w"))
    _report(out, config["classes"])
    if remove_preds:
        shutil.rmtree(path_preds)
    return out
