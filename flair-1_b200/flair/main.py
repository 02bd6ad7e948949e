"""`flair --conf x.yaml` on B200: the predict + metrics stages (mirrors src/flair/main.py).

Training (tasks.train) is out of scope: `tasks.train: True` is rejected. The predict loop
of the reference is a Lightning Trainer with batch_size 1 (data_module.py:97-104); here the patches are decoded
by a pool of host threads (the TIFF codecs release the GIL), predicted in groups by one fb_predict_patches call
each while the next group is being read, and written (LZW) by the same pool while the GPU works on the next
group. Under torchrun the test CSV is sharded round-robin across ranks; every rank writes its own PRED_* files
and rank 0 computes the metrics from the files once all ranks are done (the reference's `metrics` stage also
reads the predictions back from disk, src/flair/metrics.py:60-74).
"""
from __future__ import annotations

import argparse
import datetime
import os
import shutil
import sys
import time
from collections import deque
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import torch

from .data_loader import predict_dataset
from .metrics import metrics
from .model import FLAIR_ModelFactory, load_checkpoint
from .task_module import segmentation_task_predict
from .tasks_utils import gather_paths
from .utils import print_recap, read_config
from .writer import predictionwriter

argParser = argparse.ArgumentParser()
argParser.add_argument("--conf", help="Path to the .yaml config file", required=True)


def setup_environment(args):
    """src/flair/main.py:25-34."""
    config = read_config(args.conf)
    out_dir = Path(config["paths"]["out_folder"], config["paths"]["out_model_name"])
    out_dir.mkdir(parents=True, exist_ok=True)
    return config, out_dir


class Logger(object):
    """src/flair/main.py:36-48."""

    def __init__(self, filename="Default.log"):
        self.terminal = sys.stdout
        self.log = open(filename, "w", encoding="utf-8") if int(os.environ.get("RANK", "0")) == 0 else None
        self.encoding = self.terminal.encoding

    def write(self, message):
        self.terminal.write(message)
        if self.log:
            self.log.write(message)

    def flush(self):
        if self.log:
            self.log.flush()


def get_datasets(config):
    """src/flair/main.py:50-60 (predict split only)."""
    dict_test = gather_paths(config, split="test") if config["tasks"]["predict"] else None
    return None, None, dict_test


def copy_csv_and_config(config, out_dir, args):
    """src/flair/main.py:62-73."""
    csv_copy_dir = Path(out_dir, "used_csv_and_config")
    csv_copy_dir.mkdir(parents=True, exist_ok=True)
    if config["tasks"]["predict"]:
        shutil.copy(config["paths"]["test_csv"], csv_copy_dir)
    shutil.copy(args.conf, csv_copy_dir)


def get_segmentation_module(config, stage="predict", device=0):
    """src/flair/tasks_utils.py:65-122, predict stage."""
    if stage != "predict":
        raise NotImplementedError("only the predict stage is built for B200")
    model = FLAIR_ModelFactory(config, device)
    return segmentation_task_predict(model=model, num_classes=len(config["classes"]), use_metadata=config["use_metadata"],
                                     norm_type=config["norm_type"], means=config.get("norm_means", []),
                                     stds=config.get("norm_stds", []))


def predict(config, dict_test, seg_module, out_dir_predict, rank=0, world=1):
    """src/flair/tasks.py:113-142: loop over the test patches, write PRED_* files. Three stages overlap: host threads
    read + decode the patches of group g + 1 and encode + write the predictions of group g - 1 while the GPU predicts
    group g. A group is `batch_size` patches, raised to `patches_per_launch` (default 148, one 512^2 patch per SM) --
    the result of a patch does not depend on what it is batched with. Returns (patches, seconds)."""
    ds = predict_dataset(dict_files=dict_test, channels=config["channels"], num_classes=len(config["classes"]),
                         use_metadata=config["use_metadata"], norm_type=config["norm_type"],
                         means=config.get("norm_means", []), stds=config.get("norm_stds", []))
    writer = predictionwriter(config, out_dir_predict.as_posix(), write_interval="batch")
    group = max(1, int(config.get("batch_size", 1)), int(config.get("patches_per_launch", 148)))
    idx = list(range(rank, len(ds), world))
    groups = [idx[s:s + group] for s in range(0, len(idx), group)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=max(2, min(16, os.cpu_count() or 2))) as pool:
        reads = deque()

        def start_read(g):
            reads.append([pool.submit(ds.__getitem__, i) for i in groups[g]])

        writes = []
        for g in range(min(2, len(groups))):
            start_read(g)
        for g in range(len(groups)):
            items = [f.result() for f in reads.popleft()]
            if g + 2 < len(groups):
                start_read(g + 2)
            batch = {"img": [it["img"] for it in items], "id": [it["id"] for it in items]}
            if config["use_metadata"]:
                batch["mtd"] = [it["mtd"] for it in items]
            out = seg_module.predict_step(batch, g)
            writes.extend(writer.write_async(out, pool))
        for f in writes:
            f.result()
    return len(idx), time.perf_counter() - t0


def predict_stage(config, dict_test, out_dir_predict, device=0, rank=0, world=1):
    """src/flair/main.py:187-203."""
    seg_module = get_segmentation_module(config, stage="predict", device=device)
    if config["tasks"]["train"]:
        raise NotImplementedError("tasks.train is out of scope of the B200 hot path; train with the reference and predict here")
    load_checkpoint(config, seg_module.model)   # every rank loads (the reference's rank_zero_only here is a bug)
    if not seg_module.model.loaded:
        raise SystemExit("no usable checkpoint: refusing to predict with uninitialised weights")
    n, seconds = predict(config, dict_test, seg_module, out_dir_predict, rank, world)
    if rank == 0:
        print(f"    [x] predicted {n} patches in {seconds:.2f} s ({n / max(seconds, 1e-9):.1f} patches/s incl. read + write"
              + (f", rank 0 of {world}" if world > 1 else "") + ")")
    return seg_module


def main():
    """src/flair/main.py:206-243."""
    args = argParser.parse_args()
    config, out_dir = setup_environment(args)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("flair1_b200 runs on a B200 (sm_100a) only: no CUDA device is visible and there is no CPU fallback")
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    old = sys.stdout
    sys.stdout = Logger(Path(config["paths"]["out_folder"], config["paths"]["out_model_name"], "flair-compute.log").as_posix())
    try:
        print(datetime.datetime.now().strftime("Starting : %Y-%m-%d  %H:%M") + "\n")
        dict_train, dict_val, dict_test = get_datasets(config)
        if rank == 0:
            print_recap(config, dict_train, dict_val, dict_test)
            if config["cp_csv_and_conf_to_output"]:
                copy_csv_and_config(config, out_dir, args)
        if config["tasks"]["predict"]:
            out_dir_predict = Path(out_dir, "predictions_" + config["paths"]["out_model_name"])
            out_dir_predict.mkdir(parents=True, exist_ok=True)
            seg_module = predict_stage(config, dict_test, out_dir_predict, device=local, rank=rank, world=world)
            if world > 1:
                import torch.distributed as dist
                dist.barrier()
            if config["tasks"]["metrics"] and rank == 0:
                metrics(config, out_dir_predict, remove_preds=config["tasks"]["delete_preds"], context=seg_module.model.seg_model)
    finally:
        sys.stdout = old


if __name__ == "__main__":
    main()
