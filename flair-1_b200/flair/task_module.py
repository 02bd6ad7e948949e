"""Predict task (mirrors segmentation_task_predict of src/flair/task_module.py:174-213)."""
from __future__ import annotations

import numpy as np
import torch

from .model import FLAIR_ModelFactory


class segmentation_task_predict:
    """predict_step: logits = model(img, mtd); softmax; argmax (task_module.py:206-213). Batches are
    lists of uint8 patches [C, H, W]; the whole step is one fb_predict_patches call."""

    def __init__(self, model: FLAIR_ModelFactory, num_classes: int, use_metadata: bool = False, norm_type: str = "scaling",
                 means=(), stds=()):
        self.model = model
        self.num_classes = num_classes
        self.use_metadata = use_metadata
        model.seg_model.set_norm(norm_type, means, stds, channels=model.n_channels)

    def load_state_dict(self, state_dict, strict: bool = False):
        self.model.load_state_dict(state_dict, strict=strict)

    def predict_step(self, batch: dict, batch_idx: int = 0, dataloader_idx: int = 0) -> dict:
        imgs = np.stack(batch["img"]).astype(np.uint8, copy=False)
        n, c, h, w = imgs.shape
        if h != w or h % 32:
            raise RuntimeError(f"Wrong input shape height={h}, width={w}. Expected a square divisible by 32.")
        dev = self.model.seg_model.device
        patches = torch.from_numpy(np.ascontiguousarray(imgs)).to(dev)
        mtd = np.stack(batch["mtd"]).astype(np.float32) if self.use_metadata else None
        batch["preds"] = self.model.seg_model.predict_patches(patches, h, n, metadata=mtd)
        return batch
