"""Seeded synthetic rasters, masks and a briefly *trained* synthetic checkpoint.
TEST INFRASTRUCTURE (see oracle/__init__.py).

The shipped weights and the toy dataset are absent (.MISSING_LARGE_BLOBS:1-3), and random-init nets
are either degenerate (one class everywhere) or chaotic under bf16, so argmax
parity is measured on a checkpoint trained for a few hundred steps on a synthetic colour->class task.
Everything is seeded; the checkpoint is cached under tests/golden/_cache/ (git-ignored, travels to the
GPU box with the working tree; regenerated on demand when absent).
"""
from __future__ import annotations

import os
from pathlib import Path
from typing import Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .unet_smp033 import FlairModel, Unet

FLAIR_MEANS = [105.08, 110.87, 101.82, 106.38, 53.26]  # configs/flair-1-config.yaml:44
FLAIR_STDS = [52.17, 45.38, 44, 39.69, 79.3]           # configs/flair-1-config.yaml:45
CACHE_DIR = Path(__file__).resolve().parent.parent / "tests" / "golden" / "_cache"


def synth_raster(bands: int, height: int, width: int, seed: int) -> np.ndarray:
    """uint8 [bands, H, W]: low-frequency field (64x bicubic-upsampled N(0,1)) * 50 + 110 plus
    per-pixel N(0,1) * 10, clipped to 0..255 (SURVEY.md section 8d / Appendix E family)."""
    g = torch.Generator().manual_seed(seed)
    h64, w64 = (height + 63) // 64 + 1, (width + 63) // 64 + 1
    low = torch.randn((1, bands, h64, w64), generator=g)
    field = F.interpolate(low, size=(h64 * 64, w64 * 64), mode="bicubic", align_corners=False)[0, :, :height, :width]
    out = np.empty((bands, height, width), np.uint8)
    rows = max(1, (1 << 24) // max(width, 1))
    for r0 in range(0, height, rows):  # chunked so a 40000-wide strip never needs a float copy of the whole raster
        r1 = min(height, r0 + rows)
        noise = torch.randn((bands, r1 - r0, width), generator=g)
        out[:, r0:r1] = (field[:, r0:r1] * 50 + 110 + noise * 10).clamp_(0, 255).round_().to(torch.uint8).numpy()
    return out


def _labels_from(x_norm: torch.Tensor, P: torch.Tensor) -> torch.Tensor:
    """labels = argmax_k P[k] . blur_sigma3(x_norm): spatially coherent regions."""
    k = torch.arange(-9, 10, dtype=torch.float32)
    g1 = torch.exp(-0.5 * (k / 3.0) ** 2)
    g1 /= g1.sum()
    c = x_norm.shape[1]
    xb = F.conv2d(F.pad(x_norm, (9, 9, 0, 0), mode="reflect"), g1.view(1, 1, 1, -1).repeat(c, 1, 1, 1), groups=c)
    xb = F.conv2d(F.pad(xb, (0, 0, 9, 9), mode="reflect"), g1.view(1, 1, -1, 1).repeat(c, 1, 1, 1), groups=c)
    return torch.einsum("kc,bchw->bkhw", P, xb).argmax(1)


def class_projection(n_classes: int, bands: int, seed: int = 1234) -> torch.Tensor:
    return torch.randn((n_classes, bands), generator=torch.Generator().manual_seed(seed))


def synth_mask(raster: np.ndarray, n_classes: int, bands_used: int, seed: int = 1234) -> np.ndarray:
    """1-based uint8 label mask consistent with the training task (so mIoU is non-trivial)."""
    x = torch.from_numpy(raster[:bands_used].astype(np.float32))
    m = torch.tensor(FLAIR_MEANS[:bands_used]).view(-1, 1, 1)
    s = torch.tensor(FLAIR_STDS[:bands_used]).view(-1, 1, 1)
    lab = _labels_from(((x - m) / s)[None], class_projection(n_classes, bands_used, seed))[0]
    return (lab + 1).to(torch.uint8).numpy()


def train_synthetic_checkpoint(in_channels: int = 3, n_classes: int = 15, steps: int = 240, seed: int = 2022,
                               use_metadata: bool = False, verbose: bool = False, device: str = "cpu") -> dict:
    """Adam lr 2e-3, `steps` steps of 8 x 128^2 crops, BN in train mode (Appendix E recipe). The batches are
    drawn on the CPU from seeded generators; `device` only says where torch runs the training steps (a GPU box
    trains in seconds what takes the host cores minutes; the weights are synthetic either way)."""
    torch.manual_seed(seed)
    model = FlairModel(in_channels, n_classes, use_metadata) if use_metadata else Unet(in_channels, n_classes)
    model.to(device)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=2e-3)
    P = class_projection(n_classes, in_channels)
    m = torch.tensor(FLAIR_MEANS[:in_channels]).view(1, -1, 1, 1)
    s = torch.tensor(FLAIR_STDS[:in_channels]).view(1, -1, 1, 1)
    g = torch.Generator().manual_seed(seed + 1)
    size = 512 if use_metadata else 128  # the metadata branch is hard-wired to 512x512 (flair/model.py:59)
    bs = 2 if use_metadata else 8
    for it in range(steps):
        low = torch.randn((bs, in_channels, size // 32, size // 32), generator=g)
        img = F.interpolate(low, size=(size, size), mode="bicubic", align_corners=False) * 50 + 110
        img = (img + torch.randn(img.shape, generator=g) * 10).clamp(0, 255).round()
        x = (img - m) / s
        y = _labels_from(x, P).to(device)
        if use_metadata:
            met = torch.rand((bs, 45), generator=g)
            logits = model(x.to(device), met.to(device))
        else:
            logits = model(x.to(device))
        loss = F.cross_entropy(logits, y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        if verbose and it % 40 == 0:
            print(f"  synth-train step {it} loss {loss.item():.3f}", flush=True)
    model.eval()
    return {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}


def cached_checkpoint(in_channels: int = 3, n_classes: int = 15, use_metadata: bool = False, steps: int = 240) -> dict:
    """state_dict of the briefly trained synthetic checkpoint (bare smp keys, or FlairModel keys
    `seg_model.*` / `enc.enc_mlp.*` when use_metadata)."""
    CACHE_DIR.mkdir(parents=True, exist_ok=True)
    path = CACHE_DIR / f"synth_c{in_channels}_n{n_classes}_m{int(use_metadata)}_s{steps}.pth"
    if path.exists():
        return torch.load(path, map_location="cpu")
    if int(os.environ.get("RANK", "0")) != 0:
        # one process per GPU (torchrun): rank 0 trains and writes the cache, the others wait for the file
        import time
        deadline = time.time() + 1800
        while not path.exists():
            if time.time() > deadline:
                raise RuntimeError(f"{path} did not appear: rank 0 failed to build the synthetic checkpoint")
            time.sleep(0.5)
        return torch.load(path, map_location="cpu")
    nthreads = torch.get_num_threads()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    device = f"cuda:{os.environ.get('LOCAL_RANK', '0')}" if torch.cuda.is_available() else "cpu"
    try:
        sd = train_synthetic_checkpoint(in_channels, n_classes, steps=steps, use_metadata=use_metadata, device=device)
    finally:
        torch.set_num_threads(nthreads)
    tmp = path.with_suffix(".tmp")
    torch.save(sd, tmp)
    os.replace(tmp, path)
    return sd


def random_checkpoint(in_channels: int, n_classes: int, seed: int, use_metadata: bool = False) -> dict:
    """torch default inits, default BN statistics (configs 3 and 5 of BASELINE.json: logits tolerance only)."""
    torch.manual_seed(seed)
    model = FlairModel(in_channels, n_classes, use_metadata) if use_metadata else Unet(in_channels, n_classes)
    model.eval()
    return {k: v.detach().clone() for k, v in model.state_dict().items()}
