"""Profiling driver (not a pytest module): the exact-clipping zone loop (fb_detect_strip: active-tile lists,
fused class-map sink) on the first `n` tiles of a 1024 x 10000 zone, `passes` times, nothing else on the GPU.
For `ncu --metrics gpu__time_duration.sum` launch lists and `ncu --set full` captures:

    python tests/prof_zone.py 74 2            # plain run first (must exit 0), then the same under ncu
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import flair1_b200._native as nat  # noqa: E402
from flair1_b200.zone_detect.slicing_job import tile_table  # noqa: E402
from oracle import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 74
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
T, margin, W, H = 512, 128, 1024, 10000
raster = torch.from_numpy(synth.synth_raster(3, H, W, seed=1)).cuda()
ctx = nat.Context(0)
ctx.load_weights(synth.cached_checkpoint(3, 15), 3, 15)
ctx.set_norm("custom", synth.FLAIR_MEANS[:3], synth.FLAIR_STDS[:3])
ctx.set_raster(raster, [0, 1, 2], W, H)
tiles = tile_table(W, H, T, margin)[:n]
cls = torch.zeros((H, W), dtype=torch.uint8, device="cuda")
conf = torch.zeros((H, W), dtype=torch.uint8, device="cuda")
for _ in range(passes):
    l0, f0 = ctx.launch_count, ctx.flop_count
    ctx.detect_strip(tiles, T, n, cls, conf, W, 0)
    ctx.synchronize()
print("launches per pass", ctx.launch_count - l0, "GFLOP per tile", (ctx.flop_count - f0) / n / 1e9)
