"""Profiling driver (not a pytest module): one warm-up pass and one measured pass of the network on
`n` tiles, nothing else on the GPU, for `ncu --metrics gpu__time_duration.sum` launch lists:

    python tests/prof_forward.py 32            # plain run first (must exit 0), then the same under ncu
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import flair1_b200._native as nat  # noqa: E402
from oracle import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
T, W, H = 512, 4096, 1024
rng = np.random.default_rng(0)
raster = rng.integers(0, 256, size=(3, H, W), dtype=np.uint8)
ctx = nat.Context(0)
ctx.load_weights(synth.cached_checkpoint(3, 15), 3, 15)
ctx.set_norm("custom", synth.FLAIR_MEANS[:3], synth.FLAIR_STDS[:3])
ctx.upload_raster(raster, [0, 1, 2], W, H)
xy = np.array([[(i * 256) % (W - T), (i * 128) % (H - T)] for i in range(n)], np.int32)
for _ in range(passes):
    ctx.forward_tiles(xy, T)
ctx.synchronize()
print("launches", ctx.launch_count)
