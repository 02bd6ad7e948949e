"""Small run for compute-sanitizer (not a pytest module): the smoke zone (640 x 512, 256 px tiles, margin 32, 12 tiles)
through the exact-clipping loop with the fused class-map sink, the confusion kernel, and the `average_weights` /
`max` blended stitchings (fp32 atomics, u64 atomicMax); the conv kernels it launches include the CTA-pair forms
(cta_group::2, remote barrier arrives), the streamed-weight ring and the depth-to-space head.

    compute-sanitizer --tool memcheck  python tests/sanitize_run.py
    compute-sanitizer --tool racecheck python tests/sanitize_run.py
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import flair1_b200._native as nat  # noqa: E402
from flair1_b200.zone_detect.slicing_job import tile_table  # noqa: E402
from oracle import synth  # noqa: E402

W, H, T, margin = 640, 512, 256, 32
sd = synth.cached_checkpoint(3, 15)
raster = synth.synth_raster(3, H, W, seed=7)
ctx = nat.Context(0)
ctx.load_weights(sd, 3, 15)
ctx.set_norm("custom", synth.FLAIR_MEANS[:3], synth.FLAIR_STDS[:3])
ctx.set_raster(torch.from_numpy(raster).cuda(), [0, 1, 2], W, H)
tiles = tile_table(W, H, T, margin)
cls = torch.zeros((H, W), dtype=torch.uint8, device="cuda")
conf = torch.zeros((H, W), dtype=torch.uint8, device="cuda")
ctx.detect_strip(tiles, T, 8, cls, conf, W, 0)
truth = torch.from_numpy(synth.synth_mask(raster, 15, 3)).cuda()
cm = ctx.confusion(cls, truth, 15, truth_sub=1)
for method in ("average_weights", "max"):
    acc, wsum = ctx.blend_buffers(method, H, W)
    ctx.blend_strip(tiles, T, 8, method, acc, wsum, W, 0)
    bc = torch.zeros((H, W), dtype=torch.uint8, device="cuda")
    bf = torch.zeros((H, W), dtype=torch.uint8, device="cuda")
    ctx.blend_finalize(method, acc, wsum, bc, bf)
torch.cuda.synchronize()
print("sanitize_run OK:", int(cm.sum()), "px in the confusion matrix,", ctx.launch_count, "launches")
