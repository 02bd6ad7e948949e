"""BASELINE.json configs[4] (not a pytest module): 19-class, 5-band patch throughput sweep, batch 1 .. 256 of
512 x 512 patches through fb_predict_patches (whole tiles: every pixel of a patch is an output), against the conv
tensor-core roofline. Random-init weights (torch defaults, manual_seed(5)), uniform uint8 patches, seed 5.

    python tests/bench_patch_sweep.py [max_batch] > gpurun_out/patch_sweep.json
"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import flair1_b200._native as nat  # noqa: E402
from oracle import synth  # noqa: E402
from oracle.unet_smp033 import Unet  # noqa: E402

max_b = int(sys.argv[1]) if len(sys.argv) > 1 else 256
T = 512
torch.manual_seed(5)
sd = Unet(5, 19).state_dict()
ctx = nat.Context(0)
ctx.load_weights(sd, 5, 19)
ctx.set_norm("custom", synth.FLAIR_MEANS, synth.FLAIR_STDS)
peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
peak = peaks.get("bf16_tflops_sustained", 1400.0)
g = torch.Generator(device="cuda").manual_seed(5)
rows = []
b = 1
while b <= max_b:
    patches = torch.randint(0, 256, (b, 5, T, T), generator=g, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        ctx.predict_patches(patches, T, b)
    torch.cuda.synchronize()
    reps = max(3, min(20, 512 // b))
    f0 = ctx.flop_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ctx.predict_patches(patches, T, b)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tflops = (ctx.flop_count - f0) / reps / (ms / 1e3) / 1e12
    rows.append({"batch": b, "ms": ms, "patches_per_s": b / (ms / 1e3), "Mpixels_per_s": b * T * T / 1e6 / (ms / 1e3),
                 "conv_tflops_whole_call": tflops, "frac_of_peak": tflops / peak})
    print(f"batch {b:4d}: {ms:9.3f} ms  {rows[-1]['patches_per_s']:9.1f} patches/s  {tflops:7.1f} TFLOP/s ({100 * tflops / peak:5.1f} % of {peak})",
          file=sys.stderr, flush=True)
    del patches
    b *= 2
print(json.dumps({"workload": "5-band 19-class 512x512 patch predict, random init", "peak_tflops": peak,
                  "gflop_per_patch": (ctx.flop_count - f0) / reps / (b // 2) / 1e9, "rows": rows}))
ctx.close()
