"""Event-timed micro-benchmark of single conv launches (not a pytest module).

    python tests/bench_conv_shapes.py            # TMA producer, single-CTA MMAs
    FB_PAIR=1 python tests/bench_conv_shapes.py  # same shapes through the cta_group::2 kernel
"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import flair1_b200._native as nat  # noqa: E402

ctx = nat.Context(0)
g = torch.Generator().manual_seed(0)
for (B, H, W, C, N) in ((37, 128, 128, 64, 64), (37, 64, 64, 128, 128), (37, 32, 32, 256, 256), (74, 32, 32, 256, 256)):
    x = torch.randn((B, H, W, C), generator=g).to(torch.bfloat16).cuda()
    w = torch.randn((N, C, 3, 3), generator=g) / np.sqrt(C * 9)
    wp = nat.pack_conv_weight(w, cin_pad=C, cout_pad=N).cuda()
    bias = torch.zeros(N).cuda()
    for res in (None, torch.randn((B, H, W, N), generator=g).to(torch.bfloat16).cuda()):
        for _ in range(3):
            ctx.conv2d(x, wp, bias, 3, 3, 1, 1, relu=True, mode=1, residual=res)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ctx.conv2d(x, wp, bias, 3, 3, 1, 1, relu=True, mode=1, residual=res)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        fl = 2 * B * H * W * N * C * 9
        print(f"B={B} C={C} N={N} {H}x{W}{' +res' if res is not None else '     '}: {ms * 1e3:.1f} us  {fl / ms / 1e9:.0f} TFLOP/s  "
              f"({ms * 1e3 / B:.2f} us/tile)", flush=True)
