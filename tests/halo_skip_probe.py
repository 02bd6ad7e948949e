"""Bottleneck hunt for the halo-staged conv kernel (not a pytest module): run each layer shape with parts of
the kernel switched off (FB_HALO_SKIP bit mask: 1 producers, 2 MMAs, 4 epilogue stores, 8 whole epilogue) and
read the kernel durations from an ncu launch list:

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tests/halo_skip_probe.py

The script prints the order of the launches (shape, mask) so that the list can be matched up.
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import flair1_b200._native as nat  # noqa: E402

BATCHES = [int(b) for b in (sys.argv[1] if len(sys.argv) > 1 else "37").split(",")]
SHAPES = [
    # name, H, W, C1, C2, Cout, KH, stride, kwargs
    ("layer1 64->64 @128", 128, 128, 64, 0, 64, 3, 1, {}),
    ("layer1 64->64 @128 +res", 128, 128, 64, 0, 64, 3, 1, dict(res=True)),
    ("dec3.c1 64+64->32 @256", 256, 256, 64, 64, 32, 3, 1, {}),
    ("dec3.c2 32->32 @256", 256, 256, 32, 0, 32, 3, 1, {}),
    ("dec4.c1 phase 32->16 @256->512", 256, 256, 32, 0, 16, 3, 1, dict(phase=True)),
    ("dec4.c2 16->16 @512", 512, 512, 16, 0, 16, 3, 1, {}),
    ("head 16->16 f32 @512", 512, 512, 16, 0, 16, 3, 1, dict(out_f32=True, relu=False)),
    ("stem 8->64 s2 @512", 512, 512, 8, 0, 64, 7, 2, {}),
]
MASKS = [int(m) for m in os.environ.get("SKIP_MASKS", "0,1,2,4,8,3,9,10").split(",")]  # bit 16: no filter-bank load

ctx = nat.Context(0)
g = torch.Generator().manual_seed(0)
order = []
for (name, H, W, C1, C2, Cout, KH, stride, kw), B in [(s, b) for b in BATCHES for s in SHAPES]:
    name = f"B{B} {name}"
    x1 = torch.randn((B, H, W, C1), generator=g).to(torch.bfloat16).cuda()
    x2 = torch.randn((B, H, W, C2), generator=g).to(torch.bfloat16).cuda() if C2 else None
    w = torch.randn((Cout, C1 + C2, KH, KH), generator=g) / np.sqrt((C1 + C2) * KH * KH)
    bias = torch.zeros(Cout).cuda()
    pad = KH // 2
    Ho, Wo = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KH) // stride + 1
    res = torch.randn((B, Ho, Wo, Cout), generator=g).to(torch.bfloat16).cuda() if kw.get("res") else None
    up2 = 2 if kw.get("phase") else 0
    for m in MASKS:
        # masks >= 1000 select the direct-store epilogue (FB_DIRECT_STORE=1) with skip mask m - 1000
        os.environ["FB_HALO_SKIP"] = str(m % 1000)
        os.environ["FB_DIRECT_STORE"] = "1" if m >= 1000 else "0"
        for _ in range(2):  # the second launch of each pair is the one to read
            ctx.conv2d_halo(x1, w, bias, KH, stride, x2=x2, residual=res, relu=kw.get("relu", True), up2_out=up2,
                            out_f32=kw.get("out_f32", False))
            order.append((name, m))
    torch.cuda.synchronize()
    del x1, x2, res
os.environ["FB_HALO_SKIP"] = "0"
for i, (n, m) in enumerate(order):
    print(f"LAUNCH {i} | {n} | mask {m}")
