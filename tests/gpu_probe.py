"""Diagnostic runner for the CUDA kernels (not a pytest module): runs each kernel against a plain
PyTorch fp32 reference on the same inputs and prints one line per case, flushing as it goes so that a
hang still leaves a readable log. Usage (on a GPU box):

    timeout 600 python tests/gpu_probe.py [case-substring ...] > gpurun_out/probe.log 2>&1
"""
from __future__ import annotations

import json
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import flair1_b200._native as nat  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

RESULTS = []


def log(*a):
    print(*a, flush=True)


def ref_conv(x1, w, bias, stride, pad, x2=None, up1=False, residual=None, rowbias=None, relu=False):
    """fp32 reference on bf16-rounded operands; NHWC in / NHWC out."""
    a = x1.float().permute(0, 3, 1, 2)
    if up1:
        a = F.interpolate(a, scale_factor=2, mode="nearest")
    if x2 is not None:
        a = torch.cat([a, x2.float().permute(0, 3, 1, 2)], dim=1)
    y = F.conv2d(a, w.to(torch.bfloat16).float(), bias=bias, stride=stride, padding=pad)
    if residual is not None:
        y = y + residual.float().permute(0, 3, 1, 2)
    if relu:
        y = torch.relu(y)
    if rowbias is not None:
        y = y + rowbias[:, None, :, None]
    return y.permute(0, 2, 3, 1).contiguous()


def conv_case(ctx, name, B, H, W, C1, Cout, k, stride, pad, mode, C2=0, up1=False, res=False, rowb=False,
              relu=True, out_f32=False, seed=0, identity=False, phase=False):
    """phase=True: the sub-pixel phase form (mode 2) of the same conv on [upsampled x1 (+) x2]."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    dev = ctx.device
    up1 = up1 or phase
    h1, w1 = (H // 2, W // 2) if up1 else (H, W)
    x1 = torch.randn((B, h1, w1, C1), generator=g).to(torch.bfloat16).to(dev)
    x2 = torch.randn((B, H, W, C2), generator=g).to(torch.bfloat16).to(dev) if C2 else None
    Cin = C1 + C2
    w = torch.randn((Cout, Cin, k, k), generator=g) / np.sqrt(Cin * k * k)
    if identity:  # centre tap identity: out[..., n] = x[..., n]
        w.zero_()
        for o in range(min(Cout, Cin)):
            w[o, o, k // 2, k // 2] = 1.0
    bias = (torch.randn((Cout,), generator=g) * 0.1).to(dev)
    Ho = (H + 2 * pad - k) // stride + 1
    Wo = (W + 2 * pad - k) // stride + 1
    residual = torch.randn((B, Ho, Wo, Cout), generator=g).to(torch.bfloat16).to(dev) if res else None
    rowbias = torch.randn((B, Ho), generator=g).to(dev) if rowb else None
    wp = (nat.pack_phase_weight(w, C1, C2) if phase else nat.pack_conv_weight(w, cin_pad=Cin, cout_pad=Cout)).to(dev)
    t0 = time.time()
    try:
        y = ctx.conv2d(x1, wp, bias, k, k, stride, pad, x2=x2, up1=up1, residual=residual, rowbias=rowbias,
                       relu=relu, out_f32=out_f32, mode=2 if phase else mode)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        log(f"CASE {name}: EXCEPTION {e}")
        RESULTS.append({"case": name, "ok": False, "error": str(e)})
        return False
    dt = time.time() - t0
    yr = ref_conv(x1, w.to(dev), bias, stride, pad, x2=x2, up1=up1, residual=residual, rowbias=rowbias, relu=relu)
    err = (y.float() - yr).abs()
    scale = yr.abs().max().item() + 1e-9
    tol = 2e-5 if out_f32 else 1.0 / 128  # bf16 output: half-ulp relative 2^-9, allow 2x
    if phase:
        # the phase form rounds SUMS of taps to bf16 while the reference sums bf16-rounded taps: the two differ by
        # weight-rounding noise (~2^-9 per weight), a few output ulps on small outputs; a layout error would be O(1)
        tol *= 2.5
    rel = (err / (yr.abs() + 0.05 * scale)).max().item()
    ok = bool(rel < tol * 2 + 1e-4) and bool(torch.isfinite(y.float()).all())
    bad = (err / (yr.abs() + 0.05 * scale) > tol * 2 + 1e-4)
    info = ""
    if not ok:
        idx = bad.nonzero()
        info = f" nbad={idx.shape[0]}/{bad.numel()} first_bad={idx[:4].tolist()} " \
               f"rows_bad={sorted(set((idx[:, 1] * Wo + idx[:, 2]).tolist()))[:16]} chans_bad={sorted(set(idx[:, 3].tolist()))[:16]}"
    log(f"CASE {name}: {'OK ' if ok else 'FAIL'} maxabs={err.max().item():.4e} rel={rel:.4e} refmax={scale:.3f} t={dt*1e3:.1f}ms{info}")
    RESULTS.append({"case": name, "ok": ok, "maxabs": err.max().item(), "rel": rel})
    return ok


def halo_case(ctx, name, B, H, W, C1, Cout, KH, stride, C2=0, res=False, relu=True, up2=False, out_f32=False, seed=0,
              identity=False, phase=False, d2s=0):
    """Halo-staged kernel against the fp32 reference on bf16-rounded operands.

    phase: the sub-pixel form; x1 is the LOW-res input, the reference is the conv of its nearest x2 upsample.
    d2s: the depth-to-space forms (1: 16 channels in, 4x4 stride-2 conv over 2x2 cells; 2: 32 channels in at low
    resolution, the reference is the conv of the nearest x2 upsample)."""
    phase = phase or d2s == 2
    g = torch.Generator(device="cpu").manual_seed(seed)
    dev = ctx.device
    x1 = torch.randn((B, H, W, C1), generator=g).to(torch.bfloat16).to(dev)
    x2 = torch.randn((B, H, W, C2), generator=g).to(torch.bfloat16).to(dev) if C2 else None
    Cin = C1 + C2
    w = torch.randn((Cout, Cin, KH, KH), generator=g) / np.sqrt(Cin * KH * KH)
    if identity:
        w.zero_()
        for o in range(min(Cout, Cin)):
            w[o, o, KH // 2, KH // 2] = 1.0
    w = w.to(torch.bfloat16).float()
    bias = (torch.randn((Cout,), generator=g) * 0.1).to(dev)
    pad = KH // 2
    Ho, Wo = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KH) // stride + 1
    residual = torch.randn((B, Ho, Wo, Cout), generator=g).to(torch.bfloat16).to(dev) if res else None
    try:
        y = ctx.conv2d_halo(x1, w, bias, KH, stride, x2=x2, residual=residual, relu=relu,
                            up2_out=2 + d2s if d2s else 2 if phase else up2, out_f32=out_f32)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        log(f"CASE {name}: EXCEPTION {e}")
        RESULTS.append({"case": name, "ok": False, "error": str(e)})
        return False
    xin = x1.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2) if phase else x1
    yr = ref_conv(xin, w.to(dev), bias, stride, pad, x2=x2, residual=residual, relu=relu)
    if up2:
        yr = yr.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    err = (y.float() - yr).abs()
    scale = yr.abs().max().item() + 1e-9
    tol = 2e-5 if out_f32 else 1.0 / 128
    if phase and not identity:
        tol *= 2.5  # collapsed taps are summed before the bf16 rounding of the weights (not after, as in the reference)
    relm = err / (yr.abs() + 0.05 * scale)
    rel = relm.max().item()
    ok = bool(rel < tol * 2 + 1e-4) and bool(torch.isfinite(y.float()).all())
    info = ""
    if not ok:
        idx = (relm > tol * 2 + 1e-4).nonzero()
        info = f" nbad={idx.shape[0]}/{relm.numel()} first_bad={idx[:6].tolist()} chans_bad={sorted(set(idx[:, 3].tolist()))[:16]}" \
               f" rows_bad={sorted(set(idx[:, 1].tolist()))[:20]} cols_bad={sorted(set(idx[:, 2].tolist()))[:20]}"
    log(f"CASE {name}: {'OK ' if ok else 'FAIL'} maxabs={err.max().item():.4e} rel={rel:.4e} refmax={scale:.3f}{info}")
    RESULTS.append({"case": name, "ok": ok, "maxabs": err.max().item(), "rel": rel})
    return ok


HALO_CASES = [
    # name, B, H, W, C1, Cout, KH, stride, kwargs
    ("h_ident_16_16", 1, 16, 32, 16, 16, 3, 1, dict(identity=True, relu=False)),
    ("h_3x3_16_16", 1, 16, 32, 16, 16, 3, 1, {}),
    ("h_3x3_16_16_multi", 2, 64, 64, 16, 16, 3, 1, {}),
    ("h_head_16_16_f32", 1, 32, 32, 16, 16, 3, 1, dict(relu=False, out_f32=True)),
    ("h_3x3_32_16", 2, 32, 32, 32, 16, 3, 1, {}),
    ("h_3x3_32_32_up2", 1, 32, 32, 32, 32, 3, 1, dict(up2=True)),
    ("h_ident_64_64", 1, 16, 16, 64, 64, 3, 1, dict(identity=True, relu=False)),
    ("h_3x3_64_64_res", 2, 32, 32, 64, 64, 3, 1, dict(res=True)),
    ("h_3x3_64p64_32", 2, 32, 32, 64, 32, 3, 1, dict(C2=64)),
    ("h_stem_7x7s2", 2, 64, 64, 8, 64, 7, 2, {}),
    ("h_stem_big", 3, 256, 128, 8, 64, 7, 2, {}),
    ("h_many_tiles", 4, 256, 256, 16, 16, 3, 1, {}),
    # sub-pixel phase form (dec4.conv1): low-res input H x W -> output 2H x 2W
    ("h_phase_ident_32_16", 1, 16, 8, 32, 16, 3, 1, dict(phase=True, identity=True, relu=False)),
    ("h_phase_32_16", 2, 32, 24, 32, 16, 3, 1, dict(phase=True)),
    ("h_phase_32_16_many", 3, 128, 128, 32, 16, 3, 1, dict(phase=True)),
]


# streamed filter bank (128 -> 128 channels: layer2, dec1.conv2): the halo kernel with the weights in a bulk-copy ring
SB_CASES = [
    ("sb_ident_128_128", 1, 16, 16, 128, 128, 3, 1, dict(identity=True, relu=False)),
    ("sb_128_128", 2, 32, 32, 128, 128, 3, 1, dict(seed=11)),
    ("sb_128_128_res", 2, 64, 64, 128, 128, 3, 1, dict(res=True, seed=12)),
    ("sb_128_128_many", 37, 64, 64, 128, 128, 3, 1, dict(res=True, seed=13)),
    # 48 rows = three 16-row tiles: not a whole number of CTA pairs, so these run as single CTAs
    ("sb_128_128_single", 2, 48, 32, 128, 128, 3, 1, dict(res=True, seed=14)),
    ("h_64_64_single", 3, 48, 64, 64, 64, 3, 1, dict(res=True, seed=15)),
    # CTA pairs, 64 channels (layer1), many tiles per cluster
    ("h_64_64_pair_many", 40, 128, 128, 64, 64, 3, 1, dict(res=True, seed=16)),
    ("h_64_64_pair_nores", 3, 64, 32, 64, 64, 3, 1, dict(seed=17)),
]


# depth-to-space forms of the 16-output-channel layers (HaloArgs::d2s): tiles of 16 x 16 cells = 32 x 32 pixels
D2S_CASES = [
    ("d2s_ident_16_16", 1, 32, 32, 16, 16, 3, 1, dict(d2s=1, identity=True, relu=False)),
    ("d2s_16_16", 1, 32, 32, 16, 16, 3, 1, dict(d2s=1)),
    ("d2s_16_16_multi", 3, 96, 64, 16, 16, 3, 1, dict(d2s=1, seed=3)),
    ("d2s_head_16_16_f32", 2, 64, 64, 16, 16, 3, 1, dict(d2s=1, relu=False, out_f32=True)),
    ("d2s_many_tiles", 5, 256, 256, 16, 16, 3, 1, dict(d2s=1, seed=5)),
    ("d2s_up_ident_32_16", 1, 16, 16, 32, 16, 3, 1, dict(d2s=2, identity=True, relu=False)),
    ("d2s_up_32_16", 2, 32, 48, 32, 16, 3, 1, dict(d2s=2, seed=7)),
    ("d2s_up_32_16_many", 3, 128, 128, 32, 16, 3, 1, dict(d2s=2, seed=9)),
]


TMA_EXTRA_CASES = [
    # strided TMA boxes (elementStrides = 2) and the sub-pixel phase form of decoder conv1
    ("ts_3x3s2_64_128", 2, 32, 32, 64, 128, 3, 2, 1, 1, {}),
    ("ts_3x3s2_256_512", 3, 32, 32, 256, 512, 3, 2, 1, 1, {}),
    ("ts_1x1s2_64_128", 2, 32, 64, 64, 128, 1, 2, 0, 1, dict(relu=False)),
    ("ts_1x1s2_128_256", 2, 32, 32, 128, 256, 1, 2, 0, 1, dict(relu=False)),
    ("ph_ident_64_64", 1, 16, 32, 64, 64, 3, 1, 1, 1, dict(phase=True, identity=True, relu=False)),
    ("ph_64_64", 1, 16, 32, 64, 64, 3, 1, 1, 1, dict(phase=True)),
    ("ph_128p64_64", 2, 32, 64, 128, 64, 3, 1, 1, 1, dict(phase=True, C2=64)),
    ("ph_512p256_256", 2, 32, 32, 512, 256, 3, 1, 1, 1, dict(phase=True, C2=256)),
    ("ph_256p128_128", 1, 64, 64, 256, 128, 3, 1, 1, 1, dict(phase=True, C2=128)),
    ("ph_64p64_32", 1, 64, 64, 64, 32, 3, 1, 1, 1, dict(phase=True, C2=64)),
]


def main():
    want = sys.argv[1:]
    log("device:", torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    ctx = nat.Context(0)

    def sel(name):
        return not want or any(s in name for s in want)

    G, T = 0, 1
    cases = [
        # name, B, H, W, C1, Cout, k, stride, pad, mode, kwargs
        ("g_ident_64_64_8x16", 1, 8, 16, 64, 64, 3, 1, 1, G, dict(identity=True, relu=False)),
        ("g_3x3_64_64_8x16", 1, 8, 16, 64, 64, 3, 1, 1, G, {}),
        ("t_ident_64_64_8x16", 1, 8, 16, 64, 64, 3, 1, 1, T, dict(identity=True, relu=False)),
        ("t_3x3_64_64_8x16", 1, 8, 16, 64, 64, 3, 1, 1, T, {}),
        ("g_3x3_64_64_32x32_b2", 2, 32, 32, 64, 64, 3, 1, 1, G, {}),
        ("t_3x3_64_64_32x32_b2", 2, 32, 32, 64, 64, 3, 1, 1, T, {}),
        ("g_3x3_128_128_res", 2, 16, 16, 128, 128, 3, 1, 1, G, dict(res=True)),
        ("t_3x3_128_128_res", 2, 16, 16, 128, 128, 3, 1, 1, T, dict(res=True)),
        ("t_3x3_256_256", 3, 32, 32, 256, 256, 3, 1, 1, T, {}),
        ("t_3x3_512_512_rowb", 4, 16, 16, 512, 512, 3, 1, 1, T, dict(res=True, rowb=True)),
        ("g_3x3_512_512", 2, 16, 16, 512, 512, 3, 1, 1, G, {}),
        ("g_7x7s2_8_64", 2, 64, 64, 8, 64, 7, 2, 3, G, {}),
        ("g_3x3s2_64_128", 2, 32, 32, 64, 128, 3, 2, 1, G, {}),
        ("g_1x1s2_64_128", 2, 32, 32, 64, 128, 1, 2, 0, G, dict(relu=False)),
        ("g_up_cat_512_256_256", 1, 32, 32, 512, 256, 3, 1, 1, G, dict(C2=256, up1=True)),
        ("g_up_cat_64_64_32", 1, 64, 64, 64, 32, 3, 1, 1, G, dict(C2=64, up1=True)),
        ("g_up_32_16", 1, 64, 64, 32, 16, 3, 1, 1, G, dict(up1=True)),
        ("g_3x3_16_16", 1, 64, 64, 16, 16, 3, 1, 1, G, {}),
        ("g_3x3_32_32", 1, 64, 64, 32, 32, 3, 1, 1, G, {}),
        ("g_head_16_16_f32", 1, 64, 64, 16, 16, 3, 1, 1, G, dict(relu=False, out_f32=True)),
        ("g_ragged_m", 1, 20, 20, 64, 64, 3, 1, 1, G, {}),
        ("t_many_tiles", 8, 128, 128, 64, 64, 3, 1, 1, T, {}),
        ("g_many_tiles", 8, 128, 128, 64, 64, 3, 1, 1, G, {}),
    ]
    for c in cases:
        name = c[0]
        if sel(name):
            conv_case(ctx, name, *c[1:10], **c[10])

    for c in HALO_CASES:
        if sel(c[0]):
            halo_case(ctx, c[0], *c[1:8], **c[8])
    if sel("t_dual"):
        conv_case(ctx, "t_dual_128p64_64", 2, 32, 32, 128, 64, 3, 1, 1, T, C2=64)
        conv_case(ctx, "g_dual_128p64_64", 2, 32, 32, 128, 64, 3, 1, 1, G, C2=64)
    for c in TMA_EXTRA_CASES:
        if sel(c[0]):
            conv_case(ctx, c[0], *c[1:10], **c[10])

    if sel("confusion"):
        g = torch.Generator().manual_seed(1)
        for npx in (1, 15, 16, 1000, 1 << 20, (1 << 22) + 7):
            pred = torch.randint(0, 21, (npx,), generator=g, dtype=torch.uint8)
            truth = torch.randint(0, 21, (npx,), generator=g, dtype=torch.uint8)
            cm = ctx.confusion(pred.cuda(), truth.cuda(), 19, truth_sub=1).cpu().numpy()
            t = (truth.numpy().astype(np.int64) - 1) & 0xFF
            p = pred.numpy().astype(np.int64)
            m = (t < 19) & (p < 19)
            ref = np.bincount(t[m] * 19 + p[m], minlength=361).reshape(19, 19)
            ok = bool((cm == ref).all())
            log(f"CASE confusion_{npx}: {'OK ' if ok else 'FAIL'} sum={cm.sum()} ref={ref.sum()}")
            RESULTS.append({"case": f"confusion_{npx}", "ok": ok})

    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "probe.json").write_text(json.dumps(RESULTS, indent=1))
    nfail = sum(1 for r in RESULTS if not r["ok"])
    log(f"DONE {len(RESULTS)} cases, {nfail} failed")
    return 1 if nfail else 0


if __name__ == "__main__":
    sys.exit(main())
