"""Diagnostic runner (not a pytest module): whole-network parity of the CUDA path against the CPU
oracle with per-layer error report, then a small zone. Usage on a GPU box:

    timeout 900 python tests/gpu_probe_net.py > gpurun_out/probe_net.log 2>&1
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import flair1_b200._native as nat  # noqa: E402
from flair1_b200.zone_detect.slicing_job import tile_table  # noqa: E402
from oracle import synth  # noqa: E402
from oracle.unet_smp033 import Unet, layer_activations  # noqa: E402
from oracle.zone_detect_ref import GeoRaster, run_zone  # noqa: E402


def log(*a):
    print(*a, flush=True)


def main():
    torch.set_num_threads(max(1, torch.get_num_threads()))
    log("device:", torch.cuda.get_device_name(0))
    sd = synth.cached_checkpoint(3, 15)
    model = Unet(3, 15)
    model.load_state_dict(sd, strict=True)
    model.eval()

    W, H, T = 1000, 700, 512
    raster = synth.synth_raster(3, H, W, seed=1)
    means, stds = synth.FLAIR_MEANS[:3], synth.FLAIR_STDS[:3]

    ctx = nat.Context(0)
    ctx.load_weights(sd, 3, 15)
    ctx.set_norm("custom", means, stds)
    ctx.set_raster(torch.from_numpy(raster).cuda(), [0, 1, 2], W, H)

    # ---- per-layer parity on 2 tiles (one fully inside, one hanging over the top-left corner)
    xy = np.array([[200, 100], [-128, -128]], dtype=np.int32)
    t0 = time.time()
    logits = ctx.forward_tiles(xy, T)
    torch.cuda.synchronize()
    log(f"forward_tiles ok in {time.time() - t0:.3f}s, launches={ctx.launch_count}")

    imgs = []
    for x0, y0 in xy:
        patch = np.zeros((3, T, T), np.uint8)
        r0, r1, c0, c1 = max(y0, 0), min(y0 + T, H), max(x0, 0), min(x0 + T, W)
        patch[:, r0 - y0:r1 - y0, c0 - x0:c1 - x0] = raster[:, r0:r1, c0:c1]
        img = patch.astype(np.float64)
        for i in range(3):
            img[i] = (img[i] - means[i]) / stds[i]
        imgs.append(torch.as_tensor(img, dtype=torch.float))
    x = torch.stack(imgs)
    acts = layer_activations(model, x)
    x0 = ctx.debug_input_tiles().float().cpu().permute(0, 3, 1, 2)
    log(f"LAYER {'x0':14s} maxabs={(x0 - x).abs().max().item():.4e} (vs fp32 normalised input; bf16 rounding expected) "
        f"exact_vs_bf16={(x0 == x.to(torch.bfloat16).float()).all().item()}")
    names = ["f1", "pool"] + [f"layer{l}.{b}.out" for l, n in ((1, 3), (2, 4), (3, 6), (4, 3)) for b in range(n)] + \
            [f"dec{i}" for i in range(5)] + ["logits"]
    for nme in names:
        ref = acts[nme]
        got = ctx.debug_activation(nme).float().cpu().permute(0, 3, 1, 2)[:, :ref.shape[1]]
        if got.shape[-1] == 2 * ref.shape[-1]:
            ref = ref.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
        err = (got - ref).abs().max().item()
        log(f"LAYER {nme:14s} maxabs={err:.4e} refmax={ref.abs().max().item():.3f} rel={err / (ref.abs().max().item() + 1e-9):.4e}")
    ref = acts["logits"]
    got = logits.cpu().permute(0, 3, 1, 2)[:, :15]
    agree = (got.argmax(1) == ref.argmax(1)).float().mean().item()
    log(f"LOGITS max-abs/max|ref| = {(got - ref).abs().max().item() / ref.abs().max().item():.4e}  argmax agreement = {agree * 100:.4f}%")
    log(f"pad-class logits all zero: {(logits[..., 15:] == 0).all().item()}")

    # ---- small zone vs oracle
    config = {"img_pixels_detection": T, "margin": 128, "channels": [1, 2, 3], "n_classes": 15,
              "norma_task": [{"norm_type": "custom", "norm_means": means, "norm_stds": stds}]}
    tiles = tile_table(W, H, T, 128)
    cls = torch.zeros((H, W), dtype=torch.uint8, device="cuda")
    conf = torch.zeros((H, W), dtype=torch.uint8, device="cuda")
    t0 = time.time()
    ctx.detect_strip(tiles, T, 8, cls, conf, W, 0)
    torch.cuda.synchronize()
    log(f"detect_strip {len(tiles)} tiles in {time.time() - t0:.3f}s")
    t0 = time.time()
    ref_cls, ref_conf, rows = run_zone(model, GeoRaster(raster, 800000.0, 6500000.0 + H * 0.2, 0.2), config)
    log(f"oracle run_zone in {time.time() - t0:.1f}s ({len(rows)} tiles)")
    cls_h, conf_h = cls.cpu().numpy(), conf.cpu().numpy()
    agree = (cls_h == ref_cls).mean()
    log(f"ZONE argmax agreement = {agree * 100:.4f}%  conf agreement = {(conf_h == ref_conf).mean() * 100:.4f}%  "
        f"classes used = {np.unique(ref_cls).tolist()}")
    for mode in (0,):
        pass
    # host-in/host-out entry point
    out_cls = np.zeros((H, W), np.uint8)
    out_conf = np.zeros((H, W), np.uint8)
    ctx.detect_zone_host(raster, [0, 1, 2], W, H, 0, nat.FB_LAYOUT_CHW, tiles, T, 8, out_cls, out_conf, W, 0, H)
    log(f"detect_zone_host identical to detect_strip: {(out_cls == cls_h).all()} {(out_conf == conf_h).all()}")
    prof = ctx.profile_forward(8, T, 3)
    log("profile (8 tiles):", {k: round(v, 3) for k, v in prof.items()})
    prof = ctx.profile_forward(32, T, 3)
    log("profile (32 tiles):", {k: round(v, 3) for k, v in prof.items()})
    log("DONE")


if __name__ == "__main__":
    main()
