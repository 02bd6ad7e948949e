"""A/B timing driver (not a pytest module): the exact-clipping zone loop of the bench zone (10000 x 10000, 1600 tiles,
raster resident) under several environment settings inside ONE job, alternating, because box-to-box and minute-to-minute
spread on the power-capped parts is larger than most single optimisations.

    python tests/ab_zone.py "" "FB_NO_D2S=1"              # two configurations, 3 rounds of (warm-up + 3 steps) each
    python tests/ab_zone.py --rounds 2 --steps 2 "" "FB_EPI2=0"

Every configuration gets its own context (the library reads its switches in fb_create). Prints per round and
configuration: ms per step, Mpx/s, conv TFLOP/s of the outputs computed, and at the end the class maps' agreement
between the first configuration and each of the others (they must be identical unless a switch changes arithmetic).
"""
import argparse
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import flair1_b200._native as nat  # noqa: E402
from flair1_b200.zone_detect.slicing_job import tile_table  # noqa: E402
from oracle import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rounds", type=int, default=3)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--size", type=int, default=10000)
ap.add_argument("--batch", type=int, default=148)
ap.add_argument("--layers", action="store_true", help="print the library's per-layer CUDA-event times of the last round (FB_LAYER_TIMES=1)")
ap.add_argument("configs", nargs="+")
args = ap.parse_args()

T, margin, W, H = 512, 128, args.size, args.size
raster = torch.from_numpy(synth.synth_raster(3, H, W, seed=1)).cuda()
tiles = tile_table(W, H, T, margin)
sd = synth.cached_checkpoint(3, 15)


def make(envs: str):
    keys = []
    for kv in envs.split():
        k, v = kv.split("=", 1)
        os.environ[k] = v
        keys.append(k)
    ctx = nat.Context(0)
    ctx.load_weights(sd, 3, 15)
    ctx.set_norm("custom", synth.FLAIR_MEANS[:3], synth.FLAIR_STDS[:3])
    ctx.set_raster(raster, [0, 1, 2], W, H)
    return ctx, keys


def with_env(envs: str, fn):
    """Some switches are read at launch time (getenv in the launch helpers), so they are set around every call too."""
    keys = []
    for kv in envs.split():
        k, v = kv.split("=", 1)
        os.environ[k] = v
        keys.append(k)
    try:
        return fn()
    finally:
        for k in keys:
            os.environ.pop(k, None)


ctxs = []
for envs in args.configs:
    ctx, keys = make(envs)
    for k in keys:
        os.environ.pop(k, None)
    ctxs.append(ctx)
maps = [(torch.zeros((H, W), dtype=torch.uint8, device="cuda"), torch.zeros((H, W), dtype=torch.uint8, device="cuda")) for _ in ctxs]

for rnd in range(args.rounds):
    for i, (envs, ctx) in enumerate(zip(args.configs, ctxs)):
        cls, conf = maps[i]

        def step():
            ctx.detect_strip(tiles, T, args.batch, cls, conf, W, 0)

        with_env(envs, step)
        ctx.synchronize()
        f0 = ctx.flop_count
        layers = args.layers and rnd == args.rounds - 1
        if layers:
            os.environ["FB_LAYER_TIMES"] = "1"
            print(f"--- per-layer times [{envs or 'default'}], {args.steps} steps ---", file=sys.stderr, flush=True)
            ctx.profile_begin()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            with_env(envs, step)
        b.record()
        torch.cuda.synchronize()
        if layers:
            ctx.profile_end()
            os.environ.pop("FB_LAYER_TIMES", None)
        ms = a.elapsed_time(b) / args.steps
        fl = (ctx.flop_count - f0) / args.steps
        print(f"round {rnd} [{envs or 'default'}]: {ms:.2f} ms/step, {W * H / 1e3 / ms:.1f} Mpx/s, "
              f"{fl / ms / 1e9:.1f} TFLOP/s on {fl / len(tiles) / 1e9:.2f} GFLOP/tile", flush=True)
for i in range(1, len(ctxs)):
    same = (maps[0][0] == maps[i][0]).float().mean().item()
    same_c = (maps[0][1] == maps[i][1]).float().mean().item()
    print(f"class map [{args.configs[0] or 'default'}] vs [{args.configs[i]}]: {100 * same:.4f} % equal, confidence {100 * same_c:.4f} %")
