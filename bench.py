#!/usr/bin/env python
"""zone_detect throughput benchmark (BASELINE.json metric: zone_detect Mpixels/s, ResNet34-UNet).

    python bench.py --gpus N --steps K --warmup W            # our arm: CUDA path through the C ABI
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the CPU path (oracle port)
    python bench.py --verify-full                            # N = 1: also check EVERY tile of the zone against the oracle

One step = one pass of the whole hot path over one synthetic zone WITH its `-m` work: tile extraction +
normalisation, U-Net forward, softmax/argmax, margin clipping and stitching into the uint8 class map, the
confusion matrix of the map against a truth raster, and (N > 1) the NCCL all-reduce of the per-rank matrices.
  N = 1 : BASELINE.json configs[1] -- 10000 x 10000 RGB raster, 512 px tiles, margin 128, 15 classes
          (1600 tiles).
  N > 1 : configs[3] -- ONE 40000 x 40000 raster (24 649 tiles) strong-scaled over the N ranks (one process per
          GPU): each rank takes a contiguous range of the row-major tile order (3081 or 3082 tiles at N = 8), reads
          its raster rows itself (halo rows re-read, no exchange) and owns a disjoint set of row bands of the map.
`value` is timed with inputs resident in HBM. `e2e` is the same step through fb_detect_zone_shard with pinned
HOST buffers: raster and truth rows go up, and every rank's write rectangles come down straight into ONE
[2, H, W] output map in shared memory (the reference's single output raster, main.py:421-426), so after the
closing barrier the writer rank holds the complete map; the summed confusion matrix is read back to the host.
Every line carries tiles_per_s_per_gpu, the figure that is comparable between the two zone sizes.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

TILE, MARGIN, NCLS, BANDS = 512, 128, 15, 3
GFLOP_PER_TILE = 63.569  # 2*MAC of the 47 convolutions of one 512^2 tile, 3 bands / 15 classes
MEANS = [105.08, 110.87, 101.82]
STDS = [52.17, 45.38, 44.0]
HBM_FALLBACK_GBS = 6650.0   # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------------ data
def synth_rows_gpu(width: int, height: int, y0: int, y1: int, seed: int, device) -> torch.Tensor:
    """uint8 [3, y1-y0, width] rows of the seeded synthetic zone (low-frequency field + pixel noise,
    the family the synthetic checkpoint was trained on). Row blocks of 512 are generated from
    (seed, block index) so any rank materialises identical bytes for the same global rows."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    lh, lw = (height + 511) // 512 * 8 + 4, (width + 63) // 64 + 4
    low = torch.randn((1, BANDS, lh, lw), generator=g).to(device)
    out = torch.empty((BANDS, y1 - y0, width), dtype=torch.uint8, device=device)
    blk = 512
    for b0 in range((y0 // blk) * blk, y1, blk):
        r0, r1 = max(b0, y0), min(b0 + blk, y1)
        l0 = b0 // 64  # low-res rows needed by bicubic for this block: [l0-2, l0+blk/64+2] -> apron of 2
        crop = low[:, :, l0:l0 + blk // 64 + 4, :]
        up = torch.nn.functional.interpolate(crop, scale_factor=64, mode="bicubic", align_corners=False)[0]
        up = up[:, 128:128 + blk, 128:128 + width]
        gn = torch.Generator(device=device).manual_seed(seed * 1000003 + b0 // blk)
        noise = torch.randn((BANDS, blk, width), generator=gn, device=device)
        rows = (up * 50 + 110 + noise * 10).clamp_(0, 255).round_().to(torch.uint8)
        out[:, r0 - y0:r1 - y0] = rows[:, r0 - b0:r1 - b0]
    return out


def synth_truth_rows_gpu(width: int, y0: int, y1: int, seed: int, device) -> torch.Tensor:
    """uint8 [y1-y0, width] truth labels 1..19 (classes 16..19 fall outside the 15-class matrix and are dropped, like
    the weight-0 classes of the reference nomenclature); rows are generated per global row block so that every rank
    sees the same bytes for the same rows."""
    out = torch.empty((y1 - y0, width), dtype=torch.uint8, device=device)
    blk = 512
    for b0 in range((y0 // blk) * blk, y1, blk):
        r0, r1 = max(b0, y0), min(b0 + blk, y1)
        g = torch.Generator(device=device).manual_seed(seed * 7919 + b0 // blk)
        rows = torch.randint(1, 20, (blk, width), dtype=torch.uint8, device=device, generator=g)
        out[r0 - y0:r1 - y0] = rows[r0 - b0:r1 - b0]
    return out


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    QUERY = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.samples = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.samples:
            if ts < t0 or ts > t1:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def sample_tile_indices(n_tiles: int, nx: int, ny: int, count: int) -> list:
    """Tiles spread over the zone, always including the 4 corners and the clamped last row/column."""
    picks = {0, ny - 1, (nx - 1) * ny, nx * ny - 1, (nx - 1) * ny + ny // 2, (nx // 2) * ny + ny - 1}
    picks = {p for p in picks if 0 <= p < n_tiles}
    step = max(1, n_tiles // max(1, count - len(picks)))
    picks |= set(range(step // 2, n_tiles, step))
    return sorted(picks)[:max(count, 6)]


def cpu_zone_sample(model, raster_rows_fn, W: int, H: int, tiles: np.ndarray, idx: list, batch: int = 4):
    """The reference's CPU path (oracle restatement of main.py:398-426: float64 normalise -> fp32 forward
    -> softmax -> crop -> argmax) on the tiles `idx`. Returns (seconds, {tile: class patch}, written px)."""
    from oracle.zone_detect_ref import normalization, stitching_exact_clipping
    out, px = {}, 0
    t0 = time.perf_counter()
    for s in range(0, len(idx), batch):
        ids = idx[s:s + batch]
        imgs = []
        for i in ids:
            x0, y0 = int(tiles[i, 0]), int(tiles[i, 1])
            patch = np.zeros((BANDS, TILE, TILE), np.uint8)
            r0, r1, c0, c1 = max(y0, 0), min(y0 + TILE, H), max(x0, 0), min(x0 + TILE, W)
            patch[:, r0 - y0:r1 - y0, c0 - x0:c1 - x0] = raster_rows_fn(r0, r1)[:, :, c0:c1]
            imgs.append(torch.as_tensor(normalization(patch, "custom", MEANS, STDS), dtype=torch.float))
        with torch.no_grad():
            probs = torch.softmax(model(torch.stack(imgs)), dim=1).cpu().numpy()
        for i, p in zip(ids, probs):
            out[i] = stitching_exact_clipping(p, MARGIN, TILE, "argmax")[0].astype(np.uint8)
            px += int((tiles[i, 4] - tiles[i, 2]) * (tiles[i, 5] - tiles[i, 3]))
    return time.perf_counter() - t0, out, px


def agreement(cpu_out: dict, tiles: np.ndarray, cls_map: np.ndarray, map_row0: int = 0):
    """(#equal, #compared) between the oracle's per-tile class patches and the stitched map on each tile's write rectangle."""
    same = tot = 0
    for i, patch in cpu_out.items():
        x0, y0, wx0, wy0, wx1, wy1 = (int(v) for v in tiles[i])
        ref = patch[wy0 - (y0 + MARGIN):wy1 - (y0 + MARGIN), wx0 - (x0 + MARGIN):wx1 - (x0 + MARGIN)]
        got = cls_map[wy0 - map_row0:wy1 - map_row0, wx0:wx1]
        same += int((ref == got).sum())
        tot += ref.size
    return same, tot


def load_oracle_model():
    from oracle import synth
    from oracle.unet_smp033 import Unet
    sd = synth.cached_checkpoint(BANDS, NCLS)
    m = Unet(BANDS, NCLS)
    m.load_state_dict(sd, strict=True)
    m.eval()
    return sd, m


def workload_config(n_gpus: int, W: int, H: int, n_tiles: int) -> dict:
    return {"workload": f"zone_detect {W}x{H} RGB uint8 raster, {TILE}px tiles, margin {MARGIN}, {NCLS} classes, "
                        f"ResNet34-UNet, {n_tiles} tiles, with -m (confusion matrix vs a truth raster)"
                        + ("" if n_gpus == 1 else f", tile ranges sharded over {n_gpus} GPUs"),
            "raster": [W, H], "tile": TILE, "margin": MARGIN, "n_classes": NCLS, "bands": BANDS, "tiles": n_tiles,
            "l2_policy": "inputs larger than L2 (raster strip >= 300 MB, activations > 2 GB per batch)"}


def run_reference(args) -> int:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from flair1_b200.zone_detect.slicing_job import tile_table
    torch.set_num_threads(os.cpu_count() or 1)
    _, model = load_oracle_model()
    W = H = 10000 if args.gpus == 1 else 40000
    tiles = tile_table(W, H, TILE, MARGIN)
    ny = len(np.unique(tiles[:, 1]))
    nx = len(tiles) // ny
    per_step = args.ref_tiles
    cache = {}

    def rows_fn(r0, r1):  # the host raster is materialised lazily in 512-row blocks (same seed family)
        parts = []
        for b0 in range((r0 // 512) * 512, r1, 512):
            if b0 not in cache:
                cache[b0] = synth_rows_gpu(W, H, b0, min(b0 + 512, H), 1, "cpu").numpy()
            parts.append(cache[b0][:, max(r0, b0) - b0:min(r1, b0 + 512) - b0])
        return parts[0] if len(parts) == 1 else np.concatenate(parts, axis=1)

    idx_all = sample_tile_indices(len(tiles), nx, ny, per_step * (args.steps + args.warmup))
    times, pxs = [], []
    for step in range(args.warmup + args.steps):
        idx = idx_all[step * per_step:(step + 1) * per_step] or idx_all[:per_step]
        for i in idx:  # materialise input rows outside the timed region ("inputs resident")
            rows_fn(max(int(tiles[i, 1]), 0), min(int(tiles[i, 1]) + TILE, H))
        dt, _, px = cpu_zone_sample(model, rows_fn, W, H, tiles, idx)
        if step >= args.warmup:
            times.append(dt)
            pxs.append(px)
    value = sum(pxs) / sum(times) / 1e6
    line = {
        "impl": "reference", "metric": "zone_detect Mpixels/s", "value": value, "unit": "Mpixels/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup,
        # a step of this arm is a bounded SAMPLE of the zone (ref_tiles tiles); this is the time it really took
        "ms_per_step": 1e3 * sum(times) / len(times),
        "ms_per_zone_extrapolated": 1e3 * (W * H / 1e6) / value,
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args.gpus, W, H, len(tiles)),
        "cpu_baseline": {"value": value, "unit": "Mpixels/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{per_step} tiles per step x {args.steps} steps of the same zone (batch 4), linear in tiles"},
        "e2e": {"value": value, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "tiles_per_s_per_gpu": None, "tiles_per_s_cpu": len(times) * per_step / sum(times),
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------ extras (N = 1, outside the timed region)
def event_time_ms(fn, iters: int = 5, warm: int = 2) -> float:
    """Mean CUDA-event time of fn() on the current stream (after `warm` untimed calls)."""
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def hbm_kernel_legs(ctx, nat, dev, W: int, H: int, tiles: np.ndarray, cls_dev, truth_dev, hbm_peak: float) -> dict:
    """Achieved GB/s of the HBM-bound kernels on the 10000^2 sizes (algorithmic bytes / CUDA-event time), each as a
    fraction of the measured copy bandwidth: K9 confusion histogram, K1 extract + normalise, K6 / K6b stitching of a
    batch of logits, K7 / K8 blended accumulation and finalisation, K9b per-tile confusion."""
    out = {}
    npx = W * H
    cm = torch.zeros((NCLS, NCLS), dtype=torch.int64, device=dev)

    def leg(name, ms, nbytes, note):
        gbs = nbytes / (ms * 1e-3) / 1e9
        out[name] = {"ms": ms, "algorithmic_bytes": int(nbytes), "GBps": gbs, "frac_of_hbm_peak": gbs / hbm_peak, "what": note}

    leg("K9_confusion", event_time_ms(lambda: ctx.confusion(cls_dev, truth_dev, NCLS, truth_sub=1, out=cm)), 2 * npx,
        f"{npx / 1e6:.0f} Mpx class map vs truth, 2 B per px")
    nb = 148
    batch_tiles = tiles[:nb]
    # K1 through the public forward call is not separable; use the profile hooks around one forward pass instead
    ctx.profile_begin()
    logits = ctx.forward_tiles(batch_tiles[:, :2], TILE)
    prof = ctx.profile_end()
    x0_bytes = nb * TILE * TILE * (BANDS + 8)   # 3 band bytes read + 8 B written per px (space-to-depth form: 16 bf16 per 2x2 px)
    leg("K1_extract_normalise", prof["extract_ms"], x0_bytes, f"{nb} tiles, {BANDS} B read + 8 B written per px (2x2 space-to-depth bf16)")
    ls = ctx.logit_stride
    cls_b = torch.zeros((H, W), dtype=torch.uint8, device=dev)
    conf_b = torch.zeros((H, W), dtype=torch.uint8, device=dev)
    lib, h = ctx._lib, ctx._h
    import ctypes as C
    tiles_c = np.ascontiguousarray(batch_tiles, dtype=np.int32)
    # stitch kernels are reached through the C ABI entry points that run them on a batch of logits: the zone loops with the
    # fused sink switched off (FB_NO_FUSED_SINK is read at fb_create) are not available on this context, so the standalone
    # kernels are timed through fb_detect_strip_prob (K6b) and fb_blend_strip / fb_blend_finalize (K7 / K8) minus the
    # network time of the same batch (fb_profile_* attributes every launch to a family).
    prob = torch.zeros((NCLS, 2048, W), dtype=torch.uint8, device=dev)
    sub = tiles_c[(tiles_c[:, 3] >= 0) & (tiles_c[:, 5] <= 2048)]
    ctx.profile_begin()
    ctx.detect_strip_prob(sub, TILE, len(sub), prob, W, 0)
    p2 = ctx.profile_end()
    wpx = int(((sub[:, 4] - sub[:, 2]) * (sub[:, 5] - sub[:, 3])).sum())
    leg("K6b_prob_stitch", p2["stitch_ms"], wpx * (4 * ls + NCLS), f"{len(sub)} tiles: {4 * ls} B of logits read + {NCLS} B written per written px")
    for method in ("average_weights", "max"):
        acc, wsum = ctx.blend_buffers(method, 2048, W)
        ctx.profile_begin()
        ctx.blend_strip(sub, TILE, len(sub), method, acc, wsum, W, 0)
        p3 = ctx.profile_end()
        cpx = 0
        for t in sub:   # covered px of each whole tile, clipped to the raster and the 2048 map rows
            cpx += max(0, min(int(t[0]) + TILE, W) - max(int(t[0]), 0)) * max(0, min(int(t[1]) + TILE, 2048) - max(int(t[1]), 0))
        per_px = 4 * ls + (2 * (4 * ls + 4) if method != "max" else 16)   # logits read + accumulator read-modify-write
        leg(f"K7_blend_accumulate_{method}", p3["stitch_ms"], cpx * per_px, f"{len(sub)} tiles, {per_px} B per covered px (atomics)")
        ms8 = event_time_ms(lambda: ctx.blend_finalize(method, acc, wsum, cls_b[:2048], conf_b[:2048]))
        per_px8 = (4 * ls + 4 if method != "max" else 8) + 2
        leg(f"K8_blend_finalize_{method}", ms8, 2048 * W * per_px8, f"2048 x {W} px, {per_px8} B per px")
        del acc, wsum
    wins = np.ascontiguousarray(sub.copy())
    truth_sub = truth_dev[:2048].contiguous()
    ctx.profile_begin()
    ctx.detect_strip_metrics(sub, wins, TILE, len(sub), cls_b[:2048], conf_b[:2048], W, 0, truth_sub, truth_sub=1)
    p4 = ctx.profile_end()
    out["K6_argmax_stitch_plus_K9b_tile_confusion"] = {
        "ms": p4["stitch_ms"], "what": f"{len(sub)} whole tiles of fp32 logits: K6 ({4 * ls} B read + 2 B written per written px) and "
                                       f"K9b ({4 * ls} + 1 B per window px) share one event pair",
        "algorithmic_bytes": int(wpx * (2 * 4 * ls + 3)), "GBps": wpx * (2 * 4 * ls + 3) / (p4["stitch_ms"] * 1e-3) / 1e9,
        "frac_of_hbm_peak": wpx * (2 * 4 * ls + 3) / (p4["stitch_ms"] * 1e-3) / 1e9 / hbm_peak}
    del logits, prob, cls_b, conf_b
    return out


def patch_sweep_leg(nat, dev, peak_tf: float) -> dict:
    """BASELINE.json configs[4]: 5-band / 19-class whole-patch predict (fb_predict_patches, patches resident in HBM), batch
    1 .. 256: patches/s and conv TFLOP/s against the tensor roofline. Random-init weights (logits-tolerance config)."""
    from oracle import synth
    sd = synth.random_checkpoint(5, 19, seed=5)
    ctx = nat.Context(dev.index)
    ctx.load_weights(sd, 5, 19)
    ctx.set_norm("custom", synth.FLAIR_MEANS, synth.FLAIR_STDS)
    g = torch.Generator(device=dev).manual_seed(5)
    patches = torch.randint(0, 256, (256, 5, TILE, TILE), dtype=torch.uint8, device=dev, generator=g)
    rows = []
    for b in (1, 2, 4, 8, 16, 32, 64, 128, 148, 256):
        n = max(b, 148) if b < 148 else b
        n = (n // b) * b
        f0 = ctx.flop_count
        ms = event_time_ms(lambda: ctx.predict_patches(patches[:n], TILE, b), iters=3, warm=1)
        flops = (ctx.flop_count - f0) / 4   # 1 warm + 3 timed calls
        rows.append({"batch": b, "patches": n, "patches_per_s": n / (ms * 1e-3), "tflops": flops / (ms * 1e-3) / 1e12,
                     "frac_of_tensor_peak": flops / (ms * 1e-3) / 1e12 / peak_tf})
    ctx.close()
    return {"config": "5 bands, 19 classes, 512x512 patches, random-init, whole-call time incl. extract + argmax", "rows": rows}


def blend_leg(ctx, dev, W: int, H: int, tiles: np.ndarray) -> dict:
    """a8: the `average_weights` stitching of a 4096-row band of the zone (every tile that touches it, whole tiles, logits
    through K7 / K8): Mpx/s of finished map."""
    rows = 4096
    sub = tiles[(tiles[:, 1] < rows) & (tiles[:, 1] + TILE > 0)]
    cls_b = torch.zeros((rows, W), dtype=torch.uint8, device=dev)
    conf_b = torch.zeros((rows, W), dtype=torch.uint8, device=dev)

    def run():
        acc, wsum = ctx.blend_buffers("average_weights", rows, W)
        ctx.blend_strip(sub, TILE, 148, "average_weights", acc, wsum, W, 0)
        ctx.blend_finalize("average_weights", acc, wsum, cls_b, conf_b)

    ms = event_time_ms(run, iters=2, warm=1)
    return {"method": "average_weights", "map": [W, rows], "tiles": int(len(sub)), "ms": ms, "mpx_per_s": rows * W / 1e6 / (ms * 1e-3)}


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args) -> int:
    import flair1_b200._native as nat
    from flair1_b200 import numa
    from flair1_b200.zone_detect.shared_map import SharedHostMap
    from flair1_b200.zone_detect.slicing_job import owned_rects, split_tiles_across_ranks, tile_table

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            log(f"bench.py --gpus {args.gpus} must be launched with torch.distributed.run --nproc-per-node {args.gpus}")
            return 2
    if not torch.cuda.is_available():
        log("bench.py: no CUDA device; the product path has no CPU fallback")
        return 3
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)

    W = H = 10000 if world == 1 else 40000
    tiles_all = tile_table(W, H, TILE, MARGIN)
    shards = split_tiles_across_ranks(tiles_all, world)
    tiles = tiles_all[shards[rank]]                 # y-sorted contiguous range of the row-major tile order
    rects = owned_rects(tiles)                      # (y0, y1, x0, x1) row bands this rank writes
    own_px = int(((rects[:, 1] - rects[:, 0]) * (rects[:, 3] - rects[:, 2])).sum())
    # raster rows this rank needs (tile rows + halo), class-map rows this rank writes
    ry0, ry1 = max(int(tiles[:, 1].min()), 0), min(int(tiles[:, 1].max()) + TILE, H)
    my0, my1 = int(rects[:, 0].min()), int(rects[:, 1].max())
    sd, model = load_oracle_model() if rank == 0 else (None, None)
    if rank != 0:
        from oracle import synth
        sd = synth.cached_checkpoint(BANDS, NCLS)

    ctx = nat.Context(local)
    ctx.load_weights(sd, BANDS, NCLS)
    ctx.set_norm("custom", MEANS, STDS)
    if world > 1:
        def exchange(ident):
            box = [ident]
            dist.broadcast_object_list(box, src=0)
            return box[0]
        ctx.comm_init(rank, world, exchange)        # the library's own NCCL communicator (fb_allreduce_confusion)

    raster_dev = synth_rows_gpu(W, H, ry0, ry1, 1, dev)
    truth_dev = synth_truth_rows_gpu(W, my0, my1, 4, dev)
    cls_dev = torch.zeros((my1 - my0, W), dtype=torch.uint8, device=dev)
    conf_dev = torch.zeros((my1 - my0, W), dtype=torch.uint8, device=dev)
    cm_dev = torch.zeros((NCLS, NCLS), dtype=torch.int64, device=dev)
    # host staging: pages first-touched (and page-locked) while the process sits on the CPUs of its GPU's NUMA node
    affinity0 = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    numa_report = numa.bind_to_gpu_node(local) if world > 1 else {"node": None, "cpus": None, "bound": False}
    raster_host = torch.empty(raster_dev.shape, dtype=torch.uint8, pin_memory=True)
    raster_host.copy_(raster_dev)
    truth_host = torch.empty(truth_dev.shape, dtype=torch.uint8, pin_memory=True)
    truth_host.copy_(truth_dev)
    cm_host = torch.zeros((NCLS, NCLS), dtype=torch.int64, pin_memory=True)
    shared = None
    if world == 1:
        out_host = torch.full((2, H, W), 255, dtype=torch.uint8, pin_memory=True)
        out_map = out_host.numpy()
    else:
        box = [SharedHostMap.fresh_path(2 * H * W, "bench") if rank == 0 else None]
        if rank == 0:
            shared = SharedHostMap(box[0], 2, H, W, create=True)
            shared.array[0].fill(255)               # no valid class: every pixel must be overwritten by some rank
        dist.broadcast_object_list(box, src=0)
        if rank != 0:
            shared = SharedHostMap(box[0], 2, H, W, create=False)
        shared.pin_rows(my0, my1)
        out_map = shared.array
    if affinity0 is not None and numa_report["bound"]:
        os.sched_setaffinity(0, affinity0)
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def score_resident():
        cm_dev.zero_()
        for y0, y1, x0, x1 in rects:   # the rectangles this rank owns (<= 21 launches; whole rows but for the first / last band)
            ctx.confusion_rect(cls_dev[y0 - my0:y1 - my0, x0:x1], truth_dev[y0 - my0:y1 - my0, x0:x1], NCLS, truth_sub=1, out=cm_dev)
        if world > 1:
            ctx.allreduce_confusion(cm_dev)

    def step_resident():
        ctx.set_raster(raster_dev, [0, 1, 2], W, H, row0=ry0)
        ctx.detect_strip(tiles, TILE, args.batch, cls_dev, conf_dev, W, my0)
        score_resident()

    def step_e2e():
        t = time.perf_counter()
        cm_dev.zero_()
        ctx.detect_zone_shard(raster_host, [0, 1, 2], W, H, ry0, nat.FB_LAYOUT_CHW, tiles, TILE, args.batch,
                              out_map[0], out_map[1], W, 0, H, truth=truth_host, truth_row0=my0, truth_sub=1, cm=cm_dev)
        if world > 1:
            ctx.allreduce_confusion(cm_dev)
        cm_host.copy_(cm_dev)           # the step's result (the summed matrix) read back to the host
        torch.cuda.current_stream().synchronize()
        if rank == 0:
            log(f"[bench] e2e step {1e3 * (time.perf_counter() - t):.1f} ms")

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t1 = time.time()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), t0, t1

    # nvidia-smi needs a few hundred ms to deliver its first sample: start it before the warm-up, keep the samples that
    # fall inside the timed region
    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        step_resident()
    launches0, flops0 = ctx.launch_count, ctx.flop_count
    ctx.profile_begin()
    ms_total, t0, t1 = timed(step_resident, args.steps)
    prof = ctx.profile_end()
    launches = ctx.launch_count - launches0
    flops = ctx.flop_count - flops0   # algorithmic FLOPs of the conv outputs this rank actually computed
    clocks = sampler.stop(t0, t1) if sampler else None
    cm_resident = cm_dev.cpu().numpy().copy()
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    ms_e2e, _, _ = timed(step_e2e, args.steps)
    cm_e2e = cm_host.numpy().copy()

    # ---- checks on the 40000^2 / 10000^2 run itself: the summed matrix counts every in-range truth pixel exactly once,
    # both paths give the same matrix, and every pixel of the one output map was delivered by some rank
    inrange = torch.tensor([sum(int(((truth_dev[y0 - my0:y1 - my0, x0:x1] >= 1) & (truth_dev[y0 - my0:y1 - my0, x0:x1] <= NCLS)).sum().item())
                                for y0, y1, x0, x1 in rects), own_px], dtype=torch.int64, device=dev)
    if dist is not None:
        dist.all_reduce(inrange)
    assert int(inrange[1].item()) == W * H, "the shards' write rectangles do not partition the raster"
    assert int(cm_e2e.sum()) == int(inrange[0].item()), f"confusion matrix counts {int(cm_e2e.sum())} px, {int(inrange[0].item())} truth px are in range"
    assert np.array_equal(cm_e2e, cm_resident), "resident and end-to-end paths disagree on the confusion matrix"
    barrier()
    if rank == 0:
        assert int((np.asarray(out_map[0]) == 255).sum()) == 0, "a pixel of the shared output map was never written"
    own_cls = torch.from_numpy(np.ascontiguousarray(out_map[0][my0:my1]))
    for y0, y1, x0, x1 in rects:
        assert torch.equal(own_cls[y0 - my0:y1 - my0, x0:x1], cls_dev[y0 - my0:y1 - my0, x0:x1].cpu()), "e2e map differs from the resident map"

    h2d = raster_host.numel() + truth_host.numel()
    d2h = 2 * own_px + cm_host.numel() * 8
    lt = torch.tensor([launches, h2d, d2h], dtype=torch.int64, device=dev)
    conv_ms = torch.tensor([prof["conv_ms"]], dtype=torch.float64, device=dev)
    # per-GPU conv rate of this rank (executed FLOPs / its summed conv time); the line reports the slowest rank
    rate = torch.tensor([flops / (prof["conv_ms"] / 1e3) / 1e12, -flops / max(len(tiles), 1) / args.steps / 1e9],
                        dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(lt)
        dist.all_reduce(conv_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(rate, op=dist.ReduceOp.MIN)
    mpx = W * H / 1e6
    value = mpx * args.steps / (ms_total / 1e3)
    e2e_value = mpx * args.steps / (ms_e2e / 1e3)

    if rank == 0:
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        peak_tf, peak_src = (peaks["bf16_tflops_sustained"], "measured (MEASURED_PEAKS.json, sustained)") \
            if "bf16_tflops_sustained" in peaks else (1400.0, "fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)")
        hbm_peak = float(peaks.get("hbm_gbs", HBM_FALLBACK_GBS))
        max_tiles = max(len(s) for s in shards)
        conv_tflops = float(rate[0].item())
        gflop_per_tile = -float(rate[1].item())
        full_tile_tflops = max_tiles * GFLOP_PER_TILE * args.steps / (float(conv_ms.item()) / 1e3) / 1e3
        # CPU baseline + agreement on a bounded sample of rank 0's tiles (all of them with --verify-full)
        torch.set_num_threads(os.cpu_count() or 1)
        ny = len(np.unique(tiles_all[:, 1]))
        nx = len(tiles_all) // ny
        if args.verify_full and world == 1:
            idx = list(range(len(tiles)))
        else:
            # the CPU baseline is an N=1 figure; at N>1 only a small sample is run, for the agreement check
            idx = sample_tile_indices(len(tiles), nx if world == 1 else len(tiles), ny if world == 1 else 1,
                                      args.cpu_tiles if world == 1 else 24)
        rh = raster_host.numpy()
        cpu_s, cpu_out, cpu_px = cpu_zone_sample(model, lambda r0, r1: rh[:, r0 - ry0:r1 - ry0], W, H, tiles, idx)
        same, tot = agreement(cpu_out, tiles, np.asarray(out_map[0]))
        line = {
            "metric": "zone_detect Mpixels/s", "value": value, "unit": "Mpixels/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            # N > 1: one fixed 40000^2 raster divided over the ranks. (N = 1 runs BASELINE configs[1], the 10000^2 zone,
            # which has 16.0 tiles per output Mpx against 15.4: compare tiles_per_s_per_gpu across N, not Mpx/s.)
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(world, W, H, len(tiles_all)),
            "batch_tiles": args.batch,
            "tiles_per_s_per_gpu": max_tiles * args.steps / (ms_total / 1e3),
            "e2e_tiles_per_s_per_gpu": max_tiles * args.steps / (ms_e2e / 1e3),
            "tiles_slowest_rank": max_tiles, "shard": "contiguous ranges of the row-major tile order (split_tiles_across_ranks)",
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mpixels/s", "h2d_bytes_per_step": int(lt[1].item()),
                    "d2h_bytes_per_step": int(lt[2].item()),
                    "includes": "raster + truth upload, tiles, fused confusion matrix, NCCL all-reduce of it (N > 1), every rank's "
                                "write rectangles copied into ONE shared [2,H,W] host map, matrix read back",
                    "output_map": "pinned host tensor" if world == 1 else f"POSIX shared memory ({Path(shared.path).parent}), rows page-locked per rank",
                    "numa": numa_report},
            "gpu_launches": int(lt[0].item()),
            "roofline": {"bound": "tensor", "achieved": conv_tflops, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": conv_tflops / peak_tf,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch of the family's top kernel
                         # (conv_igemm2_kernel<256>, a layer3 3x3 conv on 148 tiles: 77.6 MB in, 77.6 MB out, 1.2 MB of
                         # weights; part of the output stays in the 126 MB L2 for the next layer; 114.2 us, tensor pipe
                         # 88 % active), from the `ncu --set full` capture in profiles/r01_ncu_full_igemm2_256_v21.csv
                         "traffic": 114.74e6, "traffic_unit": "bytes per launch (conv_igemm2_kernel<256>, 148 tiles)",
                         "peak_source": peak_src,
                         "kernel": "conv_igemm2_kernel / conv_igemm_kernel / conv_halo_kernel (47 launches per batch; the stem launch also pools); achieved = algorithmic FLOPs "
                                   "(2*MAC of the direct conv) of the outputs actually computed / summed conv time",
                         "gflop_per_tile_computed": gflop_per_tile, "gflop_per_tile_full": GFLOP_PER_TILE,
                         "note": "the decoder skips outputs that can only reach the cropped margin (bit-identical class map, "
                                 "csrc/tile_need.cuh); full_tile_equivalent_tflops counts 63.569 GFLOP per tile instead",
                         "full_tile_equivalent_tflops": full_tile_tflops,
                         "conv_share_of_step": float(conv_ms.item()) / ms_total},
            "cpu_baseline": {"value": cpu_px / cpu_s / 1e6, "unit": "Mpixels/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{len(idx)} of rank 0's {len(tiles)} tiles (batch 4, fp32 torch CPU oracle), linear in tiles"},
            "argmax_agreement_pct": 100.0 * same / max(tot, 1),
            "argmax_agreement_tiles": len(idx),
            "checks": {"cm_sum_equals_in_range_truth_px": int(cm_e2e.sum()), "e2e_cm_equals_resident_cm": True,
                       "every_map_pixel_written": True, "write_rectangles_partition_raster": True},
            "stage_ms_per_step": {k: v / args.steps for k, v in prof.items()},
        }
        if world > 1:
            line["cpu_baseline"] = None  # reported at N=1 only (torchrun pins OMP threads; see the N=1 line)
        elif not args.no_extras:
            try:
                line["hbm_kernels"] = hbm_kernel_legs(ctx, nat, dev, W, H, tiles, cls_dev, truth_dev, hbm_peak)
                line["hbm_kernels"]["peak_GBps"] = hbm_peak
                line["blend"] = blend_leg(ctx, dev, W, H, tiles_all)
                line["patch_sweep"] = patch_sweep_leg(nat, dev, peak_tf)
            except Exception as e:   # an extra must never cost the headline line
                line["extras_error"] = repr(e)
        if args.verify_full and world == 1:
            line["verify_full"] = {"tiles": len(idx), "pixels": tot, "agree": same, "agreement_pct": 100.0 * same / max(tot, 1),
                                   "cpu_seconds": cpu_s}
        emit(line)
    if dist is not None:
        dist.barrier()
    if shared is not None:
        shared.close()
        if rank == 0:
            shared.unlink()
    if dist is not None:
        dist.destroy_process_group()
    ctx.close()
    return 0


def emit(line: dict) -> None:
    """The one JSON line of this run, on the process's ORIGINAL stdout (see main())."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main() -> int:
    # Libraries may chat on stdout (NCCL prints its version line there): everything but the JSON line goes to stderr.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=296, help="tiles per forward pass (296 = two per SM; round 1: 74 / 111 / 148 measured 906 / 907 / 922 Mpx/s, round 2: 148 / 222 / 296 measured 1209 / 1231 / 1233 Mpx/s)")
    ap.add_argument("--cpu-tiles", type=int, default=160, help="tiles of the zone timed on the host cores (cpu_baseline)")
    ap.add_argument("--ref-tiles", type=int, default=64, help="tiles per step of the reference arm")
    ap.add_argument("--verify-full", action="store_true", help="N = 1: run the CPU oracle on every tile of the zone and compare the whole map")
    ap.add_argument("--no-extras", action="store_true", help="skip the hbm_kernels / blend / patch_sweep legs (N = 1)")
    args = ap.parse_args()
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
