#!/usr/bin/env python
"""zone_detect throughput benchmark (BASELINE.json metric: zone_detect Mpixels/s, ResNet34-UNet).

    python bench.py --gpus N --steps K --warmup W            # our arm: CUDA path through the C ABI
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the CPU path (oracle port)

One step = one pass of the whole hot path over one synthetic zone: tile extraction + normalisation,
U-Net forward, softmax/argmax, margin clipping and stitching into the uint8 class map.
  N = 1 : BASELINE.json configs[1] -- 10000 x 10000 RGB raster, 512 px tiles, margin 128, 15 classes
          (1600 tiles).
  N > 1 : configs[3] -- a 40000 x 40000 raster whose 157 tile rows are sharded across the N ranks (one
          process per GPU, halo rows re-read, no data-path collective); the per-rank confusion matrices
          are summed with one NCCL all-reduce per step.
`value` is timed with inputs resident in HBM; `e2e` is the same step through fb_detect_zone_host with
pinned HOST buffers (raster upload and class-map download inside the timed region).
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


def load_oracle_model():
    from oracle import synth
    from oracle.unet_smp033 import Unet
    sd = synth.cached_checkpoint(BANDS, NCLS)
    m = Unet(BANDS, NCLS)
    m.load_state_dict(sd, strict=True)
    m.eval()
    return sd, m


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
    from oracle import synth
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
    ms_per_step = 1e3 * (W * H / 1e6) / value  # one step of OUR arm = the whole zone
    line = {
        "impl": "reference", "metric": "zone_detect Mpixels/s", "value": value, "unit": "Mpixels/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, W, H, len(tiles)),
        "cpu_baseline": {"value": value, "unit": "Mpixels/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{per_step} tiles per step x {args.steps} steps of the same zone (batch 4), linear in tiles"},
        "e2e": {"value": value, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def workload_config(n_gpus: int, W: int, H: int, n_tiles: int) -> dict:
    return {"workload": f"zone_detect {W}x{H} RGB uint8 raster, {TILE}px tiles, margin {MARGIN}, {NCLS} classes, "
                        f"ResNet34-UNet, {n_tiles} tiles" + ("" if n_gpus == 1 else f", tile rows sharded over {n_gpus} GPUs"),
            "raster": [W, H], "tile": TILE, "margin": MARGIN, "n_classes": NCLS, "bands": BANDS, "tiles": n_tiles,
            "l2_policy": "inputs larger than L2 (raster strip >= 300 MB, activations > 2 GB per batch)"}


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args) -> int:
    import flair1_b200._native as nat
    from flair1_b200.zone_detect.slicing_job import split_rows_across_ranks, tile_table

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
    shard = split_rows_across_ranks(tiles_all, world)[rank]
    tiles = tiles_all[shard]
    # raster rows this rank needs (tile rows + halo), class-map rows this rank writes
    ry0, ry1 = max(int(tiles[:, 1].min()), 0), min(int(tiles[:, 1].max()) + TILE, H)
    my0, my1 = int(tiles[:, 3].min()), int(tiles[:, 5].max())
    sd, model = load_oracle_model() if rank == 0 else (None, None)
    if rank != 0:
        from oracle import synth
        sd = synth.cached_checkpoint(BANDS, NCLS)

    ctx = nat.Context(local)
    ctx.load_weights(sd, BANDS, NCLS)
    ctx.set_norm("custom", MEANS, STDS)
    raster_dev = synth_rows_gpu(W, H, ry0, ry1, 1, dev)
    truth_dev = torch.randint(1, 20, (my1 - my0, W), dtype=torch.uint8, device=dev,
                              generator=torch.Generator(device=dev).manual_seed(100 + rank))
    cls_dev = torch.zeros((my1 - my0, W), dtype=torch.uint8, device=dev)
    conf_dev = torch.zeros((my1 - my0, W), dtype=torch.uint8, device=dev)
    cm_dev = torch.zeros((NCLS, NCLS), dtype=torch.int64, device=dev)
    raster_host = torch.empty(raster_dev.shape, dtype=torch.uint8, pin_memory=True)
    raster_host.copy_(raster_dev)
    cls_host = torch.empty((my1 - my0, W), dtype=torch.uint8, pin_memory=True)
    conf_host = torch.empty((my1 - my0, W), dtype=torch.uint8, pin_memory=True)
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        ctx.set_raster(raster_dev, [0, 1, 2], W, H, row0=ry0)
        ctx.detect_strip(tiles, TILE, args.batch, cls_dev, conf_dev, W, my0)
        cm_dev.zero_()
        ctx.confusion(cls_dev, truth_dev, NCLS, truth_sub=1, out=cm_dev)
        if dist is not None:
            dist.all_reduce(cm_dev)

    def step_e2e():
        t = time.perf_counter()
        ctx.detect_zone_host(raster_host, [0, 1, 2], W, H, ry0, nat.FB_LAYOUT_CHW, tiles, TILE, args.batch,
                             cls_host, conf_host, W, my0, my1 - my0)
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
    for _ in range(min(args.warmup, 2)):
        step_e2e()
    ms_e2e, _, _ = timed(step_e2e, args.steps)

    lt = torch.tensor([launches, raster_host.numel(), cls_host.numel() + conf_host.numel()], dtype=torch.int64, device=dev)
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
        # conv kernels of the slowest rank: its tiles * 63.569 GFLOP per step over its summed conv time
        max_tiles = max(len(s) for s in split_rows_across_ranks(tiles_all, world))
        conv_tflops = float(rate[0].item())
        gflop_per_tile = -float(rate[1].item())
        full_tile_tflops = max_tiles * GFLOP_PER_TILE * args.steps / (float(conv_ms.item()) / 1e3) / 1e3
        # CPU baseline + agreement on a bounded sample of rank 0's tiles
        torch.set_num_threads(os.cpu_count() or 1)
        ny = len(np.unique(tiles[:, 1]))
        nx = len(tiles) // ny
        # the CPU baseline is an N=1 figure; at N>1 only a small sample is run, for the agreement check
        idx = sample_tile_indices(len(tiles), nx, ny, args.cpu_tiles if world == 1 else 24)
        rh = raster_host.numpy()
        cpu_s, cpu_out, cpu_px = cpu_zone_sample(model, lambda r0, r1: rh[:, r0 - ry0:r1 - ry0], W, H, tiles, idx)
        cls_np = cls_host.numpy()
        same = tot = 0
        for i, patch in cpu_out.items():
            x0, y0, wx0, wy0, wx1, wy1 = (int(v) for v in tiles[i])
            ref = patch[wy0 - (y0 + MARGIN):wy1 - (y0 + MARGIN), wx0 - (x0 + MARGIN):wx1 - (x0 + MARGIN)]
            got = cls_np[wy0 - my0:wy1 - my0, wx0:wx1]
            same += int((ref == got).sum())
            tot += ref.size
        line = {
            "metric": "zone_detect Mpixels/s", "value": value, "unit": "Mpixels/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {**workload_config(world, W, H, len(tiles_all)), "batch_tiles": args.batch},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mpixels/s", "h2d_bytes_per_step": int(lt[1].item()),
                    "d2h_bytes_per_step": int(lt[2].item())},
            "gpu_launches": int(lt[0].item()),
            "roofline": {"bound": "tensor", "achieved": conv_tflops, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": conv_tflops / peak_tf,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch of the family's top kernel
                         # (conv_igemm2_kernel<256>, a layer3 3x3 conv on 148 tiles: 77.6 MB in, 77.6 MB out, 1.2 MB of
                         # weights; part of the output stays in the 126 MB L2 for the next layer; 114.2 us, tensor pipe
                         # 88 % active), from the `ncu --set full` capture in profiles/r01_ncu_full_igemm2_256_v21.csv
                         "traffic": 114.74e6, "traffic_unit": "bytes per launch (conv_igemm2_kernel<256>, 148 tiles)",
                         "peak_source": peak_src,
                         "kernel": "conv_igemm2_kernel / conv_igemm_kernel / conv_halo_kernel (47 launches per batch); achieved = algorithmic FLOPs "
                                   "(2*MAC of the direct conv) of the outputs actually computed / summed conv time",
                         "gflop_per_tile_computed": gflop_per_tile, "gflop_per_tile_full": GFLOP_PER_TILE,
                         "note": "the decoder skips outputs that can only reach the cropped margin (bit-identical class map, "
                                 "csrc/tile_need.cuh); full_tile_equivalent_tflops counts 63.569 GFLOP per tile instead",
                         "full_tile_equivalent_tflops": full_tile_tflops,
                         "conv_share_of_step": float(conv_ms.item()) / ms_total},
            "cpu_baseline": {"value": cpu_px / cpu_s / 1e6, "unit": "Mpixels/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{len(idx)} of rank 0's {len(tiles)} tiles (batch 4, fp32 torch CPU oracle), linear in tiles"},
            "argmax_agreement_pct": 100.0 * same / max(tot, 1),
            "stage_ms_per_step": {k: v / args.steps for k, v in prof.items()},
        }
        if world > 1:
            line["cpu_baseline"] = None  # reported at N=1 only (torchrun pins OMP threads; see the N=1 line)
        emit(line)
    if dist is not None:
        dist.barrier()
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
    ap.add_argument("--batch", type=int, default=148, help="tiles per forward pass (148 = one per SM; 74 / 111 / 148 measured 906 / 907 / 922 Mpx/s)")
    ap.add_argument("--cpu-tiles", type=int, default=160, help="tiles of the zone timed on the host cores (cpu_baseline)")
    ap.add_argument("--ref-tiles", type=int, default=64, help="tiles per step of the reference arm")
    args = ap.parse_args()
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
