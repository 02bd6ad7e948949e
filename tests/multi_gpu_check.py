"""Multi-GPU check of the real pipeline (not a pytest module; needs >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
        tests/multi_gpu_check.py > gpurun_out/multi_gpu_check.log 2>&1

Every rank runs flair-detect (`run_pipeline`, `-m`, then `-c -m` over exact-clipping / average / max, then `-c -m` with
an overlapping stride) on the same synthetic zone sharded over the ranks: the `-m` run goes through the pipelined
fb_detect_zone_shard with tile-range sharding and the shared-memory output map, the confusion matrices are summed by
the library's own NCCL communicator (fb_allreduce_confusion), the compare runs keep whole tile rows per rank. Rank 0
then repeats everything alone on one context (`fb_detect_strip` over the whole tile table, blended stitchings over the
whole table) and requires the outputs to be identical byte for byte: class map, confidence band, confusion matrix,
per-patch metrics -- also where a tile's metric window reaches into the rows of the neighbouring rank (stride below
the interior) -- ("allreduced confmat == single-GPU confmat bit-exact")."""
import json
import os
import sys
import tempfile
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch
import torch.distributed as dist
import yaml

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from flair1_b200 import geotiff as gt  # noqa: E402
from flair1_b200.zone_detect import main as zmain  # noqa: E402
from flair1_b200.zone_detect.slicing_job import tile_table, tile_windows  # noqa: E402
from flair1_b200.zone_detect.utils import read_config  # noqa: E402
from oracle import synth  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
CLASSES15 = {i + 1: [1 if i < 12 else 0, f"class{i + 1}"] for i in range(15)}
W, H, T, M = 3000, 2600, 512, 128
sd = synth.cached_checkpoint(3, 15)

box = [None]
if rank == 0:
    box[0] = tempfile.mkdtemp(prefix="fb_mgpu_")
dist.broadcast_object_list(box, src=0)
tmp = Path(box[0])
if rank == 0:
    raster = synth.synth_raster(3, H, W, seed=77)
    truth = synth.synth_mask(raster, 15, 3)
    d = tmp / "D001_2021" / "Z1_UU"
    d.mkdir(parents=True)
    tags = gt.georef_tags(800000.0, 6500000.0 + H * 0.2, 0.2, 0.2)
    gt.write(d / "zone.tif", raster, geo_tags=tags, compress="lzw", tiled=True, blocksize=256)
    gt.write(d / "truth.tif", truth, geo_tags=tags, compress="deflate", tiled=False, blocksize=64)
    torch.save(sd, tmp / "weights.pth")
dist.barrier()
d = tmp / "D001_2021" / "Z1_UU"
means, stds = synth.FLAIR_MEANS[:3], synth.FLAIR_STDS[:3]
cfg = {"output_path": str(tmp / "out"), "output_name": "pred", "input_img_path": str(d / "zone.tif"),
       "truth_path": str(d / "truth.tif"), "channels": [1, 2, 3], "img_pixels_detection": T, "margin": M,
       "output_type": "argmax", "n_classes": 15, "model_weights": str(tmp / "weights.pth"),
       "model_framework": {"model_provider": "SegmentationModelsPytorch", "HuggingFace": {"org_model": None},
                           "SegmentationModelsPytorch": {"encoder_decoder": "resnet34_unet"}},
       "batch_size": 4, "use_gpu": True, "num_worker": 2, "write_dataframe": False,
       "norma_task": [{"norm_type": "custom", "norm_means": means, "norm_stds": stds}], "classes": CLASSES15,
       "overlap_strat": False,
       "strategies": {"tiling": {"enabled": False, "size_range": [], "stride_range": []},
                      "stitching": {"enabled": True, "methods": ["exact-clipping", "average", "max"], "margin": [0.25]},
                      "padding_overall": None}}
conf = tmp / f"detect_{rank}.yaml"
conf.write_text(yaml.safe_dump(cfg))
res = zmain.run_pipeline(read_config(SimpleNamespace(conf=str(conf), metrics=True, batch_mode=False, compare=False)), dev, True)
res_c = zmain.run_pipeline(read_config(SimpleNamespace(conf=str(conf), metrics=True, batch_mode=False, compare=True)), dev, True)
# stride 192 < interior 256: the margin-cropped windows of neighbouring tile rows overlap, so a window at a shard
# boundary reaches into rows the other rank owns
cfg_s = dict(cfg, overlap_strat=True,
             strategies={"tiling": {"enabled": False, "size_range": [], "stride_range": [0.375]},
                         "stitching": {"enabled": True, "methods": ["exact-clipping"], "margin": [0.25]}, "padding_overall": None})
conf_s = tmp / f"detect_s_{rank}.yaml"
conf_s.write_text(yaml.safe_dump(cfg_s))
res_s = zmain.run_pipeline(read_config(SimpleNamespace(conf=str(conf_s), metrics=True, batch_mode=False, compare=True)), dev, True)
dist.barrier()

if rank == 0:
    import flair1_b200._native as nat
    from flair1_b200.zone_detect.metrics import metrics_from_confmat
    raster = gt.read(d / "zone.tif")
    truth = gt.read(d / "truth.tif")[0]
    ctx = nat.Context(local)
    ctx.load_weights(sd, 3, 15)
    ctx.set_norm("custom", means, stds)
    ctx.set_raster(torch.from_numpy(raster).to(dev), [0, 1, 2], W, H)
    tiles, wins = tile_table(W, H, T, M), tile_windows(W, H, T, M)
    truth_dev = torch.from_numpy(truth - np.uint8(1)).to(dev)
    cls = torch.zeros((H, W), dtype=torch.uint8, device=dev)
    cnf = torch.zeros((H, W), dtype=torch.uint8, device=dev)
    cm_tiles = ctx.detect_strip_metrics(tiles, wins, T, 37, cls, cnf, W, 0, truth_dev).cpu().numpy()
    cm = ctx.confusion(cls, truth_dev, 15, truth_sub=0).cpu().numpy()
    got = gt.read(res["outputs"][0])
    checks = {"tiles": len(tiles), "world": world,
              "class map == single GPU": bool(np.array_equal(got[0], cls.cpu().numpy())),
              "confidence band == single GPU": bool(np.array_equal(got[1], cnf.cpu().numpy())),
              "all-reduced confusion matrix == single GPU": bool(np.array_equal(res["confmat"], cm))}
    method = "size=512_stride=256_margin=128_padding=no-padding_stitching=exact-clipping"
    ref_patch = [metrics_from_confmat(c, cfg, f"{method}_{int(w[2])}_{int(w[3])}") for w, c in zip(wins, cm_tiles)]
    checks["per-patch metrics == single GPU"] = json.loads(json.dumps(res_c["patch_metrics"][method])) == json.loads(json.dumps(ref_patch))
    tiles_s, wins_s = tile_table(W, H, T, M, 192), tile_windows(W, H, T, M, 192)
    cls_s = torch.zeros((H, W), dtype=torch.uint8, device=dev)
    cnf_s = torch.zeros((H, W), dtype=torch.uint8, device=dev)
    cm_tiles_s = ctx.detect_strip_metrics(tiles_s, wins_s, T, 37, cls_s, cnf_s, W, 0, truth_dev).cpu().numpy()
    method_s = "size=512_stride=192_margin=128_padding=no-padding_stitching=exact-clipping"
    ref_patch_s = [metrics_from_confmat(c, cfg, f"{method_s}_{int(w[2])}_{int(w[3])}") for w, c in zip(wins_s, cm_tiles_s)]
    checks["stride 192: per-patch metrics == single GPU"] = json.loads(json.dumps(res_s["patch_metrics"][method_s])) == json.loads(json.dumps(ref_patch_s))
    checks["stride 192: class map == single GPU"] = bool(np.array_equal(gt.read(res_s["outputs"][0])[0], cls_s.cpu().numpy()))
    for path in res_c["outputs"]:
        stitch = Path(path).stem.split("stitching=")[1]
        g = gt.read(path)
        c2 = torch.zeros((H, W), dtype=torch.uint8, device=dev)
        f2 = torch.zeros((H, W), dtype=torch.uint8, device=dev)
        if stitch == "exact-clipping":
            ctx.detect_strip(tiles, T, 37, c2, f2, W, 0)
            checks[f"compare grid {stitch}: class map == single GPU"] = bool(np.array_equal(g[0], c2.cpu().numpy()))
        else:
            acc, wsum = ctx.blend_buffers(stitch, H, W)
            ctx.blend_strip(tiles, T, 37, stitch, acc, wsum, W, 0)
            ctx.blend_finalize(stitch, acc, wsum, c2, f2)
            same = float((g[0] == c2.cpu().numpy()).mean())
            # fp32 atomics: the order of the sums is not fixed, exact ties may flip ("max" is order-free: bit-exact)
            checks[f"compare grid {stitch}: class map agreement with single GPU"] = same
            assert same >= (1.0 if stitch == "max" else 0.9999), (stitch, same)
    print(json.dumps(checks, indent=1))
    assert all(v is True or not isinstance(v, bool) for v in checks.values()), checks
    print("multi-GPU check OK")
    ctx.close()
dist.barrier()
dist.destroy_process_group()
