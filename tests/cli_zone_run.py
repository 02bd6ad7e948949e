"""The product's own CLI on the bench zone (not a pytest module): writes a synthetic 10000 x 10000 RGB GeoTIFF (LZW, tiled)
and its truth raster, runs `flair-detect --conf x.yaml -m` (run_pipeline: GeoTIFF decode -> pinned rows -> pipelined
fb_detect_zone_shard with the fused confusion matrix -> LZW BigTIFF encode) twice and prints the pipeline's own figures:
Mpx/s including file I/O, Mpx/s of read + detect, and the share of the wall time the GPU was busy (zone time of the
resident loop / wall time). Also runs `flair --conf` (patch predict + metrics) on 50 synthetic 512 x 512 patches, the
size of the reference's toy CSV, and prints its patches/s including TIFF read and LZW write.

    python tests/cli_zone_run.py [size] > profiles/r02_cli_runs.txt
"""
import re
import sys
import tempfile
import time
from pathlib import Path
from types import SimpleNamespace

import torch
import yaml

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from flair1_b200 import geotiff as gt  # noqa: E402
from flair1_b200.zone_detect import main as zmain  # noqa: E402
from flair1_b200.zone_detect.utils import read_config  # noqa: E402
from oracle import synth  # noqa: E402

SIZE = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
CLASSES15 = {i + 1: [1 if i < 12 else 0, f"class{i + 1}"] for i in range(15)}
tmp = Path(tempfile.mkdtemp(prefix="fb_cli_"))
W = H = SIZE
t0 = time.time()
raster = synth.synth_raster(3, H, W, seed=1)
truth = synth.synth_mask(raster, 15, 3)
d = tmp / "D001_2021" / "Z1_UU"
d.mkdir(parents=True)
tags = gt.georef_tags(800000.0, 6500000.0 + H * 0.2, 0.2, 0.2)
gt.write(d / "zone.tif", raster, geo_tags=tags, compress="lzw", tiled=True, blocksize=512)
gt.write(d / "truth.tif", truth, geo_tags=tags, compress="lzw", tiled=True, blocksize=512)
print(f"synthetic zone {W} x {H}: {(d / 'zone.tif').stat().st_size / 1e6:.0f} MB LZW GeoTIFF + truth, written in {time.time() - t0:.1f} s")
sd = synth.cached_checkpoint(3, 15)
torch.save(sd, tmp / "weights.pth")
means, stds = synth.FLAIR_MEANS[:3], synth.FLAIR_STDS[:3]
cfg = {"output_path": str(tmp / "out"), "output_name": "pred_zone", "input_img_path": str(d / "zone.tif"),
       "truth_path": str(d / "truth.tif"), "channels": [1, 2, 3], "img_pixels_detection": 512, "margin": 128,
       "output_type": "argmax", "n_classes": 15, "model_weights": str(tmp / "weights.pth"),
       "model_framework": {"model_provider": "SegmentationModelsPytorch", "HuggingFace": {"org_model": None},
                           "SegmentationModelsPytorch": {"encoder_decoder": "resnet34_unet"}},
       "batch_size": 4, "use_gpu": True, "num_worker": 2, "write_dataframe": False,
       "norma_task": [{"norm_type": "custom", "norm_means": means, "norm_stds": stds}], "classes": CLASSES15}
conf = tmp / "detect.yaml"
conf.write_text(yaml.safe_dump(cfg))
args = SimpleNamespace(conf=str(conf), metrics=True, batch_mode=False, compare=False)
for run in range(2):
    t = time.time()
    res = zmain.run_pipeline(read_config(args), torch.device("cuda", 0), True)
    wall = time.time() - t
    print(f"flair-detect -m, run {run}: wall {wall:.2f} s; incl. file I/O {res['mpx_per_s_incl_io'][0]:.1f} Mpx/s; "
          f"read + detect {res['mpx_per_s_read_detect'][0]:.1f} Mpx/s; GPU busy ~{100 * (W * H / 1e6 / 1000.0) / wall:.0f} % of the wall time "
          f"(zone at ~1000 Mpx/s resident); mIoU {res['metrics'][0][next(iter(res['metrics'][0]))]['Avg_metrics'][0]:.4f}")
