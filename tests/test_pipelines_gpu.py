"""End-to-end drop-in tests on the GPU: the `flair-detect` and `flair` entry points fed with YAML-shaped
configs, GeoTIFF files and .pth/.ckpt checkpoints, checked against the CPU oracle."""
import json
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import yaml

pytestmark = pytest.mark.gpu

CLASSES19 = {i + 1: [0 if i + 1 in (15, 16, 17, 19) else 1, f"class{i + 1}"] for i in range(19)}
CLASSES15 = {i + 1: [1 if i < 12 else 0, f"class{i + 1}"] for i in range(15)}


def _write_zone(tmp_path, W, H, seed):
    from flair1_b200 import geotiff as gt
    from oracle import synth
    raster = synth.synth_raster(3, H, W, seed=seed)
    truth = synth.synth_mask(raster, 15, 3)
    d = tmp_path / "D001_2021" / "Z1_UU"
    d.mkdir(parents=True)
    tags = gt.georef_tags(800000.0, 6500000.0 + H * 0.2, 0.2, 0.2)
    gt.write(d / "zone.tif", raster, geo_tags=tags, compress="lzw", tiled=True, blocksize=256)
    gt.write(d / "truth.tif", truth, geo_tags=tags, compress="deflate", tiled=False, blocksize=64)
    return raster, truth, d


def test_flair_detect_cli_end_to_end(tmp_path, trained_3_15):
    """flair-detect --conf x.yaml -m on a 900 x 700 GeoTIFF: 2-band uint8 LZW tiled BigTIFF with the
    input's georeferencing, class map >= 99.9 % equal to the oracle's, metrics JSON equal to sklearn on
    the same class map."""
    from flair1_b200 import geotiff as gt
    from flair1_b200.zone_detect import main as zmain
    from flair1_b200.zone_detect.utils import read_config
    from oracle import synth
    from oracle.metrics_ref import class_IoU, clean_confmat, overall_accuracy, patch_confusion
    from oracle.zone_detect_ref import GeoRaster, run_zone
    sd, model = trained_3_15
    W, H = 900, 700
    raster, truth, d = _write_zone(tmp_path, W, H, seed=11)
    torch.save({"model.seg_model." + k: v for k, v in sd.items()}, tmp_path / "weights.pth")   # prefixed layout
    means, stds = synth.FLAIR_MEANS[:3], synth.FLAIR_STDS[:3]
    cfg = {"output_path": str(tmp_path / "out"), "output_name": "pred_zone", "input_img_path": str(d / "zone.tif"),
           "truth_path": str(d / "truth.tif"), "channels": [1, 2, 3], "img_pixels_detection": 512, "margin": 128,
           "output_type": "argmax", "n_classes": 15, "model_weights": str(tmp_path / "weights.pth"),
           "model_framework": {"model_provider": "SegmentationModelsPytorch", "HuggingFace": {"org_model": None},
                               "SegmentationModelsPytorch": {"encoder_decoder": "resnet34_unet"}},
           "batch_size": 4, "use_gpu": True, "num_worker": 2, "write_dataframe": True,
           "norma_task": [{"norm_type": "custom", "norm_means": means, "norm_stds": stds}], "classes": CLASSES15}
    conf = tmp_path / "detect.yaml"
    conf.write_text(yaml.safe_dump(cfg))
    args = SimpleNamespace(conf=str(conf), metrics=True, batch_mode=False, compare=False)
    config = read_config(args)
    res = zmain.run_pipeline(config, torch.device("cuda", 0), True)
    out_path = Path(res["outputs"][0])
    assert out_path.name == "pred_zone.tif" and (tmp_path / "out" / "pred_zone_slicing_job.csv").exists()
    assert any(p.suffix == ".log" for p in (tmp_path / "out").iterdir())
    info = gt.read_info(out_path)
    assert (info.width, info.height, info.count, info.compression, info.tiled, info.block_w, info.bigtiff) == (W, H, 2, 5, True, 512, True)
    assert info.geo_tags[33922] == gt.read_info(d / "zone.tif").geo_tags[33922]
    got = gt.read(out_path)
    ocfg = dict(cfg)
    ref_cls, ref_conf, _ = run_zone(model, GeoRaster(raster, 800000.0, 6500000.0 + H * 0.2, 0.2), ocfg)
    agree = (got[0] == ref_cls).mean()
    print(f"flair-detect end-to-end agreement {agree * 100:.4f}%")
    assert agree >= 0.999 and (got[1] == ref_conf).mean() >= 0.995
    # metrics: bit-exact confusion matrix w.r.t. sklearn on the SAME class map, same ratios
    cm_ref = patch_confusion(truth - 1, got[0], 15)
    np.testing.assert_array_equal(res["confmat"], cm_ref)
    m = json.loads(Path(res["metrics_json"]).read_text())[0]
    (key, body), = m.items()
    assert key.startswith("size=512_stride=256_margin=128") and Path(res["metrics_json"]).name == "metrics_per-patch_D001_2021_Z1_UU.json"
    cleaned = clean_confmat(cm_ref, CLASSES15)
    assert body["Avg_metrics"][0] == class_IoU(cleaned)[1] and body["Avg_metrics"][1] == overall_accuracy(cleaned)
    # second run never overwrites
    res2 = zmain.run_pipeline(read_config(args), torch.device("cuda", 0), True)
    assert Path(res2["outputs"][0]).name == "pred_zone_1.tif"
    np.testing.assert_array_equal(gt.read(res2["outputs"][0]), got)


def test_flair_detect_class_prob_and_compare_grid(tmp_path, trained_3_15):
    """output_type class_prob -> n_classes uint8 bands; `-c` with strategies.stitching -> one output per
    stitching method (exact-clipping, average, average_weights, max), each >= 99.9 % equal to its oracle."""
    from flair1_b200 import geotiff as gt
    from flair1_b200.zone_detect import main as zmain
    from flair1_b200.zone_detect.utils import read_config
    from oracle import synth
    from oracle.zone_detect_ref import GeoRaster, run_zone, run_zone_blend, run_zone_class_prob
    sd, model = trained_3_15
    W, H = 800, 600
    raster, truth, d = _write_zone(tmp_path, W, H, seed=23)
    torch.save(sd, tmp_path / "weights.pth")
    means, stds = synth.FLAIR_MEANS[:3], synth.FLAIR_STDS[:3]
    cfg = {"output_path": str(tmp_path / "out"), "output_name": "pz", "input_img_path": str(d / "zone.tif"),
           "truth_path": str(d / "truth.tif"), "channels": [1, 2, 3], "img_pixels_detection": 512, "margin": 128,
           "output_type": "class_prob", "n_classes": 15, "model_weights": str(tmp_path / "weights.pth"),
           "model_framework": {"model_provider": "SegmentationModelsPytorch", "HuggingFace": {"org_model": None},
                               "SegmentationModelsPytorch": {"encoder_decoder": "resnet34_unet"}},
           "batch_size": 4, "use_gpu": True, "num_worker": 2, "write_dataframe": False,
           "norma_task": [{"norm_type": "custom", "norm_means": means, "norm_stds": stds}], "classes": CLASSES15}
    georaster = GeoRaster(raster, 800000.0, 6500000.0 + H * 0.2, 0.2)
    conf = tmp_path / "prob.yaml"
    conf.write_text(yaml.safe_dump(cfg))
    res = zmain.run_pipeline(read_config(SimpleNamespace(conf=str(conf), metrics=False, batch_mode=False, compare=False)),
                             torch.device("cuda", 0), True)
    got = gt.read(res["outputs"][0]).astype(np.int32)
    ref = run_zone_class_prob(model, georaster, cfg).astype(np.int32)
    assert got.shape == (15, H, W) and gt.read_info(res["outputs"][0]).count == 15
    assert (np.abs(got - ref) <= 4).mean() >= 0.999

    cfg.update({"output_type": "argmax", "output_name": "cmp", "overlap_strat": False,
                "strategies": {"tiling": {"enabled": False, "size_range": [], "stride_range": []},
                               "stitching": {"enabled": True, "methods": ["exact-clipping", "average", "average_weights", "max"],
                                             "margin": [0.25]},      # a fraction of the tile size (utils.py:87-92)
                               "padding_overall": None}})
    conf2 = tmp_path / "cmp.yaml"
    conf2.write_text(yaml.safe_dump(cfg))
    res = zmain.run_pipeline(read_config(SimpleNamespace(conf=str(conf2), metrics=True, batch_mode=False, compare=True)),
                             torch.device("cuda", 0), True)
    assert len(res["outputs"]) == 4 and len(res["metrics"]) == 4
    for path in res["outputs"]:
        method = Path(path).stem.split("stitching=")[1]
        cls = gt.read(path)[0]
        ref_cls = run_zone(model, georaster, cfg)[0] if method == "exact-clipping" else run_zone_blend(model, georaster, cfg, method)[0]
        agree = (cls == ref_cls).mean()
        print(f"compare grid, stitching={method}: agreement {agree * 100:.4f}%")
        assert agree >= 0.999


def test_flair_detect_rejects_what_it_cannot_do(tmp_path, trained_3_15):
    from flair1_b200.zone_detect.model import load_model
    sd, _ = trained_3_15
    broken = dict(sd)
    broken.pop("encoder.layer3.2.bn1.running_mean")
    torch.save(broken, tmp_path / "broken.pth")
    cfg = {"channels": [1, 2, 3], "n_classes": 15, "model_weights": str(tmp_path / "broken.pth"),
           "model_framework": {"model_provider": "SegmentationModelsPytorch",
                               "SegmentationModelsPytorch": {"encoder_decoder": "resnet34_unet"}}}
    with pytest.raises(RuntimeError, match="Missing key"):      # load_state_dict(strict=True)
        load_model(cfg, 0)
    cfg["model_framework"]["model_provider"] = "HuggingFace"
    with pytest.raises(NotImplementedError):
        load_model(cfg, 0)


def test_flair_predict_and_metrics_end_to_end(tmp_path, trained_3_15):
    """BASELINE config 1 at its stated size: flair --conf x.yaml (predict + metrics) on 50 synthetic 512 x 512 patches
    laid out like csv_toy/flair-1-paths-toy-test.csv (50 rows of image,mask): PRED_*.tif files with 0-based classes,
    metrics/confmat.npy + metrics.json equal to the oracle's on the same files; the log carries the patches/s of the run."""
    from flair1_b200 import geotiff as gt
    from flair1_b200.flair import main as fmain
    from oracle import synth
    from oracle.flair_ref import norm, predict_step
    from oracle.metrics_ref import flair_metrics, patch_confusion
    sd, model = trained_3_15
    means, stds = synth.FLAIR_MEANS[:3], synth.FLAIR_STDS[:3]
    data = tmp_path / "data"
    (data / "img").mkdir(parents=True)
    (data / "msk").mkdir()
    N = 50
    rows, imgs, msks = [], [], []
    for i in range(N):
        img5 = synth.synth_raster(5, 512, 512, seed=100 + i)
        msk = synth.synth_mask(img5[:3], 15, 3)
        msk[:8] = 19                                  # labels outside 1..15 are dropped by the confusion matrix
        ip, mp = data / "img" / f"IMG_{i:06d}.tif", data / "msk" / f"MSK_{i:06d}.tif"
        gt.write(ip, img5, geo_tags=gt.georef_tags(900000.0 + 102.4 * i, 6400000.0, 0.2, 0.2), compress="deflate", tiled=False, blocksize=64)
        gt.write(mp, msk, compress="lzw", tiled=False, blocksize=64)
        rows.append(f"{ip},{mp}")
        imgs.append(img5[:3])
        msks.append(msk)
    csv = tmp_path / "test.csv"
    csv.write_text("\n".join(rows) + "\n")
    torch.save({"state_dict": {**{"model.seg_model." + k: v for k, v in sd.items()}, "criterion.weight": torch.ones(15)}},
               tmp_path / "model.ckpt")
    cfg = {"paths": {"out_folder": str(tmp_path / "exp"), "out_model_name": "run1", "train_csv": None, "val_csv": None,
                     "test_csv": str(csv), "ckpt_model_path": str(tmp_path / "model.ckpt"), "path_metadata_aerial": None},
           "tasks": {"train": False, "train_tasks": {"init_weights_only_from_ckpt": False, "resume_training_from_ckpt": False},
                     "predict": True, "metrics": True, "delete_preds": False},
           "model_framework": {"model_provider": "SegmentationModelsPytorch",
                               "SegmentationModelsPytorch": {"encoder_decoder": "resnet34_unet"}},
           "use_augmentation": False, "use_metadata": False, "channels": [1, 2, 3], "norm_type": "custom",
           "norm_means": means, "norm_stds": stds, "seed": 2022, "batch_size": 4, "classes": CLASSES15,
           "georeferencing_output": True, "cp_csv_and_conf_to_output": True, "accelerator": "gpu", "num_nodes": 1,
           "gpus_per_node": 1, "strategy": "auto", "num_workers": 0}
    conf = tmp_path / "flair.yaml"
    conf.write_text(yaml.safe_dump(cfg))
    old = sys.argv
    sys.argv = ["flair", "--conf", str(conf)]
    try:
        fmain.main()
    finally:
        sys.argv = old
    pred_dir = tmp_path / "exp" / "run1" / "predictions_run1"
    assert (tmp_path / "exp" / "run1" / "flair-compute.log").exists()
    assert (tmp_path / "exp" / "run1" / "used_csv_and_config" / "test.csv").exists()
    import re
    log_text = (tmp_path / "exp" / "run1" / "flair-compute.log").read_text()
    rate = re.search(r"predicted (\d+) patches in ([0-9.]+) s \(([0-9.]+) patches/s", log_text)
    assert rate and int(rate.group(1)) == N
    print(f"flair CLI, {N} patches incl. TIFF read + LZW write: {rate.group(3)} patches/s")
    agree, cms = [], []
    for i in range(N):
        p = pred_dir / f"PRED_IMG_{i:06d}.tif"
        info = gt.read_info(p)
        assert (info.count, info.compression) == (1, 5) and 33922 in info.geo_tags
        got = gt.read(p)[0]
        x = torch.as_tensor(norm(imgs[i], "custom", means, stds), dtype=torch.float)[None]
        ref = predict_step(model, x)[0].numpy().astype(np.uint8)
        agree.append((got == ref).mean())
        cms.append(patch_confusion(msks[i] - 1, got, 15))
    print(f"flair predict agreement {np.mean(agree) * 100:.4f}%")
    assert np.mean(agree) >= 0.999
    confmat = np.load(tmp_path / "exp" / "run1" / "metrics" / "confmat.npy")
    np.testing.assert_array_equal(confmat, np.sum(cms, axis=0))
    m = json.loads((tmp_path / "exp" / "run1" / "metrics" / "metrics.json").read_text())
    ref_m = flair_metrics(np.sum(cms, axis=0), CLASSES15)
    assert m["Avg_metrics"] == [float(v) for v in ref_m["Avg_metrics"]] and m["classes"] == ref_m["classes"]
    assert m["per_class_iou"] == [float(v) for v in ref_m["per_class_iou"]]


def test_flair_predict_with_metadata_end_to_end(tmp_path):
    """BASELINE config 3 through the CLI: flair --conf x.yaml with use_metadata: True and path_metadata_aerial. The JSON
    goes through parsing_metadata (src/flair/tasks_utils.py:158-213) -> 45 floats per patch -> predict_dataset ->
    fb_predict_patches (MLP + broadcast add onto the bottleneck, src/flair/model.py:52-70). PRED files against the oracle's
    FlairModel on the same inputs (>= 99.9 % on a briefly trained 5-band / 13-class checkpoint with the metadata branch),
    and the logits of the same patches within tolerance."""
    from flair1_b200 import geotiff as gt
    from flair1_b200.flair import main as fmain
    from flair1_b200.flair.tasks_utils import parsing_metadata
    import flair1_b200._native as nat
    from oracle import synth
    from oracle.flair_ref import norm, predict_step
    from oracle.unet_smp033 import FlairModel
    sd = synth.cached_checkpoint(5, 13, use_metadata=True)
    model = FlairModel(5, 13, True)
    model.load_state_dict(sd, strict=True)
    model.eval()
    means, stds = synth.FLAIR_MEANS, synth.FLAIR_STDS
    data = tmp_path / "data"
    (data / "img").mkdir(parents=True)
    (data / "msk").mkdir()
    N = 6
    rows, imgs, meta = [], [], {}
    cams = ["UCE-M3-f120-s06", "UCX-2", "UCE"]
    for i in range(N):
        img5 = synth.synth_raster(5, 512, 512, seed=300 + i)
        msk = synth.synth_mask(img5[:3], 13, 3)
        ip, mp = data / "img" / f"IMG_{i:06d}.tif", data / "msk" / f"MSK_{i:06d}.tif"
        gt.write(ip, img5, compress="deflate", tiled=False, blocksize=64)
        gt.write(mp, msk, compress="lzw", tiled=False, blocksize=64)
        rows.append(f"{ip},{mp}")
        imgs.append(img5)
        meta[f"IMG_{i:06d}"] = {"patch_centroid_x": 489212.4 + 51234.5 * i, "patch_centroid_y": 6222222.2 + 77777.7 * i,
                               "patch_centroid_z": 137.5 * i, "camera": cams[i % 3], "date": f"{2018 + i % 4}-{1 + i:02d}-{3 + 4 * i:02d}",
                               "time": f"{8 + i:02d}h{7 * i:02d}"}
    csv = tmp_path / "test.csv"
    csv.write_text("\n".join(rows) + "\n")
    (tmp_path / "meta.json").write_text(json.dumps(meta))
    classes13 = {i + 1: [1, f"class{i + 1}"] for i in range(13)}
    # Lightning checkpoint layout of the flair CLI: model.seg_model.*, model.enc.enc_mlp.*, criterion.weight
    torch.save({"state_dict": {**{"model." + k: v for k, v in sd.items()}, "criterion.weight": torch.ones(13)}}, tmp_path / "model.ckpt")
    cfg = {"paths": {"out_folder": str(tmp_path / "exp"), "out_model_name": "meta", "train_csv": None, "val_csv": None,
                     "test_csv": str(csv), "ckpt_model_path": str(tmp_path / "model.ckpt"), "path_metadata_aerial": str(tmp_path / "meta.json")},
           "tasks": {"train": False, "train_tasks": {"init_weights_only_from_ckpt": False, "resume_training_from_ckpt": False},
                     "predict": True, "metrics": True, "delete_preds": False},
           "model_framework": {"model_provider": "SegmentationModelsPytorch",
                               "SegmentationModelsPytorch": {"encoder_decoder": "resnet34_unet"}},
           "use_augmentation": False, "use_metadata": True, "channels": [1, 2, 3, 4, 5], "norm_type": "custom",
           "norm_means": means, "norm_stds": stds, "seed": 2022, "batch_size": 4, "classes": classes13,
           "georeferencing_output": False, "cp_csv_and_conf_to_output": False, "accelerator": "gpu", "num_nodes": 1,
           "gpus_per_node": 1, "strategy": "auto", "num_workers": 0}
    conf = tmp_path / "flair.yaml"
    conf.write_text(yaml.safe_dump(cfg))
    old = sys.argv
    sys.argv = ["flair", "--conf", str(conf)]
    try:
        fmain.main()
    finally:
        sys.argv = old
    mtd = np.asarray(parsing_metadata([r.split(",")[0] for r in rows], cfg), dtype=np.float32)
    assert mtd.shape == (N, 45)
    x = torch.stack([torch.as_tensor(norm(im, "custom", means, stds), dtype=torch.float) for im in imgs])
    ref_cls = predict_step(model, x, torch.from_numpy(mtd)).numpy().astype(np.uint8)
    pred_dir = tmp_path / "exp" / "meta" / "predictions_meta"
    got = np.stack([gt.read(pred_dir / f"PRED_IMG_{i:06d}.tif")[0] for i in range(N)])
    agree = (got == ref_cls).mean()
    print(f"flair predict with metadata: agreement {agree * 100:.4f}%, classes present {np.unique(ref_cls).size}")
    assert agree >= 0.999
    # the metadata really matters for this checkpoint: other metadata, other logits (guards against a silently ignored MTD)
    with torch.no_grad():
        ref_logits = model(x[:2], torch.from_numpy(mtd[:2]))
        other = model(x[:2], torch.from_numpy(mtd[2:4]))
    assert (ref_logits - other).abs().max() > 1e-3
    # logits of the same patches through the library, with the vectors parsing_metadata produced
    ctx = nat.Context(0)
    ctx.load_weights({k.replace("seg_model.", "", 1) if k.startswith("seg_model.") else k: v for k, v in sd.items()}, 5, 13, use_metadata=True)
    ctx.set_norm("custom", means, stds)
    raster = torch.from_numpy(np.concatenate(imgs[:2], axis=1)).cuda()
    ctx.set_raster(raster, [0, 1, 2, 3, 4], 512, 1024)
    lg = ctx.forward_tiles(np.array([[0, 0], [0, 512]], np.int32), 512, metadata=mtd[:2]).cpu().permute(0, 3, 1, 2)[:, :13]
    rel = (lg - ref_logits).abs().max().item() / ref_logits.abs().max().item()
    print(f"flair predict with metadata: logits rel err {rel:.4e}")
    assert rel <= 2e-2
    ctx.close()
    m = json.loads((tmp_path / "exp" / "meta" / "metrics" / "metrics.json").read_text())
    assert len(m["per_class_iou"]) == 13


def test_load_checkpoint_class_count_surgery(tmp_path, trained_3_15):
    """src/flair/main.py:106-138: a 15-class checkpoint loaded into a 13-class config gets its head
    truncated and zeroed (so every logit is 0 and argmax is class 0 everywhere)."""
    from flair1_b200.flair.model import FLAIR_ModelFactory, load_checkpoint
    from oracle import synth
    sd, _ = trained_3_15
    torch.save({**sd, "criterion.weight": torch.ones(15)}, tmp_path / "m.pth")
    classes13 = {i + 1: [1, f"c{i + 1}"] for i in range(13)}
    cfg = {"paths": {"ckpt_model_path": str(tmp_path / "m.pth")}, "classes": classes13, "channels": [1, 2, 3], "use_metadata": False,
           "model_framework": {"model_provider": "SegmentationModelsPytorch",
                               "SegmentationModelsPytorch": {"encoder_decoder": "resnet34_unet"}}}
    fac = FLAIR_ModelFactory(cfg, 0)
    load_checkpoint(cfg, fac)
    assert fac.loaded
    ctx = fac.seg_model
    ctx.set_norm("custom", synth.FLAIR_MEANS[:3], synth.FLAIR_STDS[:3])
    patches = torch.from_numpy(synth.synth_raster(3, 512, 512, seed=5))[None].cuda()
    assert int(ctx.predict_patches(patches, 512, 1).max()) == 0
    cfg["paths"]["ckpt_model_path"] = str(tmp_path / "missing.pth")
    fac2 = FLAIR_ModelFactory(cfg, 0)
    load_checkpoint(cfg, fac2)                 # "Invalid checkpoint file path." and nothing loaded
    assert not fac2.loaded
    with pytest.raises(SystemExit):
        load_checkpoint(cfg, fac2, exit_on_fail=True)


def test_per_patch_confusion_is_bit_exact(ctx, trained_3_15):
    """fb_detect_strip_metrics (compute_metrics_patch of the compare loop, test/metrics.py:124-163): tile i's own
    arg-max over its margin-cropped window against truth - 1, bit-exact w.r.t. numpy on the same logits, incl. the
    clamped last row / column whose windows overlap other tiles; the maps equal the plain detect_strip."""
    from flair1_b200.zone_detect.slicing_job import tile_table, tile_windows
    from oracle import synth
    from oracle.metrics_ref import patch_confusion
    sd, _ = trained_3_15
    W, H, T, margin = 900, 700, 512, 128
    raster = synth.synth_raster(3, H, W, seed=5)
    truth = synth.synth_mask(raster, 15, 3)
    truth[::7, ::5] = 0          # wraps to 255 after - 1: dropped
    truth[3::11, 2::9] = 17      # 16 after - 1: out of range, dropped
    ctx.load_weights(sd, 3, 15)
    ctx.set_norm("custom", synth.FLAIR_MEANS[:3], synth.FLAIR_STDS[:3])
    ctx.set_raster(torch.from_numpy(raster).cuda(), [0, 1, 2], W, H)
    tiles, wins = tile_table(W, H, T, margin), tile_windows(W, H, T, margin)
    cls = torch.zeros((H, W), dtype=torch.uint8, device="cuda")
    conf = torch.zeros((H, W), dtype=torch.uint8, device="cuda")
    cm = ctx.detect_strip_metrics(tiles, wins, T, 5, cls, conf, W, 0, torch.from_numpy(truth).cuda(), truth_sub=1).cpu().numpy()
    cls2 = torch.zeros_like(cls)
    conf2 = torch.zeros_like(conf)
    ctx.detect_strip(tiles, T, 5, cls2, conf2, W, 0)
    assert torch.equal(cls, cls2) and torch.equal(conf, conf2)
    logits = ctx.forward_tiles(np.ascontiguousarray(tiles[:, :2]), T).cpu().numpy()[..., :15]
    overlap = 0
    for i, (x0, y0, a, b, c, d) in enumerate(wins):
        pred = logits[i].argmax(-1)[b - y0:d - y0, a - x0:c - x0].astype(np.uint8)
        ref = patch_confusion(truth[b:d, a:c] - np.uint8(1), pred, 15)
        np.testing.assert_array_equal(cm[i], ref)
        overlap += (d - b) * (c - a)
    assert overlap > W * H     # the clamped windows really do overlap


def test_flair_detect_batch_mode(tmp_path, trained_3_15):
    """flair-detect -c -m -b over a department directory (main.py:440-497): two zones (one directory without
    imagery and one without ground truth are skipped), two stitching methods; metrics.json holds one entry per
    method with the metrics of the confusion matrices summed over the zones, computed from the prediction rasters
    the run wrote; per-patch metrics of the compare loop land in metrics_per-patch_<dpt>_<zone>.json."""
    from flair1_b200 import geotiff as gt
    from flair1_b200.zone_detect import main as zmain
    from flair1_b200.zone_detect.utils import read_config
    from oracle import synth
    from oracle.metrics_ref import class_IoU, clean_confmat, overall_accuracy, patch_confusion
    sd, _ = trained_3_15
    torch.save(sd, tmp_path / "weights.pth")
    dpt_in, dpt_gt = tmp_path / "images" / "032_2019", tmp_path / "labels" / "032_2019"
    truths = {}
    for zi, (zone, (W, H)) in enumerate({"UA-zone_1": (700, 600), "UN_S-12_3": (640, 520)}.items()):
        raster = synth.synth_raster(3, H, W, seed=40 + zi)
        truths[zone] = synth.synth_mask(raster, 15, 3)
        (dpt_in / zone).mkdir(parents=True)
        (dpt_gt / zone).mkdir(parents=True)
        tags = gt.georef_tags(800000.0, 6500000.0 + H * 0.2, 0.2, 0.2)
        gt.write(dpt_in / zone / f"032_2019_{zone}_RGB.tif", raster, geo_tags=tags, compress="lzw", tiled=True, blocksize=256)
        gt.write(dpt_gt / zone / f"032_2019_{zone}_MSK.tif", truths[zone], geo_tags=tags, compress="deflate", tiled=False, blocksize=64)
    (dpt_in / "no_imagery").mkdir()
    (dpt_in / "no_truth").mkdir()
    gt.write(dpt_in / "no_truth" / "032_2019_no_truth_RGB.tif", synth.synth_raster(3, 64, 64, seed=1), compress="lzw", tiled=False, blocksize=64)
    first = dpt_in / "UA-zone_1" / "032_2019_UA-zone_1_RGB.tif"
    cfg = {"output_path": str(tmp_path / "out"), "output_name": "unused", "input_img_path": str(first), "input_path": str(dpt_in),
           "truth_path": str(dpt_gt / "UA-zone_1" / "032_2019_UA-zone_1_MSK.tif"), "truth_root": str(tmp_path / "labels"),
           "data_type": "RGB", "model_name": "resnet34-unet",
           "channels": [1, 2, 3], "img_pixels_detection": 256, "margin": 32, "output_type": "argmax", "n_classes": 15,
           "model_weights": str(tmp_path / "weights.pth"),
           "model_framework": {"model_provider": "SegmentationModelsPytorch", "HuggingFace": {"org_model": None},
                               "SegmentationModelsPytorch": {"encoder_decoder": "resnet34_unet"}},
           "batch_size": 4, "use_gpu": True, "num_worker": 2, "write_dataframe": False,
           "norma_task": [{"norm_type": "custom", "norm_means": synth.FLAIR_MEANS[:3], "norm_stds": synth.FLAIR_STDS[:3]}],
           "classes": CLASSES15, "overlap_strat": False,
           "strategies": {"tiling": {"enabled": False, "size_range": [], "stride_range": []},
                          "stitching": {"enabled": True, "methods": ["exact-clipping", "max"], "margin": [0.125]},
                          "padding_overall": None}}
    conf = tmp_path / "batch.yaml"
    conf.write_text(yaml.safe_dump(cfg))
    config = read_config(SimpleNamespace(conf=str(conf), metrics=True, batch_mode=True, compare=True))
    gt_dpt = Path(config["truth_root"]) / Path(config["truth_path"]).parts[-3]
    assert gt_dpt == dpt_gt
    out = zmain.batch_metrics_pipeline(config, gt_dpt, torch.device("cuda", 0), True)
    metrics = json.loads(Path(out).read_text())
    assert Path(out).name == "metrics.json" and len(metrics) == 2
    preds = sorted((tmp_path / "out").rglob("*.tif"))
    assert len(preds) == 4 and all("-ARGMAX-S_size=256_stride=192_margin=32_padding=no-padding_stitching=" in p.name for p in preds)
    for entry in metrics:
        stitch = entry["Parameters values"][5]
        assert entry["Parameters values"][:5] == ["resnet34-unet", 256, 192, 32, "no-padding"] and stitch in ("exact-clipping", "max")
        cm = np.zeros((15, 15))
        for zone, truth in truths.items():
            (pp,) = [p for p in preds if f"_{zone}_RGB" in p.name and p.name.endswith(f"stitching={stitch}.tif")]
            cm += patch_confusion(truth - np.uint8(1), gt.read(pp)[0], 15)
        cleaned = clean_confmat(cm, CLASSES15)
        assert entry["Avg_metrics"][0] == class_IoU(cleaned)[1] and entry["Avg_metrics"][1] == overall_accuracy(cleaned)
        assert entry["Avg_metrics"][3] > 0 and len(entry["per_class_iou"]) == 12
    per_patch = sorted((tmp_path / "out").rglob("metrics_per-patch_032_2019_*.json"))
    assert len(per_patch) == 2
    body = json.loads(per_patch[0].read_text())
    from flair1_b200.zone_detect.slicing_job import tile_table
    assert len(body) == len(tile_table(700, 600, 256, 32))
    (key, val), = body[0].items()
    # first tile of the write order = bottom-left corner: window rows [600 - 192, 600)
    assert key == "size=256_stride=192_margin=32_padding=no-padding_stitching=exact-clipping_0_408" and len(val["Avg_metrics"]) == 3
