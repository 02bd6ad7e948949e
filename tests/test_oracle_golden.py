"""The oracle against vectors produced by the reference's own code (tests/golden/make_golden.py),
and the product's integer tile grid against the same vectors. CPU only."""
import json
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import flair_ref, metrics_ref, tiles_ref, zone_detect_ref as zref
from flair1_b200.zone_detect import slicing_job, tiles as ptiles


def _case_rows(case):
    spec = case["spec"]
    W, H, res, ox, oy = spec
    rows = zref.slice_extent((ox, oy, ox + W * res, oy + H * res), (res, res), case["size"], case["margin"], case["stride"])
    return np.array([[r["left"], r["bottom"], r["right"], r["top"], *r["geometry"]] for r in rows], dtype=np.float64)


def test_slice_extent_oracle_matches_reference(golden):
    for case in golden["slice_extent"]:
        arr = _case_rows(case)
        assert len(arr) == case["n"], case["spec"]
        if "rows" in case:
            np.testing.assert_allclose(arr, np.array(case["rows"]), rtol=0, atol=1e-6)
        else:
            np.testing.assert_allclose(arr[:50], np.array(case["rows_head"]), rtol=0, atol=1e-6)
            np.testing.assert_allclose(arr[-50:], np.array(case["rows_tail"]), rtol=0, atol=1e-6)
            np.testing.assert_allclose(arr.sum(axis=0), np.array(case["rows_sum"]), rtol=1e-12)


def test_integer_tile_grid_matches_reference(golden):
    """The product's pixel-space grid reproduces the reference dataframe (interior boxes, tile
    geometry, order, de-duplication) for every pixel-aligned golden raster."""
    for case in golden["slice_extent"]:
        W, H, res, ox, oy = case["spec"]
        size, margin = case["size"], case["margin"]
        t = slicing_job.tile_table(int(W), int(H), size, margin, case["stride"])
        ints = slicing_job.tile_interiors(int(W), int(H), size, margin, case["stride"])
        assert len(t) == case["n"], (case["spec"], size, margin)
        ref = np.array(case["rows"]) if "rows" in case else None
        if ref is None:
            head = np.array(case["rows_head"])
            geo = np.stack([(head[:, 0] - ox) / res, (head[:, 1] - oy) / res, (head[:, 2] - ox) / res, (head[:, 3] - oy) / res], 1)
            np.testing.assert_allclose(ints[:50], geo, atol=1e-3)
            continue
        geo = np.stack([(ref[:, 0] - ox) / res, (ref[:, 1] - oy) / res, (ref[:, 2] - ox) / res, (ref[:, 3] - oy) / res], 1)
        np.testing.assert_allclose(ints, geo, atol=1e-3)
        # tile origin (top-left, y down) against the margin-expanded geometry bounds
        x0 = (ref[:, 4] - ox) / res
        y0 = H - (ref[:, 7] - oy) / res
        np.testing.assert_allclose(t[:, 0], x0, atol=1e-3)
        np.testing.assert_allclose(t[:, 1], y0, atol=1e-3)


def test_write_rects_replay_reference_write_order(golden):
    """Painting interiors in row order with later tiles overwriting (main.py:409-426) must give the
    same owner per pixel as the precomputed write rectangles."""
    for case in golden["slice_extent"]:
        W, H = int(case["spec"][0]), int(case["spec"][1])
        if W * H > 4_000_000 or "rows" not in case:
            continue
        size, margin = case["size"], case["margin"]
        t = slicing_job.tile_table(W, H, size, margin, case["stride"])
        ints = slicing_job.tile_interiors(W, H, size, margin, case["stride"])
        owner = np.full((H, W), -1, np.int32)
        for i, (l, b, r, tp) in enumerate(ints):
            owner[max(H - tp, 0):H - b, max(l, 0):r] = i
        mine = np.full((H, W), -1, np.int32)
        for i, (x0, y0, wx0, wy0, wx1, wy1) in enumerate(t):
            assert (mine[wy0:wy1, wx0:wx1] == -1).all()
            mine[wy0:wy1, wx0:wx1] = i
        assert (mine == owner).all() and (mine >= 0).all()


def test_get_stride_and_weights(golden):
    for g in golden["get_stride"]:
        assert zref.get_stride(g["config"]) == g["stride"]
        assert ptiles.get_stride(g["config"]) == g["stride"]
    z = np.load(GOLDEN / "tiles_weights.npz")
    for n, k in ((512, "w512"), (128, "w128"), (7, "w7")):
        np.testing.assert_array_equal(tiles_ref.patch_weights(n, 0.5), z[k])
        np.testing.assert_array_equal(ptiles.patch_weights(n, 0.5), z[k])
    np.testing.assert_array_equal(tiles_ref.total_weights((1000, 700), 256, [100, 600, 50, 500], 128), z["tw_a"])
    np.testing.assert_array_equal(tiles_ref.total_weights((64, 64), 16, [0, 64, 0, 64], 8), z["tw_b"])
    np.testing.assert_array_equal(tiles_ref.patch_overlap((1000, 700), 256, [100, 600, 50, 500], 128), z["ov_a"])
    np.testing.assert_array_equal(tiles_ref.patch_overlap((64, 64), 16, [0, 64, 0, 64], 8), z["ov_b"])
    for g in golden["get_tile_coord"]:
        assert sorted(tiles_ref.get_tile_coord(*g["args"])) == g["coords"]


def test_blend_oracle_properties():
    """a8: the blend restatement on a position-only toy model (probabilities depend on the tile-local pixel
    position, so overlapping tiles disagree): `average` equals the mean over the covering tiles, the weighted
    mean lies between the per-tile extremes, `max` picks the most confident tile, and class_prob is the
    exact-clipping map of uint8(p * 255)."""
    import torch

    class Toy(torch.nn.Module):
        def forward(self, x):                       # logits from the tile-local coordinates only
            n, _, h, w = x.shape
            yy = torch.linspace(-2, 2, h).view(1, 1, h, 1).expand(n, 1, h, w)
            xx = torch.linspace(-2, 2, w).view(1, 1, 1, w).expand(n, 1, h, w)
            return torch.cat([yy, xx, -yy - xx], dim=1)

    H, W, T, m = 96, 80, 32, 8
    raster = zref.GeoRaster(np.zeros((3, H, W), np.uint8), 10.0, 50.0, 0.5)
    config = {"img_pixels_detection": T, "margin": m, "channels": [1, 2, 3], "n_classes": 3,
              "norma_task": [{"norm_type": "scaling", "norm_means": [], "norm_stds": []}]}
    rows = zref.slice_extent(raster.bounds, (0.5, 0.5), T, m, T - 2 * m)
    probs = torch.softmax(Toy()(torch.zeros(1, 3, T, T)), dim=1)[0].numpy()
    cover = np.zeros((H, W), np.int32)
    psum = np.zeros((3, H, W), np.float64)
    pbest = np.zeros((H, W), np.float64)
    for r in rows:
        x0 = int(round((r["geometry"][0] - raster.min_x) / 0.5)); y0 = int(round((raster.max_y - r["geometry"][3]) / 0.5))
        ys, ye, xs, xe = max(y0, 0), min(y0 + T, H), max(x0, 0), min(x0 + T, W)
        cover[ys:ye, xs:xe] += 1
        psum[:, ys:ye, xs:xe] += probs[:, ys - y0:ye - y0, xs - x0:xe - x0]
        pbest[ys:ye, xs:xe] = np.maximum(pbest[ys:ye, xs:xe], probs[:, ys - y0:ye - y0, xs - x0:xe - x0].max(axis=0))
    assert cover.min() >= 1 and cover.max() >= 4
    cls_avg, conf_avg = zref.run_zone_blend(Toy(), raster, config, "average")
    np.testing.assert_allclose(conf_avg, (psum / cover).max(axis=0), rtol=1e-5)
    top2 = np.sort(psum / cover, axis=0)
    clear = top2[-1] - top2[-2] > 1e-5              # the toy model is symmetric: exact ties exist
    np.testing.assert_array_equal(cls_avg[clear], (psum / cover).argmax(axis=0)[clear])
    cls_w, conf_w = zref.run_zone_blend(Toy(), raster, config, "average_weights")
    assert conf_w.min() > 1 / 3 - 1e-6 and (conf_w <= pbest + 1e-6).all()
    cls_max, conf_max = zref.run_zone_blend(Toy(), raster, config, "max")
    np.testing.assert_allclose(conf_max, pbest, rtol=1e-6)
    cp = zref.run_zone_class_prob(Toy(), raster, config)
    assert cp.shape == (3, H, W) and (cp.sum(axis=0) >= 253).all() and (cp.sum(axis=0) <= 255).all()
    cls_clip, _, _ = zref.run_zone(Toy(), raster, config)
    np.testing.assert_array_equal(cp.astype(np.int32).argmax(axis=0)[cp.max(axis=0) > cp.min(axis=0) + 2],
                                  cls_clip[cp.max(axis=0) > cp.min(axis=0) + 2])


def test_convert_and_normalisation():
    z = np.load(GOLDEN / "convert_norm.npz")
    np.testing.assert_array_equal(zref.convert(z["probs"], "argmax"), z["argmax"])
    np.testing.assert_array_equal(zref.convert(z["probs"], "class_prob"), z["class_prob"])
    means, stds = list(z["means"]), list(z["stds"])
    np.testing.assert_array_equal(zref.normalization(z["img"], "custom", means, stds), z["zone_custom"])
    np.testing.assert_array_equal(zref.normalization(z["img"], "scaling", means, stds), z["zone_scaling"])
    np.testing.assert_array_equal(flair_ref.norm(z["img"].copy(), "custom", means, stds), z["flair_custom"])
    np.testing.assert_array_equal(flair_ref.norm(z["img"].copy(), "scaling"), z["flair_scaling"])
    np.testing.assert_array_equal(flair_ref.norm(z["img"].copy(), "without"), z["flair_without"])
    with pytest.raises(SystemExit):
        flair_ref.norm(z["img"].copy(), "bogus")


def test_metric_formulas(golden):
    for name in ("kat", "rand19"):
        g = golden["metrics"][name]
        cm = np.array(g["cm"])
        iou, miou = metrics_ref.class_IoU(cm)
        p, ap = metrics_ref.class_precision(cm)
        r, ar = metrics_ref.class_recall(cm)
        f, af = metrics_ref.class_fscore(p, r)
        np.testing.assert_array_equal(iou, np.array(g["iou"]))
        assert miou == g["miou"] == g["z_miou"]
        assert metrics_ref.overall_accuracy(cm) == g["oa"] == g["z_oa"]
        np.testing.assert_array_equal(p, np.array(g["precision"]))
        np.testing.assert_array_equal(r, np.array(g["recall"]))
        np.testing.assert_array_equal(f, np.array(g["fscore"]))
        np.testing.assert_array_equal(f, np.array(g["z_fscore"]))
        assert [ap, ar, af] == g["avg"]
    g = golden["metrics"]["kat"]  # hand-computable known answer
    np.testing.assert_allclose(g["iou"], [62.5, 50.0, 0.0])
    assert abs(g["miou"] - 37.5) < 1e-12 and abs(g["oa"] - 72.72727272727273) < 1e-9
    classes = {int(k): v for k, v in golden["classes19"].items()}
    cleaned = metrics_ref.clean_confmat(np.array(golden["metrics"]["rand19"]["cm"]), classes)
    np.testing.assert_array_equal(cleaned, np.array(golden["metrics"]["clean_confmat_rand19"]))


def test_confusion_matrix_restatements():
    z = np.load(GOLDEN / "confusion.npz")
    np.testing.assert_array_equal(metrics_ref.confusion_numpy(z["truth"], z["pred"], 19, 1), z["cm"])
    np.testing.assert_array_equal(metrics_ref.patch_confusion(z["truth"] - 1, z["pred"], 19), z["cm"])
    assert z["cm"].dtype == np.int64


def test_metadata_encoding(golden):
    md = json.loads((GOLDEN / "metadata_aerial.json").read_text())
    g = golden["parsing_metadata"]
    enc = flair_ref.parsing_metadata(g["images"], md)
    assert np.array(enc).shape == (3, 45)
    np.testing.assert_array_equal(np.array(enc), np.array(g["encoded"]))


def test_checkpoint_key_handling(golden):
    import torch
    g = golden["checkpoint"]
    sd = {"model.seg_model.encoder.conv1.weight": torch.ones(1), "model.seg_model.segmentation_head.0.bias": torch.zeros(2),
          "model.enc.enc_mlp.0.weight": torch.ones(3), "criterion.weight": torch.ones(2)}
    assert sorted(flair_ref.get_module(sd, False).keys()) == g["pth_prefixed"]
    assert sorted(flair_ref.get_module({"state_dict": sd}, True).keys()) == g["ckpt_prefixed"]
    assert sorted(flair_ref.get_module({"encoder.conv1.weight": torch.ones(1)}, False).keys()) == g["pth_bare"]
    s = g["surgery"]
    src = {k: torch.tensor(v) for k, v in s["src"].items()}
    dst_shapes = {"head.weight": torch.zeros(3, 4, 3, 3), "head.bias": torch.zeros(3), "criterion.weight": torch.zeros(3)}
    out = flair_ref.load_checkpoint_state(src, dst_shapes, {int(k): v for k, v in s["classes"].items()})
    for k, v in s["dst"].items():
        np.testing.assert_array_equal(np.asarray(out[k]), np.array(v, dtype=np.float32))


def test_unet_restatement_pins():
    """What can be pinned about the un-vendored network: key set, parameter count (README.md:91),
    output shape, divisibility check."""
    import torch
    from oracle.unet_smp033 import Unet
    from flair1_b200.zone_detect.model import expected_keys
    m = Unet(3, 15)
    sd = m.state_dict()
    assert len(sd) == 278
    assert sum(p.numel() for p in m.parameters()) == 24_438_399
    assert {k for k in sd if not k.endswith("num_batches_tracked")} == expected_keys()
    m.eval()
    with torch.no_grad():
        assert m(torch.zeros(1, 3, 64, 64)).shape == (1, 15, 64, 64)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 65, 64))
    assert Unet(5, 19).state_dict()["encoder.conv1.weight"].shape == (64, 5, 7, 7)


def test_batch_mode_filename_grammar_and_metrics(tmp_path, monkeypatch):
    """Batch mode `-b` host logic against the reference's own code (golden["batch"], produced by executing
    utils.info_extract / extract_method and test/metrics.collect_paths_truth / batch_metrics of /root/reference):
    file-name grammar, prediction <-> truth pairing, per-method grouping and every metric value. The one GPU
    call (confusion matrix of two rasters) is replaced by the oracle's numpy count here; the GPU suite runs
    the real thing."""
    import json as _json
    from oracle.metrics_ref import confusion_numpy
    import flair1_b200.zone_detect.metrics as zmet
    from flair1_b200.zone_detect.utils import extract_method, info_extract
    g = _json.loads((Path(__file__).parent / "golden" / "golden.json").read_text())["batch"]
    for case in g["info_extract"]:
        assert info_extract(Path(case["file"])) == case["info"]
    for case in g["extract_method"]:
        assert extract_method(case["method"]) == case["info"]
    assert g["extract_method_underscore_value"] == "IndexError"
    with pytest.raises(IndexError):
        extract_method("size=512_stride=256_margin=128_padding=no-padding_stitching=average_weights")
    with pytest.raises(ValueError, match=g["info_extract_bad_suffix"][:20]):
        info_extract(Path("/a/b/zone.png"))

    out_dir, truth_dir = tmp_path / "out", tmp_path / "truth" / "032_2019"
    arrays = {}
    for zi, zone in enumerate(g["zones"]):
        ts = out_dir / f"20250101_00000{zi}"
        ts.mkdir(parents=True)
        (truth_dir / zone).mkdir(parents=True)
        tpath = truth_dir / zone / f"032_2019_{zone}_MSK.tif"
        tpath.touch()
        arrays[str(tpath)] = np.array(g["files"][f"truth/{zone}"], np.uint8)
        for method in g["methods"]:
            ppath = ts / f"032_2019_{zone}_RGBI-ARGMAX-S_{method}.tif"
            ppath.touch()
            arrays[str(ppath)] = np.array(g["files"][f"pred/{zone}/{method}"], np.uint8)
    classes = {int(k): v for k, v in g["classes"].items()}
    config = {"output_path": str(out_dir), "classes": classes, "model_name": "resnet34-unet",
              "times": {g["methods"][0]: [10.0, 30.0]}}
    df = zmet.collect_paths_truth(config, truth_dir)
    got = sorted([[Path(r.pred_path).name, Path(r.truth_path).name, r.method] for r in df.itertuples()])
    assert got == g["collect_paths_truth"]
    monkeypatch.setattr(zmet, "confmat_of_rasters",
                        lambda model, pp, tp, n: confusion_numpy(arrays[tp], arrays[pp], n, 1))
    res = zmet.batch_metrics(config, truth_dir, model=object())
    assert _json.loads(_json.dumps(res)) == g["batch_metrics"]


def test_tile_windows_are_the_interiors():
    """tile_windows = the reference's left/bottom/right/top interior boxes (already pinned by the slice_extent
    golden cases through tile_interiors) flipped to top-left pixel coordinates; every write rectangle lies in
    its window and the windows cover the raster."""
    from flair1_b200.zone_detect.slicing_job import tile_interiors, tile_table, tile_windows
    for W, H, T, m in [(1000, 700, 512, 128), (513, 513, 512, 0), (10000, 10000, 512, 128), (300, 280, 256, 32)]:
        t, w, ints = tile_table(W, H, T, m), tile_windows(W, H, T, m), tile_interiors(W, H, T, m)
        assert (w[:, :2] == t[:, :2]).all()
        assert (w[:, 2] == ints[:, 0]).all() and (w[:, 4] == ints[:, 2]).all()
        assert (w[:, 3] == H - ints[:, 3]).all() and (w[:, 5] == H - ints[:, 1]).all()
        own = (t[:, 4] > t[:, 2]) & (t[:, 5] > t[:, 3])
        assert (t[own, 2] >= w[own, 2]).all() and (t[own, 4] <= w[own, 4]).all()
        assert (t[own, 3] >= w[own, 3]).all() and (t[own, 5] <= w[own, 5]).all()
        assert (w[:, 2] >= t[:, 0]).all() and (w[:, 4] <= t[:, 0] + T).all() and (w[:, 3] >= t[:, 1]).all() and (w[:, 5] <= t[:, 1] + T).all()
        cover = np.zeros((H, W), bool)
        for x0, y0, a, b, c, d in w:
            cover[b:d, a:c] = True
        assert cover.all()


def test_oracle_metadata_forward_equals_the_references_forward():
    """oracle.unet_smp033.FlairModel.forward (the restatement of src/flair/model.py:52-70) reproduces the fixture made by
    the reference's own FLAIR_ModelFactory.forward + MetadataMLP: same MLP, same repeat(1,512,1,16) broadcast, add."""
    import numpy as np
    import torch
    from oracle import synth
    from oracle.flair_ref import norm
    from oracle.unet_smp033 import FlairModel
    g = np.load(GOLDEN / "metadata_forward.npz")
    sd = synth.random_checkpoint(5, 13, seed=int(g["weight_seed"]), use_metadata=True)
    checksum = float(sum(v.double().abs().sum() for v in sd.values() if v.is_floating_point()))
    if abs(checksum - float(g["weights_checksum"])) > 1e-6 * float(g["weights_checksum"]):
        import pytest
        pytest.skip("torch draws other numbers from the seed than when the golden was made")
    m = FlairModel(5, 13, True)
    m.load_state_dict(sd, strict=True)
    m.eval()
    img = synth.synth_raster(5, 512, 512, seed=int(g["img_seed"]))
    x = torch.as_tensor(norm(img, "custom", synth.FLAIR_MEANS, synth.FLAIR_STDS), dtype=torch.float)[None]
    with torch.no_grad():
        y = m(x, torch.from_numpy(g["met"]))
        e = m.enc(torch.from_numpy(g["met"]))
    np.testing.assert_allclose(e.numpy(), g["enc"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(y[:, :, ::16, ::16].numpy(), g["logits_sub"], rtol=0, atol=1e-4 * float(g["logits_absmax"]))
