"""Generates tests/golden/*.npz|json by executing the REFERENCE'S OWN CODE from /root/reference.

Run in the build container only (`python tests/golden/make_golden.py`); /root/reference does not exist
on the GPU box and nothing at test time reads it. The reference cannot be imported as-is here because
its I/O and framework dependencies are missing (rasterio, geopandas, shapely, skimage,
pytorch_lightning, torchmetrics, albumentations, matplotlib, segmentation_models_pytorch), so this
script installs *stub modules* for exactly those names and then imports the real `src.*` modules.
Only I/O is faked (a raster header, a dataframe container, a box); every line of arithmetic that ends
up in a fixture is the reference's. Where a stub carries semantics it is stated:

  rasterio.open(path)            -> header of a synthetic north-up raster (bounds, res, shape, profile)
  geopandas.GeoDataFrame(rows)   -> plain list holder
  shapely.geometry.box(a,b,c,d)  -> object with .bounds = (min x, min y, max x, max y)
  skimage.img_as_float(uint8)    -> x / 255 in float64 (skimage's documented uint8 behaviour)
  pytorch_lightning rank_zero_only -> identity decorator
  smp.create_model(...)          -> oracle.unet_smp033.Unet (the restated network; the reference's
                                    OWN forward()/MetadataMLP code then runs on top of it)
"""
from __future__ import annotations

import json
import sys
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(REF))


# ------------------------------------------------------------------------------------ stubs
class _Anything(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = _Anything(f"{self.__name__}.{name}")
        setattr(self, name, m)
        return m

    def __call__(self, *a, **k):
        return _Anything("call")


def _stub(name: str) -> types.ModuleType:
    parts = name.split(".")
    for i in range(1, len(parts) + 1):
        n = ".".join(parts[:i])
        if n not in sys.modules:
            sys.modules[n] = _Anything(n)
            if i > 1:
                setattr(sys.modules[".".join(parts[:i - 1])], parts[i - 1], sys.modules[n])
    return sys.modules[name]


class _Box:
    def __init__(self, minx, miny, maxx, maxy):
        self.bounds = (min(minx, maxx), min(miny, maxy), max(minx, maxx), max(miny, maxy))


class _Frame:
    def __init__(self, rows, crs=None, geometry=None):
        self.rows = rows

    def __len__(self):
        return len(self.rows)

    def to_file(self, *a, **k):
        pass


class _FakeRaster:
    def __init__(self, spec):
        self.spec = spec
        W, H, res, ox, oy = spec
        self.bounds = (ox, oy, ox + W * res, oy + H * res)
        self.res = (res, res)
        self.profile = {"crs": "EPSG:2154", "width": W, "height": H, "count": 3, "dtype": "uint8"}

    def read(self, band):
        W, H = self.spec[0], self.spec[1]
        return np.broadcast_to(np.uint8(0), (H, W))  # slice_extent only takes .shape (slicing_job.py:30)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_RASTERS = {}
_ARRAYS = {}   # path -> uint8 [H, W] band 1 of a stubbed raster (batch_metrics reads predictions and truths)


class _ArrayRaster:
    def __init__(self, arr):
        self.arr = arr

    def read(self, band):
        return self.arr.copy()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def install_stubs():
    for n in ["rasterio", "rasterio.windows", "rasterio.enums", "rasterio.features", "rasterio.io", "rasterio._err",
              "geopandas", "shapely", "shapely.geometry", "skimage", "skimage.util", "matplotlib", "matplotlib.pyplot",
              "pytorch_lightning", "pytorch_lightning.utilities", "pytorch_lightning.utilities.rank_zero",
              "pytorch_lightning.callbacks", "pytorch_lightning.callbacks.progress.tqdm_progress",
              "pytorch_lightning.loggers", "torchmetrics", "torchmetrics.classification", "torchmetrics.aggregation",
              "albumentations", "segmentation_models_pytorch", "scipy.ndimage"]:
        if n.split(".")[0] in ("scipy",):
            continue
        _stub(n)
    sys.modules["rasterio"].open = lambda path, *a, **k: (_ArrayRaster(_ARRAYS[str(path)]) if str(path) in _ARRAYS
                                                          else _FakeRaster(_RASTERS[str(path)]))
    sys.modules["geopandas"].GeoDataFrame = _Frame
    sys.modules["shapely.geometry"].box = _Box
    sys.modules["shapely"].Polygon = _Box
    sys.modules["shapely.geometry"].mapping = lambda b: {"bounds": b.bounds}
    f = lambda a: a.astype(np.float64) / 255.0 if a.dtype == np.uint8 else a.astype(np.float64)  # noqa: E731
    sys.modules["skimage"].img_as_float = f
    sys.modules["skimage.util"].img_as_float = f
    sys.modules["pytorch_lightning.utilities.rank_zero"].rank_zero_only = lambda fn: fn
    import torch.nn as nn
    sys.modules["pytorch_lightning"].LightningModule = nn.Module
    sys.modules["pytorch_lightning"].LightningDataModule = object
    sys.modules["pytorch_lightning.callbacks"].BasePredictionWriter = object
    from oracle.unet_smp033 import Unet
    sys.modules["segmentation_models_pytorch"].create_model = \
        lambda arch, encoder_name, classes, in_channels: Unet(in_channels, classes)


def _jsonable(o):
    if isinstance(o, dict):
        return {str(k): _jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_jsonable(v) for v in o]
    if isinstance(o, (np.floating, np.integer)):
        return o.item()
    if isinstance(o, np.ndarray):
        return o.tolist()
    return o


# ------------------------------------------------------------------------------------ fixtures
def gold_slicing(out: dict):
    from src.zone_detect.slicing_job import slice_extent
    cases = []
    specs = [(1000, 700, 1.0, 0.0, 0.0), (513, 513, 0.2, 800000.0, 6500000.0), (512, 512, 1.0, 0.0, 0.0),
             (10000, 10000, 0.2, 800000.0, 6500000.0), (2048, 1536, 0.2, 812345.6, 6512345.4), (300, 300, 1.0, 0.0, 0.0),
             (40000, 40000, 0.2, 800000.0, 6500000.0)]
    for spec in specs:
        for size, margin in [(512, 0), (512, 64), (512, 128), (256, 32), (1024, 128)]:
            if spec[0] >= 40000 and (size, margin) != (512, 128):
                continue
            key = f"r{len(_RASTERS)}"
            _RASTERS[key] = spec
            stride = size - 2 * margin
            frame, profile, res, img_size = slice_extent(key, size, margin, Path("/tmp"), "out", False, stride)
            rows = frame.rows
            arr = np.array([[r["left"], r["bottom"], r["right"], r["top"], *r["geometry"].bounds] for r in rows], dtype=np.float64)
            case = {"spec": list(spec), "size": size, "margin": margin, "stride": stride, "n": len(rows),
                    "res": list(res), "img_size": list(img_size)}
            if len(rows) <= 2000:
                case["rows"] = arr.tolist()
            else:  # large grids: count + checksum + first/last rows
                case["rows_head"] = arr[:50].tolist()
                case["rows_tail"] = arr[-50:].tolist()
                case["rows_sum"] = arr.sum(axis=0).tolist()
            cases.append(case)
    out["slice_extent"] = cases


def gold_tiles(out: dict):
    from src.zone_detect.test.tiles import get_stride, get_tile_coord, patch_overlap, patch_weights, total_weights
    out["get_stride"] = [
        {"config": c, "stride": get_stride(c)} for c in [
            {"img_pixels_detection": 512, "margin": 128},
            {"img_pixels_detection": 1024, "margin": 0},
            {"img_pixels_detection": 512, "margin": 128, "overlap_strat": True,
             "strategies": {"tiling": {"stride_range": [0.25, 0.5, 1.0]}}}]]
    np.savez_compressed(HERE / "tiles_weights.npz",
                        w512=patch_weights(512, 0.5, "exp"), w128=patch_weights(128, 0.5, "exp"), w7=patch_weights(7, 0.5, "exp"),
                        tw_a=total_weights((1000, 700), 256, [100, 600, 50, 500], 128)[0],
                        tw_b=total_weights((64, 64), 16, [0, 64, 0, 64], 8)[0],
                        ov_a=patch_overlap((1000, 700), 256, [100, 600, 50, 500], 128),
                        ov_b=patch_overlap((64, 64), 16, [0, 64, 0, 64], 8))
    out["get_tile_coord"] = [{"args": list(a), "coords": sorted(get_tile_coord(*a))} for a in
                             [(0, 64, 64, 16, 8), (100, 600, 1000, 256, 128), (50, 500, 700, 256, 128), (0, 10, 10, 16, 8)]]


def gold_convert_norm(out: dict):
    from src.zone_detect.dataset import convert, Sliced_Dataset
    from src.flair.data_loader import norm
    rng = np.random.default_rng(7)
    logits = rng.normal(size=(15, 24, 24)).astype(np.float32) * 3
    probs = torch.softmax(torch.from_numpy(logits), 0).numpy()
    img = rng.integers(0, 256, size=(5, 16, 16), dtype=np.uint8)
    means, stds = [105.08, 110.87, 101.82, 106.38, 53.26], [52.17, 45.38, 44, 39.69, 79.3]
    ds = Sliced_Dataset.__new__(Sliced_Dataset)
    ds.norm_type, ds.norm_means, ds.norm_stds, ds.num_bands = "custom", means, stds, 5
    zn_custom = ds.normalization(img)
    ds.norm_type = "scaling"
    zn_scaling = ds.normalization(img)
    np.savez_compressed(HERE / "convert_norm.npz", probs=probs, argmax=convert(probs, "argmax"),
                        class_prob=convert(probs, "class_prob"), img=img,
                        zone_custom=zn_custom, zone_custom_f32=torch.as_tensor(zn_custom, dtype=torch.float).numpy(),
                        zone_scaling=zn_scaling,
                        flair_custom=norm(img.copy(), "custom", means, stds), flair_scaling=norm(img.copy(), "scaling"),
                        flair_without=norm(img.copy(), "without"), means=np.array(means), stds=np.array(stds))


def gold_metrics(out: dict):
    import src.flair.metrics as fm
    import src.zone_detect.test.metrics as zm
    rng = np.random.default_rng(11)
    classes = {i + 1: [0 if i + 1 in (15, 16, 17, 19) else 1, f"class{i + 1}"] for i in range(19)}
    cms = {"kat": np.array([[5, 1, 0], [2, 3, 0], [0, 0, 0]]), "rand19": rng.integers(0, 10000, size=(19, 19))}
    cms["rand19"][:, 5] = 0  # an all-zero column: NaN -> 0 paths
    res = {}
    for name, cm in cms.items():
        with np.errstate(divide="ignore", invalid="ignore"):
            p, ap = fm.class_precision(cm)
            r, ar = fm.class_recall(cm)
            f, af = fm.class_fscore(p, r)
            iou, miou = fm.class_IoU(cm, len(cm))
            ziou, zmiou = zm.class_IoU(cm)
            zf, zaf = zm.class_fscore(cm)
            res[name] = {"cm": cm.tolist(), "oa": fm.overall_accuracy(cm), "iou": iou.tolist(), "miou": miou,
                         "precision": p.tolist(), "recall": r.tolist(), "fscore": f.tolist(),
                         "avg": [ap, ar, af], "z_iou": ziou.tolist(), "z_miou": zmiou, "z_oa": zm.overall_accuracy(cm),
                         "z_fscore": zf.tolist(), "z_avg_fscore": zaf}
    cleaned = zm.clean_confmat(cms["rand19"], {"classes": classes})
    res["clean_confmat_rand19"] = cleaned.tolist()
    # the confusion-matrix call itself: the reference's line, sklearn present in this image
    truth = rng.integers(0, 21, size=(64, 64), dtype=np.uint8)
    pred = rng.integers(0, 20, size=(64, 64), dtype=np.uint8)
    target = truth - 1  # uint8 wrap, flair/metrics.py:62
    cm = fm.confusion_matrix(target.flatten(), pred.flatten(), labels=list(range(19)))
    np.savez_compressed(HERE / "confusion.npz", truth=truth, pred=pred, cm=cm)
    out["metrics"] = res
    out["classes19"] = {str(k): v for k, v in classes.items()}


def gold_metadata(out: dict):
    import src.flair.tasks_utils as tu
    md = {
        "IMG_000001": {"patch_centroid_x": 915984.0, "patch_centroid_y": 6458560.5, "patch_centroid_z": 412.3,
                       "camera": "UCE-M3-f120-s06", "date": "2020-07-14", "time": "11h37"},
        "IMG_000002": {"patch_centroid_x": 489212.4, "patch_centroid_y": 6812345.9, "patch_centroid_z": 0.0,
                       "camera": "UCX-2", "date": "2018-12-31", "time": "08h05"},
        "IMG_000003": {"patch_centroid_x": 1011111.1, "patch_centroid_y": 6222222.2, "patch_centroid_z": 3164.9099121094,
                       "camera": "UCE", "date": "2021-01-01", "time": "16h59"},
    }
    p = HERE / "metadata_aerial.json"
    p.write_text(json.dumps(md, indent=1))
    imgs = [f"/data/D001_2020/Z1_UU/img/{k}.tif" for k in md]
    enc = tu.parsing_metadata(imgs, {"paths": {"path_metadata_aerial": str(p)}})
    out["parsing_metadata"] = {"images": imgs, "encoded": _jsonable(enc)}


def gold_checkpoint(out: dict):
    import os
    import tempfile
    from src.zone_detect.model import get_module
    import src.flair.main as fmain
    sd = {"model.seg_model.encoder.conv1.weight": torch.ones(1), "model.seg_model.segmentation_head.0.bias": torch.zeros(2),
          "model.enc.enc_mlp.0.weight": torch.ones(3), "criterion.weight": torch.ones(2)}
    res = {}
    with tempfile.TemporaryDirectory() as d:
        torch.save(sd, os.path.join(d, "a.pth"))
        torch.save({"state_dict": sd, "epoch": 3}, os.path.join(d, "b.ckpt"))
        torch.save({"encoder.conv1.weight": torch.ones(1)}, os.path.join(d, "c.pth"))
        res["pth_prefixed"] = sorted(get_module(os.path.join(d, "a.pth")).keys())
        res["ckpt_prefixed"] = sorted(get_module(os.path.join(d, "b.ckpt")).keys())
        res["pth_bare"] = sorted(get_module(os.path.join(d, "c.pth")).keys())
        res["missing"] = get_module(os.path.join(d, "nope.pth"))

        # load_checkpoint's class-count surgery (flair/main.py:106-138) on a toy module
        class Toy(torch.nn.Module):
            def __init__(self, n):
                super().__init__()
                self.head = torch.nn.Conv2d(4, n, 3, padding=1)
                self.criterion = torch.nn.CrossEntropyLoss(weight=torch.ones(n))
        src_m = Toy(5)
        torch.manual_seed(0)
        for p_ in src_m.parameters():
            torch.nn.init.normal_(p_)
        torch.save(src_m.state_dict(), os.path.join(d, "five.pth"))
        classes = {1: [1, "a"], 2: [0, "b"], 3: [2, "c"]}
        dst = Toy(3)
        fmain.load_checkpoint({"paths": {"ckpt_model_path": os.path.join(d, "five.pth")}, "classes": classes}, dst)
        res["surgery"] = {"src": {k: v.tolist() for k, v in src_m.state_dict().items()},
                          "dst": {k: v.tolist() for k, v in dst.state_dict().items()},
                          "classes": {str(k): v for k, v in classes.items()}}
    out["checkpoint"] = res


def gold_config(out: dict):
    """gen_param_combination / check_list_type / setup_indiv_path of src/zone_detect/utils.py."""
    import tempfile
    import src.zone_detect.utils as zu
    base = {"img_pixels_detection": 512, "margin": 128}
    cfgs = [dict(base),
            dict(base, overlap_strat=True, strategies={"tiling": {"enabled": True, "size_range": [256, 512, 1024], "stride_range": [0.5, 1.0]},
                                                       "stitching": {"enabled": True, "methods": ["exact-clipping", "average"], "margin": [0.125, 0.25, 0.5]},
                                                       "padding_overall": ["no-padding"]}),
            dict(base, strategies={"tiling": {"enabled": False}, "stitching": {"enabled": True, "methods": ["exact-clipping"], "margin": [64.0, 0.25]}})]
    out["gen_param_combination"] = [{"config": c, "combi": zu.gen_param_combination(c)} for c in cfgs]
    out["check_list_type"] = [{"arg": a, "type": t.__name__, "res": zu.check_list_type(a, t)} for a, t in
                              [(512, int), ([128, 256], int), (0.5, float), ([0.25, 0.5], float), ("max", str), (["a", "b"], str)]]
    with tempfile.TemporaryDirectory() as d:
        names = []
        for _ in range(3):
            _, p = zu.setup_indiv_path({"output_name": "zone", "local_out": d}, "_id")
            open(p, "w").close()
            names.append(Path(p).name)
        out["setup_indiv_path"] = names


def gold_batch(out: dict):
    """Batch mode (-b): filename grammar (utils.py:170-217), prediction/truth collection and the per-method
    metrics of test/metrics.py:48-84, 195-287 executed on a small synthetic department: 2 zones x 2 methods,
    64 x 48 rasters. Only rasterio.open(...).read(1) is stubbed (arrays registered per path)."""
    import tempfile
    import src.zone_detect.utils as zu
    import src.zone_detect.test.metrics as zm
    names = ["032_2019_UA-zone_1_RGBI-ARGMAX-S_size=512_stride=256_margin=128_padding=no-padding_stitching=exact-clipping.tif",
             "077_2021_UN_S-12_3_IRC-ARGMAX-S_size=1024_stride=512_margin=256_padding=no-padding_stitching=average.tif",
             "D032_2019_UA-zone_1_RGBI-ARGMAX-S_size=256_stride=192_margin=32_padding=some-padding_stitching=max_extra=7.tif"]
    res = {"info_extract": [{"file": "/a/b/" + n, "info": zu.info_extract(Path("/a/b/" + n))} for n in names],
           "extract_method": [{"method": m, "info": zu.extract_method(m, {})} for m in
                              ["size=512_stride=256_margin=128_padding=no-padding_stitching=exact-clipping",
                               "size=128_stride=96_margin=16_padding=reflect_stitching=average_foo=bar"]]}
    # the "_" separated grammar cannot carry a value with an underscore: "stitching=average_weights" raises
    try:
        zu.extract_method("size=512_stride=256_margin=128_padding=no-padding_stitching=average_weights", {})
        res["extract_method_underscore_value"] = None
    except IndexError as e:
        res["extract_method_underscore_value"] = "IndexError"
    try:
        zu.info_extract(Path("/a/b/zone.png"))
        res["info_extract_bad_suffix"] = None
    except ValueError as e:
        res["info_extract_bad_suffix"] = str(e)
    rng = np.random.default_rng(23)
    classes = {i + 1: [0 if i + 1 in (13, 14, 15) else 1, f"class{i + 1}"] for i in range(15)}
    methods = ["size=512_stride=256_margin=128_padding=no-padding_stitching=exact-clipping",
               "size=256_stride=192_margin=32_padding=no-padding_stitching=max"]
    zones = ["UA-zone_1", "UN_S-12_3"]
    with tempfile.TemporaryDirectory() as d:
        d = Path(d)
        out_dir, truth_dir = d / "out", d / "truth" / "032_2019"
        files = {}
        for zi, zone in enumerate(zones):
            ts = out_dir / f"20250101_00000{zi}"
            ts.mkdir(parents=True)
            (truth_dir / zone).mkdir(parents=True)
            tpath = truth_dir / zone / f"032_2019_{zone}_MSK.tif"
            tpath.touch()
            truth = rng.integers(0, 17, size=(48, 64), dtype=np.uint8)   # 0 wraps to 255 after -1; 16 is out of range
            _ARRAYS[str(tpath)] = truth
            files[f"truth/{zone}"] = truth.tolist()
            for method in methods:
                ppath = ts / f"032_2019_{zone}_RGBI-ARGMAX-S_{method}.tif"
                ppath.touch()
                pred = rng.integers(0, 15, size=(48, 64), dtype=np.uint8)
                _ARRAYS[str(ppath)] = pred
                files[f"pred/{zone}/{method}"] = pred.tolist()
        config = {"output_path": str(out_dir), "classes": classes, "model_name": "resnet34-unet",
                  "times": {methods[0]: [10.0, 30.0]}}
        df = zm.collect_paths_truth(config, truth_dir)
        res["collect_paths_truth"] = sorted([[Path(r.pred_path).name, Path(r.truth_path).name, r.method] for r in df.itertuples()])
        with np.errstate(divide="ignore", invalid="ignore"):
            metrics = zm.batch_metrics(config, truth_dir)
        res["batch_metrics"] = metrics
        res["files"] = files
        res["zones"] = zones
        res["methods"] = methods
        res["classes"] = {str(k): v for k, v in classes.items()}
    out["batch"] = res


def gold_metadata_forward(out: dict):
    """The reference's own FLAIR_ModelFactory.forward / MetadataMLP (flair/model.py:52-96) on top of the restated
    Unet: pins the MLP + `repeat(1,512,1,16)` broadcast + add semantics. Weights and input are regenerated from
    seeds at test time (oracle.synth.random_checkpoint / synth_raster: a 98 MB state_dict cannot be committed), so
    the fixture holds the reference's OUTPUTS only, plus a checksum of the weights that tells the test whether torch
    still draws the same numbers from the seed."""
    from oracle import synth
    from oracle.flair_ref import norm as ref_norm
    from src.flair.model import FLAIR_ModelFactory, MetadataMLP
    cfg = {"model_framework": {"model_provider": "SegmentationModelsPytorch",
                               "SegmentationModelsPytorch": {"encoder_decoder": "resnet34_unet"}},
           "use_metadata": False, "channels": [1, 2, 3, 4, 5], "classes": {i: [1, str(i)] for i in range(1, 14)}}
    m = FLAIR_ModelFactory(cfg)        # use_metadata=True would hit the NameError at model.py:32
    m.enc = MetadataMLP()
    m.use_metadata = True
    sd = synth.random_checkpoint(5, 13, seed=5, use_metadata=True)   # oracle FlairModel keys: seg_model.* / enc.enc_mlp.*
    missing, unexpected = m.load_state_dict(sd, strict=True), None
    m.eval()
    img = synth.synth_raster(5, 512, 512, seed=33)
    x = torch.as_tensor(ref_norm(img, "custom", synth.FLAIR_MEANS, synth.FLAIR_STDS), dtype=torch.float)[None]
    met = torch.rand((1, 45), generator=torch.Generator().manual_seed(45))
    with torch.no_grad():
        y = m(x, met)
        e = m.enc(met)
    checksum = float(sum(v.double().abs().sum() for v in sd.values() if v.is_floating_point()))
    np.savez_compressed(HERE / "metadata_forward.npz", met=met.numpy(), enc=e.numpy(), logits_sub=y[:, :, ::16, ::16].numpy(),
                        logits_absmax=np.float32(y.abs().max()), weights_checksum=np.float64(checksum),
                        img_seed=np.int64(33), weight_seed=np.int64(5))
    out["metadata_forward"] = {"note": "tests/golden/metadata_forward.npz: outputs of the reference's FLAIR_ModelFactory.forward",
                               "enc": e.tolist(), "logits_absmax": float(y.abs().max()), "weights_checksum": checksum}


def main():
    install_stubs()
    (HERE / "_cache").mkdir(exist_ok=True)
    out = {"generated_by": "tests/golden/make_golden.py", "reference": "Draghoyns/FLAIR-1 @ /root/reference"}
    gold_slicing(out)
    gold_tiles(out)
    gold_convert_norm(out)
    gold_metrics(out)
    gold_metadata(out)
    gold_checkpoint(out)
    gold_config(out)
    gold_batch(out)
    gold_metadata_forward(out)
    (HERE / "golden.json").write_text(json.dumps(_jsonable(out), indent=1))
    print("wrote", HERE / "golden.json")


if __name__ == "__main__":
    main()
