"""Host-side logic, the GeoTIFF codec and the C-ABI surface -- everything that needs no GPU."""
import json
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


# ------------------------------------------------------------------------------------------ C ABI
def test_library_exports_every_declared_symbol():
    """Every function declared in include/flair_b200.h is exported by libflairb200.so and bound."""
    import flair1_b200._native as nat
    header = (ROOT / "include" / "flair_b200.h").read_text()
    declared = set(re.findall(r"\b(fb_[a-z0-9_]+)\s*\(", header))
    declared -= {"fb_ctx", "fb_tensor_desc", "fb_tile"}
    lib = nat.load_library()
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"{sym} declared in the header but not exported"
    assert set(nat.EXPORTED_SYMBOLS) == declared
    assert lib.fb_api_version() == 1


def test_no_cpu_fallback():
    """Without a CUDA device the product path fails loudly at every entry."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import ctypes as C
    import flair1_b200._native as nat
    lib = nat.load_library()
    h = C.c_void_p()
    rc = lib.fb_create(0, None, C.byref(h))
    assert rc == -3 and b"no CPU fallback" in lib.fb_last_error(None)
    with pytest.raises(RuntimeError):
        nat.Context(0)
    from flair1_b200.zone_detect.utils import setup_device
    with pytest.raises(RuntimeError):
        setup_device({"use_gpu": True})


def test_product_never_imports_the_oracle():
    for p in (ROOT / "flair-1_b200").rglob("*.py"):
        src = p.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{p} imports the oracle"


def test_lzw_codec_roundtrip_and_libtiff_interop(tmp_path):
    import io
    from PIL import Image
    import flair1_b200._native as nat
    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 4096, 300_000):
        for kind in ("rand", "runs", "const"):
            a = rng.integers(0, 256, n, dtype=np.uint8) if kind == "rand" else \
                ((np.arange(n) // 37) % 19).astype(np.uint8) if kind == "runs" else np.full(n, 7, np.uint8)
            np.testing.assert_array_equal(nat.lzw_decode(nat.lzw_encode(a), n), a)
    a = ((np.arange(256 * 256) // 37) % 19).astype(np.uint8).reshape(256, 256)
    buf = io.BytesIO()
    Image.fromarray(a).save(buf, format="TIFF", compression="tiff_lzw")
    raw = buf.getvalue()
    im = Image.open(io.BytesIO(raw))
    offs, cnts, rps = im.tag_v2[273], im.tag_v2[279], im.tag_v2[278]
    rows = [nat.lzw_decode(raw[o:o + c], min(rps, 256 - i * rps) * 256) for i, (o, c) in enumerate(zip(offs, cnts))]
    np.testing.assert_array_equal(np.concatenate(rows).reshape(256, 256), a)   # libtiff's stream decodes with ours


def test_geotiff_roundtrip_and_pil_interop(tmp_path):
    from PIL import Image
    from flair1_b200 import geotiff as gt
    rng = np.random.default_rng(1)
    for bands, (H, W) in ((1, (300, 500)), (2, (33, 70)), (5, (512, 512))):
        a = ((np.arange(H * W).reshape(H, W) // 97) % 19).astype(np.uint8)[None].repeat(bands, 0).copy()
        a[:, H // 4:H // 2, W // 3:W // 2] = rng.integers(0, 256, (bands, H // 2 - H // 4, W // 2 - W // 3), dtype=np.uint8)
        for comp in ("lzw", "deflate", "none"):
            for tiled, big in ((True, True), (False, False)):
                p = tmp_path / "t.tif"
                gt.write(p, a, geo_tags=gt.georef_tags(800000.0, 6500000.0 + H * 0.2, 0.2, 0.2), compress=comp, tiled=tiled,
                         blocksize=256 if tiled else 64, bigtiff=big)
                info = gt.read_info(p)
                assert (info.width, info.height, info.count, info.bigtiff) == (W, H, bands, big)
                np.testing.assert_array_equal(gt.read(p), a)
                w = gt.read(p, bands=[bands], window=(3, 5, 40, 20))
                np.testing.assert_array_equal(w[0], a[bands - 1, 5:25, 3:43])
                assert abs(info.bounds[0] - 800000.0) < 1e-6 and abs(info.bounds[3] - (6500000.0 + H * 0.2)) < 1e-6
                assert abs(info.res[0] - 0.2) < 1e-12
    a = rng.integers(0, 19, (3, 100, 160), dtype=np.uint8)
    p = tmp_path / "x.tif"
    gt.write(p, a, compress="lzw", tiled=True, blocksize=64, bigtiff=False)
    np.testing.assert_array_equal(np.array(Image.open(p)).transpose(2, 0, 1), a)       # libtiff reads ours
    Image.fromarray(a.transpose(1, 2, 0)).save(p, compression="tiff_lzw")
    np.testing.assert_array_equal(gt.read(p), a)                                       # we read libtiff's
    with pytest.raises(IndexError):
        gt.read(p, bands=[4])


# ------------------------------------------------------------------------------------------ config surface
def _detect_config(tmp_path):
    w = tmp_path / "w.pth"
    w.write_bytes(b"x")
    img = tmp_path / "img.tif"
    img.write_bytes(b"x")
    return {"output_path": str(tmp_path / "out"), "output_name": "zone", "input_img_path": str(img), "channels": [1, 2, 3],
            "img_pixels_detection": 512, "margin": 128, "output_type": "argmax", "n_classes": 15, "model_weights": str(w),
            "model_framework": {"model_provider": "SegmentationModelsPytorch",
                                "SegmentationModelsPytorch": {"encoder_decoder": "resnet34_unet"}},
            "batch_size": 4, "use_gpu": True, "num_worker": 2, "write_dataframe": False,
            "norma_task": [{"norm_type": "custom", "norm_means": [1, 2, 3], "norm_stds": [1, 1, 1]}],
            "metrics": False, "batch_mode": False, "compare": False}


def test_preprocess_config_errors_match_reference(tmp_path):
    from flair1_b200.zone_detect.utils import preprocess_config
    cfg = preprocess_config(_detect_config(tmp_path))
    assert isinstance(cfg["input_img_path"], Path) and (tmp_path / "out").is_dir()
    for key, val, msg in (("margin", 256, "Margin should be an integer"), ("output_type", "bogus", "Invalid output type"),
                          ("channels", [1, "2"], "Channels should be a list of integers"), ("n_classes", "15", "n_classes should be an integer"),
                          ("img_pixels_detection", 512.0, "img_pixels_detection should be an integer")):
        bad = _detect_config(tmp_path)
        bad[key] = val
        with pytest.raises(AssertionError, match=msg):
            preprocess_config(bad)
    bad = _detect_config(tmp_path)
    bad["norma_task"][0]["norm_type"] = "zscore"
    with pytest.raises(AssertionError, match="Invalid normalization type"):
        preprocess_config(bad)
    bad = _detect_config(tmp_path)
    (tmp_path / "w.bin").write_bytes(b"x")
    bad["model_weights"] = str(tmp_path / "w.bin")
    with pytest.raises(ValueError, match=r"should be a \.pth or \.ckpt"):
        preprocess_config(bad)
    bad = _detect_config(tmp_path)
    bad["input_img_path"] = str(tmp_path / "missing.tif")
    with pytest.raises(AssertionError, match="Input image path does not exist"):
        preprocess_config(bad)


def test_config_helpers_match_reference_golden(golden, tmp_path):
    from flair1_b200.zone_detect import utils as zu
    for g in golden["gen_param_combination"]:
        assert zu.gen_param_combination(g["config"]) == g["combi"]
    types = {"int": int, "float": float, "str": str}
    for g in golden["check_list_type"]:
        assert zu.check_list_type(g["arg"], types[g["type"]]) == g["res"]
    names = []
    for _ in range(3):
        _, p = zu.setup_indiv_path({"output_name": "zone", "local_out": str(tmp_path)}, "_id")
        open(p, "w").close()
        names.append(Path(p).name)
    assert names == golden["setup_indiv_path"]


def test_detect_yaml_of_the_reference_parses(tmp_path):
    """The key surface of configs/flair-1-config-detect.yaml is accepted as is."""
    import yaml
    from flair1_b200.zone_detect.utils import preprocess_config
    text = f"""
output_path: {tmp_path}/o
output_name: out
input_img_path: {tmp_path}/img.tif
channels: [1, 2, 3]
img_pixels_detection: 512
margin: 128
output_type: "argmax"
n_classes: 15
model_weights: {tmp_path}/w.pth
model_framework:
    model_provider: SegmentationModelsPytorch
    HuggingFace:
        org_model:
    SegmentationModelsPytorch:
        encoder_decoder: resnet34_unet
batch_size: 4
use_gpu: true
num_worker: 2
write_dataframe: False
norma_task:
  - norm_type: custom
    norm_means: [105.08, 110.87, 101.82]
    norm_stds: [52.17, 45.38, 44]
"""
    (tmp_path / "img.tif").write_bytes(b"x")
    (tmp_path / "w.pth").write_bytes(b"x")
    cfg = yaml.safe_load(text)
    cfg.update({"metrics": False, "batch_mode": False, "compare": False})
    assert preprocess_config(cfg)["margin"] == 128


# ------------------------------------------------------------------------------------------ patch path host logic
def test_product_metadata_encoding_matches_reference(golden):
    from flair1_b200.flair.tasks_utils import parsing_metadata
    g = golden["parsing_metadata"]
    enc = parsing_metadata(g["images"], {"paths": {"path_metadata_aerial": str(GOLDEN / "metadata_aerial.json")}})
    np.testing.assert_array_equal(np.array(enc), np.array(g["encoded"]))


def test_product_metric_formulas_match_reference(golden):
    from flair1_b200.flair import metrics as fm
    from flair1_b200.zone_detect import metrics as zm
    classes = {int(k): v for k, v in golden["classes19"].items()}
    for name in ("kat", "rand19"):
        g = golden["metrics"][name]
        cm = np.array(g["cm"])
        with np.errstate(divide="ignore", invalid="ignore"):
            p, ap = fm.class_precision(cm)
            r, ar = fm.class_recall(cm)
            f, af = fm.class_fscore(p, r)
            iou, miou = fm.class_IoU(cm, len(cm))
            ziou, zmiou = zm.class_IoU(cm)
            zf, zaf = zm.class_fscore(cm)
        np.testing.assert_array_equal(iou, np.array(g["iou"]))
        np.testing.assert_array_equal(ziou, np.array(g["z_iou"]))
        np.testing.assert_array_equal(f, np.array(g["fscore"]))
        np.testing.assert_array_equal(zf, np.array(g["z_fscore"]))
        assert (miou, fm.overall_accuracy(cm), zmiou, zm.overall_accuracy(cm), zaf) == \
            (g["miou"], g["oa"], g["z_miou"], g["z_oa"], g["z_avg_fscore"])
        assert [ap, ar, af] == g["avg"]
    cleaned = zm.clean_confmat(np.array(golden["metrics"]["rand19"]["cm"]), {"classes": classes})
    np.testing.assert_array_equal(cleaned, np.array(golden["metrics"]["clean_confmat_rand19"]))


def test_checkpoint_key_handling_matches_reference(golden, tmp_path):
    import torch
    from flair1_b200.zone_detect.model import check_strict, expected_keys, get_module
    g = golden["checkpoint"]
    sd = {"model.seg_model.encoder.conv1.weight": torch.ones(1), "model.seg_model.segmentation_head.0.bias": torch.zeros(2),
          "model.enc.enc_mlp.0.weight": torch.ones(3), "criterion.weight": torch.ones(2)}
    torch.save(sd, tmp_path / "a.pth")
    torch.save({"state_dict": sd, "epoch": 3}, tmp_path / "b.ckpt")
    torch.save({"encoder.conv1.weight": torch.ones(1)}, tmp_path / "c.pth")
    assert sorted(get_module(str(tmp_path / "a.pth")).keys()) == g["pth_prefixed"]
    assert sorted(get_module(str(tmp_path / "b.ckpt")).keys()) == g["ckpt_prefixed"]
    assert sorted(get_module(str(tmp_path / "c.pth")).keys()) == g["pth_bare"]
    assert get_module(str(tmp_path / "nope.pth")) == g["missing"]
    from oracle.unet_smp033 import Unet
    full = Unet(3, 15).state_dict()
    check_strict(full)                                     # strict=True accepts exactly the smp key set
    broken = dict(full)
    broken.pop("decoder.blocks.2.conv1.1.running_var")
    broken["extra.weight"] = torch.zeros(1)
    with pytest.raises(RuntimeError, match="Missing key"):
        check_strict(broken)
    assert len(expected_keys(True)) == len(expected_keys(False)) + 6


def test_flair_norm_argument_errors():
    from flair1_b200.flair.data_loader import norm
    img = np.zeros((3, 4, 4), np.uint8)
    assert norm(img, "custom", [1, 2, 3], [1, 1, 1]) is img
    with pytest.raises(SystemExit):
        norm(img, "bogus")
    with pytest.raises(SystemExit):
        norm(img, "custom", [1, 2], [1, 1, 1])


# ------------------------------------------------------------------------------------------ sharding + collectives (gloo, 2 ranks)
def test_row_sharding_covers_every_tile_once():
    from flair1_b200.zone_detect.slicing_job import split_rows_across_ranks, tile_table
    tiles = tile_table(40000, 40000, 512, 128)
    for world in (1, 2, 4, 8):
        shards = split_rows_across_ranks(tiles, world)
        allidx = np.concatenate(shards)
        assert sorted(allidx.tolist()) == list(range(len(tiles)))
        sizes = [len(s) for s in shards]
        assert max(sizes) - min(sizes) <= 157            # balanced to one tile row
        for s in shards:                                 # a shard = whole tile rows, write rects are disjoint bands
            ys = np.unique(tiles[s, 1])
            assert np.isin(tiles[:, 1], ys).sum() == len(s)


def test_tile_range_sharding_balances_to_one_tile_and_partitions_the_map():
    """split_tiles_across_ranks: contiguous ranges of the row-major tile order, sizes within one tile of each other
    (157 tile rows over 8 ranks would give 20 against 19.6 rows), and the owned rectangles of the shards partition
    the raster exactly (each pixel written by exactly one rank)."""
    from flair1_b200.zone_detect.slicing_job import owned_rects, split_tiles_across_ranks, tile_table
    tiles = tile_table(40000, 40000, 512, 128)
    for world in (1, 2, 4, 8):
        shards = split_tiles_across_ranks(tiles, world)
        assert sorted(np.concatenate(shards).tolist()) == list(range(len(tiles)))
        sizes = [len(s) for s in shards]
        assert max(sizes) - min(sizes) <= 1
        area = 0
        for s in shards:
            t = tiles[s]
            assert (np.diff(t[:, 1]) >= 0).all()                      # y-sorted: a shard needs one contiguous row span
            r = owned_rects(t)
            assert len(r) <= len(np.unique(t[:, 1])) + 1
            area += int(((r[:, 1] - r[:, 0]) * (r[:, 3] - r[:, 2])).sum())
        assert area == 40000 * 40000
    # small ragged zone, pixel-exact cover, also with more ranks than tile rows
    for (W, H, T, M, world) in ((1500, 1100, 512, 128, 3), (700, 513, 256, 32, 5), (300, 300, 512, 128, 4)):
        tiles = tile_table(W, H, T, M)
        cover = np.zeros((H, W), np.int32)
        for s in split_tiles_across_ranks(tiles, world):
            for y0, y1, x0, x1 in owned_rects(tiles[s]):
                cover[y0:y1, x0:x1] += 1
        assert (cover == 1).all()


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from flair1_b200.zone_detect.slicing_job import split_rows_across_ranks, tile_table
from oracle.metrics_ref import confusion_numpy
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
W, H, T, M, NC = 1500, 1100, 512, 128, 15
rng = np.random.default_rng(7)
pred = rng.integers(0, NC, (H, W), dtype=np.uint8)
truth = rng.integers(0, NC + 3, (H, W), dtype=np.uint8)
tiles = tile_table(W, H, T, M)
mine = tiles[split_rows_across_ranks(tiles, world)[rank]]
r0, r1 = int(mine[:, 3].min()), int(mine[:, 5].max())
# every rank owns a disjoint band of rows; bands tile the raster
bands = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
dist.all_gather(bands, torch.tensor([r0, r1]))
cover = np.zeros(H, np.int32)
for b in bands:
    cover[int(b[0]):int(b[1])] += 1
assert (cover == 1).all(), cover
cm = torch.from_numpy(confusion_numpy(truth[r0:r1], pred[r0:r1], NC, 1))
dist.all_reduce(cm)
assert np.array_equal(cm.numpy(), confusion_numpy(truth, pred, NC, 1))
# strip gather to rank 0 (what zone_detect.main._gather_strips does over NCCL)
rows_max = max(int(b[1] - b[0]) for b in bands)
padded = torch.zeros((rows_max, W), dtype=torch.uint8)
padded[:r1 - r0] = torch.from_numpy(pred[r0:r1])
bufs = [torch.empty_like(padded) for _ in range(world)] if rank == 0 else None
dist.gather(padded, bufs, dst=0)
if rank == 0:
    full = np.zeros((H, W), np.uint8)
    for b, buf in zip(bands, bufs):
        full[int(b[0]):int(b[1])] = buf[:int(b[1] - b[0])].numpy()
    assert np.array_equal(full, pred)
dist.destroy_process_group()
import sys
sys.stdout.write("rank %d ok\\n" % rank)   # one write per rank: the two ranks share the pipe
sys.stdout.flush()
"""


def test_two_rank_confusion_allreduce_and_strip_gather_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=str(ROOT)))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29613", str(script)],
                       capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout


def test_dead_output_regions_cover_the_receptive_field():
    """csrc/tile_need.cuh (host half, no GPU): for random write rectangles the region need_rect() reports for every
    decoder layer equals the exact dependency set obtained by pushing a pixel mask backwards through the decoder
    (3x3 conv = dilation by one pixel clipped to the grid, nearest x2 upsample = pixel (y, x) reads (y//2, x//2)),
    which is what makes the elimination bit-exact: nothing outside is read, nothing inside is missing."""
    import ctypes as C
    import flair1_b200._native as nat
    lib = nat.load_library(build_if_missing=False)
    rng = np.random.default_rng(7)

    def dilate(m):
        p = np.pad(m, 1)
        out = np.zeros_like(m)
        for dy in range(3):
            for dx in range(3):
                out |= p[dy:dy + m.shape[0], dx:dx + m.shape[1]]
        return out

    def bbox(m):
        ys, xs = np.nonzero(m)
        return [int(xs.min()), int(ys.min()), int(xs.max()) + 1, int(ys.max()) + 1] if ys.size else [0, 0, 0, 0]

    cases = [(512, 128, 128, 384, 384), (512, 0, 0, 256, 384), (512, 128, 0, 512, 512), (256, 32, 32, 224, 224), (64, 0, 0, 64, 64)]
    for _ in range(12):
        T = int(rng.choice([64, 128, 256, 512]))
        x0, y0 = int(rng.integers(0, T - 1)), int(rng.integers(0, T - 1))
        cases.append((T, x0, y0, int(rng.integers(x0 + 1, T + 1)), int(rng.integers(y0 + 1, T + 1))))
    for T, ax0, ay0, ax1, ay1 in cases:
        mask = np.zeros((T, T), bool)
        mask[ay0:ay1, ax0:ax1] = True
        want = {10: mask}
        cur = dilate(mask)                       # the head's 3x3 window -> dec4.conv2 output
        for d in range(4, -1, -1):
            want[2 * d + 1] = cur
            cur = dilate(cur)                    # conv2's window -> conv1 output
            want[2 * d] = cur
            cur = dilate(cur)                    # conv1's window on the upsampled grid
            S = cur.shape[0] // 2
            cur = cur.reshape(S, 2, S, 2).any(axis=(1, 3))   # -> pixels of the low-res producer that are read
        for layer, m in want.items():
            r = (C.c_int32 * 4)()
            assert lib.fb_debug_need_rect(T, layer, ax0, ay0, ax1, ay1, r) == 0
            bb = bbox(m)
            assert list(r) == bb, (T, layer, (ax0, ay0, ax1, ay1), list(r), bb)
            assert m[bb[1]:bb[3], bb[0]:bb[2]].all()          # the dependency set is exactly that rectangle
    r = (C.c_int32 * 4)()
    assert lib.fb_debug_need_rect(512, 3, 10, 10, 10, 40, r) == 0 and list(r) == [0, 0, 0, 0]   # empty write rectangle


def test_origin_shifted_tile_cover_contains_the_region_with_the_fewest_tiles():
    """csrc/tile_need.cuh need_span (host half): the kernel tiles of a launch start at the needed region instead of on
    the fixed tile grid. For every decoder layer, kernel tiling the library uses and random write rectangles the cover
    must contain the region, stay inside the grid (no store masking in the kernels) and use ceil(len / tile) tiles
    per axis -- never more than the grid-aligned cover."""
    import ctypes as C
    import flair1_b200._native as nat
    lib = nat.load_library(build_if_missing=False)
    rng = np.random.default_rng(3)
    tilings = [(1, 16, 16), (1, 16, 8), (1, 16, 32), (1, 8, 16), (2, 16, 16), (2, 16, 8), (2, 8, 16),
               (1, 4, 16), (1, 4, 8), (2, 4, 16), (2, 4, 8)]   # the last four: half / quarter boxes of the implicit GEMM
    cases = [(512, 128, 128, 384, 384), (512, 0, 0, 384, 384), (512, 128, 128, 512, 512), (512, 0, 128, 512, 384), (256, 64, 64, 192, 192)]
    for _ in range(40):
        T = int(rng.choice([256, 512, 1024]))
        x0, y0 = int(rng.integers(0, T - 1)), int(rng.integers(0, T - 1))
        cases.append((T, x0, y0, int(rng.integers(x0 + 1, T + 1)), int(rng.integers(y0 + 1, T + 1))))
    fewer = 0
    for T, ax0, ay0, ax1, ay1 in cases:
        for layer in range(11):
            S_out = T if layer >= 10 else (T // 16) << (layer // 2)
            r = (C.c_int32 * 4)()
            assert lib.fb_debug_need_rect(T, layer, ax0, ay0, ax1, ay1, r) == 0
            for scale, th, tw in tilings:
                S = S_out // scale
                cov = (C.c_int32 * 4)()
                rc = lib.fb_debug_tile_cover(T, layer, scale, th, tw, ax0, ay0, ax1, ay1, cov)
                if S % th or S % tw:
                    assert rc != 0
                    continue
                assert rc == 0
                ox, oy, nx, ny = cov
                lo_x, hi_x = r[0] // scale, (r[2] - 1) // scale + 1
                lo_y, hi_y = r[1] // scale, (r[3] - 1) // scale + 1
                assert 0 <= ox <= lo_x and hi_x <= ox + nx * tw <= S, (T, layer, scale, tw, list(r), list(cov))
                assert 0 <= oy <= lo_y and hi_y <= oy + ny * th <= S, (T, layer, scale, th, list(r), list(cov))
                assert nx == -(-(hi_x - lo_x) // tw) and ny == -(-(hi_y - lo_y) // th)
                aligned = ((hi_x - 1) // tw - lo_x // tw + 1) * ((hi_y - 1) // th - lo_y // th + 1)
                assert nx * ny <= aligned
                fewer += nx * ny < aligned
    assert fewer > 0
    # the interior tile of the bench zone: dec2.conv2 (128^2 grid, 16 x 16 tiles) needs 68 pixels = 5 tiles from column 30
    cov = (C.c_int32 * 4)()
    assert lib.fb_debug_tile_cover(512, 5, 1, 16, 16, 128, 128, 384, 384, cov) == 0 and list(cov) == [30, 30, 5, 5]
    assert lib.fb_debug_tile_cover(512, 5, 1, 16, 16, 256, 256, 512, 512, cov) == 0 and list(cov) == [48, 48, 5, 5]   # pulled back at the far edge


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the CPU path: the oracle port on the host cores) on a two-tile sample: exactly one
    line on stdout, carrying the keys of the measurement contract."""
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-tiles", "2"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "zone_detect Mpixels/s" and d["unit"] == "Mpixels/s"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "10000x10000" in d["config"]["workload"] and d["config"]["tiles"] == 1600
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


_WORKER_MAIN = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from flair1_b200.zone_detect import main as zmain
from flair1_b200.zone_detect.metrics import metrics_from_confmat
from flair1_b200.zone_detect.slicing_job import owned_rects, split_tiles_across_ranks, tile_table, tile_windows
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
W, H, T, M, NC = 1500, 1100, 512, 128, 15
rng = np.random.default_rng(11)
full = rng.integers(0, NC, (2, H, W), dtype=np.uint8)
tiles, wins = tile_table(W, H, T, M), tile_windows(W, H, T, M)
shard = split_tiles_across_ranks(tiles, world)[rank]
mine = tiles[shard]
# the product's own output map: one [bands, H, W] array in shared memory, every rank writes the rectangles it owns
# (on the GPU box fb_detect_zone_shard does these copies from device memory), rank 0 reads the whole map
out = zmain.OutputMap(2, H, W)
assert out.shared is not None
for y0, y1, x0, x1 in owned_rects(mine):
    out.array[:, y0:y1, x0:x1] = full[:, y0:y1, x0:x1]
got = out.finish()
assert (got is None) == (rank != 0)
if rank == 0:
    assert np.array_equal(got, full)
path = out.shared.path
out.close()
dist.barrier()
assert not os.path.exists(path)
# the product's own per-patch metric gather: every rank contributes the confusion matrices of its tiles
classes = {{i + 1: [1 if i < 12 else 0, "class%d" % (i + 1)] for i in range(NC)}}
all_cm = rng.integers(0, 1000, (len(tiles), NC, NC)).astype(np.int64)
cfg = {{"classes": classes, "n_classes": NC, "_my_index": shard, "_my_windows": wins[shard], "_patch_cm": torch.from_numpy(all_cm[shard])}}
res = zmain._gather_patch_metrics(cfg, "m")
if rank == 0:
    ref = [metrics_from_confmat(all_cm[i], cfg, "m_%d_%d" % (wins[i][2], wins[i][3])) for i in range(len(tiles))]
    assert res == ref and len(res) == len(tiles)
else:
    assert res is None
dist.destroy_process_group()
sys.stdout.write("rank %d ok\n" % rank)
sys.stdout.flush()
"""


def test_two_rank_strip_and_patch_metric_gather_of_the_pipeline_gloo(tmp_path):
    """zone_detect.main.OutputMap / _gather_patch_metrics (the multi-GPU tail of run_pipeline) under gloo with two
    ranks sharded by tile ranges: rank 0 ends up with the whole class map (shared memory, no gather) and with every
    tile's metrics in write order."""
    script = tmp_path / "worker_main.py"
    script.write_text(_WORKER_MAIN.format(root=str(ROOT)))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout


def test_header_is_plain_c(tmp_path):
    """include/flair_b200.h is the drop-in boundary: it must compile as C99 (plain pointers and sizes, no C++ / torch types)."""
    src = tmp_path / "hdr.c"
    src.write_text('#include "flair_b200.h"\nint main(void) { fb_tile t = {0, 0, 0, 0, 0, 0}; return t.x0 + (FB_API_VERSION != 1); }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", str(ROOT / "include"), str(src)],
                       capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
