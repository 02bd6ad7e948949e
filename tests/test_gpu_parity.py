"""GPU parity tests proper: every call goes through the C ABI (flair1_b200._native -> libflairb200.so)
and is checked against the CPU oracle (oracle/) or a plain PyTorch fp32 reference on the same seeded
inputs. Tolerances are BASELINE.json's: logits max-abs <= 2e-2 * max|ref|, argmax agreement >= 99.9 %,
confusion matrices bit-exact."""
import numpy as np
import pytest
import torch

from pathlib import Path

pytestmark = pytest.mark.gpu

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"
LOGIT_TOL = 2e-2        # relative to max|ref| (BASELINE.json north_star)
AGREE_MIN = 0.999       # argmax pixel agreement


def _nat():
    import flair1_b200._native as nat
    return nat


# ------------------------------------------------------------------------------------------ kernels
CONV_CASES = [
    # name, B, H, W, C1, Cout, k, stride, pad, mode(0 gather / 1 TMA), kwargs
    ("g_ident", 1, 8, 16, 64, 64, 3, 1, 1, 0, dict(identity=True, relu=False)),
    ("t_ident", 1, 8, 16, 64, 64, 3, 1, 1, 1, dict(identity=True, relu=False)),
    ("g_3x3_64", 2, 32, 32, 64, 64, 3, 1, 1, 0, {}),
    ("t_3x3_64", 2, 32, 32, 64, 64, 3, 1, 1, 1, {}),
    ("t_3x3_128_res", 2, 16, 16, 128, 128, 3, 1, 1, 1, dict(res=True)),
    ("t_3x3_256", 3, 32, 32, 256, 256, 3, 1, 1, 1, {}),
    ("t_3x3_512_res_rowbias", 4, 16, 16, 512, 512, 3, 1, 1, 1, dict(res=True, rowb=True)),
    ("g_3x3_512", 2, 16, 16, 512, 512, 3, 1, 1, 0, {}),
    ("g_stem_7x7s2", 2, 64, 64, 8, 64, 7, 2, 3, 0, {}),
    ("g_3x3s2", 2, 32, 32, 64, 128, 3, 2, 1, 0, {}),
    ("g_1x1s2", 2, 32, 32, 64, 128, 1, 2, 0, 0, dict(relu=False)),
    ("g_up_cat_768_256", 1, 32, 32, 512, 256, 3, 1, 1, 0, dict(C2=256, up1=True)),
    ("g_up_cat_128_32", 1, 64, 64, 64, 32, 3, 1, 1, 0, dict(C2=64, up1=True)),
    ("g_up_32_16", 1, 64, 64, 32, 16, 3, 1, 1, 0, dict(up1=True)),
    ("g_3x3_16", 1, 64, 64, 16, 16, 3, 1, 1, 0, {}),
    ("g_head_f32", 1, 64, 64, 16, 16, 3, 1, 1, 0, dict(relu=False, out_f32=True)),
    ("g_ragged_m", 1, 20, 20, 64, 64, 3, 1, 1, 0, {}),
    ("t_many_tiles", 8, 128, 128, 64, 64, 3, 1, 1, 1, {}),
    ("auto_many_tiles", 3, 64, 64, 128, 128, 3, 1, 1, -1, {}),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_kernel_vs_torch_fp32(ctx, case):
    import gpu_probe
    gpu_probe.RESULTS.clear()
    assert gpu_probe.conv_case(ctx, case[0], *case[1:10], **case[10]), gpu_probe.RESULTS[-1]


def _halo_cases():
    import gpu_probe
    return gpu_probe.HALO_CASES


@pytest.mark.parametrize("idx", range(12))
def test_halo_conv_kernel_vs_torch_fp32(ctx, idx):
    import gpu_probe
    case = gpu_probe.HALO_CASES[idx]
    gpu_probe.RESULTS.clear()
    assert gpu_probe.halo_case(ctx, case[0], *case[1:8], **case[8]), gpu_probe.RESULTS[-1]


@pytest.mark.parametrize("idx", range(8))
def test_halo_depth_to_space_forms_vs_torch_fp32(ctx, idx):
    """dec4.conv2 / head as a 4x4 stride-2 conv over 2x2 cells, dec4.conv1 as a 3x3 conv on the low-res grid with the
    four output phases as 64 accumulator columns: same results as the plain 3x3 conv (of the upsampled input)."""
    import gpu_probe
    case = gpu_probe.D2S_CASES[idx]
    gpu_probe.RESULTS.clear()
    assert gpu_probe.halo_case(ctx, case[0], *case[1:8], **case[8]), gpu_probe.RESULTS[-1]


@pytest.mark.parametrize("idx", range(8))
def test_halo_streamed_weights_vs_torch_fp32(ctx, idx):
    """128 -> 128 channels (layer2, dec1.conv2): halo-staged input, filter bank streamed through a bulk-copy ring; as
    CTA pairs (cta_group::2) where the image is a whole number of 32-row tile pairs, single CTAs otherwise; the
    64-channel pair form of layer1."""
    import gpu_probe
    case = gpu_probe.SB_CASES[idx]
    gpu_probe.RESULTS.clear()
    assert gpu_probe.halo_case(ctx, case[0], *case[1:8], **case[8]), gpu_probe.RESULTS[-1]


def test_dual_source_tma_conv(ctx):
    import gpu_probe
    gpu_probe.RESULTS.clear()
    assert gpu_probe.conv_case(ctx, "t_dual_128p64_64", 2, 32, 32, 128, 64, 3, 1, 1, 1, C2=64), gpu_probe.RESULTS[-1]
    assert gpu_probe.conv_case(ctx, "t_dual_512p256_256", 1, 32, 32, 512, 256, 3, 1, 1, 1, C2=256), gpu_probe.RESULTS[-1]


@pytest.mark.parametrize("idx", range(10))
def test_strided_tma_and_phase_form_convs(ctx, idx):
    """Stride-2 3x3 / 1x1 convs through strided TMA boxes, and decoder conv1 in sub-pixel phase form
    (2x2 taps on the low-res source + 3x3 taps on the skip) against conv(upsample(x1) (+) x2)."""
    import gpu_probe
    case = gpu_probe.TMA_EXTRA_CASES[idx]
    gpu_probe.RESULTS.clear()
    assert gpu_probe.conv_case(ctx, case[0], *case[1:10], **case[10]), gpu_probe.RESULTS[-1]


def test_cta_pair_kernel_matches(monkeypatch):
    """tcgen05.mma.cta_group::2 variant (FB_PAIR=1): a cluster of two CTAs per 16 x 16 pixel tile, weights
    split across the pair. Opt-in because it measured no faster; must give the same results."""
    import gpu_probe
    nat = _nat()
    monkeypatch.setenv("FB_PAIR", "1")
    c = nat.Context(0)
    gpu_probe.RESULTS.clear()
    try:
        for case in (("p_3x3_128_res", 2, 16, 16, 128, 128, 3, 1, 1, 1, dict(res=True)),
                     ("p_3x3_256", 3, 32, 32, 256, 256, 3, 1, 1, 1, {}),
                     ("p_3x3_512_res_rowbias", 4, 16, 16, 512, 512, 3, 1, 1, 1, dict(res=True, rowb=True)),
                     ("p_dual_512p256_256", 1, 32, 32, 512, 256, 3, 1, 1, 1, dict(C2=256)),
                     ("p_many", 5, 64, 64, 128, 128, 3, 1, 1, 1, {})):
            assert gpu_probe.conv_case(c, case[0], *case[1:10], **case[10]), gpu_probe.RESULTS[-1]
    finally:
        c.close()


def test_extract_normalise_bit_exact(ctx, trained_3_15):
    """K1 against the oracle's float64 -> float32 normalisation rounded to bf16, including tiles that
    hang over every edge of the raster (boundless zero fill before normalisation)."""
    from oracle import synth
    from oracle.zone_detect_ref import normalization
    sd, _ = trained_3_15
    W, H, T = 300, 200, 64
    raster = synth.synth_raster(5, H, W, seed=3)
    ctx.load_weights(sd, 3, 15)
    for norm_type in ("custom", "scaling"):
        means, stds = [synth.FLAIR_MEANS[i] for i in (3, 0, 2)], [synth.FLAIR_STDS[i] for i in (3, 0, 2)]
        ctx.set_norm(norm_type, means, stds)
        ctx.set_raster(torch.from_numpy(raster).cuda(), [3, 0, 2], W, H)
        xy = np.array([[-30, -20], [W - 40, H - 30], [100, 50], [-64, 10], [W, H]], np.int32)
        ctx.forward_tiles(xy, T)
        got = ctx.debug_input_tiles().float().cpu().numpy()      # [n, T, T, 3], padding lanes checked to be 0
        assert got.shape == (len(xy), T, T, 3)
        for i, (x0, y0) in enumerate(xy):
            patch = np.zeros((3, T, T), np.uint8)
            r0, r1, c0, c1 = max(y0, 0), min(y0 + T, H), max(x0, 0), min(x0 + T, W)
            if r1 > r0 and c1 > c0:
                patch[:, r0 - y0:r1 - y0, c0 - x0:c1 - x0] = raster[[3, 0, 2], r0:r1, c0:c1]
            ref = torch.as_tensor(normalization(patch, norm_type, means, stds), dtype=torch.float).to(torch.bfloat16).float().numpy()
            np.testing.assert_array_equal(got[i].transpose(2, 0, 1), ref)
    # HWC layout gives the same tiles
    ctx.set_norm("custom", means, stds)
    ctx.set_raster(torch.from_numpy(np.ascontiguousarray(raster.transpose(1, 2, 0))).cuda(), [3, 0, 2], W, H, layout=1)
    ctx.forward_tiles(xy, T)
    got_hwc = ctx.debug_input_tiles().float().cpu().numpy()
    ctx.set_raster(torch.from_numpy(raster).cuda(), [3, 0, 2], W, H)
    ctx.forward_tiles(xy, T)
    np.testing.assert_array_equal(got_hwc, ctx.debug_input_tiles().float().cpu().numpy())


def test_space_to_depth_stem_matches_plain_stem(trained_3_15, monkeypatch):
    """<= 4 bands: the stem runs as a 4x4 stride-1 conv on the 2x2 space-to-depth tile (16 instead of 28 MMA steps,
    half the input bytes). Same bf16 products as the 7x7 stride-2 form, other summation order: the stem output may
    differ by a bf16 rounding, the input tiles are bit-identical, and the logits stay within the tolerance."""
    from oracle import synth
    nat = _nat()
    sd, _ = trained_3_15
    raster = torch.from_numpy(synth.synth_raster(3, 600, 700, seed=9)).cuda()
    xy = np.array([[0, 0], [-100, 37], [300, 200], [700 - 256, 600 - 256]], np.int32)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("FB_NO_S2D", mode)
        c = nat.Context(0)
        c.load_weights(sd, 3, 15)
        c.set_norm("custom", synth.FLAIR_MEANS[:3], synth.FLAIR_STDS[:3])
        c.set_raster(raster, [0, 1, 2], 700, 600)
        logits = c.forward_tiles(xy, 512).cpu()
        res[mode] = (c.debug_input_tiles().cpu(), c.debug_activation("f1").float().cpu(), logits, tuple(c.debug_activation("x0").shape))
        c.close()
    assert res["1"][3] == (4, 512, 512, 8) and res["0"][3] == (4, 256, 256, 16)
    assert torch.equal(res["0"][0], res["1"][0])
    f_rel = (res["0"][1] - res["1"][1]).abs().max().item() / res["1"][1].abs().max().item()
    l_rel = (res["0"][2] - res["1"][2]).abs().max().item() / res["1"][2].abs().max().item()
    print(f"space-to-depth stem vs 7x7 stem: f1 rel diff {f_rel:.3e}, logits rel diff {l_rel:.3e}")
    assert f_rel < 1e-2 and l_rel < 1e-2
    assert (res["0"][2][..., :15].argmax(-1) == res["1"][2][..., :15].argmax(-1)).float().mean().item() > 0.999


def test_confusion_bit_exact_vs_sklearn(ctx):
    from oracle.metrics_ref import confusion_numpy, patch_confusion
    z = np.load(__import__("conftest").GOLDEN / "confusion.npz")
    cm = ctx.confusion(torch.from_numpy(z["pred"]).cuda(), torch.from_numpy(z["truth"]).cuda(), 19, truth_sub=1)
    np.testing.assert_array_equal(cm.cpu().numpy(), z["cm"])          # the reference's own sklearn call
    g = torch.Generator().manual_seed(5)
    for npx, ncls in ((0, 15), (1, 15), (17, 13), (4099, 19), (1 << 21, 19), ((1 << 22) + 5, 15)):
        pred = torch.randint(0, 22, (npx,), generator=g, dtype=torch.uint8)
        truth = torch.randint(0, 22, (npx,), generator=g, dtype=torch.uint8)
        got = ctx.confusion(pred.cuda(), truth.cuda(), ncls, truth_sub=1).cpu().numpy()
        ref = confusion_numpy(truth.numpy(), pred.numpy(), ncls, 1)
        np.testing.assert_array_equal(got, ref)
        assert got.dtype == np.int64
        if npx in (4099, 1 << 21):
            np.testing.assert_array_equal(got, patch_confusion(truth.numpy() - 1, pred.numpy(), ncls))
    # unaligned views take the scalar path; accumulation into an existing matrix is additive (linearity)
    pred = torch.randint(0, 19, (100003,), generator=g, dtype=torch.uint8).cuda()
    truth = torch.randint(1, 20, (100003,), generator=g, dtype=torch.uint8).cuda()
    whole = ctx.confusion(pred, truth, 19, 1)
    parts = ctx.confusion(pred[:33333].contiguous(), truth[:33333].contiguous(), 19, 1)
    parts = ctx.confusion(pred[33333:], truth[33333:], 19, 1, out=parts)   # offset 33333: not 16-byte aligned
    assert torch.equal(whole, parts)
    # a patch whose labels are all out of range contributes nothing (sklearn raises there and the
    # reference skips the patch: flair/metrics.py:73-74)
    none = ctx.confusion(torch.full((1000,), 3, dtype=torch.uint8).cuda(), torch.zeros(1000, dtype=torch.uint8).cuda(), 19, 1)
    assert int(none.sum()) == 0


# ------------------------------------------------------------------------------------------ network
def _oracle_inputs(raster, xy, T, bands, means, stds, norm_type="custom"):
    from oracle.zone_detect_ref import normalization
    H, W = raster.shape[1:]
    imgs = []
    for x0, y0 in xy:
        patch = np.zeros((len(bands), T, T), np.uint8)
        r0, r1, c0, c1 = max(y0, 0), min(y0 + T, H), max(x0, 0), min(x0 + T, W)
        if r1 > r0 and c1 > c0:
            patch[:, r0 - y0:r1 - y0, c0 - x0:c1 - x0] = raster[bands, r0:r1, c0:c1]
        imgs.append(torch.as_tensor(normalization(patch, norm_type, means, stds), dtype=torch.float))
    return torch.stack(imgs)


def test_forward_logits_and_argmax_vs_oracle(ctx, trained_3_15):
    """Config 2's model (3 bands / 15 classes, trained-like weights): logits within 2e-2 of max|ref|,
    argmax agreement >= 99.9 %, per-layer error reported."""
    from oracle import synth
    from oracle.unet_smp033 import layer_activations
    sd, model = trained_3_15
    W, H, T = 1100, 900, 512
    raster = synth.synth_raster(3, H, W, seed=21)
    means, stds = synth.FLAIR_MEANS[:3], synth.FLAIR_STDS[:3]
    ctx.load_weights(sd, 3, 15)
    ctx.set_norm("custom", means, stds)
    ctx.set_raster(torch.from_numpy(raster).cuda(), [0, 1, 2], W, H)
    xy = np.array([[300, 200], [-128, -128], [W - 384, H - 384]], np.int32)
    logits = ctx.forward_tiles(xy, T).cpu()
    x = _oracle_inputs(raster, xy, T, [0, 1, 2], means, stds)
    acts = layer_activations(model, x)
    ref = acts["logits"]
    got = logits.permute(0, 3, 1, 2)[:, :15]
    rel = (got - ref).abs().max().item() / ref.abs().max().item()
    agree = (got.argmax(1) == ref.argmax(1)).float().mean().item()
    print(f"logits rel err {rel:.4e}, argmax agreement {agree * 100:.4f}%")
    assert rel <= LOGIT_TOL
    assert agree >= AGREE_MIN
    assert (logits[..., 15:] == 0).all()
    for name in ("f1", "layer1.2.out", "layer2.3.out", "layer3.5.out", "layer4.2.out", "dec0", "dec2", "dec4"):
        r = acts[name]
        g = ctx.debug_activation(name).float().cpu().permute(0, 3, 1, 2)
        if g.shape[-1] == 2 * r.shape[-1]:  # stored 2x2-replicated for the next decoder block
            r = r.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
        e = (g - r).abs().max().item() / r.abs().max().item()
        print(f"  {name:14s} rel err {e:.4e}")
        assert e <= LOGIT_TOL, name


def test_gather_and_tma_producers_agree(trained_3_15, monkeypatch):
    """Same network through the cp.async im2col gather producer only (FB_FORCE_GATHER=1) against the
    default mix of halo-staged / TMA / gather kernels."""
    from oracle import synth
    nat = _nat()
    sd, _ = trained_3_15
    raster = torch.from_numpy(synth.synth_raster(3, 512, 512, seed=2)).cuda()
    outs = []
    for force in ("0", "1"):
        monkeypatch.setenv("FB_FORCE_GATHER", force)
        c = nat.Context(0)
        c.load_weights(sd, 3, 15)
        c.set_norm("custom", synth.FLAIR_MEANS[:3], synth.FLAIR_STDS[:3])
        c.set_raster(raster, [0, 1, 2], 512, 512)
        outs.append(c.forward_tiles(np.array([[0, 0], [-100, 37]], np.int32), 512).cpu())
        c.close()
    rel = (outs[0] - outs[1]).abs().max().item() / outs[1].abs().max().item()
    print(f"fast paths vs gather-only: logits rel diff {rel:.3e}")
    # same bf16 operands and fp32 accumulation, only the summation order inside a K loop differs
    assert rel < 1e-2
    assert (outs[0][..., :15].argmax(-1) == outs[1][..., :15].argmax(-1)).float().mean().item() > 0.999


def test_five_band_metadata_model_logits(ctx):
    """Config 3: 5-band / 13-class U-Net with the metadata MLP, random init (torch defaults,
    manual_seed(3)), batched 512^2 patches: logits tolerance only (argmax is near-degenerate)."""
    from oracle import synth
    from oracle.unet_smp033 import FlairModel
    torch.manual_seed(3)
    model = FlairModel(5, 13, True).eval()
    sd = {k.replace("seg_model.", "", 1) if k.startswith("seg_model.") else k: v for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(3)
    B, T = 3, 512
    patches = torch.randint(0, 256, (B, 5, T, T), generator=g, dtype=torch.uint8)
    met = torch.rand((B, 45), generator=g)
    ctx.load_weights(sd, 5, 13, use_metadata=True)
    ctx.set_norm("custom", synth.FLAIR_MEANS, synth.FLAIR_STDS)
    from oracle.flair_ref import norm
    x = torch.stack([torch.as_tensor(norm(p.numpy(), "custom", synth.FLAIR_MEANS, synth.FLAIR_STDS), dtype=torch.float) for p in patches])
    with torch.no_grad():
        ref = model(x, met)
    # forward through the zone entry point on a raster made of the stacked patches
    raster = patches.permute(1, 0, 2, 3).reshape(5, B * T, T).contiguous().cuda()
    ctx.set_raster(raster, [0, 1, 2, 3, 4], T, B * T)
    xy = np.array([[0, i * T] for i in range(B)], np.int32)
    got = ctx.forward_tiles(xy, T, metadata=met.numpy()).cpu().permute(0, 3, 1, 2)[:, :13]
    rel = (got - ref).abs().max().item() / ref.abs().max().item()
    print(f"5-band/13-class + metadata: logits rel err {rel:.4e}; agreement {(got.argmax(1) == ref.argmax(1)).float().mean().item() * 100:.3f}% (not asserted)")
    assert rel <= LOGIT_TOL
    # patch-predict entry point gives the argmax of the same logits
    cls = ctx.predict_patches(patches.cuda(), T, 2, metadata=met.numpy()).cpu()
    assert torch.equal(cls.long(), got.argmax(1))
    with pytest.raises(nat_error()):
        ctx.forward_tiles(xy, T)      # metadata model without metadata must fail loudly


@pytest.mark.parametrize("bands,ncls", [(1, 7), (2, 15), (4, 16), (6, 13), (8, 19)])
def test_other_band_counts_logits(ctx, bands, ncls):
    """1, 2 and 4 bands take the space-to-depth stem (one extract instantiation each), 6 and 8 the 7x7 stride-2 stem on
    the 8-channel tile; class counts at the 16 / 17 boundary. Random-init models, tiles hanging over the raster edge:
    logits within the tolerance of the fp32 oracle."""
    from oracle.unet_smp033 import Unet
    from oracle.zone_detect_ref import normalization
    torch.manual_seed(100 + bands)
    model = Unet(bands, ncls).eval()
    g = torch.Generator().manual_seed(bands)
    W, H, T = 400, 300, 256
    raster = torch.randint(0, 256, (bands, H, W), generator=g, dtype=torch.uint8)
    means = [100.0 + 3 * i for i in range(bands)]
    stds = [40.0 + 2 * i for i in range(bands)]
    ctx.load_weights(model.state_dict(), bands, ncls)
    ctx.set_norm("custom", means, stds)
    ctx.set_raster(raster.cuda(), list(range(bands)), W, H)
    xy = np.array([[-40, -30], [W - 200, H - 180], [72, 20]], np.int32)
    got = ctx.forward_tiles(xy, T).cpu()
    assert got.shape[-1] == (16 if ncls <= 16 else 32) and (got[..., ncls:] == 0).all()
    patches = []
    for x0, y0 in xy:
        patch = np.zeros((bands, T, T), np.uint8)
        r0, r1, c0, c1 = max(y0, 0), min(y0 + T, H), max(x0, 0), min(x0 + T, W)
        patch[:, r0 - y0:r1 - y0, c0 - x0:c1 - x0] = raster.numpy()[:, r0:r1, c0:c1]
        patches.append(torch.as_tensor(normalization(patch, "custom", means, stds), dtype=torch.float))
    with torch.no_grad():
        ref = model(torch.stack(patches))
    rel = (got.permute(0, 3, 1, 2)[:, :ncls] - ref).abs().max().item() / ref.abs().max().item()
    print(f"{bands} bands / {ncls} classes: logits rel err {rel:.4e}")
    assert rel <= LOGIT_TOL


def test_nineteen_class_model(ctx):
    """Configs 1/5 of BASELINE.json name the 19-class nomenclature: n_classes > 16 switches logits and blend
    accumulators to 32 floats per pixel (fb_logit_stride) and the head to a 32-column accumulator. Random-init
    5-band / 19-class model (torch defaults, manual_seed(5)): logits tolerance; patch predict = arg-max of the same
    logits; and, on a briefly trained 3-band / 19-class checkpoint, the zone loop (exact clipping, class_prob and
    the three blended stitchings) against the oracle."""
    from oracle import synth
    from oracle.flair_ref import norm
    from oracle.unet_smp033 import Unet
    from oracle.zone_detect_ref import GeoRaster, run_zone, run_zone_blend, run_zone_class_prob
    from flair1_b200.zone_detect.slicing_job import tile_table
    torch.manual_seed(5)
    model = Unet(5, 19).eval()
    g = torch.Generator().manual_seed(5)
    B, T = 2, 512
    patches = torch.randint(0, 256, (B, 5, T, T), generator=g, dtype=torch.uint8)
    ctx.load_weights(model.state_dict(), 5, 19)
    assert ctx.logit_stride == 32
    ctx.set_norm("custom", synth.FLAIR_MEANS, synth.FLAIR_STDS)
    x = torch.stack([torch.as_tensor(norm(p.numpy(), "custom", synth.FLAIR_MEANS, synth.FLAIR_STDS), dtype=torch.float) for p in patches])
    with torch.no_grad():
        ref = model(x)
    raster = patches.permute(1, 0, 2, 3).reshape(5, B * T, T).contiguous().cuda()
    ctx.set_raster(raster, [0, 1, 2, 3, 4], T, B * T)
    xy = np.array([[0, i * T] for i in range(B)], np.int32)
    full = ctx.forward_tiles(xy, T).cpu()
    assert full.shape == (B, T, T, 32) and (full[..., 19:] == 0).all()
    got = full.permute(0, 3, 1, 2)[:, :19]
    rel = (got - ref).abs().max().item() / ref.abs().max().item()
    print(f"5-band/19-class: logits rel err {rel:.4e}; agreement {(got.argmax(1) == ref.argmax(1)).float().mean().item() * 100:.3f}% (not asserted)")
    assert rel <= LOGIT_TOL
    cls = ctx.predict_patches(patches.cuda(), T, 2).cpu()
    assert torch.equal(cls.long(), got.argmax(1))

    # zone loop on a trained 19-class checkpoint
    sd = synth.cached_checkpoint(3, 19)
    m = Unet(3, 19)
    m.load_state_dict(sd, strict=True)
    m.eval()
    W, H, T, margin = 700, 560, 256, 32
    raster = synth.synth_raster(3, H, W, seed=19)
    means, stds = synth.FLAIR_MEANS[:3], synth.FLAIR_STDS[:3]
    ctx.load_weights(sd, 3, 19)
    ctx.set_norm("custom", means, stds)
    ctx.set_raster(torch.from_numpy(raster).cuda(), [0, 1, 2], W, H)
    tiles = tile_table(W, H, T, margin)
    config = {"img_pixels_detection": T, "margin": margin, "channels": [1, 2, 3], "n_classes": 19,
              "norma_task": [{"norm_type": "custom", "norm_means": means, "norm_stds": stds}]}
    geo = GeoRaster(raster, 800000.0, 6500000.0 + H * 0.2, 0.2)
    cls = torch.full((H, W), 255, dtype=torch.uint8, device="cuda")
    conf = torch.full((H, W), 255, dtype=torch.uint8, device="cuda")
    ctx.detect_strip(tiles, T, 5, cls, conf, W, 0)
    ref_cls, ref_conf, _ = run_zone(m, geo, config)
    agree = (cls.cpu().numpy() == ref_cls).mean()
    print(f"19-class zone {W}x{H}: agreement {agree * 100:.4f}%, classes present {np.unique(ref_cls).size}")
    assert agree >= AGREE_MIN and (cls.cpu().numpy() < 19).all()
    prob = torch.zeros((19, H, W), dtype=torch.uint8, device="cuda")
    ctx.detect_strip_prob(tiles, T, 4, prob, W, 0)
    diff = np.abs(prob.cpu().numpy().astype(np.int32) - run_zone_class_prob(m, geo, config).astype(np.int32))
    assert (diff <= 4).mean() >= 0.999
    for method in ("average", "average_weights", "max"):
        acc, wsum = ctx.blend_buffers(method, H, W)
        ctx.blend_strip(tiles, T, 4, method, acc, wsum, W, 0)
        bc = torch.zeros((H, W), dtype=torch.uint8, device="cuda")
        ctx.blend_finalize(method, acc, wsum, bc, None)
        a = (bc.cpu().numpy() == run_zone_blend(m, geo, config, method)[0]).mean()
        print(f"19-class blend {method}: agreement {a * 100:.4f}%")
        assert a >= AGREE_MIN


def nat_error():
    return _nat().NativeError


@pytest.mark.parametrize("W,H,T,margin", [(1000, 700, 512, 128), (700, 520, 256, 32), (640, 512, 512, 0), (300, 280, 512, 128)])
def test_zone_detect_vs_oracle(ctx, trained_3_15, W, H, T, margin):
    """Whole small zones, every tile: class map agreement >= 99.9 % with the oracle's run_zone, every
    pixel written, clamped last row/column resolved with the reference's write order."""
    from oracle import synth
    from oracle.zone_detect_ref import GeoRaster, run_zone
    from flair1_b200.zone_detect.slicing_job import tile_table
    sd, model = trained_3_15
    raster = synth.synth_raster(3, H, W, seed=W + H)
    means, stds = synth.FLAIR_MEANS[:3], synth.FLAIR_STDS[:3]
    ctx.load_weights(sd, 3, 15)
    ctx.set_norm("custom", means, stds)
    ctx.set_raster(torch.from_numpy(raster).cuda(), [0, 1, 2], W, H)
    tiles = tile_table(W, H, T, margin)
    cls = torch.full((H, W), 255, dtype=torch.uint8, device="cuda")
    conf = torch.full((H, W), 255, dtype=torch.uint8, device="cuda")
    ctx.detect_strip(tiles, T, 5, cls, conf, W, 0)          # batch 5: exercises a ragged last batch
    config = {"img_pixels_detection": T, "margin": margin, "channels": [1, 2, 3], "n_classes": 15,
              "norma_task": [{"norm_type": "custom", "norm_means": means, "norm_stds": stds}]}
    ref_cls, ref_conf, rows = run_zone(model, GeoRaster(raster, 800000.0, 6500000.0 + H * 0.2, 0.2), config)
    assert len(rows) == len(tiles)
    cls_h, conf_h = cls.cpu().numpy(), conf.cpu().numpy()
    assert (cls_h < 15).all() and (conf_h <= 1).all()
    agree = (cls_h == ref_cls).mean()
    print(f"zone {W}x{H} T={T} m={margin}: {len(tiles)} tiles, agreement {agree * 100:.4f}%, conf agreement {(conf_h == ref_conf).mean() * 100:.4f}%")
    assert agree >= AGREE_MIN
    assert (conf_h == ref_conf).mean() >= 0.995
    # idempotence + host entry point: same bytes again
    out_cls = np.zeros((H, W), np.uint8)
    ctx.detect_zone_host(raster, [0, 1, 2], W, H, 0, 0, tiles, T, 4, out_cls, None, W, 0, H)
    np.testing.assert_array_equal(out_cls, cls_h)


def test_no_writes_outside_the_caller_buffers(ctx, trained_3_15):
    """compute-sanitizer is closed on the GPU pool (profiles/r02_sanitizer_closed.txt), so the out-of-bounds check
    is done with canaries: every buffer the caller hands in (class / confidence maps, class_prob planes, blend
    accumulators, confusion matrix, truth, raster) sits inside a larger allocation filled with a pattern, and the
    bytes before and after it must come back untouched from the exact-clipping loop (fused sink, CTA-pair and
    streamed-weight convs, depth-to-space head), the class_prob loop, both blended stitchings and the confusion
    kernel. A 637 x 509 zone: every edge tile is clamped, no dimension is a multiple of a vector width."""
    from oracle import synth
    from flair1_b200.zone_detect.slicing_job import tile_table
    sd, _ = trained_3_15
    W, H, T, margin, G = 637, 509, 256, 32, 4096
    raster_h = synth.synth_raster(3, H, W, seed=5)
    ctx.load_weights(sd, 3, 15)
    ctx.set_norm("custom", synth.FLAIR_MEANS[:3], synth.FLAIR_STDS[:3])

    def guarded(shape, dtype, fill=None):
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        raw = torch.full((G + n + G,), 0xA5, dtype=torch.uint8, device="cuda")
        view = raw[G:G + n].view(dtype).view(*shape)
        if fill is not None:
            view.copy_(fill)
        return raw, view, n

    def intact(raw, n):
        return bool((raw[:G] == 0xA5).all()) and bool((raw[G + n:] == 0xA5).all())

    bufs = {}
    bufs["raster"] = guarded((3, H, W), torch.uint8, torch.from_numpy(raster_h).cuda())
    ctx.set_raster(bufs["raster"][1], [0, 1, 2], W, H)
    tiles = tile_table(W, H, T, margin)
    bufs["cls"] = guarded((H, W), torch.uint8, 0)
    bufs["conf"] = guarded((H, W), torch.uint8, 0)
    ctx.detect_strip(tiles, T, 5, bufs["cls"][1], bufs["conf"][1], W, 0)
    bufs["prob"] = guarded((15, H, W), torch.uint8, 0)
    ctx.detect_strip_prob(tiles, T, 5, bufs["prob"][1], W, 0)
    bufs["truth"] = guarded((H, W), torch.uint8, torch.from_numpy(synth.synth_mask(raster_h, 15, 3)).cuda())
    bufs["cm"] = guarded((15, 15), torch.int64, 0)
    ctx.confusion(bufs["cls"][1], bufs["truth"][1], 15, truth_sub=1, out=bufs["cm"][1])
    ls = ctx.logit_stride
    bufs["acc"] = guarded((H, W, ls), torch.float32, 0)
    bufs["wsum"] = guarded((H, W), torch.float32, 0)
    ctx.blend_strip(tiles, T, 5, "average_weights", bufs["acc"][1], bufs["wsum"][1], W, 0)
    bufs["bcls"] = guarded((H, W), torch.uint8, 0)
    bufs["bconf"] = guarded((H, W), torch.uint8, 0)
    ctx.blend_finalize("average_weights", bufs["acc"][1], bufs["wsum"][1], bufs["bcls"][1], bufs["bconf"][1])
    bufs["key"] = guarded((H, W), torch.int64, 0)
    ctx.blend_strip(tiles, T, 5, "max", bufs["key"][1], None, W, 0)
    ctx.blend_finalize("max", bufs["key"][1], None, bufs["bcls"][1], bufs["bconf"][1])
    torch.cuda.synchronize()
    broken = [k for k, (raw, _, n) in bufs.items() if not intact(raw, n)]
    assert not broken, f"bytes outside these buffers were overwritten: {broken}"
    assert int(bufs["cm"][1].sum()) > 0 and bool((bufs["cls"][1] < 15).all())


@pytest.mark.parametrize("W,H,T,margin", [(1500, 1300, 512, 128), (700, 520, 256, 32), (640, 512, 512, 0)])
def test_dead_output_elimination_and_fused_sink_are_bit_exact(trained_3_15, monkeypatch, W, H, T, margin):
    """The exact-clipping loop only computes the decoder outputs inside the receptive field of each write
    rectangle (csrc/tile_need.cuh) and lets the head's epilogue write the class / confidence bytes. Both
    must leave the maps bit-identical to the plain path (every tile in full, fp32 logits, separate K6
    kernel), for class map, confidence band and class_prob planes; the FLOP counter shows the saving and is
    63.569 GFLOP per 512^2 tile on the plain path."""
    from oracle import synth
    from flair1_b200.zone_detect.slicing_job import tile_table
    nat = _nat()
    sd, _ = trained_3_15
    raster = torch.from_numpy(synth.synth_raster(3, H, W, seed=W * 3 + H)).cuda()
    tiles = tile_table(W, H, T, margin)
    res = {}
    for mode in ("plain", "fast", "aligned"):   # aligned: active kernel tiles on the fixed tile grid (FB_ALIGNED_TILES=1)
        monkeypatch.setenv("FB_FULL_TILES", "1" if mode == "plain" else "0")
        monkeypatch.setenv("FB_NO_FUSED_SINK", "1" if mode == "plain" else "0")
        monkeypatch.setenv("FB_ALIGNED_TILES", "1" if mode == "aligned" else "0")
        c = nat.Context(0)
        c.load_weights(sd, 3, 15)
        c.set_norm("custom", synth.FLAIR_MEANS[:3], synth.FLAIR_STDS[:3])
        c.set_raster(raster, [0, 1, 2], W, H)
        cls = torch.full((H, W), 255, dtype=torch.uint8, device="cuda")
        conf = torch.full((H, W), 255, dtype=torch.uint8, device="cuda")
        f0 = c.flop_count
        c.detect_strip(tiles, T, 7, cls, conf, W, 0)
        torch.cuda.synchronize()
        flops = c.flop_count - f0
        prob = torch.zeros((15, H, W), dtype=torch.uint8, device="cuda")
        c.detect_strip_prob(tiles, T, 7, prob, W, 0)
        res[mode] = (cls.cpu().numpy(), conf.cpu().numpy(), prob.cpu().numpy(), flops)
        c.close()
    np.testing.assert_array_equal(res["fast"][0], res["plain"][0])
    np.testing.assert_array_equal(res["fast"][1], res["plain"][1])
    np.testing.assert_array_equal(res["fast"][2], res["plain"][2])
    for k in range(3):
        np.testing.assert_array_equal(res["aligned"][k], res["plain"][k])
    assert res["fast"][3] <= res["aligned"][3]   # origin-shifted tiles never compute more than grid-aligned ones
    assert (res["fast"][0] < 15).all()
    per_tile = res["plain"][3] / len(tiles) / 1e9
    print(f"zone {W}x{H} T={T} m={margin}: {len(tiles)} tiles, {per_tile:.3f} GFLOP per tile in full, "
          f"{res['aligned'][3] / len(tiles) / 1e9:.3f} with dead-output elimination on the fixed tile grid, "
          f"{res['fast'][3] / len(tiles) / 1e9:.3f} with origin-shifted tiles")
    if T == 512:
        assert abs(per_tile - 63.569) < 0.01
    assert res["fast"][3] <= res["plain"][3]
    if margin >= 128:   # small tiles / thin margins save little: the regions grow by one pixel per 3x3 conv
        assert res["fast"][3] < 0.9 * res["plain"][3]


@pytest.mark.parametrize("bands,T,margin", [(3, 128, 32), (5, 128, 32), (3, 512, 128)])
def test_fused_stem_max_pool_is_bit_exact(trained_3_15, monkeypatch, bands, T, margin):
    """With a batch that gives most SMs an image of their own, the stem's epilogue pools its own output (csrc/conv_halo.cu,
    POOL) and, in the exact-clipping loop, stores only the part of it the decoder reads. Pooled tensor, logits, class map
    and confidence band must be byte-identical to the separate max-pool kernel (FB_NO_POOL_FUSE=1) for the
    space-to-depth stem (<= 4 bands), on edge tiles and interior tiles; a 5-band model (7x7 stride-2 stem, no room for
    the pool buffers in shared memory) must be unaffected by the switch."""
    from oracle import synth
    from flair1_b200.zone_detect.slicing_job import tile_table
    nat = _nat()
    ncls = 15 if bands == 3 else 13
    if bands == 3:
        sd = trained_3_15[0]
    else:
        from oracle.unet_smp033 import Unet
        torch.manual_seed(5)
        sd = Unet(bands, ncls).state_dict()   # random init: this test compares two GPU paths, not the oracle
    n = 80 if T == 128 else 76
    W, H = (T - 2 * margin) * 10 + 37, (T - 2 * margin) * 8 + 11
    raster = torch.from_numpy(synth.synth_raster(bands, H, W, seed=bands * T)).cuda()
    tiles = tile_table(W, H, T, margin)[:n]
    assert len(tiles) == n
    res = {}
    for mode in ("separate", "fused"):
        monkeypatch.setenv("FB_NO_POOL_FUSE", "1" if mode == "separate" else "0")
        c = nat.Context(0)
        c.load_weights(sd, bands, ncls)
        c.set_norm("custom", synth.FLAIR_MEANS[:bands], synth.FLAIR_STDS[:bands])
        c.set_raster(raster, list(range(bands)), W, H)
        l0 = c.launch_count
        logits = c.forward_tiles(np.ascontiguousarray(tiles[:, :2]), T)
        launches = c.launch_count - l0
        pool = c.debug_activation("pool").clone()
        f1 = c.debug_activation("f1").clone()
        cls = torch.full((H, W), 255, dtype=torch.uint8, device="cuda")
        conf = torch.full((H, W), 255, dtype=torch.uint8, device="cuda")
        c.detect_strip(tiles, T, n, cls, conf, W, 0)
        torch.cuda.synchronize()
        res[mode] = (pool.cpu(), f1.cpu(), logits.cpu(), cls.cpu(), conf.cpu(), launches)
        c.close()
    # (the 7x7 stride-2 stem of > 4-band models has no room for the pool buffers: same launches there)
    assert res["fused"][5] == res["separate"][5] - (1 if bands <= 4 else 0), "the fused run must launch one kernel less (no max-pool)"
    for k, name in enumerate(("pool", "f1 (whole-tile forward)", "logits", "class map", "confidence")):
        assert torch.equal(res["fused"][k], res["separate"][k]), name


def test_fused_sink_for_more_than_16_classes(monkeypatch):
    """19-class nomenclature (32 logit columns): the head's fused sink takes the soft-max in two halves of 16 columns and
    rescales the first half's exponent sum, K6 sums once against the global maximum. The class byte must be identical;
    the confidence byte (round-half-up of the max probability) may only differ where that probability sits within
    rounding of 0.5."""
    from oracle import synth
    from flair1_b200.zone_detect.slicing_job import tile_table
    nat = _nat()
    sd = synth.cached_checkpoint(3, 19)
    W, H, T, margin = 900, 700, 256, 32
    raster = torch.from_numpy(synth.synth_raster(3, H, W, seed=19)).cuda()
    tiles = tile_table(W, H, T, margin)
    res = {}
    for mode in ("separate", "fused"):
        monkeypatch.setenv("FB_NO_FUSED_SINK", "1" if mode == "separate" else "0")
        c = nat.Context(0)
        c.load_weights(sd, 3, 19)
        c.set_norm("custom", synth.FLAIR_MEANS[:3], synth.FLAIR_STDS[:3])
        c.set_raster(raster, [0, 1, 2], W, H)
        cls = torch.full((H, W), 255, dtype=torch.uint8, device="cuda")
        conf = torch.full((H, W), 255, dtype=torch.uint8, device="cuda")
        c.detect_strip(tiles, T, 6, cls, conf, W, 0)
        torch.cuda.synchronize()
        res[mode] = (cls.cpu(), conf.cpu())
        c.close()
    assert torch.equal(res["fused"][0], res["separate"][0])
    assert int(res["fused"][0].max()) < 19
    differ = int((res["fused"][1] != res["separate"][1]).sum())
    assert differ <= W * H // 100000, f"{differ} confidence bytes differ"


def _zone_setup(ctx, trained_3_15, W, H, T, margin, seed):
    from oracle import synth
    from oracle.zone_detect_ref import GeoRaster
    from flair1_b200.zone_detect.slicing_job import tile_table
    sd, model = trained_3_15
    raster = synth.synth_raster(3, H, W, seed=seed)
    means, stds = synth.FLAIR_MEANS[:3], synth.FLAIR_STDS[:3]
    ctx.load_weights(sd, 3, 15)
    ctx.set_norm("custom", means, stds)
    ctx.set_raster(torch.from_numpy(raster).cuda(), [0, 1, 2], W, H)
    config = {"img_pixels_detection": T, "margin": margin, "channels": [1, 2, 3], "n_classes": 15,
              "norma_task": [{"norm_type": "custom", "norm_means": means, "norm_stds": stds}]}
    return model, GeoRaster(raster, 800000.0, 6500000.0 + H * 0.2, 0.2), config, tile_table(W, H, T, margin)


def test_zone_class_prob_vs_oracle(ctx, trained_3_15):
    """SURVEY 8(f)2, output_type class_prob: n_classes planes of uint8(p * 255). The probabilities come from
    bf16 logits, so bytes may differ by a few counts from the fp32 oracle: tolerance 4/255 on 99.9 % of the
    values, mean absolute difference below 0.5 counts, every pixel written, planes sum to ~255."""
    from oracle.zone_detect_ref import run_zone_class_prob
    W, H, T, margin = 900, 640, 512, 128
    model, georaster, config, tiles = _zone_setup(ctx, trained_3_15, W, H, T, margin, seed=77)
    prob = torch.zeros((15, H, W), dtype=torch.uint8, device="cuda")
    ctx.detect_strip_prob(tiles, T, 4, prob, W, 0)
    got = prob.cpu().numpy().astype(np.int32)
    ref = run_zone_class_prob(model, georaster, config).astype(np.int32)
    diff = np.abs(got - ref)
    print(f"class_prob {W}x{H}: mean |diff| {diff.mean():.4f} counts, max {diff.max()}, within 4: {(diff <= 4).mean() * 100:.4f}%")
    assert (diff <= 4).mean() >= 0.999 and diff.mean() < 0.5
    s = got.sum(axis=0)
    assert s.min() >= 255 - 15 and s.max() <= 255            # truncation loses < 1 count per class
    assert (got.argmax(axis=0) == ref.argmax(axis=0)).mean() >= 0.995   # ties after quantisation are common


@pytest.mark.parametrize("method", ["average", "average_weights", "max"])
def test_zone_blend_vs_oracle(ctx, trained_3_15, method):
    """blended stitching as oracle.zone_detect_ref.run_zone_blend restates the intent of
    compare.py:84-138 (weights test/tiles.py:97-108). Class-map agreement >= 99.9 %, every pixel written;
    a second strip-wise run over two row bands (halo tile rows recomputed) gives the same agreement."""
    from oracle.zone_detect_ref import run_zone_blend
    W, H, T, margin = 1000, 700, 512, 128
    model, georaster, config, tiles = _zone_setup(ctx, trained_3_15, W, H, T, margin, seed=1700)
    acc, wsum = ctx.blend_buffers(method, H, W)
    ctx.blend_strip(tiles, T, 5, method, acc, wsum, W, 0)
    cls = torch.full((H, W), 255, dtype=torch.uint8, device="cuda")
    conf = torch.full((H, W), 255, dtype=torch.uint8, device="cuda")
    ctx.blend_finalize(method, acc, wsum, cls, conf)
    ref_cls, ref_conf = run_zone_blend(model, georaster, config, method)
    cls_h = cls.cpu().numpy()
    agree = (cls_h == ref_cls).mean()
    conf_ref_u8 = np.clip(np.floor(ref_conf + 0.5), 0, 255).astype(np.uint8)
    print(f"blend {method} {W}x{H}: agreement {agree * 100:.4f}%, conf agreement {(conf.cpu().numpy() == conf_ref_u8).mean() * 100:.4f}%")
    assert (cls_h < 15).all() and agree >= AGREE_MIN
    assert (conf.cpu().numpy() == conf_ref_u8).mean() >= 0.995
    if method != "max":
        ws = wsum.cpu().numpy()
        assert ws.min() > 0                                   # every raster pixel is covered by a tile
        if method == "average":
            assert np.array_equal(ws, np.round(ws)) and ws.max() >= 4   # overlap counts (clamped tiles add to 4)
    # two row bands, each fed every tile that touches it: same map
    cls2 = torch.full((H, W), 255, dtype=torch.uint8, device="cuda")
    for r0, r1 in ((0, 300), (300, H)):
        need = tiles[(tiles[:, 1] < r1) & (tiles[:, 1] + T > r0)]
        a2, w2 = ctx.blend_buffers(method, r1 - r0, W)
        ctx.blend_strip(need, T, 4, method, a2, w2, W, r0)
        ctx.blend_finalize(method, a2, w2, cls2[r0:r1], None)
    assert (cls2.cpu().numpy() == ref_cls).mean() >= AGREE_MIN
    assert (cls2 == cls).float().mean().item() >= 0.9995    # float atomics: summation order differs


def test_zone_confusion_matches_oracle_given_same_predictions(ctx, trained_3_15):
    """a9/a14: metrics from the GPU histogram == metrics from sklearn on the same class map."""
    from oracle import synth
    from oracle.metrics_ref import class_IoU, overall_accuracy, patch_confusion
    g = torch.Generator().manual_seed(9)
    pred = torch.randint(0, 15, (512, 768), generator=g, dtype=torch.uint8)
    truth = torch.randint(0, 20, (512, 768), generator=g, dtype=torch.uint8)
    cm = ctx.confusion(pred.cuda(), truth.cuda(), 15, truth_sub=1).cpu().numpy()
    ref = patch_confusion(truth.numpy() - 1, pred.numpy(), 15)
    np.testing.assert_array_equal(cm, ref)
    assert class_IoU(cm)[1] == class_IoU(ref)[1] and overall_accuracy(cm) == overall_accuracy(ref)


def test_full_size_zone_properties(ctx, trained_3_15):
    """BASELINE configs[1] at full size (10000 x 10000, 1600 tiles of 512 / margin 128), through properties that do not
    need the oracle: every pixel written with a valid class, the run is idempotent, sharding the tile rows the way the
    multi-GPU path does (4 strips, each with its own raster rows and map band) gives the same bytes as the whole table,
    the host entry point (row-sorted, pipelined copies) gives the same bytes, and the confusion matrix of the whole
    map is the sum over the strips and counts every pixel whose truth is in range."""
    import bench
    from oracle import synth
    from flair1_b200.zone_detect.slicing_job import split_rows_across_ranks, tile_table
    sd, _ = trained_3_15
    W = H = 10000
    T, M = 512, 128
    dev = torch.device("cuda", 0)
    raster = bench.synth_rows_gpu(W, H, 0, H, 1, dev)
    ctx.load_weights(sd, 3, 15)
    ctx.set_norm("custom", synth.FLAIR_MEANS[:3], synth.FLAIR_STDS[:3])
    ctx.set_raster(raster, [0, 1, 2], W, H)
    tiles = tile_table(W, H, T, M)
    assert len(tiles) == 1600
    cls = torch.full((H, W), 255, dtype=torch.uint8, device=dev)
    conf = torch.full((H, W), 255, dtype=torch.uint8, device=dev)
    ctx.detect_strip(tiles, T, 148, cls, conf, W, 0)
    assert int(cls.max()) < 15 and int(conf.max()) <= 1                     # every pixel written
    cls2 = torch.full_like(cls, 255)
    ctx.detect_strip(tiles, T, 74, cls2, None, W, 0)                          # another batch size, no confidence band
    assert torch.equal(cls, cls2)
    truth = torch.randint(0, 21, (H, W), dtype=torch.uint8, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
    cm = ctx.confusion(cls, truth, 15, truth_sub=1)
    assert int(cm.sum()) == int(((truth >= 1) & (truth <= 15)).sum())
    # the multi-GPU sharding, emulated on one device: 4 strips, own raster rows, own map band
    cm_sum = torch.zeros_like(cm)
    for shard in split_rows_across_ranks(tiles, 4):
        mine = tiles[shard]
        ry0, ry1 = max(int(mine[:, 1].min()), 0), min(int(mine[:, 1].max()) + T, H)
        my0, my1 = int(mine[:, 3].min()), int(mine[:, 5].max())
        ctx.set_raster(raster[:, ry0:ry1].contiguous(), [0, 1, 2], W, H, row0=ry0)
        band = torch.full((my1 - my0, W), 255, dtype=torch.uint8, device=dev)
        ctx.detect_strip(mine, T, 148, band, None, W, my0)
        assert torch.equal(band, cls[my0:my1])
        cm_sum += ctx.confusion(band, truth[my0:my1].contiguous(), 15, truth_sub=1)
    assert torch.equal(cm_sum, cm)
    # host in / host out
    host_r = raster.cpu().numpy()
    out_cls, out_conf = np.zeros((H, W), np.uint8), np.zeros((H, W), np.uint8)
    ctx.detect_zone_host(host_r, [0, 1, 2], W, H, 0, 0, tiles, T, 148, out_cls, out_conf, W, 0, H)
    assert np.array_equal(out_cls, cls.cpu().numpy()) and np.array_equal(out_conf, conf.cpu().numpy())


# ------------------------------------------------------------------------------------------ round 2: shards, collectives, goldens
def test_fused_sink_is_bit_exact_for_19_classes(monkeypatch):
    """The 32-column head (17..32 classes) takes its soft-max in two 16-column halves; the standalone K6 kernel uses
    the same order and arithmetic, so class map AND confidence band are byte-identical with and without the fused
    sink / dead-output elimination, also where the two best classes sit in different halves."""
    from oracle import synth
    from flair1_b200.zone_detect.slicing_job import tile_table
    nat = _nat()
    sd = synth.cached_checkpoint(3, 19)
    W, H, T, margin = 700, 560, 256, 32
    raster = torch.from_numpy(synth.synth_raster(3, H, W, seed=19)).cuda()
    tiles = tile_table(W, H, T, margin)
    res = {}
    for mode in ("plain", "fast", "aligned"):   # aligned: active kernel tiles on the fixed tile grid (FB_ALIGNED_TILES=1)
        monkeypatch.setenv("FB_FULL_TILES", "1" if mode == "plain" else "0")
        monkeypatch.setenv("FB_NO_FUSED_SINK", "1" if mode == "plain" else "0")
        monkeypatch.setenv("FB_ALIGNED_TILES", "1" if mode == "aligned" else "0")
        c = nat.Context(0)
        c.load_weights(sd, 3, 19)
        c.set_norm("custom", synth.FLAIR_MEANS[:3], synth.FLAIR_STDS[:3])
        c.set_raster(raster, [0, 1, 2], W, H)
        cls = torch.full((H, W), 255, dtype=torch.uint8, device="cuda")
        conf = torch.full((H, W), 255, dtype=torch.uint8, device="cuda")
        c.detect_strip(tiles, T, 7, cls, conf, W, 0)
        res[mode] = (cls.cpu().numpy(), conf.cpu().numpy())
        c.close()
    np.testing.assert_array_equal(res["fast"][0], res["plain"][0])
    np.testing.assert_array_equal(res["fast"][1], res["plain"][1])
    assert (res["fast"][0] < 19).all() and (res["fast"][0] >= 16).any() and (res["fast"][0] < 16).any()


def test_confusion_rect_bit_exact(ctx):
    """fb_confusion_rect: the histogram over a rectangle cut out of two wider maps (ragged width, unaligned origin,
    different pitches) equals numpy's on the same pixels, incl. the uint8 wrap of truth - 1 and dropped labels."""
    from oracle.metrics_ref import confusion_numpy
    rng = np.random.default_rng(3)
    for (Hm, Wp, Wt, y0, y1, x0, x1, ncls, sub) in [(300, 1000, 1000, 7, 291, 13, 977, 15, 1), (64, 4099, 5003, 0, 64, 1, 4098, 19, 1),
                                                    (50, 700, 700, 10, 11, 3, 5, 7, 0), (33, 2048, 2048, 0, 33, 0, 2048, 32, 0),
                                                    (40, 600, 640, 5, 35, 16, 528, 15, 1)]:
        pred = rng.integers(0, ncls + 2, (Hm, Wp), dtype=np.uint8)
        truth = rng.integers(0, ncls + 3, (Hm, Wt), dtype=np.uint8)
        pd, td = torch.from_numpy(pred).cuda(), torch.from_numpy(truth).cuda()
        cm = ctx.confusion_rect(pd[y0:y1, x0:x1], td[y0:y1, x0:x1], ncls, truth_sub=sub).cpu().numpy()
        ref = confusion_numpy(truth[y0:y1, x0:x1], pred[y0:y1, x0:x1], ncls, sub)
        np.testing.assert_array_equal(cm, ref)
    empty = torch.zeros((0, 16), dtype=torch.uint8, device="cuda")
    assert int(ctx.confusion_rect(empty, empty, 15).sum()) == 0


def test_zone_shards_fill_one_host_map_and_sum_to_the_whole_confusion_matrix(ctx, trained_3_15):
    """fb_detect_zone_shard: three shards (contiguous ranges of the row-major tile order, so the first / last tile row
    of a shard is shared with its neighbour) write ONLY their own rectangles into one host map, which ends up
    byte-identical to fb_detect_strip over the whole tile table; the fused confusion matrices of the shards add up to
    the matrix of the whole map (bit-exact, sklearn semantics), and cells no shard owns are never touched."""
    from oracle import synth
    from oracle.metrics_ref import confusion_numpy
    from flair1_b200.zone_detect.slicing_job import owned_rects, split_tiles_across_ranks, tile_table
    sd, _ = trained_3_15
    W, H, T, margin = 1500, 1100, 512, 128
    raster = synth.synth_raster(3, H, W, seed=21)
    truth = synth.synth_mask(raster, 15, 3)
    truth[::9, ::4] = 0
    truth[5::13, 1::7] = 18
    ctx.load_weights(sd, 3, 15)
    ctx.set_norm("custom", synth.FLAIR_MEANS[:3], synth.FLAIR_STDS[:3])
    tiles = tile_table(W, H, T, margin)
    ctx.set_raster(torch.from_numpy(raster).cuda(), [0, 1, 2], W, H)
    ref_cls = torch.zeros((H, W), dtype=torch.uint8, device="cuda")
    ref_conf = torch.zeros((H, W), dtype=torch.uint8, device="cuda")
    ctx.detect_strip(tiles, T, 6, ref_cls, ref_conf, W, 0)
    ref_cls, ref_conf = ref_cls.cpu().numpy(), ref_conf.cpu().numpy()
    out = torch.full((2, H, W), 255, dtype=torch.uint8).pin_memory()
    total = torch.zeros((15, 15), dtype=torch.int64, device="cuda")
    for shard in split_tiles_across_ranks(tiles, 3):
        mine = tiles[shard]
        ry0, ry1 = max(int(mine[:, 1].min()), 0), min(int(mine[:, 1].max()) + T, H)
        rects = owned_rects(mine)
        my0, my1 = int(rects[:, 0].min()), int(rects[:, 1].max())
        before = out.clone()
        cm = torch.zeros((15, 15), dtype=torch.int64, device="cuda")
        ctx.detect_zone_shard(torch.from_numpy(raster[:, ry0:ry1].copy()).pin_memory(), [0, 1, 2], W, H, ry0, 0, mine, T, 5,
                              out[0].numpy(), out[1].numpy(), W, 0, H,
                              truth=torch.from_numpy(truth[my0:my1].copy()).pin_memory(), truth_row0=my0, truth_sub=1, cm=cm)
        mask = np.zeros((H, W), bool)
        ref_cm = np.zeros((15, 15), np.int64)
        for y0, y1, x0, x1 in rects:
            mask[y0:y1, x0:x1] = True
            ref_cm += confusion_numpy(truth[y0:y1, x0:x1], ref_cls[y0:y1, x0:x1], 15, 1)
        np.testing.assert_array_equal(cm.cpu().numpy(), ref_cm)
        assert np.array_equal(out.numpy()[:, ~mask], before.numpy()[:, ~mask])      # nothing outside the shard's rectangles
        total += cm
    np.testing.assert_array_equal(out[0].numpy(), ref_cls)
    np.testing.assert_array_equal(out[1].numpy(), ref_conf)
    np.testing.assert_array_equal(total.cpu().numpy(), confusion_numpy(truth, ref_cls, 15, 1))
    # an empty shard (more ranks than tiles) is a no-op
    ctx.detect_zone_shard(torch.from_numpy(raster[:, :T].copy()), [0, 1, 2], W, H, 0, 0, tiles[:0], T, 5, out[0].numpy(), None, W, 0, H)
    np.testing.assert_array_equal(out[0].numpy(), ref_cls)


def test_comm_single_rank_allreduce_and_gather(ctx):
    """The library's own NCCL communicator (fb_comm_*, resolved with dlopen): with one rank the all-reduce and the
    byte gather are identities; a second fb_comm_init on the same context is refused."""
    ctx.comm_init(0, 1, lambda ident: ident)
    cm = torch.arange(15 * 15, dtype=torch.int64, device="cuda").reshape(15, 15).contiguous()
    ref = cm.clone()
    ctx.allreduce_confusion(cm)
    torch.cuda.synchronize()
    assert torch.equal(cm, ref)
    send = torch.arange(1000, dtype=torch.int32, device="cuda").to(torch.uint8)
    got = ctx.gather_bytes(send, [send.numel()], root=0)
    torch.cuda.synchronize()
    assert torch.equal(got, send)
    with pytest.raises(nat_error()):
        ctx.comm_init(0, 1, lambda ident: ident)
    ctx.comm_destroy()
    with pytest.raises(nat_error()):
        ctx.allreduce_confusion(cm)


def test_metadata_forward_matches_the_references_own_forward(ctx):
    """tests/golden/metadata_forward.npz holds the OUTPUT of the reference's FLAIR_ModelFactory.forward / MetadataMLP
    (src/flair/model.py:52-96, run by tests/golden/make_golden.py) on seeded weights and a seeded 5-band tile: the
    library's logits (fb_forward_tiles with metadata) stay within the bf16 tolerance of it."""
    from oracle import synth
    g = np.load(GOLDEN_DIR / "metadata_forward.npz")
    sd = synth.random_checkpoint(5, 13, seed=int(g["weight_seed"]), use_metadata=True)
    checksum = float(sum(v.double().abs().sum() for v in sd.values() if v.is_floating_point()))
    if abs(checksum - float(g["weights_checksum"])) > 1e-6 * float(g["weights_checksum"]):
        pytest.skip("torch draws other numbers from the seed than when the golden was made")
    sd = {k.replace("seg_model.", "", 1) if k.startswith("seg_model.") else k: v for k, v in sd.items()}
    img = synth.synth_raster(5, 512, 512, seed=int(g["img_seed"]))
    ctx.load_weights(sd, 5, 13, use_metadata=True)
    ctx.set_norm("custom", synth.FLAIR_MEANS, synth.FLAIR_STDS)
    ctx.set_raster(torch.from_numpy(img).cuda(), [0, 1, 2, 3, 4], 512, 512)
    got = ctx.forward_tiles(np.array([[0, 0]], np.int32), 512, metadata=g["met"]).cpu().permute(0, 3, 1, 2)[:, :13, ::16, ::16].numpy()
    rel = np.abs(got - g["logits_sub"]).max() / float(g["logits_absmax"])
    print(f"metadata forward vs the reference's own forward(): rel err {rel:.4e}")
    assert rel <= LOGIT_TOL


_FALLBACK_SNIPPET = r"""
import hashlib, sys
sys.path.insert(0, {root!r})
import torch
import flair1_b200._native as nat
from flair1_b200.zone_detect.slicing_job import tile_table
from oracle import synth
W = H = 896
T, margin = 128, 32
raster = torch.from_numpy(synth.synth_raster(3, H, W, seed=5)).cuda()
tiles = tile_table(W, H, T, margin)
c = nat.Context(0)
c.load_weights(synth.cached_checkpoint(3, 15), 3, 15)
c.set_norm("custom", synth.FLAIR_MEANS[:3], synth.FLAIR_STDS[:3])
c.set_raster(raster, [0, 1, 2], W, H)
cls = torch.full((H, W), 255, dtype=torch.uint8, device="cuda")
conf = torch.full((H, W), 255, dtype=torch.uint8, device="cuda")
c.detect_strip(tiles, T, 148, cls, conf, W, 0)
torch.cuda.synchronize()
print("tiles", len(tiles), "digest", hashlib.sha256(cls.cpu().numpy().tobytes() + conf.cpu().numpy().tobytes()).hexdigest())
"""


def test_cp_async_halo_fallback_writes_the_same_maps():
    """FB_TMAH=0 (read once per process, hence the two subprocesses) switches every halo layer back from TMA-staged planes
    to cp.async gathers; the fused stem, which exists only in the TMA-staged form, then runs as stem + separate max-pool
    kernel. A zone of 169 tiles in batches of 148 (large enough for the fused stem in the default run) must give the
    same class and confidence bytes either way."""
    import os
    import subprocess
    import sys
    code = _FALLBACK_SNIPPET.format(root=str(Path(__file__).resolve().parent.parent))
    out = {}
    for mode in ("default", "cp.async"):
        env = dict(os.environ)
        env.pop("FB_TMAH", None)
        if mode == "cp.async":
            env["FB_TMAH"] = "0"
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        line = [ln for ln in r.stdout.splitlines() if ln.startswith("tiles")][-1]
        assert int(line.split()[1]) >= 148
        out[mode] = line.split()[-1]
    assert out["default"] == out["cp.async"]
