/* libflairb200 -- C ABI of the B200-native zone_detect / patch-predict hot path.
 *
 * The reference (Draghoyns/FLAIR-1) is pure Python and has no FFI of its own; its "operator boundary"
 * for this path is a handful of Python calls. Each entry point below replaces
 * one of those calls and cites it. A reference maintainer binds these with ctypes.
 *
 * Conventions
 *  - every function returns 0 on success, a positive cudaError_t, or a negative library code;
 *    fb_last_error(ctx) gives a human-readable message for the last failure on that context
 *    (fb_last_error(NULL) for fb_create failures). No exception ever crosses this boundary.
 *  - a context is bound to one CUDA device and one stream (a cudaStream_t passed as void*; NULL = the
 *    legacy default stream). It is NOT thread-safe; distinct contexts are independent.
 *  - "dev" pointers are device pointers on the context's device, "host" pointers are host memory.
 *  - all work is enqueued on the context's stream; calls that take host output buffers synchronise
 *    the stream before returning, the others do not.
 *  - there is no CPU fallback: every call fails with FB_ERR_NO_DEVICE when no sm_100 device exists.
 */
#ifndef FLAIR_B200_H_
#define FLAIR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FB_API_VERSION 1

#define FB_OK 0
#define FB_ERR_INVALID (-1)    /* bad argument */
#define FB_ERR_STATE (-2)      /* call order (weights / norm / raster not set) */
#define FB_ERR_NO_DEVICE (-3)  /* no CUDA device of compute capability 10.x */
#define FB_ERR_WEIGHTS (-4)    /* missing / mis-shaped tensor in the state dict */
#define FB_ERR_OOM (-5)

#define FB_NORM_CUSTOM 0  /* (x - mean) / std, float64 then float32: src/zone_detect/dataset.py:79-86 */
#define FB_NORM_SCALING 1 /* x / 255 (skimage img_as_float): dataset.py:88, src/flair/data_loader.py:28-29 */
#define FB_NORM_WITHOUT 2 /* raw values: src/flair/data_loader.py:15 ('without') */

#define FB_LAYOUT_CHW 0 /* band-planar, what rasterio's read() returns */
#define FB_LAYOUT_HWC 1 /* pixel-interleaved */

#define FB_LOGIT_STRIDE 16 /* logits are written as [n, T, T, S] fp32 with S = fb_logit_stride(ctx): 16 for models
                             with <= 16 classes, 32 above (the 19-class nomenclature); entries >= n_classes are 0 */
#define FB_METADATA_DIM 45 /* src/flair/tasks_utils.py:158-213 */

typedef struct fb_ctx fb_ctx;

/* One entry of the checkpoint: host fp32 data under its segmentation-models-pytorch key
 * ("encoder.conv1.weight", "decoder.blocks.0.conv1.1.running_var", "segmentation_head.0.bias",
 * "enc.enc_mlp.0.weight", ...) after the "model.seg_model." / "model." prefixes were stripped the way
 * src/zone_detect/model.py:61-76 does. */
typedef struct fb_tensor_desc {
  const char* name;
  const float* data;
  int32_t ndim;
  int64_t shape[4];
} fb_tensor_desc;

/* One sliding-window tile of the slicing job (src/zone_detect/slicing_job.py:54-106) in raster pixel
 * coordinates, y growing downwards. (x0, y0) is the top-left of the margin-expanded square and may
 * lie outside the raster; [wx0,wx1) x [wy0,wy1) is the part of the class map this tile owns
 * (interior box after margin clipping, minus what a later tile of the write order overwrites:
 * src/zone_detect/compare.py:66-82, src/zone_detect/main.py:409-426). */
typedef struct fb_tile {
  int32_t x0, y0;
  int32_t wx0, wy0, wx1, wy1;
} fb_tile;

/* ---- lifetime ------------------------------------------------------------------------------- */
int fb_api_version(void);
int fb_create(int device, void* cuda_stream, fb_ctx** out);
void fb_destroy(fb_ctx* ctx);
const char* fb_last_error(const fb_ctx* ctx);
int fb_synchronize(fb_ctx* ctx);

/* ---- model: replaces smp.create_model + load_state_dict(strict=True)
 *      (src/zone_detect/model.py:30-39,79-88; src/flair/model.py:20-50, src/flair/main.py:77-146).
 *      Folds eval-mode BatchNorm into the conv weights in float64, rounds once to bf16, repacks to
 *      [Cout][kh][kw][Cin] and uploads. in_channels in 1..8, n_classes in 1..32. */
int fb_load_weights(fb_ctx* ctx, const fb_tensor_desc* tensors, int n_tensors, int in_channels,
                    int n_classes, int use_metadata);
/* Floats per pixel of the logits (fb_forward_tiles) and of the blend accumulators (fb_blend_strip) for the loaded
 * model: 16 for <= 16 classes, 32 above. */
int fb_logit_stride(const fb_ctx* ctx);

/* ---- input normalisation: Sliced_Dataset.normalization (dataset.py:68-88) / norm()
 *      (data_loader.py:9-30). mean/std are per selected band (ignored unless FB_NORM_CUSTOM). */
int fb_set_norm(fb_ctx* ctx, int mode, const double* mean, const double* std, int c);

/* ---- raster: replaces rasterio.open + windowed boundless read (dataset.py:90-104).
 *      The buffer holds rows [row0, row0+rows) of a W x H raster with bands_total uint8 bands;
 *      band_idx[c] are the 0-based bands fed to the network (config "channels" minus 1). Anything
 *      outside [0,W)x[0,H) -- or outside the resident rows -- reads as raw 0 before normalisation.
 *      fb_set_raster borrows a device pointer; fb_upload_raster copies from host (pinned or not) into
 *      a context-owned device buffer on the context's stream. */
int fb_set_raster(fb_ctx* ctx, const uint8_t* dev_raster, int bands_total, const int32_t* band_idx,
                  int c, int64_t W, int64_t H, int64_t row0, int64_t rows, int layout);
int fb_upload_raster(fb_ctx* ctx, const uint8_t* host_raster, int bands_total, const int32_t* band_idx,
                     int c, int64_t W, int64_t H, int64_t row0, int64_t rows, int layout);

/* ---- logits = model(imgs [, met]) (compare.py:27-33; task_module.py:206-210; flair/model.py:52-70)
 *      for n tiles of tile x tile pixels cut from the current raster at host tile_xy[n][2] = (x0, y0).
 *      metadata: host [n][45] or NULL. logits_dev: device [n][tile][tile][fb_logit_stride] fp32. */
int fb_forward_tiles(fb_ctx* ctx, const int32_t* tile_xy, int n, int tile, const float* metadata,
                     float* logits_dev);

/* ---- the zone_detect hot loop (main.py:398-426): for every tile forward -> softmax -> margin crop ->
 *      argmax / max-probability -> write into the class map, `batch` tiles per forward pass.
 *      cls_map_dev / conf_map_dev (conf may be NULL): device uint8 [map_rows][map_w]; a raster pixel
 *      (x, y) lands at (y - map_row0) * map_w + x. */
int fb_detect_strip(fb_ctx* ctx, const fb_tile* tiles, int n, int tile, int batch, uint8_t* cls_map_dev,
                    uint8_t* conf_map_dev, int64_t map_w, int64_t map_row0);

/* ---- the same loop with the per-patch metrics of the compare loop (main.py:349-366 -> test/metrics.py:124-163,
 *      compute_metrics_patch): besides writing the maps, tile i's OWN arg-max prediction over windows[i]
 *      (x0, y0 = the tile's origin, [wx0,wx1) x [wy0,wy1) = its margin-cropped window, which for a clamped
 *      last row / column overlaps what other tiles own) is counted against the truth raster:
 *      cm_tiles_dev[i][t][p] += #{ px in window : (uint8)(truth - truth_sub) == t, pred == p, t < n_classes }.
 *      truth_dev: uint8 with the class map's geometry; cm_tiles_dev: int64 [n][n_classes][n_classes], zeroed
 *      by the caller. Whole tiles are computed (no dead-output elimination). */
int fb_detect_strip_metrics(fb_ctx* ctx, const fb_tile* tiles, const fb_tile* windows, int n, int tile, int batch,
                            uint8_t* cls_map_dev, uint8_t* conf_map_dev, int64_t map_w, int64_t map_row0,
                            const uint8_t* truth_dev, int truth_sub, int64_t* cm_tiles_dev);

/* ---- the same loop with output_type "class_prob" (dataset.py:15-21 `convert`, main.py:229, 421-426):
 *      every class probability of the margin-cropped tile as uint8(p * 255) (truncation), written to
 *      prob_map_dev uint8 [n_classes][map_rows][map_w] (band k+1 of the reference's output = plane k). */
int fb_detect_strip_prob(fb_ctx* ctx, const fb_tile* tiles, int n, int tile, int batch, uint8_t* prob_map_dev,
                         int64_t map_w, int64_t map_row0, int64_t map_rows);

/* ---- blended stitching: what the weighted branches of `stitching` (compare.py:84-138) intend, with the
 *      weights of test/tiles.py:97-108 and the normalisations of tiles.py:54-94 / 111-169. The whole
 *      tile contributes, clipped to the raster (and to the map rows). method: FB_STITCH_AVERAGE (mean
 *      probability of the covering tiles), FB_STITCH_AVERAGE_WEIGHTS (weight exp(-0.5 * Chebyshev
 *      distance to the tile centre / (tile/2))), FB_STITCH_MAX (class of the most confident tile, the
 *      later tile of the write order on ties; tile i of this call has sequence number tile_seq0 + i).
 *      acc_dev: float [map_rows][map_w][fb_logit_stride] and wsum_dev: float [map_rows][map_w] for the two averages;
 *      for FB_STITCH_MAX acc_dev is used as uint64 [map_rows][map_w] and wsum_dev may be NULL. The caller
 *      zeroes the accumulators before the first call; the write rectangles of `tiles` are ignored.
 *      fb_blend_finalize turns the accumulators of npx pixels into the class map (+ confidence band). */
#define FB_STITCH_AVERAGE 0
#define FB_STITCH_AVERAGE_WEIGHTS 1
#define FB_STITCH_MAX 2
int fb_blend_strip(fb_ctx* ctx, const fb_tile* tiles, int n, int tile, int batch, int method, float* acc_dev,
                   float* wsum_dev, int64_t map_w, int64_t map_row0, int64_t map_rows, int tile_seq0);
int fb_blend_finalize(fb_ctx* ctx, const float* acc_dev, const float* wsum_dev, int method, int64_t npx,
                      uint8_t* cls_map_dev, uint8_t* conf_map_dev);

/* ---- same loop, host buffers in and out (upload raster rows, detect, download class map); this is
 *      the call the end-to-end benchmark times. host_cls / host_conf: [map_rows][map_w] uint8. */
int fb_detect_zone_host(fb_ctx* ctx, const uint8_t* host_raster, int bands_total, const int32_t* band_idx,
                        int c, int64_t W, int64_t H, int64_t row0, int64_t rows, int layout,
                        const fb_tile* tiles, int n, int tile, int batch, uint8_t* host_cls,
                        uint8_t* host_conf, int64_t map_w, int64_t map_row0, int64_t map_rows);

/* ---- the same loop for ONE SHARD of a zone whose class map is assembled by several ranks (one process per GPU;
 *      SURVEY.md section 8e). The reference writes every tile's window into the one output raster as it goes
 *      (main.py:421-426); here host_cls / host_conf describe that one map -- rows [map_row0, map_row0 + map_rows),
 *      pitch map_w, typically a shared mapping that every rank has pinned -- and this call writes ONLY the write
 *      rectangles of its own `tiles` into it (2-D copies, row band by row band as they become final), so the
 *      ranks' pieces never pass through another rank's GPU or PCIe link. The device-side maps cover just the rows
 *      the shard writes. With host_truth (rows [truth_row0, ...), pitch map_w, NULL to skip) the truth pixels of
 *      the same rectangles are uploaded behind the raster and the shard's confusion matrix
 *      (test/metrics.py:161-163, 229-231: rows = (uint8)(truth - truth_sub), columns = prediction) is added to
 *      cm_dev (device int64 [ncls_cm][ncls_cm], zeroed by the caller, summed over ranks with
 *      fb_allreduce_confusion). Synchronises all three streams before returning. */
int fb_detect_zone_shard(fb_ctx* ctx, const uint8_t* host_raster, int bands_total, const int32_t* band_idx,
                         int c, int64_t W, int64_t H, int64_t row0, int64_t rows, int layout,
                         const fb_tile* tiles, int n, int tile, int batch, uint8_t* host_cls,
                         uint8_t* host_conf, int64_t map_w, int64_t map_row0, int64_t map_rows,
                         const uint8_t* host_truth, int64_t truth_row0, int truth_sub, int ncls_cm, int64_t* cm_dev);

/* ---- patch predict (flair/task_module.py:206-213 + data_loader.py:130-144): n whole patches,
 *      dev_patches uint8 [n][c][tile][tile] (band-planar per patch, already restricted to the selected
 *      bands), metadata host [n][45] or NULL, cls_out_dev uint8 [n][tile][tile] (0-based classes). */
int fb_predict_patches(fb_ctx* ctx, const uint8_t* dev_patches, const float* metadata, int n, int tile,
                       int batch, uint8_t* cls_out_dev);

/* ---- confusion matrix (sklearn.metrics.confusion_matrix(labels=range(ncls)) at
 *      flair/metrics.py:67-71 and zone_detect/test/metrics.py:161-163,229-231):
 *      cm_dev[t*ncls + p] += #{ i : (uint8)(truth[i] - truth_sub) == t, pred[i] == p, t,p < ncls }.
 *      int64 device accumulator, caller zeroes it; ncls <= 32. */
int fb_confusion(fb_ctx* ctx, const uint8_t* pred_dev, const uint8_t* truth_dev, int64_t npx, int ncls,
                 int truth_sub, int64_t* cm_dev);

/* The same histogram over a rows x width rectangle of two pitched uint8 maps (pitch = bytes between row starts):
 * what one rank scores when it owns a set of write rectangles rather than whole rows of the map. */
int fb_confusion_rect(fb_ctx* ctx, const uint8_t* pred_dev, const uint8_t* truth_dev, int64_t rows, int64_t width,
                      int64_t pred_pitch, int64_t truth_pitch, int ncls, int truth_sub, int64_t* cm_dev);

/* ---- multi-GPU (one process / context per GPU). The reference is single-GPU on this path (its Lightning DDP code
 *      only serves training, src/flair/tasks.py:83-142); sharding a zone needs exactly two collectives, both over
 *      NCCL (resolved with dlopen at the first call: libnccl.so.2 of the process, else of the system):
 *      fb_comm_unique_id: rank 0 creates the 128-byte rendezvous id and hands it to the other ranks by any means
 *      (torch.distributed broadcast, a file, MPI); fb_comm_init: every rank joins (collective, blocking);
 *      fb_allreduce_confusion: in-place sum over the ranks of cm_dev[ncls*ncls] (int64), on the context's stream
 *      (replaces nothing in the reference -- np.sum over patches at flair/metrics.py:76 is the single-process
 *      analogue); fb_gather_bytes: rank r's send_bytes bytes land at recv_dev + offsets[r] on `root`
 *      (offsets / counts: host arrays of world entries, the same on every rank) -- the class_prob planes and the
 *      blended maps travel to the writer rank this way. */
/* Page-lock / release host memory owned by the caller (cudaHostRegister, portable): the shared class map of a
 * sharded zone, so that fb_detect_zone_shard's copies into it run asynchronously. host_ptr and bytes are rounded
 * by the caller to the page size. Failures are reported through fb_last_error(NULL). */
int fb_host_register(void* host_ptr, int64_t bytes);
int fb_host_unregister(void* host_ptr);

#define FB_COMM_ID_BYTES 128
#define FB_ERR_NCCL_MISSING (-6) /* libnccl.so.2 could not be loaded */
#define FB_ERR_NCCL (-7)         /* an NCCL call failed; see fb_last_error */
int fb_comm_unique_id(uint8_t* id128);
int fb_comm_init(fb_ctx* ctx, const uint8_t* id128, int rank, int world);
int fb_comm_destroy(fb_ctx* ctx);
int fb_allreduce_confusion(fb_ctx* ctx, int64_t* cm_dev, int ncls);
int fb_gather_bytes(fb_ctx* ctx, const void* send_dev, int64_t send_bytes, void* recv_dev, const int64_t* offsets,
                    const int64_t* counts, int root);

/* ---- test / profiling hooks (used by tests/ and bench.py, not by the pipelines) --------------- */
/* Single convolution: NHWC bf16 in/out, weights [Cout][Kpad] bf16 (k = (kh*KW+kw)*Cin + c, Kpad a
 * multiple of 64), fp32 bias. x2/C2: optional second (skip) source concatenated after x1's channels;
 * up1: x1 is read through a nearest x2 upsample. mode: 0 = cp.async gather producer, 1 = TMA producer
 * (3x3 stride 1 only), -1 = automatic. Exactly one of out_bf16 / out_f32 is non-NULL. */
int fb_conv2d(fb_ctx* ctx, const void* x1, const void* x2, int C1, int C2, int up1, int B, int Hin,
              int Win, int KH, int KW, int stride, int pad, int Cout, const void* weights, int Kpad,
              const float* bias, const void* residual, const float* rowbias, int relu, void* out_bf16,
              float* out_f32, int mode);
/* Same convolution through the halo-staged kernel (3x3 stride 1 with <= 64 channels per source and
 * Cout in {16,32,64}, or the 7x7 stride-2 stem on 8 channels): weights are given as host fp32
 * [Cout][C1+C2][KH][KH] and packed internally. up2_out: write the bf16 result 2x2-replicated into
 * [B, 2*Hout, 2*Wout, Cout] (the decoder's nearest x2 upsample). up2_out == 2 selects the sub-pixel
 * phase form instead (32 -> 16 channels only): the result is the 3x3 conv of the x2-upsampled x1,
 * [B, 2*Hin, 2*Win, Cout], computed from the low-res x1 as four 2x2-tap phases. Synchronises the stream. */
int fb_conv2d_halo(fb_ctx* ctx, const void* x1, const void* x2, int C1, int C2, int B, int Hin, int Win,
                   int KH, int stride, int Cout, const float* w_oihw_host, const float* bias,
                   const void* residual, int relu, int up2_out, void* out_bf16, float* out_f32);
/* Copy a named intermediate of the last fb_forward_tiles call (e.g. "f1", "layer2.0.out", "dec3")
 * to a device buffer as bf16 NHWC. Returns its element count through *count (also when out is NULL). */
int fb_debug_activation(fb_ctx* ctx, const char* name, void* out_dev, int64_t* count, int32_t* dims4);
/* Run fb_forward_tiles-equivalent work `iters` times and report the mean device time per kernel
 * family in ms (CUDA events on the context's stream): [0]=extract [1]=conv(all) [2]=maxpool+mlp
 * [3]=argmax/stitch [4]=total. */
int fb_profile_forward(fb_ctx* ctx, int n, int tile, int iters, float* ms5);
/* Device-time accounting over any region of calls on this context: between fb_profile_begin and
 * fb_profile_end every kernel family is bracketed by CUDA events on the context's stream (two events
 * per family per batch, not per launch). ms4: [0]=extract [1]=conv kernels [2]=maxpool+mlp
 * [3]=argmax/stitch, summed over the region. fb_profile_end synchronises the stream. */
int fb_profile_begin(fb_ctx* ctx);
int fb_profile_end(fb_ctx* ctx, float* ms4);
/* Number of kernels this library has launched on the context since creation. */
int64_t fb_launch_count(const fb_ctx* ctx);
/* Algorithmic FLOPs (2 * MAC of the direct convolution, unpadded channels) of the conv outputs this context
 * has computed since creation. A full 512^2 tile, 3 bands / 15 classes, is 63.569 GFLOP;
 * the exact-clipping zone loop computes less per tile because the decoder skips outputs that can only reach
 * the cropped margin (compare.py:66-82; FB_FULL_TILES=1 in the environment turns that off). */
double fb_flop_count(const fb_ctx* ctx);

/* Host-only (no device needed): the output region of decoder layer `layer` (2*d = dec<d>.conv1, 2*d + 1 =
 * dec<d>.conv2, 10 = segmentation head) that a tile x tile input needs when only [ax0,ax1) x [ay0,ay1) of its
 * logits is used (csrc/tile_need.cuh, dead-output elimination). rect4 = x0, y0, x1, y1 in that layer's output grid. */
int fb_debug_need_rect(int tile, int layer, int ax0, int ay0, int ax1, int ay1, int32_t* rect4);
/* Origin-shifted kernel-tile cover of that region (csrc/tile_need.cuh, need_span): kernel tiles of th x tw on the tile
 * grid (= the layer's output grid / scale, scale 1 or 2). cover4 = origin column, origin row, tile columns, tile rows,
 * in tile-grid units; the cover contains the region, stays inside the grid and uses the fewest tiles per axis. */
int fb_debug_tile_cover(int tile, int layer, int scale, int th, int tw, int ax0, int ay0, int ax1, int ay1, int32_t* cover4);

/* ---- host-side TIFF LZW codec (compression 5, libtiff/GDAL-compatible) used by the GeoTIFF
 *      reader/writer that stands in for rasterio (main.py:218-232, 421-426; writer.py:38-50).
 *      encode/decode return the number of bytes produced, or -1 (buffer too small / corrupt). */
int64_t fb_lzw_bound(int64_t n);
int64_t fb_lzw_encode(const uint8_t* src, int64_t n, uint8_t* dst, int64_t cap);
int64_t fb_lzw_decode(const uint8_t* src, int64_t n, uint8_t* dst, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* FLAIR_B200_H_ */
