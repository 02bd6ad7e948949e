// Launchers for the HBM-bound kernels (see elementwise.cu). All return cudaError_t as int (0 = OK)
// or a negative library code.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fb {

// tile_stride != 0 (patch predict): tile t is cut at tile_xy[0..1] from its own raster at raster + t * tile_stride.
int launch_extract_normalise(const uint8_t* raster, int layout_hwc, int bands_total, const int* band_idx,
                             int c, long long W, long long H, long long row0, long long rows,
                             const int* tile_xy, int n, int T, const __nv_bfloat16* lut,
                             __nv_bfloat16* out, int num_sms, cudaStream_t stream, long long tile_stride = 0);

// <= 4 bands: the tile in 2x2 space-to-depth form, out = [n][T/2][T/2][16] bf16, channel (py*2 + px)*c + band.
int launch_extract_normalise_s2d(const uint8_t* raster, int layout_hwc, int bands_total, const int* band_idx,
                                 int c, long long W, long long H, long long row0, long long rows,
                                 const int* tile_xy, int n, int T, const __nv_bfloat16* lut,
                                 __nv_bfloat16* out, int num_sms, cudaStream_t stream, long long tile_stride = 0);

int launch_maxpool3x3s2(const __nv_bfloat16* in, __nv_bfloat16* out, int B, int H, int W, int C,
                        int num_sms, cudaStream_t stream);

// wb = {w0[64x45], b0[64], w1[32x64], b1[32], w2[16x32], b2[16]} (device fp32)
int launch_metadata_mlp(const float* met, const float* const* wb, float* out, int n, cudaStream_t stream);

// ls = floats per pixel of the logits / blend accumulators: 16 for <= 16 classes, 32 above.
// tiles: int32 [n][6] = x0, y0 (tile origin in raster px), wx0, wy0, wx1, wy1 (half-open write rect)
int launch_argmax_stitch(const float* logits, int ncls, int ls, int n, int T, const int* tiles, uint8_t* cls_map,
                         uint8_t* conf_map, long long map_w, long long map_row0, cudaStream_t stream);

// class_prob output: prob_map[k][y - map_row0][x] = uint8(softmax_k * 255) (truncation) for k < ncls inside
// the write rectangles; plane stride = map_rows * map_w.
int launch_prob_stitch(const float* logits, int ncls, int ls, int n, int T, const int* tiles, uint8_t* prob_map,
                       long long map_w, long long map_row0, long long map_rows, cudaStream_t stream);

// Blended stitching. method 0 = average (weight 1), 1 = average_weights (exp(-0.5 * chebyshev distance to the
// tile centre / (T/2))), 2 = max (highest soft-max confidence wins, later tile on ties).
// methods 0/1: acc = float [map_rows][map_w][16] (sum of weight * probability) and wsum = float
// [map_rows][map_w]; method 2: acc is read as uint64 [map_rows][map_w] keys
// (confidence bits << 32 | tile sequence number << 8 | class) and wsum is unused.
// Every pixel of every tile that lies inside the raster [0,W)x[0,H) and inside the map rows contributes.
int launch_blend_accumulate(const float* logits, int ncls, int ls, int n, int T, const int* tiles, int method, float* acc,
                            float* wsum, long long map_w, long long map_row0, long long map_rows, long long W,
                            long long H, int seq0, cudaStream_t stream);
int launch_blend_finalize(const float* acc, const float* wsum, int method, int ncls, int ls, long long npx, uint8_t* cls_map,
                          uint8_t* conf_map, int num_sms, cudaStream_t stream);

int launch_confusion(const uint8_t* pred, const uint8_t* truth, long long npx, int ncls, int truth_sub,
                     long long* cm, int num_sms, cudaStream_t stream);

// The same histogram over rows x width pixels of two pitched maps (bytes between row starts: pred_pitch / truth_pitch).
int launch_confusion_rect(const uint8_t* pred, const uint8_t* truth, long long rows, long long width,
                          long long pred_pitch, long long truth_pitch, int ncls, int truth_sub, long long* cm,
                          int num_sms, cudaStream_t stream);

// Per-tile confusion matrices from each tile's own arg-max over its metric window (compute_metrics_patch of the
// compare loop). windows: int32 [n][6] = x0, y0, then the half-open window; cm: int64 [n][ncls][ncls], accumulated.
int launch_tile_confusion(const float* logits, int ncls, int ls, int n, int T, const int* windows, const uint8_t* truth,
                          int truth_sub, long long map_w, long long map_row0, long long* cm, cudaStream_t stream);

}  // namespace fb
