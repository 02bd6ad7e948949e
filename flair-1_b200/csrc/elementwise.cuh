// Launchers for the HBM-bound kernels (see elementwise.cu). All return cudaError_t as int (0 = OK)
// or a negative library code.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fb {

int launch_extract_normalise(const uint8_t* raster, int layout_hwc, int bands_total, const int* band_idx,
                             int c, long long W, long long H, long long row0, long long rows,
                             const int* tile_xy, int n, int T, const __nv_bfloat16* lut,
                             __nv_bfloat16* out, int num_sms, cudaStream_t stream);

int launch_maxpool3x3s2(const __nv_bfloat16* in, __nv_bfloat16* out, int B, int H, int W, int C,
                        int num_sms, cudaStream_t stream);

// wb = {w0[64x45], b0[64], w1[32x64], b1[32], w2[16x32], b2[16]} (device fp32)
int launch_metadata_mlp(const float* met, const float* const* wb, float* out, int n, cudaStream_t stream);

// tiles: int32 [n][6] = x0, y0 (tile origin in raster px), wx0, wy0, wx1, wy1 (half-open write rect)
int launch_argmax_stitch(const float* logits, int ncls, int n, int T, const int* tiles, uint8_t* cls_map,
                         uint8_t* conf_map, long long map_w, long long map_row0, cudaStream_t stream);

int launch_confusion(const uint8_t* pred, const uint8_t* truth, long long npx, int ncls, int truth_sub,
                     long long* cm, int num_sms, cudaStream_t stream);

}  // namespace fb
