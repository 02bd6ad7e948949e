// HBM-bound kernels of the zone_detect / patch-predict hot path (everything that is not a conv):
//   K1  tile extraction + per-band normalisation   (src/zone_detect/dataset.py:68-113, src/flair/data_loader.py:9-30)
//   --  3x3 stride-2 max-pool of the stem           (torchvision ResNet `maxpool`)
//   K4  metadata MLP 45->64->32->16                 (src/flair/model.py:74-96)
//   K6  softmax-max / argmax / margin clip / stitch (src/zone_detect/compare.py:35,66-82, dataset.py:11-34)
//   K9  confusion-matrix histogram                  (src/flair/metrics.py:60-74, src/zone_detect/test/metrics.py:146-163)
#include "elementwise.cuh"
#include "conv_epilogue.cuh"
#include "tile_need.cuh"

#include <cuda_bf16.h>

namespace fb {

// ------------------------------------------------------------------------------------------ K1
// One thread per four consecutive output pixels of a tile row (blockIdx.y = tile): gathers their `c` band bytes
// (all loads of the four pixels are issued before the first use, coalesced along x for the planar layout),
// maps each through a 256-entry bf16 table (the host builds it in float64 -> float32 -> bf16 so the
// result is bit-identical to rounding the reference's float32 tensor), writes 4 x 8 bf16 (64 contiguous bytes).
// Pixels outside the raster read raw 0 *before* normalisation (rasterio boundless=True semantics).
__global__ void __launch_bounds__(256)
extract_normalise_kernel(const uint8_t* __restrict__ raster, int layout_hwc, int bands_total,
                         const int* __restrict__ band_idx, int c, long long W, long long H,
                         long long row0, long long rows, const int* __restrict__ tile_xy, int n,
                         int T, const __nv_bfloat16* __restrict__ lut, __nv_bfloat16* __restrict__ out,
                         long long tile_stride) {
  __shared__ __nv_bfloat16 s_lut[8 * 256];
  __shared__ int s_band[8];
  for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) s_lut[i] = lut[i];
  if (threadIdx.x < 8) s_band[threadIdx.x] = threadIdx.x < c ? band_idx[threadIdx.x] : 0;
  __syncthreads();
  const int t = blockIdx.y;
  const int TQ = T >> 2;
  const int quads = T * TQ;
  // tile_stride != 0: every tile is cut from its own little raster (patch predict), all at tile_xy[0..1]
  raster += static_cast<long long>(t) * tile_stride;
  const long long tx0 = tile_xy[tile_stride ? 0 : 2 * t], ty0 = tile_xy[tile_stride ? 1 : 2 * t + 1];
  __nv_bfloat16* tile_out = out + static_cast<long long>(t) * T * T * 8;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += gridDim.x * blockDim.x) {
    const int y = q / TQ, x = (q - y * TQ) << 2;
    const long long ry = ty0 + y, rx = tx0 + x;
    const bool row_ok = ry >= 0 && ry < H && ry >= row0 && ry < row0 + rows;
    const long long ly = ry - row0;
    unsigned raw[8][4];
#pragma unroll
    for (int ch = 0; ch < 8; ++ch) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        raw[ch][j] = 0;
        if (ch < c && row_ok && rx + j >= 0 && rx + j < W)
          raw[ch][j] = layout_hwc ? raster[(ly * W + rx + j) * bands_total + s_band[ch]]
                                  : raster[(static_cast<long long>(s_band[ch]) * rows + ly) * W + rx + j];
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __align__(16) __nv_bfloat16 v[8];
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) v[ch] = ch < c ? s_lut[ch * 256 + raw[ch][j]] : __float2bfloat16(0.f);
      *reinterpret_cast<uint4*>(tile_out + (static_cast<long long>(y) * T + x + j) * 8) = *reinterpret_cast<const uint4*>(v);
    }
  }
}

int launch_extract_normalise(const uint8_t* raster, int layout_hwc, int bands_total, const int* band_idx,
                             int c, long long W, long long H, long long row0, long long rows,
                             const int* tile_xy, int n, int T, const __nv_bfloat16* lut,
                             __nv_bfloat16* out, int num_sms, cudaStream_t stream, long long tile_stride) {
  if (n <= 0 || T <= 0) return 0;
  if (T % 4 != 0) return -2003;
  // ~8 resident blocks per SM over the whole batch, at least one block per tile
  int bx = (num_sms * 8 + n - 1) / n;
  const int need = (T * (T / 4) + 255) / 256;
  if (bx > need) bx = need;
  if (bx < 1) bx = 1;
  dim3 grid(bx, n);
  extract_normalise_kernel<<<grid, 256, 0, stream>>>(
      raster, layout_hwc, bands_total, band_idx, c, W, H, row0, rows, tile_xy, n, T, lut, out, tile_stride);
  return static_cast<int>(cudaGetLastError());
}

// Space-to-depth variant for models with <= 4 bands: the 7x7 stride-2 stem becomes a 4x4 stride-1 convolution on
// the 2x2 space-to-depth image (conv_halo.cuh), so the tile is written as [T/2][T/2][16] bf16 with channel
// (py*2 + px)*c + band = pixel (2Y + py, 2X + px) and zeros above 4*c: 8 bytes per pixel instead of 16, and a
// K=16 MMA step per filter tap with no padded lanes beyond 4*c. Same table, same boundless rule, bit-identical
// values. One thread per two horizontally adjacent space-to-depth pixels (2 rows x 4 pixels of the tile).
template <int c>
__global__ void __launch_bounds__(256)
extract_normalise_s2d_kernel(const uint8_t* __restrict__ raster, int layout_hwc, int bands_total,
                             const int* __restrict__ band_idx, long long W, long long H,
                             long long row0, long long rows, const int* __restrict__ tile_xy, int n,
                             int T, const __nv_bfloat16* __restrict__ lut, __nv_bfloat16* __restrict__ out,
                             long long tile_stride) {
  __shared__ __nv_bfloat16 s_lut[4 * 256];
  __shared__ int s_band[4];
  for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) s_lut[i] = lut[i];
  if (threadIdx.x < 4) s_band[threadIdx.x] = threadIdx.x < c ? band_idx[threadIdx.x] : 0;
  __syncthreads();
  const int t = blockIdx.y;
  const int T2 = T >> 1, TQ = T >> 2;
  const int items = T2 * TQ;
  raster += static_cast<long long>(t) * tile_stride;
  const long long tx0 = tile_xy[tile_stride ? 0 : 2 * t], ty0 = tile_xy[tile_stride ? 1 : 2 * t + 1];
  __nv_bfloat16* tile_out = out + static_cast<long long>(t) * T2 * T2 * 16;
  // rx = tx0 + 4 * (q % TQ): every row address of a band plane is a multiple of four when base, W and tx0 are
  const bool quad_ok = !layout_hwc && (W & 3) == 0 && (tx0 & 3) == 0 && (reinterpret_cast<uintptr_t>(raster) & 3) == 0;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < items; q += gridDim.x * blockDim.x) {
    const int Y = q / TQ, X = (q - Y * TQ) << 1;        // space-to-depth pixels (Y, X) and (Y, X + 1)
    const long long rx = tx0 + 2 * X;
    unsigned raw[4][2][4];                              // [band][row py][4 tile pixels]
#pragma unroll
    for (int py = 0; py < 2; ++py) {
      const long long ry = ty0 + 2 * Y + py;
      const bool row_ok = ry >= 0 && ry < H && ry >= row0 && ry < row0 + rows;
      const long long ly = ry - row0;
      if (quad_ok && row_ok && rx >= 0 && rx + 3 < W) {
        // band-planar raster, the four pixels inside the row and their address a multiple of four: one 32-bit load per
        // band and row instead of four byte loads with a bounds test each (the kernel is bound by its instruction count)
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          unsigned w = 0;
          if (ch < c) w = __ldg(reinterpret_cast<const unsigned*>(raster + (static_cast<long long>(s_band[ch]) * rows + ly) * W + rx));
#pragma unroll
          for (int j = 0; j < 4; ++j) raw[ch][py][j] = (w >> (8 * j)) & 0xFFu;
        }
        continue;
      }
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          raw[ch][py][j] = 0;
          if (ch < c && row_ok && rx + j >= 0 && rx + j < W)
            raw[ch][py][j] = layout_hwc ? raster[(ly * W + rx + j) * bands_total + s_band[ch]]
                                        : raster[(static_cast<long long>(s_band[ch]) * rows + ly) * W + rx + j];
        }
      }
      // (unaligned 32-bit loads emulated with two aligned loads + funnel shift measured slower than the byte loads,
      // 1.66 -> 1.95 ms per zone; only the aligned case above takes words)
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {                       // the two space-to-depth pixels
      __align__(16) __nv_bfloat16 v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = __float2bfloat16(0.f);
#pragma unroll
      for (int py = 0; py < 2; ++py)
#pragma unroll
        for (int px = 0; px < 2; ++px)
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)
            if (ch < c) v[(py * 2 + px) * c + ch] = s_lut[ch * 256 + raw[ch][py][2 * k + px]];
      uint4* d = reinterpret_cast<uint4*>(tile_out + (static_cast<long long>(Y) * T2 + X + k) * 16);
      d[0] = *reinterpret_cast<const uint4*>(v);
      d[1] = *reinterpret_cast<const uint4*>(v + 8);
    }
  }
}

int launch_extract_normalise_s2d(const uint8_t* raster, int layout_hwc, int bands_total, const int* band_idx,
                                 int c, long long W, long long H, long long row0, long long rows,
                                 const int* tile_xy, int n, int T, const __nv_bfloat16* lut,
                                 __nv_bfloat16* out, int num_sms, cudaStream_t stream, long long tile_stride) {
  if (n <= 0 || T <= 0) return 0;
  if (T % 4 != 0 || c < 1 || c > 4) return -2004;
  int bx = (num_sms * 8 + n - 1) / n;
  const int need = ((T / 2) * (T / 4) + 255) / 256;
  if (bx > need) bx = need;
  if (bx < 1) bx = 1;
  dim3 grid(bx, n);
#define FB_S2D(C_)                                                                                             \
  case C_:                                                                                                     \
    extract_normalise_s2d_kernel<C_><<<grid, 256, 0, stream>>>(raster, layout_hwc, bands_total, band_idx, W, H, \
                                                               row0, rows, tile_xy, n, T, lut, out, tile_stride); \
    break;
  switch (c) {
    FB_S2D(1)
    FB_S2D(2)
    FB_S2D(3)
    FB_S2D(4)
  }
#undef FB_S2D
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------ maxpool
// One thread per output pixel and 8-channel group, nine unconditional, independent 16-byte loads (out-of-range
// taps clamped into the window): 3.5 -> 4.5 TB/s against the version that skipped them with branches. (A
// rolling-window variant that fetches every input row once per column strip was measured 6 % slower: fewer loads
// in flight per thread; running stem + pool in L2-sized chunks of tiles, FB_FRONT_CHUNK, was 15-20 % slower.)
__global__ void __launch_bounds__(256)
maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int H,
                    int W, int C) {
  const int Ho = H >> 1, Wo = W >> 1, Cg = C >> 3;
  const long long total = static_cast<long long>(B) * Ho * Wo * Cg;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx % Cg);
    long long pix = idx / Cg;
    const int ow = static_cast<int>(pix % Wo);
    pix /= Wo;
    const int oh = static_cast<int>(pix % Ho);
    const int b = static_cast<int>(pix / Ho);
    // Out-of-range taps (row / column -1 only: H and W are even) are clamped to 0, which lies inside the same
    // window, so the maximum is unchanged and all nine 16-byte loads are unconditional and independent.
    const int r0 = 2 * oh > 0 ? 2 * oh - 1 : 0, c0 = 2 * ow > 0 ? 2 * ow - 1 : 0;
    const int rows[3] = {r0, 2 * oh, 2 * oh + 1}, cols[3] = {c0, 2 * ow, 2 * ow + 1};
    uint4 raw[9];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx)
        raw[dy * 3 + dx] = __ldg(reinterpret_cast<const uint4*>(
            in + ((static_cast<long long>(b) * H + rows[dy]) * W + cols[dx]) * C + g * 8));
    float m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) m[i] = -3.0e38f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const uint32_t w[4] = {raw[t].x, raw[t].y, raw[t].z, raw[t].w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
        m[2 * i] = fmaxf(m[2 * i], __low2float(b2));
        m[2 * i + 1] = fmaxf(m[2 * i + 1], __high2float(b2));
      }
    }
    uint32_t pk[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 b2 = __floats2bfloat162_rn(m[2 * i], m[2 * i + 1]);
      pk[i] = *reinterpret_cast<const uint32_t*>(&b2);
    }
    *reinterpret_cast<uint4*>(out + idx * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

int launch_maxpool3x3s2(const __nv_bfloat16* in, __nv_bfloat16* out, int B, int H, int W, int C,
                        int num_sms, cudaStream_t stream) {
  if ((H & 1) || (W & 1) || (C & 7)) return -2005;   // the clamped taps rely on even input sizes
  const long long total = static_cast<long long>(B) * (H / 2) * (W / 2) * (C / 8);
  if (total == 0) return 0;
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(num_sms) * 8;
  if (blocks > cap) blocks = cap;
  maxpool3x3s2_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(in, out, B, H, W, C);
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------ K4
// One 64-thread block per sample; fp32 FMA on CUDA cores (5.4 kMAC: far too small for tensor cores).
__global__ void __launch_bounds__(64)
metadata_mlp_kernel(const float* __restrict__ met, const float* __restrict__ w0,
                    const float* __restrict__ b0, const float* __restrict__ w1,
                    const float* __restrict__ b1, const float* __restrict__ w2,
                    const float* __restrict__ b2, float* __restrict__ out) {
  __shared__ float x[45], h0[64], h1[32];
  const int s = blockIdx.x, t = threadIdx.x;
  if (t < 45) x[t] = met[s * 45 + t];
  __syncthreads();
  {
    float acc = b0[t];
    for (int i = 0; i < 45; ++i) acc = fmaf(w0[t * 45 + i], x[i], acc);
    h0[t] = fmaxf(acc, 0.f);
  }
  __syncthreads();
  if (t < 32) {
    float acc = b1[t];
    for (int i = 0; i < 64; ++i) acc = fmaf(w1[t * 64 + i], h0[i], acc);
    h1[t] = fmaxf(acc, 0.f);
  }
  __syncthreads();
  if (t < 16) {
    float acc = b2[t];
    for (int i = 0; i < 32; ++i) acc = fmaf(w2[t * 32 + i], h1[i], acc);
    out[s * 16 + t] = fmaxf(acc, 0.f);
  }
}

int launch_metadata_mlp(const float* met, const float* const* wb, float* out, int n,
                        cudaStream_t stream) {
  if (n == 0) return 0;
  metadata_mlp_kernel<<<n, 64, 0, stream>>>(met, wb[0], wb[1], wb[2], wb[3], wb[4], wb[5], out);
  return static_cast<int>(cudaGetLastError());
}

// Logits (and blend accumulators) are stored LS = 16 floats per pixel for models with <= 16 classes, 32 above
// (the 19-class nomenclature): the kernels below are instantiated for both.
template <int LS>
__device__ __forceinline__ void load_logits(const float* p, float (&v)[LS]) {
  const float4* lp = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int k = 0; k < LS / 4; ++k) {
    const float4 f = __ldg(lp + k);
    v[4 * k] = f.x; v[4 * k + 1] = f.y; v[4 * k + 2] = f.z; v[4 * k + 3] = f.w;
  }
}
// soft-max over the first ncls entries in place (entries >= ncls become 0); returns the arg-max (first maximum)
template <int LS>
__device__ __forceinline__ int softmax_ls(float (&v)[LS], int ncls, float* pmax = nullptr) {
  float best = v[0];
  int arg = 0;
#pragma unroll
  for (int k = 1; k < LS; ++k)
    if (k < ncls && v[k] > best) { best = v[k]; arg = k; }
  float den = 0.f;
#pragma unroll
  for (int k = 0; k < LS; ++k) {
    v[k] = k < ncls ? __expf(v[k] - best) : 0.f;
    den += v[k];
  }
  const float inv = 1.f / den;
#pragma unroll
  for (int k = 0; k < LS; ++k) v[k] *= inv;
  if (pmax) *pmax = inv;   // exp(0) / den
  return arg;
}
#define FB_LS_DISPATCH(ls_, ...)                      \
  do {                                                \
    if ((ls_) == 16) { constexpr int LS = 16; __VA_ARGS__; } \
    else if ((ls_) == 32) { constexpr int LS = 32; __VA_ARGS__; } \
    else return -2006;                                \
  } while (0)

// ------------------------------------------------------------------------------------------ K6
// One thread per written pixel. argmax = first maximum (numpy semantics); confidence byte =
// round-half-up of the max soft-max probability (what a float32 -> uint8 raster write produces).
template <int LS>
__global__ void __launch_bounds__(256)
argmax_stitch_kernel(const float* __restrict__ logits, int ncls, int T, const int* __restrict__ tiles,
                     uint8_t* __restrict__ cls_map, uint8_t* __restrict__ conf_map, long long map_w,
                     long long map_row0) {
  const int t = blockIdx.y;
  const int* tt = tiles + 6 * t;  // x0, y0 (tile origin, raster px), wx0, wy0, wx1, wy1 (write rect)
  const int x0 = tt[0], y0 = tt[1], wx0 = tt[2], wy0 = tt[3], wx1 = tt[4], wy1 = tt[5];
  const int rw = wx1 - wx0, rh = wy1 - wy0;
  if (rw <= 0 || rh <= 0) return;
  const int total = rw * rh;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ry = wy0 + i / rw, rx = wx0 + i % rw;
    const int ty = ry - y0, tx = rx - x0;
    float v[LS];
    load_logits<LS>(logits + ((static_cast<long long>(t) * T + ty) * T + tx) * LS, v);
    // 16 classes at a time, in the order and with the arithmetic of the head's fused sink (conv_halo.cu), which sees
    // the logits as 16-column groups: maximum / arg-max / exponent sum of the first group, then the second group
    // rescales that sum to the joint maximum. The two paths therefore write identical bytes for any class count.
    float best, den;
    int arg;
    {
      float lo[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) lo[k] = v[k];
      softmax_max16(lo, ncls, 0, best, arg, den);
    }
    if constexpr (LS == 32) {
      if (ncls > 16) {
        float hi[16], best2, den2;
        int arg2;
#pragma unroll
        for (int k = 0; k < 16; ++k) hi[k] = v[16 + k];
        softmax_max16(hi, ncls, 16, best2, arg2, den2);
        const bool second = best2 > best;   // the earlier (lower) class wins ties
        const float joint = second ? best2 : best;
        den = den * __expf(best - joint) + den2 * __expf(best2 - joint);
        if (second) arg = arg2;
        best = joint;
      }
    }
    const float pmax = 1.f / den;
    const long long o = (static_cast<long long>(ry) - map_row0) * map_w + rx;
    cls_map[o] = static_cast<uint8_t>(arg);
    if (conf_map != nullptr) conf_map[o] = static_cast<uint8_t>(pmax + 0.5f);
  }
}

int launch_argmax_stitch(const float* logits, int ncls, int ls, int n, int T, const int* tiles, uint8_t* cls_map,
                         uint8_t* conf_map, long long map_w, long long map_row0, cudaStream_t stream) {
  if (n == 0) return 0;
  dim3 grid((T * T + 255) / 256 > 64 ? 64 : (T * T + 255) / 256, n);
  FB_LS_DISPATCH(ls, argmax_stitch_kernel<LS><<<grid, 256, 0, stream>>>(logits, ncls, T, tiles, cls_map, conf_map, map_w, map_row0));
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------ K6b
// class_prob output (zone_detect/dataset.py:15-21): every class probability as uint8(p * 255), truncated.
template <int LS>
__global__ void __launch_bounds__(256)
prob_stitch_kernel(const float* __restrict__ logits, int ncls, int T, const int* __restrict__ tiles,
                   uint8_t* __restrict__ prob_map, long long map_w, long long map_row0, long long plane) {
  const int t = blockIdx.y;
  const int* tt = tiles + 6 * t;
  const int x0 = tt[0], y0 = tt[1], wx0 = tt[2], wy0 = tt[3], wx1 = tt[4], wy1 = tt[5];
  const int rw = wx1 - wx0, rh = wy1 - wy0;
  if (rw <= 0 || rh <= 0) return;
  const int total = rw * rh;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ry = wy0 + i / rw, rx = wx0 + i % rw;
    float v[LS];
    load_logits<LS>(logits + ((static_cast<long long>(t) * T + (ry - y0)) * T + (rx - x0)) * LS, v);
    softmax_ls<LS>(v, ncls);
    const long long o = (static_cast<long long>(ry) - map_row0) * map_w + rx;
#pragma unroll
    for (int k = 0; k < LS; ++k)
      if (k < ncls) prob_map[k * plane + o] = static_cast<uint8_t>(v[k] * 255.f);
  }
}

int launch_prob_stitch(const float* logits, int ncls, int ls, int n, int T, const int* tiles, uint8_t* prob_map,
                       long long map_w, long long map_row0, long long map_rows, cudaStream_t stream) {
  if (n == 0) return 0;
  dim3 grid((T * T + 255) / 256 > 64 ? 64 : (T * T + 255) / 256, n);
  FB_LS_DISPATCH(ls, prob_stitch_kernel<LS><<<grid, 256, 0, stream>>>(logits, ncls, T, tiles, prob_map, map_w, map_row0, map_rows * map_w));
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------ K7 / K8
// Blended stitching: the intent of the weighted branches of zone_detect/compare.py:84-138 with the weights
// of test/tiles.py:97-108 (mode "exp", sigma 0.5) and the per-pixel normalisation of tiles.py:111-169
// (sum of the weights of the covering tiles) / tiles.py:54-94 (their count). The whole tile contributes,
// clipped to the raster. Sums are accumulated with floating-point atomics (order not fixed).
template <int LS>
__global__ void __launch_bounds__(256)
blend_accumulate_kernel(const float* __restrict__ logits, int ncls, int T, const int* __restrict__ tiles, int method,
                        float* __restrict__ acc, float* __restrict__ wsum, long long map_w, long long map_row0,
                        long long map_rows, long long W, long long H, int seq0) {
  const int t = blockIdx.y;
  const int x0 = tiles[6 * t], y0 = tiles[6 * t + 1];
  const int centre = T / 2;
  const float inv_dmax = 1.f / static_cast<float>(centre > T - 1 - centre ? centre : T - 1 - centre);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < T * T; i += gridDim.x * blockDim.x) {
    const int ty = i / T, tx = i % T;
    const long long ry = static_cast<long long>(y0) + ty, rx = static_cast<long long>(x0) + tx;
    if (rx < 0 || rx >= W || ry < 0 || ry >= H || ry < map_row0 || ry >= map_row0 + map_rows || rx >= map_w) continue;
    float v[LS];
    load_logits<LS>(logits + ((static_cast<long long>(t) * T + ty) * T + tx) * LS, v);
    float pmax;
    const int arg = softmax_ls<LS>(v, ncls, &pmax);
    const long long o = (ry - map_row0) * map_w + rx;
    if (method == 2) {
      const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(pmax)) << 32) |
                                     (static_cast<unsigned long long>((seq0 + t) & 0xFFFFFF) << 8) | static_cast<unsigned>(arg);
      atomicMax(reinterpret_cast<unsigned long long*>(acc) + o, key);
    } else {
      const int dy = ty > centre ? ty - centre : centre - ty, dx = tx > centre ? tx - centre : centre - tx;
      const float w = method == 0 ? 1.f : __expf(-0.5f * static_cast<float>(dy > dx ? dy : dx) * inv_dmax);
      float4* a4 = reinterpret_cast<float4*>(acc + o * LS);
#pragma unroll
      for (int k = 0; k < LS / 4; ++k)
        if (4 * k < ncls) atomicAdd(a4 + k, make_float4(v[4 * k] * w, v[4 * k + 1] * w, v[4 * k + 2] * w, v[4 * k + 3] * w));
      atomicAdd(wsum + o, w);
    }
  }
}

int launch_blend_accumulate(const float* logits, int ncls, int ls, int n, int T, const int* tiles, int method, float* acc,
                            float* wsum, long long map_w, long long map_row0, long long map_rows, long long W,
                            long long H, int seq0, cudaStream_t stream) {
  if (n == 0) return 0;
  dim3 grid((T * T + 255) / 256 > 128 ? 128 : (T * T + 255) / 256, n);
  FB_LS_DISPATCH(ls, blend_accumulate_kernel<LS><<<grid, 256, 0, stream>>>(logits, ncls, T, tiles, method, acc, wsum, map_w, map_row0,
                                                                              map_rows, W, H, seq0));
  return static_cast<int>(cudaGetLastError());
}

// class = arg-max of the blended probabilities (first maximum), confidence byte = round-half-up of the
// normalised maximum (same band-2 semantics as the exact-clipping path). Uncovered pixels stay 0.
template <int LS>
__global__ void __launch_bounds__(256)
blend_finalize_kernel(const float* __restrict__ acc, const float* __restrict__ wsum, int method, int ncls, long long npx,
                      uint8_t* __restrict__ cls_map, uint8_t* __restrict__ conf_map) {
  for (long long o = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; o < npx;
       o += static_cast<long long>(gridDim.x) * blockDim.x) {
    int arg = 0;
    float conf = 0.f;
    if (method == 2) {
      const unsigned long long key = reinterpret_cast<const unsigned long long*>(acc)[o];
      arg = static_cast<int>(key & 0xFF);
      conf = __uint_as_float(static_cast<unsigned>(key >> 32));
    } else {
      float v[LS];
      load_logits<LS>(acc + o * LS, v);
      float best = v[0];
#pragma unroll
      for (int k = 1; k < LS; ++k)
        if (k < ncls && v[k] > best) { best = v[k]; arg = k; }
      const float ws = wsum[o];
      conf = ws > 0.f ? best / ws : 0.f;
    }
    cls_map[o] = static_cast<uint8_t>(arg);
    if (conf_map != nullptr) conf_map[o] = static_cast<uint8_t>(conf + 0.5f);
  }
}

int launch_blend_finalize(const float* acc, const float* wsum, int method, int ncls, int ls, long long npx, uint8_t* cls_map,
                          uint8_t* conf_map, int num_sms, cudaStream_t stream) {
  if (npx == 0) return 0;
  FB_LS_DISPATCH(ls, blend_finalize_kernel<LS><<<num_sms * 8, 256, 0, stream>>>(acc, wsum, method, ncls, npx, cls_map, conf_map));
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------ K9
// Confusion matrix cm[truth][pred] (int64, rows = truth) over pairs with both labels < ncls; any other
// pair is dropped, which is what sklearn.confusion_matrix(labels=range(ncls)) does. `truth_sub` is
// subtracted from the truth byte with uint8 wrap-around first (the reference computes `mask - 1` on a
// uint8 array, so 0 wraps to 255 and is dropped). Per-block bins live in shared memory; lanes of a
// warp that hit the same bin are merged with __match_any_sync so hot bins cost one atomic per warp.
constexpr int kMaxCls = 32;

__global__ void __launch_bounds__(256)
confusion_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ truth, long long npx,
                 int ncls, int truth_sub, unsigned long long* __restrict__ cm) {
  __shared__ unsigned int bins[kMaxCls * kMaxCls];
  const int nb = ncls * ncls;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) bins[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long nvec = npx >> 4;  // 16 pixels per uint4
  const bool aligned = ((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(truth)) & 15) == 0;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long done = 0;
  if (aligned) {
    // every lane of a warp runs the same trip count (loop bound is rounded up per warp)
    const long long first = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    for (long long vbase = first - lane; vbase < nvec; vbase += stride) {
      const long long v = vbase + lane;
      const bool active = v < nvec;
      uint4 pv = make_uint4(0, 0, 0, 0), tv = make_uint4(0, 0, 0, 0);
      if (active) {
        pv = __ldg(reinterpret_cast<const uint4*>(pred) + v);
        tv = __ldg(reinterpret_cast<const uint4*>(truth) + v);
      }
      const uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w};
      const uint32_t tw[4] = {tv.x, tv.y, tv.z, tv.w};
#pragma unroll
      for (int w = 0; w < 4; ++w) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const unsigned pb = (pw[w] >> (8 * b)) & 0xFF;
          const unsigned tb = ((tw[w] >> (8 * b)) - truth_sub) & 0xFF;
          const bool ok = active && pb < static_cast<unsigned>(ncls) && tb < static_cast<unsigned>(ncls);
          const unsigned key = ok ? tb * ncls + pb : 0xFFFFu;
          const unsigned peers = __match_any_sync(0xFFFFFFFFu, key);
          if (ok && lane == __ffs(peers) - 1) atomicAdd(&bins[key], __popc(peers));
        }
      }
    }
    done = nvec << 4;
  }
  // scalar tail (and the whole range when the pointers are not 16-byte aligned)
  for (long long i = done + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < npx;
       i += stride) {
    const unsigned pb = pred[i];
    const unsigned tb = (static_cast<unsigned>(truth[i]) - truth_sub) & 0xFF;
    if (pb < static_cast<unsigned>(ncls) && tb < static_cast<unsigned>(ncls))
      atomicAdd(&bins[tb * ncls + pb], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += blockDim.x)
    if (bins[i] != 0) atomicAdd(&cm[i], static_cast<unsigned long long>(bins[i]));
}

int launch_confusion(const uint8_t* pred, const uint8_t* truth, long long npx, int ncls, int truth_sub,
                     long long* cm, int num_sms, cudaStream_t stream) {
  if (ncls <= 0 || ncls > kMaxCls) return -2001;
  if (npx <= 0) return 0;
  // a block handles at most 2^32-1 pixels between flushes: 256 threads * 16 px * trips stays far below
  long long blocks = (npx / 16 + 255) / 256;
  if (blocks < 1) blocks = 1;
  const long long cap = static_cast<long long>(num_sms) * 8;
  if (blocks > cap) blocks = cap;
  confusion_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(
      pred, truth, npx, ncls, truth_sub, reinterpret_cast<unsigned long long*>(cm));
  return static_cast<int>(cudaGetLastError());
}

// K9 over a rectangle of two pitched uint8 maps (the write rectangles one rank owns in a sharded zone, the union of
// which is not a contiguous byte range): work item = 512 consecutive pixels of one row (16 per lane), same warp
// aggregation as above. 16-byte loads where the lane's addresses allow it, byte loads at ragged edges.
__global__ void __launch_bounds__(256)
confusion_rect_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ truth, long long rows,
                      long long width, long long pred_pitch, long long truth_pitch, int ncls, int truth_sub,
                      unsigned long long* __restrict__ cm) {
  __shared__ unsigned int bins[kMaxCls * kMaxCls];
  const int nb = ncls * ncls;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) bins[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long chunks = (width + 511) >> 9;
  const long long items = rows * chunks;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long nwarps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long it = warp0; it < items; it += nwarps) {
    const long long r = it / chunks, x = ((it - r * chunks) << 9) + lane * 16;
    const uint8_t* pp = pred + r * pred_pitch + x;
    const uint8_t* tp = truth + r * truth_pitch + x;
    const int valid = x >= width ? 0 : (width - x >= 16 ? 16 : static_cast<int>(width - x));
    uint32_t pw[4] = {0, 0, 0, 0}, tw[4] = {0, 0, 0, 0};
    if (valid == 16 && ((reinterpret_cast<uintptr_t>(pp) | reinterpret_cast<uintptr_t>(tp)) & 15) == 0) {
      const uint4 pv = __ldg(reinterpret_cast<const uint4*>(pp)), tv = __ldg(reinterpret_cast<const uint4*>(tp));
      pw[0] = pv.x; pw[1] = pv.y; pw[2] = pv.z; pw[3] = pv.w;
      tw[0] = tv.x; tw[1] = tv.y; tw[2] = tv.z; tw[3] = tv.w;
    } else {
      for (int i = 0; i < valid; ++i) {
        pw[i >> 2] |= static_cast<uint32_t>(pp[i]) << (8 * (i & 3));
        tw[i >> 2] |= static_cast<uint32_t>(tp[i]) << (8 * (i & 3));
      }
    }
#pragma unroll
    for (int w = 0; w < 4; ++w) {
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const unsigned pb = (pw[w] >> (8 * b)) & 0xFF;
        const unsigned tb = ((tw[w] >> (8 * b)) - truth_sub) & 0xFF;
        const bool ok = 4 * w + b < valid && pb < static_cast<unsigned>(ncls) && tb < static_cast<unsigned>(ncls);
        const unsigned key = ok ? tb * ncls + pb : 0xFFFFu;
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, key);
        if (ok && lane == __ffs(peers) - 1) atomicAdd(&bins[key], __popc(peers));
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += blockDim.x)
    if (bins[i] != 0) atomicAdd(&cm[i], static_cast<unsigned long long>(bins[i]));
}

int launch_confusion_rect(const uint8_t* pred, const uint8_t* truth, long long rows, long long width,
                          long long pred_pitch, long long truth_pitch, int ncls, int truth_sub, long long* cm,
                          int num_sms, cudaStream_t stream) {
  if (ncls <= 0 || ncls > kMaxCls) return -2001;
  if (rows <= 0 || width <= 0) return 0;
  const long long items = rows * ((width + 511) >> 9);
  long long blocks = (items + 7) / 8;   // 8 warps per block
  const long long cap = static_cast<long long>(num_sms) * 8;
  if (blocks > cap) blocks = cap;
  confusion_rect_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(
      pred, truth, rows, width, pred_pitch, truth_pitch, ncls, truth_sub, reinterpret_cast<unsigned long long*>(cm));
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------ K9b
// Per-tile confusion matrices of the compare loop (zone_detect/main.py:349-366 -> test/metrics.py:124-163,
// compute_metrics_patch): tile t's OWN arg-max prediction over its metric window against the truth raster, as
// sklearn.confusion_matrix(labels=range(ncls)) counts them. windows: int32 [n][6] = x0, y0 (tile origin),
// then the half-open window in raster pixels. truth: uint8 map with the class map's geometry; `truth_sub` is
// subtracted with uint8 wrap-around first. cm: u64 [n][ncls][ncls], accumulated.
template <int LS>
__global__ void __launch_bounds__(256)
tile_confusion_kernel(const float* __restrict__ logits, int ncls, int T, const int* __restrict__ windows,
                      const uint8_t* __restrict__ truth, int truth_sub, long long map_w, long long map_row0,
                      unsigned long long* __restrict__ cm) {
  __shared__ unsigned int bins[kMaxCls * kMaxCls];
  const int nb = ncls * ncls;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) bins[i] = 0;
  __syncthreads();
  const int t = blockIdx.y;
  const int* tt = windows + 6 * t;
  const int x0 = tt[0], y0 = tt[1], wx0 = tt[2], wy0 = tt[3], wx1 = tt[4], wy1 = tt[5];
  const int rw = wx1 - wx0, rh = wy1 - wy0;
  if (rw <= 0 || rh <= 0) return;
  const int total = rw * rh;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int ry = wy0 + i / rw, rx = wx0 + i % rw;
    float v[LS];
    load_logits<LS>(logits + ((static_cast<long long>(t) * T + (ry - y0)) * T + (rx - x0)) * LS, v);
    float best = v[0];
    unsigned arg = 0;
#pragma unroll
    for (int k = 1; k < LS; ++k)
      if (k < ncls && v[k] > best) { best = v[k]; arg = k; }
    const unsigned tb = (static_cast<unsigned>(truth[(static_cast<long long>(ry) - map_row0) * map_w + rx]) - truth_sub) & 0xFF;
    if (tb < static_cast<unsigned>(ncls)) atomicAdd(&bins[tb * ncls + arg], 1u);
  }
  __syncthreads();
  unsigned long long* out = cm + static_cast<long long>(t) * nb;
  for (int i = threadIdx.x; i < nb; i += blockDim.x)
    if (bins[i] != 0) atomicAdd(&out[i], static_cast<unsigned long long>(bins[i]));
}

int launch_tile_confusion(const float* logits, int ncls, int ls, int n, int T, const int* windows, const uint8_t* truth,
                          int truth_sub, long long map_w, long long map_row0, long long* cm, cudaStream_t stream) {
  if (ncls <= 0 || ncls > ls) return -2002;
  if (n == 0) return 0;
  dim3 grid(16, n);
  FB_LS_DISPATCH(ls, tile_confusion_kernel<LS><<<grid, 256, 0, stream>>>(logits, ncls, T, windows, truth, truth_sub & 0xFF, map_w, map_row0,
                                                                         reinterpret_cast<unsigned long long*>(cm)));
  return static_cast<int>(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------ active tiles
// Expands the per-image needed regions (tile_need.cuh) into the list of kernel tiles one conv launch walks.
// One block: per-image counts, a serial scan (n is a batch, <= a few hundred), then warps write the entries.
constexpr int kListMaxImages = 1024;

__global__ void __launch_bounds__(256)
build_tile_lists_kernel(const int* __restrict__ tiles, int n, int T, const __grid_constant__ TileListPlan plan,
                        int* __restrict__ list_base) {
  __shared__ int off[kListMaxImages + 1];
  __shared__ NeedRect rng[kListMaxImages];
  const TileListSpec& sp = plan.spec[blockIdx.x];
  if (!sp.use) return;
  int* list = list_base + sp.offset;
  for (int b = threadIdx.x; b < n; b += blockDim.x) {
    const int* t = tiles + 6 * b;
    const NeedRect r = need_rect(T, sp.layer, t[2] - t[0], t[3] - t[1], t[4] - t[0], t[5] - t[1]);
    NeedRect k;
    if (sp.shifted) {
      // rng = origin (x0, y0) and tile counts (x1, y1) of the shifted cover
      const NeedRect g = need_on_tile_grid(r, sp.scale);
      const NeedSpan sy = need_span(g.y0, g.y1, sp.th, sp.gh * sp.th);
      NeedSpan sx;
      if (sp.half_x) {
        // columns in blocks of tw / 2; k.x1 = tiles per row, bit 16 = the last one is narrow
        const NeedSpan bx = need_span(g.x0, g.x1, sp.tw / 2, sp.gw * sp.tw);
        sx.o = bx.o;
        sx.n = ((bx.n + 1) / 2) | ((bx.n & 1) << 16);
      } else {
        sx = need_span(g.x0, g.x1, sp.tw, sp.gw * sp.tw);
      }
      k.x0 = sx.o; k.y0 = sy.o; k.x1 = sx.n; k.y1 = sy.n;
      off[b + 1] = (sx.n & 0xFFFF) * sy.n;
    } else {
      k = need_tile_range(r, sp.scale, sp.th, sp.tw);
      if (k.x1 > sp.gw) k.x1 = sp.gw;
      if (k.y1 > sp.gh) k.y1 = sp.gh;
      off[b + 1] = (k.x1 - k.x0) * (k.y1 - k.y0);
    }
    rng[b] = k;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    off[0] = 0;
    for (int b = 0; b < n; ++b) off[b + 1] += off[b];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = warp; b < n; b += blockDim.x >> 5) {
    const NeedRect k = rng[b];
    const int nx = sp.shifted ? (k.x1 & 0xFFFF) : k.x1 - k.x0, cnt = off[b + 1] - off[b];
    const bool last_narrow = sp.shifted && (k.x1 >> 16) != 0;
    for (int i = lane; i < cnt; i += 32)
      list[off[b] + i] = sp.shifted ? static_cast<int>(pack_tile_origin(b, k.y0 + (i / nx) * sp.th, k.x0 + (i % nx) * sp.tw,
                                                                        last_narrow && i % nx == nx - 1))
                                    : (b * sp.gh + k.y0 + i / nx) * sp.gw + k.x0 + i % nx;
  }
  if (sp.sub > 1) {
    // sub-boxes are consumed sp.sub at a time: the list is padded with copies of its last entry (computed and stored again)
    __syncthreads();
    if (threadIdx.x == 0 && off[n] > 0)
      for (int i = off[n]; i % sp.sub != 0; ++i) list[i] = list[off[n] - 1];
  }
}

int launch_build_tile_lists(const int* tiles_dev, int n, int T, const TileListPlan& plan, int* list_dev,
                            cudaStream_t stream) {
  if (n <= 0) return 0;
  if (n > kListMaxImages) return -2101;
  build_tile_lists_kernel<<<kNeedLayers, 256, 0, stream>>>(tiles_dev, n, T, plan, list_dev);
  return static_cast<int>(cudaGetLastError());
}

long long count_active_tiles(const int* tiles, int n, int T, int layer, int scale, int th, int tw, bool shifted, bool half_x,
                             long long* blocks) {
  if (blocks) *blocks = 0;
  const int S = layer >= 10 ? T : (T / 16) << (layer / 2);
  const int gh = (S / scale) / th, gw = (S / scale) / tw;
  long long total = 0;
  for (int b = 0; b < n; ++b) {
    const int* t = tiles + 6 * b;
    const NeedRect r = need_rect(T, layer, t[2] - t[0], t[3] - t[1], t[4] - t[0], t[5] - t[1]);
    if (shifted) {
      const NeedRect g = need_on_tile_grid(r, scale);
      const long long ny = need_span(g.y0, g.y1, th, gh * th).n;
      if (half_x) {
        const long long nb = need_span(g.x0, g.x1, tw / 2, gw * tw).n;
        total += (nb + 1) / 2 * ny;
        if (blocks) *blocks += nb * ny;
      } else {
        const long long nx = need_span(g.x0, g.x1, tw, gw * tw).n;
        total += nx * ny;
        if (blocks) *blocks += 2 * nx * ny;
      }
      continue;
    }
    NeedRect k = need_tile_range(r, scale, th, tw);
    if (k.x1 > gw) k.x1 = gw;
    if (k.y1 > gh) k.y1 = gh;
    total += static_cast<long long>(k.x1 - k.x0) * (k.y1 - k.y0);
    if (blocks) *blocks += 2LL * (k.x1 - k.x0) * (k.y1 - k.y0);
  }
  return total;
}

}  // namespace fb
