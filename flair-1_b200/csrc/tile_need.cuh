// Dead-output elimination for the exact-clipping zone loop.
//
// The reference runs the whole network on every margin-expanded tile and then throws the margin away
// (zone_detect/compare.py:66-82 crops `margin` pixels on each side before the window write). Only the
// write rectangle of a tile reaches the class map, so a decoder convolution only has to produce the
// outputs inside the receptive field of that rectangle: one pixel more per 3x3 conv walking back from
// the head, halved at every nearest-x2 upsample. The encoder is needed in full (its deep layers see the
// whole tile). Outputs inside the needed region are bit-identical to the full computation: a conv output
// depends on nothing but its own 3x3 input window, and that window lies in the previous layer's region.
//
// need_rect() is the one definition of those regions, shared by the host (tile counts, grid sizes, FLOP
// accounting) and the device (build_tile_list_kernel expands them into the per-launch list of active
// kernel tiles that conv_halo_kernel / conv_igemm_kernel walk instead of the full tile grid).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fb {

// layer ids: 2*d = dec<d>.conv1, 2*d + 1 = dec<d>.conv2 (d = 0..4), 10 = segmentation head
constexpr int kNeedLayers = 11;

struct NeedRect {
  int x0, y0, x1, y1;  // half-open, in the layer's output pixel grid; empty when x1 <= x0 or y1 <= y0
};

__host__ __device__ inline NeedRect need_grow(NeedRect r, int S) {
  r.x0 = r.x0 > 0 ? r.x0 - 1 : 0;
  r.y0 = r.y0 > 0 ? r.y0 - 1 : 0;
  r.x1 = r.x1 < S ? r.x1 + 1 : S;
  r.y1 = r.y1 < S ? r.y1 + 1 : S;
  return r;
}

// Output region of `layer` needed by a T x T tile whose write rectangle, relative to the tile origin, is
// [ax0, ax1) x [ay0, ay1). Decoder block d works at T / 16 * 2^d pixels per side.
__host__ __device__ inline NeedRect need_rect(int T, int layer, int ax0, int ay0, int ax1, int ay1) {
  NeedRect r;
  r.x0 = ax0 < 0 ? 0 : ax0; r.y0 = ay0 < 0 ? 0 : ay0;
  r.x1 = ax1 > T ? T : ax1; r.y1 = ay1 > T ? T : ay1;
  if (r.x1 <= r.x0 || r.y1 <= r.y0) { r.x0 = r.y0 = r.x1 = r.y1 = 0; return r; }
  if (layer >= 10) return r;
  int S = T;
  r = need_grow(r, S);                   // 3x3 window of the head -> output of dec4.conv2
  for (int d = 4; d >= 0; --d) {
    if (layer == 2 * d + 1) return r;    // output of dec<d>.conv2
    r = need_grow(r, S);                 // its 3x3 window
    if (layer == 2 * d) return r;        // output of dec<d>.conv1
    r = need_grow(r, S);                 // its 3x3 window, on the upsampled grid ...
    r.x0 >>= 1; r.y0 >>= 1;              // ... which reads these pixels of dec<d-1>'s output
    r.x1 = ((r.x1 - 1) >> 1) + 1; r.y1 = ((r.y1 - 1) >> 1) + 1;
    S >>= 1;
  }
  return r;
}

// Kernel-tile range [tx0, tx1) x [ty0, ty1) covering rect r, for kernel tiles of th x tw pixels on a tile grid
// that is the output grid divided by `scale` (2 for the sub-pixel phase kernels, which tile the low-res grid).
__host__ __device__ inline NeedRect need_tile_range(NeedRect r, int scale, int th, int tw) {
  NeedRect t;
  if (r.x1 <= r.x0 || r.y1 <= r.y0) { t.x0 = t.y0 = t.x1 = t.y1 = 0; return t; }
  const int sh = scale == 2 ? 1 : 0;
  t.x0 = (r.x0 >> sh) / tw; t.y0 = (r.y0 >> sh) / th;
  t.x1 = ((r.x1 - 1) >> sh) / tw + 1; t.y1 = ((r.y1 - 1) >> sh) / th + 1;
  return t;
}

// Origin-shifted tiling of one axis. The fixed tile grid wastes up to a tile on each side of a needed span (an
// interior zone tile needs 258 of dec4's 512 columns = 9 kernel tiles of 32, but the grid-aligned cover is 10; for
// the deep decoder layers 68 of 128 -> 5 tiles of 16 instead of 6, 36 of 64 -> 3 of 16 instead of 4). The kernels
// take their tile ORIGINS from the list, so the cover may start at the span itself: n = ceil(len / t) tiles from
// origin o = min(lo, extent - n * t) -- pulled back at the far edge so that no tile leaves the grid (extent is a
// multiple of t) and no store needs masking. lo / hi / extent in tile-grid units.
struct NeedSpan { int o, n; };
__host__ __device__ inline NeedSpan need_span(int lo, int hi, int t, int extent) {
  NeedSpan s;
  s.n = hi > lo ? (hi - lo + t - 1) / t : 0;
  s.o = lo < extent - s.n * t ? lo : extent - s.n * t;
  return s;
}
// rect r of a layer's output grid -> [lo, hi) on the tile grid (= output grid / scale) of each axis
__host__ __device__ inline NeedRect need_on_tile_grid(NeedRect r, int scale) {
  if (r.x1 <= r.x0 || r.y1 <= r.y0) { r.x0 = r.y0 = r.x1 = r.y1 = 0; return r; }
  const int sh = scale == 2 ? 1 : 0;
  r.x0 >>= sh; r.y0 >>= sh;
  r.x1 = ((r.x1 - 1) >> sh) + 1; r.y1 = ((r.y1 - 1) >> sh) + 1;
  return r;
}
// list entry of an origin-shifted tile: narrow flag | image | origin row | origin column on the tile grid
constexpr int kTileOriginBits = 11;   // grids up to 2047
constexpr int kTileImageBits = 9;     // up to 512 images per launch
constexpr uint32_t kTileNarrow = 1u << 31;   // halo kernels: only the first 8-column block of the tile is computed
__host__ __device__ inline uint32_t pack_tile_origin(int b, int y0, int x0, bool narrow = false) {
  return (narrow ? kTileNarrow : 0u) | (static_cast<uint32_t>(b) << (2 * kTileOriginBits)) |
         (static_cast<uint32_t>(y0) << kTileOriginBits) | static_cast<uint32_t>(x0);
}
__host__ __device__ inline void unpack_tile_origin(uint32_t e, int& b, int& y0, int& x0) {
  x0 = static_cast<int>(e & ((1u << kTileOriginBits) - 1));
  y0 = static_cast<int>((e >> kTileOriginBits) & ((1u << kTileOriginBits) - 1));
  b = static_cast<int>((e >> (2 * kTileOriginBits)) & ((1u << kTileImageBits) - 1));
}

// One conv launch's kernel tiling: tiles of th x tw pixels on the tile grid (= output grid / scale), gh x gw of them
// per image; its active tiles go to list[offset .. offset + count), images in order. Entries: (b * gh + ty) * gw + tx
// of the fixed grid, or -- shifted != 0 -- pack_tile_origin() of origin-shifted tiles (need_span).
struct TileListSpec {
  int layer, scale, th, tw, gh, gw;
  int offset, count;
  int use;   // 0: every tile is active, no list is built
  int shifted;
  // 1: the kernel computes a tile as two 8-column blocks and can skip the second (halo kernels, tw = 16): the columns are
  // covered in blocks of tw / 2, an odd block count ends every tile row with a narrow tile (kTileNarrow)
  int half_x;
  long long blocks;   // half_x: active blocks of the list (host bookkeeping)
  int sub;   // n > 1: entries are sub-boxes, n per kernel tile: the list is padded to a multiple of n with copies of its last entry
};
struct TileListPlan {
  TileListSpec spec[kNeedLayers];
};

// tiles: int32 [n][6] = x0, y0, wx0, wy0, wx1, wy1 (fb_tile). Builds every list of the plan in one launch (one
// block per layer). Returns cudaError_t as int.
int launch_build_tile_lists(const int* tiles_dev, int n, int T, const TileListPlan& plan, int* list_dev,
                            cudaStream_t stream);
// The count of one list on the host (tiles = host copy of the same table).
// blocks (optional): half_x lists -- the number of active 8-column blocks (tiles * 2 - narrow tiles), for the FLOP count
long long count_active_tiles(const int* tiles_host, int n, int T, int layer, int scale, int th, int tw, bool shifted,
                             bool half_x = false, long long* blocks = nullptr);

}  // namespace fb
