// Halo-staged convolution on tcgen05 (sm_100a) for the layers whose input is spatially large and
// channel-poor: the stem (7x7 stride 2 on the 8-channel-padded image), layer1 / dec2.conv2 (64 ch at
// 128^2), dec3 (64+64 / 32 ch at 256^2), dec4 and the segmentation head (32 / 16 ch at 512^2).
//
// For models with <= 4 bands the stem runs in 2x2 space-to-depth form instead: the 7x7 stride-2 filter, padded with a
// zero row and column in front, is a 4x4 stride-1 filter over [T/2][T/2][4*bands <= 16] (pixel (2Y+py, 2X+px), band c
// -> channel (py*2+px)*bands + c; filter tap (kh, kw) -> tap ((kh+1)/2, (kw+1)/2) of phase ((kh+1)%2, (kw+1)%2)),
// reading rows oh-2 .. oh+1: 16 K=16 steps instead of 28 and half the input bytes (KH = 4 instantiation).
//
// An im2col operand re-reads every input pixel KH*KW times. Here each CTA copies the (16*s+KH-s) x
// (8*s+KW-s) input halo of its 16 x 8 output tile into shared memory ONCE, as planes of 16-byte cells
// [channel-chunk][w-parity][h][w] (no swizzle), and every filter tap is just a different start address
// of the same UMMA shared-memory descriptor: 8 consecutive output columns are 8 consecutive 16-byte
// cells (the descriptor's 8-row core matrix), the next output row is `stride` plane rows further (SBO)
// and the second 8-channel chunk of a K=16 step is one plane (or, for the stem, one input row) further
// (LBO). The planes are written by TMA tensor loads (one 4-D box of 8 channels x halo width x halo height per
// plane, out-of-range pixels zero-filled = the conv padding, every second pixel through the map's element stride
// for the stride-2 forms), issued by one thread; by one cp.async per cell where a source is read through the x2
// up-sampling gather or the TMA unit already streams the filter bank. The filter bank stays resident in shared
// memory for the CTA's lifetime, or -- 128 -> 128 channels, 295 KB -- is streamed through a ring of bulk copies.
// Further forms (template parameters of conv_halo_kernel, conv_halo.cu): CTA pairs (tcgen05.mma.cta_group::2),
// depth-to-space output for the 16-channel layers, the stem with its max-pool fused into the epilogue.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fb {

constexpr int kHaloMaxSteps = 36;

struct HaloArgs {
  const __nv_bfloat16* x1;
  const __nv_bfloat16* x2;       // optional second source concatenated after x1's channels
  int C1, C2;
  int up1;                       // 1: x1 is [B, Hin/2, Win/2, C1] and is read through a nearest x2 upsample (3x3 stride 1 only)
  int B, Hin, Win, Hout, Wout;
  int Cout;                      // == BN of the instantiation
  // epilogue (same meaning as ConvArgs)
  const float* bias;
  const __nv_bfloat16* residual;
  const float* rowbias;
  int relu;
  __nv_bfloat16* out;
  float* out_f32;
  int up2_out;
  // filter bank packed by pack_halo_weights(): [group][step][chunk 0/1][n][8] bf16
  const __nv_bfloat16* wpacked;
  // streamed filter bank of a CTA pair (optional): the same bank with each stage's two 64-column halves contiguous,
  // [group][tap][half][step in tap][chunk 0/1][n < 64][8] (pack_halo_weights_pair128), so that a CTA fetches its half of
  // a stage with ONE bulk copy instead of eight 1 KB pieces
  const __nv_bfloat16* wpacked_pair;
  int groups1, groups2;          // channel groups (of NCH*8 channels) taken from x1 / x2
  int nsteps;                    // K=16 MMA steps per group
  // low word of the A descriptor of each K-step, relative to the stage base: start offset of the
  // step's first 8-channel chunk (16-byte units, bits 0-13) | distance to its second chunk << 16 (LBO)
  uint32_t a_lo[kHaloMaxSteps];
  int num_m_tiles;
  // 1: sub-pixel phase form. x1 is the LOW-res [B, Hin, Win, C1] input of a 3x3 conv on its nearest-x2
  // upsample, out is [B, 2*Hin, 2*Win, Cout]; a_lo holds 4 phases x nsteps entries and wpacked one filter
  // set per phase (pack_halo_weights_phase)
  int phase_mode;
  // Depth-to-space form of a 16-output-channel layer (64 accumulator columns = the 2x2 output pixels of one cell x 16
  // channels; the tile grid is the cell grid, out is [B, Hout, Wout, 16]):
  //   1: 3x3 stride-1 conv on [B, Hout, Wout, 16] run as a 4x4 stride-2 conv over cells (dec4.conv2, head);
  //   2: 3x3 conv of the nearest-x2 upsample of the LOW-res x1 [B, Hin, Win, 32], Hout = 2*Hin, run as a 3x3 conv on
  //      the low-res grid whose 64 outputs are the four output phases (dec4.conv1).
  // wpacked / bias come from pack_halo_weights_d2s (bias replicated four times), Cout = 64.
  int d2s;
  // 1: run as CTA pairs (tcgen05.mma.cta_group::2, M = 256 = two vertically adjacent 16-row tiles) where the shape has
  // that instantiation (3x3 stride 1, 64-channel groups, 64 or 128 output channels, no tile list); see conv_halo.cu
  int pair;
  // Fused stem max-pool (KH = 4 or 7 stem forms only): when non-null the kernel also writes MaxPool2d(3, 2, 1) of its
  // (post-ReLU) output to pool_out [B, Hout/2, Wout/2, 64]. keep_tiles (optional, int32 [B][6] = fb_tile of each image,
  // tile side keep_T): only the part of `out` that dec3.conv1 reads is stored (exact-clipping zone loop).
  __nv_bfloat16* pool_out;
  const int* keep_tiles;
  int keep_T;
  // The first 64 bias values by value (required by the fused-pool stem): kernel parameters live in the constant bank,
  // so the epilogue's bias add takes them as instruction operands instead of 32 shared-memory loads per tile and
  // thread -- in the fused stem those broadcast loads were a quarter of all LSU shared-memory wavefronts of a kernel
  // whose shared-memory pipe is saturated (ncu: LSU 60 % + tensor-core operand reads 42 %).
  int bias_in_args;
  float bias_c[64];
  int no_prefetch;               // 0 only with FB_PREFETCH=1: L2 prefetch of upcoming halos (measured neutral)
  // FB_HALO_SKIP bit mask, bottleneck hunting only (results are wrong): 1 = producers copy nothing,
  // 2 = no MMAs are issued, 4 = the epilogue does not store, 8 = the epilogue only does the barrier handshake
  int debug_skip;
  int direct_store;              // epilogue stores from registers (32 B per lane and instruction) instead of staging
  // Active-tile list (tile_need.cuh): when non-null the kernel walks tile_list[0 .. num_m_tiles) instead of the
  // full tile grid; every entry is a linear (image, tile row, tile column) index of the full grid.
  const int* tile_list;
  int tile_packed;               // 1: the entries are pack_tile_origin() words of origin-shifted tiles (tile_need.cuh, need_span)
  // Fused class-map sink of the segmentation head (Cout == 16 or 32, out_f32 set but never written): instead of
  // storing the fp32 logits of its pixel, a lane takes their soft-max maximum / arg-max (first maximum, numpy
  // semantics; confidence byte = round-half-up of the max probability) and, when the pixel lies inside the write
  // rectangle of its image, writes the two bytes straight into the class / confidence maps
  // (zone_detect/compare.py:35,66-82 + dataset.py:11-34 + the window write of main.py:421-423).
  // sink_tiles: int32 [B][6] = x0, y0, wx0, wy0, wx1, wy1 of the images of this launch.
  const int* sink_tiles;
  uint8_t* sink_cls;
  uint8_t* sink_conf;            // may be null
  long long sink_map_w, sink_map_row0;
  int sink_ncls;
};

// Geometry of one instantiation, shared by host packing and the kernel.
struct HaloGeom {
  int KH, stride, nch, bn, mb;
  int pad, np, ph, pw, kw_cells, plane16, stage_bytes, nsteps;
};
HaloGeom halo_geom(int KH, int stride, int nch, int bn);
int halo_blocks(int KH, int nch, int bn);  // 128-pixel blocks per CTA tile (tile = 16 rows x 8*blocks columns)

// True when (KH, stride, channels per group, Cout) has a compiled instantiation and the shape tiles.
bool halo_supported(int KH, int stride, int C1, int C2, int Cout, int Hout, int Wout);
int halo_group_channels(int KH, int C1, int C2);

// Fill a_lo / nsteps / groups / num_m_tiles for the given problem (host).
void halo_fill_steps(HaloArgs& a, int KH, int stride);

// Pack folded fp32 weights [Cout][Cin][KH][KW] (Cin = C1 + C2, already BN-scaled) into the step order.
// Returns the number of bf16 elements written to `dst` (dst may be null to query the size).
size_t pack_halo_weights(const float* w, int Cout, int CoutPad, int Cin, int CinPad, int KH, int stride,
                         int C1pad, int C2pad, uint16_t* dst);

// Re-orders a bank packed by pack_halo_weights for 128 output channels and 64-channel groups (`total` bf16 elements) into
// the CTA-pair stage order (HaloArgs::wpacked_pair); dst holds `total` elements.
void pack_halo_weights_pair128(const uint16_t* packed, size_t total, uint16_t* dst);

// Phase form (HaloArgs::phase_mode): single source of 32 channels -> 16 output channels.
bool halo_phase_supported(int C1, int C2, int Cout, int Hlo, int Wlo);
void halo_fill_steps_phase(HaloArgs& a);
// [phase = pa*2+pb][step = (di*2+dj)*(C/16) + kk][chunk 0/1][n][8]; taps of the 3x3 filter that read the same
// low-res pixel are summed before the single bf16 rounding.
size_t pack_halo_weights_phase(const float* w, int Cout, int CoutPad, int Cin, int CinPad, uint16_t* dst);

// Depth-to-space forms (HaloArgs::d2s). mode 1: Cin == 16, mode 2: Cin == 32 (low-res source); Cout <= 16.
bool halo_d2s_supported(int mode, int C1, int C2, int Cout, int Hout, int Wout);
void halo_fill_steps_d2s(HaloArgs& a, int mode);
// w: folded fp32 [Cout][Cin][3][3]. dst: the 64-output filter bank in the halo kernel's step order (mode 1: 16 taps of
// the 4x4 cell window, mode 2: 9 low-res taps with the taps that read the same low-res pixel summed before the single
// bf16 rounding); bias64: bias[co] at (py*2+px)*16 + co. Returns the bf16 element count (dst may be null).
size_t pack_halo_weights_d2s(int mode, const float* w, int Cout, int Cin, uint16_t* dst);

// The stem's max-pool can run inside its epilogue (HaloArgs::pool_out) only while the TMA-staged halo is on (FB_TMAH).
bool halo_pool_fusable();

int launch_conv_halo(const HaloArgs& a, int KH, int stride, int num_sms, cudaStream_t stream);

}  // namespace fb
