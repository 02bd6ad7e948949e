// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
//   out[p, n] = epilogue( sum_k A[p, k] * Wt[n, k] )          p = output pixel, n = output channel
//   k = (kh*KW + kw)*Cin + cin                                 (NHWC bf16 activations)
//
// One persistent CTA per SM walks (pixel-tile, channel-tile) pairs. Warp roles:
//   warps 0-3  producers : A tile (128 pixels x 64 k) either by TMA box loads ("TMA_A", 3x3 stride-1
//                          convs on an 8x16 spatial tile, zero fill = conv padding) or by a cp.async
//                          gather (any kernel size / stride / nearest-x2 upsample + channel concat of
//                          two sources); the weight tile always arrives by TMA.
//   warps 4-7  epilogue  : TMEM -> registers, + folded-BN bias (+ residual) (ReLU) (+ per-row metadata
//                          bias), one bf16 rounding, NHWC store (or fp32 logits).
//   warp  8    MMA issuer: single thread issues tcgen05.mma (M=128, N=BN, K=16) from the swizzled
//                          smem stages into a double-buffered TMEM accumulator.
// Replaces the cuDNN conv + BatchNorm + ReLU (+ add) calls issued by the reference's
// `model(imgs)` (src/zone_detect/compare.py:31, src/flair/model.py:57-64).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fb {

constexpr int kBM = 128;  // pixels per tile (UMMA M)
constexpr int kBK = 64;   // k elements per stage (128 bytes: one swizzle row)

struct ConvArgs {
  // sources (NHWC bf16). x2 is the skip tensor of a decoder block (may be null).
  const __nv_bfloat16* x1;
  const __nv_bfloat16* x2;
  int C1, C2;    // channels of x1 / x2; Cin = C1 + C2
  int up1;       // 1: x1 is read through a nearest x2 upsample (x1 spatial = Hin/2 x Win/2)
  int B, Hin, Win, Hout, Wout;
  int KH, KW, stride, pad;
  int Cout;      // multiple of BN
  int Ktot;      // KH*KW*Cin (weights are zero-padded to a multiple of 64 along k)
  // epilogue
  const float* bias;               // [Cout]
  const __nv_bfloat16* residual;   // [B,Hout,Wout,Cout] or null
  const float* rowbias;            // [B,Hout] added after ReLU (metadata MLP broadcast) or null
  int relu;
  __nv_bfloat16* out;              // [B,Hout,Wout,Cout] bf16 (null when out_f32 is used)
  float* out_f32;                  // [B,Hout,Wout,Cout] fp32 (logits) or null
  int up2_out;                     // 1: bf16 output is written 2x2-replicated into [B,2*Hout,2*Wout,Cout]
  // Sub-pixel ("phase") form of a 3x3 conv on [nearest-x2-upsampled x1 (+) x2]: x1 is given at LOW
  // resolution [B,Hout/2,Wout/2,C1]; the four output phases (oh%2, ow%2) are four 2x2 convs on x1 (taps
  // that hit the same low-res pixel are pre-summed in the weights) plus the plain 3x3 taps on x2 read at
  // stride 2. 4*C1 + 9*C2 instead of 9*(C1+C2) products per pixel and no materialised upsample.
  // weights: [4 phases][Cout][Kpad].
  int phase_mode;
  // TMA-producer kernel: the epilogue stores each lane's pixel from registers (32 B per instruction) instead
  // of staging through shared memory. Set by launch_conv: off unless FB_DIRECT_STORE=2 (a win in the halo
  // kernel, which is shared-memory-bandwidth bound; neutral to slower here).
  int direct_store;
  // tiling
  int M_total;       // B*Hout*Wout
  int num_m_tiles;   // gather: ceil(M_total/128); TMA: B*(tile-grid H/8)*(tile-grid W/16)
  int num_n_tiles;   // Cout / BN
  int num_k_iters;   // ceil(Ktot/64)
  // TMA producer schedule (filled by launch_conv): tiles are 8x16 boxes of a tile grid (the output, or
  // the low-res grid in phase mode); per k-iteration one box of 64 channels is loaded from source
  // (e&1) at chunk (e>>1)&0x7F and pixel offset (dx, dy) = (((e>>8)&15)-8, ((e>>12)&15)-8) after scaling
  // the box origin by tm_scale[source] (2 = stride-2 convs and the skip tensor of phase mode).
  int tgrid_h, tgrid_w;
  int tm_scale[2];
  uint32_t ktab[128];
  // Active-tile list (tile_need.cuh), TMA producer only: when non-null the kernel walks the tile_list_len 8x16
  // boxes tile_list[i] = (b * tgrid_h/8 + th) * tgrid_w/16 + tw instead of the full tile grid.
  const int* tile_list;
  int tile_list_len;
  int tile_packed;   // 1: entries are pack_tile_origin() words of origin-shifted boxes (tile_need.cuh, need_span)
  // With tile_packed: 1 = the entries are HALF boxes of 4 x 16 pixels, two per 128-row tile (rows 0-63 = entry 2i, rows
  // 64-127 = entry 2i + 1), 2 = QUARTER boxes of 4 x 8 pixels, four per tile (rows 32e .. = entry 4i + e = one epilogue
  // warp's rows); every entry is fetched by a TMA box of its own, tile_list_len counts tiles. A needed region of 20 x 20
  // pixels is then covered by 5 x 3 quarter boxes (480 pixels) instead of 3 x 2 whole ones (768).
  int tile_sub;
};

// Launch one convolution. `use_tma_a` requires conv_tma_eligible(a). Returns cudaError_t as int.
// TMA-eligible: (3x3 pad 1 or 1x1 pad 0), stride 1 or 2 (stride 2: single source), C1 % 64 == 0,
// C2 % 64 == 0, tile grid rows % 8 == 0 and columns % 16 == 0 (tile grid = output, or output / 2 in
// phase mode).
bool conv_tma_eligible(const ConvArgs& a);
int conv_pick_bn(int Cout);   // accumulator width the kernel uses for this Cout (256, 128, 64, 32 or 16)
int launch_conv(const ConvArgs& a, const __nv_bfloat16* weights /*[Cout][Kpad] bf16*/, int Kpad,
                bool use_tma_a, int num_sms, cudaStream_t stream);

// Resolve cuTensorMapEncodeTiled through the runtime (no link-time libcuda dependency).
int init_tma_encoder();
// cuTensorMapEncodeTiled for a bf16 tensor without swizzle / interleave, out-of-range elements read as zero (the halo
// kernel's TMA-staged input planes). dims / box / es: `rank` entries, innermost first; strides: rank - 1 entries in bytes.
// Returns 0 or a negative code.
int encode_tma_plain_bf16(void* tensor_map, const void* base, int rank, const unsigned long long* dims,
                          const unsigned long long* strides, const unsigned* box, const unsigned* es);

}  // namespace fb
