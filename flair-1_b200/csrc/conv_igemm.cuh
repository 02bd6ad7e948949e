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
  // tiling
  int M_total;       // B*Hout*Wout
  int num_m_tiles;   // gather: ceil(M_total/128); TMA: B*(Hout/8)*(Wout/16)
  int num_n_tiles;   // Cout / BN
  int num_k_iters;   // ceil(Ktot/64)
};

// Launch one convolution. `use_tma_a` requires KH=KW=3, stride=1, pad=1, no input upsample,
// C1 % 64 == 0 (and C2 % 64 == 0 when a second source is concatenated), Hout % 8 == 0,
// Wout % 16 == 0. Returns cudaError_t as int.
bool conv_tma_eligible(const ConvArgs& a);
int launch_conv(const ConvArgs& a, const __nv_bfloat16* weights /*[Cout][Kpad] bf16*/, int Kpad,
                bool use_tma_a, int num_sms, cudaStream_t stream);

// Resolve cuTensorMapEncodeTiled through the runtime (no link-time libcuda dependency).
int init_tma_encoder();

}  // namespace fb
