// Shared epilogue of the tcgen05 convolution kernels: one thread owns one output pixel (= one TMEM
// lane), walks its BN accumulator columns in chunks of 16, applies folded-BN bias (+ residual) (+ ReLU)
// (+ per-row metadata bias), rounds once to bf16 and stores NHWC -- optionally replicated 2x2 so that
// the consumer sees the decoder's nearest-neighbour x2 upsample (smp DecoderBlock) already materialised.
//
// The function also performs the wait on the "accumulator full" mbarrier itself so that, with
// PREFETCH, the residual row of the pixel is already in flight while the MMAs of the tile still run.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#include "ptx.cuh"

namespace fb {

// Args must provide: residual, rowbias, relu, out, out_f32, Cout, Hout, Wout, up2_out.
// bias: fp32 [>= n0 + BN]; BIAS_SMEM selects plain (shared-memory) loads instead of the read-only path.
template <int BN, bool PREFETCH, bool BIAS_SMEM, typename Args>
__device__ __forceinline__ void epilogue_pixel(const Args& p, const float* bias, uint32_t taddr, uint32_t tfull_bar,
                                               uint32_t tfull_parity, bool valid, int b, int oh, int ow, int n0) {
  using namespace ptx;
  const long long pix = (static_cast<long long>(b) * p.Hout + oh) * p.Wout + ow;
  const float rb = (p.rowbias != nullptr && valid) ? p.rowbias[b * p.Hout + oh] : 0.f;
  const size_t obase = static_cast<size_t>(pix) * p.Cout + n0;
  size_t ubase = 0, urow = 0;
  if (p.up2_out) {
    urow = static_cast<size_t>(2 * p.Wout) * p.Cout;
    ubase = ((static_cast<size_t>(b) * 2 * p.Hout + 2 * oh) * 2 * p.Wout + 2 * ow) * p.Cout + n0;
  }
  const bool has_res = p.residual != nullptr && valid;
  constexpr int NRES = PREFETCH ? BN / 8 : 1;
  uint4 res[NRES];
  if (PREFETCH) {
#pragma unroll
    for (int i = 0; i < NRES; ++i) res[i] = make_uint4(0, 0, 0, 0);
    if (has_res) {
      const uint4* rp = reinterpret_cast<const uint4*>(p.residual + obase);
#pragma unroll
      for (int i = 0; i < NRES; ++i) res[i] = __ldg(rp + i);
    }
  }
  mbar_wait(tfull_bar, tfull_parity);
  tc_fence_after_sync();

#pragma unroll(PREFETCH ? BN / 16 : 2)
  for (int c0 = 0; c0 < BN; c0 += 16) {
    uint32_t r[16];
    tmem_ld_x16(taddr + c0, r);
    uint4 rr[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
    if (PREFETCH) {
      rr[0] = res[(c0 / 8) % NRES];
      rr[1] = res[(c0 / 8 + 1) % NRES];
    } else if (has_res) {
      const uint4* rp = reinterpret_cast<const uint4*>(p.residual + obase + c0);
      rr[0] = __ldg(rp);
      rr[1] = __ldg(rp + 1);
    }
    float4 bb[4];
    const float4* bp = reinterpret_cast<const float4*>(bias + n0 + c0);
#pragma unroll
    for (int i = 0; i < 4; ++i) bb[i] = BIAS_SMEM ? bp[i] : __ldg(bp + i);
    tmem_ld_wait();
    if (valid) {
      float v[16];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        v[4 * i + 0] = __uint_as_float(r[4 * i + 0]) + bb[i].x;
        v[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + bb[i].y;
        v[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + bb[i].z;
        v[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + bb[i].w;
      }
      if (has_res) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t w[4] = {rr[h].x, rr[h].y, rr[h].z, rr[h].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
            v[8 * h + 2 * i + 0] += __low2float(b2);
            v[8 * h + 2 * i + 1] += __high2float(b2);
          }
        }
      }
      if (p.relu) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] += rb;
      if (p.out_f32 != nullptr) {
        float4* op = reinterpret_cast<float4*>(p.out_f32 + obase + c0);
#pragma unroll
        for (int i = 0; i < 4; ++i) op[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      } else {
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
          pk[i] = *reinterpret_cast<const uint32_t*>(&b2);
        }
        const uint4 lo = make_uint4(pk[0], pk[1], pk[2], pk[3]), hi = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        if (p.up2_out) {
#pragma unroll
          for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
              uint4* op = reinterpret_cast<uint4*>(p.out + ubase + dy * urow + static_cast<size_t>(dx) * p.Cout + c0);
              op[0] = lo;
              op[1] = hi;
            }
        } else {
          uint4* op = reinterpret_cast<uint4*>(p.out + obase + c0);
          op[0] = lo;
          op[1] = hi;
        }
      }
    }
  }
}

}  // namespace fb
