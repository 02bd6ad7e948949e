// Shared epilogue of the tcgen05 convolution kernels.
//
// One thread owns one output pixel (= one TMEM lane): it walks its BN accumulator columns in chunks of
// 16, applies folded-BN bias (+ residual) (+ ReLU) (+ per-row metadata bias) and rounds once to bf16.
// A lane's result row (<= 128 bytes per column group) goes to a per-warp staging buffer in shared
// memory; the warp then copies the 32 staged rows to global memory cooperatively, RUN/16 lanes per
// pixel, so every store instruction writes whole contiguous runs instead of 32 scattered 16-byte
// pieces. With up2_out each run is written to the 2x2 replicated positions, which materialises the
// decoder's nearest-neighbour x2 upsample (smp DecoderBlock).
//
// Where the tile-row -> pixel mapping does not depend on the tile (TMA 8x16 boxes, halo 16x8 blocks) the
// per-lane destination offsets of the copy-out rounds are computed once per kernel (EpiLane) so that a
// round costs one LDS, one 64-bit add and the store(s).
//
// The function performs the wait on the "accumulator full" mbarrier itself so that, with PREFETCH, the
// residual row of the pixel is already in flight while the MMAs of the tile still run.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#include "ptx.cuh"

namespace fb {

// Soft-max maximum of one pixel over the classes col0 .. col0 + 15 (those below ncls): first maximum (numpy arg-max
// semantics: the lower index wins ties) and the sum of exp(v - max), both as balanced trees over the 16 values --
// four dependent levels instead of fifteen, which is what paces the head's epilogue (ncu: its warps sit in fixed-latency
// dependency stalls). Shared by the fused class-map sinks of the halo kernel and by K6 (argmax_stitch_kernel), so the
// two paths keep writing identical bytes.
// (unmasked core: every one of the 16 values takes part)
__device__ __forceinline__ void softmax_max16_all(const float (&v)[16], int col0, float& best, int& arg, float& den) {
  float m[16];
  int ix[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    m[k] = v[k];
    ix[k] = col0 + k;
  }
#pragma unroll
  for (int w = 1; w < 16; w <<= 1) {
#pragma unroll
    for (int k = 0; k < 16; k += 2 * w) {
      const bool right = m[k + w] > m[k];   // strictly greater: the earlier class keeps a tie
      m[k] = right ? m[k + w] : m[k];
      ix[k] = right ? ix[k + w] : ix[k];
    }
  }
  best = m[0];
  arg = ix[0];
  float e[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) e[k] = __expf(v[k] - best);
#pragma unroll
  for (int w = 1; w < 16; w <<= 1) {
#pragma unroll
    for (int k = 0; k < 16; k += 2 * w) e[k] += e[k + w];
  }
  den = e[0];
}
// Columns at or beyond ncls are replaced by kSoftmaxMasked: they never win the maximum and their exponential is exactly
// zero, so a caller that already holds kSoftmaxMasked in those columns (the depth-to-space head folds it into its bias
// vector) gets the same bytes from softmax_max16_all without the 32 selects.
constexpr float kSoftmaxMasked = -3.0e38f;
__device__ __forceinline__ void softmax_max16(const float (&v)[16], int ncls, int col0, float& best, int& arg, float& den) {
  float m[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) m[k] = col0 + k < ncls ? v[k] : kSoftmaxMasked;
  softmax_max16_all(m, col0, best, arg, den);
}

constexpr int kStgPitch = 144;                 // bytes per staged pixel row (128 + 16: conflict-free 16 B stores)
constexpr int kStgWarpBytes = 32 * kStgPitch;  // staging bytes per epilogue warp

template <int BN>
struct EpiRun {  // bytes staged per pixel per copy-out
  static constexpr int GC_BF16 = BN < 64 ? BN : 64;
  static constexpr int GC_F32 = BN < 32 ? BN : 32;
};

// Tile-invariant per-lane geometry. dh/dw = position of a tile row inside the tile.
struct EpiLane {
  int own_dh, own_dw;   // this lane's own row
  int rnd_rel[8];       // copy-out round rd: pixel offset of the row this lane helps to write
                        // (dh*Wout + dw, or (2dh)*(2Wout) + 2dw when the output is 2x2 replicated)
};

// rowpos(r, dh, dw): position of tile row r. run_bytes: bytes staged per pixel (RUN).
// scale / pitch: destination pixel offset of tile position (dh, dw) = (scale*dh)*pitch + scale*dw
// (plain: 1, Wout; 2x2-replicated output: 2, 2*Wout; sub-pixel phase output: 2, Wout).
template <typename RowPos>
__device__ __forceinline__ EpiLane make_epi_lane(int warp_q, int lane, int run_bytes, int pitch, int scale, RowPos rowpos) {
  EpiLane L;
  rowpos(warp_q * 32 + lane, L.own_dh, L.own_dw);
  const int lpp = run_bytes / 16, ppr = 32 / lpp;
#pragma unroll
  for (int rd = 0; rd < 8; ++rd) {
    int dh = 0, dw = 0;
    if (rd < lpp) rowpos(warp_q * 32 + rd * ppr + lane / lpp, dh, dw);
    L.rnd_rel[rd] = (scale * dh) * pitch + scale * dw;
  }
  return L;
}

// Cooperative copy-out with precomputed offsets. dst0: byte address of (tile origin pixel, first staged column).
template <int RUN>
__device__ __forceinline__ void warp_copy_out_fast(const uint8_t* stg, int lane, const EpiLane& L, uint8_t* dst0,
                                                   size_t pixel_bytes, int up2, size_t up_row_bytes) {
  constexpr int LPP = RUN / 16, PPR = 32 / LPP;
  const int sub = lane % LPP;
  const uint8_t* s = stg + (lane / LPP) * kStgPitch + sub * 16;
  uint8_t* d0 = dst0 + sub * 16;
#pragma unroll
  for (int rd = 0; rd < LPP; ++rd) {
    const uint4 v = *reinterpret_cast<const uint4*>(s + rd * PPR * kStgPitch);
    uint8_t* d = d0 + static_cast<size_t>(L.rnd_rel[rd]) * pixel_bytes;
    *reinterpret_cast<uint4*>(d) = v;
    if (up2) {
      *reinterpret_cast<uint4*>(d + pixel_bytes) = v;
      *reinterpret_cast<uint4*>(d + up_row_bytes) = v;
      *reinterpret_cast<uint4*>(d + up_row_bytes + pixel_bytes) = v;
    }
  }
}

// Generic copy-out: rowfn(r, b, oh, ow) -> valid for tile row r (any mapping, ragged tiles).
template <int RUN, typename Args, typename RowFn>
__device__ __forceinline__ void warp_copy_out(const Args& p, const uint8_t* stg, int warp_q, int lane, int col0, int elem,
                                              uint8_t* out_bytes, RowFn rowfn) {
  constexpr int LPP = RUN / 16;   // lanes per pixel
  constexpr int PPR = 32 / LPP;   // pixels per round
  const int sub = lane % LPP;
#pragma unroll
  for (int rd = 0; rd < LPP; ++rd) {
    const int pr = rd * PPR + lane / LPP;
    int b, oh, ow;
    const bool valid = rowfn(warp_q * 32 + pr, b, oh, ow);
    if (valid) {
      const uint4 v = *reinterpret_cast<const uint4*>(stg + pr * kStgPitch + sub * 16);
      if (p.up2_out) {
        const size_t rowpitch = static_cast<size_t>(2 * p.Wout) * p.Cout * elem;
        uint8_t* d = out_bytes + (((static_cast<size_t>(b) * 2 * p.Hout + 2 * oh) * 2 * p.Wout + 2 * ow) * p.Cout + col0) * elem + sub * 16;
        const size_t px = static_cast<size_t>(p.Cout) * elem;
        *reinterpret_cast<uint4*>(d) = v;
        *reinterpret_cast<uint4*>(d + px) = v;
        *reinterpret_cast<uint4*>(d + rowpitch) = v;
        *reinterpret_cast<uint4*>(d + rowpitch + px) = v;
      } else {
        uint8_t* d = out_bytes + (((static_cast<size_t>(b) * p.Hout + oh) * p.Wout + ow) * p.Cout + col0) * elem + sub * 16;
        *reinterpret_cast<uint4*>(d) = v;
      }
    }
  }
}

// BN channels of one pixel's residual row into registers, one 32-byte sector per load (for EXT_RES).
template <int BN>
__device__ __forceinline__ void load_residual_row(const __nv_bfloat16* row, uint32_t (*res)[8]) {
#pragma unroll
  for (int i = 0; i < BN / 16; ++i) ptx::ld_global_nc_v8(row + 16 * i, res[i]);
}

// Args must provide: residual, rowbias, relu, out, out_f32, Cout, Hout, Wout, up2_out.
// bias: fp32 [>= n0 + BN]; BIAS_SMEM selects plain (shared-memory) loads instead of the read-only path.
// stg: this warp's staging buffer (kStgWarpBytes, 16-byte aligned).
// own_valid / own_pix / own_rb: this lane's row: in range?, linear output pixel index, rowbias index.
// copy(run_tag, col0, elem): cooperative copy-out of the staged column group starting at column col0.
// DIRECT: no staging; copy(col0, regs) receives each finished group of 16 columns in registers
// (const uint32_t (&)[8] = 16 packed bf16, or const float (&)[16]) and stores the lane's own pixel itself.
// EXT_RES: the caller has already loaded the residual row (ext_res[c][8] = channels 16c..16c+15 of this
// lane's pixel, see load_residual_row) so that the loads of the next block overlap this block's work.
template <int BN, bool PREFETCH, bool BIAS_SMEM, bool DIRECT = false, bool EXT_RES = false, typename Args, typename CopyFn>
__device__ __forceinline__ void epilogue_tile(const Args& p, const float* bias, uint32_t taddr, uint32_t tfull_bar,
                                              uint32_t tfull_parity, int lane, int n0, uint8_t* stg, bool own_valid,
                                              long long own_pix, int own_rb, CopyFn copy,
                                              const uint32_t (*ext_res)[8] = nullptr) {
  using namespace ptx;
  const float rb = (p.rowbias != nullptr && own_valid) ? p.rowbias[own_rb] : 0.f;
  const size_t obase = static_cast<size_t>(own_pix) * p.Cout + n0;
  const bool has_res = p.residual != nullptr && own_valid;
  if constexpr (EXT_RES) {
    // same body as below with the residual taken from the caller's registers
    mbar_wait_relaxed(tfull_bar, tfull_parity);
    tc_fence_after_sync();
    const bool f32x = p.out_f32 != nullptr;
    // (requesting chunk c + 1 before chunk c is processed measured neutral: ptxas already overlaps the next tcgen05.ld
    // with the current chunk's packing and stores)
#pragma unroll
    for (int c0 = 0; c0 < BN; c0 += 16) {
      uint32_t r[16];
      tmem_ld_x16(taddr + c0, r);
      float4 bb[4];
      const float4* bp = reinterpret_cast<const float4*>(bias + n0 + c0);
#pragma unroll
      for (int i = 0; i < 4; ++i) bb[i] = BIAS_SMEM ? bp[i] : __ldg(bp + i);
      tmem_ld_wait();
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
        for (int i = 0; i < 8; ++i) {
          const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&ext_res[c0 / 16][i]);
          v[2 * i + 0] += __low2float(b2);
          v[2 * i + 1] += __high2float(b2);
        }
      }
      if (p.relu) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] += rb;
      static_assert(!EXT_RES || DIRECT, "EXT_RES is only implemented for the direct-store epilogue");
      if (f32x) {
        copy(n0 + c0, v);
      } else {
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
          pk[i] = *reinterpret_cast<const uint32_t*>(&b2);
        }
        copy(n0 + c0, pk);
      }
    }
    return;
  }
  // residual row of this lane's pixel: 32-byte (one sector) loads of 16 channels, prefetched in groups of
  // <= 64 channels: group 0 before the accumulator wait, group g + 1 while group g is consumed
  constexpr int CPG = PREFETCH ? (BN <= 64 ? BN / 16 : 2) : 1;   // 16-channel chunks per group (registers: 2 * CPG * 8)
  constexpr int NG = PREFETCH ? BN / (16 * CPG) : 1;
  uint32_t res[NG > 1 ? 2 : 1][CPG][8];
  auto load_group = [&](int g, int buf) {
#pragma unroll
    for (int i = 0; i < CPG; ++i) ld_global_nc_v8(p.residual + obase + 16 * (g * CPG + i), res[buf][i]);
  };
  if (PREFETCH) {
#pragma unroll
    for (int b = 0; b < (NG > 1 ? 2 : 1); ++b)
#pragma unroll
      for (int i = 0; i < CPG; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) res[b][i][j] = 0u;
    if (has_res) load_group(0, 0);
  }
  mbar_wait_relaxed(tfull_bar, tfull_parity);
  tc_fence_after_sync();

  const bool f32 = p.out_f32 != nullptr;
  uint8_t* my_row = stg + lane * kStgPitch;
  constexpr int GC_BF16 = EpiRun<BN>::GC_BF16;
  constexpr int GC_F32 = EpiRun<BN>::GC_F32;
  // Unrolled body = two residual groups (so that the double-buffer index is a compile-time constant) or the
  // whole tile when it is a single group; wider tiles loop over the body to keep the code small.
  constexpr int CHUNKS = BN / 16;
  constexpr int UNR = PREFETCH ? (NG > 1 ? 2 * CPG : CHUNKS) : 2;
  static_assert(CHUNKS % UNR == 0 || CHUNKS < UNR, "tile width must be a whole number of unrolled bodies");
#pragma unroll 1
  for (int cb = 0; cb < CHUNKS; cb += UNR) {
#pragma unroll
  for (int u = 0; u < UNR; ++u) {
    if (cb + u >= CHUNKS) break;   // only BN = 16 without prefetch (UNR = 2 > CHUNKS)
    const int c0 = (cb + u) * 16;
    uint32_t r[16];
    tmem_ld_x16(taddr + c0, r);
    uint32_t rr[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    if (PREFETCH) {
      const int g = (cb + u) / CPG;
      constexpr int kBufMask = NG > 1 ? 1 : 0;
      const int i = u % CPG, buf = (u / CPG) & kBufMask;   // compile-time: cb is a multiple of 2 * CPG
      if (i == 0 && g + 1 < NG && has_res) load_group(g + 1, (buf ^ 1) & kBufMask);
#pragma unroll
      for (int j = 0; j < 8; ++j) rr[j] = res[buf][i][j];
    } else if (has_res) {
      ld_global_nc_v8(p.residual + obase + c0, rr);
    }
    float4 bb[4];
    const float4* bp = reinterpret_cast<const float4*>(bias + n0 + c0);
#pragma unroll
    for (int i = 0; i < 4; ++i) bb[i] = BIAS_SMEM ? bp[i] : __ldg(bp + i);
    tmem_ld_wait();
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
      for (int i = 0; i < 8; ++i) {
        const __nv_bfloat162 b2 = *reinterpret_cast<const __nv_bfloat162*>(&rr[i]);
        v[2 * i + 0] += __low2float(b2);
        v[2 * i + 1] += __high2float(b2);
      }
    }
    if (p.relu) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += rb;

    if constexpr (DIRECT) {
      if (f32) {
        copy(n0 + c0, v);
      } else {
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
          pk[i] = *reinterpret_cast<const uint32_t*>(&b2);
        }
        copy(n0 + c0, pk);
      }
    } else if (f32) {
      float4* sp = reinterpret_cast<float4*>(my_row + (c0 % GC_F32) * 4);
#pragma unroll
      for (int i = 0; i < 4; ++i) sp[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      if ((c0 + 16) % GC_F32 == 0) {
        __syncwarp();
        copy(ptx::Int<GC_F32 * 4>{}, n0 + c0 + 16 - GC_F32, 4);
        __syncwarp();
      }
    } else {
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const __nv_bfloat162 b2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        pk[i] = *reinterpret_cast<const uint32_t*>(&b2);
      }
      uint4* sp = reinterpret_cast<uint4*>(my_row + (c0 % GC_BF16) * 2);
      sp[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      sp[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      if ((c0 + 16) % GC_BF16 == 0) {
        __syncwarp();
        copy(ptx::Int<GC_BF16 * 2>{}, n0 + c0 + 16 - GC_BF16, 2);
        __syncwarp();
      }
    }
  }
  }
}

}  // namespace fb
