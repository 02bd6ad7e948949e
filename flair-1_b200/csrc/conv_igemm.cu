// Implicit-GEMM convolution kernel for sm_100a (tcgen05.mma + TMEM accumulators + TMA / cp.async
// operand staging). See conv_igemm.cuh for the role layout and the reference call sites it replaces.
#include "conv_igemm.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "conv_epilogue.cuh"
#include "tile_need.cuh"
#include "ptx.cuh"

namespace fb {

using namespace ptx;

namespace {

constexpr int kNumProducerThreads = 128;  // warps 0-3
constexpr int kNumEpilogueThreads = 128;  // warps 4-7
constexpr int kNumThreads = 288;          // + warp 8 (MMA issuer / TMEM owner)

// EPI = epilogue warp groups. With 2, warps 9-12 form a second group of four and every accumulator tile is
// drained by both: group g takes columns [g*BN/2, (g+1)*BN/2) (a warp may only touch the 32 TMEM lanes of its
// own quarter, warp % 4, so a group always needs one warp per quarter). For BN <= 128 the drain of a tile
// otherwise takes longer than the MMAs of the next one and the tensor pipe idles on "accumulator empty".
template <int BN, int EPI = 1>
struct Cfg {
  static constexpr int kThreads = EPI == 2 ? 416 : kNumThreads;
#ifndef FB_IGEMM64_STAGES
#define FB_IGEMM64_STAGES 6
#endif
  static constexpr int kStages = (BN >= 256) ? 4 : (BN >= 128 && EPI == 2) ? 5 : (BN == 64 ? FB_IGEMM64_STAGES : 6);
  static constexpr int kABytes = kBM * kBK * 2;  // 16384
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int kBarBytes = ((2 * kStages + 4) * 8 + 16 + 127) / 128 * 128;
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kBarBytes + 4 * EPI * kStgWarpBytes;
};

template <int BN, bool TMA_A, int EPI = 1>
__global__ void __launch_bounds__((Cfg<BN, EPI>::kThreads), 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                  const __grid_constant__ CUtensorMap tmB, const ConvArgs p) {
  using C = Cfg<BN, EPI>;
  static_assert(EPI == 1 || (TMA_A && BN % 32 == 0), "two epilogue groups: TMA producer, BN/2 a multiple of 16");
  constexpr int BNE = BN / EPI;   // accumulator columns drained by one epilogue group
  constexpr int S = C::kStages;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;  // SW128 atoms need 1024-byte alignment
  uint8_t* smem = smem_raw + (base - raw_addr);

  const uint32_t bars = base + S * C::kStageBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * S + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S * C::kStageBytes + (2 * S + 4) * 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 8) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) {
        mbar_init(full_bar(s), TMA_A ? 1 : (kNumProducerThreads + 1));
        mbar_init(empty_bar(s), 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(tfull_bar(a), 1);
        mbar_init(tempty_bar(a), kNumEpilogueThreads * EPI);
      }
      fence_mbar_init();
      if (TMA_A) {
        tma_prefetch_desc(&tmA);
        if (p.C2 > 0) tma_prefetch_desc(&tmA2);
      }
      tma_prefetch_desc(&tmB);
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), C::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch (ptx.cuh): barriers, TMEM and descriptor prefetch are set up while the
  // predecessor drains; its outputs (and the tile list) are only touched after this point
  pdl_launch_dependents();
  pdl_wait();

  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int nk = p.num_k_iters;

  if (warp < 4) {
    // ===================================================================== producers
    if (TMA_A) {
      if (warp == 0 && elect_one()) {
        const int tiles_w = p.tgrid_w >> 4, tiles_h = p.tgrid_h >> 3;
        const int phases = p.phase_mode ? 4 : 1;
        uint32_t it = 0;
        // list lookups run one iteration ahead (a global load must not sit between two tiles)
        // list entries per tile: 1 whole box, 2 half boxes (4 x 16) or 4 quarter boxes (4 x 8)
        const int sub_n = p.tile_sub == 2 ? 4 : p.tile_sub == 1 ? 2 : 1;
        auto m_tile_of = [&](int t, int e) {
          const int idx = (t / p.num_n_tiles) / phases;
          return p.tile_list != nullptr ? __ldg(p.tile_list + sub_n * idx + e) : idx;
        };
        int m_next = 0, m_nexts[3] = {0, 0, 0};
        if (static_cast<int>(blockIdx.x) < num_tiles) {
          m_next = m_tile_of(blockIdx.x, 0);
#pragma unroll
          for (int e = 1; e < 4; ++e)
            if (e < sub_n) m_nexts[e - 1] = m_tile_of(blockIdx.x, e);
        }
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
          const int n_tile = tile % p.num_n_tiles;
          const int rest = tile / p.num_n_tiles;
          const int phase = rest % phases;
          const int m_tile = m_next;
          const int m_subs[3] = {m_nexts[0], m_nexts[1], m_nexts[2]};
          if (tile + static_cast<int>(gridDim.x) < num_tiles) {
            m_next = m_tile_of(tile + gridDim.x, 0);
#pragma unroll
            for (int e = 1; e < 4; ++e)
              if (e < sub_n) m_nexts[e - 1] = m_tile_of(tile + gridDim.x, e);
          }
          const int pa = p.phase_mode ? (phase >> 1) : 0, pb = p.phase_mode ? (phase & 1) : 0;
          int b, ty0, tx0;   // image and box origin on the tile grid
          int bs[3] = {0, 0, 0}, tys[3] = {0, 0, 0}, txs[3] = {0, 0, 0};   // the further sub-boxes
          if (p.tile_packed) {
            unpack_tile_origin(static_cast<uint32_t>(m_tile), b, ty0, tx0);
#pragma unroll
            for (int e = 1; e < 4; ++e)
              if (e < sub_n) unpack_tile_origin(static_cast<uint32_t>(m_subs[e - 1]), bs[e - 1], tys[e - 1], txs[e - 1]);
          } else {
            tx0 = (m_tile % tiles_w) * 16; ty0 = ((m_tile / tiles_w) % tiles_h) * 8; b = m_tile / (tiles_w * tiles_h);
          }
          for (int kit = 0; kit < nk; ++kit, ++it) {
            const int s = it % S;
            const uint32_t ph = (it / S) & 1;
            mbar_wait(empty_bar(s), ph ^ 1);
            mbar_expect_tx(full_bar(s), C::kStageBytes);
            const uint32_t e = p.ktab[kit];
            const int src = e & 1, cc = (e >> 1) & 0x7F;
            const int dx = static_cast<int>((e >> 8) & 15) - 8 + pb, dy = static_cast<int>((e >> 12) & 15) - 8 + pa;
            const int sc = p.tm_scale[src];
            const uint32_t a_dst = base + s * C::kStageBytes;
            tma_load_4d(a_dst, src ? &tmA2 : &tmA, full_bar(s), cc * 64, tx0 * sc + dx, ty0 * sc + dy, b);
            // (sub-boxes: the map's box is 4 rows high and 16 or 8 pixels wide; entry e fills rows e * 128 / sub_n .. of the
            // A tile, 128 bytes per row)
#pragma unroll
            for (int e = 1; e < 4; ++e)
              if (e < sub_n)
                tma_load_4d(a_dst + e * (128 / sub_n) * 128, src ? &tmA2 : &tmA, full_bar(s), cc * 64, txs[e - 1] * sc + dx,
                            tys[e - 1] * sc + dy, bs[e - 1]);
            tma_load_2d(a_dst + C::kABytes, &tmB, full_bar(s), kit * kBK, phase * p.Cout + n_tile * BN);
          }
        }
      }
    } else {
      const int tid = threadIdx.x;          // 0..127
      const int chunk = tid & 7;            // 16-byte chunk (8 channels) within the 128-byte k row
      const int rbase = tid >> 3;           // rows rbase + 16*j, j = 0..7
      const int Cin = p.C1 + p.C2;
      const int HWo = p.Hout * p.Wout;
      const uint32_t sw_off = static_cast<uint32_t>((chunk ^ (rbase & 7)) << 4);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.num_n_tiles, m_tile = tile / p.num_n_tiles;
        int rb[8], rih[8], riw[8];  // batch index (-1 = row out of range), top-left input coords
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int pix = m_tile * kBM + rbase + 16 * j;
          if (pix < p.M_total) {
            const int b = pix / HWo, rem = pix - b * HWo;
            const int oh = rem / p.Wout, ow = rem - oh * p.Wout;
            rb[j] = b;
            rih[j] = oh * p.stride - p.pad;
            riw[j] = ow * p.stride - p.pad;
          } else {
            rb[j] = -1; rih[j] = 0; riw[j] = 0;
          }
        }
        for (int kit = 0; kit < nk; ++kit, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          const uint32_t a_dst = base + s * C::kStageBytes;
          if (tid == 0) {
            mbar_expect_tx(full_bar(s), C::kBBytes);
            tma_load_2d(a_dst + C::kABytes, &tmB, full_bar(s), kit * kBK, n_tile * BN);
          }
          const int k0 = kit * kBK + chunk * 8;
          const bool kvalid = k0 < p.Ktot;
          const int tap = k0 / Cin, ch = k0 - tap * Cin;
          const int kh = tap / p.KW, kw = tap - kh * p.KW;
          const bool from1 = ch < p.C1;
          const __nv_bfloat16* src = from1 ? p.x1 : p.x2;
          const int Cs = from1 ? p.C1 : p.C2;
          const int chs = from1 ? ch : ch - p.C1;
          const int sh = (from1 && p.up1) ? 1 : 0;
          const int Hs = p.Hin >> sh, Ws = p.Win >> sh;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int ih = rih[j] + kh, iw = riw[j] + kw;
            const bool ok = kvalid && rb[j] >= 0 && ih >= 0 && ih < p.Hin && iw >= 0 && iw < p.Win;
            const __nv_bfloat16* g = src;
            if (ok) {
              g = src + (static_cast<size_t>(rb[j] * Hs + (ih >> sh)) * Ws + (iw >> sh)) * Cs + chs;
            }
            const uint32_t dst = a_dst + static_cast<uint32_t>((rbase + 16 * j) * 128) + sw_off;
            cp_async_16(dst, g, ok ? 16u : 0u);
          }
          // asynchronous arrival once this thread's copies have landed: no producer-side wait, all S stages
          // can be in flight (the MMA thread issues the generic -> async proxy fence after its wait)
          cp_async_mbar_arrive_noinc(full_bar(s));
        }
      }
    }
  } else if (warp != 8) {
    // ===================================================================== epilogue
    const int q = warp & 3;            // TMEM lane quarter this warp may access (tile rows 32q..32q+31)
    const int grp = warp > 8 ? 1 : 0;  // second group (EPI == 2): the upper half of the accumulator columns
    const int tiles_w = (TMA_A ? p.tgrid_w : p.Wout) >> 4, tiles_h = (TMA_A ? p.tgrid_h : p.Hout) >> 3;
    const int phases = (TMA_A && p.phase_mode) ? 4 : 1;
    const int HWo = p.Hout * p.Wout;
    uint8_t* stg = smem + S * C::kStageBytes + C::kBarBytes + (grp * 4 + q) * kStgWarpBytes;
    const bool f32 = p.out_f32 != nullptr;
    const int elem = f32 ? 4 : 2;
    const size_t pixel_bytes = static_cast<size_t>(p.Cout) * elem;
    const size_t up_row_bytes = static_cast<size_t>(2 * p.Wout) * pixel_bytes;
    uint8_t* out_bytes = f32 ? reinterpret_cast<uint8_t*>(p.out_f32) : reinterpret_cast<uint8_t*>(p.out);
    // TMA tiles are 8 x 16 pixel boxes: row r of the tile is pixel (r / 16, r % 16) of the box. In phase
    // mode the box lives on the low-res grid and pixel (dh, dw) lands at (2*dh + pa, 2*dw + pb).
    const int osc = (p.up2_out || phases == 4) ? 2 : 1;
    // (quarter boxes, tile_sub == 2: a warp's 32 rows are pixel (r / 8, r % 8) of its own 4 x 8 box)
    const bool quarter = TMA_A && p.tile_sub == 2;
    const EpiLane L = make_epi_lane(q, lane, f32 ? EpiRun<BNE>::GC_F32 * 4 : EpiRun<BNE>::GC_BF16 * 2,
                                    p.up2_out ? 2 * p.Wout : p.Wout, osc, [quarter](int r, int& dh, int& dw) {
                                      if (quarter) { dh = (r >> 3) & 3; dw = r & 7; }
                                      else { dh = r >> 4; dw = r & 15; }
                                    });
    uint32_t tcount = 0;
    // (half boxes: warps 0-1 of a group drain rows 0-63 = entry 2i, warps 2-3 rows 64-127 = entry 2i + 1; quarter boxes:
    // warp q drains entry 4i + q)
    const int sub_n = !TMA_A ? 1 : p.tile_sub == 2 ? 4 : p.tile_sub == 1 ? 2 : 1;   // list entries per tile
    const int sub_e = sub_n == 4 ? q : sub_n == 2 ? (q >> 1) : 0;
    auto m_tile_of = [&](int t) {
      const int idx = (t / p.num_n_tiles) / phases;
      return (TMA_A && p.tile_list != nullptr) ? __ldg(p.tile_list + sub_n * idx + sub_e) : idx;
    };
    int m_next = blockIdx.x < num_tiles ? m_tile_of(blockIdx.x) : 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
      const int n_tile = tile % p.num_n_tiles;
      const int rest = tile / p.num_n_tiles;
      const int phase = rest % phases;
      const int m_tile = m_next;
      if (tile + static_cast<int>(gridDim.x) < num_tiles) m_next = m_tile_of(tile + gridDim.x);
      const int as = tcount & 1;
      const uint32_t aph = (tcount >> 1) & 1;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + grp * BNE;
      const int n0 = n_tile * BN + grp * BNE;   // first output channel this warp drains
      if (TMA_A) {
        int tb, ty0, tx0;
        if (p.tile_packed) {
          unpack_tile_origin(static_cast<uint32_t>(m_tile), tb, ty0, tx0);
          if (sub_n == 2) ty0 -= 4 * sub_e;   // this warp's tile rows 4 * sub_e .. are rows 0 .. 3 of its half box
        } else {
          tx0 = (m_tile % tiles_w) * 16; ty0 = ((m_tile / tiles_w) % tiles_h) * 8; tb = m_tile / (tiles_w * tiles_h);
        }
        const int pa = phase >> 1, pb = phase & 1;
        const int psc = phases == 4 ? 2 : 1;   // output pixel = psc * tile-grid pixel + (pa, pb)
        const int oh = psc * (ty0 + L.own_dh) + pa, ow = psc * (tx0 + L.own_dw) + pb;
        const long long pix0 = (static_cast<long long>(tb) * p.Hout + psc * ty0 + pa) * p.Wout + psc * tx0 + pb;
        const long long up0 = (static_cast<long long>(tb) * 2 * p.Hout + ty0 * 2) * (2 * p.Wout) + tx0 * 2;
        uint8_t* tile_dst = out_bytes + static_cast<size_t>(p.up2_out ? up0 : pix0) * pixel_bytes;
        if (p.direct_store) {
          // every lane stores its own pixel from registers, one 32-byte sector per instruction (no staging)
          const long long own_pix = (static_cast<long long>(tb) * p.Hout + oh) * p.Wout + ow;
          const long long dpix = p.up2_out ? (static_cast<long long>(tb) * 2 * p.Hout + 2 * oh) * (2 * p.Wout) + 2 * ow : own_pix;
          uint8_t* own_dst = out_bytes + static_cast<size_t>(dpix) * pixel_bytes;
          auto store_regs = [&](uint8_t* d, const auto& regs) {
            if constexpr (sizeof(regs) == 32) {
              st_global_v8(d, regs);
            } else {
              uint32_t lo[8], hi[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                lo[i] = __float_as_uint(regs[i]);
                hi[i] = __float_as_uint(regs[8 + i]);
              }
              st_global_v8(d, lo);
              st_global_v8(d + 32, hi);
            }
          };
          auto direct = [&](int col0, const auto& regs) {
            uint8_t* d = own_dst + static_cast<size_t>(col0) * elem;
            store_regs(d, regs);
            if (p.up2_out) {
              store_regs(d + pixel_bytes, regs);
              store_regs(d + up_row_bytes, regs);
              store_regs(d + up_row_bytes + pixel_bytes, regs);
            }
          };
          epilogue_tile<BNE, true, false, true>(p, p.bias, taddr, tfull_bar(as), aph, lane, n0, stg, true, own_pix,
                                                      tb * p.Hout + oh, direct);
          tc_fence_before_sync();
          mbar_arrive(tempty_bar(as));
          continue;
        }
        auto copy = [&](auto run, int col0, int el) {
          warp_copy_out_fast<decltype(run)::value>(stg, lane, L, tile_dst + static_cast<size_t>(col0) * el, pixel_bytes,
                                                   p.up2_out, up_row_bytes);
        };
        epilogue_tile<BNE, true, false>(p, p.bias, taddr, tfull_bar(as), aph, lane, n0, stg, true,
                                              (static_cast<long long>(tb) * p.Hout + oh) * p.Wout + ow, tb * p.Hout + oh, copy);
      } else {
        // gather tiles are 128 consecutive pixels of the flattened (b, h, w) index space, possibly ragged
        auto rowfn = [&](int r, int& b, int& oh, int& ow) -> bool {
          const long long pix = static_cast<long long>(m_tile) * kBM + r;
          b = static_cast<int>(pix / HWo);
          const int rem = static_cast<int>(pix - static_cast<long long>(b) * HWo);
          oh = rem / p.Wout;
          ow = rem - oh * p.Wout;
          return pix < p.M_total;
        };
        auto copy = [&](auto run, int col0, int el) {
          warp_copy_out<decltype(run)::value>(p, stg, q, lane, col0, el, out_bytes, rowfn);
        };
        int b, oh, ow;
        const bool valid = rowfn(q * 32 + lane, b, oh, ow);
        epilogue_tile<BNE, true, false>(p, p.bias, taddr, tfull_bar(as), aph, lane, n0, stg, valid,
                                              static_cast<long long>(m_tile) * kBM + q * 32 + lane, b * p.Hout + oh, copy);
      }
      tc_fence_before_sync();
      mbar_arrive(tempty_bar(as));
    }
  } else {
    // ===================================================================== MMA issuer
    // elect_one(), not lane == 0: the compiler then issues each tcgen05.mma once from uniform registers
    // instead of wrapping it in a per-lane serialisation loop (2.3x the issue rate at N <= 64, tests/umma_probe.cu)
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN);
      uint32_t it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
        const int as = tcount & 1;
        const uint32_t aph = (tcount >> 1) & 1;
        mbar_wait(tempty_bar(as), aph ^ 1);
        tc_fence_after_sync();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kit = 0; kit < nk; ++kit, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(full_bar(s), ph);
          if (!TMA_A) fence_proxy_async_smem();  // A was written by cp.async (generic proxy)
          tc_fence_after_sync();
          const uint32_t a_addr = base + s * C::kStageBytes;
          const uint64_t adesc = umma_desc_sw128(a_addr);
          const uint64_t bdesc = umma_desc_sw128(a_addr + C::kABytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // +32 bytes per K=16 step inside the 128-byte swizzle row (start-address field is >>4)
            umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kit | k) != 0 ? 1u : 0u);
          }
          umma_commit(empty_bar(s));
        }
        umma_commit(tfull_bar(as));
      }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------ CTA-pair kernel
// Same implicit GEMM with tcgen05.mma.cta_group::2: a cluster of two CTAs (one TPC) computes a 256-pixel
// tile (two vertically adjacent 8x16 boxes) x BN channels. Each CTA stages its own box of A and HALF of
// the weight tile, so a K=16 step costs each SM 4 KB + BN*16 B of shared-memory reads instead of
// 4 KB + BN*32 B -- the single-CTA kernel is bound by exactly that traffic for BN >= 128.
// The leader (cluster rank 0) owns the "full" barriers (TMA bytes of both CTAs are counted there), issues
// the MMAs and multicasts the commits; both CTAs run their own producer thread and epilogue warps.
template <int BN, int EPI = 1>
struct Cfg2 {
  static constexpr int kThreads = EPI == 2 ? 416 : kNumThreads;
  // 6 x 32 KB or 8 x 24 KB of operands in flight per CTA (7 when the second epilogue group needs its staging buffers)
  static constexpr int kStages = BN >= 256 ? 6 : (EPI == 2 ? 7 : 8);
  static constexpr int kABytes = kBM * kBK * 2;        // 16384: this CTA's 128 pixels
  static constexpr int kBBytes = (BN / 2) * kBK * 2;   // this CTA's half of the weight tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;
  static constexpr int kBarBytes = ((2 * kStages + 4) * 8 + 16 + 127) / 128 * 128;
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kBarBytes + 4 * EPI * kStgWarpBytes;
};

// EPI = 2: a second group of four epilogue warps (9-12) in each CTA drains the upper half of the accumulator
// columns, as in the single-CTA kernel.
template <int BN, int EPI = 1>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__((Cfg2<BN, EPI>::kThreads), 1)
conv_igemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                   const __grid_constant__ CUtensorMap tmB, const ConvArgs p) {
  using C = Cfg2<BN, EPI>;
  constexpr int S = C::kStages;
  constexpr int BNE = BN / EPI;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const uint32_t bars = base + S * C::kStageBytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };          // used in the leader only
  auto empty_bar = [&](int s) { return bars + 8u * (S + s); };   // one per CTA, released by multicast commits
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * S + 2 + a); };  // leader only: 256 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + S * C::kStageBytes + (2 * S + 4) * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 8) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) {
        mbar_init(full_bar(s), 1);
        mbar_init(empty_bar(s), 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(tfull_bar(a), 1);
        mbar_init(tempty_bar(a), 2 * kNumEpilogueThreads * EPI);
      }
      fence_mbar_init();
      tma_prefetch_desc(&tmA);
      if (p.C2 > 0) tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB);
    }
    __syncwarp();
    tmem_alloc_2sm(smem_u32(tmem_slot), C::kTmemCols);
    tmem_relinquish_2sm();
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();   // barriers of both CTAs are initialised before any remote arrive / complete_tx
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_w = p.Wout >> 4, tiles_h2 = p.Hout >> 4;   // a pair covers 16 rows x 16 columns
  const int num_pairs = p.B * tiles_h2 * tiles_w;
  const int num_tiles = num_pairs * p.num_n_tiles;
  const int nk = p.num_k_iters;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp < 4) {
    if (warp == 0 && elect_one()) {
      const int c1chunks = p.C1 >> 6;
      const int cchunks = (p.C1 + p.C2) >> 6;
      uint32_t it = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int n_tile = tile % p.num_n_tiles, pair = tile / p.num_n_tiles;
        const int tw = pair % tiles_w, th2 = (pair / tiles_w) % tiles_h2, b = pair / (tiles_w * tiles_h2);
        const int h0 = th2 * 16 + 8 * static_cast<int>(rank);
        for (int kit = 0; kit < nk; ++kit, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          if (leader) mbar_expect_tx(full_bar(s), 2 * C::kStageBytes);
          const int tap = kit / cchunks, cc = kit - tap * cchunks;
          const int kh = tap / 3, kw = tap - kh * 3;
          const uint32_t a_dst = base + s * C::kStageBytes;
          if (cc < c1chunks) {
            tma_load_4d_2sm(a_dst, &tmA, full_bar(s), cc * 64, tw * 16 + kw - 1, h0 + kh - 1, b);
          } else {
            tma_load_4d_2sm(a_dst, &tmA2, full_bar(s), (cc - c1chunks) * 64, tw * 16 + kw - 1, h0 + kh - 1, b);
          }
          tma_load_2d_2sm(a_dst + C::kABytes, &tmB, full_bar(s), kit * kBK,
                          n_tile * BN + static_cast<int>(rank) * (BN / 2));
        }
      }
    }
  } else if (warp != 8) {
    const int q = warp & 3;
    const int grp = warp > 8 ? 1 : 0;
    uint8_t* stg = smem + S * C::kStageBytes + C::kBarBytes + (grp * 4 + q) * kStgWarpBytes;
    const bool f32 = p.out_f32 != nullptr;
    const int elem = f32 ? 4 : 2;
    const size_t pixel_bytes = static_cast<size_t>(p.Cout) * elem;
    const size_t up_row_bytes = static_cast<size_t>(2 * p.Wout) * pixel_bytes;
    uint8_t* out_bytes = f32 ? reinterpret_cast<uint8_t*>(p.out_f32) : reinterpret_cast<uint8_t*>(p.out);
    const EpiLane L = make_epi_lane(q, lane, f32 ? EpiRun<BNE>::GC_F32 * 4 : EpiRun<BNE>::GC_BF16 * 2,
                                    p.up2_out ? 2 * p.Wout : p.Wout, p.up2_out ? 2 : 1,
                                    [](int r, int& dh, int& dw) { dh = r >> 4; dw = r & 15; });
    uint32_t tcount = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++tcount) {
      const int n_tile = tile % p.num_n_tiles, pair = tile / p.num_n_tiles;
      const int as = tcount & 1;
      const uint32_t aph = (tcount >> 1) & 1;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + grp * BNE;
      const int tw = pair % tiles_w, th2 = (pair / tiles_w) % tiles_h2, tb = pair / (tiles_w * tiles_h2);
      const int h0 = th2 * 16 + 8 * static_cast<int>(rank);
      const int oh = h0 + L.own_dh, ow = tw * 16 + L.own_dw;
      const long long pix0 = (static_cast<long long>(tb) * p.Hout + h0) * p.Wout + tw * 16;
      const long long up0 = (static_cast<long long>(tb) * 2 * p.Hout + 2 * h0) * (2 * p.Wout) + tw * 32;
      uint8_t* tile_dst = out_bytes + static_cast<size_t>(p.up2_out ? up0 : pix0) * pixel_bytes;
      auto copy = [&](auto run, int col0, int el) {
        warp_copy_out_fast<decltype(run)::value>(stg, lane, L, tile_dst + static_cast<size_t>(col0) * el, pixel_bytes,
                                                 p.up2_out, up_row_bytes);
      };
      epilogue_tile<BNE, true, false>(p, p.bias, taddr, tfull_bar(as), aph, lane, n_tile * BN + grp * BNE, stg, true,
                                       (static_cast<long long>(tb) * p.Hout + oh) * p.Wout + ow, tb * p.Hout + oh, copy);
      tc_fence_before_sync();
      mbar_arrive_leader(tempty_bar(as));   // local arrive in the leader, remote arrive from the peer
    }
  } else if (leader && elect_one()) {
    // M = 256 (both CTAs' pixels), N = BN; descriptors name the leader's smem, the peer's data sits at
    // the same offsets of its own shared memory.
    constexpr uint32_t idesc = umma_idesc_bf16(2 * kBM, BN);
    uint32_t it = 0, tcount = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++tcount) {
      const int as = tcount & 1;
      const uint32_t aph = (tcount >> 1) & 1;
      mbar_wait(tempty_bar(as), aph ^ 1);
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int kit = 0; kit < nk; ++kit, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        mbar_wait(full_bar(s), ph);
        tc_fence_after_sync();
        const uint32_t a_addr = base + s * C::kStageBytes;
        const uint64_t adesc = umma_desc_sw128(a_addr);
        const uint64_t bdesc = umma_desc_sw128(a_addr + C::kABytes);
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k)
          umma_bf16_2sm(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kit | k) != 0 ? 1u : 0u);
        umma_commit_2sm(empty_bar(s), 0b11);     // frees the stage in both CTAs
      }
      umma_commit_2sm(tfull_bar(as), 0b11);      // accumulator ready for both epilogues
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();   // the peer may still be read by the leader's MMAs / arrive on its barriers
  if (warp == 8) {
    tc_fence_after_sync();
    tmem_dealloc_2sm(tmem_base, C::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

template <int BN, bool TMA_A, int EPI = 1>
int launch_t(const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmB, const ConvArgs& a,
             int grid, cudaStream_t stream) {
  using C = Cfg<BN, EPI>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel<BN, TMA_A, EPI>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = true;
  }
  static const bool pdl = !(getenv("FB_NO_PDL") && getenv("FB_NO_PDL")[0] == '1');
  const cudaError_t le = launch_kernel_pdl(conv_igemm_kernel<BN, TMA_A, EPI>, dim3(grid), dim3(C::kThreads),
                                           static_cast<size_t>(C::kSmemBytes), stream, pdl, tmA, tmA2, tmB, a);
  return static_cast<int>(le != cudaSuccess ? le : cudaGetLastError());
}

template <int BN, int EPI = 1>
int launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmB, const ConvArgs& a,
                int grid, cudaStream_t stream) {
  using C = Cfg2<BN, EPI>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_igemm2_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::kSmemBytes);
    if (e != cudaSuccess) return static_cast<int>(e);
    configured = true;
  }
  conv_igemm2_kernel<BN, EPI><<<grid, C::kThreads, C::kSmemBytes, stream>>>(tmA, tmA2, tmB, a);  // cluster dims are static (2,1,1)
  return static_cast<int>(cudaGetLastError());
}

}  // namespace

int encode_tma_plain_bf16(void* tensor_map, const void* base, int rank, const unsigned long long* dims,
                          const unsigned long long* strides, const unsigned* box, const unsigned* es) {
  if (init_tma_encoder() != 0) return -1001;
  cuuint64_t d[5], st[4];
  cuuint32_t bx[5], e[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; e[i] = es[i]; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = strides[i];
  const CUresult r = g_encode(static_cast<CUtensorMap*>(tensor_map), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank),
                              const_cast<void*>(base), d, st, bx, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -1300 - static_cast<int>(r);
}

int init_tma_encoder() {
  if (g_encode != nullptr) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) return -1;
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return 0;
}

int conv_pick_bn(int Cout) {
  if (Cout % 256 == 0) return 256;
  if (Cout % 128 == 0) return 128;
  if (Cout % 64 == 0) return 64;
  if (Cout % 32 == 0) return 32;
  if (Cout % 16 == 0) return 16;
  return 0;
}

bool conv_tma_eligible(const ConvArgs& a) {
  if (a.up1 != 0 || a.C1 % 64 != 0 || a.C2 % 64 != 0 || (a.C2 > 0 && a.x2 == nullptr)) return false;
  if (a.phase_mode) {  // x1 at half resolution, x2 at full resolution
    return a.KH == 3 && a.KW == 3 && a.stride == 1 && a.pad == 1 && a.Hout % 16 == 0 && a.Wout % 32 == 0 &&
           a.residual == nullptr && a.rowbias == nullptr && a.up2_out == 0 &&
           (9 * (a.C2 / 64) + 4 * (a.C1 / 64)) <= 128;
  }
  const bool k3 = a.KH == 3 && a.KW == 3 && a.pad == 1, k1 = a.KH == 1 && a.KW == 1 && a.pad == 0;
  if (!(k3 || k1) || (a.stride != 1 && a.stride != 2)) return false;
  if (a.stride == 2 && a.C2 != 0) return false;
  if (a.Hout % 8 != 0 || a.Wout % 16 != 0 || a.Hin != a.Hout * a.stride || a.Win != a.Wout * a.stride) return false;
  return a.KH * a.KW * ((a.C1 + a.C2) / 64) <= 128;
}

int launch_conv(const ConvArgs& a_in, const __nv_bfloat16* weights, int Kpad, bool use_tma_a,
                int num_sms, cudaStream_t stream) {
  if (init_tma_encoder() != 0) return -1001;
  ConvArgs a = a_in;
  const int BN = conv_pick_bn(a.Cout);
  if (BN == 0) return -1002;
  const int Cin = a.C1 + a.C2;
  if (Cin % 8 != 0 || a.C1 % 8 != 0 || Kpad % kBK != 0 || Kpad < a.Ktot) return -1003;
  a.M_total = a.B * a.Hout * a.Wout;
  a.num_n_tiles = a.Cout / BN;
  a.num_k_iters = (a.Ktot + kBK - 1) / kBK;
  {
    const char* ds = getenv("FB_DIRECT_STORE");
    a.direct_store = ds ? atoi(ds) == 2 : 0;   // measured neutral-to-slower here (not shared-memory bound): opt-in with 2
  }
  if (use_tma_a) {
    if (!conv_tma_eligible(a)) return -1004;
    const int c1c = a.C1 / 64, c2c = a.C2 / 64;
    int nk = 0;
    auto put = [&](int src, int chunk, int dx, int dy) {
      a.ktab[nk++] = static_cast<uint32_t>(src) | (static_cast<uint32_t>(chunk) << 1) |
                     (static_cast<uint32_t>(dx + 8) << 8) | (static_cast<uint32_t>(dy + 8) << 12);
    };
    if (a.phase_mode) {
      // k order = weight column order of pack_phase_weights: 4 low-res taps x C1, then 9 taps x C2
      a.tgrid_h = a.Hout / 2; a.tgrid_w = a.Wout / 2;
      a.tm_scale[0] = 1; a.tm_scale[1] = 2;
      for (int di = 0; di < 2; ++di)
        for (int dj = 0; dj < 2; ++dj)
          for (int c = 0; c < c1c; ++c) put(0, c, dj - 1, di - 1);   // + (pb, pa) added per phase in the kernel
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw)
          for (int c = 0; c < c2c; ++c) put(1, c, kw - 1, kh - 1);
      a.num_m_tiles = a.B * (a.tgrid_h / 8) * (a.tgrid_w / 16) * 4;
    } else {
      // k order = (kh*KW + kw)*Cin + c over the concatenated channels (pack of api.cu::build_conv)
      a.tgrid_h = a.Hout; a.tgrid_w = a.Wout;
      a.tm_scale[0] = a.stride; a.tm_scale[1] = 1;
      for (int kh = 0; kh < a.KH; ++kh)
        for (int kw = 0; kw < a.KW; ++kw) {
          for (int c = 0; c < c1c; ++c) put(0, c, kw - a.pad, kh - a.pad);
          for (int c = 0; c < c2c; ++c) put(1, c, kw - a.pad, kh - a.pad);
        }
      a.num_m_tiles = a.B * (a.tgrid_h / 8) * (a.tgrid_w / 16);
    }
    if (a.tile_list != nullptr) a.num_m_tiles = a.tile_list_len * (a.phase_mode ? 4 : 1);
    if (a.tile_sub && (a.tile_list == nullptr || !a.tile_packed)) return -1009;
    a.num_k_iters = nk;
    if (nk * kBK > Kpad) return -1006;
  } else {
    if (a.phase_mode) return -1007;
    if (a.tile_list != nullptr) return -1008;
    a.num_m_tiles = (a.M_total + kBM - 1) / kBM;
  }

  alignas(64) CUtensorMap tmA;
  alignas(64) CUtensorMap tmA2;
  alignas(64) CUtensorMap tmB;
  memset(&tmA, 0, sizeof(tmA));
  memset(&tmA2, 0, sizeof(tmA2));
  memset(&tmB, 0, sizeof(tmB));
  // CTA pairs (cta_group::2) for the wide layers: 16 x 16 pixel tiles, each CTA stages half of the weights, so a
  // K = 16 step costs each SM 4 KB + BN*16 B of shared-memory operand reads instead of 4 KB + BN*32 B.
  const char* pair_env = getenv("FB_PAIR");
  // Default: the N = 256 layers (measured +2.4 % on the zone at 148 tiles per pass, profiles/r01_v19_summary.md; N = 128
  // neutral). FB_PAIR=0: never; FB_PAIR=1: every eligible layer; FB_PAIR=<BN>: the layers of that width only.
  const int pair_sel = pair_env != nullptr ? atoi(pair_env) : 256;
  const bool pair_on = pair_sel == 1 || pair_sel == BN;
  const bool use_pair = use_tma_a && BN >= 64 && a.Hout % 16 == 0 && pair_on &&
                        a.KH == 3 && a.stride == 1 && !a.phase_mode && a.tile_list == nullptr;
  {
    // weights: [Cout][Kpad] bf16 ([4][Cout][Kpad] in phase mode), box = 64 k x BN rows (BN/2 per CTA of a
    // pair), 128-byte swizzle
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(Kpad), static_cast<cuuint64_t>(a.Cout) * (a.phase_mode ? 4 : 1)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(Kpad) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(kBK), static_cast<cuuint32_t>(use_pair ? BN / 2 : BN)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = g_encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(weights),
                          dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return -1100 - static_cast<int>(r);
  }
  if (use_tma_a) {
    // activations: NHWC bf16 seen as (C, W, H, B); box = 64 channels x 16 x 8 pixels; out-of-range
    // coordinates are zero-filled by the TMA unit, which is exactly the conv's zero padding.
    // A source read at stride s (stride-2 convs; the full-resolution skip tensor of phase mode) uses
    // elementStrides = s: the box spans 16*s x 8*s pixels and every s-th one lands in shared memory.
    for (int src = 0; src < (a.C2 > 0 ? 2 : 1); ++src) {
      const int Cs = src == 0 ? a.C1 : a.C2;
      const __nv_bfloat16* base = src == 0 ? a.x1 : a.x2;
      const int sc = a.tm_scale[src];
      // spatial dims of the tensor the source lives in
      const int Ws = a.phase_mode ? (src == 0 ? a.Wout / 2 : a.Wout) : a.Win;
      const int Hs = a.phase_mode ? (src == 0 ? a.Hout / 2 : a.Hout) : a.Hin;
      cuuint64_t dims[4] = {static_cast<cuuint64_t>(Cs), static_cast<cuuint64_t>(Ws),
                            static_cast<cuuint64_t>(Hs), static_cast<cuuint64_t>(a.B)};
      cuuint64_t strides[3] = {static_cast<cuuint64_t>(Cs) * 2, static_cast<cuuint64_t>(Ws) * Cs * 2,
                               static_cast<cuuint64_t>(Hs) * Ws * Cs * 2};
      cuuint32_t box[4] = {64, static_cast<cuuint32_t>((a.tile_sub == 2 ? 8 : 16) * sc), static_cast<cuuint32_t>((a.tile_sub ? 4 : 8) * sc), 1};
      cuuint32_t es[4] = {1, static_cast<cuuint32_t>(sc), static_cast<cuuint32_t>(sc), 1};
      CUresult r = g_encode(src == 0 ? &tmA : &tmA2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                            const_cast<__nv_bfloat16*>(base), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return -1200 - static_cast<int>(r);
    }
  }

  if (use_pair) {
    const int num_pairs = a.B * (a.Hout / 16) * (a.Wout / 16) * a.num_n_tiles;
    int clusters = num_sms / 2;
    if (num_pairs < clusters) clusters = num_pairs;
    if (clusters <= 0) return 0;
    const char* e2 = getenv("FB_EPI2");
    const bool epi2 = !(e2 && e2[0] == '0');
    return BN == 256 ? launch_pair<256>(tmA, tmA2, tmB, a, 2 * clusters, stream)
           : BN == 128 ? (epi2 ? launch_pair<128, 2>(tmA, tmA2, tmB, a, 2 * clusters, stream)
                               : launch_pair<128>(tmA, tmA2, tmB, a, 2 * clusters, stream))
                       : launch_pair<64>(tmA, tmA2, tmB, a, 2 * clusters, stream);
  }

  const int num_tiles = a.num_m_tiles * a.num_n_tiles;
  const int grid = num_tiles < num_sms ? num_tiles : num_sms;
  if (grid <= 0) return 0;

  // two epilogue warp groups where the drain of a tile is longer than the next tile's MMAs (FB_EPI2=0: one group)
  {
    const char* e2 = getenv("FB_EPI2");
    const bool epi2 = use_tma_a && !(e2 && e2[0] == '0');
    if (epi2 && BN == 128) return launch_t<128, true, 2>(tmA, tmA2, tmB, a, grid, stream);
    if (epi2 && BN == 64) return launch_t<64, true, 2>(tmA, tmA2, tmB, a, grid, stream);
  }

#define FB_DISPATCH(BN_)                                                               \
  case BN_:                                                                            \
    return use_tma_a ? launch_t<BN_, true>(tmA, tmA2, tmB, a, grid, stream)            \
                     : launch_t<BN_, false>(tmA, tmA2, tmB, a, grid, stream);
  switch (BN) {
    FB_DISPATCH(16)
    FB_DISPATCH(32)
    FB_DISPATCH(64)
    FB_DISPATCH(128)
    FB_DISPATCH(256)
  }
#undef FB_DISPATCH
  return -1005;
}

}  // namespace fb
