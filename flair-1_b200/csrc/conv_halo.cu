// Halo-staged tcgen05 convolution (see conv_halo.cuh for the idea and the layers it serves).
#include "conv_halo.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include <cuda.h>

#include "conv_epilogue.cuh"
#include "conv_igemm.cuh"
#include "ptx.cuh"
#include "tile_need.cuh"

namespace fb {

using namespace ptx;

namespace {

constexpr int kTH = 16;                  // output tile: 16 rows x (8*MB) columns = MB blocks of 128 pixels (UMMA M)
// 8 warps: 0-2 producers, 3 MMA issuer / TMEM owner, 4-7 epilogue. Two such CTAs fit one SM at
// <= 128 registers per thread (16 warps, 4 per scheduler), which is what the small layers need.
constexpr int kThreads = 256;
constexpr int kProducers = 96;
constexpr int kMmaWarp = 3;

// epilogue arguments of the depth-to-space layers (conv_epilogue.cuh reads residual / rowbias / relu / out_f32 / Cout)
struct D2SEpiArgs {
  static constexpr const __nv_bfloat16* residual = nullptr;
  static constexpr const float* rowbias = nullptr;
  int relu;
  const float* out_f32;
  int Cout;
};

template <int KH, int STRIDE, int NCH, int MB, int BN = 64, int EPI = 1, bool DEEP = false, bool PARB = false>
struct Geo {
  static constexpr int TW = 8 * MB;
  // 4x4 stride 2 = the 2x2-cell form of a 3x3 stride-1 conv (D2S): cell (Y, X) reads pixels 2Y-1 .. 2Y+2
  static constexpr int PAD = (KH == 4 && STRIDE == 2) ? 1 : KH / 2;
  // w-parity planes: one per column parity for the stride-2 forms; PARB (the fused-pool stem) splits the planes of a
  // stride-1 conv the same way so that the two 8-column blocks of a tile can be its even and its odd columns
  static constexpr int NP = PARB ? 2 : STRIDE;
  static constexpr int PH = STRIDE * (kTH - 1) + KH + (NCH == 1 ? 1 : 0);  // +1: stem pairs taps vertically
  static constexpr int SPANW = STRIDE * (TW - 1) + KH;
  static constexpr int PW = (SPANW + NP - 1) / NP;
  static constexpr int KWCELLS = NP * PW;
  static constexpr int PLANE16 = (PH * PW + 7) / 8 * 8;                // 16-byte cells per plane, padded to whole 128-byte rows (TMA destinations)
  static constexpr int STAGE = ((NCH * NP * PLANE16 * 16) + 127) / 128 * 128;
  static constexpr int CELLS = PH * KWCELLS * NCH;
  static constexpr int SBO16 = STRIDE * PW;                            // next output row, in 16-byte units
  // ring depth (= stages of loads in flight) / CTAs per SM, sized so that OCC CTAs fit 227 KB of shared
  // memory (filter bank + stages + 18 KB epilogue staging): the 16/32-channel layers run two CTAs per SM
  // (4x4 stride 2: a stage is the whole 34 x 34 pixel halo of a 16 x 16 cell tile, 37 KB, next to a 32 KB filter bank)
  // (with two epilogue groups a configuration runs one CTA per SM and spends the room on a deeper ring)
  static constexpr bool TWO = ((KH == 3 && NCH <= 4) || (KH == 4 && STRIDE == 2)) && EPI == 1;
  // (DEEP: the CTA-pair forms store from registers only, the 18 KB of copy-out staging buy a fourth stage)
  static constexpr int STAGES = (KH == 4 && STRIDE == 2) ? (EPI == 2 ? 4 : 2)
                                : (KH == 3 && NCH <= 4 && EPI == 2) ? 6
                                : (KH == 7 || NCH == 2 || (NCH == 4 && BN == 16)) ? 4 : (DEEP ? 4 : 3);
  static constexpr int OCC = TWO ? 2 : 1;
};

// PH = sub-pixel phase form of a 3x3 conv on the nearest-x2-upsampled x1 (single source): the tile is a
// 16 x 8 block of the LOW-res grid, the MB = 4 accumulators are the four output phases (oh%2, ow%2), each
// a 2x2-tap conv on the same staged halo with phase-specific (pre-summed) weights; outputs land at
// (2*h + pa, 2*w + pb) of the [B, 2*Hin, 2*Win, Cout] tensor.
// EPI = epilogue warp groups (1 or 2). With 2, warps 8-11 drain the odd 128-pixel blocks of every tile while warps
// 4-7 drain the even ones (direct-store mode only; a warp may only touch the TMEM lanes of its quarter, warp % 4,
// so a group is always four warps). Used for the one-CTA-per-SM configurations, whose drain (64 accumulator
// columns + residual per block) otherwise outlasts the MMAs of the next tile.
// D2S = depth-to-space output: the BN = 64 accumulator columns of a tile row are the 2x2 output pixels of one CELL
// (column (py*2 + px)*16 + co -> pixel (2Y + py, 2X + px), channel co of a 16-channel tensor), the tile grid is the
// cell grid. Serves the 16-channel layers at full resolution, whose N = 16 MMAs are bound by the fetch of the A
// operand (4 KB per 128 pixels and tap): as a 4x4 stride-2 conv over cells (head, dec4.conv2: 16 taps with N = 64 per
// 512 pixels instead of 4 x 9 taps with N = 16) or as the 3x3 conv on the low-res input whose 64 outputs are the four
// phases of the x2-upsampled conv (dec4.conv1: 18 MMAs with N = 64 instead of 32 with N = 16).
// SB = streamed filter bank: the layer's weights (295 KB for 128 -> 128 channels) do not fit in shared memory, so one
// extra warp streams them through a ring of kSbStages stages with 1-D bulk copies (a stage = the NCH / 2 K-steps of
// one filter tap of one channel group, contiguous in pack_halo_weights' order) while the input halo is still staged
// once per tile: against the im2col implicit GEMM (every input pixel fetched nine times, weights fetched once per 128
// pixels) this moves 185 KB instead of 576 KB from L2 to shared memory per 128 output pixels of layer2.
constexpr int kSbStages = 4;
#ifndef FB_ACC_DEEP
#define FB_ACC_DEEP 4
#endif
constexpr int kAccDeep = FB_ACC_DEEP;
// Halo gather through registers (ld.global.nc 16 B -> st.shared 16 B, a batch of loads in flight per thread) instead of
// cp.async. `ncu` counts 30-42 shared-memory wavefronts per warp-wide cp.async of 16 bytes per lane against 4 for the
// same bytes stored with st.shared.v4, and switching the copies off (FB_HALO_SKIP=1) makes the halo layers 15-30 %
// faster -- but a synchronous gather needs one L2 round trip per batch, and measured per layer (CUDA events, zone
// loop, profiles/r02_ldg_gather.txt) it only wins where a stage is one batch: the space-to-depth stem, 8 cells per
// thread (762 -> 642 us with the pool fused); layer1 314 -> 346 us, the streamed-weight layers 147 -> 280 us (16 loads
// in flight). FB_LDG_PRODUCER: 1 = the stem only (default), 0 = cp.async everywhere, 2 = registers everywhere.
#ifndef FB_LDG_PRODUCER
#define FB_LDG_PRODUCER 1
#endif
#ifndef FB_LDG_BATCH
#define FB_LDG_BATCH 8
#endif
constexpr int kLdgBatchMax = FB_LDG_BATCH;   // 16-byte loads a gathering thread keeps in flight (4 registers each)
constexpr int kLdgMode = FB_LDG_PRODUCER;   // accumulator buffers of the 64-channel one-CTA-per-SM configurations (2 or 4)

// PAIR = two CTAs of a cluster (one TPC) run every MMA together (tcgen05.mma.cta_group::2, M = 256): each CTA stages
// the halo of its own 16-row tile (the pair covers two vertically adjacent tiles) and holds HALF of the filter bank
// (BN / 2 output channels), so a K = 16 step costs each SM 4 KB + BN * 16 B of shared-memory operand reads instead of
// 4 KB + BN * 32 B -- the bound of the single-CTA form. The leader (cluster rank 0) issues the
// MMAs and multicasts the commits to both CTAs' barriers; the peer's MMA warp only forwards "my stage has landed"
// (its halo is written by cp.async, which can only signal a barrier of its own CTA) to the leader. Encoder layers
// only: no active-tile lists.
// POOL = the stem with its 3x3 stride-2 max-pool (torchvision ResNet: MaxPool2d(3, 2, 1)) fused into the epilogue. A CTA
// takes whole images and walks their 16 x 16 tiles in row-major order; the first epilogue group takes the horizontal
// 3-maximum of every finished tile in registers and leaves it in shared memory (bf16; the separate pool kernel would
// read the full-resolution tensor back from HBM, 8.4 MB per 512^2 tile), the second group finishes the pool vertically
// from there -- the row above comes from a carry buffer filled by the tile row before it, the column to the left from
// the previous tile -- and writes the pooled tensor. The stem's own output is still stored, for the decoder's skip
// connection: all of it, or, in the exact-clipping zone loop, only the part dec3.conv1 reads (HaloArgs::keep_tiles).
constexpr int kPoolPitch = 144;   // bytes per pixel of the shared-memory tile (128 + 16: conflict-free 16-byte stores)

template <int KH, int STRIDE, int NCH, int BN, int MB, bool PH = false, int EPI = 1, bool D2S = false, bool SB = false,
          bool PAIR = false, bool POOL = false, bool TMAH = false>
__global__ void __launch_bounds__(kThreads + 128 * (EPI - 1), Geo<KH, STRIDE, NCH, (PH ? 1 : MB), BN, EPI, PAIR, POOL>::OCC)
conv_halo_kernel(const __grid_constant__ HaloArgs p, const __grid_constant__ CUtensorMap tm1, const __grid_constant__ CUtensorMap tm2) {
  using G = Geo<KH, STRIDE, NCH, (PH ? 1 : MB), BN, EPI, PAIR, POOL>;
  // TMAH = the halo planes are written by TMA tensor loads (one 4-D box {8 channels, PW, PH, 1} per 8-channel chunk,
  // out-of-range pixels zero-filled = the conv padding) issued by one thread, instead of one cp.async per 16-byte cell
  // from 64-96 threads: single-source stride-1 layers only
  // (stride 2: one box per w-parity plane, every second pixel through the map's element stride)
  static_assert(!TMAH || !PH, "TMA-staged halo: not for the four-phase form");
  static_assert(!POOL || (EPI == 2 && MB == 2 && BN == 64 && !PH && !D2S && !SB && !PAIR && G::OCC == 1 && TMAH && STRIDE == 1),
                "fused max-pool: 16 x 16 tiles of 64 channels, two epilogue groups, TMA-staged parity planes");
  static_assert(!PAIR || (!PH && !D2S && G::OCC == 1 && EPI == 2 && BN % 32 == 0), "CTA pairs: plain form, one CTA per SM");
  constexpr int BNH = PAIR ? BN / 2 : BN;               // filter-bank columns held by this CTA
  static_assert(!D2S || (BN == 64 && !PH), "depth-to-space output: 4 pixels x 16 channels per tile row");
  static_assert(EPI == 1 || (G::OCC == 1 && !PH && MB % 2 == 0), "two epilogue groups: one CTA per SM, even block count");
  static_assert(!SB || (!PH && !D2S && G::OCC == 1 && NCH % 2 == 0), "streamed filter bank: plain form, one CTA per SM");
  constexpr int kThreadsK = kThreads + 128 * (EPI - 1);
  // SB: warp 2 streams the weights and warps 0-1 gather the halo (a 13th warp would cap the kernel at 128 registers:
  // four warps on one scheduler share its 16 K registers)
  constexpr int kBWarp = 2;
  constexpr int kProd = SB ? 64 : kProducers;           // halo-gathering threads
  constexpr int kCellsPerThread = (G::CELLS + kProd - 1) / kProd;
  constexpr int kBSteps = NCH >= 2 ? NCH / 2 : 1;       // K-steps per weight stage
  constexpr int kBStage = kBSteps * 2 * BNH * 16;       // bytes (of this CTA's half in a pair)
  static_assert(!PH || (MB == 4 && KH == 3 && STRIDE == 1), "phase form: 4 accumulators on a 3x3 stride-1 halo");
  constexpr int S = G::STAGES;
  // pairs: + "the peer's halo stage / weight stage has landed", used in the leader only
  constexpr int ACC = MB * BN;  // TMEM columns of one accumulator buffer (MB blocks of 128 x BN)
  // accumulator buffers: four where they fit the 512 TMEM columns of a one-CTA-per-SM configuration (the 64-channel
  // layers), two otherwise. Neutral while the halo was staged by cp.async (104.0 vs 103.9 ms per zone); with TMA
  // staging the MMAs run closer to their ceiling and the deeper ring is worth 0.5 % (91.6 vs 92.1 ms, three
  // alternating pairs). FB_ACC_DEEP=2: two everywhere.
  constexpr int NACC = (G::OCC == 1 && 4 * ACC <= 512 && !PH) ? kAccDeep : 2;
  constexpr int kBars = 2 * S + 2 * NACC + (SB ? 2 * kSbStages : 0) + (PAIR ? S + (SB ? kSbStages : 0) : 0);
  constexpr int kBarBytes = (kBars * 8 + 16 + 127) / 128 * 128;
  constexpr int kBiasBytes = BN * 4 <= 256 ? 256 : BN * 4;
  constexpr int TMEM_COLS = (NACC * ACC <= 32) ? 32 : (NACC * ACC <= 64) ? 64 : (NACC * ACC <= 128) ? 128 : (NACC * ACC <= 256) ? 256 : 512;
  static_assert(NACC * ACC <= 512, "the accumulator buffers must fit the 512 TMEM columns");

  extern __shared__ __align__(128) uint8_t smem[];
  const int groups = p.groups1 + p.groups2;
  // phase form: one filter set per phase; streamed: the ring
  const int wbytes = SB ? kSbStages * kBStage : (PH ? MB : 1) * groups * p.nsteps * 2 * BNH * 16;
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;
  const uint32_t w_addr = smem_base;
  const uint32_t bias_off = ((wbytes + 127) / 128) * 128;
  const uint32_t stage_addr0 = smem_base + bias_off + kBiasBytes;      // BN fp32 of bias
  const uint32_t bars = stage_addr0 + S * G::STAGE;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (S + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * S + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * S + NACC + a); };
  auto bfull_bar = [&](int s) { return bars + 8u * (2 * S + 2 * NACC + s); };
  auto bempty_bar = [&](int s) { return bars + 8u * (2 * S + 2 * NACC + kSbStages + s); };
  auto pfull_bar = [&](int s) { return bars + 8u * (2 * S + 2 * NACC + (SB ? 2 * kSbStages : 0) + s); };
  auto pbfull_bar = [&](int s) { return bars + 8u * (2 * S + 2 * NACC + (SB ? 2 * kSbStages : 0) + S + s); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (bars - smem_base) + kBars * 8);
  float* bias_s = reinterpret_cast<float*>(smem + bias_off);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // resident filter bank + bias: plain copies, made visible to the async proxy (tcgen05) below
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.wpacked);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    if (!SB && !(p.debug_skip & 16)) {
      if (PAIR) {
        // unit i of this CTA's bank = (step-chunk sc, column n < BN / 2) <- column rank * BN / 2 + n of the full bank
        for (int i = threadIdx.x; i < wbytes / 16; i += kThreadsK) dst[i] = __ldg(src + (i / BNH) * BN + rank * BNH + i % BNH);
      } else {
        for (int i = threadIdx.x; i < wbytes / 16; i += kThreadsK) dst[i] = __ldg(src + i);
      }
    }
    if (threadIdx.x < BN) bias_s[threadIdx.x] = p.bias[threadIdx.x];
  }
  if (warp == kMmaWarp) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) {
        mbar_init(full_bar(s), TMAH ? 1 : kProd);   // TMA: expect_tx by the issuing thread; else one arrival per gathering thread
        mbar_init(empty_bar(s), 1);
      }
      for (int a = 0; a < NACC; ++a) {
        mbar_init(tfull_bar(a), 1);
        // pairs: the leader's barrier takes both CTAs' epilogues; fused pool: only the first group drains accumulators
        mbar_init(tempty_bar(a), POOL ? 128 : 128 * EPI * (PAIR ? 2 : 1));
      }
      if (SB) {
        for (int s = 0; s < kSbStages; ++s) {
          mbar_init(bfull_bar(s), 1);
          mbar_init(bempty_bar(s), 1);
        }
      }
      if (PAIR) {
        for (int s = 0; s < S; ++s) mbar_init(pfull_bar(s), 1);
        if (SB)
          for (int s = 0; s < kSbStages; ++s) mbar_init(pbfull_bar(s), 1);
      }
      fence_mbar_init();
    }
    __syncwarp();
    if (PAIR) {
      tmem_alloc_2sm(smem_u32(tmem_slot), TMEM_COLS);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
      tmem_relinquish();
    }
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch: everything above touched only this CTA's shared memory / TMEM and the static
  // filter bank; from here on the kernel reads what its predecessor in the stream wrote
  pdl_launch_dependents();
  pdl_wait();

  const int tiles_w = (PH ? p.Win : D2S ? p.Wout / 2 : p.Wout) / G::TW, tiles_h = (PH ? p.Hin : D2S ? p.Hout / 2 : p.Hout) / kTH;
  // position i of this launch's schedule -> tile of the full grid (identity unless an active-tile list is given;
  // pairs: position = pair (image, tile-row pair, tile column), this CTA takes tile row 2 * pair row + rank)
  auto tile_of = [&](int i) {
    if (PAIR) {
      const int tw = i % tiles_w, r2 = i / tiles_w;
      return (2 * r2 + static_cast<int>(rank)) * tiles_w + tw;   // (b * tiles_h + 2 * th2 + rank) * tiles_w + tw
    }
    if (POOL) {
      // whole images per CTA (blockIdx.x, blockIdx.x + gridDim.x, ...), their tiles in row-major order
      const int per_img = tiles_w * tiles_h;
      return (static_cast<int>(blockIdx.x) + (i / per_img) * static_cast<int>(gridDim.x)) * per_img + i % per_img;
    }
    return p.tile_list != nullptr ? __ldg(p.tile_list + i) : i;
  };
  // tile -> image and origin (row, column) on the tile grid; packed entries carry the origin themselves (tile_need.cuh)
  auto tile_pos = [&](int tile, int& b, int& ty0, int& tx0) {
    if (!PAIR && !POOL && p.tile_packed) {
      unpack_tile_origin(static_cast<uint32_t>(tile), b, ty0, tx0);
    } else {
      tx0 = (tile % tiles_w) * G::TW; ty0 = ((tile / tiles_w) % tiles_h) * kTH; b = tile / (tiles_w * tiles_h);
    }
  };
  // packed entries may mark a tile as narrow: only its first 8-column block is computed and stored (tile_need.cuh)
  auto tile_narrow = [&](int tile) { return !PAIR && !POOL && !PH && p.tile_packed && (static_cast<uint32_t>(tile) & kTileNarrow) != 0; };
  // this CTA's (pair's) schedule: positions sched0, sched0 + sched_step, ... < sched_end
  const int sched0 = POOL ? 0 : PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int sched_step = POOL ? 1 : PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int sched_end = POOL ? ((p.B - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)) *
                                   tiles_w * tiles_h
                             : p.num_m_tiles;

  if (warp < (SB ? kBWarp : kMmaWarp)) {
    // ===================================================================== producers (halo gather)
    // The cells a thread copies do not depend on the tile: decode them once into one word per cell,
    // byte offset inside a stage | hh << 16 | k << 24. kProducers is a multiple of NCH, so the channel
    // chunk c of every cell of a thread is the same: tid % NCH.
    const int tid = threadIdx.x;
    if constexpr (TMAH) {
      if (warp == 0 && elect_one()) {
        uint32_t it = 0;
        for (int ti = sched0; ti < sched_end; ti += sched_step) {
          const int tile = tile_of(ti);
          int b, ty0, tx0;
          tile_pos(tile, b, ty0, tx0);
          const int ih0 = ty0 * STRIDE - G::PAD, iw0 = tx0 * STRIDE - G::PAD;
          for (int g = 0; g < groups; ++g, ++it) {
            const int s = it % S;
            mbar_wait_relaxed(empty_bar(s), ((it / S) & 1) ^ 1);
            const uint32_t st = stage_addr0 + s * G::STAGE;
            if (!(p.debug_skip & 1)) {
              mbar_expect_tx(full_bar(s), NCH * G::NP * G::PH * G::PW * 16);
#pragma unroll
              for (int c = 0; c < NCH; ++c)
#pragma unroll
                for (int par = 0; par < G::NP; ++par)
                  tma_load_4d(st + (c * G::NP + par) * (G::PLANE16 * 16), g < p.groups1 ? &tm1 : &tm2, full_bar(s),
                              ((g < p.groups1 ? g : g - p.groups1) * NCH + c) * 8, iw0 + par, ih0, b);
            } else {
              mbar_arrive(full_bar(s));
            }
          }
        }
      }
    } else {
    static_assert(kProd % NCH == 0, "channel chunk must be constant per producer thread");
    const int c8 = (tid % NCH) * 8;
    // (the 4x4 stride-2 depth-to-space form runs two CTAs per SM at 128 registers with a register-hungry epilogue: a
    // 25-word table per producer thread spilled to local memory and its reloads sat in the copy loop, so that
    // instantiation decodes each cell on the fly -- divisions by compile-time constants -- instead)
    constexpr bool kLdgProducer = kLdgMode == 2 || (kLdgMode == 1 && KH == 4 && STRIDE == 1);
    constexpr bool kCellTable = !(KH == 4 && STRIDE == 2) && !kLdgProducer;   // (register gather: the registers hold data instead)
    auto cell_word = [&](int j) -> uint32_t {
      const int idx = tid + j * kProd;
      const int c = idx % NCH;
      const int k = (idx / NCH) % G::KWCELLS;
      const int hh = idx / (NCH * G::KWCELLS);
      const uint32_t dst = static_cast<uint32_t>(((c * G::NP + (k % G::NP)) * G::PLANE16 + hh * G::PW + k / G::NP) * 16);
      return idx < G::CELLS ? (dst | (static_cast<uint32_t>(hh) << 16) | (static_cast<uint32_t>(k) << 24)) : 0xFFFFFFFFu;
    };
    uint32_t cell[kCellTable ? kCellsPerThread : 1];
    if constexpr (kCellTable) {
#pragma unroll
      for (int j = 0; j < kCellsPerThread; ++j) cell[j] = cell_word(j);
    }
    auto cellw = [&](int j) -> uint32_t {
      if constexpr (kCellTable) return cell[j];
      else return cell_word(j);
    };
    // Optional (FB_PREFETCH=1; measured neutral, so off by default): pull the halo rows of the tile
    // kPrefetchDist iterations ahead into L2, one bulk prefetch per halo row and source, clipped to the image.
    constexpr int kPrefetchDist = 2;
    const int nsrc = p.C2 > 0 ? 2 : 1;
    auto prefetch_tile = [&](int t) {
      if (t >= sched_end || p.no_prefetch || p.up1) return;
      t = tile_of(t);
      int b, ty0, tx0;
      tile_pos(t, b, ty0, tx0);
      const int iw0 = tx0 * STRIDE - G::PAD;
      const int w_lo = iw0 < 0 ? 0 : iw0, w_hi = iw0 + G::KWCELLS < p.Win ? iw0 + G::KWCELLS : p.Win;
      for (int r = tid; r < G::PH * nsrc; r += kProd) {
        const int sidx = r / G::PH, ih = ty0 * STRIDE - G::PAD + r % G::PH;
        if (static_cast<unsigned>(ih) >= static_cast<unsigned>(p.Hin)) continue;
        const int Cs = sidx ? p.C2 : p.C1;
        const __nv_bfloat16* row = (sidx ? p.x2 : p.x1) + ((static_cast<long long>(b) * p.Hin + ih) * p.Win + w_lo) * Cs;
        prefetch_l2_bulk(row, static_cast<uint32_t>((w_hi - w_lo) * Cs * 2));
      }
    };
    for (int d = 0; d < kPrefetchDist; ++d) prefetch_tile(sched0 + d * sched_step);
    uint32_t it = 0;
    // the tile id of the NEXT iteration is loaded now, so that the list lookup (a global load) never sits on the
    // path between two tiles
    int tile_next = sched0 < sched_end ? tile_of(sched0) : 0;
    for (int ti = sched0; ti < sched_end; ti += sched_step) {
      prefetch_tile(ti + kPrefetchDist * sched_step);
      const int tile = tile_next;
      if (ti + sched_step < sched_end) tile_next = tile_of(ti + sched_step);
      int b, ty0, tx0;
      tile_pos(tile, b, ty0, tx0);
      const int ih0 = ty0 * STRIDE - G::PAD, iw0 = tx0 * STRIDE - G::PAD;
      const bool interior = ih0 >= 0 && iw0 >= 0 && ih0 + G::PH <= p.Hin && iw0 + G::KWCELLS <= p.Win;
      const long long origin = (static_cast<long long>(b) * p.Hin + ih0) * p.Win + iw0;  // may be "negative"
      for (int g = 0; g < groups; ++g, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        mbar_wait_relaxed(empty_bar(s), ph ^ 1);
        const bool from1 = g < p.groups1;
        const int Cs = from1 ? p.C1 : p.C2;
        const __nv_bfloat16* src = (from1 ? p.x1 : p.x2) + origin * Cs + (from1 ? g : g - p.groups1) * (NCH * 8) + c8;
        const int row_elems = p.Win * Cs;  // element distance between halo rows
        const uint32_t st = stage_addr0 + s * G::STAGE;
        if constexpr (kLdgProducer) {
          // loads of a batch of cells in flight together, then their stores; zero for cells outside the image
          constexpr int kBatch = kLdgBatchMax < kCellsPerThread ? kLdgBatchMax : kCellsPerThread;
          const int Hlo = p.Hin >> 1, Wlo = p.Win >> 1;
          const __nv_bfloat16* lo = p.x1 + static_cast<long long>(b) * Hlo * Wlo * Cs + g * (NCH * 8) + c8;
          const bool up = p.up1 && from1;
          if (!(p.debug_skip & 1)) {
#pragma unroll 1
            for (int j0 = 0; j0 < kCellsPerThread; j0 += kBatch) {
              uint4 v[kBatch];
#pragma unroll
              for (int jj = 0; jj < kBatch; ++jj) {
                if (j0 + jj < kCellsPerThread) {
                  const uint32_t cw = cellw(j0 + jj);
                  v[jj] = make_uint4(0u, 0u, 0u, 0u);
                  if (cw != 0xFFFFFFFFu) {
                    const int hh = (cw >> 16) & 0xFF, k = cw >> 24;
                    const bool ok = interior || (static_cast<unsigned>(ih0 + hh) < static_cast<unsigned>(p.Hin) &&
                                                 static_cast<unsigned>(iw0 + k) < static_cast<unsigned>(p.Win));
                    const __nv_bfloat16* gp = up ? lo + (static_cast<long long>((ih0 + hh) >> 1) * Wlo + ((iw0 + k) >> 1)) * Cs
                                                 : src + hh * row_elems + k * Cs;
                    if (ok) v[jj] = ld_global_nc_v4(gp);
                  }
                }
              }
#pragma unroll
              for (int jj = 0; jj < kBatch; ++jj) {
                if (j0 + jj < kCellsPerThread) {
                  const uint32_t cw = cellw(j0 + jj);
                  if (cw != 0xFFFFFFFFu) st_shared_v4(st + (cw & 0xFFFFu), v[jj]);
                }
              }
            }
          }
          mbar_arrive(full_bar(s));   // (release: the stores above are visible to whoever sees the phase complete)
          continue;
        }
        if (p.debug_skip & 1) {
        } else if (p.up1 && from1) {
          // x1 is read through a nearest x2 upsample: halo pixel (ih, iw) of the conv's input grid is pixel
          // (ih >> 1, iw >> 1) of the low-res tensor [B, Hin/2, Win/2, C1] (decoder block without a phase form)
          const int Hlo = p.Hin >> 1, Wlo = p.Win >> 1;
          const __nv_bfloat16* lo = p.x1 + static_cast<long long>(b) * Hlo * Wlo * Cs + g * (NCH * 8) + c8;
#pragma unroll
          for (int j = 0; j < kCellsPerThread; ++j) {
            const uint32_t cw = cellw(j);
            if (cw != 0xFFFFFFFFu) {
              const int ih = ih0 + static_cast<int>((cw >> 16) & 0xFF), iw = iw0 + static_cast<int>(cw >> 24);
              const bool ok = static_cast<unsigned>(ih) < static_cast<unsigned>(p.Hin) &&
                              static_cast<unsigned>(iw) < static_cast<unsigned>(p.Win);
              const __nv_bfloat16* gp = ok ? lo + (static_cast<long long>(ih >> 1) * Wlo + (iw >> 1)) * Cs : p.x1;
              cp_async_16(st + (cw & 0xFFFFu), gp, ok ? 16u : 0u);
            }
          }
        } else if (interior) {
#pragma unroll
          for (int j = 0; j < kCellsPerThread; ++j) {
            const uint32_t cw = cellw(j);
            if (cw != 0xFFFFFFFFu) {
              const int hh = (cw >> 16) & 0xFF, k = cw >> 24;
              cp_async_16(st + (cw & 0xFFFFu), src + hh * row_elems + k * Cs, 16u);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < kCellsPerThread; ++j) {
            const uint32_t cw = cellw(j);
            if (cw != 0xFFFFFFFFu) {
              const int hh = (cw >> 16) & 0xFF, k = cw >> 24;
              const bool ok = static_cast<unsigned>(ih0 + hh) < static_cast<unsigned>(p.Hin) &&
                              static_cast<unsigned>(iw0 + k) < static_cast<unsigned>(p.Win);
              const __nv_bfloat16* gp = ok ? src + hh * row_elems + k * Cs : p.x1;
              cp_async_16(st + (cw & 0xFFFFu), gp, ok ? 16u : 0u);
            }
          }
        }
        // asynchronous arrival: the stage's full barrier completes when every producer thread's copies have
        // landed; nobody waits here, so up to S stages of loads are in flight (the consumer issues the
        // generic -> async proxy fence after its wait)
        cp_async_mbar_arrive_noinc(full_bar(s));
      }
    }
    }   // !TMAH
  } else if (SB && warp == kBWarp) {
    // ===================================================================== weight streamer (SB)
    // same (tile, group, tap) order as the MMA issuer; a stage is refilled as soon as its MMAs have completed
    if (elect_one()) {
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpacked);
      const int nb = p.nsteps / kBSteps;
      uint32_t bit = 0;
      for (int ti = sched0; ti < sched_end; ti += sched_step)
        for (int g = 0; g < groups; ++g)
          for (int bs = 0; bs < nb; ++bs, ++bit) {
            const int s = bit % kSbStages;
            mbar_wait_relaxed(bempty_bar(s), ((bit / kSbStages) & 1) ^ 1);
            mbar_expect_tx(bfull_bar(s), kBStage);
            const uint8_t* src = wsrc + static_cast<size_t>(g * p.nsteps + bs * kBSteps) * (2 * BN * 16);
            if (PAIR && p.wpacked_pair != nullptr) {
              // pair-ordered bank: this CTA's half of the stage is contiguous
              bulk_load_1d(w_addr + s * kBStage,
                           reinterpret_cast<const uint8_t*>(p.wpacked_pair) +
                               (static_cast<size_t>(g * nb + bs) * 2 + rank) * kBStage,
                           kBStage, bfull_bar(s));
            } else if (PAIR) {
              // this CTA's BN / 2 columns of every (step, chunk) slab: 2 * kBSteps pieces of BN / 2 * 16 bytes
#pragma unroll
              for (int sc = 0; sc < 2 * kBSteps; ++sc)
                bulk_load_1d(w_addr + s * kBStage + sc * (BNH * 16), src + (static_cast<size_t>(sc) * BN + rank * BNH) * 16, BNH * 16,
                             bfull_bar(s));
            } else {
              bulk_load_1d(w_addr + s * kBStage, src, kBStage, bfull_bar(s));
            }
          }
    }
  } else if (warp >= 4) {
    // ===================================================================== epilogue
    const int q = warp & 3;
    const int grp = warp >= 8 ? 1 : 0;   // second epilogue group (EPI == 2): odd blocks, direct stores only
    uint8_t* stg = smem + (bars - smem_base) + kBarBytes + q * kStgWarpBytes;
    const bool f32 = p.out_f32 != nullptr;
    const int elem = f32 ? 4 : 2;
    const size_t pixel_bytes = static_cast<size_t>(p.Cout) * elem;
    const size_t up_row_bytes = static_cast<size_t>(2 * p.Wout) * pixel_bytes;
    uint8_t* out_bytes = f32 ? reinterpret_cast<uint8_t*>(p.out_f32) : reinterpret_cast<uint8_t*>(p.out);
    // block rows: row r is pixel (r / 8, r % 8) of the 16 x 8 block
    const EpiLane L = make_epi_lane(q, lane, f32 ? EpiRun<BN>::GC_F32 * 4 : EpiRun<BN>::GC_BF16 * 2,
                                    p.up2_out ? 2 * p.Wout : p.Wout, (p.up2_out || PH) ? 2 : 1,
                                    [](int r, int& dh, int& dw) { dh = r >> 3; dw = r & 7; });
    if constexpr (POOL) {
      // ---- fused max-pool: a two-stage pipeline inside the epilogue. The two 128-row blocks of a tile are its EVEN and
      // its ODD columns (parity planes, halo_fill_steps_pool), so a thread of group 0 (warps 4-7, one per TMEM lane
      // quarter) owns the horizontally adjacent pixels (dh, 2d) and (dh, 2d + 1). It drains both -- bias, ReLU, one bf16
      // rounding --, stores the stem's own output from its registers (only inside the keep rectangle) and takes the
      // horizontal 3-maximum H(dh, d) = max(x[2d - 1], x[2d], x[2d + 1]) in registers: the pair's own maximum and the odd
      // pixel of the lane to its left (one shuffle per register; lane d = 0 takes the previous tile's last column from a
      // carry its own warp wrote). Only H goes to shared memory (16 x 8 cells, half the tile). Group 1 (warps 8-11) takes
      // the buffer and finishes the pool vertically: a thread loads nine rows of one H column for four outputs (row -1 =
      // the carry of the tile above). Against the first version (full tile to shared memory, 25 loads per thread for a
      // 2 x 2 block of outputs) this moves 36 KB instead of 91 KB per tile through a shared-memory pipe that ncu showed
      // saturated (LSU wavefronts 60 % + tensor-core operand reads 42 % of peak). Hand-over by named barriers:
      // 1 + k = "buffer k filled", 3 + k = "buffer k free".
      constexpr int kHCols = G::TW / 2;                                                   // H cells per tile row
      uint8_t* const pool_tiles = smem + (bars - smem_base) + kBarBytes;                 // 2 x [16 x 8 cells][kPoolPitch]
      uint8_t* const pool_rows = pool_tiles + 2 * 16 * kHCols * kPoolPitch;              // 2 x [Wout / 2 cells][128]: by tile-row parity
      uint8_t* const pool_cols = pool_rows + static_cast<size_t>(p.Wout) * 128;          // 2 x [16 px][128]: by tile parity
      const int et = (warp & 3) * 32 + lane;   // 0 .. 127 inside the group
      uint32_t tc = 0;
      auto hmax2 = [](uint32_t x, uint32_t y) -> uint32_t {
        const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&x), *reinterpret_cast<const __nv_bfloat162*>(&y));
        return *reinterpret_cast<const uint32_t*>(&r);
      };
      if (grp == 0) {
        const int dh = et >> 3, d = et & 7;   // TMEM lane = row dh, column pair d of the tile
        int kx0 = 0, ky0 = 0, kx1 = p.Wout, ky1 = p.Hout, keep_tb = -1;
        // (a CTA walks whole images in row-major tile order: the tile coordinates are counted, not divided out of the
        // tile index -- the three runtime divisions per tile were ~15 % of this group's instructions)
        int tw = 0, th = 0, tb = static_cast<int>(blockIdx.x);
        for (int ti = sched0; ti < sched_end; ti += sched_step, ++tc) {
          if (tc != 0 && ++tw == tiles_w) {
            tw = 0;
            if (++th == tiles_h) { th = 0; tb += static_cast<int>(gridDim.x); }
          }
          const int as = tc % NACC, k = tc & 1;
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * ACC;
          // In the exact-clipping loop only the part of the stem's output that dec3.conv1 (layer 6 of tile_need.cuh)
          // reads for this image's live outputs is stored: its needed region and the one-pixel halo of the 3x3 window
          // (the dead outputs of that conv's edge tiles read stale pixels; nothing live depends on them).
          // (recomputed when the image changes, i.e. once per 256 tiles: the table look-up and the walk through the
          // decoder's regions were a third of this group's time when done per tile)
          if (p.keep_tiles != nullptr && tb != keep_tb) {
            keep_tb = tb;
            const int* kt = p.keep_tiles + 6 * tb;
            const int x0 = __ldg(kt), y0 = __ldg(kt + 1);
            const NeedRect r = need_rect(p.keep_T, 6, __ldg(kt + 2) - x0, __ldg(kt + 3) - y0, __ldg(kt + 4) - x0, __ldg(kt + 5) - y0);
            const bool any = r.x1 > r.x0 && r.y1 > r.y0;
            kx0 = any ? r.x0 - 1 : 0; ky0 = any ? r.y0 - 1 : 0;
            kx1 = any ? r.x1 + 1 : 0; ky1 = any ? r.y1 + 1 : 0;
          }
          const int oy = th * kTH + dh, ox = tw * G::TW + 2 * d;
          const bool row_kept = oy >= ky0 && oy < ky1 && !(p.debug_skip & 4);
          const bool kept_e = row_kept && ox >= kx0 && ox < kx1, kept_o = row_kept && ox + 1 >= kx0 && ox + 1 < kx1;
          uint8_t* const gpx = reinterpret_cast<uint8_t*>(p.out) + ((static_cast<size_t>(tb) * p.Hout + oy) * p.Wout + ox) * 128;
          mbar_wait_relaxed(tfull_bar(as), (tc / NACC) & 1);
          tc_fence_after_sync();
          uint8_t* const hpx = pool_tiles + k * (16 * kHCols * kPoolPitch) + (dh * kHCols + d) * kPoolPitch;
          // last-column carry: written by the d = 7 lanes of tile tc, read by the d = 0 lanes (same warp) of tile tc + 1
          uint8_t* const carry_w = pool_cols + (tc & 1) * (16 * 128) + dh * 128;
          const uint8_t* const carry_r = pool_cols + ((tc & 1) ^ 1) * (16 * 128) + dh * 128;
          bool buf_free = tc < 2;
#pragma unroll
          for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t re[2][16], ro[2][16];   // even pixel = block 0, odd pixel = block 1
            tmem_ld_x16(taddr + c0, re[0]);
            tmem_ld_x16(taddr + c0 + 16, re[1]);
            tmem_ld_x16(taddr + BN + c0, ro[0]);
            tmem_ld_x16(taddr + BN + c0 + 16, ro[1]);
            if (!buf_free) {   // buffer k free (the loads are already in flight)
              asm volatile("bar.sync %0, 256;" ::"r"(3 + k) : "memory");
              buf_free = true;
            }
            tmem_ld_wait();
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {
              const int cb = c0 + 16 * hlf;   // first channel of this 16-channel piece
              uint32_t pe[8], po[8], h[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                // (bias from the kernel parameters = constant-bank operands, HaloArgs::bias_c)
                // ReLU on the packed pair: rounding to bf16 is monotonic and keeps zero, so max(round(v), 0) = round(max(v, 0))
                const float b0 = p.bias_c[(cb + 2 * i) & 63], b1 = p.bias_c[(cb + 2 * i + 1) & 63];
                const __nv_bfloat162 e2 = __floats2bfloat162_rn(__uint_as_float(re[hlf][2 * i]) + b0, __uint_as_float(re[hlf][2 * i + 1]) + b1);
                const __nv_bfloat162 o2 = __floats2bfloat162_rn(__uint_as_float(ro[hlf][2 * i]) + b0, __uint_as_float(ro[hlf][2 * i + 1]) + b1);
                pe[i] = hmax2(*reinterpret_cast<const uint32_t*>(&e2), 0u);
                po[i] = hmax2(*reinterpret_cast<const uint32_t*>(&o2), 0u);
              }
              if (kept_e) st_global_v8(gpx + cb * 2, pe);
              if (kept_o) st_global_v8(gpx + 128 + cb * 2, po);
              // the odd pixel to the left: lane - 1, or the carry of the previous tile (zero at the image's left edge:
              // the values are post-ReLU, so a zero never wins against the window's centre)
              uint32_t lf[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) lf[i] = __shfl_up_sync(0xffffffffu, po[i], 1);
              if (d == 0) {
                uint4 c0v = make_uint4(0u, 0u, 0u, 0u), c1v = c0v;
                if (tw > 0) {
                  c0v = *reinterpret_cast<const uint4*>(carry_r + cb * 2);
                  c1v = *reinterpret_cast<const uint4*>(carry_r + cb * 2 + 16);
                }
                lf[0] = c0v.x; lf[1] = c0v.y; lf[2] = c0v.z; lf[3] = c0v.w;
                lf[4] = c1v.x; lf[5] = c1v.y; lf[6] = c1v.z; lf[7] = c1v.w;
              }
              if (d == 7) {
                *reinterpret_cast<uint4*>(carry_w + cb * 2) = make_uint4(po[0], po[1], po[2], po[3]);
                *reinterpret_cast<uint4*>(carry_w + cb * 2 + 16) = make_uint4(po[4], po[5], po[6], po[7]);
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) h[i] = hmax2(hmax2(pe[i], po[i]), lf[i]);
              uint4* dst = reinterpret_cast<uint4*>(hpx + cb * 2);
              dst[0] = make_uint4(h[0], h[1], h[2], h[3]);
              dst[1] = make_uint4(h[4], h[5], h[6], h[7]);
            }
          }
          __syncwarp();   // the carry written above is read by this warp's d = 0 lanes in the next tile
          tc_fence_before_sync();
          mbar_arrive(tempty_bar(as));
          __threadfence_block();
          asm volatile("bar.arrive %0, 256;" ::"r"(1 + k) : "memory");
        }
      } else {
        const int vec = et & 7;                 // 16-byte piece (8 channels) of a cell
        const int j = (et >> 3) & 7;            // H column = pooled output column of the tile
        const int half = et >> 6;               // output rows 4 * half .. 4 * half + 3
        const int Hp = p.Hout >> 1, Wp = p.Wout >> 1;
        auto vmax = [&](uint4 a, const uint4 b) -> uint4 {
          return make_uint4(hmax2(a.x, b.x), hmax2(a.y, b.y), hmax2(a.z, b.z), hmax2(a.w, b.w));
        };
        int tw = 0, th = 0, tb = static_cast<int>(blockIdx.x);
        for (int ti = sched0; ti < sched_end; ti += sched_step, ++tc) {
          if (tc != 0 && ++tw == tiles_w) {
            tw = 0;
            if (++th == tiles_h) { th = 0; tb += static_cast<int>(gridDim.x); }
          }
          const int k = tc & 1;
          asm volatile("bar.sync %0, 256;" ::"r"(1 + k) : "memory");
          const uint8_t* const buf = pool_tiles + k * (16 * kHCols * kPoolPitch);
          const uint8_t* const rows_prev = pool_rows + static_cast<size_t>((th & 1) ^ 1) * Wp * 128;
          uint8_t* const rows_cur = pool_rows + static_cast<size_t>(th & 1) * Wp * 128;
          // H rows 8 * half - 1 .. 8 * half + 7 of column j; row -1 = last H row of the tile above (zero above the image).
          // Branch-free: every load comes from an address inside the pool buffers and is masked afterwards, so the nine
          // loads are in flight together.
          uint4 v[9];
#pragma unroll
          for (int qq = 0; qq < 9; ++qq) {
            const int r = 8 * half - 1 + qq;
            const uint8_t* a_tile = buf + ((r < 0 ? 0 : r) * kHCols + j) * kPoolPitch;
            const uint8_t* a_row = rows_prev + static_cast<size_t>(tw * kHCols + j) * 128;
            const uint4 x = *reinterpret_cast<const uint4*>((qq == 0 && half == 0 ? a_row : a_tile) + vec * 16);
            v[qq] = (qq == 0 && half == 0 && th == 0) ? make_uint4(0u, 0u, 0u, 0u) : x;
          }
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            const uint4 o = vmax(vmax(v[2 * a], v[2 * a + 1]), v[2 * a + 2]);
            const int pi = 4 * half + a;
            if (!(p.debug_skip & 4))
              *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(p.pool_out) +
                                        ((static_cast<size_t>(tb) * Hp + th * 8 + pi) * Wp + tw * 8 + j) * 128 + vec * 16) = o;
          }
          // carry: this tile's last H row, for the tile below
          if (half == 1) *reinterpret_cast<uint4*>(rows_cur + static_cast<size_t>(tw * kHCols + j) * 128 + vec * 16) = v[8];
          // hand the buffer back unless nobody will fill it again (no arrival may be left pending at exit)
          if (ti + 2 * sched_step < sched_end) {
            __threadfence_block();
            asm volatile("bar.arrive %0, 256;" ::"r"(3 + k) : "memory");
          }
        }
      }
    } else {
    uint32_t tcount = 0;
    // "accumulator drained": pairs count both CTAs' epilogue threads on the leader's barrier
    auto tempty_arrive = [&](uint32_t bar) {
      if (PAIR) mbar_arrive_leader(bar);
      else mbar_arrive(bar);
    };
    int tile_next = sched0 < sched_end ? tile_of(sched0) : 0;
    for (int ti = sched0; ti < sched_end; ti += sched_step, ++tcount) {
      const int tile = tile_next;
      if (ti + sched_step < sched_end) tile_next = tile_of(ti + sched_step);
      int tb, ty0, tx0;
      tile_pos(tile, tb, ty0, tx0);
      const int nblk = tile_narrow(tile) ? 1 : MB;   // blocks of this tile that were computed
      const int as = tcount % NACC;
      const uint32_t aph = (tcount / NACC) & 1;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * ACC;
      if (p.debug_skip & 8) {
        mbar_wait(tfull_bar(as), aph);
        tc_fence_after_sync();
        tc_fence_before_sync();
        tempty_arrive(tempty_bar(as));
        continue;
      }
      // direct mode: every lane stores its own pixel straight from registers, 32 bytes (one sector) per
      // instruction, instead of going through the shared-memory staging buffer
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
      if (D2S || SB || p.direct_store) {
        const int oh = ty0 + L.own_dh, ow = tx0 + L.own_dw;
        int sink_x0 = 0, sink_y0 = 0, sink_wx0 = 0, sink_wy0 = 0, sink_wx1 = 0, sink_wy1 = 0;
        float sink_best = 0.f, sink_den = 0.f;
        int sink_arg = 0;
        if (!PH && (BN == 16 || BN == 32 || D2S) && p.sink_cls != nullptr) {
          const int* st = p.sink_tiles + 6 * tb;
          sink_x0 = __ldg(st); sink_y0 = __ldg(st + 1); sink_wx0 = __ldg(st + 2); sink_wy0 = __ldg(st + 3);
          sink_wx1 = __ldg(st + 4); sink_wy1 = __ldg(st + 5);
        }
        // residual rows are loaded one block ahead: block 0 before the accumulator wait, block m + 1 while
        // block m is processed (BN <= 32: everything up front)
        constexpr int kResBuf = (PH || D2S || SB) ? 1 : (MB * BN <= 64 ? MB : 2);
        uint32_t rbuf[kResBuf][BN / 16][8];
        // (SB: 128 channels per pixel; the residual row is prefetched in 64-channel groups inside epilogue_tile instead)
        const bool has_res = !PH && !D2S && !SB && p.residual != nullptr;
        const __nv_bfloat16* res_row0 = p.residual + ((static_cast<long long>(tb) * p.Hout + oh) * p.Wout + ow) * BN;
        if (EPI == 2) {
          // this group's blocks: m = grp, grp + 2, ...; residual rows one block (of the group) ahead
          if (has_res) load_residual_row<BN>(res_row0 + 8 * grp * BN, rbuf[0]);
        } else if (has_res) {
          if (kResBuf == MB) {
#pragma unroll
            for (int m = 0; m < MB; ++m) load_residual_row<BN>(res_row0 + 8 * m * BN, rbuf[m % kResBuf]);
          } else {
            load_residual_row<BN>(res_row0, rbuf[0]);
          }
        }
        if (EPI == 2 && grp >= nblk) {
          // narrow tile: this group's block was not computed; keep the accumulator hand-shake in step
          mbar_wait(tfull_bar(as), aph);
          tc_fence_after_sync();
          tc_fence_before_sync();
          tempty_arrive(tempty_bar(as));
          continue;
        }
#pragma unroll
        for (int mi = 0; mi < MB / EPI; ++mi) {
          const int m = EPI == 2 ? 2 * mi + grp : mi;
          if (m >= nblk) continue;
          if (EPI == 2) {
            if (has_res && mi + 1 < MB / EPI) load_residual_row<BN>(res_row0 + 8 * (m + 2) * BN, rbuf[(mi + 1) % kResBuf]);
          } else if (has_res && kResBuf != MB && m + 1 < MB) {
            load_residual_row<BN>(res_row0 + 8 * (m + 1) * BN, rbuf[(m + 1) % kResBuf]);
          }
          if constexpr (D2S) {
            if (p.sink_cls != nullptr) {
              // Fused K6 of the depth-to-space head: the lane's cell = four pixels x 16 class columns. All 64 columns are
              // read first and the four soft-max maxima are straight-line code without branches in between, so their
              // dependency chains (compare trees, exponentials, sums) interleave: with two epilogue warps per scheduler
              // the per-pixel version (one chain at a time, a branch per pixel) ran at 0.4 instructions per cycle and
              // cost the head 180 of its 293 us per 148 tiles. The bias comes from the kernel arguments (constant bank),
              // with kSoftmaxMasked in the padded class columns, so the soft-max needs no per-class selects.
              mbar_wait_relaxed(tfull_bar(as), aph);
              tc_fence_after_sync();
              uint32_t acc[4][16];
#pragma unroll
              for (int g4 = 0; g4 < 4; ++g4) tmem_ld_x16(taddr + m * BN + 16 * g4, acc[g4]);
              tmem_ld_wait();
              if (!(p.debug_skip & 4)) {
                int arg4[4];
                float den4[4];
#pragma unroll
                for (int g4 = 0; g4 < 4; ++g4) {
                  float v[16], best;
#pragma unroll
                  for (int k = 0; k < 16; ++k) v[k] = __uint_as_float(acc[g4][k]) + p.bias_c[(16 * g4 + k) & 63];
                  softmax_max16_all(v, 0, best, arg4[g4], den4[g4]);
                }
#pragma unroll
                for (int g4 = 0; g4 < 4; ++g4) {
                  // pixel (2*oh + py, 2*(ow + 8m) + px) of image tb; only inside the write rectangle
                  const int rx = sink_x0 + 2 * (ow + 8 * m) + (g4 & 1), ry = sink_y0 + 2 * oh + (g4 >> 1);
                  if (rx >= sink_wx0 && rx < sink_wx1 && ry >= sink_wy0 && ry < sink_wy1) {
                    const long long o = (static_cast<long long>(ry) - p.sink_map_row0) * p.sink_map_w + rx;
                    p.sink_cls[o] = static_cast<uint8_t>(arg4[g4]);
                    if (p.sink_conf != nullptr) p.sink_conf[o] = static_cast<uint8_t>(1.f / den4[g4] + 0.5f);
                  }
                }
              }
              continue;
            }
          }
          long long dpix, own_pix = 0;
          if (PH) {
            dpix = (static_cast<long long>(tb) * p.Hout + 2 * oh + (m >> 1)) * p.Wout + 2 * ow + (m & 1);
          } else {
            own_pix = (static_cast<long long>(tb) * p.Hout + oh) * p.Wout + ow + 8 * m;
            dpix = p.up2_out ? (static_cast<long long>(tb) * 2 * p.Hout + 2 * oh) * (2 * p.Wout) + 2 * (ow + 8 * m) : own_pix;
          }
          uint8_t* own_dst = out_bytes + static_cast<size_t>(dpix) * pixel_bytes;
          auto direct = [&](int col0, const auto& regs) {
            if (p.debug_skip & 4) return;
            if constexpr (D2S) {
              // columns col0 .. col0 + 15 = the 16 channels of pixel (2*oh + py, 2*(ow + 8m) + px) of the cell
              const int grp = col0 >> 4;
              const int y = 2 * oh + (grp >> 1), x = 2 * (ow + 8 * m) + (grp & 1);
              store_regs(out_bytes + ((static_cast<size_t>(tb) * p.Hout + y) * p.Wout + x) * (16 * elem), regs);
              return;
            }
            if constexpr (!PH && (BN == 16 || BN == 32) && sizeof(regs) == 64) {
              if (p.sink_cls != nullptr) {
                // fused K6: the fp32 logits of pixel (oh, ow + 8m) of image tb arrive 16 at a time. Arg-max = first
                // maximum; max probability = 1 / sum(exp(v - max)). With 32 columns (17..32 classes) the first half
                // leaves its maximum, arg-max and exponent sum behind and the second half rescales that sum.
                // pixels of an active kernel tile that lie outside the write rectangle are not worth the exponentials
                const int rx = sink_x0 + ow + 8 * m, ry = sink_y0 + oh;
                if (!(rx >= sink_wx0 && rx < sink_wx1 && ry >= sink_wy0 && ry < sink_wy1)) return;
                float best, den;
                int arg;
                softmax_max16(regs, p.sink_ncls, col0, best, arg, den);
                if (BN == 32 && col0 != 0) {
                  // second half of a 17 .. 32-class head: join with the first half's maximum / arg-max / exponent sum
                  // (the earlier, lower class wins ties); both sums are rescaled to the joint maximum
                  const float joint = (col0 < p.sink_ncls && best > sink_best) ? best : sink_best;
                  const float den2 = col0 < p.sink_ncls ? den * __expf(best - joint) : 0.f;
                  if (!(col0 < p.sink_ncls && best > sink_best)) arg = sink_arg;
                  den = sink_den * __expf(sink_best - joint) + den2;
                  best = joint;
                }
                if (BN == 32 && col0 == 0) {
                  sink_best = best; sink_arg = arg; sink_den = den;
                  return;
                }
                const long long o = (static_cast<long long>(ry) - p.sink_map_row0) * p.sink_map_w + rx;
                p.sink_cls[o] = static_cast<uint8_t>(arg);
                if (p.sink_conf != nullptr) p.sink_conf[o] = static_cast<uint8_t>(1.f / den + 0.5f);
                return;
              }
            }
            uint8_t* d = own_dst + static_cast<size_t>(col0) * elem;
            store_regs(d, regs);
            if (!PH && p.up2_out) {
              store_regs(d + pixel_bytes, regs);
              store_regs(d + up_row_bytes, regs);
              store_regs(d + up_row_bytes + pixel_bytes, regs);
            }
          };
          if constexpr (D2S) {
            // no residual / row bias in these layers: constant-null members let the compiler drop those paths
            const D2SEpiArgs ea{p.relu, p.out_f32, p.Cout};
            // (bias from the kernel arguments when the caller put it there: constant-bank loads instead of shared-memory ones)
            if (p.bias_in_args)
              epilogue_tile<BN, true, true, true, true>(ea, p.bias_c, taddr + m * BN, tfull_bar(as), aph, lane, 0, stg, true, own_pix,
                                                        0, direct, rbuf[0]);
            else
              epilogue_tile<BN, true, true, true, true>(ea, bias_s, taddr + m * BN, tfull_bar(as), aph, lane, 0, stg, true, own_pix,
                                                        0, direct, rbuf[0]);
          } else if constexpr (SB) {
            epilogue_tile<BN, true, true, true, false>(p, bias_s, taddr + m * BN, tfull_bar(as), aph, lane, 0, stg, true, own_pix,
                                                       tb * p.Hout + oh, direct);
          } else {
            epilogue_tile<BN, true, true, true, true>(p, bias_s, taddr + m * BN, tfull_bar(as), aph, lane, 0, stg, true, own_pix,
                                                      tb * p.Hout + oh, direct, rbuf[(EPI == 2 ? mi : m) % kResBuf]);
          }
        }
        tc_fence_before_sync();
        tempty_arrive(tempty_bar(as));
        continue;
      }
      if constexpr (D2S || SB) continue;
      if (EPI == 2 && grp == 1) {
        // staged copy-out (2x2-replicated outputs) is done by the first group alone
        mbar_wait(tfull_bar(as), aph);
        tc_fence_after_sync();
        tc_fence_before_sync();
        tempty_arrive(tempty_bar(as));
        continue;
      }
      if (PH) {
        // accumulator m = phase (pa, pb): low-res pixel (h, w) of the block -> output (2h + pa, 2w + pb)
#pragma unroll
        for (int m = 0; m < MB; ++m) {
          const int pa = m >> 1, pb = m & 1;
          uint8_t* blk_dst = out_bytes + static_cast<size_t>((static_cast<long long>(tb) * p.Hout + 2 * ty0 + pa) * p.Wout +
                                                             2 * tx0 + pb) * pixel_bytes;
          auto copy = [&](auto run, int col0, int el) {
            if (p.debug_skip & 4) return;
            warp_copy_out_fast<decltype(run)::value>(stg, lane, L, blk_dst + static_cast<size_t>(col0) * el, pixel_bytes, 0, 0);
          };
          epilogue_tile<BN, true, true>(p, bias_s, taddr + m * BN, tfull_bar(as), aph, lane, 0, stg, true, 0, 0, copy);
        }
        tc_fence_before_sync();
        tempty_arrive(tempty_bar(as));
        continue;
      }
      const int oh = ty0 + L.own_dh;
      const long long pix0 = (static_cast<long long>(tb) * p.Hout + ty0) * p.Wout + tx0;
      const long long up0 = (static_cast<long long>(tb) * 2 * p.Hout + ty0 * 2) * (2 * p.Wout) + tx0 * 2;
      uint8_t* tile_dst = out_bytes + static_cast<size_t>(p.up2_out ? up0 : pix0) * pixel_bytes;
      const long long own_pix0 = (static_cast<long long>(tb) * p.Hout + oh) * p.Wout + tx0 + L.own_dw;
#pragma unroll
      for (int m = 0; m < MB; ++m) {  // block m = columns 8m..8m+7 of the tile
        if (m >= nblk) continue;
        uint8_t* blk_dst = tile_dst + static_cast<size_t>(p.up2_out ? 16 * m : 8 * m) * pixel_bytes;
        auto copy = [&](auto run, int col0, int el) {
          if (p.debug_skip & 4) return;
          warp_copy_out_fast<decltype(run)::value>(stg, lane, L, blk_dst + static_cast<size_t>(col0) * el, pixel_bytes,
                                                   p.up2_out, up_row_bytes);
        };
        epilogue_tile<BN, true, true>(p, bias_s, taddr + m * BN, tfull_bar(as), aph, lane, 0, stg, true, own_pix0 + 8 * m,
                                      tb * p.Hout + oh, copy);
      }
      tc_fence_before_sync();
      tempty_arrive(tempty_bar(as));
    }
    }   // !POOL
  } else if (elect_one()) {
    // ===================================================================== MMA issuer
    // (elect_one(), not lane == 0: each tcgen05.mma is then issued once from uniform registers instead of
    // inside a per-lane serialisation loop, which more than halves the issue interval, tests/umma_probe.cu)
    // Only the low descriptor word (start address, LBO) changes between instructions; the high word
    // (SBO, descriptor version 1, no swizzle) is a constant per operand.
    constexpr uint32_t idesc = umma_idesc_bf16(PAIR ? 256 : 128, BN);
    constexpr uint32_t a_hi = static_cast<uint32_t>(G::SBO16) | (1u << 14);
    constexpr uint32_t b_hi = 8u | (1u << 14);
    // filter slab of one K-step: two 8-channel chunks of BNH columns (LBO = BNH cells)
    const uint32_t b_lo0 = (w_addr >> 4) | (static_cast<uint32_t>(BNH) << 16);
    auto mma = [&](uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t acc) {
      if (PAIR) umma_bf16_lohi_2sm(d, a_lo, a_hi, b_lo, b_hi, idesc, acc);
      else umma_bf16_lohi(d, a_lo, a_hi, b_lo, b_hi, idesc, acc);
    };
    auto commit = [&](uint32_t bar) {
      if (PAIR) umma_commit_2sm(bar, 0b11);   // the barrier at this offset in both CTAs
      else umma_commit(bar);
    };
    uint32_t it = 0, tcount = 0, bit = 0;
    if (PAIR && !leader) {
      // peer of a pair: forward "landed" of every halo stage (and weight stage) to the leader, in the order the
      // leader's MMAs consume them; the fence makes this CTA's cp.async writes visible to the async proxy first
      const int nb = SB ? p.nsteps / kBSteps : 0;
      for (int ti = sched0; ti < sched_end; ti += sched_step)
        for (int g = 0; g < groups; ++g, ++it) {
          const int s = it % S;
          mbar_wait(full_bar(s), (it / S) & 1);
          fence_proxy_async_smem();
          mbar_arrive_leader(pfull_bar(s));
          for (int bs = 0; bs < nb; ++bs, ++bit) {
            const int sb = bit % kSbStages;
            mbar_wait(bfull_bar(sb), (bit / kSbStages) & 1);
            mbar_arrive_leader(pbfull_bar(sb));
          }
        }
    } else {
    const bool listed = !PAIR && !POOL && !PH && p.tile_packed && p.tile_list != nullptr;
    int mtile_next = (listed && sched0 < sched_end) ? tile_of(sched0) : 0;
    for (int ti = sched0; ti < sched_end; ti += sched_step, ++tcount) {
      const int as = tcount % NACC;
      const uint32_t aph = (tcount / NACC) & 1;
      // blocks of this tile: a narrow tile of an origin-shifted list has one (the entry is read one tile ahead)
      const int nblk = (listed && tile_narrow(mtile_next)) ? 1 : MB;
      if (listed && ti + sched_step < sched_end) mtile_next = tile_of(ti + sched_step);
      mbar_wait(tempty_bar(as), aph ^ 1);
      tc_fence_after_sync();
      const uint32_t d_tmem = tmem_base + as * ACC;
      for (int g = 0; g < groups; ++g, ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        mbar_wait(full_bar(s), ph);
        if (PAIR) mbar_wait(pfull_bar(s), ph);
        fence_proxy_async_smem();   // the stage was written by cp.async (generic proxy), tcgen05.mma reads it through the async proxy
        tc_fence_after_sync();
        const uint32_t st16 = (stage_addr0 + s * G::STAGE) >> 4;
        // K-step outer, block inner: consecutive MMAs hit different accumulators (block m = output columns
        // 8m..8m+7 of the tile, 8 cells further in the stage) and share the step's filter slab
        uint32_t b_lo = b_lo0 + static_cast<uint32_t>(g * p.nsteps) * (2 * BNH);
        if constexpr (SB) {
          // one weight stage per filter tap: wait for it, issue its kBSteps x MB MMAs, hand it back
          const int nb = p.nsteps / kBSteps;
          for (int bs = 0; bs < nb; ++bs, ++bit) {
            const int sb = bit % kSbStages;
            mbar_wait(bfull_bar(sb), (bit / kSbStages) & 1);
            if (PAIR) mbar_wait(pbfull_bar(sb), (bit / kSbStages) & 1);
            tc_fence_after_sync();
            const uint32_t wb = ((w_addr + sb * kBStage) >> 4) | (static_cast<uint32_t>(BNH) << 16);
            if (!(p.debug_skip & 2)) {
#pragma unroll
              for (int kk = 0; kk < kBSteps; ++kk) {
                const int k = bs * kBSteps + kk;
                const uint32_t a_lo = p.a_lo[k] + st16;
                const uint32_t acc = (g | k) != 0 ? 1u : 0u;
#pragma unroll
                for (int m = 0; m < MB; ++m)
                  if (m < nblk) mma(d_tmem + m * BN, a_lo + 8 * m, wb + kk * (2 * BNH), acc);
              }
            }
            commit(bempty_bar(sb));
          }
        } else if (p.debug_skip & 2) {
        } else if (PH) {
          // step k of phase m: a_lo[m * nsteps + k], filter slab (m * nsteps + k); k outer so that consecutive
          // MMAs alternate between the four phase accumulators
#pragma unroll 2
          for (int k = 0; k < p.nsteps; ++k) {
            const uint32_t acc = k != 0 ? 1u : 0u;
#pragma unroll
            for (int m = 0; m < MB; ++m)
              umma_bf16_lohi(d_tmem + m * BN, p.a_lo[m * p.nsteps + k] + st16, a_hi,
                             b_lo + static_cast<uint32_t>(m * p.nsteps + k) * (2 * BN), b_hi, idesc, acc);
          }
        } else {
#pragma unroll 2
          for (int k = 0; k < p.nsteps; ++k) {
            const uint32_t a_lo = p.a_lo[k] + st16;
            const uint32_t acc = (g | k) != 0 ? 1u : 0u;
#pragma unroll
            for (int m = 0; m < MB; ++m) {
              // POOL: block m = the tile's columns of parity m, with a step table of its own (halo_fill_steps_pool)
              if (POOL) mma(d_tmem + m * BN, p.a_lo[m * p.nsteps + k] + st16, b_lo, acc);
              else if (m < nblk) mma(d_tmem + m * BN, a_lo + 8 * m, b_lo, acc);
            }
            b_lo += 2 * BNH;
          }
        }
        commit(empty_bar(s));
      }
      commit(tfull_bar(as));
    }
    }
  }

  tc_fence_before_sync();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // the leader's MMAs read the peer's shared memory, the peer arrives on the leader's barriers
  if (warp == kMmaWarp) {
    tc_fence_after_sync();
    if (PAIR) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int KH, int STRIDE, int NCH, int BN, int MB, bool PH = false, int EPI = 1, bool D2S = false, bool SB = false,
          bool PAIR = false, bool POOL = false, bool TMAH = false>
int launch_halo_t(const HaloArgs& a, int num_sms, cudaStream_t stream) {
  constexpr int kThreadsK = kThreads + 128 * (EPI - 1);
  using G = Geo<KH, STRIDE, NCH, (PH ? 1 : MB), BN, EPI, PAIR, POOL>;
  const int groups = a.groups1 + a.groups2;
  constexpr int BNH = PAIR ? BN / 2 : BN;
  const int wbytes = SB ? kSbStages * (NCH / 2) * 2 * BNH * 16 : (PH ? MB : 1) * groups * a.nsteps * 2 * BNH * 16;
  // (the depth-to-space and streamed-weight kernels only store from registers: no copy-out staging)
  const int smem = ((wbytes + 127) / 128) * 128 + (BN * 4 <= 256 ? 256 : BN * 4) + G::STAGES * G::STAGE +
                   ((2 * G::STAGES + 2 * kAccDeep + (SB ? 2 * kSbStages : 0) + (PAIR ? G::STAGES + (SB ? kSbStages : 0) : 0)) * 8 + 16 + 127) / 128 * 128 +
                   ((D2S || SB || POOL || PAIR) ? 0 : 4 * kStgWarpBytes) +
                   (POOL ? 2 * 16 * (G::TW / 2) * kPoolPitch + a.Wout * 128 + 2 * 16 * 128 : 0);
  HaloArgs pool_args;
  const HaloArgs* argp = &a;
  if (POOL) {
    // step tables of the fused-pool stem: block m = the tile's columns of parity m. Output column 2d + m reads halo
    // column 2d + m + kw = cell d + (m + kw) / 2 of parity plane (m + kw) % 2; the second 8-channel chunk is NP planes on
    pool_args = a;
    for (int m = 0; m < 2; ++m)
      for (int tap = 0; tap < KH * KH; ++tap) {
        const int kh = tap / KH, kw = tap % KH;
        const uint32_t off = static_cast<uint32_t>(((m + kw) & 1) * G::PLANE16 + kh * G::PW + ((m + kw) >> 1));
        pool_args.a_lo[m * (KH * KH) + tap] = off | (static_cast<uint32_t>(G::NP * G::PLANE16) << 16);
      }
    if (a.nsteps != KH * KH || NCH != 2) return -3008;
    argp = &pool_args;
  }
  alignas(64) CUtensorMap tm1, tm2;
  memset(&tm1, 0, sizeof tm1);
  memset(&tm2, 0, sizeof tm2);
  if (TMAH) {
    // a source [B, Hin, Win, C] bf16 seen as (C, W, H, B); one box = the halo plane of an 8-channel chunk
    for (int src = 0; src < (a.C2 > 0 ? 2 : 1); ++src) {
      const unsigned long long C = static_cast<unsigned long long>(src ? a.C2 : a.C1);
      const unsigned long long dims[4] = {C, static_cast<unsigned long long>(a.Win), static_cast<unsigned long long>(a.Hin),
                                          static_cast<unsigned long long>(a.B)};
      const unsigned long long strides[3] = {C * 2, static_cast<unsigned long long>(a.Win) * C * 2,
                                             static_cast<unsigned long long>(a.Hin) * a.Win * C * 2};
      // stride 2: a plane holds every second pixel of a row (w-parity planes); rows are staged in full
      const unsigned box[4] = {8u, static_cast<unsigned>(G::NP * G::PW), static_cast<unsigned>(G::PH), 1u};
      const unsigned es[4] = {1u, static_cast<unsigned>(G::NP), 1u, 1u};
      const int rc = encode_tma_plain_bf16(src ? &tm2 : &tm1, src ? a.x2 : a.x1, 4, dims, strides, box, es);
      if (rc) return rc;
    }
  }
  static int configured = 0;
  static int occ = 1;
  if (configured < smem) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<KH, STRIDE, NCH, BN, MB, PH, EPI, D2S, SB, PAIR, POOL, TMAH>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return static_cast<int>(e);
    // ask for the largest shared-memory carve-out so that two CTAs of the small configurations fit
    cudaFuncSetAttribute(conv_halo_kernel<KH, STRIDE, NCH, BN, MB, PH, EPI, D2S, SB, PAIR, POOL, TMAH>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    configured = smem;
    int nb = 1;
    cudaError_t qe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, conv_halo_kernel<KH, STRIDE, NCH, BN, MB, PH, EPI, D2S, SB, PAIR, POOL, TMAH>, kThreadsK, smem);
    if (getenv("FB_DEBUG")) fprintf(stderr, "[halo occupancy query] err=%d blocks/SM=%d\n", static_cast<int>(qe), nb);
    // CTAs are independent (static tile schedule, private TMEM columns <= 256): over-subscribing is safe,
    // so size the grid for the intended co-residency and let the hardware place what fits.
    occ = G::OCC;
  }
  if (PAIR) {
    // a.num_m_tiles counts pairs; clusters of two CTAs, one pair of SMs (a TPC) each
    const int clusters = a.num_m_tiles < num_sms / 2 ? a.num_m_tiles : num_sms / 2;
    if (clusters <= 0) return 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kThreadsK);
    cfg.dynamicSmemBytes = static_cast<size_t>(smem);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (getenv("FB_DEBUG")) fprintf(stderr, "[halo pair %d,%d] smem=%d clusters=%d pairs=%d\n", NCH, BN, smem, clusters, a.num_m_tiles);
    const cudaError_t le = cudaLaunchKernelEx(&cfg, conv_halo_kernel<KH, STRIDE, NCH, BN, MB, PH, EPI, D2S, SB, PAIR, POOL, TMAH>, *argp, tm1, tm2);
    return static_cast<int>(le != cudaSuccess ? le : cudaGetLastError());
  }
  const int cap = num_sms * occ;
  const int grid = POOL ? (a.B < cap ? a.B : cap) : (a.num_m_tiles < cap ? a.num_m_tiles : cap);   // POOL: whole images per CTA
  if (grid <= 0) return 0;
  if (getenv("FB_DEBUG")) fprintf(stderr, "[halo %d,%d,%d,%d,%d] smem=%d occ=%d grid=%d tiles=%d\n", KH, STRIDE, NCH, BN, MB, smem, occ, grid, a.num_m_tiles);
  static const bool pdl = !(getenv("FB_NO_PDL") && getenv("FB_NO_PDL")[0] == '1');
  const cudaError_t le = launch_kernel_pdl(conv_halo_kernel<KH, STRIDE, NCH, BN, MB, PH, EPI, D2S, SB, PAIR, POOL, TMAH>, dim3(grid), dim3(kThreadsK),
                                           static_cast<size_t>(smem), stream, pdl, *argp, tm1, tm2);
  return static_cast<int>(le != cudaSuccess ? le : cudaGetLastError());
}

}  // namespace

bool halo_pool_fusable() {
  return !(getenv("FB_TMAH") && atoi(getenv("FB_TMAH")) <= 0);
}

int halo_blocks(int KH, int nch, int bn) {
  if (KH == 7 || KH == 4) return 2;
  if (nch == 2) return 4;               // 16 -> 16 @ 512^2: 16 x 32 pixel tiles
  if (nch == 4 && bn == 16) return 2;   // 32 -> 16 @ 512^2 (two CTAs per SM)
  return 2;
}

HaloGeom halo_geom(int KH, int stride, int nch, int bn) {
  HaloGeom g;
  g.KH = KH; g.stride = stride; g.nch = nch; g.bn = bn;
  g.mb = halo_blocks(KH, nch, bn);
  const int kTW = 8 * g.mb;
  g.pad = KH / 2;
  g.np = stride;
  g.ph = stride * (kTH - 1) + KH + (nch == 1 ? 1 : 0);
  const int spanw = stride * (kTW - 1) + KH;
  g.pw = (spanw + g.np - 1) / g.np;
  g.kw_cells = g.np * g.pw;
  g.plane16 = (g.ph * g.pw + 7) / 8 * 8;
  g.stage_bytes = ((nch * g.np * g.plane16 * 16) + 127) / 128 * 128;
  g.nsteps = (nch == 1) ? KH * ((KH + 1) / 2) : KH * KH * (nch / 2);
  return g;
}

int halo_group_channels(int KH, int C1, int C2) {
  if (KH == 7) return 8;
  int cg = C1 < 64 ? C1 : 64;
  return cg;
}

bool halo_supported(int KH, int stride, int C1, int C2, int Cout, int Hout, int Wout) {
  if (Hout % kTH != 0) return false;
  if (KH == 7) return stride == 2 && C1 == 8 && C2 == 0 && Cout == 64 && Wout % (8 * halo_blocks(7, 1, 64)) == 0;
  // the stem in space-to-depth form: 4x4 stride 1 on 16 channels (taps at rows oh-2 .. oh+1)
  if (KH == 4) return stride == 1 && C1 == 16 && C2 == 0 && Cout == 64 && Wout % (8 * halo_blocks(4, 2, 64)) == 0;
  if (KH != 3 || stride != 1) return false;
  const int cg = halo_group_channels(KH, C1, C2);
  if (cg < 16 || C1 % cg != 0 || C2 % cg != 0) return false;
  const int nch = cg / 8;
  if ((C1 + C2) / cg > 2) return false;
  if (Wout % (8 * halo_blocks(KH, nch, Cout)) != 0) return false;
  if (nch == 8 && Cout == 128) return C2 == 0 && C1 == 128;   // streamed filter bank (layer2, dec1.conv2)
  return (nch == 2 && (Cout == 16 || Cout == 32)) || (nch == 4 && (Cout == 16 || Cout == 32)) ||
         (nch == 8 && (Cout == 32 || Cout == 64));
}

void halo_fill_steps(HaloArgs& a, int KH, int stride) {
  const int cg = halo_group_channels(KH, a.C1, a.C2);
  const int nch = cg / 8;
  const HaloGeom g = halo_geom(KH, stride, nch, a.Cout);
  a.groups1 = a.C1 / cg;
  a.groups2 = a.C2 / cg;
  a.nsteps = g.nsteps;
  memset(a.a_lo, 0, sizeof a.a_lo);
  if (nch == 1) {
    // stem: one 8-channel chunk per pixel; a K=16 step pairs filter rows (2*pr, 2*pr+1) of column kw
    for (int kw = 0; kw < KH; ++kw)
      for (int pr = 0; pr < (KH + 1) / 2; ++pr) {
        const int s = kw * ((KH + 1) / 2) + pr;
        const int par = kw % g.np;
        const uint32_t off = static_cast<uint32_t>(par * g.plane16 + 2 * pr * g.pw + kw / g.np);
        const uint32_t lbo = static_cast<uint32_t>(g.pw);
        a.a_lo[s] = off | (lbo << 16);
      }
  } else {
    for (int tap = 0; tap < KH * KH; ++tap)
      for (int kk = 0; kk < nch / 2; ++kk) {
        const int s = tap * (nch / 2) + kk;
        const int kh = tap / KH, kw = tap % KH;
        const int par = kw % g.np;
        const uint32_t off = static_cast<uint32_t>(((2 * kk) * g.np + par) * g.plane16 + kh * g.pw + kw / g.np);
        const uint32_t lbo = static_cast<uint32_t>(g.np * g.plane16);
        a.a_lo[s] = off | (lbo << 16);
      }
  }
  a.num_m_tiles = a.B * (a.Hout / kTH) * (a.Wout / (8 * g.mb));
  const char* np = getenv("FB_PREFETCH");
  a.no_prefetch = !(np && np[0] == '1');
  const char* sk = getenv("FB_HALO_SKIP");
  a.debug_skip = sk ? atoi(sk) : 0;
  const char* ds = getenv("FB_DIRECT_STORE");
  // 2x2-replicated outputs keep the staged copy-out (four scattered 32-byte stores per lane are slower)
  a.direct_store = ds ? atoi(ds) : (a.up2_out ? 0 : 1);
}

static uint16_t bf16_rne(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u && (u & 0x007FFFFFu)) return static_cast<uint16_t>((u >> 16) | 0x40);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}

size_t pack_halo_weights(const float* w, int Cout, int CoutPad, int Cin, int CinPad, int KH, int stride,
                         int C1pad, int C2pad, uint16_t* dst) {
  const int cg = halo_group_channels(KH, C1pad, C2pad);
  const int nch = cg / 8;
  const HaloGeom g = halo_geom(KH, stride, nch, CoutPad);
  const int groups = CinPad / cg;
  const size_t total = static_cast<size_t>(groups) * g.nsteps * 2 * CoutPad * 8;
  if (!dst) return total;
  memset(dst, 0, total * 2);
  auto W = [&](int o, int ci, int kh, int kw) -> float {
    if (o >= Cout || ci >= Cin || kh >= KH || kw >= KH) return 0.f;
    return w[((static_cast<size_t>(o) * Cin + ci) * KH + kh) * KH + kw];
  };
  for (int grp = 0; grp < groups; ++grp)
    for (int s = 0; s < g.nsteps; ++s)
      for (int j = 0; j < 2; ++j)
        for (int n = 0; n < CoutPad; ++n)
          for (int e = 0; e < 8; ++e) {
            int ci, kh, kw;
            if (nch == 1) {
              kw = s / ((KH + 1) / 2);
              kh = 2 * (s % ((KH + 1) / 2)) + j;
              ci = e;
            } else {
              const int tap = s / (nch / 2), kk = s % (nch / 2);
              kh = tap / KH; kw = tap % KH;
              ci = grp * cg + (2 * kk + j) * 8 + e;
            }
            dst[(((static_cast<size_t>(grp) * g.nsteps + s) * 2 + j) * CoutPad + n) * 8 + e] = bf16_rne(W(n, ci, kh, kw));
          }
  return total;
}

void pack_halo_weights_pair128(const uint16_t* packed, size_t total, uint16_t* dst) {
  // packed: [step (all groups)][chunk j][n < 128][8]; a stage = 4 consecutive steps (one filter tap of a 64-channel group)
  const size_t step_elems = 2 * 128 * 8, stage_elems = 4 * step_elems;
  for (size_t st = 0; st * stage_elems < total; ++st)
    for (int r = 0; r < 2; ++r)
      for (int kk = 0; kk < 4; ++kk)
        for (int j = 0; j < 2; ++j)
          for (int n = 0; n < 64; ++n)
            memcpy(dst + st * stage_elems + ((static_cast<size_t>(r) * 4 + kk) * 2 + j) * 64 * 8 + n * 8,
                   packed + st * stage_elems + (static_cast<size_t>(kk) * 2 + j) * 128 * 8 + (r * 64 + n) * 8, 8 * sizeof(uint16_t));
}

// ---- sub-pixel phase form (32 -> 16 channels): tile = 16 x 8 low-res pixels, halo 18 x 10 cells per chunk
namespace {
constexpr int kPhNch = 4, kPhPw = 10, kPhPlane16 = (18 * kPhPw + 7) / 8 * 8;   // = Geo<3, 1, 4, 1, 16>::PLANE16
// original filter taps that read low-res tap d (0/1) for output parity `parity`
int phase_taps(int parity, int d, int* out) {
  if (parity == 0) { if (d == 0) { out[0] = 0; return 1; } out[0] = 1; out[1] = 2; return 2; }
  if (d == 0) { out[0] = 0; out[1] = 1; return 2; }
  out[0] = 2; return 1;
}
}  // namespace

bool halo_phase_supported(int C1, int C2, int Cout, int Hlo, int Wlo) {
  return C1 == 8 * kPhNch && C2 == 0 && Cout == 16 && Hlo % kTH == 0 && Wlo % 8 == 0;
}

void halo_fill_steps_phase(HaloArgs& a) {
  a.groups1 = 1;
  a.groups2 = 0;
  a.nsteps = 4 * (kPhNch / 2);
  memset(a.a_lo, 0, sizeof a.a_lo);
  for (int pa = 0; pa < 2; ++pa)
    for (int pb = 0; pb < 2; ++pb)
      for (int di = 0; di < 2; ++di)
        for (int dj = 0; dj < 2; ++dj)
          for (int kk = 0; kk < kPhNch / 2; ++kk) {
            // output (2h+pa, 2w+pb), low-res tap (di, dj) -> halo cell (h + di + pa, w + dj + pb); halo origin (-1, -1)
            const uint32_t off = static_cast<uint32_t>((2 * kk) * kPhPlane16 + (di + pa) * kPhPw + (dj + pb));
            a.a_lo[(pa * 2 + pb) * a.nsteps + (di * 2 + dj) * (kPhNch / 2) + kk] = off | (static_cast<uint32_t>(kPhPlane16) << 16);
          }
  a.num_m_tiles = a.B * (a.Hin / kTH) * (a.Win / 8);
  const char* np = getenv("FB_PREFETCH");
  a.no_prefetch = !(np && np[0] == '1');
  const char* sk = getenv("FB_HALO_SKIP");
  a.debug_skip = sk ? atoi(sk) : 0;
  const char* ds = getenv("FB_DIRECT_STORE");
  a.direct_store = ds ? atoi(ds) : 1;
}

size_t pack_halo_weights_phase(const float* w, int Cout, int CoutPad, int Cin, int CinPad, uint16_t* dst) {
  const int nsteps = 4 * (CinPad / 16);
  const size_t total = static_cast<size_t>(4) * nsteps * 2 * CoutPad * 8;
  if (!dst) return total;
  memset(dst, 0, total * 2);
  for (int pa = 0; pa < 2; ++pa)
    for (int pb = 0; pb < 2; ++pb)
      for (int di = 0; di < 2; ++di)
        for (int dj = 0; dj < 2; ++dj) {
          int khs[2], kws[2];
          const int nh = phase_taps(pa, di, khs), nw = phase_taps(pb, dj, kws);
          for (int kk = 0; kk < CinPad / 16; ++kk)
            for (int j = 0; j < 2; ++j)
              for (int n = 0; n < Cout; ++n)
                for (int e = 0; e < 8; ++e) {
                  const int ci = (2 * kk + j) * 8 + e;
                  if (ci >= Cin) continue;
                  double s = 0.0;
                  for (int i = 0; i < nh; ++i)
                    for (int q = 0; q < nw; ++q) s += w[((static_cast<size_t>(n) * Cin + ci) * 3 + khs[i]) * 3 + kws[q]];
                  const size_t step = static_cast<size_t>(pa * 2 + pb) * nsteps + (di * 2 + dj) * (CinPad / 16) + kk;
                  dst[((step * 2 + j) * CoutPad + n) * 8 + e] = bf16_rne(static_cast<float>(s));
                }
        }
  return total;
}

// ---- depth-to-space forms of the 16-output-channel layers (HaloArgs::d2s)
bool halo_d2s_supported(int mode, int C1, int C2, int Cout, int Hout, int Wout) {
  if (C2 != 0 || Cout > 16 || Hout % (2 * kTH) != 0 || Wout % 32 != 0) return false;   // tiles of 16 x 16 cells
  return mode == 1 ? C1 == 16 : mode == 2 ? C1 == 32 : false;
}

void halo_fill_steps_d2s(HaloArgs& a, int mode) {
  a.d2s = mode;
  a.Cout = 64;
  if (mode == 1) halo_fill_steps(a, 4, 2);   // 16 taps x one K = 16 step on the pixel grid, two w-parity planes
  else halo_fill_steps(a, 3, 1);             // 9 taps x two K = 16 steps on the low-res grid
  a.num_m_tiles = a.B * (a.Hout / (2 * kTH)) * (a.Wout / 32);
  a.direct_store = 1;
}

size_t pack_halo_weights_d2s(int mode, const float* w, int Cout, int Cin, uint16_t* dst) {
  const int KH = mode == 1 ? 4 : 3, stride = mode == 1 ? 2 : 1;
  if (!dst) return pack_halo_weights(nullptr, 64, 64, Cin, Cin, KH, stride, Cin, 0, nullptr);
  // w64[(py*2 + px)*16 + co][ci][a][b]: the weight output pixel (py, px) of the cell applies to window position (a, b)
  std::vector<float> w64(static_cast<size_t>(64) * Cin * KH * KH, 0.f);
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px)
      for (int co = 0; co < Cout; ++co)
        for (int ci = 0; ci < Cin; ++ci)
          for (int a = 0; a < KH; ++a)
            for (int b = 0; b < KH; ++b) {
              double s = 0.0;
              if (mode == 1) {
                // window row a is pixel row 2Y - 1 + a; output row 2Y + py reads rows 2Y + py - 1 + kh: kh = a - py
                const int kh = a - py, kw = b - px;
                if (kh < 0 || kh > 2 || kw < 0 || kw > 2) continue;
                s = w[((static_cast<size_t>(co) * Cin + ci) * 3 + kh) * 3 + kw];
              } else {
                // window row a is low-res row Y - 1 + a; output row 2Y + py reads upsampled rows 2Y + d, d = py - 1 + kh
                // in -1 .. 2, i.e. low-res row Y + floor(d / 2): a = floor(d / 2) + 1 = (d + 2) >> 1
                for (int kh = 0; kh < 3; ++kh)
                  for (int kw = 0; kw < 3; ++kw)
                    if (((py + 1 + kh) >> 1) == a && ((px + 1 + kw) >> 1) == b)
                      s += w[((static_cast<size_t>(co) * Cin + ci) * 3 + kh) * 3 + kw];
              }
              w64[((static_cast<size_t>((py * 2 + px) * 16 + co) * Cin + ci) * KH + a) * KH + b] = static_cast<float>(s);
            }
  return pack_halo_weights(w64.data(), 64, 64, Cin, Cin, KH, stride, Cin, 0, dst);
}

int launch_conv_halo(const HaloArgs& a, int KH, int stride, int num_sms, cudaStream_t stream) {
  // TMA-staged halo (the input planes written by tensor loads from one thread instead of one cp.async per 16-byte cell):
  // FB_TMAH=0 never, 1 (default) wherever it measured faster, 2 also the streamed-weight 128-channel form
  static const int tmah_mode = getenv("FB_TMAH") ? atoi(getenv("FB_TMAH")) : 1;
  const bool tma_ok = tmah_mode > 0 && !a.up1;   // (a source read through the x2 up-sampling gather cannot be a TMA box)
  static const bool one_cta = !(getenv("FB_ONE_CTA") && getenv("FB_ONE_CTA")[0] == '0');
  if (a.d2s) {
    if (!halo_d2s_supported(a.d2s, a.C1, a.C2, 16, a.Hout, a.Wout) || a.Cout != 64 || a.residual || a.rowbias || a.up2_out ||
        a.up1 || a.phase_mode || (a.d2s == 1 ? (a.Hin != a.Hout || a.Win != a.Wout) : (2 * a.Hin != a.Hout || 2 * a.Win != a.Wout)))
      return -3005;
    // (the sink's class mask lives in the bias vector, which it takes from the kernel arguments)
    if (a.sink_cls != nullptr && (a.relu || a.sink_ncls < 1 || a.sink_ncls > 16 || !a.bias_in_args)) return -3006;
    if (a.nsteps != (a.d2s == 1 ? 16 : 18)) return -3002;
    // With TMA staging one thread feeds a CTA, so the 16 / 32-channel layers no longer need two CTAs per SM for
    // their copy threads: one CTA with two epilogue groups and a ring twice as deep measured 350 -> 300 us for the
    // head (FB_ONE_CTA=0: two CTAs per SM, one epilogue group each).
    if (tma_ok && one_cta)
      return a.d2s == 1 ? launch_halo_t<4, 2, 2, 64, 2, false, 2, true, false, false, false, true>(a, num_sms, stream)
                        : launch_halo_t<3, 1, 4, 64, 2, false, 2, true, false, false, false, true>(a, num_sms, stream);
    if (tma_ok)
      return a.d2s == 1 ? launch_halo_t<4, 2, 2, 64, 2, false, 1, true, false, false, false, true>(a, num_sms, stream)
                        : launch_halo_t<3, 1, 4, 64, 2, false, 1, true, false, false, false, true>(a, num_sms, stream);
    return a.d2s == 1 ? launch_halo_t<4, 2, 2, 64, 2, false, 1, true>(a, num_sms, stream)
                      : launch_halo_t<3, 1, 4, 64, 2, false, 1, true>(a, num_sms, stream);
  }
  if (a.phase_mode) {
    if (KH != 3 || stride != 1 || !halo_phase_supported(a.C1, a.C2, a.Cout, a.Hin, a.Win) || a.Hout != 2 * a.Hin ||
        a.Wout != 2 * a.Win || a.residual || a.rowbias || a.up2_out || a.out_f32)
      return -3004;
    if (a.nsteps != 4 * (kPhNch / 2)) return -3002;
    return launch_halo_t<3, 1, kPhNch, 16, 4, true>(a, num_sms, stream);
  }
  const int cg = halo_group_channels(KH, a.C1, a.C2);
  const int nch = cg / 8;
  if (!halo_supported(KH, stride, a.C1, a.C2, a.Cout, a.Hout, a.Wout)) return -3001;
  if (a.nsteps <= 0 || a.nsteps > kHaloMaxSteps) return -3002;
  // two epilogue groups for the one-CTA-per-SM configurations (FB_EPI2=0: one group, for A/B runs)
  const char* e2 = getenv("FB_EPI2");
  const bool epi2 = !(e2 && e2[0] == '0') && a.direct_store;
  // CTA pairs (cta_group::2) for the 64- and 128-channel layers that run whole images (no active-tile list)
  const bool pair = a.pair && KH == 3 && nch == 8 && (a.Cout == 64 || a.Cout == 128) && a.tile_list == nullptr && !a.up2_out &&
                    !a.out_f32 && a.Hout % (2 * kTH) == 0 && epi2 && !a.up1;
  // (64-channel layers: 313 -> 216 us per launch with TMA staging; the streamed-weight 128-channel form 141 -> 161 us)
  const bool tmah = tma_ok && KH == 3 && nch == 8 && epi2 && a.C2 == 0 && (a.Cout == 64 || (a.Cout == 128 && tmah_mode >= 2));
  if (pair) {
    HaloArgs b = a;
    b.num_m_tiles = a.B * (a.Hout / (2 * kTH)) * (a.Wout / (8 * halo_blocks(3, 8, a.Cout)));
    if (tmah)
      return a.Cout == 128 ? launch_halo_t<3, 1, 8, 128, 2, false, 2, false, true, true, false, true>(b, num_sms, stream)
                           : launch_halo_t<3, 1, 8, 64, 2, false, 2, false, false, true, false, true>(b, num_sms, stream);
    return a.Cout == 128 ? launch_halo_t<3, 1, 8, 128, 2, false, 2, false, true, true>(b, num_sms, stream)
                         : launch_halo_t<3, 1, 8, 64, 2, false, 2, false, false, true>(b, num_sms, stream);
  }
  if (KH == 3 && nch == 8 && a.Cout == 128) {
    if (a.up2_out || a.out_f32) return -3006;
    if (tmah) return launch_halo_t<3, 1, 8, 128, 2, false, 2, false, true, false, false, true>(a, num_sms, stream);
    return launch_halo_t<3, 1, 8, 128, 2, false, 2, false, true>(a, num_sms, stream);
  }
  if (tmah && a.Cout == 64) return launch_halo_t<3, 1, 8, 64, 2, false, 2, false, false, false, false, true>(a, num_sms, stream);
  // two-source 64 + 64 -> 32 (dec3.conv1 on a materialised up-sample, FB_NO_UP1=1)
  if (tma_ok && KH == 3 && nch == 8 && epi2 && a.Cout == 32)
    return launch_halo_t<3, 1, 8, 32, 2, false, 2, false, false, false, false, true>(a, num_sms, stream);
  if (a.pool_out != nullptr) {
    // the space-to-depth stem with its max-pool fused (the 7x7 stride-2 form's filter bank and stages leave no room
    // for the pool buffers in 227 KB of shared memory: models with more than four bands keep the separate kernel)
    if (KH != 4 || !epi2 || !a.relu || !a.bias_in_args || a.residual || a.rowbias || a.up2_out || a.out_f32 || a.tile_list || a.Wout > 256 || a.Hout % 16 ||
        a.Wout % 16)
      return -3007;
    if (!tma_ok) return -3007;   // (the fused form stages parity planes by TMA only; callers ask halo_pool_fusable() first)
    return launch_halo_t<4, 1, 2, 64, 2, false, 2, false, false, false, true, true>(a, num_sms, stream);
  }
  if (KH == 4 && epi2 && tma_ok) return launch_halo_t<4, 1, 2, 64, 2, false, 2, false, false, false, false, true>(a, num_sms, stream);
  if (KH == 7 && epi2 && tma_ok) return launch_halo_t<7, 2, 1, 64, 2, false, 2, false, false, false, false, true>(a, num_sms, stream);
  if (KH == 4) return epi2 ? launch_halo_t<4, 1, 2, 64, 2, false, 2>(a, num_sms, stream)
                           : launch_halo_t<4, 1, 2, 64, 2>(a, num_sms, stream);
  if (epi2) {
    if (KH == 7) return launch_halo_t<7, 2, 1, 64, 2, false, 2>(a, num_sms, stream);
    if (nch == 8 && a.Cout == 32) return launch_halo_t<3, 1, 8, 32, 2, false, 2>(a, num_sms, stream);
    if (nch == 8 && a.Cout == 64) return launch_halo_t<3, 1, 8, 64, 2, false, 2>(a, num_sms, stream);
  }
  // (2x2-replicated outputs use the staged copy-out of the one-group epilogue)
  if (tma_ok && nch == 8 && a.Cout == 64 && a.C2 == 0) return launch_halo_t<3, 1, 8, 64, 2, false, 1, false, false, false, false, true>(a, num_sms, stream);
  // (one CTA per SM with two epilogue groups, per 148 tiles: head 349 -> 296 us, dec3.conv2 152 -> 144, dec4.conv1 the
  // same; dec4.conv2 233 -> 242, so that one keeps two CTAs per SM unless FB_ONE_CTA=2)
  static const bool one_cta_all = getenv("FB_ONE_CTA") && getenv("FB_ONE_CTA")[0] == '2';
  if (tma_ok && !a.up2_out && one_cta && a.direct_store && a.sink_cls == nullptr) {
    if (nch == 2 && a.Cout == 16 && one_cta_all) return launch_halo_t<3, 1, 2, 16, 4, false, 2, false, false, false, false, true>(a, num_sms, stream);
    if (nch == 4 && a.Cout == 32) return launch_halo_t<3, 1, 4, 32, 2, false, 2, false, false, false, false, true>(a, num_sms, stream);
  }
  if (tma_ok && !a.up2_out) {
    if (nch == 2 && a.Cout == 16) return launch_halo_t<3, 1, 2, 16, 4, false, 1, false, false, false, false, true>(a, num_sms, stream);
    if (nch == 4 && a.Cout == 32) return launch_halo_t<3, 1, 4, 32, 2, false, 1, false, false, false, false, true>(a, num_sms, stream);
  }
  if (KH == 7) return launch_halo_t<7, 2, 1, 64, 2>(a, num_sms, stream);
  if (nch == 2 && a.Cout == 16) return launch_halo_t<3, 1, 2, 16, 4>(a, num_sms, stream);
  if (nch == 2 && a.Cout == 32) return launch_halo_t<3, 1, 2, 32, 4>(a, num_sms, stream);   // head of a 17..32-class model
  if (nch == 4 && a.Cout == 16) return launch_halo_t<3, 1, 4, 16, 2>(a, num_sms, stream);
  if (nch == 4 && a.Cout == 32) return launch_halo_t<3, 1, 4, 32, 2>(a, num_sms, stream);
  if (nch == 8 && a.Cout == 32) return launch_halo_t<3, 1, 8, 32, 2>(a, num_sms, stream);
  if (nch == 8 && a.Cout == 64) return launch_halo_t<3, 1, 8, 64, 2>(a, num_sms, stream);
  return -3003;
}

}  // namespace fb
