// C ABI of libflairb200 (include/flair_b200.h): context, checkpoint folding/packing, the U-Net
// (ResNet34 encoder) layer graph, and the zone_detect / patch-predict loops built from the kernels in
// conv_igemm.cu and elementwise.cu. Topology restated from segmentation-models-pytorch 0.3.3
// `Unet(resnet34)` (un-vendored dependency of the reference, setup.py:36).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <climits>
#include <functional>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/flair_b200.h"
#include "comm.cuh"
#include "conv_halo.cuh"
#include "conv_igemm.cuh"
#include "elementwise.cuh"
#include "tile_need.cuh"

namespace {

std::string g_create_error;

struct ConvLayer {
  __nv_bfloat16* w = nullptr;       // device [Cout][Kpad] (im2col order, TMA / gather producers)
  __nv_bfloat16* w_halo = nullptr;  // device, K-step order of the halo-staged kernel (or null)
  __nv_bfloat16* w_halo_pair = nullptr;  // device, the same bank in CTA-pair stage order (128 -> 128 layers; or null)
  __nv_bfloat16* w_phase = nullptr; // device [4][Cout][Kp_phase]: sub-pixel phase form of a decoder conv1 (or null)
  int Kp_phase = 0;
  __nv_bfloat16* w_halo_phase = nullptr;  // device, phase form for the halo-staged kernel (32 -> 16 channels) (or null)
  __nv_bfloat16* w_s2d = nullptr;         // device, the stem as a 4x4 filter on the 2x2 space-to-depth image (<= 4 bands)
  __nv_bfloat16* w_d2s = nullptr;         // device, depth-to-space form of a 16-output-channel layer (conv_halo.cuh, HaloArgs::d2s)
  float* bias_d2s = nullptr;              // device [64]: bias[co] at (py*2+px)*16 + co
  int d2s_mode = 0;                       // 1: 16 -> 16 as a 4x4 stride-2 conv over cells, 2: upsampled 32 -> 16 on the low-res grid
  float* bias = nullptr;            // device [Cout]
  std::vector<float> bias_host;     // the same values (HaloArgs::bias_c)
  std::string name;                 // "layer1.0.conv1", ... (per-layer timings)
  int Cin = 0, Cout = 0, KH = 0, KW = 0, stride = 1, pad = 0, Ktot = 0, Kpad = 0;
  int C1 = 0, C2 = 0;               // channel split of a two-source (decoder conv1) layer
  double flops_px = 0;              // algorithmic FLOPs per output pixel: 2 * Cout * Cin * KH * KW, unpadded channels
};

struct Act {  // a named NHWC bf16 (or fp32) activation in the arena
  void* ptr = nullptr;
  int B = 0, H = 0, W = 0, C = 0;   // stored dims (already doubled when up2)
  int elem = 2;
  bool up2 = false;                 // producer writes it 2x2-replicated (decoder nearest x2 upsample)
};

// The tiles of the batch a network pass works on, for dead-output elimination (tile_need.cuh) and the fused
// class-map sink of the head: device and host copies of the same [n][6] table.
struct NeedCtx {
  const int* tiles_dev = nullptr;
  const int* tiles_host = nullptr;
  int n = 0, T = 0;
  bool restrict_tiles = false;   // false: the table only serves the head's sink, every kernel tile is computed
};
struct HeadSink {
  uint8_t* cls = nullptr;
  uint8_t* conf = nullptr;
  int64_t map_w = 0, map_row0 = 0;
};

static_assert(sizeof(fb_tile) == 6 * sizeof(int32_t), "fb_tile is handed to the kernels as int32 [n][6]");

struct ProfRec {
  int cat;               // 0 extract, 1 convs, 2 pool / MLP, 3 stitch; 9 = one conv layer (label), not part of the sums
  cudaEvent_t a, b;
  const char* label;     // layer name (owned by the ConvLayer), or null
};

}  // namespace

struct fb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int num_sms = 148;
  std::string err;
  int64_t launches = 0;
  bool force_gather = false;  // FB_FORCE_GATHER=1: cp.async im2col producer everywhere (debug / A-B tests)
  bool no_halo = false;       // FB_NO_HALO=1: skip the halo-staged kernel
  bool no_phase = false;      // FB_NO_PHASE=1: decoder conv1 on the materialised upsample instead of sub-pixel phases
  bool dec_phase[6] = {false, false, false, false, false, false};  // per decoder block, decided by arena_plan
  bool dec_up1[6] = {false, false, false, false, false, false};    // block reads its low-res x1 through the halo kernel's x2 upsample
  bool no_up1 = true;         // FB_NO_UP1=0: such blocks read their low-res x1 through the halo kernel's x2 up-sampling gather
  int front_chunk = 0;        // FB_FRONT_CHUNK: tiles per stem + max-pool chunk (0 = the whole batch)
  bool no_s2d = false;        // FB_NO_S2D=1: 7x7 stride-2 stem on the 8-channel-padded tile also for <= 4 bands
  bool stem_s2d = false;      // decided by arena_plan: x0 is stored in space-to-depth form
  bool full_tiles = false;    // FB_FULL_TILES=1: no dead-output elimination in the exact-clipping zone loop
  bool no_fused_sink = false; // FB_NO_FUSED_SINK=1: head writes fp32 logits, K6 runs as its own kernel
  bool no_d2s = false;        // FB_NO_D2S=1: dec4 / head as N = 16 convs instead of the depth-to-space forms
  bool aligned_tiles = false; // FB_ALIGNED_TILES=1: active kernel tiles on the fixed tile grid instead of origin-shifted (tile_need.cuh)
  bool no_narrow_tiles = false; // FB_NO_NARROW_TILES=1: halo kernel tiles of an origin-shifted list always compute both 8-column blocks
  int sub_tiles = 2;          // FB_SUB_TILES=0 / 1: the implicit GEMM's list entries stay whole 8 x 16 boxes / halves of 4 x 16 (default: quarters of 4 x 8)
  bool dec_pair = false;      // FB_DEC_PAIR=1: dec0 stays on the CTA-pair kernel (no tile lists) under origin-shifted tiles
  bool no_pool_fuse = false;  // FB_NO_POOL_FUSE=1: the stem's max-pool as a kernel of its own
  bool no_hpair = false;      // FB_NO_HPAIR=1: halo kernel always as single CTAs (no cta_group::2 pairs)
  bool d2s_all = true;        // FB_D2S_ALL=0: dec4.conv2 (bf16 output) on the N = 16 kernel instead of the depth-to-space form
  bool no_sb = false;         // FB_NO_SB=1: the 128 -> 128 layers on the im2col implicit GEMM instead of the halo kernel with streamed weights
  double flops = 0;           // algorithmic FLOPs of the conv outputs actually computed since creation
  int* list_dev = nullptr;    // active-tile lists of the current network pass
  size_t list_cap = 0;        // in ints
  fb::TileListPlan list_plan; // tiling of every decoder launch of the pass (make_tile_list)
  size_t list_plan_ints = 0;
  int plan_mode = 0;          // 1: planning walk (run_conv launches nothing), 2: real walk with planned lists

  // model
  bool loaded = false;
  int in_ch = 0, ncls = 0, use_meta = 0;
  int ls = 16;                // floats per pixel of the logits: 16 for <= 16 classes, 32 above
  std::map<std::string, ConvLayer> conv;
  float* mlp[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  std::vector<void*> owned;  // device allocations freed at destroy

  // normalisation
  bool norm_set = false;
  __nv_bfloat16* lut = nullptr;  // device [8][256]

  // raster
  const uint8_t* raster = nullptr;
  uint8_t* raster_own = nullptr;
  size_t raster_own_bytes = 0;
  uint8_t* maps_own = nullptr;   // class / confidence (/ truth) maps of fb_detect_zone_host / fb_detect_zone_shard
  size_t maps_own_bytes = 0;
  fb::Comm* comm = nullptr;      // fb_comm_init
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;  // copy streams of fb_detect_zone_host (non-blocking)
  std::vector<cudaEvent_t> copy_events;                     // grow-only pool, timing disabled
  int bands_total = 0, rc = 0, layout = 0;
  int64_t W = 0, H = 0, row0 = 0, rows = 0;
  int* band_idx_dev = nullptr;
  int* ident_dev = nullptr;      // 0..7: band selection of fb_predict_patches (patches hold the selected bands only)

  // activation arena
  uint8_t* arena = nullptr;
  size_t arena_bytes = 0, arena_used = 0;
  int arena_n = 0, arena_T = 0;
  std::map<std::string, Act> acts;
  int* tile_xy_dev = nullptr;  // [cap][2]
  int* tiles_dev = nullptr;    // [cap][6]
  int tile_cap = 0;
  int* win_dev = nullptr;      // [cap][6] metric windows of fb_detect_strip_metrics
  int win_cap = 0;
  float* meta_dev = nullptr;   // [n][45]
  float* menc_dev = nullptr;   // [n][16]
  int meta_cap = 0;

  // profiling
  bool prof = false;
  bool prof_layers = false;   // FB_LAYER_TIMES=1 at fb_profile_begin: one event pair per conv launch, table on stderr at fb_profile_end
  std::vector<ProfRec> prof_recs;
};

namespace {

int fail(fb_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg; else g_create_error = msg;
  return code;
}
int cuda_fail(fb_ctx* c, cudaError_t e, const char* what) {
  return fail(c, static_cast<int>(e), std::string(what) + ": " + cudaGetErrorString(e));
}
#define FB_CUDA(c, call)                                     \
  do {                                                       \
    cudaError_t e__ = (call);                                \
    if (e__ != cudaSuccess) return cuda_fail(c, e__, #call); \
  } while (0)
#define FB_TRY(expr)              \
  do {                            \
    int rc__ = (expr);            \
    if (rc__ != 0) return rc__;   \
  } while (0)

struct ProfScope {
  fb_ctx* c;
  ProfRec r;
  bool on;
  ProfScope(fb_ctx* c_, int cat, const char* label = nullptr) : c(c_), on(c_->prof && cat >= 0 && (cat != 9 || c_->prof_layers)) {
    if (on) {
      r.cat = cat;
      r.label = label;
      cudaEventCreate(&r.a);
      cudaEventCreate(&r.b);
      cudaEventRecord(r.a, c->stream);
    }
  }
  ~ProfScope() {
    if (on) {
      cudaEventRecord(r.b, c->stream);
      c->prof_recs.push_back(r);
    }
  }
};

uint16_t f32_to_bf16_rne(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u && (u & 0x007FFFFFu)) return static_cast<uint16_t>((u >> 16) | 0x40);
  const uint32_t lsb = (u >> 16) & 1u;
  u += 0x7FFFu + lsb;
  return static_cast<uint16_t>(u >> 16);
}

typedef std::unordered_map<std::string, const fb_tensor_desc*> TensorMap;

const fb_tensor_desc* find(const TensorMap& m, const std::string& k) {
  auto it = m.find(k);
  return it == m.end() ? nullptr : it->second;
}
int64_t numel(const fb_tensor_desc* t) {
  int64_t n = 1;
  for (int i = 0; i < t->ndim; ++i) n *= t->shape[i];
  return n;
}

// conv weight [Cout][Cin][KH][KW] (+ optional eval-mode BatchNorm `bn.*`, + optional bias) ->
// bf16 [CoutPad][Kpad] with k = (kh*KW + kw)*CinPad + cin, fp32 bias [CoutPad].
int build_conv(fb_ctx* c, const TensorMap& tm, const std::string& name, const std::string& wkey,
               const std::string& bn, const std::string& bkey, int Cin, int Cout, int KH, int stride,
               int pad, int C2 = 0, bool decoder_conv1 = false) {
  const fb_tensor_desc* w = find(tm, wkey);
  if (!w) return fail(c, FB_ERR_WEIGHTS, "missing tensor " + wkey);
  if (w->ndim != 4 || w->shape[0] != Cout || w->shape[1] != Cin || w->shape[2] != KH || w->shape[3] != KH)
    return fail(c, FB_ERR_WEIGHTS, "bad shape for " + wkey);
  std::vector<double> scale(Cout, 1.0), shift(Cout, 0.0);
  if (!bn.empty()) {
    const fb_tensor_desc* g = find(tm, bn + ".weight");
    const fb_tensor_desc* b = find(tm, bn + ".bias");
    const fb_tensor_desc* mu = find(tm, bn + ".running_mean");
    const fb_tensor_desc* var = find(tm, bn + ".running_var");
    if (!g || !b || !mu || !var) return fail(c, FB_ERR_WEIGHTS, "missing BatchNorm tensors " + bn + ".*");
    if (numel(g) != Cout || numel(b) != Cout || numel(mu) != Cout || numel(var) != Cout)
      return fail(c, FB_ERR_WEIGHTS, "bad BatchNorm shape " + bn);
    for (int o = 0; o < Cout; ++o) {
      // torch: y = (x - mean) / sqrt(var + eps) * gamma + beta, eps = 1e-5
      const double s = static_cast<double>(g->data[o]) / sqrt(static_cast<double>(var->data[o]) + 1e-5);
      scale[o] = s;
      shift[o] = static_cast<double>(b->data[o]) - static_cast<double>(mu->data[o]) * s;
    }
  }
  if (!bkey.empty()) {
    const fb_tensor_desc* b = find(tm, bkey);
    if (!b || numel(b) != Cout) return fail(c, FB_ERR_WEIGHTS, "missing/bad tensor " + bkey);
    for (int o = 0; o < Cout; ++o) shift[o] += static_cast<double>(b->data[o]) * scale[o];
  }
  ConvLayer L;
  const int CinPad = (Cin + 7) / 8 * 8;
  const int CoutPad = (Cout + 15) / 16 * 16;
  L.Cin = CinPad; L.Cout = CoutPad; L.KH = KH; L.KW = KH; L.stride = stride; L.pad = pad;
  L.Ktot = KH * KH * CinPad;
  L.Kpad = (L.Ktot + 63) / 64 * 64;
  L.C2 = C2;
  L.C1 = CinPad - C2;
  L.flops_px = 2.0 * Cout * Cin * KH * KH;
  std::vector<uint16_t> packed(static_cast<size_t>(CoutPad) * L.Kpad, 0);
  std::vector<float> bias(CoutPad, 0.f);
  std::vector<float> folded(static_cast<size_t>(Cout) * Cin * KH * KH);
  for (int o = 0; o < Cout; ++o) {
    bias[o] = static_cast<float>(shift[o]);
    for (int ci = 0; ci < Cin; ++ci)
      for (int kh = 0; kh < KH; ++kh)
        for (int kw = 0; kw < KH; ++kw) {
          const size_t wi = ((static_cast<size_t>(o) * Cin + ci) * KH + kh) * KH + kw;
          const float v = static_cast<float>(static_cast<double>(w->data[wi]) * scale[o]);
          folded[wi] = v;
          packed[static_cast<size_t>(o) * L.Kpad + (kh * KH + kw) * CinPad + ci] = f32_to_bf16_rne(v);
        }
  }
  // second packing for the halo-staged kernel when the channel configuration has an instantiation
  if (fb::halo_supported(KH, stride, L.C1, L.C2, CoutPad, 16, 64)) {
    const size_t n = fb::pack_halo_weights(folded.data(), Cout, CoutPad, Cin, CinPad, KH, stride, L.C1, L.C2, nullptr);
    std::vector<uint16_t> hp(n);
    fb::pack_halo_weights(folded.data(), Cout, CoutPad, Cin, CinPad, KH, stride, L.C1, L.C2, hp.data());
    FB_CUDA(c, cudaMalloc(&L.w_halo, n * 2));
    c->owned.push_back(L.w_halo);
    FB_CUDA(c, cudaMemcpy(L.w_halo, hp.data(), n * 2, cudaMemcpyHostToDevice));
    if (KH == 3 && stride == 1 && CoutPad == 128 && L.C1 == 128 && L.C2 == 0) {
      std::vector<uint16_t> pp(n);
      fb::pack_halo_weights_pair128(hp.data(), n, pp.data());
      FB_CUDA(c, cudaMalloc(&L.w_halo_pair, n * 2));
      c->owned.push_back(L.w_halo_pair);
      FB_CUDA(c, cudaMemcpy(L.w_halo_pair, pp.data(), n * 2, cudaMemcpyHostToDevice));
    }
  }
  // the stem on the 2x2 space-to-depth image (conv_halo.cuh): w2[o][(py*2+px)*Cin + c][a][b] = w[o][c][2a+py-1][2b+px-1]
  if (name == "stem" && KH == 7 && stride == 2 && Cin <= 4 && Cout == 64) {
    std::vector<float> w2(static_cast<size_t>(Cout) * 16 * 4 * 4, 0.f);
    for (int o = 0; o < Cout; ++o)
      for (int ci = 0; ci < Cin; ++ci)
        for (int a = 0; a < 4; ++a)
          for (int b = 0; b < 4; ++b)
            for (int py = 0; py < 2; ++py)
              for (int px = 0; px < 2; ++px) {
                const int kh = 2 * a + py - 1, kw = 2 * b + px - 1;
                if (kh < 0 || kw < 0 || kh > 6 || kw > 6) continue;
                w2[((static_cast<size_t>(o) * 16 + (py * 2 + px) * Cin + ci) * 4 + a) * 4 + b] =
                    folded[((static_cast<size_t>(o) * Cin + ci) * 7 + kh) * 7 + kw];
              }
    const size_t n = fb::pack_halo_weights(w2.data(), Cout, CoutPad, 16, 16, 4, 1, 16, 0, nullptr);
    std::vector<uint16_t> hp(n);
    fb::pack_halo_weights(w2.data(), Cout, CoutPad, 16, 16, 4, 1, 16, 0, hp.data());
    FB_CUDA(c, cudaMalloc(&L.w_s2d, n * 2));
    c->owned.push_back(L.w_s2d);
    FB_CUDA(c, cudaMemcpy(L.w_s2d, hp.data(), n * 2, cudaMemcpyHostToDevice));
  }
  // third packing for decoder conv1 layers: the sub-pixel phase form (conv_igemm.cuh, ConvArgs::phase_mode).
  // x1 = the upsampled source: taps that read the same low-res pixel are summed in fp32 before the single
  // bf16 rounding. rows(a): output-row parity a, low-res tap di -> original taps {kh}.
  if (decoder_conv1 && KH == 3 && stride == 1 && L.C1 % 64 == 0 && L.C2 % 64 == 0 && Cin == CinPad) {
    const int C1 = L.C1, C2 = L.C2;
    L.Kp_phase = 4 * C1 + 9 * C2;
    std::vector<uint16_t> pp(static_cast<size_t>(4) * CoutPad * L.Kp_phase, 0);
    auto taps = [](int parity, int d, int* out) -> int {   // original taps feeding low-res tap d
      if (parity == 0) { if (d == 0) { out[0] = 0; return 1; } out[0] = 1; out[1] = 2; return 2; }
      if (d == 0) { out[0] = 0; out[1] = 1; return 2; }
      out[0] = 2; return 1;
    };
    for (int pa = 0; pa < 2; ++pa)
      for (int pb = 0; pb < 2; ++pb)
        for (int o = 0; o < Cout; ++o) {
          uint16_t* row = pp.data() + (static_cast<size_t>(pa * 2 + pb) * CoutPad + o) * L.Kp_phase;
          for (int di = 0; di < 2; ++di)
            for (int dj = 0; dj < 2; ++dj) {
              int khs[2], kws[2];
              const int nh = taps(pa, di, khs), nw = taps(pb, dj, kws);
              for (int ci = 0; ci < C1; ++ci) {
                double s = 0.0;
                for (int i = 0; i < nh; ++i)
                  for (int j = 0; j < nw; ++j)
                    s += folded[((static_cast<size_t>(o) * Cin + ci) * 3 + khs[i]) * 3 + kws[j]];
                row[(di * 2 + dj) * C1 + ci] = f32_to_bf16_rne(static_cast<float>(s));
              }
            }
          for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw)
              for (int ci = 0; ci < C2; ++ci)
                row[4 * C1 + (kh * 3 + kw) * C2 + ci] =
                    f32_to_bf16_rne(folded[((static_cast<size_t>(o) * Cin + C1 + ci) * 3 + kh) * 3 + kw]);
        }
    FB_CUDA(c, cudaMalloc(&L.w_phase, pp.size() * 2));
    c->owned.push_back(L.w_phase);
    FB_CUDA(c, cudaMemcpy(L.w_phase, pp.data(), pp.size() * 2, cudaMemcpyHostToDevice));
  }
  // the same decomposition for the channel-poor last decoder block, in the halo kernel's step order
  if (decoder_conv1 && KH == 3 && stride == 1 && fb::halo_phase_supported(L.C1, L.C2, CoutPad, 16, 8)) {
    const size_t n = fb::pack_halo_weights_phase(folded.data(), Cout, CoutPad, Cin, CinPad, nullptr);
    std::vector<uint16_t> hp(n);
    fb::pack_halo_weights_phase(folded.data(), Cout, CoutPad, Cin, CinPad, hp.data());
    FB_CUDA(c, cudaMalloc(&L.w_halo_phase, n * 2));
    c->owned.push_back(L.w_halo_phase);
    FB_CUDA(c, cudaMemcpy(L.w_halo_phase, hp.data(), n * 2, cudaMemcpyHostToDevice));
  }
  // the 16-output-channel layers at full resolution (dec4, head) in depth-to-space form
  {
    const int mode = (KH == 3 && stride == 1 && CoutPad == 16 && Cin == CinPad)
                         ? ((!decoder_conv1 && Cin == 16) ? 1 : (decoder_conv1 && L.C1 == 32 && L.C2 == 0) ? 2 : 0) : 0;
    if (mode) {
      const size_t n = fb::pack_halo_weights_d2s(mode, folded.data(), Cout, Cin, nullptr);
      std::vector<uint16_t> hp(n);
      fb::pack_halo_weights_d2s(mode, folded.data(), Cout, Cin, hp.data());
      FB_CUDA(c, cudaMalloc(&L.w_d2s, n * 2));
      c->owned.push_back(L.w_d2s);
      FB_CUDA(c, cudaMemcpy(L.w_d2s, hp.data(), n * 2, cudaMemcpyHostToDevice));
      std::vector<float> b64(64, 0.f);
      for (int q = 0; q < 4; ++q)
        for (int o = 0; o < Cout; ++o) b64[q * 16 + o] = bias[o];
      FB_CUDA(c, cudaMalloc(&L.bias_d2s, 64 * 4));
      c->owned.push_back(L.bias_d2s);
      FB_CUDA(c, cudaMemcpy(L.bias_d2s, b64.data(), 64 * 4, cudaMemcpyHostToDevice));
      L.d2s_mode = mode;
    }
  }
  FB_CUDA(c, cudaMalloc(&L.w, packed.size() * 2));
  c->owned.push_back(L.w);
  L.bias_host = bias;
  FB_CUDA(c, cudaMalloc(&L.bias, bias.size() * 4));
  c->owned.push_back(L.bias);
  FB_CUDA(c, cudaMemcpyAsync(L.w, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice, c->stream));
  FB_CUDA(c, cudaMemcpyAsync(L.bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice, c->stream));
  FB_CUDA(c, cudaStreamSynchronize(c->stream));  // host vectors die at scope exit
  L.name = name;
  c->conv[name] = L;
  return 0;
}

int upload_f32(fb_ctx* c, const TensorMap& tm, const std::string& key, int64_t n, float** out) {
  const fb_tensor_desc* t = find(tm, key);
  if (!t || numel(t) != n) return fail(c, FB_ERR_WEIGHTS, "missing/bad tensor " + key);
  FB_CUDA(c, cudaMalloc(out, n * 4));
  c->owned.push_back(*out);
  FB_CUDA(c, cudaMemcpy(*out, t->data, n * 4, cudaMemcpyHostToDevice));
  return 0;
}

// ---------------------------------------------------------------------------------- arena
const int kStageBlocks[4] = {3, 4, 6, 3};
const int kStageCh[4] = {64, 128, 256, 512};
const int kDecOut[5] = {256, 128, 64, 32, 16};

int arena_alloc(fb_ctx* c, const std::string& name, int B, int H, int W, int C, int elem, bool dry,
                bool up2 = false) {
  if (up2) { H *= 2; W *= 2; }
  const size_t bytes = (static_cast<size_t>(B) * H * W * C * elem + 1023) / 1024 * 1024;
  if (!dry) {
    Act a;
    a.ptr = c->arena + c->arena_used;
    a.B = B; a.H = H; a.W = W; a.C = C; a.elem = elem; a.up2 = up2;
    c->acts[name] = a;
  }
  c->arena_used += bytes;
  return 0;
}

void arena_plan(fb_ctx* c, int n, int T, bool dry) {
  c->arena_used = 0;
  if (!dry) c->acts.clear();
  // decoder block d runs its conv1 in sub-pixel phase form when the layer has the packing and the low-res
  // grid tiles into 8 x 16 boxes; its x1 input is then kept at low resolution, otherwise the producer
  // writes it 2x2-replicated
  for (int d = 0; d < 5; ++d) {
    char nm[32];
    snprintf(nm, sizeof nm, "dec%d.conv1", d);
    auto it = c->conv.find(nm);
    const int S_lo = (T / 32) << d;   // low-res side of block d's x1
    // block 3 (64+64 -> 32 at 256^2) is faster in the halo-staged kernel than as 4 x 13 narrow-N GEMM steps
    const char* pm = getenv("FB_PHASE_MAX");
    const int phase_max = pm ? atoi(pm) : 2;
    c->dec_phase[d] = !c->force_gather && !c->no_phase && d <= phase_max && it != c->conv.end() && it->second.w_phase != nullptr &&
                      S_lo % 16 == 0;
    // the last block (32 -> 16 at 512^2) has its own phase form inside the halo-staged kernel
    const char* hp = getenv("FB_NO_HALO_PHASE");
    if (!c->force_gather && !c->no_phase && !c->no_halo && !(hp && hp[0] == '1') && it != c->conv.end() &&
        it->second.w_halo_phase != nullptr && fb::halo_phase_supported(it->second.C1, it->second.C2, it->second.Cout, S_lo, S_lo))
      c->dec_phase[d] = true;
    // no phase form: the halo kernel can still read the low-res x1 itself (upsample in its gather) when the block
    // is one of its shapes; otherwise the producer materialises the 2x2 replication
    c->dec_up1[d] = false;
    if (!c->dec_phase[d] && d >= 1 && !c->no_up1 && !c->force_gather && !c->no_halo && it != c->conv.end() &&
        it->second.w_halo != nullptr &&
        fb::halo_supported(3, 1, it->second.C1, it->second.C2, it->second.Cout, 2 * S_lo, 2 * S_lo))
      c->dec_up1[d] = true;
  }
  {
    auto it = c->conv.find("stem");
    c->stem_s2d = !c->no_s2d && !c->force_gather && !c->no_halo && it != c->conv.end() && it->second.w_s2d != nullptr &&
                  c->in_ch <= 4 && fb::halo_supported(4, 1, 16, 0, 64, T / 2, T / 2);
  }
  if (c->stem_s2d) arena_alloc(c, "x0", n, T / 2, T / 2, 16, 2, dry);
  else arena_alloc(c, "x0", n, T, T, 8, 2, dry);
  arena_alloc(c, "f1", n, T / 2, T / 2, 64, 2, dry);
  arena_alloc(c, "pool", n, T / 4, T / 4, 64, 2, dry);
  int S = T / 4;
  for (int st = 0; st < 4; ++st) {
    const int C = kStageCh[st];
    char buf[64];
    snprintf(buf, sizeof buf, "layer%d.tmp", st + 1);
    arena_alloc(c, buf, n, S, S, C, 2, dry);
    if (st > 0) {
      snprintf(buf, sizeof buf, "layer%d.ds", st + 1);
      arena_alloc(c, buf, n, S, S, C, 2, dry);
    }
    for (int b = 0; b < kStageBlocks[st]; ++b) {
      snprintf(buf, sizeof buf, "layer%d.%d.out", st + 1, b);
      // the bottleneck feature only feeds decoder block 0, which reads it through the x2 upsample
      arena_alloc(c, buf, n, S, S, C, 2, dry, st == 3 && b == kStageBlocks[st] - 1 && !c->dec_phase[0]);
    }
    S /= 2;
  }
  S = T / 16;
  for (int d = 0; d < 5; ++d) {
    char buf[64];
    snprintf(buf, sizeof buf, "dec%d.mid", d);
    arena_alloc(c, buf, n, S, S, kDecOut[d], 2, dry);
    snprintf(buf, sizeof buf, "dec%d", d);
    // dec0..3 feed the next block through the x2 upsample: materialised here unless that block runs in phase form
    arena_alloc(c, buf, n, S, S, kDecOut[d], 2, dry, d < 4 && !c->dec_phase[d + 1] && !c->dec_up1[d + 1]);
    S *= 2;
  }
  arena_alloc(c, "logits", n, T, T, c->ls, 4, dry);
}

int ensure_arena(fb_ctx* c, int n, int T) {
  if (c->arena && c->arena_n == n && c->arena_T == T) return 0;
  arena_plan(c, n, T, true);
  const size_t need = c->arena_used;
  if (need > c->arena_bytes) {
    FB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->arena) cudaFree(c->arena);
    c->arena = nullptr;
    c->arena_bytes = 0;
    cudaError_t e = cudaMalloc(&c->arena, need);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(c, FB_ERR_OOM, "activation arena: cudaMalloc of " + std::to_string(need) + " bytes failed");
    }
    c->arena_bytes = need;
  }
  arena_plan(c, n, T, false);
  c->arena_n = n;
  c->arena_T = T;
  return 0;
}

int ensure_tile_buffers(fb_ctx* c, int n) {
  if (n <= c->tile_cap) return 0;
  FB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->tile_xy_dev) cudaFree(c->tile_xy_dev);
  if (c->tiles_dev) cudaFree(c->tiles_dev);
  c->tile_xy_dev = nullptr; c->tiles_dev = nullptr; c->tile_cap = 0;
  FB_CUDA(c, cudaMalloc(&c->tile_xy_dev, static_cast<size_t>(n) * 2 * sizeof(int)));
  FB_CUDA(c, cudaMalloc(&c->tiles_dev, static_cast<size_t>(n) * 6 * sizeof(int)));
  c->tile_cap = n;
  return 0;
}

int ensure_meta_buffers(fb_ctx* c, int n) {
  if (n <= c->meta_cap) return 0;
  FB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->meta_dev) cudaFree(c->meta_dev);
  if (c->menc_dev) cudaFree(c->menc_dev);
  c->meta_dev = nullptr; c->menc_dev = nullptr; c->meta_cap = 0;
  FB_CUDA(c, cudaMalloc(&c->meta_dev, static_cast<size_t>(n) * FB_METADATA_DIM * sizeof(float)));
  FB_CUDA(c, cudaMalloc(&c->menc_dev, static_cast<size_t>(n) * 16 * sizeof(float)));
  c->meta_cap = n;
  return 0;
}

// ---------------------------------------------------------------------------------- graph
// Active-tile lists (tile_need.cuh). The decoder of one network pass is walked twice: a planning walk in which
// run_conv only reports the tiling of the kernel it would launch (kernel tiles of th x tw pixels on the tile grid
// = output grid / scale, gh x gw of them per image) so that ONE build_tile_lists launch can expand the needed
// regions of all eleven layers, then the real walk, in which every launch picks up its list.
// *list stays null when every tile is active.
// sub (optional, in/out): the caller's kernel can take sub-boxes -- in: 1 = halves (th / 2 rows, two list entries per kernel
// tile), 2 = quarters (th / 2 rows x tw / 2 columns, four entries); out: the mode the list was built in (0 = whole tiles).
// *active then counts sub-boxes.
int make_tile_list(fb_ctx* c, const NeedCtx* need, int layer, int B, int scale, int th, int tw, int gh, int gw,
                   const int** list, long long* active, int* shifted, int* sub = nullptr, long long* blocks = nullptr) {
  // blocks (optional, halo kernels with tiles of two 8-column blocks): request narrow tiles; on return the number of
  // active blocks (for the FLOP count)
  *list = nullptr;
  *shifted = 0;
  const int want_sub = sub ? *sub : 0;
  if (sub) *sub = 0;
  const long long full = static_cast<long long>(B) * gh * gw;
  *active = full;
  if (blocks) *blocks = 2 * full;
  if (!need || !need->restrict_tiles || layer < 0 || layer >= fb::kNeedLayers || need->n != B) return 0;
  fb::TileListSpec& sp = c->list_plan.spec[layer];
  auto sub_tiling = [&](int mode, int& lth, int& ltw, int& lgh, int& lgw) {
    lth = mode ? th / 2 : th; lgh = mode ? gh * 2 : gh;
    ltw = mode == 2 ? tw / 2 : tw; lgw = mode == 2 ? gw * 2 : gw;
  };
  if (c->plan_mode == 1) {
    // origin-shifted tiles (tile_need.cuh, need_span) unless switched off or the packed entry cannot hold the launch
    const bool sh = !c->aligned_tiles && B <= (1 << fb::kTileImageBits) && gh * th < (1 << fb::kTileOriginBits) && gw * tw < (1 << fb::kTileOriginBits);
    int mode = (sh && th % 2 == 0 && tw % 2 == 0) ? want_sub : 0;
    if (mode > c->sub_tiles) mode = c->sub_tiles;
    int lth, ltw, lgh, lgw;
    sub_tiling(mode, lth, ltw, lgh, lgw);
    const bool half_x = blocks != nullptr && sh && !c->no_narrow_tiles && tw == 16 && B <= (1 << fb::kTileImageBits);
    long long nblocks = 0;
    const long long cnt = fb::count_active_tiles(need->tiles_host, B, need->T, layer, scale, lth, ltw, sh, half_x, &nblocks);
    const int per = mode == 2 ? 4 : mode == 1 ? 2 : 1;
    sp.shifted = sh ? 1 : 0;
    sp.half_x = half_x ? 1 : 0;
    sp.blocks = nblocks;
    sp.sub = per;
    sp.layer = layer; sp.scale = scale; sp.th = lth; sp.tw = ltw; sp.gh = lgh; sp.gw = lgw;
    sp.count = static_cast<int>(cnt);
    sp.offset = static_cast<int>(c->list_plan_ints);
    sp.use = cnt < static_cast<long long>(B) * lgh * lgw ? 1 : 0;
    if (sp.use) c->list_plan_ints += static_cast<size_t>((cnt + per - 1) / per * per);
    return 0;
  }
  const int mode = sp.sub == 4 ? 2 : sp.sub == 2 ? 1 : 0;
  int lth, ltw, lgh, lgw;
  sub_tiling(mode, lth, ltw, lgh, lgw);
  if (sp.scale != scale || sp.th != lth || sp.tw != ltw || sp.gh != lgh || sp.gw != lgw)
    return fail(c, FB_ERR_INVALID, "internal: tile list planned for another kernel tiling");
  if (sp.use) {
    *list = c->list_dev + sp.offset;
    *shifted = sp.shifted;
    *active = sp.count;
    if (sub) *sub = mode;
    if (blocks) *blocks = sp.blocks;
  }
  return 0;
}

// After the planning walk: one launch builds every planned list of this pass.
int build_planned_lists(fb_ctx* c, const NeedCtx* need) {
  if (c->list_plan_ints == 0) return 0;
  if (c->list_plan_ints > c->list_cap) {
    FB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->list_dev) cudaFree(c->list_dev);
    c->list_dev = nullptr; c->list_cap = 0;
    const size_t want = c->list_plan_ints + c->list_plan_ints / 4;
    FB_CUDA(c, cudaMalloc(&c->list_dev, want * sizeof(int)));
    c->list_cap = want;
  }
  const int rc = fb::launch_build_tile_lists(need->tiles_dev, need->n, need->T, c->list_plan, c->list_dev, c->stream);
  if (rc) return fail(c, rc, "tile list launch failed (code " + std::to_string(rc) + ")");
  c->launches++;
  return 0;
}

// need / layer: restrict the launch to the kernel tiles the write rectangles of the batch need (decoder layers of
// the exact-clipping zone loop). sink: the head writes class / confidence bytes instead of logits (*sunk says
// whether the kernel that ran could do it).
int run_conv(fb_ctx* c, const ConvLayer& L, const Act& x1, const Act* x2, const Act* res,
             const float* rowbias, bool relu, const Act& out, bool phase = false, const NeedCtx* need = nullptr,
             int layer = -1, const HeadSink* sink = nullptr, bool* sunk = nullptr, bool up1 = false) {
  // `out` may be stored 2x2-replicated (Act::up2): the conv itself runs at half those dims
  ProfScope layer_scope(c, c->plan_mode == 1 ? -1 : 9, L.name.c_str());
  const int Hout = out.up2 ? out.H / 2 : out.H, Wout = out.up2 ? out.W / 2 : out.W;
  const int C1 = x1.C, C2 = x2 ? x2->C : 0;
  if (C1 + C2 != L.Cin || out.C != L.Cout) return fail(c, FB_ERR_INVALID, "internal: conv channel mismatch");
  if (sunk) *sunk = false;
  const int* list = nullptr;
  long long active = 0;
  int shifted = 0;
  int rc;
  // depth-to-space form (16 output channels at full resolution): 64 accumulator columns = the 2x2 pixels of a cell
  // (measured per 148 tiles: head 368 -> 304 us, dec4.conv1 211 -> 200 us; dec4.conv2 first 301 -> 491 us with a spilled
  // producer table, then -- TMA-staged, origin-shifted tiles -- 210 -> 192 us, so it runs in this form too; FB_D2S_ALL=0
  // puts it back on the N = 16 kernel)
  if (L.d2s_mode && !c->no_d2s && !c->force_gather && !c->no_halo && !x2 && !res && !rowbias && !out.up2 && !up1 &&
      (L.d2s_mode == 2) == phase && (L.d2s_mode == 2 || out.elem == 4 || c->d2s_all) && fb::halo_d2s_supported(L.d2s_mode, C1, 0, L.Cout, Hout, Wout) &&
      (L.d2s_mode == 1 ? (x1.H == Hout && x1.W == Wout && !x1.up2) : (x1.H * 2 == Hout && x1.W * 2 == Wout))) {
    fb::HaloArgs h;
    memset(&h, 0, sizeof h);
    h.x1 = static_cast<const __nv_bfloat16*>(x1.ptr);
    h.C1 = C1;
    h.B = x1.B; h.Hin = x1.H; h.Win = x1.W; h.Hout = Hout; h.Wout = Wout;
    h.bias = L.bias_d2s;
    h.relu = relu ? 1 : 0;
    if (out.elem == 4) h.out_f32 = static_cast<float*>(out.ptr); else h.out = static_cast<__nv_bfloat16*>(out.ptr);
    h.wpacked = L.w_d2s;
    fb::halo_fill_steps_d2s(h, L.d2s_mode);
    long long blocks = 0;   // (narrow tiles: the second 8-cell block of the last tile of a row may be skipped)
    FB_TRY(make_tile_list(c, need, layer, x1.B, 2, 16, 16, Hout / 32, Wout / 32, &list, &active, &shifted, nullptr, &blocks));
    if (c->plan_mode == 1) return 0;
    if (list) { h.tile_list = list; h.tile_packed = shifted; h.num_m_tiles = static_cast<int>(active); }
    c->flops += static_cast<double>(blocks) * (32 * 16) * L.flops_px;
    if (sink && need && out.elem == 4) {
      h.sink_tiles = need->tiles_dev;
      h.sink_cls = sink->cls;
      h.sink_conf = sink->conf;
      h.sink_map_w = sink->map_w;
      h.sink_map_row0 = sink->map_row0;
      h.sink_ncls = c->ncls;
      if (sunk) *sunk = true;
    }
    // bias by value (cell order: bias[co] at (py*2+px)*16 + co); the sink wants the padded class columns masked
    h.bias_in_args = 1;
    for (int i = 0; i < 64; ++i) {
      const int co = i & 15;
      h.bias_c[i] = co < L.Cout && co < static_cast<int>(L.bias_host.size()) ? L.bias_host[co] : 0.f;
      if (h.sink_cls != nullptr && co >= c->ncls) h.bias_c[i] = -3.0e38f;
    }
    rc = fb::launch_conv_halo(h, 3, 1, c->num_sms, c->stream);
    if (rc != 0) return fail(c, rc, "depth-to-space conv launch failed (code " + std::to_string(rc) + ")");
    c->launches++;
    return 0;
  }
  if (phase) {
    // decoder conv1 in sub-pixel phase form: x1 at half the output resolution, x2 (skip) at full resolution
    if ((!L.w_phase && !L.w_halo_phase) || x1.H * 2 != Hout || x1.W * 2 != Wout || (x2 && (x2->H != Hout || x2->W != Wout)) || res || rowbias)
      return fail(c, FB_ERR_INVALID, "internal: phase-form conv shape mismatch");
    if (L.w_halo_phase && !x2 && fb::halo_phase_supported(C1, 0, L.Cout, x1.H, x1.W) && !out.up2 && out.elem == 2) {
      fb::HaloArgs h;
      memset(&h, 0, sizeof h);
      h.x1 = static_cast<const __nv_bfloat16*>(x1.ptr);
      h.C1 = C1;
      h.B = x1.B; h.Hin = x1.H; h.Win = x1.W; h.Hout = Hout; h.Wout = Wout;
      h.Cout = L.Cout;
      h.bias = L.bias;
      h.relu = relu ? 1 : 0;
      h.out = static_cast<__nv_bfloat16*>(out.ptr);
      h.wpacked = L.w_halo_phase;
      h.phase_mode = 1;
      fb::halo_fill_steps_phase(h);
      FB_TRY(make_tile_list(c, need, layer, x1.B, 2, 16, 8, x1.H / 16, x1.W / 8, &list, &active, &shifted));
      if (c->plan_mode == 1) return 0;
      if (list) { h.tile_list = list; h.tile_packed = shifted; h.num_m_tiles = static_cast<int>(active); }
      c->flops += static_cast<double>(active) * (16 * 8 * 4) * L.flops_px;
      rc = fb::launch_conv_halo(h, 3, 1, c->num_sms, c->stream);
      if (rc != 0) return fail(c, rc, "phase conv (halo) launch failed (code " + std::to_string(rc) + ")");
      c->launches++;
      return 0;
    }
    if (!L.w_phase) return fail(c, FB_ERR_INVALID, "internal: phase-form conv has no packing for this shape");
    fb::ConvArgs a;
    memset(&a, 0, sizeof a);
    a.x1 = static_cast<const __nv_bfloat16*>(x1.ptr);
    a.x2 = x2 ? static_cast<const __nv_bfloat16*>(x2->ptr) : nullptr;
    a.C1 = C1; a.C2 = C2;
    a.B = x1.B; a.Hin = Hout; a.Win = Wout; a.Hout = Hout; a.Wout = Wout;
    a.KH = 3; a.KW = 3; a.stride = 1; a.pad = 1;
    a.Cout = L.Cout;
    a.Ktot = L.Kp_phase;
    a.bias = L.bias;
    a.relu = relu ? 1 : 0;
    a.out = static_cast<__nv_bfloat16*>(out.ptr);
    a.up2_out = 0;
    a.phase_mode = 1;
    if (out.up2 || out.elem != 2 || !fb::conv_tma_eligible(a)) return fail(c, FB_ERR_INVALID, "internal: phase-form conv not eligible");
    int sub = 2;   // the implicit GEMM takes quarter boxes (4 x 8) under origin-shifted lists
    FB_TRY(make_tile_list(c, need, layer, x1.B, 2, 8, 16, Hout / 16, Wout / 32, &list, &active, &shifted, &sub));
    if (c->plan_mode == 1) return 0;
    const int per = sub == 2 ? 4 : sub == 1 ? 2 : 1;
    if (list) {
      a.tile_list = list; a.tile_packed = shifted; a.tile_sub = sub;
      a.tile_list_len = static_cast<int>((active + per - 1) / per);
    }
    c->flops += static_cast<double>(active) * (8 * 16 * 4 / per) * L.flops_px;
    rc = fb::launch_conv(a, L.w_phase, L.Kp_phase, true, c->num_sms, c->stream);
    if (rc != 0) return fail(c, rc, "phase conv launch failed (code " + std::to_string(rc) + ")");
    c->launches++;
    return 0;
  }
  if (up1) {
    if (x1.up2 || x1.H * 2 != Hout || x1.W * 2 != Wout || (x2 && (x2->H != Hout || x2->W != Wout)) || L.KH != 3 || L.stride != 1 ||
        !L.w_halo || !fb::halo_supported(3, 1, C1, C2, L.Cout, Hout, Wout))
      return fail(c, FB_ERR_INVALID, "internal: upsampling halo conv shape mismatch");
  } else if (x2 && (x2->H != x1.H || x2->W != x1.W)) {
    return fail(c, FB_ERR_INVALID, "internal: skip tensor shape mismatch");
  }
  if (up1 || (!c->force_gather && !c->no_halo && L.w_halo && L.C1 == C1 && L.C2 == C2 && !(c->no_sb && L.Cout == 128) &&
              !(out.up2 && L.Cout == 128) && fb::halo_supported(L.KH, L.stride, C1, C2, L.Cout, Hout, Wout))) {
    fb::HaloArgs h;
    memset(&h, 0, sizeof h);
    h.x1 = static_cast<const __nv_bfloat16*>(x1.ptr);
    h.x2 = x2 ? static_cast<const __nv_bfloat16*>(x2->ptr) : nullptr;
    h.C1 = C1; h.C2 = C2;
    h.up1 = up1 ? 1 : 0;
    h.B = x1.B; h.Hin = up1 ? Hout : x1.H; h.Win = up1 ? Wout : x1.W; h.Hout = Hout; h.Wout = Wout;
    h.Cout = L.Cout;
    h.bias = L.bias;
    h.residual = res ? static_cast<const __nv_bfloat16*>(res->ptr) : nullptr;
    h.rowbias = rowbias;
    h.relu = relu ? 1 : 0;
    if (out.elem == 4) h.out_f32 = static_cast<float*>(out.ptr); else h.out = static_cast<__nv_bfloat16*>(out.ptr);
    h.up2_out = out.up2 ? 1 : 0;
    h.wpacked = L.w_halo;
    h.wpacked_pair = L.w_halo_pair;
    h.pair = c->no_hpair ? 0 : 1;
    fb::halo_fill_steps(h, L.KH, L.stride);
    const int tw = 8 * fb::halo_blocks(L.KH, fb::halo_group_channels(L.KH, C1, C2) / 8, L.Cout);
    long long blocks = 0;
    FB_TRY(make_tile_list(c, need, layer, x1.B, 1, 16, tw, Hout / 16, Wout / tw, &list, &active, &shifted, nullptr, &blocks));
    if (c->plan_mode == 1) return 0;
    if (list) { h.tile_list = list; h.tile_packed = shifted; h.num_m_tiles = static_cast<int>(active); }
    c->flops += static_cast<double>(blocks) * (16 * tw / 2) * L.flops_px;
    if (sink && need && out.elem == 4 && (L.Cout == 16 || L.Cout == 32) && h.direct_store && !out.up2) {
      h.sink_tiles = need->tiles_dev;
      h.sink_cls = sink->cls;
      h.sink_conf = sink->conf;
      h.sink_map_w = sink->map_w;
      h.sink_map_row0 = sink->map_row0;
      h.sink_ncls = c->ncls;
      if (sunk) *sunk = true;
    }
    rc = fb::launch_conv_halo(h, L.KH, L.stride, c->num_sms, c->stream);
  } else {
    fb::ConvArgs a;
    memset(&a, 0, sizeof a);
    a.x1 = static_cast<const __nv_bfloat16*>(x1.ptr);
    a.x2 = x2 ? static_cast<const __nv_bfloat16*>(x2->ptr) : nullptr;
    a.C1 = C1; a.C2 = C2;
    a.up1 = 0;
    a.B = x1.B; a.Hin = x1.H; a.Win = x1.W;
    a.Hout = Hout; a.Wout = Wout;
    a.KH = L.KH; a.KW = L.KW; a.stride = L.stride; a.pad = L.pad;
    a.Cout = L.Cout;
    a.Ktot = L.Ktot;
    a.bias = L.bias;
    a.residual = res ? static_cast<const __nv_bfloat16*>(res->ptr) : nullptr;
    a.rowbias = rowbias;
    a.relu = relu ? 1 : 0;
    if (out.elem == 4) a.out_f32 = static_cast<float*>(out.ptr); else a.out = static_cast<__nv_bfloat16*>(out.ptr);
    a.up2_out = out.up2 ? 1 : 0;
    const bool tma = !c->force_gather && fb::conv_tma_eligible(a);
    const char* pair_env = getenv("FB_PAIR");
    // (the CTA-pair kernel does not take tile lists: it serves the N = 256 encoder layers, which run in full anyway)
    const int pair_sel = pair_env ? atoi(pair_env) : 256;
    const bool pair_layer = (pair_sel == 1 || pair_sel == fb::conv_pick_bn(L.Cout)) && L.KH == 3 && L.stride == 1 && Hout % 16 == 0 &&
                            fb::conv_pick_bn(L.Cout) >= 64;
    // ... except dec0 under origin-shifted tiles: an interior zone tile needs 22 of its 32 rows / columns = 3 x 2 boxes of
    // 8 x 16 instead of 4 x 2, which the single-CTA kernel can skip and the pair kernel (16 x 16 per pair) cannot
    const bool listed = need && need->restrict_tiles && layer >= 0 && !c->aligned_tiles && !c->dec_pair;
    if (tma && (!pair_layer || listed)) {
      int sub = 2;
      FB_TRY(make_tile_list(c, need, layer, x1.B, 1, 8, 16, Hout / 8, Wout / 16, &list, &active, &shifted, &sub));
      if (c->plan_mode == 1) return 0;
      const int per = sub == 2 ? 4 : sub == 1 ? 2 : 1;
      if (list) {
        a.tile_list = list; a.tile_packed = shifted; a.tile_sub = sub;
        a.tile_list_len = static_cast<int>((active + per - 1) / per);
      }
      c->flops += static_cast<double>(active) * (8 * 16 / per) * L.flops_px;
    } else {
      if (c->plan_mode == 1) return 0;
      c->flops += static_cast<double>(x1.B) * Hout * Wout * L.flops_px;
    }
    rc = fb::launch_conv(a, L.w, L.Kpad, tma, c->num_sms, c->stream);
  }
  if (rc != 0) return fail(c, rc, "conv launch failed (code " + std::to_string(rc) + ")");
  c->launches++;
  return 0;
}

// x0 (normalised tiles) must already be in the arena; enqueues encoder + decoder + head.
int run_network(fb_ctx* c, int n, int T, const float* menc_dev, const NeedCtx* need = nullptr,
                const HeadSink* sink = nullptr, bool* sunk = nullptr) {
  auto A = [&](const std::string& k) -> Act& { return c->acts[k]; };
  auto L = [&](const std::string& k) -> const ConvLayer& { return c->conv[k]; };
  // Stem + max-pool in chunks of front_chunk tiles: the stem's output (8.4 MB per 512^2 tile) is read straight back by
  // the bandwidth-bound pool kernel, and a chunk that fits the 126 MB L2 is read from there instead of from HBM.
  {
    const Act x0 = A("x0"), f1 = A("f1"), pool = A("pool");
    const size_t x0_px = static_cast<size_t>(x0.H) * x0.W * x0.C, f1_px = static_cast<size_t>(f1.H) * f1.W * f1.C,
                 pool_px = static_cast<size_t>(pool.H) * pool.W * pool.C;
    const int chunk = c->front_chunk > 0 ? c->front_chunk : n;
    for (int b0 = 0; b0 < n; b0 += chunk) {
      const int nb = n - b0 < chunk ? n - b0 : chunk;
      Act x0c = x0, f1c = f1;
      x0c.B = nb; x0c.ptr = static_cast<__nv_bfloat16*>(x0.ptr) + b0 * x0_px;
      f1c.B = nb; f1c.ptr = static_cast<__nv_bfloat16*>(f1.ptr) + b0 * f1_px;
      // the max-pool runs inside the stem's epilogue (HaloArgs::pool_out) when the stem is in space-to-depth form (<= 4
      // bands), the batch gives most SMs an image of their own (a CTA walks whole images there) and the rows fit the
      // carry buffers; FB_NO_POOL_FUSE=1 keeps the separate kernel
      const ConvLayer& S = L("stem");
      const bool stem_halo = c->stem_s2d || (!c->force_gather && !c->no_halo && S.w_halo && fb::halo_supported(7, 2, x0.C, 0, S.Cout, f1.H, f1.W));
      const bool fuse_pool = c->stem_s2d && !c->no_pool_fuse && fb::halo_pool_fusable() && c->front_chunk <= 0 && f1.W <= 256 && f1.H % 16 == 0 && f1.W % 16 == 0 &&
                             2 * nb >= c->num_sms;
      {
        ProfScope ps(c, 1);
        ProfScope ps_layer(c, 9, fuse_pool ? "stem+pool" : "stem");
        if (stem_halo) {
          fb::HaloArgs h;
          memset(&h, 0, sizeof h);
          h.x1 = static_cast<const __nv_bfloat16*>(x0c.ptr);
          h.C1 = x0.C;
          h.B = nb; h.Hin = x0.H; h.Win = x0.W; h.Hout = f1.H; h.Wout = f1.W;
          h.Cout = S.Cout;
          h.bias = S.bias;
          h.relu = 1;
          h.out = static_cast<__nv_bfloat16*>(f1c.ptr);
          h.wpacked = c->stem_s2d ? S.w_s2d : S.w_halo;
          fb::halo_fill_steps(h, c->stem_s2d ? 4 : 7, c->stem_s2d ? 1 : 2);
          if (fuse_pool) {
            h.bias_in_args = 1;
            for (int i = 0; i < 64; ++i) h.bias_c[i] = i < static_cast<int>(S.bias_host.size()) ? S.bias_host[i] : 0.f;
            h.pool_out = static_cast<__nv_bfloat16*>(pool.ptr) + b0 * pool_px;
            if (need && need->restrict_tiles && need->n == n && b0 == 0 && nb == n) {
              h.keep_tiles = need->tiles_dev;
              h.keep_T = need->T;
            }
          }
          c->flops += static_cast<double>(nb) * f1.H * f1.W * S.flops_px;
          const int rc = fb::launch_conv_halo(h, c->stem_s2d ? 4 : 7, c->stem_s2d ? 1 : 2, c->num_sms, c->stream);
          if (rc != 0) return fail(c, rc, "stem launch failed (code " + std::to_string(rc) + ")");
          c->launches++;
        } else {
          FB_TRY(run_conv(c, L("stem"), x0c, nullptr, nullptr, nullptr, true, f1c));
        }
      }
      if (!fuse_pool) {
        ProfScope ps(c, 2);
        ProfScope ps_layer(c, 9, "maxpool");
        int rc = fb::launch_maxpool3x3s2(static_cast<const __nv_bfloat16*>(f1c.ptr),
                                         static_cast<__nv_bfloat16*>(pool.ptr) + b0 * pool_px, nb, f1.H, f1.W, 64,
                                         c->num_sms, c->stream);
        if (rc) return fail(c, rc, "maxpool launch failed");
        c->launches++;
      }
    }
  }
  ProfScope ps_convs(c, 1);  // every launch from here to the end of the function is a conv
  std::string cur = "pool";
  for (int st = 0; st < 4; ++st) {
    char nm[64], tmp[64], ds[64];
    snprintf(tmp, sizeof tmp, "layer%d.tmp", st + 1);
    snprintf(ds, sizeof ds, "layer%d.ds", st + 1);
    for (int b = 0; b < kStageBlocks[st]; ++b) {
      snprintf(nm, sizeof nm, "layer%d.%d", st + 1, b);
      const std::string base(nm);
      const std::string outn = base + ".out";
      const Act* res = &A(cur);
      FB_TRY(run_conv(c, L(base + ".conv1"), A(cur), nullptr, nullptr, nullptr, true, A(tmp)));
      if (st > 0 && b == 0) {
        FB_TRY(run_conv(c, L(base + ".downsample"), A(cur), nullptr, nullptr, nullptr, false, A(ds)));
        res = &A(ds);
      }
      const bool last = (st == 3 && b == kStageBlocks[st] - 1);
      FB_TRY(run_conv(c, L(base + ".conv2"), A(tmp), nullptr, res, last ? menc_dev : nullptr, true, A(outn)));
      cur = outn;
    }
  }
  const char* skips[5] = {"layer3.5.out", "layer2.3.out", "layer1.2.out", "f1", nullptr};
  const std::string enc_out = cur;
  auto decoder = [&]() -> int {
    std::string x = enc_out;
    for (int d = 0; d < 5; ++d) {
      char nm[64];
      snprintf(nm, sizeof nm, "dec%d", d);
      const std::string base(nm);
      FB_TRY(run_conv(c, L(base + ".conv1"), A(x), skips[d] ? &A(skips[d]) : nullptr, nullptr, nullptr, true, A(base + ".mid"),
                      c->dec_phase[d], need, 2 * d, nullptr, nullptr, c->dec_up1[d]));
      FB_TRY(run_conv(c, L(base + ".conv2"), A(base + ".mid"), nullptr, nullptr, nullptr, true, A(base), false, need, 2 * d + 1));
      x = base;
    }
    return run_conv(c, L("head"), A(x), nullptr, nullptr, nullptr, false, A("logits"), false, need, 10, sink, sunk);
  };
  if (need && need->restrict_tiles) {
    memset(&c->list_plan, 0, sizeof c->list_plan);
    c->list_plan_ints = 0;
    c->plan_mode = 1;
    const int rc = decoder();
    c->plan_mode = 0;
    if (rc) return rc;
    FB_TRY(build_planned_lists(c, need));
    c->plan_mode = 2;
    const int rc2 = decoder();
    c->plan_mode = 0;
    return rc2;
  }
  return decoder();
}

int run_extract(fb_ctx* c, const uint8_t* raster, int layout, int bands_total, const int* band_idx_dev,
                int rc_, int64_t W, int64_t H, int64_t row0, int64_t rows, const int* xy_dev, int n, int T) {
  ProfScope ps(c, 0);
  int rc = c->stem_s2d
               ? fb::launch_extract_normalise_s2d(raster, layout, bands_total, band_idx_dev, rc_, W, H, row0, rows, xy_dev, n, T,
                                                  c->lut, static_cast<__nv_bfloat16*>(c->acts["x0"].ptr), c->num_sms, c->stream)
               : fb::launch_extract_normalise(raster, layout, bands_total, band_idx_dev, rc_, W, H, row0, rows, xy_dev, n, T,
                                              c->lut, static_cast<__nv_bfloat16*>(c->acts["x0"].ptr), c->num_sms, c->stream);
  if (rc) return fail(c, rc, "extract launch failed");
  c->launches++;
  return 0;
}

int run_metadata(fb_ctx* c, const float* metadata_host, int n, const float** menc) {
  *menc = nullptr;
  if (!c->use_meta) return 0;
  if (!metadata_host) return fail(c, FB_ERR_INVALID, "model was loaded with use_metadata=1 but metadata is NULL");
  FB_TRY(ensure_meta_buffers(c, n));
  FB_CUDA(c, cudaMemcpyAsync(c->meta_dev, metadata_host, static_cast<size_t>(n) * FB_METADATA_DIM * 4,
                             cudaMemcpyHostToDevice, c->stream));
  ProfScope ps(c, 2);
  int rc = fb::launch_metadata_mlp(c->meta_dev, c->mlp, c->menc_dev, n, c->stream);
  if (rc) return fail(c, rc, "metadata MLP launch failed");
  c->launches++;
  *menc = c->menc_dev;
  return 0;
}

int check_ready(fb_ctx* c, bool need_raster, int tile) {
  if (!c) return FB_ERR_INVALID;
  if (!c->loaded) return fail(c, FB_ERR_STATE, "fb_load_weights has not been called");
  if (!c->norm_set) return fail(c, FB_ERR_STATE, "fb_set_norm has not been called");
  if (need_raster && !c->raster) return fail(c, FB_ERR_STATE, "no raster set (fb_set_raster / fb_upload_raster)");
  if (tile <= 0 || tile % 32 != 0) return fail(c, FB_ERR_INVALID, "tile size must be a positive multiple of 32 (U-Net depth 5)");
  if (c->use_meta && tile != 512) return fail(c, FB_ERR_INVALID, "the metadata branch is hard-wired to 512x512 inputs (flair/model.py:59)");
  return 0;
}

int set_raster_common(fb_ctx* c, int bands_total, const int32_t* band_idx, int rc_, int64_t W, int64_t H,
                      int64_t row0, int64_t rows, int layout) {
  if (!c->loaded) return fail(c, FB_ERR_STATE, "fb_load_weights has not been called");
  if (rc_ != c->in_ch) return fail(c, FB_ERR_INVALID, "number of selected bands differs from the model's in_channels");
  if (bands_total <= 0 || W <= 0 || H <= 0 || rows < 0 || row0 < 0 || row0 + rows > H)
    return fail(c, FB_ERR_INVALID, "bad raster geometry");
  if (layout != FB_LAYOUT_CHW && layout != FB_LAYOUT_HWC) return fail(c, FB_ERR_INVALID, "bad raster layout");
  for (int i = 0; i < rc_; ++i)
    if (band_idx[i] < 0 || band_idx[i] >= bands_total) return fail(c, FB_ERR_INVALID, "band index out of range");
  if (!c->band_idx_dev) FB_CUDA(c, cudaMalloc(&c->band_idx_dev, 8 * sizeof(int)));
  int tmp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < rc_; ++i) tmp[i] = band_idx[i];
  FB_CUDA(c, cudaMemcpyAsync(c->band_idx_dev, tmp, sizeof tmp, cudaMemcpyHostToDevice, c->stream));
  c->bands_total = bands_total; c->rc = rc_; c->layout = layout;
  c->W = W; c->H = H; c->row0 = row0; c->rows = rows;
  return 0;
}

}  // namespace

extern "C" {

int fb_api_version(void) { return FB_API_VERSION; }

const char* fb_last_error(const fb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int fb_create(int device, void* cuda_stream, fb_ctx** out) {
  if (!out) return FB_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(nullptr, FB_ERR_NO_DEVICE, "no CUDA device available; libflairb200 has no CPU fallback");
  }
  if (device < 0 || device >= count) return fail(nullptr, FB_ERR_INVALID, "device index out of range");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, FB_ERR_NO_DEVICE, "cudaGetDeviceProperties failed");
  if (prop.major != 10)
    return fail(nullptr, FB_ERR_NO_DEVICE, std::string("device '") + prop.name + "' is not compute capability 10.x; this library is sm_100a only");
  if (cudaSetDevice(device) != cudaSuccess) return fail(nullptr, FB_ERR_NO_DEVICE, "cudaSetDevice failed");
  if (fb::init_tma_encoder() != 0) return fail(nullptr, FB_ERR_NO_DEVICE, "cuTensorMapEncodeTiled not available from the driver");
  fb_ctx* c = new fb_ctx();
  c->device = device;
  c->stream = static_cast<cudaStream_t>(cuda_stream);
  c->num_sms = prop.multiProcessorCount;
  const char* fg = getenv("FB_FORCE_GATHER");
  c->force_gather = fg && fg[0] == '1';
  const char* nh = getenv("FB_NO_HALO");
  c->no_halo = nh && nh[0] == '1';
  const char* np = getenv("FB_NO_PHASE");
  c->no_phase = np && np[0] == '1';
  // default since the halo planes are staged by TMA (which cannot up-sample while it copies): dec2.conv2 writes its
  // output 2x2-replicated again (120 -> 193 us) and dec3.conv1 becomes a plain two-source conv with TMA-staged planes
  // (621 -> 410 us). FB_NO_UP1=0: dec3.conv1 reads the low-res tensor through the cp.async up-sampling gather.
  const char* nu = getenv("FB_NO_UP1");
  c->no_up1 = !(nu && nu[0] == '0');
  const char* fc = getenv("FB_FRONT_CHUNK");
  c->front_chunk = fc ? atoi(fc) : 0;
  const char* ns = getenv("FB_NO_S2D");
  c->no_s2d = ns && ns[0] == '1';
  const char* ft = getenv("FB_FULL_TILES");
  c->full_tiles = ft && ft[0] == '1';
  const char* nf = getenv("FB_NO_FUSED_SINK");
  c->no_fused_sink = nf && nf[0] == '1';
  const char* nd = getenv("FB_NO_D2S");
  c->no_d2s = nd && nd[0] == '1';
  const char* npf = getenv("FB_NO_POOL_FUSE");
  c->no_pool_fuse = npf && npf[0] == '1';
  const char* nnt = getenv("FB_NO_NARROW_TILES");
  c->no_narrow_tiles = nnt && nnt[0] == '1';
  const char* sbt = getenv("FB_SUB_TILES");
  c->sub_tiles = sbt ? atoi(sbt) : 2;
  if (c->sub_tiles < 0 || c->sub_tiles > 2) c->sub_tiles = 2;
  const char* dpr = getenv("FB_DEC_PAIR");
  c->dec_pair = dpr && dpr[0] == '1';
  const char* alt = getenv("FB_ALIGNED_TILES");
  c->aligned_tiles = alt && alt[0] == '1';
  const char* nhp = getenv("FB_NO_HPAIR");
  c->no_hpair = nhp && nhp[0] == '1';
  const char* da = getenv("FB_D2S_ALL");
  c->d2s_all = !(da && da[0] == '0');
  const char* nsb = getenv("FB_NO_SB");
  c->no_sb = nsb && nsb[0] == '1';
  *out = c;
  return 0;
}

void fb_destroy(fb_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->comm) fb::comm_destroy(c->comm);
  for (void* p : c->owned) cudaFree(p);
  if (c->lut) cudaFree(c->lut);
  if (c->raster_own) cudaFree(c->raster_own);
  if (c->maps_own) cudaFree(c->maps_own);
  for (cudaEvent_t e : c->copy_events) cudaEventDestroy(e);
  if (c->h2d_stream) cudaStreamDestroy(c->h2d_stream);
  if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
  if (c->band_idx_dev) cudaFree(c->band_idx_dev);
  if (c->ident_dev) cudaFree(c->ident_dev);
  if (c->arena) cudaFree(c->arena);
  if (c->tile_xy_dev) cudaFree(c->tile_xy_dev);
  if (c->tiles_dev) cudaFree(c->tiles_dev);
  if (c->meta_dev) cudaFree(c->meta_dev);
  if (c->menc_dev) cudaFree(c->menc_dev);
  if (c->list_dev) cudaFree(c->list_dev);
  if (c->win_dev) cudaFree(c->win_dev);
  delete c;
}

int fb_synchronize(fb_ctx* c) {
  if (!c) return FB_ERR_INVALID;
  FB_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

int64_t fb_launch_count(const fb_ctx* c) { return c ? c->launches : 0; }

double fb_flop_count(const fb_ctx* c) { return c ? c->flops : 0.0; }

int fb_logit_stride(const fb_ctx* c) { return c ? c->ls : 0; }

int fb_debug_need_rect(int tile, int layer, int ax0, int ay0, int ax1, int ay1, int32_t* rect4) {
  if (!rect4 || tile <= 0 || tile % 32 != 0 || layer < 0 || layer >= fb::kNeedLayers) return FB_ERR_INVALID;
  const fb::NeedRect r = fb::need_rect(tile, layer, ax0, ay0, ax1, ay1);
  rect4[0] = r.x0; rect4[1] = r.y0; rect4[2] = r.x1; rect4[3] = r.y1;
  return 0;
}

int fb_debug_tile_cover(int tile, int layer, int scale, int th, int tw, int ax0, int ay0, int ax1, int ay1, int32_t* cover4) {
  if (!cover4 || tile <= 0 || tile % 32 != 0 || layer < 0 || layer >= fb::kNeedLayers || (scale != 1 && scale != 2) || th <= 0 || tw <= 0)
    return FB_ERR_INVALID;
  const int S = (layer >= 10 ? tile : (tile / 16) << (layer / 2)) / scale;   // tile-grid extent of the layer
  if (S % th != 0 || S % tw != 0) return FB_ERR_INVALID;
  const fb::NeedRect g = fb::need_on_tile_grid(fb::need_rect(tile, layer, ax0, ay0, ax1, ay1), scale);
  const fb::NeedSpan sx = fb::need_span(g.x0, g.x1, tw, S), sy = fb::need_span(g.y0, g.y1, th, S);
  cover4[0] = sx.o; cover4[1] = sy.o; cover4[2] = sx.n; cover4[3] = sy.n;
  return 0;
}

int fb_load_weights(fb_ctx* c, const fb_tensor_desc* tensors, int n_tensors, int in_channels,
                    int n_classes, int use_metadata) {
  if (!c || !tensors || n_tensors <= 0) return FB_ERR_INVALID;
  if (in_channels < 1 || in_channels > 8) return fail(c, FB_ERR_INVALID, "in_channels must be in 1..8");
  if (n_classes < 1 || n_classes > 32) return fail(c, FB_ERR_INVALID, "n_classes must be in 1..32");
  FB_CUDA(c, cudaSetDevice(c->device));
  TensorMap tm;
  for (int i = 0; i < n_tensors; ++i)
    if (tensors[i].name && tensors[i].data) tm[tensors[i].name] = &tensors[i];
  // a reload on a live context: nothing queued may still read the old weights when they are freed
  FB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->h2d_stream) FB_CUDA(c, cudaStreamSynchronize(c->h2d_stream));
  if (c->d2h_stream) FB_CUDA(c, cudaStreamSynchronize(c->d2h_stream));
  for (void* p : c->owned) cudaFree(p);
  c->owned.clear();
  c->conv.clear();
  for (float*& m : c->mlp) m = nullptr;
  c->loaded = false;
  c->use_meta = 0;
  c->arena_n = 0;   // the arena plan depends on the model (input layout of the stem): re-plan on the next pass

  FB_TRY(build_conv(c, tm, "stem", "encoder.conv1.weight", "encoder.bn1", "", in_channels, 64, 7, 2, 3));
  int cin = 64;
  for (int st = 0; st < 4; ++st) {
    const int C = kStageCh[st];
    for (int b = 0; b < kStageBlocks[st]; ++b) {
      char pre[96], nm[64];
      snprintf(pre, sizeof pre, "encoder.layer%d.%d", st + 1, b);
      snprintf(nm, sizeof nm, "layer%d.%d", st + 1, b);
      const std::string P(pre), N(nm);
      const int s = (st > 0 && b == 0) ? 2 : 1;
      FB_TRY(build_conv(c, tm, N + ".conv1", P + ".conv1.weight", P + ".bn1", "", cin, C, 3, s, 1));
      FB_TRY(build_conv(c, tm, N + ".conv2", P + ".conv2.weight", P + ".bn2", "", C, C, 3, 1, 1));
      if (st > 0 && b == 0)
        FB_TRY(build_conv(c, tm, N + ".downsample", P + ".downsample.0.weight", P + ".downsample.1", "", cin, C, 1, 2, 0));
      cin = C;
    }
  }
  const int dec_in[5] = {512 + 256, 256 + 128, 128 + 64, 64 + 64, 32};
  const int dec_skip[5] = {256, 128, 64, 64, 0};
  for (int d = 0; d < 5; ++d) {
    char pre[96], nm[64];
    snprintf(pre, sizeof pre, "decoder.blocks.%d", d);
    snprintf(nm, sizeof nm, "dec%d", d);
    const std::string P(pre), N(nm);
    FB_TRY(build_conv(c, tm, N + ".conv1", P + ".conv1.0.weight", P + ".conv1.1", "", dec_in[d], kDecOut[d], 3, 1, 1, dec_skip[d], true));
    FB_TRY(build_conv(c, tm, N + ".conv2", P + ".conv2.0.weight", P + ".conv2.1", "", kDecOut[d], kDecOut[d], 3, 1, 1));
  }
  FB_TRY(build_conv(c, tm, "head", "segmentation_head.0.weight", "", "segmentation_head.0.bias", 16, n_classes, 3, 1, 1));
  if (use_metadata) {
    FB_TRY(upload_f32(c, tm, "enc.enc_mlp.0.weight", 64 * 45, &c->mlp[0]));
    FB_TRY(upload_f32(c, tm, "enc.enc_mlp.0.bias", 64, &c->mlp[1]));
    FB_TRY(upload_f32(c, tm, "enc.enc_mlp.3.weight", 32 * 64, &c->mlp[2]));
    FB_TRY(upload_f32(c, tm, "enc.enc_mlp.3.bias", 32, &c->mlp[3]));
    FB_TRY(upload_f32(c, tm, "enc.enc_mlp.6.weight", 16 * 32, &c->mlp[4]));
    FB_TRY(upload_f32(c, tm, "enc.enc_mlp.6.bias", 16, &c->mlp[5]));
  }
  c->in_ch = in_channels;
  c->ncls = n_classes;
  c->ls = n_classes <= 16 ? 16 : 32;
  c->use_meta = use_metadata ? 1 : 0;
  c->loaded = true;
  return 0;
}

int fb_set_norm(fb_ctx* c, int mode, const double* mean, const double* std, int nc) {
  if (!c) return FB_ERR_INVALID;
  if (nc < 1 || nc > 8) return fail(c, FB_ERR_INVALID, "norm: channel count must be in 1..8");
  if (mode != FB_NORM_CUSTOM && mode != FB_NORM_SCALING && mode != FB_NORM_WITHOUT)
    return fail(c, FB_ERR_INVALID, "norm: unknown mode");
  if (mode == FB_NORM_CUSTOM && (!mean || !std)) return fail(c, FB_ERR_INVALID, "norm: custom mode needs means and stds");
  std::vector<uint16_t> lut(8 * 256, 0);
  for (int ch = 0; ch < nc; ++ch)
    for (int v = 0; v < 256; ++v) {
      double d;
      if (mode == FB_NORM_CUSTOM) d = (static_cast<double>(v) - mean[ch]) / std[ch];
      else if (mode == FB_NORM_SCALING) d = static_cast<double>(v) / 255.0;
      else d = static_cast<double>(v);
      lut[ch * 256 + v] = f32_to_bf16_rne(static_cast<float>(d));  // float64 -> float32 -> bf16
    }
  FB_CUDA(c, cudaSetDevice(c->device));
  if (!c->lut) FB_CUDA(c, cudaMalloc(&c->lut, lut.size() * 2));
  FB_CUDA(c, cudaMemcpyAsync(c->lut, lut.data(), lut.size() * 2, cudaMemcpyHostToDevice, c->stream));
  FB_CUDA(c, cudaStreamSynchronize(c->stream));
  c->norm_set = true;
  return 0;
}

int fb_set_raster(fb_ctx* c, const uint8_t* dev_raster, int bands_total, const int32_t* band_idx, int nc,
                  int64_t W, int64_t H, int64_t row0, int64_t rows, int layout) {
  if (!c || !dev_raster || !band_idx) return FB_ERR_INVALID;
  FB_TRY(set_raster_common(c, bands_total, band_idx, nc, W, H, row0, rows, layout));
  c->raster = dev_raster;
  return 0;
}

int fb_upload_raster(fb_ctx* c, const uint8_t* host_raster, int bands_total, const int32_t* band_idx, int nc,
                     int64_t W, int64_t H, int64_t row0, int64_t rows, int layout) {
  if (!c || !host_raster || !band_idx) return FB_ERR_INVALID;
  FB_TRY(set_raster_common(c, bands_total, band_idx, nc, W, H, row0, rows, layout));
  const size_t bytes = static_cast<size_t>(bands_total) * rows * W;
  if (bytes > c->raster_own_bytes) {
    FB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->raster_own) cudaFree(c->raster_own);
    c->raster_own = nullptr; c->raster_own_bytes = 0;
    if (cudaMalloc(&c->raster_own, bytes) != cudaSuccess) {
      cudaGetLastError();
      return fail(c, FB_ERR_OOM, "raster upload: cudaMalloc failed");
    }
    c->raster_own_bytes = bytes;
  }
  FB_CUDA(c, cudaMemcpyAsync(c->raster_own, host_raster, bytes, cudaMemcpyHostToDevice, c->stream));
  c->raster = c->raster_own;
  return 0;
}

int fb_forward_tiles(fb_ctx* c, const int32_t* tile_xy, int n, int tile, const float* metadata,
                     float* logits_dev) {
  FB_TRY(check_ready(c, true, tile));
  if (!tile_xy || n <= 0) return fail(c, FB_ERR_INVALID, "forward: no tiles");
  FB_CUDA(c, cudaSetDevice(c->device));
  FB_TRY(ensure_arena(c, n, tile));
  FB_TRY(ensure_tile_buffers(c, n));
  FB_CUDA(c, cudaMemcpyAsync(c->tile_xy_dev, tile_xy, static_cast<size_t>(n) * 2 * sizeof(int),
                             cudaMemcpyHostToDevice, c->stream));
  const float* menc = nullptr;
  FB_TRY(run_metadata(c, metadata, n, &menc));
  FB_TRY(run_extract(c, c->raster, c->layout, c->bands_total, c->band_idx_dev, c->rc, c->W, c->H, c->row0,
                     c->rows, c->tile_xy_dev, n, tile));
  FB_TRY(run_network(c, n, tile, menc));
  if (logits_dev)
    FB_CUDA(c, cudaMemcpyAsync(logits_dev, c->acts["logits"].ptr, static_cast<size_t>(n) * tile * tile * c->ls * 4,
                               cudaMemcpyDeviceToDevice, c->stream));
  return 0;
}

}  // extern "C" (the sink helpers below are C++)

namespace {
// where the logits of a batch go: the class map (argmax), the probability planes, or the blend accumulators
struct DetectSink {
  int kind = 0;                 // 0 argmax + confidence, 1 class_prob planes, 2 blend accumulation
  uint8_t* cls = nullptr;
  uint8_t* conf = nullptr;
  uint8_t* prob = nullptr;
  float* acc = nullptr;
  float* wsum = nullptr;
  int method = 0;
  int64_t map_w = 0, map_row0 = 0, map_rows = 0;
  int seq0 = 0;
  // kind 0 only: per-tile confusion matrices of each tile's own prediction over its metric window
  const fb_tile* windows = nullptr;
  const uint8_t* truth = nullptr;
  int truth_sub = 0;
  int64_t* tile_cm = nullptr;
  // called on the host right before the first launch / right after the last launch of the batch of tiles
  // [i0, i0 + nb) (fb_detect_zone_host hangs its copy-stream dependencies here)
  std::function<int(int, int)> before_batch, after_batch;
};

int check_write_rects(fb_ctx* c, const fb_tile* tiles, int n, int tile, int64_t map_w, int64_t map_row0) {
  for (int i = 0; i < n; ++i) {
    const fb_tile& t = tiles[i];
    if (t.wx1 > t.wx0 && t.wy1 > t.wy0 &&
        (t.wx0 < t.x0 || t.wy0 < t.y0 || t.wx1 > t.x0 + tile || t.wy1 > t.y0 + tile || t.wx0 < 0 || t.wx1 > map_w || t.wy0 < map_row0))
      return fail(c, FB_ERR_INVALID, "detect: write rectangle of tile " + std::to_string(i) + " is outside its tile or the map");
  }
  return 0;
}

int detect_loop(fb_ctx* c, const fb_tile* tiles, int n, int tile, int batch, const DetectSink& s);
}  // namespace

extern "C" {

int fb_detect_strip(fb_ctx* c, const fb_tile* tiles, int n, int tile, int batch, uint8_t* cls_map_dev,
                    uint8_t* conf_map_dev, int64_t map_w, int64_t map_row0) {
  FB_TRY(check_ready(c, true, tile));
  if (!tiles || n < 0 || batch <= 0 || !cls_map_dev) return fail(c, FB_ERR_INVALID, "detect: bad arguments");
  FB_TRY(check_write_rects(c, tiles, n, tile, map_w, map_row0));
  DetectSink s;
  s.kind = 0; s.cls = cls_map_dev; s.conf = conf_map_dev; s.map_w = map_w; s.map_row0 = map_row0;
  return detect_loop(c, tiles, n, tile, batch, s);
}

int fb_detect_strip_metrics(fb_ctx* c, const fb_tile* tiles, const fb_tile* windows, int n, int tile, int batch,
                            uint8_t* cls_map_dev, uint8_t* conf_map_dev, int64_t map_w, int64_t map_row0,
                            const uint8_t* truth_dev, int truth_sub, int64_t* cm_tiles_dev) {
  FB_TRY(check_ready(c, true, tile));
  if (!tiles || !windows || n < 0 || batch <= 0 || !cls_map_dev || !truth_dev || !cm_tiles_dev)
    return fail(c, FB_ERR_INVALID, "detect (metrics): bad arguments");
  FB_TRY(check_write_rects(c, tiles, n, tile, map_w, map_row0));
  FB_TRY(check_write_rects(c, windows, n, tile, map_w, map_row0));
  for (int i = 0; i < n; ++i)
    if (windows[i].x0 != tiles[i].x0 || windows[i].y0 != tiles[i].y0)
      return fail(c, FB_ERR_INVALID, "detect (metrics): window " + std::to_string(i) + " belongs to another tile origin");
  DetectSink s;
  s.kind = 0; s.cls = cls_map_dev; s.conf = conf_map_dev; s.map_w = map_w; s.map_row0 = map_row0;
  s.windows = windows; s.truth = truth_dev; s.truth_sub = truth_sub; s.tile_cm = cm_tiles_dev;
  return detect_loop(c, tiles, n, tile, batch, s);
}

int fb_detect_strip_prob(fb_ctx* c, const fb_tile* tiles, int n, int tile, int batch, uint8_t* prob_map_dev,
                         int64_t map_w, int64_t map_row0, int64_t map_rows) {
  FB_TRY(check_ready(c, true, tile));
  if (!tiles || n < 0 || batch <= 0 || !prob_map_dev || map_rows <= 0) return fail(c, FB_ERR_INVALID, "detect (class_prob): bad arguments");
  FB_TRY(check_write_rects(c, tiles, n, tile, map_w, map_row0));
  for (int i = 0; i < n; ++i)
    if (tiles[i].wy1 > tiles[i].wy0 && tiles[i].wy1 > map_row0 + map_rows)
      return fail(c, FB_ERR_INVALID, "detect (class_prob): write rectangle below the map");
  DetectSink s;
  s.kind = 1; s.prob = prob_map_dev; s.map_w = map_w; s.map_row0 = map_row0; s.map_rows = map_rows;
  return detect_loop(c, tiles, n, tile, batch, s);
}

int fb_blend_strip(fb_ctx* c, const fb_tile* tiles, int n, int tile, int batch, int method, float* acc_dev,
                   float* wsum_dev, int64_t map_w, int64_t map_row0, int64_t map_rows, int tile_seq0) {
  FB_TRY(check_ready(c, true, tile));
  if (!tiles || n < 0 || batch <= 0 || !acc_dev || map_rows <= 0 || map_w <= 0 || method < 0 || method > 2 ||
      (method != 2 && !wsum_dev))
    return fail(c, FB_ERR_INVALID, "blend: bad arguments");
  DetectSink s;
  s.kind = 2; s.acc = acc_dev; s.wsum = wsum_dev; s.method = method; s.map_w = map_w; s.map_row0 = map_row0;
  s.map_rows = map_rows; s.seq0 = tile_seq0;
  return detect_loop(c, tiles, n, tile, batch, s);
}

int fb_blend_finalize(fb_ctx* c, const float* acc_dev, const float* wsum_dev, int method, int64_t npx,
                      uint8_t* cls_map_dev, uint8_t* conf_map_dev) {
  if (!c) return FB_ERR_INVALID;
  if (!c->loaded) return fail(c, FB_ERR_STATE, "blend: weights not loaded");
  if (!acc_dev || !cls_map_dev || npx < 0 || method < 0 || method > 2 || (method != 2 && !wsum_dev))
    return fail(c, FB_ERR_INVALID, "blend finalize: bad arguments");
  FB_CUDA(c, cudaSetDevice(c->device));
  int rc = fb::launch_blend_finalize(acc_dev, wsum_dev, method, c->ncls, c->ls, npx, cls_map_dev, conf_map_dev, c->num_sms, c->stream);
  if (rc) return fail(c, rc, "blend finalize launch failed");
  c->launches++;
  return 0;
}

}  // extern "C"

namespace {
int detect_loop(fb_ctx* c, const fb_tile* tiles, int n, int tile, int batch, const DetectSink& s) {
  if (c->use_meta) return fail(c, FB_ERR_INVALID, "zone detection does not take metadata (zone_detect/model.py:52)");
  if (n == 0) return 0;
  FB_CUDA(c, cudaSetDevice(c->device));
  if (batch > n) batch = n;
  if (batch > 1024) batch = 1024;   // images per launch the tile-list builder handles (elementwise.cu)
  FB_TRY(ensure_arena(c, batch, tile));
  FB_TRY(ensure_tile_buffers(c, n));
  std::vector<int> xy(static_cast<size_t>(n) * 2);
  for (int i = 0; i < n; ++i) { xy[2 * i] = tiles[i].x0; xy[2 * i + 1] = tiles[i].y0; }
  FB_CUDA(c, cudaMemcpyAsync(c->tile_xy_dev, xy.data(), xy.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  FB_CUDA(c, cudaMemcpyAsync(c->tiles_dev, tiles, static_cast<size_t>(n) * sizeof(fb_tile), cudaMemcpyHostToDevice, c->stream));
  if (s.tile_cm) {
    // the metric windows may reach outside the write rectangles (clamped last row / column): whole tiles, fp32 logits
    if (n > c->win_cap) {
      FB_CUDA(c, cudaStreamSynchronize(c->stream));
      if (c->win_dev) cudaFree(c->win_dev);
      c->win_dev = nullptr; c->win_cap = 0;
      FB_CUDA(c, cudaMalloc(&c->win_dev, static_cast<size_t>(n) * sizeof(fb_tile)));
      c->win_cap = n;
    }
    FB_CUDA(c, cudaMemcpyAsync(c->win_dev, s.windows, static_cast<size_t>(n) * sizeof(fb_tile), cudaMemcpyHostToDevice, c->stream));
  }
  FB_CUDA(c, cudaStreamSynchronize(c->stream));  // xy is a local
  for (int i0 = 0; i0 < n; i0 += batch) {
    const int nb = (n - i0 < batch) ? n - i0 : batch;
    if (nb != c->arena_n) FB_TRY(ensure_arena(c, nb, tile));  // ragged last batch: re-plan inside the same block
    if (s.before_batch) FB_TRY(s.before_batch(i0, nb));
    FB_TRY(run_extract(c, c->raster, c->layout, c->bands_total, c->band_idx_dev, c->rc, c->W, c->H, c->row0,
                       c->rows, c->tile_xy_dev + 2 * i0, nb, tile));
    // exact clipping (kinds 0 and 1) only ever reads the logits inside the write rectangles: the decoder computes
    // nothing else (tile_need.cuh), and for the class map the head's epilogue writes the bytes itself
    NeedCtx need;
    need.tiles_dev = c->tiles_dev + 6 * i0;
    need.tiles_host = &tiles[i0].x0;
    need.n = nb; need.T = tile;
    need.restrict_tiles = s.kind != 2 && !c->full_tiles && !s.tile_cm;
    HeadSink hs;
    hs.cls = s.cls; hs.conf = s.conf; hs.map_w = s.map_w; hs.map_row0 = s.map_row0;
    bool sunk = false;
    FB_TRY(run_network(c, nb, tile, nullptr, &need, (s.kind == 0 && !c->no_fused_sink && !s.tile_cm) ? &hs : nullptr, &sunk));
    if (sunk) {
      if (s.after_batch) FB_TRY(s.after_batch(i0, nb));
      continue;
    }
    ProfScope ps(c, 3);
    const float* logits = static_cast<const float*>(c->acts["logits"].ptr);
    int rc;
    if (s.kind == 0)
      rc = fb::launch_argmax_stitch(logits, c->ncls, c->ls, nb, tile, c->tiles_dev + 6 * i0, s.cls, s.conf, s.map_w, s.map_row0, c->stream);
    else if (s.kind == 1)
      rc = fb::launch_prob_stitch(logits, c->ncls, c->ls, nb, tile, c->tiles_dev + 6 * i0, s.prob, s.map_w, s.map_row0, s.map_rows,
                                  c->stream);
    else
      rc = fb::launch_blend_accumulate(logits, c->ncls, c->ls, nb, tile, c->tiles_dev + 6 * i0, s.method, s.acc, s.wsum, s.map_w,
                                       s.map_row0, s.map_rows, c->W, c->H, s.seq0 + i0, c->stream);
    if (rc) return fail(c, rc, "stitch launch failed");
    c->launches++;
    if (s.tile_cm) {
      rc = fb::launch_tile_confusion(logits, c->ncls, c->ls, nb, tile, c->win_dev + 6 * i0, s.truth, s.truth_sub, s.map_w, s.map_row0,
                                     reinterpret_cast<long long*>(s.tile_cm) + static_cast<long long>(i0) * c->ncls * c->ncls, c->stream);
      if (rc) return fail(c, rc, "tile confusion launch failed");
      c->launches++;
    }
    if (s.after_batch) FB_TRY(s.after_batch(i0, nb));
  }
  return 0;
}

// What fb_detect_zone_shard adds to fb_detect_zone_host.
struct ZoneHostOpts {
  bool owned_only = false;              // D2H only the write rectangles of `tiles` (the host maps may be shared between ranks)
  const uint8_t* host_truth = nullptr;  // optional: truth rows [truth_row0, ...) with pitch map_w -> fused confusion matrix
  int64_t truth_row0 = 0;
  int truth_sub = 0, ncls_cm = 0;
  int64_t* cm_dev = nullptr;
};

// One maximal run of tiles of the y-sorted table that share their written rows and own adjacent columns: the unit
// in which a shard's part of the class map goes back to the host (and is scored against the truth).
struct RowRect {
  int64_t y0, y1, x0, x1;
  int last;   // index (in the sorted table) of the last tile that writes into it
};

// Host buffers in, host buffers out, software-pipelined: the raster goes up in row chunks on a copy stream while
// the compute stream already works on the tile rows whose pixels have landed, and every finished piece of the
// class map is sent back on a second copy stream as soon as no remaining tile can write it. The tiles are
// processed by rows (sorted by y0) for that; their write rectangles are disjoint by contract (fb_tile), so the
// order in which they run does not change a byte of the result.
int zone_host_impl(fb_ctx* c, const uint8_t* host_raster, int bands_total, const int32_t* band_idx, int nc, int64_t W,
                   int64_t H, int64_t row0, int64_t rows, int layout, const fb_tile* tiles, int n, int tile, int batch,
                   uint8_t* host_cls, uint8_t* host_conf, int64_t map_w, int64_t map_row0, int64_t map_rows,
                   const ZoneHostOpts& o) {
  if (!c || !host_cls || !host_raster || !band_idx || map_w <= 0 || map_rows <= 0 || (n > 0 && !tiles) || n < 0 || batch <= 0)
    return FB_ERR_INVALID;
  FB_TRY(check_ready(c, false, tile));
  FB_TRY(set_raster_common(c, bands_total, band_idx, nc, W, H, row0, rows, layout));
  FB_TRY(check_write_rects(c, tiles, n, tile, map_w, map_row0));
  for (int i = 0; i < n; ++i)
    if (tiles[i].wy1 > tiles[i].wy0 && tiles[i].wy1 > map_row0 + map_rows)
      return fail(c, FB_ERR_INVALID, "detect: write rectangle below the map");
  const bool score = o.host_truth != nullptr;
  if (score && (!o.owned_only || !o.cm_dev || o.ncls_cm < 1 || o.ncls_cm > 32 || o.truth_row0 < 0))
    return fail(c, FB_ERR_INVALID, "detect (shard): bad truth / confusion-matrix arguments");
  FB_CUDA(c, cudaSetDevice(c->device));

  // tiles by rows (stable: the columns of one tile row keep their write order, i.e. ascending x)
  std::vector<fb_tile> sorted(tiles, tiles + n);
  std::stable_sort(sorted.begin(), sorted.end(), [](const fb_tile& a, const fb_tile& b) { return a.y0 < b.y0; });

  // device maps: the rows of the host maps (whole-map mode) or just the rows this shard writes (owned-only mode)
  int64_t dev_row0 = map_row0, dev_rows = map_rows;
  std::vector<RowRect> rects;
  if (o.owned_only) {
    int64_t lo = INT64_MAX, hi = INT64_MIN;
    for (int i = 0; i < n; ++i) {
      const fb_tile& t = sorted[i];
      if (t.wx1 <= t.wx0 || t.wy1 <= t.wy0) continue;
      lo = std::min<int64_t>(lo, t.wy0);
      hi = std::max<int64_t>(hi, t.wy1);
      if (!rects.empty() && rects.back().y0 == t.wy0 && rects.back().y1 == t.wy1 && rects.back().x1 == t.wx0) {
        rects.back().x1 = t.wx1;
        rects.back().last = i;
      } else {
        rects.push_back(RowRect{t.wy0, t.wy1, t.wx0, t.wx1, i});
      }
    }
    if (rects.empty()) return 0;   // this shard writes nothing
    dev_row0 = lo;
    dev_rows = hi - lo;
    if (score)
      for (const RowRect& r : rects)
        if (r.y0 < o.truth_row0) return fail(c, FB_ERR_INVALID, "detect (shard): truth rows start below a write rectangle");
  }
  const size_t rbytes = static_cast<size_t>(bands_total) * rows * W;
  const size_t mbytes = static_cast<size_t>(map_w) * dev_rows;
  const size_t nmaps = 2 + (score ? 1 : 0);
  // context-owned, grow-only buffers (a cudaMalloc/cudaFree pair per call would serialise the device)
  if (rbytes > c->raster_own_bytes) {
    FB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->raster_own) cudaFree(c->raster_own);
    c->raster_own = nullptr; c->raster_own_bytes = 0;
    if (cudaMalloc(&c->raster_own, rbytes) != cudaSuccess) {
      cudaGetLastError();
      return fail(c, FB_ERR_OOM, "raster upload: cudaMalloc failed");
    }
    c->raster_own_bytes = rbytes;
  }
  if (mbytes * nmaps > c->maps_own_bytes) {
    FB_CUDA(c, cudaStreamSynchronize(c->stream));
    if (c->maps_own) cudaFree(c->maps_own);
    c->maps_own = nullptr; c->maps_own_bytes = 0;
    if (cudaMalloc(&c->maps_own, mbytes * nmaps) != cudaSuccess) {
      cudaGetLastError();
      return fail(c, FB_ERR_OOM, "class map: cudaMalloc failed");
    }
    c->maps_own_bytes = mbytes * nmaps;
  }
  if (!c->h2d_stream) FB_CUDA(c, cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
  if (!c->d2h_stream) FB_CUDA(c, cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
  c->raster = c->raster_own;
  uint8_t* maps = c->maps_own;
  uint8_t* conf_dev = host_conf ? maps + mbytes : nullptr;
  uint8_t* truth_dev = score ? maps + 2 * mbytes : nullptr;

  // whole-map mode: suffix minimum of the first written row -- after tile i, map rows below low_after[i + 1] are final
  std::vector<int64_t> low_after;
  if (!o.owned_only) {
    low_after.assign(static_cast<size_t>(n) + 1, map_row0 + map_rows);
    for (int i = n - 1; i >= 0; --i) {
      const bool writes = sorted[i].wx1 > sorted[i].wx0 && sorted[i].wy1 > sorted[i].wy0;
      low_after[i] = writes ? std::min<int64_t>(low_after[i + 1], sorted[i].wy0) : low_after[i + 1];
    }
  }

  // upload: chunks of raster rows, one event per chunk. Work already queued on the compute stream (an earlier
  // call still reading the raster buffer) must finish before the first byte is overwritten.
  const int64_t chunk_rows = 512;
  const int nchunks = static_cast<int>((rows + chunk_rows - 1) / chunk_rows);
  const size_t nbatches = static_cast<size_t>((n + batch - 1) / batch);
  const size_t nev = static_cast<size_t>(nchunks) + nbatches + 1 + (score ? rects.size() : 0);
  while (c->copy_events.size() < nev) {
    cudaEvent_t ev;
    FB_CUDA(c, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    c->copy_events.push_back(ev);
  }
  cudaEvent_t* ev_up = c->copy_events.data();
  cudaEvent_t* ev_batch = c->copy_events.data() + nchunks;
  cudaEvent_t ev_start = c->copy_events[static_cast<size_t>(nchunks) + nbatches];
  cudaEvent_t* ev_truth = c->copy_events.data() + nchunks + nbatches + 1;
  if (!o.owned_only) FB_CUDA(c, cudaMemsetAsync(maps, 0, mbytes * (host_conf ? 2 : 1), c->stream));
  FB_CUDA(c, cudaEventRecord(ev_start, c->stream));
  FB_CUDA(c, cudaStreamWaitEvent(c->h2d_stream, ev_start, 0));
  FB_CUDA(c, cudaStreamWaitEvent(c->d2h_stream, ev_start, 0));
  size_t truth_sent = 0;   // rects[0 .. truth_sent) have their truth pixels on the way
  auto send_truth = [&](int64_t upto_row) -> int {   // truth of every rectangle that starts above raster row `upto_row`
    while (score && truth_sent < rects.size() && rects[truth_sent].y0 < upto_row) {
      const RowRect& r = rects[truth_sent];
      FB_CUDA(c, cudaMemcpy2DAsync(truth_dev + static_cast<size_t>(r.y0 - dev_row0) * map_w + r.x0, static_cast<size_t>(map_w),
                                   o.host_truth + static_cast<size_t>(r.y0 - o.truth_row0) * map_w + r.x0, static_cast<size_t>(map_w),
                                   static_cast<size_t>(r.x1 - r.x0), static_cast<size_t>(r.y1 - r.y0), cudaMemcpyHostToDevice,
                                   c->h2d_stream));
      FB_CUDA(c, cudaEventRecord(ev_truth[truth_sent], c->h2d_stream));
      ++truth_sent;
    }
    return 0;
  };
  for (int k = 0; k < nchunks; ++k) {
    const int64_t r0 = k * chunk_rows, nr = std::min<int64_t>(chunk_rows, rows - r0);
    if (layout == FB_LAYOUT_HWC) {
      const size_t off = static_cast<size_t>(r0) * W * bands_total;
      FB_CUDA(c, cudaMemcpyAsync(c->raster_own + off, host_raster + off, static_cast<size_t>(nr) * W * bands_total,
                                 cudaMemcpyHostToDevice, c->h2d_stream));
    } else {
      for (int b = 0; b < bands_total; ++b) {
        const size_t off = (static_cast<size_t>(b) * rows + r0) * W;
        FB_CUDA(c, cudaMemcpyAsync(c->raster_own + off, host_raster + off, static_cast<size_t>(nr) * W,
                                   cudaMemcpyHostToDevice, c->h2d_stream));
      }
    }
    FB_CUDA(c, cudaEventRecord(ev_up[k], c->h2d_stream));
    FB_TRY(send_truth(row0 + r0 + nr));   // truth rows travel right behind the raster rows of the same place
  }
  FB_TRY(send_truth(INT64_MAX));

  int64_t sent = map_row0;   // whole-map mode: map rows [map_row0, sent) are already on their way to the host
  auto send_rows = [&](int64_t upto) -> int {
    if (upto <= sent) return 0;
    const size_t off = static_cast<size_t>(sent - map_row0) * map_w, len = static_cast<size_t>(upto - sent) * map_w;
    FB_CUDA(c, cudaMemcpyAsync(host_cls + off, maps + off, len, cudaMemcpyDeviceToHost, c->d2h_stream));
    if (host_conf) FB_CUDA(c, cudaMemcpyAsync(host_conf + off, conf_dev + off, len, cudaMemcpyDeviceToHost, c->d2h_stream));
    sent = upto;
    return 0;
  };
  size_t rect_sent = 0;      // owned-only mode: rects[0 .. rect_sent) are already on their way
  auto finish_rects = [&](int done_tiles, cudaEvent_t ev_done) -> int {
    // (a) score the finished rectangles on the compute stream, (b) copy them out behind ev_done on the copy stream
    size_t upto = rect_sent;
    while (upto < rects.size() && rects[upto].last < done_tiles) ++upto;
    for (size_t k = rect_sent; k < upto && score; ++k) {
      const RowRect& r = rects[k];
      FB_CUDA(c, cudaStreamWaitEvent(c->stream, ev_truth[k], 0));
      const size_t doff = static_cast<size_t>(r.y0 - dev_row0) * map_w + r.x0;
      const int rc = fb::launch_confusion_rect(maps + doff, truth_dev + doff, r.y1 - r.y0, r.x1 - r.x0, map_w, map_w, o.ncls_cm,
                                               o.truth_sub & 0xFF, reinterpret_cast<long long*>(o.cm_dev), c->num_sms, c->stream);
      if (rc) return fail(c, rc, "confusion launch failed");
      c->launches++;
    }
    for (size_t k = rect_sent; k < upto; ++k) {
      const RowRect& r = rects[k];
      const size_t doff = static_cast<size_t>(r.y0 - dev_row0) * map_w + r.x0;
      const size_t hoff = static_cast<size_t>(r.y0 - map_row0) * map_w + r.x0;
      const size_t w = static_cast<size_t>(r.x1 - r.x0), h = static_cast<size_t>(r.y1 - r.y0);
      if (k == rect_sent) FB_CUDA(c, cudaStreamWaitEvent(c->d2h_stream, ev_done, 0));
      if (w == static_cast<size_t>(map_w)) {   // full rows: one linear copy
        FB_CUDA(c, cudaMemcpyAsync(host_cls + hoff, maps + doff, w * h, cudaMemcpyDeviceToHost, c->d2h_stream));
        if (host_conf) FB_CUDA(c, cudaMemcpyAsync(host_conf + hoff, conf_dev + doff, w * h, cudaMemcpyDeviceToHost, c->d2h_stream));
      } else {
        FB_CUDA(c, cudaMemcpy2DAsync(host_cls + hoff, static_cast<size_t>(map_w), maps + doff, static_cast<size_t>(map_w), w, h,
                                     cudaMemcpyDeviceToHost, c->d2h_stream));
        if (host_conf)
          FB_CUDA(c, cudaMemcpy2DAsync(host_conf + hoff, static_cast<size_t>(map_w), conf_dev + doff, static_cast<size_t>(map_w), w, h,
                                       cudaMemcpyDeviceToHost, c->d2h_stream));
      }
    }
    rect_sent = upto;
    return 0;
  };
  DetectSink s;
  s.kind = 0; s.cls = maps; s.conf = conf_dev; s.map_w = map_w; s.map_row0 = dev_row0;
  s.before_batch = [&](int i0, int nb) -> int {
    // the batch reads raster rows up to its lowest tile's bottom edge (the table is sorted by y0)
    int64_t last = static_cast<int64_t>(sorted[i0 + nb - 1].y0) + tile - 1 - row0;
    if (last < 0) return 0;
    if (last > rows - 1) last = rows - 1;
    FB_CUDA(c, cudaStreamWaitEvent(c->stream, ev_up[last / chunk_rows], 0));
    return 0;
  };
  s.after_batch = [&](int i0, int nb) -> int {
    const size_t bi = static_cast<size_t>(i0 / batch);
    FB_CUDA(c, cudaEventRecord(ev_batch[bi], c->stream));
    if (o.owned_only) return finish_rects(i0 + nb, ev_batch[bi]);
    FB_CUDA(c, cudaStreamWaitEvent(c->d2h_stream, ev_batch[bi], 0));
    return send_rows(low_after[i0 + nb]);
  };
  int rc = detect_loop(c, sorted.data(), n, tile, batch, s);
  if (!rc && !o.owned_only) {
    // whatever no tile wrote (or an empty table): the zeroed rows still go back
    cudaError_t e = cudaEventRecord(ev_start, c->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(c->d2h_stream, ev_start, 0);
    if (e != cudaSuccess) rc = cuda_fail(c, e, "class map download");
    if (!rc) rc = send_rows(map_row0 + map_rows);
  }
  // every stream is drained before the host buffers are handed back, also on failure
  cudaError_t e1 = cudaStreamSynchronize(c->h2d_stream);
  cudaError_t e2 = cudaStreamSynchronize(c->d2h_stream);
  cudaError_t e3 = cudaStreamSynchronize(c->stream);
  if (!rc && (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess))
    rc = cuda_fail(c, e1 != cudaSuccess ? e1 : e2 != cudaSuccess ? e2 : e3, "class map download");
  return rc;
}

}  // namespace

extern "C" {

int fb_detect_zone_host(fb_ctx* c, const uint8_t* host_raster, int bands_total, const int32_t* band_idx,
                        int nc, int64_t W, int64_t H, int64_t row0, int64_t rows, int layout,
                        const fb_tile* tiles, int n, int tile, int batch, uint8_t* host_cls,
                        uint8_t* host_conf, int64_t map_w, int64_t map_row0, int64_t map_rows) {
  return zone_host_impl(c, host_raster, bands_total, band_idx, nc, W, H, row0, rows, layout, tiles, n, tile, batch, host_cls,
                        host_conf, map_w, map_row0, map_rows, ZoneHostOpts());
}

int fb_detect_zone_shard(fb_ctx* c, const uint8_t* host_raster, int bands_total, const int32_t* band_idx,
                         int nc, int64_t W, int64_t H, int64_t row0, int64_t rows, int layout,
                         const fb_tile* tiles, int n, int tile, int batch, uint8_t* host_cls,
                         uint8_t* host_conf, int64_t map_w, int64_t map_row0, int64_t map_rows,
                         const uint8_t* host_truth, int64_t truth_row0, int truth_sub, int ncls_cm, int64_t* cm_dev) {
  ZoneHostOpts o;
  o.owned_only = true;
  o.host_truth = host_truth;
  o.truth_row0 = truth_row0;
  o.truth_sub = truth_sub;
  o.ncls_cm = ncls_cm;
  o.cm_dev = cm_dev;
  return zone_host_impl(c, host_raster, bands_total, band_idx, nc, W, H, row0, rows, layout, tiles, n, tile, batch, host_cls,
                        host_conf, map_w, map_row0, map_rows, o);
}

int fb_confusion_rect(fb_ctx* c, const uint8_t* pred_dev, const uint8_t* truth_dev, int64_t rows, int64_t width,
                      int64_t pred_pitch, int64_t truth_pitch, int ncls, int truth_sub, int64_t* cm_dev) {
  if (!c) return FB_ERR_INVALID;
  if (rows == 0 || width == 0) return 0;
  if (!pred_dev || !truth_dev || !cm_dev || rows < 0 || width < 0 || pred_pitch < width || truth_pitch < width)
    return fail(c, FB_ERR_INVALID, "confusion (rect): null pointer, negative size or pitch below the width");
  if (ncls < 1 || ncls > 32) return fail(c, FB_ERR_INVALID, "confusion: ncls must be in 1..32");
  FB_CUDA(c, cudaSetDevice(c->device));
  int rc = fb::launch_confusion_rect(pred_dev, truth_dev, rows, width, pred_pitch, truth_pitch, ncls, truth_sub & 0xFF,
                                     reinterpret_cast<long long*>(cm_dev), c->num_sms, c->stream);
  if (rc) return fail(c, rc, "confusion (rect) launch failed");
  c->launches++;
  return 0;
}

// Page-lock host memory the caller owns (a shared mapping of the class map): the D2H copies of fb_detect_zone_shard
// only overlap with compute when their destination is pinned.
int fb_host_register(void* host_ptr, int64_t bytes) {
  if (!host_ptr || bytes <= 0) return FB_ERR_INVALID;
  const cudaError_t e = cudaHostRegister(host_ptr, static_cast<size_t>(bytes), cudaHostRegisterPortable);
  if (e != cudaSuccess) {
    cudaGetLastError();
    g_create_error = std::string("cudaHostRegister: ") + cudaGetErrorString(e);
    return static_cast<int>(e);
  }
  return 0;
}

int fb_host_unregister(void* host_ptr) {
  if (!host_ptr) return FB_ERR_INVALID;
  const cudaError_t e = cudaHostUnregister(host_ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    g_create_error = std::string("cudaHostUnregister: ") + cudaGetErrorString(e);
    return static_cast<int>(e);
  }
  return 0;
}

// ---- multi-GPU collectives (csrc/comm.cu): one NCCL communicator per context
int fb_comm_unique_id(uint8_t* id128) {
  std::string err;
  const int rc = fb::comm_unique_id(id128, &err);
  if (rc) g_create_error = err;
  return rc;
}

int fb_comm_init(fb_ctx* c, const uint8_t* id128, int rank, int world) {
  if (!c || !id128) return FB_ERR_INVALID;
  if (c->comm) return fail(c, FB_ERR_STATE, "fb_comm_init: this context already has a communicator");
  FB_CUDA(c, cudaSetDevice(c->device));
  std::string err;
  const int rc = fb::comm_init(&c->comm, id128, rank, world, &err);
  return rc ? fail(c, rc, err) : 0;
}

int fb_comm_destroy(fb_ctx* c) {
  if (!c) return FB_ERR_INVALID;
  if (c->comm) {
    cudaStreamSynchronize(c->stream);
    fb::comm_destroy(c->comm);
    c->comm = nullptr;
  }
  return 0;
}

int fb_allreduce_confusion(fb_ctx* c, int64_t* cm_dev, int ncls) {
  if (!c || !cm_dev || ncls < 1 || ncls > 32) return FB_ERR_INVALID;
  if (!c->comm) return fail(c, FB_ERR_STATE, "fb_allreduce_confusion: fb_comm_init has not been called");
  FB_CUDA(c, cudaSetDevice(c->device));
  std::string err;
  const int rc = fb::comm_allreduce_i64(c->comm, reinterpret_cast<long long*>(cm_dev), static_cast<size_t>(ncls) * ncls, c->stream, &err);
  return rc ? fail(c, rc, err) : 0;
}

int fb_gather_bytes(fb_ctx* c, const void* send_dev, int64_t send_bytes, void* recv_dev, const int64_t* offsets,
                    const int64_t* counts, int root) {
  if (!c || !offsets || !counts) return FB_ERR_INVALID;
  if (!c->comm) return fail(c, FB_ERR_STATE, "fb_gather_bytes: fb_comm_init has not been called");
  FB_CUDA(c, cudaSetDevice(c->device));
  std::string err;
  static_assert(sizeof(long long) == sizeof(int64_t), "int64_t is long long here");
  const int rc = fb::comm_gather_bytes(c->comm, send_dev, send_bytes, recv_dev, reinterpret_cast<const long long*>(offsets),
                                       reinterpret_cast<const long long*>(counts), root, c->stream, &err);
  return rc ? fail(c, rc, err) : 0;
}

int fb_predict_patches(fb_ctx* c, const uint8_t* dev_patches, const float* metadata, int n, int tile,
                       int batch, uint8_t* cls_out_dev) {
  FB_TRY(check_ready(c, false, tile));
  if (!dev_patches || !cls_out_dev || n < 0 || batch <= 0) return fail(c, FB_ERR_INVALID, "predict: bad arguments");
  if (n == 0) return 0;
  FB_CUDA(c, cudaSetDevice(c->device));
  if (batch > n) batch = n;
  FB_TRY(ensure_arena(c, batch, tile));
  FB_TRY(ensure_tile_buffers(c, batch));
  if (!c->band_idx_dev) FB_CUDA(c, cudaMalloc(&c->band_idx_dev, 8 * sizeof(int)));
  // every patch is its own little band-planar raster of c bands; the write rectangle is the whole tile
  if (!c->ident_dev) {
    const int ident[8] = {0, 1, 2, 3, 4, 5, 6, 7};
    FB_CUDA(c, cudaMalloc(&c->ident_dev, sizeof ident));
    FB_CUDA(c, cudaMemcpy(c->ident_dev, ident, sizeof ident, cudaMemcpyHostToDevice));
  }
  int* ident_dev = c->ident_dev;
  std::vector<int> xy(static_cast<size_t>(batch) * 2), rect(static_cast<size_t>(batch) * 6);
  for (int i = 0; i < batch; ++i) {
    xy[2 * i] = 0; xy[2 * i + 1] = i * tile;
    rect[6 * i] = 0; rect[6 * i + 1] = i * tile; rect[6 * i + 2] = 0; rect[6 * i + 3] = i * tile;
    rect[6 * i + 4] = tile; rect[6 * i + 5] = (i + 1) * tile;
  }
  FB_CUDA(c, cudaMemcpyAsync(c->tile_xy_dev, xy.data(), xy.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  FB_CUDA(c, cudaMemcpyAsync(c->tiles_dev, rect.data(), rect.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  FB_CUDA(c, cudaStreamSynchronize(c->stream));
  int rc = 0;
  for (int i0 = 0; i0 < n && !rc; i0 += batch) {
    const int nb = (n - i0 < batch) ? n - i0 : batch;
    if (nb != c->arena_n) rc = ensure_arena(c, nb, tile);
    const float* menc = nullptr;
    if (!rc) rc = run_metadata(c, metadata ? metadata + static_cast<size_t>(i0) * FB_METADATA_DIM : nullptr, nb, &menc);
    // patches are stored patch-major ([n][c][T][T]): every patch is its own band-planar raster of T rows, cut at
    // (0, 0); one launch for the batch (tile_stride = bytes per patch)
    if (!rc) {
      const uint8_t* first = dev_patches + static_cast<size_t>(i0) * c->in_ch * tile * tile;
      const long long stride = static_cast<long long>(c->in_ch) * tile * tile;
      __nv_bfloat16* x0 = static_cast<__nv_bfloat16*>(c->acts["x0"].ptr);
      ProfScope ps(c, 0);
      rc = c->stem_s2d ? fb::launch_extract_normalise_s2d(first, FB_LAYOUT_CHW, c->in_ch, ident_dev, c->in_ch, tile, tile, 0, tile,
                                                          c->tile_xy_dev /* (0,0) */, nb, tile, c->lut, x0, c->num_sms, c->stream, stride)
                       : fb::launch_extract_normalise(first, FB_LAYOUT_CHW, c->in_ch, ident_dev, c->in_ch, tile, tile, 0, tile,
                                                      c->tile_xy_dev /* (0,0) */, nb, tile, c->lut, x0, c->num_sms, c->stream, stride);
      if (rc) rc = fail(c, rc, "extract launch failed");
      c->launches++;
    }
    NeedCtx need;
    need.tiles_dev = c->tiles_dev;
    need.tiles_host = rect.data();
    need.n = nb; need.T = tile;
    HeadSink hs;
    hs.cls = cls_out_dev + static_cast<size_t>(i0) * tile * tile;
    hs.map_w = tile;
    bool sunk = false;
    if (!rc) rc = run_network(c, nb, tile, menc, &need, c->no_fused_sink ? nullptr : &hs, &sunk);
    if (!rc && !sunk) {
      ProfScope ps(c, 3);
      rc = fb::launch_argmax_stitch(static_cast<const float*>(c->acts["logits"].ptr), c->ncls, c->ls, nb, tile, c->tiles_dev,
                                    cls_out_dev + static_cast<size_t>(i0) * tile * tile, nullptr, tile, 0, c->stream);
      if (rc) rc = fail(c, rc, "argmax launch failed");
      c->launches++;
    }
  }
  cudaStreamSynchronize(c->stream);   // rect / xy are locals, and the caller reads cls_out next
  return rc;
}

int fb_confusion(fb_ctx* c, const uint8_t* pred_dev, const uint8_t* truth_dev, int64_t npx, int ncls,
                 int truth_sub, int64_t* cm_dev) {
  if (!c) return FB_ERR_INVALID;
  if (npx == 0) return 0;
  if (!pred_dev || !truth_dev || !cm_dev || npx < 0) return fail(c, FB_ERR_INVALID, "confusion: null pointer or negative size");
  if (ncls < 1 || ncls > 32) return fail(c, FB_ERR_INVALID, "confusion: ncls must be in 1..32");
  FB_CUDA(c, cudaSetDevice(c->device));
  int rc = fb::launch_confusion(pred_dev, truth_dev, npx, ncls, truth_sub & 0xFF, reinterpret_cast<long long*>(cm_dev),
                                c->num_sms, c->stream);
  if (rc) return fail(c, rc, "confusion launch failed");
  c->launches++;
  return 0;
}

int fb_conv2d(fb_ctx* c, const void* x1, const void* x2, int C1, int C2, int up1, int B, int Hin, int Win,
              int KH, int KW, int stride, int pad, int Cout, const void* weights, int Kpad,
              const float* bias, const void* residual, const float* rowbias, int relu, void* out_bf16,
              float* out_f32, int mode) {
  if (!c || !x1 || !weights || !bias || (!out_bf16 == !out_f32)) return FB_ERR_INVALID;
  FB_CUDA(c, cudaSetDevice(c->device));
  fb::ConvArgs a;
  memset(&a, 0, sizeof a);
  a.x1 = static_cast<const __nv_bfloat16*>(x1);
  a.x2 = static_cast<const __nv_bfloat16*>(x2);
  a.C1 = C1; a.C2 = x2 ? C2 : 0; a.up1 = up1;
  a.B = B; a.Hin = Hin; a.Win = Win;
  a.Hout = (Hin + 2 * pad - KH) / stride + 1;
  a.Wout = (Win + 2 * pad - KW) / stride + 1;
  a.KH = KH; a.KW = KW; a.stride = stride; a.pad = pad;
  a.Cout = Cout;
  a.Ktot = KH * KW * (a.C1 + a.C2);
  a.bias = bias;
  a.residual = static_cast<const __nv_bfloat16*>(residual);
  a.rowbias = rowbias;
  a.relu = relu;
  a.out = static_cast<__nv_bfloat16*>(out_bf16);
  a.out_f32 = out_f32;
  if (mode == 2) {  // sub-pixel phase form: x1 is [B, Hin/2, Win/2, C1], weights are [4][Cout][Kpad]
    a.up1 = 0;
    a.phase_mode = 1;
    a.Ktot = Kpad;
    if (!fb::conv_tma_eligible(a)) return fail(c, FB_ERR_INVALID, "conv2d: shape not eligible for the phase form");
    const int rc = fb::launch_conv(a, static_cast<const __nv_bfloat16*>(weights), Kpad, true, c->num_sms, c->stream);
    if (rc) return fail(c, rc, "conv2d (phase) launch failed (code " + std::to_string(rc) + ")");
    c->launches++;
    return 0;
  }
  const bool can_tma = fb::conv_tma_eligible(a);
  bool tma;
  if (mode == 1) {
    if (!can_tma) return fail(c, FB_ERR_INVALID, "conv2d: shape not eligible for the TMA producer");
    tma = true;
  } else if (mode == 0) tma = false;
  else tma = can_tma && !c->force_gather;
  const int rc = fb::launch_conv(a, static_cast<const __nv_bfloat16*>(weights), Kpad, tma, c->num_sms, c->stream);
  if (rc) return fail(c, rc, "conv2d launch failed (code " + std::to_string(rc) + ")");
  c->launches++;
  return 0;
}

int fb_conv2d_halo(fb_ctx* c, const void* x1, const void* x2, int C1, int C2, int B, int Hin, int Win, int KH,
                   int stride, int Cout, const float* w_oihw_host, const float* bias, const void* residual,
                   int relu, int up2_out, void* out_bf16, float* out_f32) {
  if (!c || !x1 || !w_oihw_host || !bias || (!out_bf16 == !out_f32)) return FB_ERR_INVALID;
  FB_CUDA(c, cudaSetDevice(c->device));
  const int pad = KH / 2;
  fb::HaloArgs h;
  memset(&h, 0, sizeof h);
  h.x1 = static_cast<const __nv_bfloat16*>(x1);
  h.x2 = static_cast<const __nv_bfloat16*>(x2);
  h.C1 = C1; h.C2 = x2 ? C2 : 0;
  h.B = B; h.Hin = Hin; h.Win = Win;
  h.Hout = (Hin + 2 * pad - KH) / stride + 1;
  h.Wout = (Win + 2 * pad - KH) / stride + 1;
  h.Cout = Cout;
  h.bias = bias;
  h.residual = static_cast<const __nv_bfloat16*>(residual);
  h.relu = relu;
  h.out = static_cast<__nv_bfloat16*>(out_bf16);
  h.out_f32 = out_f32;
  h.up2_out = up2_out;
  if (up2_out == 3 || up2_out == 4) {
    // depth-to-space forms: 3 = 16 -> <= 16 channels as a 4x4 stride-2 conv over 2x2 cells, 4 = the conv of the
    // x2-upsampled 32-channel x1 on the low-res grid (HaloArgs::d2s = 1 / 2)
    const int mode = up2_out - 2;
    if (mode == 2) { h.Hout = 2 * Hin; h.Wout = 2 * Win; }
    if (KH != 3 || stride != 1 || x2 || residual || !fb::halo_d2s_supported(mode, C1, 0, Cout, h.Hout, h.Wout))
      return fail(c, FB_ERR_INVALID, "conv2d_halo: no depth-to-space instantiation for this shape");
    h.up2_out = 0;
    const size_t n = fb::pack_halo_weights_d2s(mode, w_oihw_host, Cout, C1, nullptr);
    std::vector<uint16_t> hp(n);
    fb::pack_halo_weights_d2s(mode, w_oihw_host, Cout, C1, hp.data());
    std::vector<float> bh(Cout), b64(64, 0.f);
    FB_CUDA(c, cudaMemcpy(bh.data(), bias, Cout * 4, cudaMemcpyDeviceToHost));
    for (int q = 0; q < 4; ++q)
      for (int o = 0; o < Cout; ++o) b64[q * 16 + o] = bh[o];
    __nv_bfloat16* wdev = nullptr;
    float* bdev = nullptr;
    FB_CUDA(c, cudaMalloc(&wdev, n * 2));
    FB_CUDA(c, cudaMalloc(&bdev, 64 * 4));
    cudaMemcpy(wdev, hp.data(), n * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(bdev, b64.data(), 64 * 4, cudaMemcpyHostToDevice);
    h.wpacked = wdev;
    h.bias = bdev;
    fb::halo_fill_steps_d2s(h, mode);
    const int rc = fb::launch_conv_halo(h, 3, 1, c->num_sms, c->stream);
    cudaStreamSynchronize(c->stream);
    cudaFree(wdev);
    cudaFree(bdev);
    if (rc) return fail(c, rc, "conv2d_halo (depth-to-space) launch failed (code " + std::to_string(rc) + ")");
    c->launches++;
    return 0;
  }
  if (up2_out == 2) {  // sub-pixel phase form: the conv of the x2-upsampled x1, computed from the low-res x1
    if (KH != 3 || stride != 1 || x2 || residual || out_f32 || !fb::halo_phase_supported(C1, 0, Cout, Hin, Win))
      return fail(c, FB_ERR_INVALID, "conv2d_halo: no phase-form instantiation for this shape");
    h.up2_out = 0;
    h.phase_mode = 1;
    h.Hout = 2 * Hin; h.Wout = 2 * Win;
    const size_t n = fb::pack_halo_weights_phase(w_oihw_host, Cout, Cout, C1, C1, nullptr);
    std::vector<uint16_t> hp(n);
    fb::pack_halo_weights_phase(w_oihw_host, Cout, Cout, C1, C1, hp.data());
    __nv_bfloat16* wdev = nullptr;
    FB_CUDA(c, cudaMalloc(&wdev, n * 2));
    cudaMemcpy(wdev, hp.data(), n * 2, cudaMemcpyHostToDevice);
    h.wpacked = wdev;
    fb::halo_fill_steps_phase(h);
    const int rc = fb::launch_conv_halo(h, 3, 1, c->num_sms, c->stream);
    cudaStreamSynchronize(c->stream);
    cudaFree(wdev);
    if (rc) return fail(c, rc, "conv2d_halo (phase) launch failed (code " + std::to_string(rc) + ")");
    c->launches++;
    return 0;
  }
  if (!fb::halo_supported(KH, stride, h.C1, h.C2, Cout, h.Hout, h.Wout))
    return fail(c, FB_ERR_INVALID, "conv2d_halo: no instantiation for this shape");
  const int Cin = h.C1 + h.C2;
  const size_t n = fb::pack_halo_weights(w_oihw_host, Cout, Cout, Cin, Cin, KH, stride, h.C1, h.C2, nullptr);
  std::vector<uint16_t> hp(n);
  fb::pack_halo_weights(w_oihw_host, Cout, Cout, Cin, Cin, KH, stride, h.C1, h.C2, hp.data());
  __nv_bfloat16* wdev = nullptr;
  __nv_bfloat16* wpair = nullptr;
  FB_CUDA(c, cudaMalloc(&wdev, n * 2));
  cudaMemcpy(wdev, hp.data(), n * 2, cudaMemcpyHostToDevice);
  if (KH == 3 && stride == 1 && Cout == 128 && h.C1 == 128 && h.C2 == 0) {
    std::vector<uint16_t> pp(n);
    fb::pack_halo_weights_pair128(hp.data(), n, pp.data());
    FB_CUDA(c, cudaMalloc(&wpair, n * 2));
    cudaMemcpy(wpair, pp.data(), n * 2, cudaMemcpyHostToDevice);
  }
  h.wpacked = wdev;
  h.wpacked_pair = wpair;
  h.pair = c->no_hpair ? 0 : 1;
  fb::halo_fill_steps(h, KH, stride);
  const int rc = fb::launch_conv_halo(h, KH, stride, c->num_sms, c->stream);
  cudaStreamSynchronize(c->stream);
  cudaFree(wdev);
  if (wpair) cudaFree(wpair);
  if (rc) return fail(c, rc, "conv2d_halo launch failed (code " + std::to_string(rc) + ")");
  c->launches++;
  return 0;
}

int fb_debug_activation(fb_ctx* c, const char* name, void* out_dev, int64_t* count, int32_t* dims4) {
  if (!c || !name) return FB_ERR_INVALID;
  auto it = c->acts.find(name);
  if (it == c->acts.end()) return fail(c, FB_ERR_INVALID, std::string("no activation named ") + name);
  const Act& a = it->second;
  const int64_t n = static_cast<int64_t>(a.B) * a.H * a.W * a.C;
  if (count) *count = n;
  if (dims4) { dims4[0] = a.B; dims4[1] = a.H; dims4[2] = a.W; dims4[3] = a.C; }
  if (out_dev) FB_CUDA(c, cudaMemcpyAsync(out_dev, a.ptr, static_cast<size_t>(n) * a.elem, cudaMemcpyDeviceToDevice, c->stream));
  return 0;
}

int fb_profile_begin(fb_ctx* c) {
  if (!c) return FB_ERR_INVALID;
  for (auto& r : c->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  c->prof_recs.clear();
  c->prof = true;
  const char* lt = getenv("FB_LAYER_TIMES");
  c->prof_layers = lt && lt[0] == '1';
  return 0;
}

int fb_profile_end(fb_ctx* c, float* ms4) {
  if (!c || !ms4) return FB_ERR_INVALID;
  c->prof = false;
  FB_CUDA(c, cudaStreamSynchronize(c->stream));
  double acc[4] = {0, 0, 0, 0};
  std::vector<std::pair<std::string, std::pair<double, int>>> layers;   // first-seen order
  for (auto& r : c->prof_recs) {
    float m = 0;
    const bool ok = cudaEventElapsedTime(&m, r.a, r.b) == cudaSuccess;
    if (ok && r.cat >= 0 && r.cat < 4) acc[r.cat] += m;
    if (ok && r.cat == 9 && r.label) {
      size_t k = 0;
      while (k < layers.size() && layers[k].first != r.label) ++k;
      if (k == layers.size()) layers.push_back({r.label, {0.0, 0}});
      layers[k].second.first += m;
      layers[k].second.second++;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  c->prof_recs.clear();
  if (c->prof_layers) {
    double tot = 0;
    for (auto& l : layers) tot += l.second.first;
    fprintf(stderr, "[layer times] CUDA events around every conv launch, %zu layers, %.3f ms in total\n", layers.size(), tot);
    for (auto& l : layers)
      fprintf(stderr, "[layer times] %-22s %5d launches %9.3f ms %7.1f us/launch %5.1f %%\n", l.first.c_str(), l.second.second,
              l.second.first, 1e3 * l.second.first / l.second.second, 100.0 * l.second.first / (tot > 0 ? tot : 1));
  }
  c->prof_layers = false;
  for (int i = 0; i < 4; ++i) ms4[i] = static_cast<float>(acc[i]);
  return 0;
}

int fb_profile_forward(fb_ctx* c, int n, int tile, int iters, float* ms5) {
  FB_TRY(check_ready(c, true, tile));
  if (n <= 0 || iters <= 0 || !ms5) return fail(c, FB_ERR_INVALID, "profile: bad arguments");
  FB_CUDA(c, cudaSetDevice(c->device));
  FB_TRY(ensure_arena(c, n, tile));
  FB_TRY(ensure_tile_buffers(c, n));
  std::vector<int> xy(static_cast<size_t>(n) * 2), rect(static_cast<size_t>(n) * 6);
  for (int i = 0; i < n; ++i) {
    const int x0 = static_cast<int>((static_cast<int64_t>(i) * 256) % (c->W > tile ? c->W - tile : 1));
    const int y0 = static_cast<int>(c->row0);
    xy[2 * i] = x0; xy[2 * i + 1] = y0;
    rect[6 * i] = x0; rect[6 * i + 1] = y0; rect[6 * i + 2] = x0; rect[6 * i + 3] = y0;
    rect[6 * i + 4] = x0 + tile; rect[6 * i + 5] = y0 + tile;
  }
  FB_CUDA(c, cudaMemcpy(c->tile_xy_dev, xy.data(), xy.size() * sizeof(int), cudaMemcpyHostToDevice));
  FB_CUDA(c, cudaMemcpy(c->tiles_dev, rect.data(), rect.size() * sizeof(int), cudaMemcpyHostToDevice));
  uint8_t* scratch = nullptr;
  FB_CUDA(c, cudaMalloc(&scratch, static_cast<size_t>(c->W) * (tile + 1)));
  const float* menc = nullptr;
  if (c->use_meta) {
    std::vector<float> met(static_cast<size_t>(n) * FB_METADATA_DIM, 0.5f);
    int rc = run_metadata(c, met.data(), n, &menc);
    if (rc) { cudaFree(scratch); return rc; }
    cudaStreamSynchronize(c->stream);
  }
  double acc[5] = {0, 0, 0, 0, 0};
  int rc = 0;
  for (int itn = 0; itn < iters + 1 && !rc; ++itn) {  // first pass is a warm-up
    c->prof = itn > 0;
    c->prof_recs.clear();
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0); cudaEventCreate(&t1);
    cudaEventRecord(t0, c->stream);
    rc = run_extract(c, c->raster, c->layout, c->bands_total, c->band_idx_dev, c->rc, c->W, c->H, c->row0, c->rows,
                     c->tile_xy_dev, n, tile);
    if (!rc) rc = run_network(c, n, tile, menc);
    if (!rc) {
      ProfScope ps(c, 3);
      rc = fb::launch_argmax_stitch(static_cast<const float*>(c->acts["logits"].ptr), c->ncls, c->ls, n, tile, c->tiles_dev,
                                    scratch, nullptr, c->W, c->row0, c->stream);
      c->launches++;
    }
    cudaEventRecord(t1, c->stream);
    cudaStreamSynchronize(c->stream);
    if (itn > 0) {
      float ms = 0;
      cudaEventElapsedTime(&ms, t0, t1);
      acc[4] += ms;
      for (auto& r : c->prof_recs) {
        float m = 0;
        cudaEventElapsedTime(&m, r.a, r.b);
        acc[r.cat] += m;
      }
    }
    for (auto& r : c->prof_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    c->prof_recs.clear();
    cudaEventDestroy(t0); cudaEventDestroy(t1);
  }
  c->prof = false;
  cudaFree(scratch);
  if (rc) return rc;
  for (int i = 0; i < 5; ++i) ms5[i] = static_cast<float>(acc[i] / iters);
  return 0;
}

}  // extern "C"
