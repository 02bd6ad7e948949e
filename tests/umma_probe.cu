// Micro-benchmark (not part of the library): issue rate of tcgen05.mma kind::f16 as a function of N, M, the
// shared-memory operand layout (SWIZZLE_128B / no-swizzle dense / no-swizzle halo-style), A from TMEM, the
// number of accumulators cycled through and concurrent shared-memory traffic from other warps.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tests/umma_probe tests/umma_probe.cu && tests/umma_probe
//
// Output: one line per configuration with cycles (clock64) and ns per MMA, median over the CTAs of the grid.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#include "../flair-1_b200/csrc/ptx.cuh"

using namespace fb::ptx;

struct P {
  uint32_t idesc;
  uint32_t a_lo, a_hi, a_step;  // descriptor words relative to the operand region (16-byte units), step per MMA
  uint32_t b_lo, b_hi, b_step;
  int ncyc;                     // descriptor positions cycled through
  int iters;
  int nacc, ncols;              // accumulators cycled through, TMEM columns each
  int a_tmem;                   // 1: A operand from TMEM
  int noise;                    // 1: warps 4-7 hammer shared memory with 16-byte stores + loads while the MMAs run
};

__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(1u)
      : "memory");
}

constexpr int kOperandBytes = 96 * 1024;
constexpr int kNoiseBytes = 32 * 1024;

// ELECT: the issuing thread is chosen with elect.sync (as opposed to `tid == 0`)
template <bool ELECT>
__global__ void __launch_bounds__(256) probe(P p, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int done;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (kOperandBytes + kNoiseBytes) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    done = 0;
    mbar_init(smem_u32(&bar), 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&slot), 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = slot;
  const uint32_t base16 = smem_u32(smem) >> 4;
  bool issuer;
  if (ELECT) issuer = warp == 0 && elect_one();
  else issuer = tid == 0;
  if (issuer) {
    const long long t0 = clock64();
    int c = 0, a = 0;
    for (int i = 0; i < p.iters; ++i) {
      const uint32_t d = tmem + a * p.ncols;
      if (p.a_tmem) umma_bf16_ts(d, tmem + 480, p.b_lo + base16 + c * p.b_step, p.b_hi, p.idesc);
      else umma_bf16_lohi(d, p.a_lo + base16 + c * p.a_step, p.a_hi, p.b_lo + base16 + c * p.b_step, p.b_hi, p.idesc, 1u);
      if (++c == p.ncyc) c = 0;
      if (++a == p.nacc) a = 0;
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
    done = 1;
  } else if (p.noise && warp >= 4) {
    uint4* region = reinterpret_cast<uint4*>(smem + kOperandBytes);
    uint4 v = make_uint4(tid, 0, 0, 0);
    int k = tid - 128;
    while (!done) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        region[k] = v;
        k = (k + 128) & (kNoiseBytes / 16 - 1);
        const uint4 r = region[k];
        v.x += r.y;
      }
    }
    if (v.x == 0x12345678u) out[0] = 0;
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem, 512);
  }
}

// Latency probes for the barrier round trips of a warp-specialised pipeline:
//   mode 0: tcgen05.commit -> mbarrier completion observed by the committing thread, nothing pending
//   mode 1: the same with one M=128 N=64 MMA issued before each commit
//   mode 2: mbarrier ping-pong between lane 0 of warp 0 and lane 0 of warp 4 (arrive -> try_wait success), per hop
//   mode 3: commit by warp 0's elected thread observed by warp 4 (all 32 lanes waiting), which arrives back (32 arrivals)
__global__ void __launch_bounds__(256) latency_probe(int mode, int iters, uint32_t idesc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < kOperandBytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), mode == 3 ? 32 : 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(&slot), 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = slot;
  const uint32_t base16 = smem_u32(smem) >> 4;
  const uint32_t b0 = smem_u32(&bars[0]), b1 = smem_u32(&bars[1]);
  if (mode <= 1) {
    if (warp == 0 && elect_one()) {
      const long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        if (mode == 1)
          umma_bf16_lohi(tmem, base16 | (1u << 16), 64u | (1u << 14) | (2u << 29), (base16 + 2048) | (1u << 16),
                         64u | (1u << 14) | (2u << 29), idesc, 0u);
        umma_commit(b0);
        mbar_wait(b0, i & 1);
      }
      out[blockIdx.x] = clock64() - t0;
    }
  } else if (mode == 2) {
    if (tid == 0) {
      const long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        mbar_arrive(b0);
        mbar_wait(b1, i & 1);
      }
      out[blockIdx.x] = (clock64() - t0) / 2;
    } else if (tid == 128) {
      for (int i = 0; i < iters; ++i) {
        mbar_wait(b0, i & 1);
        mbar_arrive(b1);
      }
    }
  } else {
    if (warp == 0 && elect_one()) {
      const long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        umma_commit(b0);
        mbar_wait(b1, i & 1);
        tc_fence_after_sync();
      }
      out[blockIdx.x] = clock64() - t0;
    } else if (warp == 4) {
      for (int i = 0; i < iters; ++i) {
        mbar_wait(b0, i & 1);
        tc_fence_after_sync();
        tc_fence_before_sync();
        mbar_arrive(b1);
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after_sync();
    tmem_dealloc(tmem, 512);
  }
}

struct Cfg {
  const char* name;
  int M, N, layout, a_tmem, nacc, noise, per_sm;
};

int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaSetDevice(dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  printf("SMs %d, max clock %d MHz\n", sms, khz / 1000);
  const int smem = kOperandBytes + kNoiseBytes;
  cudaFuncSetAttribute(probe<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* out = nullptr;
  cudaMalloc(&out, 8 * 1024);
  {
    cudaFuncSetAttribute(latency_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const char* names[4] = {"commit -> own wait, nothing pending", "1 MMA (N=64) + commit -> own wait",
                            "mbarrier arrive -> other warp's try_wait (per hop)", "commit -> warp of 32 waits -> 32 arrivals -> issuer"};
    for (int mode = 0; mode < 4; ++mode) {
      const int it = 2000;
      latency_probe<<<sms, 256, smem>>>(mode, it, umma_idesc_bf16(128, 64), out);
      cudaError_t err = cudaDeviceSynchronize();
      if (err != cudaSuccess) { printf("latency probe %d: %s\n", mode, cudaGetErrorString(err)); return 1; }
      std::vector<long long> h(sms);
      cudaMemcpy(h.data(), out, sms * 8, cudaMemcpyDeviceToHost);
      std::sort(h.begin(), h.end());
      printf("latency: %-55s %8.1f cycles (min %.1f max %.1f)\n", names[mode], static_cast<double>(h[sms / 2]) / it,
             static_cast<double>(h[0]) / it, static_cast<double>(h[sms - 1]) / it);
    }
    fflush(stdout);
  }
  std::vector<Cfg> cfgs;
  for (int layout = 0; layout < 3; ++layout)
    for (int N : {16, 32, 64, 128, 256}) cfgs.push_back({layout == 0 ? "sw128" : layout == 1 ? "none-dense" : "none-halo", 128, N, layout, 0, 2, 0, 1});
  for (int N : {16, 32, 64, 128, 256}) cfgs.push_back({"A-in-TMEM", 128, N, 1, 1, 2, 0, 1});
  for (int N : {16, 64, 256}) cfgs.push_back({"sw128 M=64", 64, N, 0, 0, 2, 0, 1});
  for (int N : {16, 32, 64}) cfgs.push_back({"none-halo 1 acc", 128, N, 2, 0, 1, 0, 1});
  for (int N : {16, 32, 64}) cfgs.push_back({"none-halo 4 acc", 128, N, 2, 0, 4, 0, 1});
  for (int N : {16, 32, 64, 256}) cfgs.push_back({"none-halo +noise", 128, N, 2, 0, 2, 1, 1});
  for (int N : {16, 32, 64, 256}) cfgs.push_back({"sw128 +noise", 128, N, 0, 0, 2, 1, 1});
  const int iters = 4096;
  for (const Cfg& c : cfgs) {
    P p;
    p.idesc = umma_idesc_bf16(c.M, c.N);
    p.iters = iters;
    p.nacc = c.nacc;
    p.ncols = c.N < 32 ? 32 : c.N;
    if (p.ncols * p.nacc > 448) p.nacc = 448 / p.ncols;
    p.a_tmem = c.a_tmem;
    p.noise = c.noise;
    const uint32_t b_off16 = 32 * 1024 / 16;
    if (c.layout == 0) {  // SWIZZLE_128B, K-major, 128-byte rows: K = 64 per row, 4 K16 steps
      p.a_lo = 0 | (1u << 16);
      p.a_hi = 64u | (1u << 14) | (2u << 29);
      p.a_step = 2;
      p.b_lo = b_off16 | (1u << 16);
      p.b_hi = p.a_hi;
      p.b_step = 2;
      p.ncyc = 4;
    } else {
      if (c.layout == 1) {  // dense interleaved: core matrix 128 B, 8-row groups 128 B apart, K chunks M*16 B apart
        p.a_lo = 0 | (static_cast<uint32_t>(c.M) << 16);
        p.a_hi = 8u | (1u << 14);
        p.a_step = 2 * c.M;
        p.ncyc = 4;
      } else {  // halo-style: rows of an 8-row group are consecutive cells, next group 10 cells on, chunk = plane of 180 cells
        p.a_lo = 0 | (180u << 16);
        p.a_hi = 10u | (1u << 14);
        p.a_step = 1;
        p.ncyc = 3;
      }
      p.b_lo = b_off16 | (static_cast<uint32_t>(c.N) << 16);
      p.b_hi = 8u | (1u << 14);
      p.b_step = 2 * c.N;
      p.ncyc = c.layout == 1 ? 4 : 3;
      if (p.b_step * 16 * p.ncyc > 64 * 1024) p.b_step = 0;
    }
    const int grid = sms * c.per_sm;
    for (int elect = 0; elect < 2; ++elect) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    if (elect) probe<true><<<grid, 256, smem>>>(p, out); else probe<false><<<grid, 256, smem>>>(p, out);  // warm-up
    cudaEventRecord(e0);
    if (elect) probe<true><<<grid, 256, smem>>>(p, out); else probe<false><<<grid, 256, smem>>>(p, out);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
      printf("%-18s M=%3d N=%3d: CUDA error %s\n", c.name, c.M, c.N, cudaGetErrorString(err));
      return 1;
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), out, grid * 8, cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    const double cyc = static_cast<double>(h[grid / 2]) / iters;
    const double math = 128.0 * c.N / 256.0 * (c.M / 128.0);
    printf("%-18s %s M=%3d N=%3d acc=%d: %7.1f cycles/MMA (min %.1f max %.1f), kernel %.1f us -> %.1f ns/MMA; math-bound %.0f cycles; operand bytes/cycle %.1f\n",
           c.name, elect ? "elect " : "tid==0", c.M, c.N, p.nacc, cyc, static_cast<double>(h[0]) / iters,
           static_cast<double>(h[grid - 1]) / iters, ms * 1e3, ms * 1e6 / iters, math, ((c.a_tmem ? 0 : c.M) + c.N) * 32.0 / cyc);
    fflush(stdout);
    }
  }
  cudaFree(out);
  return 0;
}
