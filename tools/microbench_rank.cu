// microbench_rank.cu — throughput of the warp-level primitives a radix-sort ranking step can be
// built from, on sm_100a.  Developer tool (not part of the library).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench_rank tools/microbench_rank.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

constexpr int kThreads = 512;
constexpr int kIters = 2048;

__device__ __forceinline__ uint32_t Rng(uint32_t& s) {
  s ^= s << 13; s ^= s >> 17; s ^= s << 5;
  return s;
}
__device__ __forceinline__ uint32_t LaneMaskLt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

// mode 0: baseline (rng only); 1: MATCH.ANY 8-bit; 2: ballot loop 8 bits; 3: smem atomicOr + LDS + clear;
// 4: smem atomicAdd (1 per key, warp-private bins); 5: smem atomicAdd shared bins (block-wide 256 bins);
// 6: 4x atomicAdd block-wide (histogram kernel shape); 7: vote only x8; 8: match.any 4-bit
// 9: full ballot-rank step (ballot loop + LDS/STS counter update); 10: full atomicOr-rank step
template <int MODE>
__global__ void __launch_bounds__(kThreads) Bench(uint32_t* out, int entropy_mask) {
  __shared__ uint32_t sm[16 * 256 + 1024];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 16 * 256 + 1024; i += kThreads) sm[i] = 0;
  __syncthreads();
  uint32_t* wh = sm + warp * 256;
  uint32_t* bh = sm + 16 * 256;
  uint32_t s = 0x9E3779B9u * (blockIdx.x * kThreads + tid + 1);
  uint32_t acc = 0;
  const uint32_t lt = LaneMaskLt();
  for (int it = 0; it < kIters; ++it) {
    uint32_t key = Rng(s);
    uint32_t d = key & 0xFFu & entropy_mask;
    if (MODE == 0) acc += d;
    if (MODE == 1) acc += __match_any_sync(0xffffffffu, d);
    if (MODE == 8) acc += __match_any_sync(0xffffffffu, d & 15u);
    if (MODE == 2 || MODE == 9) {
      uint32_t peers = 0xffffffffu;
#pragma unroll
      for (int b = 0; b < 8; ++b) {
        const bool bit = (d >> b) & 1u;
        const uint32_t m = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? m : ~m;
      }
      if (MODE == 2) acc += peers;
      if (MODE == 9) {
        const uint32_t before = wh[d];
        const uint32_t below = __popc(peers & lt);
        __syncwarp();
        if (below == 0) wh[d] = before + __popc(peers);
        __syncwarp();
        acc += before + below;
      }
    }
    if (MODE == 7) {
#pragma unroll
      for (int b = 0; b < 8; ++b) acc += __ballot_sync(0xffffffffu, (d >> b) & 1u);
    }
    if (MODE == 3 || MODE == 10) {
      atomicOr(&wh[d], 1u << lane);
      __syncwarp();
      const uint32_t peers = wh[d];
      __syncwarp();
      if (MODE == 3) {
        if ((peers & lt) == 0) wh[d] = 0;
        __syncwarp();
        acc += peers;
      } else {
        // counters live in a second array (bh is block-wide here only for the footprint)
        uint32_t* cnt = bh + (warp & 3) * 256;
        const uint32_t before = cnt[d];
        const uint32_t below = __popc(peers & lt);
        __syncwarp();
        if (below == 0) { cnt[d] = before + __popc(peers); wh[d] = 0; }
        __syncwarp();
        acc += before + below;
      }
    }
    if (MODE == 4) acc += atomicAdd(&wh[d], 1u);
    if (MODE == 5) atomicAdd(&bh[d], 1u);
    if (MODE == 6) {
      atomicAdd(&bh[d], 1u);
      atomicAdd(&bh[256 + ((key >> 8) & 0xFFu & entropy_mask)], 1u);
      atomicAdd(&bh[512 + ((key >> 16) & 0xFFu & entropy_mask)], 1u);
      atomicAdd(&bh[768 + ((key >> 24) & entropy_mask)], 1u);
    }
  }
  __syncthreads();
  if (MODE >= 3) acc += sm[tid] + sm[16 * 256 + tid];
  out[blockIdx.x * kThreads + tid] = acc;
}

template <int MODE>
void Run(const char* name, uint32_t* out, int sms, int entropy_mask) {
  const int blocks = sms * 2;  // 2 CTAs x 16 warps per SM = 32 warps/SM
  Bench<MODE><<<blocks, kThreads>>>(out, entropy_mask);
  cudaDeviceSynchronize();
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  Bench<MODE><<<blocks, kThreads>>>(out, entropy_mask);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double keys = (double)blocks * kThreads * kIters;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double cyc_per_warp_item_per_sm = (ms * 1e-3) * (clk * 1e3) / (keys / 32 / sms);
  printf("%-44s mask=%3d  %8.3f ms  %8.1f Gkeys/s  %7.2f cyc/warp-item/SM (at %d MHz nominal)  err=%d\n", name,
         entropy_mask, ms, keys / ms / 1e6, cyc_per_warp_item_per_sm, clk / 1000, (int)cudaGetLastError());
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint32_t* out; cudaMalloc(&out, sizeof(uint32_t) * sms * 2 * kThreads);
  for (int mask : {255, 15, 0}) {
    Run<0>("0 baseline rng", out, sms, mask);
    Run<1>("1 match.any 8-bit", out, sms, mask);
    Run<8>("8 match.any 4-bit", out, sms, mask);
    Run<2>("2 ballot loop 8 bits", out, sms, mask);
    Run<7>("7 vote x8 only", out, sms, mask);
    Run<3>("3 smem atomicOr + LDS + clear", out, sms, mask);
    Run<4>("4 smem atomicAdd warp-private (returns)", out, sms, mask);
    Run<5>("5 smem atomicAdd block bins (no return)", out, sms, mask);
    Run<6>("6 4x smem atomicAdd block bins (hist kernel)", out, sms, mask);
    Run<9>("9 full rank step: ballot + counters", out, sms, mask);
    Run<10>("10 full rank step: atomicOr + counters", out, sms, mask);
  }
  return 0;
}
