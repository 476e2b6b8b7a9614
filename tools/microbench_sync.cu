// microbench_sync.cu — what one dependent step costs on B200: a chain of tiny kernels (plain launches, programmatic
// dependent launches) against grid barriers inside one co-resident kernel.  Decides whether a single cooperative
// kernel is worth building for small sorts (DESIGN "small sorts").
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench_sync tools/microbench_sync.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include <algorithm>

__device__ __forceinline__ unsigned long long Now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// one "pass": every CTA reads a word written by the previous step (all CTAs), writes one for the next
__global__ void __launch_bounds__(256) StepKernel(const unsigned* in, unsigned* out, int pdl) {
  if (pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  unsigned v = in[(blockIdx.x + 1) % gridDim.x];
  if (threadIdx.x == 0) out[blockIdx.x] = v + 1;
}

__device__ __forceinline__ void GridBarrier(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned v;
    do {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) BarrierKernel(unsigned* a, unsigned* b, unsigned* counter, int steps,
                                                     unsigned long long* t) {
  unsigned long long t0 = Now();
  for (int s = 0; s < steps; ++s) {
    const unsigned* in = (s & 1) ? b : a;
    unsigned* out = (s & 1) ? a : b;
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(in + (blockIdx.x + 1) % gridDim.x) : "memory");
    if (threadIdx.x == 0) out[blockIdx.x] = v + 1;
    GridBarrier(counter, (unsigned)(s + 1) * gridDim.x);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { t[0] = t0; t[1] = Now(); }
}

int main() {
  unsigned *a, *b, *counter;
  unsigned long long* t;
  cudaMalloc(&a, 4096 * 4); cudaMalloc(&b, 4096 * 4); cudaMalloc(&counter, 4); cudaMalloc(&t, 16);
  cudaMemset(a, 0, 4096 * 4); cudaMemset(b, 0, 4096 * 4);
  cudaStream_t st; cudaStreamCreate(&st);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int steps = 6, reps = 200;
  for (int grid : {52, 128, 256, 592}) {
    for (int pdl = 0; pdl < 2; ++pdl) {
      std::vector<float> ms;
      for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0, st);
        for (int s = 0; s < steps; ++s) {
          cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.stream = st;
          cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
          at[0].val.programmaticStreamSerializationAllowed = 1; cfg.attrs = at; cfg.numAttrs = (pdl && s > 0) ? 1 : 0;
          cudaLaunchKernelEx(&cfg, StepKernel, (const unsigned*)((s & 1) ? b : a), (s & 1) ? a : b, pdl);
        }
        cudaEventRecord(e1, st); cudaEventSynchronize(e1);
        float m; cudaEventElapsedTime(&m, e0, e1); ms.push_back(m);
      }
      std::sort(ms.begin(), ms.end());
      printf("grid %4d  chain of %d kernels, %s: median %.2f us total, %.2f us per kernel\n", grid, steps,
             pdl ? "PDL  " : "plain", ms[reps / 2] * 1e3, ms[reps / 2] * 1e3 / steps);
    }
    std::vector<double> us; std::vector<float> ms;
    for (int r = 0; r < reps; ++r) {
      cudaMemsetAsync(counter, 0, 4, st);
      cudaEventRecord(e0, st);
      void* args[] = {&a, &b, &counter, (void*)&steps, &t};
      cudaLaunchCooperativeKernel((void*)BarrierKernel, dim3(grid), dim3(256), args, 0, st);
      cudaEventRecord(e1, st); cudaEventSynchronize(e1);
      unsigned long long h[2]; cudaMemcpy(h, t, 16, cudaMemcpyDeviceToHost);
      us.push_back((h[1] - h[0]) / 1e3);
      float m; cudaEventElapsedTime(&m, e0, e1); ms.push_back(m);
    }
    std::sort(us.begin(), us.end()); std::sort(ms.begin(), ms.end());
    printf("grid %4d  one cooperative kernel, %d grid barriers: in-kernel median %.2f us (%.2f us per step), events %.2f us; err=%s\n",
           grid, steps, us[reps / 2], us[reps / 2] / steps, ms[reps / 2] * 1e3, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
