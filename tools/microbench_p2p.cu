// microbench_p2p.cu — what SM-issued stores (and loads) reach over NVLink between two B200s, by access width
// and grid size.  Developer tool for the fused partition + exchange kernel (not part of the library).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench_p2p tools/microbench_p2p.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

template <typename T>
__global__ void CopyKernel(const T* __restrict__ src, T* __restrict__ dst, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}
// every CTA writes runs of `run` words at pseudo-random run-aligned places (the partition kernel's pattern)
__global__ void RunScatterKernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, size_t n, uint32_t run) {
  const size_t runs = n / run;
  for (size_t r = blockIdx.x; r < runs; r += gridDim.x) {
    const size_t to = (r * 2654435761ull) % runs;
    for (uint32_t i = threadIdx.x; i < run; i += blockDim.x) dst[to * run + i] = src[r * run + i];
  }
}

template <typename F>
float TimeMs(F f, int reps = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main() {
  int ndev = 0; CK(cudaGetDeviceCount(&ndev));
  if (ndev < 2) { printf("needs 2 GPUs\n"); return 0; }
  const size_t bytes = 1ull << 30;
  uint32_t *local_src, *local_dst, *peer;
  CK(cudaSetDevice(1)); CK(cudaMalloc(&peer, bytes)); CK(cudaMemset(peer, 0, bytes));
  CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0));
  CK(cudaMalloc(&local_src, bytes)); CK(cudaMalloc(&local_dst, bytes)); CK(cudaMemset(local_src, 1, bytes));
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t n4 = bytes / 4, n16 = bytes / 16;
  printf("1 GiB per transfer, GPU0 -> GPU1, %d SMs\n", sms);
  {
    float ms = TimeMs([&] { cudaMemcpyPeerAsync(peer, 1, local_src, 0, bytes, 0); });
    printf("cudaMemcpyPeer (copy engine)            %7.3f ms  %7.1f GB/s\n", ms, bytes / ms / 1e6);
  }
  for (int mult : {1, 2, 4, 8, 16}) {
    const int grid = sms * mult;
    float a = TimeMs([&] { CopyKernel<uint32_t><<<grid, 512>>>(local_src, peer, n4); });
    float b = TimeMs([&] { CopyKernel<uint4><<<grid, 512>>>((const uint4*)local_src, (uint4*)peer, n16); });
    float c = TimeMs([&] { CopyKernel<uint4><<<grid, 512>>>((const uint4*)peer, (uint4*)local_dst, n16); });
    float d = TimeMs([&] { CopyKernel<uint4><<<grid, 512>>>((const uint4*)local_src, (uint4*)local_dst, n16); });
    printf("grid %5d x 512: peer store 4B %7.1f GB/s | peer store 16B %7.1f GB/s | peer load 16B %7.1f GB/s | local copy 16B %7.1f GB/s\n",
           grid, bytes / a / 1e6, bytes / b / 1e6, bytes / c / 1e6, bytes / d / 1e6);
  }
  for (uint32_t run : {64u, 256u, 1024u, 4096u}) {
    float a = TimeMs([&] { RunScatterKernel<<<sms * 8, 256>>>(local_src, peer, n4, run); });
    float b = TimeMs([&] { RunScatterKernel<<<sms * 8, 256>>>(local_src, local_dst, n4, run); });
    printf("scattered runs of %5u words (4B stores): peer %7.1f GB/s | local %7.1f GB/s\n", run, bytes / a / 1e6, bytes / b / 1e6);
  }
  return 0;
}
