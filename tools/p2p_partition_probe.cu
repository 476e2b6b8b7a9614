// p2p_partition_probe.cu — the fused partition + exchange kernel (vrdxDistCmdPartitionScatter) in ONE
// process driving TWO GPUs, so that it can run under ncu (never wrap a multi-rank command in ncu):
// half of GPU 0's keys are stored straight into a buffer on GPU 1 over NVLink.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -Iinclude -o tools/p2p_partition_probe \
//        tools/p2p_partition_probe.cu -Lvulkan_radix_sort_b200/lib -lvrdx_b200 \
//        -Xlinker -rpath -Xlinker '$ORIGIN/../vulkan_radix_sort_b200/lib'
//   ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,gpu__time_duration.sum -k regex:DistPartition tools/p2p_partition_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>
#define VRDX_FORCE_VK_SHIM 1
#include "vk_radix_sort.h"
#include "vrdx_cuda.h"
#include "vrdx_dist.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__global__ void Fill(uint32_t* k, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint64_t x = (i + 1) * 0x9E3779B97F4A7C15ull;
    x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
    k[i] = (uint32_t)x;
  }
}

int main(int argc, char** argv) {
  int ndev = 0; CK(cudaGetDeviceCount(&ndev));
  if (ndev < 2) { printf("needs 2 GPUs\n"); return 0; }
  const int dests = argc > 1 ? atoi(argv[1]) : 2;      // 2 or 8 destinations (odd ones on the peer GPU)
  const uint32_t n = 1u << 29;
  uint32_t *keys, *local_out, *peer_out, *splitters, *cursors, *counts;
  CK(cudaSetDevice(1)); CK(cudaMalloc(&peer_out, (size_t)n * 4));
  CK(cudaSetDevice(0)); CK(cudaDeviceEnablePeerAccess(1, 0));
  CK(cudaMalloc(&keys, (size_t)n * 4)); CK(cudaMalloc(&local_out, (size_t)n * 4));
  CK(cudaMalloc(&splitters, 64)); CK(cudaMalloc(&cursors, 256)); CK(cudaMalloc(&counts, 256));
  Fill<<<148 * 8, 256>>>(keys, n);
  VrdxSorterCreateInfo ci{vrdxCudaPhysicalDevice(0), vrdxCudaDevice(0), VK_NULL_HANDLE};
  VrdxSorter sorter;
  if (vrdxCreateSorter(&ci, &sorter) != VK_SUCCESS) { printf("create failed\n"); return 1; }
  const int m = dests - 1;
  std::vector<uint32_t> spl(m);
  for (int k = 0; k < m; ++k) spl[k] = (uint32_t)(((uint64_t)(k + 1) << 32) / dests);
  CK(cudaMemcpy(splitters, spl.data(), 4 * m, cudaMemcpyHostToDevice));
  CK(cudaMemset(counts, 0, 256));
  vrdxDistCmdClassCount(nullptr, sorter, n, (VkBuffer)keys, 0, m, (VkBuffer)splitters, 0, (VkBuffer)counts, 0);
  std::vector<uint32_t> cls(2 * m + 1);
  CK(cudaMemcpy(cls.data(), counts, 4 * (2 * m + 1), cudaMemcpyDeviceToHost));
  std::vector<uint32_t> starts(2 * m + 1, 0), first(dests + 1, 0);
  for (int c = 1; c < 2 * m + 1; ++c) starts[c] = starts[c - 1] + cls[c - 1];
  for (int j = 1; j < dests; ++j) first[j] = starts[2 * j - 1];   // destination j starts at the tie class of splitter j
  first[dests] = n;
  // destination table: uint64 ptr[dests] | uint32 first_pos[dests + 1]; odd destinations live on GPU 1
  std::vector<unsigned char> table(8 * dests + 4 * (dests + 1));
  uint64_t remote_bytes = 0;
  for (int j = 0; j < dests; ++j) {
    uint64_t p = (uint64_t)(uintptr_t)((j & 1) ? peer_out : local_out) + 4ull * first[j];
    memcpy(&table[8 * j], &p, 8);
    if (j & 1) remote_bytes += 4ull * (first[j + 1] - first[j]);
  }
  memcpy(&table[8 * dests], first.data(), 4 * (dests + 1));
  unsigned char* dtable; CK(cudaMalloc(&dtable, table.size()));
  CK(cudaMemcpy(dtable, table.data(), table.size(), cudaMemcpyHostToDevice));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9f;
  for (int it = 0; it < 4; ++it) {
    CK(cudaMemcpy(cursors, starts.data(), 4 * (2 * m + 1), cudaMemcpyHostToDevice));
    cudaEventRecord(a);
    vrdxDistCmdPartitionScatter(nullptr, sorter, n, (VkBuffer)keys, 0, m, (VkBuffer)splitters, 0, (VkBuffer)cursors, 0, dests,
                                (VkBuffer)dtable, 0);
    cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (it && ms < best) best = ms;
  }
  printf("dests=%d  n=2^29  partition+scatter %.3f ms  (%.1f GKeys/s)  remote %.3f GB -> %.1f GB/s over NVLink  err=%d\n", dests,
         best, n / best / 1e6, remote_bytes / 1e9, remote_bytes / best / 1e6, vrdxCudaGetLastError(sorter));
  vrdxDestroySorter(sorter);
  return 0;
}
