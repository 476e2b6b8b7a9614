// The reference's README usage (README.md:121-183 of vulkan_radix_sort), on libvrdx_b200.so: create a sorter, size
// and allocate the storage, record a key-value sort with a device-resident count, wait, check.  Also one
// vrdxCudaCmdSortEx call (float keys, descending).  Build: see examples/Makefile.  Exit code 0 = results correct.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <numeric>
#include <random>
#include <vector>

#include <vk_radix_sort.h>
#include <vrdx_cuda.h>

#define CHECK_CUDA(x)                                                                  \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                    \
      return 2;                                                                        \
    }                                                                                  \
  } while (0)

int main(int argc, char** argv) {
  const uint32_t max_count = argc > 1 ? (uint32_t)std::strtoul(argv[1], nullptr, 10) : 1000003u;
  const uint32_t count = max_count - max_count / 7;  // the device-side count is below the maximum

  VrdxSorterCreateInfo info = {vrdxCudaPhysicalDevice(0), vrdxCudaDevice(0), VK_NULL_HANDLE};
  VrdxSorter sorter = VK_NULL_HANDLE;
  if (vrdxCreateSorter(&info, &sorter) != VK_SUCCESS) {
    std::fprintf(stderr, "vrdxCreateSorter failed (needs an sm_100 device)\n");
    return 2;
  }
  VrdxSorterStorageRequirements req;
  vrdxGetSorterKeyValueStorageRequirements(sorter, max_count, &req);

  std::mt19937 rng(42);
  std::vector<uint32_t> keys(max_count), values(max_count);
  for (auto& k : keys) k = rng();
  std::iota(values.begin(), values.end(), 0u);

  uint32_t *d_keys, *d_values, *d_count;
  void* d_storage;
  CHECK_CUDA(cudaMalloc(&d_keys, 4ull * max_count));
  CHECK_CUDA(cudaMalloc(&d_values, 4ull * max_count));
  CHECK_CUDA(cudaMalloc(&d_count, 4));
  CHECK_CUDA(cudaMalloc(&d_storage, req.size));
  cudaStream_t stream;
  CHECK_CUDA(cudaStreamCreate(&stream));
  CHECK_CUDA(cudaMemcpyAsync(d_keys, keys.data(), 4ull * max_count, cudaMemcpyHostToDevice, stream));
  CHECK_CUDA(cudaMemcpyAsync(d_values, values.data(), 4ull * max_count, cudaMemcpyHostToDevice, stream));
  CHECK_CUDA(cudaMemcpyAsync(d_count, &count, 4, cudaMemcpyHostToDevice, stream));

  VkQueryPool pool = VK_NULL_HANDLE;
  vrdxCudaCreateQueryPool(vrdxCudaDevice(0), 15, &pool);
  vrdxCmdSortKeyValueIndirect(vrdxCudaCommandBuffer(stream), sorter, max_count, vrdxCudaBuffer(d_count), 0,
                              vrdxCudaBuffer(d_keys), 0, vrdxCudaBuffer(d_values), 0, vrdxCudaBuffer(d_storage), 0,
                              pool, 0);
  std::vector<uint32_t> out_k(max_count), out_v(max_count);
  CHECK_CUDA(cudaMemcpyAsync(out_k.data(), d_keys, 4ull * max_count, cudaMemcpyDeviceToHost, stream));
  CHECK_CUDA(cudaMemcpyAsync(out_v.data(), d_values, 4ull * max_count, cudaMemcpyDeviceToHost, stream));
  CHECK_CUDA(cudaStreamSynchronize(stream));
  if (int e = vrdxCudaGetLastError(sorter)) {
    std::fprintf(stderr, "sort failed: %s\n", vrdxCudaGetErrorString(e));
    return 2;
  }

  // the reference's own check (bench/bench.cc:41-64): element-wise equality with a stable CPU sort
  std::vector<uint32_t> perm(count);
  std::iota(perm.begin(), perm.end(), 0u);
  std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
  int bad = 0;
  for (uint32_t i = 0; i < count; ++i) bad += out_k[i] != keys[perm[i]] || out_v[i] != values[perm[i]];
  for (uint32_t i = count; i < max_count; ++i) bad += out_k[i] != keys[i] || out_v[i] != values[i];  // tail untouched
  uint64_t ns[15];
  if (vrdxCudaGetQueryPoolResults(pool, 0, 15, ns) == VK_SUCCESS)
    std::printf("key-value indirect sort of %u of %u pairs: %.3f ms on the GPU, %d mismatches\n", count, max_count,
                ns[14] * 1e-6, bad);

  // extension: float keys, descending, keys only, direct count
  std::vector<float> f(max_count);
  std::normal_distribution<float> nd(0.f, 1e3f);
  for (auto& x : f) x = nd(rng);
  CHECK_CUDA(cudaMemcpyAsync(d_keys, f.data(), 4ull * max_count, cudaMemcpyHostToDevice, stream));
  VrdxCudaSortKeyInfo ki = {sizeof ki, VRDX_CUDA_KEY_TYPE_FLOAT32, VRDX_CUDA_SORT_ORDER_DESCENDING, 0, 32, {0, 0, 0}};
  vrdxCudaCmdSortEx(vrdxCudaCommandBuffer(stream), sorter, &ki, max_count, VK_NULL_HANDLE, 0, vrdxCudaBuffer(d_keys), 0,
                    VK_NULL_HANDLE, 0, vrdxCudaBuffer(d_storage), 0, VK_NULL_HANDLE, 0);
  std::vector<float> out_f(max_count);
  CHECK_CUDA(cudaMemcpyAsync(out_f.data(), d_keys, 4ull * max_count, cudaMemcpyDeviceToHost, stream));
  CHECK_CUDA(cudaStreamSynchronize(stream));
  std::sort(f.begin(), f.end(), std::greater<float>());
  int bad_f = 0;
  for (uint32_t i = 0; i < max_count; ++i) bad_f += out_f[i] != f[i];
  std::printf("float32 descending sort of %u keys: %d mismatches\n", max_count, bad_f);

  vrdxCudaDestroyQueryPool(pool);
  vrdxDestroySorter(sorter);
  cudaFree(d_keys); cudaFree(d_values); cudaFree(d_count); cudaFree(d_storage);
  cudaStreamDestroy(stream);
  return (bad || bad_f) ? 1 : 0;
}
