/*
 * vrdx_cuda.h — CUDA-side extensions of the vk_radix_sort C API (libvrdx_b200.so).
 *
 * Nothing here exists in the reference; these are the pieces a CUDA caller needs because
 * the Vulkan objects the reference API mentions are opaque handles on this backend:
 * handle constructors, a query-pool object of GPU-written timestamps, an error channel (the
 * reference's vrdxCmd* return void, src/vk_radix_sort.h.in:51-81), creation options for
 * A/B measurement, and an import path for a Vulkan application's exported VkDeviceMemory.
 * Plain C ABI: pointers and integers only.
 */
#ifndef VRDX_CUDA_H
#define VRDX_CUDA_H

#include <stddef.h>
#include <stdint.h>

#include "vk_radix_sort.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ handle constructors */

/* VkDevice / VkPhysicalDevice for CUDA device `ordinal` (encoded as ordinal+1, never NULL). */
static inline VkDevice vrdxCudaDevice(int ordinal) { return (VkDevice)(uintptr_t)(ordinal + 1); }
static inline VkPhysicalDevice vrdxCudaPhysicalDevice(int ordinal) {
  return (VkPhysicalDevice)(uintptr_t)(ordinal + 1);
}
/* VkBuffer for a device pointer; the address used is (char*)devicePointer + offset. */
static inline VkBuffer vrdxCudaBuffer(const void* devicePointer) {
  return (VkBuffer)(uintptr_t)devicePointer;
}
/* VkCommandBuffer for a cudaStream_t (NULL = legacy default stream). */
static inline VkCommandBuffer vrdxCudaCommandBuffer(void* cudaStream) {
  return (VkCommandBuffer)cudaStream;
}

/* ------------------------------------------------------------------ creation options */

typedef enum VrdxCudaAlgorithm {
  VRDX_CUDA_ALGORITHM_AUTO = 0,             /* onesweep; reduce-then-scan when N >= 2^30 */
  VRDX_CUDA_ALGORITHM_ONESWEEP = 1,         /* histogram + 4 single-pass (decoupled look-back) kernels */
  VRDX_CUDA_ALGORITHM_REDUCE_THEN_SCAN = 2  /* per pass: tile histogram, spine scan, scatter (the reference's shape) */
} VrdxCudaAlgorithm;

typedef enum VrdxCudaTileLoad {
  VRDX_CUDA_TILE_LOAD_AUTO = 0,
  VRDX_CUDA_TILE_LOAD_DIRECT = 1, /* warp-striped coalesced global loads into registers */
  VRDX_CUDA_TILE_LOAD_TMA = 2     /* cp.async.bulk tile staging into shared memory (needs 16-byte aligned buffers) */
} VrdxCudaTileLoad;

typedef struct VrdxCudaSorterOptions {
  uint32_t structSize; /* = sizeof(VrdxCudaSorterOptions) */
  VrdxCudaAlgorithm algorithm;
  VrdxCudaTileLoad tileLoad;
  uint32_t reserved[5]; /* zero */
} VrdxCudaSorterOptions;

/* vrdxCreateSorter with explicit options (NULL options == vrdxCreateSorter). */
VkResult vrdxCudaCreateSorter(const VrdxSorterCreateInfo* pCreateInfo,
                              const VrdxCudaSorterOptions* pOptions, VrdxSorter* pSorter);

/* ------------------------------------------------------------------ error channel */

/* Last CUDA error (cudaError_t as int, 0 = none) raised by any vrdxCmd* on this sorter since
 * the previous call of this function; reading clears it.  vrdxCmd* never abort or throw. */
int vrdxCudaGetLastError(VrdxSorter sorter);
/* Static, human-readable description of an error code returned above. */
const char* vrdxCudaGetErrorString(int error);

/* Number of kernels + memsets the most recent vrdxCmd* on this sorter enqueued. */
uint32_t vrdxCudaGetLastLaunchCount(VrdxSorter sorter);

/* ------------------------------------------------------------------ timestamps */

/* A VkQueryPool of `queryCount` timestamp slots on `device` (>= 15 for one sort).  The slots are
 * device memory written by the sort's own kernels with the GPU global timer (nanoseconds), as
 * vkCmdWriteTimestamp would: no extra stream operations, legal inside CUDA-graph capture. */
VkResult vrdxCudaCreateQueryPool(VkDevice device, uint32_t queryCount, VkQueryPool* pQueryPool);
void vrdxCudaDestroyQueryPool(VkQueryPool queryPool);
/* After the stream has been synchronised: nanosecond timestamps of slots
 * [firstQuery, firstQuery+queryCount), relative to slot firstQuery (so pNanoseconds[0]==0). */
VkResult vrdxCudaGetQueryPoolResults(VkQueryPool queryPool, uint32_t firstQuery,
                                     uint32_t queryCount, uint64_t* pNanoseconds);

/* ------------------------------------------------------------------ Vulkan interop */

/* Import a VkDeviceMemory a Vulkan application exported as an opaque POSIX fd
 * (VK_KHR_external_memory_fd, handle type OPAQUE_FD).  On success CUDA owns the fd.
 * vrdxCudaImportedMemoryBuffer(mem, memoryOffset) is then the VkBuffer value to pass to
 * vrdxCmd* for a VkBuffer bound at `memoryOffset` inside that allocation. */
typedef struct VrdxCudaImportedMemory_T* VrdxCudaImportedMemory;
VkResult vrdxCudaImportMemoryFd(VkDevice device, int fd, VkDeviceSize allocationSize,
                                int dedicated, VrdxCudaImportedMemory* pMemory);
VkBuffer vrdxCudaImportedMemoryBuffer(VrdxCudaImportedMemory memory, VkDeviceSize memoryOffset);
void vrdxCudaReleaseImportedMemory(VrdxCudaImportedMemory memory);

/* ------------------------------------------------------------------ introspection */

typedef struct VrdxCudaSorterProperties {
  int deviceOrdinal;
  int smCount;
  int ccMajor, ccMinor;
  uint32_t keysTileSize;     /* keys per tile, keys-only onesweep kernel */
  uint32_t keyValueTileSize; /* keys per tile, key-value kernel */
  uint32_t offsetAlignment;  /* required alignment of every buffer offset (16) */
  uint32_t maxOnesweepCount; /* onesweep handles counts below this (2^30); above -> reduce-then-scan */
} VrdxCudaSorterProperties;
void vrdxCudaGetSorterProperties(VrdxSorter sorter, VrdxCudaSorterProperties* pProperties);

#ifdef __cplusplus
}
#endif
#endif /* VRDX_CUDA_H */
