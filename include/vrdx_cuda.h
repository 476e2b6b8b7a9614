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
  VRDX_CUDA_ALGORITHM_AUTO = 0,             /* onesweep below 3*2^23 keys (3*2^24 pairs), reduce-then-scan from there up */
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

/* ------------------------------------------------------------------ key types, order, bit range
 *
 * Absent from the reference (CUB exposes the same knobs,
 * bench/cuda_benchmark.cu:63).  The eight vrdxCmdSort* entry points are unchanged: uint32 keys,
 * ascending, all 32 bits.  vrdxCudaCmdSortEx is the general form. */

typedef enum VrdxCudaKeyType {
  VRDX_CUDA_KEY_TYPE_UINT32 = 0,
  VRDX_CUDA_KEY_TYPE_INT32 = 1,   /* two's complement, negative before positive */
  VRDX_CUDA_KEY_TYPE_FLOAT32 = 2  /* IEEE-754 total order of the bit patterns: -NaN < -inf < ... < -0 < +0 < ... < +inf < +NaN */
} VrdxCudaKeyType;

typedef enum VrdxCudaSortOrder {
  VRDX_CUDA_SORT_ORDER_ASCENDING = 0,
  VRDX_CUDA_SORT_ORDER_DESCENDING = 1 /* still stable: equal keys keep their input order */
} VrdxCudaSortOrder;

typedef struct VrdxCudaSortKeyInfo {
  uint32_t structSize; /* = sizeof(VrdxCudaSortKeyInfo) */
  VrdxCudaKeyType keyType;
  VrdxCudaSortOrder order;
  /* Only bits [beginBit, endBit) of the key take part in the comparison (bit positions of the
   * order-preserving unsigned image of the key, as in CUB); elements equal in that range keep
   * their input order.  0 <= beginBit <= endBit <= 32.  ceil((endBit-beginBit)/8) passes run
   * instead of 4; an empty range is a no-op. */
  uint32_t beginBit;
  uint32_t endBit;
  uint32_t reserved[3]; /* zero */
} VrdxCudaSortKeyInfo;

/* General sort.  pKeyInfo NULL = uint32 / ascending / bits [0,32).  indirectBuffer NULL = direct
 * (elementCount is the count), else elementCount is maxElementCount and the uint32 count is read
 * on the device (clamped).  valuesBuffer NULL = keys only.  Storage as for vrdxCmdSort* (size it
 * with vrdxGetSorter[KeyValue]StorageRequirements).  Same stream, timestamp and error rules. */
void vrdxCudaCmdSortEx(VkCommandBuffer commandBuffer, VrdxSorter sorter,
                       const VrdxCudaSortKeyInfo* pKeyInfo, uint32_t elementCount,
                       VkBuffer indirectBuffer, VkDeviceSize indirectOffset, VkBuffer keysBuffer,
                       VkDeviceSize keysOffset, VkBuffer valuesBuffer, VkDeviceSize valuesOffset,
                       VkBuffer storageBuffer, VkDeviceSize storageOffset, VkQueryPool queryPool,
                       uint32_t query);

/* 64-bit keys (keys only).  keyType is read as its 64-bit counterpart (UINT32 -> uint64, INT32 ->
 * int64, FLOAT32 -> float64 / double, same total order rules), order as above; beginBit / endBit
 * are ignored (all 64 bits are compared).  Implemented as two chained 32-bit key-value sorts (low
 * word, then high word), so it needs its own, larger storage: vrdxCudaGetSorterKeys64Storage-
 * Requirements.  The keys buffer + offset must be 8-byte aligned.  indirectBuffer NULL = direct. */
void vrdxCudaGetSorterKeys64StorageRequirements(VrdxSorter sorter, uint32_t maxElementCount,
                                                VrdxSorterStorageRequirements* requirements);
void vrdxCudaCmdSortKeys64(VkCommandBuffer commandBuffer, VrdxSorter sorter,
                           const VrdxCudaSortKeyInfo* pKeyInfo, uint32_t elementCount,
                           VkBuffer indirectBuffer, VkDeviceSize indirectOffset, VkBuffer keysBuffer,
                           VkDeviceSize keysOffset, VkBuffer storageBuffer, VkDeviceSize storageOffset);

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

/* Import a VkSemaphore a Vulkan application exported as an opaque POSIX fd
 * (VK_KHR_external_semaphore_fd; `timeline` non-zero for a VK_SEMAPHORE_TYPE_TIMELINE one), and
 * wait for / signal it in stream order: the Vulkan queue signals `value`, the sort's stream waits
 * for it, sorts, and signals the value the Vulkan side waits for — the reference's barrier
 * contract (README.md:150-157) across the two APIs.  `value` is ignored for binary semaphores.
 * On success CUDA owns the fd. */
typedef struct VrdxCudaImportedSemaphore_T* VrdxCudaImportedSemaphore;
VkResult vrdxCudaImportSemaphoreFd(VkDevice device, int fd, int timeline,
                                   VrdxCudaImportedSemaphore* pSemaphore);
VkResult vrdxCudaCmdWaitSemaphore(VkCommandBuffer commandBuffer, VrdxCudaImportedSemaphore semaphore,
                                  uint64_t value);
VkResult vrdxCudaCmdSignalSemaphore(VkCommandBuffer commandBuffer, VrdxCudaImportedSemaphore semaphore,
                                    uint64_t value);
void vrdxCudaReleaseImportedSemaphore(VrdxCudaImportedSemaphore semaphore);

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
