/*
 * vk_radix_sort.h — the vk_radix_sort C API, served by a B200-native CUDA implementation
 * (libvrdx_b200.so).  Same eight entry points, same parameter lists, same structs as the
 * reference header (src/vk_radix_sort.h.in:1-83 of jaesung-cs/vulkan_radix_sort v0.4.0);
 * only the meaning of the opaque handles changes:
 *
 *   VkPhysicalDevice / VkDevice  ->  CUDA device ordinal + 1       (vrdxCudaDevice(), vrdx_cuda.h)
 *   VkCommandBuffer              ->  cudaStream_t   ("record" == enqueue on the stream;
 *                                    NULL is the legacy default stream; capture-safe, so a
 *                                    CUDA graph restores record-once / replay-many)
 *   VkBuffer + VkDeviceSize      ->  device pointer base + byte offset (offsets multiple of 16,
 *                                    README.md:150 of the reference)
 *   VkQueryPool                  ->  vrdxCudaCreateQueryPool() object: device memory the sort's kernels
 *                                    stamp with the GPU global timer (like vkCmdWriteTimestamp)
 *   VkPipelineCache              ->  ignored
 *
 * Unlike the reference this is a compiled library, and the declarations are extern "C" so
 * that FFI users (ctypes, cgo, JNI, ...) can bind them.  `#define VRDX_IMPLEMENTATION` is
 * accepted and ignored for source compatibility with the reference's single-header form
 * (src/vk_radix_sort.h.in:85-86).
 */
#ifndef VK_RADIX_SORT_H
#define VK_RADIX_SORT_H

#if defined(__has_include)
#if __has_include(<vulkan/vulkan_core.h>) && !defined(VRDX_FORCE_VK_SHIM)
#include <vulkan/vulkan_core.h>
#endif
#endif
#include "vrdx_vk_shim.h"

/* Version of the reference API this library is a drop-in for (src/vk_radix_sort.h.in:6-9). */
#define VRDX_VERSION_MAJOR 0
#define VRDX_VERSION_MINOR 4
#define VRDX_VERSION_PATCH 0
#define VRDX_VERSION ((VRDX_VERSION_MAJOR << 22) | (VRDX_VERSION_MINOR << 12) | VRDX_VERSION_PATCH)

#ifdef __cplusplus
extern "C" {
#endif

struct VrdxSorter_T;

/* Replaces: VK_DEFINE_HANDLE(VrdxSorter), src/vk_radix_sort.h.in:11-16.
 * The sorter is immutable after creation; all per-sort state lives in the caller's storage. */
VK_DEFINE_HANDLE(VrdxSorter)

/* Replaces: struct VrdxSorterCreateInfo, src/vk_radix_sort.h.in:18-22. */
typedef struct VrdxSorterCreateInfo {
  VkPhysicalDevice physicalDevice; /* vrdxCudaPhysicalDevice(ordinal) */
  VkDevice device;                 /* vrdxCudaDevice(ordinal) */
  VkPipelineCache pipelineCache;   /* ignored */
} VrdxSorterCreateInfo;

/* Replaces: vrdxCreateSorter, src/vk_radix_sort.h.in:24 (body :141-265).
 * Returns VK_SUCCESS, or VK_ERROR_INITIALIZATION_FAILED (bad arguments / no such device /
 * CUDA failure), VK_ERROR_FEATURE_NOT_PRESENT (device is not compute capability 10.x),
 * VK_ERROR_OUT_OF_HOST_MEMORY.  *pSorter is left unwritten on failure, like the reference. */
VkResult vrdxCreateSorter(const VrdxSorterCreateInfo* pCreateInfo, VrdxSorter* pSorter);

/* Replaces: vrdxDestroySorter, src/vk_radix_sort.h.in:26 (body :267-277). NULL is a no-op. */
void vrdxDestroySorter(VrdxSorter sorter);

/* Replaces: struct VrdxSorterStorageRequirements, src/vk_radix_sort.h.in:28-31. */
typedef struct VrdxSorterStorageRequirements {
  VkDeviceSize size;
  VkBufferUsageFlags usage; /* STORAGE_BUFFER | TRANSFER_DST, as the reference reports */
} VrdxSorterStorageRequirements;

/* Replaces: vrdxGetSorterStorageRequirements, src/vk_radix_sort.h.in:33-34 (body :279-293).
 * Pure function of maxElementCount; 64-bit arithmetic (the reference wraps at 2^30). */
void vrdxGetSorterStorageRequirements(VrdxSorter sorter, uint32_t maxElementCount,
                                      VrdxSorterStorageRequirements* requirements);

/* Replaces: vrdxGetSorterKeyValueStorageRequirements, src/vk_radix_sort.h.in:36-37 (body :295-308). */
void vrdxGetSorterKeyValueStorageRequirements(VrdxSorter sorter, uint32_t maxElementCount,
                                              VrdxSorterStorageRequirements* requirements);

/*
 * If queryPool is not VK_NULL_HANDLE the sort records 15 timestamps into
 * [query .. query+14], keeping the reference's slot layout (src/vk_radix_sort.h.in:39-50):
 *   query + 0              start
 *   query + 1              after the transfer stage (here: after the state reset + histogram kernel)
 *   query + 2 + 3*i + 0    pass i "upsweep"   \  the three stages are ONE fused onesweep kernel
 *   query + 2 + 3*i + 1    pass i "spine"      > here: upsweep and spine slots are recorded
 *   query + 2 + 3*i + 2    pass i "downsweep" /   immediately before it, downsweep after it
 *   query + 14             end
 */

/* Replaces: vrdxCmdSort, src/vk_radix_sort.h.in:51-53 (body :310-315 -> gpuSort :344-507). */
void vrdxCmdSort(VkCommandBuffer commandBuffer, VrdxSorter sorter, uint32_t elementCount,
                 VkBuffer keysBuffer, VkDeviceSize keysOffset, VkBuffer storageBuffer,
                 VkDeviceSize storageOffset, VkQueryPool queryPool, uint32_t query);

/* Replaces: vrdxCmdSortIndirect, src/vk_radix_sort.h.in:55-58 (body :317-323).
 * Reads one uint32 element count at indirectBuffer+indirectOffset on the stream's timeline;
 * no host read-back.  A count above maxElementCount is clamped (the reference: undefined). */
void vrdxCmdSortIndirect(VkCommandBuffer commandBuffer, VrdxSorter sorter, uint32_t maxElementCount,
                         VkBuffer indirectBuffer, VkDeviceSize indirectOffset, VkBuffer keysBuffer,
                         VkDeviceSize keysOffset, VkBuffer storageBuffer,
                         VkDeviceSize storageOffset, VkQueryPool queryPool, uint32_t query);

/* Replaces: vrdxCmdSortKeyValue, src/vk_radix_sort.h.in:60-63 (body :325-331). */
void vrdxCmdSortKeyValue(VkCommandBuffer commandBuffer, VrdxSorter sorter, uint32_t elementCount,
                         VkBuffer keysBuffer, VkDeviceSize keysOffset, VkBuffer valuesBuffer,
                         VkDeviceSize valuesOffset, VkBuffer storageBuffer,
                         VkDeviceSize storageOffset, VkQueryPool queryPool, uint32_t query);

/* Replaces: vrdxCmdSortKeyValueIndirect, src/vk_radix_sort.h.in:76-81 (body :333-342). */
void vrdxCmdSortKeyValueIndirect(VkCommandBuffer commandBuffer, VrdxSorter sorter,
                                 uint32_t maxElementCount, VkBuffer indirectBuffer,
                                 VkDeviceSize indirectOffset, VkBuffer keysBuffer,
                                 VkDeviceSize keysOffset, VkBuffer valuesBuffer,
                                 VkDeviceSize valuesOffset, VkBuffer storageBuffer,
                                 VkDeviceSize storageOffset, VkQueryPool queryPool, uint32_t query);

#ifdef __cplusplus
} /* extern "C" */
#endif

#endif /* VK_RADIX_SORT_H */

#ifdef VRDX_IMPLEMENTATION
#undef VRDX_IMPLEMENTATION /* compiled library: nothing to instantiate (h.in:85-86) */
#endif
