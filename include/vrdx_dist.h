/*
 * vrdx_dist.h — device-side building blocks of the multi-GPU sort (libvrdx_b200.so).
 *
 * The reference is single-device (no NCCL/MPI/peer code anywhere); this is new
 * work named by BASELINE.json: shard by most-significant-digit ranges, exchange over NVLink,
 * sort locally with vrdxCmdSort*.  The host side (one process per GPU, torch.distributed) is
 * vulkan_radix_sort_b200/dist.py; these two entry points are the kernels it needs besides the
 * sort itself.  Same handle conventions as vk_radix_sort.h (VkCommandBuffer = cudaStream_t,
 * VkBuffer + offset = device pointer + bytes).  Plain C ABI.
 */
#ifndef VRDX_DIST_H
#define VRDX_DIST_H

#include "vk_radix_sort.h"

#ifdef __cplusplus
extern "C" {
#endif

#define VRDX_DIST_MAX_SPLITTERS 15 /* up to 16 destination ranks */

/*
 * One level of the distributed splitter search.  For every prefix j < prefixCount, counts the keys
 * whose bits above (shift + 8) equal prefixes[j] by their 8-bit digit at `shift`:
 *     histogram[j][d] += #{ i < elementCount : (key_i >> (shift+8)) == prefixes[j]  &&  ((key_i >> shift) & 255) == d }
 * (shift = 24: the prefix is empty and every key is counted).  `histogram` (prefixCount x 256
 * uint32) must be zeroed by the caller.  prefixCount <= VRDX_DIST_MAX_SPLITTERS.
 */
void vrdxDistCmdPrefixHistogram(VkCommandBuffer commandBuffer, VrdxSorter sorter, uint32_t elementCount,
                                VkBuffer keysBuffer, VkDeviceSize keysOffset, uint32_t shift,
                                uint32_t prefixCount, VkBuffer prefixesBuffer, VkDeviceSize prefixesOffset,
                                VkBuffer histogramBuffer, VkDeviceSize histogramOffset);

/*
 * Multi-split of the local keys by `splitterCount` distinct ascending splitter values u[0..m):
 *     class(key) = 2 * #{ i : key > u[i] } + (key == u[i] for some i)          (2m + 1 classes)
 * Keys are written to outBuffer grouped by class in ascending class order (order inside a class
 * is unspecified — keys-only).  cursors (2m + 1 uint32) must hold the first output slot of every
 * class on entry (exclusive prefix of the class sizes) and are advanced to the end of each class.
 */
void vrdxDistCmdPartition(VkCommandBuffer commandBuffer, VrdxSorter sorter, uint32_t elementCount,
                          VkBuffer keysBuffer, VkDeviceSize keysOffset, uint32_t splitterCount,
                          VkBuffer splittersBuffer, VkDeviceSize splittersOffset, VkBuffer cursorsBuffer,
                          VkDeviceSize cursorsOffset, VkBuffer outBuffer, VkDeviceSize outOffset);

#ifdef __cplusplus
}
#endif
#endif /* VRDX_DIST_H */
