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
 * whose bits above (shift + digitBits) equal prefixes[j] by their digitBits-wide digit at `shift`:
 *     histogram[j][d] += #{ i < elementCount : (key_i >> (shift+digitBits)) == prefixes[j]
 *                                              &&  ((key_i >> shift) & (2^digitBits - 1)) == d }
 * (shift + digitBits = 32: the prefix is empty and every key is counted).  `histogram`
 * (prefixCount x 2^digitBits uint32) must be zeroed by the caller.  digitBits in [1, 12],
 * prefixCount <= VRDX_DIST_MAX_SPLITTERS and prefixCount * 2^digitBits * 4 bytes <= 160 KiB.
 */
void vrdxDistCmdPrefixHistogram(VkCommandBuffer commandBuffer, VrdxSorter sorter, uint32_t elementCount,
                                VkBuffer keysBuffer, VkDeviceSize keysOffset, uint32_t shift, uint32_t digitBits,
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

/*
 * Class sizes for the same splitters, without moving anything: counts[c] += #{ keys of class c }
 * (2 * splitterCount + 1 uint32, zeroed by the caller).  One 4 B/key read.
 */
void vrdxDistCmdClassCount(VkCommandBuffer commandBuffer, VrdxSorter sorter, uint32_t elementCount,
                           VkBuffer keysBuffer, VkDeviceSize keysOffset, uint32_t splitterCount,
                           VkBuffer splittersBuffer, VkDeviceSize splittersOffset, VkBuffer countsBuffer,
                           VkDeviceSize countsOffset);

/*
 * Partition + exchange in ONE kernel.  Same multi-split as vrdxDistCmdPartition, but a key whose
 * class-ordered position is p, with firstPosition[j] <= p < firstPosition[j+1], is stored straight
 * into destination j's receive buffer at destPointers[j][p - firstPosition[j]] — peer memory
 * mapped with vrdxDistOpenShared(), i.e. the stores travel over NVLink / NVSwitch while the kernel
 * is still ranking other tiles.  No separate all-to-all; the caller only needs a barrier across
 * ranks before anyone reads its receive buffer.
 * destTable (device memory): uint64 destPointers[destCount] followed by uint32 firstPosition[destCount + 1].
 */
void vrdxDistCmdPartitionScatter(VkCommandBuffer commandBuffer, VrdxSorter sorter, uint32_t elementCount,
                                 VkBuffer keysBuffer, VkDeviceSize keysOffset, uint32_t splitterCount,
                                 VkBuffer splittersBuffer, VkDeviceSize splittersOffset, VkBuffer cursorsBuffer,
                                 VkDeviceSize cursorsOffset, uint32_t destCount, VkBuffer destTableBuffer,
                                 VkDeviceSize destTableOffset);

/* Device memory that other processes on the node can map (cudaMalloc + CUDA IPC). */
#define VRDX_DIST_IPC_HANDLE_BYTES 64
VkResult vrdxDistAllocShared(VkDevice device, VkDeviceSize size, VkBuffer* pBuffer,
                             unsigned char handle[VRDX_DIST_IPC_HANDLE_BYTES]);
void vrdxDistFreeShared(VkDevice device, VkBuffer buffer);
VkResult vrdxDistOpenShared(VkDevice device, const unsigned char handle[VRDX_DIST_IPC_HANDLE_BYTES],
                            VkBuffer* pBuffer);
void vrdxDistCloseShared(VkDevice device, VkBuffer buffer);

#ifdef __cplusplus
}
#endif
#endif /* VRDX_DIST_H */
