/*
 * vrdx_vk_shim.h — ABI-compatible stand-ins for the handful of Vulkan types that the
 * vk_radix_sort C API mentions, for translation units that do not have
 * <vulkan/vulkan_core.h>.  If the real Vulkan headers were included first (VULKAN_CORE_H_
 * defined) this header adds nothing, so a Vulkan application keeps its own definitions
 * and the two agree at the ABI level on every 64-bit platform:
 *
 *   dispatchable handles     (VkPhysicalDevice, VkDevice, VkCommandBuffer)  -> pointer
 *   non-dispatchable handles (VkBuffer, VkQueryPool, VkPipelineCache)        -> pointer (64-bit)
 *   VkDeviceSize -> uint64_t, VkFlags/VkBufferUsageFlags -> uint32_t, VkResult -> int32 enum
 *
 * Reference: the API includes <vulkan/vulkan_core.h> at src/vk_radix_sort.h.in:4.
 */
#ifndef VRDX_VK_SHIM_H
#define VRDX_VK_SHIM_H

#include <stdint.h>

#ifndef VULKAN_CORE_H_

#if !(defined(__LP64__) || defined(_WIN64) || defined(__x86_64__) || defined(__aarch64__))
#error "vrdx_vk_shim.h assumes 64-bit pointers (non-dispatchable Vulkan handles are pointers)"
#endif

#define VRDX_VK_SHIM_ACTIVE 1

#define VK_DEFINE_HANDLE(object) typedef struct object##_T* object;
#define VK_DEFINE_NON_DISPATCHABLE_HANDLE(object) typedef struct object##_T* object;
#define VK_NULL_HANDLE 0

typedef uint32_t VkFlags;
typedef uint64_t VkDeviceSize;
typedef VkFlags VkBufferUsageFlags;

VK_DEFINE_HANDLE(VkPhysicalDevice)
VK_DEFINE_HANDLE(VkDevice)
VK_DEFINE_HANDLE(VkCommandBuffer)
VK_DEFINE_NON_DISPATCHABLE_HANDLE(VkBuffer)
VK_DEFINE_NON_DISPATCHABLE_HANDLE(VkQueryPool)
VK_DEFINE_NON_DISPATCHABLE_HANDLE(VkPipelineCache)

/* The subset of VkResult the sorter can return; numeric values are Vulkan's. */
typedef enum VkResult {
  VK_SUCCESS = 0,
  VK_NOT_READY = 1,
  VK_ERROR_OUT_OF_HOST_MEMORY = -1,
  VK_ERROR_OUT_OF_DEVICE_MEMORY = -2,
  VK_ERROR_INITIALIZATION_FAILED = -3,
  VK_ERROR_DEVICE_LOST = -4,
  VK_ERROR_FEATURE_NOT_PRESENT = -8,
  VK_ERROR_INCOMPATIBLE_DRIVER = -9,
  VK_ERROR_UNKNOWN = -13,
  VK_RESULT_MAX_ENUM = 0x7FFFFFFF
} VkResult;

/* VkBufferUsageFlagBits used by Vrdx*StorageRequirements::usage (h.in:293, 307). */
#define VK_BUFFER_USAGE_TRANSFER_SRC_BIT 0x00000001
#define VK_BUFFER_USAGE_TRANSFER_DST_BIT 0x00000002
#define VK_BUFFER_USAGE_STORAGE_BUFFER_BIT 0x00000020

#endif /* VULKAN_CORE_H_ */
#endif /* VRDX_VK_SHIM_H */
