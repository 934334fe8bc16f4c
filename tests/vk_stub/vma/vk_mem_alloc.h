/* stand-in for VulkanMemoryAllocator (see ../vulkan/vulkan.h): the names the checked translation units use */
#pragma once
#include <vulkan/vulkan.h>
typedef struct VmaAllocator_T* VmaAllocator;
typedef struct VmaAllocation_T* VmaAllocation;
typedef enum VmaMemoryUsage { VMA_MEMORY_USAGE_UNKNOWN, VMA_MEMORY_USAGE_GPU_ONLY, VMA_MEMORY_USAGE_CPU_ONLY, VMA_MEMORY_USAGE_CPU_TO_GPU, VMA_MEMORY_USAGE_GPU_TO_CPU } VmaMemoryUsage;
typedef struct VmaAllocationCreateInfo { VkFlags flags; VmaMemoryUsage usage; } VmaAllocationCreateInfo;
typedef struct VmaAllocationInfo { VkDeviceSize size; } VmaAllocationInfo;
extern "C" {
VkResult vmaCreateBuffer(VmaAllocator, const VkBufferCreateInfo*, const VmaAllocationCreateInfo*, VkBuffer*, VmaAllocation*, VmaAllocationInfo*);
void vmaDestroyBuffer(VmaAllocator, VkBuffer, VmaAllocation);
VkResult vmaMapMemory(VmaAllocator, VmaAllocation, void**);
void vmaUnmapMemory(VmaAllocator, VmaAllocation);
}
