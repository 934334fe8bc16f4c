/* Minimal stand-in for <vulkan/vulkan.h>: ONLY so that tests/test_patches.py can run `g++ -fsyntax-only` over the patched
 * reference sources that touch libvgi (this image has no Vulkan SDK). Declares the handles, enums, structs and entry points
 * those translation units name; values are arbitrary. Not used by the product. */
#pragma once
#include <stdint.h>
#include <stddef.h>
#define VK_DEFINE_HANDLE(n) typedef struct n##_T* n;
VK_DEFINE_HANDLE(VkInstance) VK_DEFINE_HANDLE(VkPhysicalDevice) VK_DEFINE_HANDLE(VkDevice) VK_DEFINE_HANDLE(VkQueue)
VK_DEFINE_HANDLE(VkCommandBuffer) VK_DEFINE_HANDLE(VkBuffer) VK_DEFINE_HANDLE(VkImage) VK_DEFINE_HANDLE(VkImageView)
VK_DEFINE_HANDLE(VkDeviceMemory) VK_DEFINE_HANDLE(VkSemaphore) VK_DEFINE_HANDLE(VkFence) VK_DEFINE_HANDLE(VkSampler)
VK_DEFINE_HANDLE(VkSurfaceKHR) VK_DEFINE_HANDLE(VkDebugUtilsMessengerEXT) VK_DEFINE_HANDLE(VkDescriptorSet)
VK_DEFINE_HANDLE(VkDescriptorSetLayout) VK_DEFINE_HANDLE(VkDescriptorPool) VK_DEFINE_HANDLE(VkPipeline)
VK_DEFINE_HANDLE(VkPipelineLayout) VK_DEFINE_HANDLE(VkRenderPass) VK_DEFINE_HANDLE(VkFramebuffer) VK_DEFINE_HANDLE(VkCommandPool)
VK_DEFINE_HANDLE(VkQueryPool) VK_DEFINE_HANDLE(VkSwapchainKHR) VK_DEFINE_HANDLE(VkShaderModule) VK_DEFINE_HANDLE(VkBufferView)
#define VK_NULL_HANDLE nullptr
typedef uint32_t VkFlags; typedef uint32_t VkBool32; typedef uint64_t VkDeviceSize;
typedef VkFlags VkAccessFlags, VkImageAspectFlags, VkBufferUsageFlags, VkImageUsageFlags, VkPipelineStageFlags, VkDependencyFlags,
    VkMemoryPropertyFlags, VkExternalMemoryHandleTypeFlags, VkExternalSemaphoreHandleTypeFlags, VkSemaphoreCreateFlags,
    VkBufferCreateFlags, VkImageCreateFlags, VkSampleCountFlags, VkShaderStageFlags, VkDebugUtilsMessageTypeFlagsEXT, VkImageViewCreateFlags;
#define VK_TRUE 1u
#define VK_FALSE 0u
#define VK_QUEUE_FAMILY_IGNORED (~0u)
#define VK_KHR_SWAPCHAIN_EXTENSION_NAME "VK_KHR_swapchain"
#define VK_KHR_EXTERNAL_MEMORY_EXTENSION_NAME "VK_KHR_external_memory"
#define VK_KHR_EXTERNAL_MEMORY_FD_EXTENSION_NAME "VK_KHR_external_memory_fd"
#define VK_KHR_EXTERNAL_SEMAPHORE_EXTENSION_NAME "VK_KHR_external_semaphore"
#define VK_KHR_EXTERNAL_SEMAPHORE_FD_EXTENSION_NAME "VK_KHR_external_semaphore_fd"
typedef enum VkResult { VK_SUCCESS = 0, VK_NOT_READY = 1 } VkResult;
typedef enum VkStructureType {
    VK_STRUCTURE_TYPE_BUFFER_CREATE_INFO, VK_STRUCTURE_TYPE_EXTERNAL_MEMORY_BUFFER_CREATE_INFO, VK_STRUCTURE_TYPE_EXPORT_MEMORY_ALLOCATE_INFO,
    VK_STRUCTURE_TYPE_MEMORY_ALLOCATE_INFO, VK_STRUCTURE_TYPE_MEMORY_GET_FD_INFO_KHR, VK_STRUCTURE_TYPE_SEMAPHORE_CREATE_INFO,
    VK_STRUCTURE_TYPE_EXPORT_SEMAPHORE_CREATE_INFO, VK_STRUCTURE_TYPE_SEMAPHORE_GET_FD_INFO_KHR, VK_STRUCTURE_TYPE_IMAGE_MEMORY_BARRIER,
    VK_STRUCTURE_TYPE_BUFFER_MEMORY_BARRIER, VK_STRUCTURE_TYPE_IMAGE_CREATE_INFO, VK_STRUCTURE_TYPE_IMAGE_VIEW_CREATE_INFO,
    VK_STRUCTURE_TYPE_DEBUG_UTILS_LABEL_EXT, VK_STRUCTURE_TYPE_DEBUG_UTILS_OBJECT_NAME_INFO_EXT } VkStructureType;
typedef enum VkImageLayout { VK_IMAGE_LAYOUT_UNDEFINED, VK_IMAGE_LAYOUT_GENERAL, VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL, VK_IMAGE_LAYOUT_TRANSFER_DST_OPTIMAL,
    VK_IMAGE_LAYOUT_SHADER_READ_ONLY_OPTIMAL, VK_IMAGE_LAYOUT_DEPTH_STENCIL_READ_ONLY_OPTIMAL } VkImageLayout;
typedef enum VkSharingMode { VK_SHARING_MODE_EXCLUSIVE } VkSharingMode;
typedef enum VkImageTiling { VK_IMAGE_TILING_OPTIMAL } VkImageTiling;
typedef enum VkImageType { VK_IMAGE_TYPE_1D, VK_IMAGE_TYPE_2D, VK_IMAGE_TYPE_3D, VK_IMAGE_TYPE_MAX_ENUM = 0x7fffffff } VkImageType;
typedef enum VkImageViewType { VK_IMAGE_VIEW_TYPE_2D, VK_IMAGE_VIEW_TYPE_3D } VkImageViewType;
typedef enum VkFormat { VK_FORMAT_UNDEFINED, VK_FORMAT_D32_SFLOAT, VK_FORMAT_R8G8B8A8_UNORM, VK_FORMAT_R16G16B16A16_SFLOAT, VK_FORMAT_R32G32B32A32_SFLOAT, VK_FORMAT_MAX_ENUM = 0x7fffffff } VkFormat;
typedef enum VkSampleCountFlagBits { VK_SAMPLE_COUNT_1_BIT = 1 } VkSampleCountFlagBits;
typedef enum VkObjectType { VK_OBJECT_TYPE_UNKNOWN, VK_OBJECT_TYPE_BUFFER, VK_OBJECT_TYPE_IMAGE, VK_OBJECT_TYPE_IMAGE_VIEW, VK_OBJECT_TYPE_SAMPLER,
    VK_OBJECT_TYPE_PIPELINE, VK_OBJECT_TYPE_RENDER_PASS, VK_OBJECT_TYPE_FENCE, VK_OBJECT_TYPE_SEMAPHORE, VK_OBJECT_TYPE_DESCRIPTOR_SET,
    VK_OBJECT_TYPE_BUFFER_VIEW, VK_OBJECT_TYPE_COMMAND_BUFFER, VK_OBJECT_TYPE_QUEUE, VK_OBJECT_TYPE_DEVICE, VK_OBJECT_TYPE_FRAMEBUFFER,
    VK_OBJECT_TYPE_DESCRIPTOR_SET_LAYOUT, VK_OBJECT_TYPE_DESCRIPTOR_POOL, VK_OBJECT_TYPE_PIPELINE_LAYOUT, VK_OBJECT_TYPE_SHADER_MODULE,
    VK_OBJECT_TYPE_COMMAND_POOL, VK_OBJECT_TYPE_DEVICE_MEMORY, VK_OBJECT_TYPE_QUERY_POOL, VK_OBJECT_TYPE_SWAPCHAIN_KHR, VK_OBJECT_TYPE_SURFACE_KHR } VkObjectType;
typedef enum VkComponentSwizzle { VK_COMPONENT_SWIZZLE_IDENTITY, VK_COMPONENT_SWIZZLE_R, VK_COMPONENT_SWIZZLE_G, VK_COMPONENT_SWIZZLE_B, VK_COMPONENT_SWIZZLE_A } VkComponentSwizzle;
enum { VK_ACCESS_TRANSFER_READ_BIT = 1, VK_ACCESS_TRANSFER_WRITE_BIT = 2, VK_ACCESS_SHADER_READ_BIT = 4, VK_ACCESS_COLOR_ATTACHMENT_WRITE_BIT = 8,
       VK_ACCESS_DEPTH_STENCIL_ATTACHMENT_WRITE_BIT = 16, VK_ACCESS_SHADER_WRITE_BIT = 32 };
enum { VK_IMAGE_ASPECT_COLOR_BIT = 1, VK_IMAGE_ASPECT_DEPTH_BIT = 2 };
enum { VK_BUFFER_USAGE_TRANSFER_SRC_BIT = 1, VK_BUFFER_USAGE_TRANSFER_DST_BIT = 2 };
enum { VK_PIPELINE_STAGE_TOP_OF_PIPE_BIT = 1, VK_PIPELINE_STAGE_TRANSFER_BIT = 2, VK_PIPELINE_STAGE_ALL_GRAPHICS_BIT = 4, VK_PIPELINE_STAGE_FRAGMENT_SHADER_BIT = 8 };
enum { VK_MEMORY_PROPERTY_DEVICE_LOCAL_BIT = 1 };
enum { VK_IMAGE_USAGE_TRANSFER_SRC_BIT = 1, VK_IMAGE_USAGE_TRANSFER_DST_BIT = 2, VK_IMAGE_USAGE_SAMPLED_BIT = 4, VK_IMAGE_USAGE_COLOR_ATTACHMENT_BIT = 16,
       VK_IMAGE_USAGE_DEPTH_STENCIL_ATTACHMENT_BIT = 32 };
enum { VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT = 1, VK_EXTERNAL_SEMAPHORE_HANDLE_TYPE_OPAQUE_FD_BIT = 1 };
typedef VkFlags VkExternalMemoryHandleTypeFlagBits, VkExternalSemaphoreHandleTypeFlagBits;
typedef struct VkExtent2D { uint32_t width, height; } VkExtent2D;
typedef struct VkExtent3D { uint32_t width, height, depth; } VkExtent3D;
typedef struct VkOffset3D { int32_t x, y, z; } VkOffset3D;
typedef struct VkImageSubresourceRange { VkImageAspectFlags aspectMask; uint32_t baseMipLevel, levelCount, baseArrayLayer, layerCount; } VkImageSubresourceRange;
typedef struct VkImageSubresourceLayers { VkImageAspectFlags aspectMask; uint32_t mipLevel, baseArrayLayer, layerCount; } VkImageSubresourceLayers;
typedef struct VkComponentMapping { VkComponentSwizzle r, g, b, a; } VkComponentMapping;
typedef struct VkImageMemoryBarrier { VkStructureType sType; const void* pNext; VkAccessFlags srcAccessMask, dstAccessMask; VkImageLayout oldLayout, newLayout;
    uint32_t srcQueueFamilyIndex, dstQueueFamilyIndex; VkImage image; VkImageSubresourceRange subresourceRange; } VkImageMemoryBarrier;
typedef struct VkBufferMemoryBarrier { VkStructureType sType; const void* pNext; VkAccessFlags srcAccessMask, dstAccessMask;
    uint32_t srcQueueFamilyIndex, dstQueueFamilyIndex; VkBuffer buffer; VkDeviceSize offset, size; } VkBufferMemoryBarrier;
typedef struct VkMemoryBarrier { VkStructureType sType; const void* pNext; VkAccessFlags srcAccessMask, dstAccessMask; } VkMemoryBarrier;
typedef struct VkBufferImageCopy { VkDeviceSize bufferOffset; uint32_t bufferRowLength, bufferImageHeight; VkImageSubresourceLayers imageSubresource;
    VkOffset3D imageOffset; VkExtent3D imageExtent; } VkBufferImageCopy;
typedef struct VkBufferCreateInfo { VkStructureType sType; const void* pNext; VkBufferCreateFlags flags; VkDeviceSize size; VkBufferUsageFlags usage;
    VkSharingMode sharingMode; uint32_t queueFamilyIndexCount; const uint32_t* pQueueFamilyIndices; } VkBufferCreateInfo;
typedef struct VkImageCreateInfo { VkStructureType sType; const void* pNext; VkImageCreateFlags flags; VkImageType imageType; VkFormat format; VkExtent3D extent;
    uint32_t mipLevels, arrayLayers; VkSampleCountFlagBits samples; VkImageTiling tiling; VkImageUsageFlags usage; VkSharingMode sharingMode;
    uint32_t queueFamilyIndexCount; const uint32_t* pQueueFamilyIndices; VkImageLayout initialLayout; } VkImageCreateInfo;
typedef struct VkImageViewCreateInfo { VkStructureType sType; const void* pNext; VkImageViewCreateFlags flags; VkImage image; VkImageViewType viewType; VkFormat format;
    VkComponentMapping components; VkImageSubresourceRange subresourceRange; } VkImageViewCreateInfo;
typedef struct VkExternalMemoryBufferCreateInfo { VkStructureType sType; const void* pNext; VkExternalMemoryHandleTypeFlags handleTypes; } VkExternalMemoryBufferCreateInfo;
typedef struct VkExportMemoryAllocateInfo { VkStructureType sType; const void* pNext; VkExternalMemoryHandleTypeFlags handleTypes; } VkExportMemoryAllocateInfo;
typedef struct VkMemoryAllocateInfo { VkStructureType sType; const void* pNext; VkDeviceSize allocationSize; uint32_t memoryTypeIndex; } VkMemoryAllocateInfo;
typedef struct VkMemoryRequirements { VkDeviceSize size, alignment; uint32_t memoryTypeBits; } VkMemoryRequirements;
typedef struct VkMemoryType { VkMemoryPropertyFlags propertyFlags; uint32_t heapIndex; } VkMemoryType;
typedef struct VkMemoryHeap { VkDeviceSize size; VkFlags flags; } VkMemoryHeap;
typedef struct VkPhysicalDeviceMemoryProperties { uint32_t memoryTypeCount; VkMemoryType memoryTypes[32]; uint32_t memoryHeapCount; VkMemoryHeap memoryHeaps[16]; } VkPhysicalDeviceMemoryProperties;
typedef struct VkMemoryGetFdInfoKHR { VkStructureType sType; const void* pNext; VkDeviceMemory memory; VkExternalMemoryHandleTypeFlagBits handleType; } VkMemoryGetFdInfoKHR;
typedef struct VkSemaphoreCreateInfo { VkStructureType sType; const void* pNext; VkSemaphoreCreateFlags flags; } VkSemaphoreCreateInfo;
typedef struct VkExportSemaphoreCreateInfo { VkStructureType sType; const void* pNext; VkExternalSemaphoreHandleTypeFlags handleTypes; } VkExportSemaphoreCreateInfo;
typedef struct VkSemaphoreGetFdInfoKHR { VkStructureType sType; const void* pNext; VkSemaphore semaphore; VkExternalSemaphoreHandleTypeFlagBits handleType; } VkSemaphoreGetFdInfoKHR;
typedef struct VkAllocationCallbacks VkAllocationCallbacks;
typedef struct VkQueueFamilyProperties { VkFlags queueFlags; uint32_t queueCount; } VkQueueFamilyProperties;
typedef struct VkPhysicalDeviceProperties { uint32_t apiVersion; } VkPhysicalDeviceProperties;
typedef struct VkPhysicalDeviceFeatures { VkBool32 samplerAnisotropy; } VkPhysicalDeviceFeatures;
typedef struct VkDebugUtilsLabelEXT { VkStructureType sType; const void* pNext; const char* pLabelName; float color[4]; } VkDebugUtilsLabelEXT;
typedef struct VkDebugUtilsObjectNameInfoEXT { VkStructureType sType; const void* pNext; VkObjectType objectType; uint64_t objectHandle; const char* pObjectName; } VkDebugUtilsObjectNameInfoEXT;
#define VKAPI_ATTR
#define VKAPI_CALL
#define VKAPI_PTR
typedef enum VkDebugUtilsMessageSeverityFlagBitsEXT { VK_DEBUG_UTILS_MESSAGE_SEVERITY_ERROR_BIT_EXT = 1 } VkDebugUtilsMessageSeverityFlagBitsEXT;
typedef struct VkDebugUtilsMessengerCallbackDataEXT { VkStructureType sType; const void* pNext; const char* pMessage; } VkDebugUtilsMessengerCallbackDataEXT;
typedef struct VkDebugUtilsMessengerCreateInfoEXT { VkStructureType sType; const void* pNext; } VkDebugUtilsMessengerCreateInfoEXT;
typedef void (*PFN_vkVoidFunction)(void);
typedef VkResult (*PFN_vkGetMemoryFdKHR)(VkDevice, const VkMemoryGetFdInfoKHR*, int*);
typedef VkResult (*PFN_vkGetSemaphoreFdKHR)(VkDevice, const VkSemaphoreGetFdInfoKHR*, int*);
extern "C" {
PFN_vkVoidFunction vkGetDeviceProcAddr(VkDevice, const char*);
VkResult vkCreateBuffer(VkDevice, const VkBufferCreateInfo*, const VkAllocationCallbacks*, VkBuffer*);
void vkDestroyBuffer(VkDevice, VkBuffer, const VkAllocationCallbacks*);
void vkGetBufferMemoryRequirements(VkDevice, VkBuffer, VkMemoryRequirements*);
void vkGetPhysicalDeviceMemoryProperties(VkPhysicalDevice, VkPhysicalDeviceMemoryProperties*);
VkResult vkAllocateMemory(VkDevice, const VkMemoryAllocateInfo*, const VkAllocationCallbacks*, VkDeviceMemory*);
void vkFreeMemory(VkDevice, VkDeviceMemory, const VkAllocationCallbacks*);
VkResult vkBindBufferMemory(VkDevice, VkBuffer, VkDeviceMemory, VkDeviceSize);
VkResult vkCreateSemaphore(VkDevice, const VkSemaphoreCreateInfo*, const VkAllocationCallbacks*, VkSemaphore*);
void vkDestroySemaphore(VkDevice, VkSemaphore, const VkAllocationCallbacks*);
void vkCmdPipelineBarrier(VkCommandBuffer, VkPipelineStageFlags, VkPipelineStageFlags, VkDependencyFlags, uint32_t, const VkMemoryBarrier*,
                          uint32_t, const VkBufferMemoryBarrier*, uint32_t, const VkImageMemoryBarrier*);
void vkCmdCopyImageToBuffer(VkCommandBuffer, VkImage, VkImageLayout, VkBuffer, uint32_t, const VkBufferImageCopy*);
void vkCmdBeginDebugUtilsLabelEXT(VkCommandBuffer, const VkDebugUtilsLabelEXT*);
void vkCmdEndDebugUtilsLabelEXT(VkCommandBuffer);
void vkCmdInsertDebugUtilsLabelEXT(VkCommandBuffer, const VkDebugUtilsLabelEXT*);
VkResult vkSetDebugUtilsObjectNameEXT(VkDevice, const VkDebugUtilsObjectNameInfoEXT*);
void vkCmdCopyBufferToImage(VkCommandBuffer, VkBuffer, VkImage, VkImageLayout, uint32_t, const VkBufferImageCopy*);
}
