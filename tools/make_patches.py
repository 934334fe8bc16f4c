"""Regenerate patches/*.patch: the Vulkan-side change set a maintainer applies to Snowapril/vk_voxel_cone_tracing so that
the three voxel-GI passes call libvgi (include/vgi.h) and share memory with CUDA.

Works on a scratch git copy of the reference (never writes to /root/reference): every edit below is an exact-text
replacement in the reference's own files, or a new file from patches/src/; `git diff` of each step becomes one patch.
tests/test_patches.py checks (where the reference checkout exists) that the committed patches apply cleanly, in order.

  python tools/make_patches.py [--reference /root/reference]
"""
import argparse
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "patches")
SRC = os.path.join(OUT, "src")


def git(cwd, *a):
    return subprocess.run(["git", "-c", "user.email=vgi@localhost", "-c", "user.name=vgi", "-c", "core.autocrlf=false", *a],
                          cwd=cwd, check=True, capture_output=True, text=True).stdout


def edit(tree, rel, pairs):
    p = os.path.join(tree, rel)
    s = open(p, newline="").read()
    for old, new in pairs:
        assert s.count(old) == 1, f"{rel}: anchor not unique / not found: {old[:60]!r}"
        s = s.replace(old, new)
    open(p, "w", newline="").write(s)


def add(tree, rel, name):
    shutil.copy(os.path.join(SRC, name), os.path.join(tree, rel))


def step_device(t):
    edit(t, "VulkanFramework/Device.cpp", [(
        "constexpr const char* REQUIRED_EXTENSIONS[] = { VK_KHR_SWAPCHAIN_EXTENSION_NAME };",
        "constexpr const char* REQUIRED_EXTENSIONS[] = {\n"
        "\tVK_KHR_SWAPCHAIN_EXTENSION_NAME,\n"
        "\t// libvgi: buffers and semaphores shared with CUDA (cudaImportExternalMemory / cudaImportExternalSemaphore)\n"
        "\tVK_KHR_EXTERNAL_MEMORY_EXTENSION_NAME,\n"
        "\tVK_KHR_EXTERNAL_MEMORY_FD_EXTENSION_NAME,\n"
        "\tVK_KHR_EXTERNAL_SEMAPHORE_EXTENSION_NAME,\n"
        "\tVK_KHR_EXTERNAL_SEMAPHORE_FD_EXTENSION_NAME,\n"
        "};")])


def step_buffer(t):
    edit(t, "VulkanFramework/Buffers/Buffer.h", [
        ("\t\tvoid downloadData\t(void* dstData, uint64_t size);\n",
         "\t\tvoid downloadData\t(void* dstData, uint64_t size);\n"
         "\t\t// libvgi: a DEVICE_LOCAL buffer in its own exportable allocation (VkExportMemoryAllocateInfo, OPAQUE_FD);\n"
         "\t\t// getMemoryFd() hands the allocation to cudaImportExternalMemory (the fd is consumed by the importer)\n"
         "\t\tbool initializeExportable(VkDevice device, VkPhysicalDevice physicalDevice, uint64_t bufferSize, VkBufferUsageFlags bufferUsage);\n"
         "\t\tint  getMemoryFd\t(void) const;\n"),
        ("\t\tuint64_t\t\t_allocatedSize\t\t{ 0 };\n",
         "\t\tuint64_t\t\t_allocatedSize\t\t{ 0 };\n"
         "\t\tVkDevice\t\t_exportDevice\t\t{ VK_NULL_HANDLE };\n"
         "\t\tVkDeviceMemory\t_exportMemory\t\t{ VK_NULL_HANDLE };\n"),
    ])
    edit(t, "VulkanFramework/Buffers/Buffer.cpp", [
        ("\tvoid Buffer::destroyBuffer(void)\n\t{\n",
         "\tvoid Buffer::destroyBuffer(void)\n\t{\n"
         "\t\tif (_exportMemory != VK_NULL_HANDLE)\n"
         "\t\t{\n"
         "\t\t\tvkDestroyBuffer(_exportDevice, _buffer, nullptr);\n"
         "\t\t\tvkFreeMemory(_exportDevice, _exportMemory, nullptr);\n"
         "\t\t\t_buffer = VK_NULL_HANDLE;\n"
         "\t\t\t_exportMemory = VK_NULL_HANDLE;\n"
         "\t\t\treturn;\n"
         "\t\t}\n"),
        ("\tvoid Buffer::uploadData(const void* srcData, uint64_t size)\n",
         "\tbool Buffer::initializeExportable(VkDevice device, VkPhysicalDevice physicalDevice, uint64_t bufferSize, VkBufferUsageFlags bufferUsage)\n"
         "\t{\n"
         "\t\t_exportDevice = device;\n"
         "\t\t_allocatedSize = bufferSize;\n"
         "\n"
         "\t\tVkExternalMemoryBufferCreateInfo externalInfo = {};\n"
         "\t\texternalInfo.sType = VK_STRUCTURE_TYPE_EXTERNAL_MEMORY_BUFFER_CREATE_INFO;\n"
         "\t\texternalInfo.handleTypes = VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT;\n"
         "\n"
         "\t\tVkBufferCreateInfo bufferInfo = {};\n"
         "\t\tbufferInfo.sType = VK_STRUCTURE_TYPE_BUFFER_CREATE_INFO;\n"
         "\t\tbufferInfo.pNext = &externalInfo;\n"
         "\t\tbufferInfo.size = bufferSize;\n"
         "\t\tbufferInfo.sharingMode = VK_SHARING_MODE_EXCLUSIVE;\n"
         "\t\tbufferInfo.usage = bufferUsage;\n"
         "\t\tif (vkCreateBuffer(device, &bufferInfo, nullptr, &_buffer) != VK_SUCCESS)\n"
         "\t\t{\n"
         "\t\t\treturn false;\n"
         "\t\t}\n"
         "\n"
         "\t\tVkMemoryRequirements requirements;\n"
         "\t\tvkGetBufferMemoryRequirements(device, _buffer, &requirements);\n"
         "\t\tVkPhysicalDeviceMemoryProperties memoryProperties;\n"
         "\t\tvkGetPhysicalDeviceMemoryProperties(physicalDevice, &memoryProperties);\n"
         "\t\tuint32_t memoryType = UINT32_MAX;\n"
         "\t\tfor (uint32_t i = 0; i < memoryProperties.memoryTypeCount; ++i)\n"
         "\t\t{\n"
         "\t\t\tif ((requirements.memoryTypeBits & (1u << i)) &&\n"
         "\t\t\t\t(memoryProperties.memoryTypes[i].propertyFlags & VK_MEMORY_PROPERTY_DEVICE_LOCAL_BIT))\n"
         "\t\t\t{\n"
         "\t\t\t\tmemoryType = i;\n"
         "\t\t\t\tbreak;\n"
         "\t\t\t}\n"
         "\t\t}\n"
         "\t\tif (memoryType == UINT32_MAX)\n"
         "\t\t{\n"
         "\t\t\treturn false;\n"
         "\t\t}\n"
         "\n"
         "\t\tVkExportMemoryAllocateInfo exportInfo = {};\n"
         "\t\texportInfo.sType = VK_STRUCTURE_TYPE_EXPORT_MEMORY_ALLOCATE_INFO;\n"
         "\t\texportInfo.handleTypes = VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT;\n"
         "\t\tVkMemoryAllocateInfo allocateInfo = {};\n"
         "\t\tallocateInfo.sType = VK_STRUCTURE_TYPE_MEMORY_ALLOCATE_INFO;\n"
         "\t\tallocateInfo.pNext = &exportInfo;\n"
         "\t\tallocateInfo.allocationSize = requirements.size;\n"
         "\t\tallocateInfo.memoryTypeIndex = memoryType;\n"
         "\t\tif (vkAllocateMemory(device, &allocateInfo, nullptr, &_exportMemory) != VK_SUCCESS)\n"
         "\t\t{\n"
         "\t\t\treturn false;\n"
         "\t\t}\n"
         "\t\t// cudaExternalMemoryHandleDesc::size must be the allocation size\n"
         "\t\t_allocatedSize = requirements.size;\n"
         "\t\treturn vkBindBufferMemory(device, _buffer, _exportMemory, 0) == VK_SUCCESS;\n"
         "\t}\n"
         "\n"
         "\tint Buffer::getMemoryFd(void) const\n"
         "\t{\n"
         "\t\tVkMemoryGetFdInfoKHR fdInfo = {};\n"
         "\t\tfdInfo.sType = VK_STRUCTURE_TYPE_MEMORY_GET_FD_INFO_KHR;\n"
         "\t\tfdInfo.memory = _exportMemory;\n"
         "\t\tfdInfo.handleType = VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT;\n"
         "\t\tPFN_vkGetMemoryFdKHR getMemoryFd = reinterpret_cast<PFN_vkGetMemoryFdKHR>(vkGetDeviceProcAddr(_exportDevice, \"vkGetMemoryFdKHR\"));\n"
         "\t\tint fd = -1;\n"
         "\t\tif (getMemoryFd == nullptr || getMemoryFd(_exportDevice, &fdInfo, &fd) != VK_SUCCESS)\n"
         "\t\t{\n"
         "\t\t\treturn -1;\n"
         "\t\t}\n"
         "\t\treturn fd;\n"
         "\t}\n"
         "\n"
         "\tvoid Buffer::uploadData(const void* srcData, uint64_t size)\n"),
    ])


def step_semaphore(t):
    edit(t, "VulkanFramework/Sync/Semaphore.h", [
        ("\t\tbool initialize(DevicePtr device);\n",
         "\t\tbool initialize(DevicePtr device);\n"
         "\t\t// libvgi: binary semaphore exportable as an opaque fd (cudaImportExternalSemaphore); getFd() transfers ownership\n"
         "\t\tbool initializeExportable(DevicePtr device);\n"
         "\t\tint  getFd(void) const;\n")])
    edit(t, "VulkanFramework/Sync/Semaphore.cpp", [
        ("\t\treturn true;\n\t}\n\n}",
         "\t\treturn true;\n\t}\n"
         "\n"
         "\tbool Semaphore::initializeExportable(DevicePtr device)\n"
         "\t{\n"
         "\t\t_device = device;\n"
         "\n"
         "\t\tVkExportSemaphoreCreateInfo exportInfo = {};\n"
         "\t\texportInfo.sType = VK_STRUCTURE_TYPE_EXPORT_SEMAPHORE_CREATE_INFO;\n"
         "\t\texportInfo.handleTypes = VK_EXTERNAL_SEMAPHORE_HANDLE_TYPE_OPAQUE_FD_BIT;\n"
         "\n"
         "\t\tVkSemaphoreCreateInfo semaphoreInfo = {};\n"
         "\t\tsemaphoreInfo.sType = VK_STRUCTURE_TYPE_SEMAPHORE_CREATE_INFO;\n"
         "\t\tsemaphoreInfo.pNext = &exportInfo;\n"
         "\t\tsemaphoreInfo.flags = 0;\n"
         "\t\treturn vkCreateSemaphore(_device->getDeviceHandle(), &semaphoreInfo, nullptr, &_semaphore) == VK_SUCCESS;\n"
         "\t}\n"
         "\n"
         "\tint Semaphore::getFd(void) const\n"
         "\t{\n"
         "\t\tVkSemaphoreGetFdInfoKHR fdInfo = {};\n"
         "\t\tfdInfo.sType = VK_STRUCTURE_TYPE_SEMAPHORE_GET_FD_INFO_KHR;\n"
         "\t\tfdInfo.semaphore = _semaphore;\n"
         "\t\tfdInfo.handleType = VK_EXTERNAL_SEMAPHORE_HANDLE_TYPE_OPAQUE_FD_BIT;\n"
         "\t\tPFN_vkGetSemaphoreFdKHR getSemaphoreFd = reinterpret_cast<PFN_vkGetSemaphoreFdKHR>(\n"
         "\t\t\tvkGetDeviceProcAddr(_device->getDeviceHandle(), \"vkGetSemaphoreFdKHR\"));\n"
         "\t\tint fd = -1;\n"
         "\t\tif (getSemaphoreFd == nullptr || getSemaphoreFd(_device->getDeviceHandle(), &fdInfo, &fd) != VK_SUCCESS)\n"
         "\t\t{\n"
         "\t\t\treturn -1;\n"
         "\t\t}\n"
         "\t\treturn fd;\n"
         "\t}\n"
         "\n}")])


def step_bridge(t):
    add(t, "VFS/RenderPass/Clipmap/VgiBridge.h", "VgiBridge.h")
    add(t, "VFS/RenderPass/Clipmap/VgiBridge.cpp", "VgiBridge.cpp")
    add(t, "VFS/GLTFSceneVgi.cpp", "GLTFSceneVgi.cpp")
    # accessors the bridge needs: the camera UBO of the current frame, the scene arrays before they are released
    edit(t, "VFS/Camera.h", [
        ("\t\tinline glm::vec3 getOriginPos(void) const\n",
         "\t\t// libvgi: {viewProj, viewProjInv, eyePos, padding} exactly as updateCamera uploads it (== vgi_camera)\n"
         "\t\tinline void readCameraUBO(void* dst) const\n"
         "\t\t{\n"
         "\t\t\tconst CameraUBO ubo = { _projMatrix * _viewMatrix, glm::inverse(_viewMatrix) * glm::inverse(_projMatrix), _position };\n"
         "\t\t\tmemcpy(dst, &ubo, sizeof(ubo));\n"
         "\t\t}\n"
         "\t\tinline glm::vec3 getOriginPos(void) const\n"),
    ])
    edit(t, "VFS/GLTFScene.h", [
        ("\t\tbool uploadMatrixBuffer\t\t(void);\n",
         "\t\tbool uploadMatrixBuffer\t\t(void);\n"
         "#if defined(VFS_USE_VGI)\n"
         "\t\tbool submitSceneToVgi\t\t(void);\t// GLTFSceneVgi.cpp\n"
         "#endif\n")])
    edit(t, "VFS/GLTFScene.cpp", [
        ("\t\t// After uploading all required vertex data and images We can release them to free\n",
         "#if defined(VFS_USE_VGI)\n"
         "\t\tsubmitSceneToVgi();\n"
         "#endif\n"
         "\t\t// After uploading all required vertex data and images We can release them to free\n")])


def step_passes(t):
    guard_begin = "#if defined(VFS_USE_VGI)\n"
    edit(t, "VFS/RenderPass/Clipmap/VoxelizationPass.cpp", [
        ("\tvoid VoxelizationPass::onBeginRenderPass(const FrameLayout* frameLayout)\n\t{\n",
         "\tvoid VoxelizationPass::onBeginRenderPass(const FrameLayout* frameLayout)\n\t{\n"
         + guard_begin +
         "\t\t// libvgi: clear, six-level voxelization, opacity mips and border wrap are one call in onUpdate\n"
         "\t\t(void)frameLayout;\n"
         "\t\treturn;\n"
         "#endif\n"),
        ("\tvoid VoxelizationPass::onEndRenderPass(const FrameLayout* frameLayout)\n\t{\n",
         "\tvoid VoxelizationPass::onEndRenderPass(const FrameLayout* frameLayout)\n\t{\n"
         + guard_begin +
         "\t\t(void)frameLayout;\n"
         "\t\treturn;\n"
         "#endif\n"),
        ("\tvoid VoxelizationPass::onUpdate(const FrameLayout* frameLayout)\n\t{\n",
         "\tvoid VoxelizationPass::onUpdate(const FrameLayout* frameLayout)\n\t{\n"
         + guard_begin +
         "\t\t// replaces updateClipRegionBoundingBox + calculateChangeDelta + the clear / 6 x (cmdVoxelize + cmdDraw) /\n"
         "\t\t// DownSampler / BorderWrapper sequence of this pass (include/vgi.h: vgi_update_regions, vgi_voxelize_opacity)\n"
         "\t\t(void)frameLayout;\n"
         "\t\tVgiBridge* bridge = _renderPassManager->get<VgiBridge>(\"VgiBridge\");\n"
         "\t\tconst glm::vec3 cameraPos = _renderPassManager->get<Camera>(\"MainCamera\")->getOriginPos();\n"
         "\t\tbridge->setIncremental(!_fullRevoxelization);\t// the reference's own switch (VoxelizationPass.h:59), never cleared there\n"
         "\t\tbridge->voxelizeOpacity(cameraPos, &_clipmapRegions);\n"
         "\t\treturn;\n"
         "#endif\n"),
        ("#include <RenderPass/Clipmap/VoxelizationPass.h>\n",
         "#include <RenderPass/Clipmap/VoxelizationPass.h>\n"
         "#if defined(VFS_USE_VGI)\n"
         "#include <RenderPass/Clipmap/VgiBridge.h>\n"
         "#include <Camera.h>\n"
         "#endif\n"),
    ])
    edit(t, "VFS/RenderPass/Clipmap/RadianceInjectionPass.cpp", [
        ("\tvoid RadianceInjectionPass::onBeginRenderPass(const FrameLayout* frameLayout)\n\t{\n",
         "\tvoid RadianceInjectionPass::onBeginRenderPass(const FrameLayout* frameLayout)\n\t{\n"
         + guard_begin +
         "\t\t(void)frameLayout;\n"
         "\t\treturn;\n"
         "#endif\n"),
        ("\tvoid RadianceInjectionPass::onEndRenderPass(const FrameLayout* frameLayout)\n\t{\n",
         "\tvoid RadianceInjectionPass::onEndRenderPass(const FrameLayout* frameLayout)\n\t{\n"
         + guard_begin +
         "\t\t(void)frameLayout;\n"
         "\t\t_frameIndex = (_frameIndex + 1);\n"
         "\t\treturn;\n"
         "#endif\n"),
        ("\tvoid RadianceInjectionPass::onUpdate(const FrameLayout* frameLayout)\n\t{\n",
         "\tvoid RadianceInjectionPass::onUpdate(const FrameLayout* frameLayout)\n\t{\n"
         + guard_begin +
         "\t\t// replaces the cadence clear, per-level injection draws, CopyAlpha and the radiance DownSampler of this pass\n"
         "\t\t// (include/vgi.h: vgi_inject_radiance; level l is refreshed when frameIndex % 2^l == 0, as kUpdateRegionLevelOffsets)\n"
         "\t\t(void)frameLayout;\n"
         "\t\t_renderPassManager->get<VgiBridge>(\"VgiBridge\")->injectRadiance(_frameIndex);\n"
         "\t\treturn;\n"
         "#endif\n"),
        ("#include <RenderPass/Clipmap/ClipmapCleaner.h>\n",
         "#include <RenderPass/Clipmap/ClipmapCleaner.h>\n"
         "#if defined(VFS_USE_VGI)\n"
         "#include <RenderPass/Clipmap/VgiBridge.h>\n"
         "#endif\n"),
    ])
    edit(t, "VFS/RenderPass/Clipmap/VoxelConeTracingPass.cpp", [
        ("\t\tcmdBuffer.pushConstants(_pipelineLayout->getLayoutHandle(), VK_SHADER_STAGE_FRAGMENT_BIT, 0, sizeof(VoxelConeTracingDesc), &vctDesc);\n"
         "\t\tcmdBuffer.draw(4, 1, 0, 0);\n",
         "#if defined(VFS_USE_VGI)\n"
         "\t\t// replaces the full-screen draw of voxelConeTracing.frag: the 52-byte push-constant block IS vgi_vct_params;\n"
         "\t\t// the images were traced by CUDA before this command buffer was recorded (Application::run), here they are\n"
         "\t\t// copied from the shared buffers into the two attachments the SpecularFilterPass samples\n"
         "\t\tstatic_assert(sizeof(VoxelConeTracingDesc) == 52, \"vgi_vct_params layout\");\n"
         "\t\tVgiBridge* bridge = _renderPassManager->get<VgiBridge>(\"VgiBridge\");\n"
         "\t\tbridge->setConeTracingDesc(&vctDesc, sizeof(vctDesc));\n"
         "\t\tbridge->cmdCopyOutputs(cmdBuffer.getHandle(), _attachments[0].image->getImageHandle(), _attachments[1].image->getImageHandle(),\n"
         "\t\t\t\t\t\t\t   _resolution.width, _resolution.height);\n"
         "#else\n"
         "\t\tcmdBuffer.pushConstants(_pipelineLayout->getLayoutHandle(), VK_SHADER_STAGE_FRAGMENT_BIT, 0, sizeof(VoxelConeTracingDesc), &vctDesc);\n"
         "\t\tcmdBuffer.draw(4, 1, 0, 0);\n"
         "#endif\n"),
        ("#include <RenderPass/Clipmap/VoxelConeTracingPass.h>\n",
         "#include <RenderPass/Clipmap/VoxelConeTracingPass.h>\n"
         "#if defined(VFS_USE_VGI)\n"
         "#include <RenderPass/Clipmap/VgiBridge.h>\n"
         "#endif\n"),
        ("\t\t_attachments.push_back({createAttachment(attachmentExtent, VK_FORMAT_R32G32B32A32_SFLOAT, VK_IMAGE_USAGE_COLOR_ATTACHMENT_BIT, VK_SAMPLE_COUNT_1_BIT) });\n",
         "\t\t_attachments.push_back({createAttachment(attachmentExtent, VK_FORMAT_R32G32B32A32_SFLOAT, VK_IMAGE_USAGE_COLOR_ATTACHMENT_BIT | VK_IMAGE_USAGE_TRANSFER_DST_BIT, VK_SAMPLE_COUNT_1_BIT) });\n"),
        ("\t\t_attachments.push_back({ createAttachment(attachmentExtent, VK_FORMAT_R32G32B32A32_SFLOAT, VK_IMAGE_USAGE_COLOR_ATTACHMENT_BIT, VK_SAMPLE_COUNT_1_BIT)});\n",
         "\t\t_attachments.push_back({ createAttachment(attachmentExtent, VK_FORMAT_R32G32B32A32_SFLOAT, VK_IMAGE_USAGE_COLOR_ATTACHMENT_BIT | VK_IMAGE_USAGE_TRANSFER_DST_BIT, VK_SAMPLE_COUNT_1_BIT)});\n"),
    ])


def step_inputs(t):
    gb = "VFS/RenderPass/GBufferPass.cpp"
    src = open(os.path.join(t, gb), newline="").read()
    pairs = []
    for fmt in ("VK_FORMAT_R8G8B8A8_UNORM, VK_IMAGE_USAGE_COLOR_ATTACHMENT_BIT, maxSampleCount",
                "VK_FORMAT_R16G16B16A16_SFLOAT, VK_IMAGE_USAGE_COLOR_ATTACHMENT_BIT, maxSampleCount"):
        assert src.count(fmt) >= 2
    src = src.replace("VK_IMAGE_USAGE_COLOR_ATTACHMENT_BIT, maxSampleCount",
                      "VK_IMAGE_USAGE_COLOR_ATTACHMENT_BIT | VK_IMAGE_USAGE_TRANSFER_SRC_BIT, maxSampleCount")
    src = src.replace("VK_FORMAT_D32_SFLOAT, VK_IMAGE_USAGE_DEPTH_STENCIL_ATTACHMENT_BIT, maxSampleCount",
                      "VK_FORMAT_D32_SFLOAT, VK_IMAGE_USAGE_DEPTH_STENCIL_ATTACHMENT_BIT | VK_IMAGE_USAGE_TRANSFER_SRC_BIT, maxSampleCount")
    open(os.path.join(t, gb), "w", newline="").write(src)
    edit(t, gb, [
        ("\t\t_renderPassManager->put(\"GBufferSampler\",\t\t_colorSampler.get());\n",
         "\t\t_renderPassManager->put(\"GBufferSampler\",\t\t_colorSampler.get());\n"
         "\t\t// libvgi: the images themselves, for the copies into the buffers shared with CUDA (VgiBridge::cmdCopyInputs)\n"
         "\t\t_renderPassManager->put(\"DiffuseImage\",\t\t_attachments[0].image.get());\n"
         "\t\t_renderPassManager->put(\"NormalImage\",\t\t_attachments[1].image.get());\n"
         "\t\t_renderPassManager->put(\"SpecularImage\",\t\t_attachments[2].image.get());\n"
         "\t\t_renderPassManager->put(\"EmissionImage\",\t\t_attachments[3].image.get());\n"
         "\t\t_renderPassManager->put(\"DepthImage\",\t\t\t_attachments[5].image.get());\n")])
    edit(t, "VFS/DirectionalLight.cpp", [
        ("\t\timageInfo.usage\t\t= VK_IMAGE_USAGE_DEPTH_STENCIL_ATTACHMENT_BIT | VK_IMAGE_USAGE_SAMPLED_BIT;\n",
         "\t\timageInfo.usage\t\t= VK_IMAGE_USAGE_DEPTH_STENCIL_ATTACHMENT_BIT | VK_IMAGE_USAGE_SAMPLED_BIT | VK_IMAGE_USAGE_TRANSFER_SRC_BIT;\n")])


def step_application(t):
    edit(t, "VFS/Application.cpp", [
        ("#include <RenderPass/Clipmap/Voxelizer.h>\n",
         "#include <RenderPass/Clipmap/Voxelizer.h>\n"
         "#if defined(VFS_USE_VGI)\n"
         "#include <RenderPass/Clipmap/VgiBridge.h>\n"
         "#endif\n"),
        ("                _graphicsQueue->submitCmdBuffer({ preCmdBuffer }, &fence);\n"
         "                fence.waitForAllFences(UINT64_MAX);\n",
         "                _graphicsQueue->submitCmdBuffer({ preCmdBuffer }, &fence);\n"
         "                fence.waitForAllFences(UINT64_MAX);\n"
         "#if defined(VFS_USE_VGI)\n"
         "                // The pre-pass batch (G-buffer, shadow map, copies into the shared buffers) is complete on the GPU: run\n"
         "                // the cone trace on CUDA and wait for it before the frame command buffer copies the two images out.\n"
         "                // (vgi_wait_vk_semaphore / vgi_signal_vk_semaphore replace both CPU waits once the submit carries the\n"
         "                // exported semaphores: VgiBridge::useSemaphores.)\n"
         "                _renderPassManager->get<VgiBridge>(\"VgiBridge\")->coneTrace(*_mainCamera, _window->getWindowExtent());\n"
         "#endif\n"),
        ("                // 3. Radiance Injection Pass (Radiance Encoding)\n",
         "#if defined(VFS_USE_VGI)\n"
         "                // shadow depth + G-buffer attachments -> exportable linear buffers libvgi reads\n"
         "                _renderPassManager->get<VgiBridge>(\"VgiBridge\")->cmdCopyInputs(preCmdBuffer.getHandle());\n"
         "#endif\n"
         "                // 3. Radiance Injection Pass (Radiance Encoding)\n"),
        ("        updateClipRegionBoundingBox();\n        \n        {\n            _voxelizer = std::make_unique<vfs::Voxelizer>(_mainCommandPool, DEFAULT_VOXEL_RESOLUTION);\n",
         "        updateClipRegionBoundingBox();\n        \n"
         "#if defined(VFS_USE_VGI)\n"
         "        {\n"
         "            // one libvgi context for the three passes; scene, light and the shared buffers are bound inside\n"
         "            _vgiBridge = std::make_unique<vfs::VgiBridge>(_device, _renderPassManager.get());\n"
         "            if (!_vgiBridge->initialize(DEFAULT_VOXEL_RESOLUTION, DEFAULT_CLIP_REGION_COUNT, DEFAULT_VOXEL_EXTENT_L0, _window->getWindowExtent()))\n"
         "            {\n"
         "                VFS_ERROR << \"libvgi : \" << _vgiBridge->lastError();\n"
         "                return false;\n"
         "            }\n"
         "            _renderPassManager->put(\"VgiBridge\", _vgiBridge.get());\n"
         "        }\n"
         "#endif\n"
         "        {\n            _voxelizer = std::make_unique<vfs::Voxelizer>(_mainCommandPool, DEFAULT_VOXEL_RESOLUTION);\n"),
    ])
    edit(t, "VFS/Application.h", [
        ("\t\tstd::shared_ptr<Voxelizer>\t\t\t\t_voxelizer;\n",
         "\t\tstd::shared_ptr<Voxelizer>\t\t\t\t_voxelizer;\n"
         "#if defined(VFS_USE_VGI)\n\t\tstd::unique_ptr<class VgiBridge>\t\t_vgiBridge;\n#endif\n"),
    ])


STEPS = [
    ("0001-device-enable-external-memory-and-semaphore-fd.patch", step_device),
    ("0002-buffer-exportable-allocation.patch", step_buffer),
    ("0003-semaphore-exportable.patch", step_semaphore),
    ("0004-vgi-bridge.patch", step_bridge),
    ("0005-gbuffer-and-shadow-map-as-transfer-sources.patch", step_inputs),
    ("0006-clipmap-passes-call-libvgi.patch", step_passes),
    ("0007-application-create-bridge-and-trace.patch", step_application),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.environ.get("VGI_REFERENCE_ROOT", "/root/reference"))
    a = ap.parse_args()
    tree = tempfile.mkdtemp(prefix="vgi_patch_")
    for d in ("VFS", "VulkanFramework"):
        shutil.copytree(os.path.join(a.reference, d), os.path.join(tree, d))
    git(tree, "init", "-q", ".")
    git(tree, "add", "-A")
    git(tree, "commit", "-qm", "reference")
    for name, fn in STEPS:
        fn(tree)
        git(tree, "add", "-A")
        diff = git(tree, "diff", "--cached", "--no-color")
        assert diff.strip(), name
        open(os.path.join(OUT, name), "w", newline="").write(diff)
        git(tree, "commit", "-qm", name)
        print("wrote", name, f"({diff.count(chr(10))} lines)")
    shutil.rmtree(tree)


if __name__ == "__main__":
    sys.exit(main())
