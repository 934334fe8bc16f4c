// VgiBridge.cpp — see VgiBridge.h. Written against the reference's own wrapper classes (Buffer, Semaphore, Image,
// RenderPassManager blackboard) and include/vgi.h; nothing else of the application changes.
#if defined(VFS_USE_VGI)

#include <pch.h>
#include <RenderPass/Clipmap/VgiBridge.h>
#include <RenderPass/RenderPassManager.h>
#include <VulkanFramework/Device.h>
#include <VulkanFramework/Images/Image.h>
#include <Camera.h>
#include <DirectionalLight.h>
#include <Common/Logger.h>
#include <cstring>
#include <vgi.h>

namespace vfs
{
	static VgiBridge* gInstance = nullptr;

	VgiBridge::VgiBridge(DevicePtr device, RenderPassManager* renderPassManager)
		: _device(device), _renderPassManager(renderPassManager)
	{
		gInstance = this;
	}

	VgiBridge::~VgiBridge()
	{
		if (_ctx != nullptr)
		{
			SharedBuffer* all[] = { &_gbuffer[0], &_gbuffer[1], &_gbuffer[2], &_gbuffer[3], &_gbuffer[4], &_shadowDepth, &_outDiffuse, &_outSpecular };
			for (SharedBuffer* shared : all)
			{
				if (shared->importHandle != nullptr)
					vgi_release_vk_memory(_ctx, shared->importHandle);
			}
			vgi_destroy(_ctx);
		}
		if (gInstance == this)
			gInstance = nullptr;
	}

	VgiBridge* VgiBridge::instance(void)
	{
		return gInstance;
	}

	const char* VgiBridge::lastError(void) const
	{
		return _error.empty() ? vgi_last_error(_ctx) : _error.c_str();
	}

	bool VgiBridge::createShared(SharedBuffer* shared, uint64_t bytes, VkBufferUsageFlags usage)
	{
		if (!shared->buffer.initializeExportable(_device->getDeviceHandle(), _device->getPhysicalDeviceHandle(), bytes, usage))
		{
			_error = "exportable buffer allocation failed";
			return false;
		}
		const int fd = shared->buffer.getMemoryFd();
		if (fd < 0)
		{
			_error = "vkGetMemoryFdKHR failed";
			return false;
		}
		// the fd is owned by CUDA after a successful import
		return vgi_import_vk_memory(_ctx, fd, shared->buffer.getTotalSize(), &shared->devicePtr, &shared->importHandle) == VGI_OK;
	}

	bool VgiBridge::initialize(uint32_t voxelResolution, uint32_t clipRegionCount, uint32_t voxelExtentLevel0, VkExtent2D resolution)
	{
		vgi_config config;
		vgi_default_config(&config);
		config.resolution	 = voxelResolution;		// DEFAULT_VOXEL_RESOLUTION
		config.level_count	 = clipRegionCount;		// DEFAULT_CLIP_REGION_COUNT
		config.extent_level0 = static_cast<float>(voxelExtentLevel0);
		if (vgi_create(&config, &_ctx) != VGI_OK)
		{
			_error = vgi_last_error(nullptr);
			return false;
		}

		_resolution = resolution;
		const uint64_t pixels = static_cast<uint64_t>(resolution.width) * resolution.height;
		const uint64_t texelBytes[5] = { 4, 8, 4, 8, 4 };	// GBufferPass.cpp:177-194 formats
		for (int i = 0; i < 5; ++i)
		{
			if (!createShared(&_gbuffer[i], pixels * texelBytes[i], VK_BUFFER_USAGE_TRANSFER_DST_BIT))
				return false;
		}
		if (!createShared(&_outDiffuse,  pixels * 16, VK_BUFFER_USAGE_TRANSFER_SRC_BIT) ||
			!createShared(&_outSpecular, pixels * 16, VK_BUFFER_USAGE_TRANSFER_SRC_BIT))
			return false;

		DirectionalLight* light = _renderPassManager->get<DirectionalLight>("DirectionalLight");
		_shadowResolution = light->getShadowMapResolution();
		return createShared(&_shadowDepth, static_cast<uint64_t>(_shadowResolution.width) * _shadowResolution.height * 4,
							VK_BUFFER_USAGE_TRANSFER_DST_BIT);
	}

	bool VgiBridge::setTextures(const vgi_texture* textures, uint32_t count)
	{
		if (vgi_set_textures(_ctx, textures, count) != VGI_OK)
		{
			VFS_ERROR << "libvgi : " << vgi_last_error(_ctx);
			return false;
		}
		return true;
	}

	bool VgiBridge::setScene(const vgi_scene_desc& sceneDesc)
	{
		if (vgi_set_scene(_ctx, &sceneDesc) != VGI_OK)
		{
			VFS_ERROR << "libvgi : " << vgi_last_error(_ctx);
			return false;
		}
		_sceneSet = true;
		return true;
	}

	void VgiBridge::cmdImageToBuffer(VkCommandBuffer cmdBuffer, VkImage image, VkImageAspectFlags aspect, VkImageLayout layout,
									 const SharedBuffer& dst, uint32_t width, uint32_t height)
	{
		VkImageMemoryBarrier toSrc = {};
		toSrc.sType = VK_STRUCTURE_TYPE_IMAGE_MEMORY_BARRIER;
		toSrc.srcAccessMask = VK_ACCESS_COLOR_ATTACHMENT_WRITE_BIT | VK_ACCESS_DEPTH_STENCIL_ATTACHMENT_WRITE_BIT;
		toSrc.dstAccessMask = VK_ACCESS_TRANSFER_READ_BIT;
		toSrc.oldLayout = layout;
		toSrc.newLayout = VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL;
		toSrc.srcQueueFamilyIndex = VK_QUEUE_FAMILY_IGNORED;
		toSrc.dstQueueFamilyIndex = VK_QUEUE_FAMILY_IGNORED;
		toSrc.image = image;
		toSrc.subresourceRange = { aspect, 0, 1, 0, 1 };
		vkCmdPipelineBarrier(cmdBuffer, VK_PIPELINE_STAGE_ALL_GRAPHICS_BIT, VK_PIPELINE_STAGE_TRANSFER_BIT, 0, 0, nullptr, 0, nullptr, 1, &toSrc);

		VkBufferImageCopy region = {};
		region.imageSubresource = { aspect, 0, 0, 1 };
		region.imageExtent = { width, height, 1 };	// bufferRowLength 0: tightly packed rows, the layout vgi_gbuffer expects
		vkCmdCopyImageToBuffer(cmdBuffer, image, VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL, dst.buffer.getBufferHandle(), 1, &region);

		VkImageMemoryBarrier back = toSrc;
		back.srcAccessMask = VK_ACCESS_TRANSFER_READ_BIT;
		back.dstAccessMask = VK_ACCESS_SHADER_READ_BIT;
		back.oldLayout = VK_IMAGE_LAYOUT_TRANSFER_SRC_OPTIMAL;
		back.newLayout = layout;
		vkCmdPipelineBarrier(cmdBuffer, VK_PIPELINE_STAGE_TRANSFER_BIT, VK_PIPELINE_STAGE_ALL_GRAPHICS_BIT, 0, 0, nullptr, 0, nullptr, 1, &back);
	}

	void VgiBridge::cmdCopyInputs(VkCommandBuffer cmdBuffer)
	{
		// "DiffuseImage" .. "DepthImage" are put on the blackboard by the GBufferPass hunk of patch 0005
		const char* names[5] = { "DiffuseImage", "NormalImage", "SpecularImage", "EmissionImage", "DepthImage" };
		for (int i = 0; i < 5; ++i)
		{
			Image* image = _renderPassManager->get<Image>(names[i]);
			const bool depth = (i == 4);
			cmdImageToBuffer(cmdBuffer, image->getImageHandle(), depth ? VK_IMAGE_ASPECT_DEPTH_BIT : VK_IMAGE_ASPECT_COLOR_BIT,
							 depth ? VK_IMAGE_LAYOUT_DEPTH_STENCIL_READ_ONLY_OPTIMAL : VK_IMAGE_LAYOUT_SHADER_READ_ONLY_OPTIMAL,
							 _gbuffer[i], _resolution.width, _resolution.height);
		}
		DirectionalLight* light = _renderPassManager->get<DirectionalLight>("DirectionalLight");
		cmdImageToBuffer(cmdBuffer, light->getShadowMap()->getImageHandle(), VK_IMAGE_ASPECT_DEPTH_BIT,
						 VK_IMAGE_LAYOUT_DEPTH_STENCIL_READ_ONLY_OPTIMAL, _shadowDepth, _shadowResolution.width, _shadowResolution.height);
	}

	void VgiBridge::voxelizeOpacity(const glm::vec3& cameraPos, std::array<ClipmapRegion, DEFAULT_CLIP_REGION_COUNT>* regionsOut)
	{
		if (!_sceneSet)
			return;
		vgi_update_regions(_ctx, &cameraPos.x);
		// "ClipmapRegions" stays valid for the GUI and for VoxelConeTracingPass::onUpdate (ClipmapRegion == vgi_clip_region)
		static_assert(sizeof(ClipmapRegion) == sizeof(vgi_clip_region), "ClipmapRegion layout");
		vgi_get_regions(_ctx, reinterpret_cast<vgi_clip_region*>(regionsOut->data()), DEFAULT_CLIP_REGION_COUNT);
		if (_incremental)
			return;		// vgi_build_clipmap_incremental does both halves at once, in coneTrace()
		if (vgi_voxelize_opacity(_ctx, _cudaStream) != VGI_OK)
			VFS_ERROR << "libvgi : " << vgi_last_error(_ctx);
	}

	void VgiBridge::injectRadiance(uint32_t frameIndex)
	{
		if (!_sceneSet)
			return;
		// the injection samples the shadow map rendered in THIS pre-pass batch: record only the parameters here, the
		// call itself runs in coneTrace() once the batch (and its copy into _shadowDepth) has executed
		DirectionalLight* light = _renderPassManager->get<DirectionalLight>("DirectionalLight");
		vgi_dir_light lightDesc;
		vgi_dir_light_shadow shadowDesc;
		static_assert(sizeof(vgi_dir_light) == 32 && sizeof(vgi_dir_light_shadow) == 136, "light.glsl:8-20");
		light->getLightDescBuffer()->downloadData(&lightDesc, sizeof(lightDesc));
		light->getViewProjectionBuffer()->downloadData(&shadowDesc, sizeof(shadowDesc));
		// vgi_set_light invalidates the incremental build's cache: hand the light over only when it changed (the shadow map
		// itself lives in the shared buffer and is read at build time; a static scene under a static light renders the same map)
		if (!_lightValid || std::memcmp(&lightDesc, _lastLight, sizeof(lightDesc)) != 0 ||
			std::memcmp(&shadowDesc, _lastShadow, sizeof(shadowDesc)) != 0)
		{
			vgi_set_light(_ctx, &lightDesc, &shadowDesc, static_cast<const float*>(_shadowDepth.devicePtr),
						  _shadowResolution.width, _shadowResolution.height, /* is_host */ 0);
			std::memcpy(_lastLight, &lightDesc, sizeof(lightDesc));
			std::memcpy(_lastShadow, &shadowDesc, sizeof(shadowDesc));
			_lightValid = true;
		}
		_pendingInjectFrame = frameIndex;
		_pendingInject = true;
	}

	void VgiBridge::setConeTracingDesc(const void* desc, size_t size)
	{
		if (size == sizeof(_vctDesc))
		{
			std::memcpy(_vctDesc, desc, size);
			_vctDescValid = true;
		}
	}

	void VgiBridge::coneTrace(const Camera& camera, VkExtent2D resolution)
	{
		if (!_sceneSet)
			return;
		if (_cudaInputsReady != nullptr)
			vgi_wait_vk_semaphore(_ctx, _cudaInputsReady, _cudaStream);
		if (_pendingInject)
		{
			const int built = _incremental ? vgi_build_clipmap_incremental(_ctx, _pendingInjectFrame, nullptr, _cudaStream)
										   : vgi_inject_radiance(_ctx, _pendingInjectFrame, _cudaStream);
			if (built != VGI_OK)
				VFS_ERROR << "libvgi : " << vgi_last_error(_ctx);
			_pendingInject = false;
		}

		vgi_camera cam;		// Camera::CameraUBO { viewProj, viewProjInv, eyePos, padding }
		static_assert(sizeof(vgi_camera) == 144, "Camera.h:37-43");
		camera.readCameraUBO(&cam);

		vgi_vct_params params;
		if (_vctDescValid)
			std::memcpy(&params, _vctDesc, sizeof(params));		// last frame's GUI values (the pass records after the trace)
		else
			vgi_default_vct_params(_ctx, &params);
		{
			// the volume fields follow this frame's level-0 region, as VoxelConeTracingPass.cpp:88-93
			vgi_vct_params fresh;
			vgi_default_vct_params(_ctx, &fresh);
			std::memcpy(params.volume_center, fresh.volume_center, sizeof(params.volume_center));
			params.voxel_size = fresh.voxel_size;
			params.volume_dimension = fresh.volume_dimension;
		}

		vgi_gbuffer gbuffer;
		gbuffer.diffuse_rgba8	 = _gbuffer[0].devicePtr;
		gbuffer.normal_rgba16f	 = _gbuffer[1].devicePtr;
		gbuffer.specular_rgba8	 = _gbuffer[2].devicePtr;
		gbuffer.emission_rgba16f = _gbuffer[3].devicePtr;
		gbuffer.depth_f32		 = static_cast<const float*>(_gbuffer[4].devicePtr);
		gbuffer.width  = resolution.width;
		gbuffer.height = resolution.height;
		if (vgi_cone_trace(_ctx, &cam, &gbuffer, &params, _outDiffuse.devicePtr, _outSpecular.devicePtr, _cudaStream) != VGI_OK)
			VFS_ERROR << "libvgi : " << vgi_last_error(_ctx);

		if (_cudaTraceDone != nullptr)
			vgi_signal_vk_semaphore(_ctx, _cudaTraceDone, _cudaStream);	// the frame submit waits on getTraceDoneSemaphore()
		else
			vgi_synchronize(_ctx, _cudaStream);
	}

	void VgiBridge::cmdCopyOutputs(VkCommandBuffer cmdBuffer, VkImage diffuse, VkImage specular, uint32_t width, uint32_t height)
	{
		const VkImage images[2] = { diffuse, specular };
		const SharedBuffer* sources[2] = { &_outDiffuse, &_outSpecular };
		for (int i = 0; i < 2; ++i)
		{
			VkImageMemoryBarrier toDst = {};
			toDst.sType = VK_STRUCTURE_TYPE_IMAGE_MEMORY_BARRIER;
			toDst.dstAccessMask = VK_ACCESS_TRANSFER_WRITE_BIT;
			toDst.oldLayout = VK_IMAGE_LAYOUT_UNDEFINED;
			toDst.newLayout = VK_IMAGE_LAYOUT_TRANSFER_DST_OPTIMAL;
			toDst.srcQueueFamilyIndex = VK_QUEUE_FAMILY_IGNORED;
			toDst.dstQueueFamilyIndex = VK_QUEUE_FAMILY_IGNORED;
			toDst.image = images[i];
			toDst.subresourceRange = { VK_IMAGE_ASPECT_COLOR_BIT, 0, 1, 0, 1 };
			vkCmdPipelineBarrier(cmdBuffer, VK_PIPELINE_STAGE_TOP_OF_PIPE_BIT, VK_PIPELINE_STAGE_TRANSFER_BIT, 0, 0, nullptr, 0, nullptr, 1, &toDst);

			VkBufferImageCopy region = {};
			region.imageSubresource = { VK_IMAGE_ASPECT_COLOR_BIT, 0, 0, 1 };
			region.imageExtent = { width, height, 1 };
			vkCmdCopyBufferToImage(cmdBuffer, sources[i]->buffer.getBufferHandle(), images[i], VK_IMAGE_LAYOUT_TRANSFER_DST_OPTIMAL, 1, &region);

			VkImageMemoryBarrier toRead = toDst;
			toRead.srcAccessMask = VK_ACCESS_TRANSFER_WRITE_BIT;
			toRead.dstAccessMask = VK_ACCESS_SHADER_READ_BIT;
			toRead.oldLayout = VK_IMAGE_LAYOUT_TRANSFER_DST_OPTIMAL;
			toRead.newLayout = VK_IMAGE_LAYOUT_SHADER_READ_ONLY_OPTIMAL;		// what SpecularFilterPass samples
			vkCmdPipelineBarrier(cmdBuffer, VK_PIPELINE_STAGE_TRANSFER_BIT, VK_PIPELINE_STAGE_FRAGMENT_SHADER_BIT, 0, 0, nullptr, 0, nullptr, 1, &toRead);
		}
	}

	bool VgiBridge::useSemaphores(void)
	{
		if (!_inputsReady.initializeExportable(_device) || !_traceDone.initializeExportable(_device))
			return false;
		const int fdIn = _inputsReady.getFd(), fdOut = _traceDone.getFd();
		if (fdIn < 0 || fdOut < 0)
			return false;
		return vgi_import_vk_semaphore(_ctx, fdIn, &_cudaInputsReady) == VGI_OK &&
			   vgi_import_vk_semaphore(_ctx, fdOut, &_cudaTraceDone) == VGI_OK;
	}
}

#endif
