// VgiBridge — the one object through which the clipmap passes of vk_voxel_cone_tracing call libvgi
// (include/vgi.h of the libvgi repository). Added by patches/0004; compiled only with VFS_USE_VGI.
//
// Owns: the vgi_ctx, the exportable linear buffers shared with CUDA (VK_KHR_external_memory_fd ->
// vgi_import_vk_memory = cudaImportExternalMemory): five G-buffer planes + the shadow depth (inputs, written by
// vkCmdCopyImageToBuffer in the pre-pass command buffer) and the two float4 images of the cone trace (outputs,
// copied into the VoxelConeTracingPass attachments by vkCmdCopyBufferToImage), and optionally two exported
// binary semaphores that replace the CPU waits either side of the CUDA work.
#if !defined(VFS_VGI_BRIDGE_H)
#define VFS_VGI_BRIDGE_H

#include <pch.h>
#include <array>
#include <string>
#include <Util/EngineConfig.h>
#include <RenderPass/Clipmap/ClipmapRegion.h>
#include <VulkanFramework/Buffers/Buffer.h>
#include <VulkanFramework/Sync/Semaphore.h>

struct vgi_ctx;
struct vgi_scene_desc;
struct vgi_texture;

namespace vfs
{
	class Camera;
	class RenderPassManager;

	class VgiBridge : NonCopyable
	{
	public:
		explicit VgiBridge(DevicePtr device, RenderPassManager* renderPassManager);
				~VgiBridge();

		// the bridge of the running application (GLTFScene::initialize hands its arrays over through it)
		static VgiBridge* instance(void);

	public:
		bool initialize(uint32_t voxelResolution, uint32_t clipRegionCount, uint32_t voxelExtentLevel0, VkExtent2D resolution);
		const char* lastError(void) const;

		// GLTFScene::initialize, before releaseSourceData()
		bool setTextures(const vgi_texture* textures, uint32_t count);
		bool setScene(const vgi_scene_desc& sceneDesc);

		// pre-pass command buffer, after "GBuffer" and "RSMPass": attachments -> shared linear buffers
		void cmdCopyInputs(VkCommandBuffer cmdBuffer);
		// VoxelizationPass::onUpdate: incremental = !_fullRevoxelization (VoxelizationPass.h:59) - only the clip levels whose
		// region moved are rebuilt (vgi_build_clipmap_incremental); the whole build then runs in coneTrace()
		void setIncremental(bool incremental) { _incremental = incremental; }
		void voxelizeOpacity(const glm::vec3& cameraPos, std::array<ClipmapRegion, DEFAULT_CLIP_REGION_COUNT>* regionsOut);
		// RadianceInjectionPass::onUpdate
		void injectRadiance(uint32_t frameIndex);
		// VoxelConeTracingPass::onUpdate: the push-constant block of this frame (= vgi_vct_params, 52 bytes)
		void setConeTracingDesc(const void* desc, size_t size);
		// Application::run, between the pre-pass batch and the frame command buffer
		void coneTrace(const Camera& camera, VkExtent2D resolution);
		// VoxelConeTracingPass::onUpdate: shared buffers -> the two R32G32B32A32_SFLOAT attachments
		void cmdCopyOutputs(VkCommandBuffer cmdBuffer, VkImage diffuse, VkImage specular, uint32_t width, uint32_t height);

		// optional: exported semaphores instead of fence.waitForAllFences + cudaStreamSynchronize
		bool useSemaphores(void);
		inline VkSemaphore getInputsReadySemaphore(void) const { return _inputsReady.getHandle(); }
		inline VkSemaphore getTraceDoneSemaphore(void)   const { return _traceDone.getHandle();   }

	private:
		struct SharedBuffer
		{
			Buffer	buffer;
			void*	devicePtr	{ nullptr };	// CUDA address of the same allocation
			void*	importHandle{ nullptr };	// vgi_release_vk_memory
		};
		bool createShared(SharedBuffer* shared, uint64_t bytes, VkBufferUsageFlags usage);
		void cmdImageToBuffer(VkCommandBuffer cmdBuffer, VkImage image, VkImageAspectFlags aspect, VkImageLayout layout,
							  const SharedBuffer& dst, uint32_t width, uint32_t height);

	private:
		DevicePtr			_device;
		RenderPassManager*	_renderPassManager	{ nullptr };
		vgi_ctx*			_ctx				{ nullptr };
		void*				_cudaStream			{ nullptr };	// default stream
		VkExtent2D			_resolution			{ 0, 0 };
		VkExtent2D			_shadowResolution	{ 0, 0 };
		SharedBuffer		_gbuffer[5];		// diffuse, normal, specular, emission, depth
		SharedBuffer		_shadowDepth;
		SharedBuffer		_outDiffuse, _outSpecular;
		unsigned char		_vctDesc[52]		{};
		bool				_vctDescValid		{ false };
		bool				_sceneSet			{ false };
		bool				_pendingInject		{ false };
		bool				_incremental		{ false };
		bool				_lightValid			{ false };
		unsigned char		_lastLight[32]		{};		// vgi_dir_light / vgi_dir_light_shadow of the last vgi_set_light (light.glsl:8-20)
		unsigned char		_lastShadow[136]	{};
		uint32_t			_pendingInjectFrame	{ 0 };
		Semaphore			_inputsReady, _traceDone;
		void*				_cudaInputsReady	{ nullptr };
		void*				_cudaTraceDone		{ nullptr };
		std::string			_error;
	};
}

#endif
