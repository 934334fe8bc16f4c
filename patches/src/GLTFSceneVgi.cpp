// GLTFSceneVgi.cpp — hands the scene arrays to libvgi before GLTFScene::initialize releases them.
// Same data as uploadBuffer / uploadMaterialBuffer / uploadMatrixBuffer put into the Vulkan buffers, and the same
// {instanceIndex, materialIndex} pairs GLTFScene::cmdDraw pushes per primitive.
#if defined(VFS_USE_VGI)

#include <pch.h>
#include <GLTFScene.h>
#include <RenderPass/Clipmap/VgiBridge.h>
#include <vector>
#include <vgi.h>

namespace vfs
{
	bool GLTFScene::submitSceneToVgi(void)
	{
		VgiBridge* bridge = VgiBridge::instance();
		if (bridge == nullptr)
			return false;

		std::vector<vgi_material> materials;
		materials.reserve(_sceneMaterials.size());
		for (const GLTFMaterial& material : _sceneMaterials)
		{
			vgi_material m = {};		// 80-byte GltfShadeMaterial (gltf.glsl:8-26)
			for (int k = 0; k < 4; ++k) m.base_color_factor[k] = material.baseColorFactor[k];
			m.base_color_texture		 = material.baseColorTexture;
			m.metallic_factor			 = material.metallicFactor;
			m.roughness_factor			 = material.roughnessFactor;
			m.metallic_roughness_texture = material.metallicRoughnessTexture;
			m.emissive_texture			 = material.emissiveTexture;
			m.alpha_mode				 = material.alphaMode;
			m.alpha_cutoff				 = material.alphaCutoff;
			m.double_sided				 = material.doubleSided;
			for (int k = 0; k < 3; ++k) m.emissive_factor[k] = material.emissiveFactor[k];
			m.normal_texture			 = material.normalTexture;
			m.normal_texture_scale		 = material.normalTextureScale;
			m.occlusion_texture			 = material.occlusionTexture;
			m.occlusion_texture_strength = material.occlusionTextureStrength;
			materials.push_back(m);
		}

		std::vector<vgi_node_matrix> matrices;		// uploadMatrixBuffer: nodes that own primitives, in node order
		std::vector<vgi_primitive> primitives;		// cmdDraw: one entry per drawIndexed
		uint32_t instanceIndex = 0;
		for (const GLTFNode& node : _sceneNodes)
		{
			if (!node.primMeshes.empty())
			{
				vgi_node_matrix nm;
				const glm::mat4 itWorld = glm::transpose(glm::inverse(node.world));
				std::memcpy(nm.model, &node.world[0][0], sizeof(nm.model));
				std::memcpy(nm.it_model, &itWorld[0][0], sizeof(nm.it_model));
				matrices.push_back(nm);
			}
			for (uint32_t meshIdx : node.primMeshes)
			{
				const GLTFPrimMesh& prim = _scenePrimMeshes[meshIdx];
				primitives.push_back({ prim.firstIndex, prim.indexCount, prim.vertexOffset, prim.materialIndex, instanceIndex });
			}
			++instanceIndex;
		}

		// the images GLTFScene::uploadImage turns into uTextures[] (RGBA8, sampled REPEAT + LINEAR at level 0)
		std::vector<vgi_texture> textures;
		textures.reserve(_images.size());
		for (const GLTFImage& image : _images)
		{
			textures.push_back({ image.data.data(), image.width, image.height });
		}
		if (!bridge->setTextures(textures.data(), static_cast<uint32_t>(textures.size())))
			return false;

		vgi_scene_desc desc = {};
		desc.positions	= reinterpret_cast<const float*>(_positions.data());
		desc.normals	= reinterpret_cast<const float*>(_normals.data());
		desc.texcoords	= _texCoords.empty() ? nullptr : reinterpret_cast<const float*>(_texCoords.data());
		desc.indices	= _indices.data();
		desc.primitives = primitives.data();
		desc.nodes		= matrices.data();
		desc.materials	= materials.data();
		desc.vertex_count	 = static_cast<uint32_t>(_positions.size());
		desc.index_count	 = static_cast<uint32_t>(_indices.size());
		desc.primitive_count = static_cast<uint32_t>(primitives.size());
		desc.node_count		 = static_cast<uint32_t>(matrices.size());
		desc.material_count	 = static_cast<uint32_t>(materials.size());
		desc.tangents	= _tangents.empty() ? nullptr : reinterpret_cast<const float*>(_tangents.data());
		return bridge->setScene(desc);
	}
}

#endif
