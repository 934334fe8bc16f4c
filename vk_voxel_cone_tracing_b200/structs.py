"""ctypes mirrors of the POD structs in include/vgi.h (which cite the reference UBO / push-constant
structs they replace). Field order and padding are identical so the patched passes can memcpy."""
import ctypes as C

import numpy as np

VGI_MAX_LEVELS = 8
VGI_FACES = 6

VGI_OK = 0
VGI_E_INVALID = -1
VGI_E_CUDA = -2
VGI_E_STATE = -3
VGI_E_OVERFLOW = -4
VGI_E_UNSUPPORTED = -5
VGI_E_NOMEM = -6

VGI_MODE_BORDER_LITERAL = 0x1
VGI_MODE_SHADOW_COMPARE = 0x2
VGI_MODE_SVO_LITERAL = 0x4


class Config(C.Structure):
    """vgi_config — ref: VFS/Util/EngineConfig.h:28-33"""
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("resolution", C.c_uint32),
        ("level_count", C.c_uint32),
        ("downsample_band", C.c_uint32),
        ("extent_level0", C.c_float),
        ("clip_min_change", C.c_uint32 * VGI_MAX_LEVELS),
        ("max_fragments", C.c_uint32),
        ("mode_flags", C.c_uint32),
        ("device", C.c_int32),
        ("svo_max_nodes", C.c_uint32),
    ]


def default_config(resolution=128, level_count=6, **kw):
    """Reference defaults (EngineConfig.h:28-33, VoxelizationPass.h:57)."""
    c = Config()
    c.struct_size = C.sizeof(Config)
    c.resolution = resolution
    c.level_count = level_count
    c.downsample_band = 10
    c.extent_level0 = 16.0
    # VoxelizationPass.h:57 {2,2,2,2,2,1}: every level snaps by 2 voxels, the coarsest by 1
    for i in range(VGI_MAX_LEVELS):
        c.clip_min_change[i] = 2 if i < level_count - 1 else 1
    c.max_fragments = 0
    c.mode_flags = 0
    c.device = -1
    c.svo_max_nodes = 0
    for k, v in kw.items():
        setattr(c, k, v)
    return c


class Texture(C.Structure):
    """vgi_texture — one RGBA8 material texture (ref: GLTFScene::uploadImage; REPEAT + LINEAR, level 0 only, Q23)"""
    _fields_ = [("rgba8", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32)]


def texture_array(images):
    """list of (H, W, 4) uint8 arrays -> (ctypes array of vgi_texture, keep-alive list)."""
    keep = [np.ascontiguousarray(im, dtype=np.uint8) for im in images]
    arr = (Texture * max(len(keep), 1))()
    for i, im in enumerate(keep):
        assert im.ndim == 3 and im.shape[2] == 4
        arr[i].rgba8 = im.ctypes.data
        arr[i].height, arr[i].width = im.shape[:2]
    return arr, keep


class ClipRegion(C.Structure):
    """vgi_clip_region — ref: VFS/RenderPass/Clipmap/ClipmapRegion.h:8-18"""
    _fields_ = [("min_corner", C.c_int32 * 3), ("extent", C.c_uint32 * 3), ("voxel_size", C.c_float)]


class Camera(C.Structure):
    """vgi_camera — ref: VFS/Camera.h:35-41"""
    _fields_ = [("view_proj", C.c_float * 16), ("view_proj_inv", C.c_float * 16),
                ("eye_pos", C.c_float * 3), ("padding", C.c_int32)]


class DirLight(C.Structure):
    """vgi_dir_light — ref: VFS/Shaders/light.glsl:8-13"""
    _fields_ = [("direction", C.c_float * 3), ("intensity", C.c_float),
                ("color", C.c_float * 3), ("padding", C.c_int32)]


class DirLightShadow(C.Structure):
    """vgi_dir_light_shadow — ref: VFS/Shaders/light.glsl:15-20"""
    _fields_ = [("view", C.c_float * 16), ("proj", C.c_float * 16), ("z_near", C.c_float), ("z_far", C.c_float)]


class Material(C.Structure):
    """vgi_material — ref: VFS/Shaders/gltf.glsl:8-26 (80 bytes)"""
    _fields_ = [
        ("base_color_factor", C.c_float * 4),
        ("base_color_texture", C.c_int32),
        ("metallic_factor", C.c_float),
        ("roughness_factor", C.c_float),
        ("metallic_roughness_texture", C.c_int32),
        ("emissive_texture", C.c_int32),
        ("alpha_mode", C.c_int32),
        ("alpha_cutoff", C.c_float),
        ("double_sided", C.c_int32),
        ("emissive_factor", C.c_float * 3),
        ("normal_texture", C.c_int32),
        ("normal_texture_scale", C.c_float),
        ("occlusion_texture", C.c_int32),
        ("occlusion_texture_strength", C.c_float),
        ("padding", C.c_int32),
    ]


assert C.sizeof(Material) == 80

MATERIAL_DTYPE = np.dtype([
    ("base_color_factor", "<f4", 4), ("base_color_texture", "<i4"), ("metallic_factor", "<f4"),
    ("roughness_factor", "<f4"), ("metallic_roughness_texture", "<i4"), ("emissive_texture", "<i4"),
    ("alpha_mode", "<i4"), ("alpha_cutoff", "<f4"), ("double_sided", "<i4"), ("emissive_factor", "<f4", 3),
    ("normal_texture", "<i4"), ("normal_texture_scale", "<f4"), ("occlusion_texture", "<i4"),
    ("occlusion_texture_strength", "<f4"), ("padding", "<i4")])
assert MATERIAL_DTYPE.itemsize == 80

PRIMITIVE_DTYPE = np.dtype([("first_index", "<u4"), ("index_count", "<u4"), ("vertex_offset", "<u4"),
                            ("material_index", "<i4"), ("node_index", "<u4")])
NODE_DTYPE = np.dtype([("model", "<f4", 16), ("it_model", "<f4", 16)])


class Primitive(C.Structure):
    """vgi_primitive — ref: VFS/Util/GLTFLoader.h:79-90 + GLTFScene.cpp:457-490 push constants"""
    _fields_ = [("first_index", C.c_uint32), ("index_count", C.c_uint32), ("vertex_offset", C.c_uint32),
                ("material_index", C.c_int32), ("node_index", C.c_uint32)]


class NodeMatrix(C.Structure):
    """vgi_node_matrix — ref: VFS/GLTFScene.cpp:398-411"""
    _fields_ = [("model", C.c_float * 16), ("it_model", C.c_float * 16)]


class SceneDesc(C.Structure):
    """vgi_scene_desc — SoA buffers as GLTFScene uploads them (GLTFScene.cpp:55-93)"""
    _fields_ = [
        ("positions", C.c_void_p), ("normals", C.c_void_p), ("texcoords", C.c_void_p),
        ("indices", C.c_void_p), ("primitives", C.c_void_p), ("nodes", C.c_void_p),
        ("materials", C.c_void_p),
        ("vertex_count", C.c_uint32), ("index_count", C.c_uint32), ("primitive_count", C.c_uint32),
        ("node_count", C.c_uint32), ("material_count", C.c_uint32),
        ("tangents", C.c_void_p),
    ]


class GBuffer(C.Structure):
    """vgi_gbuffer — ref: VFS/RenderPass/GBufferPass.cpp:177-194"""
    _fields_ = [("diffuse_rgba8", C.c_void_p), ("normal_rgba16f", C.c_void_p), ("specular_rgba8", C.c_void_p),
                ("emission_rgba16f", C.c_void_p), ("depth_f32", C.c_void_p),
                ("width", C.c_uint32), ("height", C.c_uint32)]


class VctParams(C.Structure):
    """vgi_vct_params — ref: VFS/RenderPass/Clipmap/VoxelConeTracingPass.h:46-59 (52 bytes)"""
    _fields_ = [
        ("volume_center", C.c_float * 3), ("rendering_mode", C.c_uint32), ("voxel_size", C.c_float),
        ("volume_dimension", C.c_float), ("trace_start_offset", C.c_float),
        ("indirect_diffuse_intensity", C.c_float), ("ambient_occlusion_factor", C.c_float),
        ("min_trace_step_factor", C.c_float), ("indirect_specular_intensity", C.c_float),
        ("occlusion_decay", C.c_float), ("enable_32_cones", C.c_int32),
    ]


assert C.sizeof(VctParams) == 52


class FilterParams(C.Structure):
    """vgi_filter_params — ref: VFS/RenderPass/SpecularFilterPass.h:31-37 (16 bytes)"""
    _fields_ = [("tonemap_gamma", C.c_float), ("tonemap_exposure", C.c_float),
                ("tonemap_enable", C.c_int32), ("filter_method", C.c_int32)]


assert C.sizeof(FilterParams) == 16


def default_filter_params(filter_method=1, tonemap_enable=0):
    """Reference defaults (SpecularFilterPass.h:48-51)."""
    return FilterParams(2.2, 0.1, tonemap_enable, filter_method)


class Stats(C.Structure):
    _fields_ = [("triangles", C.c_uint64), ("clip_pairs", C.c_uint64), ("occupied_voxels", C.c_uint64),
                ("svo_fragments", C.c_uint64), ("svo_nodes", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("shaded_pairs", C.c_uint64)]


def default_vct_params(region0, resolution, rendering_mode=8):
    """Reference defaults (VoxelConeTracingPass.h:75-82) + volume fields derived from the level-0
    region (VoxelConeTracingPass.cpp:88-93)."""
    p = VctParams()
    vs = np.float32(region0.voxel_size)
    for k in range(3):
        p.volume_center[k] = float(np.float32(region0.min_corner[k]) * vs
                                   + (np.float32(region0.extent[k]) * vs) * np.float32(0.5))
    p.rendering_mode = rendering_mode
    p.voxel_size = float(vs)
    p.volume_dimension = float(resolution)
    p.trace_start_offset = 1.0
    p.indirect_diffuse_intensity = 8.0
    p.ambient_occlusion_factor = 0.5
    p.min_trace_step_factor = 1.0
    p.indirect_specular_intensity = 3.0
    p.occlusion_decay = 2.0
    p.enable_32_cones = 0
    return p


def atlas_shape(cfg):
    """(D, H, W, 4) of the reference atlas image — ref: Voxelizer.h:40-52"""
    rb = cfg.resolution + 2
    return (rb, rb * cfg.level_count, rb * VGI_FACES, 4)
