/*
 * vgi.h — C ABI of libvgi.so: the B200-native voxel-GI hot path.
 *
 * This is the drop-in boundary for the voxel-GI path of Snowapril/vk_voxel_cone_tracing.
 * The reference has no plugin/FFI interface; its seam is the C++ RenderPassBase hooks
 * (VFS/RenderPass/RenderPassBase.h:56-58) plus the helper entry points that the passes look up
 * from the RenderPassManager blackboard (VFS/RenderPass/RenderPassManager.h:42-53).  Every entry
 * point below names the reference function(s) it replaces.  All "ref:" paths are relative to the
 * reference checkout.
 *
 * Conventions
 *   - every function returns int: VGI_OK (0) or a negative VGI_E_* code; vgi_last_error() gives text;
 *     nothing aborts or throws across the ABI (the reference uses bool + side-effecting asserts).
 *   - matrices are column-major float[16] exactly as glm::mat4 lays them out in the reference UBOs.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); calls are asynchronous
 *     unless stated otherwise.  One ctx per GPU, not thread-safe (the reference host is single-threaded).
 *   - pointer arguments are DEVICE pointers unless the parameter/flag says host.
 *   - there is no CPU fallback: if no CUDA device is usable vgi_create fails with VGI_E_CUDA.
 */
#ifndef VGI_H
#define VGI_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VGI_VERSION 100

enum {
    VGI_OK            = 0,
    VGI_E_INVALID     = -1,  /* bad argument / struct */
    VGI_E_CUDA        = -2,  /* CUDA runtime error (see vgi_last_error) */
    VGI_E_STATE       = -3,  /* call order violated (e.g. inject before voxelize) */
    VGI_E_OVERFLOW    = -4,  /* a bounded device list (fragments, nodes) overflowed */
    VGI_E_UNSUPPORTED = -5,  /* feature outside the hot path (e.g. textured materials) */
    VGI_E_NOMEM       = -6
};

#define VGI_MAX_LEVELS 8
#define VGI_FACES      6

/* mode_flags: switches between the canonical semantics and
 * the literal reference behaviour where the two differ. Default 0 = canonical. */
#define VGI_MODE_BORDER_LITERAL   0x1u  /* Q4/Q5: export only the low opacity border, radiance border 0 */
#define VGI_MODE_SHADOW_COMPARE   0x2u  /* Q1 fix: depth comparison instead of averaged raw depth */
#define VGI_MODE_SVO_LITERAL      0x4u  /* Q13/Q21/Q22 literal SVO fragment shading */

/* ref: VFS/Util/EngineConfig.h:28-33 (compile-time constants there, runtime fields here). */
typedef struct vgi_config {
    uint32_t struct_size;          /* = sizeof(vgi_config) */
    uint32_t resolution;           /* R: DEFAULT_VOXEL_RESOLUTION (128); power of two, 32..512 */
    uint32_t level_count;          /* L: DEFAULT_CLIP_REGION_COUNT (6); 1..VGI_MAX_LEVELS */
    uint32_t downsample_band;      /* DEFAULT_DOWNSAMPLE_REGION_SIZE (10) */
    float    extent_level0;        /* DEFAULT_VOXEL_EXTENT_L0 (16) world units */
    uint32_t clip_min_change[VGI_MAX_LEVELS]; /* ref: VoxelizationPass.h:57 {2,2,2,2,2,1} */
    uint32_t max_fragments;        /* capacity of the (triangle,voxel) pair list; 0 = default */
    uint32_t mode_flags;           /* VGI_MODE_* */
    int32_t  device;               /* CUDA device ordinal, -1 = current */
    uint32_t svo_max_nodes;        /* SVO node pool capacity, 0 = clamp(8*Nfrag,1e6,5e8) ref: OctreeBuilder.cpp:110-113 */
} vgi_config;

/* ref: VFS/RenderPass/Clipmap/ClipmapRegion.h:8-18 */
typedef struct vgi_clip_region {
    int32_t  min_corner[3];        /* in voxels of this level */
    uint32_t extent[3];            /* = R */
    float    voxel_size;
} vgi_clip_region;

/* ref: VFS/Camera.h:35-41 (CameraUBO) */
typedef struct vgi_camera {
    float   view_proj[16];
    float   view_proj_inv[16];
    float   eye_pos[3];
    int32_t padding;
} vgi_camera;

/* ref: VFS/Shaders/light.glsl:8-13 */
typedef struct vgi_dir_light {
    float   direction[3];
    float   intensity;
    float   color[3];
    int32_t padding;
} vgi_dir_light;

/* ref: VFS/Shaders/light.glsl:15-20 */
typedef struct vgi_dir_light_shadow {
    float view[16];
    float proj[16];
    float z_near;
    float z_far;
} vgi_dir_light_shadow;

/* ref: VFS/Shaders/gltf.glsl:8-26 (80-byte GltfShadeMaterial, std430) */
typedef struct vgi_material {
    float   base_color_factor[4];
    int32_t base_color_texture;
    float   metallic_factor;
    float   roughness_factor;
    int32_t metallic_roughness_texture;
    int32_t emissive_texture;
    int32_t alpha_mode;
    float   alpha_cutoff;
    int32_t double_sided;
    float   emissive_factor[3];
    int32_t normal_texture;
    float   normal_texture_scale;
    int32_t occlusion_texture;
    float   occlusion_texture_strength;
    int32_t padding;
} vgi_material;

/* ref: VFS/Util/GLTFLoader.h:79-90 (GLTFPrimMesh) + push constants {instanceIndex, materialIndex}
 * issued per primitive by GLTFScene::cmdDraw (VFS/GLTFScene.cpp:457-490). */
typedef struct vgi_primitive {
    uint32_t first_index;
    uint32_t index_count;
    uint32_t vertex_offset;
    int32_t  material_index;
    uint32_t node_index;           /* uInstanceIndex */
} vgi_primitive;

/* ref: VFS/GLTFScene.cpp:398-411 — {world, transpose(inverse(world))} */
typedef struct vgi_node_matrix {
    float model[16];
    float it_model[16];
} vgi_node_matrix;

/* One material texture as GLTFScene::uploadImage creates it (VFS/GLTFScene.cpp:196-260): RGBA8 UNORM, row-major, sampled
 * with REPEAT addressing and LINEAR filtering at mip level 0 only (the reference builds a mip chain but creates its samplers
 * with maxLod = 0, GLTFScene.cpp:339 / Sampler.cpp:46, so every texture() / textureLod() of the path is a bi-linear read
 * of level 0: quirk Q23). HOST pointer; vgi_set_textures copies. */
typedef struct vgi_texture {
    const void* rgba8;
    uint32_t    width, height;
} vgi_texture;

/* Scene buffers in the SoA layout GLTFScene uploads (VFS/GLTFScene.cpp:55-93). HOST pointers;
 * vgi_set_scene copies what it needs. Materials may reference textures (indices into the array given to vgi_set_textures):
 * the voxelization and injection stages use base_color_texture, emissive_texture and occlusion_texture as
 * msaaVoxelizer.frag:64, msaaInjectRadiance.frag:73-82,131-136 and voxelizer.frag:52-76 do; vgi_render_gbuffer uses
 * base colour, metallic-roughness, emissive and normal textures and the alpha cutoff as gBufferPass.frag:62-116 does.
 * texcoords may be NULL when no material is textured; tangents (gBufferPass.vert:40, xyz transformed by itModel, w the
 * handedness) may be NULL unless vgi_render_gbuffer is asked to shade a material with normal_texture > -1. */
typedef struct vgi_scene_desc {
    const float*           positions;   /* vec3 f32 x vertex_count */
    const float*           normals;     /* vec3 f32 x vertex_count */
    const float*           texcoords;   /* vec2 f32 x vertex_count, may be NULL */
    const uint32_t*        indices;     /* u32 x index_count */
    const vgi_primitive*   primitives;
    const vgi_node_matrix* nodes;
    const vgi_material*    materials;
    uint32_t vertex_count, index_count, primitive_count, node_count, material_count;
    const float*           tangents;    /* vec4 f32 x vertex_count, may be NULL */
} vgi_scene_desc;

/* G-buffer in the reference formats (ref: VFS/RenderPass/GBufferPass.cpp:177-194), linear row-major
 * device buffers of width*height texels. */
typedef struct vgi_gbuffer {
    const void*  diffuse_rgba8;     /* rgb = diffuse colour, a = perceptual roughness */
    const void*  normal_rgba16f;    /* n*0.5+0.5 */
    const void*  specular_rgba8;    /* rgb = F0 colour, a = metallic */
    const void*  emission_rgba16f;
    const float* depth_f32;         /* D32, LH zero-to-one */
    uint32_t width, height;
} vgi_gbuffer;

/* ref: VFS/RenderPass/Clipmap/VoxelConeTracingPass.h:46-59 (52-byte push constant block) */
typedef struct vgi_vct_params {
    float    volume_center[3];
    uint32_t rendering_mode;        /* 0..8, ref: VoxelConeTracingPass.h:17-28 */
    float    voxel_size;
    float    volume_dimension;
    float    trace_start_offset;
    float    indirect_diffuse_intensity;
    float    ambient_occlusion_factor;
    float    min_trace_step_factor;
    float    indirect_specular_intensity;
    float    occlusion_decay;
    int32_t  enable_32_cones;
} vgi_vct_params;

/* ref: VFS/RenderPass/SpecularFilterPass.h:31-37 (16-byte push constant block of specularFilter.frag) */
typedef struct vgi_filter_params {
    float   tonemap_gamma;          /* SpecularFilterPass.h:48 (2.2) */
    float   tonemap_exposure;       /* :49 (0.1) */
    int32_t tonemap_enable;         /* :50 (0) */
    int32_t filter_method;          /* :51 (1): 0 = bilateral 15x15, 1 = gaussian 32 x 8 taps, other = bilateral */
} vgi_filter_params;

/* Counters from the last build, for roofline accounting and overflow diagnosis. */
typedef struct vgi_stats {
    uint64_t triangles;             /* triangles in the scene */
    uint64_t clip_pairs;            /* (triangle,voxel) pairs emitted over all levels */
    uint64_t occupied_voxels;       /* occupied voxels over all levels */
    uint64_t svo_fragments;
    uint64_t svo_nodes;
    uint64_t kernel_launches;       /* kernels launched by this ctx since creation */
    uint64_t shaded_pairs;          /* the pairs on the injection's work list: clip_pairs minus those whose voxel is
                                     * overwritten by the down-sample of the finer level (centre half, off the blend band) */
} vgi_stats;

typedef struct vgi_ctx vgi_ctx;

/* ---- lifetime ------------------------------------------------------------------------------- */
int         vgi_version(void);
void        vgi_default_config(vgi_config* cfg);  /* reference defaults (EngineConfig.h:28-33) */
int         vgi_create(const vgi_config* cfg, vgi_ctx** out);
int         vgi_destroy(vgi_ctx* ctx);
const char* vgi_last_error(const vgi_ctx* ctx);   /* ctx may be NULL: last global error */
int         vgi_get_stats(vgi_ctx* ctx, vgi_stats* out); /* synchronises the ctx's last stream */
/* Per-kernel device timing for roofline accounting: when enabled, CUDA events are recorded on the
 * launching stream around every kernel. vgi_get_timings synchronises and returns, per kernel name,
 * the accumulated milliseconds and launch count since the last vgi_reset_timings.
 * names: array of `capacity` const char* (owned by the library). Returns the number of entries. */
int         vgi_set_timing(vgi_ctx* ctx, int enable);
int         vgi_get_timings(vgi_ctx* ctx, const char** names, double* ms, uint64_t* launches, uint32_t capacity);
int         vgi_reset_timings(vgi_ctx* ctx);

/* ---- inputs --------------------------------------------------------------------------------- */
/* replaces: GLTFScene vertex/index/matrix/material uploads consumed by msaaVoxelizer.vert:31-36 */
int vgi_set_scene(vgi_ctx* ctx, const vgi_scene_desc* scene);
/* replaces: the per-frame world transform of msaaVoxelizer.vert:31-36 / voxelizer.vert for ANIMATED nodes: new node matrices
 * (same count and order as in vgi_set_scene) re-transform the object-space vertices on the device — no re-upload of the
 * mesh. The result equals vgi_set_scene with those matrices bit for bit. nodes: HOST pointer, copied. */
int vgi_update_nodes(vgi_ctx* ctx, const vgi_node_matrix* nodes, uint32_t count, void* stream);
/* replaces: GLTFScene::uploadImage + the uTextures[] descriptor array (set 1, binding 2). May be called before or after
 * vgi_set_scene; a build whose materials reference a texture index >= count fails with VGI_E_STATE. count 0 clears. */
int vgi_set_textures(vgi_ctx* ctx, const vgi_texture* textures, uint32_t count);
/* replaces: light UBOs + shadow-map binding of RadianceInjectionPass (set 4) and
 * VoxelConeTracingPass (set 3). shadow_depth: w*h f32 (D32), device pointer borrowed until replaced
 * (is_host != 0: copied). */
int vgi_set_light(vgi_ctx* ctx, const vgi_dir_light* light, const vgi_dir_light_shadow* shadow,
                  const float* shadow_depth, uint32_t width, uint32_t height, int is_host);
/* replaces: Application::updateClipRegionBoundingBox (Application.cpp:116-128) +
 * VoxelizationPass::createVoxelClipmap / calculateChangeDelta (VoxelizationPass.cpp:335-357,438-448) */
int vgi_update_regions(vgi_ctx* ctx, const float camera_pos[3]);
int vgi_set_regions(vgi_ctx* ctx, const vgi_clip_region* regions, uint32_t count);
int vgi_get_regions(vgi_ctx* ctx, vgi_clip_region* out, uint32_t count);

/* ---- clipmap build -------------------------------------------------------------------------- */
/* replaces: VoxelizationPass::render (VoxelizationPass.cpp:74-212): clear, per-level conservative
 * voxelization (msaaVoxelizer.*), opacity down-sample, border wrap. */
int vgi_voxelize_opacity(vgi_ctx* ctx, void* stream);
/* replaces: RadianceInjectionPass::render (RadianceInjectionPass.cpp:64-159): cadence clear,
 * injection (msaaInjectRadiance.*), copy-alpha, radiance down-sample. Levels with
 * frame_index % 2^level != 0 keep their previous radiance. */
int vgi_inject_radiance(vgi_ctx* ctx, uint32_t frame_index, void* stream);
/* both of the above in one call (the fused fast path; identical results). */
int vgi_build_clipmap(vgi_ctx* ctx, uint32_t frame_index, void* stream);
/* replaces: the `_fullRevoxelization = false` branch of VoxelizationPass (VoxelizationPass.cpp:81-99,127-150 with
 * fillRevoxelizationRegions :450-494 — present in the reference, never taken): the same frame as vgi_build_clipmap, but only
 * what a moved clip region invalidates is rebuilt. Granularity is the clip level: a level whose region, children and radiance
 * are current keeps its records in the toroidal store; a level whose region moved — or whose cadence frame arrives while its
 * radiance is stale — is rebuilt together with every coarser level (their centre halves are down-samples of it). Scene, light
 * and textures must be unchanged since the previous build; their setters (and every full or sharded build) invalidate, after
 * which the next call rebuilds everything. Called on every frame of a sequence, the store equals the one vgi_build_clipmap
 * leaves on the same sequence bit for bit. first_level_rebuilt (may be NULL): the finest level that was rebuilt, level_count
 * when nothing had to be. vgi_get_stats then describes the rebuilt levels only. */
int vgi_build_clipmap_incremental(vgi_ctx* ctx, uint32_t frame_index, uint32_t* first_level_rebuilt, void* stream);

/* Export one atlas in the reference image layout: RGBA8, x-fastest,
 * W=(R+2)*6, H=(R+2)*L, D=R+2 (ref: Voxelizer.h:40-52), borders per mode_flags.
 * which: 0 = opacity ("VoxelOpacity"), 1 = radiance ("VoxelRadiance"). dst: device pointer. */
int    vgi_export_atlas(vgi_ctx* ctx, int which, void* dst, void* stream);
size_t vgi_atlas_bytes(const vgi_ctx* ctx);
/* The internal voxel store (DESIGN.md "data layout"): L*R^3 records of 32 bytes. */
int    vgi_get_voxel_store(vgi_ctx* ctx, void** dev_ptr, size_t* bytes);
/* Bind caller-owned device memory as the voxel store (multi-GPU all-gather, external memory). */
int    vgi_bind_voxel_store(vgi_ctx* ctx, void* dev_ptr, size_t bytes);
/* Restrict vgi_build_clipmap's record writes to z in [z0,z1) (slab sharding); default [0,R). */
int    vgi_set_slab(vgi_ctx* ctx, uint32_t z0, uint32_t z1);

/* ---- slab-sharded clipmap build (one ctx per GPU; the collectives stay with the caller) --------
 * GPU g owns the texel planes z in [z0,z1) of every level (vgi_set_slab); the z-slowest layout makes a slab
 * of the occupancy words one contiguous range per level. Sequence per frame, identical on every rank:
 *   vgi_slab_build_begin   voxelize + inject the own slab
 *   all-gather of the occupancy words in place (vgi_get_occupancy; e.g. ncclAllGather per level)
 *   vgi_slab_finalize      masks of the whole volume, own-slab records written and packed for exchange
 *   all-gather of the packed records (vgi_get_slab_pack), vgi_slab_unpack for every other rank's buffer
 *   vgi_slab_build_end     opacity + radiance mips (replicated) and the tracer's empty-space masks
 * The result equals vgi_build_clipmap on one GPU bit for bit. */
int vgi_slab_build_begin(vgi_ctx* ctx, uint32_t frame_index, void* stream);
/* occupancy words: level l at dev_ptr + l * bytes_per_level; planes [z0,z1) = bytes [z0*R*R/8, z1*R*R/8) */
int vgi_get_occupancy(vgi_ctx* ctx, void** dev_ptr, size_t* bytes_per_level);
int vgi_slab_finalize(vgi_ctx* ctx, uint32_t frame_index, void* stream);
/* ids: count x u32 (level << 27 | texel index), recs: count x 32-byte records; synchronises the stream */
int vgi_get_slab_pack(vgi_ctx* ctx, void** ids, void** recs, uint32_t* count);
int vgi_slab_unpack(vgi_ctx* ctx, const void* ids, const void* recs, uint32_t count, void* stream);
int vgi_slab_build_end(vgi_ctx* ctx, uint32_t frame_index, void* stream);

/* ---- peer build: the slab-sharded build with the exchange done by the kernels over NVLink peer memory ----
 * One ctx per GPU of one node (one process per GPU). Every GPU maps the voxel store, the occupancy words and a
 * row of arrival flags of every other GPU through CUDA IPC; a build then voxelizes + injects the own slab of texel
 * planes, stores its occupancy words and finalized records straight into ALL stores and meets the other GPUs at
 * three flag barriers inside the stream: no collective library call, no host round trip, no staging buffers.
 *   vgi_peer_export   handles: 3 x VGI_IPC_HANDLE_BYTES (store, occupancy, flags) of this ctx — exchange them
 *                     between the processes (e.g. torch.distributed.all_gather_object)
 *   vgi_peer_attach   all_handles: nranks x 3 x VGI_IPC_HANDLE_BYTES in rank order; nranks a power of two <= 8;
 *                     this GPU then owns the texel planes z with z mod nranks == rank (round-robin: balanced)
 *   vgi_peer_build_clipmap   = vgi_build_clipmap, bit for bit, on every GPU; all ranks must call it for the same
 *                     frame with the same regions / scene / light (a barrier waits ~3 s for a missing peer, then
 *                     gives up and the next vgi_get_stats fails with VGI_E_OVERFLOW, mask bit 0x20)
 * The stream order of each rank must keep its readers of the store (cone traces) before its next build. */
#define VGI_IPC_HANDLE_BYTES 64
int vgi_peer_export(vgi_ctx* ctx, void* handles);
int vgi_peer_attach(vgi_ctx* ctx, uint32_t rank, uint32_t nranks, const void* all_handles);
int vgi_peer_build_clipmap(vgi_ctx* ctx, uint32_t frame_index, void* stream);
int vgi_peer_detach(vgi_ctx* ctx);

/* ---- cone tracing --------------------------------------------------------------------------- */
/* replaces: VoxelConeTracingPass::onUpdate (VoxelConeTracingPass.cpp:75-106) + voxelConeTracing.frag.
 * out_diffuse / out_specular: width*height float4 (R32G32B32A32_SFLOAT, VoxelConeTracingPass.cpp:141-144).
 * Pixels with depth == 1 are left untouched (the shader discards). */
int vgi_cone_trace(vgi_ctx* ctx, const vgi_camera* cam, const vgi_gbuffer* gbuf,
                   const vgi_vct_params* params, void* out_diffuse, void* out_specular, void* stream);
/* Same, restricted to rows [y0,y1) of the image (screen-tile sharding across GPUs). */
int vgi_cone_trace_rows(vgi_ctx* ctx, const vgi_camera* cam, const vgi_gbuffer* gbuf,
                        const vgi_vct_params* params, void* out_diffuse, void* out_specular,
                        uint32_t y0, uint32_t y1, void* stream);
/* Same, restricted to the 8-row tile rows part, part + parts, part + 2*parts, ... of the image: the balanced way
 * to shard ONE view across `parts` GPUs (contiguous row blocks put all the sky on one GPU). */
int vgi_cone_trace_interleaved(vgi_ctx* ctx, const vgi_camera* cam, const vgi_gbuffer* gbuf,
                               const vgi_vct_params* params, void* out_diffuse, void* out_specular,
                               uint32_t part, uint32_t parts, void* stream);
/* Fill params with the reference defaults (VoxelConeTracingPass.h:75-82) and the volume fields
 * derived from the ctx's level-0 region (VoxelConeTracingPass.cpp:88-93). */
int vgi_default_vct_params(vgi_ctx* ctx, vgi_vct_params* out);
/* How the two marches of one vgi_cone_trace* call share the GPU (modes 6 and 8; the reference runs both inside one
 * fragment invocation, voxelConeTracing.frag:165-216). 0 (default): the diffuse march, then the specular march, one
 * after the other on the caller's stream. spec_blocks_per_sm > 0: the pixels that need a specular cone are listed by a
 * pre-pass, the specular march runs on the ctx's own stream with that many resident blocks per SM BESIDE the diffuse
 * march, and the caller's stream joins it before the call's work ends (images identical to the serial order).
 * Measured on B200 (1080p bench frame, tools/overlap_sweep.py): serial 4.17 ms; 1 / 2 / 3 / 4 / 6 blocks beside
 * 4.68 / 5.56 / 4.92 / 4.47 / 4.19 ms — two large kernels resident on one SM evict each other's instructions, so
 * the default stays serial; the switch is kept for hosts whose specular share is small. */
int vgi_set_trace_overlap(vgi_ctx* ctx, uint32_t spec_blocks_per_sm);

/* ---- the passes before the path: image inputs without Vulkan ------------------ */
/* replaces: the depth attachment of ShadowMapPass / ReflectiveShadowMapPass (shadowPass.vert:33: proj * view *
 * model, clear 1.0) for the ctx's scene. depth: width*height f32 DEVICE buffer (D32), usable as the
 * shadow_depth of vgi_set_light. Pixel-centre sampling, depth test LESS. */
int vgi_render_shadow_map(vgi_ctx* ctx, const vgi_dir_light_shadow* shadow, uint32_t width, uint32_t height,
                          float* depth, void* stream);
/* replaces: GBufferPass::onUpdate (gBufferPass.vert:38-45, gBufferPass.frag:39-116: factors, base-colour /
 * metallic-roughness / emissive / normal textures, alpha cutoff; formats GBufferPass.cpp:177-194). target: a vgi_gbuffer
 * whose five DEVICE buffers are WRITTEN (width*height texels each); uncovered pixels get depth 1 and zero attributes.
 * VGI_E_STATE: a material references a texture vgi_set_textures has not provided, or is normal-mapped while the scene
 * came without tangents. */
int vgi_render_gbuffer(vgi_ctx* ctx, const vgi_camera* cam, const vgi_gbuffer* target, void* stream);

/* ---- the pass after cone tracing ------------------------------------------ */
/* replaces: SpecularFilterPass::onUpdate (SpecularFilterPass.cpp:73-91) + specularFilter.frag:25-53,
 * filter.glsl:9-24,44-63, tonemapping.glsl:4-26: final = diffuse + filtered specular, optional Uncharted-2
 * tonemap. diffuse / specular: the two width*height float4 images of vgi_cone_trace, sampled LINEAR +
 * CLAMP_TO_EDGE as by "VCTSampler" (VoxelConeTracingPass.cpp:147); out: width*height float4
 * ("FinalOutputImageView", R32G32B32A32_SFLOAT, SpecularFilterPass.cpp:111). params NULL = defaults. */
void vgi_default_filter_params(vgi_filter_params* out);
int  vgi_specular_filter(vgi_ctx* ctx, const void* diffuse, const void* specular, uint32_t width,
                         uint32_t height, const vgi_filter_params* params, void* out, void* stream);

/* ---- whole frame with HOST buffers ---------------------------------------------------------- */
/* replaces: one iteration of Application::run's pass sequence for this path (Application.cpp:178,192,221:
 * "VoxelizationPass", "RadianceInjectionPass", "VoxelConeTracingPass") for a host that does not share
 * memory with CUDA: every pointer in host_gbuf, host_shadow_depth (w*h f32 as given to vgi_set_light,
 * NULL = keep the current one) and the two outputs (width*height float4 each) are HOST pointers
 * (pinned memory recommended). The call uploads the inputs, updates the regions from camera_pos,
 * builds the clipmap for frame_index, cone-traces, downloads both images and synchronises `stream`.
 * params may be NULL (reference defaults, rendering mode 8). */
int vgi_frame_host(vgi_ctx* ctx, uint32_t frame_index, const float camera_pos[3], const vgi_camera* cam,
                   const vgi_gbuffer* host_gbuf, const float* host_shadow_depth,
                   const vgi_vct_params* params, void* host_out_diffuse, void* host_out_specular,
                   void* stream);

/* Same frame for a headless / batched VIEW: nothing but this call's structs goes up. The G-buffer (width x height) is
 * rasterised on the device from the ctx's scene (vgi_render_gbuffer = GBufferPass::onUpdate) and, when `shadow` is not
 * NULL, so is the shadow map at the size given to vgi_set_light (vgi_render_shadow_map = the ShadowMapPass depth
 * attachment; NULL = keep the current one). Then regions, clipmap build, cone trace; both images are downloaded to the
 * HOST pointers and `stream` is synchronised. */
int vgi_frame_view_host(vgi_ctx* ctx, uint32_t frame_index, const float camera_pos[3], const vgi_camera* cam,
                        uint32_t width, uint32_t height, const vgi_dir_light_shadow* shadow,
                        const vgi_vct_params* params, void* host_out_diffuse, void* host_out_specular,
                        void* stream);
/* The same frame split in two so that a sequence of views is pipelined: _begin enqueues everything (both downloads ride a
 * copy stream) and returns; _end waits until the OLDEST frame begun has both images in its host buffers and reports its list
 * overflows. At most two frames may be in flight: begin(i), begin(i + 1), end() [= frame i], begin(i + 2), end() ... — the
 * download of frame i then overlaps the rasterisation and build of frame i + 1. Host buffers should be pinned.
 * vgi_frame_view_host = _begin + _end + a synchronise of `stream`. */
int vgi_frame_view_host_begin(vgi_ctx* ctx, uint32_t frame_index, const float camera_pos[3], const vgi_camera* cam,
                              uint32_t width, uint32_t height, const vgi_dir_light_shadow* shadow,
                              const vgi_vct_params* params, void* host_out_diffuse, void* host_out_specular,
                              void* stream);
int vgi_frame_view_host_end(vgi_ctx* ctx);

/* ---- sparse voxel octree -------------------------------------------------------------------- */
/* replaces: SparseVoxelizer::preVoxelize + cmdVoxelize (SparseVoxelizer.cpp:248-326) — one pass,
 * no host read-back. bb_min/bb_max: scene world bounding box (voxelizer.vert:42-47). */
int vgi_svo_voxelize(vgi_ctx* ctx, uint32_t level, const float bb_min[3], const float bb_max[3],
                     void* stream);
/* replaces: OctreeBuilder::cmdBuild (OctreeBuilder.cpp:205-345). */
int vgi_svo_build(vgi_ctx* ctx, void* stream);
/* replaces: SparseVoxelizer::getFramgnetList/getFragmentCount, OctreeBuilder::getOctreeBuffer /
 * getNumOctreeNodes (synchronise the stream). Fragments and nodes are uvec2. */
int vgi_svo_get_fragments(vgi_ctx* ctx, void** dev_ptr, uint32_t* count);
int vgi_svo_get_nodes(vgi_ctx* ctx, void** dev_ptr, uint32_t* count);
/* replaces: OctreeVoxelConeTracing::onUpdate (OctreeVoxelConeTracing.cpp:74-104). */
int vgi_svo_cone_trace(vgi_ctx* ctx, const vgi_camera* cam, const vgi_gbuffer* gbuf,
                       const vgi_vct_params* params, void* out_diffuse, void* out_specular,
                       void* stream);

/* ---- helper passes on caller-owned reference-layout atlases --------------------------------- */
/* Stand-alone equivalents of the reference's compute helpers, for hosts that keep their own
 * atlases. atlas: device RGBA8 image in the reference layout for (R, L). */
/* replaces: ClipmapCleaner::cmdClear*ClipRegion (ClipmapCleaner.cpp:94-134) */
int vgi_atlas_clear_region(vgi_ctx* ctx, void* atlas, const int32_t min_corner[3],
                           const uint32_t extent[3], uint32_t level, void* stream);
/* replaces: CopyAlpha::cmdImageCopyAlpha (CopyAlpha.cpp:93-146) */
int vgi_atlas_copy_alpha(vgi_ctx* ctx, void* dst_atlas, const void* src_atlas, uint32_t level,
                         void* stream);
/* replaces: DownSampler::cmdDownSampleOpacity / cmdDownSampleRadiance (DownSampler.cpp:115-168);
 * which: 0 opacity, 1 radiance. Uses the ctx's regions. */
int vgi_atlas_downsample(vgi_ctx* ctx, void* atlas, int which, uint32_t level, void* stream);
/* replaces: BorderWrapper::cmdWrappingOpacityBorder (BorderWrapper.cpp:123-151) */
int vgi_atlas_wrap_border(vgi_ctx* ctx, void* atlas, void* stream);

/* ---- Vulkan interop ------------------------------------------------------------------------- */
/* Import memory exported by the Vulkan host (VK_KHR_external_memory_fd, OPAQUE_FD) and map it as a
 * linear device buffer (cudaImportExternalMemory + cudaExternalMemoryGetMappedBuffer). */
int vgi_import_vk_memory(vgi_ctx* ctx, int fd, size_t size, void** dev_ptr, void** handle);
int vgi_release_vk_memory(vgi_ctx* ctx, void* handle);
/* Import a Vulkan semaphore (OPAQUE_FD) and wait / signal it on a stream. */
int vgi_import_vk_semaphore(vgi_ctx* ctx, int fd, void** handle);
int vgi_wait_vk_semaphore(vgi_ctx* ctx, void* handle, void* stream);
int vgi_signal_vk_semaphore(vgi_ctx* ctx, void* handle, void* stream);
int vgi_release_vk_semaphore(vgi_ctx* ctx, void* handle);
/* Block the calling thread until everything queued on `stream` has finished (cudaStreamSynchronize for hosts that do
 * not link the CUDA runtime themselves): the CPU-side alternative to the semaphore pair above — what the reference does
 * today with fence.waitForAllFences between its pre-pass batch and the frame (Application.cpp:199-201). */
int vgi_synchronize(vgi_ctx* ctx, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VGI_H */
