/*
 * vgi_oracle.h — CPU restatement ("oracle") of the reference's voxel-GI shaders.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing in the product path (vk_voxel_cone_tracing_b200/, libvgi.so)
 * may include, link or call this. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker / CPU baseline.
 *
 * PARITY: the reference ships no tests, golden vectors or fixtures for this path and its
 * GLSL cannot run as GLSL here (no Vulkan ICD / GLSL compiler). The oracle is pinned instead against the reference's
 * own SHADER TEXT compiled for the CPU (oracle/glsl_shim -> oracle/_ref/libvgi_refshaders.so): tests/test_ref_shaders.py
 * compares them live and through tests/golden/ref_shader_golden.npz - down-sampling, border wrap, clear, copy-alpha,
 * octree build, both cone tracers and the specular filter are bit-identical, the injection shading agrees to 2^-17.
 * STILL UNPINNED: triangle coverage (fixed-function rasteriser + MSAA in the reference, quirk Q3) - the canonical
 * conservative coverage below is a stated deviation checked by closed-form known answers (tests/test_oracle_kat.py).
 * Camera / light / voxelizer matrices are pinned by golden vectors generated from the reference's vendored glm
 * (oracle/ref_glm/).
 *
 * Arithmetic contract: IEEE-754 binary32, round-to-nearest-even, NO fused multiply-add
 * (build with -ffp-contract=off), expressions evaluated left to right as written in the GLSL.
 * The CUDA kernels follow the same contract wherever results are quantised (occupancy, RGBA8).
 *
 * All atlases here are in the REFERENCE image layout (ref: Voxelizer.h:40-52, Voxelizer.cpp:153-174):
 * RGBA8, x fastest, W=(R+2)*6, H=(R+2)*L, D=R+2; texel of voxel (x,y,z), face f, level l is
 * (1+x+f*(R+2), 1+y+l*(R+2), 1+z) (ref: msaaVoxelizer.frag:43-55).
 */
#ifndef VGI_ORACLE_H
#define VGI_ORACLE_H

#include "../include/vgi.h"

#ifdef __cplusplus
extern "C" {
#endif

/* world-space triangle soup produced by vgo_scene_triangles */
typedef struct vgo_tris {
    uint32_t count;
    float*   pos;   /* count*9: p0 p1 p2 (world) */
    float*   nrm;   /* count*9: world normals (itModel * n), NOT normalised */
    int32_t* mat;   /* count: material index */
    /* textured materials (both NULL for factor-only scenes): per-triangle texture coordinates and the material array the
     * coverage stage needs for the occlusion-texture alpha test (msaaVoxelizer.frag:64) */
    const float* uv;                    /* count*6: uv0 uv1 uv2 */
    const vgi_material* materials;
} vgo_tris;

/* the texture array of the scene (uTextures[]); pointers are borrowed. ref: GLTFScene::uploadImage, REPEAT + LINEAR,
 * level 0 only (Q23) */
void vgo_set_textures(const vgi_texture* textures, uint32_t count);
/* test aid: one bi-linear REPEAT read of level 0, as every texture() / textureLod() of the path evaluates */
void vgo_texture_fetch(uint32_t texture, float u, float v, float* rgba);

size_t vgo_atlas_bytes(const vgi_config* cfg);
/* sets (n > 0) and returns the number of OpenMP threads the oracle loops use */
int vgo_set_threads(int n);

/* ref: Application.cpp:116-128, VoxelizationPass.cpp:335-357, 438-448 */
void vgo_regions(const vgi_config* cfg, const float cam[3], vgi_clip_region* out);

/* ref: msaaVoxelizer.vert:31-36 (model / itModel transform), GLTFScene.cpp:457-490 (draw order) */
uint32_t vgo_scene_triangle_count(const vgi_scene_desc* s);
void     vgo_scene_triangles(const vgi_scene_desc* s, float* pos, float* nrm, int32_t* mat);

/* test aid: dominant axis of a world-space triangle, ref: msaaVoxelizer.geom:27-32 (ties -> z, then y) */
int vgo_dominant_axis(const float* p /*9*/);
/* ref: VoxelizationPass.cpp:104-126 (vkCmdClearColorImage) */
void vgo_clear_atlas(const vgi_config* cfg, uint8_t* atlas);
/* ref: msaaVoxelizer.geom:27-49, msaaVoxelizer.frag:43-73 with canonical conservative coverage (Q3).
 * returns number of (triangle,voxel) pairs */
uint64_t vgo_voxelize_level(const vgi_config* cfg, const vgi_clip_region* regions, uint32_t level,
                            const vgo_tris* tris, uint8_t* opacity);
/* ref: clipmapCleaning.comp:17-31 */
void vgo_clear_region(const vgi_config* cfg, uint8_t* atlas, const int32_t min_corner[3],
                      const uint32_t extent[3], uint32_t level);
/* ref: msaaInjectRadiance.frag:68-155,162-216 + shadow.glsl:8-36, canonical exact-mean accumulation (Q10) */
void vgo_inject_level(const vgi_config* cfg, const vgi_clip_region* regions, uint32_t level,
                      const vgo_tris* tris, const vgi_material* materials,
                      const vgi_dir_light* light, const vgi_dir_light_shadow* shadow,
                      const float* shadow_depth, uint32_t sw, uint32_t sh, uint8_t* radiance);
/* Q3 quantified (vgi_oracle_literal.inc): a MODEL of the reference's own raster coverage — (R+2)-pixel orthographic
 * viewports in which a pixel spans two voxels (Voxelizer.cpp:266-318), n-sample MSAA at the Vulkan standard locations
 * (Utils.cpp:8-20), sample shading 0.25 (VoxelizationPass.cpp:409-413), near plane 0.1, the literal texel addressing of
 * msaaVoxelizer.frag:43-73. samples: 8 | 4 | 1; shade_at: 0 = first covered sample of the invocation, 1 = pixel centre;
 * q2_fixed: per-axis clamp bounds. Return value: fragment-shader invocations that stored. Not used by any parity test:
 * the canonical mode stays the contract; these exist to measure how far the two are apart. */
uint64_t vgo_literal_voxelize_level(const vgi_config* cfg, const vgi_clip_region* regions, uint32_t level, const vgo_tris* tris,
                                    int samples, int shade_at, int q2_fixed, uint8_t* opacity);
uint64_t vgo_literal_inject_level(const vgi_config* cfg, const vgi_clip_region* regions, uint32_t level, const vgo_tris* tris,
                                  const vgi_material* materials, const vgi_dir_light* light, const vgi_dir_light_shadow* shadow,
                                  const float* shadow_depth, uint32_t sw, uint32_t sh, int samples, int shade_at, int q2_fixed,
                                  uint8_t* radiance);

/* test aid: the samples vgo_inject_level shades (position, un-normalised normal, material, unwrapped voxel) and their
 * shading results (faces, 16-bit fixed-point rgb); capacity == 0 counts only. Returns the number of samples. */
uint64_t vgo_inject_fragments(const vgi_config* cfg, const vgi_clip_region* regions, uint32_t level,
                              const vgo_tris* tris, const vgi_material* materials,
                              const vgi_dir_light* light, const vgi_dir_light_shadow* shadow,
                              const float* shadow_depth, uint32_t sw, uint32_t sh, uint64_t capacity,
                              float* pos, float* nrm, int32_t* mat, int32_t* voxel, int32_t* nfaces,
                              int32_t* faces, uint32_t* q, float* uv /* n*2, may be NULL: texture coordinate of each sample */);
/* ref: copyAlphaImage.comp:16-29 */
void vgo_copy_alpha(const vgi_config* cfg, uint32_t level, uint8_t* dst, const uint8_t* src);
/* ref: opacityDownSample.comp:28-138 (which=0), radianceDownSample.comp:28-139 with Q6 repaired (which=1) */
void vgo_downsample(const vgi_config* cfg, const vgi_clip_region* regions, uint32_t level,
                    uint8_t* atlas, int which);
/* ref: borderWrapping.comp:14-37; literal!=0 reproduces the 16-group dispatch (Q4) */
void vgo_wrap_border(const vgi_config* cfg, uint8_t* atlas, int literal);

/* ref: VoxelizationPass.cpp:74-212 (clear, voxelize all levels, opacity mips, border wrap) */
uint64_t vgo_voxelization_pass(const vgi_config* cfg, const vgi_clip_region* regions,
                               const vgo_tris* tris, uint8_t* opacity);
/* ref: RadianceInjectionPass.cpp:64-159 (cadence clear, inject, copy alpha, radiance mips) */
void vgo_injection_pass(const vgi_config* cfg, const vgi_clip_region* regions, const vgo_tris* tris,
                        const vgi_material* materials, const vgi_dir_light* light,
                        const vgi_dir_light_shadow* shadow, const float* shadow_depth,
                        uint32_t sw, uint32_t sh, uint32_t frame_index,
                        const uint8_t* opacity, uint8_t* radiance);

/* ref: voxelConeTracing.frag:143-414, brdf.glsl:30-78, shadow.glsl:8-36. HOST pointers in gbuf.
 * taps (may be NULL) receives the number of trilinear taps executed (SURVEY 8d A_cone).
 * rows [y0,y1). */
void vgo_cone_trace(const vgi_config* cfg, const vgi_camera* cam, const vgi_gbuffer* gbuf,
                    const vgi_vct_params* prm, const vgi_dir_light* light,
                    const vgi_dir_light_shadow* shadow, const float* shadow_depth,
                    uint32_t sw, uint32_t sh, const uint8_t* radiance,
                    float* out_diffuse, float* out_specular, uint32_t y0, uint32_t y1,
                    uint64_t* taps);

/* development aid: work statistics of the cone marches. enable != 0 turns counting on for later vgo_cone_trace calls; out
 * (10 values, may be NULL) receives and resets [diffuse | specular][steps, level samples, samples in a new (level, base cell)
 * compared with the previous step, all-zero samples, all-zero steps] */
void vgo_debug_cell_stats(int enable, uint64_t* out);

/* taps of the specular cones alone in the last vgo_cone_trace call */
uint64_t vgo_last_specular_taps(void);

/* ref: specularFilter.frag:25-53, filter.glsl:9-24 (gaussian), :27-63 (bilateral), tonemapping.glsl:4-26;
 * sampler LINEAR / CLAMP_TO_EDGE (VoxelConeTracingPass.cpp:147). HOST float4 images. */
void vgo_specular_filter(const float* diffuse, const float* specular, uint32_t w, uint32_t h,
                         const vgi_filter_params* prm, float* out);

/* ---- SVO (vgi_oracle_svo.inc, compiled as part of vgi_oracle.c) ---- */
/* ref: voxelizer.vert:37-48, voxelizer.geom:39-56, voxelizer.frag:48-103. Fragments are emitted
 * in (triangle, z, y, x) order. frags may be NULL to count only. returns count. */
uint32_t vgo_svo_fragments(uint32_t level, const float bb_min[3], const float bb_max[3],
                           const vgo_tris* tris, const vgi_material* materials,
                           const vgi_dir_light* light, const vgi_dir_light_shadow* shadow,
                           const float* shadow_depth, uint32_t sw, uint32_t sh,
                           uint32_t mode_flags, uint32_t* frags);
/* test aid: per triangle what voxelizer.vert / voxelizer.geom compute before rasterisation (normalised and biased vertex
 * positions, dominant axis from the normalised positions) */
void vgo_svo_vertex_stage(uint32_t level, const float bb_min[3], const float bb_max[3], const vgo_tris* tris,
                          float* ndc, float* biased, int32_t* axis);
/* test aid: the sample behind every covered (triangle, voxel) of vgo_svo_fragments (before the shading's discard), same
 * order: world position, biased [0,1] position, un-normalised normal, material, voxel. capacity == 0 counts only. */
uint32_t vgo_svo_fragment_samples(uint32_t level, const float bb_min[3], const float bb_max[3], const vgo_tris* tris,
                                  uint32_t capacity, float* world, float* biased, float* nrm, int32_t* mat, int32_t* voxel);
/* ref: OctreeBuilder.cpp:213-345 + octreeNode{Init,Flag,Alloc,ModifyArg,LeafWrite,MipmapWrite}.comp,
 * invocations executed sequentially in gl_GlobalInvocationID order. nodes: capacity*2 u32.
 * returns number of nodes allocated (allocBegin+allocNum after the last level). */
uint32_t vgo_svo_build(uint32_t level, const uint32_t* frags, uint32_t nfrag, uint32_t* nodes,
                       uint32_t capacity, uint32_t mode_flags);
/* canonical child ordering: rewrites the pool breadth-first so that equal trees compare equal
 * regardless of allocation order. returns node count. */
uint32_t vgo_svo_canonicalize(const uint32_t* nodes, uint32_t nnodes, uint32_t* out);
/* ref: voxelConeTracing_Octree.frag:148-409 */
void vgo_svo_cone_trace(const vgi_camera* cam, const vgi_gbuffer* gbuf, const vgi_vct_params* prm,
                        const vgi_dir_light* light, const vgi_dir_light_shadow* shadow,
                        const float* shadow_depth, uint32_t sw, uint32_t sh,
                        const uint32_t* nodes, const float bb_min[3], const float bb_max[3],
                        uint32_t clip_level_count,
                        float* out_diffuse, float* out_specular, uint32_t y0, uint32_t y1);
/* same with mode_flags: VGI_MODE_SVO_LITERAL keeps the shipped sampling (no >>1, Q13) */
void vgo_svo_cone_trace_mode(const vgi_camera* cam, const vgi_gbuffer* gbuf, const vgi_vct_params* prm,
                             const vgi_dir_light* light, const vgi_dir_light_shadow* shadow,
                             const float* shadow_depth, uint32_t sw, uint32_t sh,
                             const uint32_t* nodes, const float bb_min[3], const float bb_max[3],
                             uint32_t clip_level_count, uint32_t mode_flags,
                             float* out_diffuse, float* out_specular, uint32_t y0, uint32_t y1);

#ifdef __cplusplus
}
#endif
#endif
