// vgi_internal.h — shared between the C-ABI host code (vgi_api.cpp) and the CUDA translation units.
// Not part of the public interface (include/vgi.h is).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/vgi.h"

// ---- voxel store ---------------------------------------------------------------------------------
// One 32-byte record per voxel, levels concatenated, x fastest, no border (filtering addresses
// toroidally, which is what the reference's border texels emulate for the hardware sampler):
//   bytes  0..23  radiance RGBA8 of faces 0..5 (+X,-X,+Y,-Y,+Z,-Z)      == "VoxelRadiance" texels
//   bytes 24..29  opacity alpha of faces 0..5                             == "VoxelOpacity".a (= .g)
//   byte  30      raw occupancy flag (0/1)                                == "VoxelOpacity".r (= .b) / 255
//   byte  31      0
struct alignas(32) VoxelRecord {
    uint32_t radiance[6];
    uint8_t  opacity[6];
    uint8_t  raw;
    uint8_t  pad;
};
static_assert(sizeof(VoxelRecord) == 32, "record size");

// (triangle, level, texel) pair emitted by the voxelizer, consumed by the injection kernel.
// bits 0..8 x, 9..17 y, 18..26 z (texel = voxel mod R), 27..29 level, 32..63 triangle.
typedef unsigned long long vgi_pair_t;

struct LevelParams {
    int   min_corner[3];
    float voxel_size;
};

struct BuildParams {
    LevelParams lv[VGI_MAX_LEVELS];
    int      R, L, logR;
    int      band;
    uint32_t ntri;
    uint32_t max_pairs;
    uint32_t max_occ;       // capacity of the accumulator table (occupied voxels)
    uint32_t max_large;
    uint32_t level_mask;    // levels whose radiance is (re)injected this frame (cadence)
    int      level_first;   // incremental build: levels below this one are current and stay untouched (0 = rebuild all)
    uint32_t vox_levels;    // the levels k_voxelize visits, 4 bits each, finest first (full build: 0x..543210)
    int      vox_nlev;      // ... and how many
    int      z0, z1;        // slab of records this GPU writes: texel planes z in [z0, z1) ...
    int      z_mask, z_rem; // ... with (z & z_mask) == z_rem (peer build: planes dealt round-robin; 0, 0 = every plane)
    int      shadow_compare;
};

// material textures (vgi_set_textures): RGBA8 texels of all textures back to back + one (offset, width, height) row each;
// tri_uv = 3 texture coordinates per triangle (nullptr when the scene has none)
struct TexSet {
    const uint32_t* data;
    const uint4*    table;
    const float2*   tri_uv;
    uint32_t        count;
};

struct LightParams {
    float view[16], proj[16];
    float dir_to_light[3];  // normalize(-direction)
    float color[3];
    float intensity;
    float z_near, z_far;
    const float* depth;
    int   sw, sh;
};

// device-side counters (one small buffer, zeroed at the start of each build)
struct Counters {
    uint32_t pairs;
    uint32_t large;
    uint32_t occ_total;
    uint32_t overflow;      // bit 0 pairs, bit 1 large queue, bit 2 accumulators, bit 3 svo frags, bit 4 svo nodes
    uint32_t spec_pixels;
    uint32_t svo_frags;
    uint32_t svo_counter;
    uint32_t svo_alloc_begin;
    uint32_t svo_alloc_num;
    uint32_t visit[VGI_MAX_LEVELS]; // entries in each level's visit list (k_level_masks)
    uint32_t pairs_unlisted;        // canonical pairs that only set their occupancy bit (mip_interior): stats = pairs + this
    uint32_t pad[6];
};

// peer build (vgi_peer_*): device pointers of every GPU's buffers, mapped through CUDA IPC; [rank] = this GPU's own
#define VGI_MAX_PEERS 8
struct PeerSet {
    VoxelRecord* store[VGI_MAX_PEERS];
    uint32_t*    occ[VGI_MAX_PEERS];
    uint32_t*    flags[VGI_MAX_PEERS];  // one row of VGI_MAX_PEERS arrival epochs per GPU
    int n, rank;
};

// does this GPU own texel plane z (slab-sharded builds)?
static __host__ __device__ __forceinline__ bool owns_plane(const BuildParams& bp, int z)
{
    return z >= bp.z0 && z < bp.z1 && (z & bp.z_mask) == bp.z_rem;
}

struct TraceParams {
    vgi_vct_params p;
    float    view_proj_inv[16];
    float    eye[3];
    int      R, L, logR;
    const VoxelRecord* store;
    const uint8_t* brick_mask;    // 1 bit per 4^3 brick (dilated by one voxel on the high side), see k_brick_mask; read by the tracers only with VGI_TRACE_FP_ONLY = 0
    const uint8_t* footprint;     // per voxel: which of the 8 footprint records may be non-zero (0 in empty bricks: valid everywhere)
    float    vox_scale0;          // R / (voxel_size * volume_dimension): world units -> level-0 voxels
    float    level_scale[VGI_MAX_LEVELS]; // 2^-level
    float    min_level_dd[VGI_MAX_LEVELS]; // [k]: largest dist^2 with sqrtf(dd) / minRadius <= 2^k (+inf for k >= L-1)
    const void* diffuse; const void* normal; const void* specular; const void* emission;
    const float* depth;
    int      width, height, y0, y1;
    int      tile_stride, tile_phase; // 8-row tile rows phase, phase+stride, ... of [y0,y1) (multi-GPU interleave)
    float4*  out_diffuse;
    float4*  out_specular;
    LightParams light;
    int      shadow_compare;
    uint32_t* spec_list;          // compacted pixels needing a specular cone
    uint32_t* spec_count;
    uint32_t* spec_cursor;        // next unclaimed entry of spec_list (k_trace_specular)
    int      spec_presplit;       // 1: spec_list was filled by k_spec_classify and the specular march runs beside k_trace_main,
                                  //    which then neither appends nor writes out_specular of the listed pixels
    const uint2* svo_nodes;       // SVO tracer: node pool, grid of the fragment voxelizer
    float    svo_center[3], svo_extent, svo_max_level;
    int      svo_literal;         // VGI_MODE_SVO_LITERAL: sample positions are not halved (voxelConeTracing_Octree.frag:330-333 as shipped, Q13)
    float    cone_coeff_diffuse;  // 2*tan(aperture/2), evaluated on the host
    float    diffuse_aperture;
    // specular marches (k_trace_specular_warp): the step sequence depends on the roughness byte alone, so the host
    // tabulates it per byte value: spec_tab[rb * spec_stride + k] = (step_k, log2(diameter_k / voxel_size))
    const float2* spec_tab;       // nullptr -> per-lane marcher (k_trace_specular)
    const uint32_t* spec_cnt;     // [256] steps until MAX_TRACE_DISTANCE
    const float* spec_coeff;      // [256] 2*tan(max(roughness, 0.05)/2)
    uint32_t spec_stride;
};

// optional per-kernel CUDA-event timing (vgi_set_timing): events are recorded on the launching stream
// around every kernel and resolved at the next vgi_get_timings().
#define VGI_MAX_TIMED_KERNELS 32
struct KernelTimer {
    bool enabled = false;
    struct Pending { int slot; cudaEvent_t a, b; };
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> pool;
    const char* names[VGI_MAX_TIMED_KERNELS] = {};
    double ms[VGI_MAX_TIMED_KERNELS] = {};
    uint64_t count[VGI_MAX_TIMED_KERNELS] = {};
    int nslots = 0;
    int slot_of(const char* name);
    cudaEvent_t get_event();
    void begin(const char* name, cudaStream_t s);
    void end(cudaStream_t s);
    void resolve();
    void reset();
};

struct vgi_ctx {
    vgi_config cfg;
    KernelTimer timer;
    int device = 0;
    std::string err;
    vgi_clip_region regions[VGI_MAX_LEVELS];
    bool regions_set = false;

    // scene (world space)
    uint32_t ntri = 0;
    float4*  tri_pos = nullptr;   // 3 per triangle; p0.w carries the material index bits
    float4*  tri_nrm = nullptr;   // 3 per triangle
    vgi_material* materials = nullptr;
    uint32_t nmat = 0;
    float2*  tri_uv = nullptr;    // 3 per triangle, only when a material is textured
    float4*  tri_tan = nullptr;   // 3 per triangle (itModel * tangent.xyz, handedness), only for normal-mapped scenes with tangents
    float4*  obj_tan = nullptr;
    // object-space copies for vgi_update_nodes (animated nodes): xyz + node index in w / normal xyz
    float4*  obj_pos = nullptr;
    float4*  obj_nrm = nullptr;
    vgi_node_matrix* d_nodes = nullptr;
    uint32_t nnodes = 0;
    uint32_t* tex_data = nullptr;
    uint4*    tex_table = nullptr;
    uint32_t  ntex = 0;
    int32_t   scene_max_texture = -1;   // highest texture index the build stages read (base colour, emissive, occlusion)
    int32_t   scene_max_texture_all = -1;   // ... and the G-buffer producer (also metallic-roughness, normal)
    bool      scene_normal_mapped = false;  // a material has normal_texture > -1
    bool      scene_alpha_tested = false;   // a material has alpha_mode > 0 (gBufferPass.frag:96)
    TexSet texset() const { return TexSet{ tex_data, tex_table, tri_uv, ntex }; }
    std::vector<float> h_tri_pos; // host copies for debugging / multi-GPU culling
    float scene_bb_min[3], scene_bb_max[3];

    // light
    LightParams light;
    float* shadow_owned = nullptr;
    bool light_set = false;

    // build state
    VoxelRecord* store = nullptr;
    bool store_owned = false;
    size_t store_bytes = 0;
    uint32_t* occ = nullptr;        // L * R^3/32 words
    uint32_t* occ_prefix = nullptr; // exclusive prefix of popcounts per word
    uint32_t* block_sums = nullptr;
    uint32_t* acc = nullptr;        // max_occ * 24 u32: [voxel][face][r,g,b,count]
    vgi_pair_t* pairs = nullptr;
    uint2* large = nullptr;         // (triangle, level) work items for big triangles
    uint32_t* nz[2] = { nullptr, nullptr }; // non-zero-record masks, ping-pong between frames (L * R^3/32 words each)
    int nz_cur = 0;
    // vgi_build_clipmap_incremental: what the store currently holds, per level
    bool    inc_valid = false;
    int32_t inc_corner[VGI_MAX_LEVELS][3] = {};
    bool    inc_radiance_current[VGI_MAX_LEVELS] = {};
    uint8_t* brick_mask = nullptr;
    // slab-sharded build: exchange buffer of this GPU's finalized records
    uint32_t* slab_ids = nullptr;
    uint4*    slab_recs = nullptr;
    uint32_t* slab_count = nullptr;
    uint32_t  slab_cap = 0;
    int       slab_phase = 0;        // 0 idle, 1 begun, 2 finalized
    uint32_t* visit_list = nullptr;  // L segments of visit_cap voxel ids (records to rewrite this frame)
    uint32_t visit_cap = 0;
    uint8_t* footprint = nullptr;   // L * R^3 bytes, zero-initialised and kept valid for every voxel by k_brick_mask
    Counters* counters = nullptr;
    Counters* h_counters = nullptr; // pinned
    uint32_t max_pairs = 0, max_occ = 0, max_large = 0;
    int z0 = 0, z1 = 0, z_mask = 0, z_rem = 0;
    bool voxelized = false, built = false;
    uint32_t frame_of_voxelize = 0;
    cudaStream_t last_stream = 0;
    uint64_t launches = 0;

    // peer build
    PeerSet peers = {};
    bool peers_attached = false;
    uint32_t* sync_flags = nullptr;      // this GPU's flag row (exported)
    uint32_t peer_epoch = 0;
    void* peer_opened[VGI_MAX_PEERS][3] = {};

    // software rasteriser scratch (vgi_render_shadow_map / vgi_render_gbuffer)
    void* raster_proj = nullptr;               // projected triangles
    unsigned long long* raster_keys = nullptr; // per pixel: depth bits << 32 | triangle
    uint32_t* raster_large = nullptr;          // ntri queue entries + 1 counter
    size_t raster_px_cap = 0;
    uint32_t raster_tri_cap = 0;

    // cone trace scratch
    uint32_t* spec_list = nullptr;
    size_t spec_capacity = 0;
    // specular step tables (see TraceParams::spec_tab), rebuilt when the voxel size changes
    float2* spec_tab = nullptr;
    uint32_t* spec_cnt = nullptr;
    float* spec_coeff = nullptr;
    uint32_t spec_stride = 0;
    float spec_tab_voxel_size = 0.0f;
    bool spec_tab_unfit = false;      // table would be too large for this voxel size: per-lane marcher

    // cone trace: the specular march on its own stream beside the diffuse march (vgi_set_trace_overlap)
    uint32_t trace_spec_blocks = 0;   // resident specular blocks per SM while both run; 0 (default) = one kernel after the other
    cudaStream_t spec_stream = nullptr;
    cudaEvent_t ev_spec_fork = nullptr, ev_spec_done = nullptr;

    // build: the empty-space / visit-list masks depend on the occupancy alone, so they run on a side stream next to
    // the injection and the record pass (created on first use)
    cudaStream_t side_stream = nullptr;
    cudaEvent_t ev_side_fork = nullptr, ev_side_masks = nullptr, ev_side_done = nullptr;

    // vgi_frame_host: second stream + events so that PCIe copies overlap the kernels
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_inputs = nullptr, ev_main_done = nullptr, ev_fork = nullptr, ev_copy_done = nullptr;
    cudaEvent_t mark_main_done = nullptr; // when set, vgi_launch_trace records it after k_trace_main
    // vgi_frame_host staging (device copies of the host G-buffer and of both output images)
    uint8_t* stage = nullptr;
    size_t stage_bytes = 0;
    // vgi_frame_view_host[_begin/_end]: two device staging sets (G-buffer + both output images) so that the download of
    // frame i overlaps the work of frame i + 1; a frame's slot is free again after its _end
    uint8_t* view_stage[2] = { nullptr, nullptr };
    size_t view_stage_bytes = 0;
    cudaEvent_t ev_view_done[2] = { nullptr, nullptr };
    cudaEvent_t ev_view_traced = nullptr;
    Counters* h_view_counters = nullptr;   // pinned, one block per slot
    // ... and its rasterisation (shadow map + G-buffer depend on the scene only) runs on its own stream, beside the previous
    // frame's cone trace; each slot renders into its own shadow map
    cudaStream_t raster_stream = nullptr;
    cudaEvent_t ev_raster_done = nullptr;
    cudaEvent_t ev_scene = nullptr;         // recorded after the last device-side scene change (vgi_update_nodes)
    float* view_shadow[2] = { nullptr, nullptr };
    size_t view_shadow_bytes = 0;
    int view_pending = 0;                   // frames begun and not ended (0..2)
    int view_next = 0;                      // slot the next _begin takes
    size_t shadow_owned_bytes = 0;

    // svo
    uint32_t svo_level = 0;
    uint2* svo_frags = nullptr;
    uint32_t svo_frag_capacity = 0;
    uint2* svo_nodes = nullptr;
    uint32_t svo_node_capacity = 0;
    uint32_t svo_nfrag = 0, svo_nnodes = 0;
    bool svo_voxelized = false, svo_built = false, svo_counters_fresh = false;
    float svo_bb_min[3], svo_bb_max[3];
    uint32_t* svo_scratch = nullptr;
    size_t svo_scratch_words = 0;
};

// ---- launch wrappers implemented in the .cu files (return number of kernels launched) ------------
int vgi_launch_transform_scene(vgi_ctx* c, cudaStream_t s);
int vgi_launch_voxelize(vgi_ctx* c, const BuildParams& bp, cudaStream_t s);
int vgi_launch_inject_finalize(vgi_ctx* c, const BuildParams& bp, cudaStream_t s);
int vgi_launch_slab_begin(vgi_ctx* c, const BuildParams& bp, cudaStream_t s);
int vgi_launch_slab_finalize(vgi_ctx* c, const BuildParams& bp, cudaStream_t s);
int vgi_launch_slab_unpack(vgi_ctx* c, const uint32_t* ids, const uint4* recs, uint32_t count, cudaStream_t s);
int vgi_launch_slab_end(vgi_ctx* c, const BuildParams& bp, cudaStream_t s);
int vgi_launch_peer_build(vgi_ctx* c, const BuildParams& bp, const PeerSet& ps, uint32_t* epoch, cudaStream_t s);
int vgi_launch_export(vgi_ctx* c, int which, uint8_t* dst, int literal_border, cudaStream_t s);
int vgi_launch_trace(vgi_ctx* c, const TraceParams& tp, cudaStream_t s);
int vgi_launch_trace_svo(vgi_ctx* c, const TraceParams& tp, cudaStream_t s);
int vgi_launch_atlas_clear(uint8_t* atlas, int R, int L, const int32_t* mc, const uint32_t* ext, int level, cudaStream_t s);
int vgi_launch_atlas_copy_alpha(uint8_t* dst, const uint8_t* src, int R, int L, int level, cudaStream_t s);
int vgi_launch_atlas_downsample(uint8_t* atlas, int R, int L, int band, const int32_t* prev_min, int level, int which, cudaStream_t s);
int vgi_launch_atlas_wrap(uint8_t* atlas, int R, int L, int literal, cudaStream_t s);
int vgi_launch_render_shadow(vgi_ctx* c, const float* M, uint32_t w, uint32_t h, float* depth, cudaStream_t s);
int vgi_launch_render_gbuffer(vgi_ctx* c, const float* M, const vgi_gbuffer* target, cudaStream_t s);
size_t vgi_raster_proj_bytes(uint32_t ntri);
int vgi_launch_specular_filter(vgi_ctx* c, const void* diffuse, const void* specular, uint32_t width, uint32_t height,
                               const vgi_filter_params* prm, void* out, cudaStream_t s);

void build_params_from_ctx(const vgi_ctx* c, uint32_t frame_index, BuildParams* bp);
void vgi_set_thread_error(const std::string& msg);   // sets the string vgi_last_error(NULL) returns (vgi_api.cpp)
