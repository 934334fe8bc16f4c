// vgi_api.cpp — the C ABI of libvgi.so (include/vgi.h): context, inputs, orchestration.
// Host-side float arithmetic that feeds quantised results (regions, world-space vertices) follows the
// same IEEE binary32 / no-FMA / left-to-right contract as the kernels (built with -ffp-contract=off).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>

#include "vgi_internal.h"

static thread_local std::string g_err;

static int fail(vgi_ctx* c, int code, const std::string& msg)
{
    g_err = msg;
    if (c) c->err = msg;
    return code;
}
// for the other translation units: their messages must reach vgi_last_error(NULL) too
void vgi_set_thread_error(const std::string& msg) { g_err = msg; }
#define CK(ctx, call)                                                                                         \
    do {                                                                                                      \
        cudaError_t e_ = (call);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? VGI_E_NOMEM : VGI_E_CUDA,                     \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                                  \
    } while (0)

static int ilog2(uint32_t v) { int l = 0; while ((1u << l) < v) ++l; return l; }

extern "C" {

int vgi_version(void) { return VGI_VERSION; }

void vgi_default_config(vgi_config* cfg)
{
    memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = sizeof(vgi_config);
    cfg->resolution = 128;       // DEFAULT_VOXEL_RESOLUTION
    cfg->level_count = 6;        // DEFAULT_CLIP_REGION_COUNT
    cfg->downsample_band = 10;   // DEFAULT_DOWNSAMPLE_REGION_SIZE
    cfg->extent_level0 = 16.0f;  // DEFAULT_VOXEL_EXTENT_L0
    const uint32_t mc[6] = { 2, 2, 2, 2, 2, 1 }; // VoxelizationPass.h:57
    for (int i = 0; i < VGI_MAX_LEVELS; ++i) cfg->clip_min_change[i] = i < 6 ? mc[i] : 1;
    cfg->device = -1;
}

const char* vgi_last_error(const vgi_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

static void peer_close(vgi_ctx* c);

static void free_scene(vgi_ctx* c)
{
    cudaFree(c->tri_pos); cudaFree(c->tri_nrm); cudaFree(c->materials); cudaFree(c->tri_uv);
    cudaFree(c->obj_pos); cudaFree(c->obj_nrm); cudaFree(c->d_nodes);
    c->obj_pos = c->obj_nrm = nullptr; c->d_nodes = nullptr; c->nnodes = 0;
    cudaFree(c->tri_tan); cudaFree(c->obj_tan);
    c->tri_tan = c->obj_tan = nullptr;
    c->tri_uv = nullptr; c->scene_max_texture = c->scene_max_texture_all = -1; c->scene_normal_mapped = c->scene_alpha_tested = false;
    cudaFree(c->pairs); cudaFree(c->large); cudaFree(c->acc);
    cudaFree(c->raster_proj); cudaFree(c->raster_large);
    c->raster_proj = nullptr; c->raster_large = nullptr; c->raster_tri_cap = 0;
    c->tri_pos = c->tri_nrm = nullptr; c->materials = nullptr; c->pairs = nullptr; c->large = nullptr; c->acc = nullptr;
    c->ntri = 0;
}

int vgi_create(const vgi_config* cfg, vgi_ctx** out)
{
    if (!cfg || !out) return fail(nullptr, VGI_E_INVALID, "vgi_create: null argument");
    if (cfg->struct_size != sizeof(vgi_config)) return fail(nullptr, VGI_E_INVALID, "vgi_create: struct_size mismatch");
    const uint32_t R = cfg->resolution, L = cfg->level_count;
    if (R < 32 || R > 512 || (R & (R - 1))) return fail(nullptr, VGI_E_INVALID, "vgi_create: resolution must be a power of two in [32,512]");
    if (L < 1 || L > VGI_MAX_LEVELS) return fail(nullptr, VGI_E_INVALID, "vgi_create: level_count out of range");
    if (!(cfg->extent_level0 > 0.0f)) return fail(nullptr, VGI_E_INVALID, "vgi_create: extent_level0 must be positive");
    for (uint32_t i = 0; i < L; ++i)
    {
        if (cfg->clip_min_change[i] == 0) return fail(nullptr, VGI_E_INVALID, "vgi_create: clip_min_change must be >= 1");
        // levels with a parent are down-sampled in 2x2x2 blocks addressed by min_corner >> 1 (VoxelizationPass.h:57: 2,2,2,2,2,1)
        if (i + 1 < L && (cfg->clip_min_change[i] & 1u))
            return fail(nullptr, VGI_E_INVALID, "vgi_create: clip_min_change must be even on every level that has a parent");
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, VGI_E_CUDA, "vgi_create: no CUDA device (libvgi has no CPU fallback)");
    vgi_ctx* c = new (std::nothrow) vgi_ctx();
    if (!c) return fail(nullptr, VGI_E_NOMEM, "vgi_create: out of host memory");
    c->cfg = *cfg;
    // every allocation below may fail (the store alone is 3.2 GB at 6 x 256^3): on any error the partly built ctx is torn
    // down again (all pointers start out null) and the message survives in vgi_last_error(NULL)
    const int rc = [&]() -> int {
    if (cfg->device >= 0) { CK(c, cudaSetDevice(cfg->device)); }
    CK(c, cudaGetDevice(&c->device));
    memset(&c->light, 0, sizeof(c->light));
    const size_t nvox = (size_t)R * R * R;
    const size_t nwords = (nvox >> 5) * L;
    c->store_bytes = nvox * L * sizeof(VoxelRecord);
    CK(c, cudaMalloc(&c->store, c->store_bytes));
    c->store_owned = true;
    CK(c, cudaMemset(c->store, 0, c->store_bytes));
    CK(c, cudaMalloc(&c->occ, nwords * sizeof(uint32_t)));
    CK(c, cudaMalloc(&c->occ_prefix, nwords * sizeof(uint32_t)));
    for (int k = 0; k < 2; ++k) {
        CK(c, cudaMalloc(&c->nz[k], nwords * sizeof(uint32_t)));
        CK(c, cudaMemset(c->nz[k], 0, nwords * sizeof(uint32_t)));
    }
    {
        const size_t nb = (size_t)(R >> 2) * (R >> 2) * (R >> 5) * L;
        CK(c, cudaMalloc(&c->brick_mask, nb));
        CK(c, cudaMemset(c->brick_mask, 0, nb));
        CK(c, cudaMalloc(&c->footprint, nvox * L));
        CK(c, cudaMemset(c->footprint, 0, nvox * L)); // read speculatively next to the brick bit: keep it defined
    }
    CK(c, cudaMalloc(&c->block_sums, ((nwords + 4095) / 4096 + 1) * sizeof(uint32_t)));
    CK(c, cudaMalloc(&c->counters, sizeof(Counters)));
    CK(c, cudaMemset(c->counters, 0, sizeof(Counters)));
    CK(c, cudaMallocHost(&c->h_counters, sizeof(Counters)));
    memset(c->h_counters, 0, sizeof(Counters));
    return VGI_OK;
    }();
    if (rc != VGI_OK) {
        const std::string msg = g_err;
        cudaGetLastError();     // clear the sticky allocation error before tearing down
        vgi_destroy(c);
        g_err = msg;
        *out = nullptr;
        return rc;
    }
    c->z0 = 0;
    c->z1 = (int)R;
    // default regions: camera at the origin
    const float origin[3] = { 0.f, 0.f, 0.f };
    *out = c;
    vgi_update_regions(c, origin);
    return VGI_OK;
}

int vgi_destroy(vgi_ctx* c)
{
    if (!c) return VGI_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    free_scene(c);
    if (c->store_owned) cudaFree(c->store);
    cudaFree(c->occ); cudaFree(c->occ_prefix); cudaFree(c->block_sums); cudaFree(c->counters);
    cudaFree(c->brick_mask); cudaFree(c->slab_ids); cudaFree(c->slab_recs); cudaFree(c->slab_count); cudaFree(c->visit_list); cudaFree(c->footprint); cudaFree(c->nz[0]); cudaFree(c->nz[1]); cudaFree(c->spec_list); cudaFree(c->spec_tab); cudaFree(c->spec_cnt); cudaFree(c->spec_coeff); cudaFree(c->shadow_owned); cudaFree(c->stage);
    if (c->copy_stream) { cudaStreamDestroy(c->copy_stream); cudaEventDestroy(c->ev_inputs); cudaEventDestroy(c->ev_main_done); cudaEventDestroy(c->ev_fork); cudaEventDestroy(c->ev_copy_done); }
    cudaFree(c->svo_frags); cudaFree(c->svo_nodes); cudaFree(c->svo_scratch); cudaFree(c->raster_keys);
    peer_close(c);
    cudaFree(c->sync_flags);
    if (c->spec_stream) { cudaStreamDestroy(c->spec_stream); cudaEventDestroy(c->ev_spec_fork); cudaEventDestroy(c->ev_spec_done); }
    if (c->side_stream) { cudaStreamDestroy(c->side_stream); cudaEventDestroy(c->ev_side_fork); cudaEventDestroy(c->ev_side_masks); cudaEventDestroy(c->ev_side_done); }
    cudaFreeHost(c->h_counters);
    cudaFreeHost(c->h_view_counters);
    for (int k = 0; k < 2; ++k) cudaFree(c->view_shadow[k]);
    if (c->raster_stream) { cudaStreamDestroy(c->raster_stream); cudaEventDestroy(c->ev_raster_done); }
    if (c->ev_scene) cudaEventDestroy(c->ev_scene);
    for (int k = 0; k < 2; ++k) { cudaFree(c->view_stage[k]); if (c->ev_view_done[k]) cudaEventDestroy(c->ev_view_done[k]); }
    if (c->ev_view_traced) cudaEventDestroy(c->ev_view_traced);
    c->timer.resolve();
    for (cudaEvent_t e : c->timer.pool) cudaEventDestroy(e);
    delete c;
    return VGI_OK;
}

// The bounded device lists report overflow through Counters::overflow (h_counters is refreshed by every build). One message
// per bit, so that a peer-barrier timeout is not reported as a pair-list problem. Call only after the build's stream was synchronised.
static int report_overflow(vgi_ctx* c, const char* who, const Counters* counters = nullptr)
{
    const Counters* hc = counters ? counters : c->h_counters;
    const uint32_t m = hc->overflow;
    if (!m) return VGI_OK;
    c->inc_valid = false;       // an overflowed build left records unwritten: the next incremental build starts over
    char buf[320];
    int n = snprintf(buf, sizeof buf, "%s: device list overflow (mask 0x%x):", who, m);
    auto add = [&](const char* fmt, unsigned a, unsigned b) { if (n < (int)sizeof buf) n += snprintf(buf + n, sizeof buf - n, fmt, a, b); };
    if (m & 1u) add(" (triangle, voxel) pair list full, %u capacity %u - raise vgi_config.max_fragments;", hc->pairs, c->max_pairs);
    if (m & 2u) add(" large-triangle queue full (capacity %u)%.0u;", c->max_large, 0u);
    if (m & 4u) add(" accumulator / visit / exchange list full, %u occupied voxels, capacity %u - raise vgi_config.max_fragments;", hc->occ_total, c->max_occ);
    if (m & 8u) add(" octree fragment list full, %u capacity %u;", hc->svo_frags, c->svo_frag_capacity);
    if (m & 16u) add(" octree node pool full, %u capacity %u;", hc->svo_counter, c->svo_node_capacity);
    if (m & 32u) add(" peer build: a barrier timed out waiting for another GPU (epoch %u)%.0u;", c->peer_epoch, 0u);
    return fail(c, VGI_E_OVERFLOW, buf);
}

int vgi_get_stats(vgi_ctx* c, vgi_stats* out)
{
    if (!c || !out) return fail(c, VGI_E_INVALID, "vgi_get_stats: null argument");
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaStreamSynchronize(c->last_stream));
    CK(c, cudaMemcpy(c->h_counters, c->counters, sizeof(Counters), cudaMemcpyDeviceToHost));
    out->triangles = c->ntri;
    out->clip_pairs = (uint64_t)c->h_counters->pairs + c->h_counters->pairs_unlisted;
    out->shaded_pairs = c->h_counters->pairs;
    out->occupied_voxels = c->h_counters->occ_total;
    if (c->svo_counters_fresh) { // the device counters still hold the SVO pass (no clipmap build since)
        if (c->svo_voxelized) c->svo_nfrag = c->h_counters->svo_frags;
        if (c->svo_built) c->svo_nnodes = c->h_counters->svo_counter;
    }
    out->svo_fragments = c->svo_nfrag;
    out->svo_nodes = c->svo_nnodes;
    out->kernel_launches = c->launches;
    return report_overflow(c, "vgi_get_stats");
}

// ---- per-kernel timing ---------------------------------------------------------------------------
} // extern "C"

int KernelTimer::slot_of(const char* name)
{
    for (int i = 0; i < nslots; ++i)
        if (names[i] == name || !strcmp(names[i], name)) return i;
    if (nslots >= VGI_MAX_TIMED_KERNELS) return VGI_MAX_TIMED_KERNELS - 1;
    names[nslots] = name;
    return nslots++;
}
cudaEvent_t KernelTimer::get_event()
{
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
void KernelTimer::begin(const char* name, cudaStream_t s)
{
    if (!enabled) return;
    Pending p{ slot_of(name), get_event(), get_event() };
    cudaEventRecord(p.a, s);
    pending.push_back(p);
}
void KernelTimer::end(cudaStream_t s)
{
    if (!enabled || pending.empty()) return;
    cudaEventRecord(pending.back().b, s);
}
void KernelTimer::resolve()
{
    for (auto& p : pending) {
        cudaEventSynchronize(p.b);
        float t = 0.f;
        if (cudaEventElapsedTime(&t, p.a, p.b) == cudaSuccess) { ms[p.slot] += t; count[p.slot]++; }
        pool.push_back(p.a);
        pool.push_back(p.b);
    }
    pending.clear();
}
void KernelTimer::reset()
{
    resolve();
    for (int i = 0; i < VGI_MAX_TIMED_KERNELS; ++i) { ms[i] = 0; count[i] = 0; }
}

extern "C" {

int vgi_set_timing(vgi_ctx* c, int enable)
{
    if (!c) return fail(c, VGI_E_INVALID, "vgi_set_timing: null ctx");
    c->timer.resolve();
    c->timer.enabled = enable != 0;
    return VGI_OK;
}

int vgi_get_timings(vgi_ctx* c, const char** names, double* ms, uint64_t* launches, uint32_t capacity)
{
    if (!c) return fail(c, VGI_E_INVALID, "vgi_get_timings: null ctx");
    cudaStreamSynchronize(c->last_stream);
    c->timer.resolve();
    uint32_t n = 0;
    for (int i = 0; i < c->timer.nslots && n < capacity; ++i, ++n) {
        if (names) names[n] = c->timer.names[i];
        if (ms) ms[n] = c->timer.ms[i];
        if (launches) launches[n] = c->timer.count[i];
    }
    return (int)n;
}

int vgi_reset_timings(vgi_ctx* c)
{
    if (!c) return fail(c, VGI_E_INVALID, "vgi_reset_timings: null ctx");
    c->timer.reset();
    return VGI_OK;
}

// ---- inputs -----------------------------------------------------------------------------------------

static inline void xform_point(const float* m, const float* v, float* o)
{
    for (int r = 0; r < 3; ++r) o[r] = ((m[r] * v[0] + m[4 + r] * v[1]) + m[8 + r] * v[2]) + m[12 + r];
}
static inline void xform_dir(const float* m, const float* v, float* o)
{
    for (int r = 0; r < 3; ++r) o[r] = (m[r] * v[0] + m[4 + r] * v[1]) + m[8 + r] * v[2];
}

int vgi_set_scene(vgi_ctx* c, const vgi_scene_desc* s)
{
    if (!c || !s) return fail(c, VGI_E_INVALID, "vgi_set_scene: null argument");
    if (!s->positions || !s->normals || !s->indices || !s->primitives || !s->nodes || !s->materials)
        return fail(c, VGI_E_INVALID, "vgi_set_scene: missing buffer");
    int32_t maxTex = -1, maxTexAll = -1;
    bool normalMapped = false, alphaTested = false;
    for (uint32_t m = 0; m < s->material_count; ++m) {
        const vgi_material& mt = s->materials[m];
        const int32_t used[5] = { mt.base_color_texture, mt.emissive_texture, mt.occlusion_texture,
                                  mt.metallic_roughness_texture, mt.normal_texture };
        for (int k = 0; k < 3; ++k) maxTex = used[k] > maxTex ? used[k] : maxTex;
        for (int k = 0; k < 5; ++k) maxTexAll = used[k] > maxTexAll ? used[k] : maxTexAll;
        normalMapped = normalMapped || mt.normal_texture > -1;
        alphaTested = alphaTested || mt.alpha_mode > 0;
    }
    if (maxTexAll > -1 && !s->texcoords) return fail(c, VGI_E_INVALID, "vgi_set_scene: textured materials need texcoords");
    const bool withTan = normalMapped && s->tangents;
    uint64_t ntri = 0;
    for (uint32_t p = 0; p < s->primitive_count; ++p) {
        const vgi_primitive& pr = s->primitives[p];
        if (pr.node_index >= s->node_count || pr.material_index < 0 || (uint32_t)pr.material_index >= s->material_count ||
            (uint64_t)pr.first_index + pr.index_count > s->index_count)
            return fail(c, VGI_E_INVALID, "vgi_set_scene: primitive out of range");
        ntri += pr.index_count / 3;
    }
    if (ntri > 0x7fffffffull) return fail(c, VGI_E_INVALID, "vgi_set_scene: too many triangles");
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaStreamSynchronize(c->last_stream));
    free_scene(c);
    std::vector<float4> pos(ntri * 3), nrm(ntri * 3), opos(ntri * 3), onrm(ntri * 3);
    std::vector<float2> uv(maxTexAll > -1 ? ntri * 3 : 0);
    std::vector<float4> tan(withTan ? ntri * 3 : 0), otan(withTan ? ntri * 3 : 0);
    float bbmin[3] = { INFINITY, INFINITY, INFINITY }, bbmax[3] = { -INFINITY, -INFINITY, -INFINITY };
    size_t t = 0;
    // world transform: ref msaaVoxelizer.vert:31-36; draw order: GLTFScene.cpp:457-490
    for (uint32_t p = 0; p < s->primitive_count; ++p) {
        const vgi_primitive& pr = s->primitives[p];
        const vgi_node_matrix& nm = s->nodes[pr.node_index];
        for (uint32_t i = 0; i + 2 < pr.index_count; i += 3, ++t) {
            for (int k = 0; k < 3; ++k) {
                const uint64_t vi = (uint64_t)s->indices[pr.first_index + i + k] + pr.vertex_offset;
                if (vi >= s->vertex_count) return fail(c, VGI_E_INVALID, "vgi_set_scene: vertex index out of range");
                float w[3], n[3];
                xform_point(nm.model, s->positions + 3 * vi, w);
                xform_dir(nm.it_model, s->normals + 3 * vi, n);
                int32_t mat = pr.material_index;
                float matf;
                memcpy(&matf, &mat, 4);
                pos[t * 3 + k] = make_float4(w[0], w[1], w[2], matf);
                {   // object-space copy: the node index rides in w (vgi_update_nodes)
                    const uint32_t ni = pr.node_index;
                    float nif;
                    memcpy(&nif, &ni, 4);
                    opos[t * 3 + k] = make_float4(s->positions[3 * vi], s->positions[3 * vi + 1], s->positions[3 * vi + 2], nif);
                    onrm[t * 3 + k] = make_float4(s->normals[3 * vi], s->normals[3 * vi + 1], s->normals[3 * vi + 2], matf);
                }
                nrm[t * 3 + k] = make_float4(n[0], n[1], n[2], 0.f);
                if (maxTexAll > -1) uv[t * 3 + k] = make_float2(s->texcoords[2 * vi], s->texcoords[2 * vi + 1]);
                if (withTan) {      // ref: gBufferPass.vert:40
                    float tg[3];
                    xform_dir(nm.it_model, s->tangents + 4 * vi, tg);
                    tan[t * 3 + k] = make_float4(tg[0], tg[1], tg[2], s->tangents[4 * vi + 3]);
                    otan[t * 3 + k] = make_float4(s->tangents[4 * vi], s->tangents[4 * vi + 1], s->tangents[4 * vi + 2], s->tangents[4 * vi + 3]);
                }
                for (int a = 0; a < 3; ++a) {
                    bbmin[a] = w[a] < bbmin[a] ? w[a] : bbmin[a];
                    bbmax[a] = w[a] > bbmax[a] ? w[a] : bbmax[a];
                }
            }
        }
    }
    c->ntri = (uint32_t)ntri;
    memcpy(c->scene_bb_min, bbmin, sizeof bbmin);
    memcpy(c->scene_bb_max, bbmax, sizeof bbmax);
    c->nmat = s->material_count;
    c->scene_max_texture = maxTex;
    c->scene_max_texture_all = maxTexAll;
    c->scene_normal_mapped = normalMapped;
    c->scene_alpha_tested = alphaTested;
    if (ntri && maxTexAll > -1) {
        CK(c, cudaMalloc(&c->tri_uv, uv.size() * sizeof(float2)));
        CK(c, cudaMemcpy(c->tri_uv, uv.data(), uv.size() * sizeof(float2), cudaMemcpyHostToDevice));
    }
    if (ntri && withTan) {
        CK(c, cudaMalloc(&c->tri_tan, tan.size() * sizeof(float4)));
        CK(c, cudaMemcpy(c->tri_tan, tan.data(), tan.size() * sizeof(float4), cudaMemcpyHostToDevice));
        CK(c, cudaMalloc(&c->obj_tan, otan.size() * sizeof(float4)));
        CK(c, cudaMemcpy(c->obj_tan, otan.data(), otan.size() * sizeof(float4), cudaMemcpyHostToDevice));
    }
    if (ntri) {
        CK(c, cudaMalloc(&c->tri_pos, pos.size() * sizeof(float4)));
        CK(c, cudaMalloc(&c->tri_nrm, nrm.size() * sizeof(float4)));
        CK(c, cudaMemcpy(c->tri_pos, pos.data(), pos.size() * sizeof(float4), cudaMemcpyHostToDevice));
        CK(c, cudaMemcpy(c->tri_nrm, nrm.data(), nrm.size() * sizeof(float4), cudaMemcpyHostToDevice));
        CK(c, cudaMalloc(&c->obj_pos, opos.size() * sizeof(float4)));
        CK(c, cudaMalloc(&c->obj_nrm, onrm.size() * sizeof(float4)));
        CK(c, cudaMemcpy(c->obj_pos, opos.data(), opos.size() * sizeof(float4), cudaMemcpyHostToDevice));
        CK(c, cudaMemcpy(c->obj_nrm, onrm.data(), onrm.size() * sizeof(float4), cudaMemcpyHostToDevice));
    }
    c->nnodes = s->node_count;
    CK(c, cudaMalloc(&c->d_nodes, (s->node_count ? s->node_count : 1) * sizeof(vgi_node_matrix)));
    CK(c, cudaMemcpy(c->d_nodes, s->nodes, s->node_count * sizeof(vgi_node_matrix), cudaMemcpyHostToDevice));
    CK(c, cudaMalloc(&c->materials, (s->material_count ? s->material_count : 1) * sizeof(vgi_material)));
    CK(c, cudaMemcpy(c->materials, s->materials, s->material_count * sizeof(vgi_material), cudaMemcpyHostToDevice));

    const uint32_t R = c->cfg.resolution, L = c->cfg.level_count;
    uint64_t maxPairs = c->cfg.max_fragments;
    if (!maxPairs) {
        // pairs scale with the surface area in voxels, not only with the triangle count: a few wall-sized triangles
        // cover ~R^2 voxels each on every level they span, so the default also carries an R^2 L term
        maxPairs = ntri * 48 + 24ull * R * R * L;
        if (maxPairs < (1ull << 22)) maxPairs = 1ull << 22;
        if (maxPairs > (1ull << 27)) maxPairs = 1ull << 27;
    }
    c->max_pairs = (uint32_t)maxPairs;
    uint64_t maxOcc = maxPairs / 2;
    const uint64_t allVox = (uint64_t)R * R * R * L;
    if (maxOcc > allVox) maxOcc = allVox;
    if (maxOcc < 1024) maxOcc = 1024;
    c->max_occ = (uint32_t)maxOcc;
    c->max_large = (uint32_t)((ntri ? ntri : 1) * L);
    CK(c, cudaMalloc(&c->pairs, (size_t)c->max_pairs * sizeof(vgi_pair_t)));
    CK(c, cudaMalloc(&c->large, (size_t)c->max_large * sizeof(uint2)));
    CK(c, cudaMalloc(&c->acc, (size_t)c->max_occ * 24 * sizeof(uint32_t)));
    // visit list per level: this frame's and last frame's non-zero records (each at most occupied + mip-derived)
    cudaFree(c->visit_list);
    c->visit_list = nullptr;
    c->visit_cap = c->max_occ;
    CK(c, cudaMalloc(&c->visit_list, (size_t)c->visit_cap * L * sizeof(uint32_t)));
    c->voxelized = c->built = false;
    c->inc_valid = false;
    return VGI_OK;
}

static int check_launch(vgi_ctx* c, const char* what);

int vgi_update_nodes(vgi_ctx* c, const vgi_node_matrix* nodes, uint32_t count, void* stream)
{
    if (!c || !nodes) return fail(c, VGI_E_INVALID, "vgi_update_nodes: null argument");
    if (!c->materials) return fail(c, VGI_E_STATE, "vgi_update_nodes: call vgi_set_scene first");
    if (count != c->nnodes) return fail(c, VGI_E_INVALID, "vgi_update_nodes: node count differs from the scene's");
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    // the matrices are read by the transform kernel on `s`: a synchronous copy keeps the caller's buffer free on return
    CK(c, cudaStreamSynchronize(c->last_stream));
    CK(c, cudaMemcpy(c->d_nodes, nodes, (size_t)count * sizeof(vgi_node_matrix), cudaMemcpyHostToDevice));
    if (c->ntri) c->launches += vgi_launch_transform_scene(c, s);
    if (!c->ev_scene) cudaEventCreateWithFlags(&c->ev_scene, cudaEventDisableTiming);
    cudaEventRecord(c->ev_scene, s);    // rasterisation on another stream (vgi_frame_view_host_begin) waits for the new triangles
    c->voxelized = c->built = false;
    c->inc_valid = false;
    c->svo_voxelized = false;
    c->last_stream = s;
    return check_launch(c, "vgi_update_nodes");
}

int vgi_set_textures(vgi_ctx* c, const vgi_texture* tex, uint32_t count)
{
    if (!c || (count && !tex)) return fail(c, VGI_E_INVALID, "vgi_set_textures: null argument");
    size_t total = 0;
    for (uint32_t i = 0; i < count; ++i) {
        if (!tex[i].rgba8 || !tex[i].width || !tex[i].height) return fail(c, VGI_E_INVALID, "vgi_set_textures: empty texture");
        total += (size_t)tex[i].width * tex[i].height;
    }
    if (total > 0xffffffffull) return fail(c, VGI_E_INVALID, "vgi_set_textures: more than 2^32 texels");
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaStreamSynchronize(c->last_stream));
    cudaFree(c->tex_data); cudaFree(c->tex_table);
    c->tex_data = nullptr; c->tex_table = nullptr; c->ntex = 0;
    c->voxelized = c->built = false;
    c->inc_valid = false;
    if (!count) return VGI_OK;
    std::vector<uint4> table(count);
    CK(c, cudaMalloc(&c->tex_data, total * sizeof(uint32_t)));
    size_t off = 0;
    for (uint32_t i = 0; i < count; ++i) {
        const size_t n = (size_t)tex[i].width * tex[i].height;
        CK(c, cudaMemcpy(c->tex_data + off, tex[i].rgba8, n * sizeof(uint32_t), cudaMemcpyHostToDevice));
        table[i] = make_uint4((uint32_t)off, tex[i].width, tex[i].height, 0u);
        off += n;
    }
    CK(c, cudaMalloc(&c->tex_table, count * sizeof(uint4)));
    CK(c, cudaMemcpy(c->tex_table, table.data(), count * sizeof(uint4), cudaMemcpyHostToDevice));
    c->ntex = count;
    return VGI_OK;
}

int vgi_set_light(vgi_ctx* c, const vgi_dir_light* light, const vgi_dir_light_shadow* sh, const float* depth,
                  uint32_t w, uint32_t h, int is_host)
{
    if (!c || !light || !sh || !depth || !w || !h) return fail(c, VGI_E_INVALID, "vgi_set_light: null argument");
    LightParams& lp = c->light;
    memcpy(lp.view, sh->view, sizeof lp.view);
    memcpy(lp.proj, sh->proj, sizeof lp.proj);
    // normalize(-direction), ref msaaInjectRadiance.frag:139
    const float d[3] = { -light->direction[0], -light->direction[1], -light->direction[2] };
    const float len = sqrtf((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
    for (int k = 0; k < 3; ++k) { lp.dir_to_light[k] = d[k] / len; lp.color[k] = light->color[k]; }
    lp.intensity = light->intensity;
    lp.z_near = sh->z_near;
    lp.z_far = sh->z_far;
    lp.sw = (int)w;
    lp.sh = (int)h;
    if (is_host) {
        CK(c, cudaSetDevice(c->device));
        CK(c, cudaStreamSynchronize(c->last_stream));
        cudaFree(c->shadow_owned);
        c->shadow_owned = nullptr;
        CK(c, cudaMalloc(&c->shadow_owned, (size_t)w * h * sizeof(float)));
        c->shadow_owned_bytes = (size_t)w * h * sizeof(float);
        CK(c, cudaMemcpy(c->shadow_owned, depth, (size_t)w * h * sizeof(float), cudaMemcpyHostToDevice));
        lp.depth = c->shadow_owned;
    } else {
        lp.depth = depth;
    }
    c->light_set = true;
    c->inc_valid = false;
    return VGI_OK;
}

// ref: Application.cpp:116-128 + VoxelizationPass.cpp:335-357, 438-448
int vgi_update_regions(vgi_ctx* c, const float cam[3])
{
    if (!c || !cam) return fail(c, VGI_E_INVALID, "vgi_update_regions: null argument");
    const uint32_t R = c->cfg.resolution;
    for (uint32_t i = 0; i < c->cfg.level_count; ++i) {
        vgi_clip_region& r = c->regions[i];
        const float voxelSize = (c->cfg.extent_level0 * (float)(1u << i)) / (float)R;
        const float halfSize = (c->cfg.extent_level0 * 0.5f) * (float)(1u << i);
        const int32_t mc = (int32_t)c->cfg.clip_min_change[i];
        const float minChange = voxelSize * (float)mc;
        for (int k = 0; k < 3; ++k) {
            const int32_t minCorner = -(int32_t)(R >> 1);
            const float bbMin = cam[k] - halfSize;
            const float deltaW = bbMin - ((float)minCorner * voxelSize);
            const int32_t delta = (int32_t)truncf(deltaW / minChange) * mc;
            r.min_corner[k] = minCorner + delta;
            r.extent[k] = R;
        }
        r.voxel_size = voxelSize;
    }
    c->regions_set = true;
    return VGI_OK;
}

int vgi_set_regions(vgi_ctx* c, const vgi_clip_region* regions, uint32_t count)
{
    if (!c || !regions || count != c->cfg.level_count) return fail(c, VGI_E_INVALID, "vgi_set_regions: bad argument");
    for (uint32_t i = 0; i < count; ++i) {
        for (int k = 0; k < 3; ++k) {
            if (regions[i].extent[k] != c->cfg.resolution)
                return fail(c, VGI_E_INVALID, "vgi_set_regions: every extent must equal the resolution");
            // a level that has a parent is down-sampled in 2x2x2 blocks addressed by min_corner >> 1: an odd corner would
            // shift the child blocks by one toroidal plane (the reference snaps these levels in steps of 2, VoxelizationPass.h:57)
            if (i + 1 < count && (regions[i].min_corner[k] & 1))
                return fail(c, VGI_E_INVALID, "vgi_set_regions: min_corner must be even on every level that has a parent");
        }
        if (!(regions[i].voxel_size > 0.0f)) return fail(c, VGI_E_INVALID, "vgi_set_regions: voxel_size must be positive");
    }
    for (uint32_t i = 0; i < count; ++i) c->regions[i] = regions[i];
    c->regions_set = true;
    return VGI_OK;
}

int vgi_get_regions(vgi_ctx* c, vgi_clip_region* out, uint32_t count)
{
    if (!c || !out || count > c->cfg.level_count) return fail(c, VGI_E_INVALID, "vgi_get_regions: bad argument");
    memcpy(out, c->regions, count * sizeof(vgi_clip_region));
    return VGI_OK;
}

} // extern "C"

void build_params_from_ctx(const vgi_ctx* c, uint32_t frame_index, BuildParams* bp)
{
    memset(bp, 0, sizeof(*bp));
    bp->R = (int)c->cfg.resolution;
    bp->L = (int)c->cfg.level_count;
    bp->logR = ilog2(c->cfg.resolution);
    bp->band = (int)c->cfg.downsample_band;
    bp->ntri = c->ntri;
    bp->max_pairs = c->max_pairs;
    bp->max_occ = c->max_occ;
    bp->max_large = c->max_large;
    bp->z0 = c->z0;
    bp->z1 = c->z1;
    bp->z_mask = c->z_mask;
    bp->z_rem = c->z_rem;
    bp->shadow_compare = (c->cfg.mode_flags & VGI_MODE_SHADOW_COMPARE) ? 1 : 0;
    uint32_t mask = 0;
    for (int l = 0; l < bp->L; ++l) {
        for (int k = 0; k < 3; ++k) bp->lv[l].min_corner[k] = c->regions[l].min_corner[k];
        bp->lv[l].voxel_size = c->regions[l].voxel_size;
        if (frame_index % (1u << l) == 0) mask |= 1u << l; // ref: RadianceInjectionPass.cpp:36-38,75
    }
    bp->level_mask = mask;
    bp->vox_levels = 0x76543210u;
    bp->vox_nlev = bp->L;
}

extern "C" {

static int check_launch(vgi_ctx* c, const char* what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(c, VGI_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return VGI_OK;
}

int vgi_voxelize_opacity(vgi_ctx* c, void* stream)
{
    if (!c) return fail(c, VGI_E_INVALID, "vgi_voxelize_opacity: null ctx");
    if (!c->pairs) return fail(c, VGI_E_STATE, "vgi_voxelize_opacity: call vgi_set_scene first");
    if (c->scene_max_texture >= (int32_t)c->ntex) return fail(c, VGI_E_STATE, "vgi_voxelize_opacity: a material references a texture that vgi_set_textures has not provided");
    CK(c, cudaSetDevice(c->device));
    BuildParams bp;
    build_params_from_ctx(c, 0, &bp);
    cudaStream_t s = (cudaStream_t)stream;
    c->launches += vgi_launch_voxelize(c, bp, s);
    c->last_stream = s;
    if (c->svo_counters_fresh && c->svo_built) {    // keep the octree's node count before its counters are reused
        cudaStreamSynchronize(c->last_stream);
        c->svo_nnodes = c->h_counters->svo_counter;
    }
    c->svo_counters_fresh = false;  // the counters block and the pair buffer are shared with the octree path:
    c->svo_voxelized = false;       // its fragment list is gone (vgi_svo_build then fails with VGI_E_STATE instead of building an empty tree)
    c->voxelized = true;
    return check_launch(c, "vgi_voxelize_opacity");
}

int vgi_inject_radiance(vgi_ctx* c, uint32_t frame_index, void* stream)
{
    if (!c) return fail(c, VGI_E_INVALID, "vgi_inject_radiance: null ctx");
    if (!c->voxelized) return fail(c, VGI_E_STATE, "vgi_inject_radiance: call vgi_voxelize_opacity first");
    if (!c->light_set) return fail(c, VGI_E_STATE, "vgi_inject_radiance: call vgi_set_light first");
    CK(c, cudaSetDevice(c->device));
    BuildParams bp;
    build_params_from_ctx(c, frame_index, &bp);
    cudaStream_t s = (cudaStream_t)stream;
    c->launches += vgi_launch_inject_finalize(c, bp, s);
    c->last_stream = s;
    c->voxelized = false; // the pair list is consumed: the next frame re-voxelizes (Q8/Q9)
    c->built = true;
    c->inc_valid = false;   // a full rebuild: vgi_build_clipmap_incremental starts over
    return check_launch(c, "vgi_inject_radiance");
}

int vgi_build_clipmap(vgi_ctx* c, uint32_t frame_index, void* stream)
{
    int r = vgi_voxelize_opacity(c, stream);
    if (r != VGI_OK) return r;
    return vgi_inject_radiance(c, frame_index, stream);
}

// The reference's `_fullRevoxelization = false` branch (VoxelizationPass.cpp:81-99,127-150, fillRevoxelizationRegions
// :450-494), which it never takes: rebuild only what a moved clip region invalidates. Granularity here is the clip level:
// the toroidal store keeps every record of a level whose region, children and radiance are current, and a level is rebuilt
// as a whole — together with every coarser one, whose centre half is a down-sample of it — when its region moved, or when
// its cadence frame arrives while its radiance is stale. Scene, light and textures are assumed unchanged since the last
// build (their setters invalidate). The store after this call equals the store after vgi_build_clipmap(frame) called on
// every frame of the same sequence, bit for bit.
int vgi_build_clipmap_incremental(vgi_ctx* c, uint32_t frame_index, uint32_t* first_level_rebuilt, void* stream)
{
    if (!c) return fail(c, VGI_E_INVALID, "vgi_build_clipmap_incremental: null ctx");
    if (!c->pairs) return fail(c, VGI_E_STATE, "vgi_build_clipmap_incremental: call vgi_set_scene first");
    if (!c->light_set) return fail(c, VGI_E_STATE, "vgi_build_clipmap_incremental: call vgi_set_light first");
    if (!c->regions_set) return fail(c, VGI_E_STATE, "vgi_build_clipmap_incremental: call vgi_update_regions first");
    if (c->scene_max_texture >= (int32_t)c->ntex) return fail(c, VGI_E_STATE, "vgi_build_clipmap_incremental: a material references a texture that vgi_set_textures has not provided");
    if (c->z0 != 0 || c->z1 != (int)c->cfg.resolution || c->z_mask != 0)
        return fail(c, VGI_E_STATE, "vgi_build_clipmap_incremental: not with a slab-sharded context");
    CK(c, cudaSetDevice(c->device));
    BuildParams bp;
    build_params_from_ctx(c, frame_index, &bp);
    int first = bp.L;
    if (!c->inc_valid) first = 0;
    else
        for (int l = 0; l < bp.L && first == bp.L; ++l) {
            const bool moved = c->inc_corner[l][0] != bp.lv[l].min_corner[0] || c->inc_corner[l][1] != bp.lv[l].min_corner[1] ||
                               c->inc_corner[l][2] != bp.lv[l].min_corner[2];
            const bool onCadence = (bp.level_mask >> l) & 1u;
            if (moved || (onCadence && !c->inc_radiance_current[l])) first = l;
        }
    if (first_level_rebuilt) *first_level_rebuilt = (uint32_t)first;
    if (first == bp.L) return VGI_OK;       // every record is current
    bp.level_first = first;
    // occupancy bits persist per level: a level from `first` up is voxelized again only if its region moved (new occupancy)
    // or it is on its cadence frame (the injection needs its pair list); the others only re-run their mask / record passes
    bp.vox_levels = 0u;
    bp.vox_nlev = 0;
    for (int l = first; l < bp.L; ++l) {
        const bool moved = !c->inc_valid || c->inc_corner[l][0] != bp.lv[l].min_corner[0] || c->inc_corner[l][1] != bp.lv[l].min_corner[1] ||
                           c->inc_corner[l][2] != bp.lv[l].min_corner[2];
        if (moved || ((bp.level_mask >> l) & 1u)) bp.vox_levels |= (uint32_t)l << (4 * bp.vox_nlev++);
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (c->svo_counters_fresh && c->svo_built) {    // as in vgi_voxelize_opacity: the counters block is shared with the octree path
        cudaStreamSynchronize(c->last_stream);
        c->svo_nnodes = c->h_counters->svo_counter;
    }
    c->svo_counters_fresh = false;
    c->svo_voxelized = false;
    c->launches += vgi_launch_voxelize(c, bp, s);
    c->launches += vgi_launch_inject_finalize(c, bp, s);
    c->last_stream = s;
    for (int l = first; l < bp.L; ++l) {
        for (int k = 0; k < 3; ++k) c->inc_corner[l][k] = bp.lv[l].min_corner[k];
        // rebuilt off its cadence frame: opacity is current, the radiance texels are last frame's (as in the full build)
        c->inc_radiance_current[l] = (bp.level_mask >> l) & 1u;
    }
    c->inc_valid = true;
    c->voxelized = false;
    c->built = true;
    return check_launch(c, "vgi_build_clipmap_incremental");
}

size_t vgi_atlas_bytes(const vgi_ctx* c)
{
    if (!c) return 0;
    const size_t rb = c->cfg.resolution + 2;
    return rb * 6 * rb * c->cfg.level_count * rb * 4;
}

int vgi_export_atlas(vgi_ctx* c, int which, void* dst, void* stream)
{
    if (!c || !dst || (which != 0 && which != 1)) return fail(c, VGI_E_INVALID, "vgi_export_atlas: bad argument");
    if (!c->built) return fail(c, VGI_E_STATE, "vgi_export_atlas: nothing built yet");
    CK(c, cudaSetDevice(c->device));
    c->launches += vgi_launch_export(c, which, (uint8_t*)dst, (c->cfg.mode_flags & VGI_MODE_BORDER_LITERAL) ? 1 : 0, (cudaStream_t)stream);
    c->last_stream = (cudaStream_t)stream;
    return check_launch(c, "vgi_export_atlas");
}

int vgi_get_voxel_store(vgi_ctx* c, void** dev_ptr, size_t* bytes)
{
    if (!c) return fail(c, VGI_E_INVALID, "vgi_get_voxel_store: null ctx");
    if (dev_ptr) *dev_ptr = c->store;
    if (bytes) *bytes = c->store_bytes;
    return VGI_OK;
}

int vgi_bind_voxel_store(vgi_ctx* c, void* dev_ptr, size_t bytes)
{
    if (!c || !dev_ptr) return fail(c, VGI_E_INVALID, "vgi_bind_voxel_store: null argument");
    if (bytes < c->store_bytes) return fail(c, VGI_E_INVALID, "vgi_bind_voxel_store: buffer too small");
    if (((uintptr_t)dev_ptr) & 31u) return fail(c, VGI_E_INVALID, "vgi_bind_voxel_store: buffer must be 32-byte aligned");
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaStreamSynchronize(c->last_stream));
    if (c->store_owned) cudaFree(c->store);
    c->store = (VoxelRecord*)dev_ptr;
    c->store_owned = false;
    // the sparse build only rewrites records its masks know about: start from an all-zero store
    CK(c, cudaMemset(c->store, 0, c->store_bytes));
    const size_t nwords = (((size_t)c->cfg.resolution * c->cfg.resolution * c->cfg.resolution) >> 5) * c->cfg.level_count;
    for (int k = 0; k < 2; ++k) CK(c, cudaMemset(c->nz[k], 0, nwords * sizeof(uint32_t)));
    c->built = false;
    c->inc_valid = false;
    return VGI_OK;
}

int vgi_set_slab(vgi_ctx* c, uint32_t z0, uint32_t z1)
{
    if (!c || z0 >= z1 || z1 > c->cfg.resolution) return fail(c, VGI_E_INVALID, "vgi_set_slab: bad range");
    c->z0 = (int)z0;
    c->z1 = (int)z1;
    c->z_mask = c->z_rem = 0;
    return VGI_OK;
}

// ---- slab-sharded build ----------------------------------------------------------------------------

int vgi_slab_build_begin(vgi_ctx* c, uint32_t frame_index, void* stream)
{
    if (!c) return fail(c, VGI_E_INVALID, "vgi_slab_build_begin: null ctx");
    if (!c->pairs) return fail(c, VGI_E_STATE, "vgi_slab_build_begin: call vgi_set_scene first");
    if (c->scene_max_texture >= (int32_t)c->ntex) return fail(c, VGI_E_STATE, "vgi_slab_build_begin: a material references a texture that vgi_set_textures has not provided");
    if (!c->light_set) return fail(c, VGI_E_STATE, "vgi_slab_build_begin: call vgi_set_light first");
    CK(c, cudaSetDevice(c->device));
    if (!c->slab_ids) {
        c->slab_cap = c->max_occ;
        CK(c, cudaMalloc(&c->slab_ids, (size_t)c->slab_cap * sizeof(uint32_t)));
        CK(c, cudaMalloc(&c->slab_recs, (size_t)c->slab_cap * 2 * sizeof(uint4)));
        CK(c, cudaMalloc(&c->slab_count, sizeof(uint32_t)));
    }
    BuildParams bp;
    build_params_from_ctx(c, frame_index, &bp);
    cudaStream_t s = (cudaStream_t)stream;
    if (c->svo_counters_fresh && c->svo_built) {    // keep the octree's node count before its counters are reused
        cudaStreamSynchronize(c->last_stream);
        c->svo_nnodes = c->h_counters->svo_counter;
    }
    c->svo_counters_fresh = false;  // the counters block and the pair buffer are shared with the octree path:
    c->svo_voxelized = false;       // its fragment list is gone (vgi_svo_build then fails with VGI_E_STATE instead of building an empty tree)
    c->launches += vgi_launch_slab_begin(c, bp, s);
    c->last_stream = s;
    c->slab_phase = 1;
    return check_launch(c, "vgi_slab_build_begin");
}

int vgi_get_occupancy(vgi_ctx* c, void** dev_ptr, size_t* bytes_per_level)
{
    if (!c) return fail(c, VGI_E_INVALID, "vgi_get_occupancy: null ctx");
    if (dev_ptr) *dev_ptr = c->occ;
    if (bytes_per_level) *bytes_per_level = ((size_t)c->cfg.resolution * c->cfg.resolution * c->cfg.resolution) >> 3;
    return VGI_OK;
}

int vgi_slab_finalize(vgi_ctx* c, uint32_t frame_index, void* stream)
{
    if (!c) return fail(c, VGI_E_INVALID, "vgi_slab_finalize: null ctx");
    if (c->slab_phase != 1) return fail(c, VGI_E_STATE, "vgi_slab_finalize: call vgi_slab_build_begin first");
    CK(c, cudaSetDevice(c->device));
    BuildParams bp;
    build_params_from_ctx(c, frame_index, &bp);
    cudaStream_t s = (cudaStream_t)stream;
    c->launches += vgi_launch_slab_finalize(c, bp, s);
    c->last_stream = s;
    c->slab_phase = 2;
    return check_launch(c, "vgi_slab_finalize");
}

int vgi_get_slab_pack(vgi_ctx* c, void** ids, void** recs, uint32_t* count)
{
    if (!c) return fail(c, VGI_E_INVALID, "vgi_get_slab_pack: null ctx");
    if (c->slab_phase != 2) return fail(c, VGI_E_STATE, "vgi_get_slab_pack: call vgi_slab_finalize first");
    uint32_t n = 0;
    CK(c, cudaMemcpyAsync(&n, c->slab_count, sizeof n, cudaMemcpyDeviceToHost, c->last_stream));
    CK(c, cudaMemcpyAsync(c->h_counters, c->counters, sizeof(Counters), cudaMemcpyDeviceToHost, c->last_stream)); // this build's, not a stale copy
    CK(c, cudaStreamSynchronize(c->last_stream));
    if (n > c->slab_cap)
        return fail(c, VGI_E_OVERFLOW, "vgi_get_slab_pack: exchange buffer overflow — raise vgi_config.max_fragments");
    { const int ro = report_overflow(c, "vgi_get_slab_pack"); if (ro != VGI_OK) return ro; }
    if (ids) *ids = c->slab_ids;
    if (recs) *recs = c->slab_recs;
    if (count) *count = n;
    return VGI_OK;
}

int vgi_slab_unpack(vgi_ctx* c, const void* ids, const void* recs, uint32_t count, void* stream)
{
    if (!c || (count && (!ids || !recs))) return fail(c, VGI_E_INVALID, "vgi_slab_unpack: null argument");
    if (c->slab_phase != 2) return fail(c, VGI_E_STATE, "vgi_slab_unpack: call vgi_slab_finalize first");
    CK(c, cudaSetDevice(c->device));
    c->launches += vgi_launch_slab_unpack(c, (const uint32_t*)ids, (const uint4*)recs, count, (cudaStream_t)stream);
    c->last_stream = (cudaStream_t)stream;
    return check_launch(c, "vgi_slab_unpack");
}

int vgi_slab_build_end(vgi_ctx* c, uint32_t frame_index, void* stream)
{
    if (!c) return fail(c, VGI_E_INVALID, "vgi_slab_build_end: null ctx");
    if (c->slab_phase != 2) return fail(c, VGI_E_STATE, "vgi_slab_build_end: call vgi_slab_finalize first");
    CK(c, cudaSetDevice(c->device));
    BuildParams bp;
    build_params_from_ctx(c, frame_index, &bp);
    cudaStream_t s = (cudaStream_t)stream;
    c->launches += vgi_launch_slab_end(c, bp, s);
    c->last_stream = s;
    c->slab_phase = 0;
    c->voxelized = false;
    c->built = true;
    c->inc_valid = false;
    return check_launch(c, "vgi_slab_build_end");
}

// ---- cone tracing -----------------------------------------------------------------------------------

int vgi_default_vct_params(vgi_ctx* c, vgi_vct_params* p)
{
    if (!c || !p) return fail(c, VGI_E_INVALID, "vgi_default_vct_params: null argument");
    const vgi_clip_region& r0 = c->regions[0];
    // ref: VoxelConeTracingPass.cpp:88-93
    for (int k = 0; k < 3; ++k)
        p->volume_center[k] = ((float)r0.min_corner[k] * r0.voxel_size) + ((float)r0.extent[k] * r0.voxel_size) * 0.5f;
    p->rendering_mode = 8;                 // CombinedGI, VoxelConeTracingPass.h:75
    p->voxel_size = r0.voxel_size;
    p->volume_dimension = (float)c->cfg.resolution;
    p->trace_start_offset = 1.0f;          // VoxelConeTracingPass.h:76-82
    p->indirect_diffuse_intensity = 8.0f;
    p->ambient_occlusion_factor = 0.5f;
    p->min_trace_step_factor = 1.0f;
    p->indirect_specular_intensity = 3.0f;
    p->occlusion_decay = 2.0f;
    p->enable_32_cones = 0;
    return VGI_OK;
}

int vgi_set_trace_overlap(vgi_ctx* c, uint32_t spec_blocks_per_sm)
{
    if (!c) return fail(c, VGI_E_INVALID, "vgi_set_trace_overlap: null ctx");
    if (spec_blocks_per_sm > 16u) return fail(c, VGI_E_INVALID, "vgi_set_trace_overlap: at most 16 blocks per SM");
    c->trace_spec_blocks = spec_blocks_per_sm;
    return VGI_OK;
}

static int fill_trace_params(vgi_ctx* c, const vgi_camera* cam, const vgi_gbuffer* g, const vgi_vct_params* prm,
                             void* out_diffuse, void* out_specular, uint32_t y0, uint32_t y1, TraceParams& tp)
{
    memset(&tp, 0, sizeof tp);
    tp.p = *prm;
    memcpy(tp.view_proj_inv, cam->view_proj_inv, sizeof tp.view_proj_inv);
    memcpy(tp.eye, cam->eye_pos, sizeof tp.eye);
    tp.R = (int)c->cfg.resolution;
    tp.L = (int)c->cfg.level_count;
    tp.logR = ilog2(c->cfg.resolution);
    tp.store = c->store;
    tp.brick_mask = c->brick_mask;
    tp.footprint = c->footprint;
    // ref: voxelConeTracing.frag:315-317 — extent_l = voxelSize * volumeDimension * 2^level, texel = fract(pos / extent_l) * R
    tp.vox_scale0 = (float)c->cfg.resolution / (prm->voxel_size * prm->volume_dimension);
    for (int l = 0; l < VGI_MAX_LEVELS; ++l) tp.level_scale[l] = exp2f(-(float)l);
    {
        // ref: voxelConeTracing.frag:366-368 — minLevel = ceil(log2(dist / minRadius)), clamped to L-1 by the
        // tracer: ceil(log2 x) > k  <=>  x > 2^k. x(dd) = sqrtf(dd) / minRadius is monotone in the squared
        // distance dd, so bisect the largest dd (over binary32 bit patterns) with x(dd) <= 2^k.
        const float minRadius = prm->voxel_size * prm->volume_dimension * 0.5f;
        for (int k = 0; k < VGI_MAX_LEVELS; ++k) {
            if (k >= (int)c->cfg.level_count - 1) { tp.min_level_dd[k] = INFINITY; continue; }
            const float bound = exp2f((float)k);
            uint32_t lo = 0u, hi = 0x7f7fffffu; // invariant: x(lo) <= bound; answer = largest such pattern
            while (lo < hi) {
                const uint32_t mid = lo + (hi - lo + 1u) / 2u;
                float dd;
                memcpy(&dd, &mid, 4);
                const volatile float dist = sqrtf(dd);
                const volatile float x = dist / minRadius;
                if (x <= bound) lo = mid; else hi = mid - 1u;
            }
            memcpy(&tp.min_level_dd[k], &lo, 4);
        }
    }
    tp.diffuse = g->diffuse_rgba8; tp.normal = g->normal_rgba16f; tp.specular = g->specular_rgba8;
    tp.emission = g->emission_rgba16f; tp.depth = g->depth_f32;
    tp.width = (int)g->width; tp.height = (int)g->height; tp.y0 = (int)y0; tp.y1 = (int)y1;
    tp.tile_stride = 1; tp.tile_phase = 0;
    tp.out_diffuse = (float4*)out_diffuse;
    tp.out_specular = (float4*)out_specular;
    tp.light = c->light;
    tp.shadow_compare = (c->cfg.mode_flags & VGI_MODE_SHADOW_COMPARE) ? 1 : 0;
    tp.spec_list = c->spec_list + 2;
    tp.spec_count = c->spec_list;
    tp.spec_cursor = c->spec_list + 1;
    tp.spec_tab = c->spec_tab; tp.spec_cnt = c->spec_cnt; tp.spec_coeff = c->spec_coeff; tp.spec_stride = c->spec_stride;
    // ref: voxelConeTracing.frag:80,117,344 — coneCoefficient = 2 tan(aperture / 2)
    tp.diffuse_aperture = prm->enable_32_cones ? 0.628319f : 0.872665f;
    tp.cone_coeff_diffuse = 2.0f * tanf(tp.diffuse_aperture * 0.5f);
    return VGI_OK;
}

// Step tables of the specular cones (ref: voxelConeTracing.frag:205-216, 341-392). aperture = max(roughness, 0.05) with
// the roughness an 8-bit G-buffer value, stepFactor = uVoxelSize (Q12), so the sequence step_0 = 0,
// step_{k+1} = step_k + max(diameter_k, voxelSize) * voxelSize, diameter = step * 2 tan(aperture / 2) takes one of 256
// forms. Iterated here in binary32 exactly as the shader does (this file is compiled without FMA contraction); the warp-
// per-cone marcher reads 32 consecutive steps per batch from it.
static int ensure_spec_tables(vgi_ctx* c, const vgi_vct_params* prm, cudaStream_t s)
{
    const float vs = prm->voxel_size;
    if (c->spec_tab_voxel_size == vs && (c->spec_tab || c->spec_tab_unfit)) return VGI_OK;
    const float maxDistance = 30.0f;     // MAX_TRACE_DISTANCE
    std::vector<float> coeff(256);
    std::vector<uint32_t> cnt(256);
    std::vector<std::vector<float2>> seq(256);
    uint32_t stride = 0;
    bool unfit = !(vs > 0.0f);
    for (int rb = 0; rb < 256 && !unfit; ++rb) {
        const float rough = (float)rb / 255.0f;
        const float aperture = rough > 0.05f ? rough : 0.05f;          // MIN_SPECULAR_APERTURE
        const float cc = 2.0f * tanf(aperture * 0.5f);
        coeff[rb] = cc;
        if (rb > 0 && cc == coeff[rb - 1]) { seq[rb] = seq[rb - 1]; cnt[rb] = cnt[rb - 1]; continue; }
        float step = 0.0f;
        float diameter = step * cc > vs ? step * cc : vs;
        while (step < maxDistance) {
            seq[rb].push_back(make_float2(step, log2f(diameter / vs)));
            if (seq[rb].size() > (1u << 15)) { unfit = true; break; }
            step += (diameter > vs ? diameter : vs) * vs;
            diameter = step * cc;
        }
        cnt[rb] = (uint32_t)seq[rb].size();
        if (cnt[rb] > stride) stride = cnt[rb];
    }
    CK(c, cudaStreamSynchronize(c->last_stream));      // a trace in flight may still read the old tables
    CK(c, cudaStreamSynchronize(s));
    cudaFree(c->spec_tab); cudaFree(c->spec_cnt); cudaFree(c->spec_coeff);
    c->spec_tab = nullptr; c->spec_cnt = nullptr; c->spec_coeff = nullptr;
    c->spec_tab_voxel_size = vs;
    c->spec_tab_unfit = unfit;
    if (unfit) return VGI_OK;
    stride = (stride + 31u) & ~31u;
    std::vector<float2> flat((size_t)256 * stride, make_float2(3.0e38f, 0.0f));
    for (int rb = 0; rb < 256; ++rb) std::copy(seq[rb].begin(), seq[rb].end(), flat.begin() + (size_t)rb * stride);
    CK(c, cudaMalloc(&c->spec_tab, flat.size() * sizeof(float2)));
    CK(c, cudaMalloc(&c->spec_cnt, 256 * sizeof(uint32_t)));
    CK(c, cudaMalloc(&c->spec_coeff, 256 * sizeof(float)));
    CK(c, cudaMemcpy(c->spec_tab, flat.data(), flat.size() * sizeof(float2), cudaMemcpyHostToDevice));
    CK(c, cudaMemcpy(c->spec_cnt, cnt.data(), 256 * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CK(c, cudaMemcpy(c->spec_coeff, coeff.data(), 256 * sizeof(float), cudaMemcpyHostToDevice));
    c->spec_stride = stride;
    return VGI_OK;
}

static int check_trace_args(vgi_ctx* c, const vgi_camera* cam, const vgi_gbuffer* g, const vgi_vct_params* prm,
                            void* out_diffuse, void* out_specular, uint32_t y0, uint32_t y1, const char* who)
{
    if (!c || !cam || !g || !prm || !out_diffuse || !out_specular) return fail(c, VGI_E_INVALID, std::string(who) + ": null argument");
    if (!g->diffuse_rgba8 || !g->normal_rgba16f || !g->specular_rgba8 || !g->emission_rgba16f || !g->depth_f32 || !g->width || !g->height)
        return fail(c, VGI_E_INVALID, std::string(who) + ": incomplete G-buffer");
    if (y0 > y1 || y1 > g->height) return fail(c, VGI_E_INVALID, std::string(who) + ": bad row range");
    if (prm->rendering_mode > 8) return fail(c, VGI_E_INVALID, std::string(who) + ": rendering_mode out of range");
    if (!c->light_set) return fail(c, VGI_E_STATE, std::string(who) + ": call vgi_set_light first");
    return VGI_OK;
}

int vgi_cone_trace_rows(vgi_ctx* c, const vgi_camera* cam, const vgi_gbuffer* g, const vgi_vct_params* prm,
                        void* out_diffuse, void* out_specular, uint32_t y0, uint32_t y1, void* stream)
{
    int r = check_trace_args(c, cam, g, prm, out_diffuse, out_specular, y0, y1, "vgi_cone_trace");
    if (r != VGI_OK) return r;
    if (!c->built) return fail(c, VGI_E_STATE, "vgi_cone_trace: build the clipmap first");
    if ((uint32_t)prm->volume_dimension != c->cfg.resolution)
        return fail(c, VGI_E_INVALID, "vgi_cone_trace: volume_dimension must equal the clipmap resolution");
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t need = (size_t)g->width * g->height + 2;
    if (c->spec_capacity < need) {
        CK(c, cudaStreamSynchronize(c->last_stream));
        cudaFree(c->spec_list);
        c->spec_list = nullptr;
        CK(c, cudaMalloc(&c->spec_list, need * sizeof(uint32_t)));
        c->spec_capacity = need;
    }
    if (prm->rendering_mode == 6 || prm->rendering_mode == 8) { r = ensure_spec_tables(c, prm, s); if (r != VGI_OK) return r; }
    TraceParams tp;
    fill_trace_params(c, cam, g, prm, out_diffuse, out_specular, y0, y1, tp);
    c->launches += vgi_launch_trace(c, tp, s);
    c->last_stream = s;
    return check_launch(c, "vgi_cone_trace");
}

int vgi_cone_trace_interleaved(vgi_ctx* c, const vgi_camera* cam, const vgi_gbuffer* g, const vgi_vct_params* prm,
                               void* out_diffuse, void* out_specular, uint32_t part, uint32_t parts, void* stream)
{
    int r = check_trace_args(c, cam, g, prm, out_diffuse, out_specular, 0, g ? g->height : 0, "vgi_cone_trace_interleaved");
    if (r != VGI_OK) return r;
    if (!parts || part >= parts) return fail(c, VGI_E_INVALID, "vgi_cone_trace_interleaved: part must be < parts");
    if (!c->built) return fail(c, VGI_E_STATE, "vgi_cone_trace_interleaved: build the clipmap first");
    if ((uint32_t)prm->volume_dimension != c->cfg.resolution)
        return fail(c, VGI_E_INVALID, "vgi_cone_trace_interleaved: volume_dimension must equal the clipmap resolution");
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t need = (size_t)g->width * g->height + 2;
    if (c->spec_capacity < need) {
        CK(c, cudaStreamSynchronize(c->last_stream));
        cudaFree(c->spec_list);
        c->spec_list = nullptr;
        CK(c, cudaMalloc(&c->spec_list, need * sizeof(uint32_t)));
        c->spec_capacity = need;
    }
    if (prm->rendering_mode == 6 || prm->rendering_mode == 8) { r = ensure_spec_tables(c, prm, s); if (r != VGI_OK) return r; }
    TraceParams tp;
    fill_trace_params(c, cam, g, prm, out_diffuse, out_specular, 0, g->height, tp);
    tp.tile_stride = (int)parts;
    tp.tile_phase = (int)part;
    c->launches += vgi_launch_trace(c, tp, s);
    c->last_stream = s;
    return check_launch(c, "vgi_cone_trace_interleaved");
}

// replaces: OctreeVoxelConeTracing::onUpdate (OctreeVoxelConeTracing.cpp:74-104) + voxelConeTracing_Octree.frag
int vgi_svo_cone_trace(vgi_ctx* c, const vgi_camera* cam, const vgi_gbuffer* g, const vgi_vct_params* prm,
                       void* out_diffuse, void* out_specular, void* stream)
{
    int r = check_trace_args(c, cam, g, prm, out_diffuse, out_specular, 0, g ? g->height : 0, "vgi_svo_cone_trace");
    if (r != VGI_OK) return r;
    if (!c->svo_built) return fail(c, VGI_E_STATE, "vgi_svo_cone_trace: call vgi_svo_build first");
    if ((uint32_t)prm->volume_dimension != (1u << c->svo_level))
        return fail(c, VGI_E_INVALID, "vgi_svo_cone_trace: volume_dimension must equal 2^level of the octree");
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    TraceParams tp;
    fill_trace_params(c, cam, g, prm, out_diffuse, out_specular, 0, g->height, tp);
    tp.svo_nodes = c->svo_nodes;
    {   // ref: voxelizer.vert:42-47 — the grid of the fragment voxelizer
        const float ex = c->svo_bb_max[0] - c->svo_bb_min[0], ey = c->svo_bb_max[1] - c->svo_bb_min[1], ez = c->svo_bb_max[2] - c->svo_bb_min[2];
        const float m = ex > ey ? (ex > ez ? ex : ez) : (ey > ez ? ey : ez);
        tp.svo_extent = m * 0.5f;
        for (int k = 0; k < 3; ++k) tp.svo_center[k] = (c->svo_bb_min[k] + c->svo_bb_max[k]) * 0.5f;
        tp.svo_max_level = (float)((int)c->cfg.level_count - 1);
        tp.svo_literal = (c->cfg.mode_flags & VGI_MODE_SVO_LITERAL) ? 1 : 0;
    }
    c->launches += vgi_launch_trace_svo(c, tp, s);
    c->last_stream = s;
    return check_launch(c, "vgi_svo_cone_trace");
}

int vgi_cone_trace(vgi_ctx* c, const vgi_camera* cam, const vgi_gbuffer* g, const vgi_vct_params* prm,
                   void* out_diffuse, void* out_specular, void* stream)
{
    if (!g) return fail(c, VGI_E_INVALID, "vgi_cone_trace: null G-buffer");
    return vgi_cone_trace_rows(c, cam, g, prm, out_diffuse, out_specular, 0, g->height, stream);
}

// ---- helper passes on caller-owned reference-layout atlases --------------------------------------

int vgi_atlas_clear_region(vgi_ctx* c, void* atlas, const int32_t min_corner[3], const uint32_t extent[3], uint32_t level, void* stream)
{
    if (!c || !atlas || !min_corner || !extent) return fail(c, VGI_E_INVALID, "vgi_atlas_clear_region: null argument");
    if (level >= c->cfg.level_count) return fail(c, VGI_E_INVALID, "vgi_atlas_clear_region: level out of range");
    for (int k = 0; k < 3; ++k)
        if (extent[k] > c->cfg.resolution) return fail(c, VGI_E_INVALID, "vgi_atlas_clear_region: extent exceeds the resolution");
    CK(c, cudaSetDevice(c->device));
    c->launches += vgi_launch_atlas_clear((uint8_t*)atlas, (int)c->cfg.resolution, (int)c->cfg.level_count, min_corner, extent, (int)level, (cudaStream_t)stream);
    c->last_stream = (cudaStream_t)stream;
    return check_launch(c, "vgi_atlas_clear_region");
}

int vgi_atlas_copy_alpha(vgi_ctx* c, void* dst_atlas, const void* src_atlas, uint32_t level, void* stream)
{
    if (!c || !dst_atlas || !src_atlas) return fail(c, VGI_E_INVALID, "vgi_atlas_copy_alpha: null argument");
    if (level >= c->cfg.level_count) return fail(c, VGI_E_INVALID, "vgi_atlas_copy_alpha: level out of range");
    CK(c, cudaSetDevice(c->device));
    c->launches += vgi_launch_atlas_copy_alpha((uint8_t*)dst_atlas, (const uint8_t*)src_atlas, (int)c->cfg.resolution, (int)c->cfg.level_count, (int)level, (cudaStream_t)stream);
    c->last_stream = (cudaStream_t)stream;
    return check_launch(c, "vgi_atlas_copy_alpha");
}

int vgi_atlas_downsample(vgi_ctx* c, void* atlas, int which, uint32_t level, void* stream)
{
    if (!c || !atlas || (which != 0 && which != 1)) return fail(c, VGI_E_INVALID, "vgi_atlas_downsample: bad argument");
    if (level < 1 || level >= c->cfg.level_count) return fail(c, VGI_E_INVALID, "vgi_atlas_downsample: level must be in [1, L)");
    if (c->cfg.downsample_band > c->cfg.resolution / 4) return fail(c, VGI_E_INVALID, "vgi_atlas_downsample: band exceeds R/4");
    CK(c, cudaSetDevice(c->device));
    c->launches += vgi_launch_atlas_downsample((uint8_t*)atlas, (int)c->cfg.resolution, (int)c->cfg.level_count, (int)c->cfg.downsample_band,
                                               c->regions[level - 1].min_corner, (int)level, which, (cudaStream_t)stream);
    c->last_stream = (cudaStream_t)stream;
    return check_launch(c, "vgi_atlas_downsample");
}

int vgi_atlas_wrap_border(vgi_ctx* c, void* atlas, void* stream)
{
    if (!c || !atlas) return fail(c, VGI_E_INVALID, "vgi_atlas_wrap_border: null argument");
    CK(c, cudaSetDevice(c->device));
    c->launches += vgi_launch_atlas_wrap((uint8_t*)atlas, (int)c->cfg.resolution, (int)c->cfg.level_count,
                                         (c->cfg.mode_flags & VGI_MODE_BORDER_LITERAL) ? 1 : 0, (cudaStream_t)stream);
    c->last_stream = (cudaStream_t)stream;
    return check_launch(c, "vgi_atlas_wrap_border");
}

// ---- peer build: slab-sharded build with kernel-side exchange over NVLink peer memory -----------------

static void peer_close(vgi_ctx* c)
{
    for (int r = 0; r < VGI_MAX_PEERS; ++r)
        for (int k = 0; k < 3; ++k)
            if (c->peer_opened[r][k]) { cudaIpcCloseMemHandle(c->peer_opened[r][k]); c->peer_opened[r][k] = nullptr; }
    c->peers = PeerSet{};
    c->peers_attached = false;
}

int vgi_peer_export(vgi_ctx* c, void* handles)
{
    if (!c || !handles) return fail(c, VGI_E_INVALID, "vgi_peer_export: null argument");
    if (!c->store_owned || !c->store) return fail(c, VGI_E_STATE, "vgi_peer_export: the voxel store must be the ctx's own allocation");
    CK(c, cudaSetDevice(c->device));
    if (!c->sync_flags) {
        CK(c, cudaMalloc(&c->sync_flags, VGI_MAX_PEERS * sizeof(uint32_t)));
        CK(c, cudaMemset(c->sync_flags, 0, VGI_MAX_PEERS * sizeof(uint32_t)));
        c->peer_epoch = 0;
    }
    {   // the planes other GPUs own are only ever written by them: start from a defined state
        CK(c, cudaStreamSynchronize(c->last_stream));
        const size_t nwords = (((size_t)c->cfg.resolution * c->cfg.resolution * c->cfg.resolution) >> 5) * c->cfg.level_count;
        CK(c, cudaMemset(c->occ, 0, nwords * sizeof(uint32_t)));
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == VGI_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t* h = (cudaIpcMemHandle_t*)handles;
    CK(c, cudaIpcGetMemHandle(&h[0], c->store));
    CK(c, cudaIpcGetMemHandle(&h[1], c->occ));
    CK(c, cudaIpcGetMemHandle(&h[2], c->sync_flags));
    return VGI_OK;
}

int vgi_peer_attach(vgi_ctx* c, uint32_t rank, uint32_t nranks, const void* all_handles)
{
    if (!c || !all_handles) return fail(c, VGI_E_INVALID, "vgi_peer_attach: null argument");
    if (nranks < 1 || nranks > VGI_MAX_PEERS || rank >= nranks) return fail(c, VGI_E_INVALID, "vgi_peer_attach: bad rank / nranks");
    if (nranks & (nranks - 1)) return fail(c, VGI_E_INVALID, "vgi_peer_attach: nranks must be a power of two");
    if (!c->sync_flags) return fail(c, VGI_E_STATE, "vgi_peer_attach: call vgi_peer_export first");
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaStreamSynchronize(c->last_stream));
    peer_close(c);
    const cudaIpcMemHandle_t* h = (const cudaIpcMemHandle_t*)all_handles;
    for (uint32_t r = 0; r < nranks; ++r) {
        if (r == rank) {
            c->peers.store[r] = c->store; c->peers.occ[r] = c->occ; c->peers.flags[r] = c->sync_flags;
            continue;
        }
        void* p[3] = { nullptr, nullptr, nullptr };
        for (int k = 0; k < 3; ++k) {
            cudaError_t e = cudaIpcOpenMemHandle(&p[k], h[r * 3 + k], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                peer_close(c);
                cudaGetLastError();
                return fail(c, VGI_E_CUDA, std::string("vgi_peer_attach: cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
            }
            c->peer_opened[r][k] = p[k];
        }
        c->peers.store[r] = (VoxelRecord*)p[0]; c->peers.occ[r] = (uint32_t*)p[1]; c->peers.flags[r] = (uint32_t*)p[2];
    }
    c->peers.n = (int)nranks;
    c->peers.rank = (int)rank;
    c->peers_attached = true;
    // texel planes are dealt round-robin (z mod nranks == rank): neighbouring planes carry similar geometry, so every
    // GPU gets an equal share of the pairs whatever the scene's extent along z
    c->z0 = 0;
    c->z1 = (int)c->cfg.resolution;
    c->z_mask = (int)nranks - 1;
    c->z_rem = (int)rank;
    return VGI_OK;
}

int vgi_peer_detach(vgi_ctx* c)
{
    if (!c) return fail(c, VGI_E_INVALID, "vgi_peer_detach: null ctx");
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaStreamSynchronize(c->last_stream));
    peer_close(c);
    return vgi_set_slab(c, 0, c->cfg.resolution);
}

int vgi_peer_build_clipmap(vgi_ctx* c, uint32_t frame_index, void* stream)
{
    if (!c) return fail(c, VGI_E_INVALID, "vgi_peer_build_clipmap: null ctx");
    if (!c->peers_attached) return fail(c, VGI_E_STATE, "vgi_peer_build_clipmap: call vgi_peer_attach first");
    if (!c->pairs) return fail(c, VGI_E_STATE, "vgi_peer_build_clipmap: call vgi_set_scene first");
    if (c->scene_max_texture >= (int32_t)c->ntex) return fail(c, VGI_E_STATE, "vgi_peer_build_clipmap: a material references a texture that vgi_set_textures has not provided");
    if (!c->light_set) return fail(c, VGI_E_STATE, "vgi_peer_build_clipmap: call vgi_set_light first");
    CK(c, cudaSetDevice(c->device));
    BuildParams bp;
    build_params_from_ctx(c, frame_index, &bp);
    cudaStream_t s = (cudaStream_t)stream;
    if (c->svo_counters_fresh && c->svo_built) {    // keep the octree's node count before its counters are reused
        cudaStreamSynchronize(c->last_stream);
        c->svo_nnodes = c->h_counters->svo_counter;
    }
    c->svo_counters_fresh = false;  // the counters block and the pair buffer are shared with the octree path:
    c->svo_voxelized = false;       // its fragment list is gone (vgi_svo_build then fails with VGI_E_STATE instead of building an empty tree)
    c->launches += vgi_launch_peer_build(c, bp, c->peers, &c->peer_epoch, s);
    c->last_stream = s;
    c->voxelized = false;
    c->built = true;
    c->inc_valid = false;
    return check_launch(c, "vgi_peer_build_clipmap");
}

// ---- producers of the image inputs (the passes before the path) -----------------------------------

static int raster_scratch(vgi_ctx* c, size_t npx)
{
    if (c->raster_tri_cap < c->ntri + 1u) {
        CK(c, cudaStreamSynchronize(c->last_stream));
        cudaFree(c->raster_proj); cudaFree(c->raster_large);
        c->raster_proj = nullptr; c->raster_large = nullptr;
        CK(c, cudaMalloc(&c->raster_proj, vgi_raster_proj_bytes(c->ntri)));
        CK(c, cudaMalloc(&c->raster_large, ((size_t)c->ntri + 2) * sizeof(uint32_t)));
        c->raster_tri_cap = c->ntri + 1u;
    }
    if (c->raster_px_cap < npx) {
        CK(c, cudaStreamSynchronize(c->last_stream));
        cudaFree(c->raster_keys);
        c->raster_keys = nullptr; c->raster_px_cap = 0;
        CK(c, cudaMalloc(&c->raster_keys, npx * sizeof(unsigned long long)));
        c->raster_px_cap = npx;
    }
    return VGI_OK;
}

int vgi_render_shadow_map(vgi_ctx* c, const vgi_dir_light_shadow* sh, uint32_t width, uint32_t height, float* depth, void* stream)
{
    if (!c || !sh || !depth || !width || !height) return fail(c, VGI_E_INVALID, "vgi_render_shadow_map: bad argument");
    if (!c->tri_pos && c->ntri) return fail(c, VGI_E_STATE, "vgi_render_shadow_map: call vgi_set_scene first");
    if (!c->materials) return fail(c, VGI_E_STATE, "vgi_render_shadow_map: call vgi_set_scene first");
    CK(c, cudaSetDevice(c->device));
    int r = raster_scratch(c, (size_t)width * height);
    if (r != VGI_OK) return r;
    float M[16]; // proj * view, column-major (shadowPass.vert:33)
    for (int col = 0; col < 4; ++col)
        for (int row = 0; row < 4; ++row) {
            double a = 0;
            for (int k = 0; k < 4; ++k) a += (double)sh->proj[k * 4 + row] * sh->view[col * 4 + k];
            M[col * 4 + row] = (float)a;
        }
    c->launches += vgi_launch_render_shadow(c, M, width, height, depth, (cudaStream_t)stream);
    c->last_stream = (cudaStream_t)stream;
    return check_launch(c, "vgi_render_shadow_map");
}

int vgi_render_gbuffer(vgi_ctx* c, const vgi_camera* cam, const vgi_gbuffer* target, void* stream)
{
    if (!c || !cam || !target) return fail(c, VGI_E_INVALID, "vgi_render_gbuffer: null argument");
    if (!target->diffuse_rgba8 || !target->normal_rgba16f || !target->specular_rgba8 || !target->emission_rgba16f ||
        !target->depth_f32 || !target->width || !target->height)
        return fail(c, VGI_E_INVALID, "vgi_render_gbuffer: incomplete target");
    if (!c->materials) return fail(c, VGI_E_STATE, "vgi_render_gbuffer: call vgi_set_scene first");
    if (c->scene_max_texture_all >= (int32_t)c->ntex)
        return fail(c, VGI_E_STATE, "vgi_render_gbuffer: a material references a texture that vgi_set_textures has not provided");
    if (c->scene_normal_mapped && !c->tri_tan)
        return fail(c, VGI_E_STATE, "vgi_render_gbuffer: normal-mapped materials need vgi_scene_desc.tangents");
    CK(c, cudaSetDevice(c->device));
    int r = raster_scratch(c, (size_t)target->width * target->height);
    if (r != VGI_OK) return r;
    c->launches += vgi_launch_render_gbuffer(c, cam->view_proj, target, (cudaStream_t)stream);
    c->last_stream = (cudaStream_t)stream;
    return check_launch(c, "vgi_render_gbuffer");
}

// ---- specular filter + tonemap (the pass after cone tracing) --------------------------------------

void vgi_default_filter_params(vgi_filter_params* p)
{
    if (!p) return;
    p->tonemap_gamma = 2.2f;     // ref: SpecularFilterPass.h:48-51
    p->tonemap_exposure = 0.1f;
    p->tonemap_enable = 0;
    p->filter_method = 1;
}

int vgi_specular_filter(vgi_ctx* c, const void* diffuse, const void* specular, uint32_t width, uint32_t height,
                        const vgi_filter_params* params, void* out, void* stream)
{
    if (!c || !diffuse || !specular || !out) return fail(c, VGI_E_INVALID, "vgi_specular_filter: null argument");
    if (!width || !height) return fail(c, VGI_E_INVALID, "vgi_specular_filter: empty image");
    if (out == diffuse || out == specular) return fail(c, VGI_E_INVALID, "vgi_specular_filter: out must not alias an input");
    vgi_filter_params prm;
    if (params) prm = *params;
    else vgi_default_filter_params(&prm);
    if (prm.tonemap_enable == 1 && !(prm.tonemap_gamma > 0.0f))
        return fail(c, VGI_E_INVALID, "vgi_specular_filter: tonemap_gamma must be positive");
    CK(c, cudaSetDevice(c->device));
    c->launches += vgi_launch_specular_filter(c, diffuse, specular, width, height, &prm, out, (cudaStream_t)stream);
    c->last_stream = (cudaStream_t)stream;
    return check_launch(c, "vgi_specular_filter");
}

// ---- whole frame with host buffers ----------------------------------------------------------------

int vgi_frame_host(vgi_ctx* c, uint32_t frame_index, const float camera_pos[3], const vgi_camera* cam,
                   const vgi_gbuffer* hg, const float* host_shadow_depth, const vgi_vct_params* params,
                   void* host_out_diffuse, void* host_out_specular, void* stream)
{
    if (!c || !camera_pos || !cam || !hg || !host_out_diffuse || !host_out_specular)
        return fail(c, VGI_E_INVALID, "vgi_frame_host: null argument");
    if (!hg->diffuse_rgba8 || !hg->normal_rgba16f || !hg->specular_rgba8 || !hg->emission_rgba16f || !hg->depth_f32 ||
        !hg->width || !hg->height)
        return fail(c, VGI_E_INVALID, "vgi_frame_host: incomplete G-buffer");
    if (!c->light_set) return fail(c, VGI_E_STATE, "vgi_frame_host: call vgi_set_light first");
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t npx = (size_t)hg->width * hg->height;
    // staging layout: diffuse 4, specular 4, depth 4, normal 8, emission 8, out_diffuse 16, out_specular 16 B/pixel
    const size_t need = npx * 60;
    if (c->stage_bytes < need) {
        CK(c, cudaStreamSynchronize(c->last_stream));
        cudaFree(c->stage);
        c->stage = nullptr;
        c->stage_bytes = 0;
        CK(c, cudaMalloc(&c->stage, need));
        c->stage_bytes = need;
    }
    uint8_t* d_out_d = c->stage;
    uint8_t* d_out_s = d_out_d + npx * 16;
    uint8_t* d_nrm = d_out_s + npx * 16;
    uint8_t* d_emi = d_nrm + npx * 8;
    uint8_t* d_dif = d_emi + npx * 8;
    uint8_t* d_spc = d_dif + npx * 4;
    uint8_t* d_dep = d_spc + npx * 4;
    if (!c->copy_stream) {
        CK(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        CK(c, cudaEventCreateWithFlags(&c->ev_inputs, cudaEventDisableTiming));
        CK(c, cudaEventCreateWithFlags(&c->ev_main_done, cudaEventDisableTiming));
        CK(c, cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        CK(c, cudaEventCreateWithFlags(&c->ev_copy_done, cudaEventDisableTiming));
    }
    cudaStream_t cs = c->copy_stream;
    // fork: the copy stream starts where the caller's stream is now
    CK(c, cudaEventRecord(c->ev_fork, s));
    CK(c, cudaStreamWaitEvent(cs, c->ev_fork, 0));
    // main stream: shadow map (needed by the injection) then the build; copy stream: the G-buffer, which only the
    // tracer needs, travels while the clipmap is built.
    // (Measured and rejected: tracing the image in four row bands so that band k uploads while band k-1 marches and
    // finished rows go home early — the specular kernel's tail, its longest march, is then paid once per band:
    // 8.4 ms instead of 7.5 ms per 1080p frame.)
    if (host_shadow_depth) {
        const size_t sb = (size_t)c->light.sw * c->light.sh * sizeof(float);
        if (!c->shadow_owned || c->shadow_owned_bytes < sb) {
            CK(c, cudaStreamSynchronize(c->last_stream));
            cudaFree(c->shadow_owned);
            c->shadow_owned = nullptr;
            CK(c, cudaMalloc(&c->shadow_owned, sb));
            c->shadow_owned_bytes = sb;
        }
        CK(c, cudaMemcpyAsync(c->shadow_owned, host_shadow_depth, sb, cudaMemcpyHostToDevice, s));
        c->light.depth = c->shadow_owned;
    }
    CK(c, cudaMemcpyAsync(d_nrm, hg->normal_rgba16f, npx * 8, cudaMemcpyHostToDevice, cs));
    CK(c, cudaMemcpyAsync(d_emi, hg->emission_rgba16f, npx * 8, cudaMemcpyHostToDevice, cs));
    CK(c, cudaMemcpyAsync(d_dif, hg->diffuse_rgba8, npx * 4, cudaMemcpyHostToDevice, cs));
    CK(c, cudaMemcpyAsync(d_spc, hg->specular_rgba8, npx * 4, cudaMemcpyHostToDevice, cs));
    CK(c, cudaMemcpyAsync(d_dep, hg->depth_f32, npx * 4, cudaMemcpyHostToDevice, cs));
    // discarded pixels (depth == 1) are left untouched by the tracer: give them a defined value
    CK(c, cudaMemsetAsync(d_out_d, 0, npx * 32, cs));
    CK(c, cudaEventRecord(c->ev_inputs, cs));
    int r = vgi_update_regions(c, camera_pos);
    if (r != VGI_OK) return r;
    r = vgi_build_clipmap(c, frame_index, stream);
    if (r != VGI_OK) return r;
    vgi_vct_params prm;
    if (params) prm = *params;
    else vgi_default_vct_params(c, &prm);
    vgi_gbuffer dg;
    dg.diffuse_rgba8 = d_dif; dg.normal_rgba16f = d_nrm; dg.specular_rgba8 = d_spc; dg.emission_rgba16f = d_emi;
    dg.depth_f32 = (const float*)d_dep; dg.width = hg->width; dg.height = hg->height;
    CK(c, cudaStreamWaitEvent(s, c->ev_inputs, 0));
    c->mark_main_done = c->ev_main_done;
    r = vgi_cone_trace(c, cam, &dg, &prm, d_out_d, d_out_s, stream);
    c->mark_main_done = nullptr;
    if (r != VGI_OK) return r;
    // the diffuse image is final after the main kernel: it goes home while the specular cones march
    CK(c, cudaStreamWaitEvent(cs, c->ev_main_done, 0));
    CK(c, cudaMemcpyAsync(host_out_diffuse, d_out_d, npx * 16, cudaMemcpyDeviceToHost, cs));
    CK(c, cudaEventRecord(c->ev_copy_done, cs));
    CK(c, cudaMemcpyAsync(host_out_specular, d_out_s, npx * 16, cudaMemcpyDeviceToHost, s));
    // join: the caller's stream is complete only when the side copies are
    CK(c, cudaStreamWaitEvent(s, c->ev_copy_done, 0));
    CK(c, cudaStreamSynchronize(s));
    // the build's counters came home with it: a dropped pair or record means the image is wrong, say so
    return report_overflow(c, "vgi_frame_host");
}

// Whole frame of ONE VIEW with device-rendered inputs (batched / headless views: only the camera goes up, the two
// images come home). The shadow map (when `shadow` is given) and the G-buffer are rasterised from the ctx's scene by
// vgi_render_shadow_map / vgi_render_gbuffer — bit-identical to the host rasteriser the parity tests pin — so nothing
// but this call's small structs crosses PCIe on the way in.
int vgi_frame_view_host_begin(vgi_ctx* c, uint32_t frame_index, const float camera_pos[3], const vgi_camera* cam,
                              uint32_t width, uint32_t height, const vgi_dir_light_shadow* shadow,
                              const vgi_vct_params* params, void* host_out_diffuse, void* host_out_specular, void* stream)
{
    if (!c || !camera_pos || !cam || !host_out_diffuse || !host_out_specular || !width || !height)
        return fail(c, VGI_E_INVALID, "vgi_frame_view_host: bad argument");
    if (!c->light_set) return fail(c, VGI_E_STATE, "vgi_frame_view_host: call vgi_set_light first");
    if (c->view_pending >= 2) return fail(c, VGI_E_STATE, "vgi_frame_view_host_begin: two frames are in flight - call vgi_frame_view_host_end first");
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t npx = (size_t)width * height;
    const size_t need = npx * 60;
    if (!c->copy_stream) {
        CK(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        CK(c, cudaEventCreateWithFlags(&c->ev_inputs, cudaEventDisableTiming));
        CK(c, cudaEventCreateWithFlags(&c->ev_main_done, cudaEventDisableTiming));
        CK(c, cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        CK(c, cudaEventCreateWithFlags(&c->ev_copy_done, cudaEventDisableTiming));
    }
    cudaStream_t cs = c->copy_stream;
    if (!c->ev_view_traced) {
        CK(c, cudaEventCreateWithFlags(&c->ev_view_traced, cudaEventDisableTiming));
        for (int k = 0; k < 2; ++k) CK(c, cudaEventCreateWithFlags(&c->ev_view_done[k], cudaEventDisableTiming));
        CK(c, cudaMallocHost(&c->h_view_counters, 2 * sizeof(Counters)));
        memset(c->h_view_counters, 0, 2 * sizeof(Counters));
    }
    if (c->view_stage_bytes < need) {
        if (c->view_pending) return fail(c, VGI_E_STATE, "vgi_frame_view_host_begin: the image size changed while a frame is in flight");
        CK(c, cudaStreamSynchronize(c->last_stream));
        CK(c, cudaStreamSynchronize(cs));
        for (int k = 0; k < 2; ++k) { cudaFree(c->view_stage[k]); c->view_stage[k] = nullptr; }
        c->view_stage_bytes = 0;
        for (int k = 0; k < 2; ++k) CK(c, cudaMalloc(&c->view_stage[k], need));
        c->view_stage_bytes = need;
    }
    const int slot = c->view_next;
    uint8_t* d_out_d = c->view_stage[slot];
    uint8_t* d_out_s = d_out_d + npx * 16;
    uint8_t* d_nrm = d_out_s + npx * 16;
    uint8_t* d_emi = d_nrm + npx * 8;
    uint8_t* d_dif = d_emi + npx * 8;
    uint8_t* d_spc = d_dif + npx * 4;
    uint8_t* d_dep = d_spc + npx * 4;
    int r;
    // The rasterisation reads the scene only, so it runs on its own stream: with two frames in flight it overlaps the cone
    // trace of the previous frame, which is still queued on `stream`. Per-kernel timing keeps everything on one stream.
    if (!c->raster_stream) {
        CK(c, cudaStreamCreateWithFlags(&c->raster_stream, cudaStreamNonBlocking));
        CK(c, cudaEventCreateWithFlags(&c->ev_raster_done, cudaEventDisableTiming));
    }
    cudaStream_t rs = c->timer.enabled ? s : c->raster_stream;
    if (rs != s && c->ev_scene) CK(c, cudaStreamWaitEvent(rs, c->ev_scene, 0));
    if (shadow) {
        const size_t sb = (size_t)c->light.sw * c->light.sh * sizeof(float);
        if (c->view_shadow_bytes < sb) {
            if (c->view_pending) return fail(c, VGI_E_STATE, "vgi_frame_view_host_begin: the shadow map size changed while a frame is in flight");
            CK(c, cudaStreamSynchronize(c->last_stream));
            CK(c, cudaStreamSynchronize(rs));
            for (int k = 0; k < 2; ++k) { cudaFree(c->view_shadow[k]); c->view_shadow[k] = nullptr; }
            c->view_shadow_bytes = 0;
            for (int k = 0; k < 2; ++k) CK(c, cudaMalloc(&c->view_shadow[k], sb));
            c->view_shadow_bytes = sb;
        }
        r = vgi_render_shadow_map(c, shadow, (uint32_t)c->light.sw, (uint32_t)c->light.sh, c->view_shadow[slot], rs);
        if (r != VGI_OK) return r;
        c->light.depth = c->view_shadow[slot];
        memcpy(c->light.view, shadow->view, sizeof c->light.view);
        memcpy(c->light.proj, shadow->proj, sizeof c->light.proj);
        c->light.z_near = shadow->z_near; c->light.z_far = shadow->z_far;
        c->inc_valid = false;
    }
    vgi_gbuffer dg;
    dg.diffuse_rgba8 = d_dif; dg.normal_rgba16f = d_nrm; dg.specular_rgba8 = d_spc; dg.emission_rgba16f = d_emi;
    dg.depth_f32 = (const float*)d_dep; dg.width = width; dg.height = height;
    r = vgi_render_gbuffer(c, cam, &dg, rs);
    if (r != VGI_OK) return r;
    if (rs != s) {
        CK(c, cudaEventRecord(c->ev_raster_done, rs));
        CK(c, cudaStreamWaitEvent(s, c->ev_raster_done, 0));
    }
    CK(c, cudaMemsetAsync(d_out_d, 0, npx * 32, s));
    r = vgi_update_regions(c, camera_pos);
    if (r != VGI_OK) return r;
    r = vgi_build_clipmap(c, frame_index, stream);
    if (r != VGI_OK) return r;
    // this frame's own copy of the list counters: the next frame's build overwrites the shared block
    CK(c, cudaMemcpyAsync(c->h_view_counters + slot, c->counters, sizeof(Counters), cudaMemcpyDeviceToHost, s));
    vgi_vct_params prm;
    if (params) prm = *params;
    else vgi_default_vct_params(c, &prm);
    c->mark_main_done = c->ev_main_done;
    r = vgi_cone_trace(c, cam, &dg, &prm, d_out_d, d_out_s, stream);
    c->mark_main_done = nullptr;
    if (r != VGI_OK) return r;
    // both downloads ride the copy stream: the diffuse image leaves while the specular cones still march, the specular
    // image while the NEXT frame's rasterisation and build run on `stream`
    CK(c, cudaEventRecord(c->ev_view_traced, s));
    CK(c, cudaStreamWaitEvent(cs, c->ev_main_done, 0));
    CK(c, cudaMemcpyAsync(host_out_diffuse, d_out_d, npx * 16, cudaMemcpyDeviceToHost, cs));
    CK(c, cudaStreamWaitEvent(cs, c->ev_view_traced, 0));
    CK(c, cudaMemcpyAsync(host_out_specular, d_out_s, npx * 16, cudaMemcpyDeviceToHost, cs));
    CK(c, cudaEventRecord(c->ev_view_done[slot], cs));
    c->view_next = slot ^ 1;
    ++c->view_pending;
    return VGI_OK;
}

int vgi_frame_view_host_end(vgi_ctx* c)
{
    if (!c) return fail(c, VGI_E_INVALID, "vgi_frame_view_host_end: null ctx");
    if (!c->view_pending) return fail(c, VGI_E_STATE, "vgi_frame_view_host_end: no frame in flight");
    CK(c, cudaSetDevice(c->device));
    const int slot = c->view_pending == 2 ? c->view_next : (c->view_next ^ 1);     // the oldest frame in flight
    CK(c, cudaEventSynchronize(c->ev_view_done[slot]));
    --c->view_pending;
    return report_overflow(c, "vgi_frame_view_host", c->h_view_counters + slot);
}

int vgi_frame_view_host(vgi_ctx* c, uint32_t frame_index, const float camera_pos[3], const vgi_camera* cam,
                        uint32_t width, uint32_t height, const vgi_dir_light_shadow* shadow,
                        const vgi_vct_params* params, void* host_out_diffuse, void* host_out_specular, void* stream)
{
    if (c && c->view_pending) return fail(c, VGI_E_STATE, "vgi_frame_view_host: frames begun with vgi_frame_view_host_begin are still in flight");
    int r = vgi_frame_view_host_begin(c, frame_index, camera_pos, cam, width, height, shadow, params, host_out_diffuse, host_out_specular, stream);
    if (r != VGI_OK) return r;
    r = vgi_frame_view_host_end(c);
    if (r != VGI_OK) return r;
    CK(c, cudaStreamSynchronize((cudaStream_t)stream));
    return VGI_OK;
}

// ---- Vulkan interop (VK_KHR_external_memory_fd / VK_KHR_external_semaphore_fd) --------------------

int vgi_import_vk_memory(vgi_ctx* c, int fd, size_t size, void** dev_ptr, void** handle)
{
    if (!c || !dev_ptr || !handle || fd < 0 || !size) return fail(c, VGI_E_INVALID, "vgi_import_vk_memory: bad argument");
    CK(c, cudaSetDevice(c->device));
    cudaExternalMemoryHandleDesc hd;
    memset(&hd, 0, sizeof hd);
    hd.type = cudaExternalMemoryHandleTypeOpaqueFd;
    hd.handle.fd = fd;
    hd.size = size;
    cudaExternalMemory_t em;
    CK(c, cudaImportExternalMemory(&em, &hd));
    cudaExternalMemoryBufferDesc bd;
    memset(&bd, 0, sizeof bd);
    bd.offset = 0;
    bd.size = size;
    void* p = nullptr;
    cudaError_t e = cudaExternalMemoryGetMappedBuffer(&p, em, &bd);
    if (e != cudaSuccess) {
        cudaDestroyExternalMemory(em);
        return fail(c, VGI_E_CUDA, std::string("cudaExternalMemoryGetMappedBuffer: ") + cudaGetErrorString(e));
    }
    *dev_ptr = p;
    *handle = (void*)em;
    return VGI_OK;
}

int vgi_release_vk_memory(vgi_ctx* c, void* handle)
{
    if (!c || !handle) return fail(c, VGI_E_INVALID, "vgi_release_vk_memory: bad argument");
    CK(c, cudaStreamSynchronize(c->last_stream));
    CK(c, cudaDestroyExternalMemory((cudaExternalMemory_t)handle));
    return VGI_OK;
}

int vgi_import_vk_semaphore(vgi_ctx* c, int fd, void** handle)
{
    if (!c || !handle || fd < 0) return fail(c, VGI_E_INVALID, "vgi_import_vk_semaphore: bad argument");
    CK(c, cudaSetDevice(c->device));
    cudaExternalSemaphoreHandleDesc sd;
    memset(&sd, 0, sizeof sd);
    sd.type = cudaExternalSemaphoreHandleTypeOpaqueFd;
    sd.handle.fd = fd;
    cudaExternalSemaphore_t sem;
    CK(c, cudaImportExternalSemaphore(&sem, &sd));
    *handle = (void*)sem;
    return VGI_OK;
}

int vgi_wait_vk_semaphore(vgi_ctx* c, void* handle, void* stream)
{
    if (!c || !handle) return fail(c, VGI_E_INVALID, "vgi_wait_vk_semaphore: bad argument");
    cudaExternalSemaphore_t sem = (cudaExternalSemaphore_t)handle;
    cudaExternalSemaphoreWaitParams wp;
    memset(&wp, 0, sizeof wp);
    CK(c, cudaWaitExternalSemaphoresAsync(&sem, &wp, 1, (cudaStream_t)stream));
    return VGI_OK;
}

int vgi_signal_vk_semaphore(vgi_ctx* c, void* handle, void* stream)
{
    if (!c || !handle) return fail(c, VGI_E_INVALID, "vgi_signal_vk_semaphore: bad argument");
    cudaExternalSemaphore_t sem = (cudaExternalSemaphore_t)handle;
    cudaExternalSemaphoreSignalParams sp;
    memset(&sp, 0, sizeof sp);
    CK(c, cudaSignalExternalSemaphoresAsync(&sem, &sp, 1, (cudaStream_t)stream));
    return VGI_OK;
}

int vgi_release_vk_semaphore(vgi_ctx* c, void* handle)
{
    if (!c || !handle) return fail(c, VGI_E_INVALID, "vgi_release_vk_semaphore: bad argument");
    CK(c, cudaDestroyExternalSemaphore((cudaExternalSemaphore_t)handle));
    return VGI_OK;
}

int vgi_synchronize(vgi_ctx* c, void* stream)
{
    if (!c) return fail(c, VGI_E_INVALID, "vgi_synchronize: null ctx");
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaStreamSynchronize((cudaStream_t)stream));
    return VGI_OK;
}

} // extern "C"
