/*
 * vgi_oracle.c — CPU oracle for the clipmap voxel-GI path. See vgi_oracle.h for the contract
 * (TEST INFRASTRUCTURE ONLY, parity status in vgi_oracle.h, IEEE binary32 without FMA contraction).
 *
 * Every function cites the reference file:line it restates. Where the reference leaves behaviour
 * to the Vulkan implementation (raster coverage, UNORM conversion, texture filtering) the software
 * definition used here is spelled out in the comment above the helper.
 */
#include "vgi_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------
 * small helpers
 * ---------------------------------------------------------------------------------------------- */
static inline float f_min(float a, float b) { return a < b ? a : b; }
static inline float f_max(float a, float b) { return a > b ? a : b; }
static inline float f_clamp(float x, float lo, float hi) { return f_min(f_max(x, lo), hi); }
static inline float f_fract(float x) { return x - floorf(x); }
/* GLSL mix(x,y,a) = x*(1-a) + y*a */
static inline float f_mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
/* GLSL pow(x, y): "results are undefined if x < 0". GPUs evaluate exp2(y * log2(x)), i.e. NaN for every negative base,
 * while the C library's powf(negative, integer) is finite. traceCone reaches this when the filtered alpha exceeds 1 by an
 * ulp: with NaN, clamp(1 - NaN, 0, 1) is 0 on the hardware (minNum / maxNum), with powf(-e, 1.0) it would be 1. The oracle
 * follows the hardware (so do the CUDA kernels: exp2f(c * log2f(x))) and the shim (glsl_shim.h). */
static inline float glsl_pow(float x, float y) { return x < 0.0f ? NAN : powf(x, y); }
static inline float dot3(const float* a, const float* b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

/* Vulkan UNORM8 decode: c / 255 (exact IEEE division). */
static inline float unorm8_to_f(uint8_t c) { return (float)c / 255.0f; }
/* Vulkan float -> UNORM8: clamp to [0,1], scale by 255, round to nearest (ties up); NaN -> 0. */
static inline uint8_t f_to_unorm8(float x)
{
    if (!(x > 0.0f)) return 0;
    if (x > 1.0f) x = 1.0f;
    return (uint8_t)(x * 255.0f + 0.5f);
}

static inline uint32_t cfg_R(const vgi_config* c) { return c->resolution; }
static inline uint32_t cfg_L(const vgi_config* c) { return c->level_count; }
static inline size_t atlas_W(const vgi_config* c) { return (size_t)(c->resolution + 2) * VGI_FACES; }
static inline size_t atlas_H(const vgi_config* c) { return (size_t)(c->resolution + 2) * c->level_count; }
static inline size_t atlas_D(const vgi_config* c) { return (size_t)(c->resolution + 2); }

size_t vgo_atlas_bytes(const vgi_config* cfg) { return atlas_W(cfg) * atlas_H(cfg) * atlas_D(cfg) * 4; }

static inline uint8_t* atlas_px(const vgi_config* c, uint8_t* atlas, size_t x, size_t y, size_t z)
{
    return atlas + ((z * atlas_H(c) + y) * atlas_W(c) + x) * 4;
}
static inline const uint8_t* atlas_cpx(const vgi_config* c, const uint8_t* atlas, size_t x, size_t y, size_t z)
{
    return atlas + ((z * atlas_H(c) + y) * atlas_W(c) + x) * 4;
}

/* ------------------------------------------------------------------------------------------------
 * A1 regions. ref: Application.cpp:116-128 (camera box = cam +- (extent0/2)*2^level),
 * VoxelizationPass.cpp:335-357 (minCorner = -R/2, voxelSize = extent0*2^i / R, += delta),
 * VoxelizationPass.cpp:438-448 (delta = trunc((bbMin - minCorner*voxelSize) / (voxelSize*minChange)) * minChange)
 * ---------------------------------------------------------------------------------------------- */
void vgo_regions(const vgi_config* cfg, const float cam[3], vgi_clip_region* out)
{
    const uint32_t R = cfg_R(cfg);
    for (uint32_t i = 0; i < cfg_L(cfg); ++i) {
        vgi_clip_region* r = &out[i];
        const float voxelSize = (cfg->extent_level0 * (float)(1u << i)) / (float)R;
        const float halfSize = (cfg->extent_level0 * 0.5f) * (float)(1u << i);
        const int32_t mc = (int32_t)cfg->clip_min_change[i];
        const float minChange = voxelSize * (float)mc;
        for (int k = 0; k < 3; ++k) {
            int32_t minCorner = -(int32_t)(R >> 1);
            const float bbMin = cam[k] - halfSize;
            const float deltaW = bbMin - ((float)minCorner * voxelSize);
            const int32_t delta = (int32_t)truncf(deltaW / minChange) * mc;
            r->min_corner[k] = minCorner + delta;
            r->extent[k] = R;
        }
        r->voxel_size = voxelSize;
    }
}

/* ------------------------------------------------------------------------------------------------
 * Scene -> world-space triangles. ref: msaaVoxelizer.vert:31-36:
 *   gl_Position = model * vec4(aPosition, 1); normal = (itModel * vec4(aNormal, 0)).xyz
 * mat4*vec4 evaluated column by column, left to right. Draw order: GLTFScene.cpp:457-490
 * (primitive by primitive; drawIndexed(indexCount, 1, firstIndex, vertexOffset, 0)).
 * ---------------------------------------------------------------------------------------------- */
uint32_t vgo_scene_triangle_count(const vgi_scene_desc* s)
{
    uint64_t n = 0;
    for (uint32_t p = 0; p < s->primitive_count; ++p) n += s->primitives[p].index_count / 3;
    return (uint32_t)n;
}

static inline void xform_point(const float* m, const float* v, float* o)
{
    for (int r = 0; r < 3; ++r) o[r] = ((m[0 + r] * v[0] + m[4 + r] * v[1]) + m[8 + r] * v[2]) + m[12 + r];
}
static inline void xform_dir(const float* m, const float* v, float* o)
{
    for (int r = 0; r < 3; ++r) o[r] = (m[0 + r] * v[0] + m[4 + r] * v[1]) + m[8 + r] * v[2];
}

void vgo_scene_triangles(const vgi_scene_desc* s, float* pos, float* nrm, int32_t* mat)
{
    size_t t = 0;
    for (uint32_t p = 0; p < s->primitive_count; ++p) {
        const vgi_primitive* pr = &s->primitives[p];
        const vgi_node_matrix* nm = &s->nodes[pr->node_index];
        for (uint32_t i = 0; i + 2 < pr->index_count; i += 3, ++t) {
            for (int k = 0; k < 3; ++k) {
                const uint32_t vi = s->indices[pr->first_index + i + k] + pr->vertex_offset;
                xform_point(nm->model, s->positions + 3 * (size_t)vi, pos + t * 9 + k * 3);
                xform_dir(nm->it_model, s->normals + 3 * (size_t)vi, nrm + t * 9 + k * 3);
            }
            mat[t] = pr->material_index;
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Material textures. ref: texture() / textureLod() of msaaVoxelizer.frag:64, msaaInjectRadiance.frag:73-82,131-136,
 * voxelizer.frag:52-76 on samplers created with REPEAT + LINEAR and maxLod = 0 (GLTFScene.cpp:339, Sampler.cpp:46): a
 * bi-linear read of level 0 — the same rule oracle/glsl_shim's sampler2D applies when the shaders themselves run.
 * ---------------------------------------------------------------------------------------------- */
static const vgi_texture* g_textures = NULL;
static uint32_t g_texture_count = 0;

void vgo_set_textures(const vgi_texture* textures, uint32_t count)
{
    g_textures = textures;
    g_texture_count = count;
}

static int tex_wrap(long long i, int n)
{
    if (i >= 0 && i < n) return (int)i;
    const long long m = i % n;
    return (int)(m < 0 ? m + n : m);
}

static void tex_fetch(int tex, float u, float v, float* o)
{
    const vgi_texture* t = &g_textures[tex];
    const int W = (int)t->width, H = (int)t->height;
    const uint8_t* d = (const uint8_t*)t->rgba8;
    const float ux = u * (float)W - 0.5f, uy = v * (float)H - 0.5f;
    const float fx = floorf(ux), fy = floorf(uy);
    const float wx = ux - fx, wy = uy - fy;
    const long long ix = (long long)fx, iy = (long long)fy;
    const int x0 = tex_wrap(ix, W), x1 = tex_wrap(ix + 1, W), y0 = tex_wrap(iy, H), y1 = tex_wrap(iy + 1, H);
    const uint8_t* t00 = d + ((size_t)y0 * W + x0) * 4, *t10 = d + ((size_t)y0 * W + x1) * 4;
    const uint8_t* t01 = d + ((size_t)y1 * W + x0) * 4, *t11 = d + ((size_t)y1 * W + x1) * 4;
    for (int k = 0; k < 4; ++k) {
        const float r0 = unorm8_to_f(t00[k]) * (1.0f - wx) + unorm8_to_f(t10[k]) * wx;
        const float r1 = unorm8_to_f(t01[k]) * (1.0f - wx) + unorm8_to_f(t11[k]) * wx;
        o[k] = r0 * (1.0f - wy) + r1 * wy;
    }
}

void vgo_texture_fetch(uint32_t texture, float u, float v, float* rgba) { tex_fetch((int)texture, u, v, rgba); }

/* ------------------------------------------------------------------------------------------------
 * Canonical conservative coverage: a voxel of level l is covered by a triangle iff
 * the triangle overlaps the voxel's cube, evaluated with the Schwarz-Seidel (2010) triangle/box
 * test in voxel units (q = p / voxelSize, unit cubes at integer coordinates), binary32, in exactly
 * the operation order below. Candidate voxels: floor(min(q)) .. floor(max(q)) clipped to the region
 * (ref: msaaVoxelizer.frag:58-60 discards fragments outside the region). Triangles whose world or
 * voxel-space normal is exactly zero produce no coverage (the rasteriser emits no fragments).
 * Dominant axis: ref msaaVoxelizer.geom:27-32 (world space, ties -> z, then y).
 * ---------------------------------------------------------------------------------------------- */
typedef struct tri_setup {
    float n[3], d1, d2;
    float ne[3][3][2]; /* [plane xy,yz,zx][edge][2] */
    float de[3][3];
    int lo[3], hi[3];
    int axis;          /* dominant axis (world) */
    float N[3];        /* world-space cross(p1-p0, p2-p0) */
    int valid;
} tri_setup;

/* cross(p1-p0, p2-p0) and the dominant axis of msaaVoxelizer.geom:27-32 */
static int cross_and_axis(const float* p /*9*/, float* N)
{
    const float a[3] = { p[3] - p[0], p[4] - p[1], p[5] - p[2] };
    const float b[3] = { p[6] - p[0], p[7] - p[1], p[8] - p[2] };
    N[0] = a[1] * b[2] - a[2] * b[1];
    N[1] = a[2] * b[0] - a[0] * b[2];
    N[2] = a[0] * b[1] - a[1] * b[0];
    const float ax = fabsf(N[0]), ay = fabsf(N[1]), az = fabsf(N[2]);
    return (ax > ay && ax > az) ? 0 : ((ay > az) ? 1 : 2);
}

/* test aid: the dominant axis of a world-space triangle (msaaVoxelizer.geom:27-32) */
int vgo_dominant_axis(const float* p /*9*/)
{
    float N[3];
    return cross_and_axis(p, N);
}

/* q: triangle in grid units (unit voxels at integer coordinates); candidate voxels are clipped to
 * [clipLo, clipHi] (inclusive). ts->N / ts->axis must already be set by the caller. */
static void tri_setup_grid(tri_setup* ts, float q[3][3], const int* clipLo, const int* clipHi);

static void tri_setup_init(tri_setup* ts, const float* p /*9*/, float voxelSize,
                           const int32_t* regionMin, uint32_t R)
{
    float q[3][3];
    ts->axis = cross_and_axis(p, ts->N);
    for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 3; ++k) q[i][k] = p[i * 3 + k] / voxelSize;
    const int lo[3] = { regionMin[0], regionMin[1], regionMin[2] };
    const int hi[3] = { regionMin[0] + (int)R - 1, regionMin[1] + (int)R - 1, regionMin[2] + (int)R - 1 };
    tri_setup_grid(ts, q, lo, hi);
}

static void tri_setup_grid(tri_setup* ts, float q[3][3], const int* clipLo, const int* clipHi)
{
    float e[3][3];
    for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 3; ++k) e[i][k] = q[(i + 1) % 3][k] - q[i][k];
    float* n = ts->n;
    n[0] = e[0][1] * e[1][2] - e[0][2] * e[1][1];
    n[1] = e[0][2] * e[1][0] - e[0][0] * e[1][2];
    n[2] = e[0][0] * e[1][1] - e[0][1] * e[1][0];
    ts->valid = !((n[0] == 0.0f && n[1] == 0.0f && n[2] == 0.0f) ||
                  (ts->N[0] == 0.0f && ts->N[1] == 0.0f && ts->N[2] == 0.0f));
    /* also reject non-finite input */
    for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 3; ++k)
            if (!(fabsf(q[i][k]) < 1.0e9f)) ts->valid = 0;
    if (!ts->valid) return;

    const float c[3] = { n[0] > 0.0f ? 1.0f : 0.0f, n[1] > 0.0f ? 1.0f : 0.0f, n[2] > 0.0f ? 1.0f : 0.0f };
    ts->d1 = (n[0] * (c[0] - q[0][0]) + n[1] * (c[1] - q[0][1])) + n[2] * (c[2] - q[0][2]);
    ts->d2 = (n[0] * ((1.0f - c[0]) - q[0][0]) + n[1] * ((1.0f - c[1]) - q[0][1])) +
             n[2] * ((1.0f - c[2]) - q[0][2]);

    /* plane 0: xy (u=x,v=y, sign from n.z); plane 1: yz (u=y,v=z, sign n.x); plane 2: zx (u=z,v=x, sign n.y) */
    static const int U[3] = { 0, 1, 2 }, V[3] = { 1, 2, 0 }, S[3] = { 2, 0, 1 };
    for (int pl = 0; pl < 3; ++pl) {
        const float s = n[S[pl]] >= 0.0f ? 1.0f : -1.0f;
        for (int i = 0; i < 3; ++i) {
            const float nx = -e[i][V[pl]] * s;
            const float ny = e[i][U[pl]] * s;
            ts->ne[pl][i][0] = nx;
            ts->ne[pl][i][1] = ny;
            ts->de[pl][i] = (-(nx * q[i][U[pl]] + ny * q[i][V[pl]]) + f_max(0.0f, nx)) + f_max(0.0f, ny);
        }
    }
    for (int k = 0; k < 3; ++k) {
        const float mn = f_min(q[0][k], f_min(q[1][k], q[2][k]));
        const float mx = f_max(q[0][k], f_max(q[1][k], q[2][k]));
        int lo = (int)floorf(mn), hi = (int)floorf(mx);
        ts->lo[k] = lo > clipLo[k] ? lo : clipLo[k];
        ts->hi[k] = hi < clipHi[k] ? hi : clipHi[k];
    }
}

static inline int tri_overlaps_voxel(const tri_setup* ts, int vx, int vy, int vz)
{
    const float f[3] = { (float)vx, (float)vy, (float)vz };
    const float np = (ts->n[0] * f[0] + ts->n[1] * f[1]) + ts->n[2] * f[2];
    if ((np + ts->d1) * (np + ts->d2) > 0.0f) return 0;
    static const int U[3] = { 0, 1, 2 }, V[3] = { 1, 2, 0 };
    for (int pl = 0; pl < 3; ++pl)
        for (int i = 0; i < 3; ++i)
            if ((ts->ne[pl][i][0] * f[U[pl]] + ts->ne[pl][i][1] * f[V[pl]]) + ts->de[pl][i] < 0.0f) return 0;
    return 1;
}

/* A4. ref: VoxelizationPass.cpp:104-126 */
void vgo_clear_atlas(const vgi_config* cfg, uint8_t* atlas) { memset(atlas, 0, vgo_atlas_bytes(cfg)); }

/* alpha-tested materials: is the (triangle, voxel) pair discarded by the occlusion texture at its canonical sample? A pair
 * without a sample (degenerate projection) is kept by the coverage stage and skipped by the injection. */
static int pair_alpha_discarded(const vgo_tris* tris, int64_t t, const tri_setup* ts, float voxelSize, int x, int y, int z);

/* A3. ref: msaaVoxelizer.frag:43-73 — texel = (v mod R) + 1 (+ level*(R+2) in y), imageStore(vec4(1)) to 6 faces */
uint64_t vgo_voxelize_level(const vgi_config* cfg, const vgi_clip_region* regions, uint32_t level,
                            const vgo_tris* tris, uint8_t* opacity)
{
    const uint32_t R = cfg_R(cfg);
    const vgi_clip_region* rg = &regions[level];
    uint64_t pairs = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : pairs)
    for (int64_t t = 0; t < (int64_t)tris->count; ++t) {
        tri_setup ts;
        tri_setup_init(&ts, tris->pos + t * 9, rg->voxel_size, rg->min_corner, R);
        if (!ts.valid) continue;
        for (int z = ts.lo[2]; z <= ts.hi[2]; ++z)
            for (int y = ts.lo[1]; y <= ts.hi[1]; ++y)
                for (int x = ts.lo[0]; x <= ts.hi[0]; ++x) {
                    if (!tri_overlaps_voxel(&ts, x, y, z)) continue;
                    if (pair_alpha_discarded(tris, t, &ts, rg->voxel_size, x, y, z)) continue;
                    ++pairs;
                    const size_t tx = (size_t)(x & (int)(R - 1)) + 1;
                    const size_t ty = (size_t)(y & (int)(R - 1)) + 1 + (size_t)(R + 2) * level;
                    const size_t tz = (size_t)(z & (int)(R - 1)) + 1;
                    for (uint32_t f = 0; f < VGI_FACES; ++f) {
                        uint32_t* px = (uint32_t*)atlas_px(cfg, opacity, tx + (size_t)f * (R + 2), ty, tz);
                        __atomic_store_n(px, 0xffffffffu, __ATOMIC_RELAXED);
                    }
                }
    }
    return pairs;
}

/* A7. ref: clipmapCleaning.comp:17-31 */
void vgo_clear_region(const vgi_config* cfg, uint8_t* atlas, const int32_t min_corner[3],
                      const uint32_t extent[3], uint32_t level)
{
    const int R = (int)cfg_R(cfg);
    for (uint32_t z = 0; z < extent[2]; ++z)
        for (uint32_t y = 0; y < extent[1]; ++y)
            for (uint32_t x = 0; x < extent[0]; ++x) {
                /* GLSL % on possibly negative operands is undefined; the reference only passes
                 * non-negative corners (RadianceInjectionPass.cpp:77, VoxelizationPass.cpp:139). */
                const int px = (((int)x + min_corner[0]) % R + R) % R + 1;
                const int py = (((int)y + min_corner[1]) % R + R) % R + 1 + (int)level * (R + 2);
                const int pz = (((int)z + min_corner[2]) % R + R) % R + 1;
                for (uint32_t f = 0; f < VGI_FACES; ++f)
                    memset(atlas_px(cfg, atlas, (size_t)px + (size_t)f * (R + 2), py, pz), 0, 4);
            }
}

/* ------------------------------------------------------------------------------------------------
 * Shadow "visibility". ref: shadow.glsl:8-36. Literal (Q1): the sampler has compareEnable = FALSE
 * (VulkanFramework/Images/Sampler.cpp:41-42), so textureProj returns the filtered depth texel and
 * the result is the mean of 16 bilinear depth taps; lightSpacePos.z is computed and unused.
 * Sampler: LINEAR, CLAMP_TO_BORDER, border opaque black (RadianceInjectionPass.cpp:198,
 * VoxelConeTracingPass.cpp:294). Software bilinear: unnormalised u = s*W - 0.5, i0 = floor(u),
 * a = u - i0, result = (t00*(1-a) + t10*a)*(1-b) + (t01*(1-a) + t11*a)*b.
 * VGI_MODE_SHADOW_COMPARE replaces each tap by (tapDepth >= lightSpaceDepth - 0.002 ? 1 : 0)
 * with lightSpaceDepth = (proj * view * p).z — an optional fix, off by default.
 * ---------------------------------------------------------------------------------------------- */
typedef struct shadow_ctx {
    const vgi_dir_light_shadow* d;
    const float* depth;
    uint32_t w, h;
    int compare;
} shadow_ctx;

static inline float shadow_texel(const shadow_ctx* s, int x, int y)
{
    if (x < 0 || y < 0 || x >= (int)s->w || y >= (int)s->h) return 0.0f;
    return s->depth[(size_t)y * s->w + x];
}
static inline float shadow_bilinear(const shadow_ctx* s, float u, float v, float cmpz)
{
    const float x = u * (float)s->w - 0.5f, y = v * (float)s->h - 0.5f;
    const float fx = floorf(x), fy = floorf(y);
    const float a = x - fx, b = y - fy;
    /* keep the int conversion defined for far-away positions */
    const int ix = (int)f_clamp(fx, -4.0f, (float)s->w + 4.0f), iy = (int)f_clamp(fy, -4.0f, (float)s->h + 4.0f);
    float t00 = shadow_texel(s, ix, iy), t10 = shadow_texel(s, ix + 1, iy);
    float t01 = shadow_texel(s, ix, iy + 1), t11 = shadow_texel(s, ix + 1, iy + 1);
    if (s->compare) {
        t00 = t00 >= cmpz ? 1.0f : 0.0f; t10 = t10 >= cmpz ? 1.0f : 0.0f;
        t01 = t01 >= cmpz ? 1.0f : 0.0f; t11 = t11 >= cmpz ? 1.0f : 0.0f;
    }
    return (t00 * (1.0f - a) + t10 * a) * (1.0f - b) + (t01 * (1.0f - a) + t11 * a) * b;
}
static float calc_visibility(const shadow_ctx* s, const float* worldPos)
{
    const float* V = s->d->view;
    const float* P = s->d->proj;
    float l[3];
    xform_point(V, worldPos, l);
    const float zlin = l[2];
    l[2] = l[2] / (s->d->z_far - s->d->z_near); /* ref shadow.glsl:31 — unused afterwards */
    /* proj * vec4(l.xy, 0, 1) */
    float px = ((P[0] * l[0] + P[4] * l[1]) + P[8] * 0.0f) + P[12];
    float py = ((P[1] * l[0] + P[5] * l[1]) + P[9] * 0.0f) + P[13];
    px = px * 0.5f + 0.5f;
    py = py * 0.5f + 0.5f;
    float cmpz = 0.0f;
    if (s->compare) cmpz = (P[10] * zlin + P[14]) - 0.002f;
    const float sx = 1.0f / (float)s->w, sy = 1.0f / (float)s->h;
    float sum = 0.0f;
    for (int j = 0; j < 4; ++j) {
        const float oy = -1.5f + (float)j;
        for (int i = 0; i < 4; ++i) {
            const float ox = -1.5f + (float)i;
            sum += shadow_bilinear(s, px + ox * sx, py + oy * sy, cmpz);
        }
    }
    return sum * 0.0625f;
}

/* ------------------------------------------------------------------------------------------------
 * A8 injection. ref: msaaInjectRadiance.frag:68-155 (shading), :176-183 (face index from -normal),
 * :185-216 (CAS running average). Canonical (Q3/Q10): exactly one contribution per covered
 * (triangle, voxel) pair, evaluated at the voxel centre projected along the triangle's dominant
 * axis onto the triangle plane, barycentrics clamped to >= 0 and renormalised; contributions are
 * quantised to 16 fractional bits, summed as integers, and the texel is floor(255 * mean) — the
 * "exact mean truncated once" of the CAS loop (uint() truncation in atomic.glsl:14-22).
 * ---------------------------------------------------------------------------------------------- */
typedef struct inject_sample {
    float pos[3], nrm[3];
    float b[3];         /* barycentric weights of the sample (after clamping into the triangle) */
    float uv[2];        /* texture coordinate, valid when the caller interpolated it (sample_uv) */
} inject_sample;

static int inject_sample_at(const tri_setup* ts, const float* p, const float* n9, float* c, inject_sample* o);

static int inject_sample_point(const tri_setup* ts, const float* p, const float* n9, float voxelSize,
                               int vx, int vy, int vz, inject_sample* o)
{
    float c[3] = { ((float)vx + 0.5f) * voxelSize, ((float)vy + 0.5f) * voxelSize, ((float)vz + 0.5f) * voxelSize };
    return inject_sample_at(ts, p, n9, c, o);
}

/* c: world-space voxel centre (modified) */
static int inject_sample_at(const tri_setup* ts, const float* p, const float* n9, float* c, inject_sample* o)
{
    const int a = ts->axis;
    const float Na = ts->N[a];
    if (Na == 0.0f) return 0;
    const float d[3] = { c[0] - p[0], c[1] - p[1], c[2] - p[2] };
    const float t = dot3(ts->N, d) / Na;
    c[a] = c[a] - t;
    /* 2-D projection dropping axis a. ref msaaInjectRadiance.geom:33-36 project(): a=0 -> yz, a=1 -> xz, a=2 -> xy */
    const int u = (a == 0) ? 1 : 0;
    const int v = (a == 2) ? 1 : 2;
    const float e1u = p[3 + u] - p[u], e1v = p[3 + v] - p[v];
    const float e2u = p[6 + u] - p[u], e2v = p[6 + v] - p[v];
    const float cu = c[u] - p[u], cv = c[v] - p[v];
    const float den = e1u * e2v - e2u * e1v;
    if (den == 0.0f) return 0;
    float b1 = (cu * e2v - e2u * cv) / den;
    float b2 = (e1u * cv - cu * e1v) / den;
    float b0 = (1.0f - b1) - b2;
    b0 = f_max(b0, 0.0f); b1 = f_max(b1, 0.0f); b2 = f_max(b2, 0.0f);
    const float sum = (b0 + b1) + b2;
    b0 = b0 / sum; b1 = b1 / sum; b2 = b2 / sum;
    for (int k = 0; k < 3; ++k) {
        o->pos[k] = (p[k] * b0 + p[3 + k] * b1) + p[6 + k] * b2;
        o->nrm[k] = (n9[k] * b0 + n9[3 + k] * b1) + n9[6 + k] * b2;
    }
    o->b[0] = b0; o->b[1] = b1; o->b[2] = b2;
    o->uv[0] = o->uv[1] = 0.0f;
    return 1;
}

/* material of triangle t when it samples any texture of the path, else NULL */
static const vgi_material* textured_material(const vgo_tris* tris, int64_t t)
{
    if (!tris->uv || !tris->materials || !g_texture_count) return NULL;
    const vgi_material* m = &tris->materials[tris->mat[t]];
    return (m->base_color_texture > -1 || m->emissive_texture > -1 || m->occlusion_texture > -1) ? m : NULL;
}

static void sample_uv(const vgo_tris* tris, int64_t t, inject_sample* s)
{
    const float* uv = tris->uv + t * 6;
    s->uv[0] = (uv[0] * s->b[0] + uv[2] * s->b[1]) + uv[4] * s->b[2];
    s->uv[1] = (uv[1] * s->b[0] + uv[3] * s->b[1]) + uv[5] * s->b[2];
}

/* ref: msaaVoxelizer.frag:64 / msaaInjectRadiance.frag:73 / voxelizer.frag:52 — occlusion texture .r < 0.1 -> discard,
 * at the canonical sample of the pair. s->uv must be set. */
static int alpha_discard(const vgi_material* m, const inject_sample* s)
{
    if (!m || m->occlusion_texture < 0) return 0;
    float t[4];
    tex_fetch(m->occlusion_texture, s->uv[0], s->uv[1], t);
    return t[0] < 0.1f;
}

static int pair_alpha_discarded(const vgo_tris* tris, int64_t t, const tri_setup* ts, float voxelSize, int x, int y, int z)
{
    const vgi_material* m = textured_material(tris, t);
    if (!m || m->occlusion_texture < 0) return 0;
    inject_sample s;
    if (!inject_sample_point(ts, tris->pos + t * 9, tris->nrm + t * 9, voxelSize, x, y, z, &s)) return 0;
    sample_uv(tris, t, &s);
    return alpha_discard(m, &s);
}

/* returns number of faces written (0 = discarded); faces[i], q[i][3] */
static int shade_fragment(const vgi_material* m, const vgi_dir_light* light, const float* lightDir,
                          const shadow_ctx* sh, const inject_sample* s, int faces[6], uint32_t q[6][3])
{
    if (m->emissive_factor[0] > 0.0f || m->emissive_factor[1] > 0.0f || m->emissive_factor[2] > 0.0f) {
        /* ref msaaInjectRadiance.frag:76-86 */
        float em[3] = { m->emissive_factor[0], m->emissive_factor[1], m->emissive_factor[2] };
        if (g_texture_count && m->emissive_texture > -1) {     /* :79-82 */
            float t[4];
            tex_fetch(m->emissive_texture, s->uv[0], s->uv[1], t);
            for (int k = 0; k < 3; ++k) em[k] = em[k] + t[k];
        }
        for (int f = 0; f < 6; ++f) {
            faces[f] = f;
            for (int k = 0; k < 3; ++k)
                q[f][k] = (uint32_t)(f_clamp(em[k], 0.0f, 1.0f) * 65536.0f + 0.5f);
        }
        return 6;
    }
    const float len2 = dot3(s->nrm, s->nrm);
    if (!(len2 > 0.0f)) return 0;
    const float len = sqrtf(len2);
    const float n[3] = { s->nrm[0] / len, s->nrm[1] / len, s->nrm[2] / len };
    const float NdotL = f_clamp(dot3(n, lightDir), 0.001f, 1.0f);
    const float vis = calc_visibility(sh, s->pos);
    float lc[3];
    for (int k = 0; k < 3; ++k) lc[k] = ((NdotL * vis) * light->color[k]) * light->intensity;
    if (lc[0] == 0.0f && lc[1] == 0.0f && lc[2] == 0.0f) return 0;
    float col[4] = { m->base_color_factor[0], m->base_color_factor[1], m->base_color_factor[2], m->base_color_factor[3] };
    if (g_texture_count && m->base_color_texture > -1) {        /* :131-136 */
        float t[4];
        tex_fetch(m->base_color_texture, s->uv[0], s->uv[1], t);
        for (int k = 0; k < 4; ++k) col[k] = col[k] * t[k];
    }
    float rad[3];
    for (int k = 0; k < 3; ++k)
        rad[k] = f_clamp((lc[k] * col[k]) * col[3], 0.0f, 1.0f);
    /* calculateVoxelFaceIndex(-normal) */
    faces[0] = (-n[0] > 0.0f) ? 0 : 1;
    faces[1] = (-n[1] > 0.0f) ? 2 : 3;
    faces[2] = (-n[2] > 0.0f) ? 4 : 5;
    for (int f = 0; f < 3; ++f) {
        const float w = fabsf(n[f]);
        for (int k = 0; k < 3; ++k) q[f][k] = (uint32_t)((rad[k] * w) * 65536.0f + 0.5f);
    }
    return 3;
}

static void light_dir(const vgi_dir_light* light, float* o)
{
    /* normalize(-direction) */
    const float d[3] = { -light->direction[0], -light->direction[1], -light->direction[2] };
    const float len = sqrtf(dot3(d, d));
    o[0] = d[0] / len; o[1] = d[1] / len; o[2] = d[2] / len;
}

void vgo_inject_level(const vgi_config* cfg, const vgi_clip_region* regions, uint32_t level,
                      const vgo_tris* tris, const vgi_material* materials,
                      const vgi_dir_light* light, const vgi_dir_light_shadow* shadow,
                      const float* shadow_depth, uint32_t sw, uint32_t sh_, uint8_t* radiance)
{
    const uint32_t R = cfg_R(cfg);
    const vgi_clip_region* rg = &regions[level];
    const size_t nvox = (size_t)R * R * R;
    shadow_ctx sc = { shadow, shadow_depth, sw, sh_, (cfg->mode_flags & VGI_MODE_SHADOW_COMPARE) != 0 };
    float lightDir[3];
    light_dir(light, lightDir);

    /* pass 1: which texels receive coverage (compact accumulator indices) */
    uint8_t* occ = (uint8_t*)calloc(nvox, 1);
    int32_t* idx = (int32_t*)malloc(nvox * sizeof(int32_t));
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t t = 0; t < (int64_t)tris->count; ++t) {
        tri_setup ts;
        tri_setup_init(&ts, tris->pos + t * 9, rg->voxel_size, rg->min_corner, R);
        if (!ts.valid) continue;
        for (int z = ts.lo[2]; z <= ts.hi[2]; ++z)
            for (int y = ts.lo[1]; y <= ts.hi[1]; ++y)
                for (int x = ts.lo[0]; x <= ts.hi[0]; ++x)
                    if (tri_overlaps_voxel(&ts, x, y, z) && !pair_alpha_discarded(tris, t, &ts, rg->voxel_size, x, y, z)) {
                        const size_t v = (((size_t)(z & (int)(R - 1)) * R) + (size_t)(y & (int)(R - 1))) * R + (size_t)(x & (int)(R - 1));
                        __atomic_store_n(&occ[v], 1, __ATOMIC_RELAXED);
                    }
    }
    size_t nocc = 0;
    for (size_t v = 0; v < nvox; ++v) idx[v] = occ[v] ? (int32_t)nocc++ : -1;
    uint32_t* acc = (uint32_t*)calloc(nocc * 24 + 1, sizeof(uint32_t)); /* [vox][face][r,g,b,count] */

    /* pass 2: shade and accumulate */
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t t = 0; t < (int64_t)tris->count; ++t) {
        tri_setup ts;
        const float* p = tris->pos + t * 9;
        tri_setup_init(&ts, p, rg->voxel_size, rg->min_corner, R);
        if (!ts.valid) continue;
        const vgi_material* m = &materials[tris->mat[t]];
        for (int z = ts.lo[2]; z <= ts.hi[2]; ++z)
            for (int y = ts.lo[1]; y <= ts.hi[1]; ++y)
                for (int x = ts.lo[0]; x <= ts.hi[0]; ++x) {
                    if (!tri_overlaps_voxel(&ts, x, y, z)) continue;
                    inject_sample s;
                    if (!inject_sample_point(&ts, p, tris->nrm + t * 9, rg->voxel_size, x, y, z, &s)) continue;
                    const vgi_material* tm = textured_material(tris, t);
                    if (tm) {
                        sample_uv(tris, t, &s);
                        if (alpha_discard(tm, &s)) continue;
                    }
                    int faces[6];
                    uint32_t q[6][3];
                    const int nf = shade_fragment(m, light, lightDir, &sc, &s, faces, q);
                    const size_t v = (((size_t)(z & (int)(R - 1)) * R) + (size_t)(y & (int)(R - 1))) * R + (size_t)(x & (int)(R - 1));
                    uint32_t* a = acc + (size_t)idx[v] * 24;
                    for (int f = 0; f < nf; ++f) {
                        uint32_t* af = a + faces[f] * 4;
                        for (int k = 0; k < 3; ++k) {
#pragma omp atomic
                            af[k] += q[f][k];
                        }
#pragma omp atomic
                        af[3] += 1u;
                    }
                }
    }

    /* finalize: texel.rgb = floor(255 * sum / (65536 * count)); alpha is overwritten by copy-alpha */
    for (size_t z = 0; z < R; ++z)
        for (size_t y = 0; y < R; ++y)
            for (size_t x = 0; x < R; ++x) {
                const size_t v = (z * R + y) * R + x;
                if (idx[v] < 0) continue;
                const uint32_t* a = acc + (size_t)idx[v] * 24;
                for (uint32_t f = 0; f < VGI_FACES; ++f) {
                    const uint32_t cnt = a[f * 4 + 3];
                    if (!cnt) continue;
                    uint8_t* px = atlas_px(cfg, radiance, x + 1 + (size_t)f * (R + 2), y + 1 + (size_t)(R + 2) * level, z + 1);
                    for (int k = 0; k < 3; ++k) {
                        uint64_t val = ((uint64_t)a[f * 4 + k] * 255u) / ((uint64_t)cnt * 65536u);
                        px[k] = (uint8_t)(val > 255 ? 255 : val);
                    }
                    px[3] = 0;
                }
            }
    free(acc);
    free(idx);
    free(occ);
}

/* Test aid: the samples vgo_inject_level shades, in (triangle, z, y, x) order, with their shading results
 * (faces[f], q[f][rgb] = the 16-bit fixed-point contributions). tests/test_ref_shaders.py feeds pos / nrm / mat
 * to the reference's msaaInjectRadiance.frag and compares. Arrays may be NULL when capacity == 0 (count only). */
uint64_t vgo_inject_fragments(const vgi_config* cfg, const vgi_clip_region* regions, uint32_t level,
                              const vgo_tris* tris, const vgi_material* materials,
                              const vgi_dir_light* light, const vgi_dir_light_shadow* shadow,
                              const float* shadow_depth, uint32_t sw, uint32_t sh_, uint64_t capacity,
                              float* pos, float* nrm, int32_t* mat, int32_t* voxel, int32_t* nfaces,
                              int32_t* faces, uint32_t* q, float* uv)
{
    const uint32_t R = cfg_R(cfg);
    const vgi_clip_region* rg = &regions[level];
    shadow_ctx sc = { shadow, shadow_depth, sw, sh_, (cfg->mode_flags & VGI_MODE_SHADOW_COMPARE) != 0 };
    float lightDir[3];
    light_dir(light, lightDir);
    uint64_t n = 0;
    for (uint32_t t = 0; t < tris->count; ++t) {
        tri_setup ts;
        const float* p = tris->pos + (size_t)t * 9;
        tri_setup_init(&ts, p, rg->voxel_size, rg->min_corner, R);
        if (!ts.valid) continue;
        const vgi_material* m = &materials[tris->mat[t]];
        for (int z = ts.lo[2]; z <= ts.hi[2]; ++z)
            for (int y = ts.lo[1]; y <= ts.hi[1]; ++y)
                for (int x = ts.lo[0]; x <= ts.hi[0]; ++x) {
                    if (!tri_overlaps_voxel(&ts, x, y, z)) continue;
                    inject_sample s;
                    if (!inject_sample_point(&ts, p, tris->nrm + (size_t)t * 9, rg->voxel_size, x, y, z, &s)) continue;
                    const vgi_material* tm = textured_material(tris, (int64_t)t);
                    if (tm) {
                        sample_uv(tris, (int64_t)t, &s);
                        if (alpha_discard(tm, &s)) continue;
                    }
                    if (n < capacity) {
                        int fc[6];
                        uint32_t qq[6][3];
                        const int nf = shade_fragment(m, light, lightDir, &sc, &s, fc, qq);
                        memcpy(pos + n * 3, s.pos, sizeof s.pos);
                        memcpy(nrm + n * 3, s.nrm, sizeof s.nrm);
                        if (uv) memcpy(uv + n * 2, s.uv, sizeof s.uv);
                        mat[n] = tris->mat[t];
                        voxel[n * 3] = x; voxel[n * 3 + 1] = y; voxel[n * 3 + 2] = z;
                        nfaces[n] = nf;
                        for (int f = 0; f < 6; ++f) {
                            faces[n * 6 + f] = f < nf ? fc[f] : -1;
                            for (int k = 0; k < 3; ++k) q[(n * 6 + f) * 3 + k] = f < nf ? qq[f][k] : 0u;
                        }
                    }
                    ++n;
                }
    }
    return n;
}

/* A9. ref: copyAlphaImage.comp:16-29 — dispatched over R^3 (CopyAlpha.cpp:126-127) */
void vgo_copy_alpha(const vgi_config* cfg, uint32_t level, uint8_t* dst, const uint8_t* src)
{
    const size_t R = cfg_R(cfg);
    for (size_t z = 0; z < R; ++z)
        for (size_t y = 0; y < R; ++y)
            for (size_t x = 0; x < R; ++x)
                for (size_t f = 0; f < VGI_FACES; ++f) {
                    const size_t px = x + 1 + f * (R + 2), py = y + 1 + level * (R + 2), pz = z + 1;
                    atlas_px(cfg, dst, px, py, pz)[3] = atlas_cpx(cfg, src, px, py, pz)[3];
                }
}

/* ------------------------------------------------------------------------------------------------
 * A5 / A10 down-sample. ref: opacityDownSample.comp:28-74 (main), :94-138 (directional pair-over);
 * radianceDownSample.comp:28-75, :95-139 with the out-of-scope lerpFactor (Q6) repaired to the
 * opacity shader's :57-64. Dispatch covers (R/2)^3 invocations (DownSampler.cpp:156).
 * ---------------------------------------------------------------------------------------------- */
static const int DS_OFF[8][3] = { {0,0,0},{1,0,0},{0,1,0},{1,1,0},{0,0,1},{1,0,1},{0,1,1},{1,1,1} };
/* pairs (front, back) per face: front is composited over back along the face's travel direction */
static const int DS_PAIR[6][4][2] = {
    { {0,1},{2,3},{4,5},{6,7} }, /* +x */
    { {1,0},{3,2},{5,4},{7,6} }, /* -x */
    { {0,2},{1,3},{4,6},{5,7} }, /* +y */
    { {2,0},{3,1},{6,4},{7,5} }, /* -y */
    { {0,4},{1,5},{2,6},{3,7} }, /* +z */
    { {4,0},{5,1},{6,2},{7,3} }  /* -z */
};

void vgo_downsample(const vgi_config* cfg, const vgi_clip_region* regions, uint32_t level,
                    uint8_t* atlas, int which)
{
    const int R = (int)cfg_R(cfg);
    const int res = R + 2;
    const uint32_t halfRes = (uint32_t)R >> 1;
    const int32_t* prevMin = regions[level - 1].min_corner;
    const int32_t band = (int32_t)cfg->downsample_band;
    const float invBand = 1.0f / ((float)band + 1.0f);
    const uint32_t thrU = (halfRes >> 1) - (uint32_t)band; /* uint arithmetic as in the GLSL */
    const float thr = (float)thrU;

#pragma omp parallel for collapse(2) schedule(static)
    for (int gz = 0; gz < (int)halfRes; ++gz)
        for (int gy = 0; gy < (int)halfRes; ++gy)
            for (int gx = 0; gx < (int)halfRes; ++gx) {
                const int g[3] = { gx, gy, gz };
                int cur[3], wpos[3], pstart[3];
                float center[3], dist[3];
                for (int k = 0; k < 3; ++k) {
                    cur[k] = (prevMin[k] >> 1) + g[k];
                    wpos[k] = cur[k] & (R - 1);
                    pstart[k] = (cur[k] << 1) & (R - 1);
                    center[k] = (float)(prevMin[k] >> 1) + (float)(halfRes >> 1);
                    dist[k] = fabsf(((float)cur[k] + 0.5f) - center[k]) - 0.5f;
                }
                float lerpFactor = 0.0f;
                if (dist[0] >= thr || dist[1] >= thr || dist[2] >= thr) {
                    lerpFactor = (f_max(dist[0], f_max(dist[1], dist[2])) - thr) + 1.0f;
                    lerpFactor = lerpFactor * invBand;
                }
                for (int f = 0; f < 6; ++f) {
                    float v[8][4];
                    for (int i = 0; i < 8; ++i) {
                        /* child coordinates are NOT re-wrapped after adding OFFSETS (ref :47): start is even, so +1 stays in range */
                        const uint8_t* px = atlas_cpx(cfg, atlas, (size_t)(pstart[0] + DS_OFF[i][0] + 1 + f * res),
                                                      (size_t)(pstart[1] + DS_OFF[i][1] + 1 + res * ((int)level - 1)),
                                                      (size_t)(pstart[2] + DS_OFF[i][2] + 1));
                        for (int c = 0; c < 4; ++c) v[i][c] = unorm8_to_f(px[c]);
                    }
                    uint8_t* out = atlas_px(cfg, atlas, (size_t)(wpos[0] + 1 + f * res),
                                            (size_t)(wpos[1] + 1 + res * (int)level), (size_t)(wpos[2] + 1));
                    if (which == 0) {
                        float s = 0.0f;
                        for (int pr = 0; pr < 4; ++pr) {
                            const float a0 = v[DS_PAIR[f][pr][0]][3], a1 = v[DS_PAIR[f][pr][1]][3];
                            s = s + a0;
                            s = s + (1.0f - a0) * a1;
                        }
                        /* written as one 8-term sum; the first term is v0.a, so start from it exactly */
                        const float ds = s * 0.25f;
                        float op[4];
                        for (int c = 0; c < 4; ++c) op[c] = unorm8_to_f(out[c]);
                        op[3] = f_mix(ds, op[0], lerpFactor);
                        op[1] = op[3];
                        for (int c = 0; c < 4; ++c) out[c] = f_to_unorm8(op[c]);
                    } else {
                        float ds[4];
                        for (int c = 0; c < 4; ++c) {
                            float s = 0.0f;
                            for (int pr = 0; pr < 4; ++pr) {
                                const float* v0 = v[DS_PAIR[f][pr][0]];
                                const float* v1 = v[DS_PAIR[f][pr][1]];
                                s = s + v0[c];
                                s = s + (1.0f - v0[3]) * v1[c];
                            }
                            ds[c] = s * 0.25f;
                        }
                        for (int c = 0; c < 4; ++c)
                            out[c] = f_to_unorm8(f_mix(ds[c], unorm8_to_f(out[c]), lerpFactor));
                    }
                }
            }
}

/* A6. ref: borderWrapping.comp:14-37. Canonical (Q4/Q5 fix): every texel with a coordinate in
 * {0, R+1} receives the wrapped interior texel, for all levels and faces. Literal: only invocations
 * 0 .. 8*((R+2)>>3)-1 run (BorderWrapper.cpp:139), so index R+1 is never written when R%8==0. */
void vgo_wrap_border(const vgi_config* cfg, uint8_t* atlas, int literal)
{
    const int R = (int)cfg_R(cfg);
    const int rb = R + 2;
    const int lim = literal ? (((R + 2) >> 3) << 3) : rb;
    for (int z = 0; z < lim && z < rb; ++z)
        for (int y = 0; y < lim && y < rb; ++y)
            for (int x = 0; x < lim && x < rb; ++x) {
                if (x < R + 1 && y < R + 1 && z < R + 1 && x > 0 && y > 0 && z > 0) continue;
                const int rx = ((x + R - 1) & (R - 1)) + 1;
                const int ry = ((y + R - 1) & (R - 1)) + 1;
                const int rz = ((z + R - 1) & (R - 1)) + 1;
                for (uint32_t l = 0; l < cfg_L(cfg); ++l)
                    for (int f = 0; f < VGI_FACES; ++f)
                        memcpy(atlas_px(cfg, atlas, (size_t)(x + rb * f), (size_t)(y + rb * (int)l), (size_t)z),
                               atlas_cpx(cfg, atlas, (size_t)(rx + rb * f), (size_t)(ry + rb * (int)l), (size_t)rz), 4);
            }
}

/* ref: VoxelizationPass.cpp:74-212 */
uint64_t vgo_voxelization_pass(const vgi_config* cfg, const vgi_clip_region* regions,
                               const vgo_tris* tris, uint8_t* opacity)
{
    uint64_t pairs = 0;
    vgo_clear_atlas(cfg, opacity);                                     /* onBeginRenderPass :104-126 */
    for (uint32_t l = 0; l < cfg_L(cfg); ++l)                          /* onUpdate :190-212 */
        pairs += vgo_voxelize_level(cfg, regions, l, tris, opacity);
    for (uint32_t l = 1; l < cfg_L(cfg); ++l)                          /* onEndRenderPass :166-178 */
        vgo_downsample(cfg, regions, l, opacity, 0);
    vgo_wrap_border(cfg, opacity, (cfg->mode_flags & VGI_MODE_BORDER_LITERAL) != 0); /* :181-186 */
    return pairs;
}

/* ref: RadianceInjectionPass.cpp:64-159. Canonical border handling (Q5 fix): the radiance atlas
 * border is wrapped like the opacity one unless VGI_MODE_BORDER_LITERAL. */
void vgo_injection_pass(const vgi_config* cfg, const vgi_clip_region* regions, const vgo_tris* tris,
                        const vgi_material* materials, const vgi_dir_light* light,
                        const vgi_dir_light_shadow* shadow, const float* shadow_depth,
                        uint32_t sw, uint32_t sh, uint32_t frame_index,
                        const uint8_t* opacity, uint8_t* radiance)
{
    const uint32_t R = cfg_R(cfg);
    const int32_t zero[3] = { 0, 0, 0 };
    const uint32_t ext[3] = { R, R, R };
    for (uint32_t l = 0; l < cfg_L(cfg); ++l)                          /* onBeginRenderPass :73-80 */
        if (frame_index % (1u << l) == 0) vgo_clear_region(cfg, radiance, zero, ext, l);
    for (uint32_t l = 0; l < cfg_L(cfg); ++l)                          /* onUpdate :149-157 */
        if (frame_index % (1u << l) == 0)
            vgo_inject_level(cfg, regions, l, tris, materials, light, shadow, shadow_depth, sw, sh, radiance);
    for (uint32_t l = 0; l < cfg_L(cfg); ++l)                          /* onEndRenderPass :104-114 */
        if (frame_index % (1u << l) == 0) vgo_copy_alpha(cfg, l, radiance, opacity);
    for (uint32_t l = 1; l < cfg_L(cfg); ++l)                          /* :117-130 */
        if (frame_index % (1u << l) == 0) vgo_downsample(cfg, regions, l, radiance, 1);
    if (!(cfg->mode_flags & VGI_MODE_BORDER_LITERAL)) vgo_wrap_border(cfg, radiance, 0);
}

/* ------------------------------------------------------------------------------------------------
 * Cone tracing. ref: voxelConeTracing.frag.
 * ---------------------------------------------------------------------------------------------- */
static float half_to_float(uint16_t h)
{
    const uint32_t sign = (uint32_t)(h >> 15) << 31;
    uint32_t exp = (h >> 10) & 0x1f, man = h & 0x3ff, bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else {
            int e = -1;
            do { ++e; man <<= 1; } while (!(man & 0x400));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ff) << 13);
        }
    } else if (exp == 31) bits = sign | 0x7f800000u | (man << 13);
    else bits = sign | ((exp + 112) << 23) | (man << 13);
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

static const float CONES16[16][3] = { /* ref: voxelConeTracing.frag:118-135 */
    { 0.57735f, 0.57735f, 0.57735f }, { 0.57735f, -0.57735f, -0.57735f },
    { -0.57735f, 0.57735f, -0.57735f }, { -0.57735f, -0.57735f, 0.57735f },
    { -0.903007f, -0.182696f, -0.388844f }, { -0.903007f, 0.182696f, 0.388844f },
    { 0.903007f, -0.182696f, 0.388844f }, { 0.903007f, 0.182696f, -0.388844f },
    { -0.388844f, -0.903007f, -0.182696f }, { 0.388844f, -0.903007f, 0.182696f },
    { 0.388844f, 0.903007f, -0.182696f }, { -0.388844f, 0.903007f, 0.182696f },
    { -0.182696f, -0.388844f, -0.903007f }, { 0.182696f, 0.388844f, -0.903007f },
    { -0.182696f, 0.388844f, 0.903007f }, { 0.182696f, -0.388844f, 0.903007f }
};
static const float CONES32[32][3] = { /* ref: voxelConeTracing.frag:81-114 */
    { 0.898904f, 0.435512f, 0.0479745f }, { 0.898904f, -0.435512f, -0.0479745f },
    { 0.898904f, 0.0479745f, -0.435512f }, { 0.898904f, -0.0479745f, 0.435512f },
    { -0.898904f, 0.435512f, -0.0479745f }, { -0.898904f, -0.435512f, 0.0479745f },
    { -0.898904f, 0.0479745f, 0.435512f }, { -0.898904f, -0.0479745f, -0.435512f },
    { 0.0479745f, 0.898904f, 0.435512f }, { -0.0479745f, 0.898904f, -0.435512f },
    { -0.435512f, 0.898904f, 0.0479745f }, { 0.435512f, 0.898904f, -0.0479745f },
    { -0.0479745f, -0.898904f, 0.435512f }, { 0.0479745f, -0.898904f, -0.435512f },
    { 0.435512f, -0.898904f, 0.0479745f }, { -0.435512f, -0.898904f, -0.0479745f },
    { 0.435512f, 0.0479745f, 0.898904f }, { -0.435512f, -0.0479745f, 0.898904f },
    { 0.0479745f, -0.435512f, 0.898904f }, { -0.0479745f, 0.435512f, 0.898904f },
    { 0.435512f, -0.0479745f, -0.898904f }, { -0.435512f, 0.0479745f, -0.898904f },
    { 0.0479745f, 0.435512f, -0.898904f }, { -0.0479745f, -0.435512f, -0.898904f },
    { 0.57735f, 0.57735f, 0.57735f }, { 0.57735f, 0.57735f, -0.57735f },
    { 0.57735f, -0.57735f, 0.57735f }, { 0.57735f, -0.57735f, -0.57735f },
    { -0.57735f, 0.57735f, 0.57735f }, { -0.57735f, 0.57735f, -0.57735f },
    { -0.57735f, -0.57735f, 0.57735f }, { -0.57735f, -0.57735f, -0.57735f }
};
#define MIN_TRACE_STEP_FACTOR 0.2f
#define MAX_TRACE_DISTANCE 30.0f
#define MIN_SPECULAR_APERTURE 0.05f
#define DIFFUSE_CONE_APERTURE_16 0.872665f
#define DIFFUSE_CONE_APERTURE_32 0.628319f

typedef struct trace_ctx {
    const vgi_config* cfg;
    const vgi_vct_params* prm;
    const uint8_t* atlas;
    size_t W, H, D;
    uint64_t taps;
    /* work statistics (vgo_debug_cell_stats; off unless enabled): which = 0 diffuse / 1 specular cones */
    int stats_on, which, slot;
    int64_t prev_key[2];
    uint64_t st[2][5];      /* steps, level samples, samples whose (level, base cell) differs from the previous step's, */
} trace_ctx;                /* samples that are exactly zero, steps whose whole sample is zero */

static int g_cell_stats_on = 0;
static uint64_t g_cell_stats[2][5];

/* sampler3D, LINEAR, REPEAT (Voxelizer.cpp:183). Software trilinear: unnormalised coordinate
 * c = s*size - 0.5, i0 = floor(c), w = c - i0, texels wrapped modulo size, RGBA8 decoded c/255,
 * result = lerp_z(lerp_y(lerp_x)) with lerp(a,b,w) = a*(1-w) + b*w. */
static void tex3d(const trace_ctx* t, const float* s, float* o)
{
    const float cx = s[0] * (float)t->W - 0.5f, cy = s[1] * (float)t->H - 0.5f, cz = s[2] * (float)t->D - 0.5f;
    const float fx = floorf(cx), fy = floorf(cy), fz = floorf(cz);
    const float wx = cx - fx, wy = cy - fy, wz = cz - fz;
    long ix = (long)fx, iy = (long)fy, iz = (long)fz;
    size_t X[2], Y[2], Z[2];
    const long W = (long)t->W, H = (long)t->H, D = (long)t->D;
    X[0] = (size_t)(((ix % W) + W) % W); X[1] = (size_t)((((ix + 1) % W) + W) % W);
    Y[0] = (size_t)(((iy % H) + H) % H); Y[1] = (size_t)((((iy + 1) % H) + H) % H);
    Z[0] = (size_t)(((iz % D) + D) % D); Z[1] = (size_t)((((iz + 1) % D) + D) % D);
    for (int c = 0; c < 4; ++c) {
        float v[2][2][2];
        for (int k = 0; k < 2; ++k)
            for (int j = 0; j < 2; ++j)
                for (int i = 0; i < 2; ++i)
                    v[k][j][i] = unorm8_to_f(atlas_cpx(t->cfg, t->atlas, X[i], Y[j], Z[k])[c]);
        const float x00 = v[0][0][0] * (1.0f - wx) + v[0][0][1] * wx;
        const float x10 = v[0][1][0] * (1.0f - wx) + v[0][1][1] * wx;
        const float x01 = v[1][0][0] * (1.0f - wx) + v[1][0][1] * wx;
        const float x11 = v[1][1][0] * (1.0f - wx) + v[1][1][1] * wx;
        const float y0 = x00 * (1.0f - wy) + x10 * wy;
        const float y1 = x01 * (1.0f - wy) + x11 * wy;
        o[c] = y0 * (1.0f - wz) + y1 * wz;
    }
}

/* ref: voxelConeTracing.frag:313-327 */
static void sample_clipmap(trace_ctx* t, const float* worldPos, int level, const float* faceOffset,
                           const float* weight, float* o)
{
    const vgi_vct_params* p = t->prm;
    const float L = (float)t->cfg->level_count;
    const float voxelSize = p->voxel_size * exp2f((float)level);
    const float extent = voxelSize * p->volume_dimension;
    float s[3];
    for (int k = 0; k < 3; ++k)
        s[k] = (f_fract(worldPos[k] / extent) * p->volume_dimension + 1.0f) / (p->volume_dimension + 2.0f * 1.0f);
    s[1] += (float)level;
    s[1] /= L;
    s[0] /= (float)VGI_FACES;
    float acc[4] = { 0, 0, 0, 0 };
    for (int f = 0; f < 3; ++f) {
        float sp[3] = { s[0] + faceOffset[f], s[1], s[2] }, tx[4];
        tex3d(t, sp, tx);
        t->taps++;
        for (int c = 0; c < 4; ++c) acc[c] = (f == 0) ? tx[c] * weight[0] : acc[c] + tx[c] * weight[f];
    }
    memcpy(o, acc, sizeof acc);
    if (t->stats_on) {      /* base cell of the tri-linear footprint at this level (face-independent) */
        int64_t key = level;
        for (int k = 0; k < 3; ++k) {
            const float c = f_fract(worldPos[k] / extent) * p->volume_dimension - 0.5f;
            key = key * 4096 + ((int64_t)floorf(c) + 1);
        }
        uint64_t* st = t->st[t->which];
        st[1]++;
        if (key != t->prev_key[t->slot]) st[2]++;
        t->prev_key[t->slot] = key;
        if (acc[0] == 0.0f && acc[1] == 0.0f && acc[2] == 0.0f && acc[3] == 0.0f) st[3]++;
    }
}

/* ref: voxelConeTracing.frag:329-339 */
static void sample_clipmap_linear(trace_ctx* t, const float* worldPos, float curLevel, const int* faceIndex,
                                  const float* weight, float* o)
{
    const int lower = (int)floorf(curLevel), upper = (int)ceilf(curLevel);
    const float fo[3] = { (float)faceIndex[0] / (float)VGI_FACES, (float)faceIndex[1] / (float)VGI_FACES,
                          (float)faceIndex[2] / (float)VGI_FACES };
    float lo[4], up[4];
    t->slot = 0;
    sample_clipmap(t, worldPos, lower, fo, weight, lo);
    t->slot = 1;
    sample_clipmap(t, worldPos, upper, fo, weight, up);
    const float fr = f_fract(curLevel);
    for (int c = 0; c < 4; ++c) o[c] = f_mix(lo[c], up[c], fr);
    if (t->stats_on) {
        t->st[t->which][0]++;
        if (o[0] == 0.0f && o[1] == 0.0f && o[2] == 0.0f && o[3] == 0.0f) t->st[t->which][4]++;
    }
}

/* ref: voxelConeTracing.frag:341-392 */
static void trace_cone(trace_ctx* t, const float* startPos_, const float* dir, float aperture,
                       float maxDistance, float startLevel, float stepFactor, float* out)
{
    const vgi_vct_params* p = t->prm;
    float result[4] = { 0, 0, 0, 0 };
    const float coneCoefficient = 2.0f * tanf(aperture * 0.5f);
    float curLevel = startLevel;
    float voxelSize = p->voxel_size * exp2f(curLevel);
    float startPos[3];
    for (int k = 0; k < 3; ++k) startPos[k] = startPos_[k] + ((dir[k] * voxelSize) * p->trace_start_offset) * 0.5f;
    float step = 0.0f;
    float diameter = f_max(step * coneCoefficient, p->voxel_size);
    float occlusion = 0.0f;
    const int faceIndex[3] = { dir[0] > 0.0f ? 0 : 1, dir[1] > 0.0f ? 2 : 3, dir[2] > 0.0f ? 4 : 5 };
    const float weight[3] = { dir[0] * dir[0], dir[1] * dir[1], dir[2] * dir[2] };
    float curSegmentLength = voxelSize;
    const float minRadius = (p->voxel_size * p->volume_dimension) * 0.5f;
    const float maxLevel = (float)((int)t->cfg->level_count - 1);

    while (step < maxDistance && occlusion < 1.0f) {
        float position[3], d[3];
        for (int k = 0; k < 3; ++k) {
            position[k] = startPos[k] + dir[k] * step;
            d[k] = p->volume_center[k] - position[k];
        }
        const float distanceToVoxelCenter = sqrtf(dot3(d, d));
        const float minLevel = ceilf(log2f(distanceToVoxelCenter / minRadius));
        curLevel = log2f(diameter / p->voxel_size);
        curLevel = f_min(f_max(f_max(startLevel, curLevel), minLevel), maxLevel);

        float smp[4];
        sample_clipmap_linear(t, position, curLevel, faceIndex, weight, smp);
        float radiance[3] = { smp[0], smp[1], smp[2] };
        float opacity = smp[3];
        voxelSize = p->voxel_size * exp2f(curLevel);
        const float correction = curSegmentLength / voxelSize;
        for (int k = 0; k < 3; ++k) radiance[k] = radiance[k] * correction;
        opacity = f_clamp(1.0f - glsl_pow(1.0f - opacity, correction), 0.0f, 1.0f);
        const float k1 = f_clamp(1.0f - result[3], 0.0f, 1.0f);
        result[0] += k1 * radiance[0];
        result[1] += k1 * radiance[1];
        result[2] += k1 * radiance[2];
        result[3] += k1 * opacity;
        occlusion += ((1.0f - occlusion) * opacity) / (1.0f + (step + voxelSize) * p->occlusion_decay);
        const float prevStep = step;
        step += f_max(diameter, p->voxel_size) * stepFactor;
        curSegmentLength = step - prevStep;
        diameter = step * coneCoefficient;
    }
    out[0] = result[0]; out[1] = result[1]; out[2] = result[2];
    out[3] = 1.0f - occlusion;
}

/* ref: voxelConeTracing.frag:394-414 */
static float calc_min_level(const vgi_vct_params* p, const float* worldPos)
{
    const float d[3] = { p->volume_center[0] - worldPos[0], p->volume_center[1] - worldPos[1], p->volume_center[2] - worldPos[2] };
    const float dist = sqrtf(dot3(d, d));
    const float minRadius = (p->voxel_size * p->volume_dimension) * 0.5f;
    const float minLevel = f_max(log2f(dist / minRadius), 0.0f);
    const float radius = minRadius * exp2f(ceilf(minLevel));
    const float f = dist / radius;
    const float transitionStart = 0.5f;
    const float c = 1.0f / (1.0f - transitionStart);
    if (f > transitionStart) return ceilf(minLevel) + (f - transitionStart) * c;
    return ceilf(minLevel);
}

/* ref: brdf.glsl:30-78 */
static void microfacet_brdf(float NdotL, float NdotV, float NdotH, float VdotH, float alphaRoughness,
                            const float* r0, const float* r90, const float* diffuseColor, float* o)
{
    const float M_PI_REF = 3.141592f;
    const float fw = glsl_pow(f_clamp(1.0f - VdotH, 0.0f, 1.0f), 5.0f);
    const float r = alphaRoughness;
    const float attL = 2.0f * NdotL / (NdotL + sqrtf(r * r + (1.0f - r * r) * (NdotL * NdotL)));
    const float attV = 2.0f * NdotV / (NdotV + sqrtf(r * r + (1.0f - r * r) * (NdotV * NdotV)));
    const float G = attL * attV;
    const float rsq = r * r;
    const float ff = (NdotH * rsq - NdotH) * NdotH + 1.0f;
    const float Dm = rsq / (M_PI_REF * ff * ff);
    for (int k = 0; k < 3; ++k) {
        const float F = r0[k] + (r90[k] - r0[k]) * fw;
        const float diffuseContrib = (1.0f - F) * (diffuseColor[k] / M_PI_REF);
        const float specContrib = F * G * Dm / (4.0f * NdotL * NdotV);
        o[k] = NdotL * 1.0f * (diffuseContrib + specContrib);
    }
}

static void normalize3(const float* v, float* o)
{
    const float len = sqrtf(dot3(v, v));
    o[0] = v[0] / len; o[1] = v[1] / len; o[2] = v[2] / len;
}

/* ref: voxelConeTracing.frag:143-294. The G-buffer is fetched at the pixel (texCoord is the pixel
 * centre, so the reference's LINEAR fetch degenerates to the texel). texCoord = ((x+0.5)/w, (y+0.5)/h)
 * from voxelConeTracing.vert; no y flip (Q18). */
static uint64_t g_last_specular_taps = 0;

void vgo_cone_trace(const vgi_config* cfg, const vgi_camera* cam, const vgi_gbuffer* g,
                    const vgi_vct_params* prm, const vgi_dir_light* light,
                    const vgi_dir_light_shadow* shadow, const float* shadow_depth,
                    uint32_t sw, uint32_t sh_, const uint8_t* radiance,
                    float* out_diffuse, float* out_specular, uint32_t y0, uint32_t y1, uint64_t* taps)
{
    shadow_ctx sc = { shadow, shadow_depth, sw, sh_, (cfg->mode_flags & VGI_MODE_SHADOW_COMPARE) != 0 };
    uint64_t total_taps = 0;
    const uint8_t* dif8 = (const uint8_t*)g->diffuse_rgba8;
    const uint8_t* spc8 = (const uint8_t*)g->specular_rgba8;
    const uint16_t* nrm16 = (const uint16_t*)g->normal_rgba16f;
    const uint16_t* emi16 = (const uint16_t*)g->emission_rgba16f;
    float lightV[3];
    light_dir(light, lightV);

    uint64_t spec_taps = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total_taps, spec_taps)
    for (int64_t py = (int64_t)y0; py < (int64_t)y1; ++py) {
        trace_ctx tc;
        memset(&tc, 0, sizeof tc);
        tc.cfg = cfg; tc.prm = prm; tc.atlas = radiance;
        tc.W = atlas_W(cfg); tc.H = atlas_H(cfg); tc.D = atlas_D(cfg);
        tc.stats_on = g_cell_stats_on;
        for (uint32_t px = 0; px < g->width; ++px) {
            const size_t pi = (size_t)py * g->width + px;
            const float depth = g->depth_f32[pi];
            if (depth == 1.0f) continue; /* discard */
            const float tcx = ((float)px + 0.5f) / (float)g->width, tcy = ((float)py + 0.5f) / (float)g->height;
            /* worldPosFromDepth :305-311 */
            float worldPos[3];
            {
                const float v[4] = { tcx * 2.0f - 1.0f, tcy * 2.0f - 1.0f, depth, 1.0f };
                const float* M = cam->view_proj_inv;
                float o[4];
                for (int r = 0; r < 4; ++r) o[r] = ((M[r] * v[0] + M[4 + r] * v[1]) + M[8 + r] * v[2]) + M[12 + r] * v[3];
                worldPos[0] = o[0] / o[3]; worldPos[1] = o[1] / o[3]; worldPos[2] = o[2] / o[3];
            }
            float view[3];
            {
                const float d[3] = { cam->eye_pos[0] - worldPos[0], cam->eye_pos[1] - worldPos[1], cam->eye_pos[2] - worldPos[2] };
                normalize3(d, view);
            }
            const float diffuseColor[3] = { unorm8_to_f(dif8[pi * 4]), unorm8_to_f(dif8[pi * 4 + 1]), unorm8_to_f(dif8[pi * 4 + 2]) };
            const float perceptualRoughness = unorm8_to_f(dif8[pi * 4 + 3]);
            float normal[3];
            {
                const float n[3] = { half_to_float(nrm16[pi * 4]) * 2.0f - 1.0f, half_to_float(nrm16[pi * 4 + 1]) * 2.0f - 1.0f,
                                     half_to_float(nrm16[pi * 4 + 2]) * 2.0f - 1.0f };
                normalize3(n, normal);
            }
            const float specularColor[3] = { unorm8_to_f(spc8[pi * 4]), unorm8_to_f(spc8[pi * 4 + 1]), unorm8_to_f(spc8[pi * 4 + 2]) };
            const float metallic = unorm8_to_f(spc8[pi * 4 + 3]);
            const float emission[3] = { half_to_float(emi16[pi * 4]), half_to_float(emi16[pi * 4 + 1]), half_to_float(emi16[pi * 4 + 2]) };
            const int hasEmission = emission[0] > 0.0f || emission[1] > 0.0f || emission[2] > 0.0f;

            const float minLevel = calc_min_level(prm, worldPos);
            const float voxelSize = prm->voxel_size * exp2f(minLevel);
            float startPos[3];
            for (int k = 0; k < 3; ++k) startPos[k] = worldPos[k] + (normal[k] * voxelSize) * prm->trace_start_offset;

            float indirect[4] = { 0.0f, 0.0f, 0.0f, 1.0f };
            const int ncones = prm->enable_32_cones ? 32 : 16;
            const float (*dirs)[3] = prm->enable_32_cones ? CONES32 : CONES16;
            const float aperture = prm->enable_32_cones ? DIFFUSE_CONE_APERTURE_32 : DIFFUSE_CONE_APERTURE_16;
            for (int i = 0; i < ncones; ++i) {
                const float cosTheta = dot3(normal, dirs[i]);
                if (cosTheta < 0.0f) continue;
                float c[4];
                tc.which = 0; tc.prev_key[0] = tc.prev_key[1] = -1;
                trace_cone(&tc, startPos, dirs[i], aperture, MAX_TRACE_DISTANCE, minLevel,
                           f_max(MIN_TRACE_STEP_FACTOR, prm->min_trace_step_factor), c);
                for (int k = 0; k < 4; ++k) indirect[k] += c[k] * cosTheta;
            }
            for (int k = 0; k < 4; ++k) indirect[k] /= (float)ncones;
            indirect[3] *= prm->ambient_occlusion_factor;
            for (int k = 0; k < 3; ++k) indirect[k] *= diffuseColor[k] * prm->indirect_diffuse_intensity;

            float indirectSpecular[3] = { 0, 0, 0 };
            if ((specularColor[0] > 1e-6f || specularColor[1] > 1e-6f || specularColor[2] > 1e-6f) && metallic > 1e-6f) {
                /* reflect(-view, normal) = I - 2*dot(N,I)*N with I = -view */
                const float I[3] = { -view[0], -view[1], -view[2] };
                const float dn = dot3(normal, I);
                float sdir[3];
                for (int k = 0; k < 3; ++k) sdir[k] = I[k] - (2.0f * dn) * normal[k];
                float c[4];
                const uint64_t taps_before = tc.taps;
                tc.which = 1; tc.prev_key[0] = tc.prev_key[1] = -1;
                trace_cone(&tc, startPos, sdir, f_max(perceptualRoughness, MIN_SPECULAR_APERTURE),
                           MAX_TRACE_DISTANCE, minLevel, prm->voxel_size /* Q12 */, c);
                spec_taps += tc.taps - taps_before;
                for (int k = 0; k < 3; ++k) indirectSpecular[k] += (c[k] * specularColor[k]) * prm->indirect_specular_intensity;
            }

            float direct[3] = { 0, 0, 0 };
            if (hasEmission) {
                for (int k = 0; k < 3; ++k) direct[k] += emission[k];
            } else {
                const float alphaRoughness = perceptualRoughness * perceptualRoughness;
                const float reflectance = f_max(f_max(specularColor[0], specularColor[1]), specularColor[2]);
                const float r90v = f_clamp(reflectance * 50.0f, 0.0f, 1.0f);
                const float r90[3] = { r90v, r90v, r90v };
                float h[3];
                {
                    const float s[3] = { lightV[0] + view[0], lightV[1] + view[1], lightV[2] + view[2] };
                    normalize3(s, h);
                }
                const float NdotL = f_clamp(dot3(normal, lightV), 0.001f, 1.0f);
                const float NdotV = f_clamp(fabsf(dot3(normal, view)), 0.001f, 1.0f);
                const float NdotH = f_clamp(dot3(normal, h), 0.0f, 1.0f);
                const float VdotH = f_clamp(dot3(view, h), 0.0f, 1.0f);
                float brdf[3];
                microfacet_brdf(NdotL, NdotV, NdotH, VdotH, alphaRoughness, specularColor, r90, diffuseColor, brdf);
                const float visibility = calc_visibility(&sc, worldPos);
                for (int k = 0; k < 3; ++k) direct[k] += brdf[k] * visibility;
            }

            float dc[4] = { 0, 0, 0, 1 }, scn[4] = { 0, 0, 0, 1 };
            switch (prm->rendering_mode) {
            case 0: for (int k = 0; k < 3; ++k) dc[k] = diffuseColor[k]; break;
            case 1: for (int k = 0; k < 3; ++k) dc[k] = specularColor[k]; break;
            case 2: for (int k = 0; k < 3; ++k) dc[k] = normal[k] * 0.5f + 0.5f; break;
            case 3: { /* minLevelToColor :416-431, index clamped to the 7-entry table */
                static const float colors[7][4] = { {1,0,0,1},{0,1,0,1},{0,0,1,1},{1,1,0,1},{0,1,1,1},{1,0,1,1},{1,1,1,1} };
                int lower = (int)floorf(minLevel);
                if (lower < 0) lower = 0;
                if (lower > 5) lower = 5;
                const float fr = f_fract(minLevel);
                for (int k = 0; k < 4; ++k) dc[k] = f_mix(colors[lower][k], colors[lower + 1][k], fr) * 0.5f;
                break;
            }
            case 4: for (int k = 0; k < 3; ++k) dc[k] = direct[k] * indirect[3]; break;
            case 5: for (int k = 0; k < 3; ++k) dc[k] = (0.0f + direct[k] * indirect[3]) + indirect[k]; break;
            case 6: for (int k = 0; k < 3; ++k) scn[k] += indirectSpecular[k]; break;
            case 7: for (int k = 0; k < 3; ++k) dc[k] = indirect[3]; break;
            case 8:
                for (int k = 0; k < 3; ++k) {
                    dc[k] = (0.0f + direct[k] * indirect[3]) + indirect[k];
                    scn[k] += indirectSpecular[k];
                }
                break;
            default: break;
            }
            memcpy(out_diffuse + pi * 4, dc, sizeof dc);
            memcpy(out_specular + pi * 4, scn, sizeof scn);
        }
        total_taps += tc.taps;
        if (tc.stats_on)
            for (int w = 0; w < 2; ++w)
                for (int k = 0; k < 5; ++k) __atomic_fetch_add(&g_cell_stats[w][k], tc.st[w][k], __ATOMIC_RELAXED);
    }
    if (taps) *taps = total_taps;
    g_last_specular_taps = spec_taps;
}

/* Work statistics of the cone marches (development aid for the kernels' design, tools/trace_model.py): enable != 0 switches
 * the counting on for later vgo_cone_trace calls; out (may be NULL) receives and resets the counters [diffuse|specular][steps,
 * level samples, samples in a new (level, base cell) compared with the previous step, all-zero samples, all-zero steps]. */
void vgo_debug_cell_stats(int enable, uint64_t* out)
{
    g_cell_stats_on = enable;
    if (out) {
        memcpy(out, g_cell_stats, sizeof g_cell_stats);
        memset(g_cell_stats, 0, sizeof g_cell_stats);
    }
}

/* taps of the specular cones alone in the last vgo_cone_trace call (per-kernel roofline accounting in bench.py) */
uint64_t vgo_last_specular_taps(void) { return g_last_specular_taps; }

/* ------------------------------------------------------------------------------------------------
 * Specular filter + tonemap. ref: specularFilter.frag:25-53, filter.glsl, tonemapping.glsl.
 * texture() on the RGBA32F attachments: bilinear, CLAMP_TO_EDGE, texel centres at +0.5 (Vulkan spec
 * 16.8 "texel filtering", full float weights); texCoord of the fullscreen quad = (pixel + 0.5) / size.
 * Float loop counters are kept literal (32 directions x 8 radii with binary32 accumulation).
 * ---------------------------------------------------------------------------------------------- */
static void tex2d_linear_clamp(const float* img, uint32_t w, uint32_t h, float u, float v, float* o)
{
    const float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
    const float fx = floorf(x), fy = floorf(y);
    const float a = x - fx, b = y - fy;
    int x0 = (int)fx, y0 = (int)fy, x1 = x0 + 1, y1 = y0 + 1;
    x0 = x0 < 0 ? 0 : (x0 > (int)w - 1 ? (int)w - 1 : x0);
    x1 = x1 < 0 ? 0 : (x1 > (int)w - 1 ? (int)w - 1 : x1);
    y0 = y0 < 0 ? 0 : (y0 > (int)h - 1 ? (int)h - 1 : y0);
    y1 = y1 < 0 ? 0 : (y1 > (int)h - 1 ? (int)h - 1 : y1);
    const float* t00 = img + ((size_t)y0 * w + x0) * 4;
    const float* t10 = img + ((size_t)y0 * w + x1) * 4;
    const float* t01 = img + ((size_t)y1 * w + x0) * 4;
    const float* t11 = img + ((size_t)y1 * w + x1) * 4;
    for (int k = 0; k < 4; ++k)
        o[k] = (t00[k] * (1.0f - a) + t10[k] * a) * (1.0f - b) + (t01[k] * (1.0f - a) + t11[k] * a) * b;
}

/* filter.glsl:9-24 */
static void gaussian_blur(const float* img, uint32_t w, uint32_t h, float u, float v, float blurSize, float* o)
{
    const float DOUBLE_PI = 6.28318530718f, DIRECTIONS = 32.0f, QUALITY = 8.0f;
    const float rx = blurSize / (float)w, ry = blurSize / (float)h;
    float color[4];
    tex2d_linear_clamp(img, w, h, u, v, color);
    for (float d = 0.0f; d < DOUBLE_PI; d += DOUBLE_PI / DIRECTIONS)
        for (float i = 1.0f / QUALITY; i <= 1.0f; i += 1.0f / QUALITY) {
            float t[4];
            tex2d_linear_clamp(img, w, h, u + (cosf(d) * rx) * i, v + (sinf(d) * ry) * i, t);
            for (int k = 0; k < 4; ++k) color[k] += t[k];
        }
    for (int k = 0; k < 4; ++k) o[k] = color[k] / (QUALITY * DIRECTIONS - 15.0f);
}

/* filter.glsl:27-63 */
static float normpdf(float x, float sigma) { return 0.39894f * expf(-0.5f * x * x / (sigma * sigma)) / sigma; }
static void bilateral_filter(const float* img, uint32_t w, uint32_t h, float u, float v, float* o)
{
    static const float KERNEL[15] = { 0.031225216f, 0.033322271f, 0.035206333f, 0.036826804f, 0.038138565f,
        0.039104044f, 0.039695028f, 0.039894000f, 0.039695028f, 0.039104044f, 0.038138565f, 0.036826804f,
        0.035206333f, 0.033322271f, 0.031225216f };
    const float BSIGMA = 10.0f;
    float fin[3] = { 0.f, 0.f, 0.f }, center[4];
    tex2d_linear_clamp(img, w, h, u, v, center);
    float Z = 0.0f;
    const float bZ = 1.0f / normpdf(0.0f, BSIGMA);
    const float irx = 1.0f / (float)w, iry = 1.0f / (float)h;
    for (int i = -7; i <= 7; ++i)
        for (int j = -7; j <= 7; ++j) {
            float s[4];
            tex2d_linear_clamp(img, w, h, u + (float)i * irx, v + (float)j * iry, s);
            const float dv[3] = { s[0] - center[0], s[1] - center[1], s[2] - center[2] };
            const float pdf3 = 0.39894f * expf(-0.5f * dot3(dv, dv) / (BSIGMA * BSIGMA)) / BSIGMA;
            const float factor = pdf3 * bZ * KERNEL[7 + j] * KERNEL[7 + i];
            Z += factor;
            for (int k = 0; k < 3; ++k) fin[k] += factor * s[k];
        }
    for (int k = 0; k < 3; ++k) o[k] = fin[k] / Z;
}

/* tonemapping.glsl:9-26 */
static float uncharted2(float c)
{
    const float A = 0.15f, B = 0.50f, C = 0.10f, D = 0.20f, E = 0.02f, F = 0.30f;
    return ((c * (A * c + C * B) + D * E) / (c * (A * c + B) + D * F)) - E / F;
}

void vgo_specular_filter(const float* diffuse, const float* specular, uint32_t w, uint32_t h,
                         const vgi_filter_params* prm, float* out)
{
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t py = 0; py < (int64_t)h; ++py)
        for (uint32_t px = 0; px < w; ++px) {
            const float u = ((float)px + 0.5f) / (float)w, v = ((float)py + 0.5f) / (float)h;
            float fc[4];
            tex2d_linear_clamp(diffuse, w, h, u, v, fc);
            float sc[4] = { 0.f, 0.f, 0.f, 1.f };
            if (prm->filter_method == 1) gaussian_blur(specular, w, h, u, v, 0.01f, sc);
            else bilateral_filter(specular, w, h, u, v, sc);
            for (int k = 0; k < 3; ++k) fc[k] += sc[k];
            float* o = out + ((size_t)py * w + px) * 4;
            if (prm->tonemap_enable == 1) {
                const float white = 1.0f / uncharted2(11.2f);
                for (int k = 0; k < 3; ++k)
                    o[k] = glsl_pow(uncharted2(fc[k] * prm->tonemap_exposure) * white, 1.0f / prm->tonemap_gamma);
                o[3] = fc[3];
            } else {
                memcpy(o, fc, sizeof fc);
            }
        }
}

/* SVO path: shares the helpers above (single translation unit). */
/* thread count of the OpenMP loops (torchrun exports OMP_NUM_THREADS=1; the CPU baseline uses every core) */
int vgo_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

#include "vgi_oracle_svo.inc"
#include "vgi_oracle_literal.inc"
