/*
 * synth_raster.c — host-side producers of the cone tracer's INPUTS for headless runs:
 * the shadow-map depth image and the G-buffer that the reference renders with Vulkan
 * (ref: VFS/Shaders/shadowPass.vert:33, gBufferPass.vert:45, gBufferPass.frag:62-116;
 * formats ref: VFS/RenderPass/GBufferPass.cpp:177-194). They are NOT part of the hot path
 * (SURVEY.md section 2 rows 15-16 are out of scope) and are never timed; tests and bench.py use them to
 * build identical inputs for the CUDA path and the CPU oracle. Plain C + OpenMP, no CUDA.
 *
 * Rasterisation: pixel-centre sampling, depth test LESS with ties broken by the lower triangle
 * index; a triangle with a vertex at or behind the camera plane (clip w <= 1e-4) is rasterised with 2-D homogeneous edge
 * functions inside the pixel box of its part in front of the plane (same rule and arithmetic as csrc/vgi_raster.cu),
 * fragments with z outside [0,1] are clipped. Materials: factors and textures (base colour, metallic-roughness, emissive,
 * tangent-space normal map, alpha cutoff: gBufferPass.frag:39-116), bilinear REPEAT level-0 reads like csrc/vgi_device.cuh.
 */
#include "../../include/vgi.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static uint16_t float_to_half(float f)
{
    uint32_t x;
    memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    const int32_t exp = (int32_t)((x >> 23) & 0xff) - 127 + 15;
    uint32_t man = x & 0x7fffffu;
    if (((x >> 23) & 0xff) == 0xff) return (uint16_t)(sign | 0x7c00u | (man ? 0x200u : 0));
    if (exp >= 31) return (uint16_t)(sign | 0x7c00u);
    if (exp <= 0) {
        if (exp < -10) return (uint16_t)sign;
        man |= 0x800000u;
        const int shift = 14 - exp;
        uint32_t h = man >> shift;
        const uint32_t rem = man & ((1u << shift) - 1), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (h & 1))) ++h;
        return (uint16_t)(sign | h);
    }
    uint32_t h = ((uint32_t)exp << 10) | (man >> 13);
    const uint32_t rem = man & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) ++h;
    return (uint16_t)(sign | h);
}

static uint8_t to_unorm8(float x)
{
    if (!(x > 0.0f)) return 0;
    if (x > 1.0f) x = 1.0f;
    return (uint8_t)(x * 255.0f + 0.5f);
}

typedef struct tri_world {
    float p[3][3];
    float n[3][3];
    float uv[3][2];
    float tg[3][4];     /* itModel * tangent.xyz, handedness (gBufferPass.vert:40) */
    int32_t mat;
} tri_world;

typedef struct tex_set { const vgi_texture* t; uint32_t n; } tex_set;

static int tex_wrap(long long i, int n)
{
    if (i >= 0 && i < n) return (int)i;
    const long long m = i % n;
    return (int)(m < 0 ? m + n : m);
}

/* bilinear REPEAT read of mip level 0 (quirk Q23), the arithmetic of tex_fetch in csrc/vgi_device.cuh */
static void tex_fetch(const tex_set* ts, int tex, float u, float v, float o[4])
{
    const int W = (int)ts->t[tex].width, H = (int)ts->t[tex].height;
    const uint32_t* d = (const uint32_t*)ts->t[tex].rgba8;
    const float ux = u * (float)W - 0.5f, uy = v * (float)H - 0.5f;
    const float fx = floorf(ux), fy = floorf(uy);
    const float wx = ux - fx, wy = uy - fy;
    const long long ix = (long long)fx, iy = (long long)fy;
    const int x0 = tex_wrap(ix, W), x1 = tex_wrap(ix + 1, W), y0 = tex_wrap(iy, H), y1 = tex_wrap(iy + 1, H);
    const uint32_t t00 = d[(size_t)y0 * W + x0], t10 = d[(size_t)y0 * W + x1];
    const uint32_t t01 = d[(size_t)y1 * W + x0], t11 = d[(size_t)y1 * W + x1];
    for (int k = 0; k < 4; ++k) {
        const float a = (float)((t00 >> (8 * k)) & 0xffu) / 255.0f, b = (float)((t10 >> (8 * k)) & 0xffu) / 255.0f;
        const float c = (float)((t01 >> (8 * k)) & 0xffu) / 255.0f, e = (float)((t11 >> (8 * k)) & 0xffu) / 255.0f;
        const float r0 = a * (1.0f - wx) + b * wx;
        const float r1 = c * (1.0f - wx) + e * wx;
        o[k] = r0 * (1.0f - wy) + r1 * wy;
    }
}

static void xform_point(const float* m, const float* v, float* o)
{
    for (int r = 0; r < 3; ++r) o[r] = ((m[r] * v[0] + m[4 + r] * v[1]) + m[8 + r] * v[2]) + m[12 + r];
}
static void xform_dir(const float* m, const float* v, float* o)
{
    for (int r = 0; r < 3; ++r) o[r] = (m[r] * v[0] + m[4 + r] * v[1]) + m[8 + r] * v[2];
}

static tri_world* world_tris(const vgi_scene_desc* s, uint32_t* count)
{
    uint64_t n = 0;
    for (uint32_t p = 0; p < s->primitive_count; ++p) n += s->primitives[p].index_count / 3;
    tri_world* out = (tri_world*)malloc((n ? n : 1) * sizeof(tri_world));
    size_t t = 0;
    for (uint32_t p = 0; p < s->primitive_count; ++p) {
        const vgi_primitive* pr = &s->primitives[p];
        const vgi_node_matrix* nm = &s->nodes[pr->node_index];
        for (uint32_t i = 0; i + 2 < pr->index_count; i += 3, ++t) {
            for (int k = 0; k < 3; ++k) {
                const uint32_t vi = s->indices[pr->first_index + i + k] + pr->vertex_offset;
                xform_point(nm->model, s->positions + 3 * (size_t)vi, out[t].p[k]);
                xform_dir(nm->it_model, s->normals + 3 * (size_t)vi, out[t].n[k]);
                out[t].uv[k][0] = s->texcoords ? s->texcoords[2 * (size_t)vi] : 0.0f;
                out[t].uv[k][1] = s->texcoords ? s->texcoords[2 * (size_t)vi + 1] : 0.0f;
                out[t].tg[k][0] = out[t].tg[k][1] = out[t].tg[k][2] = out[t].tg[k][3] = 0.0f;
                if (s->tangents) {
                    xform_dir(nm->it_model, s->tangents + 4 * (size_t)vi, out[t].tg[k]);
                    out[t].tg[k][3] = s->tangents[4 * (size_t)vi + 3];
                }
            }
            out[t].mat = pr->material_index;
        }
    }
    *count = (uint32_t)n;
    return out;
}

#define RASTER_W_EPS 1e-4

typedef struct proj_tri {
    /* ok == 1: pixel coords, ndc depth, 1/w. ok == 2 (crosses the camera plane): e_i = x[i]*px + y[i]*py + z[i],
     * iw[] = clip z, wc[] = clip w of the vertices, box = pixel box of the part in front (x0, x1, y0, y1) */
    double x[3], y[3], z[3], iw[3];
    double wc[3];
    int box[4];
    int ok;
} proj_tri;

static void project(const float* M, const tri_world* t, uint32_t w, uint32_t h, proj_tri* o)
{
    double c[3][4];
    int behind = 0;
    for (int k = 0; k < 3; ++k) {
        const float* p = t->p[k];
        for (int r = 0; r < 4; ++r)
            c[k][r] = (double)M[r] * p[0] + (double)M[4 + r] * p[1] + (double)M[8 + r] * p[2] + (double)M[12 + r];
        behind += c[k][3] <= RASTER_W_EPS ? 1 : 0;
    }
    o->box[0] = o->box[2] = 0; o->box[1] = o->box[3] = -1;
    o->wc[0] = o->wc[1] = o->wc[2] = 0.0;
    if (behind == 3) { o->ok = 0; return; }
    if (behind == 0) {
        o->ok = 1;
        for (int k = 0; k < 3; ++k) {
            o->iw[k] = 1.0 / c[k][3];
            o->x[k] = (c[k][0] * o->iw[k] * 0.5 + 0.5) * (double)w;
            o->y[k] = (c[k][1] * o->iw[k] * 0.5 + 0.5) * (double)h;
            o->z[k] = c[k][2] * o->iw[k];
        }
        return;
    }
    o->ok = 2;
    double X[3], Y[3], W[3];
    for (int k = 0; k < 3; ++k) {
        X[k] = (c[k][0] * 0.5 + c[k][3] * 0.5) * (double)w;
        Y[k] = (c[k][1] * 0.5 + c[k][3] * 0.5) * (double)h;
        W[k] = c[k][3];
        o->iw[k] = c[k][2];
        o->wc[k] = c[k][3];
    }
    double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
    for (int i = 0; i < 3; ++i) {
        const int j = (i + 1) % 3, k = (i + 2) % 3;
        o->x[i] = Y[j] * W[k] - W[j] * Y[k];
        o->y[i] = W[j] * X[k] - X[j] * W[k];
        o->z[i] = X[j] * Y[k] - Y[j] * X[k];
        if (W[i] > RASTER_W_EPS) {
            const double px = X[i] / W[i], py = Y[i] / W[i];
            xmin = fmin(xmin, px); xmax = fmax(xmax, px); ymin = fmin(ymin, py); ymax = fmax(ymax, py);
        }
        if ((W[i] > RASTER_W_EPS) != (W[j] > RASTER_W_EPS)) {
            const double u = (RASTER_W_EPS - W[i]) / (W[j] - W[i]);
            const double px = (X[i] + u * (X[j] - X[i])) / RASTER_W_EPS, py = (Y[i] + u * (Y[j] - Y[i])) / RASTER_W_EPS;
            xmin = fmin(xmin, px); xmax = fmax(xmax, px); ymin = fmin(ymin, py); ymax = fmax(ymax, py);
        }
    }
    if (ymax < 0 || ymin > h || xmax < 0 || xmin > w) return;
    xmin = fmax(xmin, -1.0); ymin = fmax(ymin, -1.0); xmax = fmin(xmax, (double)w + 1.0); ymax = fmin(ymax, (double)h + 1.0);
    o->box[0] = (int)floor(xmin - 0.5) > 0 ? (int)floor(xmin - 0.5) : 0;
    o->box[1] = (int)ceil(xmax - 0.5) < (int)w - 1 ? (int)ceil(xmax - 0.5) : (int)w - 1;
    o->box[2] = (int)floor(ymin - 0.5) > 0 ? (int)floor(ymin - 0.5) : 0;
    o->box[3] = (int)ceil(ymax - 0.5) < (int)h - 1 ? (int)ceil(ymax - 0.5) : (int)h - 1;
}

static int hom_bary(const proj_tri* q, double px, double py, double* b, double* z)
{
    const double e0 = (q->x[0] * px + q->y[0] * py) + q->z[0];
    const double e1 = (q->x[1] * px + q->y[1] * py) + q->z[1];
    const double e2 = (q->x[2] * px + q->y[2] * py) + q->z[2];
    const double s = (e0 + e1) + e2;
    if (s == 0.0) return 0;
    b[0] = e0 / s; b[1] = e1 / s; b[2] = e2 / s;
    if (b[0] < 0 || b[1] < 0 || b[2] < 0) return 0;
    const double wc = (b[0] * q->wc[0] + b[1] * q->wc[1]) + b[2] * q->wc[2];
    if (!(wc > RASTER_W_EPS)) return 0;
    *z = ((b[0] * q->iw[0] + b[1] * q->iw[1]) + b[2] * q->iw[2]) / wc;
    return 1;
}

/* perspective-correct barycentric weights of the pixel centre on its winning triangle (both kinds) */
static void pixel_bary(const proj_tri* q, double px, double py, double* b)
{
    if (q->ok == 2) {
        double z;
        hom_bary(q, px, py, b, &z);
        return;
    }
    const double area = (q->x[1] - q->x[0]) * (q->y[2] - q->y[0]) - (q->x[2] - q->x[0]) * (q->y[1] - q->y[0]);
    double b0 = ((q->x[1] - px) * (q->y[2] - py) - (q->x[2] - px) * (q->y[1] - py)) / area;
    double b1 = ((q->x[2] - px) * (q->y[0] - py) - (q->x[0] - px) * (q->y[2] - py)) / area;
    double b2 = 1.0 - b0 - b1;
    b0 *= q->iw[0]; b1 *= q->iw[1]; b2 *= q->iw[2];
    const double bs = b0 + b1 + b2;
    b[0] = b0 / bs; b[1] = b1 / bs; b[2] = b2 / bs;
}

static void frag_uv(const tri_world* t, const double* b, float* uv)
{
    uv[0] = (float)((b[0] * t->uv[0][0] + b[1] * t->uv[1][0]) + b[2] * t->uv[2][0]);
    uv[1] = (float)((b[0] * t->uv[0][1] + b[1] * t->uv[1][1]) + b[2] * t->uv[2][1]);
}

/* ref: gBufferPass.frag:88-99 — 1 = the fragment survives the alpha cutoff */
static int alpha_keep(const proj_tri* q, const tri_world* t, const vgi_material* m, const tex_set* ts, double px, double py)
{
    double b[3];
    float uv[2], tx[4];
    pixel_bary(q, px, py, b);
    frag_uv(t, b, uv);
    tex_fetch(ts, m->base_color_texture, uv[0], uv[1], tx);
    return !(m->base_color_factor[3] * tx[3] < m->alpha_cutoff);
}

/* depth + winning triangle id per pixel. materials != NULL: the G-buffer pass (alpha cutoff discards before the depth write);
 * NULL: the shadow pass (no fragment shader). */
static void raster_ids(const float* M, const tri_world* tris, uint32_t ntri, uint32_t w, uint32_t h,
                       float* depth, int32_t* ids, const vgi_material* materials, const tex_set* ts)
{
    for (size_t i = 0; i < (size_t)w * h; ++i) { depth[i] = 1.0f; ids[i] = -1; }
    proj_tri* pt = (proj_tri*)malloc((ntri ? ntri : 1) * sizeof(proj_tri));
    uint8_t* alpha = (uint8_t*)calloc(ntri ? ntri : 1, 1);
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < (int64_t)ntri; ++t) {
        project(M, &tris[t], w, h, &pt[t]);
        if (materials) {
            const vgi_material* m = &materials[tris[t].mat];
            if (m->alpha_mode > 0) {
                if (m->base_color_texture > -1) alpha[t] = 1;
                else if (m->base_color_factor[3] < m->alpha_cutoff) pt[t].ok = 0;
            }
        }
    }

    const int band = 16;
    const int nbands = (int)((h + band - 1) / band);
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < nbands; ++b) {
        const int by0 = b * band, by1 = (by0 + band < (int)h) ? by0 + band : (int)h;
        for (uint32_t t = 0; t < ntri; ++t) {
            const proj_tri* q = &pt[t];
            if (!q->ok) continue;
            if (q->ok == 2) {
                const int hy0 = q->box[2] > by0 ? q->box[2] : by0, hy1 = q->box[3] < by1 - 1 ? q->box[3] : by1 - 1;
                for (int y = hy0; y <= hy1; ++y)
                    for (int x = q->box[0]; x <= q->box[1]; ++x) {
                        double b[3], z;
                        if (!hom_bary(q, x + 0.5, y + 0.5, b, &z)) continue;
                        if (z < 0.0 || z > 1.0) continue;
                        const float zf = (float)z;
                        const size_t pi = (size_t)y * w + x;
                        if (!(zf < depth[pi])) continue;
                        if (alpha[t] && !alpha_keep(q, &tris[t], &materials[tris[t].mat], ts, x + 0.5, y + 0.5)) continue;
                        depth[pi] = zf; ids[pi] = (int32_t)t;
                    }
                continue;
            }
            const double ymin = fmin(q->y[0], fmin(q->y[1], q->y[2])), ymax = fmax(q->y[0], fmax(q->y[1], q->y[2]));
            if (ymax < by0 || ymin > by1) continue;
            const double xmin = fmin(q->x[0], fmin(q->x[1], q->x[2])), xmax = fmax(q->x[0], fmax(q->x[1], q->x[2]));
            if (xmax < 0 || xmin > w) continue;
            const double area = (q->x[1] - q->x[0]) * (q->y[2] - q->y[0]) - (q->x[2] - q->x[0]) * (q->y[1] - q->y[0]);
            if (area == 0.0) continue;
            int x0 = (int)floor(xmin - 0.5), x1 = (int)ceil(xmax - 0.5);
            int y0 = (int)floor(ymin - 0.5), y1 = (int)ceil(ymax - 0.5);
            if (x0 < 0) x0 = 0;
            if (x1 > (int)w - 1) x1 = (int)w - 1;
            if (y0 < by0) y0 = by0;
            if (y1 > by1 - 1) y1 = by1 - 1;
            for (int y = y0; y <= y1; ++y)
                for (int x = x0; x <= x1; ++x) {
                    const double px = x + 0.5, py = y + 0.5;
                    const double w0 = ((q->x[1] - px) * (q->y[2] - py) - (q->x[2] - px) * (q->y[1] - py)) / area;
                    const double w1 = ((q->x[2] - px) * (q->y[0] - py) - (q->x[0] - px) * (q->y[2] - py)) / area;
                    const double w2 = 1.0 - w0 - w1;
                    if (w0 < 0 || w1 < 0 || w2 < 0) continue;
                    const double z = w0 * q->z[0] + w1 * q->z[1] + w2 * q->z[2];
                    if (z < 0.0 || z > 1.0) continue;
                    const float zf = (float)z;
                    const size_t pi = (size_t)y * w + x;
                    if (!(zf < depth[pi])) continue;
                    if (alpha[t] && !alpha_keep(q, &tris[t], &materials[tris[t].mat], ts, px, py)) continue;
                    depth[pi] = zf; ids[pi] = (int32_t)t;
                }
        }
    }
    free(alpha);
    free(pt);
}

/* ref: shadowPass.vert:33 — depth-only render of the scene with proj*view; clear 1.0 */
int vgs_shadow_depth(const vgi_scene_desc* scene, const vgi_dir_light_shadow* sh, uint32_t w, uint32_t h, float* depth)
{
    uint32_t ntri;
    tri_world* tris = world_tris(scene, &ntri);
    float M[16];
    for (int c = 0; c < 4; ++c)
        for (int r = 0; r < 4; ++r) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += (double)sh->proj[k * 4 + r] * sh->view[c * 4 + k];
            M[c * 4 + r] = (float)s;
        }
    int32_t* ids = (int32_t*)malloc((size_t)w * h * sizeof(int32_t));
    raster_ids(M, tris, ntri, w, h, depth, ids, NULL, NULL);
    free(ids);
    free(tris);
    return 0;
}

/* ref: gBufferPass.frag:62-116; textures: the array vgi_set_textures would get (may be NULL / 0 for factor-only scenes) */
int vgs_gbuffer_tex(const vgi_scene_desc* scene, const vgi_texture* textures, uint32_t texture_count, const vgi_camera* cam,
                    uint32_t w, uint32_t h, uint8_t* diffuse, uint16_t* normal, uint8_t* specular, uint16_t* emission, float* depth)
{
    uint32_t ntri;
    tri_world* tris = world_tris(scene, &ntri);
    const tex_set ts = { textures, texture_count };
    int32_t* ids = (int32_t*)malloc((size_t)w * h * sizeof(int32_t));
    raster_ids(cam->view_proj, tris, ntri, w, h, depth, ids, scene->materials, &ts);
#pragma omp parallel for schedule(static)
    for (int64_t y = 0; y < (int64_t)h; ++y)
        for (uint32_t x = 0; x < w; ++x) {
            const size_t pi = (size_t)y * w + x;
            memset(diffuse + pi * 4, 0, 4);
            memset(specular + pi * 4, 0, 4);
            memset(normal + pi * 4, 0, 8);
            memset(emission + pi * 4, 0, 8);
            if (ids[pi] < 0) continue;
            const tri_world* t = &tris[ids[pi]];
            proj_tri q;
            project(cam->view_proj, t, w, h, &q);
            double bb[3];
            pixel_bary(&q, x + 0.5, y + 0.5, bb);
            const double b0 = bb[0], b1 = bb[1], b2 = bb[2];
            double n[3];
            for (int k = 0; k < 3; ++k) n[k] = b0 * t->n[0][k] + b1 * t->n[1][k] + b2 * t->n[2][k];
            double ln = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            const vgi_material* m = &scene->materials[t->mat];
            float uv[2], tx[4];
            frag_uv(t, bb, uv);
            float rough = m->roughness_factor, metal = m->metallic_factor;
            if (m->metallic_roughness_texture > -1) {   /* no clamp in this branch (gBufferPass.frag:77-82) */
                tex_fetch(&ts, m->metallic_roughness_texture, uv[0], uv[1], tx);
                rough *= tx[1];
                metal *= tx[2];
            } else {
                rough = rough < 0.04f ? 0.04f : (rough > 1.0f ? 1.0f : rough); /* MIN_ROUGHNESS clamp */
                metal = metal < 0.0f ? 0.0f : (metal > 1.0f ? 1.0f : metal);
            }
            float base[3] = { m->base_color_factor[0], m->base_color_factor[1], m->base_color_factor[2] };
            if (m->base_color_texture > -1) {
                tex_fetch(&ts, m->base_color_texture, uv[0], uv[1], tx);
                for (int k = 0; k < 3; ++k) base[k] *= tx[k];
            }
            if (m->normal_texture > -1) {               /* gBufferPass.frag:39-60 */
                tex_fetch(&ts, m->normal_texture, uv[0], uv[1], tx);
                double T[3], B[3], r[3];
                for (int k = 0; k < 3; ++k) T[k] = b0 * t->tg[0][k] + b1 * t->tg[1][k] + b2 * t->tg[2][k];
                const double hw = b0 * t->tg[0][3] + b1 * t->tg[1][3] + b2 * t->tg[2][3];
                B[0] = (n[1] * T[2] - n[2] * T[1]) * hw;
                B[1] = (n[2] * T[0] - n[0] * T[2]) * hw;
                B[2] = (n[0] * T[1] - n[1] * T[0]) * hw;
                const double lt = sqrt(T[0] * T[0] + T[1] * T[1] + T[2] * T[2]);
                const double lb = sqrt(B[0] * B[0] + B[1] * B[1] + B[2] * B[2]);
                const double sx = (double)(2.0f * tx[0] - 1.0f), sy = (double)(2.0f * tx[1] - 1.0f), sz = (double)(2.0f * tx[2] - 1.0f);
                for (int k = 0; k < 3; ++k)
                    r[k] = (sx * (lt > 0 ? T[k] / lt : 0.0) + sy * (lb > 0 ? B[k] / lb : 0.0)) + sz * (ln > 0 ? n[k] / ln : 0.0);
                n[0] = r[0]; n[1] = r[1]; n[2] = r[2];
                ln = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
            }
            float emi[3] = { m->emissive_factor[0], m->emissive_factor[1], m->emissive_factor[2] };
            if (m->emissive_texture > -1) {             /* SRGBtoLinear(texel, 2.2), gBufferPass.frag:111-114 */
                tex_fetch(&ts, m->emissive_texture, uv[0], uv[1], tx);
                for (int k = 0; k < 3; ++k) emi[k] *= (float)pow((double)tx[k], 2.2);
            }
            for (int k = 0; k < 3; ++k) {
                diffuse[pi * 4 + k] = to_unorm8(base[k] * (1.0f - 0.04f) * (1.0f - metal));
                specular[pi * 4 + k] = to_unorm8(0.04f * (1.0f - metal) + base[k] * metal);
                normal[pi * 4 + k] = float_to_half((float)((ln > 0 ? n[k] / ln : 0.0) * 0.5 + 0.5));
                emission[pi * 4 + k] = float_to_half(emi[k]);
            }
            diffuse[pi * 4 + 3] = to_unorm8(rough);
            specular[pi * 4 + 3] = to_unorm8(metal);
            normal[pi * 4 + 3] = float_to_half(1.0f);
            emission[pi * 4 + 3] = float_to_half(1.0f);
        }
    free(ids);
    free(tris);
    return 0;
}

int vgs_gbuffer(const vgi_scene_desc* scene, const vgi_camera* cam, uint32_t w, uint32_t h,
                uint8_t* diffuse, uint16_t* normal, uint8_t* specular, uint16_t* emission, float* depth)
{
    return vgs_gbuffer_tex(scene, NULL, 0, cam, w, h, diffuse, normal, specular, emission, depth);
}

/* Test aid: what the rasteriser hands the G-buffer fragment stage at every pixel - the material index (-1 = not covered),
 * the perspective-correctly interpolated, un-normalised world normal, the texture coordinate and the tangent (the `fs_in`
 * block of gBufferPass.frag), same visibility (alpha cutoff included) and interpolation as vgs_gbuffer_tex.
 * tests/test_ref_shaders.py feeds them to the reference's gBufferPass.frag. uv / tangent may be NULL. */
int vgs_gbuffer_attributes_tex(const vgi_scene_desc* scene, const vgi_texture* textures, uint32_t texture_count,
                               const vgi_camera* cam, uint32_t w, uint32_t h, int32_t* material, float* normal, float* uv,
                               float* tangent)
{
    uint32_t ntri;
    tri_world* tris = world_tris(scene, &ntri);
    const tex_set ts = { textures, texture_count };
    int32_t* ids = (int32_t*)malloc((size_t)w * h * sizeof(int32_t));
    float* depth = (float*)malloc((size_t)w * h * sizeof(float));
    raster_ids(cam->view_proj, tris, ntri, w, h, depth, ids, scene->materials, &ts);
#pragma omp parallel for schedule(static)
    for (int64_t y = 0; y < (int64_t)h; ++y)
        for (uint32_t x = 0; x < w; ++x) {
            const size_t pi = (size_t)y * w + x;
            material[pi] = -1;
            normal[pi * 3] = normal[pi * 3 + 1] = normal[pi * 3 + 2] = 0.0f;
            if (uv) uv[pi * 2] = uv[pi * 2 + 1] = 0.0f;
            if (tangent) tangent[pi * 4] = tangent[pi * 4 + 1] = tangent[pi * 4 + 2] = tangent[pi * 4 + 3] = 0.0f;
            if (ids[pi] < 0) continue;
            const tri_world* t = &tris[ids[pi]];
            proj_tri q;
            project(cam->view_proj, t, w, h, &q);
            double bb[3];
            pixel_bary(&q, x + 0.5, y + 0.5, bb);
            const double b0 = bb[0], b1 = bb[1], b2 = bb[2];
            for (int k = 0; k < 3; ++k) normal[pi * 3 + k] = (float)(b0 * t->n[0][k] + b1 * t->n[1][k] + b2 * t->n[2][k]);
            if (uv) frag_uv(t, bb, uv + pi * 2);
            if (tangent)
                for (int k = 0; k < 4; ++k) tangent[pi * 4 + k] = (float)(b0 * t->tg[0][k] + b1 * t->tg[1][k] + b2 * t->tg[2][k]);
            material[pi] = t->mat;
        }
    free(depth);
    free(ids);
    free(tris);
    return 0;
}

int vgs_gbuffer_attributes(const vgi_scene_desc* scene, const vgi_camera* cam, uint32_t w, uint32_t h,
                           int32_t* material, float* normal)
{
    return vgs_gbuffer_attributes_tex(scene, NULL, 0, cam, w, h, material, normal, NULL, NULL);
}
