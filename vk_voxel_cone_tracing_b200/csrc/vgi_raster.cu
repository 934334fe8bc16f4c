// vgi_raster.cu — producers of the hot path's image inputs: the directional light's shadow
// depth map and the camera G-buffer, rendered from the ctx's scene without Vulkan so that batched-view runs
// (BASELINE configs[4]: 64 cameras at 4K) need no host rasteriser.
// ref: VFS/Shaders/shadowPass.vert:33 (depth-only, proj * view * model), gBufferPass.vert:38-42, gBufferPass.frag:39-116
// (factors and textures: base colour, metallic-roughness, emissive, tangent-space normal map, alpha cutoff), formats
// VFS/RenderPass/GBufferPass.cpp:177-194.
//
// Rasterisation rule (the software definition the test producers pin: pixel-centre sampling, depth test LESS with ties
// to the lower triangle index, z clipped to [0,1]): visibility is one 64-bit atomicMin per fragment on
// (depth bits << 32 | triangle), then a resolve pass shades the winning triangle. A triangle with a vertex at or behind the
// camera plane (clip w <= 1e-4: floors and walls next to the viewer) is NOT dropped: it is rasterised with 2-D homogeneous
// edge functions (Olano & Greer 1997) — e_i(px, py) = det[V_j; V_k; (px, py, 1)] on V = (x_h, y_h, w), whose ratios
// e_i / (e_0 + e_1 + e_2) are the perspective-correct barycentric weights of the point the pixel's ray hits on the
// triangle's plane, valid on both sides of w = 0 — inside the pixel box of its part in front of the plane; the hardware
// rasteriser of the reference clips such triangles against the frustum and shades the same pixels.
// Edge functions and interpolation are evaluated in binary64 without FMA contraction (this file is compiled with
// -fmad=false), so depth and every quantised attribute equal the host producer bit for bit.
#include <cuda_fp16.h>

#include "vgi_device.cuh"

struct ProjTri {
    // ok == 1: pixel coordinates, NDC depth, 1 / w of the vertices.
    // ok == 2 (crosses the camera plane): e_i = x[i] * px + y[i] * py + z[i]; iw[] = clip z, wc[] = clip w of the vertices
    double x[3], y[3], z[3], iw[3];
    double wc[3];
    int box[4];                     // ok == 2: pixel box of the part in front of the plane (x0, x1, y0, y1), may be empty
    int ok;
    int alpha;                      // G-buffer pass: 1 = every fragment takes the alpha-cutoff test against the base-colour texture
};

#ifndef RASTER_SMALL_MAX
#define RASTER_SMALL_MAX 128   // bounding boxes up to this many pixels are walked by one thread (measured on the bench scene, shadow map + G-buffer: 256 -> 0.86 ms, 128 -> 0.81, 64 -> 0.84, 32 -> 0.90)
#endif
#define RASTER_HUGE_MIN (128 * 128)    // boxes above this many pixels are spread over the whole grid

#define RASTER_W_EPS 1e-4

DEVFN void project_tri(const float* M, const float4* tri_pos, uint32_t t, int w, int h, ProjTri& o)
{
    double c[3][4];
    int behind = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float4 p = tri_pos[(size_t)t * 3 + k];
#pragma unroll
        for (int r = 0; r < 4; ++r)
            c[k][r] = (double)M[r] * p.x + (double)M[4 + r] * p.y + (double)M[8 + r] * p.z + (double)M[12 + r];
        behind += c[k][3] <= RASTER_W_EPS ? 1 : 0;
    }
    o.box[0] = o.box[2] = 0; o.box[1] = o.box[3] = -1;
    o.wc[0] = o.wc[1] = o.wc[2] = 0.0;
    if (behind == 3) { o.ok = 0; return; }
    if (behind == 0) {
        o.ok = 1;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            o.iw[k] = 1.0 / c[k][3];
            o.x[k] = (c[k][0] * o.iw[k] * 0.5 + 0.5) * (double)w;
            o.y[k] = (c[k][1] * o.iw[k] * 0.5 + 0.5) * (double)h;
            o.z[k] = c[k][2] * o.iw[k];
        }
        return;
    }
    // crosses the camera plane: homogeneous pixel coordinates V = (x_h, y_h, w) with x_h / w = pixel x
    o.ok = 2;
    double X[3], Y[3], W[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        X[k] = (c[k][0] * 0.5 + c[k][3] * 0.5) * (double)w;
        Y[k] = (c[k][1] * 0.5 + c[k][3] * 0.5) * (double)h;
        W[k] = c[k][3];
        o.iw[k] = c[k][2];
        o.wc[k] = c[k][3];
    }
    double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int j = (i + 1) % 3, k = (i + 2) % 3;
        o.x[i] = Y[j] * W[k] - W[j] * Y[k];
        o.y[i] = W[j] * X[k] - X[j] * W[k];
        o.z[i] = X[j] * Y[k] - Y[j] * X[k];
        // pixel box of the polygon in front of w = eps: its vertices are the triangle's vertices in front and the points
        // where an edge crosses the plane
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
    if (ymax < 0 || ymin > h || xmax < 0 || xmin > w) return;       // empty box
    xmin = fmax(xmin, -1.0); ymin = fmax(ymin, -1.0); xmax = fmin(xmax, (double)w + 1.0); ymax = fmin(ymax, (double)h + 1.0);
    o.box[0] = max((int)floor(xmin - 0.5), 0);
    o.box[1] = min((int)ceil(xmax - 0.5), w - 1);
    o.box[2] = max((int)floor(ymin - 0.5), 0);
    o.box[3] = min((int)ceil(ymax - 0.5), h - 1);
}

// ok == 2: perspective-correct barycentric weights and NDC depth of pixel centre (px, py); false outside the triangle
// or at / behind the camera plane
DEVFN bool hom_bary(const ProjTri& q, double px, double py, double* b, double* z)
{
    const double e0 = (q.x[0] * px + q.y[0] * py) + q.z[0];
    const double e1 = (q.x[1] * px + q.y[1] * py) + q.z[1];
    const double e2 = (q.x[2] * px + q.y[2] * py) + q.z[2];
    const double s = (e0 + e1) + e2;
    if (s == 0.0) return false;
    b[0] = e0 / s; b[1] = e1 / s; b[2] = e2 / s;
    if (b[0] < 0 || b[1] < 0 || b[2] < 0) return false;
    const double wc = (b[0] * q.wc[0] + b[1] * q.wc[1]) + b[2] * q.wc[2];
    if (!(wc > RASTER_W_EPS)) return false;
    *z = ((b[0] * q.iw[0] + b[1] * q.iw[1]) + b[2] * q.iw[2]) / wc;
    return true;
}

struct RasterParams {
    float M[16];
    const float4* tri_pos;
    const float4* tri_nrm;
    const float4* tri_tan;      // nullptr unless the scene is normal-mapped
    const vgi_material* materials;
    TexSet tex;
    uint32_t ntri;
    int w, h;
    ProjTri* proj;
    unsigned long long* keys;   // per pixel: depth bits << 32 | triangle
    uint32_t* large;            // triangles whose bounding box one thread should not walk
    uint32_t* large_count;
};

struct Box { int x0, x1, y0, y1; double area; bool any; };

DEVFN Box tri_box(const ProjTri& q, int w, int h)
{
    Box b;
    b.any = false;
    b.area = 1.0;
    if (!q.ok) return b;
    if (q.ok == 2) {
        b.x0 = q.box[0]; b.x1 = q.box[1]; b.y0 = q.box[2]; b.y1 = q.box[3];
        b.any = b.x0 <= b.x1 && b.y0 <= b.y1;
        return b;
    }
    const double ymin = fmin(q.y[0], fmin(q.y[1], q.y[2])), ymax = fmax(q.y[0], fmax(q.y[1], q.y[2]));
    const double xmin = fmin(q.x[0], fmin(q.x[1], q.x[2])), xmax = fmax(q.x[0], fmax(q.x[1], q.x[2]));
    if (ymax < 0 || ymin > h || xmax < 0 || xmin > w) return b;
    b.area = (q.x[1] - q.x[0]) * (q.y[2] - q.y[0]) - (q.x[2] - q.x[0]) * (q.y[1] - q.y[0]);
    if (b.area == 0.0) return b;
    b.x0 = max((int)floor(xmin - 0.5), 0);
    b.x1 = min((int)ceil(xmax - 0.5), w - 1);
    b.y0 = max((int)floor(ymin - 0.5), 0);
    b.y1 = min((int)ceil(ymax - 0.5), h - 1);
    b.any = b.x0 <= b.x1 && b.y0 <= b.y1;
    return b;
}

// perspective-correct barycentric weights of pixel centre (px, py) on a triangle that covers it (both kinds)
DEVFN void persp_bary(const ProjTri& q, double px, double py, double* b)
{
    if (q.ok == 2) {
        double z;
        hom_bary(q, px, py, b, &z);     // the weights of the homogeneous edge functions are perspective-correct already
        return;
    }
    const double area = (q.x[1] - q.x[0]) * (q.y[2] - q.y[0]) - (q.x[2] - q.x[0]) * (q.y[1] - q.y[0]);
    double b0 = ((q.x[1] - px) * (q.y[2] - py) - (q.x[2] - px) * (q.y[1] - py)) / area;
    double b1 = ((q.x[2] - px) * (q.y[0] - py) - (q.x[0] - px) * (q.y[2] - py)) / area;
    double b2 = 1.0 - b0 - b1;
    b0 *= q.iw[0]; b1 *= q.iw[1]; b2 *= q.iw[2];
    const double bs = b0 + b1 + b2;
    b[0] = b0 / bs; b[1] = b1 / bs; b[2] = b2 / bs;
}

// the interpolated texture coordinate the fragment stage receives (gBufferPass.vert:38)
DEVFN void frag_uv(const RasterParams& rp, uint32_t t, const double* b, float* uv)
{
    const float2 a = rp.tex.tri_uv[3 * (size_t)t], c = rp.tex.tri_uv[3 * (size_t)t + 1], e = rp.tex.tri_uv[3 * (size_t)t + 2];
    uv[0] = (float)((b[0] * a.x + b[1] * c.x) + b[2] * e.x);
    uv[1] = (float)((b[0] * a.y + b[1] * c.y) + b[2] * e.y);
}

// ref: gBufferPass.frag:88-99 — alphaMode > 0 and baseColorFactor.a * texture.a < alphaCutoff: the fragment is discarded
// before it writes depth. true = keep.
DEVFN bool alpha_keep(const RasterParams& rp, const ProjTri& q, uint32_t t, double px, double py)
{
    double b[3];
    float uv[2], tx[4];
    persp_bary(q, px, py, b);
    frag_uv(rp, t, b, uv);
    const vgi_material& m = rp.materials[__float_as_int(rp.tri_pos[(size_t)t * 3].w)];
    tex_fetch(rp.tex, m.base_color_texture, uv[0], uv[1], tx);
    return !(m.base_color_factor[3] * tx[3] < m.alpha_cutoff);
}

template <bool ALPHA>
DEVFN void raster_pixel(const RasterParams& rp, const ProjTri& q, double area, uint32_t t, int x, int y)
{
    const double px = x + 0.5, py = y + 0.5;
    if (q.ok == 2) {
        double b[3], z;
        if (!hom_bary(q, px, py, b, &z)) return;
        if (z < 0.0 || z > 1.0) return;
        const float zf = (float)z;
        if (!(zf < 1.0f)) return;
        if (ALPHA && q.alpha && !alpha_keep(rp, q, t, px, py)) return;
        atomicMin(rp.keys + (size_t)y * rp.w + x, ((unsigned long long)__float_as_uint(zf) << 32) | t);
        return;
    }
    const double e0 = (q.x[1] - px) * (q.y[2] - py) - (q.x[2] - px) * (q.y[1] - py);
    const double e1 = (q.x[2] - px) * (q.y[0] - py) - (q.x[0] - px) * (q.y[2] - py);
    // e / area < 0 <=> the signs differ (e != 0): most pixels of a bounding box leave before the two divisions
    if ((e0 != 0.0 && (e0 < 0.0) != (area < 0.0)) || (e1 != 0.0 && (e1 < 0.0) != (area < 0.0))) return;
    const double w0 = e0 / area;
    const double w1 = e1 / area;
    const double w2 = 1.0 - w0 - w1;
    if (w0 < 0 || w1 < 0 || w2 < 0) return;
    const double z = w0 * q.z[0] + w1 * q.z[1] + w2 * q.z[2];
    if (z < 0.0 || z > 1.0) return;
    const float zf = (float)z;
    if (!(zf < 1.0f)) return; // the cleared depth is 1 and the test is LESS
    if (ALPHA && q.alpha && !alpha_keep(rp, q, t, px, py)) return;
    const unsigned long long key = ((unsigned long long)__float_as_uint(zf) << 32) | t;
    atomicMin(rp.keys + (size_t)y * rp.w + x, key);
}

__global__ void __launch_bounds__(256) k_raster_clear(unsigned long long* keys, size_t n, uint32_t* large_count)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = ~0ull;
    if (i == 0) { large_count[0] = 0u; large_count[1] = 0u; }
}

template <bool ALPHA>
__global__ void __launch_bounds__(128) k_raster_small(const __grid_constant__ RasterParams rp)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= rp.ntri) return;
    ProjTri q;
    project_tri(rp.M, rp.tri_pos, t, rp.w, rp.h, q);
    q.alpha = 0;
    if (ALPHA) {
        const vgi_material& m = rp.materials[__float_as_int(rp.tri_pos[(size_t)t * 3].w)];
        if (m.alpha_mode > 0) {
            if (m.base_color_texture > -1) q.alpha = 1;
            else if (m.base_color_factor[3] < m.alpha_cutoff) q.ok = 0;     // every fragment of it is discarded
        }
    }
    rp.proj[t] = q;
    const Box b = tri_box(q, rp.w, rp.h);
    if (!b.any) return;
    const long long boxPx = (long long)(b.x1 - b.x0 + 1) * (b.y1 - b.y0 + 1);
    if (boxPx > RASTER_HUGE_MIN) {          // huge boxes queue up from the end of the same array
        rp.large[rp.ntri - 1u - atomicAdd(rp.large_count + 1, 1u)] = t;
        return;
    }
    if (boxPx > RASTER_SMALL_MAX) {
        rp.large[atomicAdd(rp.large_count, 1u)] = t;
        return;
    }
    for (int y = b.y0; y <= b.y1; ++y)
        for (int x = b.x0; x <= b.x1; ++x) raster_pixel<ALPHA>(rp, q, b.area, t, x, y);
}

// medium boxes (257 .. RASTER_HUGE_MIN pixels): one warp per triangle, lanes stride over the box row-major (the walk
// advances (x, y) by 32 pixels with one conditional wrap instead of a division per pixel). Measured and rejected: one
// lane per ROW walking only the span the triangle crosses in it (343 -> 1050 us for the 4096^2 shadow map: the boxes of
// this size class are mostly a few rows high, so most lanes idle).
template <bool ALPHA>
__global__ void __launch_bounds__(256) k_raster_large(const __grid_constant__ RasterParams rp)
{
    const uint32_t n = *rp.large_count;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warpsPerGrid = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warpsPerGrid) {
        const uint32_t t = rp.large[i];
        const ProjTri& q = rp.proj[t];
        const Box b = tri_box(q, rp.w, rp.h);
        const int bw = b.x1 - b.x0 + 1, bh = b.y1 - b.y0 + 1;
        const int dq = 32 / bw, dr = 32 % bw;
        int x = (int)lane % bw, y = (int)lane / bw;
        while (y < bh) {
            raster_pixel<ALPHA>(rp, q, b.area, t, b.x0 + x, b.y0 + y);
            x += dr; y += dq;
            if (x >= bw) { x -= bw; ++y; }
        }
    }
}

// the few triangles that cover a large part of the image (floors, walls): every block of the grid takes pieces
template <bool ALPHA>
__global__ void __launch_bounds__(256) k_raster_huge(const __grid_constant__ RasterParams rp)
{
    const uint32_t n = rp.large_count[1];
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t t = rp.large[rp.ntri - 1u - i];
        const ProjTri& q = rp.proj[t];
        const Box b = tri_box(q, rp.w, rp.h);
        const int tilesX = (b.x1 - b.x0 + 32) / 32, tilesY = (b.y1 - b.y0 + 32) / 32;
        for (int tile = blockIdx.x; tile < tilesX * tilesY; tile += gridDim.x) {
            const int tx = b.x0 + (tile % tilesX) * 32, ty = b.y0 + (tile / tilesX) * 32;
            const int tw = min(32, b.x1 - tx + 1), th = min(32, b.y1 - ty + 1);
            // thread = (row p >> 5, column p & 31) of the 32 x 32 tile: no division per pixel
            for (int p = threadIdx.x; p < 32 * th; p += 256)
                if ((p & 31) < tw) raster_pixel<ALPHA>(rp, q, b.area, t, tx + (p & 31), ty + (p >> 5));
        }
    }
}

__global__ void __launch_bounds__(256) k_raster_resolve_depth(const unsigned long long* keys, size_t n, float* depth)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long k = keys[i];
    depth[i] = (k == ~0ull) ? 1.0f : __uint_as_float((uint32_t)(k >> 32));
}

DEVFN uint8_t to_unorm8(float x)
{
    if (!(x > 0.0f)) return 0;
    if (x > 1.0f) x = 1.0f;
    return (uint8_t)(x * 255.0f + 0.5f);
}

struct GBufferTarget {
    uchar4* diffuse; uint2* normal; uchar4* specular; uint2* emission; float* depth;
};

DEVFN uint2 pack_half4(float a, float b, float c, float d)
{
    const __half2 lo = __halves2half2(__float2half_rn(a), __float2half_rn(b));
    const __half2 hi = __halves2half2(__float2half_rn(c), __float2half_rn(d));
    return make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
}

// ref: gBufferPass.frag:62-116; TEX = false: factor-only scene
template <bool TEX>
__global__ void __launch_bounds__(256) k_raster_resolve_gbuffer(const __grid_constant__ RasterParams rp, const GBufferTarget g)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= rp.w || y >= rp.h) return;
    const size_t pi = (size_t)y * rp.w + x;
    const unsigned long long k = rp.keys[pi];
    if (k == ~0ull) {
        g.diffuse[pi] = make_uchar4(0, 0, 0, 0);
        g.specular[pi] = make_uchar4(0, 0, 0, 0);
        g.normal[pi] = make_uint2(0u, 0u);
        g.emission[pi] = make_uint2(0u, 0u);
        g.depth[pi] = 1.0f;
        return;
    }
    g.depth[pi] = __uint_as_float((uint32_t)(k >> 32));
    const uint32_t t = (uint32_t)k;
    const ProjTri q = rp.proj[t];
    const double px = x + 0.5, py = y + 0.5;
    double b[3];
    persp_bary(q, px, py, b);
    const double b0 = b[0], b1 = b[1], b2 = b[2];
    const float4 n0 = rp.tri_nrm[(size_t)t * 3], n1 = rp.tri_nrm[(size_t)t * 3 + 1], n2 = rp.tri_nrm[(size_t)t * 3 + 2];
    double n[3] = { b0 * n0.x + b1 * n1.x + b2 * n2.x, b0 * n0.y + b1 * n1.y + b2 * n2.y, b0 * n0.z + b1 * n1.z + b2 * n2.z };
    double ln = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    const int mi = __float_as_int(rp.tri_pos[(size_t)t * 3].w);
    const vgi_material& m = rp.materials[mi];
    float uv[2] = { 0.0f, 0.0f }, tx[4];
    if (TEX) frag_uv(rp, t, b, uv);
    float rough = m.roughness_factor, metal = m.metallic_factor;
    if (TEX && m.metallic_roughness_texture > -1) {        // g = roughness, b = metallic; this branch does not clamp (gBufferPass.frag:77-82)
        tex_fetch(rp.tex, m.metallic_roughness_texture, uv[0], uv[1], tx);
        rough *= tx[1];
        metal *= tx[2];
    } else {
        rough = rough < 0.04f ? 0.04f : (rough > 1.0f ? 1.0f : rough); // MIN_ROUGHNESS clamp
        metal = metal < 0.0f ? 0.0f : (metal > 1.0f ? 1.0f : metal);
    }
    float base[3] = { m.base_color_factor[0], m.base_color_factor[1], m.base_color_factor[2] };
    if (TEX && m.base_color_texture > -1) {
        tex_fetch(rp.tex, m.base_color_texture, uv[0], uv[1], tx);
#pragma unroll
        for (int c = 0; c < 3; ++c) base[c] *= tx[c];
    }
    if (TEX && m.normal_texture > -1) {
        // ref: gBufferPass.frag:39-60 — tangent-space sample, bitangent = cross(N, T) * handedness, each axis normalised
        tex_fetch(rp.tex, m.normal_texture, uv[0], uv[1], tx);
        const float4 t0 = rp.tri_tan[(size_t)t * 3], t1 = rp.tri_tan[(size_t)t * 3 + 1], t2 = rp.tri_tan[(size_t)t * 3 + 2];
        const double T[3] = { b0 * t0.x + b1 * t1.x + b2 * t2.x, b0 * t0.y + b1 * t1.y + b2 * t2.y, b0 * t0.z + b1 * t1.z + b2 * t2.z };
        const double hw = b0 * t0.w + b1 * t1.w + b2 * t2.w;
        const double B[3] = { (n[1] * T[2] - n[2] * T[1]) * hw, (n[2] * T[0] - n[0] * T[2]) * hw, (n[0] * T[1] - n[1] * T[0]) * hw };
        const double lt = sqrt(T[0] * T[0] + T[1] * T[1] + T[2] * T[2]);
        const double lb = sqrt(B[0] * B[0] + B[1] * B[1] + B[2] * B[2]);
        const double sx = (double)(2.0f * tx[0] - 1.0f), sy = (double)(2.0f * tx[1] - 1.0f), sz = (double)(2.0f * tx[2] - 1.0f);
        double r[3];
#pragma unroll
        for (int c = 0; c < 3; ++c)
            r[c] = (sx * (lt > 0 ? T[c] / lt : 0.0) + sy * (lb > 0 ? B[c] / lb : 0.0)) + sz * (ln > 0 ? n[c] / ln : 0.0);
        n[0] = r[0]; n[1] = r[1]; n[2] = r[2];
        ln = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
    }
    float emi[3] = { m.emissive_factor[0], m.emissive_factor[1], m.emissive_factor[2] };
    if (TEX && m.emissive_texture > -1) {      // SRGBtoLinear(texel, 2.2), gBufferPass.frag:111-114; binary64 pow on both producers
        tex_fetch(rp.tex, m.emissive_texture, uv[0], uv[1], tx);
#pragma unroll
        for (int c = 0; c < 3; ++c) emi[c] *= (float)pow((double)tx[c], 2.2);
    }
    uint8_t dif[3], spc[3];
    float nn[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        dif[c] = to_unorm8(base[c] * (1.0f - 0.04f) * (1.0f - metal));
        spc[c] = to_unorm8(0.04f * (1.0f - metal) + base[c] * metal);
        nn[c] = (float)((ln > 0 ? n[c] / ln : 0.0) * 0.5 + 0.5);
    }
    g.diffuse[pi] = make_uchar4(dif[0], dif[1], dif[2], to_unorm8(rough));
    g.specular[pi] = make_uchar4(spc[0], spc[1], spc[2], to_unorm8(metal));
    g.normal[pi] = pack_half4(nn[0], nn[1], nn[2], 1.0f);
    g.emission[pi] = pack_half4(emi[0], emi[1], emi[2], 1.0f);
}

template <bool ALPHA>
static int launch_visibility(vgi_ctx* c, const RasterParams& rp, cudaStream_t s)
{
    const size_t npx = (size_t)rp.w * rp.h;
    c->timer.begin("k_raster_clear", s);
    k_raster_clear<<<(unsigned)((npx + 255) / 256), 256, 0, s>>>(rp.keys, npx, rp.large_count);
    c->timer.end(s);
    int n = 1;
    if (rp.ntri) {
        c->timer.begin("k_raster_small", s);
        k_raster_small<ALPHA><<<(rp.ntri + 127) / 128, 128, 0, s>>>(rp);
        c->timer.end(s);
        c->timer.begin("k_raster_large", s);
        k_raster_large<ALPHA><<<148 * 8, 256, 0, s>>>(rp);
        c->timer.end(s);
        c->timer.begin("k_raster_huge", s);
        k_raster_huge<ALPHA><<<148 * 8, 256, 0, s>>>(rp);
        c->timer.end(s);
        n += 3;
    }
    return n;
}

int vgi_launch_render_shadow(vgi_ctx* c, const float* M, uint32_t w, uint32_t h, float* depth, cudaStream_t s)
{
    RasterParams rp;
    memcpy(rp.M, M, sizeof rp.M);
    rp.tri_pos = c->tri_pos; rp.tri_nrm = c->tri_nrm; rp.tri_tan = nullptr; rp.materials = c->materials; rp.ntri = c->ntri;
    rp.tex = c->texset();
    rp.w = (int)w; rp.h = (int)h;
    rp.proj = (ProjTri*)c->raster_proj; rp.keys = c->raster_keys; rp.large = c->raster_large; rp.large_count = c->raster_large + c->ntri;
    int n = launch_visibility<false>(c, rp, s);     // no fragment shader in the shadow pass: nothing is discarded
    const size_t npx = (size_t)w * h;
    c->timer.begin("k_raster_resolve_depth", s);
    k_raster_resolve_depth<<<(unsigned)((npx + 255) / 256), 256, 0, s>>>(rp.keys, npx, depth);
    c->timer.end(s);
    return n + 1;
}

int vgi_launch_render_gbuffer(vgi_ctx* c, const float* M, const vgi_gbuffer* target, cudaStream_t s)
{
    RasterParams rp;
    memcpy(rp.M, M, sizeof rp.M);
    rp.tri_pos = c->tri_pos; rp.tri_nrm = c->tri_nrm; rp.tri_tan = c->tri_tan; rp.materials = c->materials; rp.ntri = c->ntri;
    rp.tex = c->texset();
    rp.w = (int)target->width; rp.h = (int)target->height;
    rp.proj = (ProjTri*)c->raster_proj; rp.keys = c->raster_keys; rp.large = c->raster_large; rp.large_count = c->raster_large + c->ntri;
    // gBufferPass.frag:96 discards by alpha cutoff: only scenes with such a material pay for the test
    int n = c->scene_alpha_tested ? launch_visibility<true>(c, rp, s) : launch_visibility<false>(c, rp, s);
    GBufferTarget g;
    g.diffuse = (uchar4*)target->diffuse_rgba8; g.normal = (uint2*)target->normal_rgba16f;
    g.specular = (uchar4*)target->specular_rgba8; g.emission = (uint2*)target->emission_rgba16f;
    g.depth = (float*)target->depth_f32;
    c->timer.begin("k_raster_resolve_gbuffer", s);
    if (c->scene_max_texture_all > -1) k_raster_resolve_gbuffer<true><<<dim3((rp.w + 31) / 32, (rp.h + 7) / 8), 256, 0, s>>>(rp, g);
    else k_raster_resolve_gbuffer<false><<<dim3((rp.w + 31) / 32, (rp.h + 7) / 8), 256, 0, s>>>(rp, g);
    c->timer.end(s);
    return n + 1;
}

size_t vgi_raster_proj_bytes(uint32_t ntri) { return (size_t)(ntri ? ntri : 1) * sizeof(ProjTri); }
